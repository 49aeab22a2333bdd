"""Buffer table of one step call: names, shapes and dtypes of every leaf that crosses the C-ABI, and the packing of
those arrays (numpy on the host, torch on the device) into ``LobStepBuffers`` pointers.

Leaf order follows the reference's state pytrees (gymnax_exchange/jaxen/StatesandParams.py:14-122):
``MultiAgentState = WorldState + [MMEnvState | ExecEnvState per agent type]``.
"""
import ctypes as C

import numpy as np

from . import abi
from .config import num_action_msgs, num_msgs_per_step

WORLD_I32 = ("asks", "bids", "trades", "init_time", "window_index", "max_steps", "start_index", "step_counter",
             "best_bids", "best_asks", "time", "order_id_counter")
WORLD_F32 = ("mid_price", "delta_time")
PARAMS = ("message_data", "init_asks", "init_bids", "init_trades", "init_init_time", "init_max_steps",
          "init_start_index")


def leaf_specs(cfg: abi.LobStepConfig, batch: int):
    """name -> (shape, dtype, role) for state ('s'), input ('i'), output ('o') and workspace ('w') leaves."""
    B, No, Nt, N = batch, cfg.book.n_orders, cfg.book.n_trades, num_msgs_per_step(cfg)
    T = cfg.n_agent_types
    sp = {
        "asks": ((B, No, 6), np.int32, "s"), "bids": ((B, No, 6), np.int32, "s"),
        "trades": ((B, Nt, 8), np.int32, "s"), "init_time": ((B, 2), np.int32, "s"),
        "window_index": ((B,), np.int32, "s"), "max_steps": ((B,), np.int32, "s"),
        "start_index": ((B,), np.int32, "s"), "step_counter": ((B,), np.int32, "s"),
        "best_bids": ((B, N, 2), np.int32, "s"), "best_asks": ((B, N, 2), np.int32, "s"),
        "time": ((B, 2), np.int32, "s"), "order_id_counter": ((B,), np.int32, "s"),
        "mid_price": ((B,), np.float32, "s"), "delta_time": ((B,), np.float32, "s"),
        "perm": ((B, max(num_action_msgs(cfg), 1)), np.int32, "i"),
        "reset_window": ((B,), np.int32, "i"), "reset_is_sell": ((B, max(T, 1)), np.int32, "i"),
        "done_all": ((B,), np.uint8, "o"),
        "info_world_i32": ((B, len(abi.WINFO_I32)), np.int32, "o"),
        "info_world_f32": ((B, len(abi.WINFO_F32)), np.float32, "o"),
    }
    # the piped step's workspace (env records, then the agents' messages: csrc/lob_pipe.cuh); a flat buffer of
    # lob_split_workspace_words(cfg, B) words, declared [B, words per env] so that a vmapped custom call sizes it right
    sp["work_split"] = ((B, max(abi.split_workspace_words(cfg, 1), 4)), np.int32, "w")
    if No > 128:   # workspace of lob_step_launch's window pass for deep books (scratch, not state; see include/lobstep.h)
        sp["work_redo_list"] = ((B,), np.int32, "w")
        sp["work_redo_count"] = ((4,), np.int32, "w")
    if cfg.book.cancel_mode >= 2:   # job:142-164: the uniform draws of the two random-cancel fallbacks, per message
        sp["cancel_u"] = ((B, N, 2), np.float32, "i")
    for t in range(T):
        a = cfg.agent[t]
        n = a.n_agents
        li, lf = abi.state_leaves(a.kind)
        for name in li:
            sp[f"a{t}_{name}"] = ((B, n), np.int32, "s")
        for name in lf:
            sp[f"a{t}_{name}"] = ((B, n), np.float32, "s")
        ki, kf = abi.info_cols(a.kind)
        aw = abi.action_width(a)
        sp[f"actions{t}"] = ((B, n) if aw == 1 else (B, n, aw), np.int32, "i")
        sp[f"obs{t}"] = ((B, n, abi.obs_dim(a.kind, a.observation_space, bool(cfg.ep_type_fixed_time))), np.float32, "o")
        sp[f"reward{t}"] = ((B, n), np.float32, "o")
        sp[f"done_agents{t}"] = ((B, n), np.uint8, "o")
        sp[f"info_i32_{t}"] = ((B, n, len(ki)), np.int32, "o")
        sp[f"info_f32_{t}"] = ((B, n, len(kf)), np.float32, "o")
    return sp


def state_names(cfg: abi.LobStepConfig):
    return [k for k, v in leaf_specs(cfg, 1).items() if v[2] == "s"]


def alloc_numpy(cfg: abi.LobStepConfig, batch: int):
    out = {}
    for k, (shape, dt, _) in leaf_specs(cfg, batch).items():
        out[k] = np.zeros(shape, dt)
    return out


def alloc_torch(cfg: abi.LobStepConfig, batch: int, device):
    import torch
    tdt = {np.int32: torch.int32, np.float32: torch.float32, np.uint8: torch.uint8}
    return {k: torch.zeros(shape, dtype=tdt[dt], device=device) for k, (shape, dt, _) in leaf_specs(cfg, batch).items()}


def _ptr(a, ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("buffers must be C-contiguous")
        return C.cast(a.ctypes.data, C.POINTER(ctype))
    if not a.is_contiguous():
        raise ValueError("buffers must be contiguous")
    return C.cast(a.data_ptr(), C.POINTER(ctype))


def pack_buffers(cfg: abi.LobStepConfig, arrays: dict, params: dict) -> abi.LobStepBuffers:
    """Raw pointers of ``arrays`` (leaf_specs names) and ``params`` (PARAMS names) -> LobStepBuffers.
    The caller keeps the arrays alive for the duration of the call."""
    b = abi.LobStepBuffers()
    for name in WORLD_I32:
        setattr(b, name, _ptr(arrays[name], C.c_int32))
    for name in WORLD_F32:
        setattr(b, name, _ptr(arrays[name], C.c_float))
    for t in range(cfg.n_agent_types):
        a = cfg.agent[t]
        li, lf = abi.state_leaves(a.kind)
        for j, name in enumerate(li):
            b.agent_i32[t][j] = _ptr(arrays[f"a{t}_{name}"], C.c_int32)
        for j, name in enumerate(lf):
            b.agent_f32[t][j] = _ptr(arrays[f"a{t}_{name}"], C.c_float)
        b.actions[t] = _ptr(arrays[f"actions{t}"], C.c_int32)
        b.obs[t] = _ptr(arrays[f"obs{t}"], C.c_float)
        b.reward[t] = _ptr(arrays[f"reward{t}"], C.c_float)
        b.done_agents[t] = _ptr(arrays[f"done_agents{t}"], C.c_uint8)
        b.info_agent_i32[t] = _ptr(arrays[f"info_i32_{t}"], C.c_int32)
        b.info_agent_f32[t] = _ptr(arrays[f"info_f32_{t}"], C.c_float)
    b.perm = _ptr(arrays["perm"], C.c_int32)
    b.reset_window = _ptr(arrays["reset_window"], C.c_int32)
    b.reset_is_sell = _ptr(arrays["reset_is_sell"], C.c_int32)
    b.cancel_u = _ptr(arrays.get("cancel_u"), C.c_float)
    for name in PARAMS:
        setattr(b, name, _ptr(params[name], C.c_int32))
    b.done_all = _ptr(arrays["done_all"], C.c_uint8)
    b.info_world_i32 = _ptr(arrays["info_world_i32"], C.c_int32)
    b.info_world_f32 = _ptr(arrays["info_world_f32"], C.c_float)
    b.work_redo_list = _ptr(arrays.get("work_redo_list"), C.c_int32)
    b.work_redo_count = _ptr(arrays.get("work_redo_count"), C.c_int32)
    b.work_split = _ptr(arrays.get("work_split"), C.c_int32)
    return b


# ---- the reference's pytrees <-> leaf names (for a binding that sits under the reference's own MARLEnv) ------------
# StatesandParams.py:14-44 field -> leaf name of leaf_specs / pack_buffers
WORLD_FIELD_OF_LEAF = {"asks": "ask_raw_orders", "bids": "bid_raw_orders", "trades": "trades", "init_time": "init_time",
                       "window_index": "window_index", "max_steps": "max_steps_in_episode", "start_index": "start_index",
                       "step_counter": "step_counter", "best_bids": "best_bids", "best_asks": "best_asks", "time": "time",
                       "order_id_counter": "order_id_counter", "mid_price": "mid_price", "delta_time": "delta_time"}
_PARAM_FIELD = {"init_asks": "ask_raw_orders", "init_bids": "bid_raw_orders", "init_trades": "trades",
                "init_init_time": "init_time", "init_max_steps": "max_steps_in_episode", "init_start_index": "start_index"}


def _get(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


def flatten(cfg: abi.LobStepConfig, state) -> dict:
    """``MultiAgentState`` (the reference's flax struct, or this package's view) -> {leaf name: array} in the order of
    ``leaf_specs``' state leaves: ``state.world_state.<field>`` then ``state.agent_states[t].<field>``."""
    out = {leaf: _get(state.world_state, field) for leaf, field in WORLD_FIELD_OF_LEAF.items()}
    for t in range(cfg.n_agent_types):
        li, lf = abi.state_leaves(cfg.agent[t].kind)
        for name in li + lf:
            out[f"a{t}_{name}"] = _get(state.agent_states[t], name)
    return out


def unflatten(cfg: abi.LobStepConfig, state, leaves: dict):
    """The inverse: a copy of ``state`` (any object with ``.replace``: flax structs, or dataclasses via
    ``dataclasses.replace``) whose leaves are taken from ``leaves``."""
    import dataclasses

    def repl(obj, **kw):
        if isinstance(obj, dict):
            return {**obj, **kw}
        return obj.replace(**kw) if hasattr(obj, "replace") else dataclasses.replace(obj, **kw)
    ws = repl(state.world_state, **{field: leaves[leaf] for leaf, field in WORLD_FIELD_OF_LEAF.items()})
    agents = []
    for t in range(cfg.n_agent_types):
        li, lf = abi.state_leaves(cfg.agent[t].kind)
        agents.append(repl(state.agent_states[t], **{n: leaves[f"a{t}_{n}"] for n in li + lf}))
    return repl(state, world_state=ws, agent_states=agents)


def flatten_params(params) -> dict:
    """``MultiAgentParams`` (marl_env.py:96-127) -> {PARAMS name: array}: the day tensor and the stacked reset states."""
    lp = params.loaded_params
    init = lp.init_states_array
    if isinstance(init, dict) and "init_asks" in init:     # this package's own LoadedEnvParams
        return {"message_data": lp.message_data, **{k: init[k] for k in PARAMS if k != "message_data"}}
    return {"message_data": lp.message_data, **{k: _get(init, f) for k, f in _PARAM_FIELD.items()}}


def field_offset(cfg: abi.LobStepConfig, name: str) -> int:
    """Byte offset, inside ``LobStepBuffers``, of the pointer field that the leaf ``name`` (a ``leaf_specs`` or ``PARAMS``
    name) fills -- the table a table-driven binding (the XLA-FFI handler) passes next to its operands."""
    import re
    F, P = abi.LobStepBuffers, C.sizeof(C.c_void_p)
    if name in WORLD_I32 or name in WORLD_F32 or name in PARAMS or name in (
            "perm", "reset_window", "reset_is_sell", "cancel_u", "done_all", "info_world_i32", "info_world_f32",
            "work_redo_list", "work_redo_count", "work_split"):
        return getattr(F, name).offset
    m = re.fullmatch(r"a(\d+)_(\w+)", name)
    if m:
        t, leaf = int(m.group(1)), m.group(2)
        li, lf = abi.state_leaves(cfg.agent[t].kind)
        if leaf in li:
            return F.agent_i32.offset + (t * abi.LOB_MAX_AGENT_I32 + li.index(leaf)) * P
        return F.agent_f32.offset + (t * abi.LOB_MAX_AGENT_F32 + lf.index(leaf)) * P
    m = re.fullmatch(r"(actions|obs|reward|done_agents|info_i32_|info_f32_)(\d+)", name)
    if not m:
        raise KeyError(name)
    field = {"info_i32_": "info_agent_i32", "info_f32_": "info_agent_f32"}.get(m.group(1), m.group(1))
    return getattr(F, field).offset + int(m.group(2)) * P


def pack_replay(asks, bids, trades, msgs, start, n_msgs, best_out=None, cancel_u=None) -> abi.LobReplayBuffers:
    r = abi.LobReplayBuffers()
    r.asks, r.bids, r.trades = _ptr(asks, C.c_int32), _ptr(bids, C.c_int32), _ptr(trades, C.c_int32)
    r.msgs = _ptr(msgs, C.c_int32)
    r.start = _ptr(start, C.c_int64)
    r.n_msgs_total = int(msgs.shape[0])
    r.n_msgs = int(n_msgs)
    r.best_out = _ptr(best_out, C.c_int32)
    r.cancel_u = _ptr(cancel_u, C.c_float)
    return r
