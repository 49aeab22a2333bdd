"""Host-side mirror of the reference's environment API for the LOB step, natively batched.

Reference call sites replaced (gymnax_exchange/jaxen):
  ``MARLEnv(key, MultiAgentConfig)``                              marl_env.py:46
  ``env.default_params``                                          marl_env.py:96-127, base_env.py:180-187
  ``vmap(env.reset, (0, None))(keys, params)``                     marl_env.py:764, ippo_rnn_JAXMARL.py:571
  ``vmap(env.step, (0, 0, 0, None))(keys, state, actions, params)`` marl_env.py:776, ippo_rnn_JAXMARL.py:616
  ``BaseLOBEnv.step_env`` (pure book replay)                       base_env.py:189-216

Differences that are deliberate: the batch is native (one launch steps every env: the leading ``B`` the trainer
gets from ``vmap``), the state leaves are updated in place (donated), and the PRNG products the reference draws
with ``jax.random`` inside the step (action-message permutation marl:294-295, reset window base:222-225,
``is_sell_task`` exe:221) are drawn on the device by the host layer and handed to the kernel as buffers.

The compute is ONLY the CUDA library (csrc/liblobstep.so, through include/lobstep.h); nothing here falls back to
the CPU.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Any, List

import numpy as np

from . import _lib, abi, lobster, states
from .config import (Execution_EnvironmentConfig, MarketMaking_EnvironmentConfig, MultiAgentConfig,
                     World_EnvironmentConfig, book_config, num_action_msgs, num_msgs_per_step, to_step_config)
from .spaces import Box, Discrete, MultiDiscrete


# ---- state / params containers (field names of StatesandParams.py:14-122) ---------------------------------
@dataclass
class WorldState:
    ask_raw_orders: Any
    bid_raw_orders: Any
    trades: Any
    init_time: Any
    window_index: Any
    max_steps_in_episode: Any
    start_index: Any
    step_counter: Any
    best_bids: Any
    best_asks: Any
    time: Any
    order_id_counter: Any
    mid_price: Any
    delta_time: Any


_WORLD_FIELDS = dict(ask_raw_orders="asks", bid_raw_orders="bids", trades="trades", init_time="init_time",
                     window_index="window_index", max_steps_in_episode="max_steps", start_index="start_index",
                     step_counter="step_counter", best_bids="best_bids", best_asks="best_asks", time="time",
                     order_id_counter="order_id_counter", mid_price="mid_price", delta_time="delta_time")


@dataclass
class MultiAgentState:
    """``world_state`` + one dict of leaves per agent type (MMEnvState / ExecEnvState fields).  ``arrays`` is the flat
    buffer table the kernel works on (the leaves are views of it)."""
    world_state: WorldState
    agent_states: List[dict]
    arrays: dict


@dataclass
class LoadedEnvState:
    """base_env.py:45-58 (batched: every leaf has a leading ``[B]``)."""
    ask_raw_orders: Any
    bid_raw_orders: Any
    trades: Any
    init_time: Any
    window_index: Any
    step_counter: Any
    max_steps_in_episode: Any
    start_index: Any


@dataclass
class LoadedEnvParams:
    message_data: Any
    book_data: Any
    init_states_array: dict


@dataclass
class MultiAgentParams:
    loaded_params: LoadedEnvParams
    agent_params: List[dict]


def _state_view(cfg, arrays):
    ws = WorldState(**{f: arrays[k] for f, k in _WORLD_FIELDS.items()})
    agents = []
    for t in range(cfg.n_agent_types):
        li, lf = abi.state_leaves(cfg.agent[t].kind)
        agents.append({n: arrays[f"a{t}_{n}"] for n in li + lf})
    return MultiAgentState(world_state=ws, agent_states=agents, arrays=arrays)


def build_reset_params(loaded: lobster.LoadedDay, world, replay_fn):
    """base_env.py:298-333 ``_init_states``: one precomputed reset state per window.  ``replay_fn(asks, bids, trades,
    msgs, start, n_msgs)`` scans the 2*depth initial limit orders of each window into an empty book, in place
    (product: the CUDA replay kernel; tests: the oracle).  Returns the numpy params dict (states.PARAMS)."""
    W = loaded.starts.shape[0]
    No, Nt, depth = world.nOrders, world.nTrades, world.book_depth
    first = loaded.msgs[loaded.starts]
    init_msgs = lobster.init_messages_from_books(loaded.books, first, depth, world.init_id).reshape(W * 2 * depth, 8)
    asks = np.full((W, No, 6), -1, np.int32)
    bids = np.full((W, No, 6), -1, np.int32)
    trades = np.full((W, Nt, 8), -1, np.int32)
    start = (np.arange(W, dtype=np.int64) * 2 * depth)
    asks, bids, trades = replay_fn(asks, bids, trades, np.ascontiguousarray(init_msgs), start, 2 * depth)
    if world.ep_type == "fixed_time":   # base:288-291: the window's nominal start on the time grid, not a message time
        span = world.day_end - world.day_start - world.episode_time + world.start_resolution
        t0 = (np.arange(W, dtype=np.int64) * world.start_resolution) % span + world.day_start
        init_time = np.stack([t0, np.zeros_like(t0)], axis=1).astype(np.int32)
    else:
        init_time = np.ascontiguousarray(first[:, 6:8], np.int32)
    return {
        "message_data": np.ascontiguousarray(loaded.msgs, np.int32),
        "init_asks": asks, "init_bids": bids, "init_trades": trades,
        "init_init_time": init_time,                                                     # base:288-291
        "init_max_steps": (loaded.max_msgs // world.n_data_msg_per_step + 1).astype(np.int32),  # base:323-324
        "init_start_index": loaded.starts.astype(np.int32),
    }


# ---- device-side replay -------------------------------------------------------------------------------------
def replay_books(book_cfg: abi.LobBookConfig, asks, bids, trades, msgs, start, n_msgs, best_out=None, cancel_u=None,
                 grouped=False):
    """``job.scan_through_entire_array`` for every book (torch CUDA tensors, in place).  ``cancel_u`` float32
    [B, n_msgs, 2]: the uniform draws of the random cancel fallbacks (cancel_mode 2/3, job:142-164).  ``grouped`` selects
    the 4-books-per-warp measurement variant (same results, slower)."""
    L = _lib.lib()
    r = states.pack_replay(asks, bids, trades, msgs, start, n_msgs, best_out, cancel_u)
    fn = L.lob_replay_launch_grouped if grouped else L.lob_replay_launch
    with _lib.on_device(asks.device):
        _lib.check(fn(C.byref(book_cfg), C.byref(r), asks.shape[0], _lib.current_stream_ptr(asks.device)),
                   "lob_replay_launch")


def l2_state(book_cfg: abi.LobBookConfig, asks, bids, n_levels):
    """``job.get_L2_state`` for every book -> int32 [B, 4*n_levels]."""
    import torch
    L = _lib.lib()
    out = torch.empty((asks.shape[0], 4 * n_levels), dtype=torch.int32, device=asks.device)
    p = lambda t: C.cast(t.data_ptr(), abi.p_i32)
    with _lib.on_device(asks.device):
        _lib.check(L.lob_l2_launch(C.byref(book_cfg), p(asks), p(bids), p(out), n_levels, asks.shape[0],
                                   _lib.current_stream_ptr(asks.device)), "lob_l2_launch")
    return out


def limit_only_book_config(book_cfg: abi.LobBookConfig) -> abi.LobBookConfig:
    """The book config for replaying streams that hold only limit orders (the reset-state precompute, base:245-281):
    no cancel is processed, so no random-cancel draws (cancel_mode 2/3) have to be supplied."""
    c = abi.LobBookConfig.from_buffer_copy(book_cfg)
    c.cancel_mode = min(int(c.cancel_mode), 1)
    return c


def _cuda_replay_fn(book_cfg, device):
    import torch
    book_cfg = limit_only_book_config(book_cfg)

    def fn(asks, bids, trades, msgs, start, n_msgs):
        ta, tb, tt = (torch.from_numpy(x).to(device) for x in (asks, bids, trades))
        tm, ts = torch.from_numpy(msgs).to(device), torch.from_numpy(start).to(device)
        replay_books(book_cfg, ta, tb, tt, tm, ts, n_msgs)
        torch.cuda.synchronize(device)
        return ta.cpu().numpy(), tb.cpu().numpy(), tt.cpu().numpy()
    return fn


class BaseLOBEnv:
    """base_env.py:84-411: data window bookkeeping + the pure book replay (``step_env`` without agents)."""

    def __init__(self, cfg: World_EnvironmentConfig, key=None, *, loaded=None, device="cuda", synth=None):
        self.cfg = cfg
        self.device = device
        self.book_cfg = book_config(cfg)
        self.n_data_msg_per_step = cfg.n_data_msg_per_step
        if cfg.ep_type not in ("fixed_steps", "fixed_time"):
            raise NotImplementedError('Use either "fixed_time" or "fixed_steps"')   # ldr:998
        self.loaded = loaded if loaded is not None else lobster.load_or_generate(cfg, device=device, **(synth or {}))
        self.n_windows = int(self.loaded.starts.shape[0])
        self.start_indeces, self.end_indeces = self.loaded.starts, self.loaded.ends
        self.max_messages_in_episode_arr = self.loaded.max_msgs
        self._params_np = build_reset_params(self.loaded, cfg, _cuda_replay_fn(self.book_cfg, device))
        self._params_dev = None

    def device_params(self):
        """The whole day + reset states, resident in HBM once."""
        if self._params_dev is None:
            import torch
            md = getattr(self.loaded, "msgs_device", None)     # preprocessed on the device: already resident
            keep = md is not None and md.device == torch.device(self.device) and md.is_contiguous()
            self._params_dev = {k: (md if (k == "message_data" and keep) else torch.from_numpy(v).to(self.device))
                                for k, v in self._params_np.items()}
        return self._params_dev

    @property
    def default_params(self) -> LoadedEnvParams:
        p = self.device_params()
        init = {k: p[k] for k in states.PARAMS if k.startswith("init_")}
        return LoadedEnvParams(message_data=p["message_data"], book_data=self.loaded.books, init_states_array=init)

    def replay(self, asks, bids, trades, start, n_msgs, best_out=None):
        """base_env.py:189-216 for a batch of books: scan ``n_msgs`` data messages from ``start[b]``."""
        replay_books(self.book_cfg, asks, bids, trades, self.device_params()["message_data"], start, n_msgs, best_out)

    # ---- the reference's own method pair (base_env.py:189-234), batched ----
    def reset_env(self, key=None, params: LoadedEnvParams = None, num_envs: int = 1, seed: int = 0):
        """base_env.py:218-227 -> (0, LoadedEnvState): every environment starts from the precomputed state of a data
        window -- ``cfg.window_selector`` if it is >= 0, else drawn uniformly (a torch CUDA generator seeded with
        ``seed`` stands in for ``jax.random.randint(key, ...)``)."""
        import torch
        p = self.device_params() if params is None else {**self.device_params(), **params.init_states_array}
        if self.cfg.window_selector >= 0:
            w = torch.full((num_envs,), int(self.cfg.window_selector), dtype=torch.int64, device=self.device)
        else:
            g = torch.Generator(device=self.device)
            g.manual_seed(int(seed))
            w = torch.randint(0, self.n_windows, (num_envs,), generator=g, device=self.device)
        state = LoadedEnvState(
            ask_raw_orders=p["init_asks"][w].clone(), bid_raw_orders=p["init_bids"][w].clone(),
            trades=p["init_trades"][w].clone(), init_time=p["init_init_time"][w].clone(),
            window_index=w.to(torch.int32), step_counter=torch.zeros(num_envs, dtype=torch.int32, device=self.device),
            max_steps_in_episode=p["init_max_steps"][w].clone(), start_index=p["init_start_index"][w].clone())
        return 0, state

    def step_env(self, key, state: "LoadedEnvState", action=None, params: LoadedEnvParams = None):
        """base_env.py:189-216 -> (obs 0, state, reward 0, done [B] bool, {"info": 0}): the next ``n_data_msg_per_step``
        data messages of every environment go through the book (one ``lob_replay_launch``; the state's buffers are
        updated in place), ``done`` = the last message is ``episode_time`` seconds past ``init_time`` (base:232-234)."""
        import torch
        msgs = self.device_params()["message_data"] if params is None else params.message_data
        Nd, M = self.n_data_msg_per_step, msgs.shape[0]
        off = (state.start_index.to(torch.int64) + Nd * state.step_counter.to(torch.int64)).clamp(0, max(M - Nd, 0))
        if self.cfg.ep_type == "fixed_time":     # base:358-368: messages past the episode end keep only their time stamp
            idx = off[:, None] + torch.arange(Nd, device=off.device)[None, :]
            sl = msgs[idx]                                                           # [B, Nd, 8]
            late = sl[:, :, 6] >= (state.init_time[:, 0] + self.cfg.episode_time)[:, None]
            sl = torch.where(late[:, :, None] & (torch.arange(8, device=off.device) < 6)[None, None, :],
                             torch.zeros_like(sl), sl).reshape(-1, 8).contiguous()
            start = torch.arange(off.shape[0], dtype=torch.int64, device=off.device) * Nd
            replay_books(self.book_cfg, state.ask_raw_orders, state.bid_raw_orders, state.trades, sl, start, Nd)
        else:
            replay_books(self.book_cfg, state.ask_raw_orders, state.bid_raw_orders, state.trades, msgs, off, Nd)
        t_last = msgs[off + Nd - 1, 6]
        state.step_counter += 1
        done = (t_last - state.init_time[:, 0]) >= self.cfg.episode_time
        return 0, state, 0, done, {"info": 0}


class MARLEnv:
    """marl_env.py:45-805, batched.  ``num_envs`` plays the role of the trainer's vmap width."""

    def __init__(self, key, multi_agent_config: MultiAgentConfig, *, num_envs=1, loaded=None, device="cuda",
                 synth=None, seed=0):
        import torch
        self.multi_agent_config = multi_agent_config
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.num_agents = sum(multi_agent_config.number_of_agents_per_type)
        self.base_env = BaseLOBEnv(multi_agent_config.world_config, key, loaded=loaded, device=device, synth=synth)
        self.list_of_agents_configs = list(multi_agent_config.dict_of_agents_configs.values())
        self.type_names = [c.short_name for c in self.list_of_agents_configs]
        self.instance_list = self.list_of_agents_configs  # the trainer only takes len() / indexes alongside configs
        self.cfg = to_step_config(multi_agent_config, self.base_env.n_windows, self.base_env.loaded.msgs.shape[0])
        self.num_msgs_per_step = num_msgs_per_step(self.cfg)
        self.num_action_msgs_per_step_by_all_agents = num_action_msgs(self.cfg)
        self.action_spaces = [MultiDiscrete([c.fixed_quant_value] * c.n_actions)   # exec_env.py:2167-2171
                              if getattr(c, "action_space", "") == "fixed_prices" and isinstance(c, Execution_EnvironmentConfig)
                              else Discrete(c.n_actions) for c in self.list_of_agents_configs]
        self.observation_spaces = [
            Box(-1000 if isinstance(c, MarketMaking_EnvironmentConfig) else -10000,
                1000 if isinstance(c, MarketMaking_EnvironmentConfig) else 10000,
                (abi.obs_dim(self.cfg.agent[i].kind, self.cfg.agent[i].observation_space,
                             bool(self.cfg.ep_type_fixed_time)),), np.float32)
            for i, c in enumerate(self.list_of_agents_configs)]
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._counter_dev = None
        self._cache = {}

    # -- API of the reference ------------------------------------------------------------------------------
    def action_space(self):
        return self.action_spaces

    def observation_space(self):
        return self.observation_spaces

    @property
    def default_params(self) -> MultiAgentParams:
        agent_params = []
        for t in range(self.cfg.n_agent_types):
            a = self.cfg.agent[t]
            agent_params.append({"trader_id": np.arange(a.trader_id_start, a.trader_id_start - a.n_agents, -1)})
        return MultiAgentParams(loaded_params=self.base_env.default_params, agent_params=agent_params)

    # -- per-buffer-table caches: the state leaves are updated in place, so pointers and views never change ----------
    def _bound(self, arrays):
        key = id(arrays)
        hit = self._cache.get(key)
        if hit is None or hit[0] is not arrays:
            bufs = states.pack_buffers(self.cfg, arrays, self.base_env.device_params())
            T = self.cfg.n_agent_types
            obs = [arrays[f"obs{t}"] for t in range(T)]
            rewards = [arrays[f"reward{t}"] for t in range(T)]
            import torch
            dones = {"__all__": arrays["done_all"].view(torch.bool),
                     "agents": [arrays[f"done_agents{t}"].view(torch.bool) for t in range(T)]}
            hit = (arrays, bufs, obs, rewards, dones, self.unpack_info(arrays), _state_view(self.cfg, arrays))
            self._cache = {key: hit}
        return hit

    def _draw(self, arrays, bufs):
        """The PRNG products of one step (see module docstring): lob_draw_launch, counter-based, on the device."""
        if self._counter_dev is None:
            import torch
            self._counter_dev = torch.ones(1, dtype=torch.int64, device=self.device)   # device-resident: graph-capturable
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().lob_draw_launch_dev(C.byref(self.cfg), C.byref(bufs), self.num_envs,
                                                      int(self.multi_agent_config.world_config.window_selector), self._seed,
                                                      C.c_void_p(self._counter_dev.data_ptr()),
                                                      _lib.current_stream_ptr(self.device)), "lob_draw_launch_dev")

    def reset(self, key=None, params: MultiAgentParams = None, arrays=None, draw=True):
        """marl_env.py:764 -> (obs list [B,n_i,d_i], MultiAgentState)."""
        if params is None:
            raise ValueError("Params must be provided to reset the environment.")
        L = _lib.lib()
        if arrays is None:
            arrays = states.alloc_torch(self.cfg, self.num_envs, self.device)
        _, bufs, obs, _, _, _, state = self._bound(arrays)
        if draw:
            self._draw(arrays, bufs)
        with _lib.on_device(self.device):
            _lib.check(L.lob_reset_launch(C.byref(self.cfg), C.byref(bufs), self.num_envs,
                                          _lib.current_stream_ptr(self.device)), "lob_reset_launch")
        return obs, state

    def step(self, key, state: MultiAgentState, actions, params: MultiAgentParams = None, draw=True):
        """marl_env.py:776 -> (obs, state, rewards, dones, infos); ``state``'s buffers are donated (updated in place)
        and the returned tensors are views of them."""
        L = _lib.lib()
        arrays = state.arrays
        _, bufs, obs, rewards, dones, info, view = self._bound(arrays)
        if len(actions) != self.cfg.n_agent_types:
            raise ValueError(f"actions: one entry per agent type expected ({self.cfg.n_agent_types}), got {len(actions)}")
        for t, a in enumerate(actions):
            dst = arrays[f"actions{t}"]
            dst.copy_(a.reshape(dst.shape), non_blocking=True)
        if draw:
            self._draw(arrays, bufs)
        with _lib.on_device(self.device):
            _lib.check(L.lob_step_launch(C.byref(self.cfg), C.byref(bufs), self.num_envs,
                                         _lib.current_stream_ptr(self.device)), "lob_step_launch")
        return obs, view, rewards, dones, info

    def rollout(self, state: MultiAgentState, actions, n_steps: int, params: MultiAgentParams = None, draw=True):
        """``n_steps`` env steps in ONE call of the C ABI (``lob_rollout_launch``: the piped step ``n_steps`` times over,
        trajectory rows written in place; for deep books one launch of the fused kernel with the books resident in shared
        memory for the whole rollout) -- the trainer's ``jit(lax.scan(vmap(env.step)))`` (ippo_rnn_JAXMARL.py:616-661) with
        the actions given up front (a pre-sampled / open-loop policy: Speed_test.py:165-214).  ``actions``: one int32 CUDA
        tensor ``[T, B, n_i]`` per agent type.  Equal, leaf for leaf, to ``n_steps`` calls of ``step`` (same PRNG counter
        sequence).  Returns (traj, state) with ``traj = {"obs": [[T,B,n_i,d_i] per type], "reward": [[T,B,n_i]],
        "done_agents": [[T,B,n_i] uint8], "done": [T,B] uint8}``; the state, the last step's outputs and info are in
        ``state.arrays`` as after ``step``."""
        import torch
        T_, B, dev = int(n_steps), self.num_envs, self.device
        nt = self.cfg.n_agent_types
        if len(actions) != nt:
            raise ValueError(f"actions: one [T, B, n_i] tensor per agent type expected ({nt}), got {len(actions)}")
        arrays = state.arrays
        _, bufs, _, _, _, _, view = self._bound(arrays)
        key = ("roll", T_, id(arrays))
        io = self._roll_cache.get(key) if hasattr(self, "_roll_cache") else None
        if io is None:
            z = lambda name, dt: torch.empty((T_,) + tuple(arrays[name].shape), dtype=dt, device=dev)
            io = {"perm": z("perm", torch.int32), "reset_window": z("reset_window", torch.int32),
                  "reset_is_sell": z("reset_is_sell", torch.int32),
                  "obs": [z(f"obs{t}", torch.float32) for t in range(nt)],
                  "reward": [z(f"reward{t}", torch.float32) for t in range(nt)],
                  "done_agents": [z(f"done_agents{t}", torch.uint8) for t in range(nt)], "done": z("done_all", torch.uint8)}
            if "cancel_u" in arrays:
                io["cancel_u"] = z("cancel_u", torch.float32)
            io["draw_bufs"] = []      # step k's rows of the draw buffers, packed once (pointers never change)
            for k in range(T_):
                row = {**arrays, "perm": io["perm"][k], "reset_window": io["reset_window"][k],
                       "reset_is_sell": io["reset_is_sell"][k]}
                if "cancel_u" in io:
                    row["cancel_u"] = io["cancel_u"][k]
                io["draw_bufs"].append(states.pack_buffers(self.cfg, row, self.base_env.device_params()))
            self._roll_cache = {key: io}
        acts = []
        for t, a in enumerate(actions):
            want = (T_,) + tuple(arrays[f"actions{t}"].shape)
            a = a.reshape(want).to(torch.int32).contiguous()
            if a.device != arrays["asks"].device:
                raise ValueError("actions must live on the environment's device")
            acts.append(a)
        if draw:      # the same counter sequence n_steps calls of step() would consume
            for k in range(T_):
                self._draw(arrays, io["draw_bufs"][k])
        rb = abi.LobRolloutBuffers()
        rb.n_steps, rb.batch = T_, B
        p32, pf, p8 = (lambda t: C.cast(t.data_ptr(), abi.p_i32)), (lambda t: C.cast(t.data_ptr(), abi.p_f32)), \
            (lambda t: C.cast(t.data_ptr(), abi.p_u8))
        for t in range(nt):
            rb.actions[t] = p32(acts[t]); rb.obs[t] = pf(io["obs"][t]); rb.reward[t] = pf(io["reward"][t])
            rb.done_agents[t] = p8(io["done_agents"][t])
        rb.perm, rb.reset_window, rb.reset_is_sell = p32(io["perm"]), p32(io["reset_window"]), p32(io["reset_is_sell"])
        if "cancel_u" in io:
            rb.cancel_u = pf(io["cancel_u"])
        rb.done_all = p8(io["done"])
        with _lib.on_device(dev):
            _lib.check(_lib.lib().lob_rollout_launch(C.byref(self.cfg), C.byref(bufs), C.byref(rb), B,
                                                     _lib.current_stream_ptr(dev)), "lob_rollout_launch")
        self._roll_keepalive = (acts, rb)     # the launch is asynchronous: keep the operands alive
        traj = {"obs": io["obs"], "reward": io["reward"], "done_agents": io["done_agents"], "done": io["done"]}
        return traj, view

    def capture_step(self, state: MultiAgentState, actions, params: MultiAgentParams = None, pre=None, post=None):
        """One step as a CUDA graph: [pre()] + actions copy + PRNG draw + step launches [+ post()] captured on the current
        stream (the rollout-fusion row of SURVEY 8f-4: one graph launch per step instead of 4+ launches and their Python).
        ``actions`` are the device tensors the graph reads every replay; ``pre`` / ``post`` are optional callables that
        enqueue copies (e.g. pinned host -> ``actions``, results -> pinned host).  Returns (graph, outputs) where outputs
        is what ``step`` returns (views of the in-place buffers)."""
        import torch
        with torch.cuda.device(self.device):
            snap = self._snapshot(state.arrays)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):                  # warm-up outside capture (allocations, attribute queries)
                if pre:
                    pre()
                out = self.step(None, state, actions, params)
                if post:
                    post(out)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._restore(state.arrays, snap)           # the warm-up transition is undone: the caller's episode is untouched
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if pre:
                    pre()
                out = self.step(None, state, actions, params)
                if post:
                    post(out)
        return g, out

    def _snapshot(self, arrays):
        """Copies of every leaf of the buffer table and of the device PRNG counter (the warm-up of a graph capture runs a
        real step on them)."""
        if self._counter_dev is None:
            import torch
            self._counter_dev = torch.ones(1, dtype=torch.int64, device=self.device)
        return {k: v.clone() for k, v in arrays.items()}, self._counter_dev.clone()

    def _restore(self, arrays, snap):
        leaves, counter = snap
        for k, v in leaves.items():
            arrays[k].copy_(v)
        self._counter_dev.copy_(counter)

    def capture_rollout(self, state: MultiAgentState, policy, n_steps: int, params: MultiAgentParams = None):
        """``n_steps`` env steps with the policy in between as ONE CUDA graph -- the ``jit(lax.scan(vmap(env.step)))`` of
        the trainer / Speed_test.py (ippo_rnn_JAXMARL.py:616-661, Speed_test.py:185-224) with the state resident in HBM
        for the whole rollout.  ``policy(t, obs) -> list of int32 action tensors`` is called while capturing, so it must
        consist of capturable device work (torch ops, e.g. a network forward + sampling with a CUDA generator).
        Returns (graph, traj) with ``traj = {"obs": [list per type of [T,B,n_i,d_i]], "reward": [...[T,B,n_i]],
        "done": [T,B] uint8, "done_agents": [...[T,B,n_i]]}``: step t's outputs land in row t on every replay."""
        import torch
        T = self.cfg.n_agent_types
        arrays = state.arrays
        B = self.num_envs
        traj = {"obs": [torch.empty((n_steps,) + tuple(arrays[f"obs{t}"].shape), dtype=torch.float32, device=self.device)
                        for t in range(T)],
                "reward": [torch.empty((n_steps,) + tuple(arrays[f"reward{t}"].shape), dtype=torch.float32,
                                       device=self.device) for t in range(T)],
                "done_agents": [torch.empty((n_steps,) + tuple(arrays[f"done_agents{t}"].shape), dtype=torch.uint8,
                                            device=self.device) for t in range(T)],
                "done": torch.empty((n_steps, B), dtype=torch.uint8, device=self.device)}

        def body(steps):
            obs = [arrays[f"obs{t}"] for t in range(T)]
            for k in range(steps):
                acts = policy(k, obs)
                obs, _, rewards, dones, _ = self.step(None, state, acts, params)
                for t in range(T):
                    traj["obs"][t][k].copy_(obs[t])
                    traj["reward"][t][k].copy_(rewards[t])
                    traj["done_agents"][t][k].copy_(arrays[f"done_agents{t}"])
                traj["done"][k].copy_(arrays["done_all"])

        with torch.cuda.device(self.device):
            snap = self._snapshot(arrays)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body(1)                                   # warm-up outside capture: runs ONE step ...
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._restore(arrays, snap)                   # ... which is undone (state, outputs and the PRNG counter)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body(n_steps)
        return g, traj

    def unpack_info(self, arrays):
        """Packed info columns -> the reference's dict keys (marl:624-639, mm:2695-2730, exe:1809-1829)."""
        wi, wf = arrays["info_world_i32"], arrays["info_world_f32"]
        world = {k: wi[:, j] for j, k in enumerate(abi.WINFO_I32) if k not in ("time_s", "time_ns")}
        world["time"] = wi[:, 2:4]
        world.update({k: wf[:, j] for j, k in enumerate(abi.WINFO_F32)})
        agents = []
        for t in range(self.cfg.n_agent_types):
            ki, kf = abi.info_cols(self.cfg.agent[t].kind)
            d = {k: arrays[f"info_i32_{t}"][..., j] for j, k in enumerate(ki)}
            d.update({k: arrays[f"info_f32_{t}"][..., j] for j, k in enumerate(kf)})
            agents.append(d)
        return {"world": world, "agents": agents}
