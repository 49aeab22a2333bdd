"""The C-ABI library (csrc/liblobstep.so) without a GPU: it loads, exports every symbol include/lobstep.h declares, its
structs match the ctypes mirror, and its host-side validation rejects bad calls before touching the device."""
import ctypes as C
import numpy as np
import os
import re

import pytest

from jaxmarl_hft_b200 import _lib, abi, config as Cfg
import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lobstep.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_\*\s]*?\b(lob_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("lob_step_launch", "lob_reset_launch", "lob_replay_launch", "lob_l2_launch", "lob_host_replay_run",
                 "lob_last_error", "lob_abi_version", "lob_launch_count"):
        assert must in names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _declared_functions():
        assert hasattr(L, name), f"liblobstep.so does not export {name}"


def test_struct_sizes_match_ctypes_mirror():
    L = _lib.lib()
    abi.check_sizes(L)
    assert L.lob_abi_version() == abi.LOB_ABI_VERSION


def test_derived_sizes_match_reference_formula():
    """marl_env.py:85-94: N = Nd + sum_i n_i * num_messages_by_agent_i."""
    L = _lib.lib()
    for name, N, n_act in (("2_player_fq_fqc", 112, 6), ("exec_longrun_fixed_quants_complex", 108, 4),
                           ("hetero_deep_book", 136, 18)):
        mac = H.load_mac(name)
        cfg = Cfg.to_step_config(mac, 4, 30000)
        assert L.lob_num_msgs_per_step(C.byref(cfg)) == N == Cfg.num_msgs_per_step(cfg)
        assert L.lob_num_action_msgs(C.byref(cfg)) == n_act == Cfg.num_action_msgs(cfg)
        assert L.lob_num_cancel_msgs(C.byref(cfg)) == n_act


def test_bad_calls_fail_loudly_before_the_device():
    """Validation happens on the host: these return an error code with a message and never reach CUDA."""
    L = _lib.lib()
    mac = H.load_mac("2_player_fq_fqc")
    cfg = Cfg.to_step_config(mac, 4, 30000)
    bufs = abi.LobStepBuffers()     # all NULL
    assert L.lob_step_launch(C.byref(cfg), C.byref(bufs), 8, None) == abi.LOB_E_INVALID
    assert b"null" in L.lob_last_error()
    bad = Cfg.to_step_config(mac, 4, 30000)
    bad.book.n_orders = 100000
    assert L.lob_step_launch(C.byref(bad), C.byref(bufs), 8, None) == abi.LOB_E_INVALID
    bad = Cfg.to_step_config(mac, 4, 30000)
    bad.book.cancel_mode = 7
    assert L.lob_reset_launch(C.byref(bad), C.byref(bufs), 8, None) == abi.LOB_E_INVALID
    assert b"cancel_mode" in L.lob_last_error()
    bad = Cfg.to_step_config(mac, 4, 30000)
    bad.ep_type_fixed_time = 1       # supported since ABI 5: validation proceeds to the (null) buffer table
    assert L.lob_step_launch(C.byref(bad), C.byref(bufs), 8, None) == abi.LOB_E_INVALID
    assert L.lob_obs_dim(C.byref(bad), 0) == 2 and L.lob_obs_dim(C.byref(bad), 1) == 15   # MM basic, EXE engineered
    rb = abi.LobReplayBuffers()
    bc = Cfg.book_config(mac.world_config)
    assert L.lob_replay_launch(C.byref(bc), C.byref(rb), 8, None) == abi.LOB_E_INVALID
    bc.cancel_mode = 3               # random cancel fallbacks: the host-buffer handle has no draws to offer
    assert not L.lob_host_replay_create(C.byref(bc), 4, 100, 0)
    assert b"cancel_mode" in L.lob_last_error()


def test_host_config_rejects_what_the_reference_rejects():
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    agents["MarketMaking"] = dataclasses.replace(agents["MarketMaking"], reward_function="nonsense")
    with pytest.raises(ValueError):
        Cfg.to_step_config(H.with_agents(mac, agents, [1, 1]), 4, 30000)
    with pytest.raises(ValueError):
        dataclasses.replace(agents["MarketMaking"], action_space="fixed_quants", tenth_action="bogus")


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under jaxmarl-hft_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "jaxmarl-hft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "lob_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/config/env_configs"), reason="reference not mounted")
def test_shipped_reference_env_configs_lower():
    """11 of the reference's 12 env JSONs lower to the POD config; the stale twelfth is rejected loudly."""
    import glob
    ok, bad = [], []
    for f in sorted(glob.glob("/root/reference/config/env_configs/*.json")):
        try:
            Cfg.to_step_config(Cfg.load_config_from_file(f), 4, 30000)
            ok.append(os.path.basename(f))
        except NotImplementedError:
            bad.append(os.path.basename(f))
    assert bad == ["exec_discrete_steps.json"] and len(ok) == 11


def test_field_offsets_agree_with_pack_buffers():
    """states.field_offset (the table of the table-driven XLA-FFI binding) against states.pack_buffers:
    storing each leaf's address at its offset reproduces the packed struct byte for byte."""
    from jaxmarl_hft_b200 import states
    mac = H.load_mac("hetero_deep_book", cancel_mode=3)
    cfg = Cfg.to_step_config(mac, 6, 30000)
    arrays = states.alloc_numpy(cfg, 3)
    params = {k: np.zeros(8, np.int32) for k in states.PARAMS}
    packed = states.pack_buffers(cfg, arrays, params)
    mine = abi.LobStepBuffers()
    raw = (C.c_char * C.sizeof(mine)).from_address(C.addressof(mine))
    for name, arr in list(arrays.items()) + list(params.items()):
        off = states.field_offset(cfg, name)
        C.c_void_p.from_address(C.addressof(mine) + off).value = arr.ctypes.data
    assert bytes(raw) == bytes((C.c_char * C.sizeof(packed)).from_address(C.addressof(packed)))
    with pytest.raises(KeyError):
        states.field_offset(cfg, "nonsense")


def test_graft_entry_build_runs():
    """The driver's build check: __graft_entry__.build() compiles (or finds up to date) the CUDA library and the oracle,
    loads the library and checks its ABI version."""
    import importlib
    g = importlib.import_module("__graft_entry__")
    g.build()


def test_reference_pytree_flatten_roundtrip():
    """states.flatten / unflatten / flatten_params on objects shaped like the reference's flax structs
    (StatesandParams.py): the helpers the INTEGRATION.md stub uses to feed the table-driven FFI call."""
    import dataclasses
    from types import SimpleNamespace as NS
    from jaxmarl_hft_b200 import states
    mac = H.load_mac("2_player_fq_fqc")
    cfg = Cfg.to_step_config(mac, 6, 30000)
    arrays = states.alloc_numpy(cfg, 2)

    @dataclasses.dataclass(frozen=True)
    class Rec:                       # stands in for a flax struct: attribute access + .replace
        d: dict
        def __getattr__(self, k):
            return self.d[k]
        def replace(self, **kw):
            return Rec({**self.d, **kw})
    ws = Rec({f: arrays[l] for l, f in states.WORLD_FIELD_OF_LEAF.items()})
    ags = []
    for t in range(cfg.n_agent_types):
        li, lf = abi.state_leaves(cfg.agent[t].kind)
        ags.append(Rec({n: arrays[f"a{t}_{n}"] for n in li + lf}))
    state = Rec({"world_state": ws, "agent_states": ags})
    flat = states.flatten(cfg, state)
    want = [k for k, (_, _, role) in states.leaf_specs(cfg, 2).items() if role == "s"]
    assert sorted(flat) == sorted(want) and all(flat[k] is arrays[k] for k in want)
    new = {k: v + 1 for k, v in flat.items()}
    st2 = states.unflatten(cfg, state, new)
    assert (st2.world_state.ask_raw_orders == arrays["asks"] + 1).all()
    assert (st2.agent_states[1].quant_executed == arrays["a1_quant_executed"] + 1).all()
    assert (state.world_state.ask_raw_orders == arrays["asks"]).all()          # functional
    init = NS(ask_raw_orders=1, bid_raw_orders=2, trades=3, init_time=4, max_steps_in_episode=5, start_index=6,
              window_index=7, step_counter=8)
    p = states.flatten_params(NS(loaded_params=NS(message_data=0, book_data=None, init_states_array=init)))
    assert list(p) == list(states.PARAMS) and p["init_trades"] == 3 and p["init_start_index"] == 6
