"""The XLA-FFI boundary (jaxmarl_hft_b200/ffi_stub.py + csrc/lob_ffi.cc) without JAX: the operand / result offset tables
the stub hands to the table-driven handler must fill every pointer of ``LobStepBuffers`` that ``states.pack_buffers`` fills
-- each exactly once per side -- and nothing else; state leaves are aliased operand -> result."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

import helpers as H
from jaxmarl_hft_b200 import abi, config as Cfg, ffi_stub, states


def _nonnull_offsets(bufs):
    raw = bytes(bufs)
    P = C.sizeof(C.c_void_p)
    return {o for o in range(0, len(raw), P) if int.from_bytes(raw[o:o + P], "little") != 0}


def _cfgs():
    mac = H.load_mac("2_player_fq_fqc")
    yield "2player", Cfg.to_step_config(mac, 62, 400000)
    yield "exec", Cfg.to_step_config(H.load_mac("exec_longrun_fixed_quants_complex"), 62, 400000)
    yield "deep", Cfg.to_step_config(H.load_mac("hetero_deep_book"), 62, 400000)            # workspace leaves
    yield "cancel3", Cfg.to_step_config(H.load_mac("2_player_fq_fqc", cancel_mode=3), 62, 400000)   # cancel_u input
    ex = dataclasses.replace(mac.dict_of_agents_configs["Execution"], action_space="fixed_prices", n_actions=3)
    yield "vector_actions", Cfg.to_step_config(H.with_agents(mac, {"MarketMaking": mac.dict_of_agents_configs["MarketMaking"],
                                                                   "Execution": ex}, [2, 3]), 62, 400000)


@pytest.mark.parametrize("name,cfg", list(_cfgs()), ids=[n for n, _ in _cfgs()])
def test_step_table_fills_every_pointer_once(name, cfg):
    arrays = states.alloc_numpy(cfg, 3)
    params = {p: np.zeros((2, 8), np.int32) for p in states.PARAMS}
    want = _nonnull_offsets(states.pack_buffers(cfg, arrays, params))
    tab = ffi_stub.step_table(cfg)
    a, r = tab.arg_off.tolist(), tab.ret_off.tolist()
    assert len(set(a)) == len(a) and len(set(r)) == len(r), "a pointer field is filled twice"
    assert set(a) | set(r) == want, (sorted(want - (set(a) | set(r))), sorted((set(a) | set(r)) - want))
    n_state = len(states.state_names(cfg))
    assert tab.aliases == {i: i for i in range(n_state)}
    assert tab.names_in[:n_state] == tab.names_out[:n_state] == states.state_names(cfg)
    assert set(a) & set(r) == set(a[:n_state]), "only the state leaves are both operand and result"
    specs = states.leaf_specs(cfg, 3)
    assert all(specs[n][2] in ("o", "w") for n in tab.names_out[n_state:])
    assert all(n in states.PARAMS or specs[n][2] in ("s", "i") for n in tab.names_in)
    # the handler rejects offsets outside the struct
    assert max(a + r) + C.sizeof(C.c_void_p) <= C.sizeof(abi.LobStepBuffers) and min(a + r) >= 0


def test_reset_table_is_a_subset():
    cfg = Cfg.to_step_config(H.load_mac("2_player_fq_fqc"), 62, 400000)
    step, reset = ffi_stub.step_table(cfg), ffi_stub.step_table(cfg, reset_only=True)
    assert set(reset.names_in) < set(step.names_in) and set(reset.names_out) < set(step.names_out)
    assert not any(n.startswith("actions") or n == "message_data" for n in reset.names_in)
    assert [n for n in reset.names_out if n.startswith("obs")] == ["obs0", "obs1"]


def test_replay_table_matches_pack_replay():
    B, T = 4, 10
    a = np.zeros((B, 100, 6), np.int32); t = np.zeros((B, 100, 8), np.int32)
    msgs = np.zeros((50, 8), np.int32); start = np.zeros(B, np.int64)
    best = np.zeros((B, 4), np.int32); cu = np.zeros((B, T, 2), np.float32)
    r = states.pack_replay(a, a.copy(), t, msgs, start, T, best, cu)
    tab = ffi_stub.replay_table(with_best=True, with_cancel_u=True)
    P = C.sizeof(C.c_void_p)
    raw = bytes(r)
    scalar = {abi.LobReplayBuffers.n_msgs_total.offset, abi.LobReplayBuffers.n_msgs.offset}   # attributes, not operands
    want = {o for o in range(0, len(raw), P) if int.from_bytes(raw[o:o + P], "little") != 0} - scalar
    assert set(tab.arg_off.tolist()) | set(tab.ret_off.tolist()) == want


def test_ffi_needs_jax_and_says_so():
    try:
        import jax  # noqa: F401
        pytest.skip("jax is installed: the binding is exercised by its own GPU test")
    except ImportError:
        with pytest.raises(RuntimeError, match="jax"):
            ffi_stub.build_ffi()


def test_lob_ffi_cc_compiles_and_fills_the_struct(tmp_path):
    """csrc/lob_ffi.cc against a MOCK of xla/ffi/api/ffi.h (tests/ffi_mock: same names and call shapes; the real header
    ships with jaxlib) linked with csrc/liblobstep.so: the table-driven Fill() puts every fake operand / result pointer at
    the offset the stub's table names, and a bad table is an error, not a wild store."""
    import os
    import shutil
    import subprocess
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "jaxmarl-hft_b200", "csrc")
    if not (shutil.which("g++") and os.path.exists(os.path.join(csrc, "liblobstep.so"))):
        pytest.skip("needs g++ and the built liblobstep.so")
    cfg = Cfg.to_step_config(H.load_mac("2_player_fq_fqc"), 62, 400000)
    tab = ffi_stub.step_table(cfg)
    drv = tmp_path / "drv.cc"
    drv.write_text(textwrap.dedent(f'''
        #include <cstdio>
        #include <cstring>
        #include "{os.path.join(csrc, "lob_ffi.cc")}"
        int main() {{
          const int arg_off[] = {{{", ".join(map(str, tab.arg_off.tolist()))}}};
          const int ret_off[] = {{{", ".join(map(str, tab.ret_off.tolist()))}}};
          const size_t na = sizeof(arg_off) / 4, nr = sizeof(ret_off) / 4;
          ffi::RemainingArgs args; ffi::RemainingRets rets;
          for (size_t i = 0; i < na; ++i) args.ptrs.push_back(reinterpret_cast<void*>(0x1000 + 16 * i));
          for (size_t i = 0; i < nr; ++i) rets.ptrs.push_back(reinterpret_cast<void*>(0x900000 + 16 * i));
          LobStepBuffers b;
          ffi::Error e = Fill(&b, ffi::Span<const int32_t>(arg_off, na), ffi::Span<const int32_t>(ret_off, nr), args, rets);
          if (e.failure()) {{ std::printf("fill failed: %s\\n", e.message().c_str()); return 1; }}
          const char* raw = reinterpret_cast<const char*>(&b);
          for (size_t i = 0; i < nr; ++i) {{   // results win where a state leaf is both operand and result (aliased)
            void* p; std::memcpy(&p, raw + ret_off[i], sizeof(p));
            if (p != rets.ptrs[i]) {{ std::printf("result %zu misplaced\\n", i); return 2; }}
          }}
          if (b.message_data == nullptr || b.asks != rets.ptrs[0]) return 3;
          const int bad[] = {{(int)sizeof(LobStepBuffers)}};
          ffi::RemainingArgs one; one.ptrs.push_back(nullptr);
          ffi::RemainingRets none;
          if (!Fill(&b, ffi::Span<const int32_t>(bad, 1), ffi::Span<const int32_t>(bad, 0), one, none).failure()) return 4;
          if (!Fill(&b, ffi::Span<const int32_t>(arg_off, 2), ffi::Span<const int32_t>(ret_off, 0), one, none).failure()) return 5;
          std::printf("ok %zu %zu\\n", na, nr);
          return LobStep() && LobReplay() ? 0 : 6;
        }}'''))
    exe = tmp_path / "drv"
    cmd = ["g++", "-std=c++17", "-O1", f"-I{os.path.join(root, 'tests', 'ffi_mock')}", f"-I{os.path.join(root, 'include')}",
           "-I/usr/local/cuda/include", str(drv), f"-L{csrc}", "-llobstep", f"-Wl,-rpath,{csrc}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout, r.stderr)
