"""The jax.ffi side of the boundary: the operand / result tables of the XLA custom calls in csrc/lob_ffi.cc, the build of
that handler library where JAX is installed, and ``env.step`` / ``env.reset`` as ``jax.ffi.ffi_call``s a maintainer drops
into ``gymnax_exchange/jaxen/marl_env.py`` (replacing marl_env.py:764-804; the trainer's ``vmap(env.step)`` at
ippo_rnn_JAXMARL.py:616-618 inside the ``lax.scan`` of :661 is untouched).

Nothing here needs JAX to be IMPORTED: the tables are plain Python over ``states.leaf_specs`` / ``states.field_offset``
(tests/test_ffi_tables.py checks that they cover every pointer of ``LobStepBuffers`` exactly once); ``build_ffi`` and the
``ffi_*`` functions import ``jax`` when called and raise a clear error where it is absent (this image)."""
import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import List

import numpy as np

from . import abi, states

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
FFI_SO = os.path.join(_CSRC, "liblob_ffi.so")
PRNG_INPUTS = ("perm", "reset_window", "reset_is_sell", "cancel_u")


@dataclass(frozen=True)
class CallTable:
    """One custom call: operand names in call order, result names in result order, the byte offset of the
    ``LobStepBuffers`` pointer each fills, and ``input_output_aliases`` (operand index -> result index)."""
    names_in: List[str]
    names_out: List[str]
    arg_off: np.ndarray
    ret_off: np.ndarray
    aliases: dict


def step_table(cfg: abi.LobStepConfig, reset_only: bool = False) -> CallTable:
    """Operands: the state leaves (``MultiAgentState`` in ``states.flatten`` order), the step's inputs (actions, the
    PRNG products the reference draws inside the step: marl_env.py:294-295, base_env.py:222-225, exec_env.py:221) and the
    params (day tensor + stacked reset states, unbatched).  Results: the state leaves again (aliased: updated in place),
    then the outputs (obs, reward, dones, packed info) and the scratch workspace."""
    specs = states.leaf_specs(cfg, 1)
    state = [n for n, (_, _, role) in specs.items() if role == "s"]
    inputs = [n for n, (_, _, role) in specs.items() if role == "i" and not (reset_only and n.startswith("actions"))]
    outputs = [n for n, (_, _, role) in specs.items() if role in ("o", "w")]
    if reset_only:   # lob_reset_launch writes the state and the observations only (marl_env.py:130-207)
        outputs = [n for n in outputs if n.startswith("obs")]
    params = [p for p in states.PARAMS if not (reset_only and p == "message_data")]
    names_in = state + inputs + params
    names_out = state + outputs
    off = lambda names: np.array([states.field_offset(cfg, n) for n in names], np.int32)
    return CallTable(names_in, names_out, off(names_in), off(names_out), {i: i for i in range(len(state))})


def replay_table(with_best=False, with_cancel_u=False) -> CallTable:
    """lob_replay: asks / bids / trades in place, msgs, start [, cancel_u] -> [best_out]."""
    F = abi.LobReplayBuffers
    names_in = ["asks", "bids", "trades", "msgs", "start"] + (["cancel_u"] if with_cancel_u else [])
    names_out = ["asks", "bids", "trades"] + (["best_out"] if with_best else [])
    off = lambda names: np.array([getattr(F, n).offset for n in names], np.int32)
    return CallTable(names_in, names_out, off(names_in), off(names_out), {0: 0, 1: 1, 2: 2})


# ---- needs JAX from here on ----------------------------------------------------------------------------------
def _jax():
    try:
        import jax
        return jax
    except ImportError as e:   # this image: no jax / jaxlib wheels, no network
        raise RuntimeError("the jax.ffi binding needs jax + jaxlib (jax.ffi.include_dir() supplies the XLA FFI headers)") from e


def build_ffi(force=False) -> str:
    """g++ csrc/lob_ffi.cc against ``jax.ffi.include_dir()`` and csrc/liblobstep.so -> csrc/liblob_ffi.so."""
    jax = _jax()
    src = os.path.join(_CSRC, "lob_ffi.cc")
    if force or not os.path.exists(FFI_SO) or os.path.getmtime(FFI_SO) < os.path.getmtime(src):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", f"-I{jax.ffi.include_dir()}",
               f"-I{os.path.join(os.path.dirname(_HERE), 'include')}", f"-I{cuda_inc}", src, f"-L{_CSRC}", "-llobstep",
               f"-Wl,-rpath,{_CSRC}", "-o", FFI_SO]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building lob_ffi.cc failed: {' '.join(cmd)}\n{r.stderr}")
    return FFI_SO


_registered = False


def register():
    """``jax.ffi.register_ffi_target`` for "lob_step" (also serves reset) and "lob_replay" on the CUDA platform."""
    global _registered
    if not _registered:
        jax = _jax()
        lib = ctypes.CDLL(build_ffi())
        jax.ffi.register_ffi_target("lob_step", jax.ffi.pycapsule(lib.LobStep), platform="CUDA")
        jax.ffi.register_ffi_target("lob_replay", jax.ffi.pycapsule(lib.LobReplay), platform="CUDA")
        _registered = True


def ffi_step(cfg: abi.LobStepConfig, state_leaves: dict, inputs: dict, params: dict, reset_only=False):
    """One un-batched ``env.step`` (or ``env.reset``) as a custom call; ``jax.vmap`` adds the batch
    (``vmap_method="expand_dims"``: batched operands arrive with their leading B, the params with a leading 1 -- plain
    row-major buffers, the layout LobStepBuffers documents).  ``state_leaves`` = ``states.flatten(cfg, state)``,
    ``inputs`` = {"actions<t>", "perm", "reset_window", "reset_is_sell"[, "cancel_u"]}, ``params`` =
    ``states.flatten_params(params)``.  Returns {result name: array} (state leaves first)."""
    jax = _jax()
    register()
    tab = step_table(cfg, reset_only)
    specs = states.leaf_specs(cfg, 1)
    operands = [{**state_leaves, **inputs, **params}[n] for n in tab.names_in]
    n_state = len(tab.aliases)
    out_specs = [jax.ShapeDtypeStruct(operands[i].shape, operands[i].dtype) for i in range(n_state)] + \
                [jax.ShapeDtypeStruct((4,) if n == "work_redo_count" else specs[n][0][1:], specs[n][1])   # (per env: vmap adds B)
                 for n in tab.names_out[n_state:]]
    batch = int(np.prod(operands[0].shape[:-2])) if operands[0].ndim > 2 else 1     # asks [..., No, 6]
    outs = jax.ffi.ffi_call("lob_step", out_specs, vmap_method="expand_dims", input_output_aliases=tab.aliases)(
        *operands, cfg=np.frombuffer(bytes(cfg), np.uint8), batch=np.int64(batch), reset_only=bool(reset_only),
        arg_off=tab.arg_off, ret_off=tab.ret_off)
    return dict(zip(tab.names_out, outs))


def draw_prng_jax(env, key, cfg: abi.LobStepConfig):
    """The PRNG products of one step, drawn with the reference's own key discipline so that they stay bit-identical to
    what marl_env.py / base_env.py / exec_env.py draw inside the step: returns {"perm", "reset_window", "reset_is_sell"}."""
    jax = _jax()
    jnp = jax.numpy
    key, key_reset = jax.random.split(key)                                                    # marl_env.py:787
    _, shuffle_key = jax.random.split(key)                                                    # marl_env.py:294
    n_act = int(sum(cfg.agent[t].n_agents * cfg.agent[t].num_action_messages_by_agent for t in range(cfg.n_agent_types)))
    perm = jax.random.permutation(shuffle_key, n_act)        # permutation(key, x, axis=0) == x[permutation(key, n)], :295
    keys = jax.random.split(key_reset, cfg.n_agent_types + 1)                                 # marl_env.py:135
    reset_window = jax.random.randint(keys[0], (), 0, cfg.n_windows)                          # base_env.py:224
    reset_is_sell = jnp.stack([jax.random.randint(k, (), 0, 2) for k in keys[1:]])            # exec_env.py:221
    return {"perm": perm.astype(jnp.int32), "reset_window": reset_window.astype(jnp.int32),
            "reset_is_sell": reset_is_sell.astype(jnp.int32)}
