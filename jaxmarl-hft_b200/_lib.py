"""Loader of the CUDA C-ABI library ``csrc/liblobstep.so`` (built in-tree by build.py).

There is NO CPU fallback: if the library is missing or a launch fails, the call raises."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LOB_SO") or os.path.join(_HERE, "csrc", "liblobstep.so")   # LOB_SO: a variant build (development)
_lib = None


class LobError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LobError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback for the LOB step.")
        L = C.CDLL(SO_PATH)
        abi.check_sizes(L)
        L.lob_abi_version.restype = C.c_int
        if L.lob_abi_version() != abi.LOB_ABI_VERSION:
            raise LobError("liblobstep.so ABI version mismatch; rebuild")
        L.lob_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.lob_step_launch.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64, vp]
        L.lob_reset_launch.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64, vp]
        L.lob_rollout_launch.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers),
                                         C.POINTER(abi.LobRolloutBuffers), C.c_int64, vp]
        L.lob_draw_launch.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64, C.c_int32,
                                      C.c_uint64, C.c_uint64, vp]
        L.lob_draw_launch_dev.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64, C.c_int32,
                                          C.c_uint64, vp, vp]
        L.lob_replay_launch.argtypes = [C.POINTER(abi.LobBookConfig), C.POINTER(abi.LobReplayBuffers), C.c_int64, vp]
        L.lob_replay_launch_grouped.argtypes = L.lob_replay_launch.argtypes
        L.lob_l2_launch.argtypes = [C.POINTER(abi.LobBookConfig), abi.p_i32, abi.p_i32, abi.p_i32, C.c_int32,
                                    C.c_int64, vp]
        L.lob_host_replay_create.argtypes = [C.POINTER(abi.LobBookConfig), C.c_int64, C.c_int64, C.c_int]
        L.lob_host_replay_create.restype = vp
        L.lob_host_replay_set_messages.argtypes = [vp, abi.p_i32, C.c_int64]
        L.lob_host_replay_run.argtypes = [vp, abi.p_i32, abi.p_i32, abi.p_i32, abi.p_i64, C.c_int32, C.c_int64,
                                          abi.p_i64, abi.p_i64]
        L.lob_host_replay_destroy.argtypes = [vp]
        L.lob_host_replay_destroy.restype = None
        p_f64, p64 = C.POINTER(C.c_double), abi.p_i64
        L.lob_loader_flags_launch.argtypes = [p_f64, C.c_int64, C.c_int32, C.c_int32, p64, p64, p64, abi.p_i32, vp]
        L.lob_loader_scatter_launch.argtypes = [p_f64, C.c_int64, C.c_int32, C.c_int32, p64, p64, p64, p64, abi.p_i32, p_f64,
                                                p64, abi.p_i32, vp]
        L.lob_launch_count.restype = C.c_int64
        L.lob_launch_count_reset.restype = None
        for fn in ("lob_num_msgs_per_step", "lob_num_action_msgs", "lob_num_cancel_msgs"):
            getattr(L, fn).argtypes = [C.POINTER(abi.LobStepConfig)]
            getattr(L, fn).restype = C.c_int32
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().lob_last_error()
        raise LobError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def current_stream_ptr(device=None):
    """The current torch stream of ``device`` (default: the current device) as the ``void*`` the C ABI takes."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(device):
    """Context manager: make ``device`` the current CUDA device for the launches inside (the C entry points launch on the
    current device and refuse buffers that live on another one)."""
    import torch
    return torch.cuda.device(torch.device(device))
