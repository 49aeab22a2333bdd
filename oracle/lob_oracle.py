"""ctypes wrapper of oracle/lob_oracle.c (TEST INFRASTRUCTURE -- never imported by the product package).

``load()`` builds ``oracle/_build/liblob_oracle.so`` with the committed Makefile when it is missing or stale and
returns an ``Oracle`` whose methods take numpy arrays.  The interface structs are the ones of include/lobstep.h
(mirrored in jaxmarl_hft_b200/abi.py), so a test can hand the same buffer table to the oracle (host pointers) and to
the CUDA library (device pointers)."""
import ctypes as C
import os
import subprocess

import numpy as np

from jaxmarl_hft_b200 import abi, states

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblob_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "lob_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "lobstep.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        abi.check_sizes(lib)
        lib.lob_oracle_step.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64, C.c_int]
        lib.lob_oracle_reset.argtypes = [C.POINTER(abi.LobStepConfig), C.POINTER(abi.LobStepBuffers), C.c_int64]
        lib.lob_oracle_replay.argtypes = [C.POINTER(abi.LobBookConfig), C.POINTER(abi.LobReplayBuffers), C.c_int64, C.c_int]
        lib.lob_oracle_scan_save_bidask.argtypes = [C.POINTER(abi.LobBookConfig)] + [abi.p_i32] * 4 + [C.c_int32] + [abi.p_i32] * 2 + [abi.p_f32]
        lib.lob_oracle_l2.argtypes = [C.POINTER(abi.LobBookConfig), abi.p_i32, abi.p_i32, abi.p_i32, C.c_int32, C.c_int64]
        lib.lob_oracle_max_threads.restype = C.c_int

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed with code {rc}")

    def step(self, cfg, arrays, params, n_threads=1):
        """In-place MARLEnv.step on numpy arrays (names of states.leaf_specs)."""
        batch = arrays["asks"].shape[0]
        bufs = states.pack_buffers(cfg, arrays, params)
        self._check(self.lib.lob_oracle_step(C.byref(cfg), C.byref(bufs), batch, n_threads), "lob_oracle_step")

    def reset(self, cfg, arrays, params):
        batch = arrays["asks"].shape[0]
        bufs = states.pack_buffers(cfg, arrays, params)
        self._check(self.lib.lob_oracle_reset(C.byref(cfg), C.byref(bufs), batch), "lob_oracle_reset")

    def replay(self, book_cfg, asks, bids, trades, msgs, start, n_msgs, best_out=None, n_threads=1, cancel_u=None):
        r = states.pack_replay(asks, bids, trades, msgs, start, n_msgs, best_out, cancel_u)
        self._check(self.lib.lob_oracle_replay(C.byref(book_cfg), C.byref(r), asks.shape[0], n_threads), "lob_oracle_replay")

    def scan_save_bidask(self, book_cfg, asks, bids, trades, msgs, cancel_u=None):
        """job.scan_through_entire_array_save_bidask on ONE book; returns (bestasks[n,2], bestbids[n,2])."""
        msgs = np.ascontiguousarray(msgs, np.int32)
        n = msgs.shape[0]
        ba, bb = np.zeros((n, 2), np.int32), np.zeros((n, 2), np.int32)
        p = lambda a: a.ctypes.data_as(abi.p_i32)
        self._check(self.lib.lob_oracle_scan_save_bidask(C.byref(book_cfg), p(asks), p(bids), p(trades), p(msgs), n,
                                                         p(ba), p(bb),
                                                         None if cancel_u is None else cancel_u.ctypes.data_as(abi.p_f32)),
                    "lob_oracle_scan_save_bidask")
        return ba, bb

    def l2(self, book_cfg, asks, bids, n_levels):
        nb = asks.shape[0]
        out = np.zeros((nb, 4 * n_levels), np.int32)
        p = lambda a: a.ctypes.data_as(abi.p_i32)
        self._check(self.lib.lob_oracle_l2(C.byref(book_cfg), p(asks), p(bids), p(out), n_levels, nb), "lob_oracle_l2")
        return out

    def max_threads(self):
        """Host threads the OpenMP legs may use: the CPUs this process may run on (libgomp alone reports 1 when the
        environment pins OMP_NUM_THREADS)."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        return max(int(self.lib.lob_oracle_max_threads()), n)


def load():
    return Oracle(C.CDLL(build()))
