"""Shared test scaffolding: configs, a numpy env driven by the CPU oracle, seeded PRNG products."""
import os
import sys

import numpy as np

from jaxmarl_hft_b200 import abi, config as C, env as E, lobster, states

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.harness import OracleEnv, draw_actions, draw_prng, oracle_replay_fn  # noqa: F401  (re-exported)

CONFIG_DIR = C.CONFIG_DIR
load_mac = C.load_named_config
with_agents = C.with_agents


_DAY_CACHE = {}


def small_day(seed=5, n_events=30000, stress=False, levels=10):
    key = (seed, n_events, stress, levels)
    if key not in _DAY_CACHE:
        _DAY_CACHE[key] = lobster.generate_day(seed=seed, n_events=n_events, stress=stress, levels=levels)
    return _DAY_CACHE[key]


def load_for(mac, day):
    w = mac.world_config
    return lobster.load_days([day], w.episode_time, w.n_data_msg_per_step, w.start_resolution, w.day_start, w.day_end,
                             window_type=w.ep_type)


def copy_inputs(src, dst):
    for k in src:
        if k.startswith("actions") or k in ("perm", "reset_window", "reset_is_sell", "cancel_u"):
            dst[k][...] = src[k]


def assert_arrays_match(ref: dict, got: dict, cfg, rtol=1e-5, atol=1e-6, float_exact=False, skip=()):
    """ints / bytes bit-exact; floats where ``float_exact``: the same BIT PATTERN (so +0.0 != -0.0), any NaN == any NaN
    (market_share is 0/0 when nothing traded, mm:2408); else rel 1e-5 (north_star's tolerance)."""
    for k, r in ref.items():
        if k in skip or k.startswith("work_"):     # (workspace of the CUDA launch: scratch, not state)
            continue
        g = got[k]
        if r.dtype.kind in "iu":
            np.testing.assert_array_equal(g, r, err_msg=k)
        elif float_exact:
            g32, r32 = np.ascontiguousarray(g, np.float32), np.ascontiguousarray(r, np.float32)
            nan = np.isnan(g32) & np.isnan(r32)
            np.testing.assert_array_equal(np.where(nan, 0, g32.view(np.uint32)), np.where(nan, 0, r32.view(np.uint32)),
                                          err_msg=f"{k} (bit patterns)")
        else:
            np.testing.assert_allclose(g, r, rtol=rtol, atol=atol, equal_nan=True, err_msg=k)


def to_numpy(arrays):
    return {k: v.detach().cpu().numpy() for k, v in arrays.items()}


def empty_book(n_orders=100, n_trades=100):
    return (np.full((n_orders, 6), -1, np.int32), np.full((n_orders, 6), -1, np.int32),
            np.full((n_trades, 8), -1, np.int32))


# ---- CUDA side (only imported by -m gpu tests) ----------------------------------------------------------------
class CudaEnv:
    """The same buffer table on cuda:0, stepped by csrc/liblobstep.so through the C ABI (lob_*_launch)."""

    def __init__(self, mac, loaded, num_envs, params_np, device="cuda:0"):
        import torch
        from jaxmarl_hft_b200 import _lib
        self.torch, self._lib, self.L = torch, _lib, _lib.lib()
        self.mac, self.B, self.device = mac, num_envs, torch.device(device)
        self.cfg = C.to_step_config(mac, loaded.starts.shape[0], loaded.msgs.shape[0])
        self.params = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in params_np.items()}
        self.arrays = states.alloc_torch(self.cfg, num_envs, self.device)

    def load(self, arrays_np):
        for k, v in arrays_np.items():
            self.arrays[k].copy_(self.torch.from_numpy(v))

    def set_inputs(self, arrays_np):
        for k, v in arrays_np.items():
            if k.startswith("actions") or k in ("perm", "reset_window", "reset_is_sell", "cancel_u"):
                self.arrays[k].copy_(self.torch.from_numpy(v))

    def _bufs(self):
        return states.pack_buffers(self.cfg, self.arrays, self.params)

    def reset(self):
        import ctypes
        bufs = self._bufs()
        self._lib.check(self.L.lob_reset_launch(ctypes.byref(self.cfg), ctypes.byref(bufs), self.B,
                                                self._lib.current_stream_ptr()), "lob_reset_launch")

    def step(self):
        import ctypes
        bufs = self._bufs()
        self._lib.check(self.L.lob_step_launch(ctypes.byref(self.cfg), ctypes.byref(bufs), self.B,
                                               self._lib.current_stream_ptr()), "lob_step_launch")

    def numpy(self):
        self.torch.cuda.synchronize()
        return to_numpy(self.arrays)


def cuda_replay(book_cfg, asks, bids, trades, msgs, start, n_msgs, want_best=False, device="cuda:0", cancel_u=None):
    """numpy in -> numpy out through lob_replay_launch."""
    import torch
    from jaxmarl_hft_b200 import env as E2
    dev = torch.device(device)
    ta, tb, tt = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (asks, bids, trades))
    tm = torch.from_numpy(np.ascontiguousarray(msgs, np.int32)).to(dev)
    ts = torch.from_numpy(np.ascontiguousarray(start, np.int64)).to(dev)
    best = torch.zeros((asks.shape[0], 4), dtype=torch.int32, device=dev) if want_best else None
    tu = None if cancel_u is None else torch.from_numpy(np.ascontiguousarray(cancel_u, np.float32)).to(dev)
    E2.replay_books(book_cfg, ta, tb, tt, tm, ts, n_msgs, best, cancel_u=tu)
    torch.cuda.synchronize()
    out = (ta.cpu().numpy(), tb.cpu().numpy(), tt.cpu().numpy())
    return out + ((best.cpu().numpy(),) if want_best else ())


def random_messages(rng, n, book_cfg, price_lo=99_000, price_hi=101_000, tick=100, id_pool=400, weird=0.02):
    """A raw message stream that exercises every branch of job:556-637: limits on both sides (many crossing),
    cancels of known / unknown ids, type-4 executions, no-ops, and a few malformed (type, side) pairs."""
    m = np.zeros((n, 8), np.int32)
    t = rng.choice([1, 2, 3, 4, 0], size=n, p=[0.5, 0.2, 0.1, 0.18, 0.02])
    side = rng.choice([-1, 1], size=n)
    side[t == 0] = 0
    m[:, 0], m[:, 1] = t, side
    m[:, 2] = rng.integers(1, 300, size=n)
    m[:, 3] = rng.integers(price_lo // tick, price_hi // tick, size=n) * tick
    m[:, 4] = rng.integers(1, id_pool, size=n)          # small id pool: cancels hit live orders often
    m[:, 5] = m[:, 4]
    init_rows = rng.random(n) < 0.05                    # orders carrying the INITID (get_init_id_match job:121)
    m[init_rows, 4] = book_cfg.init_id
    m[init_rows, 5] = book_cfg.init_id - rng.integers(0, 2 * book_cfg.book_depth, size=init_rows.sum())
    ts = 34200 + np.cumsum(rng.integers(0, 2, size=n))
    m[:, 6] = ts
    m[:, 7] = rng.integers(0, 1_000_000_000, size=n)
    bad = rng.random(n) < weird                         # unexpected combos dispatch to ask_lim (quirk Q8)
    m[bad, 0] = rng.choice([0, 1, 5, 7], size=bad.sum())
    m[bad, 1] = rng.choice([0, -1, 1, 2], size=bad.sum())
    m[t == 0, 2:] = 0
    return m


def adversarial_messages(rng, n, book_cfg, tick=100):
    """Streams no market would produce: -1 in every field position, zero / negative quantities and prices, order id -1,
    time -1, INT32 extremes, ids colliding with the INITID range.  They drive the fast paths' bail-out conditions and
    the literal generic path (rows holding stray -1 fields, non-positive quantities, negative prices)."""
    m = random_messages(rng, n, book_cfg, price_lo=99_500, price_hi=100_500, tick=tick, id_pool=60, weird=0.05)
    def hit(p):
        return rng.random(n) < p
    for col, vals, p in ((2, [0, -1, -7, 1, 2**31 - 1], 0.06), (3, [-1, 0, -300, 2**31 - 1, 100_000], 0.05),
                         (4, [-1, 0, book_cfg.init_id, book_cfg.init_id - 3], 0.05), (5, [-1, 0], 0.03),
                         (6, [-1, 0, 2**31 - 1], 0.03), (7, [-1, 2**31 - 1, 0], 0.03)):
        sel = hit(p)
        m[sel, col] = rng.choice(vals, size=int(sel.sum()))
    return m
