"""Deterministic synthetic LOBSTER-format day + loader (host side).

The reference's data path is ``LoadLOBSTER_resample`` (gymnax_exchange/jaxlobster/lobster_loader.py:516-1132):
LOBSTER message/orderbook CSV pairs -> one concatenated ``msgs[M,8]`` int32 tensor, window ``starts/ends``, one
L2 row per window.  No real LOBSTER data is available offline, so this module provides

* ``generate_day``   -- a seeded generator that keeps its own price-time-priority book and emits a LOBSTER
  *message* table (time, type, order_id, size, price, direction) plus the *orderbook* table (ask_p1, ask_s1,
  bid_p1, bid_s1, ... x levels, state AFTER each message), i.e. exactly what the two CSV files hold;
* ``write_lobster_csv`` -- writes the pair in LOBSTER naming so the *reference's own* loader can ingest it;
* ``preprocess_day`` / ``load_days`` -- the loader contract restated in numpy (ldr:891-945, 971-1071, 1073-1132,
  664-679): same output arrays as ``LoadLOBSTER_resample.run_loading``.

The result (`LoadedDay`) is what gets preloaded into HBM once (see env.py).
"""
import itertools
import os
from dataclasses import dataclass

import numpy as np
from sortedcontainers import SortedDict


@dataclass
class RawDay:
    """One LOBSTER day as the two CSV tables."""
    messages: np.ndarray    # float64 [R,6]  time, type, order_id, size, price, direction
    orderbook: np.ndarray   # int64   [R,4*levels]
    levels: int


@dataclass
class LoadedDay:
    """Output of the loader (== run_loading's tuple, ldr:695)."""
    msgs: np.ndarray        # int32 [M,8]  type, direction, qty, price, trader_id, order_id, time_s, time_ns
    starts: np.ndarray      # int64 [W]
    ends: np.ndarray        # int64 [W]
    books: np.ndarray       # int64 [W,4*levels]  L2 row BEFORE msgs[starts[w]]
    max_msgs: np.ndarray    # int64 [W]
    msgs_device: object = None   # the same msgs as a CUDA tensor when the day was preprocessed on the device (stays in HBM)


class _SynthBook:
    """Generator-side truth book: price -> FIFO list of [oid, qty] (+ running level totals)."""

    def __init__(self):
        self.side = {1: SortedDict(), -1: SortedDict()}  # 1 = bids, -1 = asks
        self.tot = {1: {}, -1: {}}
        self.orders = {}  # oid -> (side, price)

    def best(self, s):
        d = self.side[s]
        if not d:
            return None
        return d.peekitem(-1)[0] if s == 1 else d.peekitem(0)[0]

    def add(self, s, price, oid, qty, known=True):
        self.side[s].setdefault(price, []).append([oid, qty])
        self.tot[s][price] = self.tot[s].get(price, 0) + qty
        if known:
            self.orders[oid] = (s, price)

    def reduce(self, s, price, oid, qty):
        q = self.side[s][price]
        for i, o in enumerate(q):
            if o[0] == oid:
                o[1] -= qty
                self.tot[s][price] -= qty
                if o[1] <= 0:
                    q.pop(i)
                    self.orders.pop(oid, None)
                break
        if not q:
            del self.side[s][price]
            del self.tot[s][price]

    def l2_row(self, levels):
        row = [0] * (4 * levels)
        ta, tb = self.tot[-1], self.tot[1]
        k = 0
        for p in itertools.islice(self.side[-1], levels):
            row[4 * k] = p
            row[4 * k + 1] = ta[p]
            k += 1
        while k < levels:  # LOBSTER's empty-level filler would be +-9999999999; keep int32-safe values instead
            row[4 * k] = 2_000_000_000
            k += 1
        k = 0
        for p in itertools.islice(reversed(self.side[1]), levels):
            row[4 * k + 2] = p
            row[4 * k + 3] = tb[p]
            k += 1
        return row


def generate_day(seed=20220103, n_events=400_000, levels=10, mid=1_500_000, tick=100,
                 day_start=34200, day_end=57600, stress=False) -> RawDay:
    """Seeded synthetic day.  Event mix ~ 48 % limit adds (price = touch -+ Geom ticks, a few inside the spread),
    38 % cancels/deletes of live orders (2 % of them against initial-book liquidity, exercising
    get_init_id_match job:121), 14 % market orders emitted as type-4 executions of the resting orders they hit
    (multi-order sweeps share one timestamp, exercising merge_market_orders ldr:1073).  ``stress`` biases the flow
    towards adds so that a 100-row book side fills up (exercises the eviction / last-row quirks)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    book = _SynthBook()
    init_levels = levels + 6
    next_hidden = -1  # ids of initial-book liquidity never appear in an add message
    init_qty = {1: {}, -1: {}}
    for k in range(init_levels):
        for s in (1, -1):
            price = mid - s * (tick * (k + 1))
            q = int(rng.integers(1, 501))
            book.add(s, price, next_hidden, q, known=False)
            init_qty[s][price] = q
            next_hidden -= 1
    next_oid = 10_000_000
    # event times: sorted uniforms in the open interval, 9 decimals (what a LOBSTER file stores)
    t = np.sort(rng.uniform(day_start + 1e-3, day_end - 1e-3, size=n_events))
    t = np.round(t, 9)
    msgs, books = [], []
    live = []  # oids of known live orders (swap-remove list)
    live_pos = {}

    def live_add(oid):
        live_pos[oid] = len(live)
        live.append(oid)

    def live_del(oid):
        i = live_pos.pop(oid, None)
        if i is None:
            return
        last = live.pop()
        if last != oid:
            live[i] = last
            live_pos[last] = i

    def emit(ti, typ, oid, size, price, direction):
        msgs.append((ti, typ, oid, size, price, direction))
        books.append(book.l2_row(levels))

    p_add, p_cancel = (0.70, 0.18) if stress else (0.48, 0.38)
    target_live = 2000 if stress else 160  # keeps the env-side book (100 rows/side) mostly below capacity
    for i in range(n_events):
        ti = float(t[i])
        u = rng.random()
        if len(live) > target_live and u < p_add and rng.random() < 0.5:
            u = p_add  # mean reversion of the live-order count: turn this add into a cancel
        thin = min(len(book.side[1]), len(book.side[-1])) < levels + 2
        if u < p_add or thin or not live:
            s = 1 if rng.random() < 0.5 else -1
            if thin and len(book.side[1]) != len(book.side[-1]):
                s = 1 if len(book.side[1]) < len(book.side[-1]) else -1
            bb, ba = book.best(1), book.best(-1)
            own, opp = (bb, ba) if s == 1 else (ba, bb)
            if thin:
                # rebuild the ladder: first unoccupied tick behind the own touch
                k = 1
                while (own - s * k * tick) in book.side[s]:
                    k += 1
                price = own - s * k * tick
            else:
                # distance from the OPPOSITE touch keeps the spread tight: d = 0 quotes one tick from crossing
                d = int(rng.geometric(0.3)) - 1
                price = opp - s * (d + 1) * tick
            if price <= 0:
                price = tick
            q = int(rng.integers(1, 301))
            oid = next_oid
            next_oid += 1
            book.add(s, price, oid, q)
            live_add(oid)
            emit(ti, 1, oid, q, price, s)
        elif u < p_add + p_cancel:
            if rng.random() < 0.02:
                # cancel initial-book liquidity under an id the env has never seen
                s = 1 if rng.random() < 0.5 else -1
                cands = [p for p, q in init_qty[s].items() if q > 0 and p in book.side[s]]
                if cands:
                    price = cands[int(rng.integers(0, len(cands)))]
                    hid = [o for o in book.side[s][price] if o[0] < 0]
                    if hid:
                        o = hid[0]
                        q = int(rng.integers(1, o[1] + 1))
                        typ = 3 if q == o[1] else 2
                        fake_oid = next_oid
                        next_oid += 1
                        init_qty[s][price] -= q
                        book.reduce(s, price, o[0], q)
                        emit(ti, typ, fake_oid, q, price, s)
                        continue
            oid = live[int(rng.integers(0, len(live)))]
            s, price = book.orders[oid]
            q_live = next(o[1] for o in book.side[s][price] if o[0] == oid)
            if q_live > 1 and rng.random() < 0.2:
                q = int(rng.integers(1, q_live))
                typ = 2
            else:
                q = q_live
                typ = 3
                live_del(oid)
            book.reduce(s, price, oid, q)
            emit(ti, typ, oid, q, price, s)
        else:
            # market order against side s (the resting side), FIFO at the touch
            s = 1 if rng.random() < 0.5 else -1
            remaining = int(rng.integers(1, 301))
            n_hit = 0
            while remaining > 0 and n_hit < 4:
                price = book.best(s)
                if price is None or len(book.side[s]) <= levels:
                    break
                o = book.side[s][price][0]
                q = min(remaining, o[1])
                oid = o[0]
                shown = oid if oid > 0 else next_oid  # hidden initial liquidity executes under a fresh id
                if oid < 0:
                    next_oid += 1
                    init_qty[s][price] -= q
                elif q == o[1]:
                    live_del(oid)
                book.reduce(s, price, oid, q)
                emit(ti, 4, shown, q, price, s)
                remaining -= q
                n_hit += 1
    m = np.array(msgs, dtype=np.float64).reshape(-1, 6)
    ob = np.array(books, dtype=np.int64).reshape(-1, 4 * levels)
    return RawDay(messages=m, orderbook=ob, levels=levels)


def write_lobster_csv(day: RawDay, directory, stock="GOOG", date="2022-01-03", day_start=34200, day_end=57600):
    """``<dir>/<stock>_<date>_<start ms>_<end ms>_{message,orderbook}_<levels>.csv`` without header (ldr:587-617)."""
    os.makedirs(directory, exist_ok=True)
    base = f"{stock}_{date}_{day_start * 1000}_{day_end * 1000}"
    mpath = os.path.join(directory, f"{base}_message_{day.levels}.csv")
    opath = os.path.join(directory, f"{base}_orderbook_{day.levels}.csv")
    m = day.messages
    with open(mpath, "w") as f:
        for r in m:
            f.write(f"{r[0]:.9f},{int(r[1])},{int(r[2])},{int(r[3])},{int(r[4])},{int(r[5])}\n")
    np.savetxt(opath, day.orderbook, fmt="%d", delimiter=",")
    return mpath, opath


def merge_market_orders(typ, qty, price, direction, time_s, time_ns):
    """ldr:1073-1132: type-4 rows sharing (time_s, time_ns, direction) collapse into the LAST row of the group:
    qty = sum, price = max if direction == -1 else min.  Returns (keep_mask, qty, price)."""
    qty = qty.copy()
    price = price.copy()
    keep = np.ones(typ.shape[0], bool)
    ex = np.nonzero(typ == 4)[0]
    if ex.size == 0:
        return keep, qty, price
    key = np.stack([time_s[ex], time_ns[ex], direction[ex]], axis=1)
    _, inv, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    ng = counts.shape[0]
    last = np.full(ng, -1, np.int64)
    np.maximum.at(last, inv, ex)
    qsum = np.zeros(ng, np.int64)
    np.add.at(qsum, inv, qty[ex])
    pmax = np.full(ng, np.iinfo(np.int64).min)
    pmin = np.full(ng, np.iinfo(np.int64).max)
    np.maximum.at(pmax, inv, price[ex])
    np.minimum.at(pmin, inv, price[ex])
    multi = counts > 1
    gdir = np.zeros(ng, np.int64)
    gdir[inv] = direction[ex]
    sel = last[multi]
    qty[sel] = qsum[multi]
    price[sel] = np.where(gdir[multi] == -1, pmax[multi], pmin[multi])
    drop = ex[(ex != last[inv]) & multi[inv]]
    keep[drop] = False
    return keep, qty, price


def preprocess_day(day: RawDay, day_start=34200, day_end=57600, return_time=False):
    """ldr:891-945 ``_pre_process_msg_ob``.  Returns (msgs int64 [M,8] in the OUTPUT column order of ldr:1068-1070,
    orderbook int64 [M,4*levels]) with ``book[i]`` = state before ``msgs[i]``; with ``return_time`` also the float64
    ``time`` column (the one the fixed_time window filter of ldr:1040-1049 compares against)."""
    m = day.messages
    t = m[:, 0]
    time_s = t.astype(np.int64)
    time_ns = ((t - time_s) * 1_000_000_000).astype(np.int64)          # ldr:901-903 (float64, truncating)
    rows = np.arange(m.shape[0])
    typ = m[:, 1].astype(np.int64)
    mask = (time_s >= day_start) & (time_s <= day_end) & np.isin(typ, (1, 2, 3, 4))   # ldr:907-919
    rows = rows[mask]
    typ, oid = typ[mask], m[mask, 2].astype(np.int64)
    qty, price, direction = m[mask, 3].astype(np.int64), m[mask, 4].astype(np.int64), m[mask, 5].astype(np.int64)
    time_s, time_ns = time_s[mask], time_ns[mask]
    keep, qty, price = merge_market_orders(typ, qty, price, direction, time_s, time_ns)
    rows, typ, oid, qty, price, direction, time_s, time_ns = (
        a[keep] for a in (rows, typ, oid, qty, price, direction, time_s, time_ns))
    typ = np.where(typ == 3, 2, typ)                                    # ldr:932
    book = day.orderbook[rows]                                          # ldr:938
    out = np.stack([typ, direction, qty, price, oid, oid, time_s, time_ns], axis=1)   # trader_id := order_id ldr:935
    if return_time:
        return out[1:], book[:-1], t[rows][1:]
    return out[1:], book[:-1]                                           # ldr:941-942


def preprocess_day_cuda(day: RawDay, day_start=34200, day_end=57600, device="cuda"):
    """``preprocess_day`` on the device (csrc/lob_loader.cu through lob_loader_*_launch): the parsed message table goes to
    HBM once and the day's message tensor is produced there.  Returns (msgs int32 CUDA tensor [M,8], rows int64 CUDA tensor
    [M]: original row of the message BEFORE which ``orderbook[rows[j]]`` is the book state, time float64 CUDA tensor [M]).
    The table must be time-sorted (LOBSTER files are); an unsorted one is refused."""
    import ctypes as C

    import torch

    from . import _lib, abi
    L = _lib.lib()
    dev = torch.device(device)
    raw = torch.from_numpy(np.ascontiguousarray(day.messages[:, :6], np.float64)).to(dev)
    n = raw.shape[0]
    keep = torch.empty(n, dtype=torch.int64, device=dev)
    mq, mp = torch.empty_like(keep), torch.empty_like(keep)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    p64 = lambda t: C.cast(t.data_ptr(), abi.p_i64)
    pf = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    p32 = lambda t: C.cast(t.data_ptr(), abi.p_i32)
    with _lib.on_device(dev):
        st = _lib.current_stream_ptr(dev)
        _lib.check(L.lob_loader_flags_launch(pf(raw), n, int(day_start), int(day_end), p64(keep), p64(mq), p64(mp), p32(flags), st),
                   "lob_loader_flags_launch")
        pos = torch.cumsum(keep, 0) - keep                     # exclusive prefix sum (plumbing)
        M = int((pos[-1] + keep[-1]).item()) if n else 0
        if M < 2:
            raise ValueError("the day holds fewer than two usable messages")
        msgs = torch.empty((M - 1, 8), dtype=torch.int32, device=dev)
        tm = torch.empty(M - 1, dtype=torch.float64, device=dev)
        rows = torch.empty(M, dtype=torch.int64, device=dev)
        _lib.check(L.lob_loader_scatter_launch(pf(raw), n, int(day_start), int(day_end), p64(keep), p64(pos), p64(mq), p64(mp),
                                               p32(msgs), pf(tm), p64(rows), p32(flags), st), "lob_loader_scatter_launch")
        f = int(flags[0].item())
    if f & 1:
        raise ValueError("the message table is not time-sorted: merge_market_orders groups by time stamp (ldr:1073-1132)")
    if f & 2:
        raise ValueError("message field does not fit int32 (base:184 narrows silently; refuse instead)")
    return msgs, rows[:-1], tm


def window_indices(n_msgs, window_length, n_data_msg_per_step, window_resolution):
    """fixed_steps branch of ldr:971-1002 / :1019-1038."""
    if n_data_msg_per_step <= 0:
        raise ValueError("n_data_msg_per_step must be positive for 'fixed_steps'")
    d_end = n_msgs - window_length * n_data_msg_per_step
    end_index = d_end // n_data_msg_per_step * n_data_msg_per_step + 1
    starts = np.arange(0, end_index, n_data_msg_per_step * window_resolution, dtype=np.int64)
    if starts.shape[0] < 2:
        raise ValueError("Not enough range to get a slice")
    ends = starts + n_data_msg_per_step * window_length
    return starts, ends


def window_indices_fixed_time(time, window_length, window_resolution, day_start, day_end):
    """fixed_time branch of ldr:996 / :1040-1053: one window per ``range(day_start, day_end+1, resolution)[:-1]``
    start holding the messages with ``start <= time < start + window_length``; ``ends`` is the index of the LAST
    message in the window (inclusive, ldr:1049) and windows without data are dropped (ldr:1052-1053)."""
    grid = np.arange(day_start, day_end + 1, window_resolution, dtype=np.int64)
    if grid.shape[0] < 2:
        raise ValueError("Not enough range to get a slice")
    lo = np.searchsorted(time, grid[:-1].astype(np.float64), side="left")                  # first time >= start
    hi = np.searchsorted(time, (grid[:-1] + window_length).astype(np.float64), side="left")  # first time >= end
    ok = hi > lo
    return lo[ok].astype(np.int64), (hi[ok] - 1).astype(np.int64)


def load_days(days, window_length, n_data_msg_per_step, window_resolution, day_start=34200, day_end=57600,
              window_type="fixed_steps", device=None) -> LoadedDay:
    """ldr:626-695 ``run_loading`` over one or more days, concatenated with cumulative message offsets
    (ldr:664-679).  ``window_type`` is the reference's ``type_`` ("fixed_steps" | "fixed_time").  With ``device`` (a CUDA
    device) the per-day preprocessing runs there (``preprocess_day_cuda``) and the message tensor stays in HBM
    (``LoadedDay.msgs_device``); the window bookkeeping (a few hundred indices) is host arithmetic either way."""
    if window_type not in ("fixed_steps", "fixed_time"):
        raise NotImplementedError('Use either "fixed_time" or "fixed_steps"')      # ldr:998
    if device is not None:
        return _load_days_cuda(days, window_length, n_data_msg_per_step, window_resolution, day_start, day_end, window_type,
                               device)
    all_m, all_s, all_e, all_b, all_x = [], [], [], [], []
    offset = 0
    for d in days:
        if window_type == "fixed_time":
            m, ob, tm = preprocess_day(d, day_start, day_end, return_time=True)
            if np.any(np.diff(tm) < 0):
                raise ValueError("fixed_time windows need a time-sorted message file")
            s, e = window_indices_fixed_time(tm, window_length, window_resolution, day_start, day_end)
        else:
            m, ob = preprocess_day(d, day_start, day_end)
            s, e = window_indices(m.shape[0], window_length, n_data_msg_per_step, window_resolution)
        all_b.append(ob[s])
        all_x.append(e - s)
        all_s.append(s + offset)
        all_e.append(e + offset)
        all_m.append(m)
        offset += m.shape[0]
    msgs = np.concatenate(all_m, 0)
    if np.abs(msgs).max() > np.iinfo(np.int32).max:
        raise ValueError("message field does not fit int32 (base:184 narrows silently; refuse instead)")
    return LoadedDay(msgs=msgs.astype(np.int32), starts=np.concatenate(all_s), ends=np.concatenate(all_e),
                     books=np.concatenate(all_b, 0), max_msgs=np.concatenate(all_x))


def _load_days_cuda(days, window_length, n_data_msg_per_step, window_resolution, day_start, day_end, window_type, device):
    import torch
    all_m, all_s, all_e, all_b, all_x = [], [], [], [], []
    offset = 0
    for d in days:
        m, rows, tm = preprocess_day_cuda(d, day_start, day_end, device)
        if window_type == "fixed_time":
            tmh = tm.cpu().numpy()
            s, e = window_indices_fixed_time(tmh, window_length, window_resolution, day_start, day_end)
        else:
            s, e = window_indices(m.shape[0], window_length, n_data_msg_per_step, window_resolution)
        all_b.append(d.orderbook[rows[torch.from_numpy(s).to(rows.device)].cpu().numpy()])   # W rows of the book table
        all_x.append(e - s)
        all_s.append(s + offset)
        all_e.append(e + offset)
        all_m.append(m)
        offset += m.shape[0]
    msgs_dev = torch.cat(all_m, 0) if len(all_m) > 1 else all_m[0]
    return LoadedDay(msgs=msgs_dev.cpu().numpy(), starts=np.concatenate(all_s), ends=np.concatenate(all_e),
                     books=np.concatenate(all_b, 0), max_msgs=np.concatenate(all_x), msgs_device=msgs_dev)


def cache_suffix(world) -> str:
    """base:398-411 ``_get_filename_suffix``."""
    return "_".join(str(x) for x in (world.stock, world.timePeriod, world.book_depth, world.ep_type,
                                     world.episode_time, world.start_resolution, world.n_data_msg_per_step,
                                     world.day_start, world.day_end))


def load_or_generate(world, seed=20220103, n_events=400_000, stress=False, cache_dir=None, device=None) -> LoadedDay:
    """Loader entry point used by the env: npz cache with the reference's key names (ldr:686-693), else generate (and, with
    a CUDA ``device``, preprocess the day there)."""
    path = None
    if cache_dir is not None:
        os.makedirs(cache_dir, exist_ok=True)
        path = os.path.join(cache_dir, f"loaded_lobster_synth{seed}_{n_events}_{int(stress)}_{cache_suffix(world)}.npz")
        if os.path.exists(path):
            z = np.load(path)
            return LoadedDay(msgs=z["msgs"], starts=z["starts"], ends=z["ends"], books=z["obs"],
                             max_msgs=z["max_msgs_in_windows_arr"])
    day = generate_day(seed=seed, n_events=n_events, levels=world.book_depth, tick=world.tick_size,
                       day_start=world.day_start, day_end=world.day_end, stress=stress)
    ld = load_days([day], world.episode_time, world.n_data_msg_per_step, world.start_resolution,
                   world.day_start, world.day_end, window_type=world.ep_type, device=device)
    if path is not None:   # atomic: several ranks of one job may generate the same day at the same time
        tmp = f"{path}.{os.getpid()}.tmp.npz"
        np.savez_compressed(tmp, msgs=ld.msgs, starts=ld.starts, ends=ld.ends, obs=ld.books,
                            max_msgs_in_windows_arr=ld.max_msgs)
        os.replace(tmp, path)
    return ld


def init_messages_from_books(books, first_msgs, book_depth, init_id):
    """base:245-274 ``get_initial_orders``: every window's L2 row -> 2*depth limit-order messages
    ``[1, -+1, size, price, init_id, init_id-k, t_s, t_ns]`` (row 2k = ask level k, row 2k+1 = bid level k) stamped
    with the time of the window's first message.  Returns int32 [W, 2*depth, 8]."""
    W = books.shape[0]
    data = books.reshape(W, 2 * book_depth, 2)
    out = np.zeros((W, 2 * book_depth, 8), np.int64)
    out[:, :, 3] = data[:, :, 0]
    out[:, :, 2] = data[:, :, 1]
    out[:, :, 0] = 1
    out[:, 0::2, 1] = -1
    out[:, 1::2, 1] = 1
    out[:, :, 4] = init_id
    out[:, :, 5] = init_id - np.arange(2 * book_depth)[None, :]
    out[:, :, 6] = first_msgs[:, 6][:, None]
    out[:, :, 7] = first_msgs[:, 7][:, None]
    return out.astype(np.int32)
