"""``OrderBook``: the reference's object wrapper around the book functions (gymnax_exchange/jaxob/jorderbook.py:25-268),
over the CUDA kernels of this package.

Same method names, argument meaning and return layouts as the reference; the difference is that the state is natively
BATCHED: ``LobState.asks / bids`` are int32 CUDA tensors ``[B, nOrders, 6]`` and ``trades`` ``[B, nTrades, 8]`` (the
reference vmaps the single-book object).  Like the reference the methods are functional -- the state passed in is left
untouched and a new ``LobState`` is returned -- unless ``donate=True`` is given to the constructor, in which case the
buffers are updated in place (what XLA does with donated arguments).  Every message method is ONE launch of ``lob_replay_launch`` (``job.scan_through_entire_array``);
``get_L2_state`` is ``lob_l2_launch``; the point look-ups (an order by id / time, volume at a price) are index arithmetic
on the device tensors.  There is no CPU path: without the CUDA library the constructor raises.
"""
import ctypes as C
from typing import NamedTuple, Optional

import numpy as np

from . import _lib, abi, states
from .config import JAXLOB_Configuration, book_config

NEGATIVE_RETURN_ID = -99     # jaxob_constants.py:10


class LobState(NamedTuple):
    """jorderbook.py:17-22 (the PRNG key is replaced by the draw counter of the random cancel modes)."""
    asks: "torch.Tensor"      # noqa: F821
    bids: "torch.Tensor"      # noqa: F821
    trades: "torch.Tensor"    # noqa: F821
    key: int = 0


def init_msgs_from_l2(cfg, book_l2, time=None):
    """job:1000-1028: one L2 row ``[ask_p, ask_q, bid_p, bid_q] x levels`` -> 2*levels limit orders (numpy int32 [2L,8]).
    Order ids count down from ``init_id``; the trader id is ``init_id`` (note: the env's own precompute, base:260-270,
    has the two columns the other way round)."""
    l2 = np.asarray(book_l2).reshape(-1)
    levels = l2.shape[0] // 4
    data = l2.reshape(levels * 2, 2)
    t = (34200, 0) if time is None else (int(time[0]), int(time[1]))
    out = np.zeros((levels * 2, 8), np.int64)
    out[:, 3] = data[:, 0]
    out[:, 2] = data[:, 1]
    out[:, 0] = 1
    out[0::2, 1] = -1
    out[1::2, 1] = 1
    out[:, 4] = cfg.init_id - np.arange(levels * 2)
    out[:, 5] = cfg.init_id
    out[:, 6], out[:, 7] = t
    return out.astype(np.int32)


class OrderBook:
    def __init__(self, cfg: Optional[JAXLOB_Configuration] = None, n_books: int = 1, device="cuda", seed: int = 0,
                 donate: bool = False):
        self.cfg = cfg if cfg is not None else JAXLOB_Configuration()
        self.book_cfg = book_config(self.cfg)
        self.n_books = int(n_books)
        self.device = device
        self.seed = int(seed)
        self.donate = bool(donate)
        _lib.lib()            # fail loudly when the CUDA library is missing

    # ---- state --------------------------------------------------------------------------------------------------
    def init(self) -> LobState:
        """jorderbook.py:33-39: empty sides and trade log (all -1)."""
        import torch
        B, c = self.n_books, self.cfg
        mk = lambda n, w: torch.full((B, n, w), -1, dtype=torch.int32, device=self.device)
        return LobState(mk(c.nOrders, 6), mk(c.nOrders, 6), mk(c.nTrades, 8), 0)

    def reset(self, l2_book=None, time=None) -> LobState:
        """jorderbook.py:41-53: empty book, optionally filled from one L2 row (the same row for every book, or one row
        per book ``[B, 4*levels]``)."""
        import torch
        state = self.init()
        if l2_book is not None:
            l2 = np.asarray(l2_book)
            rows = l2.reshape(1, -1) if l2.ndim == 1 else l2
            if rows.shape[0] not in (1, self.n_books):
                raise ValueError(f"l2_book: one row or one row per book ({self.n_books}) expected, got {rows.shape[0]}")
            time = (0, 0) if time is None else time                                       # jorderbook.py:49-50
            m = np.stack([init_msgs_from_l2(self.cfg, r, time) for r in rows])            # [R, 2L, 8]
            msgs = torch.from_numpy(np.ascontiguousarray(m.reshape(-1, 8))).to(self.device)
            n = m.shape[1]
            start = torch.arange(self.n_books, dtype=torch.int64, device=self.device) * (n if rows.shape[0] > 1 else 0)
            state = self._scan(state, msgs, start, n, inplace=True)
        return state

    # ---- message processing ---------------------------------------------------------------------------------------
    def _scan(self, state: LobState, msgs, start, n_msgs, best_out=None, inplace=False):
        import torch
        if not (inplace or self.donate):
            state = LobState(state.asks.clone(), state.bids.clone(), state.trades.clone(), state.key)
        cu = None
        if self.book_cfg.cancel_mode >= 2:        # job:142-164: the uniform draws of the random cancel fallbacks
            g = torch.Generator(device=self.device)
            g.manual_seed(self.seed * 1_000_003 + state.key)
            cu = (torch.randint(0, 2 ** 23, (self.n_books, n_msgs, 2), generator=g, device=self.device)
                  .to(torch.float32) / 2.0 ** 23)
        r = states.pack_replay(state.asks, state.bids, state.trades, msgs, start, n_msgs, best_out, cu)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().lob_replay_launch(C.byref(self.book_cfg), C.byref(r), self.n_books,
                                                    _lib.current_stream_ptr(self.device)), "lob_replay_launch")
        return state._replace(key=state.key + 1)

    def _as_msgs(self, msgs):
        """[N,8] (every book processes the same stream) or [B,N,8] -> (flat int32 CUDA tensor, start offsets, N)."""
        import torch
        t = msgs if isinstance(msgs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(msgs, np.int32))
        t = t.to(device=self.device, dtype=torch.int32).contiguous()
        if t.dim() == 2:
            return t, torch.zeros(self.n_books, dtype=torch.int64, device=self.device), t.shape[0]
        if t.dim() != 3 or t.shape[0] != self.n_books:
            raise ValueError("messages must be [N,8] or [n_books,N,8]")
        n = t.shape[1]
        return t.reshape(-1, 8), torch.arange(self.n_books, dtype=torch.int64, device=self.device) * n, n

    def process_order(self, state: LobState, quote: dict, from_data: bool = False, verbose: bool = False) -> LobState:
        """jorderbook.py:55-96: a dict quote -> one message.  'market' is a limit order on the opposite side."""
        inttype, intside = 5, -1
        if quote["side"] == "bid":
            intside = 1
        if quote["type"] == "limit":
            inttype = 1
        elif quote["type"] in ("cancel", "delete"):
            inttype = 2
        elif quote["type"] == "market":
            inttype = 1
            intside = -intside
        s, ns = quote["timestamp"].split(".")
        msg = np.array([[inttype, intside, quote["quantity"], quote["price"], quote["trade_id"], quote["order_id"],
                         int(s), int(ns)]], np.int32)
        return self.process_orders_array(state, msg)

    def process_order_array(self, state: LobState, quote, from_data: bool = False, verbose: bool = False) -> LobState:
        """jorderbook.py:98-110: one message as an array [8] (or one per book, [B,8])."""
        import torch
        q = quote if isinstance(quote, torch.Tensor) else torch.from_numpy(np.asarray(quote, np.int32))
        return self.process_orders_array(state, q.reshape(1, 8) if q.dim() == 1 else q.reshape(self.n_books, 1, 8))

    def process_orders_array(self, state: LobState, msgs) -> LobState:
        """jorderbook.py:112-120: ``job.scan_through_entire_array`` of [N,8] messages (or [B,N,8]: one stream per book)."""
        flat, start, n = self._as_msgs(msgs)
        return self._scan(state, flat, start, n)

    def process_orders_array_l2(self, state: LobState, msgs, n_levels: int):
        """jorderbook.py:123-136: the same, plus the L2 snapshot after EVERY message -> (state, int32 [B,N,4*n_levels]).
        One replay launch and one L2 launch per message (a debugging aid in the reference too)."""
        import torch
        flat, start, n = self._as_msgs(msgs)
        out = torch.empty((self.n_books, n, 4 * n_levels), dtype=torch.int32, device=self.device)
        if not self.donate:
            state = LobState(state.asks.clone(), state.bids.clone(), state.trades.clone(), state.key)
        for i in range(n):
            state = self._scan(state, flat, start + i, 1, inplace=True)
            out[:, i] = self.get_L2_state(state, n_levels)
        return state, out

    # ---- queries ------------------------------------------------------------------------------------------------
    @staticmethod
    def _side(state: LobState, side: int):
        if side not in (0, 1):
            raise ValueError("Side must be 0 or 1")
        return state.bids if side == 1 else state.asks

    def get_volume_at_price(self, state: LobState, side: int, price: int, init_only: bool = False):
        """jorderbook.py:138-157 / job:907-917, :1030-1047 -> int32 [B]."""
        import torch
        a = self._side(state, side)
        m = a[..., 0] == price
        if init_only:
            m = m & (a[..., 2] <= self.cfg.init_id) & (a[..., 2] >= self.cfg.init_id - self.cfg.book_depth * 2)
        return torch.where(m, a[..., 1], torch.zeros_like(a[..., 1])).sum(-1, dtype=torch.int32)

    def get_best_bid_and_ask_inclQuants(self, state: LobState):
        """jorderbook.py:186-191 / job:968-984 -> (best_ask [B,2], best_bid [B,2]) as [price, volume at it]: a zero-message
        replay launch with the ``best_out`` output."""
        import torch
        best = torch.empty((self.n_books, 4), dtype=torch.int32, device=self.device)
        r = states.pack_replay(state.asks, state.bids, state.trades, state.asks.new_zeros((1, 8)),
                               torch.zeros(self.n_books, dtype=torch.int64, device=self.device), 0, best)
        cfg1 = abi.LobBookConfig.from_buffer_copy(self.book_cfg)
        cfg1.cancel_mode = min(int(cfg1.cancel_mode), 1)      # nothing is processed: no draws needed
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().lob_replay_launch(C.byref(cfg1), C.byref(r), self.n_books,
                                                    _lib.current_stream_ptr(self.device)), "lob_replay_launch")
        return best[:, 0:2], best[:, 2:4]

    def get_best_ask(self, state: LobState):
        return self.get_best_bid_and_ask_inclQuants(state)[0][:, 0]

    def get_best_bid(self, state: LobState):
        return self.get_best_bid_and_ask_inclQuants(state)[1][:, 0]

    def get_best_price(self, state: LobState, side: int):
        """jorderbook.py:159-170."""
        return self.get_best_bid(state) if side == 1 else self.get_best_ask(state)

    def get_L2_state(self, state: LobState, n_levels: int):
        """jorderbook.py:193-199 / job:1232-1264 -> int32 [B, 4*n_levels]."""
        import torch
        out = torch.empty((self.n_books, 4 * n_levels), dtype=torch.int32, device=self.device)
        p = lambda t: C.cast(t.data_ptr(), abi.p_i32)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().lob_l2_launch(C.byref(self.book_cfg), p(state.asks), p(state.bids), p(out), n_levels,
                                                self.n_books, _lib.current_stream_ptr(self.device)), "lob_l2_launch")
        return out

    def get_side_ids(self, state: LobState, side: int):
        """jorderbook.py:201-213 / job:1201-1209: sorted unique order ids per book, padded with 1 -> int32 [B,nOrders]."""
        import torch
        ids = self._side(state, side)[..., 2].to(torch.int64)
        s, _ = torch.sort(ids, dim=-1)
        first = torch.ones_like(s, dtype=torch.bool)
        first[:, 1:] = s[:, 1:] != s[:, :-1]
        big = torch.full_like(s, 1 << 40)
        u, _ = torch.sort(torch.where(first, s, big), dim=-1)      # the distinct ids in order, then the padding
        return torch.where(u == big, torch.ones_like(u), u).to(torch.int32)

    def _first_row(self, a, mask, missing_is_last_row=False):
        """``side_array[jnp.where(mask, size=1, fill_value=-1)]``; not found -> a row of NEGATIVE_RETURN_ID, or (the
        id look-ups, see get_order) row -1 = the LAST row of the side."""
        import torch
        n = a.shape[1]
        idx = torch.where(mask, torch.arange(n, device=a.device)[None, :], torch.full((1, 1), n, device=a.device)).amin(-1)
        found = idx < n
        rows = a[torch.arange(a.shape[0], device=a.device), idx.clamp(max=n - 1)]
        if missing_is_last_row:
            return rows
        return torch.where(found[:, None], rows, torch.full_like(rows, NEGATIVE_RETURN_ID))

    def get_order(self, state: LobState, side: int, order_id: int, price: Optional[int] = None):
        """jorderbook.py:215-233 / job:1049-1124 -> int32 [B,6].  As in the reference an id that is not in the book
        yields the side's LAST row (job:1064-1071 compares the *tuple* returned by ``jnp.where`` with -1, which is never
        true, and then indexes with the fill value -1), not the documented row of NEGATIVE_RETURN_ID."""
        a = self._side(state, side)
        m = a[..., 2] == order_id
        if price is not None:
            m = m & (a[..., 0] == price)
        return self._first_row(a, m, missing_is_last_row=True)

    def get_order_at_time(self, state: LobState, side: int, time_s: int, time_ns: int, price: Optional[int] = None):
        """jorderbook.py:235-255 / job:1128-1199 -> int32 [B,6] (NEGATIVE_RETURN_ID when no order carries the time).  With
        ``price`` the match is tried at that price first and falls back to the time alone (job:1185-1193)."""
        import torch
        a = self._side(state, side)
        m = (a[..., 4] == time_s) & (a[..., 5] == time_ns)
        if price is not None:
            mp = m & (a[..., 0] == price)
            m = torch.where(mp.any(-1, keepdim=True), mp, m)
        return self._first_row(a, m)

    def get_next_executable_order(self, state: LobState, side: int):
        """jorderbook.py:257-268 / job:1212-1229, :242-268: the order with price-time priority (extreme price, then
        earliest time_s, then earliest time_ns, then first row) -> int32 [B,6]."""
        import torch
        a = self._side(state, side).to(torch.int64)
        mx = int(self.cfg.maxint)
        p = a[..., 0]
        if side == 1:
            ext = p.amax(-1, keepdim=True)
        else:
            ext = torch.where(p == -1, torch.full_like(p, mx), p).amin(-1, keepdim=True)
        times = torch.where(p == ext, a[..., 4], torch.full_like(p, mx))
        mts = times.amin(-1, keepdim=True)
        tns = torch.where(times == mts, a[..., 5], torch.full_like(p, mx))
        m = tns == tns.amin(-1, keepdim=True)
        n = a.shape[1]
        idx = torch.where(m, torch.arange(n, device=a.device)[None, :], torch.full((1, 1), n, device=a.device)).amin(-1)
        idx = torch.where(idx == n, torch.full_like(idx, n - 1), idx)   # where(size=1, fill=-1)[0] -> row -1 = last
        return self._side(state, side)[torch.arange(a.shape[0], device=a.device), idx]
