"""jaxmarl_hft_b200.jorderbook.OrderBook (the reference's object API, jaxob/jorderbook.py:25-268) against golden vectors
produced by the reference's own OrderBook class (tests/golden/make_golden.py ``orderbook``): its __main__ scenario
(jorderbook.py:288-318) and every query method on a book driven by a random stream."""
import os

import numpy as np
import pytest

from jaxmarl_hft_b200 import config as C
from jaxmarl_hft_b200.jorderbook import LobState, OrderBook

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orderbook_api.npz")


def _eq(state: LobState, z, prefix, b=0):
    for k in ("asks", "bids", "trades"):
        np.testing.assert_array_equal(getattr(state, k)[b].cpu().numpy(), z[prefix + k], err_msg=prefix + k)


@pytest.mark.parametrize("n_books", [1, 5])
def test_orderbook_object_matches_reference(n_books):
    z = np.load(GOLDEN)
    ob = OrderBook(C.JAXLOB_Configuration(nOrders=int(z["no"]), nTrades=int(z["nt"])), n_books=n_books, device="cuda:0")
    state = ob.reset(z["l2init"])
    for b in range(n_books):
        _eq(state, z, "reset/", b)
    quote = {"type": "limit", "side": "bid", "quantity": 99, "price": 346000, "trade_id": 8888, "order_id": 8888,
             "timestamp": "3400.005000000"}
    _eq(ob.process_order(state, quote), z, "dict/")
    _eq(ob.process_order(state, dict(quote, type="market", price=355000, quantity=500)), z, "market/")
    _eq(ob.process_order(state, dict(quote, type="cancel", price=344000, quantity=150, order_id=-5)), z, "cancel/")
    _eq(state, z, "reset/")                                                   # functional: the input state is untouched
    msgs2 = np.array([[1, 1, 99, 346000, 8888, 8888, 3400, 5000000], [1, -1, 2, 346000, 8777, 8777, 3401, 5060000]], np.int32)
    _eq(ob.process_order_array(state, msgs2[0]), z, "one/")
    st2, l2s = ob.process_orders_array_l2(state, msgs2, 10)
    _eq(st2, z, "two/", n_books - 1)
    np.testing.assert_array_equal(l2s[0].cpu().numpy(), z["two/l2"])
    g = lambda t: t[0].cpu().numpy()
    assert int(g(ob.get_volume_at_price(state, 1, 344000, True))) == int(z["q/vol_init"])
    assert int(g(ob.get_volume_at_price(st2, 1, 346000, False))) == int(z["q/vol"])
    np.testing.assert_array_equal(g(ob.get_next_executable_order(state, 1)), z["q/next_bid"])
    np.testing.assert_array_equal(g(ob.get_next_executable_order(state, 0)), z["q/next_ask"])
    assert int(g(ob.get_best_price(state, 1))) == int(z["q/best_bid"])
    assert int(g(ob.get_best_price(state, 0))) == int(z["q/best_ask"])
    ba, bb = ob.get_best_bid_and_ask_inclQuants(state)
    np.testing.assert_array_equal(g(ba), z["q/best_ask_q"]); np.testing.assert_array_equal(g(bb), z["q/best_bid_q"])
    with pytest.raises(ValueError):
        ob.get_volume_at_price(state, 2, 1)
    # the livelier book
    st3 = ob.process_orders_array(state, z["stream"])
    _eq(st3, z, "stream/", n_books - 1)
    np.testing.assert_array_equal(g(ob.get_L2_state(st3, 7)), z["stream/l2"])
    for side in (0, 1):
        np.testing.assert_array_equal(g(ob.get_side_ids(st3, side)), z[f"stream/ids{side}"])
        price0 = int(z[f"stream/vol_prices{side}"][0])
        for j, oid in enumerate(z[f"stream/probe_ids{side}"]):
            np.testing.assert_array_equal(g(ob.get_order(st3, side, int(oid))), z[f"stream/order{side}"][j])
            np.testing.assert_array_equal(g(ob.get_order(st3, side, int(oid), price0)), z[f"stream/order_p{side}"][j])
        arr = z["stream/bids" if side == 1 else "stream/asks"]
        last_price = int(arr[arr[:, 0] != -1][-1, 0])
        for j, (a, b) in enumerate(z[f"stream/probe_times{side}"]):
            np.testing.assert_array_equal(g(ob.get_order_at_time(st3, side, int(a), int(b))), z[f"stream/at_time{side}"][j])
            np.testing.assert_array_equal(g(ob.get_order_at_time(st3, side, int(a), int(b), last_price)),
                                          z[f"stream/at_time_p{side}"][j])
        np.testing.assert_array_equal(g(ob.get_next_executable_order(st3, side)), z[f"stream/next{side}"])
        for p, v in zip(z[f"stream/vol_prices{side}"], z[f"stream/vol{side}"]):
            assert int(g(ob.get_volume_at_price(st3, side, int(p)))) == int(v)


def test_orderbook_per_book_streams_and_donation(oracle):
    """[B,N,8] streams (one per book) against the oracle; donate=True updates the buffers in place."""
    import helpers as H
    cfg = C.JAXLOB_Configuration(nOrders=48, nTrades=16)
    bc = C.book_config(cfg)
    B, T = 7, 300
    msgs = H.random_messages(np.random.default_rng(5), B * T, bc, price_lo=99_500, price_hi=100_500).reshape(B, T, 8)
    ob = OrderBook(cfg, n_books=B, device="cuda:0", donate=True)
    st = ob.init()
    st2 = ob.process_orders_array(st, msgs)
    assert st2.asks.data_ptr() == st.asks.data_ptr()
    ra = np.full((B, 48, 6), -1, np.int32); rb = ra.copy(); rt = np.full((B, 16, 8), -1, np.int32)
    oracle.replay(bc, ra, rb, rt, msgs.reshape(-1, 8), np.arange(B, dtype=np.int64) * T, T)
    np.testing.assert_array_equal(st2.asks.cpu().numpy(), ra); np.testing.assert_array_equal(st2.bids.cpu().numpy(), rb)
    np.testing.assert_array_equal(st2.trades.cpu().numpy(), rt)
    np.testing.assert_array_equal(ob.get_L2_state(st2, 5).cpu().numpy(), oracle.l2(bc, ra, rb, 5))
