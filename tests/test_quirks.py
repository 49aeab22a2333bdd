"""Hand-worked known-answer cases for the quirks of the reference's array order book (SURVEY.md 8a "quirk ledger",
gymnax_exchange/jaxob/JaxOrderBookArrays.py = "job").  Each expected book below was derived by hand from the cited
reference lines on a 4-row book / 3-row trade log, and is checked against the CPU oracle (always) and the CUDA path
(-m gpu).  Message = [type, side, qty, price, order_id, trader_id, time_s, time_ns]; order row = [price, qty, order_id,
trader_id, time_s, time_ns]; trade row = [price, -aggressor_side*qty, passive_oid, aggressor_oid, time_s, time_ns,
passive_tid, aggressor_tid]."""
import numpy as np
import pytest

import helpers as H
from jaxmarl_hft_b200 import config as C

E6, E8 = [-1] * 6, [-1] * 8


def _cfg(**kw):
    return C.book_config(C.World_EnvironmentConfig(nOrders=4, nTrades=3, **kw))


def _run(engine, oracle, bc, msgs):
    msgs = np.asarray(msgs, np.int32).reshape(-1, 8)
    a = np.full((1, 4, 6), -1, np.int32); b = a.copy(); t = np.full((1, 3, 8), -1, np.int32)
    start = np.zeros(1, np.int64)
    if engine == "oracle":
        best = np.zeros((1, 4), np.int32)
        oracle.replay(bc, a, b, t, msgs, start, msgs.shape[0], best_out=best)
    else:
        a, b, t, best = H.cuda_replay(bc, a, b, t, msgs, start, msgs.shape[0], want_best=True)
    return a[0].tolist(), b[0].tolist(), t[0].tolist(), best[0].tolist()


ENGINES = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("engine", ENGINES)
def test_q1_slot_is_first_row_holding_any_minus_one(engine, oracle):
    """job:73: add_order writes into the first row that CONTAINS a -1 in any field -- a live order whose trader id is -1
    is overwritten by the next add."""
    msgs = [[1, 1, 10, 100, 1, -1, 5, 0],      # bid, trader id -1 -> row 0 = [100,10,1,-1,5,0]
            [1, 1, 7, 99, 2, 22, 6, 0]]        # next bid: row 0 still "contains -1" -> overwritten
    a, b, t, _ = _run(engine, oracle, _cfg(), msgs)
    assert b == [[99, 7, 2, 22, 6, 0], E6, E6, E6] and a == [E6] * 4 and t == [E8] * 3


@pytest.mark.parametrize("engine", ENGINES)
def test_q1_full_side_without_eviction_overwrites_last_row(engine, oracle):
    """job:73: no row holds a -1 -> where(..., fill_value=-1) -> index -1 -> the LAST row is overwritten
    (check_book_fill=False, so nothing is evicted first)."""
    msgs = [[1, 1, 1, 100 - k, 10 + k, 7, k, 0] for k in range(4)] + [[1, 1, 9, 50, 99, 8, 9, 0]]
    a, b, t, _ = _run(engine, oracle, _cfg(check_book_fill=False), msgs)
    assert b == [[100, 1, 10, 7, 0, 0], [99, 1, 11, 7, 1, 0], [98, 1, 12, 7, 2, 0], [50, 9, 99, 8, 9, 0]]


@pytest.mark.parametrize("engine", ENGINES)
def test_q2_unmatched_cancel_hits_the_last_row(engine, oracle):
    """job:110-117: a cancel matching no order id and no initial liquidity gives index -1, which JAX normalises to the
    last row: on a full side that order loses the quantity (and is blanked at <= 0)."""
    fill = [[1, -1, 5, 100 + k, 10 + k, 7, k, 0] for k in range(4)]                  # asks 100..103, qty 5
    a, _, _, _ = _run(engine, oracle, _cfg(), fill + [[2, -1, 3, 555, 999, 9, 9, 0]])
    assert a == [[100, 5, 10, 7, 0, 0], [101, 5, 11, 7, 1, 0], [102, 5, 12, 7, 2, 0], [103, 2, 13, 7, 3, 0]]
    a, _, _, _ = _run(engine, oracle, _cfg(), fill + [[2, -1, 5, 555, 999, 9, 9, 0]])
    assert a[3] == E6 and a[:3] == [[100, 5, 10, 7, 0, 0], [101, 5, 11, 7, 1, 0], [102, 5, 12, 7, 2, 0]]
    # on a side with blank rows the last row is blank: -1 - q stays <= 0 and is blanked again -> no visible effect
    a, _, _, _ = _run(engine, oracle, _cfg(), fill[:2] + [[2, -1, 3, 555, 999, 9, 9, 0]])
    assert a == [[100, 5, 10, 7, 0, 0], [101, 5, 11, 7, 1, 0], E6, E6]


@pytest.mark.parametrize("engine", ENGINES)
def test_cancel_falls_back_to_initial_liquidity_at_the_price(engine, oracle):
    """job:121-139: unknown order id -> the first row at the message's price whose order id lies in
    [init_id - 2*book_depth, init_id] and that holds at least the cancelled quantity."""
    msgs = [[1, 1, 50, 100, C.INITID, C.INITID - 1, 0, 0],       # initial-liquidity order (id = init_id)
            [1, 1, 8, 100, 77, 7, 1, 0],                          # an ordinary order at the same price
            [2, 1, 20, 100, 555, 9, 2, 0],                        # cancel of an id that is not in the book
            [2, 1, 40, 100, 556, 9, 3, 0]]                        # 40 > 30 left: no candidate -> last row (blank): no effect
    _, b, _, _ = _run(engine, oracle, _cfg(), msgs)
    assert b == [[100, 30, C.INITID, C.INITID - 1, 0, 0], [100, 8, 77, 7, 1, 0], E6, E6]


@pytest.mark.parametrize("engine", ENGINES)
def test_q3_trade_slot_is_first_row_with_time_minus_one(engine, oracle):
    """job:205: the trade is written to the first row whose column 4 (time_s on the trade layout) is -1 -- a trade
    stamped time_s = -1 is overwritten by the next one; with the log full the LAST row is overwritten."""
    asks = [[1, -1, 5, 100, 1, 31, 0, 0], [1, -1, 5, 100, 2, 32, 1, 0]]
    msgs = asks + [[1, 1, 5, 100, 50, 60, -1, 4],                 # buy 5 @100 stamped time_s = -1: trade lands in row 0
                   [1, 1, 5, 100, 51, 61, 7, 0]]                  # next trade: row 0 still has time_s == -1 -> overwritten
    a, b, t, _ = _run(engine, oracle, _cfg(), msgs)
    assert a == [E6] * 4 and b == [E6] * 4                        # both buys fully matched, zero remainders never rest
    assert t == [[100, -5, 2, 51, 7, 0, 32, 61], E8, E8]
    # four trades into a 3-row log: the 4th overwrites row 2
    asks = [[1, -1, 1, 100, 10 + k, 30 + k, k, 0] for k in range(4)]
    buys = [[1, 1, 1, 100, 50 + k, 60 + k, 10 + k, 0] for k in range(4)]
    _, _, t, _ = _run(engine, oracle, _cfg(), asks + buys)
    assert t == [[100, -1, 10, 50, 10, 0, 30, 60], [100, -1, 11, 51, 11, 0, 31, 61], [100, -1, 13, 53, 13, 0, 33, 63]]


@pytest.mark.parametrize("engine", ENGINES)
def test_q5_eviction_on_a_full_side_survives_the_ioc_discard(engine, oracle):
    """job:395-401, :415-418: before a limit / IOC order is handled, a FULL own side loses every row at its worst price --
    whatever the order's remaining quantity -- and a type-4 (IOC) order then discards its own remainder but not the
    eviction.  Type 4 flips the side (job:575): [4, -1, ...] is a buy."""
    bids = [[1, 1, 2, 100 - k, 10 + k, 7, k, 0] for k in range(4)]                    # bids 100, 99, 98, 97: side full
    _, b, t, _ = _run(engine, oracle, _cfg(), bids + [[4, -1, 3, 50, 99, 8, 9, 0]])  # IOC buy @50: nothing to match
    assert b == [[100, 2, 10, 7, 0, 0], [99, 2, 11, 7, 1, 0], [98, 2, 12, 7, 2, 0], E6] and t == [E8] * 3
    # a limit order is then added into the freed row
    _, b, _, _ = _run(engine, oracle, _cfg(), bids + [[1, 1, 3, 50, 99, 8, 9, 0]])
    assert b[3] == [50, 3, 99, 8, 9, 0]
    # type_4_interpretation = LIM keeps the IOC remainder
    _, b, _, _ = _run(engine, oracle, _cfg(type_4_interpretation=1), bids + [[4, -1, 3, 50, 99, 8, 9, 0]])
    assert b[3] == [50, 3, 99, 8, 9, 0]


@pytest.mark.parametrize("engine", ENGINES)
def test_q6_zero_quantity_add_is_written_then_blanked(engine, oracle):
    """job:76-83, :86-90: max(0, qty) is written and _removeZeroNegQuant blanks every row with qty <= 0."""
    a, b, t, _ = _run(engine, oracle, _cfg(), [[1, 1, 0, 100, 1, 7, 0, 0], [1, -1, -4, 101, 2, 7, 1, 0]])
    assert a == [E6] * 4 and b == [E6] * 4 and t == [E8] * 3


@pytest.mark.parametrize("engine", ENGINES)
def test_q7_empty_side_reports_a_negative_volume(engine, oracle):
    """job:968-984: best ask of an empty side is -1 (best bid = max price = -1) and the volume "at that price" is the
    sum of the blank rows' quantities: -nOrders."""
    _, _, _, best = _run(engine, oracle, _cfg(), [[0, 0, 0, 0, 0, 0, 0, 0]])
    assert best == [-1, -4, -1, -4]
    _, _, _, best = _run(engine, oracle, _cfg(), [[1, -1, 5, 100, 1, 7, 0, 0], [1, -1, 6, 100, 2, 7, 1, 0], [1, 1, 2, 90, 3, 7, 2, 0]])
    assert best == [100, 11, 90, 2]


@pytest.mark.parametrize("engine", ENGINES)
def test_q8_unexpected_type_side_pairs_are_ask_limits(engine, oracle):
    """job:588-596: the lax.switch index is a sum of products that is 0 for every (type, side) outside the table, and
    index 0 is ask_lim -- a "type 7, side +1" message rests on the ASK side; (0, 0) is the only doNothing."""
    a, b, _, _ = _run(engine, oracle, _cfg(), [[7, 1, 5, 100, 1, 9, 1, 0], [0, 0, 9, 9, 9, 9, 9, 9], [2, 0, 1, 100, 1, 9, 2, 0]])
    # the third message (cancel, side 0) is ask_lim too: a sell of 1 @100 that finds no bid and rests
    assert a == [[100, 5, 1, 9, 1, 0], [100, 1, 1, 9, 2, 0], E6, E6] and b == [E6] * 4


@pytest.mark.parametrize("engine", ENGINES)
def test_price_time_priority_and_partial_fill(engine, oracle):
    """job:242-268, :173-220: best price first, then earliest time_s, then earliest time_ns, then lowest row; the passive
    order keeps its remainder, the aggressor's remainder rests."""
    msgs = [[1, -1, 4, 101, 1, 31, 5, 0],       # worse price
            [1, -1, 4, 100, 2, 32, 5, 9],       # best price, later ns
            [1, -1, 4, 100, 3, 33, 5, 1],       # best price, earliest -> matched first
            [1, 1, 6, 100, 50, 60, 8, 0]]       # buy 6 @100: 4 from oid 3, 2 from oid 2; nothing rests
    a, b, t, _ = _run(engine, oracle, _cfg(), msgs)
    assert a == [[101, 4, 1, 31, 5, 0], [100, 2, 2, 32, 5, 9], E6, E6] and b == [E6] * 4
    assert t == [[100, -4, 3, 50, 8, 0, 33, 60], [100, -2, 2, 50, 8, 0, 32, 60], E8]
    # a buy through the book: 2 @100, 4 @101, remainder 3 rests at 102
    a, b, t, _ = _run(engine, oracle, _cfg(), msgs + [[1, 1, 9, 102, 51, 61, 9, 0]])
    assert a == [E6] * 4 and b == [[102, 3, 51, 61, 9, 0], E6, E6, E6]
    # 3-row log: the third trade (2 @100 from oid 2) took the free row 2, the fourth (4 @101 from oid 1) found no free
    # row and overwrote the LAST row (quirk Q3)
    assert t == [[100, -4, 3, 50, 8, 0, 33, 60], [100, -2, 2, 50, 8, 0, 32, 60], [101, -4, 1, 51, 9, 0, 31, 61]]
