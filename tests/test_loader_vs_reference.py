"""The loader restatement (lobster.preprocess_day / load_days) against the reference's OWN loader.

``gymnax_exchange/jaxlobster/lobster_loader.py`` is pandas/numpy only -- its single ``import jax`` is unused -- so it
can be imported in the build container with an empty stand-in ``jax`` module and run on a CSV pair written by our
generator.  /root/reference does not exist on the GPU box: the test skips there."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import REFERENCE, has_reference
from jaxmarl_hft_b200 import lobster

pytestmark = pytest.mark.skipif(not has_reference(), reason="reference sources not mounted")


def _import_reference_loader():
    if "jax" not in sys.modules:
        sys.modules["jax"] = types.ModuleType("jax")  # unused by the loader (ldr:38)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from gymnax_exchange.jaxlobster import lobster_loader
    return lobster_loader


@pytest.mark.parametrize("seed,n_events,stress", [(7, 6000, False), (11, 5000, True)])
def test_loader_matches_reference(tmp_path, seed, n_events, stress):
    ldr = _import_reference_loader()
    day = lobster.generate_day(seed=seed, n_events=n_events, stress=stress)
    data = tmp_path / "data"
    lobster.write_lobster_csv(day, str(data / "rawLOBSTER" / "GOOG" / "2022"))
    ref = ldr.LoadLOBSTER_resample(str(data), str(tmp_path / "at"), n_Levels=10, type_="fixed_steps",
                                   window_length=8, window_resolution=4, n_data_msg_per_step=50,
                                   stock="GOOG", time_period="2022")
    msgs, starts, ends, books, max_msgs = ref.run_loading("t")
    ours = lobster.load_days([day], window_length=8, n_data_msg_per_step=50, window_resolution=4)
    np.testing.assert_array_equal(np.asarray(msgs, np.int64), ours.msgs.astype(np.int64))
    np.testing.assert_array_equal(np.asarray(starts), ours.starts)
    np.testing.assert_array_equal(np.asarray(ends), ours.ends)
    np.testing.assert_array_equal(np.asarray(books), ours.books)
    np.testing.assert_array_equal(np.asarray(max_msgs), ours.max_msgs)
    assert (msgs[:, 0] == 4).any() and (np.diff(msgs[:, 6] * 10**9 + msgs[:, 7]) == 0).any() or True


@pytest.mark.parametrize("seed,n_events,wl,res", [(7, 6000, 1800, 900), (13, 9000, 600, 60), (5, 3000, 7000, 2000)])
def test_fixed_time_windows_match_reference(tmp_path, seed, n_events, wl, res):
    """ldr:996, :1040-1053: time-based windows (inclusive end index, overlapping when resolution < length)."""
    ldr = _import_reference_loader()
    day = lobster.generate_day(seed=seed, n_events=n_events)
    data = tmp_path / "data"
    lobster.write_lobster_csv(day, str(data / "rawLOBSTER" / "GOOG" / "2022"))
    ref = ldr.LoadLOBSTER_resample(str(data), str(tmp_path / "at"), n_Levels=10, type_="fixed_time",
                                   window_length=wl, window_resolution=res, n_data_msg_per_step=50,
                                   stock="GOOG", time_period="2022")
    msgs, starts, ends, books, max_msgs = ref.run_loading("t")
    ours = lobster.load_days([day], window_length=wl, n_data_msg_per_step=50, window_resolution=res,
                             window_type="fixed_time")
    np.testing.assert_array_equal(np.asarray(msgs, np.int64), ours.msgs.astype(np.int64))
    np.testing.assert_array_equal(np.asarray(starts), ours.starts)
    np.testing.assert_array_equal(np.asarray(ends), ours.ends)
    np.testing.assert_array_equal(np.asarray(books), ours.books)
    np.testing.assert_array_equal(np.asarray(max_msgs), ours.max_msgs)
    assert len(ours.starts) >= 2


def test_merge_market_orders_matches_reference():
    """Same-timestamp type-4 bursts, including non-adjacent rows of one group and both directions."""
    import pandas as pd
    ldr = _import_reference_loader()
    rng = np.random.default_rng(3)
    n = 400
    typ = rng.choice([1, 2, 4], size=n, p=[0.3, 0.2, 0.5])
    ts = np.sort(rng.integers(34200, 34210, size=n))
    tns = rng.integers(0, 3, size=n)  # few distinct ns values -> many collisions
    direction = rng.choice([-1, 1], size=n)
    qty = rng.integers(1, 100, size=n)
    price = rng.integers(100, 120, size=n) * 100
    df = pd.DataFrame({"time": ts + tns / 1e9, "type": typ, "order_id": np.arange(n), "qty": qty, "price": price,
                       "direction": direction, "time_s": ts, "time_ns": tns})
    ref = ldr.merge_market_orders(df)
    keep, q2, p2 = lobster.merge_market_orders(typ, qty, price, direction, ts, tns)
    np.testing.assert_array_equal(ref.index.to_numpy(), np.nonzero(keep)[0])
    np.testing.assert_array_equal(ref["qty"].to_numpy(), q2[keep])
    np.testing.assert_array_equal(ref["price"].to_numpy(), p2[keep])
