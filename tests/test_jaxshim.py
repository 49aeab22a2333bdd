"""Pins the NumPy-backed JAX emulation (tests/golden/jaxshim) that the golden vectors were generated on: one assertion per
JAX semantic the reference's environment code relies on, each with the JAX behaviour it restates and the reference line
that depends on it.  (jax / jaxlib cannot be installed here; what stays UNPINNED is listed at the end of this file and in
DESIGN.md 4: threefry bit streams and XLA's float reduction order.)

Sources of the stated behaviour: the JAX documentation -- "JAX - The Sharp Bits: Out-of-bounds indexing", the docstrings of
``jax.numpy.where`` / ``unique`` (``size=``, ``fill_value=``), ``jax.numpy.ndarray.at``, ``jax.lax.switch`` /
``dynamic_slice``, "Type promotion semantics" (weak types, x64 disabled), ``jax.flatten_util.ravel_pytree`` /
pytree dict ordering -- and ``jax/_src/numpy/ufuncs.py`` (``_float_divmod``, ``floor_divide``)."""
import os
import sys
import types

import numpy as np
import pytest

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jaxshim")
_OWNED = ("jax", "flax", "chex", "gymnax", "matplotlib")


@pytest.fixture(scope="module")
def J():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _OWNED}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, SHIM_DIR)
    try:
        import jax
        import jax.numpy as jnp
        from jax import lax
        from jax.flatten_util import ravel_pytree
        assert getattr(jax, "SHIM", False), "the real jax shadows the shim"
        yield types.SimpleNamespace(jax=jax, jnp=jnp, lax=lax, ravel_pytree=ravel_pytree)
    finally:
        sys.path.remove(SHIM_DIR)
        for k in [k for k in sys.modules if k.split(".")[0] in _OWNED]:
            del sys.modules[k]
        sys.modules.update(saved)


def _np(x):
    return np.asarray(x)


def test_negative_index_wraps_in_gather_and_scatter(J):
    """NumPy-style negative indices are normalised (also DYNAMIC ones): JaxOrderBookArrays.py:110-117 -- an unmatched
    cancel yields idx = -1 and ``orderside.at[idx, 1].set`` hits the LAST row (quirk Q2)."""
    x = J.jnp.arange(5, dtype=J.jnp.int32)
    idx = J.jnp.int32(-1)
    assert int(x[idx]) == 4
    assert _np(x.at[idx].set(9)).tolist() == [0, 1, 2, 3, 9]
    assert _np(x.at[J.jnp.int32(-5)].add(7)).tolist() == [7, 1, 2, 3, 4]


def test_out_of_bounds_gather_clamps_scatter_drops(J):
    """Sharp bits: "for retrieval (x[idx]) the index is clamped to the bounds of the array ... for updates (x.at[idx].set)
    out-of-bounds updates are skipped".  getCancelMsgs (JaxOrderBookArrays.py:842-853) and the trade-slot search (:205)
    index with fill values."""
    x = J.jnp.arange(5, dtype=J.jnp.int32) * 10
    assert int(x[J.jnp.int32(7)]) == 40 and int(x[J.jnp.int32(5)]) == 40
    assert _np(x.at[J.jnp.int32(5)].set(1)).tolist() == [0, 10, 20, 30, 40]
    # a negative index wraps ONCE; still out of range after that -> clamped (gather) / dropped (scatter)
    assert int(x[J.jnp.int32(-7)]) == 0
    assert _np(x.at[J.jnp.int32(-7)].set(1)).tolist() == [0, 10, 20, 30, 40]
    m = J.jnp.arange(6, dtype=J.jnp.int32).reshape(3, 2)
    assert _np(m[J.jnp.int32(3)]).tolist() == [4, 5]


def test_where_size_fill_value(J):
    """jnp.where(cond, size=k, fill_value=f)[0]: the first k true indices in order, padded with f.
    add_order's slot (JaxOrderBookArrays.py:73), cancel's id search (:110), the trade slot (:205)."""
    m = J.jnp.array([False, True, False, True])
    assert _np(J.jnp.where(m, size=1, fill_value=-1)[0]).tolist() == [1]
    assert _np(J.jnp.where(m, size=3, fill_value=-1)[0]).tolist() == [1, 3, -1]
    assert _np(J.jnp.where(J.jnp.zeros(4, bool), size=1, fill_value=-1)[0]).tolist() == [-1]
    r, c = J.jnp.where(J.jnp.array([[0, 0], [0, -1], [-1, 0]]) == -1, size=1, fill_value=-1)   # 2-D: row-major order
    assert (int(r[0]), int(c[0])) == (1, 1)


def test_unique_size_fill_value(J):
    """jnp.unique(x, size=k, fill_value=f): sorted unique values, truncated / padded to k (get_L2_state,
    JaxOrderBookArrays.py:1245-1251)."""
    x = J.jnp.array([5, 3, 5, -1, 3], dtype=J.jnp.int32)
    assert _np(J.jnp.unique(x, size=4, fill_value=-7)).tolist() == [-1, 3, 5, -7]
    assert _np(J.jnp.unique(x, size=2, fill_value=-7)).tolist() == [-1, 3]


def test_integer_floor_division_and_remainder(J):
    """``//`` on int32 is floor division (lax.div truncates, jnp.floor_divide fixes the sign up); ``%`` has the sign of the
    divisor.  Prices are floored to ticks everywhere (mm_env.py:988-1052, exec_env.py:864-878)."""
    a = J.jnp.array([7, -7, 7, -7, 0, -1], dtype=J.jnp.int32)
    b = J.jnp.array([2, 2, -2, -2, 5, 100], dtype=J.jnp.int32)
    assert _np(a // b).tolist() == [3, -4, -4, 3, 0, -1]
    assert _np(a % b).tolist() == [1, 1, -1, -1, 0, 99]
    assert (a // b).dtype == np.int32


def test_float_floor_division_follows_float_divmod(J):
    """jax/_src/numpy/ufuncs.py ``_float_divmod``: mod = fmod(x, y); div = (x - mod) / y; sign fix-up; round(div).
    The MM half-spread (mm_env.py:1028-1029) and exe's mid price (exec_env.py:871-878) go through float32 ``//``."""
    x = J.jnp.array([7.5, -7.5, 7.5, -7.5, 1500050.0, 0.0], dtype=J.jnp.float32)
    y = J.jnp.array([2.0, 2.0, -2.0, -2.0, 100.0, 3.0], dtype=J.jnp.float32)
    q = x // y
    assert q.dtype == np.float32
    assert _np(q).tolist() == [3.0, -4.0, -4.0, 3.0, 15000.0, 0.0]
    assert _np(x % y).tolist() == [1.5, 0.5, -0.5, -1.5, 50.0, 0.0]
    assert float(J.jnp.float32(1e9) // J.jnp.float32(1.0)) == float(np.float32(1e9))


def test_weak_type_promotion_with_x64_disabled(J):
    """Type promotion semantics: Python scalars are weakly typed; int32 (+) Python float -> float32; int / int ->
    float32 true division; int64 / float64 inputs are canonicalised to 32 bits (marl_env.py:22 runs with x64 off)."""
    i = J.jnp.array([3, 4], dtype=J.jnp.int32)
    f = J.jnp.array([1.5, 2.5], dtype=J.jnp.float32)
    assert (i + 1).dtype == np.int32 and (i * 2.0).dtype == np.float32 and (i / 2).dtype == np.float32
    assert (i / i).dtype == np.float32 and (f + 1e-9).dtype == np.float32 and (f * i).dtype == np.float32
    assert J.jnp.asarray(np.array([1, 2], np.int64)).dtype == np.int32
    assert J.jnp.asarray(np.array([1.0], np.float64)).dtype == np.float32
    assert J.jnp.mean(i).dtype == np.float32 and float(J.jnp.mean(i)) == 3.5
    # (a + b) / 2 on int32 prices: the sum is int32, the division float32 (marl_env.py:160, :495)
    p = J.jnp.array([1500100, 1500300], dtype=J.jnp.int32)
    assert ((p[0] + p[1]) / 2).dtype == np.float32 and float((p[0] + p[1]) / 2) == 1500200.0


def test_int32_arithmetic_wraps(J):
    """XLA integer arithmetic is two's complement; jnp.asarray narrows int64 silently with x64 off (base_env.py:184)."""
    big = J.jnp.int32(2 ** 31 - 1)
    assert int(big + 1) == -2 ** 31
    assert int(J.jnp.int32(-2 ** 31) - 1) == 2 ** 31 - 1
    assert int(J.jnp.int32(65536) * J.jnp.int32(65536)) == 0
    assert _np(J.jnp.asarray(np.array([2 ** 31 + 5], np.int64))).tolist() == [-2 ** 31 + 5]


def test_ravel_pytree_sorts_dict_keys_and_promotes(J):
    """Dict pytrees flatten in SORTED key order and ravel_pytree promotes the leaves to one dtype: the observation vectors
    (mm_env.py:2999, exec_env.py:2077; quirk Q13)."""
    flat, unravel = J.ravel_pytree({"spread": J.jnp.float32(3.0), "inventory": J.jnp.int32(2), "mid": J.jnp.float32(1.0)})
    assert flat.dtype == np.float32 and _np(flat).tolist() == [2.0, 1.0, 3.0]     # inventory, mid, spread
    leaves = J.jax.tree_util.tree_leaves({"b": 1, "a": 2, "c": {"z": 3, "y": 4}})
    assert [int(x) for x in leaves] == [2, 1, 4, 3]


def test_lax_switch_clamps_and_dynamic_slice_clamps(J):
    """lax.switch: "index ... is clamped to the range" (cond_type_side, JaxOrderBookArrays.py:596, :725);
    lax.dynamic_slice: start indices are clamped so that the slice fits (get_data_messages, base_env.py:350-352)."""
    branches = [lambda x: x + 10, lambda x: x + 20, lambda x: x + 30]
    assert int(J.lax.switch(J.jnp.int32(7), branches, J.jnp.int32(1))) == 31
    assert int(J.lax.switch(J.jnp.int32(-3), branches, J.jnp.int32(1))) == 11
    x = J.jnp.arange(10, dtype=J.jnp.int32)
    assert _np(J.lax.dynamic_slice(x, (J.jnp.int32(8),), (4,))).tolist() == [6, 7, 8, 9]
    # a negative start is first normalised NumPy-style (jax/_src/lax/slicing.py:_dynamic_slice_indices), then clamped
    assert _np(J.lax.dynamic_slice(x, (J.jnp.int32(-2),), (3,))).tolist() == [7, 8, 9]


def test_argsort_is_stable_and_float_to_int_truncates(J):
    """jnp.argsort is stable by default (the rank pairing of _filter_messages, mm_env.py:520-582); astype(int32) on floats
    truncates toward zero (lax.convert_element_type; mm_env.py:1043-1052)."""
    k = J.jnp.array([2, 1, 2, 1, 0], dtype=J.jnp.int32)
    assert _np(J.jnp.argsort(k)).tolist() == [4, 1, 3, 0, 2]
    f = J.jnp.array([2.9, -2.9, 0.5, -0.5], dtype=J.jnp.float32)
    assert _np(f.astype(J.jnp.int32)).tolist() == [2, -2, 0, 0]


def test_while_loop_cond_scan_are_the_pure_functions_they_trace(J):
    """Eager lax.while_loop / cond / scan == the traced program for pure bodies (the matching loop,
    JaxOrderBookArrays.py:295-331; the message scan :818)."""
    out = J.lax.while_loop(lambda s: s[0] < 5, lambda s: (s[0] + 1, s[1] * 2), (J.jnp.int32(0), J.jnp.int32(1)))
    assert (int(out[0]), int(out[1])) == (5, 32)
    assert int(J.lax.cond(J.jnp.array(True), lambda a: a + 1, lambda a: a - 1, J.jnp.int32(4))) == 5
    carry, ys = J.lax.scan(lambda c, x: (c + x, c * x), J.jnp.int32(0), J.jnp.arange(4, dtype=J.jnp.int32))
    assert int(carry) == 6 and _np(ys).tolist() == [0, 0, 2, 9]


def test_float_sum_order_is_the_documented_shim_choice(J):
    """UNPINNED against XLA (which leaves reduction order unspecified): the shim sums float32 strictly left to right, and the
    oracle's ``set_sum_order(True)`` mode reproduces every golden float leaf with it (tests/test_golden.py)."""
    x = J.jnp.array([1e8, 1.0, -1e8, 1.0], dtype=J.jnp.float32)
    assert float(J.jnp.sum(x)) == 1.0      # ((1e8 + 1) - 1e8) + 1 in float32 = 0 + 1; a pairwise order would give 0 or 2
