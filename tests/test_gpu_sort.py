"""GPU radix sort (Onesweep) against the oracle's stable LSD sort: bit-exact keys AND values (stability)."""
import numpy as np
import pytest

from godotgaussiansplatting_b200.rasterizer import sort_pairs
from godotgaussiansplatting_b200.synthetic import radix_keys
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 32, 33, 255, 4095, 4096, 4097, 8191, 12345, 65536, 100001, 1 << 20, (1 << 21) + 7]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("kind", ["tile_depth", "uniform32"])
def test_sort_pairs_matches_oracle(n, kind):
    keys = radix_keys(n, 1000 + n, kind)
    vals = np.arange(n, dtype=np.uint32)[::-1].copy()
    k, v = sort_pairs(keys, vals)
    rk, rv = orc.sort_pairs(keys, vals)
    np.testing.assert_array_equal(k, rk)
    np.testing.assert_array_equal(v, rv)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k, keys[order])
    np.testing.assert_array_equal(v, vals[order])


@pytest.mark.parametrize("n", [1, 4097, 1 << 20])
def test_sort_keys_only(n):
    keys = radix_keys(n, 7 + n, "uniform32")
    k = sort_pairs(keys)
    np.testing.assert_array_equal(k, np.sort(keys, kind="stable"))


def test_sort_degenerate_keys():
    for keys in (np.zeros(50000, np.uint32), np.full(50000, 0xFFFFFFFF, np.uint32),
                 np.arange(50000, dtype=np.uint32)[::-1].copy(), (np.arange(50000, dtype=np.uint32) % 3) << 24):
        vals = np.arange(keys.size, dtype=np.uint32)
        k, v = sort_pairs(keys, vals)
        order = np.argsort(keys, kind="stable")
        np.testing.assert_array_equal(k, keys[order])
        np.testing.assert_array_equal(v, vals[order])


def test_sort_large_property():
    """Full-size property check (2^25 pairs): sorted, a permutation, stable -- without the oracle."""
    n = 1 << 25
    keys = radix_keys(n, 99, "tile_depth")
    vals = np.arange(n, dtype=np.uint32)
    k, v = sort_pairs(keys, vals)
    assert np.all(k[1:] >= k[:-1])
    np.testing.assert_array_equal(keys[v], k)            # permutation consistent with the values
    ties = k[1:] == k[:-1]
    assert np.all(v[1:][ties] > v[:-1][ties])             # stability: input order kept among equal keys
    assert np.unique(v).size == n
