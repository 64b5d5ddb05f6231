"""libgsr's sort, tile-range and compositor kernels, compiled for the CPU (tests/kernel_emu), against the oracle -- bit for bit.

This is NOT a CPU path of the product (libgsr has none; see tests/test_abi.py): it is a pre-flight check of kernel LOGIC.
csrc/radix_sort.cu, csrc/ranges.cu and csrc/compositor.cu are compiled by g++ under a thin shim of the CUDA execution model (threads = fibers,
__syncthreads / warp collectives real, __shared__ = block-shared, packed f32x2 PTX = two IEEE binary32 operations), and one
persistent block works through every tile: staging, blend, tile-stop vote, quantum, spill, re-queue, resume.
It lets a kernel variant that has never seen a GPU (GSR_COMP_V2) prove its indexing and buffering before GPU minutes are spent.
What it cannot show: memory-model behaviour (fences, races between blocks) and timing.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import oracle as orc
from tests.scenes import make_scene

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_emu"))
try:
    import build as _emu_build
finally:
    sys.path.pop(0)

SHIPPED, V2, HWEXP, P4 = 0, 1, 2, 3
_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(_emu_build.build())
        L.emu_composite.restype = C.c_int
        L.emu_composite.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_uint32, C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint)]
        L.emu_tile_ranges.restype = C.c_int
        L.emu_tile_ranges.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_int]
        L.emu_sort_pairs.restype = C.c_int
        L.emu_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.emu_band_fixup.restype = C.c_int
        L.emu_band_fixup.argtypes = [C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _L = L
    return _L


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def emu_ranges(keys, T, quirks=1, sharded=0, global_last=-1, grid=7):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    bounds = np.full((T, 2), 0xDEADBEEF, dtype=np.uint32)
    sync = np.zeros(1, dtype=np.int32)
    assert lib().emu_tile_ranges(keys.ctypes.data, keys.size, bounds.ctypes.data, T, quirks, sharded, global_last,
                                 sync.ctypes.data if sharded == 2 else None, grid) == 0
    return bounds, int(sync[0])


def emu_composite(variant, records, values, bounds, w, h, heat=0.0, target=0xFFFFFFFF, tile_begin=0, row_step=1, num_tiles=None, out=None):
    gx, gy = (w + 15) // 16, (h + 15) // 16
    out = np.zeros((h, w, 4), dtype=np.float32) if out is None else out
    pick = np.zeros(4, dtype=np.float32)
    staged, pushes = C.c_ulonglong(0), C.c_uint(0)
    vals = np.concatenate([np.asarray(values, dtype=np.uint32), np.zeros(512, dtype=np.uint32)])   # the kernels never read past a range
    recs, bnds = np.ascontiguousarray(records), np.ascontiguousarray(bounds, dtype=np.uint32)
    rc = lib().emu_composite(variant, recs.ctypes.data, vals.ctypes.data, bnds.ctypes.data, out.ctypes.data, w, h, tile_begin, row_step,
                             gx * gy if num_tiles is None else num_tiles, heat, target, pick.ctypes.data, C.byref(staged), C.byref(pushes))
    assert rc == 0, "not every tile was finished"
    return out, int(staged.value), int(pushes.value), pick


def oracle_frame(n, seed, w, h, heat=0.0, **kw):
    splat60, vp, ub = make_scene(n, seed, w, h, **kw)
    return orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), heatmap=heat, cap=40 * n)


#                n      seed w    h    heat kwargs
CASES = {
    "long_lists": (60000, 3, 320, 208, 0.0, dict(scale_boost=1.5)),     # 14 000-instance lists: many chunks, quantum, re-queues
    "ragged_heatmap": (9000, 5, 250, 130, 1.0, dict(frame=37, scale_boost=0.5)),
    "load_in": (6000, 7, 192, 160, 0.0, dict(time=0.6, scale_boost=1.0)),
    "tiny": (40, 9, 33, 17, 0.0, dict(scale_boost=-1.0)),
}


@pytest.mark.parametrize("variant", [SHIPPED, V2, P4], ids=["shipped", "v2", "p4"])
@pytest.mark.parametrize("case", list(CASES))
def test_compositor_kernels_reproduce_the_oracle(case, variant):
    n, seed, w, h, heat, kw = CASES[case]
    fr = oracle_frame(n, seed, w, h, heat, **kw)
    assert not fr.overflow
    out, staged, pushes, _ = emu_composite(variant, fr.records, fr.values, fr.bounds, w, h, heat)
    np.testing.assert_array_equal(bits(out), bits(fr.rgba))
    assert staged == fr.staged
    if case == "long_lists":
        assert pushes > 10           # the hand-back path really ran


def test_hwexp_variant_stays_inside_the_tolerance():
    n, seed, w, h, heat, kw = CASES["long_lists"]
    fr = oracle_frame(n, seed, w, h, heat, **kw)
    out, staged, _, _ = emu_composite(HWEXP, fr.records, fr.values, fr.bounds, w, h, heat)   # exp2f stands in for MUFU.EX2
    assert np.abs(out - fr.rgba).max() <= 1e-4 and staged == fr.staged


@pytest.mark.parametrize("variant", [SHIPPED, V2, P4], ids=["shipped", "v2", "p4"])
def test_pick_and_row_interleave(variant):
    n, seed, w, h = 20000, 15, 320, 240
    fr = oracle_frame(n, seed, w, h, scale_boost=1.0)
    counts = fr.bounds[:, 1].astype(np.int64) - fr.bounds[:, 0]
    busy = int(np.argmax(counts))
    _, _, _, pick = emu_composite(variant, fr.records, fr.values, fr.bounds, w, h, target=busy)
    _, _, want = orc.render(fr.records, fr.values, fr.bounds, w, h, target_tile=busy, pick=np.zeros(4, np.float32))
    np.testing.assert_array_equal(bits(pick), bits(want))
    # cyclic tile rows (gsr_set_row_interleave): three "ranks" fill one frame
    gx, gy = (w + 15) // 16, (h + 15) // 16
    out = np.zeros((h, w, 4), dtype=np.float32)
    for rem in range(3):
        rows = len(range(rem, gy, 3))
        emu_composite(variant, fr.records, fr.values, fr.bounds, w, h, tile_begin=rem * gx, row_step=3, num_tiles=rows * gx, out=out)
    np.testing.assert_array_equal(bits(out), bits(fr.rgba))


@pytest.mark.parametrize("grid", [1, 7, 64])
def test_tile_ranges_kernel(grid):
    fr = oracle_frame(20000, 3, 320, 208, scale_boost=1.0)
    T = fr.bounds.shape[0]
    got, _ = emu_ranges(fr.keys, T, quirks=1, grid=grid)
    np.testing.assert_array_equal(got, fr.bounds)
    fixed, _ = emu_ranges(fr.keys, T, quirks=0, grid=grid)
    np.testing.assert_array_equal(fixed, orc.boundaries(fr.keys, T, quirks=False))
    # last grid tile occupied: the M-1 rule (gsplat_boundaries.glsl:47-49)
    keys = np.sort(np.concatenate([fr.keys, np.full(5, ((T - 1) << 16) | 7, dtype=np.uint32)]))
    got, _ = emu_ranges(keys, T, quirks=1, grid=grid)
    np.testing.assert_array_equal(got, orc.boundaries(keys, T, quirks=True))
    # empty list
    got, _ = emu_ranges(np.zeros(0, np.uint32), T, grid=grid)
    assert not got.any()
    # fast sharded mode: local last tile -> sync word, range end written
    got, sync = emu_ranges(fr.keys, T, quirks=1, sharded=2, grid=grid)
    last = int(fr.keys[-1] >> 16)
    assert sync == last + 1 and got[last, 1] == fr.keys.size


def test_band_fixup_kernel_blanks_only_the_owned_last_tile():
    w, h = 100, 70
    gx = (w + 15) // 16
    img = np.ones((h, w, 4), dtype=np.float32)
    L = 2 * gx + 3                                     # tile row 2, column 3
    lib().emu_band_fixup(L + 1, img.ctypes.data, w, h, 0, 5, 3, 1)     # row 2 % 3 != 1: another rank owns it
    assert (img == 1).all()
    lib().emu_band_fixup(L + 1, img.ctypes.data, w, h, 0, 5, 3, 2)
    blank = np.zeros_like(img, dtype=bool)
    blank[32:48, 48:64] = True
    assert (img[blank[..., 0]] == [0, 0, 0, 1]).all() and (img[~blank[..., 0]] == 1).all()
    T = gx * ((h + 15) // 16)
    img[:] = 1
    lib().emu_band_fixup(T, img.ctypes.data, w, h, 0, 5, 1, 0)         # last occupied tile == T-1: the other rule applies, nothing blanked
    assert (img == 1).all()


@pytest.mark.parametrize("n", [1, 31, 100, 6143, 6144, 6145, 20000, 100000])
@pytest.mark.parametrize("pairs", [True, False], ids=["pairs", "keys"])
def test_onesweep_kernels_are_a_stable_sort(n, pairs):
    """sort_hist_kernel + 4 x onesweep_kernel<512, 12>: tile = 6144 keys, ragged last tile, padding keys, look-back chain."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 2] &= np.uint32(0x00FF00FF)                      # heavy ties and two constant digits
        keys[n // 2: n // 2 + 50] = 0xFFFFFFFF                        # real keys equal to the padding key
    n_max = n + 777
    k = np.zeros(n_max, dtype=np.uint32)
    k[:n] = keys
    k[n:] = 0x12345678                                               # beyond n: must not be touched or read as data
    v = np.arange(n_max, dtype=np.uint32) if pairs else None
    assert lib().emu_sort_pairs(k.ctypes.data, v.ctypes.data if pairs else None, n, n_max, 5) == 0
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k[:n], keys[order])
    assert (k[n:] == 0x12345678).all()
    if pairs:
        np.testing.assert_array_equal(v[:n], order.astype(np.uint32))


def test_sorted_frame_through_the_emulated_kernels():
    """oracle projection -> emulated sort -> emulated ranges -> emulated compositor == the oracle frame."""
    n, w, h = 15000, 256, 144
    splat60, vp, ub = make_scene(n, 21, w, h, scale_boost=1.0)
    u = orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8))
    fr = orc.frame(splat60, vp, u)
    pr = orc.project(splat60, vp, u)
    cap = 10 * n
    k = np.zeros(cap, dtype=np.uint32)
    v = np.zeros(cap, dtype=np.uint32)
    k[: pr.duplicates], v[: pr.duplicates] = pr.keys, pr.values
    assert lib().emu_sort_pairs(k.ctypes.data, v.ctypes.data, pr.duplicates, cap, 3) == 0
    np.testing.assert_array_equal(k[: pr.duplicates], fr.keys)
    np.testing.assert_array_equal(v[: pr.duplicates], fr.values)
    bounds, _ = emu_ranges(k[: pr.duplicates], fr.bounds.shape[0])
    np.testing.assert_array_equal(bounds, fr.bounds)
    out, staged, _, _ = emu_composite(SHIPPED, pr.records, v[: pr.duplicates], bounds, w, h)
    np.testing.assert_array_equal(bits(out), bits(fr.rgba))
    assert staged == fr.staged
