"""Edge cases of the hot path through the C-ABI: ragged sizes, single elements, limits, error behaviour."""
import ctypes as C

import numpy as np
import pytest

from godotgaussiansplatting_b200 import _lib
from oracle import oracle as orc
from tests.gsr_direct import Ctx
from tests.scenes import make_scene

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def both(n, seed, w, h, **kw):
    splat60, vp, ub = make_scene(n, seed, w, h, **kw)
    factor = max(10, 200000 // n + 1)  # capacity floor: a handful of big splats must not overflow 10*N
    ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), cap=factor * n)
    assert not ref.overflow
    with Ctx(n, w, h, factor=factor) as c:
        c.upload(splat60)
        img = c.render(vp, ub)
        t = c.taps()
    return ref, img, t


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (15, 15), (16, 16), (17, 17), (31, 33), (1920, 16), (16, 1080), (257, 1)])
def test_ragged_resolutions(w, h):
    ref, img, t = both(3000, 50, w, h, scale_boost=1.0)
    assert t["stats"].duplicates == ref.duplicates
    np.testing.assert_array_equal(t["keys"], ref.keys)
    np.testing.assert_array_equal(t["bounds"], ref.bounds)
    np.testing.assert_array_equal(bits(img), bits(ref.rgba))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 257, 1000])
def test_splat_counts_around_warp_and_cta_sizes(n):
    ref, img, t = both(n, 60 + n, 320, 240, scale_boost=2.0)
    assert t["stats"].duplicates == ref.duplicates and t["stats"].visible == ref.visible
    np.testing.assert_array_equal(t["keys"], ref.keys)
    np.testing.assert_array_equal(t["values"], ref.values)
    np.testing.assert_array_equal(bits(img), bits(ref.rgba))


def test_single_huge_splat_covers_every_tile():
    """One splat whose rect is the whole grid: the warp-cooperative emit path with thousands of tiles from one lane."""
    w, h = 1920, 1080
    splat60, vp, ub = make_scene(4, 70, w, h)
    splat60[:, 0:3] = (0.0, 0.0, 2.5)
    splat60[:, 4:10] = 0.0
    splat60[0, 4], splat60[0, 7], splat60[0, 9] = 4.0, 4.0, 4.0   # sigma = 2 units at distance 2.5: covers the screen
    splat60[1:, 4], splat60[1:, 7], splat60[1:, 9] = 1e-6, 1e-6, 1e-6
    splat60[:, 10] = 0.5
    ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), cap=40000)
    assert ref.duplicates >= 8160 and not ref.overflow
    with Ctx(4, w, h, factor=10000) as c:
        c.upload(splat60)
        img = c.render(vp, ub)
        t = c.taps()
    np.testing.assert_array_equal(t["keys"], ref.keys)
    np.testing.assert_array_equal(t["values"], ref.values)
    np.testing.assert_array_equal(t["bounds"], ref.bounds)
    np.testing.assert_array_equal(bits(img), bits(ref.rgba))


def test_tile_id_limit_and_argument_errors():
    L = _lib.lib()
    with Ctx(100, 64, 64) as c:
        assert L.gsr_resize(c.h, 4096, 4096) == _lib.GSR_OK             # 65536 tiles: the 16-bit tile id is full
        assert L.gsr_resize(c.h, 4112, 4096) == _lib.GSR_ERR_INVALID     # 65792 tiles would alias in the sort key
        assert L.gsr_resize(c.h, 0, 10) == _lib.GSR_ERR_INVALID
        assert L.gsr_resize(c.h, 64, 64) == _lib.GSR_OK
        s = np.zeros((10, 60), dtype=np.float32)
        fp = s.ctypes.data_as(C.POINTER(C.c_float))
        assert L.gsr_upload_splats_aos(c.h, fp, 95, 10) == _lib.GSR_ERR_INVALID      # past max_splats
        assert L.gsr_upload_splats_aos(c.h, fp, 2**64 - 1, 2) == _lib.GSR_ERR_INVALID  # first + count wraps around
        out4 = (C.c_float * 4)()
        assert L.gsr_pick(c.h, 0, 0.0, out4) == _lib.GSR_ERR_STATE                    # no frame rendered at this size yet
        assert L.gsr_upload_ply_raw(c.h, fp, 61, 0, 5, 0.0) == _lib.GSR_ERR_INVALID   # fewer than the 62 standard properties
        vp = np.zeros(32, dtype=np.float32)
        ub = bytearray(32)
        ub[16:24] = np.array([65, 64], dtype=np.int32).tobytes()                       # dims differ from gsr_resize
        assert L.gsr_render(c.h, vp.ctypes.data_as(C.POINTER(C.c_float)), bytes(ub), 0.0, None) == _lib.GSR_ERR_INVALID
        assert b"differ" in L.gsr_last_error()
        assert L.gsr_set_band(c.h, 3, 2) == _lib.GSR_ERR_INVALID
        assert L.gsr_set_row_interleave(c.h, 4, 4) == _lib.GSR_ERR_INVALID
    ctx = C.c_void_p()
    cfg = _lib.GsrConfig(99, 0, 100, 10, 0)
    assert L.gsr_create(C.byref(cfg), C.byref(ctx)) == _lib.GSR_ERR_INVALID           # no such device ordinal


def test_render_before_resize_is_a_state_error():
    L = _lib.lib()
    ctx = C.c_void_p()
    cfg = _lib.GsrConfig(0, 0, 100, 10, 0)
    _lib.check(L.gsr_create(C.byref(cfg), C.byref(ctx)), "create")
    try:
        vp = np.zeros(32, dtype=np.float32)
        assert L.gsr_render(ctx, vp.ctypes.data_as(C.POINTER(C.c_float)), bytes(32), 0.0, None) == _lib.GSR_ERR_STATE
        out = (C.c_float * 4)()
        assert L.gsr_pick(ctx, 0, 0.0, out) == _lib.GSR_ERR_STATE
    finally:
        L.gsr_destroy(ctx)


def test_nan_and_inf_inputs_do_not_break_the_frame():
    """Reference behaviour for NaN positions is undefined (int(NaN)); the gsr spec culls non-finite image positions.  The
    frame must stay finite where the oracle's is, keys must match, and nothing may hang."""
    n, w, h = 2000, 320, 240
    splat60, vp, ub = make_scene(n, 80, w, h)
    splat60[5, 0] = np.nan
    splat60[6, 1] = np.inf
    splat60[7, 2] = -np.inf
    splat60[8, 10] = 0.0          # opacity exactly 0 -> pow(0, 0.2) = 0 -> radius 0
    splat60[9, 4:10] = 0.0        # degenerate covariance (only the +0.3 dilation remains)
    ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
    with Ctx(n, w, h) as c:
        c.upload(splat60)
        img = c.render(vp, ub)
        t = c.taps()
    np.testing.assert_array_equal(t["keys"], ref.keys)
    np.testing.assert_array_equal(t["values"], ref.values)
    np.testing.assert_array_equal(bits(img), bits(ref.rgba))
    assert not np.isin([5, 6, 7], t["values"]).any()
