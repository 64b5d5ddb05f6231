"""The pin: the reference's OWN compute shaders, executed on the CPU, against the oracle -- and against the CUDA path.

oracle/_ref/libgsr_refshaders.so is the six .glsl files of the reference compiled for the CPU (oracle/glsl_cpu:
declarations rewrapped, every statement and expression the shader author's; workgroups as fibers with real barriers and
32-wide subgroup collectives).  oracle/refshaders.py issues the dispatches of `rasterize()` (rasterizer.gd:122-160).

What is asserted
  * projection: the 48-byte records, the emitted (key, value) pairs and M are bit-identical to gsr_oracle.c;
  * the three radix-sort shaders x 4 passes are the stable LSD sort the oracle and libgsr implement;
  * tile ranges are identical, including the reference's quirks, and the uninitialised `shared` read found this way (Q20);
  * pixels: bit-identical to the oracle's uncontracted evaluation, and within the north-star 1e-4 of the gsr spec (the five
    explicit contractions the CUDA compositor uses) -- both are legal evaluations of the GLSL text;
  * (-m gpu) the CUDA frame through the C-ABI against the reference-shader frame directly.
The libraries are built here when /root/reference is present and travel prebuilt (git-ignored) to the GPU box.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refshaders
from tests.scenes import make_scene

pytestmark = pytest.mark.skipif(not refshaders.available(), reason="oracle/_ref not built and /root/reference absent")

RGBA_TOL = 1e-4

#        n      seed  w    h    frame time  model_scale heatmap creation scale_boost
SCENES = {
    "default_camera": (20000, 3, 320, 208, None, 10.0, 1.0, 0.0, 0.0, 1.0),
    "orbit_ragged_size": (12000, 5, 250, 130, 37, 10.0, 1.0, 0.0, 0.0, 0.5),
    "load_in_animation": (8000, 7, 192, 160, None, 0.6, 1.0, 0.0, 0.0, 1.0),       # Q14: time - splat.time = 0.6
    "scaled_heatmap": (8000, 9, 224, 128, 120, 10.0, 0.5, 1.0, 0.0, 1.5),
    "three_splats": (3, 11, 64, 48, None, 10.0, 1.0, 0.0, 0.0, 2.0),
}


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def build(name):
    n, seed, w, h, frame, time, ms, heat, creation, boost = SCENES[name]
    splat60, vp, ub = make_scene(n, seed, w, h, frame=frame, time=time, model_scale=ms, creation_time=creation, scale_boost=boost)
    return splat60, vp, ub, w, h, heat


def reference_frame(splat60, vp, ub, w, h, heat, first_tile, libm=False, target_tile=-1):
    # the shared word gsplat_boundaries.glsl:36 reads uninitialised (Q20) holds the first key's tile: the author's intent
    refshaders.set_shared_fill(first_tile, libm=libm)
    return refshaders.ReferencePipeline(splat60, w, h, libm=libm).rasterize(vp, ub, heatmap=heat, target_tile=target_tile)


def oracle_frames(splat60, vp, ub, heat):
    u = orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8))
    spec = orc.frame(splat60, vp, u, heatmap=heat)
    orc.set_blend_contraction(False)
    try:
        strict = orc.frame(splat60, vp, u, heatmap=heat)
    finally:
        orc.set_blend_contraction(True)
    return spec, strict, u


def assert_stages_equal(ref, spec, splat60, vp, u):
    """ref: ReferenceFrame from the shaders; spec: oracle Frame."""
    assert not spec.overflow
    assert ref.duplicates == spec.duplicates
    pr = orc.project(splat60, vp, u)
    np.testing.assert_array_equal(ref.keys_unsorted, pr.keys)
    np.testing.assert_array_equal(ref.values_unsorted, pr.values)
    vis = np.unique(pr.values)
    assert vis.size == spec.visible
    for f in orc.RECORD_DTYPE.names:
        np.testing.assert_array_equal(bits(ref.records[f][vis]), bits(pr.records[f][vis]), err_msg=f"record field {f}")
    np.testing.assert_array_equal(ref.keys, spec.keys)
    np.testing.assert_array_equal(ref.values, spec.values)
    np.testing.assert_array_equal(ref.bounds, spec.bounds)
    m = spec.duplicates
    assert ref.grid_dims[0] == max(1, -(-m // 4096)) and ref.grid_dims[3] == max(1, -(-m // 256))   # :212-213


@pytest.mark.parametrize("name", list(SCENES))
def test_reference_shaders_equal_oracle_bit_for_bit(name):
    splat60, vp, ub, w, h, heat = build(name)
    spec, strict, u = oracle_frames(splat60, vp, ub, heat)
    assert spec.duplicates > 0
    ref = reference_frame(splat60, vp, ub, w, h, heat, int(spec.keys[0] >> 16))
    assert_stages_equal(ref, spec, splat60, vp, u)
    # pixels: the shader text without contraction == the oracle without contraction, bit for bit ...
    np.testing.assert_array_equal(bits(ref.rgba), bits(strict.rgba))
    # ... and the gsr spec (explicit contractions, what libgsr computes) is inside the north-star tolerance of it
    assert np.abs(ref.rgba - spec.rgba).max() <= RGBA_TOL
    assert np.all(ref.rgba[..., 3] == 1.0)


def test_boundaries_uninitialised_shared_word():
    """Q20: invocation 0 of workgroup 0 returns before storing local[1]; invocation 1 reads it as its left neighbour."""
    splat60, vp, ub, w, h, heat = build("orbit_ragged_size")
    spec, _, _ = oracle_frames(splat60, vp, ub, heat)
    T = spec.bounds.shape[0]
    first = int(spec.keys[0] >> 16)
    for garbage in (0xFFFFFFFF, 0, first + 1, T - 1, first):
        ref = reference_frame(splat60, vp, ub, w, h, heat, garbage)
        np.testing.assert_array_equal(ref.bounds, orc.boundaries_uninit(spec.keys, T, garbage), err_msg=f"garbage={garbage:#x}")
        if garbage == first:
            np.testing.assert_array_equal(ref.bounds, spec.bounds)       # the defined behaviour of orc_boundaries / libgsr
        elif garbage != int(spec.keys[1] >> 16):
            assert ref.bounds[int(spec.keys[1] >> 16), 0] == 1          # a range that starts at 1: instance 0 is dropped


@pytest.mark.parametrize("n", [1, 2, 255, 4095, 4096, 4097, 12289, 50000])
def test_sort_shaders_are_a_stable_lsd_sort(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 2] = keys[: n // 2] & np.uint32(0xFFFF00FF)       # many ties: stability is observable
    values = np.arange(n, dtype=np.uint32)
    k, v = refshaders.sort_pairs(keys, values, cap=n + 17)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k, keys[order])
    np.testing.assert_array_equal(v, values[order])
    ok, ov = orc.sort_pairs(keys, values)
    np.testing.assert_array_equal(k, ok)
    np.testing.assert_array_equal(v, ov)
    ek, ev = orc.sort_pairs_shader_emulation(keys, values, cap=n + 17)
    np.testing.assert_array_equal(k, ek)
    np.testing.assert_array_equal(v, ev)


def test_pick_tile():
    """gsplat_render.glsl:105-110 -> tile_splat_pos (rasterizer.gd:162-171)."""
    splat60, vp, ub, w, h, heat = build("default_camera")
    spec, _, u = oracle_frames(splat60, vp, ub, heat)
    counts = (spec.bounds[:, 1].astype(np.int64) - spec.bounds[:, 0].astype(np.int64))
    tile = int(np.argmax(counts))
    ref = reference_frame(splat60, vp, ub, w, h, heat, int(spec.keys[0] >> 16), target_tile=tile)
    _, _, pick = orc.render(spec.records, spec.values, spec.bounds, w, h, heatmap=heat, target_tile=tile)
    np.testing.assert_array_equal(bits(ref.pick), bits(pick))
    assert ref.pick[3] == counts[tile]


def test_libm_builtins_stay_inside_tolerance():
    """exp()/pow() are implementation-defined in GLSL: with glibc's expf/powf instead of the spec's polynomials the frame
    stays within the north-star tolerance (a 1-ulp pow() may move a tile rect; allow a vanishing fraction of pixels)."""
    splat60, vp, ub, w, h, heat = build("default_camera")
    spec, _, _ = oracle_frames(splat60, vp, ub, heat)
    ref = reference_frame(splat60, vp, ub, w, h, heat, int(spec.keys[0] >> 16), libm=True)
    assert abs(ref.duplicates - spec.duplicates) <= max(4, spec.duplicates // 10000)
    bad = (np.abs(ref.rgba - spec.rgba).max(axis=2) > RGBA_TOL).mean()
    assert bad <= 1e-3, bad


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["default_camera", "orbit_ragged_size", "load_in_animation", "scaled_heatmap"])
def test_cuda_path_against_reference_shaders(name):
    """libgsr (CUDA, through the C-ABI) against the reference's shaders themselves: integers bit-exact, pixels 1e-4."""
    from godotgaussiansplatting_b200 import _lib
    from tests.gsr_direct import Ctx

    n, seed, w, h, frame, time, ms, heat, creation, boost = SCENES[name]
    splat60, vp, ub, w, h, heat = build(name)
    with Ctx(n, w, h) as c:
        c.upload(splat60)
        c.keep_unsorted()
        rgba = c.render(vp, ub, heatmap=heat)
        t = c.taps()
        ukeys = c.copy(_lib.GSR_BUF_KEYS_UNSORTED, t["m"], np.uint32)
        uvals = c.copy(_lib.GSR_BUF_VALUES_UNSORTED, t["m"], np.uint32)
    assert t["m"] > 0
    ref = reference_frame(splat60, vp, ub, w, h, heat, int(t["keys"][0] >> 16))
    assert t["m"] == ref.duplicates
    np.testing.assert_array_equal(ukeys, ref.keys_unsorted)
    np.testing.assert_array_equal(uvals, ref.values_unsorted)
    np.testing.assert_array_equal(t["keys"], ref.keys)
    np.testing.assert_array_equal(t["values"], ref.values)
    np.testing.assert_array_equal(t["bounds"], ref.bounds)
    vis = np.unique(ref.values)
    for f in orc.RECORD_DTYPE.names:
        np.testing.assert_array_equal(bits(t["records"][f][vis]), bits(ref.records[f][vis]), err_msg=f"record field {f}")
    assert np.abs(rgba - ref.rgba).max() <= RGBA_TOL
