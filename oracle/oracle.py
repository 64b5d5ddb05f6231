"""ctypes binding of oracle/libgsr_oracle.so (the CPU restatement of the reference pipeline).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under godotgaussiansplatting_b200/ imports this.
Parity status: the shader path is pinned against the reference's own shaders run on the CPU (oracle/refshaders.py,
tests/test_refshaders.py); the GDScript host functions remain unpinned -- see the gsr_oracle.c header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgsr_oracle.so")
_lib = None

RECORD_DTYPE = np.dtype(
    [("image_pos", "<f4", 2), ("pos_xy", "<f4", 2), ("conic", "<f4", 3), ("pos_z", "<f4"), ("color", "<f4", 4)]
)
assert RECORD_DTYPE.itemsize == 48


class _Uniforms(C.Structure):
    _fields_ = [("camera_pos", C.c_float * 3), ("model_scale", C.c_float), ("dims", C.c_int32 * 2),
                ("time", C.c_float), ("_pad", C.c_float)]


class _FrameStats(C.Structure):
    _fields_ = [("visible", C.c_int64), ("duplicates", C.c_int64), ("staged", C.c_int64), ("last_tile", C.c_int64),
                ("ms_projection", C.c_double), ("ms_sort", C.c_double), ("ms_boundaries", C.c_double),
                ("ms_render", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gsr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, u32p, i64 = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int64
        L.orc_preprocess_ply.argtypes = [fp, i64, C.c_int, C.c_float, fp]
        L.orc_pack_camera.argtypes = [fp, fp, fp]
        L.orc_project.restype = i64
        L.orc_project.argtypes = [fp, i64, fp, C.POINTER(_Uniforms), C.c_int, C.c_int, C.c_void_p, u32p, u32p, i64,
                                  C.POINTER(i64), C.POINTER(i64)]
        L.orc_sort_pairs.argtypes = [u32p, u32p, i64]
        L.orc_sort_pairs_shader_emulation.argtypes = [u32p, u32p, i64, i64]
        L.orc_boundaries.argtypes = [u32p, i64, i64, u32p, C.c_int, i64]
        L.orc_boundaries_uninit.argtypes = [u32p, i64, i64, u32p, C.c_uint32]
        L.orc_render.argtypes = [C.c_void_p, u32p, u32p, C.c_int, C.c_int, C.c_float, C.c_uint32, C.c_int, C.c_int, fp, fp,
                                 C.POINTER(i64), u32p]
        L.orc_frame.restype = C.c_int
        L.orc_frame.argtypes = [fp, i64, fp, C.POINTER(_Uniforms), C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, u32p,
                                u32p, i64, u32p, fp, C.POINTER(_FrameStats)]
        for n in ("orc_test_exp", "orc_test_log2"):
            getattr(L, n).restype = C.c_float
            getattr(L, n).argtypes = [C.c_float]
        L.orc_test_pow.restype = C.c_float
        L.orc_test_pow.argtypes = [C.c_float, C.c_float]
        L.orc_present.argtypes = [fp, i64, C.c_int, C.c_void_p]
        L.orc_set_blend_contraction.argtypes = [C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def make_uniforms(camera_pos, model_scale, width, height, time) -> _Uniforms:
    u = _Uniforms()
    u.camera_pos[:] = [float(np.float32(c)) for c in camera_pos]
    u.model_scale = float(model_scale)
    u.dims[:] = [int(width), int(height)]
    u.time = float(time)
    u._pad = 0.0
    return u


def uniforms_from_bytes(buf) -> _Uniforms:
    raw = bytes(np.asarray(buf).tobytes())
    assert len(raw) == 32
    return _Uniforms.from_buffer_copy(raw)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def set_blend_contraction(on: bool) -> None:
    """True (default): the gsr spec (five explicit fmaf in the blend, what the CUDA compositor computes);
    False: no contraction at all -- the evaluation oracle/glsl_cpu gives the reference's own shader text."""
    lib().orc_set_blend_contraction(int(bool(on)))


def preprocess_ply(ply: np.ndarray, creation_time: float = 0.0) -> np.ndarray:
    """util/ply_file.gd:44-69. ply: (n, nprops>=62) float32 -> (n, 60) float32."""
    ply = np.ascontiguousarray(ply, dtype=np.float32)
    n, nprops = ply.shape
    out = np.empty((n, 60), dtype=np.float32)
    lib().orc_preprocess_ply(_f(ply), n, nprops, float(creation_time), _f(out))
    return out


def pack_camera(cam16: np.ndarray, proj16: np.ndarray) -> np.ndarray:
    """util/gaussian_splatting_rasterizer.gd:181-193. Inputs: Godot Projection column-major 16 floats."""
    cam16 = np.ascontiguousarray(cam16, dtype=np.float32).reshape(16)
    proj16 = np.ascontiguousarray(proj16, dtype=np.float32).reshape(16)
    out = np.empty(32, dtype=np.float32)
    lib().orc_pack_camera(_f(cam16), _f(proj16), _f(out))
    return out


@dataclass
class Projection:
    records: np.ndarray  # (n,) RECORD_DTYPE, indexed by splat id (culled entries are zero-filled here)
    keys: np.ndarray     # (M,) uint32, emission order
    values: np.ndarray   # (M,) uint32
    visible: int
    duplicates: int      # M (may exceed capacity)
    last_tile: int


def project(splat60, vp32, uniforms: _Uniforms, band=None, cap=None) -> Projection:
    splat60 = np.ascontiguousarray(splat60, dtype=np.float32).reshape(-1, 60)
    vp32 = np.ascontiguousarray(vp32, dtype=np.float32).reshape(32)
    n = splat60.shape[0]
    gy = (uniforms.dims[1] + 15) // 16
    y0, y1 = (0, gy) if band is None else band
    cap = int(cap if cap is not None else 10 * n)
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    keys = np.empty(max(cap, 1), dtype=np.uint32)
    vals = np.empty(max(cap, 1), dtype=np.uint32)
    vis, last = C.c_int64(0), C.c_int64(-1)
    m = lib().orc_project(_f(splat60), n, _f(vp32), C.byref(uniforms), int(y0), int(y1), recs.ctypes.data, _u(keys), _u(vals),
                          cap, C.byref(vis), C.byref(last))
    mm = min(int(m), cap)
    return Projection(recs, keys[:mm].copy(), vals[:mm].copy(), int(vis.value), int(m), int(last.value))


def sort_pairs(keys, values=None):
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = None if values is None else np.array(values, dtype=np.uint32, copy=True)
    lib().orc_sort_pairs(_u(k), None if v is None else _u(v), k.size)
    return (k, v) if v is not None else k


def sort_pairs_shader_emulation(keys, values, cap=None):
    n = len(keys)
    cap = int(cap if cap is not None else max(n, 1))
    k = np.zeros(2 * cap, dtype=np.uint32)
    v = np.zeros(2 * cap, dtype=np.uint32)
    k[:n] = keys
    v[:n] = values
    lib().orc_sort_pairs_shader_emulation(_u(k), _u(v), n, cap)
    return k[:n].copy(), v[:n].copy()


def boundaries(sorted_keys, num_tiles, quirks=True, global_last_tile=-1) -> np.ndarray:
    k = np.ascontiguousarray(sorted_keys, dtype=np.uint32)
    b = np.zeros((int(num_tiles), 2), dtype=np.uint32)
    lib().orc_boundaries(_u(k), k.size, int(num_tiles), _u(b), int(bool(quirks)), int(global_last_tile))
    return b


def boundaries_uninit(sorted_keys, num_tiles, garbage: int) -> np.ndarray:
    """gsplat_boundaries.glsl with the uninitialised shared word of :36 (quirk Q20) holding `garbage`."""
    k = np.ascontiguousarray(sorted_keys, dtype=np.uint32)
    b = np.zeros((int(num_tiles), 2), dtype=np.uint32)
    lib().orc_boundaries_uninit(_u(k), k.size, int(num_tiles), _u(b), int(garbage) & 0xFFFFFFFF)
    return b


def render(records, sorted_values, bounds, width, height, heatmap=0.0, target_tile=0xFFFFFFFF, band=None, pick=None,
           tile_staged=None):
    """Returns (rgba[H,W,4] float32, staged C, pick[4]).  tile_staged: optional uint32[T] filled with the
    number of instances each tile consumed before its stop rule fired."""
    recs = np.ascontiguousarray(records)
    assert recs.dtype == RECORD_DTYPE
    v = np.ascontiguousarray(sorted_values, dtype=np.uint32)
    if v.size == 0:
        v = np.zeros(1, dtype=np.uint32)
    b = np.ascontiguousarray(bounds, dtype=np.uint32)
    gy = (height + 15) // 16
    y0, y1 = (0, gy) if band is None else band
    out = np.zeros((height, width, 4), dtype=np.float32)
    pk = np.zeros(4, dtype=np.float32) if pick is None else np.array(pick, dtype=np.float32)
    staged = C.c_int64(0)
    lib().orc_render(recs.ctypes.data, _u(v), _u(b), int(width), int(height), float(heatmap), int(target_tile) & 0xFFFFFFFF,
                     int(y0), int(y1), _f(out), _f(pk), C.byref(staged), None if tile_staged is None else _u(tile_staged))
    return out, int(staged.value), pk


@dataclass
class Frame:
    rgba: np.ndarray
    records: np.ndarray
    keys: np.ndarray
    values: np.ndarray
    bounds: np.ndarray
    visible: int
    duplicates: int
    staged: int
    last_tile: int
    overflow: bool
    stage_ms: dict


def frame(splat60, vp32, uniforms: _Uniforms, heatmap=0.0, quirks=True, band=None, cap=None) -> Frame:
    """One full frame (rasterizer.gd:122-160): projection -> sort -> boundaries -> render."""
    splat60 = np.ascontiguousarray(splat60, dtype=np.float32).reshape(-1, 60)
    vp32 = np.ascontiguousarray(vp32, dtype=np.float32).reshape(32)
    n = splat60.shape[0]
    W, H = uniforms.dims[0], uniforms.dims[1]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    y0, y1 = (0, gy) if band is None else band
    cap = int(cap if cap is not None else 10 * n)
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    keys = np.zeros(max(cap, 1), dtype=np.uint32)
    vals = np.zeros(max(cap, 1), dtype=np.uint32)
    bounds = np.zeros((gx * gy, 2), dtype=np.uint32)
    out = np.zeros((H, W, 4), dtype=np.float32)
    st = _FrameStats()
    rc = lib().orc_frame(_f(splat60), n, _f(vp32), C.byref(uniforms), float(heatmap), int(bool(quirks)), int(y0), int(y1),
                         recs.ctypes.data, _u(keys), _u(vals), cap, _u(bounds), _f(out), C.byref(st))
    m = min(int(st.duplicates), cap)
    return Frame(out, recs, keys[:m].copy(), vals[:m].copy(), bounds, int(st.visible), int(st.duplicates), int(st.staged),
                 int(st.last_tile), rc != 0,
                 {"Projection": st.ms_projection, "Sort": st.ms_sort, "Boundaries": st.ms_boundaries, "Render": st.ms_render})


def present(rgba: np.ndarray, fmt: int) -> np.ndarray:
    """include/gsr.h GSR_OUT_* applied to an RGBA32F frame (main.gdshader:7-11 when fmt has 0x100): the expected bytes."""
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    px = rgba.size // 4
    shape = {0: ((px, 4), np.float32), 1: ((px, 3), np.float32), 2: ((px, 4), np.uint16), 3: ((px,), np.uint32)}[fmt & 0xFF]
    out = np.zeros(shape[0], dtype=shape[1])
    lib().orc_present(_f(rgba), px, int(fmt), out.ctypes.data)
    return out


def det_exp(x):
    return np.array([lib().orc_test_exp(float(v)) for v in np.atleast_1d(x)], dtype=np.float32)


def det_pow(x, y):
    return np.array([lib().orc_test_pow(float(v), float(y)) for v in np.atleast_1d(x)], dtype=np.float32)


def det_log2(x):
    return np.array([lib().orc_test_log2(float(v)) for v in np.atleast_1d(x)], dtype=np.float32)
