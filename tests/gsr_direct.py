"""Thin test helper that drives the C-ABI directly (no Python mirror in between)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from godotgaussiansplatting_b200 import _lib

REC_DTYPE = np.dtype([("image_pos", "<f4", 2), ("pos_xy", "<f4", 2), ("conic", "<f4", 3), ("pos_z", "<f4"), ("color", "<f4", 4)])


class Ctx:
    def __init__(self, max_splats, width, height, flags=0, factor=10, device=0):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        cfg = _lib.GsrConfig(device, flags, max_splats, factor, 0)
        _lib.check(self.L.gsr_create(C.byref(cfg), C.byref(self.h)), "gsr_create")
        self.max_splats, self.w, self.hgt = max_splats, width, height
        _lib.check(self.L.gsr_resize(self.h, width, height), "gsr_resize")

    def close(self):
        if self.h:
            self.L.gsr_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, splat60, first=0):
        s = np.ascontiguousarray(splat60, dtype=np.float32)
        _lib.check(self.L.gsr_upload_splats_aos(self.h, s.ctypes.data_as(C.POINTER(C.c_float)), first, s.shape[0]), "upload")

    def upload_ply_raw(self, table, first=0, creation_time=0.0):
        t = np.ascontiguousarray(table, dtype=np.float32)
        _lib.check(self.L.gsr_upload_ply_raw(self.h, t.ctypes.data_as(C.POINTER(C.c_float)), t.shape[1], first, t.shape[0], float(creation_time)), "upload raw")

    def resize(self, w, h):
        _lib.check(self.L.gsr_resize(self.h, w, h), "gsr_resize")
        self.w, self.hgt = w, h

    def set_band(self, y0, y1):
        _lib.check(self.L.gsr_set_band(self.h, y0, y1), "gsr_set_band")

    def set_row_interleave(self, rem, mod):
        _lib.check(self.L.gsr_set_row_interleave(self.h, rem, mod), "gsr_set_row_interleave")

    def band_fixup(self):
        _lib.check(self.L.gsr_band_fixup(self.h), "gsr_band_fixup")
        _lib.check(self.L.gsr_sync(self.h), "gsr_sync")

    def sync_word(self, value=None):
        """Read (or overwrite: emulates the all-reduce) the int32 at gsr_band_sync_word()."""
        rt = C.CDLL("libcudart.so")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _lib.check(self.L.gsr_sync(self.h), "gsr_sync")
        ptr = self.L.gsr_band_sync_word(self.h)
        v = C.c_int32(0 if value is None else int(value))
        if value is None:
            assert rt.cudaMemcpy(C.byref(v), C.c_void_p(ptr), 4, 2) == 0
        else:
            assert rt.cudaMemcpy(C.c_void_p(ptr), C.byref(v), 4, 1) == 0
        return int(v.value)

    def keep_unsorted(self):
        _lib.check(self.L.gsr_debug_keep_unsorted(self.h, 1), "keep_unsorted")

    def render(self, vp, uniforms, heatmap=0.0, readback=True):
        vp = np.ascontiguousarray(vp, dtype=np.float32)
        out = np.empty((self.hgt, self.w, 4), dtype=np.float32) if readback else None
        _lib.check(self.L.gsr_render(self.h, vp.ctypes.data_as(C.POINTER(C.c_float)), uniforms, float(heatmap),
                                     None if out is None else C.c_void_p(out.ctypes.data)), "gsr_render")
        return out

    # ---- multi-GPU shard group (gsr_group_*) ----
    def group_export(self) -> bytes:
        buf = (C.c_ubyte * _lib.GSR_GROUP_BLOB_BYTES)()
        _lib.check(self.L.gsr_group_export(self.h, buf), "gsr_group_export")
        return bytes(buf)

    def group_attach(self, rank, world, blobs: bytes):
        buf = (C.c_ubyte * len(blobs)).from_buffer_copy(blobs)
        _lib.check(self.L.gsr_group_attach(self.h, rank, world, buf), "gsr_group_attach")

    def group_set_present(self, rows_local):
        _lib.check(self.L.gsr_group_set_present(self.h, int(bool(rows_local))), "gsr_group_set_present")

    def readback_rows_async(self, host_ptr):
        _lib.check(self.L.gsr_readback_rows_async(self.h, C.c_void_p(host_ptr)), "gsr_readback_rows_async")

    def render_async(self, vp, uniforms, heatmap=0.0, host_ptr=None):
        vp = np.ascontiguousarray(vp, dtype=np.float32)
        _lib.check(self.L.gsr_render_async(self.h, vp.ctypes.data_as(C.POINTER(C.c_float)), uniforms, float(heatmap),
                                           None if host_ptr is None else C.c_void_p(host_ptr)), "gsr_render_async")

    def readback_async(self, host_ptr, fmt=0):
        _lib.check(self.L.gsr_readback_async(self.h, C.c_void_p(host_ptr), int(fmt)), "gsr_readback_async")

    def render_async_fmt(self, vp, uniforms, host_ptr, fmt, heatmap=0.0):
        vp = np.ascontiguousarray(vp, dtype=np.float32)
        _lib.check(self.L.gsr_render_async_fmt(self.h, vp.ctypes.data_as(C.POINTER(C.c_float)), uniforms, float(heatmap), C.c_void_p(host_ptr), int(fmt)),
                   "gsr_render_async_fmt")

    def sync(self):
        _lib.check(self.L.gsr_sync(self.h), "gsr_sync")

    def stats(self):
        st = _lib.GsrStats()
        _lib.check(self.L.gsr_get_stats(self.h, C.byref(st)), "gsr_get_stats")
        return st

    def copy(self, which, count, dtype):
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            _lib.check(self.L.gsr_debug_copy(self.h, which, C.c_void_p(out.ctypes.data), out.nbytes), "gsr_debug_copy")
        return out

    def pick(self, tile_id, heatmap=0.0):
        out = (C.c_float * 4)()
        _lib.check(self.L.gsr_pick(self.h, tile_id, float(heatmap), out), "gsr_pick")
        return np.array(list(out), dtype=np.float32)

    def taps(self):
        """All stage outputs of the last frame."""
        st = self.stats()
        m = int(min(st.duplicates, st.capacity))
        T = st.tiles_x * st.tiles_y
        return dict(stats=st, m=m,
                    records=self.copy(_lib.GSR_BUF_RECORDS, self.max_splats, REC_DTYPE),
                    keys=self.copy(_lib.GSR_BUF_KEYS, m, np.uint32), values=self.copy(_lib.GSR_BUF_VALUES, m, np.uint32),
                    bounds=self.copy(_lib.GSR_BUF_BOUNDS, T * 2, np.uint32).reshape(T, 2))
