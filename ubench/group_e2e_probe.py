"""Does a frame read-back in flight (D2H DMA) slow the group path's kernels?  G in-process ranks on one GPU, c3: per-stage GPU times of
rank 0 without any read-back, with rows-local read-back, and with the frame read back from rank 0.
    CUDA_DEVICE_MAX_CONNECTIONS=32 [GSR_LIB_PATH=...] python ubench/group_e2e_probe.py [G]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tests.gsr_direct import Ctx
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = dict(bench.WORKLOADS["c3"]); n, w, h = wl["n"], wl["w"], wl["h"]
ctxs = [Ctx(n, w, h) for _ in range(G)]
for lo, blk in bench.raw_chunks(wl):
    for c in ctxs: c.upload_ply_raw(blk, first=lo)
blobs = b"".join(c.group_export() for c in ctxs)
for r, c in enumerate(ctxs): c.group_attach(r, G, blobs)
frames = bench.frame_params(wl, 40, first=10)
hosts = [torch.zeros((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
def run(mode):
    for c in ctxs: c.sync()
    for c in ctxs: c.group_set_present(mode == "rows")
    for i, (vp, ub) in enumerate(frames):
        for c in ctxs: c.render_async(vp, ub)
        if mode == "rows":
            for c in ctxs: c.readback_rows_async(hosts[i & 1].data_ptr())
        elif mode == "root":
            ctxs[0].readback_async(hosts[i & 1].data_ptr())
    for c in ctxs: c.sync()
    import ctypes as C
    from godotgaussiansplatting_b200 import _lib
    res = []
    for c in ctxs[:2]:
        buf = (_lib.GsrFrameRecord * 30)(); got = C.c_uint32(0)
        _lib.check(c.L.gsr_get_frame_history(c.h, 30, buf, C.byref(got)), "hist")
        ms = np.array([[buf[i].stage_ms[k] for k in range(5)] for i in range(got.value)])
        res.append(ms.mean(axis=0).round(3).tolist())
    probe = ""
    if hasattr(ctxs[0].L, "gsr_debug_group_probe"):   # ubench builds (-DGSR_GROUP_PROBE): scatter kernel / segment wait / gather
        out = (C.c_float * 3)()
        ctxs[0].L.gsr_debug_group_probe.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        ps = []
        for c in ctxs[:2]:
            ctxs[0].L.gsr_debug_group_probe(c.h, out); ps.append([round(out[k], 3) for k in range(3)])
        probe = f"  scatter/wait/gather ms: rank0 {ps[0]} rank1 {ps[1]}"
    print(f"[{os.environ.get('GSR_LIB_PATH', 'libgsr.so').split('/')[-1]}] G={G} read-back {mode:5s}: rank0 stages {res[0]}  rank1 {res[1]}{probe}", flush=True)
for mode in ("none", "rows", "root", "none"):
    run(mode)
