"""Front / back overlap of consecutive frames (gsr_debug_pipeline) A/B on one GPU: c3 orbit frames through gsr_render_async, device-resident
and with the RGBA32F read-back, overlap off / on; every frame of the overlapped run is compared bit for bit with the serial run.
    python ubench/pipeline_ab.py [frames] [workload]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tests.gsr_direct import Ctx
import ctypes as C
from godotgaussiansplatting_b200 import _lib

F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
wl = dict(bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "c3"]); n, w, h = wl["n"], wl["w"], wl["h"]
ctx = Ctx(n, w, h)
for lo, blk in bench.raw_chunks(wl):
    ctx.upload_ply_raw(blk, first=lo)
frames = bench.frame_params(wl, F + 10)
hosts = [torch.zeros((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]


def hist(k):
    buf = (_lib.GsrFrameRecord * k)(); got = C.c_uint32(0)
    _lib.check(ctx.L.gsr_get_frame_history(ctx.h, k, buf, C.byref(got)), "hist")
    ms = np.array([[buf[i].stage_ms[j] for j in range(5)] + [buf[i].front_ms] for i in range(got.value)])
    return ms.mean(axis=0).round(3).tolist()


def run(overlap, e2e):
    _lib.check(ctx.L.gsr_debug_pipeline(ctx.h, overlap), "pipeline")
    for i in range(10):
        ctx.render_async(*frames[i], host_ptr=hosts[i & 1].data_ptr() if e2e else None)
    ctx.sync()
    t = time.perf_counter()
    for i in range(10, 10 + F):
        ctx.render_async(*frames[i], host_ptr=hosts[i & 1].data_ptr() if e2e else None)
    t_enq = time.perf_counter() - t
    ctx.sync()
    t = time.perf_counter() - t
    print(f"overlap {overlap} {'e2e RGBA32F' if e2e else 'device    '}: {1e3 * t / F:.3f} ms/frame = {F / t:7.1f} fps   host enqueue {1e3 * t_enq / F:.3f} ms/frame   "
          f"stages [Projection, Sort, Ranges, Render, sum, front] {hist(min(F, 256))}", flush=True)


for rep in range(2):
    for e2e in (False, True):
        for overlap in (0, 1):
            run(overlap, e2e)

# bit-exactness of the overlapped pipeline: 24 back-to-back frames read back, against the same frames rendered serially
outs = {}
for overlap in (0, 1):
    _lib.check(ctx.L.gsr_debug_pipeline(ctx.h, overlap), "pipeline")
    bufs = [torch.zeros((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(24)]
    for i in range(24):
        ctx.render_async(*frames[10 + 7 * i], host_ptr=bufs[i].data_ptr())
    ctx.sync()
    outs[overlap] = [b.numpy().copy() for b in bufs]
same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(outs[0], outs[1]))
print("overlapped frames bit-identical to serial frames:", same, flush=True)
assert same
