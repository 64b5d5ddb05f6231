"""Experiment: throughput with 1 vs 2 frames in flight (two contexts on two streams, alternating orbit frames)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from godotgaussiansplatting_b200.camera import default_camera
from godotgaussiansplatting_b200.ply_file import PlyFile
from godotgaussiansplatting_b200.rasterizer import GaussianSplattingRasterizer, RenderTexture
wl = dict(bench.WORKLOADS["c3"])
stub = PlyFile(); stub.size = wl["n"]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rs, streams = [], []
for k in range(K):
    r = GaussianSplattingRasterizer(stub, (wl["w"], wl["h"]), RenderTexture(), default_camera())
    r.init_gpu(load=False)
    st = torch.cuda.Stream(); r.set_stream(st.cuda_stream)
    rs.append(r); streams.append(st)
for lo, s60 in bench.scene_chunks(wl):
    for r in rs: r.upload_splats(s60, lo)
frames = bench.frame_params(wl, 130)
def run(nctx, steps=120, warm=10):
    for i in range(warm):
        vp, ub = frames[i]; rs[i % nctx].render_raw(vp, ub, 0.0, None, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(warm, warm + steps):
        vp, ub = frames[i]; rs[i % nctx].render_raw(vp, ub, 0.0, None, True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f"{nctx} frame(s) in flight: {dt:.3f} ms/frame  {1000/dt:.1f} fps", flush=True)
for n in range(1, K + 1): run(n)
for n in range(1, K + 1): run(n)
