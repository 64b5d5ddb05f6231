"""A shard group of G contexts in ONE process on ONE GPU at the c3 size: the kernels of all ranks run on the one device, so an
`ncu --metrics gpu__time_duration.sum` launch list of this script gives the per-rank cost of every kernel of the group path
(extent kernel, table-mode projection, sort, compositor, flag kernels) without a multi-GPU box.
    CUDA_DEVICE_MAX_CONNECTIONS=32 python ubench/group_inprocess_c3.py G [frames] [workload]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tests.gsr_direct import Ctx

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = dict(bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "c3"])
n, w, h = wl["n"], wl["w"], wl["h"]
ctxs = [Ctx(n, w, h) for _ in range(G)]
for lo, blk in bench.raw_chunks(wl):
    for c in ctxs:
        c.upload_ply_raw(blk, first=lo)
blobs = b"".join(c.group_export() for c in ctxs)
for r, c in enumerate(ctxs):
    c.group_attach(r, G, blobs)
frames = bench.frame_params(wl, nframes, first=10)
for vp, ub in frames:
    for c in ctxs:
        c.render_async(vp, ub)
    for c in ctxs:
        c.sync()
for r, c in enumerate(ctxs):
    st = c.stats()
    print(f"rank {r}: M {st.duplicates} V {st.visible} stage ms {[round(x, 3) for x in st.stage_ms]}")
for c in ctxs:
    c.close()
