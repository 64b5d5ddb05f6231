"""One GPU emulating rank r of G (cyclic tile rows): per-stage times of the c3 frame.  python ubench/emulate_rank.py G [r]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from godotgaussiansplatting_b200.camera import default_camera
from godotgaussiansplatting_b200.ply_file import PlyFile
from godotgaussiansplatting_b200.rasterizer import GaussianSplattingRasterizer, RenderTexture
G = int(sys.argv[1]); r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wl = dict(bench.WORKLOADS["c3"])
stub = PlyFile(); stub.size = wl["n"]
r = GaussianSplattingRasterizer(stub, (wl["w"], wl["h"]), RenderTexture(), default_camera())
r.init_gpu(load=False)
for lo, blk in bench.raw_chunks(wl):
    r.upload_ply_raw(blk, lo, 0.0)
r.set_row_interleave(r0, G)
frames = bench.frame_params(wl, 70)
for vp, ub in frames:
    r.render_raw(vp, ub, 0.0, None, True)
r.sync()
h = r.frame_history(60)
print(f"G={G} rank={r0} GSR_SH_BULK_MIN={os.environ.get('GSR_SH_BULK_MIN','default')}", {n: round(float(np.mean([x.stage_ms[i] for x in h])), 4) for i, n in enumerate(["proj", "sort", "ranges", "render", "total"])})
