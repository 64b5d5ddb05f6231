"""Scheduling sweep of the compositor's persistent grid on c3 frames (and optionally one emulated rank of a G-GPU group):
resident CTAs/SM x longest-chain-first ticket order -- gsr_debug_compositor_config, results never depend on it.
One scene upload, many configurations:   python ubench/compositor_sweep.py [rows_mod]"""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from godotgaussiansplatting_b200 import _lib
from godotgaussiansplatting_b200.camera import default_camera
from godotgaussiansplatting_b200.ply_file import PlyFile
from godotgaussiansplatting_b200.rasterizer import GaussianSplattingRasterizer, RenderTexture

rows_mod = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = dict(bench.WORKLOADS[os.environ.get("SWEEP_WORKLOAD", "c3")])
stub = PlyFile(); stub.size = wl["n"]
r = GaussianSplattingRasterizer(stub, (wl["w"], wl["h"]), RenderTexture(), default_camera())
r.init_gpu(load=False)
for lo, blk in bench.raw_chunks(wl):
    r.upload_ply_raw(blk, lo, 0.0)
if rows_mod > 1:
    r.set_row_interleave(3 % rows_mod, rows_mod)   # one rank's share of the tile rows (cyclic), like a group member
frames = bench.frame_params(wl, 40)
L = _lib.lib()
res = []
for ctas, lpt, sparse in [(2, 1, 3), (2, 1, 0), (2, 0, 0), (1, 1, 0), (1, 0, 0), (3, 1, 0), (0, 0, 0)]:
    _lib.check(L.gsr_debug_compositor_config(r._ctx, ctas, lpt, sparse), "config")
    for vp_, ub_ in frames:
        r.render_raw(vp_, ub_, 0.0, None, asynchronous=True)
    r.sync()
    hist = r.frame_history()[-30:]
    ms = np.array([[f.stage_ms[i] for i in range(5)] for f in hist])
    res.append((ms[:, 3].mean(), ctas, lpt, sparse, ms[:, 4].mean()))
    print(f"[rows/{rows_mod}] ctas/SM {ctas or 'max'} longest-first {lpt} sparse-rule {sparse}: compositor {ms[:, 3].mean():.3f} ms  frame {ms[:, 4].mean():.3f} ms", flush=True)
best = min(res)
print(f"[rows/{rows_mod}] BEST compositor {best[0]:.3f} ms: ctas {best[1]} longest-first {best[2]} sparse-rule {best[3]}")
