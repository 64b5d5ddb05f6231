"""Round-2 experiment harness for the opt-in compositor variants: accuracy against the oracle on a test scene, and the
compositor's stage time on c3 frames.  The knobs are read once per process, so run one process per variant:

    for v in "" GSR_COMP_V2=5 GSR_COMP_V2=6 GSR_COMP_V2=8 GSR_COMP_P4=1 GSR_COMP_HWEXP=1; do env $v python ubench/compositor_variant.py; done

  (none)           composite_kernel<false>: the shipped, GPU-verified kernel (106 registers, 4 CTAs/SM)
  GSR_COMP_V2=5|6|8  cp.async staging into a 24 KB double buffer, one barrier per chunk; 82 / 70 / 62 registers -> 5 / 6 / 8 CTAs/SM
  GSR_COMP_P4=1    v2 staging + four pixels per thread, 64 threads per tile (120 registers, 8 CTAs/SM)
  GSR_COMP_HWEXP=1 exp() on the SFU (MUFU.EX2): not bit-reproducible, pixels within 1e-4 except at stop-rule flips
V2 and P4 are bit-identical by construction; their logic is checked on the CPU by tests/test_kernel_emu.py.  None of the
variants has run on a GPU yet (written after the round-1 GPU budget was spent).  To validate one fully:
    GSR_COMP_V2=6 python -m pytest tests -m gpu -q
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from godotgaussiansplatting_b200 import _lib  # noqa: E402
from godotgaussiansplatting_b200.camera import default_camera  # noqa: E402
from godotgaussiansplatting_b200.ply_file import PlyFile  # noqa: E402
from godotgaussiansplatting_b200.rasterizer import GaussianSplattingRasterizer, RenderTexture  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.gsr_direct import Ctx  # noqa: E402
from tests.scenes import make_scene  # noqa: E402

mode = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("GSR_COMP_")) or "shipped"

# ---- accuracy: 200k splats, 1280x720, against the oracle (the gsr spec) ----
n, w, h = 200_000, 1280, 720
splat60, vp, ub = make_scene(n, 4, w, h, scale_boost=0.5)
ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
with Ctx(n, w, h) as c:
    c.upload(splat60)
    img = c.render(vp, ub)
    t = c.taps()
d = np.abs(img - ref.rgba).max(axis=2)
print(f"[{mode}] accuracy: bit-identical {np.array_equal(img.view(np.uint32), ref.rgba.view(np.uint32))}, keys equal {np.array_equal(t['keys'], ref.keys)}, ranges equal {np.array_equal(t['bounds'], ref.bounds)}, "
      f"max |rgba - oracle| {d.max():.3g}, pixels > 1e-4: {(d > 1e-4).sum()} of {d.size} ({(d > 1e-4).mean():.2e}), "
      f"> 1e-5: {(d > 1e-5).mean():.2e}, staged C {t['stats'].staged} vs {ref.staged}")

# ---- speed: c3, 40 orbit frames, per-stage CUDA events from the library ----
wl = dict(bench.WORKLOADS["c3"])
stub = PlyFile(); stub.size = wl["n"]
r = GaussianSplattingRasterizer(stub, (wl["w"], wl["h"]), RenderTexture(), default_camera())
r.init_gpu(load=False)
for lo, s60 in bench.scene_chunks(wl):
    r.upload_splats(s60, lo)
frames = bench.frame_params(wl, 50)
for vp_, ub_ in frames:
    r.render_raw(vp_, ub_, 0.0, None, asynchronous=True)
r.sync()
hist = r.frame_history()[-40:]
ms = np.array([[f.stage_ms[i] for i in range(5)] for f in hist])
print(f"[{mode}] c3 stage ms (mean of 40): projection {ms[:, 0].mean():.3f} sort {ms[:, 1].mean():.3f} ranges {ms[:, 2].mean():.3f} "
      f"compositor {ms[:, 3].mean():.3f} total {ms[:, 4].mean():.3f}")
