"""Worker of tests/test_gpu_group.py: a shard group of G contexts living in ONE process on ONE GPU (each on its own stream) --
the same device-side protocol as one context per GPU (pairs and records scattered with "peer" stores, flag words, rows composited
into rank 0's frames), exercised without a multi-GPU box.  Rank 0's frames must equal the oracle's bit for bit and every rank's
sorted pairs must be exactly the single-GPU pairs of the tile rows it owns.
CUDA_DEVICE_MAX_CONNECTIONS is raised by the caller: 2 streams per context must not share a hardware queue, or a spinning
wait kernel of one context could sit in front of the kernel it waits for (only an issue when one GPU hosts the whole group)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from godotgaussiansplatting_b200 import _lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.gsr_direct import Ctx  # noqa: E402
from tests.scenes import make_scene  # noqa: E402


def pinned(shape):
    import torch
    return torch.zeros(shape, dtype=torch.float32).pin_memory()


def main():
    G = int(sys.argv[1])
    n, w, h = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (30000, 640, 360)
    boost = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    overlap = int(sys.argv[6]) if len(sys.argv) > 6 else -1    # gsr_debug_pipeline: front/back overlap of consecutive frames
    gx = (w + 15) // 16
    frames = [make_scene(n, 41, w, h, frame=f, scale_boost=boost) for f in (0, 45, 90, 135, 180, 225)]
    splat60 = frames[0][0]
    ctxs = [Ctx(n, w, h, factor=40 if boost else 10) for _ in range(G)]
    try:
        for c in ctxs:
            c.upload(splat60)
        blobs = b"".join(c.group_export() for c in ctxs)
        for r, c in enumerate(ctxs):
            c.group_attach(r, G, blobs)
            _lib.check(c.L.gsr_debug_pipeline(c.h, overlap), "gsr_debug_pipeline")
        hosts = [pinned((h, w, 4)) for _ in frames]
        for k, (_, vp, ub) in enumerate(frames):
            for c in ctxs:                      # every rank enqueues the frame; nothing blocks on the host
                c.render_async(vp, ub)
            ctxs[0].readback_async(hosts[k].data_ptr())   # two frames in flight: slot reuse is ordered by the released flag
        for c in ctxs:
            c.sync()
        _, vp, ub = frames[-1]
        ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), cap=(40 if boost else 10) * n)
        rows = (ref.keys >> 16) // gx
        for r, c in enumerate(ctxs):            # per-rank sorted pairs of the LAST frame = the owned rows of the full frame
            t = c.taps()
            sel = rows % G == r
            assert np.array_equal(t["keys"], ref.keys[sel]), f"rank {r}: keys differ"
            assert np.array_equal(t["values"], ref.values[sel]), f"rank {r}: values differ"
            assert t["stats"].last_tile == ref.last_tile, f"rank {r}: frame-global last tile {t['stats'].last_tile} != {ref.last_tile}"
        for k, (_, vp, ub) in enumerate(frames):
            ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), cap=(40 if boost else 10) * n)
            got = hosts[k].numpy()
            assert np.array_equal(got.view(np.uint32), ref.rgba.view(np.uint32)), f"frame {k} differs from the oracle (max abs {np.abs(got - ref.rgba).max()})"
        # ---- rows-local presentation: every rank reads its own tile rows back into ONE host frame (a PCIe link per rank) ----
        for c in ctxs:
            c.group_set_present(True)
        hosts2 = [pinned((h, w, 4)) for _ in range(4)]
        for k in range(4):
            _, vp, ub = frames[k]
            for c in ctxs:
                c.render_async(vp, ub)
            for c in ctxs:
                c.readback_rows_async(hosts2[k].data_ptr())
        for c in ctxs:
            c.sync()
        for k in range(4):
            _, vp, ub = frames[k]
            ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)), cap=(40 if boost else 10) * n)
            assert np.array_equal(hosts2[k].numpy().view(np.uint32), ref.rgba.view(np.uint32)), f"rows-local frame {k} differs from the oracle"
        print(f"GROUP_INPROCESS_OK G={G}", flush=True)
    finally:
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
