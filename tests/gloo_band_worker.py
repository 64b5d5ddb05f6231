"""Worker of tests/test_host.py::test_two_rank_gloo_band_gather_reproduces_the_full_frame (launched by torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from godotgaussiansplatting_b200 import sharding  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.scenes import make_scene  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, w, h = 6000, 320, 200
    splat60, vp, ub = make_scene(n, 31, w, h)
    u = orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8))
    tiles_y = (h + 15) // 16
    band = sharding.band_partition(tiles_y, world)[rank]
    fr = orc.frame(splat60, vp, u, band=band)
    frame = torch.zeros((sharding.padded_height(h, world), w, 4), dtype=torch.float32)
    rows = sharding.slab_rows(h, world)
    y0, y1 = band[0] * 16, min(band[1] * 16, h)
    frame[y0:y1] = torch.from_numpy(fr.rgba[y0:y1])
    assert y1 - y0 <= rows
    sharding.gather_bands(frame, rank, world, dst=0)
    # sorted keys of the bands concatenate to the full sorted list
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([fr.keys.size], dtype=torch.int64))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros(mx, dtype=torch.int64)
    pad[:fr.keys.size] = torch.from_numpy(fr.keys.astype(np.int64))
    allk = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allk, pad)
    if rank == 0:
        full = orc.frame(splat60, vp, u)
        got = frame[:h].numpy()
        assert np.array_equal(got.view(np.uint32), full.rgba.view(np.uint32)), "gathered bands differ from the full frame"
        cat = np.concatenate([allk[r][:int(sizes[r].item())].numpy() for r in range(world)]).astype(np.uint32)
        assert np.array_equal(cat, full.keys), "band keys do not concatenate to the full sorted keys"
        print("BAND_GATHER_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
