"""Worker of tests/test_gpu_multi.py (torchrun, one process per GPU): tile-row bands with (a) the fused peer-memory
compositor store and (b) the NCCL gather; rank 0 checks both full frames bit for bit against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from godotgaussiansplatting_b200 import _lib, sharding  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.gsr_direct import Ctx  # noqa: E402
from tests.scenes import make_scene  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    n, w, h = 30000, 800, 600
    frames = [make_scene(n, 41, w, h, frame=f) for f in (0, 45, 90, 135)]
    splat60 = frames[0][0]
    tiles_y = (h + 15) // 16
    band = sharding.band_partition(tiles_y, world)[rank]
    L = _lib.lib()
    with Ctx(n, w, h, device=local) as c:
        _lib.check(L.gsr_set_stream(c.h, C.c_void_p(stream.cuda_stream)), "stream")
        c.upload(splat60)
        c.set_row_interleave(rank, world)   # balanced cyclic rows + fast sharded mode
        # ---- (a) peer-memory mode ----
        handles = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(L.gsr_peer_export_framebuffers(c.h, buf), "export")
            handles.copy_(torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8))
        dist.broadcast(handles, src=0)
        if rank != 0:
            buf = (C.c_ubyte * 128).from_buffer_copy(handles.cpu().numpy().tobytes())
            _lib.check(L.gsr_peer_import_framebuffers(c.h, buf), "import")
        class _Word:  # the library's int32 sync word as a torch tensor (all-reduced in place)
            __cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (int(L.gsr_band_sync_word(c.h)), False), "version": 2}
        word = torch.as_tensor(_Word(), device="cuda")
        hosts = [torch.zeros((h, w, 3), dtype=torch.float32).pin_memory() for _ in frames]
        for k, (_, vp, ub) in enumerate(frames):
            vpc = np.ascontiguousarray(vp, dtype=np.float32)
            _lib.check(L.gsr_render_async(c.h, vpc.ctypes.data_as(C.POINTER(C.c_float)), ub, 0.0, None), "render")
            if rank == 0:
                _lib.check(L.gsr_stream_join(c.h), "join")
            dist.all_reduce(word, op=dist.ReduceOp.MAX)   # completion sync + frame-global last occupied tile
            _lib.check(L.gsr_band_fixup(c.h), "fixup")
            if rank == 0:
                _lib.check(L.gsr_readback_async(c.h, C.c_void_p(hosts[k].data_ptr()), 1), "readback")
        _lib.check(L.gsr_sync(c.h), "sync")
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            for k, (_, vp, ub) in enumerate(frames):
                ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
                got = hosts[k].numpy()
                assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref.rgba[..., :3]).view(np.uint32)), f"peer frame {k} differs"
            print("PEER_MODE_OK", flush=True)
    dist.barrier()
    # ---- (b) NCCL gather mode (fresh context: external torch frame) ----
    with Ctx(n, w, h, device=local) as c:
        _lib.check(L.gsr_set_stream(c.h, C.c_void_p(stream.cuda_stream)), "stream")
        c.upload(splat60)
        c.set_band(*band)
        fb = torch.zeros((sharding.padded_height(h, world), w, 4), dtype=torch.float32, device="cuda")
        _lib.check(L.gsr_set_framebuffer_external(c.h, C.c_void_p(fb.data_ptr())), "ext")
        _, vp, ub = frames[1]
        vpc = np.ascontiguousarray(vp, dtype=np.float32)
        _lib.check(L.gsr_render_async(c.h, vpc.ctypes.data_as(C.POINTER(C.c_float)), ub, 0.0, None), "render")
        sharding.gather_bands(fb, rank, world, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
            got = fb[:h].cpu().numpy()
            assert np.array_equal(got.view(np.uint32), ref.rgba.view(np.uint32)), "gathered frame differs"
            print("NCCL_GATHER_OK", flush=True)
    dist.barrier()
    # ---- (c) shard group: NCCL-free frame path (pairs, records and rows over NVLink peer memory, device-side flags) ----
    gx = (w + 15) // 16
    with Ctx(n, w, h, device=local) as c:
        _lib.check(L.gsr_set_stream(c.h, C.c_void_p(stream.cuda_stream)), "stream")
        c.upload(splat60)
        mine = torch.frombuffer(bytearray(c.group_export()), dtype=torch.uint8).cuda()
        blobs = torch.zeros(world * _lib.GSR_GROUP_BLOB_BYTES, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(blobs, mine)          # set-up only: NCCL is not on the frame path
        c.group_attach(rank, world, blobs.cpu().numpy().tobytes())
        dist.barrier()
        seq = [frames[k % len(frames)] for k in range(7)]   # > 2 frames in flight: exercises slot release + flag parity
        hosts = [torch.zeros((h, w, 4), dtype=torch.float32).pin_memory() for _ in seq]
        for k, (_, vp, ub) in enumerate(seq):
            c.render_async(vp, ub)
            if rank == 0:
                c.readback_async(hosts[k].data_ptr())
        c.sync()
        t = c.taps()
        _, vp, ub = seq[-1]
        ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
        sel = ((ref.keys >> 16) // gx) % world == rank
        assert np.array_equal(t["keys"], ref.keys[sel]) and np.array_equal(t["values"], ref.values[sel]), f"rank {rank}: pairs differ"
        assert t["stats"].last_tile == ref.last_tile
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            for k, (_, vp, ub) in enumerate(seq):
                ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
                assert np.array_equal(hosts[k].numpy().view(np.uint32), ref.rgba.view(np.uint32)), f"group frame {k} differs"
            print("GROUP_MODE_OK", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
