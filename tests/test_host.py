"""Host-side logic without a GPU: PLY parsing, the GDScript-mirror's bookkeeping, synthetic scenes, sharding."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from godotgaussiansplatting_b200 import camera as cam
from godotgaussiansplatting_b200 import sharding
from godotgaussiansplatting_b200.ply_file import PlyFile, default_properties, load_gaussian_splats, swizzle_splats
from godotgaussiansplatting_b200.rasterizer import GaussianSplattingRasterizer, RenderTexture
from godotgaussiansplatting_b200.synthetic import radix_keys, synthetic_ply, synthetic_ply_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_ply(path, table, big_endian=False):
    names = default_properties(table.shape[1])
    with open(path, "wb") as f:
        f.write(b"ply\n")
        f.write(b"format binary_big_endian 1.0\n" if big_endian else b"format binary_little_endian 1.0\n")
        f.write(f"element vertex {table.shape[0]}\n".encode())
        for n in names:
            f.write(f"property float {n}\n".encode())
        f.write(b"end_header\n")
        f.write(table.astype(">f4" if big_endian else "<f4").tobytes())


@pytest.mark.parametrize("big", [False, True])
def test_plyfile_parse_roundtrip(tmp_path, big):
    table = synthetic_ply_table(321, 5)
    p = tmp_path / "t.ply"
    write_ply(p, table, big_endian=big)
    ply = PlyFile(str(p))
    assert ply.size == 321 and ply.properties == default_properties(62)
    np.testing.assert_array_equal(ply.table, table)
    assert ply.get_vertex(7)["opacity"] == float(table[7, 54])


def test_plyfile_truncated_body_raises(tmp_path):
    table = synthetic_ply_table(10, 5)
    p = tmp_path / "t.ply"
    write_ply(p, table)
    data = p.read_bytes()
    p.write_bytes(data[:-40])
    with pytest.raises(ValueError):
        PlyFile(str(p))


def test_load_gaussian_splats_chunks_and_terminates():
    ply = synthetic_ply(2500, 3)
    got, loaded, done = {}, [0], []
    load_gaussian_splats(ply, 1000, lambda first, blk: got.__setitem__(first, blk), [False], loaded, lambda: done.append(1), clock=lambda: 2.5)
    assert sorted(got) == [0, 1000, 2000] and loaded[0] == 2500 and done == [1]
    full = np.concatenate([got[k] for k in sorted(got)])
    np.testing.assert_array_equal(full, swizzle_splats(ply.table, 2.5))
    assert np.all(full[:, 3] == np.float32(2.5))  # creation time stamp (ply_file.gd:40,47)
    stop = [False]
    seen = []

    def upload(first, blk):
        seen.append(first)
        stop[0] = True

    load_gaussian_splats(ply, 1000, upload, stop, [0], None)
    assert seen == [0]  # cooperative cancel (rasterizer.gd:117)


def test_swizzle_layout():
    t = np.zeros((1, 62), dtype=np.float32)
    t[0, 0:3] = (1, 2, 3)
    t[0, 6:9] = (10, 11, 12)                      # f_dc
    t[0, 9:54] = np.arange(45) + 100              # f_rest: R 0..14, G 15..29, B 30..44
    t[0, 54] = 0.0                                # sigmoid -> 0.5
    t[0, 55:58] = np.log([1.0, 2.0, 3.0])
    t[0, 58:62] = (1, 0, 0, 0)                    # identity quaternion (w first in the file)
    s = swizzle_splats(t, 9.0)[0]
    assert s[:4].tolist() == [1, 2, 3, 9]
    np.testing.assert_allclose(s[4:10], [1, 0, 0, 4, 0, 9], rtol=1e-6)  # R S^2 R^T upper triangle
    assert s[10] == 0.5 and s[11] == 0
    assert s[12:15].tolist() == [10, 11, 12]
    assert s[15:18].tolist() == [100, 115, 130] and s[57:60].tolist() == [114, 129, 144]


def test_rasterizer_mirror_bookkeeping_without_gpu():
    ply = synthetic_ply(100, 1)
    c = cam.default_camera(aspect=16 / 9)
    r = GaussianSplattingRasterizer(ply, (1920, 1080), RenderTexture(), c)
    assert r.texture_size == (1920, 1080) and r.tile_dims == (120, 68)
    r.render_scale[0] = 0.5
    r.texture_size = (1920, 1080)                  # rasterizer.gd:28: scaled by render_scale, min 1
    assert r.texture_size == (960, 540) and r.tile_dims == (60, 34)
    r.render_scale[0] = 1e-9
    r.texture_size = (100, 100)
    assert r.texture_size == (1, 1) and r.tile_dims == (1, 1)
    assert r.update_camera_matrices() is True and r.update_camera_matrices() is False
    c.global_position = np.array([0.5, 0.25, -1.0], dtype=np.float32)
    assert r.update_camera_matrices() is True
    u = r.uniforms_bytes(time=3.5)
    f = np.frombuffer(u, dtype=np.float32)
    i = np.frombuffer(u, dtype=np.int32)
    assert len(u) == 32 and f[:4].tolist() == [-0.5, -0.25, -1.0, 1.0] and i[4:6].tolist() == [1, 1] and f[6] == 3.5
    assert r.camera_push_constants.shape == (32,) and r.camera_push_constants[31] == 0.0 and r.camera_push_constants[27] == -1.0


def test_synthetic_scene_is_deterministic_and_in_front_of_the_camera():
    a, b = synthetic_ply_table(3000, 9), synthetic_ply_table(3000, 9)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, synthetic_ply_table(3000, 10))
    r = np.linalg.norm(a[:, :3] - np.array([0, 0, 2.5], dtype=np.float32), axis=1)
    assert r.max() <= 1.0 + 1e-5
    np.testing.assert_allclose(np.linalg.norm(a[:, 58:62], axis=1), 1.0, rtol=1e-5)
    k = radix_keys(10000, 1)
    assert (k >> 16).max() < 8160 and 52000 <= (k & 0xFFFF).min() and (k & 0xFFFF).max() < 61500


def test_band_partition():
    assert sharding.band_partition(68, 1) == [(0, 68)]
    assert sharding.band_partition(68, 8) == [(0, 9), (9, 18), (18, 27), (27, 36), (36, 45), (45, 54), (54, 63), (63, 68)]
    assert sharding.band_partition(3, 8)[3:] == [(3, 3)] * 5           # more ranks than rows: empty bands
    assert sharding.padded_height(1080, 8) == 9 * 16 * 8 and sharding.slab_rows(1080, 8) == 144
    assert sharding.padded_height(2160, 4) >= 2160


def test_two_rank_gloo_band_gather_reproduces_the_full_frame():
    """world_size=2 on CPU (gloo): each rank renders its tile-row band with the ORACLE, the bands are gathered with
    sharding.gather_bands, and rank 0 checks the result equals the oracle's full frame bit for bit."""
    script = os.path.join(ROOT, "tests", "gloo_band_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", script], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "BAND_GATHER_OK" in res.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c2", "--splats", "30000",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ref = cb["reference_shaders"]       # the reference's own shaders on a bounded sample (oracle/_ref), when available
    assert "unavailable" in ref or (ref["kind"] == "reference" and ref["keys_and_ranges_identical_to_port"])


def test_bench_config_is_a_function_of_the_command_line_only():
    """Both bench arms must print the SAME `config` for the same workload and GPU count (the driver compares them), and the default
    multi-GPU path is the shard group with the rows-local read-back."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    old = sys.argv
    try:
        cfgs = {}
        for impl in ("gsr", "reference"):
            for n in (1, 2, 8):
                sys.argv = ["bench.py", "--impl", impl, "--gpus", str(n)]
                args = bench.parse_args()
                assert args.mgpu == "group" and args.present == "rows" and args.overlap == -1
                cfgs[impl, n] = bench.make_config(args, dict(bench.WORKLOADS[args.workload]))
    finally:
        sys.argv = old
    for n in (1, 2, 8):
        assert cfgs["gsr", n] == cfgs["reference", n]
    assert cfgs["gsr", 1]["parallelism"] == "single GPU" and cfgs["gsr", 1]["workload"].startswith("c3")
    assert "scattered to the row owners" in cfgs["gsr", 8]["parallelism"] and "x8" in cfgs["gsr", 8]["parallelism"]


def test_free_look_camera_orbits_like_util_camera_gd():
    """Scope row f4: FreeLookCamera restates util/camera.gd's orbit (set_focused_position :144-149, ORBIT mouse motion :52-60, the
    t = 1 steady state of _update_movement :127-141): the camera stays on a sphere about the focused position, always looks at it,
    turns by offset.x * cos(pitch) about the target's own y axis per event, and refuses pitches outside [-80, 70] degrees."""
    from godotgaussiansplatting_b200 import camera as cam
    focus = np.array([0.3, -0.2, 2.5])
    c = cam.FreeLookCamera()
    c.set_focused_position(focus)
    np.testing.assert_allclose(np.linalg.norm(c.global_position - focus), 2.0, atol=1e-6)        # :147 two units along the view axis
    for k in range(200):
        c.orbit_mouse_motion(7.0, -1.5 if k < 40 else 0.3)
        p = c.global_position.astype(np.float64)
        np.testing.assert_allclose(np.linalg.norm(p - focus), 2.0, atol=2e-5)
        view = -c.basis[2].astype(np.float64)
        np.testing.assert_allclose(view, (focus - p) / np.linalg.norm(focus - p), atol=2e-5)       # look_at_from_position(rotated, orbit_position)
        assert -80.5 <= c._target_pitch_deg() <= 70.5
    c2 = cam.FreeLookCamera()
    c2.set_focused_position(focus)
    for _ in range(400):
        c2.orbit_mouse_motion(0.0, -5.0)    # keep dragging up: the pitch guard of :55 stops the rotation, the camera never flips
    assert 69.0 <= c2._target_pitch_deg() <= 70.5 or -80.5 <= c2._target_pitch_deg() <= -79.0
    sweep = cam.reference_orbit_sweep(360, focus=(0.0, 0.0, 2.5))
    pos = np.array([p for _, _, p in sweep], dtype=np.float64)
    ang = np.unwrap(np.arctan2(pos[:, 0], -(pos[:, 2] - 2.5)))
    assert 355.0 <= np.degrees(ang[-1] - ang[0]) * 360.0 / 359.0 <= 370.0                            # one turn of the orbit
    vp = cam.pack_camera_push_constants(sweep[5][0], sweep[5][1])
    assert vp.shape == (32,) and np.isfinite(vp).all()
