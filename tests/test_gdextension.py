"""The GDExtension entry (godot/gsr_gdextension.c -> libgsr_godot.so, manifest addons/gsr/gsr.gdextension) loaded the way the engine
loads it, by a fake host (tests/gdext_fake_host.c) that implements the engine's side of the interface: entry symbol, get_proc_address
table, initialization levels, class and method registration, ptr-calls and variant calls.  No Godot exists in this image; the
interface declarations are hand-written (godot/gdextension_min.h)."""
import json
import os
import subprocess

import numpy as np
import pytest

from godotgaussiansplatting_b200 import build as gsr_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SRC = os.path.join(ROOT, "tests", "gdext_fake_host.c")
HOST_BIN = os.path.join(ROOT, "tests", "gdext_fake_host")


def build_host():
    shim = gsr_build.build_godot_shim()
    if not os.path.exists(HOST_BIN) or os.path.getmtime(HOST_BIN) < max(os.path.getmtime(HOST_SRC), os.path.getmtime(shim)):
        subprocess.run(["/usr/bin/gcc", "-std=gnu11", "-O1", "-Wall", HOST_SRC, "-o", HOST_BIN, "-ldl"], check=True)
    return shim


def run_host(*args):
    shim = build_host()
    res = subprocess.run([HOST_BIN, shim, *args], capture_output=True, text=True, timeout=300)
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    return res, lines


def test_manifest_names_the_entry_symbol():
    text = open(os.path.join(ROOT, "addons", "gsr", "gsr.gdextension")).read()
    assert 'entry_symbol = "gsr_gdext_init"' in text and "libgsr_godot.so" in text


def test_shim_registers_the_class_and_its_methods():
    res, lines = run_host("register")
    assert res.returncode == 0, res.stderr
    cls = lines[0]
    assert cls == {"class": "GsrRasterizer", "parent": "RefCounted", "min_level": 2, "exposed": 1, "has_create": 1, "has_free": 1}
    methods = {l["method"]: l for l in lines if "method" in l}
    assert set(methods) == {"create", "destroy", "resize", "upload_ply_raw", "upload_splats", "render", "pick", "stats", "framebuffer_ptr"}
    assert methods["render"]["types"] == [29, 29, 3, 29] and methods["upload_ply_raw"]["types"] == [29, 2, 2, 2, 3]   # PackedByteArray / int / float
    assert all(m["ptrcall"] == 1 and m["call"] == 1 for m in methods.values())
    last = lines[-1]
    from godotgaussiansplatting_b200 import _lib
    if _lib.lib().gsr_device_count() == 0:   # no GPU here: create fails loudly (no CPU fallback) and tells the engine through print_error
        assert last["create_rc"] == _lib.GSR_ERR_CUDA and last["engine_errors"] >= 1
    else:
        assert last["create_rc"] == 0


@pytest.mark.gpu
def test_shim_renders_the_oracle_frame_through_ptrcalls(tmp_path):
    from oracle import oracle as orc
    from tests.scenes import make_scene
    n, w, h = 20000, 400, 300
    splat60, vp, ub = make_scene(n, 31, w, h)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([n, w, h], dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(vp, dtype=np.float32).tobytes()); f.write(ub); f.write(np.ascontiguousarray(splat60, dtype=np.float32).tobytes())
    res, lines = run_host("render", str(inp), str(outp))
    assert res.returncode == 0 and lines[-1]["render_rc"] == 0 and lines[-1]["engine_errors"] == 0, res.stdout + res.stderr
    ref = orc.frame(splat60, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
    raw = np.fromfile(outp, dtype=np.uint8)
    rgba = raw[: w * h * 16].view(np.float32).reshape(h, w, 4)
    np.testing.assert_array_equal(rgba.view(np.uint32), ref.rgba.view(np.uint32))
    assert lines[-1]["duplicates"] == ref.duplicates and lines[-1]["fb_ptr_nonzero"] == 1
    gx = (w + 15) // 16
    tile = ((h // 2) // 16) * gx + (w // 2) // 16
    _, _, want = orc.render(ref.records, ref.values, ref.bounds, w, h, target_tile=tile, pick=np.zeros(4, np.float32))
    np.testing.assert_array_equal(raw[w * h * 16: w * h * 16 + 16].view(np.float32).view(np.uint32), want.view(np.uint32))
