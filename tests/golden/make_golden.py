"""Mints tests/golden/demo_subset.npz from the reference's only fixture, resources/demo.ply (run in the build
container where /root/reference is mounted):  python tests/golden/make_golden.py

Contents: every 33rd splat of demo.ply (8216 vertices, the raw 62 floats each), the default camera of
util/camera.gd:151-153 at 320x240, and the oracle's outputs for that frame (sorted keys/values, tile bounds, RGBA).
The reference itself cannot run here (Godot 4.3 + Vulkan), so these vectors are minted by the oracle, not by the
reference: they pin the oracle build and give the GPU parity tests a real-data case on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from godotgaussiansplatting_b200 import camera as cam  # noqa: E402
from godotgaussiansplatting_b200.ply_file import PlyFile, swizzle_splats  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.scenes import uniforms_bytes  # noqa: E402

ply = PlyFile("/root/reference/resources/demo.ply")
sub = np.ascontiguousarray(ply.table[::33])
W, H = 320, 240
c = cam.default_camera(aspect=W / H)
vp = cam.pack_camera_push_constants(c.get_camera_transform(), c.get_camera_projection())
ub = uniforms_bytes(c.global_position, 1.0, W, H, 10.0)
s = swizzle_splats(sub, 0.0)
fr = orc.frame(s, vp, orc.uniforms_from_bytes(np.frombuffer(ub, dtype=np.uint8)))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demo_subset.npz")
np.savez_compressed(out, ply62=sub, splat60=s, vp=vp, uniforms=np.frombuffer(ub, dtype=np.uint8), keys=fr.keys, values=fr.values,
                    bounds=fr.bounds, rgba=fr.rgba, duplicates=fr.duplicates, visible=fr.visible, width=W, height=H)
print(out, os.path.getsize(out), "bytes; N", sub.shape[0], "V", fr.visible, "M", fr.duplicates, "C", fr.staged)
