"""Host-side camera maths that the reference takes from the Godot engine.

* `perspective` restates Godot 4.3 `Projection::set_perspective` (engine source not vendored in the
  reference; stated from knowledge of Godot 4.x) -- what `Camera3D.get_camera_projection()` returns
  for the defaults fov 75, near 0.05, far 4000, keep_aspect = KEEP_HEIGHT (main.tscn:36-38).
* `pack_camera_push_constants` follows util/gaussian_splatting_rasterizer.gd:175-195 literally.
* `Camera3D` mirrors the handful of members the rasterizer touches (global_position,
  get_camera_transform, get_camera_projection); `reset()` follows util/camera.gd:151-153.
* `orbit_camera` generates the 1-degree-per-frame orbit of BASELINE.json config c3.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def perspective(fovy_degrees: float, aspect: float, z_near: float, z_far: float) -> np.ndarray:
    """Godot Projection columns x,y,z,w flattened (16 float32, column-major)."""
    radians = F(math.radians(F(fovy_degrees) / F(2.0)))
    delta_z = F(z_far) - F(z_near)
    sine = F(math.sin(radians))
    cotangent = F(F(math.cos(radians)) / sine)
    m = np.zeros((4, 4), dtype=np.float32)  # m[c][r]
    m[0][0] = cotangent / F(aspect)
    m[1][1] = cotangent
    m[2][2] = -(F(z_far) + F(z_near)) / delta_z
    m[2][3] = F(-1.0)
    m[3][2] = F(-2.0) * F(z_near) * F(z_far) / delta_z
    m[3][3] = F(0.0)
    return m.reshape(16)


def transform_to_projection(basis_cols: np.ndarray, origin: np.ndarray) -> np.ndarray:
    """Godot Projection(Transform3D): columns (x,0),(y,0),(z,0),(origin,1)."""
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, :3], m[1, :3], m[2, :3] = basis_cols[0], basis_cols[1], basis_cols[2]
    m[3, :3] = origin
    m[3, 3] = 1.0
    return m.reshape(16)


def pack_camera_push_constants(view16: np.ndarray, proj16: np.ndarray) -> np.ndarray:
    """util/gaussian_splatting_rasterizer.gd:181-193 -> 32 float32 (view_matrix, projection_matrix)."""
    v = np.asarray(view16, dtype=np.float32).reshape(4, 4)
    p = np.asarray(proj16, dtype=np.float32).reshape(4, 4)
    x, y, z, w = v[0], v[1], v[2], v[3]

    def dot4(a, b):
        return ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]

    out = np.array([
        -x[0], y[0], -z[0], 0.0,
        -x[1], y[1], -z[1], 0.0,
        x[2], -y[2], z[2], 0.0,
        -dot4(w, x), -dot4(w, -y), -dot4(w, z), 1.0,
        p[0][0], p[0][1], p[0][2], 0.0,
        p[1][0], p[1][1], p[1][2], 0.0,
        p[2][0], p[2][1], p[2][2], -1.0,
        p[3][0], p[3][1], p[3][2], 0.0], dtype=np.float32)
    return out


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


class Camera3D:
    """Minimal stand-in for Godot's Camera3D as used by the rasterizer (fov/near/far defaults of the engine)."""

    def __init__(self, fov: float = 75.0, near: float = 0.05, far: float = 4000.0):
        self.fov, self.near, self.far = float(fov), float(near), float(far)
        self.basis = np.eye(3, dtype=np.float32)  # rows of this array are the basis COLUMNS x, y, z
        self.global_position = np.zeros(3, dtype=np.float32)
        self.aspect = 16.0 / 9.0

    def reset(self) -> None:  # util/camera.gd:151-153: position = 0, rotation = UP * -PI
        self.global_position = np.zeros(3, dtype=np.float32)
        self.set_yaw(-math.pi)

    def set_yaw(self, yaw: float) -> None:
        c, s = F(math.cos(F(yaw))), F(math.sin(F(yaw)))
        # Basis rows [[c,0,s],[0,1,0],[-s,0,c]] -> columns x=(c,0,-s), y=(0,1,0), z=(s,0,c)
        self.basis = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=np.float32)

    def look_at_from_position(self, position, target, up=(0.0, 1.0, 0.0)) -> None:
        """Node3D.look_at_from_position: -Z looks at target."""
        position = np.asarray(position, dtype=np.float64)
        vz = _normalize(position - np.asarray(target, dtype=np.float64))
        vx = _normalize(np.cross(np.asarray(up, dtype=np.float64), vz))
        vy = np.cross(vz, vx)
        self.basis = np.array([vx, vy, vz], dtype=np.float32)
        self.global_position = position.astype(np.float32)

    def get_camera_transform(self) -> np.ndarray:
        return transform_to_projection(self.basis, self.global_position)

    def get_camera_projection(self) -> np.ndarray:
        return perspective(self.fov, self.aspect, self.near, self.far)


def orbit_camera(frame: int, center=(0.0, 0.0, 2.5), radius: float = 2.5, pitch_deg: float = 10.0,
                 fov: float = 75.0, aspect: float = 16.0 / 9.0, step_deg: float = 1.0) -> Camera3D:
    """Camera for frame `frame` of the c3 orbit sweep: `step_deg` yaw steps about `center` at fixed radius
    and pitch, looking at the centroid.  Frame 0 sits at the c2 camera side (the world origin for the
    defaults), so frame 0 of the orbit sees the same view direction as the default camera."""
    yaw = math.radians(frame * step_deg)
    pitch = math.radians(pitch_deg)
    c = np.asarray(center, dtype=np.float64)
    # Godot world space shows the scene mirrored in x,y (rasterizer.gd:181-193); the centroid's x,y are
    # mirrored accordingly.  With the default centre (0,0,2.5) the mirror is a no-op.
    cw = np.array([-c[0], -c[1], c[2]])
    pos = cw + radius * np.array([math.sin(yaw) * math.cos(pitch), math.sin(pitch), -math.cos(yaw) * math.cos(pitch)])
    cam = Camera3D(fov=fov)
    cam.aspect = aspect
    cam.look_at_from_position(pos, cw)
    return cam


def default_camera(aspect: float = 16.0 / 9.0, fov: float = 75.0) -> Camera3D:
    cam = Camera3D(fov=fov)
    cam.aspect = aspect
    cam.reset()
    return cam
