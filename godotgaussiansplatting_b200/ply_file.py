"""Host-side mirror of the reference's `PlyFile` resource (util/ply_file.gd).

`PlyFile.parse` follows util/ply_file.gd:10-19 (header scan, whole body as one float32 array) and
`load_gaussian_splats` follows util/ply_file.gd:28-77: per splat exp(scale), quaternion -> rotation,
Sigma = R S^2 R^T (upper triangle), sigmoid(opacity) and the SH re-interleave into the 60-float
std430 `Splat` struct (gsplat_projection.glsl:33-40) -- vectorised numpy instead of a GDScript loop,
then chunked uploads through the C-ABI (`gsr_upload_splats_aos`), mirroring the ~1000 chunked
`buffer_update`s of the reference.  In the reference this step also runs on the host CPU.

All float32 operations are written one IEEE op at a time in the order Godot's `Basis`/`Quaternion`
code evaluates them, so the result is bit-identical to the C oracle's restatement (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

STRUCT_SIZE = 60  # floats (util/ply_file.gd:29)
F = np.float32


class PlyFile:
    """util/ply_file.gd:1-26.  Attributes: size, vertices (flat float32), properties (names)."""

    def __init__(self, path: str = ""):
        self.size = 0
        self.vertices = np.zeros(0, dtype=np.float32)
        self.properties: list[str] = []
        if path:
            self.parse(path)

    @classmethod
    def from_array(cls, vertices: np.ndarray, properties: list[str] | None = None) -> "PlyFile":
        """Build a PlyFile from an (n, nprops) float32 array (synthetic scenes)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        self = cls()
        self.size = int(v.shape[0])
        self.properties = list(properties) if properties is not None else default_properties(v.shape[1])
        self.vertices = v.reshape(-1)
        return self

    def parse(self, path: str) -> None:  # util/ply_file.gd:10-19
        big_endian = False
        with open(path, "rb") as f:
            while True:
                raw = f.readline()
                if not raw:
                    raise ValueError("PLY header has no end_header")
                line = raw.decode("ascii", "replace").strip().split(" ")
                if line[0] == "end_header":
                    break
                if line[0] == "format":
                    big_endian = line[1] == "binary_big_endian"
                elif line[0] == "element":
                    self.size = int(line[2])
                elif line[0] == "property":
                    self.properties.append(line[2])
            count = self.size * len(self.properties)
            body = np.fromfile(f, dtype=">f4" if big_endian else "<f4", count=count)
        if body.size != count:
            raise ValueError(f"PLY body truncated: {body.size} of {count} floats")
        self.vertices = body.astype(np.float32, copy=False)

    def get_vertex(self, index: int) -> dict:  # util/ply_file.gd:21-26
        n = len(self.properties)
        return {self.properties[i]: float(self.vertices[n * index + i]) for i in range(n)}

    @property
    def table(self) -> np.ndarray:
        return self.vertices.reshape(self.size, len(self.properties))


def default_properties(nprops: int = 62) -> list[str]:
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return names[:nprops]


def _basis_mul(a, b):
    """Godot Basis*Basis, rows[i][j] = b[0][j]*a[i][0] + b[1][j]*a[i][1] + b[2][j]*a[i][2] (float32)."""
    o = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            o[i][j] = (b[0][j] * a[i][0] + b[1][j] * a[i][1]) + b[2][j] * a[i][2]
    return o


def swizzle_splats(p: np.ndarray, creation_time: float) -> np.ndarray:
    """util/ply_file.gd:41-69 for a block of vertices. p: (m, nprops>=62) float32 -> (m, 60) float32."""
    p = np.asarray(p, dtype=np.float32)
    m = p.shape[0]
    out = np.zeros((m, STRUCT_SIZE), dtype=np.float32)
    out[:, 0:3] = p[:, 0:3]
    out[:, 3] = F(creation_time)
    # exp() is a GDScript float (float64); narrowed when stored in Vector3 (real_t = float32)
    sc = [np.exp(p[:, 55 + k].astype(np.float64)).astype(np.float32) for k in range(3)]
    qx, qy, qz, qw = p[:, 59], p[:, 60], p[:, 61], p[:, 58]  # Quaternion(rot_1, rot_2, rot_3, rot_0)
    d = ((qx * qx + qy * qy) + qz * qz) + qw * qw
    s = F(2.0) / d
    xs, ys, zs = qx * s, qy * s, qz * s
    wx, wy, wz = qw * xs, qw * ys, qw * zs
    xx, xy, xz = qx * xs, qx * ys, qx * zs
    yy, yz, zz = qy * ys, qy * zs, qz * zs
    one = F(1.0)
    B = [[one - (yy + zz), xy - wz, xz + wy], [xy + wz, one - (xx + zz), yz - wx], [xz - wy, yz + wx, one - (xx + yy)]]
    R = [[B[c][r] for c in range(3)] for r in range(3)]  # .transposed()
    zero = np.zeros(m, dtype=np.float32)
    S = [[sc[0], zero, zero], [zero, sc[1], zero], [zero, zero, sc[2]]]
    M = _basis_mul(S, R)
    Mt = [[M[c][r] for c in range(3)] for r in range(3)]
    Cv = _basis_mul(Mt, M)
    out[:, 4], out[:, 5], out[:, 6] = Cv[0][0], Cv[0][1], Cv[0][2]
    out[:, 7], out[:, 8], out[:, 9] = Cv[1][1], Cv[1][2], Cv[2][2]
    with np.errstate(over="ignore"):
        out[:, 10] = (1.0 / (1.0 + np.exp(-p[:, 54].astype(np.float64)))).astype(np.float32)
    out[:, 12:15] = p[:, 6:9]
    rest = p[:, 9:54].reshape(m, 3, 15)  # channel-major in the file
    out[:, 15:60] = rest.transpose(0, 2, 1).reshape(m, 45)  # coefficient-major RGB
    return out


def load_gaussian_splats(point_cloud: PlyFile, stride: int, upload, should_terminate=None, num_points_loaded=None,
                         callback=None, clock=None) -> None:
    """util/ply_file.gd:28-77.  `upload(first_splat, block60)` plays the role of device.buffer_update
    (:71); `clock()` returns seconds (Time.get_ticks_msec()*1e-3, :40) and stamps each chunk."""
    assert stride >= 1, "stride must be >= 1 (the reference requires size >= 1000)"
    table = point_cloud.table
    n = point_cloud.size
    i = 0
    while i * stride < n:
        if should_terminate is not None and should_terminate[0]:
            return
        lo, hi = i * stride, min(n, (i + 1) * stride)
        block = swizzle_splats(table[lo:hi], clock() if clock else 0.0)
        if should_terminate is not None and should_terminate[0]:
            return
        upload(lo, block)
        if num_points_loaded is not None:
            num_points_loaded[0] += hi - lo
        i += 1
    if callback:
        callback()
