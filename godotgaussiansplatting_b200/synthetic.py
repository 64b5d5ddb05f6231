"""Deterministic synthetic splat clouds in the on-disk 3DGS PLY vertex layout.

positions : mixture of 64 anisotropic Gaussian clusters inside a unit-radius ball centred 2.5 units in
            front of the default camera (the orbit centre of config c3);
log-scale : N(-5.5, 0.9) per axis + log((271123/N)^(1/3)) so screen coverage stays roughly constant;
rotation  : normalised N(0,1)^4; opacity logit N(-2, 1.8); f_dc N(-0.4, 0.8); f_rest N(0, 0.1).
The output is the 62-float vertex table a `.ply` of that scene would hold, so it runs through the same
ingest path (`ply_file.swizzle_splats`) as `resources/demo.ply`.
"""
from __future__ import annotations

import numpy as np

from .ply_file import PlyFile, default_properties

DEMO_SPLATS = 271123
CENTER = (0.0, 0.0, 2.5)


def synthetic_ply_table(n: int, seed: int, sh_degree: int = 3, clusters: int = 64, ball_radius: float = 1.0,
                        chunk: int = 1 << 20) -> np.ndarray:
    out = np.empty((n, 62), dtype=np.float32)
    for lo, blk in synthetic_ply_chunks(n, seed, sh_degree, clusters, ball_radius, chunk):
        out[lo:lo + blk.shape[0]] = blk
    return out


def synthetic_ply_chunks(n: int, seed: int, sh_degree: int = 3, clusters: int = 64, ball_radius: float = 1.0,
                         chunk: int = 1 << 20):
    """Generator form of `synthetic_ply_table`: yields (first_index, (m, 62) float32 block) so that multi-million
    splat scenes can be swizzled and uploaded chunk by chunk without holding the whole table."""
    rng = np.random.default_rng(seed)
    # cluster parameters
    dirs = rng.standard_normal((clusters, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    centers = dirs * (0.8 * ball_radius * rng.random((clusters, 1)) ** (1.0 / 3.0))
    sig = np.exp(rng.uniform(np.log(0.03), np.log(0.25), size=(clusters, 3))) * ball_radius
    q = rng.standard_normal((clusters, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    weights = rng.dirichlet(np.full(clusters, 2.0))
    scale_shift = np.log((DEMO_SPLATS / float(n)) ** (1.0 / 3.0))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        cid = rng.choice(clusters, size=m, p=weights)
        local = rng.standard_normal((m, 3)) * sig[cid]
        pos = np.einsum("mij,mj->mi", R[cid], local) + centers[cid]
        # keep the cloud inside the ball: fold stragglers back radially
        r = np.linalg.norm(pos, axis=1)
        far = r > ball_radius
        pos[far] *= (ball_radius * (0.9 + 0.1 * rng.random(far.sum())) / r[far])[:, None]
        pos += np.asarray(CENTER)
        blk = np.empty((m, 62), dtype=np.float32)
        blk[:, 0:3] = pos
        blk[:, 3:6] = 0.0
        blk[:, 6:9] = rng.normal(-0.4, 0.8, size=(m, 3))
        rest = rng.normal(0.0, 0.1, size=(m, 45))
        if sh_degree < 3:
            keep = {0: 0, 1: 3, 2: 8}[sh_degree]
            rest.reshape(m, 3, 15)[:, :, keep:] = 0.0
        blk[:, 9:54] = rest
        blk[:, 54] = rng.normal(-2.0, 1.8, size=m)
        blk[:, 55:58] = rng.normal(-5.5, 0.9, size=(m, 3)) + scale_shift
        rot = rng.standard_normal((m, 4))
        rot /= np.linalg.norm(rot, axis=1, keepdims=True)
        blk[:, 58:62] = rot
        yield lo, blk


def synthetic_ply(n: int, seed: int, **kw) -> PlyFile:
    return PlyFile.from_array(synthetic_ply_table(n, seed, **kw), default_properties(62))


def radix_keys(n: int, seed: int, kind: str = "tile_depth", tiles: int = 8160) -> np.ndarray:
    """c5 microbench keys: `(tile<<16)|depth16` with tile ~ U[0,tiles) and depth16 from a clustered
    ~9k-value band (the measured depth-code occupancy of demo.ply), or uniform random 32-bit."""
    rng = np.random.default_rng(seed)
    if kind == "uniform32":
        return rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    tile = rng.integers(0, tiles, size=n, dtype=np.uint32)
    band = np.sort(rng.choice(np.arange(52000, 61500), size=9000, replace=False)).astype(np.uint32)
    depth = band[np.clip(rng.normal(4500, 1800, size=n).astype(np.int64), 0, 8999)]
    return (tile << np.uint32(16)) | depth
