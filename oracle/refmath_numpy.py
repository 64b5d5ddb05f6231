"""Independent float64 transliteration of gsplat_projection.glsl / gsplat_boundaries.glsl / gsplat_render.glsl.

TEST INFRASTRUCTURE (see gsr_oracle.c header).  Purpose: a second, independently written reading of the shaders
in "real-number" arithmetic (numpy float64, libm exp/pow, no attention to float32 rounding or FMA placement), used by
tests/test_oracle.py to bound how far the deterministic float32 oracle is from the shaders' mathematical meaning:
  * cull decisions / tile rects / depth codes agree except for splats that sit within rounding distance of a
    boundary (the tests assert >= 99.9 % identical keys);
  * pixels agree to ~1e-5 wherever both sides blend the same instance lists.
It deliberately shares no code with gsr_oracle.c.
"""
from __future__ import annotations

import numpy as np

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, 1.0925484305920792, 0.31539156525252005, 1.0925484305920792, 0.5462742152960396]
SH_C3 = [0.5900435899266435, 2.890611442640554, 0.4570457994644658, 0.3731763325901154, 0.4570457994644658,
         1.445305721320277, 0.5900435899266435]


def project(splat60, vp32, camera_pos, model_scale, W, H, time):
    """gsplat_projection.glsl:150-227 in float64. Returns dict of per-splat arrays (visible mask, rect, depth, record)."""
    s = np.asarray(splat60, dtype=np.float64).reshape(-1, 60)
    V = np.asarray(vp32[:16], dtype=np.float64).reshape(4, 4).T   # V[r][c] (GLSL column-major -> row-major matrix)
    P = np.asarray(vp32[16:], dtype=np.float64).reshape(4, 4).T
    n = s.shape[0]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    p = s[:, 0:3] * model_scale
    view = np.concatenate([p, np.ones((n, 1))], 1) @ V.T
    clip = view @ P.T
    vb = clip[:, 3] * 1.2
    culled = ((clip[:, 0] < -vb) | (clip[:, 1] < -vb) | (clip[:, 2] < 0) | (clip[:, 0] > vb) | (clip[:, 1] > vb) | (clip[:, 2] > clip[:, 3]))

    def ease(x):
        a = 1.0 - x
        return 1.0 - a * a * a

    st = time - s[:, 3]
    tf = ease(np.clip(st, 0, 1))
    tfl = ease(np.clip(st - 0.35, 0, 1))
    opacity = s[:, 10] * tfl * tfl
    scale = model_scale * (2.0 * (1 - tfl) + tfl)
    c = s[:, 4:10]
    cov3 = np.stack([np.stack([c[:, 0], c[:, 1], c[:, 2]], -1), np.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                     np.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], 1) * (scale * scale)[:, None, None]
    tfi = np.array([P[0, 0], P[1, 1]])
    focal = np.array([W, H]) * 0.5 * tfi
    tanfov = 1.0 / tfi
    with np.errstate(divide="ignore", invalid="ignore"):
        zinv = 1.0 / view[:, 2]
        f = focal[None, :] * zinv[:, None]
        m = np.clip(view[:, 0:2] * zinv[:, None], -tanfov * 1.3, tanfov * 1.3)
        # J (2x3, usual orientation): rows (fx, 0, -fy*mx), (0, fy, -fy*my)   [Q1: fy in both]
        J = np.zeros((n, 2, 3))
        J[:, 0, 0] = f[:, 0]; J[:, 0, 2] = -f[:, 1] * m[:, 0]
        J[:, 1, 1] = f[:, 1]; J[:, 1, 2] = -f[:, 1] * m[:, 1]
        Wm = V[:3, :3]
        T = J @ Wm[None]
        cov2 = T @ cov3 @ np.transpose(T, (0, 2, 1))
        cx, cy, cz = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
        det = cx * cz - cy * cy
        mid = 0.5 * (cx + cz)
        sq = np.sqrt(np.maximum(0.1, mid * mid - det))
        e1, e2 = mid + sq, mid - sq
        culled |= (det == 0) | (e1 < 0) | (e2 < 0)
        ndc = clip[:, :3] / clip[:, 3:4]
        ipx = ((ndc[:, 0] + 1) * 0.5 - 1.0 * (1 - tf)) * (W - 1)
        ipy = ((ndc[:, 1] + 1) * 0.5 - 0.75 * (1 - tf)) * (H - 1)
        radius = np.power(np.maximum(opacity, 0), 0.2) * 2.5 * np.sqrt(np.maximum(e1, e2))
        x0 = np.trunc(np.clip((ipx - radius) / 16, 0, gx)); y0 = np.trunc(np.clip((ipy - radius) / 16, 0, gy))
        x1 = np.trunc(np.clip(np.ceil((ipx + radius) / 16), 0, gx)); y1 = np.trunc(np.clip(np.ceil((ipy + radius) / 16), 0, gy))
        bad = ~np.isfinite(ipx) | ~np.isfinite(ipy) | ~np.isfinite(radius)
    culled |= bad
    x0, y0, x1, y1 = [np.where(culled, 0, v).astype(np.int64) for v in (x0, y0, x1, y1)]
    ntiles = (x1 - x0) * (y1 - y0)
    ntiles[culled] = 0
    d = p - np.asarray(camera_pos, dtype=np.float64)[None]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    sh = s[:, 12:60].reshape(n, 16, 3)
    basis = np.stack([np.full(n, SH_C0), -SH_C1 * y, SH_C1 * z, -SH_C1 * x,
                      SH_C2[0] * xy, -SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), -SH_C2[3] * xz, SH_C2[4] * (xx - yy),
                      -SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * x * yz, -SH_C3[2] * y * (4 * zz - xx - yy),
                      SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), -SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
                      -SH_C3[6] * x * (xx - 3 * yy)], 1)
    color = np.maximum(0.0, 0.5 + np.einsum("nk,nkc->nc", basis, sh))
    with np.errstate(invalid="ignore", divide="ignore"):
        conic = np.stack([cz, -cy, cx], 1) / det[:, None]
        depth = (np.clip(ndc[:, 2], 0, 1) ** 3 * 65535.0).astype(np.int64) & 0xFFFF
    return dict(ntiles=ntiles, rect=np.stack([x0, y0, x1, y1], 1), depth=depth, image_pos=np.stack([ipx, ipy], 1), conic=conic,
                color=color, opacity=opacity, gx=gx, gy=gy)


def emit_and_sort(pr):
    """Duplication (:219-226) in splat-id order + stable sort by key."""
    keys, vals = [], []
    gx = pr["gx"]
    for i in np.nonzero(pr["ntiles"])[0]:
        x0, y0, x1, y1 = pr["rect"][i]
        ys, xs = np.meshgrid(np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
        t = (ys * gx + xs).reshape(-1)
        keys.append((t << 16) | pr["depth"][i])
        vals.append(np.full(t.size, i))
    if not keys:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    k, v = np.concatenate(keys), np.concatenate(vals)
    o = np.argsort(k, kind="stable")
    return k[o], v[o]


def render_pixels(pr, values, start, count, px, py):
    """gsplat_render.glsl:79-91 for ONE pixel over instances [start, start+count) without the tile-stop rule
    (callers pass the instance count the oracle consumed, so both sides blend the same list)."""
    col = np.zeros(3)
    t = 1.0
    for k in range(start, start + count):
        if not (t > 1.0 / 255.0):
            break
        i = values[k]
        ox, oy = pr["image_pos"][i, 0] - px, pr["image_pos"][i, 1] - py
        cn = pr["conic"][i]
        power = -0.5 * (cn[0] * ox * ox + cn[2] * oy * oy) - cn[1] * ox * oy
        alpha = pr["opacity"][i] * np.exp(power)
        col += pr["color"][i] * alpha * t
        t *= 1.0 - alpha
    return col, t
