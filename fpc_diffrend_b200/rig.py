"""Synthetic rig / camera / reference-frame generator (numpy + scipy, host only).

Follows SURVEY.md §8(d) "Synthetic inputs": a closed genus-0 "head" with exactly V vertices and
2V-4 triangles (so 1k/2k, 20k/40k and 50k/100k of BASELINE.json:configs come out exactly), randomised
vertex order, a spherical UV unwrap with one seam (Vt > V, uv_idx != pos_idx as data.py:33-34 allows),
Gaussian-bump blendshape deltas D [3V,B] (row-major, xyz-interleaved rows as fit.py:219 builds them),
and 9 cameras in the calibration.json schema (calibrate.py:71-72) placed like the real pods
(SURVEY App. C.2).
"""
from dataclasses import dataclass, field

import numpy as np

from . import camera as cam

HEAD_RADII = (8.0, 11.0, 9.5)


@dataclass
class Rig:
    v_base: np.ndarray      # [3V] f32, (x,y,z,x,...) like MeshData.vertices (data.py:36)
    pos_idx: np.ndarray     # [T,3] i32
    uv: np.ndarray          # [Vt,2] f32
    uv_idx: np.ndarray      # [T,3] i32
    D: np.ndarray           # [3V,B] f32 blendshape deltas (fit.py:219)
    vcol: np.ndarray        # [V,3] f32 smooth vertex colours in [0,1] (config 2 shading)
    tex: np.ndarray         # [Ht,Wt,1] f32 texture in [0,1]
    calib: dict             # calibration.json-schema dict, insertion order = camera order
    P: np.ndarray = field(default=None)   # [C,4,4] f32
    A: np.ndarray = field(default=None)   # [C,4,4] f32  (MV @ T170)

    @property
    def V(self):
        return self.v_base.shape[0] // 3

    @property
    def T(self):
        return self.pos_idx.shape[0]

    @property
    def B(self):
        return self.D.shape[1]

    @property
    def C(self):
        return self.P.shape[0]


def _fibonacci_sphere(n):
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], axis=1)


def make_head_mesh(n_vertices, rng):
    """Closed genus-0 triangulation with exactly n_vertices / 2n-4 triangles (convex hull of sphere points)."""
    from scipy.spatial import ConvexHull
    s = _fibonacci_sphere(n_vertices)
    hull = ConvexHull(s)
    tri = hull.simplices.astype(np.int64)
    assert tri.shape[0] == 2 * n_vertices - 4, tri.shape
    # orient outward (counter-clockwise seen from outside)
    a, b, c = s[tri[:, 0]], s[tri[:, 1]], s[tri[:, 2]]
    flip = np.einsum('ij,ij->i', np.cross(b - a, c - a), a + b + c) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    # real assets are not cache coherent: shuffle vertex and triangle order once
    perm = rng.permutation(n_vertices)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(n_vertices)
    s = s[perm]
    tri = inv[tri]
    tri = tri[rng.permutation(tri.shape[0])]
    pos = s * np.asarray(HEAD_RADII)[None, :]
    return s, pos.astype(np.float32), tri.astype(np.int32)


def make_uv(sphere_pts, tri):
    """Spherical unwrap with one seam: per-triangle uv indices, duplicated uv rows along the seam."""
    u = np.arctan2(sphere_pts[:, 2], sphere_pts[:, 0]) / (2 * np.pi) + 0.5
    v = np.arccos(np.clip(sphere_pts[:, 1], -1, 1)) / np.pi
    u = 0.02 + 0.96 * u * 0.5          # keep the wrapped copies (u + 0.48) inside (0,1)
    v = 0.02 + 0.96 * v
    uv = [np.stack([u, v], axis=1)]
    uv_idx = tri.astype(np.int64).copy()
    n = sphere_pts.shape[0]
    tu = u[tri]
    crosses = (tu.max(axis=1) - tu.min(axis=1)) > 0.24
    dup = {}
    extra = []
    for t in np.nonzero(crosses)[0]:
        for k in range(3):
            vi = int(tri[t, k])
            if u[vi] < 0.25:
                if vi not in dup:
                    dup[vi] = n + len(extra)
                    extra.append((u[vi] + 0.48, v[vi]))
                uv_idx[t, k] = dup[vi]
    if extra:
        uv.append(np.asarray(extra))
    return np.concatenate(uv, axis=0).astype(np.float32), uv_idx.astype(np.int32)


def make_blendshapes(pos, n_shapes, rng):
    V = pos.shape[0]
    centres = pos[rng.integers(0, V, n_shapes)].astype(np.float64)
    sigma = rng.uniform(1.5, 4.0, n_shapes)
    dirs = rng.normal(size=(n_shapes, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    amp = rng.uniform(0.3, 1.0, n_shapes)
    D = np.empty((V * 3, n_shapes), dtype=np.float32)
    p = pos.astype(np.float64)
    for b0 in range(0, n_shapes, 16):
        b1 = min(n_shapes, b0 + 16)
        d2 = ((p[:, None, :] - centres[None, b0:b1, :]) ** 2).sum(-1)
        g = np.exp(-0.5 * d2 / sigma[None, b0:b1] ** 2) * amp[None, b0:b1]
        D[:, b0:b1] = (g[:, :, None] * dirs[None, b0:b1, :]).transpose(0, 2, 1).reshape(V * 3, b1 - b0)
    return D


def make_texture(size, rng):
    """Band-limited noise in [0,1], [size,size,1]."""
    f = np.fft.rfft2(rng.normal(size=(size, size)))
    ky = np.fft.fftfreq(size)[:, None]
    kx = np.fft.rfftfreq(size)[None, :]
    f *= np.exp(-((kx ** 2 + ky ** 2) / (2 * 0.02 ** 2)))
    t = np.fft.irfft2(f, s=(size, size))
    t = 0.03 + 0.5 * (t - t.min()) / (t.max() - t.min())   # 255 * t inside the [0,140] clip of fit.py:531
    return t[..., None].astype(np.float32)


def _look_at_opencv(centre, target):
    fwd = target - centre
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=0)   # rows: camera x (right), y (down), z (forward)
    t = -R @ centre
    return R, t


def make_cameras(n_cams, width, height, dist=171.5):
    """Cameras in the calibration.json schema; 3 azimuths x 3 elevations like the real pods."""
    az = [-25.0, 0.0, 25.0]
    el = [-8.0, 0.0, 8.0]
    names = ['pod%d%s' % (i + 1, s) for i in range(3) for s in ('primary', 'secondary', 'texture')]
    target = np.array([0.0, cam.MODEL_Y_OFFSET, 0.0])
    calib = {}
    k = 0
    for a in az:
        for e in el:
            if k >= n_cams:
                break
            ar, er = np.radians(a), np.radians(e)
            d = np.array([np.sin(ar) * np.cos(er), np.sin(er), np.cos(ar) * np.cos(er)])
            centre = target + dist * d
            R, t = _look_at_opencv(centre, target)
            fx = 11.0 * (width / 2.0)
            fy = 11.0 * (height / 2.0)
            calib[names[k]] = {
                'distortion': [[0.0]] * 5,
                'intrinsic': [[fx, 0.0, width / 2.0], [0.0, fy, height / 2.0], [0.0, 0.0, 1.0]],
                'rotation': R.tolist(),
                'translation': [[float(x)] for x in t],
            }
            k += 1
    return calib


def make_rig(n_vertices=1000, n_shapes=16, n_cams=1, width=128, height=128, tex_size=64, seed=0):
    rng = np.random.default_rng(seed)
    sph, pos, tri = make_head_mesh(n_vertices, rng)
    uv, uv_idx = make_uv(sph, tri)
    D = make_blendshapes(pos, n_shapes, rng)
    # 255 * colour stays inside the reference's [0,140] clip (fit.py:531)
    vcol = (0.28 + 0.25 * np.sin(sph * np.array([3.0, 5.0, 4.0]) + np.array([0.0, 1.0, 2.0]))).astype(np.float32)
    tex = make_texture(tex_size, rng)
    calib = make_cameras(n_cams, width, height)
    P, A = cam.camera_constants(list(calib.values()))
    return Rig(v_base=pos.reshape(-1).copy(), pos_idx=tri, uv=uv, uv_idx=uv_idx, D=D, vcol=vcol, tex=tex,
               calib=calib, P=P, A=A)


def make_targets(n_frames, n_shapes, seed=1):
    """Ground-truth activations / poses the reference frames are rendered from."""
    rng = np.random.default_rng(seed)
    active = rng.random((n_shapes,)) < 0.1
    if not active.any():
        active[rng.integers(0, n_shapes)] = True
    w = np.zeros((n_frames, n_shapes), dtype=np.float64)
    cur = rng.uniform(0.0, 0.6, n_shapes) * active
    for f in range(n_frames):
        cur = np.clip(cur + rng.normal(scale=0.05, size=n_shapes) * active, 0.0, 1.0)
        w[f] = cur
    t = rng.normal(scale=0.5, size=(n_frames, 3))
    axis = rng.normal(size=(n_frames, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = np.radians(rng.uniform(0.0, 3.0, n_frames))
    q = np.concatenate([axis * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    return w.astype(np.float32), t.astype(np.float32), q.astype(np.float32)


def write_obj(path, v_flat, uv, pos_idx, uv_idx):
    """OBJ with the reference's conventions (data.py:17-39): v / vt / f v/vt, 1-based, triangles only."""
    with open(path, 'w') as f:
        for x, y, z in np.asarray(v_flat).reshape(-1, 3):
            f.write('v %r %r %r\n' % (float(x), float(y), float(z)))
        for u, v in uv:
            f.write('vt %r %r\n' % (float(u), float(v)))
        for a, b in zip(pos_idx, uv_idx):
            f.write('f %d/%d %d/%d %d/%d\n' % (a[0] + 1, b[0] + 1, a[1] + 1, b[1] + 1, a[2] + 1, b[2] + 1))
