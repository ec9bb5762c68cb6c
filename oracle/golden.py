"""CPU oracle for the fit hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (fpc_diffrend_b200/) never does.

PARITY UNPINNED for the four rendering ops (see the header of oracle/golden.c): nvdiffrast is an
un-vendored, un-pinned dependency of the reference and is absent here; the reference holds no tests.
The torch stages below restate the reference's own code and are pinned by the camera known-answer
vectors in tests/golden/camera_kat.json (generated from /root/reference/src/torch/camera.py by
tests/golden/make_camera_kat.py).

Restated reference code (file:line under /root/reference/src/torch):
  blend()            fit.py:103-129 (prior-mode branch :115-122, north-star form V = base + D w)
  mvp_chain()        fit.py:541-553 with camera.py:27-66,108-132 and roma.unitquat_to_rotmat (fit.py:548,550)
  transform_clip()   camera.py:11-23
  render()           fit.py:134-162 (background constant 45/255 at :161)
  image_loss()       fit.py:579 (first term)
  Adam + LambdaLR    fit.py:493-505,610-613 (torch.optim used directly)
  mesh_regularisers() fit.py:578-582: pytorch3d.loss mesh_laplacian_smoothing(method='uniform') (squared by the
                     reference), mesh_edge_loss, mesh_normal_consistency.  pytorch3d (unpinned, v0.7-era API) is not
                     installed here: its published algorithms are restated below -> PARITY UNPINNED for these terms too.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BG = 45.0 / 255.0  # fit.py:161


def build():
    """Compile oracle/golden.c -> oracle/libgolden.so (gcc, see oracle/Makefile)."""
    subprocess.run(['make', '-s', '-C', _HERE, 'libgolden.so'], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libgolden.so')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, 'golden.c')):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape):
    a = np.empty(shape, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------------------
# the four ops (numpy in / numpy out)
# ---------------------------------------------------------------------------------------------------------

def rasterize_fwd(pos, tri, resolution, with_db=True, with_second=False):
    pos, ppos = _f(pos)
    tri, ptri = _i(tri)
    N, V, _ = pos.shape
    H, W = resolution
    rast, prast = _out((N, H, W, 4))
    db, pdb = _out((N, H, W, 4)) if with_db else (None, None)
    sec, psec = _out((N, H, W, 2)) if with_second else (None, None)
    lib().gold_rasterize_fwd(ppos, ptri, N, V, tri.shape[0], H, W, prast, pdb, psec)
    return rast, db, sec


def rasterize_bwd(pos, tri, rast, dy):
    pos, ppos = _f(pos)
    tri, ptri = _i(tri)
    rast, prast = _f(rast)
    dy, pdy = _f(dy)
    N, V, _ = pos.shape
    _, H, W, _ = rast.shape
    g, pg = _out((N, V, 4))
    lib().gold_rasterize_bwd(ppos, ptri, prast, pdy, N, V, tri.shape[0], H, W, pg)
    return g


def interpolate_fwd(attr, rast, tri):
    attr, pattr = _f(attr)
    rast, prast = _f(rast)
    tri, ptri = _i(tri)
    Na, Vt, A = attr.shape
    N, H, W, _ = rast.shape
    out, pout = _out((N, H, W, A))
    lib().gold_interpolate_fwd(pattr, Na, Vt, A, prast, ptri, N, tri.shape[0], H, W, pout)
    return out


def interpolate_bwd(attr, rast, tri, dy):
    attr, pattr = _f(attr)
    rast, prast = _f(rast)
    tri, ptri = _i(tri)
    dy, pdy = _f(dy)
    Na, Vt, A = attr.shape
    N, H, W, _ = rast.shape
    ga, pga = _out(attr.shape)
    gr, pgr = _out(rast.shape)
    lib().gold_interpolate_bwd(pattr, Na, Vt, A, prast, ptri, pdy, N, tri.shape[0], H, W, pga, pgr)
    return ga, gr


def texture_linear_fwd(tex, uv):
    tex, ptex = _f(tex)
    uv, puv = _f(uv)
    Nt, Ht, Wt, C = tex.shape
    N, H, W, _ = uv.shape
    out, pout = _out((N, H, W, C))
    lib().gold_texture_linear_fwd(ptex, Nt, Ht, Wt, C, puv, N, H, W, pout)
    return out


def texture_linear_bwd(tex, uv, dy):
    tex, ptex = _f(tex)
    uv, puv = _f(uv)
    dy, pdy = _f(dy)
    Nt, Ht, Wt, C = tex.shape
    N, H, W, _ = uv.shape
    gt, pgt = _out(tex.shape)
    guv, pguv = _out(uv.shape)
    lib().gold_texture_linear_bwd(ptex, Nt, Ht, Wt, C, puv, pdy, N, H, W, pgt, pguv)
    return gt, guv


def topology_build(tri):
    """tri_opp [T,3] i32: opposite vertex across edge e (e0=(v1,v2), e1=(v2,v0), e2=(v0,v1)), -1 if none.
    Non-manifold edges: the other triangle with the lowest (index, corner) code."""
    tri = np.asarray(tri, dtype=np.int64)
    T = tri.shape[0]
    t_idx = np.repeat(np.arange(T), 3)
    k_idx = np.tile(np.arange(3), T)
    va = tri[t_idx, (k_idx + 1) % 3]
    vb = tri[t_idx, (k_idx + 2) % 3]
    lo, hi = np.minimum(va, vb), np.maximum(va, vb)
    key = lo * (tri.max() + 2) + hi
    code = t_idx * 4 + k_idx
    order = np.lexsort((code, key))
    key_s, code_s = key[order], code[order]
    start = np.r_[True, key_s[1:] != key_s[:-1]]
    grp = np.cumsum(start) - 1
    first_pos = np.nonzero(start)[0]
    cnt = np.diff(np.r_[first_pos, key_s.shape[0]])
    c0 = code_s[first_pos]                                   # lowest code of each edge group
    c1 = np.where(cnt > 1, code_s[np.minimum(first_pos + 1, key_s.shape[0] - 1)], -1)  # second lowest
    my = code_s
    other = np.where(my == c0[grp], c1[grp], c0[grp])
    opp_v = np.where(other >= 0, tri[np.maximum(other, 0) // 4, np.maximum(other, 0) % 4], -1)
    out = np.full((T * 3,), -1, dtype=np.int32)
    out[order] = opp_v.astype(np.int32)
    return out.reshape(T, 3)


def antialias_fwd(color, rast, pos, tri, tri_opp=None):
    color, pc = _f(color)
    rast, pr = _f(rast)
    pos, pp = _f(pos)
    tri, pt = _i(tri)
    opp, po = _i(topology_build(tri) if tri_opp is None else tri_opp)
    N, H, W, C = color.shape
    out, pout = _out(color.shape)
    lib().gold_antialias_fwd(pc, pr, pp, pt, po, N, pos.shape[1], tri.shape[0], H, W, C, pout)
    return out


def antialias_bwd(color, rast, pos, tri, dy, tri_opp=None):
    color, pc = _f(color)
    rast, pr = _f(rast)
    pos, pp = _f(pos)
    tri, pt = _i(tri)
    dy, pdy = _f(dy)
    opp, po = _i(topology_build(tri) if tri_opp is None else tri_opp)
    N, H, W, C = color.shape
    gc, pgc = _out(color.shape)
    gp, pgp = _out(pos.shape)
    lib().gold_antialias_bwd(pc, pr, pp, pt, po, pdy, N, pos.shape[1], tri.shape[0], H, W, C, pgc, pgp)
    return gc, gp


# ---------------------------------------------------------------------------------------------------------
# autograd wrappers so the torch stages can differentiate through the golden ops on CPU
# ---------------------------------------------------------------------------------------------------------

def _np(t):
    return t.detach().cpu().numpy()


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tri, resolution):
        rast, db, _ = rasterize_fwd(_np(pos), _np(tri), resolution)
        rast = torch.from_numpy(rast)
        ctx.save_for_backward(pos, tri, rast)
        return rast, torch.from_numpy(db)

    @staticmethod
    def backward(ctx, dy, ddb):
        pos, tri, rast = ctx.saved_tensors
        return torch.from_numpy(rasterize_bwd(_np(pos), _np(tri), _np(rast), _np(dy))), None, None


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        ctx.save_for_backward(attr, rast, tri)
        return torch.from_numpy(interpolate_fwd(_np(attr), _np(rast), _np(tri)))

    @staticmethod
    def backward(ctx, dy):
        attr, rast, tri = ctx.saved_tensors
        ga, gr = interpolate_bwd(_np(attr), _np(rast), _np(tri), _np(dy))
        return torch.from_numpy(ga), torch.from_numpy(gr), None


class _Texture(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv):
        ctx.save_for_backward(tex, uv)
        return torch.from_numpy(texture_linear_fwd(_np(tex), _np(uv)))

    @staticmethod
    def backward(ctx, dy):
        tex, uv = ctx.saved_tensors
        gt, guv = texture_linear_bwd(_np(tex), _np(uv), _np(dy))
        return torch.from_numpy(gt), torch.from_numpy(guv)


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, tri_opp):
        ctx.save_for_backward(color, rast, pos, tri, tri_opp)
        return torch.from_numpy(antialias_fwd(_np(color), _np(rast), _np(pos), _np(tri), _np(tri_opp)))

    @staticmethod
    def backward(ctx, dy):
        color, rast, pos, tri, tri_opp = ctx.saved_tensors
        gc, gp = antialias_bwd(_np(color), _np(rast), _np(pos), _np(tri), _np(dy), _np(tri_opp))
        return torch.from_numpy(gc), None, torch.from_numpy(gp), None, None


def rasterize(pos, tri, resolution):
    return _Rasterize.apply(pos, tri, tuple(resolution))


def interpolate(attr, rast, tri):
    return _Interpolate.apply(attr, rast, tri)


def texture(tex, uv):
    return _Texture.apply(tex, uv)


def antialias(color, rast, pos, tri, tri_opp):
    return _Antialias.apply(color, rast, pos, tri, tri_opp)


# ---------------------------------------------------------------------------------------------------------
# torch stages (CPU), restating the reference
# ---------------------------------------------------------------------------------------------------------

def blend(v_base, D, w):
    """fit.py:103-129 in the north-star form: V = base + D @ w.  v_base [3V], D [3V,B], w [B] -> [3V]"""
    return torch.add(v_base, torch.matmul(D, w))


def unitquat_to_rotmat(q):
    """roma.unitquat_to_rotmat (XYZW, not normalised) as used at fit.py:548,550."""
    x, y, z, w = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)]),
        torch.stack([2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)]),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w]),
    ])


def rigid(tvec, rotmat):
    """camera.py:128-132 (rigid_grad) without the hard-coded device."""
    rt = torch.cat((rotmat, tvec.reshape(3, 1)), 1)
    br = torch.tensor([[0, 0, 0, 1]], dtype=rt.dtype)
    return torch.cat((rt, br), 0)


def mvp_chain(P, A, t_frame, q_frame, t_cam=None, q_cam=None):
    """fit.py:546-553: mvp = P @ (T_frame @ (T_cam @ A)),  A = MV @ translate(0,170,0)."""
    tr = A
    if t_cam is not None:
        tr = torch.matmul(rigid(t_cam, unitquat_to_rotmat(q_cam)), tr)
    tr_pose = torch.matmul(rigid(t_frame, unitquat_to_rotmat(q_frame)), tr)
    return torch.matmul(P, tr_pose)


def transform_clip(mvp, pos):
    """camera.py:11-23."""
    posw = torch.cat([pos, torch.ones([pos.shape[0], 1], dtype=pos.dtype)], axis=1)
    return torch.matmul(posw, mvp.t())[None, ...]


def render(mvp, pos, pos_idx, resolution, *, uv=None, uv_idx=None, tex=None, vcol=None, tri_opp=None,
           use_antialias=True):
    """fit.py:134-162.  Textured path (uv/uv_idx/tex) or the vertex-colour path of BASELINE config 2 (vcol)."""
    pos_clip = transform_clip(mvp, pos)
    rast_out, _ = rasterize(pos_clip, pos_idx, resolution)
    if vcol is not None:
        colour = interpolate(vcol[None, ...], rast_out, pos_idx)
    else:
        texc = interpolate(uv[None, ...], rast_out, uv_idx)
        colour = texture(tex[None, ...], texc)
    if use_antialias:
        colour = antialias(colour, rast_out, pos_clip, pos_idx, tri_opp)
    colour = torch.where(rast_out[..., 3:] > 0, colour, torch.tensor(BG))
    return colour[0]


def image_loss(ref, colour):
    """fit.py:579, first term: mean((ref - 255 colour)^2)."""
    return torch.mean((ref - colour * 255) ** 2)


# ---------------------------------------------------------------------------------------------------------
# mesh regularisers (fit.py:578-582), pytorch3d.loss semantics restated in plain torch (autograd gives the gradients)
# ---------------------------------------------------------------------------------------------------------

def mesh_laplacian_uniform(verts, edges):
    """pytorch3d mesh_laplacian_smoothing(method='uniform') for one mesh: L = D^-1 A - I over the unique edges,
    loss = (1/V) sum_i |(L v)_i|_2.  verts [V,3], edges [E,2] int64."""
    V = verts.shape[0]
    e0, e1 = edges[:, 0], edges[:, 1]
    deg = torch.zeros(V, dtype=verts.dtype).index_add_(0, e0, torch.ones(e0.shape[0], dtype=verts.dtype))
    deg = deg.index_add_(0, e1, torch.ones(e1.shape[0], dtype=verts.dtype))
    nsum = torch.zeros_like(verts).index_add(0, e0, verts[e1]).index_add(0, e1, verts[e0])
    inv = torch.where(deg > 0, 1.0 / deg.clamp(min=1), torch.zeros_like(deg))
    lv = torch.where(deg[:, None] > 0, nsum * inv[:, None] - verts, torch.zeros_like(verts))
    return lv.norm(dim=1).sum() / V


def mesh_edge_loss(verts, edges, target_length=0.0):
    """pytorch3d mesh_edge_loss: mean over unique edges of (|v0 - v1| - target)^2."""
    d = (verts[edges[:, 0]] - verts[edges[:, 1]]).norm(dim=1, p=2)
    return ((d - target_length) ** 2).sum() / edges.shape[0]


def mesh_normal_consistency(verts, edge_quads):
    """pytorch3d mesh_normal_consistency: faces (v0,v1,a), (v0,v1,b) sharing an edge,
    n0 = (v1-v0) x (a-v0), n1 = (v1-v0) x (b-v0), loss = mean(1 - cos(n0, -n1))."""
    if edge_quads.shape[0] == 0:
        return verts.sum() * 0.0
    v0, v1, a, b = (verts[edge_quads[:, k]] for k in range(4))
    n0 = torch.cross(v1 - v0, a - v0, dim=1)
    n1 = torch.cross(v1 - v0, b - v0, dim=1)
    return (1.0 - torch.cosine_similarity(n0, -n1, dim=1)).sum() / edge_quads.shape[0]


def mesh_regularisers(verts, edges, edge_quads, w_lap=5000.0, w_edge=0.0, edge_target=0.1, w_nc=0.0):
    """The three mesh terms of fit.py:579-582 for one frame (weights: main.py:37-40; the reference passes 0.1 as the edge
    target at fit.py:580).  Returns (total, (lap, edge, nc))."""
    lap = mesh_laplacian_uniform(verts, edges)
    edge = mesh_edge_loss(verts, edges, edge_target)
    nc = mesh_normal_consistency(verts, edge_quads)
    return w_lap * lap ** 2 + w_edge * edge + w_nc * nc, (lap, edge, nc)
