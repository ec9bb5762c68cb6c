"""Float64 torch-autograd restatement of the differentiable parts of the four ops — TEST INFRASTRUCTURE ONLY.

Purpose: pin the *analytic* backward passes of oracle/golden.c (and through them the CUDA kernels)
against automatic differentiation of the forward formulas of SURVEY.md Appendix A.  Visibility
(tri_id) is an input here: it comes from the golden rasterizer.  All functions are vectorised torch
ops, dtype follows the inputs (use float64 for gradient checks).
"""
import torch


def _pixel_grid(N, H, W, dtype):
    py, px = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing='ij')
    return px[None].expand(N, H, W), py[None].expand(N, H, W)


def barycentrics(pos, tri, tri_id, H, W):
    """(u, v, z/w) of App. A.1 for the given per-pixel triangle ids (tri_id [N,H,W] int64, -1 = background).
    Clamps are straight-through (the backward ignores them, App. A.1 'Backward')."""
    N = pos.shape[0]
    dt = pos.dtype
    px, py = _pixel_grid(N, H, W, dt)
    fx = (2.0 * px + 1.0) / W - 1.0
    fy = (2.0 * py + 1.0) / H - 1.0
    mask = tri_id >= 0
    t = tri_id.clamp(min=0)
    vi = tri.long()[t]                                         # [N,H,W,3]
    n_idx = torch.arange(N)[:, None, None, None].expand_as(vi)
    p = pos[n_idx, vi]                                         # [N,H,W,3,4]
    qx = p[..., 0] - fx[..., None] * p[..., 3]
    qy = p[..., 1] - fy[..., None] * p[..., 3]
    a0 = qx[..., 1] * qy[..., 2] - qy[..., 1] * qx[..., 2]
    a1 = qx[..., 2] * qy[..., 0] - qy[..., 2] * qx[..., 0]
    a2 = qx[..., 0] * qy[..., 1] - qy[..., 0] * qx[..., 1]
    at = a0 + a1 + a2
    at = torch.where(mask, at, torch.ones_like(at))
    u = a0 / at
    v = a1 / at
    z = p[..., 0, 2] * a0 + p[..., 1, 2] * a1 + p[..., 2, 2] * a2
    w = p[..., 0, 3] * a0 + p[..., 1, 3] * a1 + p[..., 2, 3] * a2
    zw = z / torch.where(mask, w, torch.ones_like(w))
    u = u + (u.clamp(0, 1) - u).detach()
    v = v + (v.clamp(0, 1) - v).detach()
    zero = torch.zeros_like(u)
    return torch.where(mask, u, zero), torch.where(mask, v, zero), torch.where(mask, zw, zero)


def interpolate(attr, rast, tri):
    """App. A.2.  attr [Na,Vt,A], rast [N,H,W,4], tri [T,3]"""
    N = rast.shape[0]
    tid = rast[..., 3].long() - 1
    mask = tid >= 0
    vi = tri.long()[tid.clamp(min=0)]
    if attr.shape[0] == 1:
        a = attr[0][vi]                                        # [N,H,W,3,A]
    else:
        n_idx = torch.arange(N)[:, None, None, None].expand_as(vi)
        a = attr[n_idx, vi]
    u, v = rast[..., 0:1], rast[..., 1:2]
    out = u * a[..., 0, :] + v * a[..., 1, :] + (1.0 - u - v) * a[..., 2, :]
    return torch.where(mask[..., None], out, torch.zeros_like(out))


def texture_linear(tex, uv):
    """App. A.3: bilinear, wrap.  tex [Nt,Ht,Wt,C], uv [N,H,W,2]"""
    Nt, Ht, Wt, C = tex.shape
    N = uv.shape[0]
    u = uv[..., 0] - torch.floor(uv[..., 0]).detach()
    v = uv[..., 1] - torch.floor(uv[..., 1]).detach()
    x = u * Wt - 0.5
    y = v * Ht - 0.5
    x0 = torch.floor(x).detach()
    y0 = torch.floor(y).detach()
    fx = (x - x0)[..., None]
    fy = (y - y0)[..., None]
    ix0 = x0.long() % Wt
    iy0 = y0.long() % Ht
    ix1 = (ix0 + 1) % Wt
    iy1 = (iy0 + 1) % Ht
    n_idx = torch.arange(N)[:, None, None].expand_as(ix0) if Nt > 1 else torch.zeros_like(ix0)
    t00, t10 = tex[n_idx, iy0, ix0], tex[n_idx, iy0, ix1]
    t01, t11 = tex[n_idx, iy1, ix0], tex[n_idx, iy1, ix1]
    a = t00 + (t10 - t00) * fx
    b = t01 + (t11 - t01) * fx
    return a + (b - a) * fy


def _same_sign(a, b):
    # sign-bit equality as in the golden (zero counts by its sign bit)
    return torch.signbit(a) == torch.signbit(b)


def _rational_gt(n0, n1, d0, d1):
    p0, p1 = n0 * d1, n1 * d0
    return torch.where(_same_sign(d0, d1), p0 > p1, p0 < p1)


def antialias(color, rast, pos, tri, tri_opp):
    """App. A.4, vectorised over all right / down pixel pairs.  Differentiable in color and pos."""
    N, H, W, C = color.shape
    dt = pos.dtype
    out = color.clone()
    tid = rast[..., 3].long() - 1
    zw = rast[..., 2]
    tri_l, opp_l = tri.long(), tri_opp.long()
    FMAX = torch.finfo(torch.float32).max
    for d in (0, 1):
        if d == 0:
            s0 = (slice(None), slice(None), slice(0, W - 1))
            s1 = (slice(None), slice(None), slice(1, W))
        else:
            s0 = (slice(None), slice(0, H - 1), slice(None))
            s1 = (slice(None), slice(1, H), slice(None))
        t0, t1 = tid[s0], tid[s1]
        z0, z1 = zw[s0], zw[s1]
        px, py = _pixel_grid(N, H, W, dt)
        px, py = px[s0], py[s0]
        t = torch.where(t0 >= 0, t0, t1)
        t = torch.where((t0 >= 0) & (t1 >= 0), torch.where(z0 < z1, t0, t1), t)
        use1 = (t == t1)
        px = px + use1.to(dt) * (1 - d)
        py = py + use1.to(dt) * d
        active = (t0 != t1) & (t >= 0)
        tc = t.clamp(min=0)
        vi = tri_l[tc]                                   # [...,3]
        op = opp_l[tc]
        n_idx = torch.arange(N)[:, None, None, None].expand_as(vi)
        p = pos[n_idx, vi]                               # [...,3,4]
        o = pos[n_idx, torch.where(op < 0, vi, op)]
        xh, yh = 0.5 * W, 0.5 * H
        fx = (px + 0.5 - xh)[..., None]
        fy = (py + 0.5 - yh)[..., None]
        x = p[..., 0] / p[..., 3] * xh - fx
        y = p[..., 1] / p[..., 3] * yh - fy
        ox = o[..., 0] / o[..., 3] * xh - fx
        oy = o[..., 1] / o[..., 3] * yh - fy
        x0, x1, x2 = x.unbind(-1)
        y0, y1, y2 = y.unbind(-1)
        bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        a0 = (x1 - ox[..., 0]) * (y2 - oy[..., 0]) - (x2 - ox[..., 0]) * (y1 - oy[..., 0])
        a1 = (x2 - ox[..., 1]) * (y0 - oy[..., 1]) - (x0 - ox[..., 1]) * (y2 - oy[..., 1])
        a2 = (x0 - ox[..., 2]) * (y1 - oy[..., 2]) - (x1 - ox[..., 2]) * (y0 - oy[..., 2])
        sil0, sil1, sil2 = _same_sign(a0, bb), _same_sign(a1, bb), _same_sign(a2, bb)
        active = active & (sil0 | sil1 | sil2)
        if d:
            x0, y0, x1, y1, x2, y2 = y0, x0, y1, x1, y2, x2
        dx0, dx1, dx2 = x2 - x1, x0 - x2, x1 - x0
        dy0, dy1, dy2 = y2 - y1, y0 - y2, y1 - y0
        ds = torch.where(t == t0, torch.ones_like(bb), -torch.ones_like(bb))
        d0 = ds * (x1 * dy0 - y1 * dx0)
        d1 = ds * (x2 * dy1 - y2 * dx1)
        d2 = ds * (x0 * dy2 - y0 * dx2)
        k0, k1, k2 = _same_sign(y1, y2), _same_sign(y2, y0), _same_sign(y0, y1)
        neg = torch.full_like(bb, -FMAX)
        one = torch.ones_like(bb)
        d0m, dy0m = torch.where(k0, neg, d0), torch.where(k0, one, dy0)
        d1m, dy1m = torch.where(k1, neg, d1), torch.where(k1, one, dy1)
        d2m, dy2m = torch.where(k2, neg, d2), torch.where(k2, one, dy2)
        g10 = _rational_gt(d1m, d0m, dy1m, dy0m)
        g20 = _rational_gt(d2m, d0m, dy2m, dy0m)
        g21 = _rational_gt(d2m, d1m, dy2m, dy1m)
        di = torch.where(g20 & g21, 2, torch.where(g10, 1, 0))
        dc = neg.clone()
        ok0 = (di == 0) & sil0 & (dy0m.abs() >= dx0.abs())
        ok1 = (di == 1) & sil1 & (dy1m.abs() >= dx1.abs())
        ok2 = (di == 2) & sil2 & (dy2m.abs() >= dx2.abs())
        dc = torch.where(ok0, d0m / dy0m, dc)
        dc = torch.where(ok1, d1m / dy1m, dc)
        dc = torch.where(ok2, d2m / dy2m, dc)
        eps = 0.0625
        hit = active & (dc > -eps) & (dc < 1.0 + eps)
        dc = torch.where(hit, dc, torch.zeros_like(dc))
        dc = dc + (dc.clamp(0, 1) - dc).detach()
        alpha = torch.where(hit, ds * (0.5 - dc), torch.zeros_like(dc))
        c0, c1 = color[s0], color[s1]
        contrib = alpha[..., None] * (c1 - c0)
        to0 = (alpha > 0)[..., None]
        add0 = torch.where(to0, contrib, torch.zeros_like(contrib))
        add1 = torch.where(to0, torch.zeros_like(contrib), contrib)
        pad0 = torch.zeros_like(out)
        pad1 = torch.zeros_like(out)
        pad0[s0] = add0
        pad1[s1] = add1
        out = out + pad0 + pad1
    return out


# ---------------------------------------------------------------------------------------------------------
# mip-mapped texturing path (reference fit.py:153-155 with enable_mip: rasterize -> rast_db, interpolate(...,
# rast_db, diff_attrs='all') -> uv_da, texture(..., uv_da, filter_mode='linear-mipmap-linear', max_mip_level)).
# PARITY UNPINNED like the other rendering ops: restated from the nvdiffrast paper (Laine et al. 2020, sec. 3.3-3.4:
# analytic screen-space attribute derivatives, trilinear mip-mapped lookup with the footprint's major axis selecting
# the level) and the recalled upstream conventions listed in SURVEY App. A.
# ---------------------------------------------------------------------------------------------------------

def barycentric_diffs(pos, tri, tri_id, H, W):
    """rast_db = (du/dX, du/dY, dv/dX, dv/dY), X/Y in pixels (App. A.1), for given per-pixel triangle ids (-1 = bg).
    a_k = C_k + A_k fx + B_k fy  ->  du/dfx = (A_0 - u sum A) / sum a,  fx = (2 X + 1)/W - 1  ->  d fx / dX = 2 / W."""
    N = pos.shape[0]
    dt = pos.dtype
    px, py = _pixel_grid(N, H, W, dt)
    fx = (2.0 * px + 1.0) / W - 1.0
    fy = (2.0 * py + 1.0) / H - 1.0
    mask = tri_id >= 0
    vi = tri.long()[tri_id.clamp(min=0)]
    n_idx = torch.arange(N)[:, None, None, None].expand_as(vi)
    p = pos[n_idx, vi]
    x, y, w = p[..., 0], p[..., 1], p[..., 3]
    nxt, prv = [1, 2, 0], [2, 0, 1]
    A = y[..., nxt] * w[..., prv] - w[..., nxt] * y[..., prv]
    B = w[..., nxt] * x[..., prv] - x[..., nxt] * w[..., prv]
    C = x[..., nxt] * y[..., prv] - y[..., nxt] * x[..., prv]
    a = C + A * fx[..., None] + B * fy[..., None]
    at = torch.where(mask, a.sum(-1), torch.ones_like(fx))
    u, v = a[..., 0] / at, a[..., 1] / at
    At, Bt = A.sum(-1), B.sum(-1)
    xs, ys = 2.0 / W, 2.0 / H
    db = torch.stack([xs * (A[..., 0] - u * At) / at, ys * (B[..., 0] - u * Bt) / at,
                      xs * (A[..., 1] - v * At) / at, ys * (B[..., 1] - v * Bt) / at], dim=-1)
    return torch.where(mask[..., None], db, torch.zeros_like(db))


def interpolate_da(attr, rast, tri, rast_db, diff_attrs='all'):
    """App. A.2 with pixel differentials: out_da [N,H,W,2k] = (da/dX, da/dY) per selected attribute."""
    N = rast.shape[0]
    tid = rast[..., 3].long() - 1
    mask = tid >= 0
    vi = tri.long()[tid.clamp(min=0)]
    if attr.shape[0] == 1:
        a = attr[0][vi]
    else:
        n_idx = torch.arange(N)[:, None, None, None].expand_as(vi)
        a = attr[n_idx, vi]
    sel = list(range(attr.shape[2])) if diff_attrs == 'all' else list(diff_attrs)
    a = a[..., sel]                                           # [N,H,W,3,k]
    e0, e1 = a[..., 0, :] - a[..., 2, :], a[..., 1, :] - a[..., 2, :]
    dadx = rast_db[..., 0:1] * e0 + rast_db[..., 2:3] * e1
    dady = rast_db[..., 1:2] * e0 + rast_db[..., 3:4] * e1
    da = torch.stack([dadx, dady], dim=-1).reshape(a.shape[:3] + (2 * len(sel),))
    return torch.where(mask[..., None], da, torch.zeros_like(da))


def texture_construct_mip(tex, max_mip_level=None):
    """Mip stack [level 0 = tex, level 1, ...]: every texel of a level is the mean of its 2x2 parents.  Levels are added
    while both extents are even (and > 1) and max_mip_level is not exceeded."""
    levels = [tex]
    while (max_mip_level is None or len(levels) - 1 < max_mip_level):
        t = levels[-1]
        h, w = t.shape[1], t.shape[2]
        if h < 2 or w < 2 or h % 2 or w % 2:
            break
        levels.append(0.25 * (t[:, 0::2, 0::2] + t[:, 0::2, 1::2] + t[:, 1::2, 0::2] + t[:, 1::2, 1::2]))
    return levels


def mip_level(uv_da, Ht, Wt, bias=None):
    """Mip level from the texture-space footprint of a pixel: half the log2 of the squared major axis of the ellipse
    spanned by (ds/dX, dt/dX), (ds/dY, dt/dY) in texels."""
    dsdx, dsdy = uv_da[..., 0] * Wt, uv_da[..., 1] * Wt
    dtdx, dtdy = uv_da[..., 2] * Ht, uv_da[..., 3] * Ht
    A = dsdx * dsdx + dtdx * dtdx
    B = dsdy * dsdy + dtdy * dtdy
    C = dsdx * dsdy + dtdx * dtdy
    l2b = 0.5 * (A + B)
    l2n = 0.25 * (A - B) * (A - B) + C * C
    major = l2b + torch.sqrt(l2n)
    lev = 0.5 * torch.log2(major)
    return lev if bias is None else lev + bias


def texture_mip(tex, uv, uv_da=None, mip_level_bias=None, max_mip_level=None, filter_mode='linear-mipmap-linear'):
    """Trilinear (or nearest-level) mip-mapped lookup, boundary wrap.  The level is clamped to [0, L]; at or below 0 the
    lookup is plain bilinear on level 0, at L bilinear on the coarsest level."""
    levels = texture_construct_mip(tex, max_mip_level)
    L = len(levels) - 1
    Ht, Wt = tex.shape[1], tex.shape[2]
    if uv_da is not None:
        lev = mip_level(uv_da, Ht, Wt, mip_level_bias)
    else:
        lev = mip_level_bias
    lev = torch.nan_to_num(lev, nan=0.0, neginf=0.0, posinf=float(L))
    levc = lev.clamp(0.0, float(L))
    if filter_mode == 'linear-mipmap-nearest':
        l0 = torch.floor(levc + 0.5).long().clamp(max=L)
        out = torch.zeros(uv.shape[:3] + (tex.shape[3],), dtype=tex.dtype)
        for l in range(L + 1):
            out = torch.where((l0 == l)[..., None], texture_linear(levels[l], uv), out)
        return out
    l0 = torch.floor(levc).long().clamp(max=L)
    l1 = (l0 + 1).clamp(max=L)
    f = (levc - l0.to(levc.dtype))[..., None]
    c0 = torch.zeros(uv.shape[:3] + (tex.shape[3],), dtype=tex.dtype)
    c1 = torch.zeros_like(c0)
    for l in range(L + 1):
        s = texture_linear(levels[l], uv)
        c0 = torch.where((l0 == l)[..., None], s, c0)
        c1 = torch.where((l1 == l)[..., None], s, c1)
    return c0 + (c1 - c0) * f
