"""CPU tests that pin the oracle: golden rasterizer semantics (known-answer cases), and its analytic backward
passes against float64 torch autograd of the forward formulas (oracle/torch_ref.py)."""
import numpy as np
import pytest
import torch

from conftest import clip_positions
from oracle import golden as G
from oracle import torch_ref as TR


def _quad(z0=0.0, z1=0.0):
    # two triangles sharing the diagonal of the square [-0.5,0.5]^2, w = 1
    pos = np.array([[[-0.5, -0.5, z0, 1], [0.5, -0.5, z0, 1], [0.5, 0.5, z1, 1], [-0.5, 0.5, z1, 1]]], np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return pos, tri


def test_raster_watertight_and_layout():
    pos, tri = _quad()
    rast, db, _ = G.rasterize_fwd(pos, tri, (16, 16))
    ids = rast[0, ..., 3]
    # square covers pixel centres in (-0.5,0.5): pixels 4..11 in both axes; shared diagonal drawn exactly once
    assert (ids[4:12, 4:12] > 0).all()
    assert (ids > 0).sum() == 64
    assert set(np.unique(ids)) == {0.0, 1.0, 2.0}
    # background is all zeros in all four channels and in db
    assert (rast[0][ids == 0] == 0).all() and (db[0][ids == 0] == 0).all()
    # barycentrics: weight of vertex 2 is 1-u-v; at pixel (x=11,y=4) (near vertex 1 of triangle 0) v ~ 1
    assert rast[0, 4, 11, 1] > 0.8 and rast[0, 4, 11, 3] == 1.0
    # row 0 is the bottom: triangle 0 (0,1,2) owns the lower-right half
    assert ids[5, 10] == 1.0 and ids[10, 5] == 2.0


def test_raster_depth_order_and_tie_break():
    pos, tri = _quad()
    # duplicate the quad behind (z=0.5) and in front (z=-0.5): the front one must win everywhere
    far = pos.copy(); far[..., 2] = 0.5
    near = pos.copy(); near[..., 2] = -0.5
    P = np.concatenate([pos, far, near], axis=1)
    T = np.concatenate([tri, tri + 4, tri + 8])
    rast, _, sec = G.rasterize_fwd(P, T, (16, 16), with_second=True)
    ids = rast[0, ..., 3]
    assert set(np.unique(ids[ids > 0])) == {5.0, 6.0}
    np.testing.assert_allclose(rast[0, ..., 2][ids > 0], -0.5)
    np.testing.assert_allclose(sec[0][..., 0][ids > 0], -0.5)      # winner / runner-up plane depths
    np.testing.assert_allclose(sec[0][..., 1][ids > 0], 0.0)
    # exact depth tie: the lower triangle index wins (in-order LESS)
    P2 = np.concatenate([pos, pos], axis=1)
    T2 = np.concatenate([tri + 4, tri])           # triangles 0,1 use the second copy, 2,3 the first
    rast2, _, _ = G.rasterize_fwd(P2, T2, (16, 16))
    assert set(np.unique(rast2[0, ..., 3])) == {0.0, 1.0, 2.0}


def test_raster_drops_behind_camera_and_depth_range():
    pos, tri = _quad()
    behind = pos.copy(); behind[0, 1, 3] = -1.0      # vertex 1 belongs to triangle 0 only; with z = 0 the whole triangle lies beyond the near plane: nothing survives the clipper
    rast, _, _ = G.rasterize_fwd(behind, tri, (16, 16))
    assert (rast[0, ..., 3] == 2.0).sum() > 0 and (rast[0, ..., 3] == 1.0).sum() == 0
    outside = pos.copy(); outside[..., 2] = 1.5      # z/w beyond the far plane: discarded
    rast, _, _ = G.rasterize_fwd(outside, tri, (16, 16))
    assert (rast == 0).all()
    degenerate = pos.copy(); degenerate[0, 2] = degenerate[0, 1]
    rast, _, _ = G.rasterize_fwd(degenerate, np.array([[0, 1, 2]], np.int32), (16, 16))
    assert (rast == 0).all()


def test_raster_matches_float64_barycentrics(tiny_rig):
    pc = clip_positions(tiny_rig, w=np.linspace(0, 0.5, 16))
    rast, db, sec = G.rasterize_fwd(pc, tiny_rig.pos_idx, (128, 128), with_second=True)
    cov = (rast[..., 3] > 0).mean()
    assert 0.2 < cov < 0.4                                       # 25-30 % coverage
    tid = torch.tensor(rast[..., 3]).long() - 1
    u, v, zw = TR.barycentrics(torch.tensor(pc).double(), torch.tensor(tiny_rig.pos_idx), tid, 128, 128)
    assert np.abs(u.numpy() - rast[..., 0]).max() < 2e-5
    assert np.abs(v.numpy() - rast[..., 1]).max() < 2e-5
    assert np.abs(zw.numpy() - rast[..., 2]).max() < 1e-6
    # db = analytic pixel differentials of (u, v): check against central differences of the fp64 formula
    H = W = 128
    fg = rast[..., 3] > 0
    du_dx = np.zeros_like(rast[..., 0]); dv_dy = np.zeros_like(du_dx)
    inner = fg[:, 1:-1, 1:-1] & (rast[:, 1:-1, 2:, 3] == rast[:, 1:-1, 1:-1, 3]) & (rast[:, 1:-1, :-2, 3] == rast[:, 1:-1, 1:-1, 3])
    un = u.numpy()
    du_dx_fd = (un[:, 1:-1, 2:] - un[:, 1:-1, :-2]) / 2
    sel = inner & (un[:, 1:-1, 2:] > 0) & (un[:, 1:-1, 2:] < 1) & (un[:, 1:-1, :-2] > 0) & (un[:, 1:-1, :-2] < 1)
    assert sel.sum() > 50
    np.testing.assert_allclose(db[:, 1:-1, 1:-1, 0][sel], du_dx_fd[sel], atol=2e-4)


def _rand_like(a, seed):
    return np.random.default_rng(seed).normal(size=a.shape).astype(np.float32)


def test_rasterize_bwd_matches_autograd(tiny_rig):
    pc = clip_positions(tiny_rig, w=np.linspace(0, 0.5, 16))
    rast, _, _ = G.rasterize_fwd(pc, tiny_rig.pos_idx, (128, 128))
    dy = _rand_like(rast, 1)
    g = G.rasterize_bwd(pc, tiny_rig.pos_idx, rast, dy)
    pos = torch.tensor(pc).double().requires_grad_(True)
    tid = torch.tensor(rast[..., 3]).long() - 1
    u, v, zw = TR.barycentrics(pos, torch.tensor(tiny_rig.pos_idx), tid, 128, 128)
    (u * torch.tensor(dy[..., 0]).double() + v * torch.tensor(dy[..., 1]).double()).sum().backward()
    ref = pos.grad.numpy()
    assert np.abs(ref[..., 2]).max() == 0 and np.abs(g[..., 2]).max() == 0   # no gradient to clip z
    scale = np.abs(ref).max()
    assert scale > 0
    assert np.abs(g - ref).max() / scale < 1e-4


@pytest.mark.parametrize('A,bc', [(2, True), (3, True), (5, False)])
def test_interpolate_matches_autograd(small_rig3, A, bc):
    r = small_rig3
    pc = clip_positions(r)
    H, W = 152, 200
    rast, _, _ = G.rasterize_fwd(pc, r.pos_idx, (H, W))
    N = rast.shape[0]
    rng = np.random.default_rng(2)
    if A == 2:
        attr, idx = r.uv[None], r.uv_idx
    else:
        attr, idx = rng.normal(size=(1 if bc else N, r.V, A)).astype(np.float32), r.pos_idx
    out = G.interpolate_fwd(attr, rast, idx)
    at = torch.tensor(attr).double().requires_grad_(True)
    ra = torch.tensor(rast).double().requires_grad_(True)
    ref = TR.interpolate(at, ra, torch.tensor(idx))
    assert np.abs(out - ref.detach().numpy()).max() < 1e-5
    dy = _rand_like(out, 3)
    ga, gr = G.interpolate_bwd(attr, rast, idx, dy)
    (ref * torch.tensor(dy).double()).sum().backward()
    assert np.abs(ga - at.grad.numpy()).max() / np.abs(at.grad.numpy()).max() < 1e-4
    assert np.abs(gr[..., :2] - ra.grad.numpy()[..., :2]).max() / np.abs(ra.grad.numpy()).max() < 1e-4
    assert (gr[..., 2:] == 0).all()
    assert (out[rast[..., 3] == 0] == 0).all()


@pytest.mark.parametrize('C,Nt', [(1, 1), (3, 1), (2, 2)])
def test_texture_matches_autograd(C, Nt):
    rng = np.random.default_rng(4)
    N, H, W, Ht, Wt = 2, 9, 11, 8, 16
    tex = rng.random((Nt, Ht, Wt, C)).astype(np.float32)
    uv = rng.uniform(-1.5, 2.5, size=(N, H, W, 2)).astype(np.float32)   # exercises wrap on both sides
    uv[0, 0, 0] = (0.0, 0.0)                                             # background texel corner (App. A.3)
    uv[0, 0, 1] = (1.0 - 1e-7, 0.5 / Ht)
    out = G.texture_linear_fwd(tex, uv)
    tt = torch.tensor(tex).double().requires_grad_(True)
    tu = torch.tensor(uv).double().requires_grad_(True)
    ref = TR.texture_linear(tt, tu)
    assert np.abs(out - ref.detach().numpy()).max() < 1e-5
    # uv = (0,0) samples the wrapped corner average
    np.testing.assert_allclose(out[0, 0, 0], 0.25 * (tex[0, 0, 0] + tex[0, 0, -1] + tex[0, -1, 0] + tex[0, -1, -1]), rtol=1e-6)
    dy = _rand_like(out, 5)
    gt, guv = G.texture_linear_bwd(tex, uv, dy)
    (ref * torch.tensor(dy).double()).sum().backward()
    assert np.abs(gt - tt.grad.numpy()).max() / np.abs(tt.grad.numpy()).max() < 1e-4
    # fp32 rounding of the fractional position changes d/duv by ~1e-5 relative; compare at 1e-3 of the max
    assert np.abs(guv - tu.grad.numpy()).max() / np.abs(tu.grad.numpy()).max() < 1e-3


def test_topology(tiny_rig):
    opp = G.topology_build(tiny_rig.pos_idx)
    assert opp.shape == (tiny_rig.T, 3) and (opp >= 0).all()       # closed mesh: every edge has a wing
    tri = tiny_rig.pos_idx
    # brute force on a few triangles
    for t in (0, 7, 1234):
        for e in range(3):
            va, vb = tri[t, (e + 1) % 3], tri[t, (e + 2) % 3]
            cands = [s for s in range(tri.shape[0]) if s != t and va in tri[s] and vb in tri[s]]
            assert len(cands) == 1
            third = [x for x in tri[cands[0]] if x != va and x != vb][0]
            assert opp[t, e] == third
    # open mesh: boundary edges have no opposite vertex
    opp2 = G.topology_build(np.array([[0, 1, 2], [0, 2, 3]], np.int32))
    assert opp2.tolist() == [[-1, 3, -1], [-1, -1, 1]]


def test_antialias_matches_autograd(small_rig3):
    r = small_rig3
    pc = clip_positions(r, w=np.linspace(0, 0.4, 8))
    H, W = 152, 200
    rast, _, _ = G.rasterize_fwd(pc, r.pos_idx, (H, W))
    rng = np.random.default_rng(6)
    col = rng.random(rast.shape[:3] + (3,)).astype(np.float32)
    opp = G.topology_build(r.pos_idx)
    out = G.antialias_fwd(col, rast, pc, r.pos_idx, opp)
    changed = (np.abs(out - col).max(axis=-1) > 0)
    assert 50 < changed.sum() < 0.1 * changed.size                 # only silhouette pixels are touched
    tc = torch.tensor(col).double().requires_grad_(True)
    tp = torch.tensor(pc).double().requires_grad_(True)
    ref = TR.antialias(tc, torch.tensor(rast).double(), tp, torch.tensor(r.pos_idx), torch.tensor(opp))
    # the fp64 restatement can flip a marginal silhouette decision; allow a handful of pixels
    bad = np.abs(out - ref.detach().numpy()).max(axis=-1) > 1e-4
    assert bad.sum() <= 4, bad.sum()
    dy = _rand_like(out, 7)
    dy[bad] = 0
    gc, gp = G.antialias_bwd(col, rast, pc, r.pos_idx, dy, opp)
    (ref * torch.tensor(dy).double()).sum().backward()
    assert np.abs(gc - tc.grad.numpy()).max() < 1e-3
    ref_gp = tp.grad.numpy()
    # position gradient: golden uses the 1e-3 px regulariser on 1/dy (App. A.4), autograd the exact derivative
    denom = np.abs(ref_gp).max()
    assert denom > 0
    assert np.abs(gp - ref_gp).max() / denom < 2e-2
    assert np.abs(gp[..., 2]).max() == 0


def test_torch_stages(tiny_rig):
    r = tiny_rig
    w = torch.linspace(0, 1, r.B)
    v = G.blend(torch.tensor(r.v_base), torch.tensor(r.D), w)
    np.testing.assert_allclose(v.numpy(), r.v_base + r.D @ w.numpy(), rtol=1e-5, atol=1e-5)
    ident = G.mvp_chain(torch.tensor(r.P[0]), torch.tensor(r.A[0]), torch.zeros(3), torch.tensor([0., 0, 0, 1]))
    np.testing.assert_allclose(ident.numpy(), r.P[0] @ r.A[0], rtol=1e-6)
    img = G.render(ident, v.reshape(-1, 3), torch.tensor(r.pos_idx), (128, 128), vcol=torch.tensor(r.vcol),
                   use_antialias=False)
    assert img.shape == (128, 128, 3)
    assert torch.allclose(img[0, 0], torch.tensor(G.BG))            # corner pixel is background 45/255
    ref = torch.clamp(img * 255, 0, 140)
    assert float(G.image_loss(ref, img)) < 1e-6


def test_mesh_topology_and_regularisers(small_rig3):
    """Static topology of a closed genus-0 mesh (E = 3V - 6, every edge has two faces) and the torch restatement of the
    pytorch3d mesh terms (fit.py:578-582): known values on simple shapes + float64 gradcheck."""
    import torch
    from fpc_diffrend_b200 import topology
    from oracle import golden as G
    rig = small_rig3
    tp = topology.build_topology(rig.pos_idx, rig.V)
    assert tp.E == 3 * rig.V - 6 and tp.E2 == tp.E and tp.nbr_off[-1] == 2 * tp.E
    assert (tp.edges[:, 0] < tp.edges[:, 1]).all()
    # neighbour lists = the reference's own builder (data.py:44-66) without its padding
    ref_nb = [set() for _ in range(rig.V)]
    for a, b, c in rig.pos_idx:
        ref_nb[a].update((b, c)); ref_nb[b].update((a, c)); ref_nb[c].update((a, b))
    for i in (0, 1, rig.V // 2, rig.V - 1):
        assert set(tp.nbr_idx[tp.nbr_off[i]:tp.nbr_off[i + 1]].tolist()) == ref_nb[i]
    # a flat regular patch has zero normal-consistency loss; a unit square's diagonal edge pair likewise
    quad = torch.tensor([[0., 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=torch.float64)
    tq = topology.build_topology(np.array([[0, 1, 2], [0, 2, 3]]), 4)
    assert tq.E == 5 and tq.E2 == 1
    assert abs(float(G.mesh_normal_consistency(quad, torch.tensor(tq.edge_quads).long()))) < 1e-12
    folded = quad.clone(); folded[3, 2] = 1.0; folded[3, 1] = 0.0      # fold the second face by 90 degrees about... the diagonal changes
    assert float(G.mesh_normal_consistency(folded, torch.tensor(tq.edge_quads).long())) > 0.1
    # edge loss: all edges of the unit square + diagonal against their own mean
    el = float(G.mesh_edge_loss(quad, torch.tensor(tq.edges).long(), 1.0))
    assert abs(el - (2 ** 0.5 - 1) ** 2 / 5) < 1e-12
    # uniform Laplacian of a regular polygon fan centre is 0; of a displaced centre it is the displacement / V
    n = 6
    ring = torch.tensor([[np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), 0.0] for k in range(n)], dtype=torch.float64)
    fan = torch.cat([torch.zeros(1, 3, dtype=torch.float64), ring])
    tf = topology.build_topology(np.array([[0, 1 + k, 1 + (k + 1) % n] for k in range(n)]), n + 1)
    lv0 = G.mesh_laplacian_uniform(fan, torch.tensor(tf.edges).long())
    fan2 = fan.clone(); fan2[0, 2] = 0.5
    lv1 = G.mesh_laplacian_uniform(fan2, torch.tensor(tf.edges).long())
    rim = (fan[[2, 6, 0]].mean(0) - fan[1]).norm()          # every rim vertex: mean of its 3 neighbours minus itself
    assert abs(float(lv0) - float(n * rim / (n + 1))) < 1e-12
    assert float(lv1) > float(lv0)
    # gradcheck of the combined term on the rig (float64)
    v = torch.tensor(rig.v_base, dtype=torch.float64).reshape(-1, 3)[:rig.V].clone().requires_grad_(True)
    e, q = torch.tensor(tp.edges).long(), torch.tensor(tp.edge_quads).long()
    f = lambda x: G.mesh_regularisers(x, e, q, 5000.0, 70.0, 0.05, 400.0)[0]
    tot = f(v); tot.backward()
    rng = np.random.default_rng(0)
    d = torch.tensor(rng.normal(size=v.shape))
    h = 1e-6
    fd = (f(v.detach() + h * d) - f(v.detach() - h * d)) / (2 * h)
    assert abs(float(fd) - float((v.grad * d).sum())) <= 1e-6 * abs(float(fd)) + 1e-8


def _push_towards_camera(pc, s, zn=0.01, zf=200.0):
    """Clip-space positions of the same geometry moved s units towards the camera (standard projection of camera.py:27-41:
    w = -z_eye, z_clip = A z_eye + B): the near plane then cuts through the mesh and part of it lies behind the camera."""
    A = -(zf + zn) / (zf - zn)
    out = pc.copy()
    out[..., 2] = pc[..., 2] + np.float32(A * s)
    out[..., 3] = pc[..., 3] - np.float32(s)
    return out


def _push_depth(pc, tri):
    """Push distance that puts the surface point seen at the image centre of view 0 exactly on the near plane (w = zn), so
    that the clip line of the triangle around it runs through the middle of the image."""
    rast, _, _ = G.rasterize_fwd(pc[:1], tri, (65, 65), with_db=False)
    u, v, _, idp1 = rast[0, 32, 32]
    assert idp1 > 0
    w3 = pc[0, tri[int(idp1) - 1], 3].astype(np.float64)
    return float(u * w3[0] + v * w3[1] + (1 - u - v) * w3[2]) - 0.01


def test_raster_near_plane_clipper(small_rig3):
    """Triangles with a vertex at w <= 0 are clipped against the near plane, the pieces keep the parent id and are shaded
    with the parent's vertices.  Properties: (1) perspective-correct barycentrics reproduce the pixel
    centre for every covered pixel, including pixels of clipped triangles; (2) the coverage of a clipped triangle equals
    the union of its pre-clipped pieces rasterized as ordinary triangles."""
    rig = small_rig3
    H, W = 96, 128
    pc = clip_positions(rig)
    near = _push_towards_camera(pc, _push_depth(pc, rig.pos_idx))   # the near plane cuts the front of the head at the image centre
    assert (near[..., 3] <= 0).any() and (near[..., 3] > 0).any()
    rast, _, _ = G.rasterize_fwd(near, rig.pos_idx, (H, W), with_db=False)
    ids = rast[..., 3].astype(np.int64) - 1
    clipped_tri = (near[:, rig.pos_idx, 3] <= 0).any(axis=2) & (near[:, rig.pos_idx, 3] > 0).any(axis=2)       # [C,T]
    seen_clipped = 0
    for n in range(near.shape[0]):
        ys, xs = np.nonzero(ids[n] >= 0)
        t = ids[n][ys, xs]
        seen_clipped += int(clipped_tri[n][t].sum())
        p = near[n][rig.pos_idx[t]].astype(np.float64)             # [K,3,4]
        u, v = rast[n, ys, xs, 0].astype(np.float64), rast[n, ys, xs, 1].astype(np.float64)
        b = np.stack([u, v, 1 - u - v], axis=1)[..., None]
        q = (b * p).sum(axis=1)
        fx, fy = (2 * xs + 1) / W - 1, (2 * ys + 1) / H - 1
        inner = (u > 1e-6) & (v > 1e-6) & (u + v < 1 - 1e-6)         # the clamps bend the relation on the boundary
        assert np.abs(q[inner, 0] / q[inner, 3] - fx[inner]).max() < 2e-4
        assert np.abs(q[inner, 1] / q[inner, 3] - fy[inner]).max() < 2e-4
        assert (q[inner, 3] > 0).all()                               # only the part in front of the camera is drawn
    assert seen_clipped > 50, seen_clipped                            # pixels of clipped triangles do show up
    # (2) one triangle, clipped by hand in float32 with the golden's formula
    v = np.array([[[-0.5, -0.5, 0.2, 1.0], [0.5, -0.5, 0.2, 1.0], [0.1, 0.9, -1.5, -0.5]]], np.float32)
    r1, _, _ = G.rasterize_fwd(v, np.array([[0, 1, 2]], np.int32), (64, 64), with_db=False)
    d = v[0, :, 2] + v[0, :, 3]
    lerp = lambda a, da, b, db: (a + (np.float32(da) / (np.float32(da) - np.float32(db))) * (b - a)).astype(np.float32)
    c12, c20 = lerp(v[0, 1], d[1], v[0, 2], d[2]), lerp(v[0, 0], d[0], v[0, 2], d[2])
    pieces = np.stack([v[0, 0], v[0, 1], c12, c20])[None]
    r2, _, _ = G.rasterize_fwd(pieces, np.array([[0, 1, 2], [0, 2, 3]], np.int32), (64, 64), with_db=False)
    assert (r1[..., 3] > 0).sum() > 100
    assert np.array_equal(r1[..., 3] > 0, r2[..., 3] > 0)
    assert np.abs(r1[..., 2] - r2[..., 2]).max() < 1e-5


def _blend_kat():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'blend_kat.json')) as f:
        return json.load(f)


def test_blend_modes_match_reference_kat():
    """The oracle's blend_prior / blend_free / blend_combined against vectors produced by the reference's own
    functions (tests/golden/make_blend_kat.py executes fit.py:47-129 unchanged): values and autograd gradients."""
    k = _blend_kat()
    T = lambda name: torch.tensor(k[name], dtype=torch.float32)
    v_base, D, dy = T('v_base'), T('D'), T('dy')
    for frame, rec in k['frames'].items():
        e = torch.zeros(k['F'])
        e[int(frame)] = 1.0
        for mode in ('prior', 'free', 'combined'):
            P = {n: T(n).requires_grad_(True) for n in ('M1', 'M2', 'm1', 'm2', 'm3')}
            if mode == 'prior':
                v = G.blend_prior(v_base, D, P['M1'], P['M2'], e)
            elif mode == 'free':
                v = G.blend_free(v_base, P['m1'], P['m2'], P['m3'], e)
            else:
                v = G.blend_combined(v_base, D, P['M1'], P['M2'], P['m1'], P['m2'], P['m3'], e, learned_coefficient=0.5)
            ref = torch.tensor(rec[mode]['vtx_pos'])
            assert (v.detach() - ref).abs().max() <= 1e-5 * ref.abs().max()
            (v * dy).sum().backward()
            for n, g in rec[mode]['grad'].items():
                if g is None:
                    assert P[n].grad is None
                else:
                    g = torch.tensor(g)
                    assert (P[n].grad - g).abs().max() <= 1e-5 * max(float(g.abs().max()), 1e-6), (mode, n)
            # the north-star form V = base + D w with w = M2 M1 e_f is the same map
            if mode == 'prior':
                w = torch.tensor(k['M2']) @ torch.tensor(k['M1'])[:, int(frame)]
                assert (G.blend(v_base, D, w) - ref).abs().max() <= 1e-5 * ref.abs().max()


def test_mip_path_oracle_is_self_consistent(small_rig3):
    """The torch restatement of the mip path: (a) barycentric_diffs equals the golden rasterizer's rast_db, (b) they are the
    pixel-to-pixel differences of (u, v) inside a triangle, (c) texture_mip degenerates to bilinear lookups of the right
    level at the clamps and interpolates in between, (d) the mip chain preserves the mean."""
    rig, H, W = small_rig3, 152, 200
    pc = clip_positions(rig)
    rast, db, _ = G.rasterize_fwd(pc, rig.pos_idx, (H, W))
    tid = torch.tensor(rast[..., 3]).long() - 1
    p64 = torch.tensor(pc, dtype=torch.float64)
    d64 = TR.barycentric_diffs(p64, torch.tensor(rig.pos_idx), tid, H, W).numpy()
    assert np.abs(d64 - db).max() <= 1e-4 * np.abs(db).max()        # fp32 golden vs fp64: sliver triangles cancel digits
    # (b) u is a ratio of affine functions: across one pixel step inside the same triangle the mean of the two end-point
    # derivatives matches the finite difference to second order
    same = (rast[:, :, 1:, 3] == rast[:, :, :-1, 3]) & (rast[:, :, 1:, 3] > 0)
    fd = rast[:, :, 1:, 0] - rast[:, :, :-1, 0]
    mid = 0.5 * (db[:, :, 1:, 0] + db[:, :, :-1, 0])
    inner = same & (rast[:, :, 1:, 0] > 0) & (rast[:, :, 1:, 0] < 1) & (rast[:, :, :-1, 0] > 0) & (rast[:, :, :-1, 0] < 1)
    assert inner.sum() > 1000 and np.abs(fd - mid)[inner].max() < 1e-3 * np.abs(fd[inner]).max()
    # (c), (d)
    g = torch.Generator().manual_seed(3)
    tex = torch.rand(1, 16, 32, 2, generator=g, dtype=torch.float64)
    uv = torch.rand(1, 9, 11, 2, generator=g, dtype=torch.float64) * 2 - 0.5
    levels = TR.texture_construct_mip(tex, 3)
    assert len(levels) == 4 and levels[3].shape == (1, 2, 4, 2)
    for l in levels:
        assert torch.allclose(l.mean(dim=(1, 2)), tex.mean(dim=(1, 2)))
    zeros = torch.zeros(1, 9, 11)
    for lev, expect in ((-2.0, TR.texture_linear(levels[0], uv)), (1.0, TR.texture_linear(levels[1], uv)),
                        (7.0, TR.texture_linear(levels[3], uv)),
                        (1.25, 0.75 * TR.texture_linear(levels[1], uv) + 0.25 * TR.texture_linear(levels[2], uv))):
        out = TR.texture_mip(tex, uv, None, zeros + lev, max_mip_level=3)
        assert torch.allclose(out, expect, atol=1e-12)
    # an isotropic footprint of 2^k texels selects level k
    da = torch.zeros(1, 9, 11, 4, dtype=torch.float64)
    da[..., 0] = 4.0 / 32
    da[..., 3] = 4.0 / 16
    assert torch.allclose(TR.mip_level(da, 16, 32), torch.full((1, 9, 11), 2.0, dtype=torch.float64))
