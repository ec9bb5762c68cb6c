"""oracle/nvdiffrast_arm.py — the reference's OWN GPU path (torch + nvdiffrast), when one is installed.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and by bench.py's reference / parity
legs, never by fpc_diffrend_b200/.

The north-star defines correctness "against the reference's own nvdiffrast CUDA path" (/root/reference/src/torch/
fit.py:13,151-160,484).  nvdiffrast is an un-vendored, un-pinned dependency that cannot be installed offline here
(SURVEY §8(c)), so everything in this file is gated on `probe()`: it looks for an importable `nvdiffrast.torch` (site
packages, or a copy under baseline/_ref/) and returns None otherwise — callers then report
`"reference_gpu": "unavailable"` instead of a guess.  When it IS importable:

  * `fit_step_factory()` restates one iteration of the reference's loop (fit.py:524-642) for ALL views of one frame
    under the north-star parameterisation, calling nvdiffrast exactly as fit.render() does (fit.py:134-162) but with
    `RasterizeCudaContext` (the north-star's "CUDA path") — this is what `bench.py --impl reference` times;
  * `compare_render()` runs the same tensors through nvdiffrast and through this repo's drop-in ops and reports the
    figures the north-star states tolerances for (tri_id mismatches outside depth ties, max |d| of rast / colour,
    relative gradient error).
"""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BG = 45.0 / 255.0      # fit.py:161


def probe():
    """`nvdiffrast.torch` module or None.  Never raises (a broken install counts as absent; the reason is kept in
    probe.reason)."""
    ref = os.path.join(_ROOT, 'baseline', '_ref')
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import nvdiffrast.torch as dr       # noqa: the reference's own import, fit.py:13
        probe.reason = 'nvdiffrast %s' % getattr(sys.modules.get('nvdiffrast'), '__version__', '?')
        return dr
    except Exception as e:                  # ImportError, or a plugin that fails to build / load offline
        probe.reason = '%s: %s' % (type(e).__name__, str(e).splitlines()[0] if str(e) else '')
        return None


probe.reason = 'not probed'


def _context(dr, device):
    """RasterizeCudaContext (nvdiffrast >= 0.3.0), else the GL context the reference ships with (fit.py:484)."""
    if hasattr(dr, 'RasterizeCudaContext'):
        return dr.RasterizeCudaContext(device=device), 'RasterizeCudaContext'
    return dr.RasterizeGLContext(device=device), 'RasterizeGLContext'


def quat_to_rotmat(q):
    """roma.unitquat_to_rotmat (XYZW), fit.py:548,550."""
    x, y, z, w = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)]),
        torch.stack([2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)]),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w]),
    ])


def rigid(tvec, rotmat):
    """camera.py:128-132."""
    rt = torch.cat((rotmat, tvec.reshape(3, 1)), 1)
    br = torch.tensor([[0, 0, 0, 1]], dtype=rt.dtype, device=rt.device)
    return torch.cat((rt, br), 0)


def transform_clip(mvp, pos):
    """camera.py:11-23."""
    posw = torch.cat([pos, torch.ones([pos.shape[0], 1], dtype=pos.dtype, device=pos.device)], axis=1)
    return torch.matmul(posw, mvp.t())[None, ...]


def render(dr, ctx, mvp, pos, pos_idx, resolution, uv=None, uv_idx=None, tex=None, vcol=None, antialias=True, want_rast=False):
    """fit.py:134-162 (bilinear branch), or the vertex-colour variant of BASELINE config 2."""
    pos_clip = transform_clip(mvp, pos)
    rast_out, _ = dr.rasterize(ctx, pos_clip, pos_idx, resolution=(resolution[0], resolution[1]))
    if vcol is not None:
        colour, _ = dr.interpolate(vcol[None, ...], rast_out, pos_idx)
    else:
        texc, _ = dr.interpolate(uv[None, ...], rast_out, uv_idx)
        colour = dr.texture(tex[None, ...], texc, filter_mode='linear')
    if antialias:
        colour = dr.antialias(colour, rast_out, pos_clip, pos_idx)
    colour = torch.where(rast_out[..., 3:] > 0, colour, torch.tensor(BG, device=colour.device))
    return (colour[0], rast_out, pos_clip) if want_rast else colour[0]


def fit_step_factory(dr, rig, resolution, shading, antialias, ref, device='cuda', lr=(1e-3, 1e-5, 1e-5)):
    """One frame, all views.  ref [C,H,W,Ch] float32 on `device` (0..255 scale).  Returns (step, params): step() runs
    forward + backward + Adam + quaternion renorm and returns the loss tensor (no host sync)."""
    ctx, ctx_name = _context(dr, device)
    f32 = dict(dtype=torch.float32, device=device)
    base, D = torch.tensor(rig.v_base, **f32), torch.tensor(rig.D, **f32)
    P, A = torch.tensor(rig.P, **f32), torch.tensor(rig.A, **f32)
    tri = torch.tensor(rig.pos_idx, dtype=torch.int32, device=device)
    kw = {}
    if shading == 'vcol':
        kw['vcol'] = torch.tensor(rig.vcol, **f32)
    else:
        kw.update(uv=torch.tensor(rig.uv, **f32), uv_idx=torch.tensor(rig.uv_idx, dtype=torch.int32, device=device),
                  tex=torch.tensor(rig.tex, **f32))
    C = P.shape[0]
    w = torch.zeros(D.shape[1], requires_grad=True, **f32)
    t = torch.zeros(3, requires_grad=True, **f32)
    q = torch.tensor([0., 0, 0, 1], requires_grad=True, **f32)
    opt = torch.optim.Adam([{'params': w, 'lr': lr[0]}, {'params': t, 'lr': lr[1]}, {'params': q, 'lr': lr[2]}])

    def step():
        verts = torch.add(base, torch.matmul(D, w)).reshape(-1, 3)          # fit.py:103-129
        T_frame = rigid(t, quat_to_rotmat(q))
        loss = 0.0
        for c in range(C):                                                  # the reference renders one view per iteration
            mvp = torch.matmul(P[c], torch.matmul(T_frame, A[c]))           # fit.py:546-553
            img = render(dr, ctx, mvp, verts, tri, resolution, antialias=antialias, **kw)
            loss = loss + torch.mean((ref[c] - img * 255) ** 2)             # fit.py:579
        loss = loss / C
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            q /= torch.sum(q ** 2) ** 0.5                                   # fit.py:616-618
        return loss.detach()

    step.context = ctx_name
    return step, (w, t, q)


def compare_render(dr, ours, pos_clip, tri, resolution, attr, attr_idx, tex=None, antialias=True, seed=0):
    """The same tensors through nvdiffrast and through `ours` (fpc_diffrend_b200.ops).  pos_clip [N,V,4], tri [T,3] i32,
    attr [1,Va,A] (vertex colours or uv), all on one CUDA device.  Returns a dict of parity figures."""
    dev = pos_clip.device
    out = {}
    res = {}
    for name, mod in (('ref', dr), ('ours', ours)):
        ctx = _context(mod, dev)[0] if name == 'ref' else mod.RasterizeCudaContext(device=dev)
        pos = pos_clip.clone().requires_grad_(True)
        rast, _ = mod.rasterize(ctx, pos, tri, resolution=resolution)
        col, _ = mod.interpolate(attr, rast, attr_idx)
        if tex is not None:
            col = mod.texture(tex, col, filter_mode='linear')
        if antialias:
            col = mod.antialias(col, rast, pos, tri)
        g = torch.Generator(device='cpu').manual_seed(seed)
        wgt = torch.randn(col.shape, generator=g).to(dev)
        (col * wgt).sum().backward()
        res[name] = (rast.detach(), col.detach(), pos.grad.detach())
    (r0, c0, g0), (r1, c1, g1) = res['ref'], res['ours']
    id0, id1 = r0[..., 3], r1[..., 3]
    diff = id0 != id1
    # a depth tie = the two candidates' z/w agree to within a few fp32 ulps at that pixel
    near = (r0[..., 2] - r1[..., 2]).abs() <= 4e-7 * r0[..., 2].abs().clamp_min(1e-3)
    out['pixels'] = int(id0.numel())
    out['tri_id_mismatch'] = int(diff.sum())
    out['tri_id_mismatch_outside_depth_ties'] = int((diff & ~near).sum())
    same = ~diff
    out['rast_uvz_max_abs'] = float((r0[..., :3] - r1[..., :3])[same].abs().max()) if bool(same.any()) else 0.0
    out['colour_max_abs'] = float((c0 - c1)[same].abs().max()) if bool(same.any()) else 0.0
    out['grad_pos_rel'] = float((g0 - g1).abs().max() / g0.abs().max().clamp_min(1e-30))
    return out
