"""Small multi-mode exercise of every kernel family (fused with / without antialias, texture gradient, camera corrections,
op-level chain incl. the mip path, free / combined modes, L2 terms, mesh regularisers, tensor-core blend, band split) on a ragged
resolution: a crash / NaN check, and the target to run under `compute-sanitizer --tool memcheck|racecheck` where the tool is
available.   python scripts/sanitize_target.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpc_diffrend_b200 import rig as rigmod  # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference  # noqa: E402
import fpc_diffrend_b200.ops as dr  # noqa: E402

H, W = 72, 104          # ragged: not a multiple of the 32-px bin
rig = rigmod.make_rig(n_vertices=300, n_shapes=8, n_cams=2, width=W, height=H, tex_size=32, seed=1)
F = 2
w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
for kw in (dict(shading='vcol'), dict(shading='texture', antialias=True, optimize_texture=True, optimize_cam_pose=True),
           dict(shading='texture', antialias=True, fused=False), dict(shading='vcol', mode='free'),
           dict(shading='texture', antialias=True, mode='combined', corrective_start=0, regularize_correctives=True, weight_laplacian=10.0),
           dict(shading='vcol', regularize_prior=True, fused_geometry=False, tc_blend=True, weight_meshedge=1.0, weight_normalconsistency=1.0),
           dict(shading='texture', antialias=True, cam_slice=(0, 2), cam_band=(1, 2))):
    cfg = FitConfig(resolution=(H, W), lr_base=1e-2, max_iter=10, **kw)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    l0 = None
    for _ in range(3):
        s.iteration()
        l0 = float(s.loss) if l0 is None else l0
    torch.cuda.synchronize()
    assert torch.isfinite(s.params).all() and torch.isfinite(s.loss).all(), kw
    print('ok', kw, l0, '->', float(s.loss))

ctx = dr.RasterizeGLContext(device='cuda')
sess = FitSession(rig, 1, FitConfig(resolution=(H, W)))
sess.forward(with_loss=False)
pos = sess.pos_clip.clone().requires_grad_(True)
tri = sess.pos_idx
tex = torch.tensor(rig.tex, device='cuda')[None].requires_grad_(True)
rast, db = dr.rasterize(ctx, pos, tri, resolution=(H, W))
uv, uvd = dr.interpolate(torch.tensor(rig.uv, device='cuda')[None], rast, torch.tensor(rig.uv_idx, device='cuda'), rast_db=db, diff_attrs='all')
col = dr.texture(tex, uv, uvd, filter_mode='linear-mipmap-linear', max_mip_level=3)
col = dr.antialias(col, rast, pos, tri)
col.sum().backward()
torch.cuda.synchronize()
assert torch.isfinite(pos.grad).all() and torch.isfinite(tex.grad).all()
print('ok mip chain', float(pos.grad.abs().sum()), float(tex.grad.abs().sum()))
