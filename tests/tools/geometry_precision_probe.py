"""Precision of the transpose chain (d pos_clip -> d mvp -> pose gradients; d pos_clip -> d verts -> d w) of the fused geometry
kernels and of the separate kernels against FLOAT64 evaluated from the session's own d pos_clip (developer tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fpc_diffrend_b200 import rig as rigmod  # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference  # noqa: E402
from oracle import golden as G  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), 1e-30)


V, B, Cc, H, W, F = 2000, 4, 2, 149, 165, 3
for seed in (0, 1, 2):
    rng = np.random.default_rng(seed)
    rig = rigmod.make_rig(n_vertices=V, n_shapes=B, n_cams=Cc, width=W, height=H, tex_size=32, seed=int(rng.integers(1 << 30)))
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=int(rng.integers(1 << 30)))
    base = dict(resolution=(H, W), shading='vcol', antialias=False, loss='l1', optimize_cam_pose=True)
    ref = synthesize_reference(rig, w_true, 0.3 * t_true, q_true, FitConfig(**base))
    w0 = (0.3 * rng.random((F, rig.B))).astype(np.float32)
    t0 = (0.3 * rng.normal(size=(F, 3))).astype(np.float32)
    tc = (0.1 * rng.normal(size=(Cc, 3))).astype(np.float32)
    res = {}
    for name, kw in (('fused geometry', dict(fused_geometry=True)), ('separate kernels', dict(fused_geometry=False, tc_blend=False))):
        s = FitSession(rig, F, FitConfig(**base, **kw))
        s.set_reference(ref)
        s.set_parameters(w=w0, t=t0)
        s.t_cam.copy_(torch.tensor(tc))
        s.forward(); s.backward()
        torch.cuda.synchronize()
        # float64 transposes from THIS session's d pos_clip
        g = s.g_pos.cpu().double().reshape(F, Cc, V, 4)
        verts = s.verts.cpu().double().reshape(F, V, 3)
        vh = torch.cat([verts, torch.ones(F, V, 1, dtype=torch.float64)], dim=2)
        d_mvp = torch.einsum('fcvi,fvj->fcij', g, vh)                                      # mvp[i][j]: clip_i = sum_j mvp_ij vh_j
        mvp = s.mvp.cpu().double().reshape(F, Cc, 4, 4)
        d_verts = torch.einsum('fcvi,fcij->fvj', g, mvp)[..., :3]
        d_w = torch.einsum('fr,rb->fb', d_verts.reshape(F, V * 3), torch.tensor(rig.D).double())
        tf = s.t.cpu().double().clone().requires_grad_(True)
        qf = s.q.cpu().double().clone().requires_grad_(True)
        tct = torch.tensor(tc).double().requires_grad_(True)
        qct = s.q_cam.cpu().double().clone().requires_grad_(True)
        tot = 0.0
        for f in range(F):
            for c in range(Cc):
                m = G.mvp_chain(torch.tensor(rig.P[c]).double(), torch.tensor(rig.A[c]).double(), tf[f], qf[f], tct[c], qct[c])
                tot = tot + (m * d_mvp[f, c]).sum()
        tot.backward()
        res[name] = (s.d_w.cpu().clone(), s.d_t.cpu().clone(), s.d_t_cam.cpu().clone())
        print('seed %d %-17s vs float64: d_w %.1e  d_t %.1e  d_q %.1e  d_t_cam %.1e  d_q_cam %.1e   (max |d_t| %.3g, sum |terms| / |sum| of d_mvp ~ %.0f)' %
              (seed, name, rel(s.d_w.cpu(), d_w), rel(s.d_t.cpu(), tf.grad), rel(s.d_q.cpu(), qf.grad), rel(s.d_t_cam.cpu(), tct.grad),
               rel(s.d_q_cam.cpu(), qct.grad), float(tf.grad.abs().max()),
               float(torch.einsum('fcvi,fvj->fcij', g.abs(), vh.abs()).max() / d_mvp.abs().max())))
    a, b = res['fused geometry'], res['separate kernels']
    print('        fused vs separate: d_w %.1e d_t %.1e d_t_cam %.1e' % (rel(a[0], b[0]), rel(a[1], b[1]), rel(a[2], b[2])))
