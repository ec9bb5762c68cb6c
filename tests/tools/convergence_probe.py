"""Probe of the fit dynamics (developer tool): loss / vertex-error trajectories for a few settings and a finite-difference check
of d loss / d w along the gradient direction at a mid-fit state.   python tests/tools/convergence_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fpc_diffrend_b200 import rig as rigmod  # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference  # noqa: E402

H = W = 192
F = 2
rig = rigmod.make_rig(n_vertices=1500, n_shapes=12, n_cams=3, width=W, height=H, tex_size=128, seed=5)
_, t_true, q_true = rigmod.make_targets(F, rig.B, seed=7)
w_true = np.random.default_rng(7).uniform(0.3, 0.9, size=(F, rig.B)).astype(np.float32)
v_true = torch.tensor(rig.v_base)[None] + torch.tensor(w_true) @ torch.tensor(rig.D).t()
q_id = q_true * 0 + np.array([0, 0, 0, 1], np.float32)
for shading, aa, lr, iters, b2 in (('vcol', True, 2e-2, 400, 0.999), ('vcol', True, 2e-2, 400, 0.9), ('texture', True, 5e-3, 800, 0.9)):
    cfg = FitConfig(resolution=(H, W), shading=shading, antialias=aa, lr_base=lr, lr_ramp=0.1, max_iter=iters, optimize_pose=False, beta2=b2)
    ref = synthesize_reference(rig, w_true, 0.0 * t_true, q_id, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    err0 = float((s.result_vertices().cpu() - v_true).abs().mean())
    s.iteration()
    torch.cuda.synchronize()
    loss0 = float(s.loss)
    s.capture()
    hist = []
    for i in range(iters):
        s.replay()
        if i % (iters // 8) == 0:
            hist.append(round(float(s.loss), 3))
        if i == iters // 3:
            # finite differences along the (negative) gradient at this state; forward()/backward() do not touch the parameters
            s.forward(); s.backward(); torch.cuda.synchronize()
            w0, g = s.w.clone(), s.d_w.clone()
            d = g / g.norm()
            out = []
            for eps in (1e-2, 3e-3, 1e-3):
                s.w.copy_(w0 + eps * d); s.forward(); torch.cuda.synchronize(); lp = float(s.loss)
                s.w.copy_(w0 - eps * d); s.forward(); torch.cuda.synchronize(); lm = float(s.loss)
                out.append('eps %g: fd %.4f' % (eps, (lp - lm) / (2 * eps)))
            s.w.copy_(w0)
            print('   directional derivative: analytic %.4f | %s' % (float((g * d).sum()), ' | '.join(out)))
    torch.cuda.synchronize()
    err1 = float((s.result_vertices().cpu() - v_true).abs().mean())
    print(shading, 'aa' if aa else 'no-aa', 'lr', lr, 'beta2', b2, 'loss %.3f -> %.3f' % (loss0, float(s.loss)), 'vertex err %.4f -> %.4f' % (err0, err1),
          'activation err -> %.3f' % float((s.w.cpu() - torch.tensor(w_true)).abs().mean()), hist)
