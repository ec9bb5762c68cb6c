"""Diagnostic: per-iteration agreement of the GPU fit with the CPU oracle at BASELINE config 2 size.
usage: python tests/tools/fit_traj_probe.py [iters] [eps_fraction]
Prints, per iteration, (a) rel. error of the activation gradient with the oracle evaluated at the GPU's own parameters
(pure kernel parity, no trajectory effects), (b) rel. difference of the free-running trajectories."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fpc_diffrend_b200 import rig as rigmod                                     # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference     # noqa: E402
from oracle import golden as G                                                    # noqa: E402
from dataclasses import replace                                                   # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
epsf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
H = W = 1024
rig = rigmod.make_rig(n_vertices=20000, n_shapes=200, n_cams=9, width=W, height=H, tex_size=64, seed=0)
cfg0 = FitConfig(resolution=(H, W), shading='vcol', antialias=False)
w_true, t_true, q_true = rigmod.make_targets(1, rig.B, seed=1)
ref = synthesize_reference(rig, w_true, t_true, q_true, cfg0)
ref_cpu = ref.cpu()
tri = torch.tensor(rig.pos_idx)
base, D, vcol = torch.tensor(rig.v_base), torch.tensor(rig.D), torch.tensor(rig.vcol)
Ps, As = torch.tensor(rig.P), torch.tensor(rig.A)


def oracle_grad(w, t, q):
    w = w.clone().requires_grad_(True); t = t.clone().requires_grad_(True); q = q.clone().requires_grad_(True)
    verts = G.blend(base, D, w[0]).reshape(-1, 3)
    pcs = torch.cat([G.transform_clip(G.mvp_chain(Ps[c], As[c], t[0], q[0]), verts) for c in range(9)])
    rast, _ = G.rasterize(pcs, tri, (H, W))
    col = G.interpolate(vcol[None], rast, tri)
    img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))
    loss = sum(G.image_loss(ref_cpu[0, c], img[c]) for c in range(9)) / 9
    loss.backward()
    return float(loss.detach()), w.grad, t.grad, q.grad


def staged(sess):
    """Oracle render + loss + d pos_clip fed the GPU's OWN pos_clip bits (no geometry rounding in between), and the oracle's
    sensitivity to a 1-ulp perturbation of those positions."""
    pc = sess.pos_clip.cpu().clone()
    out = []
    for pos in (pc, torch.nextafter(pc, torch.full_like(pc, float('inf')))):
        p = pos.clone().requires_grad_(True)
        rast, _ = G.rasterize(p, tri, (H, W))
        col = G.interpolate(vcol[None], rast, tri)
        img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))
        loss = sum(G.image_loss(ref_cpu[0, c], img[c]) for c in range(9)) / 9
        loss.backward()
        out.append((float(loss.detach()), p.grad.clone()))
    return out


rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-30))
w = torch.zeros(1, rig.B); t = torch.zeros(1, 3); q = torch.tensor([[0., 0, 0, 1]])
_, g0, _, _ = oracle_grad(w, t, q)
eps = epsf * float(g0.abs().max()) if epsf > 0 else 1e-8
print('max |g_w| at the start %.4g, eps %.4g' % (float(g0.abs().max()), eps))
cfg = replace(cfg0, eps=eps)
s = FitSession(rig, 1, cfg)
s.set_reference(ref)
wp = w.clone().requires_grad_(True); tp = t.clone().requires_grad_(True); qp = q.clone().requires_grad_(True)
opt = torch.optim.Adam([{'params': wp, 'lr': cfg.lr_base}, {'params': tp, 'lr': cfg.lr_t}, {'params': qp, 'lr': cfg.lr_q}], eps=eps)
sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))
for it in range(iters):
    gw_params = (s.w.cpu().clone(), s.t.cpu().clone(), s.q.cpu().clone())
    s.iteration()
    torch.cuda.synchronize()
    lo, gw, gt, gq = oracle_grad(*gw_params)
    print('it %d: same-params  loss gpu %.6f oracle %.6f | rel d_w %.2e  d_t %.2e  d_q %.2e' %
          (it, float(s.loss), lo, rel(s.d_w.cpu(), gw), rel(s.d_t.cpu(), gt), rel(s.d_q.cpu(), gq)))
    (l0, g0), (l1, g1) = staged(s)
    gp = s.g_pos.cpu()
    print('       staged on the GPU pos_clip: loss rel %.2e, d_pos rel %.2e | oracle vs ITSELF at pos_clip + 1 ulp: d_pos rel %.2e' %
          (abs(float(s.loss) - l0) / l0, rel(gp, g0), rel(g1, g0)))
    lf, gwf, gtf, gqf = oracle_grad(wp.detach(), tp.detach(), qp.detach())
    opt.zero_grad()
    wp.grad, tp.grad, qp.grad = gwf, gtf, gqf
    opt.step(); sched.step()
    with torch.no_grad():
        qp /= qp.norm(dim=1, keepdim=True)
    dw = np.abs(s.w.cpu().numpy() - wp.detach().numpy())[0]
    print('       free-running rel w %.2e (worst component %d, |g| there %.3g)  rel t %.2e  max|dq| %.2e' %
          (rel(s.w.cpu(), wp.detach()), int(dw.argmax()), float(gwf[0, int(dw.argmax())].abs()), rel(s.t.cpu(), tp.detach()),
           float(np.abs(s.q.cpu().numpy() - qp.detach().numpy()).max())))
