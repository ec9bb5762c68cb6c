"""Diagnostic: worst-vertex position-gradient discrepancy at BASELINE config 2 size after `its` fit iterations — fused path,
op-level path (per-pixel kernels) and the float64-accumulating oracle on the SAME pos_clip bits.
usage: python tests/tools/debug_sliver_grad.py [its]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fpc_diffrend_b200 import rig as rigmod                                     # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference     # noqa: E402
from oracle import golden as G                                                    # noqa: E402
from dataclasses import replace                                                   # noqa: E402

its = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = W = 1024
rig = rigmod.make_rig(n_vertices=20000, n_shapes=200, n_cams=9, width=W, height=H, tex_size=64, seed=0)
cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=False, eps=1.208)
w_true, t_true, q_true = rigmod.make_targets(1, rig.B, seed=1)
ref = synthesize_reference(rig, w_true, t_true, q_true, cfg)
s = FitSession(rig, 1, cfg)
s.set_reference(ref)
for _ in range(its):
    s.iteration()
s.forward(); s.backward()
torch.cuda.synchronize()
o = FitSession(rig, 1, replace(cfg, fused=False))
o.set_reference(ref)
o.params.copy_(s.params)
o.forward(); o.backward()
torch.cuda.synchronize()
assert torch.equal(o.pos_clip, s.pos_clip)
pc = s.pos_clip.cpu().clone().requires_grad_(True)
tri = torch.tensor(rig.pos_idx)
rast, _ = G.rasterize(pc, tri, (H, W))
col = G.interpolate(torch.tensor(rig.vcol)[None], rast, tri)
img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))
refc = ref.cpu()
loss = sum(G.image_loss(refc[0, c], img[c]) for c in range(9)) / 9
loss.backward()
go = pc.grad.numpy(); gf = s.g_pos.cpu().numpy(); gp = o.g_pos.cpu().numpy()
mx = np.abs(go).max()
print('max |d_pos| oracle %.4g' % mx)
for name, g in (('fused', gf), ('op-level', gp)):
    e = np.abs(g - go).max(axis=-1)
    n, v = np.unravel_index(e.argmax(), e.shape)
    print('%-9s vs oracle: rel %.3e at view %d vertex %d: %s vs oracle %s' % (name, e.max() / mx, n, v, g[n, v], go[n, v]))
e = np.abs(gf - gp).max(axis=-1)
n, v = np.unravel_index(e.argmax(), e.shape)
print('fused vs op-level: rel %.3e at view %d vertex %d' % (e.max() / mx, n, v))
e = np.abs(gf - go).max(axis=-1)
n, v = np.unravel_index(e.argmax(), e.shape)
inc = np.where((rig.pos_idx == v).any(axis=1))[0]
pcn = s.pos_clip[n].cpu().numpy()
r = rast[n].detach().numpy()
for t in inc:
    px = int((r[..., 3] == t + 1).sum())
    p = pcn[rig.pos_idx[t]]
    ndc = p[:, :2] / p[:, 3:4]
    area = 0.5 * abs((ndc[1, 0] - ndc[0, 0]) * (ndc[2, 1] - ndc[0, 1]) - (ndc[2, 0] - ndc[0, 0]) * (ndc[1, 1] - ndc[0, 1])) * (W / 2) * (H / 2)
    print('  incident triangle %d: %d visible px, screen area %.4f px^2' % (t, px, area))
