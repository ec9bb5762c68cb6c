import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from fpc_diffrend_b200 import rig as rigmod
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
from oracle import golden as G

rig = rigmod.make_rig(n_vertices=600, n_shapes=8, n_cams=3, width=200, height=152, tex_size=32, seed=3)
H, W, F = 152, 200, 2
cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=False)
w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
s = FitSession(rig, F, cfg); s.set_reference(ref)
rng = np.random.default_rng(0)
w0 = (0.05 * rng.random((F, rig.B))).astype(np.float32)
s.set_parameters(w=w0); s.forward(); s.backward(); torch.cuda.synchronize()
refc = ref.cpu()
def rel(a, b):
    a = np.asarray(a); b = np.asarray(b); return np.abs(a-b).max()/max(np.abs(b).max(),1e-30)
w = torch.tensor(w0, requires_grad=True)
C = 3
tot = 0
keep = {}
for f in range(F):
    verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), w[f]).reshape(-1,3); verts.retain_grad(); keep[('v',f)] = verts
    for c in range(C):
        mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), torch.zeros(3), torch.tensor([0.,0,0,1])).requires_grad_(True)
        keep[('m',f,c)] = mvp
        pc = G.transform_clip(mvp, verts); pc.retain_grad(); keep[('pc',f,c)] = pc
        rast, _ = G.rasterize(pc, torch.tensor(rig.pos_idx), (H, W)); rast.retain_grad(); keep[('r',f,c)] = rast
        col = G.interpolate(torch.tensor(rig.vcol)[None], rast, torch.tensor(rig.pos_idx)); col.retain_grad(); keep[('c',f,c)] = col
        img = torch.where(rast[...,3:]>0, col, torch.tensor(G.BG))[0]
        tot = tot + G.image_loss(refc[f,c], img)/C
tot.backward()
print('loss', float(s.loss), float(tot))
for f in range(F):
    for c in range(C):
        n = f*C+c
        pcg = s.pos_clip[n].cpu().numpy(); pco = keep[('pc',f,c)][0].detach().numpy()
        print(f, c, 'pos_clip maxabs diff', np.abs(pcg-pco).max(), 'mvp diff', np.abs(s.mvp[n].cpu().numpy().reshape(4,4)-keep[('m',f,c)].detach().numpy()).max())
        rg = s.rast[n].cpu().numpy(); ro = keep[('r',f,c)][0].detach().numpy()
        print('   id mismatches', (rg[...,3]!=ro[...,3]).sum(), 'uv diff', np.abs(rg[...,:2]-ro[...,:2]).max())
        print('   d_colour rel', rel(s.d_colour[n].cpu(), keep[('c',f,c)].grad[0]), 'g_rast rel', rel(s.g_rast[n].cpu(), keep[('r',f,c)].grad[0]),
              'g_pos rel', rel(s.g_pos[n].cpu(), keep[('pc',f,c)].grad[0]), 'd_mvp rel', rel(s.d_mvp[n].cpu().reshape(4,4), keep[('m',f,c)].grad))
    print(f, 'd_verts rel', rel(s.d_verts[f].cpu().reshape(-1,3), keep[('v',f)].grad))
print('d_w rel', rel(s.d_w.cpu(), w.grad))
print(s.d_w.cpu().numpy()); print(w.grad.numpy())
# cross-check blend_bwd with torch on GPU
dw2 = s.d_verts @ s.D
print('blend_bwd vs torch gpu', rel(s.d_w.cpu(), dw2.cpu()))
dw3 = torch.stack([keep[('v',f)].grad.reshape(-1) for f in range(F)]) @ torch.tensor(rig.D)
print('oracle dverts @ D vs w.grad', rel(dw3, w.grad))
