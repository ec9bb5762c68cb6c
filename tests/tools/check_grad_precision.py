"""Precision of d loss / d pos_clip at the shipped resolution: fused kernel and op-level chain against float64 autograd
(oracle/torch_ref.py) for one view.  usage: python tests/tools/check_grad_precision.py [H W V]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import clip_positions  # noqa: E402
from fpc_diffrend_b200 import _lib, rig as rigmod  # noqa: E402
import fpc_diffrend_b200.ops as dr  # noqa: E402
from oracle import golden as G, torch_ref as TR  # noqa: E402

H, W, V = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1600, 1200, 20000)
aa = '--no-aa' not in sys.argv
rig = rigmod.make_rig(n_vertices=V, n_shapes=4, n_cams=1, width=W, height=H, tex_size=256, seed=0)
pc = clip_positions(rig)
N, T, C = 1, rig.T, 1
rng = np.random.default_rng(21)
ref = np.round(rng.uniform(0, 140, size=(N, H, W, C))).astype(np.uint8)
cu = lambda a: torch.as_tensor(a).cuda().contiguous()
ctx = dr.RasterizeGLContext(device='cuda')
pos = cu(pc).requires_grad_(True)
tex = cu(rig.tex)[None]
rast, _ = dr.rasterize(ctx, pos, cu(rig.pos_idx), resolution=(H, W))
texc, _ = dr.interpolate(cu(rig.uv)[None], rast, cu(rig.uv_idx))
col = dr.texture(tex, texc, filter_mode='linear')
if aa:
    col = dr.antialias(col, rast, pos, cu(rig.pos_idx))
comp = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG, device='cuda'))
((cu(ref).float() - 255.0 * comp) ** 2).mean(dim=(1, 2, 3)).sum().backward()
g_ops = pos.grad.cpu().numpy()

P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
opp = dr.antialias_construct_topology_hash(cu(rig.pos_idx)).tri_opp
loss = torch.zeros(1, device='cuda')
g_pos = torch.empty(N, rig.V, 4, device='cuda')
scratch = torch.empty(int(_lib.load().fpc_render_loss_fused_scratch_bytes(N, T, H, W)), dtype=torch.uint8, device='cuda')
d_uv, d_uvi, d_tri, d_ref = cu(rig.uv), cu(rig.uv_idx), cu(rig.pos_idx), cu(ref)
head = (P(pos.detach()), P(d_tri)) + ((P(opp),) if aa else ())
adj = _lib.vertex_adjacency(d_tri, rig.V)      # kept alive until the synchronize below
_lib.call('fpc_render_loss_fused_aa' if aa else 'fpc_render_loss_fused', *head, P(d_uv), P(d_uvi), rig.uv.shape[0], 2, P(tex),
          rig.tex.shape[0], rig.tex.shape[1], P(d_ref), 1, N, rig.V, T, H, W, C, G.BG, 1.0, 0, P(loss), P(g_pos), None, None, None,
          P(adj[0]), P(adj[1]), P(scratch), scratch.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
g_fused = g_pos.cpu().numpy()

# float64 autograd on the GPU's own visibility
tid = rast[..., 3].detach().cpu().long() - 1
p64 = torch.tensor(pc, dtype=torch.float64, requires_grad=True)
tri, uvi = torch.tensor(rig.pos_idx), torch.tensor(rig.uv_idx)
u, v, zw = TR.barycentrics(p64, tri, tid, H, W)
r64 = torch.stack([u, v, zw, rast[..., 3].detach().cpu().double()], dim=-1)
c = TR.texture_linear(torch.tensor(rig.tex, dtype=torch.float64)[None], TR.interpolate(torch.tensor(rig.uv, dtype=torch.float64)[None], r64, uvi))
if aa:
    c = TR.antialias(c, r64, p64, tri, torch.tensor(G.topology_build(rig.pos_idx)))
c = torch.where(r64[..., 3:] > 0, c, torch.tensor(G.BG, dtype=torch.float64))
((torch.tensor(ref, dtype=torch.float64) - 255.0 * c) ** 2).mean(dim=(1, 2, 3)).sum().backward()
g64 = p64.grad.numpy()
mx = np.abs(g64).max()
for name, g in (('op-level chain', g_ops), ('fused kernel', g_fused)):
    err = np.abs(g - g64)
    i = np.unravel_index(err.argmax(), err.shape)
    print('%-15s max |err| / max |grad| = %.3e   (at vertex %d comp %d: %.6e vs %.6e)   rms rel = %.3e' %
          (name, err.max() / mx, i[1], i[2], g[i], g64[i], np.sqrt((err ** 2).mean()) / np.sqrt((g64 ** 2).mean())))
print('fused vs op-level: %.3e' % (np.abs(g_fused - g_ops).max() / mx))
for name, g in (('op-level chain', g_ops), ('fused kernel', g_fused)):
    ev = np.abs(g - g64).max(axis=(0, 2)) / mx
    print('%-15s vertices with err > 1e-4: %d, > 1e-3: %d of %d; worst five: %s' % (name, (ev > 1e-4).sum(), (ev > 1e-3).sum(), ev.size,
          ', '.join('%d:%.1e' % (i, ev[i]) for i in np.argsort(-ev)[:5])))
