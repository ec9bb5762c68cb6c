"""Generates tests/golden/blend_kat.json by running the REFERENCE's own blend functions
(/root/reference/src/torch/fit.py: blend_free :47-62, blend_combined :66-99, blend :103-129) on seeded inputs,
with torch autograd supplying the gradients of sum(vtx_pos * dy) w.r.t. every learnable matrix.

`import fit` fails in this container (nvdiffrast / roma / pytorch3d are absent), so the three function definitions
are taken from the reference file's AST and executed unchanged in a namespace that only holds `torch`.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_blend_kat.py
"""
import ast
import json
import os

import torch

REF_FIT = '/root/reference/src/torch/fit.py'
WANTED = ('blend_free', 'blend_combined', 'blend')

with open(REF_FIT) as f:
    tree = ast.parse(f.read(), REF_FIT)
funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
assert sorted(f.name for f in funcs) == sorted(WANTED)
ns = {'torch': torch}
exec(compile(ast.Module(body=funcs, type_ignores=[]), REF_FIT, 'exec'), ns)

g = torch.Generator().manual_seed(20261018)
R, B, F = 36, 4, 5          # 3V, blendshapes, frames of the take
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).float()
v_base = rnd(R, scale=10.0)
D = rnd(R, B)
M1 = rnd(F, F, scale=0.3)                      # maps['local']               (fit.py:219-224 initialises to 0 / eye)
M2 = torch.eye(B, F) + rnd(B, F, scale=0.2)    # maps_intermediate['local']
m1 = torch.eye(F) + rnd(F, F, scale=0.1)       # fit.py:175-177 initialises to eye / eye / 0
m2 = torch.eye(F) + rnd(F, F, scale=0.1)
m3 = rnd(R, F, scale=0.5)
dy = rnd(R)

out = {'source': 'reference fit.py blend / blend_free / blend_combined run on seeded inputs', 'R': R, 'B': B, 'F': F,
       'v_base': v_base.tolist(), 'D': D.tolist(), 'M1': M1.tolist(), 'M2': M2.tolist(), 'm1': m1.tolist(), 'm2': m2.tolist(),
       'm3': m3.tolist(), 'dy': dy.tolist(), 'frames': {}}
for frame in range(F):
    e = torch.zeros(F)
    e[frame] = 1.0
    rec = {}
    for mode in ('prior', 'free', 'combined'):
        P = {k: v.clone().requires_grad_(True) for k, v in (('M1', M1), ('M2', M2), ('m1', m1), ('m2', m2), ('m3', m3))}
        maps, maps_i, datasets = {'local': P['M1']}, {'local': P['M2']}, {'local': D}
        if mode == 'prior':
            v = ns['blend'](v_base, maps, maps_i, datasets, e)
        elif mode == 'free':
            v = ns['blend_free'](v_base, P['m1'], P['m2'], P['m3'], e)
        else:
            v = ns['blend_combined'](v_base, P['m1'], P['m2'], P['m3'], maps, maps_i, datasets, e, learned_coefficient=0.5)
        (v * dy).sum().backward()
        rec[mode] = {'vtx_pos': v.detach().tolist(),
                     'grad': {k: (p.grad.tolist() if p.grad is not None else None) for k, p in P.items()}}
    out['frames'][str(frame)] = rec
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'blend_kat.json')
with open(dst, 'w') as f:
    json.dump(out, f)
print('wrote', dst, os.path.getsize(dst), 'bytes')
