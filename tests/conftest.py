import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200, sm_100a); run with -m gpu')


@pytest.fixture(scope='session')
def tiny_rig():
    """BASELINE config 1: 1k-vertex / 2k-triangle head, 16 blendshapes, 1 camera 128x128."""
    from fpc_diffrend_b200 import rig
    return rig.make_rig(n_vertices=1000, n_shapes=16, n_cams=1, width=128, height=128, tex_size=64, seed=0)


@pytest.fixture(scope='session')
def small_rig3():
    """3 cameras, non-square ragged resolution (not a multiple of the 64-px bin)."""
    from fpc_diffrend_b200 import rig
    return rig.make_rig(n_vertices=600, n_shapes=8, n_cams=3, width=200, height=152, tex_size=32, seed=3)


def clip_positions(rig, w=None, t=None, q=None, cams=None):
    """pos_clip [C,V,4] float32 through the ORACLE's torch stages (blend, MVP chain, transform_clip)."""
    import torch
    from oracle import golden as G
    B = rig.D.shape[1]
    w = torch.zeros(B) if w is None else torch.as_tensor(w, dtype=torch.float32)
    t = torch.zeros(3) if t is None else torch.as_tensor(t, dtype=torch.float32)
    q = torch.tensor([0.0, 0.0, 0.0, 1.0]) if q is None else torch.as_tensor(q, dtype=torch.float32)
    verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), w).reshape(-1, 3)
    out = []
    for c in (range(rig.P.shape[0]) if cams is None else cams):
        mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), t, q)
        out.append(G.transform_clip(mvp, verts)[0])
    return torch.stack(out).numpy().astype(np.float32)
