"""Time the tcgen05 blend GEMM (forward + backward) at R = 3V, B, F on the GPU box; L2 flushed between launches.
    python scripts/time_blend_tc.py [V B F]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from fpc_diffrend_b200 import _lib  # noqa: E402

V, B, F = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (20000, 200, 64)
R = 3 * V
g = torch.Generator(device='cuda').manual_seed(0)
D = torch.randn(R, B, device='cuda', generator=g)
DT = D.t().contiguous()
base = torch.randn(R, device='cuda', generator=g)
w = torch.randn(F, B, device='cuda', generator=g)
dv = torch.randn(F, R, device='cuda', generator=g)
verts = torch.empty(F, R, device='cuda')
d_w = torch.empty(F, B, device='cuda')
L = _lib.load()
sc = torch.empty(int(L.fpc_blend_bwd_tc_scratch_bytes(R, B, F)), dtype=torch.uint8, device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
fwd = lambda: _lib.call('fpc_blend_fwd_tc', P(D), P(base), P(w), R, B, F, P(verts), st)
bwd = lambda: _lib.call('fpc_blend_bwd_tc', P(DT), P(dv), R, B, F, P(d_w), P(sc), sc.numel(), st)
fwd(); bwd(); torch.cuda.synchronize()
ref = base[None].double() + w.double() @ D.double().t()
err_f = float((verts.double() - ref).abs().max() / (w.double().abs() @ D.double().abs().t()).max())
refb = dv.double() @ D.double()
err_b = float((d_w.double() - refb).abs().max() / (dv.double().abs() @ D.double().abs()).max())
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = {}
for name, fn in (('fwd', fwd), ('bwd', bwd)):
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
    ts.sort()
    out[name] = ts[len(ts) // 2]
bytes_f = 4 * (R * B + F * B + R + F * R)
print('V=%d B=%d F=%d  fwd %.1f us (%.2f TB/s algorithmic)  bwd %.1f us  rel.err fwd %.2e bwd %.2e' %
      (V, B, F, out['fwd'], bytes_f / out['fwd'] / 1e6, out['bwd'], err_f, err_b))
