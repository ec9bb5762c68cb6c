"""ncu target: config-2 (or --workload X) session, warm-up, then ONE profiled eager iteration between
cudaProfilerStart/Stop.  Run as
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/X python scripts/prof_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference  # noqa: E402

workload = sys.argv[sys.argv.index('--workload') + 1] if '--workload' in sys.argv else 'config2'
iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 1
wl = bench.WORKLOADS[workload]
F = int(sys.argv[sys.argv.index('--frames') + 1]) if '--frames' in sys.argv else wl['F']
rig, w_all, t_all, q_all = bench.make_inputs(wl, F)
cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'], ref_dtype='u8', reorder_vertices=True)   # as bench.py
ref = synthesize_reference(rig, w_all, t_all, q_all, cfg, out_dtype=torch.uint8)
sess = FitSession(rig, F, cfg)
sess.set_reference(ref)
for _ in range(3):
    sess.iteration()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(iters):
    sess.iteration()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled %d iteration(s), loss %.4f' % (iters, float(sess.loss)))
