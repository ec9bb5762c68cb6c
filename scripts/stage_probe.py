"""Per-stage times (eager, CUDA events) of one workload on every rank: where does a camera-split iteration go?
    torchrun --nproc-per-node N scripts/stage_probe.py [--workload config5]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

workload = sys.argv[sys.argv.index('--workload') + 1] if '--workload' in sys.argv else 'config5'
job = bench.Job()
ctx = bench.build_session(job, bench.WORKLOADS[workload])
sess = ctx['sess']
for _ in range(5):
    sess.iteration()
job.barrier()
sess.stage_events = {}
K = 20
for _ in range(K):
    sess.iteration()
job.barrier()
stages = {k: round(sum(a.elapsed_time(b) for a, b in v) / K * 1000, 1) for k, v in sess.stage_events.items()}
sess.stage_events = None
sess.capture()
for _ in range(10):
    sess.replay()
job.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    sess.replay()
e1.record()
job.barrier()
step = e0.elapsed_time(e1) / 50 * 1000
rows = job.gather(stages.get('render_loss_fused', 0.0))
if job.rank == 0:
    print(json.dumps({'world': job.world, 'graph_step_us': round(step, 1), 'rank0_stage_us': stages, 'render_us_per_rank': [round(x, 1) for x in rows],
                      'launches': sess.launches_per_iteration, 'cam': [str(sess.cfg.cam_slice), str(sess.cfg.cam_band)]}))
sess.invalidate_graphs()
del sess
torch.cuda.synchronize()
if job.world > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
