#!/usr/bin/env python
"""bench.py — fit iterations/s of the B200-native hot path (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config1|config3]

A "step" is one fit iteration (forward + backward + Adam) over ALL 9 views of every frame of the local batch.
N = 1 workload: BASELINE.json configs[1] — 20k-vertex / 40k-triangle rig, 200 blendshapes, 9 calibrated cameras
at 1024x1024, single frame, vertex-colour shading.  N > 1: one process per GPU (torchrun), every rank fits its
own frame(s) of the sequence (frames are independent units: no data-path collective) -> weak scaling.

value   : steps/s x N with all inputs resident in HBM, K replays of the captured CUDA graph, CUDA events,
          barrier + synchronize on both sides, max over ranks.
e2e     : the same metric through FitSession.iteration_from_host(): every step uploads that step's reference
          frames from pinned host memory and reads the loss back.
roofline: the dominant kernel group (largest share of the step), timed per launch with CUDA events in an eager
          pass of the same K steps; achieved = algorithmic bytes of that op (SURVEY §8(d) formulas)
          / its mean duration; peak = MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the CPU oracle (oracle/: torch CPU stages + scalar golden rasterizer, OpenMP
          over views) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'fit iters/s (fwd+bwd+Adam, 9 views, 1024^2)'
UNIT = 'iters/s'

WORKLOADS = {
    # name: (V, B, cams, H, W, frames per GPU, shading, antialias, tex)
    'config1': dict(V=1000, B=16, C=1, H=128, W=128, F=1, shading='vcol', aa=False, tex=64,
                    desc='BASELINE configs[0]: 1k-vertex/2k-tri rig, 16 blendshapes, 1 camera 128x128, 1 frame'),
    'config2': dict(V=20000, B=200, C=9, H=1024, W=1024, F=1, shading='vcol', aa=False, tex=64,
                    desc='BASELINE configs[1]: 20k-vertex/40k-tri rig, 200 blendshapes, 9 cameras 1024x1024, single frame, vertex-colour shading'),
    'config3': dict(V=20000, B=200, C=9, H=1024, W=1024, F=64, shading='texture', aa=True, tex=1024,
                    desc='BASELINE configs[2]: same rig, UV-textured shading + antialias, 64-frame batch'),
    'config4': dict(V=20000, B=200, C=9, H=1024, W=1024, F=512, shading='texture', aa=True, tex=1024, total_frames=True,
                    desc='BASELINE configs[3]: 512-frame sequence sharded by frame batches across the GPUs, 9 views 1024x1024, pose + activation Adam'),
    'config5': dict(V=50000, B=400, C=9, H=2048, W=2048, F=1, shading='texture', aa=True, tex=2048, split='cameras',
                    desc='BASELINE configs[4]: 50k-vertex/100k-tri rig, 400 blendshapes, 9 views 2048x2048, cameras split across GPUs, NCCL all-reduce of activation/pose grads'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=None, help='override the frames per GPU of the workload (parity-case workloads only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# ---------------------------------------------------------------------------------------------------------

def make_inputs(wl, n_frames, frame_seed=1):
    from fpc_diffrend_b200 import rig as rigmod
    rig = rigmod.make_rig(n_vertices=wl['V'], n_shapes=wl['B'], n_cams=wl['C'], width=wl['W'], height=wl['H'],
                          tex_size=wl['tex'], seed=0)
    w_true, t_true, q_true = rigmod.make_targets(n_frames, wl['B'], seed=frame_seed)
    return rig, w_true, t_true, q_true


# ---------------------------------------------------------------------------------------------------------
# algorithmic bytes per op (SURVEY §8(d) "ALGORITHMIC bytes"; restated in DESIGN.md)
# ---------------------------------------------------------------------------------------------------------

def algorithmic_bytes(wl, F, Vt, C=None):
    V, B, H, W = wl['V'], wl['B'], wl['H'], wl['W']
    C = C or wl['C']          # views rendered by this rank
    T = 2 * V - 4
    N = F * C
    px = N * H * W
    geo = N * 16 * V + 12 * T
    Ch = 3 if wl['shading'] == 'vcol' else 1
    A = Ch if wl['shading'] == 'vcol' else 2
    Va = V if wl['shading'] == 'vcol' else Vt
    R = 3 * V
    b = {
        'blend_fwd': 4 * (R * B + R + B * F + R * F),
        'blend_bwd': 4 * (R * B + R * F + B * F),
        'project_fwd': 4 * (R * F) + 64 * N + 16 * N * V,
        'project_bwd': 4 * (R * F) + 64 * N + 16 * N * V + 12 * V * F,
        'rasterize_fwd': 16 * px + geo,
        'rasterize_bwd': 32 * px + geo + 16 * N * V,
        'interpolate_fwd': (16 + 4 * A) * px + 12 * T + 4 * A * Va,
        'interpolate_bwd': (4 * A + 16 + 16) * px + 12 * T + 8 * A * Va,
        'image_loss': (4 * Ch + 4 * Ch + 4 + 4 * Ch) * px,
        # fused render(+antialias)+loss+gradient call.  SURVEY §8(d) "fused-path algorithmic bytes": 56+20C B/px + geometry
        'render_loss_fused': (56 + 20 * Ch) * px + geo + 12 * T + 4 * A * Va,
        # ... and what the kernels as built must move: u8 reference frame + geometry + attributes
        # + gradient zero/accumulate + moment zero/read
        'render_loss_fused_as_built': Ch * px + geo + 12 * T + 4 * A * Va + 2 * 16 * N * V + 2 * 36 * N * T,
        'adam': 28 * F * (B + 7),
        'pose_mvp_fwd': 64 * 3 * N,
        'pose_mvp_bwd': 64 * 3 * N,
    }
    # single-frame path: pose -> MVP, blend (an HBM / L2-bound GEMV over D) and the clip transform in one kernel per direction
    b['geometry_fwd'] = b['blend_fwd'] + b['project_fwd'] + b['pose_mvp_fwd']
    b['geometry_bwd'] = b['blend_bwd'] + b['project_bwd'] + b['pose_mvp_bwd']
    if wl['shading'] == 'texture':
        b['texture_fwd'] = (8 + 4 * Ch) * px + 4 * Ch * wl['tex'] ** 2
        b['texture_bwd'] = (4 * Ch + 8 + 8) * px + 4 * Ch * wl['tex'] ** 2
    if wl['aa']:
        b['antialias_fwd'] = (4 * Ch + 16 + 4 * Ch) * px + geo
        b['antialias_bwd'] = (4 * Ch + 4 * Ch + 16 + 4 * Ch) * px + geo + 16 * N * V
    return b


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads (and with them its pinned staging buffers, first-touch) to the NUMA node of its GPU, so
    that the per-step uploads of the `e2e` leg do not cross the socket interconnect.  Best effort: returns a description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        dev = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % dev) as f:
            node = int(f.read().strip())
        if node < 0:
            return 'gpu %s: no NUMA node reported' % dev
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 'gpu %s: NUMA node %d has no allowed cpu' % (dev, node)
        os.sched_setaffinity(0, cpus)
        return 'gpu %s -> NUMA node %d (%d cpus)' % (dev, node, len(cpus))
    except (OSError, ValueError, AttributeError) as e:
        return 'not bound (%s)' % e


def measured_traffic(workload, stage, F):
    """dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernels for ONE launch, from the committed ncu --set full
    captures (profiles/traffic.json, written by hand from profiles/*_ncu_full_*.txt); None when no capture covers it."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            t = json.load(f)
        e = t[workload][stage]
        return int(e['bytes_per_frame'] * F), e['source']
    except (OSError, KeyError, ValueError):
        return None, None


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.tmp = None

    def start(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        os.unlink(self.tmp.name)
        if sm:
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------

def cpu_oracle_rate(wl, budget_s=15.0, max_iters=60):
    """Fit iterations/s of the CPU oracle (torch CPU blend/project/loss + golden rasterizer with autograd +
    torch Adam) on one frame of the workload; bounded sample: 1 warm-up + up to max_iters timed iterations."""
    import torch
    from oracle import golden as G
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    rig, w_true, t_true, q_true = make_inputs(wl, 1)
    H, W, C = wl['H'], wl['W'], wl['C']
    tri = torch.tensor(rig.pos_idx)
    opp = torch.tensor(G.topology_build(rig.pos_idx)) if wl['aa'] else None
    base, D = torch.tensor(rig.v_base), torch.tensor(rig.D)
    Ps, As = torch.tensor(rig.P), torch.tensor(rig.A)
    vcol = torch.tensor(rig.vcol)
    uv, uv_idx, tex = torch.tensor(rig.uv), torch.tensor(rig.uv_idx), torch.tensor(rig.tex)

    def render_all(w, t, q):
        """All C views of one frame as ONE batched golden call per op (OpenMP over views inside golden.c)."""
        verts = G.blend(base, D, w).reshape(-1, 3)
        pcs = torch.cat([G.transform_clip(G.mvp_chain(Ps[c], As[c], t, q), verts) for c in range(C)])
        rast, _ = G.rasterize(pcs, tri, (H, W))
        if wl['shading'] == 'vcol':
            col = G.interpolate(vcol[None], rast, tri)
        else:
            col = G.texture(tex[None], G.interpolate(uv[None], rast, uv_idx))
        if wl['aa']:
            col = G.antialias(col, rast, pcs, tri, opp)
        return torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))

    with torch.no_grad():
        ref = torch.clamp(render_all(torch.tensor(w_true[0]), torch.tensor(t_true[0]), torch.tensor(q_true[0])) * 255, 0, 140)
    w = torch.zeros(wl['B'], requires_grad=True)
    t = torch.zeros(3, requires_grad=True)
    q = torch.tensor([0., 0, 0, 1], requires_grad=True)
    opt = torch.optim.Adam([{'params': w, 'lr': 1e-3}, {'params': t, 'lr': 1e-5}, {'params': q, 'lr': 1e-5}])

    def step():
        img = render_all(w, t, q)
        loss = sum(G.image_loss(ref[c], img[c]) for c in range(C)) / C
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            q.div_(q.norm())
        return float(loss.detach())

    step()
    t0 = time.perf_counter()
    n = 0
    while n < max_iters and (n == 0 or time.perf_counter() - t0 < budget_s):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, '%d fit iteration(s) of one frame (all %d views) after 1 warm-up, %.1f s' % (n, C, dt)


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rate, cores, sample = cpu_oracle_rate(wl, budget_s=40.0, max_iters=max(1, min(args.steps, 5)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000.0 / rate, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'note': 'CPU oracle port (nvdiffrast has no CPU backend; reference GPU path not installable offline)'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------

def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the fit hot path has no CPU fallback (use --impl reference for the CPU oracle)')
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 and not os.environ.get('FPC_NO_NUMA_BIND') else 'single process: not bound'
    sys.stderr.write('rank %d: %s\n' % (rank, numa))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from fpc_diffrend_b200 import shard
    cam_split = wl.get('split') == 'cameras'
    if wl.get('total_frames'):
        # a fixed-length sequence sharded by frame batches (strong scaling): this rank fits frames [f0, f1)
        n_total = args.frames or wl['F']
        f0, f1 = shard.frame_shard(n_total, rank, world)
    elif cam_split:
        # every rank holds the same frames and renders its share of the views
        n_total = args.frames or wl['F']
        f0, f1 = 0, n_total
    else:
        # weak scaling: every rank fits its own F frames of the sequence, rank r owns frames [r*F, (r+1)*F)
        Fg = args.frames or wl['F']
        n_total = Fg * world
        f0, f1 = rank * Fg, (rank + 1) * Fg
    F = f1 - f0
    rig, w_all, t_all, q_all = make_inputs(wl, n_total)
    sl = slice(f0, f1)
    # camera split cut at bin-row granularity: 9 views balance over 2/4/8 ranks (whole views would give 5+4, 3+2+2+2, 2+1x7)
    cam_slice, cam_band = shard.view_band_shard(wl['C'], wl['H'], rank, world) if cam_split else (None, None)
    # reference frames are stored as 8-bit grey levels like the reference's camera TIFFs (fit.py:530)
    ref_dtype = 'u8'
    cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'], ref_dtype=ref_dtype, cam_slice=cam_slice, cam_band=cam_band)
    ref = synthesize_reference(rig, w_all[sl], t_all[sl], q_all[sl], cfg, out_dtype=torch.uint8)
    sess = FitSession(rig, F, cfg)
    sess.set_reference(ref)
    ref_host = ref.cpu().pin_memory()
    del ref
    torch.cuda.empty_cache()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    launches = sess.iteration()           # first eager iteration (also warms the allocator)
    use_graph = not args.no_graph
    if use_graph:
        sess.capture()
    step = sess.replay if use_graph else sess.iteration

    # ---- value: inputs resident in HBM ----
    # Timing hygiene: the fused kernels materialise nothing per pixel, so one iteration's inputs (u8 frames, D, geometry:
    # ~100 MB at config 2) would FIT the 126 MB L2.  A buffer larger than L2 is therefore written between timed steps and
    # every step is bracketed by its own CUDA-event pair on the launching stream (the flush lies outside the pairs).  The
    # back-to-back figure without the flush is reported next to it as `steady_state` (the same frame re-read every iteration
    # is what a real single-frame fit does; it is not the headline).
    for _ in range(max(args.warmup, 3)):
        step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    for a, b in pairs:
        flush.zero_()
        a.record()
        step()
        b.record()
    barrier()
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))
    ms_per_step = ms_total / args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_steady = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    del flush
    if wl.get('total_frames') or cam_split:
        # one job-wide iteration updates all n_total frames (frame batches or views are spread over the ranks)
        value = 1000.0 / ms_per_step
        frames_per_s = n_total * 1000.0 / ms_per_step
        scaling = 'strong'
    else:
        value = world * 1000.0 / ms_per_step         # iterations/s summed over ranks (each rank iterates its own frames)
        frames_per_s = world * F * 1000.0 / ms_per_step
        scaling = 'weak'

    # ---- e2e: host buffers in, loss out, every step ----
    def host_frames(n):
        for _ in range(n):
            yield ref_host                      # this step's frames, in pinned host memory

    for _ in sess.fit_stream(host_frames(4), use_graph=use_graph):
        pass
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_losses = [l for l in sess.fit_stream(host_frames(args.steps), use_graph=use_graph)]
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)) / args.steps
    e2e = {'value': (1.0 if scaling == 'strong' else world) * 1000.0 / e2e_ms, 'unit': UNIT, 'h2d_bytes_per_step': int(ref_host.numel() * ref_host.element_size()),
           'd2h_bytes_per_step': 4, 'api': 'FitSession.fit_stream (double-buffered upload of the next step overlaps the current step)',
           'h2d_GBps': ref_host.numel() * ref_host.element_size() / (e2e_ms * 1e-3) / 1e9,
           'loss_last': e2e_losses[-1]}

    # ---- roofline: per-op CUDA-event timing over an eager pass of the same K steps ----
    sess.stage_events = {}
    for _ in range(args.steps):
        sess.iteration()
    torch.cuda.synchronize()
    stage_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in sess.stage_events.items()}
    sess.stage_events = None
    total_stage = sum(stage_ms.values())
    top = max(stage_ms, key=stage_ms.get)
    alg = algorithmic_bytes(wl, F, rig.uv.shape[0], sess.C)
    peak, peak_src = measured_peak()
    achieved = alg[top] / (stage_ms[top] * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(args.workload, top, F)
    roofline = {'bound': 'hbm', 'kernel': top, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg[top],
                'avg_ms_per_launch': stage_ms[top], 'share_of_step': stage_ms[top] / total_stage,
                'timing': 'CUDA events around the C-ABI call on its stream, eager pass of the same K steps'}
    if top == 'render_loss_fused':
        ab = alg['render_loss_fused_as_built']
        roofline['byte_model'] = ('SURVEY 8(d) fused-path algorithmic bytes: (56+20C) B/px + geometry; the call spans 4 launches '
                                  '(k_setup, k_fill, k_fused[_aa], k_tri_grad with the loss reduction riding along)')
        roofline['as_built_bytes_per_launch'] = ab
        roofline['as_built_GBps'] = ab / (stage_ms[top] * 1e-3) / 1e9
        roofline['note'] = ('the kernels keep every per-pixel intermediate on chip, so their compulsory HBM traffic (as_built_*) is ~10x '
                            'below the op-boundary model the roofline is quoted on; the kernel itself is issue/latency-bound (profiles/)')
    stages = {k: {'ms': round(v, 4), 'share': round(v / total_stage, 4),
                  'GBps_algorithmic': round(alg[k] / (v * 1e-3) / 1e9, 1) if k in alg and v > 0 else None}
              for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}

    def shutdown():
        # captured graphs hold NCCL work (camera-split mode): drop them before the process group, and never let a stuck
        # teardown keep the launcher alive
        if world > 1:
            import threading
            sess.graph = None
            sess._stream_graphs = [None, None]
            barrier()
            t = threading.Timer(15.0, os._exit, (0,))
            t.daemon = True
            t.start()
            dist.destroy_process_group()
            t.cancel()

    if rank != 0:
        shutdown()
        return

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': wl['desc'], 'frames_per_gpu': F, 'frames_fitted_per_s': frames_per_s,
                   'sharding': ('single GPU' if world == 1 else
                                'views over ranks (cut at 32-px bin rows), NCCL all-reduce of the packed (B+7)*F gradient vector per iteration' if cam_split else
                                'frames over ranks, no data-path collective'),
                   'cache': 'L2 flushed between timed steps (a 256 MB buffer is written outside the per-step CUDA-event pairs); `steady_state` = the same K steps back to back without the flush',
                   'reference_frames': ref_dtype + ' grey levels, resident in HBM for `value`, pinned host memory for `e2e`',
                   'launch': 'CUDA graph replay' if use_graph else 'eager', 'loss_final': float(sess.loss), 'host_affinity': numa},
        'steady_state': {'ms_per_step': ms_steady, 'value': value * ms_per_step / ms_steady,
                         'note': 'no L2 flush: the iteration re-reads the same frames, D and geometry, part of which the 126 MB L2 retains'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches * args.steps), 'roofline': roofline, 'stages': stages,
    }
    if not args.no_cpu_baseline and world == 1:
        rate, cores, sample = cpu_oracle_rate(wl)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}
    elif world > 1:
        line['cpu_baseline'] = None
    emit(line)
    shutdown()


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG is set): from here on file descriptor 1 is routed to stderr and the JSON line goes to the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
