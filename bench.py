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
    ap.add_argument('--no-extra-legs', action='store_true', help='skip the config3/4/5 legs and the op-level stage timing')
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
            # sysfs reports no NUMA node (single-node hosts, some VMs): fall back to the CPU affinity NVML reports for the GPU
            # (the `CPU Affinity` column of `nvidia-smi topo -m`)
            try:
                import pynvml as nv
                nv.nvmlInit()
                h = nv.nvmlDeviceGetHandleByPciBusId(dev.encode())
                words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
                cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1} & os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    return 'gpu %s: no NUMA node in sysfs; bound to the %d cpus of its NVML cpu affinity' % (dev, len(cpus))
            except Exception as e:
                return 'gpu %s: no NUMA node reported, NVML affinity unavailable (%s)' % (dev, type(e).__name__)
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
        inst = e.get('warp_inst_per_frame')
        return int(e['bytes_per_frame'] * F), e['source'], (int(inst * F) if inst else None)
    except (OSError, KeyError, ValueError):
        return None, None, None


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
    """SM clock + throttle reasons sampled DURING the timed region.  NVML from a host thread every 5 ms (the timed region of
    the bench line is a few milliseconds long: nvidia-smi's 100 ms period would miss it); nvidia-smi -lms as the fallback
    when the NVML python binding is missing.  mark() / unmark() delimit the timed region: `sm_mhz` is the median of the
    samples taken inside it (or, if the region was shorter than one period, of the samples taken under load right around it)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period_s=0.005):
        self.index = index
        self.period = period_s
        self.proc = None
        self.tmp = None
        self.thread = None
        self.samples = []          # (t, sm_mhz, reasons_bitmask, in_region)
        self.in_region = False
        self._stop = False
        self.sm_max = None
        self.how = None

    def _nvml_loop(self, nv, h):
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), float(sm), int(rs), self.in_region))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.index
            if vis:
                parts = [p.strip() for p in vis.split(',') if p.strip()]
                if self.index < len(parts) and parts[self.index].isdigit():
                    idx = int(parts[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nv = nv
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.how = 'NVML, %g ms period' % (self.period * 1e3)
            return
        except Exception:
            self.thread = None
        try:
            self.tmp = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.tmp, stderr=subprocess.DEVNULL)
            self.how = 'nvidia-smi -lms 100'
        except Exception:
            self.proc = None

    def mark(self):
        self.in_region = True

    def unmark(self):
        self.in_region = False

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self._nv
            names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8)),
                     'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40)),
                     'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20)),
                     'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4))}
            inside = [x for x in self.samples if x[3]]
            use, where = inside, 'inside the timed region'
            if len(inside) < 3:
                # region shorter than a few periods: the samples taken under load (warm-up and the steady-state pass that
                # bracket it) stand in — idle samples (before the first launch) are excluded by taking the upper half
                allsm = sorted(x[1] for x in self.samples)
                use = [x for x in self.samples if x[1] >= allsm[len(allsm) // 2]] if allsm else []
                where = 'under load around the timed region (region shorter than 3 sampling periods)'
            if use:
                bits = 0
                for x in use:
                    bits |= x[2]
                out = {'sm_mhz': statistics.median(x[1] for x in use), 'sm_max_mhz': self.sm_max,
                       'reasons': sorted(k for k, m in names.items() if bits & m), 'samples': len(use),
                       'samples_in_timed_region': len(inside), 'window': where, 'how': self.how}
            return out
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
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm), 'how': self.how}
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------

def cpu_oracle_rate(wl, budget_s=15.0, max_iters=60):
    """Fit iterations/s of the CPU oracle (torch CPU blend/project/loss + golden rasterizer with autograd +
    torch Adam) on one frame of the workload; bounded sample: 1 warm-up + up to max_iters timed iterations.
    Returns (rate, threads, sample description): `threads` is what the run actually used — the OpenMP team of golden.c
    (views in parallel, forward and backward passes) and torch's intra-op pool."""
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
    omp, tth = G.omp_threads(), torch.get_num_threads()
    threads = max(min(omp, C), 1)            # the rendering ops parallelise over the C views of the frame
    return n / dt, threads, ('%d fit iteration(s) of one frame (all %d views) after 1 warm-up, %.1f s; golden.c OpenMP team %d over %d views '
                             '(forward and backward), torch intra-op threads %d, host cpus %d' % (n, C, dt, omp, C, tth, cores))


def nvdiffrast_rate(wl, steps, warmup):
    """Fit iterations/s of the reference's own GPU path (torch + nvdiffrast, RasterizeCudaContext) on one frame of the
    workload — the arm the north-star wants beaten.  Returns None (and the reason) when nvdiffrast is not importable."""
    from oracle import nvdiffrast_arm as NA
    dr = NA.probe()
    if dr is None:
        return None, NA.probe.reason
    import torch
    if not torch.cuda.is_available():
        return None, 'nvdiffrast importable but no CUDA device'
    try:
        from fpc_diffrend_b200.fit import FitConfig, synthesize_reference
        rig, w_true, t_true, q_true = make_inputs(wl, 1)
        cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'])
        ref = synthesize_reference(rig, w_true, t_true, q_true, cfg)[0]            # [C,H,W,Ch] float32, same frames as our arm
        step, _ = NA.fit_step_factory(dr, rig, (wl['H'], wl['W']), wl['shading'], wl['aa'], ref)
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {'value': 1000.0 / ms, 'ms_per_step': ms, 'context': step.context, 'loss_last': float(loss), 'module': NA.probe.reason}, None
    except Exception as e:      # an install that imports but cannot build / load its plugin offline
        return None, 'nvdiffrast present but failed: %s: %s' % (type(e).__name__, str(e).splitlines()[0] if str(e) else '')


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    gpu_ref, why = nvdiffrast_rate(wl, args.steps, args.warmup)
    rate, cores, sample = cpu_oracle_rate(wl, budget_s=40.0, max_iters=max(1, min(args.steps, 5)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000.0 / rate, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'note': 'CPU oracle port (nvdiffrast has no CPU backend); see reference_gpu for the nvdiffrast CUDA path'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'reference_gpu': gpu_ref if gpu_ref is not None else 'unavailable',
        'reference_gpu_reason': why,
    }
    if gpu_ref is not None:
        # the reference's own GPU path exists on this box: it IS the reference arm (the CPU port stays in cpu_baseline)
        line.update(value=gpu_ref['value'], ms_per_step=gpu_ref['ms_per_step'])
        line['e2e'] = {'value': gpu_ref['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
        line['cpu_baseline']['kind'] = 'port (timed beside the nvdiffrast-cuda arm this line reports)'
        line['config']['note'] = 'unmodified nvdiffrast (%s) driven as fit.py:134-162,524-642 does, all 9 views of one frame per step' % gpu_ref['context']
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------

class Job:
    """Process-wide plumbing of one bench run (rank, world, barrier, max-over-ranks)."""

    def __init__(self):
        import torch
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a CUDA device: the fit hot path has no CPU fallback (use --impl reference for the CPU oracle)')
        torch.cuda.set_device(self.local_rank)
        self.numa = bind_to_gpu_numa(self.local_rank) if self.world > 1 and not os.environ.get('FPC_NO_NUMA_BIND') else 'single process: not bound'
        sys.stderr.write('rank %d: %s\n' % (self.rank, self.numa))
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        import torch
        import torch.distributed as dist
        tt = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    def gather(self, x):
        """[x of rank 0, x of rank 1, ...] on every rank."""
        if self.world == 1:
            return [x]
        import torch
        import torch.distributed as dist
        tt = torch.zeros(self.world, dtype=torch.float64, device='cuda')
        tt[self.rank] = x
        dist.all_reduce(tt)
        return [float(v) for v in tt]


def build_session(job, wl, n_frames_arg=None, keep_host=False):
    """Synthetic inputs + the FitSession of this rank for workload `wl` (frames / views sharded as the workload says)."""
    import torch
    from fpc_diffrend_b200 import shard
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rank, world = job.rank, job.world
    cam_split = wl.get('split') == 'cameras' and world > 1
    if wl.get('total_frames'):
        # a fixed-length sequence sharded by frame batches (strong scaling): this rank fits frames [f0, f1)
        n_total = n_frames_arg or wl['F']
        f0, f1 = shard.frame_shard(n_total, rank, world)
    elif wl.get('split') == 'cameras':
        # every rank holds the same frames and renders its share of the views
        n_total = n_frames_arg or wl['F']
        f0, f1 = 0, n_total
    else:
        # weak scaling: every rank fits its own F frames of the sequence, rank r owns frames [r*F, (r+1)*F)
        Fg = n_frames_arg or wl['F']
        n_total = Fg * world
        f0, f1 = rank * Fg, (rank + 1) * Fg
    F = f1 - f0
    rig, w_all, t_all, q_all = make_inputs(wl, n_total)
    sl = slice(f0, f1)
    # camera split cut at bin-row granularity: 9 views balance over 2/4/8 ranks (whole views would give 5+4, 3+2+2+2, 2+1x7)
    cam_slice, cam_band = shard.view_band_shard(wl['C'], wl['H'], rank, world) if cam_split else (None, None)
    # reference frames are stored as 8-bit grey levels like the reference's camera TIFFs (fit.py:530)
    cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'], ref_dtype='u8', cam_slice=cam_slice, cam_band=cam_band,
                    reorder_vertices=True)
    ref = synthesize_reference(rig, w_all[sl], t_all[sl], q_all[sl], cfg, out_dtype=torch.uint8)
    sess = FitSession(rig, F, cfg)
    sess.set_reference(ref)
    ref_host = ref.cpu().pin_memory() if keep_host else None
    del ref
    torch.cuda.empty_cache()
    scaling = 'strong' if (wl.get('total_frames') or wl.get('split') == 'cameras') else 'weak'
    return dict(rig=rig, sess=sess, cfg=cfg, F=F, n_total=n_total, ref_host=ref_host, scaling=scaling, cam_split=cam_split,
                targets=(w_all[sl], t_all[sl], q_all[sl]))


def timed_steps(job, step, steps, warmup, sampler=None, min_warm_s=0.0):
    """W untimed steps, then EXACTLY `steps` steps, each with its own CUDA-event pair on the launching stream and the L2
    flushed before it (a 256 MB buffer written outside the pairs); barrier + synchronize on both sides; max over ranks.
    Returns (ms_per_step flushed, ms_per_step back to back, warm-up steps actually run)."""
    import torch
    nwarm = 0
    t0 = time.perf_counter()
    while nwarm < max(warmup, 3) or (time.perf_counter() - t0) < min_warm_s:
        step()
        nwarm += 1
        if nwarm % 16 == 0:
            torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    job.barrier()
    if sampler is not None:
        sampler.mark()
    for a, b in pairs:
        flush.zero_()
        a.record()
        step()
        b.record()
    job.barrier()
    if sampler is not None:
        sampler.unmark()
    ms_flushed = job.max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs)) / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    job.barrier()
    ms_steady = job.max_over_ranks(e0.elapsed_time(e1)) / steps
    del flush
    return ms_flushed, ms_steady, nwarm


def allreduce_us(job, sess, reps=50):
    """Mean duration of the camera-split mode's only exchange (all-reduce of the packed gradient vector), CUDA events."""
    import torch
    from fpc_diffrend_b200.shard import allreduce_gradients
    g = sess.grads.clone()
    for _ in range(5):
        allreduce_gradients(g)
    job.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        allreduce_gradients(g)
    e1.record()
    job.barrier()
    return job.max_over_ranks(e0.elapsed_time(e1)) / reps * 1000.0


def _half_pose(q):
    """A pose between the identity and the ground truth (not normalised: the kernels take any quaternion, as roma does)."""
    import numpy as np
    return 0.5 * np.asarray(q, dtype=np.float32) + 0.5 * np.array([0., 0., 0., 1.], dtype=np.float32)


def split_parity(job, wl, ctx):
    """In-run parity of the camera split: packed gradient [d_w | d_t | d_q] and loss, all-reduced over the N ranks that each
    rendered their band of the views, against rank 0 rendering ALL views of the same frames alone.  Parameters are the
    ground truth scaled by 0.5 (a non-trivial state).  Returns {'grad_rel_err', 'loss_rel_err'} on every rank."""
    import torch
    import torch.distributed as dist
    from fpc_diffrend_b200.fit import FitSession, synthesize_reference
    from dataclasses import replace
    sess = ctx['sess']
    w, t, q = ctx['targets']
    qh = _half_pose(q)
    sess.set_parameters(0.5 * w, 0.5 * t, qh)
    sess.forward(); sess.backward()
    torch.cuda.synchronize()
    g = sess.grads.clone()
    l = sess.loss.clone()
    dist.all_reduce(g); dist.all_reduce(l)
    out = torch.zeros(2, dtype=torch.float64, device='cuda')
    if job.rank == 0:
        cfg = replace(ctx['cfg'], cam_slice=None, cam_band=None)
        ref = synthesize_reference(ctx['rig'], w, t, q, cfg, out_dtype=torch.uint8)
        full = FitSession(ctx['rig'], ctx['F'], cfg)
        full.set_reference(ref)
        full.set_parameters(0.5 * w, 0.5 * t, qh)
        full.forward(); full.backward()
        torch.cuda.synchronize()
        out[0] = float((g - full.grads).abs().max() / full.grads.abs().max().clamp_min(1e-30))
        out[1] = float((l - full.loss).abs() / full.loss.abs().clamp_min(1e-30))
        del full, ref
    dist.all_reduce(out)
    sess.reset_state()
    torch.cuda.empty_cache()
    return {'grad_rel_err': float(out[0]), 'loss_rel_err': float(out[1]),
            'what': 'all-reduced packed gradient / loss of the N band-split ranks vs rank 0 rendering all views alone (max |d| / max |g|)'}


def shard_parity(job, wl, ctx):
    """In-run parity of the frame sharding: the gradient row of the LAST rank's first frame, computed inside that rank's
    batch, against rank 0 fitting that single frame alone (frames are independent units)."""
    import torch
    import torch.distributed as dist
    from fpc_diffrend_b200 import shard
    from fpc_diffrend_b200.fit import FitSession, synthesize_reference
    sess = ctx['sess']
    w, t, q = ctx['targets']
    sess.set_parameters(0.5 * w, 0.5 * t, _half_pose(q))
    sess.forward(); sess.backward()
    row = torch.cat([sess.d_w[0], sess.d_t[0], sess.d_q[0]]).clone()
    last = job.world - 1
    dist.broadcast(row, src=last)
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    if job.rank == 0:
        f0, _ = shard.frame_shard(ctx['n_total'], last, job.world)
        _, w_all, t_all, q_all = make_inputs(wl, ctx['n_total'])
        w1, t1, q1 = w_all[f0:f0 + 1], t_all[f0:f0 + 1], q_all[f0:f0 + 1]
        ref = synthesize_reference(ctx['rig'], w1, t1, q1, ctx['cfg'], out_dtype=torch.uint8)
        one = FitSession(ctx['rig'], 1, ctx['cfg'])
        one.set_reference(ref)
        one.set_parameters(0.5 * w1, 0.5 * t1, _half_pose(q1))
        one.forward(); one.backward()
        torch.cuda.synchronize()
        out[0] = float((row - one.grads).abs().max() / one.grads.abs().max().clamp_min(1e-30))
        del one, ref
    dist.all_reduce(out)
    sess.reset_state()
    torch.cuda.empty_cache()
    return {'grad_rel_err': float(out[0]),
            'what': ("gradient row of the last rank's first frame (inside its batch) vs rank 0 fitting that frame alone (max |d| / max |g|); "
                     "the two sides blend on different paths (3xTF32 tensor-core GEMM for the batch, fp32 GEMV for the single frame), so "
                     "pos_clip differs in its last bits and with it the few silhouette pixels on sliver triangles that carry the largest "
                     "position gradients (DESIGN.md section 6): 1e-4 .. 2e-3 is that effect, not a sharding error")}


def run_leg(job, name, steps, warmup, frames=None):
    """One extra, driver-visible measurement of another BASELINE config with the same timing method as the bench line
    (CUDA-graph replay, L2 flushed between steps, CUDA events, max over ranks).  All ranks call it; returns a dict."""
    import torch
    wl = WORKLOADS[name]
    ctx = build_session(job, wl, frames)
    sess = ctx['sess']
    out = {'workload': wl['desc'], 'n_gpus': job.world, 'frames_per_gpu': ctx['F'], 'frames_total': ctx['n_total'], 'steps': steps}
    try:
        if job.world > 1 and ctx['cam_split']:
            out['parity'] = split_parity(job, wl, ctx)
        elif job.world > 1 and wl.get('total_frames'):
            out['parity'] = shard_parity(job, wl, ctx)
        launches = sess.iteration()
        sess.capture()
        ms, ms_steady, nwarm = timed_steps(job, sess.replay, steps, warmup)
        if ctx['scaling'] == 'strong':
            out.update(value=1000.0 / ms, unit='job iters/s (one iteration updates all %d frame(s))' % ctx['n_total'],
                       frames_fitted_per_s=ctx['n_total'] * 1000.0 / ms)
        else:
            out.update(value=job.world * 1000.0 / ms, unit='batch iters/s summed over ranks (one iteration updates %d frames per rank)' % ctx['F'],
                       frames_fitted_per_s=job.world * ctx['F'] * 1000.0 / ms)
        out.update(ms_per_step=ms, ms_per_step_steady=ms_steady, warmup=nwarm, scaling=ctx['scaling'], launches_per_iteration=launches,
                   loss_final=sess.total_loss())
        if job.world > 1 and ctx['cam_split']:
            out['exchange'] = ('NVLink peer memory: blend + all-gather in one kernel, peer sums, 3 device-side barriers per iteration (csrc/peer.cu)'
                               if getattr(sess, 'peer', None) is not None else
                               'NCCL: all-gather of the vertices, reduce-scatter of their gradients, all-reduce of the packed gradient'
                               if sess.row_shard is not None else 'NCCL all-reduce of the packed gradient (D replicated)')
            out['nccl_allreduce_us'] = allreduce_us(job, sess)
            out['allreduce_floats'] = int(sess.grads.numel())
        # fused-design bound of SURVEY 8(d) for this workload, per rank
        alg = algorithmic_bytes(wl, ctx['F'], ctx['rig'].uv.shape[0], sess.C)
        peak, _ = measured_peak()
        share = 1.0
        if ctx['cam_split']:
            share = 1.0 / job.world          # the views are cut evenly at bin-row granularity
        out['fused_bound_frac'] = alg['render_loss_fused'] * share / (ms * 1e-3) / 1e9 / peak
    finally:
        sess.invalidate_graphs()
        del sess, ctx
        torch.cuda.empty_cache()
    return out


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    job = Job()
    rank, world, local_rank = job.rank, job.world, job.local_rank
    barrier, max_over_ranks = job.barrier, job.max_over_ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                   # before the warm-up, so that the line's own `clocks` is never null
    ctx = build_session(job, wl, args.frames, keep_host=True)
    sess, rig, F, n_total, ref_host, scaling, cam_split = ctx['sess'], ctx['rig'], ctx['F'], ctx['n_total'], ctx['ref_host'], ctx['scaling'], ctx['cam_split']
    ref_dtype = 'u8'

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
    # is what a real single-frame fit does; it is not the headline).  The warm-up runs for at least 0.3 s so that the clocks
    # have ramped and the sampler has seen the GPU under load before the (few-millisecond) timed region starts.
    ms_per_step, ms_steady, nwarm = timed_steps(job, step, args.steps, args.warmup, sampler if rank == 0 else None, min_warm_s=0.3)
    clocks = sampler.stop() if rank == 0 else None
    if scaling == 'strong':
        # one job-wide iteration updates all n_total frames (frame batches or views are spread over the ranks)
        value = 1000.0 / ms_per_step
        frames_per_s = n_total * 1000.0 / ms_per_step
    else:
        value = world * 1000.0 / ms_per_step         # iterations/s summed over ranks (each rank iterates its own frames)
        frames_per_s = world * F * 1000.0 / ms_per_step

    # ---- e2e: host buffers in, loss out, every step ----
    def host_frames(n):
        for _ in range(n):
            yield ref_host                      # this step's frames, in pinned host memory

    for _ in sess.fit_stream(host_frames(4), use_graph=use_graph):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    e2e_losses = [l for l in sess.fit_stream(host_frames(args.steps), use_graph=use_graph)]
    e1.record()
    torch.cuda.synchronize()
    my_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0) / args.steps
    barrier()
    e2e_ms = max_over_ranks(my_ms)
    h2d = int(ref_host.numel() * ref_host.element_size())
    e2e = {'value': (1.0 if scaling == 'strong' else world) * 1000.0 / e2e_ms, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': 4, 'api': 'FitSession.fit_stream (double-buffered upload of the next step overlaps the current step)',
           'h2d_GBps': h2d / (e2e_ms * 1e-3) / 1e9, 'h2d_GBps_per_rank': [round(h2d / (m * 1e-3) / 1e9, 2) for m in job.gather(my_ms)],
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
    traffic, traffic_src, inst = measured_traffic(args.workload, top, F)
    roofline = {'bound': 'hbm', 'kernel': top, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg[top],
                'avg_ms_per_launch': stage_ms[top], 'share_of_step': stage_ms[top] / total_stage,
                'timing': 'CUDA events around the C-ABI call on its stream, eager pass of the same K steps'}
    if top == 'render_loss_fused':
        ab = alg['render_loss_fused_as_built']
        roofline['byte_model'] = ('SURVEY 8(d) fused-path algorithmic bytes: (56+20C) B/px + geometry; the call spans its binning, '
                                  'fused render+loss+gradient and per-vertex gradient gather launches (k_setup, k_fill, k_fused[_aa], k_vtx_gather)')
        roofline['as_built_bytes_per_launch'] = ab
        roofline['as_built_GBps'] = ab / (stage_ms[top] * 1e-3) / 1e9
        roofline['frac_hbm_model'] = achieved / peak
        if inst:
            # instruction roofline: warp instructions the stage executes (ncu smsp__inst_executed.sum, profiles/) against the
            # issue peak of the chip = SMs x 4 schedulers x 1 warp instruction per cycle at the SM clock measured in this run
            sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
            issue_peak = 148 * 4 * sm_mhz * 1e6
            ifrac = inst / (stage_ms[top] * 1e-3) / issue_peak
            roofline['issue'] = {'warp_inst_per_launch': inst, 'peak_warp_inst_per_s': issue_peak, 'frac': ifrac, 'sm_mhz': sm_mhz,
                                 'note': 'achieved issue rate / peak issue rate; 1 - frac is the stall share (latency, barriers, tail)'}
            roofline['frac'] = min(roofline['frac_hbm_model'], ifrac)
        roofline['note'] = ('the kernels keep every per-pixel intermediate on chip, so their compulsory HBM traffic (as_built_*) is ~10x '
                            'below the op-boundary model `achieved` is quoted on; the stage is instruction-issue bound: `frac` is the smaller '
                            'of the HBM-model fraction and the instruction-issue fraction (`issue`)')
    stages = {k: {'ms': round(v, 4), 'share': round(v / total_stage, 4),
                  'GBps_algorithmic': round(alg[k] / (v * 1e-3) / 1e9, 1) if k in alg and v > 0 else None}
              for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}
    loss_final = float(sess.loss)
    launches_total = int(launches * args.steps)

    # ---- the op-level (drop-in) chain: what the reference's unchanged fit.py calls, op by op ----
    stages_oplevel = None
    if not args.no_extra_legs and world == 1 and args.workload == 'config2':
        try:
            stages_oplevel = oplevel_stages(wl, rig, ctx['targets'], args.steps)
        except Exception as e:
            stages_oplevel = {'error': '%s: %s' % (type(e).__name__, e)}

    def shutdown():
        # captured graphs hold NCCL work (camera-split mode): drop them before the process group, and never let a stuck
        # teardown keep the launcher alive
        if world > 1:
            import threading
            barrier()
            t = threading.Timer(15.0, os._exit, (0,))
            t.daemon = True
            t.start()
            dist.destroy_process_group()
            t.cancel()

    line = None
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': nwarm,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': wl['desc'], 'frames_per_gpu': F, 'frames_fitted_per_s': frames_per_s,
                       'sharding': ('single GPU' if world == 1 else
                                    'views over ranks (cut at 32-px bin rows), NCCL all-reduce of the packed (B+7)*F gradient vector per iteration' if cam_split else
                                    'frames over ranks, no data-path collective'),
                       'cache': 'L2 flushed between timed steps (a 256 MB buffer is written outside the per-step CUDA-event pairs); `steady_state` = the same K steps back to back without the flush',
                       'reference_frames': ref_dtype + ' grey levels, resident in HBM for `value`, pinned host memory for `e2e`',
                       'launch': 'CUDA graph replay' if use_graph else 'eager', 'loss_final': loss_final, 'host_affinity': job.numa,
                       'mesh': 'rig in shuffled (authoring) order; the session renumbers its vertices along a Morton curve at set-up '
                               '(FitConfig.reorder_vertices, as fit_take does); triangle order untouched'},
            'steady_state': {'ms_per_step': ms_steady, 'value': value * ms_per_step / ms_steady,
                             'note': 'no L2 flush: the iteration re-reads the same frames, D and geometry, part of which the 126 MB L2 retains'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches_total, 'launches_per_iteration': int(launches), 'roofline': roofline, 'stages': stages,
        }
        if stages_oplevel is not None:
            line['stages_oplevel'] = stages_oplevel
        if world > 1:
            line['cpu_baseline'] = None

    # ---- extra legs: the other BASELINE configs, same timing method (driver-visible; the bench line stays config 2) ----
    sess.invalidate_graphs()
    del sess, ctx, ref_host
    torch.cuda.empty_cache()
    if not args.no_extra_legs and args.workload == 'config2':
        import threading

        def give_up():
            # a leg that hangs (a rank died inside a collective) must not take the bench line with it
            if rank == 0:
                line['legs_error'] = 'extra legs exceeded their time budget; line emitted without them'
                emit(line)
            os._exit(0)

        dog = threading.Timer(420.0 if rank == 0 else 430.0, give_up)
        dog.daemon = True
        dog.start()
        names = ['config3', 'config4', 'config5'] if world == 1 else ['config4', 'config5']
        for nme in names:
            try:
                leg = run_leg(job, nme, steps=min(args.steps, 20 if nme == 'config5' else 5), warmup=3)
            except Exception as e:           # a leg must never take the bench line down with it
                leg = {'error': '%s: %s' % (type(e).__name__, e)}
                if world > 1:                # the other ranks are inside the leg's collectives: nothing left to do together
                    if rank == 0:
                        line[nme] = leg
                        emit(line)
                    os._exit(0)
            if rank == 0:
                line[nme] = leg
        dog.cancel()

    if rank != 0:
        shutdown()
        return
    if not args.no_cpu_baseline and world == 1:
        rate, cores, sample = cpu_oracle_rate(wl)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}
        from oracle import nvdiffrast_arm as NA
        line['reference_gpu'] = 'unavailable' if NA.probe() is None else 'available: see the --impl reference arm'
        line['reference_gpu_reason'] = NA.probe.reason
    emit(line)
    shutdown()


def oplevel_stages(wl, rig, targets, steps):
    """Per-op timing of the drop-in chain at this workload: FitSession(fused=False) runs rasterize -> interpolate -> [texture]
    -> [antialias] -> image loss and their backward passes as separate C-ABI calls with every per-pixel tensor in HBM (what
    the reference's unchanged fit.py drives through `import fpc_diffrend_b200.ops as dr`).  GB/s on the per-op algorithmic
    bytes of SURVEY 8(d)."""
    import torch
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'], fused=False)
    w, t, q = targets
    ref = synthesize_reference(rig, w, t, q, cfg)
    s = FitSession(rig, 1, cfg)
    s.set_reference(ref)
    for _ in range(3):
        s.iteration()
    s.stage_events = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = {k: sum(a.elapsed_time(b) for a, b in v) / steps for k, v in s.stage_events.items()}
    alg = algorithmic_bytes(wl, 1, rig.uv.shape[0], s.C)
    peak, _ = measured_peak()
    out = {k: {'ms': round(v, 4), 'GBps_algorithmic': round(alg[k] / (v * 1e-3) / 1e9, 1) if k in alg else None,
               'frac_of_hbm_peak': round(alg[k] / (v * 1e-3) / 1e9 / peak, 3) if k in alg else None}
           for k, v in sorted(ms.items(), key=lambda kv: -kv[1])}
    out['_iteration'] = {'ms': round(e0.elapsed_time(e1) / steps, 4), 'iters_per_s': round(1000.0 * steps / e0.elapsed_time(e1), 1),
                         'note': 'eager op-level iteration (no CUDA graph), host launch overhead included'}
    del s, ref
    torch.cuda.empty_cache()
    return out


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
