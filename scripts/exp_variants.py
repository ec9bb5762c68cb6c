"""Kernel-variant experiments on the GPU box: time the config-2 fit iteration for several builds of libfpc_b200.

    python scripts/exp_variants.py build  name:DEF1,DEF2 name2:DEF ...     (here: compiles libfpc_b200_<name>.so)
    python scripts/exp_variants.py run [--workload config2] name name2 ... (GPU box: one subprocess per variant)

'base' is the untagged product library.  Results (JSON lines) go to stdout and gpurun_out/exp_variants.jsonl.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lib_path(name):
    base = os.path.join(ROOT, 'fpc_diffrend_b200', 'libfpc_b200')
    return base + ('.so' if name == 'base' else '_%s.so' % name)


def one(workload):
    import torch
    import bench
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    wl = dict(bench.WORKLOADS[workload])
    if os.environ.get('FPC_EXP_SHADING'):
        wl['shading'] = os.environ['FPC_EXP_SHADING']
    F = int(os.environ.get('FPC_EXP_FRAMES', '0')) or min(wl['F'], 4)
    rig, w_all, t_all, q_all = bench.make_inputs(wl, F)
    # zero learning rates: the geometry (and with it the work per iteration) stays the same for every variant
    cfg = FitConfig(resolution=(wl['H'], wl['W']), shading=wl['shading'], antialias=wl['aa'], ref_dtype='u8',
                    lr_base=0.0, lr_t=0.0, lr_q=0.0, reorder_vertices=bool(int(os.environ.get('FPC_EXP_REORDER', '0'))))
    ref = synthesize_reference(rig, w_all, t_all, q_all, cfg, out_dtype=torch.uint8)
    sess = FitSession(rig, F, cfg)
    sess.set_reference(ref)
    sess.iteration()
    torch.cuda.synchronize()
    loss1 = float(sess.loss)
    gsum = float(sess.grads.double().abs().sum())
    sess.capture()
    for _ in range(10):
        sess.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            sess.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    sess.stage_events = {}
    for _ in range(20):
        sess.iteration()
    torch.cuda.synchronize()
    stages = {k: round(sum(a.elapsed_time(b) for a, b in v) / 20 * 1000, 1) for k, v in sess.stage_events.items()}
    print(json.dumps({'lib': os.path.basename(os.environ.get('FPC_B200_LIB', 'base')), 'step_us': round(best * 1000, 1),
                      'it_per_s': round(1000 / best, 1), 'loss_first': loss1, 'grad_abs_sum': gsum, 'loss_after': float(sess.loss),
                      'stage_us': stages}))


def main():
    cmd = sys.argv[1]
    if cmd == 'build':
        from fpc_diffrend_b200 import build as b
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition(':')
            print(b.build(defines=[d for d in defs.split(',') if d], tag='_' + name))
    elif cmd == 'one':
        one(sys.argv[2])
    elif cmd == 'run':
        args = sys.argv[2:]
        workload = 'config2'
        if args and args[0] == '--workload':
            workload, args = args[1], args[2:]
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'exp_variants.jsonl'), 'a') as log:
            for name in args:
                env = dict(os.environ)
                if name != 'base':
                    env['FPC_B200_LIB'] = lib_path(name)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), 'one', workload], env=env, capture_output=True, text=True)
                line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else json.dumps(
                    {'lib': name, 'error': (r.stderr or r.stdout)[-600:]})
                print(name, line, flush=True)
                log.write(line + '\n')


if __name__ == '__main__':
    main()
