"""Diagnostic (torchrun, >= 2 GPUs): does torch symmetric memory work on this box, step by step, with progress lines.
   timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/tools/peer_probe.py"""
import ctypes
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
faulthandler.dump_traceback_later(90, exit=True)
import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])


def say(*a):
    print('[rank %d %.1fs]' % (rank, time.time() - T0), *a, flush=True)


T0 = time.time()
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
say('process group up')
import torch.distributed._symmetric_memory as symm   # noqa: E402
if hasattr(symm, 'enable_symm_mem_for_group'):
    try:
        symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
        say('enable_symm_mem_for_group ok')
    except Exception as e:
        say('enable_symm_mem_for_group raised', type(e).__name__, e)
buf = symm.empty(1024, dtype=torch.float32, device='cuda')
buf.fill_(float(rank + 1))
say('symm.empty ok')
hdl = symm.rendezvous(buf, dist.group.WORLD)
say('rendezvous ok: world', hdl.world_size, 'ptrs', [hex(int(p)) for p in hdl.buffer_ptrs])
hdl.barrier(channel=0, timeout_ms=5000)
torch.cuda.synchronize()
say('barrier ok')
from fpc_diffrend_b200 import _lib              # noqa: E402
out = torch.zeros(1024, device='cuda')
tab = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
_lib.call('fpc_peer_sum', tab, world, 1024, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
say('peer_sum ->', float(out[0]), 'expected', world * (world + 1) / 2)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    hdl.barrier(channel=1, timeout_ms=5000)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    hdl.barrier(channel=1, timeout_ms=5000)
    _lib.call('fpc_peer_sum', tab, world, 1024, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
say('captured barrier + kernel')
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
say('graph replays ok ->', float(out[0]))
dist.barrier()
dist.destroy_process_group()
say('done')
