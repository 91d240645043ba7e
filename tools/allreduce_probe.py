#!/usr/bin/env python
"""Where does a multi-GPU J/K build spend the time that is not class kernels?  Under torchrun: times the
[vj|vk] all_reduce alone (same size as the Taxol build), a full build, and a build whose all_reduce is
replaced by a no-op, with CUDA events after warm-up.  usage: torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from gpu4pyscf_b200.scf import jk

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


mol = bench.load_mol('113_Taxol', 'def2-tzvp')
opt = jk._VHFOpt(mol).build()
nao = opt.sorted_mol.nao
buf = torch.zeros(2 * nao * nao, dtype=torch.float64, device=dev)
t_ar = timed(lambda: dist.all_reduce(buf), 10)
dm = torch.as_tensor(bench.decaying_density(mol), device=dev)
dms = opt.apply_coeff_C_mat_CT(dm.reshape(1, mol.nao, mol.nao)).contiguous()
t_full = timed(lambda: opt.get_jk(dms, 1), 3)
real = dist.all_reduce
dist.all_reduce = lambda *a, **k: None
torch.distributed.all_reduce = dist.all_reduce
t_noar = timed(lambda: opt.get_jk(dms, 1), 3)
torch.distributed.all_reduce = real
# the collective inside the build, bracketed by events on the launching stream, and by the host clock
import time
spans = []
def wrapped(t, *a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    r = real(t, *a, **k)
    e1.record()
    spans.append((e0, e1, time.perf_counter() - h0, t.numel()))
    return r
torch.distributed.all_reduce = wrapped
t_wr = timed(lambda: opt.get_jk(dms, 1), 3)
torch.distributed.all_reduce = real
torch.cuda.synchronize()
big = [(a.elapsed_time(b), h * 1e3) for a, b, h, n in spans if n > 1000]
fresh = []
for _ in range(5):
    x = torch.zeros(2 * nao * nao + 7, dtype=torch.float64, device=dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); real(x); e1.record(); e1.synchronize()
    fresh.append(e0.elapsed_time(e1))
if rank == 0:
    print('world %d: all_reduce of %.0f MB alone %.2f ms; build %.1f ms; build without all_reduce %.1f ms'
          % (world, buf.numel() * 8 / 1e6, t_ar, t_full, t_noar), flush=True)
    print('in-build all_reduce (device ms between events, host ms in the call):', ['%.1f/%.1f' % b for b in big])
    print('all_reduce on freshly allocated buffers (ms):', ['%.2f' % f for f in fresh])
dist.barrier()
dist.destroy_process_group()
