"""world_size-2 run of the multi-GPU path on CPU: two processes, gloo backend,
kernels through the SIMT emulator (tests/emul).  Checks that the rank partition
of the X-pair lists plus ONE all_reduce of the [J|K] partials reproduces the
single-rank result and the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
os.environ['GPU4PYSCF_B200_EMUL'] = '1'
sys.path.insert(0, %r)
import numpy as np
import torch
import torch.distributed as dist
from gpu4pyscf_b200.csrc import build
rank = int(os.environ['RANK'])
if rank == 0:
    build.build(emul=True)
dist.init_process_group('gloo')
dist.barrier()
from gpu4pyscf_b200.scf import jk
from tests import molecules
mol = molecules.mol1_lmax(2)
np.random.seed(3)
nao = mol.nao
dm = np.random.rand(nao, nao)
dm = dm + dm.T
opt = jk._VHFOpt(mol).build()
# first build: work items of every launch dealt over the 2 ranks, launch times measured and reduced with [J|K]
vj0, vk0 = jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
assert opt._launch_cost is not None and len(opt._launch_cost) == len(opt.class_tasks())
owners = jk._lpt_owners(opt._launch_cost, 2)
assert set(owners) == {0, 1}                               # both ranks own launches
# second build: light launches whole on one rank (longest first), heavy ones item-dealt; same single all_reduce
os.environ['B200_SPLIT_MIN_MS'] = str(float(np.median(opt._launch_cost)))
owners = jk._lpt_owners(opt._launch_cost, 2, float(os.environ['B200_SPLIT_MIN_MS']))
assert -1 in owners and 0 in owners and 1 in owners
vj, vk = jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
assert abs(vj - vj0).max() < 1e-11 and abs(vk - vk0).max() < 1e-11
# every launch whole on one rank
os.environ['B200_PARTITION'] = 'cost'
vj, vk = jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
assert abs(vj - vj0).max() < 1e-11 and abs(vk - vk0).max() < 1e-11
opt1 = jk._VHFOpt(mol).build()
opt1.rank, opt1.world = 0, 1                               # same build on one rank, no collective
vj1, vk1 = jk.get_jk(mol, dm, hermi=1, vhfopt=opt1)
assert abs(vj - vj1).max() < 1e-11 and abs(vk - vk1).max() < 1e-11
if rank == 0:
    from oracle import oracle
    rj, rk = oracle.get_jk(mol, dm, hermi=1)
    assert abs(vj.numpy() - rj).max() < 1e-10 and abs(vk.numpy() - rk).max() < 1e-10
# short-range K (two operator terms in one pass, its own screening lists and launch costs) over 2 ranks
os.environ['B200_PARTITION'] = 'hybrid'
ks = jk.get_k(mol, dm, hermi=1, vhfopt=opt, omega=-0.4)      # measuring build of the SR lists
ks2 = jk.get_k(mol, dm, hermi=1, vhfopt=opt, omega=-0.4)     # partitioned by cost
ks1 = jk.get_k(mol, dm, hermi=1, vhfopt=opt1, omega=-0.4)
assert abs(ks - ks1).max() < 1e-11 and abs(ks2 - ks1).max() < 1e-11
# every rank must hold the full result
t = torch.stack([vj.abs().sum(), vk.abs().sum()])
ts = [torch.zeros_like(t) for _ in range(2)]
dist.all_gather(ts, t)
assert torch.allclose(ts[0], ts[1], rtol=1e-13)
dist.barrier()
dist.destroy_process_group()
print('OK', rank)
''' % ROOT


def test_two_rank_partition_and_allreduce():
    procs = []
    for rank in range(2):
        env = dict(os.environ)
        env.update(RANK=str(rank), WORLD_SIZE='2', LOCAL_RANK=str(rank),
                   MASTER_ADDR='127.0.0.1', MASTER_PORT='29417')
        procs.append(subprocess.Popen([sys.executable, '-c', SCRIPT], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=1500) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and 'OK' in o, o[-2000:] + e[-3000:]
