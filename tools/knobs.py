#!/usr/bin/env python
"""Time one Taxol J/K build for a few values of a host-side knob (environment variable).
usage: python tools/knobs.py B200_STREAMS 1 4 8"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch
sys.path.insert(0, %r)
import bench
from gpu4pyscf_b200.scf import jk
mol = bench.load_mol('113_Taxol', 'def2-tzvp')
dm = torch.as_tensor(bench.decaying_density(mol), device='cuda')
opt = jk._VHFOpt(mol).build()
for _ in range(2):
    jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); jk.get_jk(mol, dm, hermi=1, vhfopt=opt); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print('%%.1f ms' %% min(ts))
''' % ROOT
knob = sys.argv[1]
for v in sys.argv[2:]:
    env = dict(os.environ)
    env[knob] = v
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(knob, '=', v, ':', r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
