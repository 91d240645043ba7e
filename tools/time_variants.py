import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import load_mol, decaying_density
from gpu4pyscf_b200.scf import jk
mol = load_mol('113_Taxol', 'def2-tzvp')
d1 = torch.as_tensor(decaying_density(mol, 0), device='cuda')
d2 = torch.as_tensor(decaying_density(mol, 1), device='cuda')
opt = jk._VHFOpt(mol).build()
for dm, name in ((d1, 'n_dm=1'), (torch.stack([d1, d2]), 'n_dm=2 hermi=1'), (d1 + 0.1 * torch.triu(d2), 'n_dm=1 hermi=0')):
    hermi = 0 if 'hermi=0' in name else 1
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        vj, vk = jk.get_jk(mol, dm, hermi=hermi, vhfopt=opt)
        e1.record(); e1.synchronize()
    print(name, '%.1f ms' % e0.elapsed_time(e1), flush=True)
vk1 = jk.get_k(mol, d1, hermi=1, vhfopt=opt)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); vk1 = jk.get_k(mol, d1, hermi=1, vhfopt=opt); e1.record(); e1.synchronize()
print('K only %.1f ms' % e0.elapsed_time(e1))
e0.record(); vj1 = jk.get_j(mol, d1, hermi=1, vhfopt=opt); e1.record(); e1.synchronize()
print('J only %.1f ms' % e0.elapsed_time(e1))
