"""The callers on the J/K path (gpu4pyscf_b200/scf/hf.py: SCF.get_jk/get_j/get_k, RHF/UHF.get_veff with
the incremental dm_last/vhf_last build, kernel) -- reference: gpu4pyscf/scf/hf.py:91-109, 933-972,
uhf.py:386-406.  One-electron integrals come from the CPU oracle (test infrastructure); the J/K builds
are the product.  CPU: through the SIMT emulator; GPU: the golden water/cc-pVDZ energy
(gpu4pyscf/scf/tests/test_hf_lowmem.py:42) and an open-shell UHF against the oracle-J/K UHF."""
import os
import subprocess
import sys
import numpy as np
import pytest

from tests import molecules

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_WATER_CCPVDZ = -76.02676567311744

SCRIPT = r'''
import os, sys
os.environ['GPU4PYSCF_B200_EMUL'] = '1'
sys.path.insert(0, %r)
import numpy as np, torch
from gpu4pyscf_b200.csrc import build
build.build(emul=True)
from gpu4pyscf_b200.scf import hf
from oracle import oracle
from tests import molecules, scf_driver
mol = molecules.water('sto-3g')
s, t, v = oracle.int1e(mol)
mf = mol.RHF()
mf.hcore, mf.ovlp = t + v, s
e = mf.kernel()
e_ref, dm_ref, _ = scf_driver.rhf(mol, lambda dm: oracle.get_jk(mol, dm, hermi=1))
assert mf.converged and abs(e - e_ref) < 1e-9, (e, e_ref)
# get_veff: full build == incremental build from an arbitrary previous point
dm = torch.as_tensor(dm_ref)
np.random.seed(2)
dm0 = np.random.rand(*dm_ref.shape); dm0 = torch.as_tensor(dm0 + dm0.T)
v_full = mf.get_veff(mol, dm)
v_inc = mf.get_veff(mol, dm, dm_last=dm0, vhf_last=mf.get_veff(mol, dm0))
assert abs(v_full - v_inc).max() < 1e-11
rj, rk = oracle.get_jk(mol, dm_ref, hermi=1)
assert abs(v_full.numpy() - (rj - .5 * rk)).max() < 1e-10
# closed-shell UHF == RHF; UHF.get_veff against the oracle for two different spin densities
mu = mol.UHF()
mu.hcore, mu.ovlp = t + v, s
assert abs(mu.kernel() - e_ref) < 1e-9
da, db = .5 * dm_ref + .05 * np.asarray(dm0), .5 * dm_ref - .02 * np.asarray(dm0)
vu = mu.get_veff(mol, torch.as_tensor(np.stack([da, db])))
ja, ka = oracle.get_jk(mol, da, hermi=1)
jb, kb = oracle.get_jk(mol, db, hermi=1)
assert abs(vu[0].numpy() - (ja + jb - ka)).max() < 1e-10 and abs(vu[1].numpy() - (ja + jb - kb)).max() < 1e-10
# one _VHFOpt per omega is cached on the mean-field object (hf.py:936-944)
k_lr = mf.get_k(mol, dm, hermi=1, omega=0.3)
assert set(mf._opt_gpu) == {0.0, 0.3}
assert abs(k_lr.numpy() - oracle.get_jk(mol, dm_ref, hermi=1, omega=0.3)[1]).max() < 1e-10
print('OK')
''' % ROOT


def test_callers_through_emulator():
    r = subprocess.run([sys.executable, '-c', SCRIPT], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_rhf_kernel_water_ccpvdz_golden():
    from oracle import oracle
    mol = molecules.water('cc-pvdz')
    s, t, v = oracle.int1e(mol)
    mf = mol.RHF()
    mf.hcore, mf.ovlp = t + v, s
    e = mf.kernel()
    assert mf.converged and abs(e - E_WATER_CCPVDZ) < 1e-9, e


@pytest.mark.gpu
def test_uhf_kernel_water_cation_vs_oracle_uhf():
    """open shell (n_dm = 2 K pass, J of the total density): H2O+ doublet, def2-TZVP"""
    from oracle import oracle
    from tests import scf_driver
    mol = molecules.water('def2-tzvp', charge=1, spin=1)
    s, t, v = oracle.int1e(mol)
    mf = mol.UHF()
    mf.hcore, mf.ovlp = t + v, s
    mf.max_cycle = 120
    e = mf.kernel()

    def jk_cpu(dm):
        ja, ka = oracle.get_jk(mol, dm[0], hermi=1)
        jb, kb = oracle.get_jk(mol, dm[1], hermi=1)
        return np.stack([ja, jb]), np.stack([ka, kb])
    e_ref, _, _ = scf_driver.uhf(mol, jk_cpu)
    assert mf.converged and abs(e - e_ref) < 1e-9, (e, e_ref)
