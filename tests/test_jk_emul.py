"""CPU-only checks of the CUDA kernel SOURCES through the SIMT emulator
(tests/emul): the same .cu/.cuh files compiled with g++, every CUDA thread a
fiber.  This is a debugging aid for a GPU-less container, not a product path."""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
os.environ['GPU4PYSCF_B200_EMUL'] = '1'
sys.path.insert(0, %r)
import numpy as np
from gpu4pyscf_b200.csrc import build
build.build(emul=True)
from gpu4pyscf_b200.scf import jk
from oracle import oracle
from tests import molecules
case, hermi = sys.argv[1], int(sys.argv[2])
if case == 'sto':
    mol = molecules.water('sto-3g')
elif case == 'tz':
    mol = molecules.water('def2-tzvp')      # contracted p shells: multi-pass primitive batches
else:
    mol = molecules.mol1_lmax(int(case))
np.random.seed(1)
nao = mol.nao
dm = np.random.rand(2, nao, nao)
if hermi == 1:
    dm = dm + dm.transpose(0, 2, 1)
elif hermi == 2:
    dm = dm - dm.transpose(0, 2, 1)
omega = float(os.environ.get('TEST_OMEGA', '0'))
vj, vk = jk.get_jk(mol, dm, hermi=hermi, omega=omega)
rj, rk = oracle.get_jk(mol, dm, hermi=0, omega=omega)
if omega != 0:      # omega shapes the exchange operator only (reference jk.py:128-135); J stays 1/r
    rj = oracle.get_jk(mol, dm, hermi=0)[0]
if hermi != 2:
    assert abs(vj.numpy() - rj).max() < 1e-10, abs(vj.numpy() - rj).max()
assert abs(vk.numpy() - rk).max() < 1e-10, abs(vk.numpy() - rk).max()
print('OK')
''' % ROOT


@pytest.mark.parametrize('case,hermi,ychunk', [('sto', 1, 0), ('sto', 0, 3), ('sto', 2, 0), ('2', 1, 7), ('3', 1, 0), ('tz', 1, 5), ('rs', 1, 0)])
def test_emulated_kernels_match_oracle(case, hermi, ychunk):
    env = dict(os.environ)
    if case == 'rs':
        case = 'tz'
        env['TEST_OMEGA'] = '-0.35'             # short range = full - long range (two builds)
    if ychunk:
        env['B200_YCHUNK'] = str(ychunk)      # tiny Y chunks: many (X pair, Y chunk) work items
    r = subprocess.run([sys.executable, '-c', SCRIPT, case, str(hermi)], env=env,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


RS_SCRIPT = r'''
import os, sys
os.environ['GPU4PYSCF_B200_EMUL'] = '1'
sys.path.insert(0, %r)
import numpy as np
from gpu4pyscf_b200.csrc import build
build.build(emul=True)
from gpu4pyscf_b200.scf import jk
from oracle import oracle
from tests import molecules
mol = molecules.mol1_lmax(2)
np.random.seed(5)
nao = mol.nao
for hermi in (0, 2):
    dm = np.random.rand(2, nao, nao)
    if hermi == 2:
        dm = dm - dm.transpose(0, 2, 1)
    for om, lr, sr in ((-0.35, None, None), (0.3, 1.0, 0.2)):
        vk = jk.get_k(mol, dm, hermi=hermi, omega=om, lr_factor=lr, sr_factor=sr).numpy()
        if lr is None:
            rk = oracle.get_jk(mol, dm, hermi=0, omega=om)[1]
        else:
            rk = (lr * oracle.get_jk(mol, dm, hermi=0, omega=abs(om))[1]
                  + sr * oracle.get_jk(mol, dm, hermi=0, omega=-abs(om))[1])
        assert abs(vk - rk).max() < 1e-10, (hermi, om, abs(vk - rk).max())
print('OK')
''' % ROOT


def test_range_separated_single_pass_hermi0_hermi2():
    """two operator terms per quartet (one build) for non-symmetric and anti-symmetric densities, n_dm = 2:
    pure short range and an lr/sr mix, against the oracle"""
    r = subprocess.run([sys.executable, '-c', RS_SCRIPT], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
