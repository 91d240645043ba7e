"""GPU J/K against golden CPU fixtures for systems the oracle needs minutes to hours for
(tests/golden/jk_*.npz, written by tests/golden/make_golden_jk.py from the CPU oracle without
any screening): PySCF-style fingerprints, 20000 sampled matrix elements and J @ V, K @ V --
all at the north star's 1e-10 absolute."""
import glob
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, 'golden', 'jk_*.npz')))


def _parse(path):
    """jk_<mol>_<basis>[_<natoms>].npz -> (mol, basis, natoms); <mol> may contain underscores"""
    parts = os.path.basename(path)[3:-4].split('_')
    natoms = None
    if parts[-1].isdigit() and not parts[-2].isdigit():      # trailing atom count after the basis name
        natoms = int(parts.pop())
    basis = parts.pop()
    return '_'.join(parts), basis, natoms


@pytest.mark.parametrize('path', FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_jk_matches_golden_fixture(path):
    import torch
    import bench
    from gpu4pyscf_b200.scf import jk
    from oracle import oracle
    g = np.load(path)
    name, basis, natoms = _parse(path)
    mol = bench.load_mol(name, basis, natoms)
    assert mol.nao == int(g['nao'])
    dm = bench.decaying_density(mol)
    vj, vk = jk.get_jk(mol, torch.as_tensor(dm, device='cuda'), hermi=1)
    vj, vk = vj.cpu().numpy(), vk.cpu().numpy()
    idx = g['idx']
    ej = abs(vj.ravel()[idx] - g['j_idx']).max()
    ek = abs(vk.ravel()[idx] - g['k_idx']).max()
    print('%s: max|dJ| %.2e max|dK| %.2e over %d sampled elements (max|J| %.1f max|K| %.1f)'
          % (os.path.basename(path), ej, ek, len(idx), float(g['max_j']), float(g['max_k'])))
    assert ej < 1e-10 and ek < 1e-10
    # every element enters the contractions with four fixed vectors and the fingerprints
    nv = np.abs(g['vec']).sum(axis=0).max()
    assert abs(vj @ g['vec'] - g['j_vec']).max() < 1e-10 * nv
    assert abs(vk @ g['vec'] - g['k_vec']).max() < 1e-10 * nv
    assert abs(oracle.fp(vj) - float(g['fp_j'])) < 1e-10 * np.sqrt(vj.size)
    assert abs(oracle.fp(vk) - float(g['fp_k'])) < 1e-10 * np.sqrt(vk.size)


def test_fixtures_present():
    assert FIXTURES, 'no golden J/K fixture under tests/golden'
