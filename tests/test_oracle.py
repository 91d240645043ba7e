"""Pin the CPU oracle (oracle/md_jk.c, McMurchie-Davidson) against the golden
fingerprints the reference's own tests assert for this path."""
import numpy as np
import pytest
from oracle import oracle
from tests import molecules


def test_fp_definition():
    a = np.arange(6.).reshape(2, 3)
    assert abs(oracle.fp(a) - np.dot(np.cos(np.arange(6)), a.ravel())) < 1e-14


def test_golden_inline_basis_hermi1():
    # gpu4pyscf/scf/tests/test_rhf.py:166-179 (l up to 5, no basis file needed)
    mol = molecules.mol1()
    np.random.seed(1)
    nao = mol.nao
    dm = np.random.random((2, nao, nao))
    dm = dm + dm.transpose(0, 2, 1)
    vj, vk = oracle.get_jk(mol, dm, hermi=1)
    assert abs(oracle.fp(vj) - 179.14526555375858) < 1e-7
    assert abs(oracle.fp(vk) - -34.851182918643005) < 1e-7


def test_golden_inline_basis_hermi0():
    # gpu4pyscf/scf/tests/test_rhf.py:181-194
    mol = molecules.mol1()
    np.random.seed(1)
    nao = mol.nao
    dm = np.random.random((2, nao, nao))
    vj, vk = oracle.get_jk(mol, dm, hermi=0)
    assert abs(oracle.fp(vj) - 89.57263277687994) < 1e-7
    assert abs(oracle.fp(vk) - -26.36969769724246) < 1e-7


def test_golden_def2tzvp_hermi1():
    # gpu4pyscf/scf/tests/test_scf_jk.py:24-47 -- also pins the embedded def2-TZVP H,C,O
    mol = molecules.jkmol()
    np.random.seed(9)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm.dot(dm.T)
    vj, vk = oracle.get_jk(mol, dm, hermi=1)
    assert abs(oracle.fp(vj) - -2327.4715195591784) < 1e-8
    assert abs(oracle.fp(vk) - -4069.3170008260583) < 1e-8


def test_golden_def2tzvp_cart():
    # gpu4pyscf/scf/tests/test_scf_jk.py:82-114
    mol = molecules.jkmol(cart=True)
    np.random.seed(9)
    nao = mol.nao
    dm = np.random.rand(nao, nao) * .1 - .03
    dm = dm.dot(dm.T)
    vj, vk = oracle.get_jk(mol, dm, hermi=1)
    assert abs(oracle.fp(vj) - 88.88500592206657) < 1e-8
    assert abs(oracle.fp(vk) - 48.57434458906684) < 1e-8


def test_screened_matches_unscreened():
    mol = molecules.water('def2-tzvp')
    np.random.seed(3)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm + dm.T
    vj0, vk0 = oracle.get_jk(mol, dm, hermi=1)
    vj1, vk1 = oracle.get_jk(mol, dm, hermi=1, screen=True, tol=1e-13)
    assert abs(vj0 - vj1).max() < 1e-10
    assert abs(vk0 - vk1).max() < 1e-10
