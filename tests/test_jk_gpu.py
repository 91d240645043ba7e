"""GPU parity tests of the sm_100a J/K path against the CPU oracle (oracle/md_jk.c,
pinned to the reference's golden fingerprints in tests/test_oracle.py) and against
the reference tests' own fingerprints.  Tolerance: BASELINE.json north_star --
J/K within 1e-10 max-abs (scaled by max|J| for the random test densities of the
reference tests, whose own tolerance is 1e-9 absolute, test_scf_jk.py:46-47)."""
import numpy as np
import pytest
import torch

from tests import molecules

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import oracle
    return oracle


def _check(mol, dm, hermi, tol=1e-10):
    from gpu4pyscf_b200.scf import jk
    oracle = _oracle()
    vj, vk = jk.get_jk(mol, dm, hermi=hermi)
    assert vj.is_cuda and vk.is_cuda
    rj, rk = oracle.get_jk(mol, dm, hermi=0)
    vj, vk = vj.cpu().numpy(), vk.cpu().numpy()
    sj = max(1., abs(rj).max())
    sk = max(1., abs(rk).max())
    if hermi != 2:
        assert abs(vj - rj).max() / sj < tol
    assert abs(vk - rk).max() / sk < tol
    return vj, vk


def test_library_is_cuda():
    from gpu4pyscf_b200.lib import loader
    lib = loader.load_library()
    assert not loader.emulated()
    assert lib.cuda_version() >= 12000
    assert lib.RYS_b200_lmax() == 4


def test_water_sto3g():
    mol = molecules.water('sto-3g')
    np.random.seed(1)
    dm = np.random.rand(mol.nao, mol.nao)
    _check(mol, dm + dm.T, 1)


def test_golden_def2tzvp_hermi1():
    # gpu4pyscf/scf/tests/test_scf_jk.py:24-47
    oracle = _oracle()
    mol = molecules.jkmol()
    np.random.seed(9)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm.dot(dm.T)
    vj, vk = _check(mol, dm, 1)
    assert abs(oracle.fp(vj) - -2327.4715195591784) < 1e-9
    assert abs(oracle.fp(vk) - -4069.3170008260583) < 1e-9


def test_golden_def2tzvp_cart():
    # gpu4pyscf/scf/tests/test_scf_jk.py:82-108
    oracle = _oracle()
    mol = molecules.jkmol(cart=True)
    np.random.seed(9)
    nao = mol.nao
    dm = np.random.rand(nao, nao) * .1 - .03
    dm = dm.dot(dm.T)
    vj, vk = _check(mol, dm, 1)
    assert abs(oracle.fp(vj) - 88.88500592206657) < 1e-10
    assert abs(oracle.fp(vk) - 48.57434458906684) < 1e-10


@pytest.mark.parametrize('hermi', [0, 1, 2])
def test_inline_basis_up_to_g(hermi):
    # molecule of gpu4pyscf/scf/tests/test_rhf.py:41-74 without its h shell
    # (no CPU fallback here; l <= 4 on the GPU), two density matrices
    mol = molecules.mol1_lmax(4)
    np.random.seed(1)
    nao = mol.nao
    dm = np.random.random((2, nao, nao))
    if hermi == 1:
        dm = dm + dm.transpose(0, 2, 1)
    elif hermi == 2:
        dm = dm - dm.transpose(0, 2, 1)
    _check(mol, dm, hermi)


def test_three_density_matrices():
    mol = molecules.water('def2-tzvp')
    np.random.seed(5)
    nao = mol.nao
    dm = np.random.random((3, nao, nao))
    dm = dm + dm.transpose(0, 2, 1)
    _check(mol, dm, 1)


def test_l5_raises():
    from gpu4pyscf_b200.scf import jk
    mol = molecules.mol1()
    with pytest.raises(NotImplementedError):
        jk._VHFOpt(mol).build()


def test_j_only_k_only():
    from gpu4pyscf_b200.scf import jk
    mol = molecules.water('def2-tzvp')
    np.random.seed(2)
    dm = np.random.rand(mol.nao, mol.nao)
    dm = dm + dm.T
    vj, vk = jk.get_jk(mol, dm, hermi=1)
    vj1 = jk.get_j(mol, dm, hermi=1)
    vk1 = jk.get_k(mol, dm, hermi=1)
    assert abs(vj - vj1).max() < 1e-11
    assert abs(vk - vk1).max() < 1e-11


def test_reference_signature_entry():
    """RYS_build_jk with the reference's argument list (jk.py:564-583): per
    (i,j,k,l) group launch, pair-id de-duplication, binned-order margin."""
    import ctypes
    import math
    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    lib = loader.load_library()
    mol = molecules.water('def2-tzvp')
    np.random.seed(7)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm + dm.T
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    dms = opt.apply_coeff_C_mat_CT(torch.as_tensor(dm, device=opt.device)[None]).contiguous()
    ref_vj, ref_vk = opt.get_jk(dms, 1)
    nbas = smol.nbas
    naos = smol.nao
    dm_cond = torch.empty((nbas, nbas), dtype=torch.float32, device=opt.device)
    lib.B200_condense(jk._ptr(dm_cond), jk._ptr(dms), 1, naos, jk._ptr(opt.rys_envs._keep[3]), nbas, None)
    vj = torch.zeros_like(dms)
    vk = torch.zeros_like(dms)
    pool = torch.zeros(16, dtype=torch.int32, device=opt.device)
    offs = np.append(0, np.cumsum(smol.l_ctr_counts))
    ng = len(smol.l_ctr_counts)
    bas_host = np.ascontiguousarray(smol._bas, dtype=np.int32)
    env_host = np.ascontiguousarray(smol._env)
    cutoff = math.log(1e-13)
    torch.cuda.synchronize()
    for i in range(ng):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    pij, qij = opt.bas_pair_cache[i, j][:2]
                    pkl, qkl = opt.bas_pair_cache[k, l][:2]
                    shls = (ctypes.c_int * 8)(offs[i], offs[i + 1], offs[j], offs[j + 1],
                                              offs[k], offs[k + 1], offs[l], offs[l + 1])
                    err = lib.RYS_build_jk(
                        jk._ptr(vj), jk._ptr(vk), jk._ptr(dms), 1, naos, ctypes.byref(opt.rys_envs),
                        shls, 0, len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl),
                        jk._ptr(qij), jk._ptr(qkl), jk._ptr(qij), jk._ptr(qkl), None,
                        jk._ptr(dm_cond), cutoff, float(dm_cond.max()), jk._ptr(pool),
                        bas_host.ctypes.data_as(ctypes.c_void_p), env_host.ctypes.data_as(ctypes.c_void_p))
                    assert err == 0
    torch.cuda.synchronize()
    vj = vj * 2
    vj = vj + vj.transpose(1, 2)
    vk = vk + vk.transpose(1, 2)
    assert abs(vj - ref_vj).max() < 1e-10
    assert abs(vk - ref_vk).max() < 1e-10
