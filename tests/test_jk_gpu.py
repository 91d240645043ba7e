"""GPU parity tests of the sm_100a J/K path against the CPU oracle (oracle/md_jk.c,
pinned to the reference's golden fingerprints in tests/test_oracle.py) and against
the reference tests' own fingerprints.  Tolerance: BASELINE.json north_star --
J/K within 1e-10 max-ABS (the reference tests' own tolerance is 1e-9 absolute,
test_scf_jk.py:46-47); the relative figure is printed beside it."""
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
    ej = abs(vj - rj).max() if hermi != 2 else 0.
    ek = abs(vk - rk).max()
    print('max|dJ| %.2e (rel %.1e)  max|dK| %.2e (rel %.1e)' % (
        ej, ej / max(1e-300, abs(rj).max()), ek, ek / max(1e-300, abs(rk).max())))
    assert ej < tol, ej          # absolute, as the north star states it
    assert ek < tol, ek
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


def _taxol(natoms=None, basis='def2-tzvp'):
    import os
    from gpu4pyscf_b200.gto.mole import M, read_xyz
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xyz = read_xyz(os.path.join(root, 'gpu4pyscf_b200', 'data', 'molecules', '113_Taxol.xyz'))
    if natoms:
        xyz = '\n'.join(xyz.splitlines()[:natoms])
    return M(atom=xyz, basis=basis)


def _decaying_density(mol, seed=0, beta=0.8):
    rng = np.random.RandomState(seed)
    nao = mol.nao
    ao_atom = np.repeat(mol._bas[:, 0], np.diff(mol.ao_loc_nr()))
    R = mol.atom_coords()
    dist = np.linalg.norm(R[ao_atom][:, None] - R[ao_atom][None], axis=-1)
    d = (rng.rand(nao, nao) - .5) * np.exp(-beta * dist)
    return d + d.T + np.eye(nao)


def test_qcond_matches_oracle():
    """Schwarz bounds of B200_qcond (through _VHFOpt.build) against the oracle's
    sqrt(max|(ij|ij)|); the stored value is 1/2 log of the same maximum (float32)."""
    from gpu4pyscf_b200.scf import jk
    oracle = _oracle()
    mol = molecules.water('def2-tzvp')
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    q_ref = oracle.q_cond(smol)                     # sorted, decontracted Cartesian shells
    nbas = smol.nbas
    n = 0
    for key, val in opt.bas_pair_cache.items():
        ids, q = val[0].cpu().numpy().astype(np.int64), val[1].cpu().numpy()
        ish, jsh = ids // nbas, ids % nbas
        ref = np.log(q_ref[ish, jsh])
        assert abs(q - ref).max() < 2e-5, (key, abs(q - ref).max())
        n += len(ids)
    assert n > 100


def test_fragment_with_screening_matches_oracle():
    """8-atom Taxol fragment / def2-TZVP with a decaying density, so that the Schwarz +
    density screening, the q-sorted early exit and the Y-chunk work items are all active;
    oracle evaluated WITHOUT screening.  Tolerance 1e-10 (BASELINE.json north_star)."""
    mol = _taxol(8)
    dm = _decaying_density(mol)
    _check(mol, dm, 1, tol=1e-10)


def test_full_size_properties():
    """Taxol / def2-TZVP (BASELINE.json configs[2], 2228 AOs): size-independent properties
    of J/K -- symmetry, linearity in D, the exchange of densities
    tr(D1 J[D2]) = tr(D2 J[D1]) (same for K), and invariance to the multi-GPU partition
    (two half builds, ranks 0 and 1 of 2, sum to the full build)."""
    from gpu4pyscf_b200.scf import jk
    mol = _taxol()
    nao = mol.nao
    d1 = torch.as_tensor(_decaying_density(mol, 0), device='cuda')
    d2 = torch.as_tensor(_decaying_density(mol, 1, beta=0.6), device='cuda')
    opt = jk._VHFOpt(mol).build()
    j1, k1 = jk.get_jk(mol, d1, hermi=1, vhfopt=opt)
    j2, k2 = jk.get_jk(mol, d2, hermi=1, vhfopt=opt)
    sj, sk = float(j1.abs().max()), float(k1.abs().max())
    assert float((j1 - j1.T).abs().max()) < 1e-10 * sj and float((k1 - k1.T).abs().max()) < 1e-10 * sk
    # linearity (screening is tighter than the tolerance: direct_scf_tol = 1e-13)
    j12, k12 = jk.get_jk(mol, d1 + d2, hermi=1, vhfopt=opt)
    assert float((j12 - j1 - j2).abs().max()) < 1e-9 * sj
    assert float((k12 - k1 - k2).abs().max()) < 1e-9 * sk
    # density exchange
    t12, t21 = float((d1 * j2).sum()), float((d2 * j1).sum())
    assert abs(t12 - t21) < 1e-10 * abs(t12)
    t12, t21 = float((d1 * k2).sum()), float((d2 * k1).sum())
    assert abs(t12 - t21) < 1e-10 * abs(t12)
    # partition invariance
    parts = []
    for rank in range(2):
        o = jk._VHFOpt(mol)
        o.__dict__.update({k: v for k, v in opt.__dict__.items() if k not in ('rank', 'world')})
        o.rank, o.world = rank, 2
        dms = o.apply_coeff_C_mat_CT(d1.reshape(1, nao, nao)).contiguous()
        # raw partial sums of this rank (no all_reduce: torch.distributed is not initialised)
        parts.append(o.get_jk(dms, 1))
    full = opt.get_jk(opt.apply_coeff_C_mat_CT(d1.reshape(1, nao, nao)).contiguous(), 1)
    # get_jk symmetrises per call, and symmetrisation is linear
    assert float((parts[0][0] + parts[1][0] - full[0]).abs().max()) < 1e-10 * sj
    assert float((parts[0][1] + parts[1][1] - full[1]).abs().max()) < 1e-10 * sk


@pytest.mark.parametrize('omega', [0.3, -0.3])
def test_range_separated(omega):
    """omega > 0: long-range erf(omega r)/r; omega < 0: short-range erfc(|omega| r)/r
    (libcint sign convention; reference gpu4pyscf/scf/jk.py:1122-1168, tests
    test_scf_jk.py:358-386)."""
    from gpu4pyscf_b200.scf import jk
    oracle = _oracle()
    mol = molecules.water('def2-tzvp')
    np.random.seed(11)
    dm = np.random.rand(mol.nao, mol.nao)
    dm = dm + dm.T
    # explicit omega: the operator of K only; J keeps the molecule's own 1/r (reference jk.py:98-100,128-135)
    vj, vk = jk.get_jk(mol, dm, hermi=1, omega=omega)
    rj, rk = oracle.get_jk(mol, dm, hermi=1, omega=omega)
    rj_full = oracle.get_jk(mol, dm, hermi=1, omega=0.)[0]
    assert abs(vj.cpu().numpy() - rj_full).max() < 1e-10
    assert abs(vk.cpu().numpy() - rk).max() < 1e-10
    # omega carried by the molecule (mol.with_range_coulomb in PySCF): J and K both use it
    mol_rs = mol.copy()
    mol_rs.omega = omega
    vj, vk = jk.get_jk(mol_rs, dm, hermi=1)
    assert abs(vj.cpu().numpy() - rj).max() < 1e-10
    assert abs(vk.cpu().numpy() - rk).max() < 1e-10
    # K with lr == sr at omega = 0 scales full-range K (reference _check_rsh_factors, jk.py:1155-1161)
    vj, vk = jk.get_jk(mol, dm, hermi=1, omega=0., lr_factor=.25, sr_factor=.25)
    rk_full = oracle.get_jk(mol, dm, hermi=1, omega=0.)[1]
    assert abs(vj.cpu().numpy() - rj_full).max() < 1e-10
    assert abs(vk.cpu().numpy() - .25 * rk_full).max() < 1e-10


def test_range_separated_k_factors():
    from gpu4pyscf_b200.scf import jk
    oracle = _oracle()
    mol = molecules.mol1_lmax(3)
    np.random.seed(12)
    dm = np.random.rand(mol.nao, mol.nao)
    dm = dm + dm.T
    vk = jk.get_k(mol, dm, hermi=1, omega=0.4, lr_factor=0.65, sr_factor=0.2)
    rk = (0.65 * oracle.get_jk(mol, dm, hermi=1, omega=0.4)[1]
          + 0.2 * oracle.get_jk(mol, dm, hermi=1, omega=-0.4)[1])
    assert abs(vk.cpu().numpy() - rk).max() < 1e-10


def test_rys_build_k_reference_signature():
    """RYS_build_k with the reference's argument list (jk.py:845-877): K only, range
    separation through (omega, lr_factor, sr_factor), per-group launch loop."""
    import ctypes
    import math
    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    lib = loader.load_library()
    mol = molecules.water('def2-tzvp')
    np.random.seed(8)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm + dm.T
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    dms = opt.apply_coeff_C_mat_CT(torch.as_tensor(dm, device=opt.device)[None]).contiguous()
    nbas, naos = smol.nbas, smol.nao
    dm_cond = torch.empty((nbas, nbas), dtype=torch.float32, device=opt.device)
    lib.B200_condense(jk._ptr(dm_cond), jk._ptr(dms), 1, naos, jk._ptr(opt.rys_envs._keep[3]), nbas, None)
    offs = np.append(0, np.cumsum(smol.l_ctr_counts))
    ng = len(smol.l_ctr_counts)
    bas_host = np.ascontiguousarray(smol._bas, dtype=np.int32)
    cutoff = math.log(1e-13)
    pool = torch.zeros(16, dtype=torch.int32, device=opt.device)
    omega, lr, sr = 0.4, 0.65, 0.2
    vk = torch.zeros_like(dms)
    dmax = float(dm_cond.max())
    torch.cuda.synchronize()
    for i in range(ng):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    pij, qij = opt.bas_pair_cache[i, j][:2]
                    pkl, qkl = opt.bas_pair_cache[k, l][:2]
                    shls = (ctypes.c_int * 8)(offs[i], offs[i + 1], offs[j], offs[j + 1],
                                              offs[k], offs[k + 1], offs[l], offs[l + 1])
                    err = lib.RYS_build_k(
                        jk._ptr(vk), jk._ptr(dms), 1, naos, omega, lr, sr, ctypes.byref(opt.rys_envs), shls, 0,
                        len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl),
                        jk._ptr(qij), jk._ptr(qkl), None, jk._ptr(dm_cond), cutoff, dmax, jk._ptr(pool),
                        bas_host.ctypes.data_as(ctypes.c_void_p))
                    assert err == 0
    torch.cuda.synchronize()
    vk = vk + vk.transpose(1, 2)
    ref = opt.get_k(dms, 1, omega=omega, lr_factor=lr, sr_factor=sr)
    assert abs(vk - ref).max() < 1e-10


def test_rys_build_j_reference_name_vs_oracle():
    """RYS_build_j / RYS_init_rysj_constant (reference: rys_contract_j.cu:824-956, jk.py:681-746) driven by
    the reference's per-group loop; the result (after the reference's post-processing 2*(J + J^T)) is
    compared with the CPU ORACLE's J, not with another GPU path."""
    import ctypes
    import math
    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    oracle = _oracle()
    lib = loader.load_library()
    mol = molecules.water('def2-tzvp')
    np.random.seed(3)
    nao = mol.nao
    dm = np.random.rand(nao, nao)
    dm = dm + dm.T
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    dms = opt.apply_coeff_C_mat_CT(torch.as_tensor(dm, device=opt.device)[None]).contiguous()
    nbas, naos = smol.nbas, smol.nao
    dm_cond = torch.empty((nbas, nbas), dtype=torch.float32, device=opt.device)
    lib.B200_condense(jk._ptr(dm_cond), jk._ptr(dms), 1, naos, jk._ptr(opt.rys_envs._keep[3]), nbas, None)
    offs = np.append(0, np.cumsum(smol.l_ctr_counts))
    ng = len(smol.l_ctr_counts)
    bas_host = np.ascontiguousarray(smol._bas, dtype=np.int32)
    env_host = np.ascontiguousarray(smol._env, dtype=np.float64)
    pool = torch.zeros(16, dtype=torch.int32, device=opt.device)
    vj = torch.zeros_like(dms)
    dmax = float(dm_cond.max())
    assert lib.RYS_init_rysj_constant() == 0
    torch.cuda.synchronize()
    for i in range(ng):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    pij, qij = opt.bas_pair_cache[i, j][:2]
                    pkl, qkl = opt.bas_pair_cache[k, l][:2]
                    shls = (ctypes.c_int * 8)(offs[i], offs[i + 1], offs[j], offs[j + 1],
                                              offs[k], offs[k + 1], offs[l], offs[l + 1])
                    err = lib.RYS_build_j(
                        jk._ptr(vj), jk._ptr(dms), 1, naos, ctypes.byref(opt.rys_envs), None, shls,
                        len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl),
                        jk._ptr(qij), jk._ptr(qkl), None, jk._ptr(dm_cond), math.log(1e-13), dmax, jk._ptr(pool),
                        bas_host.ctypes.data_as(ctypes.c_void_p), env_host.ctypes.data_as(ctypes.c_void_p))
                    assert err == 0
    torch.cuda.synchronize()
    vj = 2. * vj
    vj = opt.apply_coeff_CT_mat_C(vj + vj.transpose(1, 2))[0].cpu().numpy()
    rj = oracle.get_jk(mol, dm, hermi=1)[0]
    assert abs(vj - rj).max() < 1e-10
    # get_j (the public J-only entry) against the oracle as well
    vj2 = jk.get_j(mol, dm, hermi=1).cpu().numpy()
    assert abs(vj2 - rj).max() < 1e-10


def test_int2e_qcond_estimator_reference_signature():
    """int2e_qcond_estimator with the reference's argument list (nr_sr_estimator.cu:416-430) for the full-range,
    long-range and short-range operators, against the oracle's sqrt(max|(ij|ij)|) (reference tolerance:
    1e-3 log units, test_scf_jk.py:328-356; here 2e-5, float32 storage)."""
    import ctypes
    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    oracle = _oracle()
    lib = loader.load_library()
    mol = molecules.mol1_lmax(4)
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    nbas = smol.nbas
    vals = list(opt.bas_pair_cache.values())
    ids = torch.cat([v[0] for v in vals]).contiguous()
    o = np.append(0, np.cumsum([len(v[0]) for v in vals]))
    spo = np.hstack([np.arange(o[g], o[g + 1], 5) for g in range(len(vals))] + [[o[-1]]]).astype(np.int32)
    spo = torch.as_tensor(spo, device=opt.device)
    idn = ids.cpu().numpy().astype(np.int64)
    for om, lrf, srf in ((0., 1., 1.), (0.4, 1., 0.), (-0.4, 0., 1.)):
        q = torch.empty(len(ids), dtype=torch.float32, device=opt.device)
        torch.cuda.synchronize()
        assert lib.int2e_qcond_estimator(jk._ptr(q), ctypes.byref(opt.rys_envs), 0, len(spo) - 1, jk._ptr(ids),
                                         jk._ptr(spo), None, om, lrf, srf) == 0
        torch.cuda.synchronize()
        ref = np.log(oracle.q_cond(smol, omega=om)[idn // nbas, idn % nbas])
        assert abs(q.cpu().numpy() - ref).max() < 2e-5, (om, abs(q.cpu().numpy() - ref).max())


def test_pair_prescreen_is_conservative():
    """B200_pair_prescreen (overlap estimate, reference overlap.cu:2163-2223) on a 40-atom Taxol fragment:
    every pair it drops has a Schwarz bound that could not survive against the largest bound of the
    molecule at direct_scf_tol = 1e-13, i.e. the prescreen only removes what the q filter removes anyway."""
    import ctypes
    import math
    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    lib = loader.load_library()
    mol = _taxol(40)
    opt = jk._VHFOpt(mol).build()
    smol = opt.sorted_mol
    nbas = smol.nbas
    kept = torch.cat([v[0] for v in opt.bas_pair_cache.values()]).to(torch.int64)
    qmax = max(float(v[1].max()) for v in opt.bas_pair_cache.values() if len(v[1]))
    mask = torch.ones(nbas * nbas, dtype=torch.bool, device=opt.device)
    mask[kept] = False
    ii, jj = torch.meshgrid(torch.arange(nbas, device=opt.device), torch.arange(nbas, device=opt.device), indexing='ij')
    dropped = torch.nonzero(mask & (ii >= jj).reshape(-1)).reshape(-1)
    assert len(dropped) > 1000 and len(kept) > 1000
    offs = np.append(0, np.cumsum(smol.l_ctr_counts))
    grp = torch.as_tensor(np.repeat(np.arange(len(smol.l_ctr_counts)), smol.l_ctr_counts), device=opt.device)
    uniq = smol.uniq_l_ctr
    worst = -1e30
    for gi in range(len(offs) - 1):
        for gj in range(gi + 1):
            sel = dropped[(grp[dropped // nbas] == gi) & (grp[dropped % nbas] == gj)]
            if len(sel) == 0:
                continue
            ids = sel.to(torch.int32).contiguous()
            q = torch.empty(len(ids), dtype=torch.float32, device=opt.device)
            assert lib.B200_qcond(jk._ptr(q), ctypes.byref(opt.rys_envs), jk._ptr(ids), len(ids),
                                  int(uniq[gi, 0]), int(uniq[gj, 0]), int(uniq[gi, 1]), int(uniq[gj, 1]),
                                  loader.current_stream_ptr()) == 0
            worst = max(worst, float(q.max()))
    assert worst + qmax < math.log(1e-13), (worst, qmax)


def test_short_range_single_pass_with_g_shells_and_screening():
    """Short-range K in ONE build (two operator terms per quartet, the counterpart of the reference's doubled
    root set, rys_roots_for_k.cu:45-60) screened with the short-range Schwarz bounds: (a) the inline l <= 4
    molecule, where the block kernel of the (gx|gg) classes takes one launch per term; (b) an 8-atom Taxol
    fragment with a decaying density, so that the SR-specific pair lists and the density screening are active.
    Oracle evaluated without screening."""
    from gpu4pyscf_b200.scf import jk
    oracle = _oracle()
    mol = molecules.mol1_lmax(4)
    np.random.seed(21)
    dm = np.random.rand(mol.nao, mol.nao)
    dm = dm + dm.T
    vk = jk.get_k(mol, dm, hermi=1, omega=-0.3).cpu().numpy()
    rk = oracle.get_jk(mol, dm, hermi=1, omega=-0.3)[1]
    assert abs(vk - rk).max() < 1e-10

    mol = _taxol(8)
    dm = _decaying_density(mol)
    opt = jk._VHFOpt(mol).build()
    vk = jk.get_k(mol, dm, hermi=1, vhfopt=opt, omega=-0.4).cpu().numpy()
    rk = oracle.get_jk(mol, dm, hermi=1, omega=-0.4)[1]
    assert abs(vk - rk).max() < 1e-10
    # the SR lists are screened with the SR bounds: never more pairs than the full-range lists
    full = sum(len(v[0]) for v in opt._lists[((1., 0.),)].cache.values())
    sr = sum(len(v[0]) for v in opt._lists[((1., 0.), (-1., 0.4))].cache.values())
    assert 0 < sr <= full
    # and bas_pair_cache keeps pointing at the molecule's own operator
    assert sum(len(v[0]) for v in opt.bas_pair_cache.values()) == full
    # mixed factors (omega-B97-type: lr 1.0, sr 0.157 -> sr*1/r + (lr - sr)*erf): one pass as well
    vk = jk.get_k(mol, dm, hermi=1, vhfopt=opt, omega=0.3, lr_factor=1., sr_factor=0.157706).cpu().numpy()
    rk = (oracle.get_jk(mol, dm, hermi=1, omega=0.3)[1] + 0.157706 * oracle.get_jk(mol, dm, hermi=1, omega=-0.3)[1])
    assert abs(vk - rk).max() < 1e-10
