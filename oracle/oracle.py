"""ctypes driver for the CPU oracle liboracle_jk.so (McMurchie-Davidson J/K).

Restates ``pyscf.scf.hf.get_jk(mol, dm, hermi)`` -- the function the reference
tests compare against (gpu4pyscf/scf/tests/test_scf_jk.py:22,45) -- for the
repo's own ``Mole``.  TEST INFRASTRUCTURE ONLY.
"""
import os
import ctypes
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'liboracle_jk.so')
    src = os.path.join(_HERE, 'md_jk.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_jk.restype = ctypes.c_long
        _LIB.oracle_last_flops.restype = ctypes.c_double
    return _LIB


def fp(a):
    """PySCF's lib.fp fingerprint."""
    a = np.asarray(a)
    return float(np.dot(np.cos(np.arange(a.size)), a.ravel()))


def _c2s_matrix(mol):
    """(nao_cart, nao) block-diagonal cart->AO map for the *original* mol."""
    from gpu4pyscf_b200.gto.cart2sph import cart2sph, ncart
    ls = mol._bas[:, 1]
    nctr = mol._bas[:, 3]
    nc = int(sum(ncart(l) * c for l, c in zip(ls, nctr)))
    C = np.zeros((nc, mol.nao))
    p = q = 0
    for l, c in zip(ls, nctr):
        m = np.eye(ncart(l)) if mol.cart else cart2sph(int(l))
        for _ in range(c):
            C[p:p + m.shape[0], q:q + m.shape[1]] = m
            p += m.shape[0]
            q += m.shape[1]
    return C


def get_jk(mol, dm, hermi=0, omega=None, screen=False, tol=1e-13,
           with_j=True, with_k=True, return_count=False):
    """J_ij = sum_kl (ij|kl) D_lk, K_il = sum_jk (ij|kl) D_jk in mol's AO basis."""
    lib = _lib()
    if omega is None:
        omega = mol.omega
    dm = np.asarray(dm, dtype=float)
    shape = dm.shape
    nao = shape[-1]
    dms = dm.reshape(-1, nao, nao)
    C = _c2s_matrix(mol)
    dmc = np.ascontiguousarray(np.einsum('pi,nij,qj->npq', C, dms, C))
    nc = C.shape[0]
    vj = np.zeros_like(dmc)
    vk = np.zeros_like(dmc)
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.ascontiguousarray(mol._env, dtype=np.float64)
    n = lib.oracle_jk(
        vj.ctypes.data_as(ctypes.c_void_p), vk.ctypes.data_as(ctypes.c_void_p),
        dmc.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(dmc)), ctypes.c_int(nc),
        atm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(atm)),
        bas.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(bas)),
        env.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(omega),
        ctypes.c_int(int(screen)), ctypes.c_double(tol),
        ctypes.c_int(int(with_j)), ctypes.c_int(int(with_k)))
    if n < 0:
        raise RuntimeError('oracle_jk failed')
    vj = np.einsum('pi,npq,qj->nij', C, vj, C).reshape(shape)
    vk = np.einsum('pi,npq,qj->nij', C, vk, C).reshape(shape)
    if return_count:
        return vj, vk, n
    return vj, vk


def last_flops():
    """Algorithmic flops (SURVEY.md 8d model) of the quartets the last get_jk call evaluated."""
    return float(_lib().oracle_last_flops())


def q_cond(mol, omega=0.):
    """sqrt(max|(ij|ij)|) per shell pair of mol (Cartesian components)."""
    lib = _lib()
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.ascontiguousarray(mol._env, dtype=np.float64)
    nbas = len(bas)
    q = np.zeros((nbas, nbas))
    lib.oracle_q_cond(q.ctypes.data_as(ctypes.c_void_p),
                      atm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(atm)),
                      bas.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nbas),
                      env.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(omega))
    return q


def int1e(mol):
    """(S, T, V): overlap, kinetic and nuclear-attraction matrices in mol's AO basis -- what
    mol.intor('int1e_ovlp' / 'int1e_kin' / 'int1e_nuc') returns in PySCF.  Used only by the
    test-side SCF loop (tests/scf_driver.py)."""
    lib = _lib()
    C = _c2s_matrix(mol)
    nc = C.shape[0]
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.ascontiguousarray(mol._env, dtype=np.float64)
    mats = [np.zeros((nc, nc)) for _ in range(3)]
    lib.oracle_int1e(*[m.ctypes.data_as(ctypes.c_void_p) for m in mats], ctypes.c_int(nc),
                     atm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(atm)),
                     bas.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(bas)),
                     env.ctypes.data_as(ctypes.c_void_p))
    return tuple(C.T @ m @ C for m in mats)


def energy_nuc(mol):
    """nuclear repulsion energy (Hartree)"""
    z = np.asarray(mol._atm[:, 0], dtype=float)
    r = mol.atom_coords()
    e = 0.
    for i in range(len(z)):
        for j in range(i):
            e += z[i] * z[j] / np.linalg.norm(r[i] - r[j])
    return e
