"""TEST INFRASTRUCTURE: a minimal restricted Hartree-Fock loop (DIIS) around a pluggable
``get_jk``.  The product replaces the J/K build of gpu4pyscf's SCF drivers, not the drivers
themselves (gpu4pyscf/scf/hf.py:206-342 stay the reference's); this loop exists so that the
tests can assert the north star's second parity criterion -- the SCF total energy obtained with
the GPU J/K build agrees with the one obtained with the CPU oracle's J/K (and with the
reference's golden energies) -- without the reference's drivers.  One-electron integrals come
from the CPU oracle (oracle.int1e)."""
import numpy as np
from oracle import oracle


def rhf(mol, get_jk, conv_tol=1e-11, max_cycle=80, verbose=False):
    '''get_jk(dm) -> (vj, vk) numpy arrays for a symmetric density matrix.
    Returns (e_tot, dm, n_iter).'''
    s, t, v = oracle.int1e(mol)
    h = t + v
    enuc = oracle.energy_nuc(mol)
    nocc = int(mol._atm[:, 0].sum() - getattr(mol, 'charge', 0)) // 2
    se, sv = np.linalg.eigh(s)
    x = sv / np.sqrt(se)                 # symmetric orthogonalisation (no linear dependencies here)

    def diag(f):
        e, c = np.linalg.eigh(x.T @ f @ x)
        c = x @ c
        return e, c

    e, c = diag(h)
    dm = 2. * c[:, :nocc] @ c[:, :nocc].T
    fs, es = [], []
    e_last = 0.
    for it in range(max_cycle):
        vj, vk = get_jk(dm)
        f = h + vj - .5 * vk
        e_tot = .5 * np.sum(dm * (h + f)) + enuc
        err = x.T @ (f @ dm @ s - s @ dm @ f) @ x
        fs.append(f)
        es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:
            n = len(fs)
            b = np.zeros((n + 1, n + 1))
            b[-1, :n] = b[:n, -1] = 1.
            for i in range(n):
                for j in range(n):
                    b[i, j] = np.sum(es[i] * es[j])
            rhs = np.zeros(n + 1)
            rhs[-1] = 1.
            try:
                coef = np.linalg.solve(b, rhs)[:n]
                f = sum(ci * fi for ci, fi in zip(coef, fs))
            except np.linalg.LinAlgError:
                pass
        if verbose:
            print('cycle %d  E = %.12f  dE = %.2e  |err| = %.2e' % (it, e_tot, e_tot - e_last, np.abs(err).max()))
        if it > 0 and abs(e_tot - e_last) < conv_tol and np.abs(err).max() < 1e-6:
            return e_tot, dm, it + 1
        e_last = e_tot
        e, c = diag(f)
        dm = 2. * c[:, :nocc] @ c[:, :nocc].T
    raise RuntimeError('SCF not converged')


def uhf(mol, get_jk, conv_tol=1e-11, max_cycle=120):
    '''Unrestricted counterpart of rhf(): get_jk(dm) is called with dm of shape (2, nao, nao) and must
    return (vj, vk) of the same shape.  Returns (e_tot, dm, n_iter).'''
    s, t, v = oracle.int1e(mol)
    h = t + v
    enuc = oracle.energy_nuc(mol)
    na, nb = mol.nelec
    se, sv = np.linalg.eigh(s)
    x = sv / np.sqrt(se)

    def new_dm(f):
        out = []
        for k, n in enumerate((na, nb)):
            e, c = np.linalg.eigh(x.T @ f[k] @ x)
            c = x @ c
            out.append(c[:, :n] @ c[:, :n].T)
        return np.stack(out)

    dm = new_dm(np.stack([h, h]))
    fs, es = [], []
    e_last = 0.
    for it in range(max_cycle):
        vj, vk = get_jk(dm)
        f = h[None] + (vj[0] + vj[1])[None] - vk
        e_tot = .5 * np.sum(dm * (h[None] + f)) + enuc
        err = np.stack([x.T @ (f[k] @ dm[k] @ s - s @ dm[k] @ f[k]) @ x for k in range(2)])
        fs.append(f)
        es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:
            n = len(fs)
            b = np.zeros((n + 1, n + 1))
            b[-1, :n] = b[:n, -1] = 1.
            for i in range(n):
                for j in range(n):
                    b[i, j] = np.sum(es[i] * es[j])
            rhs = np.zeros(n + 1)
            rhs[-1] = 1.
            try:
                coef = np.linalg.solve(b, rhs)[:n]
                f = sum(ci * fi for ci, fi in zip(coef, fs))
            except np.linalg.LinAlgError:
                pass
        if it > 0 and abs(e_tot - e_last) < conv_tol and np.abs(err).max() < 1e-6:
            return e_tot, dm, it + 1
        e_last = e_tot
        dm = new_dm(f)
    raise RuntimeError('UHF not converged')
