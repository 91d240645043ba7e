"""Minimal molecule/basis host layer + the sorted, decontracted shell view the
J/K kernels consume.

PySCF is not installed in this image, so ``Mole`` restates the parts of
``pyscf.gto.Mole`` that the reference's J/K plugin reads: the libcint-layout
``_atm/_bas/_env`` tables (slot layout mirrored in the reference at
lib/gvhf-rys/vhf.cuh:5-25), ``ao_loc_nr``, ``nao``, ``cart``, ``omega``.
``SortedGTO`` mirrors gpu4pyscf/gto/mole.py:583-1082 (``SortedGTO.from_mol``,
``apply_C_mat_CT`` / ``apply_CT_mat_C``, ``RysIntEnvVars``): general
contractions are split into nctr=1 shells, primitives with exponent <=
``diffuse_cutoff`` become their own shells, shells are sorted by (l ascending,
nprim descending) and a coefficient map back to the original AO basis is kept.
"""
import ctypes
import contextlib
import numpy as np
from . import basis as _basis
from .cart2sph import cart2sph, ncart

BOHR = 0.52917721092

# libcint slots
CHARGE_OF, PTR_COORD, NUC_MOD_OF, PTR_ZETA, PTR_FRAC_CHARGE = 0, 1, 2, 3, 4
ATM_SLOTS = 6
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF = 0, 1, 2, 3, 4, 5, 6
PTR_BAS_COORD = 7
BAS_SLOTS = 8
PTR_RANGE_OMEGA = 8
PTR_ENV_START = 20


def _parse_atoms(atom, unit):
    if isinstance(atom, str):
        rows = []
        for ln in atom.replace(';', '\n').splitlines():
            ln = ln.strip().replace(',', ' ')
            if not ln:
                continue
            t = ln.split()
            rows.append((t[0], [float(t[1]), float(t[2]), float(t[3])]))
    else:
        rows = [(a[0], [float(x) for x in (a[1] if len(a) == 2 else a[1:4])]) for a in atom]
    u = unit.upper()
    scale = 1.0 if (u.startswith('B') or u == 'AU') else 1.0 / BOHR
    return [(s, np.asarray(c) * scale) for s, c in rows]


def read_xyz(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[0].split()[0])
    return '\n'.join(lines[2:2 + n])


class Mole:
    """Subset of pyscf.gto.Mole sufficient for the J/K plugin and a minimal SCF."""

    def __init__(self, atom=None, basis='sto-3g', unit='Angstrom', cart=False,
                 charge=0, spin=0, verbose=0, output=None, **kw):
        self.atom = atom
        self.basis = basis
        self.unit = unit
        self.cart = cart
        self.charge = charge
        self.spin = 0 if spin is None else spin
        self.verbose = verbose
        self.stdout = None
        self._atm = self._bas = self._env = None
        self._built = False

    # ---- construction -------------------------------------------------
    def build(self):
        atoms = _parse_atoms(self.atom, self.unit)
        self._atom = atoms
        env = [np.zeros(PTR_ENV_START)]
        ptr = PTR_ENV_START
        atm = []
        for sym, xyz in atoms:
            atm.append([_basis.charge_of(sym), ptr, 1, ptr + 3, 0, 0])
            env.append(np.append(xyz, 0.))
            ptr += 4
        # per-element basis templates
        templates = {}
        bas = []
        for ia, (sym, _) in enumerate(atoms):
            key = _basis.pure_symbol(sym)
            if key not in templates:
                shells = self._shells_for(sym)
                tmpl = []
                for sh in shells:
                    l = sh[0]
                    rows = np.asarray(sorted(sh[1:], key=lambda r: -r[0]), dtype=float)
                    es = rows[:, 0]
                    cs = rows[:, 1:]
                    nprim, nctr = cs.shape
                    cs = cs * _basis.gto_norm(l, es)[:, None]
                    cs = _basis.normalize_contracted(l, es, cs)
                    env.append(es)
                    env.append(cs.T.ravel())
                    tmpl.append([0, l, nprim, nctr, 0, ptr, ptr + nprim, 0])
                    ptr += nprim + nprim * nctr
                templates[key] = tmpl
            for b in templates[key]:
                b = list(b)
                b[ATOM_OF] = ia
                bas.append(b)
        self._atm = np.asarray(atm, dtype=np.int32).reshape(-1, ATM_SLOTS)
        self._bas = np.asarray(bas, dtype=np.int32).reshape(-1, BAS_SLOTS)
        self._env = np.hstack(env)
        self._built = True
        return self

    def _shells_for(self, sym):
        b = self.basis
        key = _basis.pure_symbol(sym)
        if isinstance(b, dict):
            b = b.get(sym, b.get(key, b.get('default')))
            if b is None:
                raise KeyError(f'no basis for {sym}')
        if isinstance(b, str):
            return _basis.load(b, key)
        # inline list [[l, [e, c...], ...], ...] used as given
        return [list(s) for s in b]

    # ---- properties ---------------------------------------------------
    @property
    def natm(self):
        return len(self._atm)

    @property
    def nbas(self):
        return len(self._bas)

    @property
    def omega(self):
        return float(self._env[PTR_RANGE_OMEGA])

    @omega.setter
    def omega(self, x):
        self._env[PTR_RANGE_OMEGA] = x

    @contextlib.contextmanager
    def with_range_coulomb(self, omega):
        old = self._env[PTR_RANGE_OMEGA]
        self._env[PTR_RANGE_OMEGA] = 0. if omega is None else omega
        try:
            yield
        finally:
            self._env[PTR_RANGE_OMEGA] = old

    def atom_coords(self):
        p = self._atm[:, PTR_COORD]
        return np.stack([self._env[q:q + 3] for q in p])

    def atom_charges(self):
        return self._atm[:, CHARGE_OF].copy()

    def atom_symbol(self, i):
        return self._atom[i][0]

    @property
    def nelectron(self):
        return int(self.atom_charges().sum() - self.charge)

    @property
    def nelec(self):
        ne = self.nelectron
        na = (ne + self.spin) // 2
        return na, ne - na

    def energy_nuc(self):
        z = self.atom_charges().astype(float)
        r = self.atom_coords()
        d = np.linalg.norm(r[:, None] - r[None], axis=-1)
        iu = np.triu_indices(len(z), 1)
        return float((z[:, None] * z[None])[iu] @ (1. / d[iu])) if len(z) > 1 else 0.

    def ao_loc_nr(self, cart=None):
        if cart is None:
            cart = self.cart
        l = self._bas[:, ANG_OF]
        n = ((l + 1) * (l + 2) // 2 if cart else 2 * l + 1) * self._bas[:, NCTR_OF]
        return np.append(0, np.cumsum(n)).astype(np.int32)

    @property
    def ao_loc(self):
        return self.ao_loc_nr()

    @property
    def nao(self):
        return int(self.ao_loc_nr()[-1])

    def nao_nr(self, cart=None):
        return int(self.ao_loc_nr(cart)[-1])

    def aoslice_by_atom(self):
        ao_loc = self.ao_loc_nr()
        out = np.zeros((self.natm, 4), dtype=int)
        at = self._bas[:, ATOM_OF]
        for ia in range(self.natm):
            idx = np.where(at == ia)[0]
            if len(idx) == 0:
                continue
            out[ia] = idx[0], idx[-1] + 1, ao_loc[idx[0]], ao_loc[idx[-1] + 1]
        return out

    def copy(self):
        import copy
        m = copy.copy(self)
        m._atm, m._bas, m._env = self._atm.copy(), self._bas.copy(), self._env.copy()
        return m

    def bas_exps(self, ib):
        p, n = self._bas[ib, PTR_EXP], self._bas[ib, NPRIM_OF]
        return self._env[p:p + n]

    def bas_ctr_coeff(self, ib):
        '''(nprim, nctr) normalised contraction coefficients'''
        p, n, c = self._bas[ib, PTR_COEFF], self._bas[ib, NPRIM_OF], self._bas[ib, NCTR_OF]
        return self._env[p:p + n * c].reshape(c, n).T

    def RHF(self):
        from ..scf import hf
        return hf.RHF(self)

    def UHF(self):
        from ..scf import hf
        return hf.UHF(self)


def M(**kw):
    """Equivalent of ``pyscf.M(...)``: build a molecule."""
    return Mole(**kw).build()


# ======================================================================
# ctypes struct handed to the C-ABI by reference
# ======================================================================
class RysIntEnvVars(ctypes.Structure):
    """Host struct of DEVICE pointers; C mirror in include/gvhf_rys_b200.h.
    Same field order as the reference (gpu4pyscf/gto/mole.py:1054-1082,
    lib/gvhf-rys/vhf.cuh:47-54) so existing ctypes callers bind unchanged."""
    _fields_ = [
        ('natm', ctypes.c_int),
        ('nbas', ctypes.c_int),
        ('atm', ctypes.c_void_p),
        ('bas', ctypes.c_void_p),
        ('env', ctypes.c_void_p),
        ('ao_loc', ctypes.c_void_p),
    ]

    @classmethod
    def new(cls, natm, nbas, atm, bas, env, ao_loc):
        '''atm, bas, env, ao_loc: device arrays (anything with data_ptr())'''
        obj = cls(natm, nbas, _ptr(atm), _ptr(bas), _ptr(env), _ptr(ao_loc))
        obj._keep = (atm, bas, env, ao_loc)
        return obj

    @property
    def device(self):
        return self._keep[0].device


def _ptr(a):
    if hasattr(a, 'data_ptr'):
        return a.data_ptr()
    if hasattr(a, 'data') and hasattr(a.data, 'ptr'):   # CuPy
        return a.data.ptr
    raise TypeError('device array expected')


def _scale_sp_ctr_coeff(mol):
    """Fold libcint's s/p angular factor sqrt((2l+1)/4pi) into the
    coefficients (reference: gto/mole.py:1120-1133)."""
    env = mol._env.copy()
    done = set()
    for b in mol._bas:
        l, p, n = b[ANG_OF], b[PTR_COEFF], b[NPRIM_OF] * b[NCTR_OF]
        if l < 2 and p not in done:
            env[p:p + n] *= ((2 * l + 1) / (4 * np.pi)) ** .5
            done.add(p)
    return env


MAX_NPRIM = 6        # primitives per sorted shell (csrc/jk_launch.cuh: JK_MAXPP = 36 primitive pairs)


class SortedGTO:
    """Decontracted, (l, nprim)-sorted Cartesian view of a ``Mole``.

    Public attributes follow the reference (gto/mole.py:585-655):
    ``mol``, ``_atm/_bas/_env``, ``uniq_l_ctr``, ``l_ctr_counts``,
    ``sorted_idx``, ``ao_loc``, ``p_ao_loc``, ``ctr_coeff``.  The recontraction
    map is kept as one dense matrix ``ctr_coeff`` (nao_sorted_cart, nao_orig):
    D_sorted = C D C^T and V_orig = C^T V_sorted C are plain FP64 GEMMs
    (cuBLAS via torch) -- < 1 % of a J/K build.
    """
    cart = True

    @classmethod
    def from_mol(cls, mol, *, group_size=None, decontract=True, diffuse_cutoff=0.3):
        if isinstance(mol, SortedGTO):
            return mol
        self = cls()
        self.mol = mol
        cutoff = 0. if diffuse_cutoff is None else diffuse_cutoff
        env = [mol._env.copy()]
        ptr = len(mol._env)
        pbas = []        # primitive/segment shells
        pmap = []        # for each new shell: (orig shell id, per-ctr coefficient vector)
        for ib, b in enumerate(mol._bas):
            l, nprim, nctr = int(b[ANG_OF]), int(b[NPRIM_OF]), int(b[NCTR_OF])
            es = mol._env[b[PTR_EXP]:b[PTR_EXP] + nprim]
            cs = mol._env[b[PTR_COEFF]:b[PTR_COEFF] + nprim * nctr].reshape(nctr, nprim)
            if nctr == 1:
                diffuse = (es <= cutoff) if decontract else np.zeros(nprim, bool)
                if nprim == 1 or not diffuse.any():
                    groups = [np.arange(nprim)]
                else:
                    groups = ([np.where(~diffuse)[0]] if (~diffuse).any() else []) + \
                             [np.array([i]) for i in np.where(diffuse)[0]]
                # the kernels hold at most 36 primitive pairs of a shell pair (JK_MAXPP): longer
                # segmented contractions are cut into pieces of <= 6 primitives whose functions add up
                groups = [g[i:i + MAX_NPRIM] for g in groups for i in range(0, len(g), MAX_NPRIM)]
                for g in groups:
                    env.append(es[g])
                    env.append(cs[0, g])
                    pbas.append([b[ATOM_OF], l, len(g), 1, 0, ptr, ptr + len(g), 0])
                    ptr += 2 * len(g)
                    pmap.append((ib, np.ones(1)))
            else:
                # general contraction: every primitive becomes its own
                # normalised shell, original functions are rebuilt by ctr_coeff
                norm = _basis.gto_norm(l, es)
                for ip in range(nprim):
                    env.append(es[ip:ip + 1])
                    env.append(norm[ip:ip + 1])
                    pbas.append([b[ATOM_OF], l, 1, 1, 0, ptr, ptr + 1, 0])
                    ptr += 2
                    pmap.append((ib, cs[:, ip] / norm[ip]))
        pbas = np.asarray(pbas, dtype=np.int32).reshape(-1, BAS_SLOTS)
        self._atm = mol._atm.copy()
        self._env = np.hstack(env)

        # sort by (l ascending, nprim descending), stable
        l_ctrs = pbas[:, [ANG_OF, NPRIM_OF]].copy()
        l_ctrs[:, 1] = -l_ctrs[:, 1]
        uniq, inv, counts = np.unique(l_ctrs, return_inverse=True, return_counts=True, axis=0)
        uniq[:, 1] = -uniq[:, 1]
        if group_size is not None:
            uniq, counts = _split_groups(uniq, counts, group_size)
        sorted_idx = np.argsort(inv.ravel(), kind='stable')
        self._bas = pbas[sorted_idx].copy()
        self._bas[:, PTR_BAS_COORD] = self._atm[self._bas[:, ATOM_OF], PTR_COORD]
        self.uniq_l_ctr = uniq
        self.l_ctr_counts = counts
        self.sorted_idx = sorted_idx
        self.p_ao_loc = self.ao_loc_nr(cart=True)

        # dense recontraction x cart2sph map
        o_ao_loc = mol.ao_loc_nr()
        nao_o = int(o_ao_loc[-1])
        nao_p = int(self.p_ao_loc[-1])
        C = np.zeros((nao_p, nao_o))
        for new_id, old_id in enumerate(sorted_idx):
            ib, coef = pmap[old_id]
            l = int(mol._bas[ib, ANG_OF])
            nctr = int(mol._bas[ib, NCTR_OF])
            c2s = np.eye(ncart(l)) if mol.cart else cart2sph(l)
            nf_o = c2s.shape[1]
            p0 = self.p_ao_loc[new_id]
            for ic in range(nctr):
                q0 = o_ao_loc[ib] + ic * nf_o
                C[p0:p0 + ncart(l), q0:q0 + nf_o] = c2s * coef[ic]
        self.ctr_coeff = C
        self._dev_cache = {}
        return self

    # -- Mole-like accessors
    @property
    def natm(self):
        return len(self._atm)

    @property
    def nbas(self):
        return len(self._bas)

    def ao_loc_nr(self, cart=True):
        l = self._bas[:, ANG_OF]
        n = (l + 1) * (l + 2) // 2
        return np.append(0, np.cumsum(n)).astype(np.int32)

    @property
    def ao_loc(self):
        return self.p_ao_loc

    @property
    def nao(self):
        return int(self.p_ao_loc[-1])

    @property
    def omega(self):
        return self.mol.omega

    @property
    def verbose(self):
        return self.mol.verbose

    def atom_coords(self):
        return self.mol.atom_coords()

    # -- AO-basis transforms (reference: gto/mole.py:667-939)
    def _coeff_on(self, device):
        import torch
        key = str(device)
        if key not in self._dev_cache:
            self._dev_cache[key] = torch.as_tensor(self.ctr_coeff, device=device)
        return self._dev_cache[key]

    def _csr_on(self, device, transpose):
        '''CSR form (row_ptr, col, val) of C (transpose=False) or C^T on `device`'''
        import torch
        key = (str(device), 'csrT' if transpose else 'csr')
        if key not in self._dev_cache:
            A = self.ctr_coeff.T if transpose else self.ctr_coeff
            rows, cols = np.nonzero(A)
            row_ptr = np.zeros(A.shape[0] + 1, dtype=np.int32)
            np.cumsum(np.bincount(rows, minlength=A.shape[0]), out=row_ptr[1:])
            self._dev_cache[key] = (torch.as_tensor(row_ptr, device=device),
                                    torch.as_tensor(cols.astype(np.int32), device=device),
                                    torch.as_tensor(np.ascontiguousarray(A[rows, cols]), device=device))
        return self._dev_cache[key]

    def _sparse_A_mat_AT(self, mat, transpose):
        import ctypes
        import torch
        from ..lib import loader
        lib = loader.load_library()
        shape = mat.shape
        ncol = shape[-1]
        mats = mat.reshape(-1, ncol, ncol).contiguous()
        nrow = self.ctr_coeff.shape[1] if transpose else self.ctr_coeff.shape[0]
        row_ptr, col, val = self._csr_on(mat.device, transpose)
        n = mats.shape[0]
        out = torch.empty((n, nrow, nrow), dtype=torch.float64, device=mat.device)
        tmp = torch.empty((n, nrow, ncol), dtype=torch.float64, device=mat.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        if lib.B200_sparse_A_mat_AT(p(out), p(tmp), p(mats), n, nrow, ncol, p(row_ptr), p(col), p(val),
                                    loader.current_stream_ptr()):
            raise RuntimeError('B200_sparse_A_mat_AT failed')
        return out.reshape(shape[:-2] + (nrow, nrow))

    def apply_C_mat_CT(self, mat):
        '''C . mat . C^T : original (sph/contracted) -> sorted Cartesian, through the sparse transform
        kernels (reference: SortedGTO.apply_C_mat_CT, gto/mole.py:906-921, bra_/ket_ kernels of
        mole_helper.cu:1735-1908)'''
        assert mat.shape[-1] == self.ctr_coeff.shape[1] and mat.dtype.is_floating_point
        return self._sparse_A_mat_AT(mat, False)

    def apply_CT_mat_C(self, mat):
        '''C^T . mat . C : sorted Cartesian -> original (reference: gto/mole.py:923-938)'''
        assert mat.shape[-1] == self.ctr_coeff.shape[0]
        return self._sparse_A_mat_AT(mat, True)

    def apply_C_dot(self, mat):
        C = self._coeff_on(mat.device)
        return C @ mat

    def apply_CT_dot(self, mat):
        C = self._coeff_on(mat.device)
        return C.T @ mat


def _split_groups(uniq, counts, group_size):
    u, c = [], []
    for row, n in zip(uniq, counts):
        l = row[0]
        nf = (l + 1) * (l + 2) // 2
        per = max(1, group_size // nf)
        while n > 0:
            m = min(per, n)
            u.append(row)
            c.append(m)
            n -= m
    return np.asarray(u), np.asarray(c)
