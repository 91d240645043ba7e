"""NWChem-format basis parser and PySCF-compatible shell normalisation.

Behaviour follows what the reference's tests rely on from PySCF's ``gto.basis``
(third-party, absent here): shells grouped by angular momentum in file order,
``'unc'`` prefix = every primitive its own shell with exponents descending and
duplicates dropped (used by gpu4pyscf/scf/tests/test_rhf.py:41-74).
"""
import math
import numpy as np
from . import basis_data

ANGULAR = 'spdfghik'
_ELEMENTS = ['X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne',
             'Na', 'Mg', 'Al', 'Si', 'P', 'S', 'Cl', 'Ar', 'K', 'Ca']


def charge_of(symbol):
    s = ''.join(c for c in symbol if c.isalpha())
    s = s[0].upper() + s[1:].lower()
    return _ELEMENTS.index(s)


def pure_symbol(symbol):
    s = ''.join(c for c in symbol if c.isalpha())
    return s[0].upper() + s[1:].lower()


def parse(text, symb=None):
    """Parse NWChem-format text -> list of shells ``[l, [exp, c1, c2..], ...]``.

    If ``symb`` is given only that element's blocks are kept.  Shells are
    returned grouped by l (stable), as PySCF's parse_nwchem does.
    """
    shells = []
    cur = None          # list of shells fed by the current block (1, or 2 for SP)
    keep = False
    for raw in text.splitlines():
        line = raw.split('#')[0].split('!')[0].strip()
        if not line:
            continue
        tok = line.split()
        up = tok[0].upper()
        if up in ('BASIS', 'END', 'SPHERICAL', 'CARTESIAN') or up.startswith('BASIS'):
            continue
        if len(tok) == 2 and tok[0].isalpha() and tok[1].isalpha():
            sym, typ = pure_symbol(tok[0]), tok[1].upper()
            keep = symb is None or sym == pure_symbol(symb)
            if not keep:
                cur = None
                continue
            if typ == 'SP':
                cur = [[0], [1]]
            else:
                cur = [[ANGULAR.index(typ.lower())]]
            shells.extend(cur)
            continue
        if cur is None:
            continue
        vals = [float(x.replace('D', 'E').replace('d', 'e')) for x in tok]
        if len(cur) == 2:   # SP shell: exp, c_s, c_p
            cur[0].append([vals[0], vals[1]])
            cur[1].append([vals[0], vals[2]])
        else:
            cur[0].append(vals)
    shells = [s for s in shells if len(s) > 1]
    out = []
    for l in range(len(ANGULAR)):
        out.extend(s for s in shells if s[0] == l)
    return out


def uncontract(shells):
    """Every primitive becomes its own shell; per l exponents descending,
    duplicates (to 9 decimals) dropped."""
    out = []
    for l in range(len(ANGULAR)):
        es = [p[0] for s in shells if s[0] == l for p in s[1:]]
        if not es:
            continue
        es = np.unique(np.asarray(es).round(9))
        for e in es[::-1]:
            out.append([l, [float(e), 1.0]])
    return out


def load(name_or_text, symb):
    """Resolve a basis specification for one element.

    ``name_or_text``: a library name ('def2-tzvp'), NWChem text, either
    optionally prefixed by 'unc'.
    """
    spec = name_or_text
    unc = False
    if spec[:3].lower() == 'unc':
        unc = True
        spec = spec[3:].lstrip('-')
    key = basis_data.normalize_name(spec) if '\n' not in spec else None
    if key is not None:
        if key not in basis_data.LIBRARY:
            raise KeyError(f'basis set {spec!r} is not embedded in gpu4pyscf_b200')
        text = basis_data.LIBRARY[key]
    else:
        text = spec
    shells = parse(text, symb)
    if not shells:
        raise KeyError(f'no basis functions for element {symb} in basis spec')
    if unc:
        shells = uncontract(shells)
    return shells


def gaussian_int(n, alpha):
    """int_0^inf x^n exp(-alpha x^2) dx"""
    n1 = (n + 1) * .5
    return math.gamma(n1) / (2. * np.asarray(alpha, dtype=float) ** n1)


def gto_norm(l, expnt):
    """Radial normalisation: 1/sqrt(int r^{2l+2} exp(-2 a r^2) dr)."""
    return 1. / np.sqrt(gaussian_int(l * 2 + 2, 2 * np.asarray(expnt, dtype=float)))


def normalize_contracted(l, es, cs):
    """cs: (nprim, nctr) already multiplied by gto_norm; rescale each
    contracted function to unit radial norm."""
    ee = es[:, None] + es[None, :]
    ee = gaussian_int(l * 2 + 2, ee)
    s1 = 1. / np.sqrt(np.einsum('pi,pq,qi->i', cs, ee, cs))
    return cs * s1
