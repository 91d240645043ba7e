"""Molecules used by the reference's hot-path tests (inputs restated, see citations)."""
from gpu4pyscf_b200.gto.mole import M

# gpu4pyscf/scf/tests/test_rhf.py:41-74  (inline uncontracted basis, l up to 5)
MOL1_ATOM = '''
C  -1.20806619, -0.34108413, -0.00755148
C   1.28636081, -0.34128013, -0.00668648
H   2.53407081,  1.81906387, -0.00736748
H   1.28693681,  3.97963587, -0.00925948
'''
MOL1_BASIS = '''unc
#BASIS SET:
H    S
      1.815041   1
      0.591063   1
H    P
      2.305000   1
#BASIS SET:
C    S
      8.383976   1
      3.577015   1
      1.547118   1
C    P
      2.305000   1
      1.098827   1
      0.806750   1
      0.282362   1
C    D
      1.81900    1
      0.72760    1
      0.29104    1
C    F
      0.970109   1
C    G
      0.625000   1
C    H
      0.4        1
'''


def mol1():
    return M(atom=MOL1_ATOM, basis=MOL1_BASIS)


def mol1_lmax(lmax):
    """mol1 with shells above lmax dropped (the CUDA path supports l <= 4)."""
    keep = []
    skip = False
    for line in MOL1_BASIS.splitlines():
        t = line.split()
        if len(t) == 2 and t[0] in ('H', 'C') and t[1].isalpha():
            skip = 'SPDFGHI'.index(t[1].upper()) > lmax
        if not skip:
            keep.append(line)
    return M(atom=MOL1_ATOM, basis='\n'.join(keep))


# gpu4pyscf/scf/tests/test_scf_jk.py:24-36  (unit='B')
JK_ATOM = '''
O   0.000   -0.    0.1174
H  -0.757    4.   -0.4696
H   0.757    4.   -0.4696
C   1.      1.    0.
H   4.      0.    3.
H   0.      1.    .6
'''


def jkmol(basis='def2-tzvp', **kw):
    return M(atom=JK_ATOM, basis=basis, unit='B', **kw)


WATER = '''
O       0.0000000000    -0.0000000000     0.1174000000
H      -0.7570000000    -0.0000000000    -0.4696000000
H       0.7570000000     0.0000000000    -0.4696000000
'''


def water(basis='def2-tzvpp', **kw):
    return M(atom=WATER, basis=basis, **kw)
