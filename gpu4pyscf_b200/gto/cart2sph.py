"""Cartesian -> real-spherical transformation matrices (libcint conventions).

Cartesian components are in lexical order (xx, xy, xz, yy, yz, zz; ...), the
order the reference lists in lib/gvhf-rys/rys_constant.cu:21-112.  Spherical
components are m = -l..l, except p which is (x, y, z).  For l >= 2 the matrix
carries the full angular normalisation (Cartesian d, f, ... carry none); for
s and p it is the identity because libcint folds sqrt((2l+1)/4pi) into the
radial coefficient (reference mirror: gto/mole.py:1120-1133).
The numbers are generated from the closed-form real solid harmonics
(Helgaker, Jorgensen, Olsen, eq. 6.4.47-50) and checked against the values
the reference hard-codes (lib/gint/cart2sph.cu:19-167) in tests/test_gto.py.
"""
import math
from functools import lru_cache
import numpy as np


def cart_powers(l):
    """Lexical Cartesian exponents [(lx, ly, lz)] for shell l."""
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1)
            for ly in range(l - lx, -1, -1)]


def ncart(l):
    return (l + 1) * (l + 2) // 2


def _binom(n, k):
    # generalised to half-integers is not needed; integer only
    if k < 0 or k > n:
        return 0
    return math.comb(n, k)


@lru_cache(maxsize=None)
def cart2sph(l):
    """(ncart, 2l+1) matrix c with |sph_m> = sum_c c[c, m] |cart_c>."""
    if l == 0:
        return np.ones((1, 1))
    if l == 1:
        return np.eye(3)
    pw = cart_powers(l)
    index = {p: i for i, p in enumerate(pw)}
    out = np.zeros((len(pw), 2 * l + 1))
    for m in range(-l, l + 1):
        am = abs(m)
        # N_lm of Helgaker (Racah-normalised solid harmonics)
        nlm = (1. / (2 ** am * math.factorial(l))) * math.sqrt(
            2. * math.factorial(l + am) * math.factorial(l - am) / (2. if m == 0 else 1.))
        # twice v_m to stay in integers
        vm2 = 0 if m >= 0 else 1
        for t in range((l - am) // 2 + 1):
            for u in range(t + 1):
                # v runs over v_m, v_m+1, ... <= floor(|m|/2 - v_m) + v_m
                nv = int(math.floor(am / 2. - vm2 / 2.))
                for iv in range(nv + 1):
                    v2 = 2 * iv + vm2            # 2*v
                    coef = ((-1) ** (t + iv) * 0.25 ** t
                            * _binom(l, t) * _binom(l - t, am + t)
                            * _binom(t, u) * _binom(am, v2))
                    px = 2 * t + am - 2 * u - v2
                    py = 2 * u + v2
                    pz = l - 2 * t - am
                    out[index[(px, py, pz)], m + l] += nlm * coef
    # Racah normalisation -> unit-sphere normalisation
    out *= math.sqrt((2 * l + 1) / (4 * math.pi))
    return out
