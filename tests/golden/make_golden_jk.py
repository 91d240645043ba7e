#!/usr/bin/env python
"""Generate a golden J/K fixture with the CPU oracle (oracle/md_jk.c, itself pinned to the
reference's golden fingerprints by tests/test_oracle.py) for a system that is too slow for the
oracle inside the test-suite (minutes to an hour on 8 cores).

    python tests/golden/make_golden_jk.py 020_Vitamin_C def2-tzvpp [natoms]

writes tests/golden/jk_<mol>_<basis>[_<natoms>].npz with, for the bench density
(bench.decaying_density, seed 0, hermi = 1): PySCF-style fingerprints fp(J), fp(K), max|J|, max|K|,
20000 sampled matrix elements (fixed index set) and J @ V, K @ V for four fixed random vectors.
The GPU tests (tests/test_jk_golden_gpu.py) compare against all of them at 1e-10 absolute.
No screening is applied in the oracle run (tol = 0): it is the exact reference."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import oracle

name, basis = sys.argv[1], sys.argv[2]
natoms = int(sys.argv[3]) if len(sys.argv) > 3 else None
mol = bench.load_mol(name, basis, natoms)
dm = bench.decaying_density(mol)
t0 = time.time()
vj, vk = oracle.get_jk(mol, dm, hermi=1)
dt = time.time() - t0
nao = mol.nao
rng = np.random.RandomState(1234)
idx = rng.randint(0, nao * nao, 20000)
vec = rng.rand(nao, 4) - .5
tag = '%s_%s%s' % (name, basis.replace('-', '').lower(), '_%d' % natoms if natoms else '')
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'jk_%s.npz' % tag)
np.savez_compressed(out, nao=nao, fp_j=oracle.fp(vj), fp_k=oracle.fp(vk), max_j=abs(vj).max(), max_k=abs(vk).max(),
                    idx=idx, j_idx=vj.ravel()[idx], k_idx=vk.ravel()[idx], vec=vec, j_vec=vj @ vec, k_vec=vk @ vec,
                    oracle_seconds=dt)
print(out, 'nao', nao, 'oracle %.0f s' % dt)
