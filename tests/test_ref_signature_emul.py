"""Reference-signature entry points RYS_build_jk / RYS_build_k (gpu4pyscf/scf/jk.py:564-583,
845-877) driven exactly like the reference's per-group launch loop, through the SIMT
emulator build of the kernels (CPU only)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, ctypes, math
os.environ['GPU4PYSCF_B200_EMUL'] = '1'
sys.path.insert(0, %r)
import numpy as np
import torch
from gpu4pyscf_b200.csrc import build
build.build(emul=True)
from gpu4pyscf_b200.scf import jk
from gpu4pyscf_b200.lib import loader
from tests import molecules
lib = loader.load_library()
mol = molecules.water('def2-svp')
np.random.seed(7)
nao = mol.nao
dm = np.random.rand(nao, nao)
dm = dm + dm.T
opt = jk._VHFOpt(mol).build()
smol = opt.sorted_mol
dms = opt.apply_coeff_C_mat_CT(torch.as_tensor(dm)[None]).contiguous()
nbas, naos = smol.nbas, smol.nao
dm_cond = torch.empty((nbas, nbas), dtype=torch.float32)
lib.B200_condense(jk._ptr(dm_cond), jk._ptr(dms), 1, naos, jk._ptr(opt.rys_envs._keep[3]), nbas, None)
offs = np.append(0, np.cumsum(smol.l_ctr_counts))
ng = len(smol.l_ctr_counts)
bas_host = np.ascontiguousarray(smol._bas, dtype=np.int32)
cutoff = math.log(1e-13)
pool = torch.zeros(16, dtype=torch.int32)
omega, lr, sr = 0.4, 0.65, 0.2

def loop(fn):
    for i in range(ng):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    pij, qij = opt.bas_pair_cache[i, j][:2]
                    pkl, qkl = opt.bas_pair_cache[k, l][:2]
                    shls = (ctypes.c_int * 8)(offs[i], offs[i + 1], offs[j], offs[j + 1],
                                              offs[k], offs[k + 1], offs[l], offs[l + 1])
                    assert fn(shls, pij, qij, pkl, qkl) == 0

vk = torch.zeros_like(dms)
loop(lambda shls, pij, qij, pkl, qkl: lib.RYS_build_k(
    jk._ptr(vk), jk._ptr(dms), 1, naos, omega, lr, sr, ctypes.byref(opt.rys_envs), shls, 0,
    len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl), jk._ptr(qij), jk._ptr(qkl),
    None, jk._ptr(dm_cond), cutoff, float(dm_cond.max()), jk._ptr(pool),
    bas_host.ctypes.data_as(ctypes.c_void_p)))
vk = vk + vk.transpose(1, 2)
ref = opt.get_k(dms, 1, omega=omega, lr_factor=lr, sr_factor=sr)
assert abs(vk - ref).max() < 1e-11, abs(vk - ref).max()

# omega == 0: the weights are scaled by lr_factor (reference: rys_roots_for_k.cu:25-32), so
# (0, 1, 1) is plain K, (0, .25, .25) a quarter of it and (0, 0, 0) nothing
def k_at_omega0(fac):
    out = torch.zeros_like(dms)
    loop(lambda shls, pij, qij, pkl, qkl: lib.RYS_build_k(
        jk._ptr(out), jk._ptr(dms), 1, naos, 0., fac, fac, ctypes.byref(opt.rys_envs), shls, 0,
        len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl), jk._ptr(qij), jk._ptr(qkl),
        None, jk._ptr(dm_cond), cutoff, float(dm_cond.max()), jk._ptr(pool),
        bas_host.ctypes.data_as(ctypes.c_void_p)))
    return out + out.transpose(1, 2)
ref0 = opt.get_k(dms, 1)
assert abs(k_at_omega0(1.) - ref0).max() < 1e-11
assert abs(k_at_omega0(.25) - .25 * ref0).max() < 1e-11
assert abs(k_at_omega0(0.)).max() == 0.
assert abs(opt.get_k(dms, 1, omega=0., lr_factor=.25, sr_factor=.25) - .25 * ref0).max() < 1e-11

# RYS_build_jk with env[8] < 0: short-range J and K (libcint convention), one call per group
env_host = np.ascontiguousarray(smol._env, dtype=np.float64).copy()
env_host[8] = -0.35
vj = torch.zeros_like(dms)
vk = torch.zeros_like(dms)
loop(lambda shls, pij, qij, pkl, qkl: lib.RYS_build_jk(
    jk._ptr(vj), jk._ptr(vk), jk._ptr(dms), 1, naos, ctypes.byref(opt.rys_envs), shls, 0,
    len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl), jk._ptr(qij), jk._ptr(qkl),
    None, jk._ptr(dm_cond), cutoff, float(dm_cond.max()), jk._ptr(pool),
    bas_host.ctypes.data_as(ctypes.c_void_p), env_host.ctypes.data_as(ctypes.c_void_p)))
vk = vk + vk.transpose(1, 2)
vj = 2. * vj
vj = vj + vj.transpose(1, 2)
full_j, full_k = opt._jk_pass(dms, 1, True, True, 0.)
lr_j, lr_k = opt._jk_pass(dms, 1, True, True, 0.35)
assert abs(vk - (full_k - lr_k)).max() < 1e-11
assert abs(vj - (full_j - lr_j)).max() < 1e-11

# RYS_build_j (reference name; plain sorted-Cartesian matrices, see include/gvhf_rys_b200.h)
assert lib.RYS_init_rysj_constant() == 0
env_host[8] = 0.
vj = torch.zeros_like(dms)
loop(lambda shls, pij, qij, pkl, qkl: lib.RYS_build_j(
    jk._ptr(vj), jk._ptr(dms), 1, naos, ctypes.byref(opt.rys_envs), None, shls,
    len(pij), len(pkl), jk._ptr(pij), jk._ptr(pkl), jk._ptr(qij), jk._ptr(qkl), jk._ptr(qij), jk._ptr(qkl),
    None, jk._ptr(dm_cond), cutoff, float(dm_cond.max()), jk._ptr(pool),
    bas_host.ctypes.data_as(ctypes.c_void_p), env_host.ctypes.data_as(ctypes.c_void_p)))
vj = 2. * vj
vj = vj + vj.transpose(1, 2)
assert abs(vj - opt.get_j(dms)).max() < 1e-11

# int2e_qcond_estimator (reference signature) for the three operators against the oracle's
# sqrt(max|(ij|ij)|): full range, long range (omega > 0, sr = 0), short range (omega < 0, lr = 0)
from oracle import oracle
ids = torch.cat([v[0] for v in opt.bas_pair_cache.values()]).contiguous()
cnts = [len(v[0]) for v in opt.bas_pair_cache.values()]
o = np.append(0, np.cumsum(cnts))
spo = torch.as_tensor(np.hstack([np.arange(o[g], o[g + 1], 7) for g in range(len(cnts))] + [[o[-1]]]).astype(np.int32))
idn = ids.numpy().astype(np.int64)
for om, lrf, srf in ((0., 1., 1.), (0.4, 1., 0.), (-0.4, 0., 1.)):
    q = torch.empty(len(ids), dtype=torch.float32)
    assert lib.int2e_qcond_estimator(jk._ptr(q), ctypes.byref(opt.rys_envs), 0, len(spo) - 1, jk._ptr(ids),
                                     jk._ptr(spo), None, om, lrf, srf) == 0
    ref = np.log(oracle.q_cond(smol, omega=om)[idn // nbas, idn %% nbas])
    assert abs(q.numpy() - ref).max() < 2e-5, (om, abs(q.numpy() - ref).max())

# fill_s_estimator (reference signature, nr_sr_estimator.cu:400-414) against a numpy restatement of the formula
from gpu4pyscf_b200.scf.jk import _diffuse_pgto_params
ex, cs = _diffuse_pgto_params(smol)
s_out = torch.empty(len(ids), dtype=torch.float32)
assert lib.fill_s_estimator(jk._ptr(s_out), ctypes.byref(opt.rys_envs), jk._ptr(ids), jk._ptr(torch.as_tensor(ex)),
                            jk._ptr(torch.as_tensor(cs)), len(ids), -0.4) == 0
ish, jsh = idn // nbas, idn %% nbas
R = np.array([smol._env[b[7]:b[7] + 3] for b in smol._bas])
ll = smol._bas[:, 1].astype(float)
ai, aj = ex[ish].astype(float), ex[jsh].astype(float)
aij = ai + aj
rr = ((R[ish] - R[jsh]) ** 2).sum(axis=1)
rg = 6. / np.sqrt(aij * 0.16 / (0.16 + aij))
norm = np.where(ll[ish] >= 2, (2 * ll[ish] + 1) / (4 * np.pi), 1.) * np.where(ll[jsh] >= 2, (2 * ll[jsh] + 1) / (4 * np.pi), 1.)
ref_s = (.5 * ll[ish] * np.log(.5 * ll[ish] / aij + (aj / aij * rg) ** 2 + 1e-9)
         + .5 * ll[jsh] * np.log(.5 * ll[jsh] / aij + (ai / aij * rg) ** 2 + 1e-9)
         - ai * aj / aij * rr + np.log(cs[ish] * cs[jsh] * np.sqrt(norm)) + 1.7171 - 1.5 * np.log(aij) + .5 - np.log(0.16) / 4)
assert abs(s_out.numpy() - ref_s).max() < 2e-4, abs(s_out.numpy() - ref_s).max()

# pair prescreen: the group-pair lists hold exactly the ish >= jsh pairs of their two groups, once each
seen = set()
grp = np.repeat(np.arange(ng), smol.l_ctr_counts)
for (gi, gj), v in opt.bas_pair_cache.items():
    a = v[0].numpy().astype(np.int64)
    assert (grp[a // nbas] == gi).all() and (grp[a %% nbas] == gj).all() and (a // nbas >= a %% nbas).all()
    assert len(set(a.tolist())) == len(a) and not (seen & set(a.tolist()))
    seen |= set(a.tolist())
assert len(seen) == nbas * (nbas + 1) // 2          # water: nothing is far enough to be dropped
print('OK')
''' % ROOT


def test_rys_build_k_reference_signature():
    r = subprocess.run([sys.executable, '-c', SCRIPT], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
