#!/bin/bash
# compute-sanitizer runs of the J/K path (the kernels rely on warp-level queues,
# named barriers, shared-memory transposes and FP64 atomics).  usage: tools/sanitize.sh [outdir]
# memcheck + racecheck + synccheck on (a) water/def2-TZVP (thread + row kernels, contracted p shells)
# and (b) the inline l <= 4 molecule of the reference's test_rhf.py (warp-team and block kernels, g shells).
OUT=${1:-gpurun_out}
mkdir -p $OUT
cat > /tmp/sanitize_case.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from gpu4pyscf_b200.scf import jk
from tests import molecules
mol = molecules.water('def2-tzvp') if sys.argv[1] == 'water' else molecules.mol1_lmax(4)
np.random.seed(1)
dm = np.random.rand(2, mol.nao, mol.nao)
dm = dm + dm.transpose(0, 2, 1)
vj, vk = jk.get_jk(mol, dm, hermi=1)
vk2 = jk.get_k(mol, dm[0], hermi=1, omega=-0.3)
torch.cuda.synchronize()
print('ran', sys.argv[1], float(vj.abs().max()), float(vk.abs().max()), float(vk2.abs().max()))
PY
for tool in memcheck racecheck synccheck; do
  for case in water lmax4; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_case.py $case \
        > $OUT/sanitizer_${tool}_${case}.log 2>&1
    echo "== $tool $case: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_${case}.log | tail -1)"
  done
done
