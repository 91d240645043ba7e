"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, and exports every symbol include/gvhf_rys_b200.h declares (no compute
calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'gvhf_rys_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\bint\s+([A-Za-z_][A-Za-z0-9_]*)\s*\(', src)))


def test_header_declares_the_reference_entry():
    syms = _declared_symbols()
    assert 'RYS_build_jk' in syms and 'B200_build_jk_class' in syms and 'B200_qcond' in syms


def test_library_exports_every_declared_symbol():
    from gpu4pyscf_b200.csrc import build
    so = build.build(emul=False)
    lib = ctypes.CDLL(so)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.RYS_b200_lmax.restype = ctypes.c_int
    assert lib.RYS_b200_lmax() == 4


def test_product_path_fails_loudly_without_cuda():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import torch\n"
            "from gpu4pyscf_b200.lib import loader\n"
            "if torch.cuda.is_available():\n"
            "    print('HAS_GPU'); sys.exit(0)\n"
            "try:\n"
            "    loader.device()\n"
            "except RuntimeError as e:\n"
            "    print('RAISED', e)\n" % ROOT)
    env = dict(os.environ)
    env.pop('GPU4PYSCF_B200_EMUL', None)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
    assert 'RAISED' in r.stdout or 'HAS_GPU' in r.stdout, r.stdout + r.stderr
