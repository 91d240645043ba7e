"""ctypes loader for libgvhf_rys_b200.so (reference: gpu4pyscf/lib/utils.py:31-36
``load_library`` + gpu4pyscf/scf/jk.py:42-47).

There is NO CPU fallback: if the CUDA library is missing the import of the J/K
plugin fails.  The only exception is the SIMT *emulator* build of the very same
kernel sources (tests/emul), which the CPU-only test-suite opts into with
GPU4PYSCF_B200_EMUL=1 to debug the kernels without a GPU.
"""
import os
import ctypes

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None


def emulated():
    return os.environ.get('GPU4PYSCF_B200_EMUL', '0') == '1'


def device():
    import torch
    if emulated():
        return torch.device('cpu')
    if not torch.cuda.is_available():
        raise RuntimeError('gpu4pyscf_b200 needs a CUDA device (sm_100a); none is visible')
    return torch.device('cuda', torch.cuda.current_device())


def load_library(name='libgvhf_rys_b200'):
    global _LIB
    if _LIB is not None:
        return _LIB
    if emulated():
        path = os.path.join(_ROOT, 'tests', 'emul', name + '_emul.so')
    else:
        # GPU4PYSCF_B200_VARIANT selects an experimental build of the same CUDA sources
        variant = os.environ.get('GPU4PYSCF_B200_VARIANT')
        path = os.path.join(_HERE, name + ('_' + variant if variant else '') + '.so')
    if not os.path.exists(path):
        raise ImportError(
            f'{path} not found: build it with `python -m gpu4pyscf_b200.csrc.build` '
            '(the J/K path has no CPU fallback)')
    lib = ctypes.CDLL(path)
    c_int, c_float, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    lib.RYS_b200_init.restype = c_int
    lib.RYS_b200_lmax.restype = c_int
    lib.cuda_version.restype = c_int
    lib.B200_build_jk_class.restype = c_int
    lib.B200_build_jk_class.argtypes = [
        vp, vp, vp, c_int, c_int, vp, vp,
        vp, vp, c_int, vp, vp, c_int,
        vp, c_float, c_float, c_float, c_int, c_int, c_int, vp, vp, vp,
        ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, ctypes.c_double]
    lib.B200_build_jk_rows.restype = c_int
    lib.B200_build_jk_rows.argtypes = [
        vp, vp, vp, vp, vp, c_int, c_int, vp, vp,
        vp, vp, c_int,
        vp, vp, vp, vp, vp, vp, c_int,
        vp, c_float, c_float, c_int, c_int, c_int, vp, vp,
        ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, ctypes.c_double]
    lib.B200_sr_pair_data.restype = c_int
    lib.B200_sr_pair_data.argtypes = [vp, vp, vp, vp, vp, vp, c_int, ctypes.c_double, vp]
    lib.fill_s_estimator.restype = c_int
    lib.fill_s_estimator.argtypes = [vp, vp, vp, vp, vp, c_int, ctypes.c_double]
    lib.B200_class_two_terms.restype = c_int
    lib.B200_class_two_terms.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.B200_class_kind.restype = c_int
    lib.B200_class_kind.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.B200_has_row_class.restype = c_int
    lib.B200_has_row_class.argtypes = [c_int, c_int, c_int, c_int]
    lib.B200_pair_rec_stride.restype = c_int
    lib.B200_pair_rec_stride.argtypes = [c_int, c_int]
    lib.B200_pair_records.restype = c_int
    lib.B200_pair_records.argtypes = [vp, vp, vp, c_int, c_int, c_int, vp]
    lib.RYS_build_jk.restype = c_int
    lib.RYS_build_jk.argtypes = [
        vp, vp, vp, c_int, c_int, vp, vp, c_int, c_int, c_int,
        vp, vp, vp, vp, vp, vp, vp, vp, c_float, c_float, vp, vp, vp]
    lib.B200_has_class.restype = c_int
    lib.B200_has_class.argtypes = [c_int, c_int, c_int, c_int]
    lib.RYS_build_k.restype = c_int
    lib.RYS_build_k.argtypes = [
        vp, vp, c_int, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp, c_int, c_int, c_int,
        vp, vp, vp, vp, vp, vp, vp, vp, c_float, c_float, vp, vp]
    lib.B200_qcond.restype = c_int
    lib.B200_qcond.argtypes = [vp, vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.int2e_qcond_estimator.restype = c_int
    lib.int2e_qcond_estimator.argtypes = [vp, vp, c_int, c_int, vp, vp, vp,
                                          ctypes.c_double, ctypes.c_double, ctypes.c_double]
    lib.B200_pair_prescreen.restype = c_int
    lib.B200_pair_prescreen.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_float, vp]
    lib.RYS_build_j.restype = c_int
    lib.RYS_build_j.argtypes = [
        vp, vp, c_int, c_int, vp, vp, vp, c_int, c_int,
        vp, vp, vp, vp, vp, vp, vp, vp, c_float, c_float, vp, vp, vp]
    lib.RYS_init_rysj_constant.restype = c_int
    lib.B200_sparse_A_mat_AT.restype = c_int
    lib.B200_sparse_A_mat_AT.argtypes = [vp, vp, vp, c_int, c_int, c_int, vp, vp, vp, vp]
    lib.B200_set_counter.restype = c_int
    lib.B200_set_counter.argtypes = [vp]
    lib.B200_condense.restype = c_int
    lib.B200_condense.argtypes = [vp, vp, c_int, c_int, vp, c_int, vp]
    lib.B200_dfma_burn.restype = c_int
    lib.B200_dfma_burn.argtypes = [vp, c_int, c_int, c_int, vp]
    _LIB = lib
    return lib


def current_stream_ptr():
    import torch
    if emulated():
        return None
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
