#!/usr/bin/env python
"""Build libgvhf_rys_b200.so for sm_100a (nvcc), in-tree.

    python -m gpu4pyscf_b200.csrc.build            # CUDA library -> gpu4pyscf_b200/lib/
    python -m gpu4pyscf_b200.csrc.build --emul     # CPU SIMT emulation (tests only) -> tests/emul/

Only the standard toolchain is used (nvcc / g++); objects are cached under
gpu4pyscf_b200/csrc/build/ and rebuilt when a source or header is newer.
"""
import os
import sys
import glob
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def _generate(variant=None, genenv=None):
    gen = os.path.join(HERE, 'gen' + ('_' + variant if variant else ''))
    script = os.path.join(ROOT, 'tools', 'gen_instances.py')
    stamp = os.path.join(gen, 'jk_dispatch.inc')
    if not os.path.exists(stamp) or os.path.getmtime(stamp) < os.path.getmtime(script):
        env = dict(os.environ)
        env.update(genenv or {})
        subprocess.check_call([sys.executable, script, gen], stdout=subprocess.DEVNULL, env=env)
    return gen, sorted(glob.glob(os.path.join(gen, 'jk_inst_*.cu')))


# headers only capi.cu includes: a change there does not rebuild the class instantiation units
CAPI_ONLY = ('qcond.cuh', 'rys_tables.h', 'ao_transform.cuh')


def _newest_header(capi=True):
    hs = glob.glob(os.path.join(HERE, '*.cuh')) + glob.glob(os.path.join(HERE, '*.h'))
    if not capi:
        hs = [h for h in hs if os.path.basename(h) not in CAPI_ONLY]
    return max(os.path.getmtime(h) for h in hs)


def build(emul=False, verbose=False, jobs=None, variant=None, defs=(), genenv=None):
    gen, insts = _generate(variant, genenv)
    srcs = insts + [os.path.join(HERE, 'capi.cu')]
    tag = 'emul' if emul else 'sm100a'
    if variant:
        tag += '_' + variant
    objdir = os.path.join(HERE, 'build', tag)
    os.makedirs(objdir, exist_ok=True)
    if emul:
        shim = os.path.join(ROOT, 'tests', 'emul')
        srcs = srcs + [os.path.join(shim, 'cuda_shim.cpp')]
        out = os.path.join(shim, 'libgvhf_rys_b200_emul.so')
        shim_time = os.path.getmtime(os.path.join(shim, 'cuda_shim.h'))
        hdr_time = max(_newest_header(), shim_time)
        inst_time = max(_newest_header(False), shim_time)
    else:
        out = os.path.join(ROOT, 'gpu4pyscf_b200', 'lib',
                           'libgvhf_rys_b200%s.so' % ('_' + variant if variant else ''))
        hdr_time = _newest_header()
        inst_time = _newest_header(False)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src).rsplit('.', 1)[0] + '.o')
        dep_time = hdr_time if os.path.basename(src) == 'capi.cu' else inst_time
        if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), dep_time):
            return obj, False
        if emul:
            cmd = ['g++', '-O1', '-g', '-std=c++17', '-fPIC', '-DCPU_EMUL', '-x', 'c++',
                   '-I', HERE, '-I', gen, '-I', os.path.join(ROOT, 'tests', 'emul'),
                   '-Wno-unused-variable', '-c', src, '-o', obj]
        else:
            cmd = [NVCC] + ARCH + ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
                                  '-I', HERE, '-I', gen, '-c', src, '-o', obj] + ['-D' + d for d in defs]
            if verbose:
                cmd += ['-Xptxas', '-v']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('compile failed: %s\n%s' % (' '.join(cmd), r.stderr[-4000:]))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj, True

    jobs = jobs or os.cpu_count() or 4
    with ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(out):
        if emul:
            cmd = ['g++', '-shared', '-o', out] + objs
        else:
            cmd = [NVCC] + ARCH + ['-shared', '-o', out] + objs
        subprocess.check_call(cmd)
    return out


if __name__ == '__main__':
    variant = None
    defs = []
    genenv = {}
    for a in sys.argv[1:]:
        if a.startswith('--gen='):
            k, v = a[6:].split('=', 1)
            genenv[k] = v
        if a.startswith('--variant='):
            variant = a.split('=', 1)[1]
        elif a.startswith('-D'):
            defs.append(a[2:])
    print(build(emul='--emul' in sys.argv, verbose='-v' in sys.argv, variant=variant, defs=defs, genenv=genenv))
