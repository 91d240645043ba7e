#!/usr/bin/env python
"""Headline benchmark: seconds per direct-SCF J/K build (BASELINE.json metric)
for Taxol / def2-TZVP (2228 spherical AOs, 2538 sorted Cartesian) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full J/K build (`_VHFOpt.get_jk`, hermi=1, n_dm=1): density
screening (dm_cond), every angular-class launch, the all-reduce of the [J|K]
partials when N > 1 and the symmetrisation.

  value   : s/build with the (sorted-basis) density resident in HBM, timed with
            CUDA events on the launching stream, max over ranks
  e2e     : s/build through the public plugin call `scf.jk.get_jk(mol, dm_host)`
            with the density in pinned host memory and J, K read back to the host
            inside the timed region
  roofline: FP64-FMA bound.  achieved = min(F_alg, F_hw) / t with F_alg = SURVEY.md 8d
            flop model x surviving quartets counted by the kernels in this run and
            F_hw = 2*DFMA + DMUL + DADD thread instructions of one build from the
            committed ncu capture (profiles/ncu_r2/build_counters_*.json, which also
            supplies the measured DRAM traffic); peak = DFMA burn measured in this
            very run (B200_dfma_burn); frac_of_nominal is against 37.2 TFLOP/s
  cpu_baseline: the CPU oracle (oracle/md_jk.c, OpenMP) timed on a fragment of
            the same molecule and scaled by algorithmic flops, plus the measured
            full-molecule oracle runs of the golden fixtures with the GPU time of
            the same build beside them

`--impl reference` times only the CPU oracle (PySCF/libcvhf cannot be installed
here: see DESIGN.md), on the same bounded sample.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'J/K build s/iter (direct SCF, Taxol def2-TZVP)'      # other --mol/--basis: named in config.workload
# F_alg of one full build (SURVEY.md 8d flop model x quartets counted by the GPU
# kernels for the bench density; profiles/bench_r1_baseline.json).  Only used to
# scale the CPU arm's bounded sample to the full workload.
F_ALG_FULL = {('113_Taxol', 'def2-tzvp'): 4.8760375e13}
MOLDIR = os.path.join(ROOT, 'gpu4pyscf_b200', 'data', 'molecules')


def decaying_density(mol, seed=0, beta=0.8):
    '''Synthetic symmetric density with the exponential inter-atomic decay of an
    insulator's density matrix: D_ab = (u_ab - 1/2) exp(-beta |R_A - R_B|) + delta_ab.'''
    import numpy as np
    rng = np.random.RandomState(seed)
    nao = mol.nao
    ao_atom = np.repeat(mol._bas[:, 0], np.diff(mol.ao_loc_nr()))
    R = mol.atom_coords()
    dist = np.linalg.norm(R[ao_atom][:, None] - R[ao_atom][None], axis=-1)
    d = (rng.rand(nao, nao) - .5) * np.exp(-beta * dist)
    return d + d.T + np.eye(nao)


def load_mol(name, basis, natoms=None):
    from gpu4pyscf_b200.gto.mole import M, read_xyz
    xyz = read_xyz(os.path.join(MOLDIR, name + '.xyz'))
    if natoms:
        xyz = '\n'.join(xyz.splitlines()[:natoms])
    return M(atom=xyz, basis=basis)


# ----------------------------------------------------------------------
class ClockSampler(threading.Thread):
    '''SM clock and throttle reasons sampled DURING the timed region (NVML).'''

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.sm_max,
                'reasons': sorted(self.reasons), 'samples': len(s)}


def physical_gpu_index(local):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local])
        except Exception:
            return local
    return local


def metric_name(args):
    if (args.mol, args.basis.lower()) == ('113_Taxol', 'def2-tzvp'):
        return METRIC
    return 'J/K build s/iter (direct SCF, %s %s)' % (args.mol, args.basis)


def workload_config(mol_name, basis, world, mol=None):
    """`config` of the JSON line -- the same keys and values for both arms (built from the molecule alone)."""
    from gpu4pyscf_b200.gto.mole import SortedGTO
    mol = load_mol(mol_name, basis) if mol is None else mol
    smol = SortedGTO.from_mol(mol)
    return {'workload': '%s RHF/%s direct-SCF J/K build (hermi=1, n_dm=1, direct_scf_tol=1e-13)' % (mol_name, basis),
            'density': 'synthetic symmetric, exp(-0.8 R_AB/Bohr) decay + identity (fixed seed; no SCF driver in the '
                       'product scope, the test-side RHF loop of tests/scf_driver.py covers converged densities)',
            'n_gpus': world,
            'nao': mol.nao, 'nao_sorted_cart': smol.nao, 'nbas_sorted': smol.nbas,
            'parallelism': 'light class launches whole on one rank (by measured cost), work items (X pair, Y chunk / '
                           'row block) of the heavy launches dealt round-robin over %d rank(s), one all-reduce of [J|K]'
                           % world,
            'l2': 'inputs+outputs (D,J,K = %.0f MB) exceed L2; kernels are FP64-compute bound'
                  % (3 * smol.nao * smol.nao * 8 / 1e6)}


def committed_counters(mol_name, basis):
    """Hardware counters of ONE build of this workload summed over all class launches, captured with ncu
    (tools/ncu_build_counters.py -> profiles/ncu_r2/build_counters_<mol>_<basis>.json).  bench.py cannot read
    DRAM or SASS counters itself; the file states the command it came from.  None when absent."""
    path = os.path.join(ROOT, 'profiles', 'ncu_r2', 'build_counters_%s_%s.json' % (mol_name, basis.lower()))
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def measured_cpu_runs(dev=None):
    """Full-molecule CPU-oracle runs that were actually measured (not extrapolated): the golden fixtures
    under tests/golden record the oracle's wall time (8 cores of the build container, no screening).
    With a device, the same J/K build is timed on the GPU next to it."""
    import glob
    import numpy as np
    out = []
    for path in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'jk_*.npz'))):
        g = np.load(path)
        parts = os.path.basename(path)[3:-4].split('_')
        natoms = int(parts.pop()) if parts[-1].isdigit() and not parts[-2].isdigit() else None
        basis = parts.pop()
        name = '_'.join(parts)
        row = {'mol': name, 'basis': basis, 'natoms': natoms, 'nao': int(g['nao']),
               'cpu_oracle_seconds': float(g['oracle_seconds']), 'cpu_cores': 8,
               'cpu_note': 'oracle/md_jk.c, OpenMP, no screening, build container (tests/golden/make_golden_jk.py)'}
        if dev is not None:
            try:
                import torch
                from gpu4pyscf_b200.scf import jk
                mol = load_mol(name, basis, natoms)
                dm = torch.as_tensor(decaying_density(mol), device=dev)
                opt = jk._VHFOpt(mol).build()
                opt.world, opt.rank = 1, 0
                for _ in range(2):
                    jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
                e1.record()
                e1.synchronize()
                row['gpu_seconds'] = e0.elapsed_time(e1) * 1e-3
            except Exception as e:          # a missing basis must not kill the bench line
                row['gpu_error'] = str(e)[:200]
        out.append(row)
    return out


# ----------------------------------------------------------------------
def cpu_oracle_sample(natoms, basis, mol_name, tol=1e-13):
    '''Time the CPU oracle on the first `natoms` atoms of the workload molecule.
    Returns (seconds, quartets, threads, nao).'''
    import numpy as np
    from oracle import oracle
    mol = load_mol(mol_name, basis, natoms)
    dm = decaying_density(mol)
    t0 = time.time()
    _, _, n = oracle.get_jk(mol, dm, hermi=1, screen=True, tol=tol, return_count=True)
    dt = time.time() - t0
    return dt, int(n), int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1)), mol.nao, oracle.last_flops()


def algorithmic_flops(opt, dmt_sorted, n_dm=1):
    '''One counting build: per-class quartet counts (kernel counter) and live
    per-class kernel times -> F_alg and the per-class table.'''
    import torch
    from gpu4pyscf_b200.lib import loader
    from gpu4pyscf_b200.scf import flops
    lib = loader.load_library()
    counter = torch.zeros(1, dtype=torch.int64, device=dmt_sorted.device)
    opt.counter = counter
    lib.B200_set_counter(ctypes.c_void_p(counter.data_ptr()))
    opt.class_timing = {}
    opt.get_jk(dmt_sorted, 1)
    torch.cuda.synchronize()
    rows = opt.class_timing
    opt.class_timing = None
    opt.counter = None
    lib.B200_set_counter(None)
    f_alg = 0.
    table = {}
    for k, v in rows.items():
        fl = v[3] * flops.f_quartet(k, n_dm)
        f_alg += fl
        a = table.setdefault(k[:4], [0., 0, 0.])
        a[0] += v[0]
        a[1] += v[3]
        a[2] += fl
    return f_alg, table, len(rows)


def dfma_peak(device, lib):
    '''FP64 FMA peak of this GPU right now (TFLOP/s): best of 5 burns.'''
    import torch
    sink = torch.zeros(8, dtype=torch.float64, device=device)
    blocks, threads, iters = 148 * 8, 256, 1 << 15
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 0.
    for rep in range(6):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.B200_dfma_burn(ctypes.c_void_p(sink.data_ptr()), iters, blocks, threads, stream)
        e1.record()
        e1.synchronize()
        if rep:
            best = max(best, 16. * iters * blocks * threads / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# ----------------------------------------------------------------------
def run_reference(args, rank):
    '''CPU arm: PySCF's get_jk is not installable here (no pyscf wheel, no
    network); the restated CPU oracle is timed instead, on a bounded sample.'''
    if rank != 0:
        return
    if 'TORCHELASTIC_RUN_ID' in os.environ or os.environ.get('OMP_NUM_THREADS') == '1':
        os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)     # torchrun pins it to 1
    times = []
    cpu_oracle_sample(2, args.basis, args.mol)       # warm-up: loads the library, spins up OpenMP
    for _ in range(max(1, min(args.steps, 3))):
        t, n, threads, nao, f_c = cpu_oracle_sample(args.cpu_atoms, args.basis, args.mol)
        times.append(t)
    t = sum(times) / len(times)
    f_full = F_ALG_FULL.get((args.mol, args.basis.lower()))
    scale = f_full / f_c if f_full else 1.
    sample = ('CPU oracle (oracle/md_jk.c, OpenMP; restated libcvhf-style direct J/K, PySCF itself is not installable '
              'here) on the first %d atoms of %s/%s (nao %d, %d quartets after Schwarz+density screening at 1e-13, '
              'F_alg %.3e) took %.2f s per build; value = that time x F_alg(full build)/F_alg(sample) = x %.1f'
              % (args.cpu_atoms, args.mol, args.basis, nao, n, f_c, t, scale))
    v = t * scale
    out = {
        'impl': 'reference', 'metric': metric_name(args), 'value': v, 'unit': 's', 'n_gpus': args.gpus,
        'steps': len(times), 'warmup': 1, 'ms_per_step': v * 1e3, 'higher_is_better': False,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.mol, args.basis, args.gpus),
        'extrapolated_from_sample': bool(f_full),
        'cpu_baseline': {'value': v, 'unit': 's', 'cores': threads, 'kind': 'port', 'sample': sample,
                         'sample_seconds': t, 'sample_gflops': f_c / t / 1e9},
        'e2e': {'value': v, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--mol', default='113_Taxol')
    ap.add_argument('--basis', default='def2-tzvp')
    ap.add_argument('--cpu-atoms', type=int, default=6)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--table', default=None, help='write the per-class table (json) here')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: the J/K path has no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    from gpu4pyscf_b200.scf import jk
    from gpu4pyscf_b200.lib import loader
    lib = loader.load_library()

    mol = load_mol(args.mol, args.basis)
    nao = mol.nao
    dm_host = decaying_density(mol)
    opt = jk._VHFOpt(mol).build()
    nao_s = opt.sorted_mol.nao
    dm_pinned = torch.as_tensor(dm_host).pin_memory()
    dmt = dm_pinned.to(dev)
    dms = opt.apply_coeff_C_mat_CT(dmt.reshape(1, nao, nao)).contiguous()

    rank_ms = {}      # per-rank device time of the timed region (ms per step), N > 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.synchronize()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            rank_ms[fn.__name__] = [float(x) / steps for x in allms]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        return opt.get_jk(dms, 1)

    out_j = torch.empty((nao, nao), dtype=torch.float64).pin_memory()
    out_k = torch.empty((nao, nao), dtype=torch.float64).pin_memory()

    def step_e2e():
        d = dm_pinned.to(dev, non_blocking=True)
        vj, vk = jk.get_jk(mol, d, hermi=1, vhfopt=opt)
        out_j.copy_(vj, non_blocking=True)
        out_k.copy_(vk, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # ---- warm-up (also compiles nothing: all kernels are prebuilt) ----
    for _ in range(max(args.warmup, 3)):
        step_resident()
    peak = dfma_peak(dev, lib)

    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps

    # ---- flop accounting + per-class live timing (one extra, untimed build) ----
    f_alg, table, nlaunch = algorithmic_flops(opt, dms)
    fl = torch.tensor([f_alg], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(fl)
    f_alg_total = float(fl.item())
    kern_ms = sum(a[0] for a in table.values())
    top = max(table.items(), key=lambda kv: kv[1][0])
    # the rys class kernels are >99% of the step; "dominant kernel" = the class family
    kern_tf = f_alg / (kern_ms * 1e-3) / 1e12 if kern_ms else 0.

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            t, n, threads, nao_c, f_c = cpu_oracle_sample(args.cpu_atoms, args.basis, args.mol)
            cpu = {'value': t * f_alg_total / f_c, 'unit': 's', 'cores': threads, 'kind': 'port',
                   'sample': 'CPU oracle (oracle/md_jk.c, OpenMP, restated libcvhf-style direct J/K, not PySCF) on the '
                             'first %d atoms of %s/%s (nao %d, %d quartets, F_alg %.3e) took %.2f s; value = that time x '
                             'F_alg(full)/F_alg(sample) = x %.1f' % (args.cpu_atoms, args.mol, args.basis, nao_c, n, f_c, t,
                                                                     f_alg_total / f_c),
                   'sample_seconds': t, 'sample_gflops': f_c / t / 1e9,
                   'measured_full_molecule_runs': measured_cpu_runs(dev)}
        nbytes = nao * nao * 8
        # F_hw and DRAM traffic come from a committed ncu capture of the same workload (per build, all launches)
        ctr = committed_counters(args.mol, args.basis) if world == 1 else None
        f_hw = ctr.get('f_hw') if ctr else None
        f_eff = min(f_alg_total, f_hw) if f_hw else f_alg_total
        achieved = f_eff / (ms_step * 1e-3) / 1e12
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12                 # 148 SMs x 64 DFMA lanes x 2 flop x 1.965 GHz
        out = {
            'metric': metric_name(args), 'value': ms_step * 1e-3, 'unit': 's', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': False,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.mol, args.basis, world, mol),
            'tflops_fp64': achieved,
            'f_alg': f_alg_total,
            'f_hw': f_hw,
            'roofline': {'bound': 'fp64_fma', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / (peak * world) if peak else None,
                         'frac_of_nominal': achieved / (nominal * world),
                         'nominal_peak': nominal,
                         'achieved_is': 'min(F_alg, F_hw) / t; F_alg = SURVEY 8d flop model x quartets counted by the '
                                        'kernels in this run, F_hw = 2*DFMA + DMUL + DADD thread instructions of one '
                                        'build (ncu, committed capture)' if f_hw else
                                        'F_alg / t (no committed ncu capture for this workload / rank count)',
                         'traffic': ctr.get('dram_bytes') if ctr else None,
                         'traffic_unit': 'DRAM bytes per build, all class launches (ncu dram__bytes_read+write)',
                         'traffic_source': ctr.get('source') if ctr else None,
                         'algorithmic_bytes': 3 * nao_s * nao_s * 8,
                         'peak_source': 'DFMA burn measured in this run on rank 0 (MEASURED_PEAKS.json has no FP64 figure)',
                         'kernel': 'rys_jk_{thread,rowt,warp}_kernel class family (%d launches, %.1f ms of the %.1f ms step on rank 0)'
                                   % (nlaunch, kern_ms, ms_step),
                         'kernel_achieved': kern_tf,
                         'top_class': {'class': '(%d%d|%d%d)' % top[0], 'ms': top[1][0], 'quartets': top[1][1],
                                       'tflops': top[1][2] / (top[1][0] * 1e-3) / 1e12 if top[1][0] else None}},
            'cpu_baseline': cpu,
            'e2e': {'value': ms_e2e * 1e-3, 'unit': 's', 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': 2 * nbytes},
            'rank_ms_per_step': ({'max': max(rank_ms['step_resident']), 'min': min(rank_ms['step_resident']),
                                  'mean': sum(rank_ms['step_resident']) / world} if rank_ms.get('step_resident') else None),
            'gpu_launches': (nlaunch + 1) * args.steps,   # class kernels + condense per build
            'clocks': clocks,
        }
        print(json.dumps(out), flush=True)
        if args.table:
            with open(args.table, 'w') as f:
                json.dump({'(%d%d|%d%d)' % k: v for k, v in sorted(table.items(), key=lambda kv: -kv[1][0])}, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
