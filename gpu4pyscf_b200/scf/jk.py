"""Direct-SCF Coulomb/exchange plugin -- drop-in for gpu4pyscf/scf/jk.py.

Public surface kept from the reference (gpu4pyscf/scf/jk.py:60-230, 407-927):
``get_jk / get_j / get_k`` and ``_VHFOpt`` with ``build``, ``get_jk``, ``get_j``,
``get_k``, ``apply_coeff_C_mat_CT``, ``apply_coeff_CT_mat_C``, ``sorted_mol``,
``bas_pair_cache``, ``rys_envs``, ``h_shls``, ``coeff``, ``direct_scf_tol``.

What is new underneath: host arrays are torch CUDA tensors (CuPy is not in the
image; the C-ABI takes raw device pointers so it is array-library agnostic),
the screening data are exactly q-sorted shell-pair lists per group pair, and one
launch handles an unordered pair of group pairs {(gi,gj),(gk,gl)} -- which of the
two is the register-tile side of the kernel is the library's choice
(B200_has_class).  Multi-GPU: one process per GPU; every rank takes the work
items (X pair, chunk of the q-sorted Y list) with ``item % world == rank`` of each
launch (neighbouring items have near-equal cost, so the round-robin deal is
balanced) and the partial ``[vj|vk]`` buffers are summed with ONE ``all_reduce``
per build.
"""
import ctypes
import math
import os
import numpy as np
import torch

from ..gto.mole import SortedGTO, RysIntEnvVars, _scale_sp_ctr_coeff, ANG_OF, NPRIM_OF
from ..lib import loader

__all__ = ['get_jk', 'get_j', 'get_k', '_VHFOpt']

LMAX = 4


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _as_device_tensor(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64)
    if hasattr(a, '__cuda_array_interface__') and device.type == 'cuda':
        return torch.as_tensor(a, device=device).to(torch.float64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)


def get_jk(mol, dm, hermi=0, vhfopt=None, with_j=True, with_k=True,
           omega=None, lr_factor=None, sr_factor=None, verbose=None):
    '''Coulomb (J) and exchange (K) matrices for one or several density
    matrices; same contract as gpu4pyscf.scf.jk.get_jk (jk.py:60-151).
    Returns arrays shaped like ``dm`` living on the device.'''
    assert with_j or with_k
    if vhfopt is None:
        vhfopt = _VHFOpt(mol).build()
    if omega is None:
        omega = mol.omega
    dm = _as_device_tensor(dm, vhfopt.device)
    nao_orig = dm.shape[-1]
    dms = dm.reshape(-1, nao_orig, nao_orig)
    dms = vhfopt.apply_coeff_C_mat_CT(dms).contiguous()
    # Range separation (reference: jk.py:1122-1168).  The kernels evaluate the full
    # operator (omega = 0) or the long-range one erf(|omega| r)/r; everything else is
    # a linear combination:  SR = full - LR,  lr_factor*LR + sr_factor*SR =
    # sr_factor*full + (lr_factor - sr_factor)*LR.
    if omega == 0 or omega is None:
        c_full, c_lr = 1., 0.
        if lr_factor is not None or sr_factor is not None:
            raise ValueError('lr_factor/sr_factor need omega != 0')
    elif lr_factor is None and sr_factor is None:
        c_full, c_lr = (0., 1.) if omega > 0 else (1., -1.)
    else:
        lr = 0. if lr_factor is None else float(lr_factor)
        sr = 0. if sr_factor is None else float(sr_factor)
        c_full, c_lr = sr, lr - sr
    vj = vk = None
    for coef, w in ((c_full, 0.), (c_lr, abs(float(omega or 0.)))):
        if coef == 0:
            continue
        tj, tk = vhfopt.get_jk(dms, hermi, verbose, with_j=with_j, with_k=with_k, omega=w)
        if tj is not None:
            vj = coef * tj if vj is None else vj + coef * tj
        if tk is not None:
            vk = coef * tk if vk is None else vk + coef * tk
    if with_k:
        vk = vhfopt.apply_coeff_CT_mat_C(vk).reshape(dm.shape) if vk is not None else torch.zeros_like(dm)
    if with_j:
        if vj is None:      # hermi == 2 (J of an anti-symmetric density vanishes) or zero coefficients
            vj = torch.zeros_like(dm)
        else:
            vj = vhfopt.apply_coeff_CT_mat_C(vj).reshape(dm.shape)
    else:
        vj = None
    return vj, vk


def get_k(mol, dm, hermi=0, vhfopt=None, omega=None, lr_factor=None, sr_factor=None, verbose=None):
    '''K only (reference: jk.py:153-204)'''
    return get_jk(mol, dm, hermi, vhfopt, False, True, omega, lr_factor, sr_factor, verbose)[1]


def get_j(mol, dm, hermi=0, vhfopt=None, verbose=None):
    '''J only (reference: jk.py:206-230)'''
    return get_jk(mol, dm, hermi, vhfopt, True, False, None, None, None, verbose)[0]


class _VHFOpt:
    '''Screening object of the direct-SCF J/K build (reference: jk.py:407-927).'''

    def __init__(self, mol, direct_scf_tol=1e-13, tile=1):
        self.mol = mol
        self.direct_scf_tol = direct_scf_tol
        self.tile = tile
        self.sorted_mol = None
        self.h_shls = []
        self.bas_pair_cache = None
        self._rys_envs = None
        self.device = None
        # multi-GPU partition (set by the caller or picked up from torch.distributed)
        self.rank = None
        self.world = None
        # set to a dict to collect per-class kernel times (ms), like the
        # reference's _TimingCollector (jk.py:1170-1188); costs a sync per class
        self.class_timing = None
        self.n_streams = int(os.environ.get('B200_STREAMS', '4'))
        # width (log units) of the q bins of the pair lists; 0 = exactly q-sorted lists
        self.q_bin = float(os.environ.get('B200_Q_BIN', '0.5'))
        self._streams = None

    def reset(self, mol):
        self.mol = mol
        self.sorted_mol = None
        self.bas_pair_cache = None
        self._rys_envs = None
        self._tasks = None
        self._qmax_cache = None
        return self

    # ------------------------------------------------------------------
    def build(self, group_size=None, verbose=None):
        lib = loader.load_library()
        self.device = loader.device()
        mol = self.mol
        smol = SortedGTO.from_mol(mol, group_size=group_size)
        self.sorted_mol = smol
        ls = smol._bas[:, ANG_OF]
        if ls.max() > LMAX:
            self.h_shls = np.where(ls > LMAX)[0].tolist()
            raise NotImplementedError(
                f'shells with l > {LMAX} are not supported on the GPU and this '
                'implementation has no CPU fallback')
        self._make_rys_envs()
        self.bas_pair_cache = _cache_q_cond_and_non0pairs(
            smol, self._rys_envs, self.direct_scf_tol, self.device, lib, self.q_bin)
        return self

    def _make_rys_envs(self):
        smol = self.sorted_mol
        dev = self.device
        env = _scale_sp_ctr_coeff(smol)
        atm = torch.as_tensor(np.ascontiguousarray(smol._atm, dtype=np.int32), device=dev)
        bas = torch.as_tensor(np.ascontiguousarray(smol._bas, dtype=np.int32), device=dev)
        envd = torch.as_tensor(np.ascontiguousarray(env, dtype=np.float64), device=dev)
        ao_loc = torch.as_tensor(np.ascontiguousarray(smol.ao_loc, dtype=np.int32), device=dev)
        self._rys_envs = RysIntEnvVars.new(smol.natm, smol.nbas, atm, bas, envd, ao_loc)

    @property
    def rys_envs(self):
        return self._rys_envs

    @property
    def coeff(self):
        return self.sorted_mol._coeff_on(self.device)

    def apply_coeff_C_mat_CT(self, mat):
        return self.sorted_mol.apply_C_mat_CT(mat)

    def apply_coeff_CT_mat_C(self, mat):
        return self.sorted_mol.apply_CT_mat_C(mat)

    def _partition(self):
        if self.world is not None:
            return self.rank or 0, self.world
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_rank(), torch.distributed.get_world_size()
        return 0, 1

    # ------------------------------------------------------------------
    def class_tasks(self):
        '''Launch list: (group pair Y, group pair X, dedupe) for every unordered
        pair of group pairs.  Y is the lane side and X the register-tile side of
        the kernel; the library says which orientation of a class it instantiates.'''
        if getattr(self, '_tasks', None) is not None:
            return self._tasks
        lib = loader.load_library()
        smol = self.sorted_mol
        uniq = smol.uniq_l_ctr
        keys = sorted(self.bas_pair_cache.keys())

        def ls(key):
            return int(uniq[key[0], 0]), int(uniq[key[1], 0])

        tasks = []
        for ia, ka in enumerate(keys):
            for kb in keys[:ia + 1]:
                if ka == kb:
                    tasks.append((ka, ka, 1))
                elif lib.B200_has_class(*ls(kb), *ls(ka)):
                    tasks.append((kb, ka, 0))      # (Y, X)
                else:
                    tasks.append((ka, kb, 0))
        # optional longest-processing-time order; measured slower than the natural group order on
        # Taxol (6.77 s vs 6.63 s), so it is off by default
        if os.environ.get('B200_SORT_TASKS', '0') == '1':
            from . import flops

            def cost(task):
                ky, kx, _ = task
                lprim = tuple(int(v) for v in uniq[[ky[0], ky[1], kx[0], kx[1]], 0]) + \
                    tuple(int(v) for v in uniq[[ky[0], ky[1], kx[0], kx[1]], 1])
                return len(self.bas_pair_cache[ky][0]) * len(self.bas_pair_cache[kx][0]) * flops.f_quartet(lprim)
            tasks.sort(key=cost, reverse=True)
        mode = os.environ.get('B200_SORT_TASKS', '0')
        if mode == '2':
            tasks.reverse()
        elif mode == '3':       # launches of the same kernel instantiation next to each other
            tasks.sort(key=lambda tk: ls(tk[0]) + ls(tk[1]))
        self._tasks = tasks
        return tasks

    def get_jk(self, dms, hermi=0, verbose=None, with_j=True, with_k=True, omega=0.):
        '''J/K in the sorted Cartesian basis.  dms: (n_dm, nao, nao) device
        tensor.  Returns (vj, vk) (None where not requested).  omega = 0: 1/r,
        omega > 0: long-range erf(omega r)/r.'''
        if omega < 0:
            raise ValueError('short range is composed by the caller as full - long range')
        lib = loader.load_library()
        smol = self.sorted_mol
        dev = self.device
        nao = smol.nao
        nbas = smol.nbas
        dms = dms.reshape(-1, nao, nao)
        n_dm = dms.shape[0]
        if hermi == 0:
            # K(D) needs the four "other half" quartet images: evaluate with D^T too
            # (same trick as the reference, jk.py:534-536, 594-605)
            dms = torch.cat([dms, dms.transpose(1, 2)], dim=0)
        dms = dms.contiguous()
        ndm_k = dms.shape[0]
        if hermi == 2:
            with_j = False
        envs = self._rys_envs
        stream = loader.current_stream_ptr()
        # Class launches are independent (they only meet in the J/K atomics), so
        # they are dealt over a few side streams: the tail of one launch, when
        # most SMs have run out of work items, overlaps the start of the next.
        side = None
        if dev.type == 'cuda' and self.class_timing is None and self.n_streams > 1:
            if self._streams is None:
                self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.n_streams)]
            side = self._streams
            main = torch.cuda.current_stream()

        dm_cond = torch.empty((nbas, nbas), dtype=torch.float32, device=dev)
        err = lib.B200_condense(_ptr(dm_cond), _ptr(dms), ndm_k, nao, _ptr(envs._keep[3]), nbas, stream)
        if err:
            raise RuntimeError('B200_condense failed')
        dm_penalty = float(dm_cond.max())
        cutoff = math.log(self.direct_scf_tol)

        vj = torch.zeros_like(dms) if with_j else None
        vk = torch.zeros_like(dms) if with_k else None
        tasks = self.class_tasks()
        if getattr(self, 'counter', None) is not None:
            self._counter_last = int(self.counter.item())
        heads = torch.zeros(len(tasks) + 1, dtype=torch.int32, device=dev)
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            for st in side:
                st.wait_event(ev)
        rank, world = self._partition()
        uniq = smol.uniq_l_ctr
        lprim = (ctypes.c_int * 8)()
        only = getattr(self, 'only_classes', None)   # measurement: restrict to some classes
        for it, (ky, kx, dedupe) in enumerate(tasks):
            if only is not None:
                l4 = tuple(int(v) for v in uniq[[ky[0], ky[1], kx[0], kx[1]], 0])
                p4 = tuple(int(v) for v in uniq[[ky[0], ky[1], kx[0], kx[1]], 1])
                if l4 not in only and l4 + p4 not in only:
                    continue
            pair_x, q_x = self.bas_pair_cache[kx][:2]
            pair_y, q_y = self.bas_pair_cache[ky][:2]
            rec_y = self.bas_pair_cache[ky][3]
            if len(pair_x) == 0 or len(pair_y) == 0:
                continue
            # cheap class-level early exit
            if float(self._qmax[kx]) + float(self._qmax[ky]) + dm_penalty < cutoff:
                continue
            lprim[0], lprim[1] = int(uniq[ky[0], 0]), int(uniq[ky[1], 0])
            lprim[2], lprim[3] = int(uniq[kx[0], 0]), int(uniq[kx[1], 0])
            lprim[4], lprim[5] = int(uniq[ky[0], 1]), int(uniq[ky[1], 1])
            lprim[6], lprim[7] = int(uniq[kx[0], 1]), int(uniq[kx[1], 1])
            head = ctypes.c_void_p(heads.data_ptr() + 4 * it)
            if self.class_timing is not None and dev.type == 'cuda':
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            if side is not None:
                stream = ctypes.c_void_p(side[it % len(side)].cuda_stream)
            err = lib.B200_build_jk_class(
                _ptr(vj), _ptr(vk), _ptr(dms), ndm_k, nao, ctypes.byref(envs), lprim,
                _ptr(pair_x), _ptr(q_x), len(pair_x), _ptr(pair_y), _ptr(q_y), len(pair_y),
                _ptr(dm_cond), cutoff, dm_penalty, float(self.q_bin), dedupe, rank, world, head, stream, _ptr(rec_y), float(omega))
            if err:
                raise RuntimeError('B200_build_jk_class failed for class %s' % (tuple(lprim),))
            if self.class_timing is not None and dev.type == 'cuda':
                ev1.record()
                ev1.synchronize()
                key = tuple(lprim)
                rec = self.class_timing.setdefault(key, [0., 0, 0, 0])
                rec[0] += ev0.elapsed_time(ev1)
                rec[1] += len(pair_x)
                rec[2] += len(pair_y)
                if getattr(self, 'counter', None) is not None:
                    n = int(self.counter.item())
                    rec[3] += n - self._counter_last
                    self._counter_last = n

        if side is not None:
            for st in side:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)

        if world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            # one collective per build over the concatenated [vj|vk] partials
            bufs = [b for b in (vj, vk) if b is not None]
            flat = torch.cat([b.reshape(-1) for b in bufs])
            torch.distributed.all_reduce(flat)
            off = 0
            for b in bufs:
                b.copy_(flat[off:off + b.numel()].reshape(b.shape))
                off += b.numel()

        # symmetrisation of the partial sums (reference: jk.py:594-616)
        if with_k:
            if hermi == 1:
                vk = vk + vk.transpose(1, 2)
            elif hermi == 2:
                vk = vk - vk.transpose(1, 2)
            else:
                vk = vk[:n_dm] + vk[n_dm:].transpose(1, 2)
        if with_j:
            if hermi == 1:
                vj = vj * 2.
            else:
                vj = vj[:n_dm] + vj[n_dm:]
            vj = vj + vj.transpose(1, 2)
        return vj, vk

    def get_j(self, dms, verbose=None):
        return self.get_jk(dms, 1, verbose, with_j=True, with_k=False)[0]

    def get_k(self, dms, hermi=0, verbose=None, omega=0, lr_factor=None, sr_factor=None):
        if not omega:
            return self.get_jk(dms, hermi, verbose, with_j=False, with_k=True)[1]
        if lr_factor is None and sr_factor is None:
            c_full, c_lr = (0., 1.) if omega > 0 else (1., -1.)
        else:
            lr = 0. if lr_factor is None else float(lr_factor)
            sr = 0. if sr_factor is None else float(sr_factor)
            c_full, c_lr = sr, lr - sr
        vk = None
        for coef, w in ((c_full, 0.), (c_lr, abs(float(omega)))):
            if coef == 0:
                continue
            tk = self.get_jk(dms, hermi, verbose, with_j=False, with_k=True, omega=w)[1]
            vk = coef * tk if vk is None else vk + coef * tk
        return vk

    @property
    def _qmax(self):
        if getattr(self, '_qmax_cache', None) is None:
            self._qmax_cache = {k: (float(v[1].max()) if len(v[1]) else -1e30)
                                for k, v in self.bas_pair_cache.items()}
        return self._qmax_cache


def _cache_q_cond_and_non0pairs(smol, rys_envs, precision, device, lib, q_bin=0.):
    '''Per group pair (gi >= gj): shell-pair ids ``ish*nbas+jsh`` (ish >= jsh),
    their Schwarz bounds q = 1/2 log max|(ij|ij)| and the same again as
    ``s_cond`` -- each list sorted by q descending (exactly; the reference bins
    by 4 log-units, jk.py:984-1120).  Pairs that cannot survive against the
    largest bound of the molecule are dropped.'''
    nbas = smol.nbas
    uniq = smol.uniq_l_ctr
    counts = np.asarray(smol.l_ctr_counts)
    offs = np.append(0, np.cumsum(counts))
    ngroups = len(counts)
    log_cutoff = math.log(precision)
    stream = loader.current_stream_ptr()
    raw = {}
    for gi in range(ngroups):
        for gj in range(gi + 1):
            ish = torch.arange(offs[gi], offs[gi + 1], device=device, dtype=torch.int64)
            jsh = torch.arange(offs[gj], offs[gj + 1], device=device, dtype=torch.int64)
            ii, jj = torch.meshgrid(ish, jsh, indexing='ij')
            ids = (ii * nbas + jj).reshape(-1)
            if gi == gj:
                ids = ids[(ii >= jj).reshape(-1)]
            ids = ids.to(torch.int32).contiguous()
            q = torch.empty(len(ids), dtype=torch.float32, device=device)
            err = lib.B200_qcond(_ptr(q), ctypes.byref(rys_envs), _ptr(ids), len(ids),
                                 int(uniq[gi, 0]), int(uniq[gj, 0]),
                                 int(uniq[gi, 1]), int(uniq[gj, 1]), stream)
            if err:
                raise RuntimeError('B200_qcond failed')
            raw[gi, gj] = (ids, q)
    qmax = max(float(q.max()) for _, q in raw.values() if len(q))
    cache = {}
    for key, (ids, q) in raw.items():
        keep = q + qmax > log_cutoff
        ids, q = ids[keep], q[keep]
        if q_bin > 0:
            # descending q bins of width q_bin, shell-pair order (ish, jsh) inside a bin:
            # neighbours in the list share the bra shell, so the D rows they load and the
            # K rows they update coincide (the kernels get q_margin = q_bin for the early exit)
            bins = torch.floor(q.double() / q_bin).to(torch.int64)
            order = torch.argsort((bins.max() - bins) * (int(nbas) * int(nbas)) + ids.to(torch.int64))
        else:
            order = torch.argsort(q, descending=True, stable=True)
        ids = ids[order].contiguous()
        q = q[order].contiguous()
        # Gaussian-product records of the pairs, in list order
        npi, npj = int(uniq[key[0], 1]), int(uniq[key[1], 1])
        rec = torch.empty(len(ids) * lib.B200_pair_rec_stride(npi, npj), dtype=torch.float64, device=device)
        err = lib.B200_pair_records(_ptr(rec), ctypes.byref(rys_envs), _ptr(ids), len(ids), npi, npj, stream)
        if err:
            raise RuntimeError('B200_pair_records failed')
        cache[key] = (ids, q, q, rec)
    return cache
