"""Direct-SCF Coulomb/exchange plugin -- drop-in for gpu4pyscf/scf/jk.py.

Public surface kept from the reference (gpu4pyscf/scf/jk.py:60-230, 407-927):
``get_jk / get_j / get_k`` and ``_VHFOpt`` with ``build``, ``get_jk``, ``get_j``,
``get_k``, ``apply_coeff_C_mat_CT``, ``apply_coeff_CT_mat_C``, ``sorted_mol``,
``bas_pair_cache``, ``rys_envs``, ``h_shls``, ``coeff``, ``direct_scf_tol``.

What is new underneath: host arrays are torch CUDA tensors (CuPy is not in the
image; the C-ABI takes raw device pointers so it is array-library agnostic).
Every group pair carries two views of its significant shell pairs: a flat list
sorted by the Schwarz bound q (the resident X side of a launch) and a ROW list
grouped by one shell of the pair (the streamed Y side of the row-stream kernels:
three shells of every quartet a warp evaluates are warp-uniform).  One launch
handles an unordered pair of group pairs; which of the two is streamed, and
through which kernel family, is decided per launch (``class_tasks``).
Screening lists exist per OPERATOR (``_VHFOpt._use_lists``): 1/r, erf(w r)/r, or a sum
of both evaluated in one pass (range-separated exchange), each with its own Schwarz
bounds; pure short-range K builds add the distance-dependent quartet filter.
Multi-GPU: one process per GPU; the work items of the heavy launches are dealt
``item % world == rank``, light launches run whole on one rank (assigned by measured
cost), and the partial ``[vj|vk]`` buffer is summed with ONE ``all_reduce`` per build.
"""
import ctypes
import math
import os
import time
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


def _check_rsh_factors(mol, omega, lr_factor, sr_factor):
    '''(omega, lr_factor, sr_factor) of the exchange operator
    lr_factor * erf(|omega| r)/r + sr_factor * erfc(|omega| r)/r, normalised exactly like
    the reference (gpu4pyscf/scf/jk.py:1122-1168): omega < 0 flags that the short-range
    part is needed, lr == sr collapses to omega = 0 with that factor.'''
    if omega is None:
        omega = mol.omega
    if lr_factor is None and sr_factor is None:
        if omega == 0:
            lr_factor = sr_factor = 1
        elif omega < 0:
            lr_factor, sr_factor = 0, 1
        else:
            lr_factor, sr_factor = 1, 0
    elif lr_factor is None:
        lr_factor = sr_factor if omega == 0 else 0
    elif sr_factor is None:
        if omega == 0:
            sr_factor = lr_factor
        else:
            assert omega > 0
            sr_factor = 0
    else:
        if omega == 0:
            assert lr_factor == sr_factor
        elif lr_factor == sr_factor:
            omega = 0
    if sr_factor != 0 and omega != 0:
        omega = -abs(omega)
    return omega, lr_factor, sr_factor


def _operator_terms(omega, lr_factor, sr_factor):
    '''[(coefficient, w)] with w = 0 for 1/r and w > 0 for erf(w r)/r -- the two operators the
    kernels evaluate (reference: rys_roots_for_k.cu:25-60).'''
    if omega == 0:
        terms = [(float(lr_factor), 0.)]
    elif sr_factor == 0:
        terms = [(float(lr_factor), abs(float(omega)))]
    else:
        terms = [(float(sr_factor), 0.), (float(lr_factor) - float(sr_factor), abs(float(omega)))]
    return [(c, w) for c, w in terms if c != 0]


def get_jk(mol, dm, hermi=0, vhfopt=None, with_j=True, with_k=True,
           omega=None, lr_factor=None, sr_factor=None, verbose=None):
    '''Coulomb (J) and exchange (K) matrices for one or several density
    matrices; same contract as gpu4pyscf.scf.jk.get_jk (jk.py:60-151): omega, lr_factor and
    sr_factor define the operator of K only; J always uses the molecule's own operator.
    Returns arrays shaped like ``dm`` living on the device.'''
    assert with_j or with_k
    if vhfopt is None:
        vhfopt = _VHFOpt(mol).build()
    dm = _as_device_tensor(dm, vhfopt.device)
    nao_orig = dm.shape[-1]
    dms = dm.reshape(-1, nao_orig, nao_orig)
    dms = vhfopt.apply_coeff_C_mat_CT(dms).contiguous()
    omega, lr_factor, sr_factor = _check_rsh_factors(mol, omega, lr_factor, sr_factor)
    vj = vk = None
    if omega == 0 and lr_factor == 1 and mol.omega == 0:
        vj, vk = vhfopt._jk_pass(dms, hermi, with_j, with_k, 0.)
    else:
        if with_k:
            vk = vhfopt.get_k(dms, hermi, verbose, omega, lr_factor, sr_factor)
        if with_j:
            vj = vhfopt.get_j(dms if hermi == 1 else (dms + dms.transpose(1, 2)) * .5, verbose)
    if with_k:
        vk = vhfopt.apply_coeff_CT_mat_C(vk).reshape(dm.shape)
    else:
        vk = None
    if with_j:
        # hermi == 2: J of an anti-symmetric density vanishes
        vj = vhfopt.apply_coeff_CT_mat_C(vj).reshape(dm.shape) if vj is not None else torch.zeros_like(dm)
    else:
        vj = None
    return vj, vk


def get_k(mol, dm, hermi=0, vhfopt=None, omega=None, lr_factor=None, sr_factor=None, verbose=None):
    '''K only (reference: jk.py:153-204)'''
    if vhfopt is None:
        vhfopt = _VHFOpt(mol).build()
    dm = _as_device_tensor(dm, vhfopt.device)
    nao_orig = dm.shape[-1]
    dms = vhfopt.apply_coeff_C_mat_CT(dm.reshape(-1, nao_orig, nao_orig)).contiguous()
    vk = vhfopt.get_k(dms, hermi, verbose, omega, lr_factor, sr_factor)
    return vhfopt.apply_coeff_CT_mat_C(vk).reshape(dm.shape)


def get_j(mol, dm, hermi=0, vhfopt=None, verbose=None):
    '''J only (reference: jk.py:206-230)'''
    if vhfopt is None:
        vhfopt = _VHFOpt(mol).build()
    dm = _as_device_tensor(dm, vhfopt.device)
    nao_orig = dm.shape[-1]
    dms = vhfopt.apply_coeff_C_mat_CT(dm.reshape(-1, nao_orig, nao_orig)).contiguous()
    if hermi != 1:
        dms = (dms + dms.transpose(1, 2)) * .5
    vj = vhfopt.get_j(dms, verbose)
    return vhfopt.apply_coeff_CT_mat_C(vj).reshape(dm.shape)


def _nf(l):
    return (l + 1) * (l + 2) // 2


def _lpt_owners(cost, world, split_min=float('inf')):
    '''Longest-processing-time-first assignment of launches to ranks: heaviest launch first, each to
    the least loaded rank (ties by index, so every rank derives the same table).  Launches of cost
    >= split_min get owner -1: their work items are dealt over all ranks instead.  Plays the role of
    the reference's dynamic per-device task hand-out (lib/multi_gpu.py:126-206, jk.py:520-524).'''
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    load = [0.] * world
    owners = [0] * len(cost)
    for i in order:
        if cost[i] >= split_min:
            owners[i] = -1
            continue
        r = min(range(world), key=lambda k: (load[k], k))
        owners[i] = r
        load[r] += cost[i]
    return owners


class _PairLists:
    '''Screening data of one operator: per group pair (ids, q, s, records), the row views, the launch
    plan and the measured launch times of the multi-GPU partition.'''
    def __init__(self):
        self.cache = None
        self.rows = {}
        self.tasks = None
        self.qmax = None
        self.launch_cost = None
        self.sr = None            # short-range filter data: {'omega': w, 'flat': {key: (s, geom)}}


def _estimator_operator(terms):
    '''(omega, lr_factor, sr_factor) in the convention of int2e_qcond_estimator for the operator
    sum_k coef_k O[w_k], O[0] = 1/r, O[w] = erf(w r)/r: one term, or (c_full, 0) + (c_lr, w).'''
    if len(terms) == 1:
        c, w = terms[0]
        return (0., abs(c), abs(c)) if w == 0 else (w, abs(c), 0.)
    (c0, w0), (c1, w1) = terms
    assert w0 == 0 and w1 > 0
    return (-w1, c1 + c0, c0)          # sr = c_full, lr = c_lr + c_full


def _psd_pair(terms):
    '''can the two terms be screened with the Schwarz bound of their SUM?  Needs a positive
    semi-definite operator: lr*erf/r + sr*erfc/r with lr, sr >= 0'''
    (c0, w0), (c1, w1) = terms
    return w0 == 0 and w1 > 0 and c0 > 0 and c1 + c0 >= 0


class _VHFOpt:
    '''Screening object of the direct-SCF J/K build (reference: jk.py:407-927).'''

    def __init__(self, mol, direct_scf_tol=1e-13, tile=1):
        self.mol = mol
        self.direct_scf_tol = direct_scf_tol
        self.tile = tile
        self.sorted_mol = None
        self.h_shls = []
        self._rys_envs = None
        self.device = None
        # screening lists per operator: key = ((coef, w), ...) of the operator terms -> _PairLists; `_cur`
        # is the set the current pass works on (the molecule's own operator after build())
        self._lists = {}
        self._cur = _PairLists()
        self._default = None
        self._diffuse_dev = None
        # multi-GPU partition (set by the caller or picked up from torch.distributed)
        self.rank = None
        self.world = None
        # set to a dict to collect per-class kernel times (ms), like the
        # reference's _TimingCollector (jk.py:1170-1188); costs a sync per class
        self.class_timing = None
        self.n_streams = int(os.environ.get('B200_STREAMS', '4'))
        # width (log units) of the q bins of the flat pair lists; 0 = exactly q-sorted lists
        self.q_bin = float(os.environ.get('B200_Q_BIN', '0.5'))
        # row-stream kernels are used when the streamed list has rows at least this long on average
        self.row_min = float(os.environ.get('B200_ROW_MIN', '24'))
        self._streams = None

    # the attributes of the reference object (bas_pair_cache) and the launch plan derived from it
    # live in the _PairLists of the operator in use
    # bas_pair_cache (reference attribute): the lists of the molecule's own operator, whatever pass ran last
    bas_pair_cache = property(lambda self: self._default.cache if self._default is not None else None)
    _rows = property(lambda self: self._cur.rows, lambda self, v: setattr(self._cur, 'rows', v))
    _tasks = property(lambda self: self._cur.tasks, lambda self, v: setattr(self._cur, 'tasks', v))
    _qmax_cache = property(lambda self: self._cur.qmax, lambda self, v: setattr(self._cur, 'qmax', v))
    _launch_cost = property(lambda self: self._cur.launch_cost,
                            lambda self, v: setattr(self._cur, 'launch_cost', v))

    def reset(self, mol):
        self.mol = mol
        self.sorted_mol = None
        self._rys_envs = None
        self._lists = {}
        self._cur = _PairLists()
        self._default = None
        self._diffuse_dev = None
        return self

    def _use_lists(self, terms):
        '''Make the screening lists of the operator sum_k coef_k O[w_k] current, building them on first
        use: Schwarz bounds of THAT operator (reference: _VHFOpt.build under mol.with_range_coulomb(omega),
        hf.py:936-944, computes q_cond with the attenuated operator), pair lists, launch plan.'''
        key = tuple((float(c), float(w)) for c, w in terms)
        pl = self._lists.get(key)
        if pl is None:
            lib = loader.load_library()
            pl = _PairLists()
            pl.cache = _cache_q_cond_and_non0pairs(
                self.sorted_mol, self._rys_envs, self.direct_scf_tol, self.device, lib, self.q_bin,
                _estimator_operator(key))
            if len(key) == 2 and key[0][1] == 0 and key[1][0] == -key[0][0]:
                # pure short-range operator c * erfc(w r)/r: pair estimators for the quartet filter
                # (reference: fill_s_estimator under omega < 0, scf/jk.py:1089-1105)
                w = key[1][1]
                pl.sr = {'omega': w, 'flat': {k: self._sr_pair_data(v[0], w) for k, v in pl.cache.items()}}
            self._lists[key] = pl
        self._cur = pl
        return pl

    def _sr_pair_data(self, ids, omega):
        '''(s, geom) of a pair list for the short-range quartet filter (B200_sr_pair_data)'''
        lib = loader.load_library()
        if getattr(self, '_diffuse_dev', None) is None:
            exps, cs = _diffuse_pgto_params(self.sorted_mol)
            self._diffuse_dev = (torch.as_tensor(exps, device=self.device), torch.as_tensor(cs, device=self.device))
        n = len(ids)
        s_est = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
        geom = torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.device)
        if n and lib.B200_sr_pair_data(_ptr(s_est), _ptr(geom), ctypes.byref(self._rys_envs), _ptr(ids),
                                       _ptr(self._diffuse_dev[0]), _ptr(self._diffuse_dev[1]), n, float(omega),
                                       loader.current_stream_ptr()):
            raise RuntimeError('B200_sr_pair_data failed')
        return s_est, geom

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
        self._lists = {}
        self._default = self._use_lists(self._own_terms() or [(1., 0.)])
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
    def row_list(self, key):
        '''Row view of group pair ``key = (gi, gj)``: the pairs grouped by one shell p (the row
        shell) with the partner q running over the row; q is the shell with the lower angular
        momentum (for equal l: the one from the larger group, i.e. the longer rows).  Pair ids
        are ``p*nbas+q`` in (p, q) = kernel order, records follow that order.'''
        if key in self._rows:
            return self._rows[key]
        lib = loader.load_library()
        smol = self.sorted_mol
        nbas = smol.nbas
        uniq = smol.uniq_l_ctr
        counts = np.asarray(smol.l_ctr_counts)
        gi, gj = key
        ids, q = self._cur.cache[key][:2]
        li, lj = int(uniq[gi, 0]), int(uniq[gj, 0])
        stream_second = li > lj or counts[gj] >= counts[gi]
        gp, gq = (gi, gj) if stream_second else (gj, gi)
        ish = torch.div(ids.to(torch.int64), nbas, rounding_mode='floor')
        jsh = ids.to(torch.int64) - ish * nbas
        p, qs = (ish, jsh) if stream_second else (jsh, ish)
        order = torch.argsort(p * nbas + qs)
        p, qs, qv = p[order], qs[order], q[order].contiguous()
        pair = (p * nbas + qs).to(torch.int32).contiguous()
        if len(pair):
            row_p, cnt = torch.unique_consecutive(p, return_counts=True)
            row_ptr = torch.zeros(len(row_p) + 1, dtype=torch.int32, device=ids.device)
            row_ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
            row_qmax = torch.segment_reduce(qv, 'max', lengths=cnt).contiguous()
            row_p = row_p.to(torch.int32).contiguous()
        else:
            row_p = torch.zeros(0, dtype=torch.int32, device=ids.device)
            row_ptr = torch.zeros(1, dtype=torch.int32, device=ids.device)
            row_qmax = torch.zeros(0, dtype=torch.float32, device=ids.device)
        npp, npq = int(uniq[gp, 1]), int(uniq[gq, 1])
        rec = torch.empty(max(1, len(pair) * lib.B200_pair_rec_stride(npp, npq)), dtype=torch.float64,
                          device=ids.device)
        if len(pair):
            err = lib.B200_pair_records(_ptr(rec), ctypes.byref(self._rys_envs), _ptr(pair), len(pair), npp, npq,
                                        loader.current_stream_ptr())
            if err:
                raise RuntimeError('B200_pair_records failed')
        sr_s = sr_g = None
        if self._cur.sr is not None:
            sr_s, sr_g = self._sr_pair_data(pair, self._cur.sr['omega'])
        out = dict(pair=pair, q=qv, rec=rec, row_ptr=row_ptr, row_p=row_p, row_qmax=row_qmax, sr_s=sr_s, sr_g=sr_g,
                   nrows=len(row_p), mean_len=len(pair) / max(1, len(row_p)),
                   lp=int(uniq[gp, 0]), lq=int(uniq[gq, 0]), nprim_p=npp, nprim_q=npq)
        self._rows[key] = out
        return out

    def class_tasks(self):
        '''Launch list: one entry per unordered pair of group pairs:
        ``('flat', Y key, X key, dedupe)`` -- both lists flat and q-sorted (rys_jk_thread /
        rys_jk_warp kernels), or ``('row', Y key, X key, dedupe)`` -- Y streamed through its row
        view.  Y is the lane side and X the resident side of the kernel.'''
        if self._tasks is not None:
            return self._tasks
        lib = loader.load_library()
        smol = self.sorted_mol
        uniq = smol.uniq_l_ctr
        keys = sorted(self._cur.cache.keys())
        use_rows = os.environ.get('B200_ROWS', '1') == '1'

        def ls(key):
            return int(uniq[key[0], 0]), int(uniq[key[1], 0])

        def row_score(ky, kx):
            '''cost estimate of streaming ky through its row view against resident kx (lower is
            better), or None when no row-stream kernel serves the launch / rows are too short'''
            if not use_rows or len(self._cur.cache[ky][0]) == 0:
                return None
            rl = self.row_list(ky)
            if rl['mean_len'] < self.row_min:
                return None
            lr, ls_ = ls(kx)
            kind = lib.B200_class_kind(rl['lp'], rl['lq'], lr, ls_, 1)
            if kind == 0:
                return None
            force = os.environ.get('B200_ROW_FORCE', '0') == '1'      # measurement: skip the exclusions below
            contracted = rl['nprim_p'] * rl['nprim_q'] != 1
            if kind == 1 and contracted and rl['mean_len'] < 2 * self.row_min and not force:
                return None                 # few, fat quartets per row: partially filled batches cost too much
            if kind == 2:
                # the row-stream warp-team kernel is specialised for uncontracted quartets
                if contracted or int(uniq[kx[0], 1]) * int(uniq[kx[1], 1]) != 1:
                    return None
                # measured on Taxol: it loses against the flat kernels when the streamed shell is a
                # d shell or higher (long unrolled b loops), and against the thread-per-quartet flat
                # kernel where the resident register tile is small (< 18 components)
                if not force and rl['lq'] >= 2:
                    return None
                if not force and _nf(lr) * _nf(ls_) < 18 and 1 in (lib.B200_class_kind(*ls(ky), lr, ls_, 0),
                                                                    lib.B200_class_kind(lr, ls_, *ls(ky), 0)):
                    return None
            scattered = _nf(rl['lq']) * (_nf(rl['lp']) + _nf(lr) + _nf(ls_))
            return scattered * (1. + 16. / rl['mean_len'])

        tasks = []
        for ia, ka in enumerate(keys):
            for kb in keys[:ia + 1]:
                cands = [(ka, kb)] if ka == kb else [(ka, kb), (kb, ka)]
                best = None
                for ky, kx in cands:
                    sc = row_score(ky, kx)
                    if sc is not None and (best is None or sc < best[0]):
                        best = (sc, ky, kx)
                if best is not None:
                    tasks.append(('row', best[1], best[2], 1 if ka == kb else 0))
                elif ka == kb:
                    tasks.append(('flat', ka, ka, 1))
                elif lib.B200_has_class(*ls(kb), *ls(ka)):
                    tasks.append(('flat', kb, ka, 0))      # (Y, X)
                else:
                    tasks.append(('flat', ka, kb, 0))
        self._tasks = tasks
        return tasks

    # ------------------------------------------------------------------
    def _jk_pass(self, dms, hermi, with_j, with_k, terms):
        '''One pass over all quartets with the operator sum_k coef_k O[w_k] (O[0] = 1/r, O[w > 0] =
        erf(w r)/r; one term, or 1/r plus an erf term: range-separated exchange in ONE build, the
        counterpart of the reference's doubled root set, rys_roots_for_k.cu:45-60), screened with the
        Schwarz bounds of that operator: J/K in the sorted Cartesian basis, symmetrised as the reference
        does (jk.py:594-616).  dms: (n_dm, nao, nao) device tensor.  Returns (vj, vk), None where not
        requested.'''
        if isinstance(terms, (int, float)):
            terms = [(1., float(terms))]
        assert 1 <= len(terms) <= 2 and all(w >= 0 for _, w in terms)
        self._use_lists(terms)
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
        if not (with_j or with_k):
            return None, None
        envs = self._rys_envs
        stream = loader.current_stream_ptr()
        tasks = self.class_tasks()
        any_rows = any(tk[0] == 'row' for tk in tasks)
        # transposed views for the row-stream kernels: dmT[k] = dms[k]^T; K blocks written through
        # vkT[k] are added back transposed.  Symmetric D: both alias the originals.
        if hermi == 1 or not any_rows:
            dmT = dms
        elif hermi == 0:
            dmT = torch.cat([dms[n_dm:], dms[:n_dm]], dim=0).contiguous()
        else:
            dmT = -dms
        sep_kT = with_k and any_rows and hermi != 1
        # one contiguous output buffer [vj | vk | vkT]: a single all_reduce, no concatenation
        nj = ndm_k if with_j else 0
        nk = ndm_k if with_k else 0
        nt = ndm_k if sep_kT else 0
        rank, world = self._partition()
        # Multi-GPU partition (B200_PARTITION).  'items': the work items (X pair, Y chunk / row block) of every
        # launch are dealt round-robin over the ranks -- balanced by construction, but a launch that lasts
        # < 1 ms on 8 GPUs pays its fixed costs (launch gap, table staging, tail) on every rank.
        # 'hybrid' (default): launches whose measured GPU time is below B200_SPLIT_MIN_MS run WHOLE on one rank,
        # assigned longest-processing-time-first; the heavy launches stay item-dealt.  'cost': every launch
        # whole on one rank (measured on 8 B200: 0.735 s vs 0.684 s for 'items' -- the measured costs of the
        # heavy launches are not accurate enough to balance them).  The first build of a _VHFOpt measures the
        # launch times (items dealt, launches serialised, CUDA events) and ships them in the tail of the
        # [vj|vk] buffer, so there is still ONE all_reduce per build.
        part = os.environ.get('B200_PARTITION', 'hybrid')
        # (a caller that sets rank/world by hand drives the ranks itself: plain item dealing, no shared state)
        spmd = self.world is None and torch.distributed.is_available() and torch.distributed.is_initialized()
        cost_mode = world > 1 and spmd and part in ('cost', 'hybrid') and self.class_timing is None
        measuring = cost_mode and (self._launch_cost is None or len(self._launch_cost) != len(tasks))
        owners = None
        if cost_mode and not measuring:
            split_min = float(os.environ.get('B200_SPLIT_MIN_MS', '20')) if part == 'hybrid' else float('inf')
            owners = _lpt_owners(self._launch_cost, world, split_min)
        nmat = nj + nk + nt
        flat = torch.zeros(nmat * nao * nao + (len(tasks) if measuring else 0), dtype=torch.float64, device=dev)
        out = flat[:nmat * nao * nao].view(nmat, nao, nao)
        vj = out[:nj] if with_j else None
        vk = out[nj:nj + nk] if with_k else None
        vkT = out[nj + nk:] if sep_kT else vk

        # Class launches are independent (they only meet in the J/K atomics), so
        # they are dealt over a few side streams: the tail of one launch, when
        # most SMs have run out of work items, overlaps the start of the next.
        side = None
        if dev.type == 'cuda' and self.class_timing is None and self.n_streams > 1 and not measuring:
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

        if getattr(self, 'counter', None) is not None:
            self._counter_last = int(self.counter.item())
        heads = torch.zeros(2 * len(tasks) + 1, dtype=torch.int64, device=dev)
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            for st in side:
                st.wait_event(ev)
        uniq = smol.uniq_l_ctr
        lprim = (ctypes.c_int * 8)()
        timers = []
        only = getattr(self, 'only_classes', None)   # measurement: restrict to some classes
        for it, (mode, ky, kx, dedupe) in enumerate(tasks):
            pair_x, q_x = self._cur.cache[kx][:2]
            if len(pair_x) == 0 or len(self._cur.cache[ky][0]) == 0:
                continue
            # cheap class-level early exit
            if float(self._qmax[kx]) + float(self._qmax[ky]) + dm_penalty < cutoff:
                continue
            lprim[2], lprim[3] = int(uniq[kx[0], 0]), int(uniq[kx[1], 0])
            lprim[6], lprim[7] = int(uniq[kx[0], 1]), int(uniq[kx[1], 1])
            if mode == 'row':
                rl = self.row_list(ky)
                lprim[0], lprim[1], lprim[4], lprim[5] = rl['lp'], rl['lq'], rl['nprim_p'], rl['nprim_q']
            else:
                lprim[0], lprim[1] = int(uniq[ky[0], 0]), int(uniq[ky[1], 0])
                lprim[4], lprim[5] = int(uniq[ky[0], 1]), int(uniq[ky[1], 1])
            if only is not None:
                if tuple(lprim[:4]) not in only and tuple(lprim) not in only:
                    continue
            l_rank, l_world = rank, world
            if owners is not None and owners[it] >= 0:
                if owners[it] != rank:
                    continue
                l_rank, l_world = 0, 1
            if measuring:
                if dev.type == 'cuda':
                    tm0 = torch.cuda.Event(enable_timing=True)
                    tm1 = torch.cuda.Event(enable_timing=True)
                    tm0.record()
                else:
                    tm0 = time.perf_counter()
            if self.class_timing is not None and dev.type == 'cuda':
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            if side is not None:
                stream = ctypes.c_void_p(side[it % len(side)].cuda_stream)
            # operator of the launch: both terms at once where the kernel family takes them, else one
            # launch per term into the same outputs
            if len(terms) == 2 and lib.B200_class_two_terms(lprim[0], lprim[1], lprim[2], lprim[3], int(mode == 'row')):
                ops = [(terms[0][1], terms[0][0], terms[1][0] / terms[0][0], terms[1][1])]
            else:
                ops = [(w, c, 0., 0.) for c, w in terms]
            # short-range exchange filter: K-only builds of the pure short-range operator
            srd = self._cur.sr if (not with_j and os.environ.get('B200_SR_FILTER', '1') == '1') else None
            sr_x = srd['flat'][kx] if srd else (None, None)
            sr_y = ((rl['sr_s'], rl['sr_g']) if mode == 'row' else srd['flat'][ky]) if srd else (None, None)
            sr_w = float(srd['omega']) if srd else 0.
            for iop, (op_w, op_scale, op_ratio2, op_w2) in enumerate(ops):
                head = ctypes.c_void_p(heads.data_ptr() + 8 * (2 * it + iop))
                if mode == 'row':
                    err = lib.B200_build_jk_rows(
                        _ptr(vj), _ptr(vk), _ptr(vkT), _ptr(dms), _ptr(dmT), ndm_k, nao, ctypes.byref(envs), lprim,
                        _ptr(pair_x), _ptr(q_x), len(pair_x),
                        _ptr(rl['pair']), _ptr(rl['q']), _ptr(rl['rec']),
                        _ptr(rl['row_ptr']), _ptr(rl['row_p']), _ptr(rl['row_qmax']), rl['nrows'],
                        _ptr(dm_cond), cutoff, dm_penalty, dedupe, l_rank, l_world, head, stream,
                        float(op_w), float(op_scale), float(op_ratio2), float(op_w2),
                        _ptr(sr_x[0]), _ptr(sr_x[1]), _ptr(sr_y[0]), _ptr(sr_y[1]), sr_w)
                else:
                    pair_y, q_y = self._cur.cache[ky][:2]
                    rec_y = self._cur.cache[ky][3]
                    err = lib.B200_build_jk_class(
                        _ptr(vj), _ptr(vk), _ptr(dms), ndm_k, nao, ctypes.byref(envs), lprim,
                        _ptr(pair_x), _ptr(q_x), len(pair_x), _ptr(pair_y), _ptr(q_y), len(pair_y),
                        _ptr(dm_cond), cutoff, dm_penalty, float(self.q_bin), dedupe, l_rank, l_world, head, stream,
                        _ptr(rec_y), float(op_w), float(op_scale), float(op_ratio2), float(op_w2),
                        _ptr(sr_x[0]), _ptr(sr_x[1]), _ptr(sr_y[0]), _ptr(sr_y[1]), sr_w)
                if err:
                    break
            if err:
                raise RuntimeError('J/K launch failed for class %s (%s)' % (tuple(lprim), mode))
            if measuring:
                if dev.type == 'cuda':
                    tm1.record()
                    timers.append((it, tm0, tm1))
                else:
                    timers.append((it, tm0, time.perf_counter()))
            if self.class_timing is not None and dev.type == 'cuda':
                ev1.record()
                ev1.synchronize()
                key = tuple(lprim) + (mode,)
                rec = self.class_timing.setdefault(key, [0., 0, 0, 0])
                rec[0] += ev0.elapsed_time(ev1)
                rec[1] += len(pair_x)
                rec[2] += len(self._cur.cache[ky][0])
                if getattr(self, 'counter', None) is not None:
                    n = int(self.counter.item())
                    rec[3] += n - self._counter_last
                    self._counter_last = n

        if side is not None:
            for st in side:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)

        if measuring:
            cost = np.zeros(len(tasks))
            if dev.type == 'cuda':
                torch.cuda.current_stream().synchronize()
            for it, a, b in timers:
                cost[it] = a.elapsed_time(b) if dev.type == 'cuda' else (b - a) * 1e3
            flat[nmat * nao * nao:] = torch.as_tensor(cost, device=dev)
        if world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(flat)      # one collective per build over [vj|vk|vkT] (+ launch times)
        if measuring:
            # summed over ranks = GPU time of the whole launch; identical on every rank after the reduce
            self._launch_cost = flat[nmat * nao * nao:].cpu().numpy().copy()

        # symmetrisation of the partial sums (reference: jk.py:594-616)
        if with_k:
            if sep_kT:
                vk = vk + vkT.transpose(1, 2)
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

    def _combine(self, dms, hermi, with_j, with_k, terms):
        '''J/K of the operator sum_k coef_k O[w_k]: ONE pass when the sum is positive semi-definite (its own
        Schwarz bounds are valid for it), else one pass per term'''
        if len(terms) == 0:
            return None, None
        if len(terms) == 1 or _psd_pair(terms):
            return self._jk_pass(dms, hermi, with_j, with_k, list(terms))
        vj = vk = None
        for term in terms:
            tj, tk = self._jk_pass(dms, hermi, with_j, with_k, [term])
            if tj is not None:
                vj = tj if vj is None else vj + tj
            if tk is not None:
                vk = tk if vk is None else vk + tk
        return vj, vk

    def _own_terms(self):
        '''operator of the molecule itself (mol.omega, libcint sign convention)'''
        return _operator_terms(*_check_rsh_factors(self.mol, None, None, None))

    def get_jk(self, dms, hermi=0, verbose=None, with_j=True, with_k=True):
        '''J and K in the sorted Cartesian basis with the molecule's own Coulomb operator
        (reference: _VHFOpt.get_jk, jk.py:492-641, which reads omega from the sorted mol's env).'''
        dms = dms.reshape(-1, self.sorted_mol.nao, self.sorted_mol.nao)
        vj, vk = self._combine(dms, hermi, with_j, with_k, self._own_terms())
        if with_j and vj is None:
            vj = torch.zeros_like(dms)
        if with_k and vk is None:
            vk = torch.zeros_like(dms)
        return vj, vk

    def get_j(self, dms, verbose=None):
        '''J of symmetric density matrices (reference: jk.py:643-783)'''
        return self.get_jk(dms, 1, verbose, with_j=True, with_k=False)[0]

    def get_k(self, dms, hermi=0, verbose=None, omega=None, lr_factor=None, sr_factor=None):
        '''K with the operator lr_factor*erf(|omega| r)/r + sr_factor*erfc(|omega| r)/r (reference:
        jk.py:785-927; omega = None: the molecule's own operator)'''
        dms = dms.reshape(-1, self.sorted_mol.nao, self.sorted_mol.nao)
        terms = _operator_terms(*_check_rsh_factors(self.mol, omega, lr_factor, sr_factor))
        vk = self._combine(dms, hermi, False, True, terms)[1]
        return vk if vk is not None else torch.zeros_like(dms)

    @property
    def _qmax(self):
        if getattr(self, '_qmax_cache', None) is None:
            self._qmax_cache = {k: (float(v[1].max()) if len(v[1]) else -1e30)
                                for k, v in self._cur.cache.items()}
        return self._qmax_cache


def _diffuse_pgto_params(smol):
    '''Exponent and |coefficient| of the most diffuse primitive of every (sorted) shell, picked
    like the reference's extract_pgto_params(mol, 'diffuse') (gto/mole.py:348-382).'''
    exps = np.empty(smol.nbas, dtype=np.float32)
    cs = np.empty(smol.nbas, dtype=np.float32)
    precision = 1e-8
    for ib, b in enumerate(smol._bas):
        l, nprim, nctr = int(b[ANG_OF]), int(b[NPRIM_OF]), int(b[3])
        e = smol._env[b[5]:b[5] + nprim]
        c = np.abs(smol._env[b[6]:b[6] + nprim * nctr].reshape(nctr, nprim)).max(axis=0)
        tmp = (2 * l + 1) / (4 * np.pi) / precision**2 * c**2 * 1e2**l
        tmp[tmp <= 0] = 1e-200
        k = int(np.argmax(np.log(tmp) / (2 * e)))
        exps[ib], cs[ib] = e[k], c[k]
    return exps, cs


def _cache_q_cond_and_non0pairs(smol, rys_envs, precision, device, lib, q_bin=0., operator=(0., 1., 1.)):
    '''Per group pair (gi >= gj): shell-pair ids ``ish*nbas+jsh`` (ish >= jsh), their Schwarz
    bounds q = 1/2 log max|(ij|ij)|, the same again as ``s_cond`` and the Gaussian-product
    records -- same role and call sequence as the reference (scf/jk.py:984-1120):

    1. overlap prescreen + enumeration of all ish >= jsh pairs in one kernel
       (``B200_pair_prescreen``: count pass, prefix sum, fill pass),
    2. ONE ``int2e_qcond_estimator`` launch sequence over the concatenated list (reference signature) with
       ``operator`` = (omega, lr_factor, sr_factor) of the operator the lists will screen,
    3. one global sort by (group pair, q bin descending, pair id) -- the reference bins by 4 log
       units (Q_COND_MARGIN); here the bins are ``q_bin`` wide (0: exactly q-sorted), and pairs
       that cannot survive against the largest bound of the molecule are dropped,
    4. one ``B200_pair_records`` launch per group pair.'''
    nbas = smol.nbas
    uniq = smol.uniq_l_ctr
    counts = np.asarray(smol.l_ctr_counts)
    ngroups = len(counts)
    ngp = ngroups * (ngroups + 1) // 2
    log_cutoff = math.log(precision)
    stream = loader.current_stream_ptr()
    i32 = dict(dtype=torch.int32, device=device)

    exps, cs = _diffuse_pgto_params(smol)
    exps_d = torch.as_tensor(exps, device=device)
    logc_d = torch.as_tensor(np.log(np.maximum(cs, 1e-30)).astype(np.float32), device=device)
    grp_d = torch.as_tensor(np.repeat(np.arange(ngroups), counts).astype(np.int32), device=device)
    cnt_d = torch.zeros(ngp, **i32)
    args = (ctypes.byref(rys_envs), _ptr(exps_d), _ptr(logc_d), _ptr(grp_d), 2. * log_cutoff, stream)
    if lib.B200_pair_prescreen(None, _ptr(cnt_d), None, *args):
        raise RuntimeError('B200_pair_prescreen failed')
    cnt = cnt_d.cpu().numpy().astype(np.int64)
    offs = np.append(0, np.cumsum(cnt))
    ntot = int(offs[-1])
    keys = [(gi, gj) for gi in range(ngroups) for gj in range(gi + 1)]
    if ntot == 0:
        empty = torch.zeros(0, **i32)
        emptyq = torch.zeros(0, dtype=torch.float32, device=device)
        return {k: (empty, emptyq, emptyq, torch.zeros(0, dtype=torch.float64, device=device)) for k in keys}
    offs_d = torch.as_tensor(offs[:-1].astype(np.int32), device=device)
    ids = torch.empty(ntot, **i32)
    cnt_d.zero_()
    if lib.B200_pair_prescreen(_ptr(ids), _ptr(cnt_d), _ptr(offs_d), *args):
        raise RuntimeError('B200_pair_prescreen failed')

    # batches of one launch block each; no batch crosses a group-pair boundary
    batch = 256
    spo = np.hstack([np.arange(offs[g], offs[g + 1], batch) for g in range(ngp)] + [[ntot]]).astype(np.int32)
    spo_d = torch.as_tensor(spo, device=device)
    q = torch.empty(ntot, dtype=torch.float32, device=device)
    # the estimator runs on the legacy default stream, like the reference's
    if device.type == 'cuda':
        torch.cuda.current_stream().synchronize()
    if lib.int2e_qcond_estimator(_ptr(q), ctypes.byref(rys_envs), 0, len(spo) - 1, _ptr(ids), _ptr(spo_d),
                                 None, float(operator[0]), float(operator[1]), float(operator[2])):
        raise RuntimeError('int2e_qcond_estimator kernel failed')
    if device.type == 'cuda':
        torch.cuda.synchronize()

    gp = torch.repeat_interleave(torch.arange(ngp, device=device), torch.as_tensor(cnt, device=device))
    keep = q + q.max() > log_cutoff
    ids, q, gp = ids[keep], q[keep], gp[keep]
    ids64 = ids.to(torch.int64)
    if q_bin > 0:
        # descending q bins of width q_bin, shell-pair order (ish, jsh) inside a bin: neighbours in the
        # list share the bra shell, so the D rows they load and the K rows they update coincide
        # (the kernels get q_margin = q_bin for the early exit)
        bins = torch.floor(q.double() / q_bin).to(torch.int64)
        bins = bins.max() - bins
        order = torch.argsort((gp * (int(bins.max()) + 1) + bins) * (int(nbas) * int(nbas)) + ids64)
    else:
        order = torch.argsort(q, descending=True, stable=True)
        order = order[torch.argsort(gp[order], stable=True)]
    ids = ids[order].contiguous()
    q = q[order].contiguous()
    kept = torch.bincount(gp, minlength=ngp).cpu().numpy()
    koffs = np.append(0, np.cumsum(kept))
    cache = {}
    for g, key in enumerate(keys):
        a, b = int(koffs[g]), int(koffs[g + 1])
        ids_g, q_g = ids[a:b], q[a:b]
        # Gaussian-product records of the pairs, in list order
        npi, npj = int(uniq[key[0], 1]), int(uniq[key[1], 1])
        rec = torch.empty((b - a) * lib.B200_pair_rec_stride(npi, npj), dtype=torch.float64, device=device)
        if b > a and lib.B200_pair_records(_ptr(rec), ctypes.byref(rys_envs), _ptr(ids_g), b - a, npi, npj, stream):
            raise RuntimeError('B200_pair_records failed')
        cache[key] = (ids_g, q_g, q_g, rec)
    return cache
