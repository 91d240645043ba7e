#!/usr/bin/env python
"""Generate the per-class instantiation units gpu4pyscf_b200/csrc/gen/jk_inst_*.cu
and the dispatch table.  A class is (Y pair type | X pair type) with
nf(X) >= nf(Y): the X pair is the register tile, see rys_jk_kernel.cuh."""
import os
import sys

LMAX = int(os.environ.get('B200_LMAX', '4'))
NUNITS = int(os.environ.get('B200_NUNITS', '16'))
# classes with nf + nf_rs*n_dm below this use the thread-per-quartet register kernel
THREAD_KERNEL_MAX = int(os.environ.get('B200_THREAD_KERNEL_MAX', '120'))


def nf(l):
    return (l + 1) * (l + 2) // 2


def ncs_for(lr, ls, ndm, g_y):
    nfr, nfs = nf(lr), nf(ls)
    for ncs in range(1, nfr + 1):
        nc = (nfr + ncs - 1) // ncs
        if nc * nfs * (1 + ndm) <= 110:
            break
    while g_y * ncs > 1024:
        ncs -= 1
    return ncs


def warp_cfg(y, x, ndm):
    """(score, ncs) of the warp-team kernel for lanes = Y pair, register tile = X pair."""
    nfpq = nf(y[0]) * nf(y[1])
    nfr, nfs = nf(x[0]), nf(x[1])
    sub = (x[0] + 1) * (x[1] + 1)
    best = None
    for ncs in range(1, nfr + 1):
        nc = -(-nfr // ncs)
        if -(-nfr // nc) != ncs:
            continue
        tile = (1 + ndm) * nc * nfs
        g = nfpq * ncs
        if g > 256:
            break
        tw = -(-g // 32)
        nsw = 32 // g if tw == 1 else 1
        eff = nsw * nfpq * nfr / (tw * 32 * nc)
        pen = 1.0 if tile <= 40 else (0.85 if tile <= 64 else (0.5 if tile <= 90 else 0.2))
        lds = 3 * sub if ncs == 1 else 3 * (x[1] + 1) * nc
        dp = 2 * nc * nfs
        sc = eff * pen * min(1., dp / (1.5 * lds)) * (1. if tw == 1 else 0.9)
        if best is None or sc > best[0] + 1e-9:
            best = (sc, ncs)
    return best


WARP_KERNEL = os.environ.get('B200_WARP_KERNEL', '1') == '1'
CHOOSER = os.environ.get('B200_CHOOSER', 'cost')
REG_DOUBLES = int(os.environ.get('B200_REG_DOUBLES', '180'))


BL_ENABLE = os.environ.get('B200_BL', '1') == '1'
BL_GAIN = float(os.environ.get('B200_BL_GAIN', '0.85'))   # use the b-loop variant when its modelled cost is below this fraction


# register budget (doubles) of the b-loop variant in ROW kernels.  Measured on Taxol (round 2):
# raising it so that (21|32) / (11|32) rows take the b-loop variant (5 quartets per warp instead of
# 1 or 3) LOSES -- 147 vs 143 ms and 60 vs 43 ms: one g-set per lane slot and 7 instead of 8 warps/SM
BL_ROW_REG = int(os.environ.get('B200_BL_ROW_REG', str(REG_DOUBLES)))


def warp_cfg_cost(y, x, ndm, bl=False, row=False):
    """Instruction-count model: estimated warp instructions per quartet for lanes =
    Y pair, register tile = X pair (split NCS ways); returns (-cost, ncs) so that
    larger is better, like warp_cfg."""
    import math
    lp, lq = y
    lr, ls = x
    nfpq = nf(lp) * nf(lq)
    nfr, nfs = nf(lr), nf(ls)
    lpq, lrs = lp + lq, lr + ls
    nroots = (lpq + lrs) // 2 + 1
    sub = (lr + 1) * (ls + 1)
    best = None
    for ncs in range(1, nfr + 1):
        nc = -(-nfr // ncs)
        if -(-nfr // nc) != ncs:
            continue
        tile = nc * nfs
        g = (nf(lp) if bl else nfpq) * ncs
        if g > 256:
            break
        tw = -(-g // 32)
        nsw = 32 // g if tw == 1 else 1
        seg = lrs + 1 if ncs == 1 else sub
        regs = tile * (1 + ndm) + (3 * sub + 3 * seg if ncs == 1 else 3 * (ls + 1)) + (bl and (nc + nfs) * ndm)
        if regs > (BL_ROW_REG if bl and row else REG_DOUBLES):
            continue
        ld = 3 * ((seg + 1) // 2) if ncs == 1 else 3 * (ls + 1) * nc
        hrr = 3 * sum(lrs - d for d in range(ls)) if ncs == 1 else 0
        dig = 6 * tile * ndm + 3 * (2 * nc + 2 * nfs + 1) * ndm
        w = nroots * (ld + hrr + 2 * tile) + dig + 20
        if bl:
            w *= nf(lq)          # the lane loops over the nf_q components itself
        k = max(1, min(4, (32 * tw) // (nsw * 3 * nroots)))
        nab2 = (lp + 1) * (lq + 1)
        p2 = 60 + 3 * (lpq + 1) * (lrs + 1) + nab2 * ((seg + 1) // 2 + 2)
        f = (150 + 100 * math.ceil(2 * nroots * nsw * k / (32 * tw))
             + p2 * math.ceil(3 * nroots * nsw * k / (32 * tw)) + 100)
        cost = tw * (f / k + w) / nsw * (1.6 if tw > 1 else 1.)   # named-barrier teams measured slower than modelled
        if best is None or -cost > best[0]:
            best = (-cost, ncs)
    return best



_warp_cfg_score = warp_cfg


def warp_cfg(y, x, ndm):
    if CHOOSER == 'cost':
        return warp_cfg_cost(y, x, ndm)
    return _warp_cfg_score(y, x, ndm)


def classes():
    """One (Y pair type | X pair type) orientation per unordered class.  Thread-kernel
    classes keep nf(X) >= nf(Y); for the warp-team kernel the orientation with
    the better lane/register-tile score wins."""
    types = [(a, b) for a in range(LMAX + 1) for b in range(a + 1)]
    out = []
    for y in types:
        for x in types:
            ny, nx = nf(y[0]) * nf(y[1]), nf(x[0]) * nf(x[1])
            if nx > ny or (nx == ny and x >= y):
                yy, xx = y, x
                if WARP_KERNEL and x != y and ny * nx + nx > THREAD_KERNEL_MAX:
                    a, b = warp_cfg(y, x, 1), warp_cfg(x, y, 1)
                    if b is not None and (a is None or b[0] > a[0] + 1e-9):
                        yy, xx = x, y
                out.append((yy[0], yy[1], xx[0], xx[1]))
    return out


def cost(c):
    lp, lq, lr, ls = c
    return nf(lp) * nf(lq) * nf(lr) * nf(ls)


# ---- row-stream kernels: Y = (p,q) rows of a uniform shell p, q streamed over lanes ----
ROWT_MAX_NF = int(os.environ.get('B200_ROWT_MAX_NF', '54'))        # integrals per lane
ROWT_MAX_LIVE = int(os.environ.get('B200_ROWT_MAX_LIVE', '100'))   # + row/item accumulators (doubles)


def rowt_ok(lp, lq, lr, ls, ndm):
    n = nf(lp) * nf(lq) * nf(lr) * nf(ls)
    acc = ndm * (nf(lp) * (nf(lr) + nf(ls)) + nf(lr) * nf(ls))
    return n <= ROWT_MAX_NF and n + acc <= ROWT_MAX_LIVE


def row_classes():
    """(lp, lq, lr, ls, kind) with lp >= lq (q = streamed shell), lr >= ls.  kind 't': thread per
    quartet (every orientation that fits is instantiated, the host picks per launch); kind 'w':
    warp-team kernel, instantiated with the lowest angular momentum of the class as the streamed
    shell (for equal lq, ls: the orientation the lane/tile model prefers)."""
    types = [(a, b) for a in range(LMAX + 1) for b in range(a + 1)]
    out = []
    for y in types:
        for x in types:
            if rowt_ok(y[0], y[1], x[0], x[1], 1):
                out.append((y[0], y[1], x[0], x[1], 't'))
                continue
            if not WARP_KERNEL or y[1] > x[1]:
                continue
            a = warp_cfg(y, x, 1)
            if a is None:
                continue
            if y[1] == x[1] and y != x and not rowt_ok(x[0], x[1], y[0], y[1], 1):
                b = warp_cfg(x, y, 1)
                if b is not None and (b[0] > a[0] + 1e-9 or (abs(b[0] - a[0]) <= 1e-9 and x > y)):
                    continue            # the other orientation is the instantiated one
            out.append((y[0], y[1], x[0], x[1], 'w'))
    return out


def warp_variants(lp, lq, lr, ls, ndm):
    """(ncs of the general kernel, ncs of the b-loop variant or None) of a ROW class"""
    wc = warp_cfg((lp, lq), (lr, ls), ndm)
    if wc is None:
        return None
    bl = None
    if BL_ENABLE and CHOOSER == 'cost' and lq > 0:
        bl = warp_cfg_cost((lp, lq), (lr, ls), ndm, bl=True, row=True)
        # measured on Taxol: pays off for unsplit register tiles only
        if bl is not None and (-bl[0] > BL_GAIN * -wc[0] or bl[1] != 1):
            bl = None
    return wc[1], (bl[1] if bl is not None else None)


def main(dst):
    os.makedirs(dst, exist_ok=True)
    cls = sorted(classes(), key=cost, reverse=True)
    units = [[] for _ in range(NUNITS)]
    loads = [0] * NUNITS
    for c in cls:
        i = loads.index(min(loads))
        units[i].append(c)
        loads[i] += cost(c) + 2000
    decls = []
    for i, u in enumerate(units):
        with open(os.path.join(dst, 'jk_inst_%02d.cu' % i), 'w') as f:
            f.write('// GENERATED by tools/gen_instances.py\n#include "jk_launch.cuh"\n')
            for (lp, lq, lr, ls) in u:
                for ndm in (1, 2):
                    ncs = ncs_for(lr, ls, ndm, nf(lp) * nf(lq))
                    name = 'jk_launch_%d%d%d%d_%d' % (lp, lq, lr, ls, ndm)
                    nfq = nf(lp) * nf(lq) * nf(lr) * nf(ls)
                    if nfq + nf(lr) * nf(ls) * ndm <= THREAD_KERNEL_MAX:
                        f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                                '{ return jk_launch_thread<%d, %d, %d, %d, %d>(t, nsm, smem_max, s); }\n'
                                % (name, lp, lq, lr, ls, ndm))
                        decls.append((lp, lq, lr, ls, ndm, name, 1))
                        continue
                    wc = warp_cfg((lp, lq), (lr, ls), ndm) if WARP_KERNEL else None
                    if wc is not None and (CHOOSER == 'cost' or wc[0] >= 0.3):
                        # uncontracted launches may use the b-loop variant (a lane owns a p
                        # component and loops over the q components): fewer lanes per quartet
                        bl = None
                        if BL_ENABLE and CHOOSER == 'cost' and lq > 0:
                            bl = warp_cfg_cost((lp, lq), (lr, ls), ndm, bl=True)
                            # measured on Taxol: pays off for unsplit register tiles only
                            if bl is not None and (-bl[0] > BL_GAIN * -wc[0] or bl[1] != 1):
                                bl = None
                        if bl is not None:
                            f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                                    '{\n    if (t->nprim_p * t->nprim_q * t->nprim_r * t->nprim_s == 1 && t->ratio2 == 0.)\n'
                                    '        return jk_launch_warp<%d, %d, %d, %d, %d, %d, 1>(t, nsm, smem_max, s);\n'
                                    '    return jk_launch_warp<%d, %d, %d, %d, %d, %d>(t, nsm, smem_max, s);\n}\n'
                                    % (name, lp, lq, lr, ls, bl[1], ndm, lp, lq, lr, ls, wc[1], ndm))
                        else:
                            f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                                    '{ return jk_launch_warp<%d, %d, %d, %d, %d, %d>(t, nsm, smem_max, s); }\n'
                                    % (name, lp, lq, lr, ls, wc[1], ndm))
                        decls.append((lp, lq, lr, ls, ndm, name, 2))
                        continue
                    f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                            '{ return jk_launch<%d, %d, %d, %d, %d, %d>(t, nsm, smem_max, s); }\n'
                            % (name, lp, lq, lr, ls, ncs, ndm))
                    decls.append((lp, lq, lr, ls, ndm, name, 3))
    # row-stream instantiations, dealt over the same units
    rdecls = []
    rcls = sorted(row_classes(), key=lambda c: cost(c[:4]), reverse=True)
    for n, (lp, lq, lr, ls, kind) in enumerate(rcls):
        with open(os.path.join(dst, 'jk_inst_%02d.cu' % (n % NUNITS)), 'a') as f:
            for ndm in (1, 2):
                name = 'jk_launch_row_%d%d%d%d_%d' % (lp, lq, lr, ls, ndm)
                if kind == 't' and rowt_ok(lp, lq, lr, ls, ndm):
                    f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                            '{ return jk_launch_rowt<%d, %d, %d, %d, %d>(t, nsm, smem_max, s); }\n'
                            % (name, lp, lq, lr, ls, ndm))
                    rdecls.append((lp, lq, lr, ls, ndm, name, 1))
                    continue
                if kind == 't':
                    continue        # thread class whose ndm = 2 form is too fat: two ndm = 1 launches
                wv = warp_variants(lp, lq, lr, ls, ndm) if WARP_KERNEL else None
                if wv is None:
                    continue
                if wv[1] is not None:
                    f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                            '{\n    if (t->nprim_p * t->nprim_q * t->nprim_r * t->nprim_s == 1 && t->ratio2 == 0.)\n'
                            '        return jk_launch_warp<%d, %d, %d, %d, %d, %d, 1, 1>(t, nsm, smem_max, s);\n'
                            '    return jk_launch_warp<%d, %d, %d, %d, %d, %d, 0, 1>(t, nsm, smem_max, s);\n}\n'
                            % (name, lp, lq, lr, ls, wv[1], ndm, lp, lq, lr, ls, wv[0], ndm))
                else:
                    f.write('int %s(JKTask *t, int nsm, int smem_max, cudaStream_t s)\n'
                            '{ return jk_launch_warp<%d, %d, %d, %d, %d, %d, 0, 1>(t, nsm, smem_max, s); }\n'
                            % (name, lp, lq, lr, ls, wv[0], ndm))
                rdecls.append((lp, lq, lr, ls, ndm, name, 2))
    with open(os.path.join(dst, 'jk_dispatch.inc'), 'w') as f:
        f.write('// GENERATED by tools/gen_instances.py\n')
        for d in rdecls:
            f.write('int %s(JKTask *, int, int, cudaStream_t);\n' % d[5])
        f.write('static jk_launcher_t jk_find_row_launcher(int lp, int lq, int lr, int ls, int ndm)\n{\n'
                '    switch ((((lp * 8 + lq) * 8 + lr) * 8 + ls) * 4 + ndm) {\n')
        for (lp, lq, lr, ls, ndm, name, kind) in rdecls:
            f.write('    case %d: return %s;\n' % ((((lp * 8 + lq) * 8 + lr) * 8 + ls) * 4 + ndm, name))
        f.write('    }\n    return nullptr;\n}\n')
        # kernel family of an instantiation: 1 thread per quartet, 2 warp team, 3 block slots
        f.write('static int jk_class_kind(int lp, int lq, int lr, int ls, int row)\n{\n'
                '    switch (((((lp * 8 + lq) * 8 + lr) * 8 + ls) * 2) + row) {\n')
        for row, dl in ((0, decls), (1, rdecls)):
            for d in dl:
                if d[4] == 1:
                    f.write('    case %d: return %d;\n' % ((((d[0] * 8 + d[1]) * 8 + d[2]) * 8 + d[3]) * 2 + row, d[6]))
        f.write('    }\n    return 0;\n}\n')
        for d in decls:
            f.write('int %s(JKTask *, int, int, cudaStream_t);\n' % d[5])
        f.write('static jk_launcher_t jk_find_launcher(int lp, int lq, int lr, int ls, int ndm)\n{\n'
                '    switch ((((lp * 8 + lq) * 8 + lr) * 8 + ls) * 4 + ndm) {\n')
        for (lp, lq, lr, ls, ndm, name, kind) in decls:
            f.write('    case %d: return %s;\n' % ((((lp * 8 + lq) * 8 + lr) * 8 + ls) * 4 + ndm, name))
        f.write('    }\n    return nullptr;\n}\n')
    print(len(cls), 'classes in', NUNITS, 'units')


if __name__ == '__main__':
    main(sys.argv[1])
