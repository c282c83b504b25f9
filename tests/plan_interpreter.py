"""Test infrastructure: a numpy interpreter for the pass plan that `k_pass` executes.

`psum_plan_export` (include/psum.h) hands out, per pass, the very arrays the kernel reads -- group
records, pattern z-masks, inline slots, term lists, chunk records, the free/non-free bit maps.
`run_pass` below walks them the way the kernel does (csrc/psum_kernels.cuh: tile addressing, the
fold of out-of-tile sign bits, bundles with one partner set per bundle and a register permutation
per member, fast members with inline slots, general members with class entry lists, remote
bundles, Hermitian pairing), so planner output can be checked against
the oracle on the CPU.  It is a checker for the PLANNER's data, not a product path: nothing in
qiskit_aqua_b200/ imports it.
"""
import ctypes as C

import numpy as np

from qiskit_aqua_b200._lib import lib

GROUP_DT = np.dtype([('xt', '<u4'), ('meta', '<u4'), ('zt0', '<u4'), ('zt1', '<u4'), ('cf0', '<f8'),
                     ('cf1', '<f8'), ('xl', '<u8'), ('ent', '<u2', (4,))])
BUNDLE_DT = np.dtype([('xo', '<u4'), ('meta', '<u4'), ('xlo', '<u8')])
CHUNK_DT = np.dtype([('g0', '<i4'), ('ng', '<i4'), ('p0', '<i4'), ('np', '<i4'), ('b0', '<i4'), ('nb', '<i4'),
                     ('ns', '<u2'), ('nq', '<u2'), ('q0', '<u4'), ('s_cnt', 'u1', (32,))])
assert GROUP_DT.itemsize == 48 and CHUNK_DT.itemsize == 64 and BUNDLE_DT.itemsize == 16
VAR_GENERAL = 63
VAR_QUAD = 62


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def export_pass(engine, g, p):
    """Arrays of pass p of global part g, or None when there is no such pass."""
    counts = (C.c_int64 * 12)()
    rc = lib().psum_plan_export(engine._h, g, p, counts, None, None, None, None, None, None, None, None, None, None, None)
    if rc != 0:
        return None
    ng, npat, nterm, nch, L, nnf, has_im, hi_bit, nbun, rb, nquad, big = [int(v) for v in counts]
    out = dict(L=L, n_nonfree=nnf, has_im=bool(has_im), hi_bit=hi_bit, rb=rb, big=bool(big),
               quad_idx=np.zeros(nquad, np.uint16),
               groups=np.zeros(ng, GROUP_DT), bundles=np.zeros(nbun, BUNDLE_DT),
               pat_zt=np.zeros(npat, np.uint16), pat_inl=np.zeros(npat, np.uint16),
               pat_tptr=np.zeros(npat + 1, np.uint32), term_zb=np.zeros(nterm, np.uint64),
               term_val=np.zeros(nterm, np.float64), chunks=np.zeros(nch, CHUNK_DT),
               fbit=np.zeros(16, np.uint8), nbit=np.zeros(64, np.uint8))
    rc = lib().psum_plan_export(engine._h, g, p, counts, _ptr(out['groups']), _ptr(out['pat_zt']), _ptr(out['pat_inl']),
                                _ptr(out['pat_tptr']), _ptr(out['term_zb']), _ptr(out['term_val']), _ptr(out['chunks']),
                                _ptr(out['fbit']), _ptr(out['nbit']), _ptr(out['bundles']), _ptr(out['quad_idx']))
    assert rc == 0
    return out


def export_part(engine, g):
    passes = []
    while True:
        ps = export_pass(engine, g, len(passes))
        if ps is None:
            return passes
        passes.append(ps)


def _parity(a):
    return (np.bitwise_count(a) & 1).astype(np.int64)


def _sign(a):
    return 1.0 - 2.0 * _parity(a)


def _deposit(values, positions):
    """Bit j of every value goes to bit positions[j]."""
    values = np.asarray(values, dtype=np.uint64)
    out = np.zeros_like(values)
    for j, pos in enumerate(positions):
        out |= ((values >> np.uint64(j)) & np.uint64(1)) << np.uint64(pos)
    return out


def run_pass(ps, src, y=None, herm=False, check=True, tiles=None):
    """Emulates one k_pass launch over all tiles (or the tile-id range `tiles` = (begin, end)).  Adds the pass's contribution to y (if given)
    and returns sum conj(src_i) * contribution_i (with `herm`: the paired 2 Re z evaluation of the
    Hermitian expectation kernels, imaginary part forced to zero as the kernel does)."""
    L, nnf, rb = ps['L'], ps['n_nonfree'], ps['rb']
    NA = 1 << rb
    regmask = np.uint64((NA - 1) << 5)
    t = np.arange(1 << L, dtype=np.uint64)
    t0 = t & ~regmask                              # the kernel's popc never sees the register bits
    reg = (t >> np.uint64(5)) & np.uint64(NA - 1)
    off = _deposit(t, ps['fbit'][:L])
    regq = 0                                       # register qubits as a mask over local qubits
    for j in range(5, 5 + rb):
        regq |= 1 << int(ps['fbit'][j])
    groups, chunks, bundles = ps['groups'], ps['chunks'], ps['bundles']
    total = 0.0 + 0.0j
    for tile_id in (range(1 << nnf) if tiles is None else range(*tiles)):
        base = _deposit(np.array([tile_id], dtype=np.uint64), ps['nbit'][:nnf])[0]
        idx = base | off
        idx0 = idx & ~np.uint64(regq)              # i0 | lane/warp bits, register qubits clear
        tile = src[idx]
        acc = np.zeros(1 << L, dtype=np.complex128)

        for ch in chunks:
            ng = int(ch['ng'])
            p0, npat = int(ch['p0']), int(ch['np'])
            grp = groups[int(ch['g0']):int(ch['g0']) + ng].copy()
            cf = np.zeros(npat)
            zt = np.zeros(npat, dtype=np.uint64)
            for p in range(npat):                                             # the fold
                gp = p0 + p
                ta, tb = int(ps['pat_tptr'][gp]), int(ps['pat_tptr'][gp + 1])
                v = float(np.sum(ps['term_val'][ta:tb] * _sign(ps['term_zb'][ta:tb] & base)))
                zraw = int(ps['pat_zt'][gp])
                if herm and zraw & 0x8000:
                    v += v
                cf[p], zt[p] = v, zraw & 0x7fff
                inl = int(ps['pat_inl'][gp])
                if inl != 0xffff:
                    grp[inl >> 1]['cf1' if inl & 1 else 'cf0'] = v
            def member(G, g, xo, remote, hbit, xlo):
                """Contribution of group record G whose bundle has (xo, remote, hbit, xlo)."""
                meta = int(G['meta'])
                xt = int(G['xt']) & 0x7fffffff
                xr = (xt >> 5) & (NA - 1)
                var = (meta >> 22) & 63
                k = reg ^ np.uint64(xr)                                   # pv[r ^ xr]
                if remote:
                    roff = _deposit(k, ps['fbit'][5:5 + rb])
                    pv = src[(idx0 ^ np.uint64(xlo)) | roff]
                else:
                    pv = tile[(t0 ^ np.uint64(xo)) | (k << np.uint64(5))]
                if check:
                    assert int(G['xt']) >> 31 == remote and (meta >> 28) == hbit
                    assert int(G['xl']) & ~regq == xlo
                    assert np.array_equal(pv, src[idx ^ np.uint64(G['xl'])])
                    if not remote:
                        assert xt & ~int(regmask) == xo
                        assert np.array_equal(idx[(t ^ np.uint64(xt)).astype(np.int64)], idx ^ np.uint64(G['xl']))
                if var == VAR_QUAD:
                    contrib = quad_member(G, meta) * pv
                    if herm and hbit:
                        contrib = np.where(((t >> np.uint64(hbit - 1)) & np.uint64(1)) == 0, contrib, 0.0)
                    return contrib
                # general evaluation from the class entry lists (every group carries them)
                p, nent, flat = meta & 0xffff, (meta >> 16) & 7, (meta >> 21) & 1
                D = np.zeros(1 << L, dtype=np.complex128)
                any_im = False
                for e in range(nent):
                    ent = int(G['ent'][e])
                    cnt, cls = ent & 0x7ff, ent >> 11
                    sacc = np.zeros(1 << L)
                    for j in range(cnt):
                        if check:                   # class == register bits of the z-pattern
                            assert (int(zt[p + j]) >> 5) & (NA - 1) == cls & (NA - 1)
                        sacc += cf[p + j] * _sign(t0 & zt[p + j])
                    p += cnt
                    any_im |= cls >= NA
                    D += (1j if cls >= NA else 1.0) * sacc * _sign(reg & np.uint64(cls & (NA - 1)))
                if check:
                    assert bool(flat) == (nent == 1 and int(G['ent'][0]) >> 11 == 0)
                    assert bool((meta >> 20) & 1) == any_im and (ps['has_im'] or not any_im)
                if var != VAR_GENERAL:
                    # fast member: two inline slots, compile-time sign pattern
                    im1 = (meta >> 19) & 1
                    assert var < 2 * NA and bool(im1) == any_im and (var & (NA - 1)) == xr
                    s0 = G['cf0'] * _sign(t0 & np.uint64(G['zt0']))
                    s1 = G['cf1'] * _sign(t0 & np.uint64(G['zt1'])) * (1j if im1 else 1.0)
                    Df = s0 + s1 if var < NA else s0 + s1 * _sign(reg & np.uint64(xr))
                    if check:
                        assert np.allclose(Df, D, rtol=0, atol=1e-13 * (1.0 + np.abs(D).max()))
                    D = Df.astype(np.complex128)
                elif check:
                    assert not (meta >> 19) & 1
                contrib = D * pv
                if herm and hbit:
                    contrib = np.where(((t >> np.uint64(hbit - 1)) & np.uint64(1)) == 0, contrib, 0.0)
                return contrib

            def quad_member(G, meta):
                """D of a quadratic group through the separable tables the kernel rebuilds per tile,
                checked against the direct sum over its patterns."""
                nW = L - 5 - rb
                raw = G.tobytes()
                area = np.frombuffer(raw[8:32], dtype=np.uint8)
                toff = int(np.frombuffer(raw[30:32], dtype='<u2')[0])
                nseg = 2 + nW + 2 * rb + 1
                pfirst = meta & 0xffff
                npat_g = int(area[nseg])
                zs = zt[pfirst:pfirst + npat_g].astype(np.uint64)
                cs = cf[pfirst:pfirst + npat_g]
                direct = np.zeros(1 << L)
                for zz, cc in zip(zs, cs):
                    direct += cc * _sign(t & zz)
                if check:
                    assert area[0] == 0 and all(area[i] <= area[i + 1] for i in range(nseg))
                    assert toff % ((1 + nW + rb) * 32 + (1 + rb) * (1 << nW) + NA) == 0
                    assert all(bin(int(zz)).count('1') <= 2 for zz in zs)

                def seg_sum(seg, M, filt=None):
                    v = 0.0
                    for q in range(int(area[seg]), int(area[seg + 1])):
                        if filt is not None and (int(zs[q]) >> 5) & (NA - 1) != filt:
                            continue
                        v += cs[q] * (1.0 - 2.0 * (bin(int(zs[q]) & M).count('1') & 1))
                    return v
                wsz = 1 << nW
                lane_tabs = [[seg_sum(0 if tb == 0 else tb + 1, l) for l in range(32)] for tb in range(1 + nW + rb)]
                w_tabs = [[seg_sum(1 if tb == 0 else 2 + nW + rb + tb - 1, w << (5 + rb)) for w in range(wsz)]
                          for tb in range(1 + rb)]
                ktab = [seg_sum(2 + nW + 2 * rb, 0, filt=m) for m in range(NA)]
                lane_i = (t & np.uint64(31)).astype(np.int64)
                w_i = (t >> np.uint64(5 + rb)).astype(np.int64)
                r_i = reg.astype(np.int64)
                LT, WT, KT = np.array(lane_tabs), np.array(w_tabs), np.array(ktab)
                D = LT[0][lane_i] + WT[0][w_i]
                for j in range(nW):
                    D = D + LT[1 + j][lane_i] * (1.0 - 2.0 * ((w_i >> j) & 1))
                for i in range(rb):
                    D = D + (LT[1 + nW + i][lane_i] + WT[1 + i][w_i]) * (1.0 - 2.0 * ((r_i >> i) & 1))
                for m in range(NA):
                    if bin(m).count('1') == 2:
                        D = D + KT[m] * (1.0 - 2.0 * (np.bitwise_count((r_i & m).astype(np.uint64)) & 1))
                if check:
                    assert np.allclose(D, direct, rtol=0, atol=1e-12 * (1.0 + np.abs(direct).max()))
                return D.astype(np.complex128)

            def hbit_check(hbit, xo, remote):
                if hbit:
                    assert not remote and xo and hbit - 1 >= 5 + rb and (xo >> (hbit - 1)) == 1
                else:
                    assert remote or xo == 0 or xo.bit_length() - 1 < 5 + rb

            # singles: sorted by s = im*NA + xr, no bundle record
            ns = int(ch['ns'])
            g = 0
            for sidx in range(32):
                for _ in range(int(ch['s_cnt'][sidx])):
                    G = grp[g]
                    meta = int(G['meta'])
                    xt = int(G['xt'])
                    if check:
                        assert sidx < 2 * NA and (ps['has_im'] or sidx < NA)
                        assert xt >> 31 == 0 and ((meta >> 22) & 63) != VAR_GENERAL
                        assert (xt >> 5) & (NA - 1) == sidx & (NA - 1) and (meta >> 19) & 1 == sidx // NA
                        hbit_check(meta >> 28, xt & ~int(regmask), 0)
                    acc += member(G, g, xt & ~int(regmask), 0, meta >> 28, int(G['xl']) & ~regq)
                    g += 1
            assert g == ns
            if check:       # the chunk's quadratic groups are listed in quad_idx[q0 : q0 + nq)
                q0 = int(ch['q0']) & 0x7fffffff
                listed = [int(v) for v in ps['quad_idx'][q0:q0 + int(ch['nq'])]]
                longest = max([int(ps['pat_tptr'][p0 + q + 1]) - int(ps['pat_tptr'][p0 + q]) for q in range(npat)] or [0])
                assert (int(ch['q0']) >> 31) == (1 if longest > 24 else 0)
                actual = [gi for gi in range(ng) if (int(grp[gi]['meta']) >> 22) & 63 == VAR_QUAD]
                assert listed == actual
            seen = ns
            for B in bundles[int(ch['b0']):int(ch['b0']) + int(ch['nb'])]:
                xo, remote = int(B['xo']) & 0x7fffffff, int(B['xo']) >> 31
                nmem, hbit, first = int(B['meta']) & 0xff, (int(B['meta']) >> 8) & 0xff, int(B['meta']) >> 16
                if check:
                    assert first == seen and nmem >= 1 and xo & int(regmask) == 0 and int(B['xlo']) & regq == 0
                    hbit_check(hbit, xo, remote)
                seen += nmem
                for g in range(first, first + nmem):
                    acc += member(grp[g], g, xo, remote, hbit, int(B['xlo']))
            assert seen == ng
        if y is not None:
            y[idx] += acc
        total += np.vdot(tile, acc)
    return complex(total.real, 0.0) if herm else total


def run_part(passes, src, y=None, herm=False):
    total = 0.0 + 0.0j
    for ps in passes:
        total += run_pass(ps, src, y, herm=herm)
    return total
