"""CPU transliteration of psn_wide4_fwd_kernel's index arithmetic (csrc/psnode_wide4_fwd.cu): shared-memory tiles as byte arrays,
TMEM as a (128 lanes x 512 columns) array, tcgen05.mma as `D[lane(r)][d + n] (+)= sum_k A[r][k] B[n][k]` over one K = 8 slice with the
canonical no-swizzle K-major operand addressing of psnode_tc.cuh, one group of 16 trajectories at a time.  It checks the folding, the
weight-tile / descriptor pairs, the K-partial bookkeeping, the M = 64 row -> lane map of the output layer and the held-input / event
staging against the oracle WITHOUT a GPU (the hardware semantics it assumes are the ones the shipped wide / tc8 kernels rely on).
    python tests/wide4_emulate.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import psnode_oracle as O  # noqa: E402

H, TN, XP, ZMAX, M4 = 128, 16, 16, 8, 64
LBO, LBO_W = 144, 128
SBO_ACT = (H // 4) * LBO
ACT_TILE = (TN // 8) * SBO_ACT
SBO_W = (H // 4) * LBO_W
SBO_F = (XP // 4) * LBO_W
TM_W2_HI, TM_W2_LO, TM_W3_HI, TM_ACC = 0, 128, 256, 384
NP, KPI = 4, 4


def tile_byte(row, k, lbo, sbo):
    return (row >> 3) * sbo + (k >> 2) * lbo + (row & 7) * 16 + (k & 3) * 4


def split(v):
    v = np.float32(v)
    hi = ((v.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, np.float32(v - hi)


class Smem:
    def __init__(self, nbytes):
        self.f = np.full(nbytes // 4, np.nan, dtype=np.float32)     # NaN = never written: any read of it poisons the result

    def st(self, byte, v):
        assert byte % 4 == 0 and 0 <= byte // 4 < self.f.size, byte
        self.f[byte // 4] = v

    def ld(self, byte):
        return self.f[byte // 4]


def mma(tmem, d_col, M, a, b, accumulate):
    """a: ('s', smem, start, lbo, sbo) or ('t', col); b: (smem, start, lbo, sbo).  One K = 8 slice, N = 16."""
    A = np.zeros((M, 8), dtype=np.float64)
    for r in range(M):
        for k in range(8):
            A[r, k] = a[1].ld(a[2] + tile_byte(r, k, a[3], a[4])) if a[0] == 's' else tmem[r, a[1] + k]
    Bm = np.zeros((TN, 8), dtype=np.float64)
    for n in range(TN):
        for k in range(8):
            Bm[n, k] = b[0].ld(b[1] + tile_byte(n, k, b[2], b[3]))
    D = A @ Bm.T
    for r in range(M):
        lane = r if M == 128 else (r % 16) + 32 * (r // 16)
        for n in range(TN):
            tmem[lane, d_col + n] = (tmem[lane, d_col + n] if accumulate else 0.0) + D[r, n]


def emulate_group(method, W, b, t, x, z, a0, event_idx, z_jump, b0, tape=None):
    """W, b: the four layers (numpy float32); series time-major numpy; returns x_sol rows of trajectories b0..b0+15 (T, 16, X)."""
    T, B = t.shape[0], t.shape[1]
    X, Z = x.shape[2], z.shape[2]
    S = X + Z
    nst = {"euler": 1, "midpoint": 2, "rk4": 4}[method]
    W1, W2, W3, W4 = W
    Hh = W2.shape[0]                # the net's hidden width; neurons Hh..127 are zero padding
    tmem = np.full((128, 512), np.nan, dtype=np.float64)
    w3lo, w4hi, w4lo = Smem(H * H * 4), Smem(M4 * H * 4), Smem(M4 * H * 4)
    fxhi, fxlo = Smem(H * XP * 4), Smem(H * XP * 4)
    act_hi, act_lo = Smem(ACT_TILE), Smem(ACT_TILE)
    for m in range(H):
        for k in range(H):
            inb = m < Hh and k < Hh
            h2, l2 = split(W2[m, k] if inb else 0.0); tmem[m, TM_W2_HI + k] = h2; tmem[m, TM_W2_LO + k] = l2
            h3, l3 = split(W3[m, k] if inb else 0.0); tmem[m, TM_W3_HI + k] = h3
            w3lo.st(tile_byte(m, k, LBO_W, SBO_W), l3)
            if m < M4:
                h4, l4 = split(W4[m, k] if (m < X and k < Hh) else 0.0)
                w4hi.st(tile_byte(m, k, LBO_W, SBO_W), h4); w4lo.st(tile_byte(m, k, LBO_W, SBO_W), l4)
        for k in range(XP):
            fh, fl = split(np.float32(W1[m, S + k]) + np.float32(W1[m, 2 * S + k]) if (m < Hh and k < X) else 0.0)
            fxhi.st(tile_byte(m, k, LBO_W, SBO_F), fh); fxlo.st(tile_byte(m, k, LBO_W, SBO_F), fl)
    bb = [min(b0 + n, B - 1) for n in range(TN)]
    fz = np.zeros((H, ZMAX)); cst = np.zeros((H, TN))
    for m in range(Hh):
        for k in range(Z):
            fz[m, k] = np.float32(W1[m, S + X + k]) + np.float32(W1[m, 2 * S + X + k])
        for n in range(TN):
            acc = np.float32(b[0][m])
            for k in range(S):
                acc = np.float32(acc + (np.float32(W1[m, k]) - np.float32(W1[m, S + k])) * a0[bb[n], k])
            cst[m, n] = acc

    def store_tile(m, n, v):        # thread (m, h = n // 8), element i = n % 8: off0 + 16 i
        off0 = (n // 8) * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4
        hi, lo = split(v)
        act_hi.st(off0 + 16 * (n % 8), hi); act_lo.st(off0 + 16 * (n % 8), lo)

    KB, KW = 2 * LBO, 2 * LBO_W     # descriptor K-step advance, in bytes
    acc_base = TM_ACC               # group 0

    def layer(kind):
        for wq in range(NP):
            my = acc_base + wq * TN
            if kind == 1:
                if wq >= 2:
                    continue
                for t_i, (a_s, b_s) in enumerate(((fxlo, act_hi), (fxhi, act_lo), (fxhi, act_hi))):
                    mma(tmem, my, 128, ('s', a_s, KW * wq, LBO_W, SBO_F), (b_s, KB * wq, LBO, SBO_ACT), t_i > 0)
                continue
            first = True
            for term in range(3):
                b_s = act_lo if term == 1 else act_hi
                for kk in range(KPI):
                    ks = KPI * wq + kk
                    if kind == 2:
                        a = ('t', (TM_W2_LO if term == 0 else TM_W2_HI) + 8 * ks)
                        M = 128
                    elif kind == 3:
                        a = ('s', w3lo, KW * ks, LBO_W, SBO_W) if term == 0 else ('t', TM_W3_HI + 8 * ks)
                        M = 128
                    else:
                        a = ('s', w4lo if term == 0 else w4hi, KW * ks, LBO_W, SBO_W)
                        M = M4
                    mma(tmem, my, M, a, (b_s, KB * ks, LBO, SBO_ACT), not first)
                    first = False

    def collect(npart):             # thread (m, n) reads TMEM lane m, column acc_base + p * 16 + n
        d = np.zeros((H, TN))
        for p in range(npart):
            d += tmem[:, acc_base + p * TN: acc_base + (p + 1) * TN]
        return d

    elu = lambda v: np.where(v > 0, v, np.expm1(np.minimum(v, 0)))
    sol = np.zeros((T, TN, X), dtype=np.float32)
    x0 = np.zeros((XP, TN))
    for m in range(X):
        for n in range(TN):
            x0[m, n] = x[0, bb[n], m]
    sol[0] = x0[:X].T
    for m in range(XP):
        for n in range(TN):
            store_tile(m, n, x0[m, n])
    c13 = np.float32(1.0 / 3.0)
    for j in range(1, T):
        ek = int(event_idx[j - 1]) if event_idx is not None else -1
        zh = np.zeros((TN, ZMAX))
        for n in range(TN):
            zh[n, :Z] = z_jump[bb[n], ek, :] if ek >= 0 else z[j - 1, bb[n], :]
        dt = np.array([t[j, bb[n], 0] - t[j - 1, bb[n], 0] for n in range(TN)], dtype=np.float32)
        pre = cst + fz @ zh.T
        k1 = k2 = k3 = xn = None
        for e in range(nst):
            rec = {"y": (xn if e > 0 else x0).astype(np.float32).copy()} if tape is not None else None     # [16 state rows][16 trajectories]
            layer(1); a = elu(collect(2) + pre)
            if rec is not None: rec["a1"] = a.astype(np.float32)
            for m in range(H):
                for n in range(TN):
                    store_tile(m, n, a[m, n])
            for kind, bias in ((2, b[1]), (3, b[2])):
                layer(kind); a = elu(collect(4) + np.concatenate([bias, np.zeros(H - Hh)])[:, None])
                if rec is not None: rec["a%d" % kind] = a.astype(np.float32)
                for m in range(H):
                    for n in range(TN):
                        store_tile(m, n, a[m, n])
            if rec is not None: tape[(j, e)] = rec
            layer(4)
            d = collect(4)[:XP]                    # the state threads: lanes 0..15
            kk = d + np.concatenate([b[3], np.zeros(XP - X)])[:, None]
            if method == "euler":
                xn = x0 + dt * kk
            elif method == "midpoint":
                xn = x0 + kk * (0.5 * dt) if e == 0 else x0 + dt * kk
            else:
                if e == 0: k1 = kk; xn = x0 + dt * kk * c13
                elif e == 1: k2 = kk; xn = x0 + dt * (kk - k1 * c13)
                elif e == 2: k3 = kk; xn = x0 + dt * (k1 - k2 + kk)
                else: xn = x0 + (k1 + 3.0 * (k2 + k3) + kk) * dt * 0.125
            xn[X:] = 0.0
            for m in range(XP):
                for n in range(TN):
                    store_tile(m, n, xn[m, n])
            if e == nst - 1:
                x0 = xn
                sol[j] = xn[:X].T
    return sol


F_DFX, F_DFZ, F_DCA, F_DB1, F_DB2, F_DB3, F_DW4, F_DB4, SLAB_FIELDS = 0, 16, 24, 48, 49, 50, 51, 67, 68


def emulate_group_bwd(method, W, t, z, a0, event_idx, z_jump, gx, b0, tape):
    """Transliteration of psn_wide4_bwd_kernel for one group: returns (slabs [2][68][128], delta records {(j, e): (d2, d3)}, d_x0 [16][X],
    d_a0 [16][S])."""
    T, B = t.shape[0], t.shape[1]
    W1, W2, W3, W4 = W
    Hh, X = W2.shape[0], W4.shape[0]
    Z = z.shape[2]
    S = X + Z
    nst = {"euler": 1, "midpoint": 2, "rk4": 4}[method]
    TM_A_HI, TM_A_LO, TM_B_HI = 0, 128, 256
    tmem = np.full((128, 512), np.nan, dtype=np.float64)
    w2t_lo, fxt_hi, fxt_lo = Smem(H * H * 4), Smem(M4 * H * 4), Smem(M4 * H * 4)
    w4t_hi, w4t_lo = Smem(H * XP * 4), Smem(H * XP * 4)
    act_hi, act_lo = Smem(ACT_TILE), Smem(ACT_TILE)
    for m in range(H):
        for k in range(H):
            inb = m < Hh and k < Hh
            ah, al = split(W3[k, m] if inb else 0.0); tmem[m, TM_A_HI + k] = ah; tmem[m, TM_A_LO + k] = al
            bh, bl = split(W2[k, m] if inb else 0.0); tmem[m, TM_B_HI + k] = bh
            w2t_lo.st(tile_byte(m, k, LBO_W, SBO_W), bl)
            if m < M4:
                fh, fl = split(np.float32(W1[k, S + m]) + np.float32(W1[k, 2 * S + m]) if (m < X and k < Hh) else 0.0)
                fxt_hi.st(tile_byte(m, k, LBO_W, SBO_W), fh); fxt_lo.st(tile_byte(m, k, LBO_W, SBO_W), fl)
        for k in range(XP):
            wh, wl = split(W4[k, m] if (m < Hh and k < X) else 0.0)
            w4t_hi.st(tile_byte(m, k, LBO_W, SBO_F), wh); w4t_lo.st(tile_byte(m, k, LBO_W, SBO_F), wl)
    bb = [min(b0 + n, B - 1) for n in range(TN)]

    def store_tile(m, n, v):
        off0 = (n // 8) * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4
        hi, lo = split(v)
        act_hi.st(off0 + 16 * (n % 8), hi); act_lo.st(off0 + 16 * (n % 8), lo)

    KB, KW = 2 * LBO, 2 * LBO_W
    acc_base = TM_ACC

    def layer(kind):
        for wq in range(NP):
            my = acc_base + wq * TN
            if kind == "w4t":
                if wq >= 2:
                    continue
                for t_i, (a_s, b_s) in enumerate(((w4t_lo, act_hi), (w4t_hi, act_lo), (w4t_hi, act_hi))):
                    mma(tmem, my, 128, ('s', a_s, KW * wq, LBO_W, SBO_F), (b_s, KB * wq, LBO, SBO_ACT), t_i > 0)
                continue
            first = True
            if kind == "w2t":
                for kk in range(KPI):
                    ks = KPI * wq + kk
                    mma(tmem, my, 128, ('s', w2t_lo, KW * ks, LBO_W, SBO_W), (act_hi, KB * ks, LBO, SBO_ACT), kk > 0)
                for term in (1, 2):
                    b_s = act_lo if term == 1 else act_hi
                    for kk in range(KPI):
                        ks = KPI * wq + kk
                        mma(tmem, my, 128, ('t', TM_B_HI + 8 * ks), (b_s, KB * ks, LBO, SBO_ACT), True)
                continue
            for term in range(3):
                b_s = act_lo if term == 1 else act_hi
                for kk in range(KPI):
                    ks = KPI * wq + kk
                    if kind == "w3t":
                        a, M = ('t', (TM_A_LO if term == 0 else TM_A_HI) + 8 * ks), 128
                    else:
                        a, M = ('s', fxt_lo if term == 0 else fxt_hi, KW * ks, LBO_W, SBO_W), M4
                    mma(tmem, my, M, a, (b_s, KB * ks, LBO, SBO_ACT), not first)
                    first = False

    def collect(npart):
        d = np.zeros((H, TN))
        for p in range(npart):
            d += tmem[:, acc_base + p * TN: acc_base + (p + 1) * TN]
        return d

    elu_g = lambda a: np.where(a > 0, 1.0, a + 1.0)

    def load_gx(j):
        v = np.zeros((XP, TN))
        for n in range(TN):
            if b0 + n < B:
                v[:X, n] = gx[j, b0 + n, :]
        return v

    lam = load_gx(T - 1)
    dw4 = np.zeros((H, TN, XP)); dfx = np.zeros((H, TN, XP)); dfz = np.zeros((H, TN, ZMAX))      # per thread (m, n) terms, reduced per half below
    dcs = np.zeros((H, TN)); db2 = np.zeros((H, TN)); db3 = np.zeros((H, TN)); db4 = np.zeros((XP, TN))
    recs = {}
    for j in range(T - 1, 0, -1):
        ek = int(event_idx[j - 1]) if event_idx is not None else -1
        zh = np.zeros((TN, ZMAX))
        for n in range(TN):
            zh[n, :Z] = z_jump[bb[n], ek, :] if ek >= 0 else z[j - 1, bb[n], :]
        dt = np.array([t[j, bb[n], 0] - t[j - 1, bb[n], 0] for n in range(TN)], dtype=np.float32)
        sum1 = np.zeros((H, TN)); dysum = np.zeros((XP, TN)); dyA = np.zeros((XP, TN)); dyB = np.zeros((XP, TN))
        for e in range(nst - 1, -1, -1):
            rec = tape[(j, e)]
            if method == "euler":
                dk = lam * dt
            elif method == "midpoint":
                dk = lam * dt if e == 1 else 0.5 * dt * dyA
            else:
                l8 = lam * (dt * 0.125)
                dk = l8 if e == 3 else (dt * dyA + 3.0 * l8 if e == 2 else (dt * (dyB - dyA) + 3.0 * l8 if e == 1 else l8 + dyB))
            db4 += dk
            dkt = dk.T.copy()                      # [n][k]
            yt = rec["y"].T.copy()                 # [n][k]
            for m in range(XP):
                for n in range(TN):
                    store_tile(m, n, dk[m, n])
            layer("w4t")
            a3 = rec["a3"]
            dw4 += a3[:, :, None] * dkt[None, :, :]
            d3 = collect(2) * elu_g(a3); db3 += d3
            for m in range(H):
                for n in range(TN):
                    store_tile(m, n, d3[m, n])
            layer("w3t")
            d2 = collect(4) * elu_g(rec["a2"]); db2 += d2
            for m in range(H):
                for n in range(TN):
                    store_tile(m, n, d2[m, n])
            layer("w2t")
            d1 = collect(4) * elu_g(rec["a1"]); sum1 += d1
            dfx += d1[:, :, None] * yt[None, :, :]
            for m in range(H):
                for n in range(TN):
                    store_tile(m, n, d1[m, n])
            layer("fxt")
            dy = collect(4)[:XP]                   # state threads: lanes 0..15 of sub-partition 0
            dysum += dy
            if method == "midpoint":
                dyA = dy
            if method == "rk4":
                if e == 3: dyA = dy
                elif e == 2: dyB = dy
                elif e == 1: dyB = dt * (1.0 / 3.0) * (dy - dyB) + dt * dyA
            recs[(j, e)] = (d2.astype(np.float32), d3.astype(np.float32))
        dfz += sum1[:, :, None] * zh[None, :, :]
        dcs += sum1
        lam = (lam + dysum) + load_gx(j - 1)
    slabs = np.zeros((2, SLAB_FIELDS, H))
    for h in range(2):
        ns = slice(8 * h, 8 * h + 8)
        slabs[h, F_DFX:F_DFX + XP] = dfx[:, ns, :].sum(axis=1).T
        slabs[h, F_DW4:F_DW4 + XP] = dw4[:, ns, :].sum(axis=1).T
        slabs[h, F_DFZ:F_DFZ + ZMAX] = dfz[:, ns, :].sum(axis=1).T
        for k in range(S):
            slabs[h, F_DCA + k] = sum(dcs[:, n] * a0[bb[n], k] for n in range(8 * h, 8 * h + 8))
        slabs[h, F_DB1] = dcs[:, ns].sum(axis=1)
        slabs[h, F_DB2] = db2[:, ns].sum(axis=1)
        slabs[h, F_DB3] = db3[:, ns].sum(axis=1)
        slabs[h, F_DB4, :XP] = db4[:, ns].sum(axis=1)
    d_a0 = np.zeros((TN, S))
    for n in range(TN):
        for k in range(S):
            d_a0[n, k] = sum(dcs[mm, n] * (np.float64(W1[mm, k]) - np.float64(W1[mm, S + k])) for mm in range(Hh))
    return slabs, recs, lam[:X].T.copy(), d_a0


def assemble(slabs, g2, g3, X, Z, Hh):
    """psn_wide4_assemble_kernel: slabs [nslab][68][128]; g2 / g3 = the 128 x 128 products of the block-GEMM pass."""
    S = X + Z
    fs = slabs.sum(axis=0)
    out = []
    dW1 = np.zeros((Hh, 3 * S))
    for m in range(Hh):
        for col in range(3 * S):
            blk, k = divmod(col, S)
            a = fs[F_DCA + k, m]
            gsum = fs[F_DFX + k, m] if k < X else fs[F_DFZ + (k - X), m]
            dW1[m, col] = a if blk == 0 else (gsum - a if blk == 1 else gsum)
    out += [dW1, fs[F_DB1, :Hh], g2[:Hh, :Hh], fs[F_DB2, :Hh], g3[:Hh, :Hh], fs[F_DB3, :Hh],
            np.stack([fs[F_DW4 + k, :Hh] for k in range(X)]), fs[F_DB4, :X]]
    return out


BWD_CASES = (("rk4", 16, 2, 16, 3, 0, 128), ("midpoint", 5, 3, 21, 3, 1, 100), ("euler", 16, 8, 7, 3, 1, 72))
FWD_CASES = (("rk4", 16, 2, 16, 3, 0, 128), ("midpoint", 5, 3, 21, 4, 1, 100), ("euler", 16, 8, 7, 3, 1, 72))


def main_bwd(cases=BWD_CASES):
    from py_psnode_b200.neural_base import DE_Func
    for method, X, Z, B, N, events, Hn in cases:
        torch.manual_seed(5)
        T = N + 1
        de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=Hn)
        params = [(m.weight.detach(), m.bias.detach()) for m in de.x_dot if isinstance(m, torch.nn.Linear)]
        t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
        x = torch.randn(T, B, X) * 0.1
        z = torch.randn(T, B, Z) * 0.1
        a0 = torch.cat((x[0], z[0]), dim=-1)
        gx = torch.randn(T, B, X)
        event_t = z_jump = event_idx = None
        if events:
            event_t = t[1, :, 0].view(B, 1, 1).clone()
            z_jump = torch.randn(B, 1, Z) * 0.1
            event_idx = np.full(T - 1, -1, dtype=np.int32)
            event_idx[1] = 0
        # float64 autograd through the oracle
        p64 = [(Wt.double().requires_grad_(True), bt.double().requires_grad_(True)) for Wt, bt in params]
        x64 = x.double().requires_grad_(True)
        a64 = a0.double().requires_grad_(True)
        sol = O.integrate_ode(method, p64, t.double(), x64, z.double(), a64, None if event_t is None else event_t.double(),
                              None if z_jump is None else z_jump.double())
        (sol * gx.double()).sum().backward()
        want = [g.grad.numpy() for pr in p64 for g in pr]
        W = [p[0].numpy() for p in params]
        bnp = [p[1].numpy().astype(np.float64) for p in params]
        all_slabs, g2, g3 = [], np.zeros((H, H)), np.zeros((H, H))
        worst = 0.0
        for b0 in range(0, B, TN):
            tape = {}
            emulate_group(method, W, bnp, t.numpy(), x.numpy(), z.numpy(), a0.numpy(), event_idx, None if z_jump is None else z_jump.numpy(), b0, tape)
            slabs, recs, d_x0, d_a0 = emulate_group_bwd(method, W, t.numpy(), z.numpy(), a0.numpy(), event_idx,
                                                        None if z_jump is None else z_jump.numpy(), gx.numpy(), b0, tape)
            all_slabs += [slabs[0], slabs[1]]
            for key, (d2, d3) in recs.items():     # psn_wide_grad_pairs: (delta2, a1) and (delta3, a2) records
                g2 += d2.astype(np.float64) @ tape[key]["a1"].astype(np.float64).T
                g3 += d3.astype(np.float64) @ tape[key]["a2"].astype(np.float64).T
            nlive = min(TN, B - b0)
            ex = np.abs(d_x0[:nlive] - x64.grad[0, b0:b0 + nlive].numpy()).max()
            ea = np.abs(d_a0[:nlive] - a64.grad[b0:b0 + nlive].numpy()).max()
            worst = max(worst, ex / np.abs(x64.grad[0].numpy()).max(), ea / np.abs(a64.grad.numpy()).max())
        got = assemble(np.stack(all_slabs), g2, g3, X, Z, Hn)
        for name, gg, ww in zip(("dW1", "db1", "dW2", "db2", "dW3", "db3", "dW4", "db4"), got, want):
            assert gg.shape == ww.shape, (name, gg.shape, ww.shape)
            rel = np.abs(gg - ww).max() / np.abs(ww).max()
            worst = max(worst, rel)
            print(f"  {name}: max rel err {rel:.2e}", flush=True)
        print(f"{method} X={X} Z={Z} H={Hn} B={B}: worst relative gradient error {worst:.2e}", flush=True)
        assert worst < 2e-6, worst
    print("bwd ok")


def main(cases=None):
    from py_psnode_b200.neural_base import DE_Func
    worst = 0.0
    for method, X, Z, B, N, events, Hn in (FWD_CASES if cases is None else cases):
        torch.manual_seed(3)
        T = N + 1
        de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=Hn)
        params = [(m.weight.detach(), m.bias.detach()) for m in de.x_dot if isinstance(m, torch.nn.Linear)]
        t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
        x = torch.randn(T, B, X) * 0.1
        z = torch.randn(T, B, Z) * 0.1
        a0 = torch.cat((x[0], z[0]), dim=-1)
        event_t = z_jump = event_idx = None
        if events:
            event_t = t[1, :, 0].view(B, 1, 1).clone()
            z_jump = torch.randn(B, 1, Z) * 0.1
            event_idx = np.full(T - 1, -1, dtype=np.int32)
            event_idx[1] = 0
        want = O.integrate_ode(method, params, t, x, z, a0, event_t, z_jump).numpy()
        W = [p[0].numpy() for p in params]
        b = [p[1].numpy().astype(np.float64) for p in params]
        for b0 in range(0, B, TN):
            got = emulate_group(method, W, b, t.numpy(), x.numpy(), z.numpy(), a0.numpy(), event_idx,
                                None if z_jump is None else z_jump.numpy(), b0)
            nlive = min(TN, B - b0)
            err = np.abs(got[:, :nlive] - want[:, b0:b0 + nlive]).max()
            assert np.isfinite(got[:, :nlive]).all(), "a never-written operand was read"
            worst = max(worst, err)
            print(f"{method} X={X} Z={Z} H={Hn} B={B} group@{b0}: max|emulated - oracle| = {err:.2e}", flush=True)
    assert worst < 2e-6, worst
    print("ok")


if __name__ == "__main__":
    main()
    main_bwd()
