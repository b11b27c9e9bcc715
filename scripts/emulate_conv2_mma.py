"""Desk-check of gennbv_b200/csrc/conv2_mma.cu's index logic on the CPU: a literal Python transliteration of the per-lane
fragment assembly (weight staging order, k-slot permutation, row -> voxel decoding, parity classes / tap sets / validity bits
of the data gradient, accumulator -> output mapping) around an emulated m16n8k8 MMA with the documented PTX fragment
layouts, compared with torch conv3d and its autograd gradient in float64.  (The hi/lo TF32 split is not emulated: the lo
fragments are zero and the arithmetic exact.)   python scripts/emulate_conv2_mma.py [G1 ...]"""
import sys
import numpy as np, torch
torch.manual_seed(0)
C, NT = 16, 27
MT, WARPS = 4, 4
RPW, RPI = 16 * MT, 16 * MT * WARPS

def mma(acc, a, b):
    # acc[lane][4], a[lane][4], b[lane][2]
    A = np.zeros((16, 8)); Bm = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
        Bm[t, g], Bm[t + 4, g] = b[lane]
    D = A @ Bm
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += D[g, 2 * t]; acc[lane][1] += D[g, 2 * t + 1]; acc[lane][2] += D[g + 8, 2 * t]; acc[lane][3] += D[g + 8, 2 * t + 1]

def stage_weights(w, dgrad):
    wsm = np.zeros((NT * 4 * 32, 4))
    for idx in range(NT * 128):
        tap, q, ln = idx >> 7, (idx >> 5) & 3, idx & 31
        gg, tt, j = ln >> 2, ln & 3, q >> 1
        for e in range(4):
            if not dgrad:
                co = 8 * j + gg
                wsm[idx, e] = w[(co * C + (4 * tt + e)) * NT + tap]
            else:
                ci = 8 * j + gg
                wsm[idx, e] = w[((4 * tt + e) * C + ci) * NT + tap]
        if q & 1: wsm[idx] = 0  # lo part: zero in this exact emulation
    return wsm

def fwd(y1, sc, sh, w, bias, G1, G2, B):
    P1, P2 = G1 ** 3, G2 ** 3
    chunks = -(-P2 // RPI)
    y2 = np.full((B, C, P2), np.nan)
    wsm = stage_weights(w, False)
    for item in range(B * chunks):
        b, chunk = divmod(item, chunks)
        for warp in range(WARPS):
            row0 = chunk * RPI + warp * RPW
            acc = np.zeros((MT, 2, 32, 4))
            for lane in range(32):
                t = lane & 3
                for m in range(MT):
                    for j in range(2):
                        acc[m, j, lane, 0] = acc[m, j, lane, 2] = bias[8 * j + 2 * t]
                        acc[m, j, lane, 1] = acc[m, j, lane, 3] = bias[8 * j + 2 * t + 1]
            for ij in range(9):
                i, jy = divmod(ij, 3)
                tap_off = ((i * G1 + jy) * G1) * C
                for l in range(3):
                    tap = ij * 3 + l
                    for m in range(MT):
                        vals = np.zeros((32, 2, 4))
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            for h in range(2):
                                p = min(row0 + m * 16 + g + 8 * h, P2 - 1)
                                z2 = p % G2; r = p // G2; yy2 = r % G2; x2 = r // G2
                                off = (((2 * x2) * G1 + 2 * yy2) * G1 + 2 * z2) * C
                                base = b * P1 * C + 4 * t + off + tap_off + l * C
                                raw = y1[base:base + 4]
                                vals[lane, h] = np.maximum(sc[4 * t:4 * t + 4] * raw + sh[4 * t:4 * t + 4], 0)
                        for ks in range(2):
                            e0, e1 = 2 * ks, 2 * ks + 1
                            for j in range(2):
                                a = [(vals[ln, 0, e0], vals[ln, 1, e0], vals[ln, 0, e1], vals[ln, 1, e1]) for ln in range(32)]
                                bfr = [(wsm[tap * 128 + (2 * j) * 32 + ln, e0], wsm[tap * 128 + (2 * j) * 32 + ln, e1]) for ln in range(32)]
                                mma(acc[m, j], a, bfr)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for m in range(MT):
                    for h in range(2):
                        p = row0 + m * 16 + g + 8 * h
                        if p < P2:
                            for j in range(2):
                                for e in range(2):
                                    y2[b, 8 * j + 2 * t + e, p] = acc[m, j, lane, 2 * h + e]
    return y2

def dgrad_class(cls, G1):
    NE, NO = (G1 + 1) // 2, G1 // 2
    cx, cy, cz = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
    nx, ny, nz = (NO if cx else NE), (NO if cy else NE), (NO if cz else NE)
    nv = nx * ny * nz
    return cx, cy, cz, nx, ny, nz, nv, -(-nv // RPI)

def dgrad(dy2cl, w, G1, G2, B):
    P1, P2 = G1 ** 3, G2 ** 3
    ips = sum(dgrad_class(c, G1)[7] for c in range(8))
    out = np.full((B, P1, C), np.nan)
    wsm = stage_weights(w, True)
    for item in range(B * ips):
        b, r = divmod(item, ips)
        for cls in range(8):
            cx, cy, cz, nx, ny, nz, nvox, chunks = dgrad_class(cls, G1)
            if r < chunks: break
            r -= chunks
        for warp in range(WARPS):
            row0 = r * RPI + warp * RPW
            acc = np.zeros((MT, 2, 32, 4))
            info = {}
            for lane in range(32):
                g = lane >> 2
                for m in range(MT):
                    for h in range(2):
                        v = row0 + m * 16 + g + 8 * h
                        inn = v < nvox
                        vv = v if inn else 0
                        zi = vv % nz; q = vv // nz; yi = q % ny; xi = q // ny
                        src = ((xi * G2 + yi) * G2 + zi) * C
                        dst = ((2 * xi + cx) * G1 + (2 * yi + cy)) * G1 + (2 * zi + cz) if inn else -1
                        bits = 0
                        if inn:
                            bits = (1 if xi < G2 else 0) | (2 if xi >= 1 else 0) | (4 if yi < G2 else 0) | (8 if yi >= 1 else 0) | (16 if zi < G2 else 0) | (32 if zi >= 1 else 0)
                        info[(lane, m, h)] = (src, dst, bits)
            nix, niy, niz = (1 if cx else 2), (1 if cy else 2), (1 if cz else 2)
            for ax in range(nix):
                for ay in range(niy):
                    for az in range(niz):
                        i = 1 if cx else 2 * ax; jy = 1 if cy else 2 * ay; l = 1 if cz else 2 * az
                        tap = (i * 3 + jy) * 3 + l
                        delta = ((ax * G2 + ay) * G2 + az) * C
                        need = (1 << ax) | (4 << ay) | (16 << az)
                        for m in range(MT):
                            vals = np.zeros((32, 2, 4))
                            for lane in range(32):
                                t = lane & 3
                                for h in range(2):
                                    src, dst, bits = info[(lane, m, h)]
                                    if (bits & need) == need:
                                        base = b * P2 * C + 4 * t + src - delta
                                        assert base >= b * P2 * C and base + 4 <= (b + 1) * P2 * C
                                        vals[lane, h] = dy2cl[base:base + 4]
                            for ks in range(2):
                                e0, e1 = 2 * ks, 2 * ks + 1
                                for j in range(2):
                                    a = [(vals[ln, 0, e0], vals[ln, 1, e0], vals[ln, 0, e1], vals[ln, 1, e1]) for ln in range(32)]
                                    bfr = [(wsm[tap * 128 + (2 * j) * 32 + ln, e0], wsm[tap * 128 + (2 * j) * 32 + ln, e1]) for ln in range(32)]
                                    mma(acc[m, j], a, bfr)
            for lane in range(32):
                t = lane & 3
                for m in range(MT):
                    for h in range(2):
                        src, dst, bits = info[(lane, m, h)]
                        if dst < 0: continue
                        for j in range(2):
                            for e in range(2):
                                assert np.isnan(out[b, dst, 8 * j + 2 * t + e])
                                out[b, dst, 8 * j + 2 * t + e] = acc[m, j, lane, 2 * h + e]
    return out


def wgrad(y1, sc, sh, dy2cl, G1, G2, B, nblocks):
    """conv2_wgrad_mma_kernel: returns (dW [co,ci,tap] flat, db [16]) summed over the per-block records."""
    P1, P2 = G1 ** 3, G2 ** 3
    total_rows = B * G2 * G2
    rpb = -(-total_rows // nblocks)
    nblk = -(-total_rows // rpb)
    TPW = -(-NT // WARPS)
    rec = np.zeros((nblk, C * C * NT + C))
    ksteps = (G2 + 7) // 8
    for blk in range(nblk):
        r0, r1 = blk * rpb, min(total_rows, blk * rpb + rpb)
        for warp in range(WARPS):
            acc = np.zeros((TPW, 2, 32, 4))
            db_lo, db_hi = np.zeros(32), np.zeros(32)
            for row in range(r0, r1):
                b, rem = divmod(row, G2 * G2)
                x2, yy2 = divmod(rem, G2)
                dyrow = (b * P2 + (x2 * G2 + yy2) * G2) * C
                xin = (b * P1 + ((2 * x2) * G1 + 2 * yy2) * G1) * C
                for ks in range(ksteps):
                    a = []
                    zc = []
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        za, zb = 8 * ks + t, 8 * ks + t + 4
                        va, vb = za < G2, zb < G2
                        zac, zbc = min(za, G2 - 1), min(zb, G2 - 1)
                        a0 = dy2cl[dyrow + zac * C + g] if va else 0.0
                        a1 = dy2cl[dyrow + zac * C + 8 + g] if va else 0.0
                        a2 = dy2cl[dyrow + zbc * C + g] if vb else 0.0
                        a3 = dy2cl[dyrow + zbc * C + 8 + g] if vb else 0.0
                        db_lo[lane] += a0 + a2; db_hi[lane] += a1 + a3
                        a.append((a0, a1, a2, a3)); zc.append((zac, zbc))
                    for k in range(TPW):
                        tap = warp + WARPS * k
                        if tap >= NT: continue
                        i, jy, l = tap // 9, (tap // 3) % 3, tap % 3
                        line = xin + ((i * G1 + jy) * G1 + l) * C
                        for hf in range(2):
                            bfr = []
                            for lane in range(32):
                                g = lane >> 2
                                zac, zbc = zc[lane]
                                ci = 8 * hf + g
                                xa = max(sc[ci] * y1[line + 2 * zac * C + ci] + sh[ci], 0.0)
                                xb = max(sc[ci] * y1[line + 2 * zbc * C + ci] + sh[ci], 0.0)
                                bfr.append((xa, xb))
                            mma(acc[k, hf], a, bfr)
            for k in range(TPW):
                tap = warp + WARPS * k
                if tap >= NT: continue
                for hf in range(2):
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        for e in range(4):
                            co, ci = g + 8 * (e >> 1), 8 * hf + 2 * t + (e & 1)
                            rec[blk, (co * C + ci) * NT + tap] = acc[k, hf, lane, e]
            if warp == 0:
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    if t == 0:
                        rec[blk, C * C * NT + g] = sum(db_lo[4 * g + tt] for tt in range(4))
                        rec[blk, C * C * NT + 8 + g] = sum(db_hi[4 * g + tt] for tt in range(4))
    tot = rec.sum(0)
    return tot[:C * C * NT], tot[C * C * NT:]


def conv1_fwd(x, w, bias, G, G1, B, RBF):
    """conv1_fwd_mma_kernel: x [B,G,G,G] tri-class grid, w [16,27] -> y1 [B, G1^3, 16] channels-last (index logic only)."""
    TAPS = 27
    P1 = G1 ** 3
    NYB = -(-G1 // RBF)
    TL = (2 * RBF + 1) * G
    y1 = np.full((B, P1, C), np.nan)
    for rb in range(B * G1 * NYB):
        b, rem = divmod(rb, G1 * NYB)
        x1, yb = divmod(rem, NYB)
        y0 = yb * RBF
        nr = min(RBF, G1 - y0)
        ts = np.zeros(3 * TL + 64)
        for i in range(3):                     # the three TMA slabs: rows 2*y0 .. 2*y0 + 2*nr of plane 2*x1 + i
            ts[i * TL:i * TL + (2 * nr + 1) * G] = x[b, 2 * x1 + i, 2 * y0:2 * y0 + 2 * nr + 1, :].reshape(-1)
        out_base = (x1 * G1 + y0) * G1
        ZT = (G1 + 15) // 16
        for warp in range(8):
            for tile in range(warp, nr * ZT, 8):
                r, zt = divmod(tile, ZT)
                z0 = 16 * zt
                acc = np.zeros((2, 32, 4))
                for lane in range(32):
                    t = lane & 3
                    for j in range(2):
                        acc[j, lane, 0] = acc[j, lane, 2] = bias[8 * j + 2 * t]
                        acc[j, lane, 1] = acc[j, lane, 3] = bias[8 * j + 2 * t + 1]
                for s in range(4):
                    a = []
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        pa = (2 * r) * G + 2 * min(z0 + g, G1 - 1)
                        pb = (2 * r) * G + 2 * min(z0 + g + 8, G1 - 1)
                        toff = []
                        for q in range(2):
                            tc = min(8 * s + t + 4 * q, TAPS - 1)
                            toff.append((tc // 9) * TL + ((tc // 3) % 3) * G + tc % 3)
                        a.append((ts[pa + toff[0]], ts[pb + toff[0]], ts[pa + toff[1]], ts[pb + toff[1]]))
                    for j in range(2):
                        bfr = []
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            vals = []
                            for q in range(2):
                                tap = 8 * s + t + 4 * q
                                vals.append(w[(8 * j + g) * TAPS + tap] if tap < TAPS else 0.0)
                            bfr.append(tuple(vals))
                        mma(acc[j], a, bfr)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for h in range(2):
                        z1 = z0 + g + 8 * h
                        if z1 < G1:
                            for j in range(2):
                                for e in range(2):
                                    idx = out_base + r * G1 + z1
                                    assert np.isnan(y1[b, idx, 8 * j + 2 * t + e])
                                    y1[b, idx, 8 * j + 2 * t + e] = acc[j, lane, 2 * h + e]
    return y1


def check_conv1(G, B=1):
    G1 = (G - 3) // 2 + 1
    RBF = max(1, min(G1, 256 // ((G1 + 3) // 4)))
    x = torch.randint(-1, 2, (B, 1, G, G, G)).double()
    wt = torch.randn(C, 1, 3, 3, 3, dtype=torch.float64)
    bias = torch.randn(C, dtype=torch.float64)
    ref = torch.nn.functional.conv3d(x, wt, bias, stride=2).permute(0, 2, 3, 4, 1).reshape(B, -1, C).numpy()
    got = conv1_fwd(x[:, 0].numpy(), wt.numpy().reshape(-1), bias.numpy(), G, G1, B, RBF)
    err = np.abs(got - ref).max()
    print("G", G, "conv1 fwd max err", err, "unwritten", np.isnan(got).sum())
    assert err < 1e-12 and not np.isnan(got).any()


def conv1_wgrad(x, dy1, G, G1, B, nblocks=3):
    """conv1_wgrad_mma_kernel index logic (BN1 backward omitted: dy1 given): x [B,G,G,G], dy1 [B,P1,16] -> dW [16*27], db [16]."""
    TAPS, RB = 27, 8
    P1 = G1 ** 3
    NYB = -(-G1 // RB)
    TL = (2 * RB + 1) * G
    total_rb = B * G1 * NYB
    rpb = -(-total_rb // nblocks)
    rec = np.zeros((-(-total_rb // rpb), C * TAPS + C))
    KS = (G1 + 7) // 8
    for blk in range(rec.shape[0]):
        acc = np.zeros((8, 4, 32, 4))
        dbs = np.zeros((8, 32, 2))
        for rb in range(blk * rpb, min(total_rb, blk * rpb + rpb)):
            b, rem = divmod(rb, G1 * NYB)
            x1, yb = divmod(rem, NYB)
            y0 = yb * RB
            nr = min(RB, G1 - y0)
            ts = np.zeros(3 * TL + 64)
            for i in range(3):
                ts[i * TL:i * TL + (2 * nr + 1) * G] = x[b, 2 * x1 + i, 2 * y0:2 * y0 + 2 * nr + 1, :].reshape(-1)
            gs = dy1[b, (x1 * G1 + y0) * G1:(x1 * G1 + y0 + nr) * G1].reshape(-1)
            for warp in range(8):
                for ks in range(warp, nr * KS, 8):
                    r, kk = divmod(ks, KS)
                    z0 = 8 * kk
                    a, zz = [], []
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        za, zb = z0 + t, z0 + t + 4
                        zac, zbc = min(za, G1 - 1), min(zb, G1 - 1)
                        av = []
                        for q in range(4):
                            h = q & 1
                            zc, valid = (zbc, zb < G1) if (q >> 1) else (zac, za < G1)
                            av.append(gs[(r * G1 + zc) * C + g + 8 * h] if valid else 0.0)
                        dbs[warp, lane, 0] += av[0] + av[2]; dbs[warp, lane, 1] += av[1] + av[3]
                        a.append(tuple(av)); zz.append((zac, zbc))
                    for j in range(4):
                        bfr = []
                        for lane in range(32):
                            g = lane >> 2
                            tc = min(8 * j + g, TAPS - 1)
                            boff = (tc // 9) * TL + ((tc // 3) % 3) * G + tc % 3
                            zac, zbc = zz[lane]
                            bfr.append((ts[(2 * r) * G + 2 * zac + boff], ts[(2 * r) * G + 2 * zbc + boff]))
                        mma(acc[warp, j], a, bfr)
        for warp in range(8):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for j in range(4):
                    for e in range(4):
                        co, tap = g + 8 * (e >> 1), 8 * j + 2 * t + (e & 1)
                        if tap < TAPS:
                            rec[blk, co * TAPS + tap] += acc[warp, j, lane, e]
                if t == 0:
                    rec[blk, C * TAPS + g] += sum(dbs[warp, 4 * g + tt, 0] for tt in range(4))
                    rec[blk, C * TAPS + 8 + g] += sum(dbs[warp, 4 * g + tt, 1] for tt in range(4))
    tot = rec.sum(0)
    return tot[:C * TAPS], tot[C * TAPS:]


def check_conv1_wgrad(G, B=1):
    G1 = (G - 3) // 2 + 1
    x = torch.randint(-1, 2, (B, 1, G, G, G)).double()
    wt = torch.randn(C, 1, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    bias = torch.zeros(C, dtype=torch.float64, requires_grad=True)
    out = torch.nn.functional.conv3d(x, wt, bias, stride=2)
    dy = torch.randn_like(out)
    (out * dy).sum().backward()
    dy_cl = dy.permute(0, 2, 3, 4, 1).reshape(B, -1, C).numpy()
    dW, db = conv1_wgrad(x[:, 0].numpy(), dy_cl, G, G1, B)
    werr, berr = np.abs(dW - wt.grad.numpy().reshape(-1)).max(), np.abs(db - bias.grad.numpy()).max()
    print("G", G, "conv1 wgrad max err", werr, "db err", berr)
    assert werr < 1e-10 and berr < 1e-10


def pad_mod32(n, r):
    v = (n // 32) * 32 + r
    return v if v >= n else v + 32


def wgrad_staged(y1, sc, sh, dy2cl, G1, G2, B, nblocks):
    """conv2_wgrad_staged_kernel: groups of 4 output rows staged as 27 y1 lines + 4 dy2 lines; slot t <-> (row t, z),
    slot t+4 <-> (row t, z+1); n-tiles dealt round-robin to the 18 consumer warps (3 each)."""
    P1, P2 = G1 ** 3, G2 ** 3
    LP, DP = pad_mod32(G1 * C, 4), pad_mod32(G2 * C, 8)
    NYG = -(-G2 // 4)
    total = B * G2 * NYG
    gpb = -(-total // nblocks)
    nblk = -(-total // gpb)
    NW = 18
    NTW = -(-2 * NT // NW)
    rec = np.zeros((nblk, C * C * NT + C))
    nzp = (G2 + 1) // 2
    for blk in range(nblk):
        acc = np.zeros((NW, NTW, 32, 4))
        db_lo, db_hi = np.zeros(32), np.zeros(32)
        for grp in range(blk * gpb, min(total, blk * gpb + gpb)):
            b, rem = divmod(grp, G2 * NYG)
            x2, yg = divmod(rem, NYG)
            y20 = yg * 4
            nrows = min(4, G2 - y20)
            xs = np.full(27 * LP + 4 * DP, np.nan)                       # unwritten smem must never be read
            for i in range(3):
                for yl in range(2 * nrows + 1):
                    src = (b * P1 + ((2 * x2 + i) * G1 + (2 * y20 + yl)) * G1) * C
                    xs[(i * 9 + yl) * LP:(i * 9 + yl) * LP + G1 * C] = y1[src:src + G1 * C]
            for r in range(nrows):
                src = (b * P2 + (x2 * G2 + y20 + r) * G2) * C
                xs[27 * LP + r * DP:27 * LP + r * DP + G2 * C] = dy2cl[src:src + G2 * C]
            for zp in range(nzp):
                za, zb = 2 * zp, 2 * zp + 1
                vb, zbc = zb < G2, min(zb, G2 - 1)
                a = []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    vrow, tc = t < nrows, min(t, nrows - 1)
                    da = 27 * LP + tc * DP + za * C + g
                    dbb = 27 * LP + tc * DP + zbc * C + g
                    a0 = xs[da] if vrow else 0.0
                    a1 = xs[da + 8] if vrow else 0.0
                    a2 = xs[dbb] if (vrow and vb) else 0.0
                    a3 = xs[dbb + 8] if (vrow and vb) else 0.0
                    db_lo[lane] += a0 + a2; db_hi[lane] += a1 + a3
                    a.append((a0, a1, a2, a3))
                for warp in range(NW):
                    hf = warp & 1
                    for k in range(NTW):
                        nt = warp + NW * k
                        if nt >= 2 * NT: continue
                        tap = nt >> 1
                        i, r9 = divmod(tap, 9)
                        jy, l = divmod(r9, 3)
                        bfr = []
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            tc = min(t, nrows - 1)
                            line = (i * 9 + 2 * tc + jy) * LP + l * C + 8 * hf + g
                            ci = 8 * hf + g
                            x0 = max(sc[ci] * xs[line + 2 * za * C] + sh[ci], 0.0)
                            x1 = max(sc[ci] * xs[line + 2 * zbc * C] + sh[ci], 0.0)
                            assert not (np.isnan(x0) or np.isnan(x1))
                            bfr.append((x0, x1))
                        mma(acc[warp, k], a, bfr)
        for warp in range(NW):
            hf = warp & 1
            for k in range(NTW):
                nt = warp + NW * k
                if nt >= 2 * NT: continue
                tap = nt >> 1
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for e in range(4):
                        co, ci = g + 8 * (e >> 1), 8 * hf + 2 * t + (e & 1)
                        rec[blk, (co * C + ci) * NT + tap] = acc[warp, k, lane, e]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            if t == 0:
                rec[blk, C * C * NT + g] = sum(db_lo[4 * g + tt] for tt in range(4))
                rec[blk, C * C * NT + 8 + g] = sum(db_hi[4 * g + tt] for tt in range(4))
    tot = rec.sum(0)
    return tot[:C * C * NT], tot[C * C * NT:]

def check(G1, B=1):
    G2 = (G1 - 3) // 2 + 1
    x = torch.randn(B, C, G1, G1, G1, dtype=torch.float64)           # pre-BN y1, NCDHW
    sc, sh = torch.rand(C, dtype=torch.float64) + 0.5, torch.randn(C, dtype=torch.float64) * 0.3
    wt = torch.randn(C, C, 3, 3, 3, dtype=torch.float64) * 0.1
    bias = torch.randn(C, dtype=torch.float64)
    act = torch.relu(x * sc.view(1, C, 1, 1, 1) + sh.view(1, C, 1, 1, 1)).requires_grad_(True)
    y2_ref = torch.nn.functional.conv3d(act, wt, bias, stride=2)
    assert y2_ref.shape[-1] == G2
    y1_cl = x.permute(0, 2, 3, 4, 1).contiguous().numpy().reshape(-1)
    y2 = fwd(y1_cl, sc.numpy(), sh.numpy(), wt.numpy().reshape(-1), bias.numpy(), G1, G2, B)
    err = np.abs(y2 - y2_ref.detach().numpy().reshape(B, C, -1)).max()
    print("G1", G1, "fwd max err", err)
    assert err < 1e-12
    dy2 = torch.randn_like(y2_ref)
    (y2_ref * dy2).sum().backward()
    dact_ref = act.grad.permute(0, 2, 3, 4, 1).reshape(B, -1, C).numpy()
    dy2cl = dy2.permute(0, 2, 3, 4, 1).contiguous().numpy().reshape(-1)
    got = dgrad(dy2cl, wt.numpy().reshape(-1), G1, G2, B)
    derr = np.abs(got - dact_ref).max()
    print("G1", G1, "dgrad max err", derr, "unwritten", np.isnan(got).sum())
    assert derr < 1e-12 and not np.isnan(got).any()
    wt_ref = torch.randn(C, C, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    b_ref = torch.zeros(C, dtype=torch.float64, requires_grad=True)
    (torch.nn.functional.conv3d(act.detach(), wt_ref, b_ref, stride=2) * dy2).sum().backward()
    dW, db = wgrad(y1_cl, sc.numpy(), sh.numpy(), dy2cl, G1, G2, B, nblocks=3)
    werr = np.abs(dW - wt_ref.grad.numpy().reshape(-1)).max()
    berr = np.abs(db - b_ref.grad.numpy()).max()
    print("G1", G1, "wgrad max err", werr, "db err", berr)
    assert werr < 1e-11 and berr < 1e-11
    dW, db = wgrad_staged(y1_cl, sc.numpy(), sh.numpy(), dy2cl, G1, G2, B, nblocks=3)
    werr = np.abs(dW - wt_ref.grad.numpy().reshape(-1)).max()
    berr = np.abs(db - b_ref.grad.numpy()).max()
    print("G1", G1, "staged wgrad max err", werr, "db err", berr)
    assert werr < 1e-11 and berr < 1e-11


if __name__ == "__main__":
    for G1 in ([int(a) for a in sys.argv[1:]] or [9, 10, 15]):
        check(G1, B=2 if G1 < 12 else 1)
    check_conv1(20, B=2)
    check_conv1(36, B=1)
    check_conv1_wgrad(20, B=2)
    check_conv1_wgrad(36, B=1)
