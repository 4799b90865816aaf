"""CPU check of the ALGORITHM behind the third-generation Canny kernels (eel_unet_b200/csrc/edge.cu: canny_rows_kernel,
canny_flood_sweep_kernel): a lane-level numpy transcription of the two kernels -- same band / lane / register layout, same sign-bit
masks, same column-strip sweeps with Jacobi exchange between threads -- must reproduce the oracle (= cv2.Canny) bit for bit.
The kernels themselves are compared with cv2 on the GPU in tests/test_edges_gpu.py; this file pins the design (border handling,
band overlap, mask algebra, flood-fill fixpoint) where no GPU is needed.  Test infrastructure only."""
import numpy as np
import pytest

from oracle import edge_np, synth

M32 = 0xFFFFFFFF


def _shfl_up(v):
    r = v.copy(); r[1:] = v[:-1]; return r


def _shfl_down(v):
    r = v.copy(); r[:-1] = v[1:]; return r


def rows_kernel(gray, BR, low, high):
    """canny_rows_kernel: one 'warp' (32 lanes x 16 pixels) per band of BR rows -> weak / strong bitmaps as uint16 per lane."""
    gray = gray.astype(np.int64)
    H, W = gray.shape
    nl = W >> 4
    lanes = np.arange(32)
    act = lanes < nl
    first, last = lanes == 0, lanes == nl - 1
    weak = np.zeros((H, W // 16), np.uint16)
    strong = np.zeros((H, W // 16), np.uint16)
    Z = lambda: np.zeros(32, np.int64)
    push = lambda mask, diff: (mask << 1) | (diff < 0)          # __funnelshift_l(diff, mask, 1): sign bit in at bit 0
    for band in range(-(-H // BR)):
        y0 = band * BR
        g = {}
        MB = np.zeros((32, 18), np.int64)
        mB = dict(cand=Z(), strong=Z(), hor=Z(), ver=Z(), neg=Z(), gt_l=Z(), gt_u=Z(), gt_ul=Z(), gt_ur=Z())
        for t in range(BR + 4):
            yy = min(max(y0 - 2 + t, 0), H - 1)
            gC = np.stack([gray[yy, (l if act[l] else 0) * 16:(l if act[l] else 0) * 16 + 16] for l in range(32)])
            g[t] = gC
            if t < 2:
                continue
            gA, gB = g[t - 2], g[t - 1]
            s = np.zeros((32, 18), np.int64); d = np.zeros((32, 18), np.int64)
            s[:, 1:17] = gA + 2 * gB + gC
            d[:, 1:17] = gC - gA
            s[:, 0] = np.where(first, s[:, 1], _shfl_up(s[:, 16])); d[:, 0] = np.where(first, d[:, 1], _shfl_up(d[:, 16]))
            s[:, 17] = np.where(last, s[:, 16], _shfl_down(s[:, 1])); d[:, 17] = np.where(last, d[:, 16], _shfl_down(d[:, 1]))
            ry = y0 + t - 3
            rowmask = -1 if 0 <= ry < H else 0
            MC = np.zeros((32, 18), np.int64)
            cand, strong_, hor, ver, neg = Z(), Z(), Z(), Z(), Z()
            for j in range(15, -1, -1):
                dx = s[:, j + 2] - s[:, j]
                dy = d[:, j] + 2 * d[:, j + 1] + d[:, j + 2]
                ax, ay = np.abs(dx), np.abs(dy)
                m = (ax + ay) & rowmask
                MC[:, j + 1] = m
                dh = (ay << 15) - ax * 13573
                dv = (ax << 16) - dh
                cand = push(cand, low - m); strong_ = push(strong_, high - m)
                hor = push(hor, dh); ver = push(ver, dv); neg = push(neg, dx ^ dy)
            MC[:, 0] = np.where(first, 0, _shfl_up(MC[:, 16])); MC[:, 17] = np.where(last, 0, _shfl_down(MC[:, 1]))
            gt_l, gt_u, gt_ul, gt_ur, gt_dl, gt_dr = Z(), Z(), Z(), Z(), Z(), Z()
            for j in range(16, -1, -1):
                gt_l = push(gt_l, MC[:, j] - MC[:, j + 1])
            for j in range(15, -1, -1):
                gt_u = push(gt_u, MB[:, j + 1] - MC[:, j + 1])
                gt_ul = push(gt_ul, MB[:, j] - MC[:, j + 1])
                gt_ur = push(gt_ur, MB[:, j + 2] - MC[:, j + 1])
                gt_dl = push(gt_dl, MC[:, j] - MB[:, j + 1])
                gt_dr = push(gt_dr, MC[:, j + 2] - MB[:, j + 1])
            gt_d = gt_u
            mC = dict(cand=cand, strong=strong_, hor=hor, ver=ver & ~hor, neg=neg, gt_l=gt_l, gt_u=gt_u, gt_ul=gt_ul, gt_ur=gt_ur)
            if t >= 4 and y0 + t - 4 < H:
                keep_h = mB["gt_l"] & ~(mB["gt_l"] >> 1)
                keep_v = mB["gt_u"] & ~gt_d
                keep_d2 = mB["gt_ul"] & gt_dr
                keep_d3 = mB["gt_ur"] & gt_dl
                diag = ~(mB["hor"] | mB["ver"])
                wk = mB["cand"] & 0xFFFF & ((mB["hor"] & keep_h) | (mB["ver"] & keep_v) |
                                            (diag & ((~mB["neg"] & keep_d2) | (mB["neg"] & keep_d3))))
                st = wk & mB["strong"]
                weak[y0 + t - 4, :nl] = wk[:nl]
                strong[y0 + t - 4, :nl] = st[:nl]
            MB, mB = MC, mC
    return weak, strong


def _brev(x):
    return int("{:032b}".format(x)[::-1], 2)


def _fill(w, sd):
    """a seed runs through its run of candidates inside the word, both directions (the carry trick of the kernels)"""
    f = ((w & ~((w + sd) & M32)) | sd) & M32
    rw, rs = _brev(w), _brev(f)
    return _brev(((rw & ~((rw + rs) & M32)) | rs) & M32)


def flood_sweep(weak16, strong16, H, W, R=16):
    """canny_flood_sweep_kernel: thread (g, k) = rows [16 g, 16 g + 16) of word column k; returns (edge map, iterations)."""
    WW, K = W // 32, 16
    words = lambda a: a.reshape(H, WW, 2)[..., 0].astype(np.uint64) | (a.reshape(H, WW, 2)[..., 1].astype(np.uint64) << np.uint64(16))
    Wb, Sb = words(weak16), words(strong16)
    nseg = -(-H // R)
    w = [[[int(Wb[g * R + r, k]) if g * R + r < H and k < WW else 0 for r in range(R)] for k in range(K)] for g in range(nseg)]
    s = [[[int(Sb[g * R + r, k]) if g * R + r < H and k < WW else 0 for r in range(R)] for k in range(K)] for g in range(nseg)]

    def published():
        T = [[s[g][k][0] for k in range(K)] for g in range(nseg)]
        B = [[s[g][k][R - 1] for k in range(K)] for g in range(nseg)]
        Mm = [[sum(((s[g][k][r] >> 31) & 1) << r for r in range(R)) for k in range(K)] for g in range(nseg)]
        Lm = [[sum((s[g][k][r] & 1) << r for r in range(R)) for k in range(K)] for g in range(nseg)]
        return T, B, Mm, Lm

    its = 0
    while True:
        its += 1
        T, B, Mm, Lm = published()                  # what the other threads published at the end of the previous iteration
        changed = False
        for g in range(nseg):
            for k in range(K):
                hasU, hasD, hasL, hasR = g > 0, g + 1 < nseg, k > 0, k < K - 1
                up = B[g - 1][k] if hasU else 0
                dn = T[g + 1][k] if hasD else 0
                upL = (B[g - 1][k - 1] >> 31) if hasU and hasL else 0
                upR = (B[g - 1][k + 1] & 1) if hasU and hasR else 0
                dnL = (T[g + 1][k - 1] >> 31) if hasD and hasL else 0
                dnR = (T[g + 1][k + 1] & 1) if hasD and hasR else 0
                mLx = ((Mm[g][k - 1] if hasL else 0) << 1) | upL | (dnL << (R + 1))
                mRx = ((Lm[g][k + 1] if hasR else 0) << 1) | upR | (dnR << (R + 1))
                cL = mLx | (mLx >> 1) | (mLx >> 2)
                cR = mRx | (mRx >> 1) | (mRx >> 2)
                sk, wk = s[g][k], w[g][k]
                for sweep in (range(R), range(R - 1, -1, -1)):
                    for r in sweep:
                        above = up if r == 0 else sk[r - 1]
                        below = dn if r == R - 1 else sk[r + 1]
                        nb = above | sk[r] | below
                        n3 = (nb | (nb << 1) | (nb >> 1) | ((cL >> r) & 1) | (((cR >> r) & 1) << 31)) & M32
                        nf = _fill(wk[r], sk[r] | (wk[r] & n3))
                        changed |= nf != sk[r]
                        sk[r] = nf
        if not changed:
            break
    out = np.zeros((H, W), np.uint8)
    for g in range(nseg):
        for k in range(WW):
            for r in range(R):
                if g * R + r < H:
                    bits = s[g][k][r]
                    out[g * R + r, k * 32:(k + 1) * 32] = [255 if (bits >> b) & 1 else 0 for b in range(32)]
    return out, its


def _snake(h, w):
    g = np.zeros((h, w), np.uint8)
    y0, x0, y1, x1 = 2, 2, h - 3, w - 3
    while y1 - y0 > 8 and x1 - x0 > 8:
        g[y0, x0:x1] = 30; g[y0:y1, x1] = 30; g[y1, x0 + 4:x1 + 1] = 30; g[y0 + 4:y1 + 1, x0 + 4] = 30
        y0 += 4; x0 += 4; y1 -= 4; x1 -= 4
    g[2, 2:6] = 255
    return g


@pytest.mark.parametrize("h,w,br", [(64, 64, 8), (37, 96, 16), (100, 128, 32), (33, 32, 8), (5, 512, 8), (1, 32, 8), (70, 256, 16)])
def test_row_bands_and_strip_sweeps_reproduce_the_oracle(h, w, br):
    rng = np.random.default_rng(h * 1000 + w)
    imgs, _ = synth.tooth_images(1, h, w, seed=h + w)
    for gray in (edge_np.gray_u8(imgs[0]), rng.integers(0, 256, size=(h, w), dtype=np.uint8)):
        for low, high in ((100, 200), (30, 400)):
            wk, st = rows_kernel(gray, br, low, high)
            nms = edge_np.canny_nms(gray, low, high)          # 0 weak candidate, 1 none, 2 strong
            px = lambda a: np.unpackbits(a.view(np.uint8), bitorder="little").reshape(h, w).astype(bool)
            assert np.array_equal(px(wk), nms != 1) and np.array_equal(px(st), nms == 2)
            out, _ = flood_sweep(wk, st, h, w)
            assert np.array_equal(out, edge_np.canny(gray, low, high))


def test_flood_follows_a_long_weak_chain_across_strips():
    """a one-pixel spiral of WEAK candidates hanging on one strong run: the sweeps must follow it through every strip border"""
    g = _snake(96, 128)
    nms = edge_np.canny_nms(g, 100, 200)
    assert (nms == 0).sum() > 100 * (nms == 2).sum() > 0           # the chain really is weak: magnitude 4 x 30 (corners 6 x 30) < 200
    wk, st = rows_kernel(g, 16, 100, 200)
    out, its = flood_sweep(wk, st, 96, 128)
    assert np.array_equal(out, edge_np.canny(g, 100, 200))
    assert its > 30                                                # it did have to cross strip borders, one per iteration
