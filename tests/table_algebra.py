"""numpy model of the table path's algebra (umpa_b200/csrc/table_path.cu), any float dtype.

cost_tables(sam, ref, Nw, max_shift, df, dtype) returns cost/T/df for every pixel of the valid
region and every integer shift, computed the way the CUDA kernels do it: centred frames,
frame-summed unwindowed correlation + one separable window filter per shift, per-frame filtered
images for the mean term, FP64 solve.  With dtype=float64 it must agree with the oracle's direct
evaluation to rounding; with float32 it predicts the FP32 error of the kernels."""
import numpy as np


def _filt(img, g, dtype):
    """valid separable correlation with the 1-D factor g along the last two axes"""
    K = len(g)
    H, W = img.shape[-2:]
    tmp = np.zeros(img.shape[:-1] + (W - K + 1,), dtype=dtype)
    for v in range(K):
        tmp = (tmp + g[v] * img[..., :, v:v + W - K + 1]).astype(dtype)
    out = np.zeros(img.shape[:-2] + (H - K + 1, W - K + 1), dtype=dtype)
    for u in range(K):
        out = (out + g[u] * tmp[..., u:u + H - K + 1, :]).astype(dtype)
    return out


def cost_tables(sam, ref, Nw, max_shift, df=True, dtype=np.float32):
    sam, ref = np.asarray(sam, np.float64), np.asarray(ref, np.float64)
    Na, H, W = sam.shape
    h = max_shift - 1
    S = 2 * h + 1
    pad = max_shift + Nw
    ham = np.hamming(2 * Nw + 1)
    g64 = ham / ham.sum()
    g = g64.astype(dtype)
    sw = float(np.sum(g.astype(np.float64))) ** 2
    c = ref.mean(axis=(1, 2)); d = sam.mean(axis=(1, 2))
    Rc = (ref - c[:, None, None]).astype(dtype); Sc = (sam - d[:, None, None]).astype(dtype)
    cf, dfm = c.astype(dtype), d.astype(dtype)
    # filtered images live on [Nw, H-Nw) x [Nw, W-Nw): index (y - Nw, x - Nw)
    a = _filt(Rc, g, dtype); b = _filt(Sc, g, dtype)
    T3 = _filt(np.sum((Rc * Rc).astype(dtype), axis=0, dtype=dtype), g, dtype)
    T1 = _filt(np.sum((Sc * Sc).astype(dtype), axis=0, dtype=dtype), g, dtype)
    P3 = np.sum(cf[:, None, None] * a, axis=0, dtype=dtype); U = np.sum(dfm[:, None, None] * a, axis=0, dtype=dtype)
    P1 = np.sum(dfm[:, None, None] * b, axis=0, dtype=dtype); V = np.sum(cf[:, None, None] * b, axis=0, dtype=dtype)
    M2 = np.sum((a * a).astype(dtype), axis=0, dtype=dtype)
    cd, cc, dd = float(np.sum(c * d)), float(np.sum(c * c)), float(np.sum(d * d))
    N0, N1 = H - 2 * pad, W - 2 * pad
    cost = np.zeros((S, S, N0, N1)); Tm = np.zeros_like(cost); Dm = np.zeros_like(cost)
    f8 = np.float64
    o = pad - Nw                    # offset of output pixel 0 inside the filtered images
    t1 = T1[o:o + N0, o:o + N1].astype(f8) + 2 * P1[o:o + N0, o:o + N1].astype(f8) + sw * dd
    Vp = V[o:o + N0, o:o + N1].astype(f8)
    for si in range(-h, h + 1):
        for sj in range(-h, h + 1):
            # unwindowed frame correlation on the window-extended region, then ONE filter
            ys, xs = pad - Nw, pad - Nw
            A = Rc[:, ys + si:ys + si + N0 + 2 * Nw, xs + sj:xs + sj + N1 + 2 * Nw]
            B = Sc[:, ys:ys + N0 + 2 * Nw, xs:xs + N1 + 2 * Nw]
            C = np.zeros(A.shape[1:], dtype=dtype)
            for k in range(Na):
                C = (C + A[k] * B[k]).astype(dtype)
            X = _filt(C, g, dtype).astype(f8)
            sl = (slice(o + si, o + si + N0), slice(o + sj, o + sj + N1))
            t3 = T3[sl].astype(f8) + 2 * P3[sl].astype(f8) + sw * cc
            lin = U[sl].astype(f8) + Vp + sw * cd
            t5 = X + lin
            if df:
                Mt = np.zeros((N0, N1), dtype=dtype)
                for k in range(Na):
                    Mt = (Mt + a[k][sl] * b[k, o:o + N0, o:o + N1]).astype(dtype)
                t2 = M2[sl].astype(f8) / sw ** 2 + 2 * P3[sl].astype(f8) / sw + cc
                t6 = sw * t2
                t4 = Mt.astype(f8) / sw + lin
                den = t2 * t3 - t6 * t6
                Kc = (t2 * t5 - t4 * t6) / den
                beta = (t3 * t4 - t5 * t6) / den
                T = beta + Kc
                Tm[si + h, sj + h] = T; Dm[si + h, sj + h] = Kc / T
                cost[si + h, sj + h] = (t1 + beta * beta * t2 + Kc * Kc * t3 - 2 * beta * t4 - 2 * Kc * t5 + 2 * beta * Kc * t6) / Na
            else:
                T = t5 / t3
                Tm[si + h, sj + h] = T
                cost[si + h, sj + h] = (t1 - t5 * T) / Na
    return cost, Tm, Dm
