"""numpy twin of the C oracle -- an independent restatement used to cross-check it.

TEST INFRASTRUCTURE ONLY.  Vectorised over t; taps are visited in ascending order with
a separate multiply and add per tap (numpy ufuncs never fuse), so results are
bit-identical to the reference's Java loops.  Works on the
sparse (a trous) form: zero taps of the dense upsampled filter add +-0.0 only.

Cites: CORE/internal/ScalarOps.java:700-723,790-808,818-835,909-916;
CORE/util/MathUtils.java:30-51; CORE/modwt/MODWTTransform.java:139-175,244-296,672-684;
CORE/modwt/MultiLevelMODWTTransform.java:244-251,554-645,795-806;
CORE/modwt/SymmetricAlignmentStrategy.java:43-117;
CORE/modwt/MutableMultiLevelMODWTResult.java:97-118;
CORE/swt/VectorWaveSwtAdapter.java:505-520,627-645.
"""
import math

import numpy as np

PERIODIC, ZERO_PADDING, SYMMETRIC = 0, 1, 2
S = 1.0 / math.sqrt(2.0)


def mirror(idx, n):
    m = np.mod(idx, 2 * n)
    return np.where(m < n, m, 2 * n - 1 - m)


def _gather(x, idx, mode):
    """x[ext(idx)] and a validity mask (False => term skipped, ZERO_PADDING only)."""
    n = x.shape[-1]
    if mode == PERIODIC:
        return x[..., np.mod(idx, n)], None
    if mode == SYMMETRIC:
        return x[..., mirror(idx, n)], None
    valid = (idx >= 0) & (idx < n)
    return x[..., np.clip(idx, 0, n - 1)], valid


def conv(x, f, mode, dilation=1):
    """out[t] = sum_k x[ext(t - k*dilation)] * f[k]; x may be [..., N]."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    t = np.arange(n)
    acc = np.zeros_like(x)
    for k in range(len(f)):
        xv, valid = _gather(x, t - k * dilation, mode)
        term = xv * f[k]
        acc = acc + term if valid is None else np.where(valid, acc + term, acc)
    return acc


def forward_single(x, h, g, mode):
    hs, gs = np.asarray(h) * S, np.asarray(g) * S
    return conv(x, hs, mode), conv(x, gs, mode)


def inverse_single(v, w, hr, gr, mode, batch_variant=False):
    v, w = np.asarray(v, dtype=np.float64), np.asarray(w, dtype=np.float64)
    hs, gs = np.asarray(hr) * S, np.asarray(gr) * S
    n = v.shape[-1]
    t = np.arange(n)
    acc = np.zeros_like(v)
    for k in range(len(hs)):
        if mode == SYMMETRIC and not batch_variant:
            idx = t - k
        else:
            idx = t + k
        vv, valid = _gather(v, idx, mode)
        wv, _ = _gather(w, idx, mode)
        term = hs[k] * vv + gs[k] * wv
        acc = acc + term if valid is None else np.where(valid, acc + term, acc)
    return acc


def decompose(x, h, g, levels, mode):
    hs, gs = np.asarray(h) * S, np.asarray(g) * S
    cur = np.asarray(x, dtype=np.float64)
    ws = []
    for level in range(1, levels + 1):
        d = 1 << (level - 1)
        if (len(hs) - 1) * d + 1 > cur.shape[-1]:
            raise ValueError("VAL_TOO_LARGE")
        ws.append(conv(cur, gs, mode, d))
        cur = conv(cur, hs, mode, d)
    return np.stack(ws, axis=0), cur


def alignment(wavelet_id, l0, level):
    detail_plus = True
    if l0 <= 2:
        return True, (0 if level <= 1 else -1), True, 0
    approx_plus = False
    if wavelet_id == 1:
        dh, dg = (0 if level <= 1 else -1), (1 if level >= 3 else 0)
    elif wavelet_id == 2:
        dh, dg = (0 if level <= 1 else 1), (1 if level >= 2 else 0)
    elif wavelet_id == 3:
        approx_plus, detail_plus, dh, dg = True, False, 0, 0
    elif wavelet_id == 4:
        dh, dg = (0, 0) if level <= 1 else ((1, 0) if level == 2 else (1, 1))
    elif wavelet_id == 5:
        approx_plus, detail_plus, dh, dg = True, False, (0 if level <= 1 else 1), 0
    elif wavelet_id == 6:
        detail_plus = False
        dh, dg = (0, 0) if level <= 1 else (-1, 1)
    elif l0 >= 12:
        if level <= 1:
            dh, dg = 0, 0
        else:
            dh = dg = 0 if level % 2 == 0 else -1
    else:
        dh, dg = (0, 0) if level <= 1 else (-1, 0)
    return approx_plus, dh, detail_plus, dg


def tau_j(l0, level):
    return ((l0 - 1) * (1 << (level - 1))) // 2


def synth_level(a, dcoef, hr, gr, level, mode, wavelet_id=0):
    a, dcoef = np.asarray(a, dtype=np.float64), np.asarray(dcoef, dtype=np.float64)
    hs, gs = np.asarray(hr) * S, np.asarray(gr) * S
    n = a.shape[-1]
    d = 1 << (level - 1)
    t = np.arange(n)
    acc = np.zeros_like(a)
    if mode == PERIODIC:
        for k in range(len(hs)):
            acc = acc + hs[k] * a[..., np.mod(t + k * d, n)]
        for k in range(len(gs)):
            acc = acc + gs[k] * dcoef[..., np.mod(t + k * d, n)]
    elif mode == ZERO_PADDING:
        for k in range(len(hs)):
            idx = t + k * d
            valid = idx < n
            idc = np.clip(idx, 0, n - 1)
            acc = np.where(valid, acc + (hs[k] * a[..., idc] + gs[k] * dcoef[..., idc]), acc)
    else:
        ap, dh, dp, dg = alignment(wavelet_id, len(hs), level)
        th, tg = tau_j(len(hs), level) + dh, tau_j(len(gs), level) + dg
        for k in range(len(hs)):
            idx = t + k * d - th if ap else t - k * d + th
            acc = acc + hs[k] * a[..., mirror(idx, n)]
        for k in range(len(gs)):
            idx = t + k * d - tg if dp else t - k * d + tg
            acc = acc + gs[k] * dcoef[..., mirror(idx, n)]
    return acc


def reconstruct(w, v, hr, gr, mode, wavelet_id=0, detail_mask=None, use_approx=True):
    levels = w.shape[0]
    if detail_mask is None:
        detail_mask = (1 << levels) - 1
    cur = np.array(v, dtype=np.float64) if use_approx else np.zeros_like(v)
    for level in range(levels, 0, -1):
        dj = w[level - 1] if (detail_mask >> (level - 1)) & 1 else np.zeros_like(v)
        cur = synth_level(cur, dj, hr, gr, level, mode, wavelet_id)
    return cur


def threshold(c, thr, soft):
    c = np.asarray(c, dtype=np.float64)
    a = np.abs(c)
    if soft:
        return np.where(a > thr, np.sign(c) * (a - thr), 0.0)
    return np.where(a <= thr, 0.0, c)


def universal_threshold(w1):
    a = np.sort(np.abs(np.asarray(w1, dtype=np.float64)))
    n = a.size
    med = (a[n // 2 - 1] + a[n // 2]) / 2.0 if n % 2 == 0 else a[n // 2]
    return (med / 0.6745) * math.sqrt(2 * math.log(n))


# ---- streaming (blockwise) batch analysis with carried history ------------------------------------------------------
class StreamingOracle:
    """Restates EXT/extensions/modwt/BatchStreamingMODWT.java:55-163,183-276 (+ history kernel
    BatchSIMDMODWT.java:447-507) for ZERO_PADDING / SYMMETRIC: per level a left history of L_j - 1 samples of the
    level's input; first block: zeros / symmetricBoundaryExtension of the block itself (:326-334); afterwards the
    last L_j - 1 samples of [history | block] (:336-350).  Blocks are [B][n] (AoS)."""

    def __init__(self, h, g, levels, mode):
        self.levels, self.mode = levels, mode
        self.low = [np.zeros((len(h) - 1) * (1 << j) + 1) for j in range(levels)]
        self.high = [np.zeros((len(h) - 1) * (1 << j) + 1) for j in range(levels)]
        for j in range(levels):   # ScalarOps.upsampleAndScaleForIMODWTSynthesis :909-916
            self.low[j][::1 << j] = np.asarray(h) * S
            self.high[j][::1 << j] = np.asarray(g) * S
        self.hist = [None] * levels

    def _level(self, j, cur, update=True):
        b, n = cur.shape
        hl = self.low[j].size - 1
        if self.hist[j] is None:
            if self.mode == 1:
                self.hist[j] = np.zeros((b, hl))
            else:
                self.hist[j] = cur[:, mirror(np.arange(hl) - hl, n)]
        ext = np.concatenate([self.hist[j], cur], axis=1)
        a = np.zeros((b, n))
        d = np.zeros((b, n))
        for l in range(hl + 1):    # ascending taps, separate multiply and add, zeros of the dense filter included
            s = ext[:, hl - l:hl - l + n]
            a = a + s * self.low[j][l]
            d = d + s * self.high[j][l]
        if update:
            self.hist[j] = ext[:, n:n + hl].copy()
        return a, d

    def process(self, block, update=True):
        cur = np.asarray(block, dtype=np.float64)
        ws = []
        for j in range(self.levels):
            cur, d = self._level(j, cur, update)
            ws.append(d)
        return np.stack(ws), cur

    def flush(self, tail_length):
        h0 = self.hist[0]
        tail = np.zeros((h0.shape[0], tail_length)) if self.mode == 1 else h0[:, ::-1][:, :tail_length]
        return self.process(tail, update=False)


# ---- WaveletDenoiser (CORE/denoising/WaveletDenoiser.java) ------------------------------------------------------------
def _seq_sum(a):
    """Java's `for (double c : coeffs) s += c` -- sequential, not numpy's pairwise summation"""
    s = 0.0
    for c in np.asarray(a, dtype=np.float64):
        s += float(c)
    return s


def denoiser_threshold(coeffs, sigma, method):
    """calculateThreshold (:391-436) for UNIVERSAL / SURE / MINIMAX / BAYES"""
    import math
    n = len(coeffs)
    if method == "UNIVERSAL":
        return sigma * math.sqrt(2.0 * math.log(n))
    if method == "MINIMAX":                                 # :497-509
        log_n = math.log(n)
        if n <= 32:
            return 0.0
        if n <= 64:
            return sigma * 0.3936 + 0.1829 * sigma * log_n
        return sigma * (0.4745 + 0.1148 * log_n)
    if method == "BAYES":                                   # :521-552
        sigma2 = sigma * sigma
        mean = _seq_sum(coeffs) / n
        variance = _seq_sum((np.asarray(coeffs) - mean) ** 2) / n
        return sigma2 / math.sqrt(max(0.0, variance - sigma2) + 1e-10)
    if method == "SURE":
        return sure_threshold(coeffs, sigma)[0]
    raise ValueError(method)


def sure_threshold(coeffs, sigma):
    """calculateSUREThreshold + calculateSURERisk (:441-492): every candidate t = sorted |c|[k] scored with a sequential
    pass over the coefficients in their stored order (vectorised over the candidates, so each risk sees exactly the
    reference's additions in the reference's order); first strict minimum in ascending k; capped at the universal
    threshold.  Returns (threshold, minimal risk)."""
    import math
    c = np.asarray(coeffs, dtype=np.float64)
    n = c.size
    t = np.sort(np.abs(c))                                  # :445-449
    sigma2 = sigma * sigma
    risk = np.full(n, float(-n) * sigma2)                   # :480
    with np.errstate(invalid="ignore", over="ignore"):
        for cj in c:                                        # :482-489
            a = abs(cj)
            d = a - t
            risk = risk + np.where(a <= t, cj * cj, sigma2 + d * d)
        risk = risk / n                                     # :491
    min_risk, best = math.inf, 0.0                          # :452-453
    ok = risk < math.inf
    if ok.any():
        min_risk = float(risk[ok].min())
        best = float(t[np.flatnonzero(risk == min_risk)[0]])   # `risk < minRisk` keeps the first
    universal = sigma * math.sqrt(2.0 * math.log(n))        # :465-469
    return (universal if best > universal else best), min_risk


def denoiser_sigma(detail):
    """estimateNoiseSigma (:376-387) + calculateMedian (:587-597)"""
    a = np.sort(np.abs(np.asarray(detail, dtype=np.float64)))
    n = a.size
    med = (a[n // 2 - 1] + a[n // 2]) / 2.0 if n % 2 == 0 else a[n // 2]
    return med / 0.6745


def denoiser_single(x, h, g, mode, method, soft, fixed=None):
    """denoise (:124-145) / denoiseFixed (:354-366): single level, MODWTTransform.inverse (pair-added, SYMMETRIC t-l)"""
    v, w = forward_single(x, h, g, mode)
    thr = fixed if fixed is not None else denoiser_threshold(w, denoiser_sigma(w), method)
    return inverse_single(v, threshold(w, thr, soft), h, g, mode), thr


def denoiser_multilevel(x, h, g, levels, mode, wavelet_id, method, soft):
    """denoiseMultiLevel (:155-171): sigma from level 1; level j thresholded with calculateThreshold(W_j, sigma/sqrt(2^j))"""
    import math
    w, v = decompose(x, h, g, levels, mode)
    sigma = denoiser_sigma(w[0])
    thrs = []
    wd = np.empty_like(w)
    for j in range(levels):
        thr = denoiser_threshold(w[j], sigma / math.sqrt(1 << (j + 1)), method)
        thrs.append(thr)
        wd[j] = threshold(w[j], thr, soft)
    return reconstruct(wd, v, h, g, mode, wavelet_id), thrs


class CoreStreamingOracle:
    """MODWTStreamingTransformImpl (:139-256) / MultiLevelMODWTStreamingTransform (:117-218) restated sample by sample:
    the ring, the `samplesInBuffer >= bufferSize` trigger, the last-bufferSize-samples window, the L-1 overlap kept after
    a single-level window, the zero-padded flush.  `windows` collects every window handed to the transform."""

    def __init__(self, buffer_size, filter_length, multilevel):
        self.bs, self.multilevel = buffer_size, multilevel
        self.overlap = 0 if multilevel else filter_length - 1
        self.ring = np.zeros(buffer_size + self.overlap)
        self.write = 0
        self.count = 0
        self.windows = []

    def process(self, data):
        for sample in data:
            self.ring[self.write] = sample
            self.write = (self.write + 1) % self.ring.size
            self.count += 1
            if self.count >= self.bs:
                read = (self.write - self.bs + self.ring.size) % self.ring.size
                self.windows.append(np.array([self.ring[(read + i) % self.ring.size] for i in range(self.bs)]))
                self.count -= self.bs - self.overlap

    def flush(self):
        if self.count > 0:
            final = np.zeros(self.bs)
            read = (self.write - self.count + self.ring.size) % self.ring.size
            for i in range(self.count):
                final[i] = self.ring[(read + i) % self.ring.size]
            self.windows.append(final)
            self.count = 0
            self.write = 0
