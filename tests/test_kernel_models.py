"""CPU models of two kernels' DECOMPOSITION, checked against the oracle (no GPU, no product code executed).

The CUDA kernels themselves are verified on the B200 (tests/test_iqcorr_gpu.py, tests/test_demod_sdriq.py); these numpy
restatements of their block / thread arithmetic pin the index algebra the kernels are built on, so that a change of tile
size, field layout or twiddle table can be checked here before it costs GPU time:
  * K7 dc_correct_kernel (sdrangel_b200/csrc/iqcorr_api.cu): 3072-sample tile + 1024-sample halo per block, 16 samples per
    thread, window sum of thread t at element k = totals of the 64 threads before t + sum_{m<=k} (x_t[m] - x_{t-64}[m])
    (reference: DSPDeviceSourceEngine::iqCorrections, dspdevicesourceengine.cpp:254-259; MovingAverageUtil, util/movingaverage.h);
  * K11 fftfilt_fast_kernel<LOG2N> (sdrangel_b200/csrc/fftfilt_api.cu): four radix-2 stages per pass on the 16 elements
    that differ in index bits [P, P+4), padded element index e + (e >> 4), per-stage twiddle tables, compile-time twiddles
    in field 0, the overlap kept per thread (reference: fftfilt::runSSB, fftfilt.cpp:285-325)."""
import numpy as np
import pytest

C64 = np.complex64
DC_N, DC_TILE, DC_THREADS, DC_PER = 1024, 3072, 256, 16


def dc_kernel_model(x, hist):
    """One launch of dc_correct_kernel on x (n, 2) int16 with the carried history (1024, 2) int16."""
    n = len(x)
    out = np.zeros_like(x)
    for b in range((n + DC_TILE - 1) // DC_TILE):
        t0 = b * DC_TILE
        raw = np.zeros((DC_THREADS, DC_PER, 2), np.int64)
        rems = np.zeros(DC_THREADS, int)
        for t in range(DC_THREADS):
            i0 = t0 - DC_N + DC_PER * t
            if i0 < 0:
                raw[t] = hist[DC_N + i0:DC_N + i0 + DC_PER]
                rems[t] = DC_PER
            else:
                rems[t] = min(DC_PER, max(0, n - i0))
                raw[t, :rems[t]] = x[i0:i0 + rems[t]]
        tot = raw.sum(1)
        incl = np.zeros_like(tot)
        wt = np.zeros((DC_THREADS // 32, 2), np.int64)
        for w in range(DC_THREADS // 32):
            c = np.cumsum(tot[32 * w:32 * w + 32], 0)
            incl[32 * w:32 * w + 32] = c
            wt[w] = c[-1]
        ex = incl - tot
        for t in range(64, DC_THREADS):
            if rems[t] <= 0:
                continue
            i0, w, pt = t0 - DC_N + DC_PER * t, t >> 5, t - 64
            S = wt[w - 2] + wt[w - 1] + ex[t] - ex[pt]
            for k in range(rems[t]):
                S = S + raw[t, k] - raw[pt, k]
                q = np.where(S >= 0, S // DC_N, -((-S) // DC_N))        # C++ division: toward zero
                out[i0 + k] = (raw[t, k] - q).astype(np.int16)
    return out, np.concatenate([hist, x])[-DC_N:]


def test_dc_correct_decomposition_matches_oracle(port):
    rs = np.random.RandomState(11)
    n = 40_001
    x = (rs.randint(-20000, 20000, size=(n, 2)) + np.array([5000, -7000])).clip(-32768, 32767).astype(np.int16)
    x[20_000:21_500] = -32768
    cuts = [0, 1, 1023, 1025, 9_000, 9_001, 30_000, n]
    o = port.PortIQCorrections()
    hist = np.zeros((DC_N, 2), np.int16)
    for a, b in zip(cuts[:-1], cuts[1:]):
        got, hist = dc_kernel_model(x[a:b], hist)
        want = o.run(x[a:b])
        assert np.array_equal(got, want), (a, b, int(np.argmax(np.any(got != want, axis=1))))


# ---------------------------------------------------------------------------------------------------- K11
W16 = [1, 0.92387953251128674 - 0.38268343236508977j, 0.70710678118654752 - 0.70710678118654752j, 0.38268343236508977 - 0.92387953251128674j,
       -1j, -0.38268343236508977 - 0.92387953251128674j, -0.70710678118654752 - 0.70710678118654752j, -0.92387953251128674 - 0.38268343236508977j]


def ff_idx(e):
    return e + (e >> 4)


def ff_stages(v, tws, low, inv, P, s_lo, s_hi):
    for bb in range(4):
        b = bb if inv else 3 - bb
        s = P + b
        if s < s_lo or s > s_hi:
            continue
        for m in range(16):
            if m & (1 << b):
                continue
            r = (m & ((1 << b) - 1))
            w = C64(W16[r << (3 - s)]) if P == 0 else tws[(1 << s) - 1 + low + (r << P)]
            a, c = v[m].copy(), v[m | (1 << b)].copy()
            if not inv:
                v[m] = a + c
                v[m | (1 << b)] = ((a - c) * w).astype(C64)
            else:
                u = (c * np.conj(w)).astype(C64)
                v[m] = a + u
                v[m | (1 << b)] = a - u


def ff_e0(G, P):
    return ((G >> P) << (P + 4)) | (G & ((1 << P) - 1))


def ff_load(x, G, P):
    return np.stack([x[ff_idx(ff_e0(G, P) + (m << P))] for m in range(16)])


def ff_store(x, G, P, v):
    for m in range(16):
        x[ff_idx(ff_e0(G, P) + (m << P))] = v[m]


def fftfilt_fast_model(log2n, stream, ovl_in, mult_br, tw, b0, b1):
    """Blocks [b0, b1) of one call as ONE CTA of fftfilt_fast_kernel<log2n> computes them (G = all threads at once)."""
    N = 1 << log2n
    N2, NT, TOP = N // 2, N // 16, log2n - 4
    f0_hi = 3 if log2n % 4 == 0 else log2n % 4 - 1
    sl = 4
    while log2n - sl > 4:
        sl += 4
    x = np.zeros(N + N // 16, C64)
    tws = np.zeros(N, C64)
    for s in range(log2n):
        for r in range(1 << s):
            tws[(1 << s) - 1 + r] = tw[r << (log2n - 1 - s)]
    G = np.arange(NT)
    ov = np.stack([ovl_in[m * NT + G] if b0 == 0 else np.zeros(NT, C64) for m in range(8)])
    inv = np.float32(1.0 / N)
    res = np.zeros((b1 - b0) * N2, C64)
    for b in range(0 if b0 == 0 else b0 - 1, b1):
        v = np.zeros((16, NT), C64)
        for m in range(8):
            v[m] = stream[b * N2 + m * NT + G]
        for m in range(8):
            v[m + 8] = (v[m] * tws[(1 << (log2n - 1)) - 1 + G + (m << TOP)]).astype(C64)
        ff_stages(v, tws, G, False, TOP, TOP, log2n - 2)
        ff_store(x, G, TOP, v)
        P = TOP - 4
        while P > 0:
            v = ff_load(x, G, P)
            ff_stages(v, tws, G & ((1 << P) - 1), False, P, P, P + 3)
            ff_store(x, G, P, v)
            P -= 4
        v = ff_load(x, G, 0)
        ff_stages(v, tws, 0 * G, False, 0, 0, f0_hi)
        for m in range(16):
            v[m] = (v[m] * mult_br[16 * G + m]).astype(C64)
        ff_stages(v, tws, 0 * G, True, 0, 0, 3)
        ff_store(x, G, 0, v)
        S = 4
        while log2n - S > 4:
            v = ff_load(x, G, S)
            ff_stages(v, tws, G & ((1 << S) - 1), True, S, S, S + 3)
            ff_store(x, G, S, v)
            S += 4
        v = ff_load(x, G, TOP)
        ff_stages(v, tws, G, True, TOP, sl, log2n - 1)
        for m in range(8):
            if b >= b0:
                res[(b - b0) * N2 + m * NT + G] = (ov[m] + v[m] * inv).astype(C64)
            ov[m] = (v[m + 8] * inv).astype(C64)
    return res


def bitrev(k, n):
    r = 0
    for b in range(n):
        if k & (1 << b):
            r |= 1 << (n - 1 - b)
    return r


@pytest.mark.parametrize("flen,log2n", [(1024, 10), (2048, 11)])
def test_fftfilt_fast_decomposition_matches_oracle(port, flen, log2n):
    rs = np.random.RandomState(3)
    N, N2 = flen, flen // 2
    o = port.PortFftFilt(0, 300 / 48000.0, 3000 / 48000.0, flen)
    filt = np.asarray(o.filter()).astype(C64)
    tw = np.exp(-2j * np.pi * np.arange(N2) / N).astype(C64)
    m = filt.copy()                      # runSSB, usb, bin 0 rejected: ff_upload_mult
    m[0] = 0
    m[N2 + 1:] = 0
    m[N2] = 1.0
    br = np.array([m[bitrev(k, log2n)] for k in range(N)], C64)
    n = N2 * 7
    xin = ((rs.randn(n) + 1j * rs.randn(n)) * 8000).astype(C64)
    want = o.run(1, xin, True, False)
    ovl = np.zeros(N2, C64)
    # two ranges, the second starting inside the call (it recomputes block 3 with its stores off)
    got = np.concatenate([fftfilt_fast_model(log2n, xin, ovl, br, tw, 0, 4), fftfilt_fast_model(log2n, xin, ovl, br, tw, 4, 7)])
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
