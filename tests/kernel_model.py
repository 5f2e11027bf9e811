"""
Numpy model of the device algorithms in pygmu2_b200/csrc (index-for-index):
Stockham radix-4/2 passes, the packed real-FFT pre/post-processing, the
reversed+doubled filter-spectrum layout and the ring / partial-block schedule of
the bank.  CPU-only test aid: it lets the host-side schedule and the kernels'
index math be checked against the oracle without a GPU.  Not a product path.
"""
from __future__ import annotations

import numpy as np


def twiddle_table(M: int) -> np.ndarray:
    """tw[k] = exp(-2*pi*i*k/M), k in [0, M): the device table (computed in double, stored float)."""
    k = np.arange(M, dtype=np.float64)
    return np.exp(-2j * np.pi * k / M).astype(np.complex64)


def stockham(z: np.ndarray, twM: np.ndarray, inverse: bool) -> np.ndarray:
    """n-point complex FFT, n = len(z) power of two >= 2, radix-4 passes then one radix-2 if needed.
    twM is the table for M = 2n, so exp(-2*pi*i*m/n) = twM[2m]."""
    n = z.shape[0]
    a = z.astype(np.complex64).copy()
    b = np.empty_like(a)
    Ns = 1
    while Ns < n:
        R = 4 if n // Ns >= 4 else 2
        nj = n // R
        for j in range(nj):
            k = j % Ns
            m = k * (n // (Ns * R))
            v = [a[j + r * nj] for r in range(R)]
            for r in range(1, R):
                w = twM[2 * m * r]
                if inverse:
                    w = np.conj(w)
                v[r] = np.complex64(v[r] * w)
            if R == 4:
                s = 1j if inverse else -1j
                t0, t1 = v[0] + v[2], v[0] - v[2]
                t2, t3 = v[1] + v[3], (v[1] - v[3]) * s
                o = [t0 + t2, t1 + t3, t0 - t2, t1 - t3]
            else:
                o = [v[0] + v[1], v[0] - v[1]]
            j0 = (j // Ns) * Ns * R + k
            for r in range(R):
                b[j0 + r * Ns] = np.complex64(o[r])
        a, b = b, a
        Ns *= R
    return a


def r2c_packed(w: np.ndarray, twM: np.ndarray) -> np.ndarray:
    """Real window w (2B floats) -> packed half spectrum (B complex): slot[0] = (X[0].re, X[B].re)."""
    B = w.shape[0] // 2
    z = (w[0::2] + 1j * w[1::2]).astype(np.complex64)
    Z = stockham(z, twM, inverse=False)
    out = np.empty(B, np.complex64)
    out[0] = (Z[0].real + Z[0].imag) + 1j * (Z[0].real - Z[0].imag)
    for k in range(1, B):
        zk, zc = Z[k], np.conj(Z[B - k])
        e = 0.5 * (zk + zc)
        o = -0.5j * (zk - zc)
        out[k] = np.complex64(e + twM[k] * o)
    return out


def c2r_packed(Y: np.ndarray, twM: np.ndarray) -> np.ndarray:
    """Packed half spectrum (B complex) -> 2B real samples, UNSCALED by 1/B (folded into the filter)."""
    B = Y.shape[0]
    Z = np.empty(B, np.complex64)
    x0, xn = Y[0].real, Y[0].imag
    Z[0] = 0.5 * (x0 + xn) + 0.5j * (x0 - xn)
    for k in range(1, B):
        xk, xc = Y[k], np.conj(Y[B - k])
        e = 0.5 * (xk + xc)
        o = 0.5 * np.conj(twM[k]) * (xk - xc)
        Z[k] = np.complex64(e + 1j * o)
    z = stockham(Z, twM, inverse=True)
    out = np.empty(2 * B, np.float32)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def mac_packed(acc, x, h):
    """acc += x*h on packed spectra: bin 0 carries two independent real bins."""
    acc[1:] += x[1:] * h[1:]
    acc[0] += (x[0].real * h[0].real) + 1j * (x[0].imag * h[0].imag)


class ModelBank:
    """One stream, one channel pair; same state machine as csrc/pgx_engine (ring head, fill, halves)."""

    def __init__(self, h: np.ndarray, B: int, vectorized: bool = True):
        self.B, self.L = B, h.shape[0]
        self.P = -(-self.L // B)
        self.tw = twiddle_table(2 * B)
        self.vec = vectorized
        P = self.P
        self.Hd = np.zeros((2 * P, B), np.complex64)
        for p in range(P):
            w = np.zeros(2 * B, np.float32)
            seg = h[p * B:(p + 1) * B]
            w[:seg.shape[0]] = seg
            Hp = self._r2c(w) / np.float32(B)
            q = P - 1 - p
            self.Hd[q] = Hp
            self.Hd[q + P] = Hp
        self.reset()

    def _r2c(self, w):
        if not self.vec:
            return r2c_packed(w, self.tw)
        B = self.B
        X = np.fft.rfft(w.astype(np.float64))
        out = X[:B].astype(np.complex64)
        out[0] = X[0].real + 1j * X[B].real
        return out

    def _c2r(self, Y):
        if not self.vec:
            return c2r_packed(Y, self.tw)
        B = self.B
        X = np.empty(B + 1, np.complex128)
        X[:B] = Y
        X[0] = Y[0].real
        X[B] = Y[0].imag
        return (np.fft.irfft(X, n=2 * B) * B).astype(np.float32)

    def reset(self):
        self.fdl = np.zeros((self.P, self.B), np.complex64)
        self.hist = np.zeros((2, self.B), np.float32)
        self.half = 0      # which half is "cur"
        self.fill = 0
        self.head = 0

    def process(self, x: np.ndarray) -> np.ndarray:
        B, P = self.B, self.P
        y = np.empty_like(x)
        pos, n = 0, x.shape[0]
        while pos < n:
            take = min(B - self.fill, n - pos)
            cur, prev = self.hist[self.half], self.hist[self.half ^ 1]
            cur[self.fill:self.fill + take] = x[pos:pos + take]
            m_new = self.fill + take
            w = np.zeros(2 * B, np.float32)
            w[:B] = prev
            w[B:B + m_new] = cur[:m_new]
            self.fdl[self.head] = self._r2c(w)
            acc = np.zeros(B, np.complex64)
            q0 = P - 1 - self.head
            for j in range(P):
                mac_packed(acc, self.fdl[j], self.Hd[q0 + j])
            yt = self._c2r(acc)
            y[pos:pos + take] = yt[B + self.fill:B + m_new]
            self.fill = m_new
            pos += take
            if self.fill == B:
                self.head = (self.head + 1) % P
                self.half ^= 1
                self.fill = 0
        return y


# ---------------------------------------------------------------------------
# v2 transform (csrc/fft.cuh, register-resident radix-8 passes): thread j of T8 = n/8 holds the 8
# elements at positions j + m*T8 before the first and after the last pass.
def _bfly(v, R, inverse):
    s = 1j if inverse else -1j
    if R == 2:
        return [v[0] + v[1], v[0] - v[1]]
    if R == 4:
        c0, c1, c2, c3 = v[0] + v[2], v[0] - v[2], v[1] + v[3], (v[1] - v[3]) * s
        return [c0 + c2, c1 + c3, c0 - c2, c1 - c3]
    w = np.complex64(np.exp(s * np.pi / 4))  # w8 (forward: e^{-i pi/4})
    a = [v[0] + v[4], v[1] + v[5], v[2] + v[6], v[3] + v[7]]
    b = [v[0] - v[4], (v[1] - v[5]) * w, (v[2] - v[6]) * s, (v[3] - v[7]) * (w * s)]
    e, o = _bfly(a, 4, inverse), _bfly(b, 4, inverse)
    return [e[0], o[0], e[1], o[1], e[2], o[2], e[3], o[3]]


def stockham8(z: np.ndarray, twM: np.ndarray, inverse: bool) -> np.ndarray:
    """n-point complex FFT (n = 2^b >= 8): radix-8 passes, then one radix-4 / radix-2 pass if b % 3 != 0.
    Twiddles: w1, w2, w4 from the table, the other powers by multiplication (as the kernel does)."""
    n = z.shape[0]
    lg = n.bit_length() - 1
    radices = [8] * (lg // 3) + ([1 << (lg % 3)] if lg % 3 else [])
    a = z.astype(np.complex64).copy()
    b = np.empty_like(a)
    T8 = n // 8
    Ns = 1
    for R in radices:
        nb = n // R                      # butterflies in this pass
        for j in range(T8):
            for u in range(8 // R):
                jj = j + u * T8
                k = jj % Ns
                v = [a[jj + r * nb] for r in range(R)]
                if Ns > 1:
                    m = 2 * k * (n // (Ns * R))
                    w1 = twM[m]
                    w2 = twM[2 * m] if R > 2 else None
                    w4 = twM[4 * m] if R > 4 else None
                    if inverse:
                        w1 = np.conj(w1)
                        w2 = np.conj(w2) if w2 is not None else None
                        w4 = np.conj(w4) if w4 is not None else None
                    ws = [None, w1]
                    if R > 2:
                        ws += [w2, np.complex64(w1 * w2)]
                    if R > 4:
                        ws += [w4, np.complex64(w4 * w1), np.complex64(w4 * w2), np.complex64(w4 * np.complex64(w1 * w2))]
                    for r in range(1, R):
                        v[r] = np.complex64(v[r] * ws[r])
                o = _bfly(v, R, inverse)
                j0 = (jj - k) * R + k
                for r in range(R):
                    b[j0 + r * Ns] = np.complex64(o[r])
        a, b = b, a
        Ns *= R
    return a


# ---------------------------------------------------------------------------------------------------------
# Register-level model of the radix-16 kernel (csrc/k_fft16.cu): 256 threads x 16 registers, the same index maps,
# exchanges (shared-memory gather / 16 x 16 half-warp transpose) and twiddles, vectorised over the threads.  The
# CUDA kernel was transcribed from this model once it agreed with numpy.fft.
FFT16_N = N = 4096; T = 256
P=lambda q: ((q&3)<<2)|(q>>2)
Pv=np.array([P(q) for q in range(16)])
def bfly16(v, inv):
    # v[:, n] natural logical input; output logical k at phys P(k)
    s = 1 if inv else -1
    n=np.arange(16); W=np.exp(s*2j*np.pi*np.outer(n,n)/16)
    X = v @ W.T   # X[:,k] = sum_n v[:,n] W[k,n]
    out=np.empty_like(v); out[:,Pv]=X
    return out
def tw(i): return np.exp(-2j*np.pi*(i%(2*N))/(2*N))   # table of 2N-th roots
def fft16_forward(z):
    j=np.arange(T); lo=j&15; hi=j>>4
    v=np.stack([z[j+256*a] for a in range(16)],axis=1)
    v=bfly16(v,False)
    A=np.empty(N,complex)
    for k0 in range(16): A[256*k0+j]=v[:,P(k0)]
    v=np.stack([A[256*hi+16*b+lo] for b in range(16)],axis=1)
    for b in range(16): v[:,b]*=tw(32*b*hi)
    v=bfly16(v,False)
    # transpose within 16-lane groups on physical regs
    nv=np.empty_like(v)
    for l in range(16):
        for p in range(16):
            nv[hi*16+l if False else (np.arange(16)*16+l), p]=v[np.arange(16)*16+p, l]
    v=nv
    k0=hi; k1=Pv[lo]
    e=k0+16*k1
    for c in range(16): v[:,c]*=tw(2*c*e)
    v=bfly16(v,False)
    Z=np.empty(N,complex)
    for k2 in range(16): Z[k0+16*k1+256*k2]=v[:,P(k2)]
    return Z
def fft16_inverse(Wk):
    j=np.arange(T); lo=j&15; hi=j>>4   # lo=b, hi=c
    v=np.stack([Wk[256*a+16*lo+hi] for a in range(16)],axis=1)
    v=bfly16(v,True)
    for n0 in range(16): v[:,P(n0)]*=np.conj(tw(32*lo*n0))
    nv=np.empty_like(v)
    for l in range(16):
        for p in range(16):
            nv[np.arange(16)*16+l, p]=v[np.arange(16)*16+p, l]
    v=nv   # lane l <-> n0=P(l), reg p <-> b
    v=bfly16(v,True)  # logical n1 at phys P(n1)
    n0=Pv[lo]
    Bf=np.empty(N,complex)
    for n1 in range(16): Bf[256*hi+16*n1+n0]=v[:,P(n1)]
    v=np.stack([Bf[256*c+j] for c in range(16)],axis=1)
    for c in range(16): v[:,c]*=np.conj(tw(2*c*j))
    v=bfly16(v,True)
    z=np.empty(N,complex)
    for n2 in range(16): z[j+256*n2]=v[:,P(n2)]
    return z


# ---------------------------------------------------------------------------------------------------------------
# Time-tiled accumulate pass: the index algebra of issue_tile (pgx_api.cu), k_fdl_mac_tile (k_mac.cu) and the recent-row
# loop of k_c2r (k_fft.cu), on plain complex rows.  Ring of R = P + n_spare rows, block t in slot t % R; filter table of 2R
# rows with partition p at rows R-1-p and 2R-1-p (the others stay zero).
def tiled_pass_terms(P: int, T: int, head0: int):
    """(off, skip, nskip, n_terms) exactly as issue_tile sets them for a pass whose first block sits in slot head0."""
    n_spare = (T - 1 if T > 2 else 1) if P > 1 else 0
    R = P + n_spare
    if head0 + n_spare < R:
        off, skip, nskip = 0, head0, 1 + n_spare
    else:
        off, skip, nskip = head0 + n_spare + 1 - R, R, 0
    return R, n_spare, off, skip, nskip, P - 1


def tiled_pass(ring: np.ndarray, Hd: np.ndarray, P: int, T: int, head0: int, n_split: int = 1) -> np.ndarray:
    """The T result sets of one pass: S[kappa] = sum over committed slots j of ring[j] * Hd[row(j - kappa)], walked the
    way the kernel walks them (term range per split, at most two runs of consecutive slots, sliding filter window)."""
    R, n_spare, off, skip, nskip, n_terms = tiled_pass_terms(P, T, head0)
    qb = R - 1 - head0
    S = np.zeros((T,) + ring.shape[1:], complex)
    tps = -(-n_terms // n_split)
    rsk = skip - off
    for sp in range(n_split):
        r0, r1 = sp * tps, min((sp + 1) * tps, n_terms)
        for run in range(2):
            rb, re = (r0, min(r1, rsk)) if run == 0 else (max(r0, rsk), r1)
            if rb >= re:
                continue
            jbeg, jend = off + rb + (nskip if run else 0), off + re + (nskip if run else 0)
            window = {}                                    # d = j - kappa -> filter row, as the register window holds it
            for w in range(T - 1):
                d = jbeg - (T - 1) + w
                q = qb + d
                window[d] = Hd[q + R if q < 0 else q]
            p0 = (head0 - jbeg) % R
            for j in range(jbeg, jend):
                assert p0 == (head0 - j) % R and 1 <= p0 <= P - 1, "a run never crosses the open or a spare slot"
                window[j] = Hd[qb + j]                     # the slot's one new filter row
                for kp in range(T):
                    if p0 + kp <= P - 1:
                        S[kp] += ring[j] * window[j - kp]
                p0 -= 1
    return S


def tiled_output(ring: np.ndarray, Hd: np.ndarray, S_kappa: np.ndarray, R: int, head_u: int, n_recent: int) -> np.ndarray:
    """k_c2r on a tiled bank: the result set of the block + its present term + the n_recent rows committed after the pass."""
    y = S_kappa + ring[head_u] * Hd[R - 1]
    for r in range(1, n_recent + 1):
        y = y + ring[(head_u - r) % R] * Hd[R - 1 - r]
    return y
