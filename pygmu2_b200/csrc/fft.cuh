// fft.cuh -- register-resident Stockham FFT building blocks (sm_100a).
//
// A real transform of 2B samples is computed as a B-point complex transform of
// z[n] = w[2n] + i*w[2n+1] plus an O(B) split step, and stored as a *packed*
// half spectrum of exactly B complex bins: bin 0 carries (Re X[0], Re X[B]) (both
// purely real), bins 1..B-1 are X[k].  One delay-line row is therefore B*8 bytes,
// a power of two, 16-byte aligned for float4 / bulk-copy streaming.
//
// The complex transform: N = 2^LOG2N points, N/8 threads per transform, each thread owning the
// 8 elements at positions j + m*N/8.  Passes are radix 8 (in-register 8-point butterflies) with one
// final radix-4 / radix-2 pass when LOG2N is not a multiple of 3; between passes the data is
// exchanged through padded, ping-pong shared-memory buffers (Stockham autosort order, one
// __syncthreads per pass).  The first pass takes its inputs straight from registers (loaded from
// global by the caller) and the last pass leaves its outputs in registers, again at j + m*N/8.
//
// Replaces numpy.fft.rfft / irfft at reference convolve_pe.py:237-239,313,318.
#pragma once
#include <cuda_runtime.h>

namespace pgx {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 d) {
  return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
}
// GLOBAL_TW: the table is read from global memory through the read-only path; otherwise it was staged in shared
template <bool INV, bool GLOBAL_TW>
__device__ __forceinline__ float2 tw_get(const float2* tw, int idx) {
  float2 w = GLOBAL_TW ? __ldg(tw + idx) : tw[idx];
  if (INV) w.y = -w.y;
  return w;
}

// padded shared-memory index: one pad element per 16 keeps both the contiguous reads and the
// stride-8 writes of the first pass free of 8-byte bank conflicts
__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }
// The exchange after the SECOND pass scatters, per half-warp, two runs of 8 elements that lie 64 apart; with
// the pad above they land 68 slots apart (4 banks off: a 2-way conflict that no additive pad common to all
// exchanges removes).  That one exchange therefore uses its own pad, 8 slots per 64 elements (72 apart = 8
// banks off), under which its contiguous reads stay conflict-free too.  E = index of the exchange.
// Only the large transforms take it (FftCfg::PAD2): the small ones run beside the resident accumulate CTAs, where
// the extra shared memory per CTA costs more than the conflicts (measured A/B: C2 +11 % step time with it).
template <int E>
__device__ __forceinline__ int phys_x(int i) {
  return E == 1 ? i + ((i >> 6) << 3) : i + (i >> 4);
}

template <int LOG2N>
struct FftCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int T8 = N / 8;                   // threads per transform
  static constexpr int NP8 = LOG2N / 3;              // radix-8 passes
  static constexpr int RLAST = 1 << (LOG2N % 3);     // 1 = none, else a final radix-2 / radix-4 pass
  static constexpr bool PAD2 = LOG2N >= 11;          // second exchange uses its own pad (see phys_x)
  static constexpr int PADN = N + N / (PAD2 ? 8 : 16);  // padded buffer length (elements)
  // Small CTAs on purpose: a 64-thread CTA needs ~3K registers, so the latency-critical FFT kernels fit into
  // what the resident accumulate CTAs leave free on an SM instead of displacing one of them.
  static constexpr int CTA = T8 < 64 ? 64 : T8;      // threads per CTA
  static constexpr int FPB = CTA / T8;               // transforms per CTA
  static constexpr bool SMEM_TW = FPB >= 4;          // stage the 2N-entry twiddle table in shared memory when
                                                     // several transforms share it; else read it through L1
  static constexpr int SMEM_BYTES = (SMEM_TW ? 2 * N : 0) * 8 + FPB * 2 * PADN * 8;
};

template <bool INV>
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
  const float2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

// 4-point DFT, natural order in and out
template <bool INV>
__device__ __forceinline__ void bfly4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 c0 = cadd(v0, v2), c1 = csub(v0, v2), c2 = cadd(v1, v3), c3 = mul_mi<INV>(csub(v1, v3));
  v0 = cadd(c0, c2);
  v1 = cadd(c1, c3);
  v2 = csub(c0, c2);
  v3 = csub(c1, c3);
}

// 8-point DFT, natural order in and out
template <bool INV>
__device__ __forceinline__ void bfly8(float2 (&v)[8]) {
  constexpr float kH = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  // b1 *= w8, b2 *= w8^2 = -+i, b3 *= w8^3 with w8 = exp(-+ i pi/4)
  b1 = INV ? make_float2(kH * (b1.x - b1.y), kH * (b1.x + b1.y)) : make_float2(kH * (b1.x + b1.y), kH * (b1.y - b1.x));
  b2 = mul_mi<INV>(b2);
  b3 = INV ? make_float2(-kH * (b3.x + b3.y), kH * (b3.x - b3.y)) : make_float2(kH * (b3.y - b3.x), -kH * (b3.x + b3.y));
  bfly4<INV>(a0, a1, a2, a3);
  bfly4<INV>(b0, b1, b2, b3);
  v[0] = a0; v[1] = b0; v[2] = a1; v[3] = b1; v[4] = a2; v[5] = b2; v[6] = a3; v[7] = b3;
}

// All passes of one N-point transform.  v[m] = element j + m*T8 on entry and on exit.
// sA / sB: this transform's two padded buffers; `first` selects which one the first exchange uses
// (the other may still be read by the caller's pre-processing).  tw[k] = exp(-2*pi*i*k/(2N)).
// Every thread of the CTA must call this (it contains __syncthreads()).
template <int LOG2N, bool INV>
__device__ __forceinline__ void fft_passes(float2 (&v)[8], float2* sA, float2* sB, const int first, const int j,
                                           const float2* tw) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  constexpr bool GTW = !C::SMEM_TW;
  float2* cur = nullptr;
  int Ns = 1;
  float2 w1 = make_float2(1.f, 0.f), w2 = w1, w4 = w1;  // twiddles of the coming pass, fetched one barrier early
#pragma unroll
  for (int p = 0; p < C::NP8; ++p) {
    const int k = j & (Ns - 1);
    if (p > 0) {
      __syncthreads();
      if (!GTW) {  // table staged in shared memory by the caller: readable only after the first barrier
        const int m = 2 * k * (N / (8 * Ns));
        w1 = tw_get<INV, GTW>(tw, m);
        w2 = tw_get<INV, GTW>(tw, 2 * m);
        w4 = tw_get<INV, GTW>(tw, 4 * m);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = cur[(C::PAD2 && p == 2) ? phys_x<1>(j + r * T8) : phys_x<0>(j + r * T8)];
      const float2 w3 = cmul(w1, w2);
      v[1] = cmul(v[1], w1);
      v[2] = cmul(v[2], w2);
      v[3] = cmul(v[3], w3);
      v[4] = cmul(v[4], w4);
      v[5] = cmul(v[5], cmul(w4, w1));
      v[6] = cmul(v[6], cmul(w4, w2));
      v[7] = cmul(v[7], cmul(w4, w3));
    }
    if (GTW && p + 1 < C::NP8) {  // issue the next pass's twiddle loads now: their latency hides behind this pass
      const int Nn = Ns * 8;
      const int m = 2 * (j & (Nn - 1)) * (N / (8 * Nn));
      w1 = tw_get<INV, GTW>(tw, m);      // one scattered load; w^2 and w^4 by squaring (2 ulp, far inside 1e-5)
      w2 = cmul(w1, w1);
      w4 = cmul(w2, w2);
    }
    bfly8<INV>(v);
    if (p < C::NP8 - 1 || C::RLAST > 1) {
      float2* d = ((p + first) & 1) ? sB : sA;
      const int j0 = (j - k) * 8 + k;
#pragma unroll
      for (int r = 0; r < 8; ++r) d[(C::PAD2 && p == 1) ? phys_x<1>(j0 + r * Ns) : phys_x<0>(j0 + r * Ns)] = v[r];
      cur = d;
    }
    Ns *= 8;
  }
  if (C::RLAST == 4) {
    constexpr int NB = N / 4;  // butterflies in this pass; Ns == NB, so k == jj
    float2 a1[2], a2[2];
    if (!GTW) __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {  // global table: twiddles are requested before the barrier
      const int m = 2 * (j + u * T8);  // 2*k*(N/(4*Ns)) with Ns = N/4
      a1[u] = tw_get<INV, GTW>(tw, m);
      a2[u] = tw_get<INV, GTW>(tw, 2 * m);
    }
    if (GTW) __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int jj = j + u * T8;
      constexpr int E = (C::PAD2 && C::NP8 - 1 == 1) ? 1 : 0;  // pad of the exchange that filled `cur`
      float2 t0 = cur[phys_x<E>(jj)], t1 = cur[phys_x<E>(jj + NB)], t2 = cur[phys_x<E>(jj + 2 * NB)],
             t3 = cur[phys_x<E>(jj + 3 * NB)];
      t1 = cmul(t1, a1[u]);
      t2 = cmul(t2, a2[u]);
      t3 = cmul(t3, cmul(a1[u], a2[u]));
      bfly4<INV>(t0, t1, t2, t3);
      v[u] = t0; v[u + 2] = t1; v[u + 4] = t2; v[u + 6] = t3;  // position jj + r*NB = j + (u + 2r)*T8
    }
  } else if (C::RLAST == 2) {
    constexpr int NB = N / 2;
    float2 a1[4];
    if (!GTW) __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) a1[u] = tw_get<INV, GTW>(tw, 2 * (j + u * T8));
    if (GTW) __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + u * T8;
      constexpr int E = (C::PAD2 && C::NP8 - 1 == 1) ? 1 : 0;
      float2 t0 = cur[phys_x<E>(jj)], t1 = cur[phys_x<E>(jj + NB)];
      t1 = cmul(t1, a1[u]);
      bfly2<INV>(t0, t1);
      v[u] = t0; v[u + 4] = t1;  // position jj + r*NB = j + (u + 4r)*T8
    }
  }
}

// tw[j + m*N/8] for m = 0..7 from ONE table read: tw[j] times the constant 16th roots exp(-2*pi*i*m/16)
// (tw is the table of 2N-th roots, so a step of N/8 entries is a 16th of a turn).
__device__ __forceinline__ void split_twiddles(const float2 t0, float2 (&t)[8]) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  t[0] = t0;
  t[1] = cmul(t0, make_float2(c1, -s1));
  t[2] = cmul(t0, make_float2(h, -h));
  t[3] = cmul(t0, make_float2(s1, -c1));
  t[4] = make_float2(t0.y, -t0.x);  // times -i
  t[5] = cmul(t0, make_float2(-s1, -c1));
  t[6] = cmul(t0, make_float2(-h, -h));
  t[7] = cmul(t0, make_float2(-c1, -s1));
}

// Split step after the forward transform: packed half-spectrum bin k from Z[k] and Z[N-k].
__device__ __forceinline__ float2 r2c_bin(const float2 zk, const float2 znk, const float2 w) {
  const float2 zc = make_float2(znk.x, -znk.y);
  const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
  const float2 d = csub(zk, zc);
  const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * d
  return cadd(e, cmul(w, o));
}

// Merge step before the inverse transform: Z[k] (unscaled) from packed bins Y[k], Y[N-k].
__device__ __forceinline__ float2 c2r_bin(const float2 xk, const float2 xnk, const float2 w) {
  const float2 xc = make_float2(xnk.x, -xnk.y);
  const float2 e = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y + xc.y));
  const float2 d = csub(xk, xc);
  const float2 o = cmul(make_float2(0.5f * w.x, -0.5f * w.y), d);  // conj(w)/2 * d
  return make_float2(e.x - o.y, e.y + o.x);                        // e + i*o
}

}  // namespace pgx
