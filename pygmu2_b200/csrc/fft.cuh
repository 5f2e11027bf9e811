// fft.cuh -- shared-memory Stockham FFT building blocks (sm_100a).
//
// A real transform of 2B samples is computed as a B-point complex transform of
// z[n] = w[2n] + i*w[2n+1] plus an O(B) split step, and stored as a *packed*
// half spectrum of exactly B complex bins: bin 0 carries (Re X[0], Re X[B]) (both
// purely real), bins 1..B-1 are X[k].  One delay-line row is therefore B*8 bytes,
// a power of two, 16-byte aligned for float4 / bulk-copy streaming.
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
template <bool INV>
__device__ __forceinline__ float2 tw_load(const float2* __restrict__ tw, int idx) {
  float2 w = __ldg(tw + idx);
  if (INV) w.y = -w.y;
  return w;
}

// Autosort Stockham passes, radix 4 while the remaining factor allows, then one radix 2.
// n = number of complex points (power of two >= 2); threads t in [0, T) of one transform
// cooperate; `a` holds the input, `b` is scratch; returns the buffer holding the result.
// twM[k] = exp(-2*pi*i*k/(2n)), so exp(-2*pi*i*m/n) = twM[2m].
// Every thread of the CTA must call this (it contains __syncthreads()).
template <bool INV>
__device__ __forceinline__ float2* stockham_passes(float2* a, float2* b, const int n, const int t, const int T,
                                                   const float2* __restrict__ twM) {
  int Ns = 1;
  while (Ns < n) {
    if (n / Ns >= 4) {
      const int nj = n >> 2;
      const int tws = 2 * (n / (Ns * 4));
      for (int j = t; j < nj; j += T) {
        const int k = j & (Ns - 1);
        float2 v0 = a[j], v1 = a[j + nj], v2 = a[j + 2 * nj], v3 = a[j + 3 * nj];
        if (k != 0) {
          const int m = k * tws;
          v1 = cmul(v1, tw_load<INV>(twM, m));
          v2 = cmul(v2, tw_load<INV>(twM, 2 * m));
          v3 = cmul(v3, tw_load<INV>(twM, 3 * m));
        }
        const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3);
        const float2 d = csub(v1, v3);
        const float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // (v1-v3) * (+/- i)
        const int j0 = ((j - k) << 2) + k;
        b[j0] = cadd(t0, t2);
        b[j0 + Ns] = cadd(t1, t3);
        b[j0 + 2 * Ns] = csub(t0, t2);
        b[j0 + 3 * Ns] = csub(t1, t3);
      }
      Ns <<= 2;
    } else {
      const int nj = n >> 1;  // here Ns == n/2
      for (int j = t; j < nj; j += T) {
        const float2 v0 = a[j];
        float2 v1 = a[j + nj];
        if (j != 0) v1 = cmul(v1, tw_load<INV>(twM, 2 * j));
        b[j] = cadd(v0, v1);
        b[j + Ns] = csub(v0, v1);
      }
      Ns <<= 1;
    }
    __syncthreads();
    float2* tmp = a;
    a = b;
    b = tmp;
  }
  return a;
}

// Split step after the forward transform: Z (n complex) -> packed half spectrum bin k.
__device__ __forceinline__ float2 r2c_bin(const float2* Z, const int n, const int k,
                                          const float2* __restrict__ twM) {
  if (k == 0) {
    const float2 z0 = Z[0];
    return make_float2(z0.x + z0.y, z0.x - z0.y);
  }
  const float2 zk = Z[k];
  float2 zc = Z[n - k];
  zc.y = -zc.y;
  const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
  const float2 d = csub(zk, zc);
  const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * d
  const float2 w = __ldg(twM + k);
  return cadd(e, cmul(w, o));
}

// Merge step before the inverse transform: packed half spectrum Y -> Z bin k (unscaled).
__device__ __forceinline__ float2 c2r_bin(const float2* Y, const int n, const int k,
                                          const float2* __restrict__ twM) {
  if (k == 0) {
    const float2 y0 = Y[0];
    return make_float2(0.5f * (y0.x + y0.y), 0.5f * (y0.x - y0.y));
  }
  const float2 xk = Y[k];
  float2 xc = Y[n - k];
  xc.y = -xc.y;
  const float2 e = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y + xc.y));
  const float2 d = csub(xk, xc);
  float2 w = __ldg(twM + k);
  w.y = -w.y;
  const float2 o = cmul(make_float2(0.5f * w.x, 0.5f * w.y), d);
  return make_float2(e.x - o.y, e.y + o.x);  // e + i*o
}

}  // namespace pgx
