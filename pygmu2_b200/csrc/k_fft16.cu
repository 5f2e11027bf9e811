// k_fft16.cu -- radix-16 variant of the fused single-partition step (k_conv1) for B = 4096 (C1: 4096-tap FIRs).
//
// Same arithmetic as k_conv1<12, false, false> in k_fft.cu (ingest -> real FFT of the 2B window -> X * H -> inverse
// real FFT -> emit; replaces np.fft.rfft / irfft + X*H of reference convolve_pe.py:294-322), different machine
// mapping: 4096 = 16 * 16 * 16, ONE transform per CTA of 256 threads, 16 complex elements per thread, three
// radix-16 passes per direction -- two exchanges per transform instead of the three of the radix-8 kernel, half
// the barrier participants, and (SHFL) the exchange whose partners sit in one half-warp done by warp shuffles
// (a 16 x 16 register transpose: xor-butterfly, 4 stages) instead of through shared memory:
//
//   forward  n = 256a + 16b + c  ->  k = k0 + 16 k1 + 256 k2
//     thread (c, b) regs a : DFT16 over a            -> regs k0        (inputs straight from global, coalesced)
//     exchange 1 (shared)  : regs k0 <-> thread digit b (the HIGH digit: partners in other warps)
//     twiddle W256^(b k0), DFT16 over b              -> regs k1
//     exchange 2 (SHUFFLE) : regs k1 <-> thread digit c (the LOW digit: partners in the same half-warp)
//     twiddle W4096^(c (k0 + 16 k1)), DFT16 over c   -> regs k2        thread (k1, k0) holds Z[k0 + 16 k1 + 256 k2]
//   split   (shared, XOR-swizzled so that the stride-16 writes are conflict-free) -> natural layout, X[k] * H[k]
//   merge   (shared)       -> thread (b, c) holds W[256a + 16b + c]
//   inverse mirrors the forward: DFT16 over a, twiddle, exchange 1 (SHUFFLE), DFT16 over b, exchange 2 (shared),
//           twiddle, DFT16 over c -> thread j holds z[j + 256 n2]: coalesced emit
//
// Barriers per block step: 4 (SHFL) or 6, against 8 in the radix-8 kernel.  Shared memory: two 32 KB buffers, no
// padding (swizzle) -> 3 CTAs per SM.  The filter row is prefetched into L2 when the CTA starts and read with plain
// coalesced loads at the product.  The index maps were checked against numpy.fft in a register-level numpy model
// before any GPU time was spent (tests/kernel_model.py::fft16_*).
#include <cstdlib>

#include "fft.cuh"
#include "kernels.h"

namespace pgx {

namespace {

constexpr int kN = 4096, kT = 256;

__host__ __device__ constexpr int P16(int q) { return ((q & 3) << 2) | (q >> 2); }

// 16-point DFT, natural order in; logical output k is left in register P16(k)
template <bool INV>
__device__ __forceinline__ void bfly16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
  for (int i = 0; i < 4; ++i) bfly4<INV>(v[i], v[i + 4], v[i + 8], v[i + 12]);
  // v[i + 4q] *= W16^(i q)   (forward: exp(-2 pi i e / 16); inverse: conjugate)
  auto rot = [](float2 a, float wr, float wi) {  // a * (wr + i wi), wi already carries the direction
    return make_float2(fmaf(a.x, wr, -a.y * wi), fmaf(a.x, wi, a.y * wr));
  };
  constexpr float sg = INV ? 1.f : -1.f;
  v[1 + 4] = rot(v[1 + 4], c1, sg * s1);    // e = 1
  v[1 + 8] = rot(v[1 + 8], h, sg * h);      // e = 2
  v[1 + 12] = rot(v[1 + 12], s1, sg * c1);  // e = 3
  v[2 + 4] = rot(v[2 + 4], h, sg * h);      // e = 2
  v[2 + 8] = mul_mi<INV>(v[2 + 8]);         // e = 4: -i (forward) / +i
  v[2 + 12] = rot(v[2 + 12], -h, sg * h);   // e = 6
  v[3 + 4] = rot(v[3 + 4], s1, sg * c1);    // e = 3
  v[3 + 8] = rot(v[3 + 8], -h, sg * h);     // e = 6
  v[3 + 12] = rot(v[3 + 12], -c1, -sg * s1);  // e = 9: exp(-+ 2 pi i 9/16) = (-c1, +-s1)
#pragma unroll
  for (int q = 0; q < 4; ++q) bfly4<INV>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// v[c] *= w^c for c = 1..15 given w^1 and w^4 (w^2, w^8 by squaring: at most 3 roundings deep)
__device__ __forceinline__ void mul_powers(float2 (&v)[16], const float2 w1, const float2 w4) {
  const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w8 = cmul(w4, w4);
  v[1] = cmul(v[1], w1);
  v[2] = cmul(v[2], w2);
  v[3] = cmul(v[3], w3);
  v[4] = cmul(v[4], w4);
  const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  v[5] = cmul(v[5], w5);
  v[6] = cmul(v[6], w6);
  v[7] = cmul(v[7], w7);
  v[8] = cmul(v[8], w8);
  v[9] = cmul(v[9], cmul(w8, w1));
  v[10] = cmul(v[10], cmul(w8, w2));
  v[11] = cmul(v[11], cmul(w8, w3));
  v[12] = cmul(v[12], cmul(w8, w4));
  v[13] = cmul(v[13], cmul(w8, w5));
  v[14] = cmul(v[14], cmul(w8, w6));
  v[15] = cmul(v[15], cmul(w8, w7));
}
// same for registers holding LOGICAL index q in physical register P16(q)
__device__ __forceinline__ void mul_powers_perm(float2 (&v)[16], const float2 w1, const float2 w4) {
  const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w8 = cmul(w4, w4);
  const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  v[P16(1)] = cmul(v[P16(1)], w1);
  v[P16(2)] = cmul(v[P16(2)], w2);
  v[P16(3)] = cmul(v[P16(3)], w3);
  v[P16(4)] = cmul(v[P16(4)], w4);
  v[P16(5)] = cmul(v[P16(5)], w5);
  v[P16(6)] = cmul(v[P16(6)], w6);
  v[P16(7)] = cmul(v[P16(7)], w7);
  v[P16(8)] = cmul(v[P16(8)], w8);
  v[P16(9)] = cmul(v[P16(9)], cmul(w8, w1));
  v[P16(10)] = cmul(v[P16(10)], cmul(w8, w2));
  v[P16(11)] = cmul(v[P16(11)], cmul(w8, w3));
  v[P16(12)] = cmul(v[P16(12)], cmul(w8, w4));
  v[P16(13)] = cmul(v[P16(13)], cmul(w8, w5));
  v[P16(14)] = cmul(v[P16(14)], cmul(w8, w6));
  v[P16(15)] = cmul(v[P16(15)], cmul(w8, w7));
}

// 16 x 16 transpose of (lane within a half-warp, physical register): new[lane l][p] = old[lane p][l]
__device__ __forceinline__ void transpose16_shfl(float2 (&v)[16], const int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if ((r & s) == 0) {
        const float2 send = up ? v[r] : v[r | s];
        float2 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, s);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, s);
        if (up) v[r] = recv; else v[r | s] = recv;
      }
    }
  }
}

// element index -> shared-memory slot: low nibble XORed with the second nibble.  Conflict-free for 16 lanes that
// differ in the low nibble (natural order) AND for 16 lanes that differ in the second nibble (stride 16).
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 4) & 15); }

// tw[j + 256 m] = tw[j] * exp(-2 pi i m / 32): the 16 twiddles of a thread's bins from one table read
__device__ __forceinline__ float2 w32(const float2 t0, const int m) {
  constexpr float C[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                           0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                           0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                           -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                           -0.92387953251128675613f, -0.98078528040323044913f};
  constexpr float S[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                           0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                           0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                           0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                           0.38268343236508977173f, 0.19509032201612826785f};
  return cmul(t0, make_float2(C[m], -S[m]));
}

}  // namespace

template <bool SHFL, bool PAIR>
__global__ void __launch_bounds__(kT, 3) k_conv1_r16(const R2CArgs a, const C2RArgs k) {
  extern __shared__ float2 sm[];
  float2* sA = sm;
  float2* sB = sm + kN;
  const float2* __restrict__ tw = a.tw;  // [2N] exp(-2 pi i t / 2N), read through L1
  const int j = threadIdx.x, lo = j & 15, hi = j >> 4;
  const int64_t f = blockIdx.x;          // one transform per CTA: f = stream * c_x + channel
  const int s = (int)(f / a.c_x), cx = (int)(f - (int64_t)s * a.c_x);
  const int fc = (k.c_f == 1) ? 0 : cx;
  const float2* __restrict__ hrow = k.Hd + ((size_t)(__ldg(k.fmap + s) * k.c_f + fc) * 2 * k.R + (k.R - 1)) * kN;
  // the filter row (32 KB) on its way into L2 while the forward transform runs: one 128-byte line per thread
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(reinterpret_cast<const char*>(hrow) + (size_t)j * 128));

  // ---- ingest: element q = j + 256 m = samples (2q, 2q+1) of the window [previous block | new block]
  float2 v[16];
  {
    float* cur = a.hist + ((size_t)f * 2 + a.half) * kN;
    const float* prev = a.hist + ((size_t)f * 2 + (a.half ^ 1)) * kN;
    const float* xn = a.x + (int64_t)s * a.xs + (int64_t)cx * a.xc + a.x_off;
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = *reinterpret_cast<const float2*>(prev + 2 * (j + m * kT));
#pragma unroll
    for (int m = 8; m < 16; ++m) v[m] = __ldg(reinterpret_cast<const float2*>(xn + 2 * (j + m * kT) - kN));
#pragma unroll
    for (int m = 8; m < 16; ++m) *reinterpret_cast<float2*>(cur + 2 * (j + m * kT) - kN) = v[m];
  }

  // ================= forward =================
  bfly16<false>(v);                                                 // over a -> k0 (register P16(k0))
#pragma unroll
  for (int k0 = 0; k0 < 16; ++k0) sA[256 * k0 + j] = v[P16(k0)];    // exchange 1: regs k0 <-> thread digit b
  float2 t2[16];                                                    // W256^(b * k0'), k0' = hi: broadcast reads
#pragma unroll
  for (int b = 1; b < 16; ++b) t2[b] = __ldg(tw + 32 * b * hi);
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 16; ++b) v[b] = sA[256 * hi + 16 * b + lo];
#pragma unroll
  for (int b = 1; b < 16; ++b) v[b] = cmul(v[b], t2[b]);
  bfly16<false>(v);                                                 // over b -> k1 (register P16(k1))
  int k1;                                                           // the k1 this thread holds from here on
  if (SHFL) {
    transpose16_shfl(v, lo);                                        // exchange 2: physical transpose ...
    k1 = P16(lo);                                                   // ... so lane l holds logical k1 = P16(l), regs c
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) sB[swz(256 * hi + 16 * q + lo)] = v[P16(q)];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = sB[swz(256 * hi + 16 * lo + c)];
    k1 = lo;
  }
  {
    const int e = hi + 16 * k1;                                     // W4096^(c e) = tw[2 c e]
    mul_powers(v, __ldg(tw + 2 * e), __ldg(tw + 8 * e));
  }
  bfly16<false>(v);                                                 // over c -> k2: Z[hi + 16 k1 + 256 k2]

  // ================= split: packed half spectrum X[k], natural layout k = j + 256 m =================
  float2* sS = SHFL ? sB : sA;   // (no barrier since exchange 1's reads in the SHFL variant: A may still be read)
  if constexpr (PAIR) {
    // pair layout: bins below N/2 at swz(k), bins from N/2 up at N/2 + swz(N - k) -- the mirror partner of bin k then
    // sits at N/2 + swz(k): both reads of a pair are natural-order runs (no run straddles a swizzle group, which cost
    // every mirrored read a 2-way bank conflict), and the stride-16 writes stay conflict-free
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) sS[swz(hi + 16 * k1 + 256 * k2)] = v[P16(k2)];
#pragma unroll
    for (int k2 = 8; k2 < 16; ++k2) sS[kN / 2 + swz((kN - (hi + 16 * k1 + 256 * k2)) & (kN / 2 - 1))] = v[P16(k2)];
  } else {
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) sS[swz(hi + 16 * k1 + 256 * k2)] = v[P16(k2)];
  }
  const float2 twj = __ldg(tw + j);
  float2* sM = SHFL ? sA : sB;
  const int kb = 16 * lo + hi;
  const float2 twb = __ldg(tw + kb);
  float2 hA[8], hB[8];
  if constexpr (PAIR) {  // the filter row (in L2 by now): requested before the barrier, consumed after it
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
      const int kA = j + 256 * sl, kB = (sl == 0 && j == 0) ? kN / 2 : kN - kA;
      hA[sl] = __ldg(hrow + kA);
      hB[sl] = __ldg(hrow + kB);
    }
  }
  __syncthreads();
  if constexpr (PAIR) {
    // ---- split, product and merge on mirror PAIRS (k, N-k), all in registers: thread j owns the 8 pairs
    // (j + 256 s, N - j - 256 s), s < 8 (each pair exactly once over the CTA; thread 0 owns bins 0 and N/2 instead of a
    // pair).  X[k] = E + w O and X[N-k] = conj(E - w O) share E, O and the twiddle; so do W[k] = e + i o and
    // W[N-k] = conj(e - i o): half the shared-memory reads of the one-bin-per-thread form, and no merge exchange --
    // only the redistribution of W into the inverse transform's input layout.
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
      const int kA = j + 256 * sl;
      const float2 zA = sS[swz(kA)], zB = sS[kN / 2 + swz(kA)];
      float2 wa, wb;
      if (sl == 0 && j == 0) {
        // bin 0 (packed DC / Nyquist, two real products) and bin N/2 (its own mirror, twiddle -i)
        const float2 x0 = make_float2(zA.x + zA.y, zA.x - zA.y);
        const float2 y0 = make_float2(x0.x * hA[sl].x, x0.y * hA[sl].y);
        wa = make_float2(0.5f * (y0.x + y0.y), 0.5f * (y0.x - y0.y));
        const float2 wq = make_float2(0.f, -1.f);
        const float2 yq = cmul(r2c_bin(zB, zB, wq), hB[sl]);
        wb = c2r_bin(yq, yq, wq);
      } else {
        const float2 w = w32(twj, sl);
        const float2 E = make_float2(0.5f * (zA.x + zB.x), 0.5f * (zA.y - zB.y));
        const float2 O = make_float2(0.5f * (zA.y + zB.y), -0.5f * (zA.x - zB.x));   // -i/2 (zA - conj zB)
        const float2 t = cmul(w, O);
        const float2 xA = make_float2(E.x + t.x, E.y + t.y), xB = make_float2(E.x - t.x, -(E.y - t.y));
        const float2 yA = cmul(xA, hA[sl]), yB = cmul(xB, hB[sl]);
        const float2 e = make_float2(0.5f * (yA.x + yB.x), 0.5f * (yA.y - yB.y));
        const float2 d = make_float2(yA.x - yB.x, yA.y + yB.y);                       // yA - conj yB
        const float2 o = cmul(make_float2(0.5f * w.x, -0.5f * w.y), d);
        wa = make_float2(e.x - o.y, e.y + o.x);                                       // e + i o
        wb = make_float2(e.x + o.y, -(e.y - o.x));                                    // conj(e - i o)
      }
      sM[swz(kA)] = wa;                    // same pair layout: W[k] at swz(k), W[N-k] at N/2 + swz(k)
      sM[kN / 2 + swz(kA)] = wb;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = sM[swz(kb + 256 * q)];
#pragma unroll
    for (int q = 8; q < 16; ++q) v[q] = sM[kN / 2 + swz((kN - (kb + 256 * q)) & (kN / 2 - 1))];
  } else {
  float2 X[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int kk = j + m * kT;
    const float2 zk = sS[swz(kk)];
    if (m == 0 && j == 0) {
      X[m] = make_float2(zk.x + zk.y, zk.x - zk.y);
    } else {
      X[m] = r2c_bin(zk, sS[swz(kN - kk)], w32(twj, m));
    }
  }
  // ================= product with the filter row (pre-scaled by 1/B) =================
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float2 hh[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) hh[m] = __ldg(hrow + j + (g * 8 + m) * kT);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float2 x = X[g * 8 + m];
      X[g * 8 + m] = (g == 0 && m == 0 && j == 0) ? make_float2(x.x * hh[m].x, x.y * hh[m].y)  // bin 0: two real bins
                                                  : cmul(x, hh[m]);
    }
  }
  // ================= merge: W[k] for the inverse transform, thread (b = lo, c = hi), k = 256a + 16b + c ==========
#pragma unroll
  for (int m = 0; m < 16; ++m) sM[swz(j + m * kT)] = X[m];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int kk = kb + 256 * q;
    const float2 yk = sM[swz(kk)];
    if (q == 0 && kb == 0) {
      v[q] = make_float2(0.5f * (yk.x + yk.y), 0.5f * (yk.x - yk.y));
    } else {
      v[q] = c2r_bin(yk, sM[swz(kN - kk)], w32(twb, q));
    }
  }
  }

  // ================= inverse =================
  bfly16<true>(v);                                                  // over a -> n0 (register P16(n0))
  {                                                                 // twiddle conj W256^(b n0), b = lo
    float2 w1 = __ldg(tw + 32 * lo), w4 = __ldg(tw + 128 * lo);
    w1.y = -w1.y;
    w4.y = -w4.y;
    mul_powers_perm(v, w1, w4);
  }
  int n0;
  float2* sX = SHFL ? sB : sA;
  if (SHFL) {
    transpose16_shfl(v, lo);                                        // exchange 1: regs n0 <-> thread digit b
    n0 = P16(lo);
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) sA[swz(256 * hi + 16 * q + lo)] = v[P16(q)];
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 16; ++b) v[b] = sA[swz(256 * hi + 16 * lo + b)];
    n0 = lo;
    sX = sB;
  }
  bfly16<true>(v);                                                  // over b -> n1 (register P16(n1))
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) sX[256 * hi + 16 * n1 + n0] = v[P16(n1)];   // exchange 2: regs n1 <-> thread digit c
  float2 w1 = __ldg(tw + 2 * j), w4 = __ldg(tw + 8 * j);            // conj W4096^(c j)
  w1.y = -w1.y;
  w4.y = -w4.y;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = sX[256 * c + j];
  mul_powers(v, w1, w4);
  bfly16<true>(v);                                                  // over c -> n2: z[j + 256 n2] in register P16(n2)

  // ================= emit: the new block = elements n >= N/2 (n2 >= 8), samples (2n, 2n+1) - N =================
  {
    const int c = cx;  // !FAN: output channel = source channel
    float* y = k.y + (int64_t)s * k.ys + (int64_t)c * k.yc + k.y_off;
    const float* xd = k.xdry ? k.xdry + (int64_t)s * k.xs + (int64_t)c * k.xc + k.x_off : nullptr;
    const float* ad = k.add ? k.add + (int64_t)s * k.as + (int64_t)c * k.ac + k.y_off : nullptr;
    const bool gains = (k.wet != 1.0f) || xd;
#pragma unroll
    for (int n2 = 8; n2 < 16; ++n2) {
      const int i = 2 * (j + n2 * kT) - kN;
      float2 val = v[P16(n2)];
      if (ad) {
        const float2 t = *reinterpret_cast<const float2*>(ad + i);
        val.x += t.x;
        val.y += t.y;
      }
      if (gains) {
        val.x = __fmul_rn(val.x, k.wet);
        val.y = __fmul_rn(val.y, k.wet);
        if (xd) {
          const float2 d = __ldg(reinterpret_cast<const float2*>(xd + i));
          val.x = __fadd_rn(__fmul_rn(d.x, k.dry), val.x);
          val.y = __fadd_rn(__fmul_rn(d.y, k.dry), val.y);
        }
      }
      *reinterpret_cast<float2*>(y + i) = val;
    }
  }
}

// PGX_FFT16 = 0 (radix-8 kernel of k_fft.cu), 1 (radix-16, both exchanges through shared memory), 2 (radix-16 with the
// half-warp exchange done by shuffles), 3 / 4 = 1 / 2 with split, product and merge on mirror pairs in registers.  Read when a bank is created, so one process can A/B the variants.
int conv1_r16_default() {
  const char* e = getenv("PGX_FFT16");
  const int v = e ? atoi(e) : 3;  // measured A/B on one box (profiles/r02_ab_fft_c1.txt)
  return (v < 0 || v > 4) ? 0 : v;
}

bool describe_conv1_r16(const R2CArgs& a, const C2RArgs& k, LaunchDesc* d) {
  const int variant = k.fft16;
  // B = 4096, one partition, no fan-out, whole aligned blocks: everything else keeps the general kernel
  if (variant == 0 || a.B != kN || k.n_past > 0 || k.n_split > 0 || (a.c_x == 1 && k.c_out > 1) || !a.fast || !k.fast || a.mixdown)
    return false;
  constexpr int smem = 2 * kN * (int)sizeof(float2);
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    cudaFuncSetAttribute(k_conv1_r16<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_conv1_r16<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_conv1_r16<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_conv1_r16<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    done[dev & 63] = true;
  }
  d->func = variant == 1 ? reinterpret_cast<const void*>(k_conv1_r16<false, false>)
          : variant == 2 ? reinterpret_cast<const void*>(k_conv1_r16<true, false>)
          : variant == 3 ? reinterpret_cast<const void*>(k_conv1_r16<false, true>)
                         : reinterpret_cast<const void*>(k_conv1_r16<true, true>);
  d->grid = dim3((unsigned)a.n_fft);
  d->block = dim3(kT);
  d->smem = smem;
  return true;
}

bool launch_conv1_r16(const R2CArgs& a, const C2RArgs& k, cudaStream_t st) {
  LaunchDesc d;
  if (!describe_conv1_r16(a, k, &d)) return false;
  void* params[] = {const_cast<R2CArgs*>(&a), const_cast<C2RArgs*>(&k)};
  cudaLaunchKernel(d.func, d.grid, d.block, params, d.smem, st);
  return true;
}

}  // namespace pgx
