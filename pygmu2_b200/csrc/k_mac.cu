// k_mac.cu -- K3/K4: frequency-domain delay-line complex multiply-accumulate.
//
//   Y[o][k] = sum over terms r of  X[xrow(o,r)][k] * H[hrow(o,r)][k]
//
// conv mode (mix=0): o = (stream s, out channel c), terms j = 0..P-1 are the delay-line slots of
//   (s, g(c)); slot j pairs with filter row q0 + j of the reversed+doubled spectrum set, q0 = P-1-head,
//   i.e. partition p = (head - j) mod P.  This is the uniformly partitioned form of the reference's
//   single X*H product (convolve_pe.py:314-317).
// mix mode (mix=1): o = c and the terms run over (s, j): the MixPE sum over streams (mix_pe.py:92-94)
//   and the per-source HRTF multiply (spatial_pe.py:503-504) are the same accumulation.
//
// HBM-bound: each term streams one B*8-byte row of X (and of H when filters are distinct) exactly
// once, as 16-byte vector loads with U rows in flight per thread; 8 flop per 16 bytes.
#include "kernels.h"

namespace pgx {

static constexpr int kMacThreads = 128;
static constexpr int kMacUnroll = 8;

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cmac2(float4& acc, const float4 x, const float4 h) {
  acc.x = fmaf(x.x, h.x, acc.x);
  acc.x = fmaf(-x.y, h.y, acc.x);
  acc.y = fmaf(x.x, h.y, acc.y);
  acc.y = fmaf(x.y, h.x, acc.y);
  acc.z = fmaf(x.z, h.z, acc.z);
  acc.z = fmaf(-x.w, h.w, acc.z);
  acc.w = fmaf(x.z, h.w, acc.w);
  acc.w = fmaf(x.w, h.z, acc.w);
}

template <bool MIX>
__global__ void __launch_bounds__(kMacThreads) k_fdl_mac(const MacArgs a) {
  __shared__ float4 red[kMacThreads];
  __shared__ float2 red0[kMacThreads];
  const int lanes = a.W4 < kMacThreads ? a.W4 : kMacThreads;  // threads covering one row segment
  const int G = kMacThreads / lanes;                          // term-parallel groups
  const int g = threadIdx.x / lanes, lane = threadIdx.x - g * lanes;
  const int kv = blockIdx.y * lanes + lane;                   // float4 index within the row
  const int o = blockIdx.x, sp = blockIdx.z;
  const int r0 = sp * a.terms_per_split;
  const int r1 = min(r0 + a.terms_per_split, a.n_terms);
  const int c = MIX ? o : o % a.c_out;
  const int s_fixed = MIX ? 0 : o / a.c_out;
  const int gx = (a.c_x == 1) ? 0 : c;
  const int fc = (a.c_f == 1) ? 0 : c;
  const bool bin0 = (kv == 0);  // packed bin 0 = two independent real bins (DC, Nyquist)
  const size_t xrow_stride = (size_t)a.W4;
  const float4* hbase_fixed = a.Hd + ((size_t)(a.fmap[s_fixed] * a.c_f + fc) * 2 * a.P + a.q0) * xrow_stride + kv;
  const float4* xbase_fixed = a.fdl + ((size_t)(s_fixed * a.c_x + gx) * a.P) * xrow_stride + kv;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 acc0 = make_float2(0.f, 0.f);

  for (int r = r0 + g; r < r1; r += G * kMacUnroll) {
    float4 xv[kMacUnroll], hv[kMacUnroll];
#pragma unroll
    for (int u = 0; u < kMacUnroll; ++u) {
      const int rr = r + u * G;
      if (rr < r1) {
        const float4 *xp, *hp;
        if (MIX) {
          const int s = rr / a.P, j = rr - s * a.P;
          xp = a.fdl + ((size_t)(s * a.c_x + gx) * a.P + j) * xrow_stride + kv;
          hp = a.Hd + ((size_t)(__ldg(a.fmap + s) * a.c_f + fc) * 2 * a.P + a.q0 + j) * xrow_stride + kv;
        } else {
          xp = xbase_fixed + (size_t)rr * xrow_stride;
          hp = hbase_fixed + (size_t)rr * xrow_stride;
        }
        xv[u] = ld_stream(xp);
        hv[u] = __ldg(hp);
      } else {
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < kMacUnroll; ++u) {
      cmac2(acc, xv[u], hv[u]);
      if (bin0) {
        acc0.x = fmaf(xv[u].x, hv[u].x, acc0.x);
        acc0.y = fmaf(xv[u].y, hv[u].y, acc0.y);
      }
    }
  }
  if (bin0) {
    acc.x = acc0.x;
    acc.y = acc0.y;
  }
  if (G > 1) {  // rows narrower than the CTA: groups took interleaved terms, fold them
    red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0) {
      for (int gg = 1; gg < G; ++gg) {
        const float4 v = red[gg * lanes + lane];
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
    }
  }
  if (g == 0) a.yspec[((size_t)sp * a.n_out + o) * xrow_stride + kv] = acc;
  (void)red0;
}

void launch_fdl_mac(const MacArgs& a, cudaStream_t st) {
  const int lanes = a.W4 < kMacThreads ? a.W4 : kMacThreads;
  dim3 grid(a.n_out, a.W4 / lanes, a.n_split);
  if (a.mix)
    k_fdl_mac<true><<<grid, kMacThreads, 0, st>>>(a);
  else
    k_fdl_mac<false><<<grid, kMacThreads, 0, st>>>(a);
}

// K5 -- MixPE: out = ((in0 + in1) + in2) + ... per element, float32, input order (bit-exact with
// the reference's sequential "+=", mix_pe.py:92-94).  One element per thread, inputs strided by
// n_elems, so every load is a coalesced row segment.
__global__ void __launch_bounds__(256) k_mix_sum(const float* __restrict__ in, const int n_inputs,
                                                 const int64_t n_elems, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
    float acc = in[e];
    for (int i = 1; i < n_inputs; ++i) acc = __fadd_rn(acc, in[(int64_t)i * n_elems + e]);
    out[e] = acc;
  }
}

void launch_mix_sum(const float* in, int32_t n_inputs, int64_t n_elems, float* out, cudaStream_t st) {
  int64_t blocks = (n_elems + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_mix_sum<<<(int)blocks, 256, 0, st>>>(in, n_inputs, n_elems, out);
}

}  // namespace pgx
