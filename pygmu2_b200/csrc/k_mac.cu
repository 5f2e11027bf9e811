// k_mac.cu -- K3/K4: frequency-domain delay-line complex multiply-accumulate.
//
//   Y[o][k] = sum over terms r of  X[xrow(o,r)][k] * H[hrow(o,r)][k]
//
// conv mode (mix=0): o = (stream s, out channel c), terms j = 0..P-1 are the delay-line slots of
//   (s, g(c)); slot j pairs with filter row q0 + j of the reversed+doubled spectrum set, q0 = R-1-head,
//   i.e. partition p = (head - j) mod R (R = ring rows).  This is the uniformly partitioned form of the reference's
//   single X*H product (convolve_pe.py:314-317).
// mix mode (mix=1): o = c and the terms run over (s, j): the MixPE sum over streams (mix_pe.py:92-94)
//   and the per-source HRTF multiply (spatial_pe.py:503-504) are the same accumulation.
//
// HBM-bound: each term streams one B*8-byte row of X (and of H when filters are distinct) exactly
// once, as 16-byte vector loads with U rows in flight per thread; 8 flop per 16 bytes.  When all
// streams share one filter (ST > 1) a CTA covers ST streams so each filter row is fetched once per
// ST delay-line rows.  The work list (out tile x bin tile x term split) is a flat 1-D grid whose
// size the host picks as a near-integer number of full waves (see mac_plan()).
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace pgx {

static constexpr int kMacThreads = 128;

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

// ST = streams per CTA that share the filter rows (conv mode, one resident filter); U = rows in flight.
template <bool MIX, int ST, int U>
__global__ void __launch_bounds__(kMacThreads) k_fdl_mac(const MacArgs a) {
  __shared__ float4 red[kMacThreads];
  const int lanes = a.W4 < kMacThreads ? a.W4 : kMacThreads;  // threads covering one row segment
  const int G = kMacThreads / lanes;                          // term-parallel groups
  const int g = threadIdx.x / lanes, lane = threadIdx.x - g * lanes;
  // flat work index -> (split, out tile, bin tile)
  const int ktiles = a.W4 / lanes;
  int w = blockIdx.x;
  const int kt = w % ktiles;
  w /= ktiles;
  const int ot = w % a.n_otiles;
  const int sp = w / a.n_otiles;
  const int kv = kt * lanes + lane;  // float4 index within the row
  const int r0 = sp * a.terms_per_split;
  const int r1 = min(r0 + a.terms_per_split, a.n_terms);
  // out tile -> channel c and first stream
  int c, s0;
  if (MIX) {
    c = ot;
    s0 = 0;
  } else if (ST == 1) {
    c = ot % a.c_out;
    s0 = ot / a.c_out;
  } else {
    c = ot % a.c_out;
    s0 = (ot / a.c_out) * ST;
  }
  const int gx = (a.c_x == 1) ? 0 : c;
  const int fc = (a.c_f == 1) ? 0 : c;
  const bool bin0 = (kv == 0);  // packed bin 0 = two independent real bins (DC, Nyquist)
  const size_t rs = (size_t)a.W4;
  const size_t stream_stride = (size_t)a.c_x * a.R * rs;  // delay-line rows of one stream
  const float4* hbase = a.Hd + ((size_t)(a.fmap[s0] * a.c_f + fc) * 2 * a.R + a.q0) * rs + kv;
  const float4* xbase = a.fdl + ((size_t)(s0 * a.c_x + gx) * a.R) * rs + kv;

  float4 acc[ST];
  float2 acc0[ST];
#pragma unroll
  for (int t = 0; t < ST; ++t) {
    acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc0[t] = make_float2(0.f, 0.f);
  }
  int nst = ST;
  if (ST > 1) nst = min(ST, a.N - s0);

  for (int r = r0 + g; r < r1; r += G * U) {
    float4 xv[U][ST], hv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = r + u * G;
      if (rr < r1) {
        const float4 *xp, *hp;
        if (MIX) {
          const int s = rr / a.Pt, jj = rr - s * a.Pt;
          int j = a.off + jj;
          j = (a.jfix >= 0) ? a.jfix : j + (j >= a.skip ? a.nskip : 0);
          xp = a.fdl + ((size_t)(s * a.c_x + gx) * a.R + j) * rs + kv;
          hp = a.Hd + ((size_t)(__ldg(a.fmap + s) * a.c_f + fc) * 2 * a.R + a.q0 + j) * rs + kv;
        } else {
          int j = a.off + rr;
          j = (a.jfix >= 0) ? a.jfix : j + (j >= a.skip ? a.nskip : 0);
          xp = xbase + (size_t)j * rs;
          hp = hbase + (size_t)j * rs;
        }
        hv[u] = (ST > 1) ? __ldg(hp) : ld_stream(hp);
#pragma unroll
        for (int t = 0; t < ST; ++t)
          xv[u][t] = (t < nst) ? ld_stream(xp + (size_t)t * stream_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < ST; ++t) xv[u][t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int t = 0; t < ST; ++t) {
        cmac2(acc[t], xv[u][t], hv[u]);
        if (bin0) {
          acc0[t].x = fmaf(xv[u][t].x, hv[u].x, acc0[t].x);
          acc0[t].y = fmaf(xv[u][t].y, hv[u].y, acc0[t].y);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < ST; ++t) {
    if (bin0) {
      acc[t].x = acc0[t].x;
      acc[t].y = acc0[t].y;
    }
    if (G > 1) {  // rows narrower than the CTA: groups took interleaved terms, fold them
      __syncthreads();
      red[threadIdx.x] = acc[t];
      __syncthreads();
      if (g == 0) {
        for (int gg = 1; gg < G; ++gg) {
          const float4 v = red[gg * lanes + lane];
          acc[t].x += v.x;
          acc[t].y += v.y;
          acc[t].z += v.z;
          acc[t].w += v.w;
        }
      }
    }
    if (g == 0 && t < nst) {
      size_t prow;  // partial row: [split][out] -- or, per-stream mix (mix == 2), [stream][split][channel]
      if (MIX) prow = (size_t)sp * a.n_out + c;
      else if (a.mix == 2) prow = ((size_t)(s0 + t) * a.n_split + sp) * a.c_out + c;
      else prow = (size_t)sp * a.n_out + (s0 + t) * a.c_out + c;
      a.yspec[prow * rs + kv] = acc[t];
    }
  }
}

template <bool MIX, int ST, int U>
static int mac_occupancy() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fdl_mac<MIX, ST, U>, kMacThreads, 0);
  return nb > 0 ? nb : 1;
}

// Which accumulate kernel: PGX_MAC=ldg|tma overrides the default.
static bool want_tma(int W4) {
  const char* e = getenv("PGX_MAC");
  if (e && !strcmp(e, "ldg")) return false;
  if (e && !strcmp(e, "tma")) return tma_supported(W4);
  return false;  // default (measured A/B)
}

// Pick the kernel variant, the work layout, the stream tile and the term split so that the work list is a
// near-integer number of full waves over the SMs.
//   mix = false          : one out row per (stream, channel), terms = the stream's Pt partitions
//   mix, Pt >= 32        : "per-stream" mix (layout 2): the same items as conv mode -- the filter row base is
//                          resolved once per CTA, no per-row stream lookup -- writing N*n_split partial rows per
//                          channel that the fold kernel sums (the MixPE sum over streams, mix_pe.py:92-94)
//   mix, short Pt        : "flattened" mix (layout 1): terms run over (stream, partition) pairs
MacPlan mac_plan(int N, int c_out, int W4, int n_terms, bool mix, bool shared_filter, int sm_count) {
  MacPlan p{};
  const int Pt = mix ? (n_terms / (N > 0 ? N : 1)) : n_terms;
  p.layout = !mix ? 0 : (Pt >= 32 ? 2 : 1);
  const bool flat = p.layout == 1;
  if (p.layout == 2) n_terms = Pt;
  p.st = (!mix && shared_filter && N >= 4) ? 4 : 1;
  p.variant = want_tma(W4) && p.layout != 2 ? 1 : 0;
  const int lanes = p.variant ? 256 : (W4 < kMacThreads ? W4 : kMacThreads);
  const int groups = p.variant ? 1 : kMacThreads / lanes, ktiles = W4 / lanes;
  int occ;
  if (p.variant) occ = tma_occupancy_of(flat, p.st);
  else if (flat) occ = mac_occupancy<true, 1, 8>();
  else if (p.st == 4) occ = mac_occupancy<false, 4, 4>();
  else occ = mac_occupancy<false, 1, 8>();
  p.n_otiles = flat ? c_out : ((N + p.st - 1) / p.st) * c_out;
  const long resident = (long)sm_count * occ;
  const long base = (long)p.n_otiles * ktiles;
  const int min_terms = 16 * groups;  // at least two unrolled batches per group and split
  int max_split = n_terms / min_terms;
  if (max_split < 1) max_split = 1;
  if (max_split > 4096) max_split = 4096;
  // cost model: waves x (terms per CTA + fixed per-item overhead expressed in row-terms)
  const int overhead = p.variant ? 2 : 6 * groups;
  double best = 1e300;
  int best_s = 1;
  for (int s = 1; s <= max_split; ++s) {
    const long items = base * s;
    const long waves = (items + resident - 1) / resident;
    const int tps = (n_terms + s - 1) / s;
    const double cost = (double)waves * (double)(tps + overhead);
    if (cost < best * 0.995) {
      best = cost;
      best_s = s;
    }
  }
  p.n_split = best_s;
  p.terms_per_split = (n_terms + best_s - 1) / best_s;
  p.n_split = (n_terms + p.terms_per_split - 1) / p.terms_per_split;
  p.n_partials = p.layout == 2 ? N * p.n_split : p.n_split;
  p.grid = (int)(base * p.n_split);
  p.occupancy = occ;
  p.persistent_ctas = (int)resident;
  return p;
}

bool describe_fdl_mac(const MacArgs& a, LaunchDesc* d) {
  d->func = nullptr;
  if (a.variant == 1) return false;  // the bulk-async kernel has its own launcher
  const int lanes = a.W4 < kMacThreads ? a.W4 : kMacThreads;
  d->grid = dim3((unsigned)(a.n_otiles * (a.W4 / lanes) * a.n_split));
  d->block = dim3(kMacThreads);
  d->smem = 0;
  if (a.mix == 1) d->func = reinterpret_cast<const void*>(k_fdl_mac<true, 1, 8>);
  else if (a.st == 4) d->func = reinterpret_cast<const void*>(k_fdl_mac<false, 4, 4>);
  else d->func = reinterpret_cast<const void*>(k_fdl_mac<false, 1, 8>);
  return true;
}

void launch_fdl_mac(const MacArgs& a, cudaStream_t st) {
  if (a.variant == 1) {
    launch_fdl_mac_tma(a, a.persistent_ctas, st);
    return;
  }
  LaunchDesc d;
  describe_fdl_mac(a, &d);
  void* params[] = {const_cast<MacArgs*>(&a)};
  cudaLaunchKernel(d.func, d.grid, d.block, params, d.smem, st);
}

// Second stage of the split accumulation: out[e] = sum_sp in[sp][e], e over n_out*W4 float4 columns.
// 32 columns x 8 split-groups per CTA; each group walks its splits with 4 loads in flight, then the
// groups are folded through shared memory in a fixed order (deterministic result).
__global__ void __launch_bounds__(256) k_reduce_partials(const float4* __restrict__ in, float4* __restrict__ out,
                                                         const int n_split, const int64_t n_cols) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < n_cols) {
    int sp = grp;
    for (; sp + 24 < n_split; sp += 32) {
      const float4 a = __ldg(in + (int64_t)sp * n_cols + col);
      const float4 b = __ldg(in + (int64_t)(sp + 8) * n_cols + col);
      const float4 c = __ldg(in + (int64_t)(sp + 16) * n_cols + col);
      const float4 d = __ldg(in + (int64_t)(sp + 24) * n_cols + col);
      acc.x += (a.x + b.x) + (c.x + d.x);
      acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z);
      acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; sp < n_split; sp += 8) {
      const float4 a = __ldg(in + (int64_t)sp * n_cols + col);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  red[grp][lane] = acc;
  __syncthreads();
  if (grp == 0 && col < n_cols) {
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 v = red[g][lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[col] = acc;
  }
}

bool describe_reduce_partials(int n_out, int W4, LaunchDesc* d) {
  const int64_t n_cols = (int64_t)n_out * W4;
  d->func = reinterpret_cast<const void*>(k_reduce_partials);
  d->grid = dim3((unsigned)((n_cols + 31) / 32));
  d->block = dim3(256);
  d->smem = 0;
  return true;
}

void launch_reduce_partials(const float4* in, float4* out, int n_split, int n_out, int W4, cudaStream_t st) {
  const int64_t n_cols = (int64_t)n_out * W4;
  const int grid = (int)((n_cols + 31) / 32);
  k_reduce_partials<<<grid, 256, 0, st>>>(in, out, n_split, n_cols);
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

// Same sum when there are many inputs and few elements (a 1024-voice bank at 64-sample pulls): one thread per
// element would walk 1024 dependent global loads.  Here a CTA owns 32 elements; warps 1..31 stage chunks of
// input rows in shared memory (every load of a chunk in flight at once, double-buffered) while warp 0 adds
// the previous chunk's rows in input order -- the float32 result is the same left-to-right sum, bit for bit.
static constexpr int kTallRows = 192;
__global__ void __launch_bounds__(1024) k_mix_sum_tall(const float* __restrict__ in, const int n_inputs,
                                                       const int64_t n_elems, float* __restrict__ out) {
  __shared__ float rows[2][kTallRows][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  const bool live = e < n_elems;
  const int n_chunks = (n_inputs + kTallRows - 1) / kTallRows;
  auto stage = [&](int c) {  // every load of the warp's rows is issued before the first shared-memory store
    const int r0 = c * kTallRows, nr = min(kTallRows, n_inputs - r0);
    constexpr int kPer = (kTallRows + 30) / 31;
    float t[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int r = warp - 1 + 31 * i;
      t[i] = (live && r < nr) ? __ldcg(in + (int64_t)(r0 + r) * n_elems + e) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int r = warp - 1 + 31 * i;
      if (r < nr) rows[c & 1][r][lane] = t[i];
    }
  };
  if (warp > 0) stage(0);
  __syncthreads();
  float acc = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    if (warp > 0) {
      if (c + 1 < n_chunks) stage(c + 1);
    } else {
      const int nr = min(kTallRows, n_inputs - c * kTallRows);
      const float(*buf)[32] = rows[c & 1];
      int r = 0;
      if (c == 0) {  // the sum starts from the first input itself (first.data.copy(), mix_pe.py:92)
        acc = buf[0][lane];
        r = 1;
      }
      for (; r + 8 <= nr; r += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = buf[r + u][lane];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, t[u]);
      }
      for (; r < nr; ++r) acc = __fadd_rn(acc, buf[r][lane]);
    }
    __syncthreads();
  }
  if (warp == 0 && live) out[e] = acc;
}

void launch_mix_sum(const float* in, int32_t n_inputs, int64_t n_elems, float* out, cudaStream_t st) {
  if (n_inputs >= 32 && n_elems <= 8192) {
    k_mix_sum_tall<<<(int)((n_elems + 31) / 32), 1024, 0, st>>>(in, n_inputs, n_elems, out);
    return;
  }
  int64_t blocks = (n_elems + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_mix_sum<<<(int)blocks, 256, 0, st>>>(in, n_inputs, n_elems, out);
}

}  // namespace pgx
