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

// Time-tiled pass (conv layout): the past sums of T consecutive output blocks from ONE read of the delay line.
//   S_kappa[o][k] = sum over committed rows j of X[j][k] * H[partition p0(j) + kappa][k],  kappa = 0..T-1,
// p0(j) = (head - j) mod R = the row's partition for the first output block; a term exists while p0 >= 1 (the row is
// committed) and p0 + kappa <= P-1 (inside the filter).  With the reversed + doubled filter rows the partition
// p0(j) + kappa sits in row (R-1-head) + (j - kappa) (+R when negative): it depends on d = j - kappa only, so over a run
// of consecutive slots the T filter rows a slot needs are a WINDOW that slides by one row per slot.  The kernel keeps the
// window in registers: per slot ONE delay-line row (the HBM stream, loaded once per T output blocks) and ONE new filter
// row, T complex MACs per loaded pair.  The committed slots are at most two runs of consecutive slots (the open slot and
// the spare slot are skipped); each CTA walks the part of its term range that falls into each run.
// The rows younger than the pass (at most T-1) are added by the output stage (C2RArgs.n_recent).
// Packed bin 0 (DC and Nyquist: two real products) is lane kv = 0's first complex: that lane multiplies with three
// operands swapped for 0 / the other component (selects on the loaded values), every lane runs the same FMAs.
// CHECK = false: a full batch whose every (slot, kappa) term exists -- no predicates in the body.
template <int ST, int T, int U, bool CHECK>
__device__ __forceinline__ void mac_tile_batch(const float4* __restrict__ xp, const float4* __restrict__ hp, const size_t rs,
                                               const size_t stream_stride, const int nst, const bool bin0, const int n_slots,
                                               const int p0, const int P, float4 (&hw)[U + T - 1], float4 (&acc)[T][ST]) {
  float4 xv[U][ST];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const bool in = !CHECK || u < n_slots;
    hw[T - 1 + u] = in ? __ldg(hp + (size_t)u * rs) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < ST; ++t)
      xv[u][t] = (in && (!CHECK || t < nst)) ? ld_stream(xp + (size_t)u * rs + (size_t)t * stream_stride)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float xim[ST], xre[ST];   // first complex: x.x as used by the imaginary sum, -x.y as used by the real sum
#pragma unroll
    for (int t = 0; t < ST; ++t) {
      xim[t] = bin0 ? 0.f : xv[u][t].x;
      xre[t] = bin0 ? 0.f : -xv[u][t].y;
    }
#pragma unroll
    for (int kp = 0; kp < T; ++kp) {
      if (!CHECK || (u < n_slots && p0 - u + kp <= P - 1)) {   // (uniform over the CTA)
        const float4 h = hw[T - 1 + u - kp];
        const float hsel = bin0 ? h.y : h.x;
#pragma unroll
        for (int t = 0; t < ST; ++t) {
          const float4 x = xv[u][t];
          float4& c = acc[kp][t];
          c.x = fmaf(x.x, h.x, c.x);
          c.x = fmaf(xre[t], h.y, c.x);
          c.y = fmaf(xim[t], h.y, c.y);
          c.y = fmaf(x.y, hsel, c.y);
          c.z = fmaf(x.z, h.z, c.z);
          c.z = fmaf(-x.w, h.w, c.z);
          c.w = fmaf(x.z, h.w, c.w);
          c.w = fmaf(x.w, h.z, c.w);
        }
      }
    }
  }
#pragma unroll
  for (int w = 0; w < T - 1; ++w) hw[w] = hw[w + U];
}

template <int ST, int T, int U>
__global__ void __launch_bounds__(kMacThreads, ST == 1 ? 5 : 4) k_fdl_mac_tile(const MacArgs a) {
  const int ktiles = a.W4 / kMacThreads;   // tiled banks have B >= 256: a row is at least one CTA wide
  int w = blockIdx.x;
  const int kt = w % ktiles;
  w /= ktiles;
  const int ot = w % a.n_otiles;
  const int sp = w / a.n_otiles;
  const int kv = kt * kMacThreads + threadIdx.x;
  const int r0 = sp * a.terms_per_split;
  const int r1 = min(r0 + a.terms_per_split, a.n_terms);
  const int c = ot % a.c_out;
  const int s0 = (ot / a.c_out) * ST;
  const int gx = (a.c_x == 1) ? 0 : c;
  const int fc = (a.c_f == 1) ? 0 : c;
  const bool bin0 = (kv == 0);
  const size_t rs = (size_t)a.W4;
  const size_t stream_stride = (size_t)a.c_x * a.R * rs;
  const float4* xbase = a.fdl + ((size_t)(s0 * a.c_x + gx) * a.R) * rs + kv;
  const float4* hfil = a.Hd + ((size_t)(a.fmap[s0] * a.c_f + fc) * 2 * a.R) * rs + kv;
  const int nst = min(ST, a.N - s0);
  const int qb = a.R - 1 - a.head;

  float4 acc[T][ST];
#pragma unroll
  for (int kp = 0; kp < T; ++kp)
#pragma unroll
    for (int t = 0; t < ST; ++t) acc[kp][t] = make_float4(0.f, 0.f, 0.f, 0.f);
  // term r -> slot off + r, + nskip from slot `skip` on: terms below rsk = skip - off are the first run
  const int rsk = a.skip - a.off;
  for (int run = 0; run < 2; ++run) {
    const int rb = run == 0 ? r0 : max(r0, rsk), re = run == 0 ? min(r1, rsk) : r1;
    if (rb >= re) continue;
    const int jbeg = a.off + rb + (run ? a.nskip : 0), jend = a.off + re + (run ? a.nskip : 0);
    float4 hw[U + T - 1];     // hw[w] <-> d = jb - (T-1) + w
#pragma unroll
    for (int w2 = 0; w2 < T - 1; ++w2) {
      int q = qb + jbeg - (T - 1) + w2;
      q += (q < 0) ? a.R : 0;
      hw[w2] = __ldg(hfil + (size_t)q * rs);
    }
    const float4* xp = xbase + (size_t)jbeg * rs;
    const float4* hp = hfil + (size_t)(qb + jbeg) * rs;
    int p0 = a.head - jbeg;
    p0 += (p0 < 0) ? a.R : 0;   // the run never crosses the open slot: p0 falls by one per slot from here
    for (int jb = jbeg; jb < jend; jb += U, p0 -= U, xp += (size_t)U * rs, hp += (size_t)U * rs) {
      if (jb + U <= jend && p0 + T - 1 <= a.P - 1 && nst == ST)
        mac_tile_batch<ST, T, U, false>(xp, hp, rs, stream_stride, nst, bin0, U, p0, a.P, hw, acc);
      else
        mac_tile_batch<ST, T, U, true>(xp, hp, rs, stream_stride, nst, bin0, jend - jb, p0, a.P, hw, acc);
    }
  }
#pragma unroll
  for (int kp = 0; kp < T; ++kp)
#pragma unroll
    for (int t = 0; t < ST; ++t)
      if (t < nst) {
        const size_t prow = (size_t)sp * a.n_out + (s0 + t) * a.c_out + c;
        a.yspec[(size_t)((a.tile_set0 + kp) % a.tile_nsets) * a.tile_stride + prow * rs + kv] = acc[kp][t];
      }
}

// The instantiations of the tiled pass: (streams per CTA, tile, rows in flight).  PGX_TILE_ST / PGX_TILE_U pick
// another one than the plan's default for A/B runs.
struct TileVariant { int st, tile, u; const void* func; };
static const TileVariant kTileVariants[] = {
    {1, 2, 4, reinterpret_cast<const void*>(k_fdl_mac_tile<1, 2, 4>)}, {1, 2, 8, reinterpret_cast<const void*>(k_fdl_mac_tile<1, 2, 8>)},
    {2, 2, 4, reinterpret_cast<const void*>(k_fdl_mac_tile<2, 2, 4>)}, {4, 2, 2, reinterpret_cast<const void*>(k_fdl_mac_tile<4, 2, 2>)},
    {1, 4, 2, reinterpret_cast<const void*>(k_fdl_mac_tile<1, 4, 2>)}, {1, 4, 4, reinterpret_cast<const void*>(k_fdl_mac_tile<1, 4, 4>)},
    {2, 4, 2, reinterpret_cast<const void*>(k_fdl_mac_tile<2, 4, 2>)}, {2, 4, 4, reinterpret_cast<const void*>(k_fdl_mac_tile<2, 4, 4>)},
};
static const TileVariant* tile_variant(int st, int tile, int u) {
  for (const TileVariant& v : kTileVariants)
    if (v.st == st && v.tile == tile && v.u == u) return &v;
  return nullptr;
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

MacPlan mac_plan_tiled(int N, int c_out, int W4, int n_terms, bool shared_filter, int sm_count, int tile) {
  MacPlan p{};
  p.layout = 0;
  p.variant = 0;
  p.tile = tile;
  p.st = (shared_filter && N >= 2) ? 2 : 1;
  p.tile_u = (tile == 2 && p.st == 1) ? 8 : 4;
  if (const char* e = getenv("PGX_TILE_ST")) {
    const int v = atoi(e);
    if (shared_filter && N >= v) p.st = v;
  }
  if (const char* e = getenv("PGX_TILE_U")) p.tile_u = atoi(e);
  const TileVariant* tv = tile_variant(p.st, tile, p.tile_u);
  if (!tv) {
    p.st = (shared_filter && N >= 2) ? 2 : 1;
    p.tile_u = (tile == 2 && p.st == 1) ? 8 : 4;
    tv = tile_variant(p.st, tile, p.tile_u);
  }
  const int lanes = kMacThreads;
  int ktiles = W4 / lanes;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tv->func, kMacThreads, 0);
  if (occ < 1) occ = 1;
  {
    // rows staged through shared memory by bulk-async copies (k_mac_tile_tma.cu): PGX_TILE_TMA=0|1
    bool tma = false;
    if (const char* e = getenv("PGX_TILE_TMA")) tma = atoi(e) != 0;
    int st2, tps, stages, occ2;
    if (tma && tile_tma_supported(W4) && tile_tma_config(shared_filter, N, W4, tile, &st2, &tps, &stages, &occ2)) {
      p.variant = 1;
      p.st = st2; p.tile_u = tps; p.tile_stages = stages;
      occ = occ2;
      ktiles = tile_tma_ktiles(W4);
      // persistent CTAs per SM: fewer than fit, so that the ingest / output kernels of the blocks in flight find room
      // beside the pass (PGX_TILE_CTAS)
      int cap = 2;
      if (const char* e = getenv("PGX_TILE_CTAS")) cap = atoi(e);
      if (cap >= 1 && cap < occ) occ = cap;
    }
  }
  p.n_otiles = ((N + p.st - 1) / p.st) * c_out;
  const long resident = (long)sm_count * occ;
  const long base = (long)p.n_otiles * ktiles;
  int max_split = n_terms / 16;
  if (max_split < 1) max_split = 1;
  if (max_split > 8) max_split = 8;     // the output stage folds the split rows itself: keep them few (<= kFoldAbove)
  int force_split = 0;
  if (const char* e = getenv("PGX_TILE_SPLIT")) force_split = atoi(e);
  if (force_split >= 1 && force_split <= max_split) max_split = force_split;
  double best = 1e300;
  int best_s = 1;
  for (int s = 1; s <= max_split; ++s) {
    const long items = base * s;
    const long waves = (items + resident - 1) / resident;
    const int tps = (n_terms + s - 1) / s;
    const double cost = (double)waves * (double)(tps + 6);
    if (cost < best * 0.995 || s == force_split) {
      best = cost;
      best_s = s;
    }
  }
  p.terms_per_split = (n_terms + best_s - 1) / best_s;
  p.n_split = (n_terms + p.terms_per_split - 1) / p.terms_per_split;
  p.n_partials = p.n_split;
  p.grid = (int)(base * p.n_split);
  p.occupancy = occ;
  p.persistent_ctas = (int)resident;
  return p;
}

bool describe_fdl_mac(const MacArgs& a, LaunchDesc* d) {
  d->func = nullptr;
  if (a.variant == 1) return false;  // the bulk-async kernel has its own launcher
  if (a.tile > 1) {
    d->grid = dim3((unsigned)(a.n_otiles * (a.W4 / kMacThreads) * a.n_split));
    d->block = dim3(kMacThreads);
    d->smem = 0;
    const TileVariant* tv = tile_variant(a.st, a.tile, a.tile_u);
    if (!tv) return false;
    d->func = tv->func;
    return true;
  }
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
    if (a.tile > 1) launch_fdl_mac_tile_tma(a, st);
    else launch_fdl_mac_tma(a, a.persistent_ctas, st);
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
