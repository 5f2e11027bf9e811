// k_mac_tile_tma.cu -- the time-tiled accumulate pass (k_fdl_mac_tile, k_mac.cu) with its rows staged through shared
// memory by bulk-async copies.
//
// Same sums, same term order per output, same result layout as k_fdl_mac_tile: T output blocks per pass over the
// committed delay-line rows, the T filter rows of a slot a register window that slides by one row per slot.  What
// changes is how the rows reach the SM: the tiled pass does T times the arithmetic per loaded byte, so the LDG kernel
// needs its registers for accumulators and cannot also keep enough loads in flight to fill HBM.  Here a persistent
// CTA = 1 producer warp + SEG/32 consumer warps walks a static list of work items (out tile x bin tile x term split);
// the producer streams row segments into a ring of stages with cp.async.bulk (SASS: UBLKCP), completion counted in
// bytes on an mbarrier per stage, and the bytes in flight live in shared memory (STAGES x stage bytes per CTA).
// A stage = TPS consecutive slots x (ST delay-line rows + the slot's one new filter row); the first stage of a run of
// consecutive slots also carries the T-1 older filter rows that seed the window.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "bulk.cuh"
#include "kernels.h"

namespace pgx {

// SEG = float4 per row segment = consumer threads (one float4 column each): 128 (2 KB) or 256 (4 KB).  When a segment is
// the whole row (W4 == SEG) the rows of consecutive slots are contiguous in HBM and a stage's TPS rows of one stream (or
// of the filter) travel as ONE bulk copy; otherwise one copy per row segment.

struct TtItem {
  int kt, ot, sp, r0, r1, c, s0, gx, fc, nst;
};

template <int ST>
__device__ __forceinline__ TtItem tt_decode(const MacArgs& a, int w, int ktiles) {
  TtItem it;
  it.kt = w % ktiles;
  w /= ktiles;
  it.ot = w % a.n_otiles;
  it.sp = w / a.n_otiles;
  it.r0 = it.sp * a.terms_per_split;
  it.r1 = min(it.r0 + a.terms_per_split, a.n_terms);
  it.c = it.ot % a.c_out;
  it.s0 = (it.ot / a.c_out) * ST;
  it.gx = (a.c_x == 1) ? 0 : it.c;
  it.fc = (a.c_f == 1) ? 0 : it.c;
  it.nst = min(ST, a.N - it.s0);
  return it;
}

// the two runs of consecutive slots of a term range (see k_fdl_mac_tile): run -> [jbeg, jend)
__device__ __forceinline__ bool tt_run(const MacArgs& a, const TtItem& it, int run, int* jbeg, int* jend) {
  const int rsk = a.skip - a.off;
  const int rb = run == 0 ? it.r0 : max(it.r0, rsk), re = run == 0 ? min(it.r1, rsk) : it.r1;
  if (rb >= re) return false;
  *jbeg = a.off + rb + (run ? a.nskip : 0);
  *jend = a.off + re + (run ? a.nskip : 0);
  return true;
}

template <int ST, int T, int TPS, int STAGES, int SEG>
__global__ void __launch_bounds__(SEG + 32) k_fdl_mac_tile_tma(const MacArgs a, const int n_items) {
  constexpr int kTtConsumers = SEG, kTtSegF4 = SEG, kTtSegBytes = SEG * 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int ROWS = TPS * (ST + 1) + (T - 1);       // + the window seed rows (first stage of a run)
  constexpr int STAGE_F4 = ROWS * kTtSegF4;
  float4* stage_base = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_F4 * 16);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ktiles = a.W4 / kTtSegF4;
  const bool contig = (ktiles == 1);
  const size_t rs = (size_t)a.W4;
  const size_t stream_stride = (size_t)a.c_x * a.R * rs;
  const int qb = a.R - 1 - a.head;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTtConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTtConsumers / 32) {
    // ===== producer warp: lane l issues row l of the stage =====
    const uint64_t pol_x = l2_policy_evict_first();
    const uint64_t pol_h = (ST > 1) ? l2_policy_evict_last() : pol_x;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const TtItem it = tt_decode<ST>(a, w, ktiles);
      const int kv0 = it.kt * kTtSegF4;
      const float4* xbase = a.fdl + ((size_t)(it.s0 * a.c_x + it.gx) * a.R) * rs + kv0;
      const float4* hfil = a.Hd + ((size_t)(__ldg(a.fmap + it.s0) * a.c_f + it.fc) * 2 * a.R) * rs + kv0;
      for (int run = 0; run < 2; ++run) {
        int jbeg, jend;
        if (!tt_run(a, it, run, &jbeg, &jend)) continue;
        for (int jb = jbeg; jb < jend; jb += TPS) {
          const int nterm = min(TPS, jend - jb);
          const bool first = (jb == jbeg);
          if (lane == 0) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], (uint32_t)((nterm * (it.nst + 1) + (first ? T - 1 : 0)) * kTtSegBytes));
          }
          __syncwarp();
          float4* sb = stage_base + (size_t)stage * STAGE_F4;
          // stage layout: [q = 0..ST-1 delay-line rows of stream q | q = ST filter rows][slot u], then the window seed
          if (contig) {
            if (lane <= ST && (lane == ST || lane < it.nst)) {
              const float4* src = (lane == ST) ? hfil + (size_t)(qb + jb) * rs
                                               : xbase + (size_t)jb * rs + (size_t)lane * stream_stride;
              bulk_g2s(sb + (size_t)lane * TPS * kTtSegF4, src, (uint32_t)(nterm * kTtSegBytes), &full[stage],
                       lane == ST ? pol_h : pol_x);
            } else if (first && lane == ST + 1) {
              int q = qb + jbeg - (T - 1);
              q += (q < 0) ? a.R : 0;      // rows q and q + R hold the same partition: shift the whole seed run
              bulk_g2s(sb + (size_t)(ST + 1) * TPS * kTtSegF4, hfil + (size_t)q * rs, (uint32_t)((T - 1) * kTtSegBytes),
                       &full[stage], pol_h);
            }
          } else if (lane < TPS * (ST + 1)) {
            const int q = lane / TPS, u = lane - q * TPS;   // row kind, slot within the stage
            if (u < nterm) {
              const int j = jb + u;
              if (q == ST) bulk_g2s(sb + (size_t)lane * kTtSegF4, hfil + (size_t)(qb + j) * rs, kTtSegBytes, &full[stage], pol_h);
              else if (q < it.nst)
                bulk_g2s(sb + (size_t)lane * kTtSegF4, xbase + (size_t)j * rs + (size_t)q * stream_stride, kTtSegBytes,
                         &full[stage], pol_x);
            }
          } else if (first && lane < ROWS) {
            const int w2 = lane - TPS * (ST + 1);                      // window seed: d = jbeg - (T-1) + w2
            int q = qb + jbeg - (T - 1);
            q += (q < 0) ? a.R : 0;
            bulk_g2s(sb + (size_t)lane * kTtSegF4, hfil + (size_t)(q + w2) * rs, kTtSegBytes, &full[stage], pol_h);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== consumer warps =====
    const int tid = threadIdx.x;   // float4 column within the segment
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const TtItem it = tt_decode<ST>(a, w, ktiles);
      const int kv = it.kt * kTtSegF4 + tid;
      const bool bin0 = (kv == 0);
      float4 acc[T][ST];
#pragma unroll
      for (int kp = 0; kp < T; ++kp)
#pragma unroll
        for (int t = 0; t < ST; ++t) acc[kp][t] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int run = 0; run < 2; ++run) {
        int jbeg, jend;
        if (!tt_run(a, it, run, &jbeg, &jend)) continue;
        int p0 = a.head - jbeg;
        p0 += (p0 < 0) ? a.R : 0;
        float4 hw[TPS + T - 1];     // hw[w] <-> d = jb - (T-1) + w
        for (int jb = jbeg; jb < jend; jb += TPS, p0 -= TPS) {
          const int nterm = min(TPS, jend - jb);
          mbar_wait(&full[stage], phase);
          const float4* sb = stage_base + (size_t)stage * STAGE_F4 + tid;
          if (jb == jbeg) {
#pragma unroll
            for (int w2 = 0; w2 < T - 1; ++w2) hw[w2] = sb[(size_t)(TPS * (ST + 1) + w2) * kTtSegF4];
          }
          const bool fast = (nterm == TPS && p0 + T - 1 <= a.P - 1 && it.nst == ST);
#pragma unroll
          for (int u = 0; u < TPS; ++u) {
            if (fast || u < nterm) {
              hw[T - 1 + u] = sb[(size_t)(ST * TPS + u) * kTtSegF4];
              float4 x[ST];
              float xim[ST], xre[ST];   // first complex: x.x as the imaginary sum uses it, -x.y as the real sum does
#pragma unroll
              for (int t = 0; t < ST; ++t) {
                x[t] = (fast || t < it.nst) ? sb[(size_t)(t * TPS + u) * kTtSegF4] : make_float4(0.f, 0.f, 0.f, 0.f);
                xim[t] = bin0 ? 0.f : x[t].x;   // packed bin 0 (lane kv = 0): two real products
                xre[t] = bin0 ? 0.f : -x[t].y;
              }
#pragma unroll
              for (int kp = 0; kp < T; ++kp) {
                if (fast || p0 - u + kp <= a.P - 1) {
                  const float4 h = hw[T - 1 + u - kp];
                  const float hsel = bin0 ? h.y : h.x;
#pragma unroll
                  for (int t = 0; t < ST; ++t) {
                    float4& c = acc[kp][t];
                    c.x = fmaf(x[t].x, h.x, c.x);
                    c.x = fmaf(xre[t], h.y, c.x);
                    c.y = fmaf(xim[t], h.y, c.y);
                    c.y = fmaf(x[t].y, hsel, c.y);
                    c.z = fmaf(x[t].z, h.z, c.z);
                    c.z = fmaf(-x[t].w, h.w, c.z);
                    c.w = fmaf(x[t].z, h.w, c.w);
                    c.w = fmaf(x[t].w, h.z, c.w);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int w2 = 0; w2 < T - 1; ++w2) hw[w2] = hw[w2 + TPS];
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
#pragma unroll
      for (int kp = 0; kp < T; ++kp)
#pragma unroll
        for (int t = 0; t < ST; ++t)
          if (t < it.nst) {
            const size_t prow = (size_t)it.sp * a.n_out + (it.s0 + t) * a.c_out + it.c;
            a.yspec[(size_t)((a.tile_set0 + kp) % a.tile_nsets) * a.tile_stride + prow * rs + kv] = acc[kp][t];
          }
    }
  }
}

struct TtVariant {
  int st, tile, tps, stages, seg;
  const void* func;
  int smem;
};
#define TT_V(ST, T, TPS, STAGES, SEG)                                                                 \
  {ST, T, TPS, STAGES, SEG, reinterpret_cast<const void*>(k_fdl_mac_tile_tma<ST, T, TPS, STAGES, SEG>), \
   STAGES * (TPS * (ST + 1) + (T - 1)) * SEG * 16 + 2 * STAGES * (int)sizeof(uint64_t)}
static const TtVariant kTtVariants[] = {
    TT_V(2, 4, 2, 3, 128), TT_V(2, 4, 2, 4, 128), TT_V(2, 4, 4, 3, 128), TT_V(1, 4, 4, 3, 128), TT_V(1, 4, 4, 4, 128),
    TT_V(2, 2, 2, 4, 128), TT_V(1, 2, 4, 4, 128),
    TT_V(2, 4, 2, 3, 256), TT_V(2, 4, 4, 2, 256), TT_V(2, 4, 4, 3, 256), TT_V(1, 4, 4, 3, 256), TT_V(1, 4, 4, 4, 256),
    TT_V(1, 4, 8, 2, 256), TT_V(2, 2, 4, 3, 256), TT_V(1, 2, 4, 4, 256),
};

static const TtVariant* tt_variant(int st, int tile, int tps, int stages, int seg) {
  for (const TtVariant& v : kTtVariants)
    if (v.st == st && v.tile == tile && v.tps == tps && v.stages == stages && v.seg == seg) return &v;
  return nullptr;
}

bool tile_tma_supported(int W4) { return W4 >= 128 && (W4 % 128) == 0; }

static int tt_seg(int W4) {
  int seg = (W4 % 256) == 0 ? 256 : 128;
  if (const char* e = getenv("PGX_TILE_SEG")) {
    const int v = atoi(e);
    if ((v == 128 || v == 256) && (W4 % v) == 0) seg = v;
  }
  return seg;
}

// default configuration per (shared filter?, tile); PGX_TILE_ST / PGX_TILE_TPS / PGX_TILE_STAGES / PGX_TILE_SEG override
bool tile_tma_config(bool shared_filter, int N, int W4, int tile, int* st, int* tps, int* stages, int* occupancy) {
  const int seg = tt_seg(W4);
  const int s0 = (shared_filter && N >= 2) ? 2 : 1, p0 = 4, g0 = 3;
  int s = s0, p = p0, g = g0;
  if (const char* e = getenv("PGX_TILE_ST")) {
    const int v = atoi(e);
    if (v == 1 || (shared_filter && N >= v)) s = v;
  }
  if (const char* e = getenv("PGX_TILE_TPS")) p = atoi(e);
  if (const char* e = getenv("PGX_TILE_STAGES")) g = atoi(e);
  const TtVariant* v = tt_variant(s, tile, p, g, seg);
  if (!v) {
    s = s0; p = p0; g = g0;
    v = tt_variant(s, tile, p, g, seg);
  }
  if (!v) return false;
  cudaFuncSetAttribute(v->func, cudaFuncAttributeMaxDynamicSharedMemorySize, v->smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, v->func, seg + 32, v->smem);
  *st = s; *tps = p; *stages = g; *occupancy = nb > 0 ? nb : 1;
  return true;
}

int tile_tma_ktiles(int W4) { return W4 / tt_seg(W4); }

bool describe_fdl_mac_tile_tma(const MacArgs& a, LaunchDesc* d, int* n_items) {
  const int seg = tt_seg(a.W4);
  const TtVariant* v = tt_variant(a.st, a.tile, a.tile_u, a.tile_stages, seg);
  if (!v) return false;
  *n_items = a.n_otiles * (a.W4 / seg) * a.n_split;
  const int grid = *n_items < a.persistent_ctas ? *n_items : a.persistent_ctas;
  d->func = v->func;
  d->grid = dim3((unsigned)grid);
  d->block = dim3(seg + 32);
  d->smem = (size_t)v->smem;
  return true;
}

void launch_fdl_mac_tile_tma(const MacArgs& a, cudaStream_t st) {
  LaunchDesc d;
  int n_items = 0;
  if (!describe_fdl_mac_tile_tma(a, &d, &n_items)) return;
  void* params[] = {const_cast<MacArgs*>(&a), &n_items};
  cudaLaunchKernel(d.func, d.grid, d.block, params, d.smem, st);
}

}  // namespace pgx
