// k_mac_tma.cu -- K3/K4, bulk-async (TMA) staged variant of the delay-line multiply-accumulate.
//
// Same arithmetic, term order and output layout as k_fdl_mac (k_mac.cu); what changes is how the rows
// reach the SM.  A persistent CTA = 1 producer warp + 8 consumer warps walks a static list of work items
// (out tile x bin tile x term split).  The producer streams 4 KB row segments of X (and H) from HBM into
// a ring of shared-memory stages with cp.async.bulk (SASS: UBLKCP), completion counted in bytes on an
// mbarrier per stage; X is tagged L2 evict-first (read once per step), H evict-last when it is shared by
// all streams.  Consumers wait on the stage's "full" barrier, read one float4 per thread per row
// (conflict-free), accumulate in registers and release the stage through its "empty" barrier.  The
// bytes in flight live in shared memory (STAGES x stage bytes per CTA), not in registers, so the CTA
// needs ~40 registers per thread and leaves room on the SM for the latency-critical FFT kernels.
// Used when a row segment is a full 4 KB (B >= 512); narrower rows stay on the LDG kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bulk.cuh"
#include "kernels.h"

namespace pgx {

static constexpr int kTmaConsumers = 256;             // 8 consumer warps, one float4 column each
static constexpr int kTmaThreads = kTmaConsumers + 32;
static constexpr int kSegF4 = 256;                    // float4 per row segment (4 KB)
static constexpr int kSegBytes = kSegF4 * 16;

__device__ __forceinline__ void cmac2t(float4& acc, const float4 x, const float4 h) {
  acc.x = fmaf(x.x, h.x, acc.x);
  acc.x = fmaf(-x.y, h.y, acc.x);
  acc.y = fmaf(x.x, h.y, acc.y);
  acc.y = fmaf(x.y, h.x, acc.y);
  acc.z = fmaf(x.z, h.z, acc.z);
  acc.z = fmaf(-x.w, h.w, acc.z);
  acc.w = fmaf(x.z, h.w, acc.w);
  acc.w = fmaf(x.w, h.z, acc.w);
}

// work item -> rows.  Shared between producer and consumers so both walk the same sequence.
struct TmaItem {
  int ot, kt, sp, r0, r1, c, s0, gx, fc, nst;
};

template <bool MIX, int ST>
__device__ __forceinline__ TmaItem decode_item(const MacArgs& a, int w, int ktiles) {
  TmaItem it;
  it.kt = w % ktiles;
  w /= ktiles;
  it.ot = w % a.n_otiles;
  it.sp = w / a.n_otiles;
  it.r0 = it.sp * a.terms_per_split;
  it.r1 = min(it.r0 + a.terms_per_split, a.n_terms);
  if (MIX) {
    it.c = it.ot;
    it.s0 = 0;
  } else {
    it.c = it.ot % a.c_out;
    it.s0 = (it.ot / a.c_out) * ST;
  }
  it.gx = (a.c_x == 1) ? 0 : it.c;
  it.fc = (a.c_f == 1) ? 0 : it.c;
  it.nst = (ST > 1) ? min(ST, a.N - it.s0) : 1;
  return it;
}

// ST streams per item share the filter rows; TPS terms per pipeline stage; STAGES stages.
template <bool MIX, int ST, int TPS, int STAGES>
__global__ void __launch_bounds__(kTmaThreads) k_fdl_mac_tma(const MacArgs a, const int n_items) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int ROWS = TPS * (ST + 1);                 // row segments per stage: TPS x (ST X rows + 1 H row)
  constexpr int STAGE_BYTES = ROWS * kSegBytes;
  float4* stage_base = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ktiles = a.W4 / kSegF4;
  const size_t rs = (size_t)a.W4;
  const size_t stream_stride = (size_t)a.c_x * a.R * rs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumers / 32) {
    // ===== producer warp =====
    const uint64_t pol_x = l2_policy_evict_first();
    const uint64_t pol_h = (ST > 1) ? l2_policy_evict_last() : pol_x;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const TmaItem it = decode_item<MIX, ST>(a, w, ktiles);
      const int kv0 = it.kt * kSegF4;
      const int fid0 = MIX ? 0 : __ldg(a.fmap + it.s0);
      for (int r = it.r0; r < it.r1; r += TPS) {
        const int nterm = min(TPS, it.r1 - r);
        if (lane == 0) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (uint32_t)(nterm * (it.nst + 1) * kSegBytes));
        }
        __syncwarp();
        // lanes 0 .. nterm*(nst+1)-1 each issue one row-segment copy
        const int u = lane / (ST + 1), q = lane - u * (ST + 1);  // term within stage, row within term (ST = H row)
        if (u < nterm && (q == ST || q < it.nst)) {
          const int rr = r + u;
          int s, jj;
          if (MIX) {
            s = rr / a.Pt;
            jj = rr - s * a.Pt;
          } else {
            s = it.s0;
            jj = rr;
          }
          int j = a.off + jj;
          j = (a.jfix >= 0) ? a.jfix : j + (j >= a.skip ? a.nskip : 0);
          float4* dst = stage_base + (size_t)stage * (STAGE_BYTES / 16) + (size_t)(u * (ST + 1) + q) * kSegF4;
          if (q == ST) {
            const int fid = MIX ? __ldg(a.fmap + s) : fid0;
            const float4* src = a.Hd + ((size_t)(fid * a.c_f + it.fc) * 2 * a.R + a.q0 + j) * rs + kv0;
            bulk_g2s(dst, src, kSegBytes, &full[stage], pol_h);
          } else {
            const float4* src = a.fdl + ((size_t)(s * a.c_x + it.gx) * a.R + j) * rs + (size_t)q * stream_stride + kv0;
            bulk_g2s(dst, src, kSegBytes, &full[stage], pol_x);
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===== consumer warps =====
    const int tid = threadIdx.x;  // float4 column within the segment
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const TmaItem it = decode_item<MIX, ST>(a, w, ktiles);
      const int kv = it.kt * kSegF4 + tid;
      const bool bin0 = (kv == 0);
      float4 acc[ST];
      float2 acc0[ST];
#pragma unroll
      for (int t = 0; t < ST; ++t) {
        acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc0[t] = make_float2(0.f, 0.f);
      }
      for (int r = it.r0; r < it.r1; r += TPS) {
        const int nterm = min(TPS, it.r1 - r);
        mbar_wait(&full[stage], phase);
        const float4* sb = stage_base + (size_t)stage * (STAGE_BYTES / 16);
#pragma unroll
        for (int u = 0; u < TPS; ++u) {
          if (u < nterm) {
            const float4 h = sb[(size_t)(u * (ST + 1) + ST) * kSegF4 + tid];
#pragma unroll
            for (int t = 0; t < ST; ++t) {
              if (t < it.nst) {
                const float4 x = sb[(size_t)(u * (ST + 1) + t) * kSegF4 + tid];
                cmac2t(acc[t], x, h);
                if (bin0) {
                  acc0[t].x = fmaf(x.x, h.x, acc0[t].x);
                  acc0[t].y = fmaf(x.y, h.y, acc0[t].y);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
#pragma unroll
      for (int t = 0; t < ST; ++t) {
        if (t < it.nst) {
          if (bin0) {
            acc[t].x = acc0[t].x;
            acc[t].y = acc0[t].y;
          }
          const int o = MIX ? it.c : (it.s0 + t) * a.c_out + it.c;
          a.yspec[((size_t)it.sp * a.n_out + o) * rs + kv] = acc[t];
        }
      }
    }
  }
}

template <bool MIX, int ST, int TPS, int STAGES>
static int tma_smem_bytes() {
  return STAGES * TPS * (ST + 1) * kSegBytes + 2 * STAGES * (int)sizeof(uint64_t);
}

template <bool MIX, int ST, int TPS, int STAGES>
static int tma_occupancy() {
  const int smem = tma_smem_bytes<MIX, ST, TPS, STAGES>();
  cudaFuncSetAttribute(k_fdl_mac_tma<MIX, ST, TPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fdl_mac_tma<MIX, ST, TPS, STAGES>, kTmaThreads, smem);
  return nb > 0 ? nb : 1;
}

// configurations: shared filter -> 4 streams per item, 1 term per stage (20 KB), 4 stages (80 KB / CTA);
// distinct / mix -> 1 stream, 2 terms per stage (16 KB), 5 stages (80 KB / CTA)
#define TMA_SHARED false, 4, 1, 4
#define TMA_CONV1 false, 1, 2, 5
#define TMA_MIX true, 1, 2, 5

bool tma_supported(int W4) { return W4 >= kSegF4 && (W4 % kSegF4) == 0; }

int tma_occupancy_of(bool mix, int st) {
  if (mix) return tma_occupancy<TMA_MIX>();
  if (st == 4) return tma_occupancy<TMA_SHARED>();
  return tma_occupancy<TMA_CONV1>();
}

void launch_fdl_mac_tma(const MacArgs& a, int persistent_ctas, cudaStream_t st) {
  const int n_items = a.n_otiles * (a.W4 / kSegF4) * a.n_split;
  const int grid = n_items < persistent_ctas ? n_items : persistent_ctas;
  if (a.mix == 1)
    k_fdl_mac_tma<TMA_MIX><<<grid, kTmaThreads, tma_smem_bytes<TMA_MIX>(), st>>>(a, n_items);
  else if (a.st == 4)
    k_fdl_mac_tma<TMA_SHARED><<<grid, kTmaThreads, tma_smem_bytes<TMA_SHARED>(), st>>>(a, n_items);
  else
    k_fdl_mac_tma<TMA_CONV1><<<grid, kTmaThreads, tma_smem_bytes<TMA_CONV1>(), st>>>(a, n_items);
}

}  // namespace pgx
