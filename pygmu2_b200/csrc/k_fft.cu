// k_fft.cu -- K1 (ingest + real FFT), K2 (inverse real FFT + emit) and filter preparation.
//
// K1/K2 replace np.fft.rfft / np.fft.irfft in the reference's overlap-save loop
// (convolve_pe.py:294-322): K1 also does what :294-310 (build [tail | segment | zeros]) and
// :325-336 (tail update) do, K2 what :321-322 (keep the valid output samples) does, plus the
// present-partition product X_t * H_0 of :314-317.
//
// One transform per N/8 threads (see fft.cuh); a CTA hosts FPB transforms.  K1 loads its window
// straight from global memory into registers (first pass needs no staging), K2 leaves its result in
// registers and writes only the samples that are new.  Kernels are instantiated per LOG2N = 4..13.
#include "bulk.cuh"
#include "fft.cuh"
#include "kernels.h"

namespace pgx {

template <int LOG2N>
__device__ __forceinline__ const float2* stage_twiddles(float2* smem, const float2* __restrict__ tw_global) {
  using C = FftCfg<LOG2N>;
  if constexpr (!C::SMEM_TW) {
    return tw_global;
  } else {
  for (int i = threadIdx.x; i < 2 * C::N; i += C::CTA) smem[i] = __ldg(tw_global + i);
    return smem;  // visible after the first __syncthreads(); the first pass uses no twiddles
  }
}

// K1's input stage: window w = [prev block (N) | open block cur[0:fill+take) | zeros], element q = (w[2q], w[2q+1]);
// the new samples are also stored into the open half of hist.  f = stream * c_x + channel.
template <int LOG2N>
__device__ __forceinline__ void ingest_window(const R2CArgs& a, const int64_t f, const int j, float2 (&v)[8]) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  // window w = [prev block (N) | open block cur[0:m_new) | zeros]; element q is (w[2q], w[2q+1])
  const int s = (int)f / a.c_x, cx = (int)f - s * a.c_x;
  float* cur = a.hist + ((size_t)f * 2 + a.half) * N;
  const float* prev = a.hist + ((size_t)f * 2 + (a.half ^ 1)) * N;
  const int m_new = a.fill + a.take;
  const float* xs = a.x + (int64_t)s * a.xs + (a.mixdown ? 0 : (int64_t)cx * a.xc);
  if (a.fast) {
    // a whole block of contiguous, 8-byte aligned samples: elements 0..N/2-1 (m < 4) are the previous block,
    // N/2..N-1 (m >= 4) the new one -- vector loads, and the new block goes to hist as vector stores
    const float* xn = xs + a.x_off;
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = *reinterpret_cast<const float2*>(prev + 2 * (j + m * T8));
#pragma unroll
    for (int m = 4; m < 8; ++m) v[m] = __ldg(reinterpret_cast<const float2*>(xn + 2 * (j + m * T8) - N));
#pragma unroll
    for (int m = 4; m < 8; ++m) *reinterpret_cast<float2*>(cur + 2 * (j + m * T8) - N) = v[m];
  } else
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int i0 = 2 * (j + m * T8);
    if (i0 < N) {
      v[m] = *reinterpret_cast<const float2*>(prev + i0);
    } else {
      const int c0 = i0 - N;
      float e[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + h;
        float val = 0.f;
        if (c < a.fill) {
          val = cur[c];                               // already ingested by an earlier partial pull
        } else if (c < m_new) {
          const int64_t off = (int64_t)(a.x_off + c - a.fill) * a.xi;
          if (a.mixdown) {
            float acc = 0.f;
            for (int ch = 0; ch < a.c_in; ++ch) acc += xs[off + (int64_t)ch * a.xc];
            val = acc / (float)a.c_in;
          } else {
            val = xs[off];
          }
          cur[c] = val;                               // each sample has exactly one owner thread
        }
        e[h] = val;
      }
      v[m] = make_float2(e[0], e[1]);
    }
  }
}

// MODE 0: stream ingest -> one delay-line row.  MODE 1: filter partition -> two (reversed, doubled) rows.
template <int LOG2N, int MODE>
__global__ void __launch_bounds__(FftCfg<LOG2N>::CTA) k_r2c(const R2CArgs a, const FilterPrepArgs fp) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  extern __shared__ float2 sm[];
  const float2* tw = stage_twiddles<LOG2N>(sm, MODE == 0 ? a.tw : fp.tw);
  float2* bufs = sm + (C::SMEM_TW ? 2 * N : 0);
  const int g = threadIdx.x / T8, j = threadIdx.x - g * T8;
  float2* sA = bufs + (size_t)g * 2 * C::PADN;
  float2* sB = sA + C::PADN;
  const int64_t f = (int64_t)blockIdx.x * C::FPB + g;
  const int64_t total = (MODE == 0) ? (int64_t)a.n_fft : (int64_t)fp.n_rows * fp.P;
  const bool active = f < total;

  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = make_float2(0.f, 0.f);
  int frow = 0, fpart = 0;
  if (active) {
    if (MODE == 0) {
      ingest_window<LOG2N>(a, f, j, v);
    } else {
      frow = (int)(f / fp.P);
      fpart = (int)(f - (int64_t)frow * fp.P);
      const float* h = fp.h + (size_t)frow * fp.L;
      const int base = fpart * N;  // window = [partition (N taps) | zeros (N)]
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int i0 = 2 * (j + m * T8);
        if (i0 < N) {
          if (base + i0 < fp.L) v[m].x = h[base + i0];
          if (base + i0 + 1 < fp.L) v[m].y = h[base + i0 + 1];
        }
      }
    }
  }

  fft_passes<LOG2N, false>(v, sA, sB, 0, j, tw);

  // split step needs Z[N-k]: exchange through shared memory, natural order and NOT padded: the natural-order
  // writes and the mirrored reads are both contiguous runs, conflict-free as they are (with the pad of the FFT
  // passes a mirrored run straddles a pad slot and two lanes meet in one bank)
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
  float2 tws[8];  // one table read (requested before the barrier), the other seven by constant 16th roots
  split_twiddles(tw[j], tws);
  __syncthreads();
  if (active) {
    float2 *row0, *row1 = nullptr;
    float scale = 1.f;
    if (MODE == 0) {
      row0 = a.fdl + ((size_t)f * a.R + a.slot) * N;
    } else {
      // reversed + doubled layout: partition p at rows R-1-p and 2R-1-p (see k_mac.cu); 1/N of the inverse
      // transform is folded in here
      row0 = fp.Hd + ((size_t)frow * 2 * fp.R + (fp.R - 1 - fpart)) * N;
      row1 = row0 + (size_t)fp.R * N;
      scale = 1.0f / (float)N;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + m * T8;
      float2 o;
      if (k == 0) {
        o = make_float2(v[m].x + v[m].y, v[m].x - v[m].y);
      } else {
        o = r2c_bin(v[m], sA[N - k], tws[m]);
      }
      if (MODE == 1) {
        o.x *= scale;
        o.y *= scale;
        row1[k] = o;
      }
      row0[k] = o;
    }
  }
}

// K2's output stage.  o = stream * c_out + channel.
template <int LOG2N>
__device__ __forceinline__ void emit_block(const C2RArgs& a, const int o, const int j, const float2 (&v)[8]) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  // overlap-save: the open block's output samples live at window positions [N+fill, N+fill+take);
  // element q = j + m*T8 carries window samples 2q (re) and 2q+1 (im)
  const int s = o / a.c_out, c = o - s * a.c_out;
  float* y = a.y + (int64_t)s * a.ys + (int64_t)c * a.yc;
  const float* xd = a.xdry ? a.xdry + (int64_t)s * a.xs + (int64_t)c * a.xc : nullptr;
  const float* ad = a.add ? a.add + (int64_t)s * a.as + (int64_t)c * a.ac : nullptr;  // tail level's contribution
  const bool gains = (a.wet != 1.0f) || xd;
  const int lo = N + a.fill, hi = lo + a.take;
  if (a.fast) {  // a whole block into contiguous, 8-byte aligned output: elements N/2.. (m >= 4), vector stores
    float* yo = y + a.y_off;
#pragma unroll
    for (int m = 4; m < 8; ++m) {
      const int i = 2 * (j + m * T8) - N;
      float2 val = v[m];
      if (ad) {
        const float2 t = *reinterpret_cast<const float2*>(ad + a.y_off + i);
        val.x += t.x;
        val.y += t.y;
      }
      if (gains) {
        val.x = __fmul_rn(val.x, a.wet);
        val.y = __fmul_rn(val.y, a.wet);
        if (xd) {
          const float2 d = __ldg(reinterpret_cast<const float2*>(xd + a.x_off + i));
          val.x = __fadd_rn(__fmul_rn(d.x, a.dry), val.x);
          val.y = __fadd_rn(__fmul_rn(d.y, a.dry), val.y);
        }
      }
      *reinterpret_cast<float2*>(yo + i) = val;
    }
  } else
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int i0 = 2 * (j + m * T8);
    const float e[2] = {v[m].x, v[m].y};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = i0 + h;
      if (i >= lo && i < hi) {
        float val = e[h];
        if (ad) val += ad[(int64_t)(a.y_off + i - lo) * a.ai];
        if (gains) {
          val = __fmul_rn(val, a.wet);
          if (xd) val = __fadd_rn(__fmul_rn(xd[(int64_t)(a.x_off + i - lo) * a.xi], a.dry), val);
        }
        y[(int64_t)(a.y_off + i - lo) * a.yi] = val;
      }
    }
  }
  }

// PART: split partials are summed in (P > 1 or mix mode); the P = 1 conv instantiation carries no such registers
template <int LOG2N, bool PART>
__global__ void __launch_bounds__(FftCfg<LOG2N>::CTA, FftCfg<LOG2N>::CTA == 512 ? 3 : 1) k_c2r(const C2RArgs a) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  extern __shared__ float2 sm[];
  const float2* tw = stage_twiddles<LOG2N>(sm, a.tw);
  float2* bufs = sm + (C::SMEM_TW ? 2 * N : 0);
  const int g = threadIdx.x / T8, j = threadIdx.x - g * T8;
  float2* sA = bufs + (size_t)g * 2 * C::PADN;
  float2* sB = sA + C::PADN;
  const int o = blockIdx.x * C::FPB + g;
  const bool active = o < a.n_out;

  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = make_float2(0.f, 0.f);
  if (active) {
    const float2 *xrow = nullptr, *hrow = nullptr;
    if (a.fdl) {  // conv mode: present term = delay-line slot `head` x filter partition 0 (row R-1 of Hd)
      const int s = o / a.c_out, c = o - s * a.c_out;
      const int gx = (a.c_x == 1) ? 0 : c, fc = (a.c_f == 1) ? 0 : c;
      xrow = a.fdl + ((size_t)(s * a.c_x + gx) * a.R + a.head) * N;
      hrow = a.Hd + ((size_t)(__ldg(a.fmap + s) * a.c_f + fc) * 2 * a.R + (a.R - 1)) * N;
    }
    if (xrow) {  // all 16 loads in flight before the first product
      float2 xx[8], hh[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        xx[m] = __ldcg(xrow + j + m * T8);   // written by K1 moments ago: L2
        hh[m] = __ldg(hrow + j + m * T8);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) v[m] = cmul(xx[m], hh[m]);
      if (j == 0) v[0] = make_float2(xx[0].x * hh[0].x, xx[0].y * hh[0].y);  // packed bin 0: two real bins
      // time-tiled banks: the rows committed after the tiled pass that covers this block, X[head - r] * H_r
      for (int r = 1; r <= a.n_recent; ++r) {
        int sl = a.head - r;
        sl += (sl < 0) ? a.R : 0;
        const float2* xr = xrow + ((ptrdiff_t)sl - a.head) * N;
        const float2* hr = hrow - (size_t)r * N;             // partition r lives r rows below partition 0 (row R-1)
        float2 x2[8], h2[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          x2[m] = __ldcg(xr + j + m * T8);
          h2[m] = __ldg(hr + j + m * T8);
        }
        const float2 b0 = make_float2(fmaf(x2[0].x, h2[0].x, v[0].x), fmaf(x2[0].y, h2[0].y, v[0].y));
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          v[m].x = fmaf(x2[m].x, h2[m].x, fmaf(-x2[m].y, h2[m].y, v[m].x));
          v[m].y = fmaf(x2[m].x, h2[m].y, fmaf(x2[m].y, h2[m].x, v[m].y));
        }
        if (j == 0) v[0] = b0;
      }
    }
    // split partials: 4 splits x 8 bins = 32 independent loads in flight per thread (fixed summation order)
    auto add_partials = [&](const float2* __restrict__ part, const int n) {
      const float2* p0 = part + (size_t)o * N + j;
      const size_t ss = (size_t)a.n_out * N;
      int sp = 0;
      for (; sp + 4 <= n; sp += 4) {
        float2 t[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int m = 0; m < 8; ++m) t[u][m] = p0[(size_t)(sp + u) * ss + m * T8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            v[m].x += t[u][m].x;
            v[m].y += t[u][m].y;
          }
      }
      for (; sp < n; ++sp) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float2 t = p0[(size_t)sp * ss + m * T8];
          v[m].x += t.x;
          v[m].y += t.y;
        }
      }
    };
    if (PART) {
      add_partials(a.yspec, a.n_split);
      add_partials(a.ynow, a.n_split_now);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
  }
  float2 tws[8];
  float2 tw0 = make_float2(1.f, 0.f);
  if (!C::SMEM_TW) tw0 = __ldg(tw + j);  // global table: requested before the barrier
  __syncthreads();
  if (C::SMEM_TW) tw0 = tw[j];
  split_twiddles(tw0, tws);
  if (active) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + m * T8;
      if (k == 0) {
        v[m] = make_float2(0.5f * (v[m].x + v[m].y), 0.5f * (v[m].x - v[m].y));
      } else {
        v[m] = c2r_bin(v[m], sA[N - k], tws[m]);
      }
    }
  }

  fft_passes<LOG2N, true>(v, sA, sB, 1, j, tw);  // first exchange goes to sB: sA may still be read above

  if (active) {
    emit_block<LOG2N>(a, o, j, v);
  }
}

// K1 + K2 fused for single-partition banks (P = 1, conv mode): ingest -> forward FFT -> split -> X * H ->
// merge -> inverse FFT -> emit, one kernel, one transform group per (stream, source channel).  The spectrum
// never goes to the delay line (nothing ever reads it back when P = 1): per stream-step the traffic is
// 2B*4 (window) + B*4 (hist) + B*8 (H) + B*4 (y) instead of that plus a B*8 row written and read again.
// FAN: a mono source fanned out to c_out filter channels (convolve_pe.py:300-310): one forward transform,
// c_out inverse transforms.
// PAST: the bank has a few partitions (2 <= P <= kFusedMaxP, conv mode): the same kernel also writes X to its
// delay-line row and adds the past partitions sum_{p>=1} X_{t-p} H_p, read straight from the stream's own ring
// rows (written by earlier launches of this kernel on the same stream) -- one launch per block step instead of
// K1 + background pass + K2 on three streams, which is what bounds short filters and the head level of a
// two-level bank.
template <int LOG2N, bool FAN, bool PAST>
__global__ void __launch_bounds__(FftCfg<LOG2N>::CTA, FftCfg<LOG2N>::CTA == 512 ? 2 : 1)
    k_conv1(const R2CArgs a, const C2RArgs k) {
  using C = FftCfg<LOG2N>;
  constexpr int N = C::N, T8 = C::T8;
  extern __shared__ float2 sm[];
  const float2* tw = stage_twiddles<LOG2N>(sm, a.tw);
  float2* bufs = sm + (C::SMEM_TW ? 2 * N : 0);
  const int g = threadIdx.x / T8, j = threadIdx.x - g * T8;
  float2* sA = bufs + (size_t)g * 2 * C::PADN;
  float2* sB = sA + C::PADN;
  const int64_t f = (int64_t)blockIdx.x * C::FPB + g;
  const bool active = f < (int64_t)a.n_fft;
  // One transform per CTA and one filter row per transform: the row (B*8 bytes of H) is fetched by ONE bulk-async
  // copy issued before the forward transform starts and waited for on an mbarrier just before the product, so
  // its HBM latency hides behind the forward FFT and no load instruction is spent on it.
  constexpr bool HPRE = !FAN && C::FPB == 1;
  float2* hbuf = bufs + 2 * C::PADN;                                  // [N] staged filter row (HPRE only)
  uint64_t* hbar = reinterpret_cast<uint64_t*>(hbuf + N);
  if (HPRE && threadIdx.x == 0) {
    mbar_init(hbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (active) {
      const int s0 = (int)(f / a.c_x), c0 = (int)(f - (int64_t)s0 * a.c_x);
      const float2* hrow = k.Hd + ((size_t)(__ldg(k.fmap + s0) * k.c_f + ((k.c_f == 1) ? 0 : c0)) * 2 * k.R + (k.R - 1)) * N;
      mbar_expect_tx(hbar, (uint32_t)(N * sizeof(float2)));
      bulk_g2s(hbuf, hrow, (uint32_t)(N * sizeof(float2)), hbar, l2_policy_evict_first());
    }
  }

  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = make_float2(0.f, 0.f);
  if (active) ingest_window<LOG2N>(a, f, j, v);
  fft_passes<LOG2N, false>(v, sA, sB, 0, j, tw);

  // split: packed half spectrum X[k], k = j + m*T8, from Z[k] and Z[N-k]
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
  float2 tws[8];
  split_twiddles(tw[j], tws);
  __syncthreads();
  float2 X[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int kk = j + m * T8;
    X[m] = (kk == 0) ? make_float2(v[m].x + v[m].y, v[m].x - v[m].y) : r2c_bin(v[m], sA[N - kk], tws[m]);
  }

  if (PAST && active) {  // the open block's spectrum goes to its ring row for the steps to come
    float2* row = a.fdl + ((size_t)f * a.R + a.slot) * N;
#pragma unroll
    for (int m = 0; m < 8; ++m) row[j + m * T8] = X[m];
  }
  const int s = active ? (int)(f / a.c_x) : 0, cx = active ? (int)(f - (int64_t)s * a.c_x) : 0;
  const int n_c = FAN ? k.c_out : 1;
  for (int ci = 0; ci < n_c; ++ci) {
    const int c = FAN ? ci : cx;
    const int fc = (k.c_f == 1) ? 0 : c;
    __syncthreads();  // every read of sA / sB above (or by the previous channel's inverse passes) is done
    if (active) {
      float2 hh[8];
      if (HPRE) {
        mbar_wait(hbar, 0);  // the staged row has landed (the barriers above made the mbarrier's init visible)
#pragma unroll
        for (int m = 0; m < 8; ++m) hh[m] = hbuf[j + m * T8];
      } else {
        const float2* hrow = k.Hd + ((size_t)(__ldg(k.fmap + s) * k.c_f + fc) * 2 * k.R + (k.R - 1)) * N;
#pragma unroll
        for (int m = 0; m < 8; ++m) hh[m] = __ldg(hrow + j + m * T8);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) v[m] = cmul(X[m], hh[m]);
      float2 b0 = make_float2(X[0].x * hh[0].x, X[0].y * hh[0].y);  // packed bin 0: two real bins
      if (PAST) {
        const float2* xring = k.fdl + (size_t)f * k.R * N + j;
        const float2* hring = k.Hd + ((size_t)(__ldg(k.fmap + s) * k.c_f + fc) * 2 * k.R + k.q0) * N + j;
        for (int jj = 0; jj < k.n_past; ++jj) {
          int slot = k.p_off + jj;
          slot += (slot >= k.p_skip) ? k.p_nskip : 0;
          float2 xx[8], hp[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            xx[m] = __ldcg(xring + (size_t)slot * N + m * T8);
            hp[m] = __ldg(hring + (size_t)slot * N + m * T8);
          }
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            v[m].x = fmaf(xx[m].x, hp[m].x, fmaf(-xx[m].y, hp[m].y, v[m].x));
            v[m].y = fmaf(xx[m].x, hp[m].y, fmaf(xx[m].y, hp[m].x, v[m].y));
          }
          b0.x = fmaf(xx[0].x, hp[0].x, b0.x);
          b0.y = fmaf(xx[0].y, hp[0].y, b0.y);
        }
      }
      if (j == 0) v[0] = b0;
      if (PAST && k.n_split > 0) {
        // past partitions summed OUTSIDE the kernel (the background pass of a many-partition bank, folded to
        // n_split <= 8 rows): K2's job, done here so that a small bank's step is one launch on the latency chain
        const float2* p0 = k.yspec + (size_t)(s * k.c_out + c) * N + j;
        const size_t ss = (size_t)k.n_out * N;
        for (int sp = 0; sp < k.n_split; ++sp) {
          float2 t[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) t[m] = p0[(size_t)sp * ss + m * T8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            v[m].x += t[m].x;
            v[m].y += t[m].y;
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int kk = j + m * T8;
        v[m] = (kk == 0) ? make_float2(0.5f * (v[m].x + v[m].y), 0.5f * (v[m].x - v[m].y))
                         : c2r_bin(v[m], sA[N - kk], tws[m]);
      }
    }
    fft_passes<LOG2N, true>(v, sA, sB, 1, j, tw);  // first exchange goes to sB: sA may still be read above
    if (active) emit_block<LOG2N>(k, s * k.c_out + c, j, v);
  }
}

// K1 + present-slot accumulate fused for single-partition MIXES of mono transforms (SpatialHRTF banks, C3): a CTA
// holds G sources; each group ingests and transforms its source, multiplies the spectrum by that source's
// filter row per output channel (the HRTF pair chosen by fmap) and the G products are summed through shared
// memory -- FFT, HRTF multiply and the MixPE sum over the CTA's sources in one kernel.  Writes one partial row per
// (CTA, channel) that K2 folds; replaces k_r2c + k_fdl_mac<MIX> of the three-kernel step.
template <int LOG2N, bool WIDE = false>
struct Mix1Cfg {
  using C = FftCfg<LOG2N>;
  // sources per CTA.  Narrow (the latency case, e.g. the named 256-source mix): at most 512 threads, so that a thread
  // may keep the filter rows of two output channels in flight in registers behind the forward transform (128 registers
  // per thread) and a small mix spreads over more SMs.  WIDE (thousands of sources: a throughput case): up to 1024
  // threads, half as many partial rows to write and fold, no register prefetch (64 registers per thread).
  static constexpr int kMaxT = WIDE ? 1024 : 512;
  static constexpr int G = (kMaxT / C::T8) < 16 ? ((kMaxT / C::T8) < 1 ? 1 : (kMaxT / C::T8)) : 16;
  static constexpr int CTA = G * C::T8;
  static constexpr int SMEM_BYTES = (C::SMEM_TW ? 2 * C::N : 0) * 8 + G * 2 * C::PADN * 8;
};

// LAST: the CTA that finishes last (ticket counter) also folds the partial rows of all CTAs, runs the c_out
// inverse transforms and emits -- the whole block step of a small mix is then ONE launch.
// The kernel is a latency chain (ncu, profiles/r02_mix1_c3_ncu_full.txt: 16 CTAs, 4.1 long-scoreboard and 2.7 barrier
// stall cycles per issue), so its global-memory round trips are overlapped instead of queued: the filter-map entry and
// the filter rows of the first two output channels are requested BEFORE the forward transform and consumed after
// it, and the last CTA folds the partial rows with all its thread groups at once (one round trip per channel) instead
// of two groups walking them four rows at a time.
template <int LOG2N, bool LAST, bool WIDE = false>
__global__ void __launch_bounds__(Mix1Cfg<LOG2N, WIDE>::CTA) k_mix1(const R2CArgs a, const C2RArgs k, float2* __restrict__ ynow,
                                                                    unsigned int* __restrict__ ticket) {
  using C = FftCfg<LOG2N>;
  using M = Mix1Cfg<LOG2N, WIDE>;
  constexpr int N = C::N, T8 = C::T8;
  extern __shared__ float2 sm[];
  const float2* tw = stage_twiddles<LOG2N>(sm, a.tw);
  float2* bufs = sm + (C::SMEM_TW ? 2 * N : 0);
  const int g = threadIdx.x / T8, j = threadIdx.x - g * T8;
  float2* sA = bufs + (size_t)g * 2 * C::PADN;
  float2* sB = sA + C::PADN;
  const int64_t f = (int64_t)blockIdx.x * M::G + g;   // c_x == 1: transform f is stream f
  const bool active = f < (int64_t)a.n_fft;
  const int fid = active ? __ldg(k.fmap + (int)f) : 0;

  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = make_float2(0.f, 0.f);
  if (active) ingest_window<LOG2N>(a, f, j, v);
  // the source's filter rows of the first two output channels: in flight while the forward transform runs
  float2 hpre[WIDE ? 1 : 2][8];
  if constexpr (!WIDE) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int fc = (k.c_f == 1) ? 0 : c;
      const float2* hrow = k.Hd + ((size_t)(fid * k.c_f + fc) * 2 * k.R + (k.R - 1)) * N + j;
#pragma unroll
      for (int m = 0; m < 8; ++m) hpre[c][m] = (active && c < k.c_out) ? __ldg(hrow + m * T8) : make_float2(0.f, 0.f);
    }
  }
  fft_passes<LOG2N, false>(v, sA, sB, 0, j, tw);
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
  float2 tws[8];
  split_twiddles(tw[j], tws);
  __syncthreads();
  float2 X[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int kk = j + m * T8;
    X[m] = (kk == 0) ? make_float2(v[m].x + v[m].y, v[m].x - v[m].y) : r2c_bin(v[m], sA[N - kk], tws[m]);
  }
  auto channel = [&](const int c, const float2 (&hh)[8]) {
    __syncthreads();  // the reads of sA above / by the previous channel's sum are done
    float2 y[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) y[m] = active ? cmul(X[m], hh[m]) : make_float2(0.f, 0.f);
    if (active && j == 0) y[0] = make_float2(X[0].x * hh[0].x, X[0].y * hh[0].y);  // packed bin 0: two real bins
#pragma unroll
    for (int m = 0; m < 8; ++m) sA[j + m * T8] = y[m];
    __syncthreads();
    // the MixPE sum over this CTA's sources, in source order
    float2* out = ynow + ((size_t)blockIdx.x * k.c_out + c) * N;
    for (int bin = threadIdx.x; bin < N; bin += M::CTA) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < M::G; ++gg) {
        const float2 t = bufs[(size_t)gg * 2 * C::PADN + bin];
        acc.x += t.x;
        acc.y += t.y;
      }
      out[bin] = acc;
    }
  };
  if constexpr (!WIDE) {
    channel(0, hpre[0]);
    if (k.c_out > 1) channel(1, hpre[1]);
  }
  for (int c = WIDE ? 0 : 2; c < k.c_out; ++c) {
    const int fc = (k.c_f == 1) ? 0 : c;
    const float2* hrow = k.Hd + ((size_t)(fid * k.c_f + fc) * 2 * k.R + (k.R - 1)) * N + j;
    float2 hh[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) hh[m] = active ? __ldg(hrow + m * T8) : make_float2(0.f, 0.f);
    channel(c, hh);
  }
  if (!LAST) return;
  // ---- last CTA: fold + inverse transforms + emit (K2's work), groups 0..c_out-1 carry one output channel each
  __shared__ unsigned int s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch (launches of one bank are stream-ordered)
  __threadfence();
  const bool emit = g < k.c_out;
  const int nr = (int)gridDim.x;
  const size_t rs = (size_t)k.c_out * N;
  float2 mine[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) mine[m] = make_float2(0.f, 0.f);
  for (int c = 0; c < k.c_out; ++c) {
    // group g sums rows g, g+G, ... of channel c (rows written by other CTAs: read through L2), every load of a batch
    // of 4 rows in flight at once; then the G group sums are added in group order -- a fixed order, so the result
    // does not depend on which CTA happened to be last
    float2 acc[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = make_float2(0.f, 0.f);
    const float2* part = ynow + (size_t)c * N + j;
    int r = g;
    for (; r + 3 * M::G < nr; r += 4 * M::G) {
      float2 t[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int m = 0; m < 8; ++m) t[u][m] = __ldcg(part + (size_t)(r + u * M::G) * rs + m * T8);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          acc[m].x += t[u][m].x;
          acc[m].y += t[u][m].y;
        }
    }
    for (; r < nr; r += M::G) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const float2 t = __ldcg(part + (size_t)r * rs + m * T8);
        acc[m].x += t.x;
        acc[m].y += t.y;
      }
    }
    __syncthreads();  // the previous channel's group sums have been read
#pragma unroll
    for (int m = 0; m < 8; ++m) bufs[(size_t)g * 2 * C::PADN + j + m * T8] = acc[m];
    __syncthreads();
    if (g == c) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        float2 sacc = make_float2(0.f, 0.f);
#pragma unroll
        for (int gg = 0; gg < M::G; ++gg) {
          const float2 t = bufs[(size_t)gg * 2 * C::PADN + j + m * T8];
          sacc.x += t.x;
          sacc.y += t.y;
        }
        mine[m] = sacc;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = mine[m];
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) sA[j + m * T8] = v[m];
  __syncthreads();
  if (emit) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int kk = j + m * T8;
      v[m] = (kk == 0) ? make_float2(0.5f * (v[m].x + v[m].y), 0.5f * (v[m].x - v[m].y))
                       : c2r_bin(v[m], sA[N - kk], tws[m]);
    }
  }
  fft_passes<LOG2N, true>(v, sA, sB, 1, j, tw);
  if (emit) emit_block<LOG2N>(k, g, j, v);
}

// ---- launchers ---------------------------------------------------------------------------------
// cudaFuncSetAttribute is per device: remember it per (kernel instantiation, device), not once per process
static bool need_smem_attr(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

int fft_smem_bytes(int B) {
  switch (ilog2(B)) {
#define PGX_CASE(L) case L: return FftCfg<L>::SMEM_BYTES;
    PGX_CASE(4) PGX_CASE(5) PGX_CASE(6) PGX_CASE(7) PGX_CASE(8) PGX_CASE(9) PGX_CASE(10) PGX_CASE(11) PGX_CASE(12) PGX_CASE(13)
#undef PGX_CASE
    default: return 0;
  }
}

template <int LOG2N, int MODE>
static void describe_r2c_t(int64_t total, LaunchDesc* d) {
  using C = FftCfg<LOG2N>;
  static bool attr_done[64] = {};
  if (C::SMEM_BYTES > 48 * 1024 && need_smem_attr(attr_done))
    cudaFuncSetAttribute(k_r2c<LOG2N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  d->func = reinterpret_cast<const void*>(k_r2c<LOG2N, MODE>);
  d->grid = dim3((unsigned)((total + C::FPB - 1) / C::FPB));
  d->block = dim3(C::CTA);
  d->smem = C::SMEM_BYTES;
}

template <int LOG2N, bool PART>
static void describe_c2r_t(const C2RArgs& a, LaunchDesc* d) {
  using C = FftCfg<LOG2N>;
  static bool attr_done[64] = {};
  if (C::SMEM_BYTES > 48 * 1024 && need_smem_attr(attr_done))
    cudaFuncSetAttribute(k_c2r<LOG2N, PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  d->func = reinterpret_cast<const void*>(k_c2r<LOG2N, PART>);
  d->grid = dim3((unsigned)((a.n_out + C::FPB - 1) / C::FPB));
  d->block = dim3(C::CTA);
  d->smem = C::SMEM_BYTES;
}

template <int LOG2N, bool FAN, bool PAST>
static void describe_conv1_t(const R2CArgs& a, LaunchDesc* d) {
  using C = FftCfg<LOG2N>;
  // + the staged filter row and its mbarrier when the CTA holds one transform (see HPRE in the kernel)
  constexpr int smem = C::SMEM_BYTES + ((!FAN && C::FPB == 1) ? C::N * 8 + 16 : 0);
  static bool attr_done[64] = {};
  if (smem > 48 * 1024 && need_smem_attr(attr_done))
    cudaFuncSetAttribute(k_conv1<LOG2N, FAN, PAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  d->func = reinterpret_cast<const void*>(k_conv1<LOG2N, FAN, PAST>);
  d->grid = dim3((unsigned)((a.n_fft + C::FPB - 1) / C::FPB));
  d->block = dim3(C::CTA);
  d->smem = smem;
}

#define PGX_DISPATCH(LOG, CALL)                                                                              \
  switch (LOG) {                                                                                             \
    case 4: { constexpr int L_ = 4; CALL; } break;                                                           \
    case 5: { constexpr int L_ = 5; CALL; } break;                                                           \
    case 6: { constexpr int L_ = 6; CALL; } break;                                                           \
    case 7: { constexpr int L_ = 7; CALL; } break;                                                           \
    case 8: { constexpr int L_ = 8; CALL; } break;                                                           \
    case 9: { constexpr int L_ = 9; CALL; } break;                                                           \
    case 10: { constexpr int L_ = 10; CALL; } break;                                                         \
    case 11: { constexpr int L_ = 11; CALL; } break;                                                         \
    case 12: { constexpr int L_ = 12; CALL; } break;                                                         \
    case 13: { constexpr int L_ = 13; CALL; } break;                                                         \
    default: break;                                                                                          \
  }

static void launch_desc(const LaunchDesc& d, void** params, cudaStream_t st) {
  if (d.func) cudaLaunchKernel(d.func, d.grid, d.block, params, d.smem, st);
}

bool describe_r2c_ingest(const R2CArgs& a, LaunchDesc* d) {
  d->func = nullptr;
  PGX_DISPATCH(ilog2(a.B), (describe_r2c_t<L_, 0>(a.n_fft, d)));
  return d->func != nullptr;
}

void launch_r2c_ingest(const R2CArgs& a, cudaStream_t st) {
  FilterPrepArgs fp{};
  LaunchDesc d;
  describe_r2c_ingest(a, &d);
  void* params[] = {const_cast<R2CArgs*>(&a), &fp};
  launch_desc(d, params, st);
}

void launch_filter_prep(const FilterPrepArgs& fp, cudaStream_t st) {
  R2CArgs a{};
  LaunchDesc d;
  PGX_DISPATCH(ilog2(fp.B), (describe_r2c_t<L_, 1>((int64_t)fp.n_rows * fp.P, &d)));
  void* params[] = {&a, const_cast<FilterPrepArgs*>(&fp)};
  launch_desc(d, params, st);
}

bool describe_conv1(const R2CArgs& a, const C2RArgs& k, LaunchDesc* d) {
  d->func = nullptr;
  if (describe_conv1_r16(a, k, d)) return true;
  const bool fan = (a.c_x == 1 && k.c_out > 1);
  if (k.n_past > 0 || k.n_split > 0) {
    if (fan) { PGX_DISPATCH(ilog2(a.B), (describe_conv1_t<L_, true, true>(a, d))); }
    else { PGX_DISPATCH(ilog2(a.B), (describe_conv1_t<L_, false, true>(a, d))); }
  } else {
    if (fan) { PGX_DISPATCH(ilog2(a.B), (describe_conv1_t<L_, true, false>(a, d))); }
    else { PGX_DISPATCH(ilog2(a.B), (describe_conv1_t<L_, false, false>(a, d))); }
  }
  return d->func != nullptr;
}

void launch_conv1(const R2CArgs& a, const C2RArgs& k, cudaStream_t st) {
  LaunchDesc d;
  describe_conv1(a, k, &d);
  void* params[] = {const_cast<R2CArgs*>(&a), const_cast<C2RArgs*>(&k)};
  launch_desc(d, params, st);
}

template <int LOG2N, bool LAST, bool WIDE>
static void describe_mix1_t(const R2CArgs& a, LaunchDesc* d) {
  using M = Mix1Cfg<LOG2N, WIDE>;
  static bool attr_done[64] = {};
  if (M::SMEM_BYTES > 48 * 1024 && need_smem_attr(attr_done))
    cudaFuncSetAttribute(k_mix1<LOG2N, LAST, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, M::SMEM_BYTES);
  d->func = reinterpret_cast<const void*>(k_mix1<LOG2N, LAST, WIDE>);
  d->grid = dim3((unsigned)((a.n_fft + M::G - 1) / M::G));
  d->block = dim3(M::CTA);
  d->smem = M::SMEM_BYTES;
}

// Wide CTAs from this many sources on: the mix is then a throughput problem (measured: 4096 sources 0.0348 ms wide vs
// 0.0450 ms narrow per step; 256 sources 0.0198 ms wide vs 0.0164 ms narrow)
static bool mix1_wide(int n_sources) { return n_sources >= 1024; }

int mix1_sources_per_cta(int B, int n_sources) {
  const bool w = mix1_wide(n_sources);
#define PGX_G(L) (w ? Mix1Cfg<L, true>::G : Mix1Cfg<L, false>::G)
  switch (ilog2(B)) {
    case 4: return PGX_G(4); case 5: return PGX_G(5); case 6: return PGX_G(6);
    case 7: return PGX_G(7); case 8: return PGX_G(8); case 9: return PGX_G(9);
    case 10: return PGX_G(10);
    default: return 0;  // larger transforms keep the three-kernel step
  }
#undef PGX_G
}

bool describe_mix1(const R2CArgs& a, bool last, LaunchDesc* d) {
  d->func = nullptr;
  const bool w = mix1_wide(a.n_fft);
#define PGX_MIX1_CASE(L)                                                \
  case L:                                                               \
    if (last) { if (w) describe_mix1_t<L, true, true>(a, d); else describe_mix1_t<L, true, false>(a, d); }    \
    else { if (w) describe_mix1_t<L, false, true>(a, d); else describe_mix1_t<L, false, false>(a, d); }      \
    break;
  switch (ilog2(a.B)) {
    PGX_MIX1_CASE(4) PGX_MIX1_CASE(5) PGX_MIX1_CASE(6) PGX_MIX1_CASE(7) PGX_MIX1_CASE(8) PGX_MIX1_CASE(9) PGX_MIX1_CASE(10)
    default: break;
  }
#undef PGX_MIX1_CASE
  return d->func != nullptr;
}

void launch_mix1(const R2CArgs& a, const C2RArgs& k, float2* ynow, unsigned int* ticket, cudaStream_t st) {
  LaunchDesc d;
  describe_mix1(a, ticket != nullptr, &d);
  void* params[] = {const_cast<R2CArgs*>(&a), const_cast<C2RArgs*>(&k), &ynow, &ticket};
  launch_desc(d, params, st);
}

bool describe_c2r_emit(const C2RArgs& a, LaunchDesc* d) {
  d->func = nullptr;
  if (a.n_split > 0 || a.n_split_now > 0) {
    PGX_DISPATCH(ilog2(a.B), (describe_c2r_t<L_, true>(a, d)));
  } else {
    PGX_DISPATCH(ilog2(a.B), (describe_c2r_t<L_, false>(a, d)));
  }
  return d->func != nullptr;
}

void launch_c2r_emit(const C2RArgs& a, cudaStream_t st) {
  LaunchDesc d;
  describe_c2r_emit(a, &d);
  void* params[] = {const_cast<C2RArgs*>(&a)};
  launch_desc(d, params, st);
}

}  // namespace pgx
