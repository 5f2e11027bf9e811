// k_fft.cu -- K1 (ingest + real FFT), K2 (inverse real FFT + emit) and filter preparation.
//
// K1/K2 replace np.fft.rfft / np.fft.irfft in the reference's overlap-save loop
// (convolve_pe.py:294-322): K1 also does what :294-310 (build [tail | segment | zeros]) and
// :325-336 (tail update) do, K2 what :321-322 (keep the valid output samples) does.
#include "fft.cuh"
#include "kernels.h"

namespace pgx {

static constexpr int kFftThreads = 256;

// threads per transform: one radix-4 butterfly per thread per pass when the CTA allows it
__host__ __device__ inline int threads_per_fft(int B) { return (B / 4 < kFftThreads) ? B / 4 : kFftThreads; }

int fft_smem_bytes(int B) {
  const int T = threads_per_fft(B);
  const int fpb = kFftThreads / T;
  return fpb * 2 * B * (int)sizeof(float2);
}

__global__ void __launch_bounds__(kFftThreads) k_r2c_ingest(const R2CArgs a) {
  extern __shared__ float2 sm[];
  const int n = a.B;
  const int T = threads_per_fft(n);
  const int fpb = kFftThreads / T;
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  const int f = blockIdx.x * fpb + g;
  const bool active = f < a.n_fft;
  float2* bufA = sm + (size_t)g * 2 * n;
  float2* bufB = bufA + n;

  float* cur = nullptr;
  const float* prev = nullptr;
  const int m_new = a.fill + a.take;
  if (active) {
    const int s = f / a.c_x, cx = f - s * a.c_x;
    cur = a.hist + ((size_t)f * 2 + a.half) * n;
    prev = a.hist + ((size_t)f * 2 + (a.half ^ 1)) * n;
    // ingest: open block [fill, fill+take) := new samples (optionally the mean over source channels)
    for (int i = t; i < a.take; i += T) {
      const int64_t base = (int64_t)s * a.xs + (int64_t)(a.x_off + i) * a.xi;
      float v;
      if (a.mixdown) {
        float acc = 0.f;
        for (int c = 0; c < a.c_in; ++c) acc += a.x[base + c * a.xc];
        v = acc / (float)a.c_in;
      } else {
        v = a.x[base + cx * a.xc];
      }
      cur[a.fill + i] = v;
    }
  }
  __syncthreads();
  if (active) {
    // z[q] = w[2q] + i*w[2q+1] over the window w = [prev (B) | cur[0:m_new) | zeros]
    for (int q = t; q < n; q += T) {
      const int i0 = 2 * q;
      float2 z;
      if (i0 < n) {
        z = *reinterpret_cast<const float2*>(prev + i0);
      } else {
        const int c0 = i0 - n;
        z.x = (c0 < m_new) ? cur[c0] : 0.f;
        z.y = (c0 + 1 < m_new) ? cur[c0 + 1] : 0.f;
      }
      bufA[q] = z;
    }
  }
  __syncthreads();
  const float2* Z = stockham_passes<false>(bufA, bufB, n, t, T, a.tw);
  if (active) {
    float2* row = a.fdl + ((size_t)f * a.P + a.slot) * n;
    for (int k = t; k < n; k += T) row[k] = r2c_bin(Z, n, k, a.tw);
  }
}

void launch_r2c_ingest(const R2CArgs& a, cudaStream_t st) {
  const int T = threads_per_fft(a.B);
  const int fpb = kFftThreads / T;
  const int grid = (a.n_fft + fpb - 1) / fpb;
  const int smem = fft_smem_bytes(a.B);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_r2c_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_r2c_ingest<<<grid, kFftThreads, smem, st>>>(a);
}

__global__ void __launch_bounds__(kFftThreads) k_filter_prep(const FilterPrepArgs a) {
  extern __shared__ float2 sm[];
  const int n = a.B;
  const int T = threads_per_fft(n);
  const int fpb = kFftThreads / T;
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  const int64_t f = (int64_t)blockIdx.x * fpb + g;  // (row, partition)
  const int64_t total = (int64_t)a.n_rows * a.P;
  const bool active = f < total;
  float2* bufA = sm + (size_t)g * 2 * n;
  float2* bufB = bufA + n;
  int row = 0, p = 0;
  if (active) {
    row = (int)(f / a.P);
    p = (int)(f - (int64_t)row * a.P);
    const float* h = a.h + (size_t)row * a.L;
    const int base = p * n;
    for (int q = t; q < n; q += T) {
      const int i0 = 2 * q;  // window = [partition (B taps) | zeros (B)]
      float2 z = make_float2(0.f, 0.f);
      if (i0 < n) {
        if (base + i0 < a.L) z.x = h[base + i0];
        if (base + i0 + 1 < a.L) z.y = h[base + i0 + 1];
      }
      bufA[q] = z;
    }
  }
  __syncthreads();
  const float2* Z = stockham_passes<false>(bufA, bufB, n, t, T, a.tw);
  if (active) {
    // reversed + doubled layout: partition p at rows P-1-p and 2P-1-p, so that the rows paired with
    // delay-line slots 0..P-1 are the contiguous run starting at P-1-head (see k_mac.cu).
    const float scale = 1.0f / (float)n;  // the inverse transform's 1/B, folded in here
    float2* r0 = a.Hd + ((size_t)row * 2 * a.P + (a.P - 1 - p)) * n;
    float2* r1 = r0 + (size_t)a.P * n;
    for (int k = t; k < n; k += T) {
      float2 v = r2c_bin(Z, n, k, a.tw);
      v.x *= scale;
      v.y *= scale;
      r0[k] = v;
      r1[k] = v;
    }
  }
}

void launch_filter_prep(const FilterPrepArgs& a, cudaStream_t st) {
  const int T = threads_per_fft(a.B);
  const int fpb = kFftThreads / T;
  const int64_t total = (int64_t)a.n_rows * a.P;
  const int grid = (int)((total + fpb - 1) / fpb);
  const int smem = fft_smem_bytes(a.B);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_filter_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_filter_prep<<<grid, kFftThreads, smem, st>>>(a);
}

__global__ void __launch_bounds__(kFftThreads) k_c2r_emit(const C2RArgs a) {
  extern __shared__ float2 sm[];
  const int n = a.B;
  const int T = threads_per_fft(n);
  const int fpb = kFftThreads / T;
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  const int o = blockIdx.x * fpb + g;
  const bool active = o < a.n_out;
  float2* bufA = sm + (size_t)g * 2 * n;
  float2* bufB = bufA + n;
  if (active) {
    const float2 *xrow = nullptr, *hrow = nullptr;
    if (a.fdl) {  // conv mode: present term = delay-line slot `head` x filter partition 0 (row P-1 of Hd)
      const int s = o / a.c_out, c = o - s * a.c_out;
      const int gx = (a.c_x == 1) ? 0 : c, fc = (a.c_f == 1) ? 0 : c;
      xrow = a.fdl + ((size_t)(s * a.c_x + gx) * a.P + a.head) * n;
      hrow = a.Hd + ((size_t)(__ldg(a.fmap + s) * a.c_f + fc) * 2 * a.P + (a.P - 1)) * n;
    }
    for (int k = t; k < n; k += T) {
      float2 acc = make_float2(0.f, 0.f);
      if (xrow) {
        const float2 x = xrow[k], h = __ldg(hrow + k);
        acc = (k == 0) ? make_float2(x.x * h.x, x.y * h.y) : cmul(x, h);
      }
      for (int sp = 0; sp < a.n_split; ++sp) {
        const float2 v = a.yspec[((size_t)sp * a.n_out + o) * n + k];
        acc.x += v.x;
        acc.y += v.y;
      }
      for (int sp = 0; sp < a.n_split_now; ++sp) {
        const float2 v = a.ynow[((size_t)sp * a.n_out + o) * n + k];
        acc.x += v.x;
        acc.y += v.y;
      }
      bufB[k] = acc;
    }
  }
  __syncthreads();
  if (active) {
    for (int k = t; k < n; k += T) bufA[k] = c2r_bin(bufB, n, k, a.tw);
  }
  __syncthreads();
  const float2* z = stockham_passes<true>(bufA, bufB, n, t, T, a.tw);
  if (active) {
    // overlap-save: output samples of the open block live at window positions [B+fill, B+fill+take)
    const int s = o / a.c_out, c = o - s * a.c_out;
    float* y = a.y + (int64_t)s * a.ys + (int64_t)c * a.yc;
    for (int i = t; i < a.take; i += T) {
      const int idx = n + a.fill + i;
      const float2 zz = z[idx >> 1];
      y[(int64_t)(a.y_off + i) * a.yi] = (idx & 1) ? zz.y : zz.x;
    }
  }
}

void launch_c2r_emit(const C2RArgs& a, cudaStream_t st) {
  const int T = threads_per_fft(a.B);
  const int fpb = kFftThreads / T;
  const int grid = (a.n_out + fpb - 1) / fpb;
  const int smem = fft_smem_bytes(a.B);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_c2r_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_c2r_emit<<<grid, kFftThreads, smem, st>>>(a);
}

}  // namespace pgx
