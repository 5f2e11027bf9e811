// k_osc.cu -- device-resident sources for the convolution path.
//
//   k_sine_bank : N constant-parameter SinePE streams        (reference sine_pe.py:135-175)
//   k_blit_bank : V voices x U band-limited sawtooth oscillators, i.e. BlitSawPE (U = 1) and SuperSawPE
//                 (reference blit_saw_pe.py:152-264, super_saw_pe.py:282-305)
//
// All arithmetic is float64 like the reference (float64 phase, float64 Dirichlet kernel, float64 leaky
// integrator whose gain 1/(1-leak) = 1000 would amplify float32 noise), rounded to float32 exactly where the
// reference rounds (every oscillator's Snippet, then the voice sum).  There is no HBM traffic to speak of:
// the kernels are FP64-transcendental bound and tiny next to the convolution they feed; what they buy is
// that the inputs of the convolution never exist on the host.
//
// The BLIT recurrence  saw[k] = x[k] + leak * saw[k-1]  (scipy.signal.lfilter in the reference) is a linear
// scan: a warp owns one oscillator and walks the pull in tiles of 128 samples; each lane integrates 4
// consecutive samples, the 32 lane carries are combined with a weighted warp-shuffle scan, and the carry
// into the tile is the integrator state.
#include "kernels.h"

namespace pgx {

__global__ void __launch_bounds__(256) k_sine_bank(const SineArgs a) {
  const int64_t total = (int64_t)a.n_streams * a.n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(e / a.n), i = (int)(e - (int64_t)s * a.n);
    const double f = a.params[3 * s], amp = a.params[3 * s + 1], ph = a.params[3 * s + 2];
    const double time = (double)(a.start + i) / (double)a.sample_rate;    // sine_pe.py:173-174
    const double phase = ph + ((2.0 * 3.141592653589793) * f) * time;     // :175 (left-to-right product)
    const float v = (float)(amp * sin(phase));                            // :146,157
    for (int c = 0; c < a.channels; ++c) a.out[(int64_t)s * a.os + (int64_t)c * a.oc + (int64_t)i * a.oi] = v;
  }
}

// Modulated SinePE: the reference's STATEFUL branch (sine_pe.py:139-142,188-232), taken as soon as any of
// frequency / amplitude / phase is a PE.  One CTA per voice.  phase[i] = cumsum(2 pi f[i] / sr)[i] + initial
// (+ phase_mod[i]); np.cumsum is a left-to-right float64 sum, so ONE thread walks the increments (staged by all
// threads first: the walk is a chain of dependent DADDs, ~2 us for a 512-sample pull) -- bit-identical phases, not a
// re-associated scan.  The accumulated phase handed to the next pull is the last sample's phase INCLUDING its phase
// modulation, exactly like :229.
__global__ void __launch_bounds__(256) k_sine_mod(const SineModArgs a) {
  const int v = blockIdx.x;
  double* ph = a.scratch + (size_t)v * a.max_pull;
  const float* fq = a.freq ? a.freq + (size_t)v * a.n : nullptr;
  const float* am = a.amp ? a.amp + (size_t)v * a.n : nullptr;
  const float* pm = a.phase ? a.phase + (size_t)v * a.n : nullptr;
  const double f0 = a.params[3 * v], a0 = a.params[3 * v + 1], p0 = a.params[3 * v + 2];
  const double sr = (double)a.sample_rate;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const double f = fq ? (double)fq[i] : f0;
    ph[i] = ((2.0 * 3.141592653589793) * f) / sr;                 // :198 phase_increment (left-to-right product, then / sr)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < a.n; ++i) {                                // :216 np.cumsum
      acc += ph[i];
      ph[i] = acc;
    }
  }
  __syncthreads();
  // :203-213 first render: the constant phase (0 when the phase is a PE); afterwards the accumulated phase
  const double initial = a.first ? (pm ? 0.0 : p0) : a.state[v];
  __syncthreads();                                                 // everyone has read the carried phase before it is replaced
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    double p = ph[i] + initial;                                    // :216
    p = p + (pm ? (double)pm[i] : p0);                             // :219-222 (a constant phase is a constant ARRAY there:
                                                                   // processing_element.py:360-365 -- it is added to every
                                                                   // sample and, through :229, carried into the next pull)
    const double amp = am ? (double)am[i] : a0;
    const float y = (float)(amp * sin(p));                         // :146,157
    for (int c = 0; c < a.channels; ++c) a.out[(int64_t)v * a.os + (int64_t)c * a.oc + (int64_t)i * a.oi] = y;
    if (i == a.n - 1) a.state[v] = p;                              // :229
  }
}

void launch_sine_mod(const SineModArgs& a, cudaStream_t st) { k_sine_mod<<<a.n_voices, 256, 0, st>>>(a); }

void launch_sine_bank(const SineArgs& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.n_streams * a.n;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_sine_bank<<<(int)blocks, 256, 0, st>>>(a);
}


__device__ __forceinline__ double shfl_up_d(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, d);
  hi = __shfl_up_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

// One CTA per voice, one warp per oscillator of the voice (U <= 32 warps).  kOscT = consecutive samples per
// lane (a warp tile is 32*kOscT samples): 1 or 2 for the short pulls of a low-latency voice bank so that all
// lanes work, 4 for long pulls.
template <int kOscT>
__global__ void __launch_bounds__(1024) k_blit_bank(const BlitArgs a) {
  constexpr int kOscTile = 32 * kOscT;
  extern __shared__ float tile_sm[];  // [U][kOscTile] float32 oscillator outputs of the current tile
  const int v = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = a.unison;
  const int o = v * U + w;
  // per-oscillator constants, computed once on the host with the reference's own float64 expressions
  // (blit_saw_pe.py:167-177,189,198,218): phase increment, period, 1/period, harmonic count
  const double gain = a.gain[o];
  const double inc = a.consts[4 * o], P = a.consts[4 * o + 1], invP = a.consts[4 * o + 2], M = a.consts[4 * o + 3];
  const double leak = a.leak;
  double lk[kOscT + 1];  // leak^q
  lk[0] = 1.0;
#pragma unroll
  for (int q = 1; q <= kOscT; ++q) lk[q] = lk[q - 1] * leak;
  // weight of the tile's carry-in at this lane, leak^(kOscT*lane), by binary exponentiation over the lane bits
  double lane_pow = 1.0, sq = lk[kOscT];
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    if (lane & (1 << b)) lane_pow *= sq;
    sq *= sq;
  }
  double ph0 = a.st_phase[o], y0 = a.st_int[o];               // state at the start of the pull
  if (a.snap_phase && lane == 0) {                            // speculative pull: keep it for pgx_osc_rollback
    a.snap_phase[o] = ph0;
    a.snap_int[o] = y0;
  }
  double ph_last = ph0;

  for (int t0 = 0; t0 < a.n; t0 += kOscTile) {
    const int n_here = min(kOscTile, a.n - t0);
    const int klast = t0 + n_here - 1;
    // ---- this lane's 4 samples: BLIT minus DC, integrated from zero
    double loc[kOscT];
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kOscT; ++q) {
      const int k = t0 + lane * kOscT + q;
      double x = 0.0;
      if (k < a.n) {
        double ph = ph0 + (double)(k + 1) * inc;   // :192 (cumsum of a constant increment)
        ph -= floor(ph);                           // :195 np.mod(phase, 1.0)
        if (k == a.n - 1) ph_last = ph;
        const double theta = 3.141592653589793 * ph;
        const double sd = sin(theta);
        const double blit = (fabs(sd) < 1e-9) ? (M / P) : (sin(M * theta) / (P * sd));  // :203-214
        x = blit - invP;
      }
      s = x + leak * s;
      loc[q] = s;
    }
    // ---- weighted inclusive scan of the lane carries: B[l] = sum_{j<=l} leak^(4(l-j)) s_j
    double B = s, pw = lk[kOscT];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double up = shfl_up_d(B, d);
      if (lane >= d) B = fma(pw, up, B);
      pw *= pw;
    }
    double carry = shfl_up_d(B, 1);
    if (lane == 0) carry = 0.0;
    carry = fma(lane_pow, y0, carry);  // integrator value just before this lane's first sample
    float* row = tile_sm + w * kOscTile + lane * kOscT;
    double ylast = 0.0;
#pragma unroll
    for (int q = 0; q < kOscT; ++q) {
      const double y = fma(lk[q + 1], carry, loc[q]);
      if (t0 + lane * kOscT + q == klast) ylast = y;
      row[q] = (float)(y * 2.0 * gain);  // :256-262 each oscillator's Snippet is float32
    }
    // new integrator state: the last valid sample of the tile
    y0 = shfl_d(ylast, (n_here - 1) / kOscT);
    __syncthreads();
    // ---- voice output: float64 sum of the float32 oscillator outputs, x amplitude (super_saw_pe.py:292-303)
    for (int i = threadIdx.x; i < n_here; i += blockDim.x) {
      double acc = 0.0;
      for (int u = 0; u < U; ++u) acc += (double)tile_sm[u * kOscTile + i];
      const float val = (float)(acc * a.amp[v]);
      for (int c = 0; c < a.channels; ++c)  // np.tile over channels (super_saw_pe.py:304-305)
        a.out[(int64_t)v * a.os + (int64_t)c * a.oc + (int64_t)(t0 + i) * a.oi] = val;
    }
    __syncthreads();
  }
  // ---- carry the state to the next pull (:249-250)
  const int last_lane = ((a.n - 1) % kOscTile) / kOscT;
  ph_last = shfl_d(ph_last, last_lane);
  if (lane == 0) {
    a.st_phase[o] = ph_last;
    a.st_int[o] = y0;
  }
}

// Modulated BLIT oscillators: BlitSawPE / SuperSawPE with a PE-valued frequency and / or amplitude
// (blit_saw_pe.py:161-262 with per-sample control vectors; super_saw_pe.py:223-246: every oscillator's frequency is
// GainPE(frequency_pe, ratio) -- a FLOAT32 product -- and :287-303: the voice amplitude multiplies the float64 sum).
// Same structure as k_blit_bank (one CTA per voice, one warp per oscillator, 128-sample tiles), but the harmonic
// count, the period and the phase increment change from sample to sample, so the two recurrences are walked in the
// reference's own order by lane 0 -- the phase is np.cumsum of the increments over the whole pull (:186-192), the
// leaky integrator is lfilter's  y[k] = x[k] + (leak * y[k-1])  (:225-236; product and sum rounded separately) --
// while all lanes evaluate the Dirichlet kernel in between.
__global__ void __launch_bounds__(1024) k_blit_mod(const BlitModArgs a) {
  constexpr int kTile = 128;
  extern __shared__ double mod_sm[];           // [U][2][kTile] doubles, then [U][kTile] floats
  const int v = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = a.unison;
  const int o = v * U + w;
  double* s_a = mod_sm + (size_t)w * 2 * kTile;   // increments -> cumulative phase
  double* s_b = s_a + kTile;                      // blit_ac -> saw
  float* tile = reinterpret_cast<float*>(mod_sm + (size_t)U * 2 * kTile);
  const float* fq = a.freq ? a.freq + (size_t)v * a.n : nullptr;
  const float* am = a.amp ? a.amp + (size_t)v * a.n : nullptr;
  const float* mc = a.m_ctl ? a.m_ctl + (size_t)v * a.n : nullptr;
  const double sr = (double)a.sample_rate, leak = a.leak;
  const float ratio = (float)a.osc_freq[o];       // detune ratio when the frequency is a PE, else the frequency itself
  const double f_const = a.osc_freq[o];
  const double gain = a.gain[o];
  const int m_fixed = a.m_fixed ? a.m_fixed[o] : 0;
  const double ph0 = a.st_phase[o];
  double acc = 0.0, y = a.st_int[o], ph_last = ph0;
  auto freq_at = [&](int k) -> double {
    return fq ? (double)(__fmul_rn(fq[k], ratio)) : f_const;      // GainPE: float32 product (gain_pe.py:123-125)
  };
  for (int t0 = 0; t0 < a.n; t0 += kTile) {
    const int nt = min(kTile, a.n - t0);
    for (int i = lane; i < nt; i += 32) s_a[i] = freq_at(t0 + i) / sr;                 // :186
    __syncwarp();
    if (lane == 0)
      for (int i = 0; i < nt; ++i) { acc += s_a[i]; s_a[i] = acc; }                    // :189 np.cumsum
    __syncwarp();
    for (int i = lane; i < nt; i += 32) {
      const double f = freq_at(t0 + i);
      double ph = ph0 + s_a[i];                                                        // :189
      ph = fmod(ph, 1.0);                                                              // :192 np.mod: sign of the divisor
      if (ph < 0.0) ph += 1.0;
      const double fm = fmax(f, 1.0);
      double m;
      if (mc) {                                                                        // :175-177 a PE-valued m
        const int mi = (int)(double)mc[t0 + i];                                        // astype(int32): toward zero
        m = (double)(mi < 1 ? 1 : mi);
      } else if (m_fixed > 0) {
        m = (double)m_fixed;
      } else {
        int mi = (int)floor(sr / (2.0 * fm));                                          // :169-174
        mi = mi - (1 - mi % 2);
        if (mi < 1) mi = 1;
        m = (double)mi;
      }
      const double P = sr / fm;                                                        // :195
      const double theta = 3.141592653589793 * ph, sden = sin(theta);
      const double blit = fabs(sden) < 1e-9 ? m / P : sin(m * theta) / (P * sden);     // :200-211
      s_b[i] = blit - 1.0 / P;                                                         // :215
      if (t0 + i == a.n - 1) ph_last = ph;
    }
    __syncwarp();
    if (lane == 0)
      for (int i = 0; i < nt; ++i) { y = __dadd_rn(s_b[i], __dmul_rn(leak, y)); s_b[i] = y; }   // :225-236 lfilter
    __syncwarp();
    for (int i = lane; i < nt; i += 32) {
      const double amp_o = (am && a.amp_per_osc) ? (double)am[t0 + i] : gain;
      tile[w * kTile + i] = (float)((s_b[i] * 2.0) * amp_o);                           // :252-259 the oscillator's float32 Snippet
    }
    __syncthreads();
    // voice output: float64 sum of the float32 oscillator outputs, x the voice amplitude (super_saw_pe.py:287-303)
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
      double sum = 0.0;
      for (int u = 0; u < U; ++u) sum += (double)tile[u * kTile + i];
      const double va = (am && !a.amp_per_osc) ? (double)am[t0 + i] : a.vamp[v];
      const float val = (float)(sum * va);
      for (int c = 0; c < a.channels; ++c) a.out[(int64_t)v * a.os + (int64_t)c * a.oc + (int64_t)(t0 + i) * a.oi] = val;
    }
    __syncthreads();
  }
  ph_last = shfl_d(ph_last, (a.n - 1) % 32);     // the lane that owned the last sample (i = lane + 32 q)
  y = shfl_d(y, 0);
  if (lane == 0) {                               // :249-250
    a.st_phase[o] = ph_last;
    a.st_int[o] = y;
  }
}

void launch_blit_mod(const BlitModArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.unison * (2 * 128 * sizeof(double) + 128 * sizeof(float));
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_blit_mod, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // unison > 18
  k_blit_mod<<<a.n_voices, a.unison * 32, smem, st>>>(a);
}

void launch_blit_bank(const BlitArgs& a, cudaStream_t st) {
  const int threads = a.unison * 32;
  if (a.n <= 32) k_blit_bank<1><<<a.n_voices, threads, (size_t)a.unison * 32 * sizeof(float), st>>>(a);
  else if (a.n <= 64) k_blit_bank<2><<<a.n_voices, threads, (size_t)a.unison * 64 * sizeof(float), st>>>(a);
  else k_blit_bank<4><<<a.n_voices, threads, (size_t)a.unison * 128 * sizeof(float), st>>>(a);
}

// ---- PCM16 <-> float32 staging for WAV I/O -----------------------------------------
// in : float32 = int16 / 32768, what soundfile.read(dtype="float32") returns for a PCM_16 file
//      (wav_reader_pe.py:127-132; libsndfile pcm.c s2f_array, normfact 1/0x8000)
// out: int16 = clip(lrintf(x * 32768)) with round-half-even, libsndfile's float -> PCM_16 conversion when
//      clipping is on, which python-soundfile always enables (wav_writer_pe.py:153; libsndfile pcm.c
//      f2s_clip_array).  soundfile / libsndfile are not in this image: parity for this pair is anchored on
//      the formulas and round-trip properties, not on golden files.
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const int16_t* __restrict__ in, float* __restrict__ out,
                                                      const int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = (float)in[e] * (1.0f / 32768.0f);
}

__global__ void __launch_bounds__(256) k_f32_to_pcm16(const float* __restrict__ in, int16_t* __restrict__ out,
                                                      const int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float scaled = in[e] * 32768.0f;
    int v;
    if (scaled >= 32767.0f) v = 32767;
    else if (scaled <= -32768.0f) v = -32768;
    else v = __float2int_rn(scaled);
    out[e] = (int16_t)v;
  }
}

__global__ void __launch_bounds__(256) k_copy_block(const float* __restrict__ src, const int64_t ss, const int64_t sc,
                                                    const int64_t si, float* __restrict__ dst, const int64_t ds,
                                                    const int64_t dc, const int64_t di, const int rows, const int chans,
                                                    const int n) {
  const int64_t total = (int64_t)rows * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), i = (int)(e - (int64_t)r * n);
    const int s = r / chans, c = r - s * chans;
    dst[s * ds + c * dc + i * di] = src[s * ss + c * sc + i * si];
  }
}

static int pcm_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}
void launch_copy_block(const float* src, int64_t ss, int64_t sc, int64_t si, float* dst, int64_t ds, int64_t dc,
                       int64_t di, int32_t n_streams, int32_t chans, int32_t n, cudaStream_t st) {
  const int rows = n_streams * chans;
  k_copy_block<<<pcm_grid((int64_t)rows * n), 256, 0, st>>>(src, ss, sc, si, dst, ds, dc, di, rows, chans, n);
}
void launch_pcm16_to_f32(const int16_t* in, float* out, int64_t n, cudaStream_t st) {
  k_pcm16_to_f32<<<pcm_grid(n), 256, 0, st>>>(in, out, n);
}
void launch_f32_to_pcm16(const float* in, int16_t* out, int64_t n, cudaStream_t st) {
  k_f32_to_pcm16<<<pcm_grid(n), 256, 0, st>>>(in, out, n);
}

}  // namespace pgx
