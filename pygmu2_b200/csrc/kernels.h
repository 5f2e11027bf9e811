// kernels.h -- launch interfaces of the sm_100a kernels behind libpgx.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgx {

// K1: ingest new samples, window [prev block | open block | 0], real FFT, write one delay-line row.
struct R2CArgs {
  const float* x;        // device samples: element (s, c, i) at x[s*xs + c*xc + (x_off+i)*xi]
  int64_t xs, xc, xi;
  int32_t x_off;
  float* hist;           // [n_fft][2][B] time-domain halves (previous block / open block)
  float2* fdl;           // [n_fft][R][B] packed input spectra (frequency-domain delay line, R ring rows)
  const float2* tw;      // [2B] exp(-2*pi*i*k/2B)
  int32_t n_fft;         // N * c_x transforms
  int32_t c_in, c_x, B, R;
  int32_t slot, half, fill, take;
  int32_t mixdown;
  int32_t fast;          // whole block (fill 0, take B) of contiguous, 8-byte aligned samples, no mix-down
};
void launch_r2c_ingest(const R2CArgs& a, cudaStream_t st);

// Filter preparation: partition p < P of filter row f -> spectrum rows (R-1-p) and (2R-1-p), scaled 1/B.
// (R = ring rows >= P; rows of partitions P..R-1 stay zero.)
struct FilterPrepArgs {
  const float* h;        // device [n_rows][L]
  float2* Hd;            // [n_rows][2R][B]
  const float2* tw;
  int32_t n_rows, L, B, P, R;
};
void launch_filter_prep(const FilterPrepArgs& a, cudaStream_t st);

// K3/K4: complex multiply-accumulate over delay-line rows x filter-spectrum rows.
struct MacArgs {
  const float4* fdl;     // rows of W4 float4
  const float4* Hd;
  float4* yspec;         // [n_split][n_out][W4]
  const int32_t* fmap;   // [N] filter of stream
  int32_t N, c_x, c_out, c_f, R, W4, q0;  // R = ring rows per (stream, channel); q0 = R-1-head
  int32_t n_out, n_terms, terms_per_split, n_split;
  int32_t n_otiles, st;  // out tiles in the flat grid; streams per CTA sharing the filter rows (1 or 4)
  int32_t mix;           // 0: out o=(s,c), terms j<P.  1: out o=c, terms (s,j) (fused MixPE / HRTF stereo mix).
                         // 2: items as 0, partial rows [stream][split][channel] folded over streams afterwards
  // which ring slots are terms: Pt slots per stream; slot = jfix if jfix >= 0 (only that slot), else
  // j = off + jj, plus nskip when j >= skip, for jj in [0, Pt): the background pass leaves out the open
  // slot `head` and the spare slot head+1 (being written by the next block's K1).
  // n_terms = Pt (conv) or N*Pt (mix).
  int32_t Pt, off, skip, nskip, jfix;
  int32_t variant;       // 0 = LDG kernel (k_mac.cu), 1 = bulk-async staged kernel (k_mac_tma.cu)
  int32_t persistent_ctas;
  // time tiling (conv layout only): ONE pass over the committed rows yields the past sums of `tile` consecutive
  // output blocks, head, head+1, ... (slot of output kappa = (head + kappa) mod R); yspec then holds `tile` result
  // sets of tile_stride float4 each.  P = partitions of the filter, head = slot of the first output block.
  int32_t tile, P, head;
  int32_t tile_u;                  // rows in flight per thread / slots per stage (selects the instantiation)
  int32_t tile_stages;             // bulk-async variant: stages of the shared-memory ring
  int32_t tile_set0, tile_nsets;   // output kappa goes to result set (tile_set0 + kappa) % tile_nsets
  int64_t tile_stride;
};
struct MacPlan {
  int32_t st, n_otiles, n_split, terms_per_split, grid, occupancy, variant, persistent_ctas;
  int32_t tile;        // > 1: time-tiled pass (k_fdl_mac_tile): `tile` output blocks per pass
  int32_t tile_u;      // its rows in flight per thread (LDG) / slots per stage (bulk-async)
  int32_t tile_stages; // bulk-async variant: ring stages
  int32_t layout;      // MacArgs.mix value: 0 conv, 1 flattened mix, 2 per-stream mix
  int32_t n_partials;  // partial rows per out row that the consumer has to sum
};
// bulk-async (TMA) staged variant, k_mac_tma.cu
bool tma_supported(int W4);
int tma_occupancy_of(bool mix, int st);
void launch_fdl_mac_tma(const MacArgs& a, int persistent_ctas, cudaStream_t st);
// Fold split partial sums: out[o][k] = sum_sp in[sp][o][k] (rows of W4 float4), deterministic order.
void launch_reduce_partials(const float4* in, float4* out, int n_split, int n_out, int W4, cudaStream_t st);
MacPlan mac_plan(int N, int c_out, int W4, int n_terms, bool mix, bool shared_filter, int sm_count);
// bulk-async staged variant of the time-tiled pass, k_mac_tile_tma.cu
bool tile_tma_supported(int W4);
bool tile_tma_config(bool shared_filter, int N, int W4, int tile, int* st, int* tps, int* stages, int* occupancy);
int tile_tma_ktiles(int W4);
void launch_fdl_mac_tile_tma(const MacArgs& a, cudaStream_t st);
// same for the time-tiled pass (conv layout): tile = 2 or 4 output blocks per pass
MacPlan mac_plan_tiled(int N, int c_out, int W4, int n_terms, bool shared_filter, int sm_count, int tile);
void launch_fdl_mac(const MacArgs& a, cudaStream_t st);

// K2: sum split partials, inverse real FFT, emit the new output samples.
struct C2RArgs {
  const float2* yspec;   // [n_split][n_out][B]   partial sums of the background (past-partition) pass
  int32_t n_split, n_out;
  const float2* ynow;    // [n_split_now][n_out][B] partial sums of the present pass (mix mode), or NULL
  int32_t n_split_now;
  // conv mode: the present term X[slot head] * H[partition 0] is folded in here (one row pair per out)
  const float2* fdl;     // NULL when ynow carries the present term
  const float2* Hd;
  const int32_t* fmap;
  int32_t c_x, c_f, R, head;
  float* y;              // element (s, c, i) at y[s*ys + c*yc + (y_off+i)*yi]; o = s*c_out + c
  int64_t ys, yc, yi;
  int32_t y_off;
  int32_t c_out, B, fill, take;
  const float2* tw;
  // fused wet/dry output stage (ReverbPE: GainPE(dry) + GainPE(wet) -> MixPE, reverb_pe.py:82-95):
  // y = dry * x + wet * conv, each product and the sum rounded to float32 like the reference's three PEs.
  // xdry = the pull's input samples (same addressing as R2CArgs.x), NULL when dry == 0.
  const float* xdry;
  int64_t xs, xc, xi;
  int32_t x_off;
  float wet, dry;
  int32_t fast;          // whole block into contiguous, 8-byte aligned y (and xdry, add): vector stores
  // addend (two-level partitioning: the tail level's contribution to these output samples), added to the
  // convolution before the output gains; element (s, c, i) at add[s*as + c*ac + (y_off... see emit_block) ], or NULL
  const float* add;
  int64_t as, ac, ai;
  // fused small-P step (k_conv1<.., PAST>): the past partitions are ring slots p_off + jj (+ p_nskip from p_skip on),
  // jj < n_past, paired with filter rows q0 + slot -- the same slot walk as MacArgs
  int32_t n_past, p_off, p_skip, p_nskip, q0;
  int32_t n_recent;      // time-tiled banks: besides the present term X[head] H_0, the output stage adds the n_recent newest
                         // committed rows X[head - r] H_r, r = 1..n_recent (rows younger than the tiled pass that covers
                         // this block)
  int32_t fft16;         // B = 4096 fused step: 0 radix-8 kernel, 1 radix-16 (shared-memory exchanges), 2 radix-16 + shuffles
};
void launch_c2r_emit(const C2RArgs& a, cudaStream_t st);
// K1 + K2 in one kernel for single-partition conv banks (P = 1): the spectrum never visits the delay line.
void launch_conv1(const R2CArgs& a, const C2RArgs& k, cudaStream_t st);
// Radix-16 variant of the same step for B = 4096 (k_fft16.cu): 16 elements per thread, two exchanges per transform, the
// half-warp exchange by warp shuffles.  Returns false when the pull is not its shape (the general kernel runs then).
bool launch_conv1_r16(const R2CArgs& a, const C2RArgs& k, cudaStream_t st);
int conv1_r16_default();  // PGX_FFT16 or the measured default, read when a bank is created

// K5: MixPE left-to-right float32 sum of n_inputs dense arrays.
void launch_mix_sum(const float* in, int32_t n_inputs, int64_t n_elems, float* out, cudaStream_t st);

// Device-resident sources (k_osc.cu).  Output element (voice/stream s, channel c, sample i) at
// out[s*os + c*oc + i*oi].
struct SineArgs {
  const double* params;  // [n_streams][3] = frequency, amplitude, phase offset
  float* out;
  int64_t os, oc, oi;
  int64_t start;         // sample index of the first sample
  int32_t n_streams, channels, n, sample_rate;
};
void launch_sine_bank(const SineArgs& a, cudaStream_t st);

struct SineModArgs {
  const double* params;     // [V][3] constants: frequency, amplitude, phase (used where the control vector is NULL)
  const float* freq;        // [V][n] per-sample controls (device), or NULL
  const float* amp;
  const float* phase;
  double* state;            // [V] accumulated phase carried between pulls
  double* scratch;          // [V][max_pull]
  float* out;
  int64_t os, oc, oi;
  int32_t n_voices, channels, n, sample_rate, max_pull, first;
};
void launch_sine_mod(const SineModArgs& a, cudaStream_t st);

struct BlitArgs {
  const double* consts;    // [V*U][4] phase increment f/sr, period sr/max(f,1), 1/period, harmonic count M
  const double* gain;      // [V*U] oscillator amplitude
  const double* amp;       // [V] voice amplitude
  double* st_phase;        // [V*U] state: phase in [0,1) at the end of the previous pull
  double* st_int;          // [V*U] state: leaky-integrator output at the end of the previous pull
  float* out;
  int64_t os, oc, oi;
  double leak;
  int32_t n_voices, unison, channels, n, sample_rate;
  double* snap_phase;      // speculative pull: the state this pull starts from is copied here first (or NULL)
  double* snap_int;
};
void launch_blit_bank(const BlitArgs& a, cudaStream_t st);

struct BlitModArgs {
  const double* osc_freq;  // [V*U] the oscillator's frequency -- its detune RATIO when a frequency control is given
  const double* gain;      // [V*U] oscillator amplitude
  const double* vamp;      // [V] voice amplitude
  const int32_t* m_fixed;  // [V*U] fixed harmonic count (0 = auto)
  const float* freq;       // [V][n] frequency control of the voice (what the frequency PE rendered), or NULL
  const float* amp;        // [V][n] amplitude control, or NULL: replaces the oscillator amplitude when amp_per_osc
                           // (BlitSawPE), else the voice amplitude (SuperSawPE)
  const float* m_ctl;      // [V][n] harmonic-count control (a PE-valued m, blit_saw_pe.py:175-177), or NULL
  double* st_phase;        // [V*U] state: wrapped phase / integrator output at the end of the previous pull
  double* st_int;
  float* out;
  int64_t os, oc, oi;
  double leak;
  int32_t n_voices, unison, channels, n, sample_rate, amp_per_osc;
};
void launch_blit_mod(const BlitModArgs& a, cudaStream_t st);

// PCM16 <-> float32 staging (k_osc.cu): dense arrays of n elements
void launch_pcm16_to_f32(const int16_t* in, float* out, int64_t n, cudaStream_t st);
void launch_f32_to_pcm16(const float* in, int16_t* out, int64_t n, cudaStream_t st);

// Strided gather of (streams x chans x n) samples: dst[s*ds + c*dc + i*di] = src[s*ss + c*sc + i*si]
void launch_copy_block(const float* src, int64_t ss, int64_t sc, int64_t si, float* dst, int64_t ds, int64_t dc,
                       int64_t di, int32_t n_streams, int32_t chans, int32_t n, cudaStream_t st);

// K1 + present-slot accumulate in one kernel for single-partition mixes of mono transforms (k_mix1): writes
// ceil(n_fft / mix1_sources_per_cta(B, n_fft)) partial rows per channel to ynow; 0 sources per CTA = not available for B.
int mix1_sources_per_cta(int B, int n_sources);
// ticket != NULL: the last CTA to finish also does K2's work (fold, inverse transforms, emit): one launch per step.
// Needs c_out <= mix1_sources_per_cta(B, n_fft); *ticket must be 0 before the first launch (the kernel re-arms it).
void launch_mix1(const R2CArgs& a, const C2RArgs& k, float2* ynow, unsigned int* ticket, cudaStream_t st);

int fft_smem_bytes(int B);

// What a launch_* call launches: the kernel instantiation, its grid and its dynamic shared memory.  The launchers
// are describe_* + cudaLaunchKernel; the CUDA-graph replay of a whole pull (pgx_api.cu) builds its kernel nodes
// from the same descriptions, so both paths run identical kernels.  describe_* also raises the kernel's dynamic
// shared-memory limit when it needs more than 48 KB.  false = no such kernel for these arguments.
struct LaunchDesc {
  const void* func = nullptr;
  dim3 grid{1, 1, 1}, block{1, 1, 1};
  unsigned smem = 0;
};
bool describe_r2c_ingest(const R2CArgs& a, LaunchDesc* d);                    // params: (R2CArgs, FilterPrepArgs)
bool describe_c2r_emit(const C2RArgs& a, LaunchDesc* d);                      // params: (C2RArgs)
bool describe_conv1(const R2CArgs& a, const C2RArgs& k, LaunchDesc* d);       // params: (R2CArgs, C2RArgs)
bool describe_mix1(const R2CArgs& a, bool last, LaunchDesc* d);               // params: (R2CArgs, C2RArgs, float2*, unsigned*)
bool describe_fdl_mac(const MacArgs& a, LaunchDesc* d);                       // params: (MacArgs); LDG kernel only
bool describe_reduce_partials(int n_out, int W4, LaunchDesc* d);              // params: (in, out, int n_split, int64 n_cols)
bool describe_conv1_r16(const R2CArgs& a, const C2RArgs& k, LaunchDesc* d);   // params: (R2CArgs, C2RArgs)

}  // namespace pgx
