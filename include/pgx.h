/*
 * pgx.h -- C ABI of libpgx.so, the B200 (sm_100a) device path for pygmu2's
 * ConvolvePE / SpatialHRTF / MixPE hot path.
 *
 * pygmu2 is pure Python and has no FFI of its own: its only
 * plugin surface is the ProcessingElement protocol.  The entry points below are
 * therefore what a maintainer-added binding for this path would call; each one
 * cites the reference code it replaces (paths relative to the reference tree).
 * The ctypes stub that binds them is shown in INTEGRATION.md and shipped as
 * pygmu2_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers, ints, no torch / numpy types in any signature;
 *   - every function returns PGX_OK (0) or a negative pgx_status; the message is
 *     available per thread from pgx_last_error();
 *   - a handle may be used by one thread at a time, from any thread (the library
 *     calls cudaSetDevice itself: reference audio_renderer.py:214-236 pulls from
 *     PortAudio's thread);
 *   - "stream" below means an audio stream (one ConvolvePE / one HRTF source),
 *     "cuda_stream" is a cudaStream_t passed as void* (NULL = the bank's own; the legacy default stream, whose
 *     handle is 0, therefore cannot be named: callers on it use a side stream, see dist.ShardedMix).
 */
#ifndef PGX_H_
#define PGX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: pgx_bank_config grew tail_block (round 1, unversioned at the time); multi-GPU mix reduce (pgx_comm_*,
 *    pgx_mix_reduce), filter trajectories, graph replay.  The binding refuses a library of another version. */
#define PGX_ABI_VERSION 2

#if defined(__GNUC__)
#define PGX_API __attribute__((visibility("default")))
#else
#define PGX_API
#endif

typedef enum pgx_status {
  PGX_OK = 0,
  PGX_ERR_INVALID = -1, /* contract violation  -> Python ValueError   */
  PGX_ERR_CUDA = -2,    /* CUDA runtime failure -> Python RuntimeError */
  PGX_ERR_NO_DEVICE = -3,
  PGX_ERR_NOMEM = -4
} pgx_status;

/* flags for pgx_bank_config.flags */
#define PGX_FLAG_MIXDOWN_INPUT 1u /* mean over c_in channels -> 1 internal channel (spatial_pe.py:483) */

typedef struct pgx_bank pgx_bank; /* opaque: device-resident state of N lock-stepped audio streams */

/*
 * Element (stream s, channel c, sample i) of a caller buffer lives at
 *   base[s*stream + c*chan + i*samp]   (strides in floats).
 * Snippet data (samples, channels) of one PE is {0, 1, C}; a planar bank buffer
 * [N][C][n] is {C*n, n, 1}.  The addressed elements must tile one dense block.
 */
typedef struct pgx_layout {
  int64_t stream;
  int64_t chan;
  int64_t samp;
} pgx_layout;

typedef struct pgx_bank_config {
  int32_t device;          /* CUDA device ordinal */
  int32_t n_streams;       /* N >= 1 independent streams processed in lockstep */
  int32_t c_in;            /* source channels per stream */
  int32_t c_out;           /* output channels per stream (convolve_pe.py:114-144, :207-223) */
  int32_t filter_len;      /* L >= 1 taps */
  int32_t filter_channels; /* 1 (mono filter on every channel) or c_out */
  int32_t n_filters;       /* F distinct filters resident on the device (1 = shared) */
  int32_t block;           /* partition size B: power of two in [16, 8192]; FFT size is 2B */
  int32_t max_pull;        /* largest n accepted by one process call (staging size) */
  uint32_t flags;          /* PGX_FLAG_* */
  int32_t tail_block;      /* 0 = uniform partitions of `block` (the default).  > 0: two-level partitioning for long
                              filters pulled in small blocks: the first tail_block taps at `block`, the rest at
                              tail_block (power of two in [block, 8192]) - same result, the long tail is streamed once
                              per tail_block samples instead of once per block.  Such a bank keeps its filters and
                              its filter map, and does not switch between per-stream and mixed pulls without a reset. */
} pgx_bank_config;

typedef struct pgx_bank_info {
  int32_t n_streams, c_in, c_x, c_out, filter_len, filter_channels, n_filters;
  int32_t block, partitions, max_pull, device;
  int32_t head, fill;       /* ring position: current delay-line slot, samples in the open block */
  int64_t state_bytes;      /* device bytes held (delay line + filter spectra + staging) */
  int64_t kernel_launches;  /* kernels launched by this handle since creation */
  int64_t block_steps;      /* FFT->MAC->IFFT steps executed since creation */
  int32_t mac_grid, mac_split, mac_stream_tile, mac_occupancy; /* launch plan of the accumulate kernel */
  int32_t tail_block, tail_partitions; /* two-level partitioning: block and partition count of the tail level (0 = off) */
  int32_t mac_tile;         /* time tiling of the conv pass: past sums of this many consecutive blocks per pass over the
                               delay line (1 = one pass per block; PGX_TILE) */
  int32_t submit_depth;     /* host-buffer pulls this bank keeps in flight (pgx_bank_submit): 3, PGX_SUBMIT_DEPTH when tiled */
  int64_t graph_pulls;      /* host pulls that ran as ONE CUDA-graph replay (small whole-block pulls in the steady state:
                               copies and kernels of the pull as one launch; PGX_GRAPH=0 disables) */
} pgx_bank_info;

/* ---- library ------------------------------------------------------------ */
PGX_API int pgx_abi_version(void);
/* sizeof() of the structs of this header as the library was compiled, so that a binding can check its own
 * mirror at load time: which = 0 pgx_layout, 1 pgx_bank_config, 2 pgx_bank_info, 3 pgx_profile, 4 pgx_osc_config;
 * -1 for an unknown index. */
PGX_API int pgx_struct_size(int32_t which);
PGX_API const char* pgx_last_error(void);
PGX_API int pgx_device_count(int* count);
/* pinned host memory for the e2e path (cudaHostAlloc / cudaFreeHost) */
PGX_API int pgx_host_alloc(void** ptr, int64_t bytes);
/* Same with flags: PGX_HOST_WRITE_COMBINED = write-combined pinned memory for buffers the host only WRITES and the device
 * reads (input staging): not snooped on its way over PCIe, at the price of very slow host reads. */
#define PGX_HOST_WRITE_COMBINED 1
PGX_API int pgx_host_alloc_flags(void** ptr, int64_t bytes, int32_t flags);
PGX_API int pgx_host_free(void* ptr);

/* device memory for inputs that stay resident (ArrayPE sources uploaded once; see PGX_PULL_X_DEVICE).
 * Synchronous, like cudaMalloc / cudaMemcpy / cudaMemset. */
PGX_API int pgx_device_alloc(int32_t device, int64_t bytes, void** ptr);
PGX_API int pgx_device_free(int32_t device, void* ptr);
PGX_API int pgx_device_upload(int32_t device, void* dst_dev, const void* src_host, int64_t bytes);
PGX_API int pgx_device_zero(int32_t device, void* dst_dev, int64_t bytes);

/* ---- bank: replaces ConvolvePE state + _render (convolve_pe.py:185-342) and
 *      SpatialHRTF.render (spatial_pe.py:465-518) for N streams at once ------ */

/*
 * Prepare filters and allocate state.  Replaces ConvolvePE._ensure_filter_prepared
 * (convolve_pe.py:185-248: render the FIR once, rfft it, zero the tail) with a
 * uniformly partitioned spectrum set per filter.
 *   h                 host, [n_filters][filter_channels][filter_len] float32
 *   filter_of_stream  host, [n_streams] indices into the filter set, or NULL for
 *                     stream s -> filter (s % n_filters)
 */
PGX_API int pgx_bank_create(pgx_bank** out, const pgx_bank_config* cfg, const float* h,
                    const int32_t* filter_of_stream);
PGX_API int pgx_bank_destroy(pgx_bank* bank);
PGX_API int pgx_bank_get_info(pgx_bank* bank, pgx_bank_info* info);

/*
 * History := 0.  k == 0 resets every stream and re-anchors the block grid
 * (ConvolvePE._reset_state / non-contiguous pull, convolve_pe.py:146-154,255-256;
 * SpatialHRTF._reset_tail_if_noncontiguous, spatial_pe.py:461-463); k > 0 clears
 * only the listed streams' history.
 */
PGX_API int pgx_bank_reset(pgx_bank* bank, const int32_t* stream_ids, int32_t k);

/*
 * Replace resident filter `filter_index` by new taps h [filter_channels][filter_len]
 * (host) and re-derive its partition spectra.  A per-PE SpatialHRTF whose azimuth /
 * elevation attributes were mutated between pulls uses this (spatial_pe.py:446-459).
 */
PGX_API int pgx_bank_load_filter(pgx_bank* bank, int32_t filter_index, const float* h);

/*
 * Re-select the filter of every stream for the following pulls (a moving HRTF
 * source: spatial_pe.py:446-449 re-resolves the IR on every render).
 */
PGX_API int pgx_bank_set_filter_map(pgx_bank* bank, const int32_t* filter_of_stream);
/* Same, but the [n_streams] map already lives on the device (e.g. one row of a resident
 * trajectory table); NULL switches back to the bank's own map.  No copy, no synchronisation. */
PGX_API int pgx_bank_use_filter_map_device(pgx_bank* bank, const int32_t* filter_of_stream_dev);

/*
 * One pull of n samples for all streams: y = x * h with carried history.
 * Replaces ConvolvePE._render (convolve_pe.py:250-342) / SpatialHRTF.render
 * (spatial_pe.py:465-518).  Host buffers; H2D and D2H copies happen inside and the
 * call returns when y is complete.  1 <= n <= max_pull.
 */
PGX_API int pgx_bank_process(pgx_bank* bank, const float* x, pgx_layout x_layout, float* y,
                     pgx_layout y_layout, int32_t n);

/*
 * Same pull, fused with the MixPE sum over streams (mix_pe.py:92-94):
 * y_mix[c*chan + i*samp] = sum_s y[s][c][i]  (y_layout.stream is ignored).
 */
PGX_API int pgx_bank_process_mix(pgx_bank* bank, const float* x, pgx_layout x_layout, float* y_mix,
                         pgx_layout y_layout, int32_t n);

/*
 * Fused output stage for the following pulls: y = dry * x + wet * (x * h), each product and the sum rounded to
 * float32.  Replaces the GainPE(dry) + GainPE(wet) -> MixPE tail of ReverbPE (reverb_pe.py:82-95,
 * gain_pe.py:123-125, mix_pe.py:92-94) - bit-identical to those three float32 steps given the same
 * convolution output.  dry != 0 needs c_in == c_out without mix-down; in a fused-mix pull only wet applies.
 * Default wet = 1, dry = 0.
 */
PGX_API int pgx_bank_set_output_gains(pgx_bank* bank, float wet, float dry);

/* flags of a pull */
#define PGX_PULL_MIX 1u            /* fused MixPE sum over streams, as pgx_bank_process_mix */
#define PGX_PULL_INPUT_RESIDENT 2u /* x is already complete in memory (not produced by work still queued on
                                      cuda_stream): the ingest of this pull may overlap earlier pulls' output stage */
#define PGX_PULL_X_DEVICE 4u       /* pgx_bank_submit only: x is a DEVICE pointer whose producer was enqueued on the
                                      bank's own stream (pgx_bank_stream) - e.g. pgx_osc_render_device - so there is
                                      no H2D copy; y is still delivered to host memory */
#define PGX_PULL_X_PCM16 8u        /* pgx_bank_submit only: host x holds int16 PCM samples (same layout, in int16
                                      elements); float32 = int16/32768 on the device - what soundfile.read(dtype=
                                      "float32") gives WavReaderPE (wav_reader_pe.py:127-132) - half the H2D bytes */
#define PGX_PULL_Y_PCM16 16u       /* pgx_bank_submit only: y is delivered as int16 PCM, clip(lrintf(y*32768)) on the
                                      device - libsndfile's float -> PCM_16 rule with clipping on, as python-soundfile
                                      writes for WavWriterPE (wav_writer_pe.py:153) - half the D2H bytes */

#define PGX_PULL_REDUCE 32u        /* with PGX_PULL_MIX on a bank that has a communicator attached (pgx_bank_attach_comm):
                                      this rank's mix is a partial one; it is summed over the ranks onto the root behind
                                      the pull's output stage (pgx_mix_reduce) and only the root delivers y */

#define PGX_SUBMIT_DEPTH 8

/*
 * Pipelined host-buffer pulls (the batched renderer loop, renderer.py:297-327, with more than one pull in
 * flight): submit stages x (H2D on a copy stream), enqueues the pull and the D2H of y, and returns a ticket
 * without waiting; pgx_bank_wait(ticket) returns when that pull's y is complete in host memory.  Pulls
 * execute in submission order.  x and y must stay valid (and should be pinned, pgx_host_alloc) until the
 * wait returns; at most pgx_bank_info.submit_depth pulls are in flight (3; PGX_SUBMIT_DEPTH on a time-tiled bank,
 * whose passes cover several blocks) - a further submit first waits for the oldest.
 * flags: PGX_PULL_MIX, PGX_PULL_X_DEVICE, PGX_PULL_X_PCM16, PGX_PULL_Y_PCM16, PGX_PULL_REDUCE (on the ranks
 * other than the root y is not written: there is no D2H copy).  pgx_bank_process[_mix] = submit + wait.
 */
PGX_API int pgx_bank_submit(pgx_bank* bank, const float* x, pgx_layout x_layout, float* y, pgx_layout y_layout,
                    int32_t n, int32_t flags, int64_t* ticket);
PGX_API int pgx_bank_wait(pgx_bank* bank, int64_t ticket);
/* pgx_bank_submit + pgx_bank_wait of that ticket in one call (one FFI crossing per pull for callers that do not
 * pipeline: the PE shims' render()).  Same flags. */
PGX_API int pgx_bank_pull(pgx_bank* bank, const float* x, pgx_layout x_layout, float* y, pgx_layout y_layout,
                  int32_t n, int32_t flags);
/* The bank's own (critical) CUDA stream as a cudaStream_t: producers of device-resident input enqueue on it. */
PGX_API void* pgx_bank_stream(pgx_bank* bank);

/* Device-resident variant: x / y are device pointers; work is enqueued and NOT synchronised; y is
 * complete when cuda_stream (NULL = the bank's stream) drains.  flags: PGX_PULL_*. */
PGX_API int pgx_bank_process_device(pgx_bank* bank, const float* x_dev, pgx_layout x_layout, float* y_dev,
                            pgx_layout y_layout, int32_t n, int32_t flags, void* cuda_stream);
PGX_API int pgx_bank_synchronize(pgx_bank* bank);

/* ---- multi-GPU: the one exchange step of the path, the MixPE sum (mix_pe.py:92-94) over stream shards ------------
 * Streams are sharded over the GPUs of one box (one process per GPU, or several devices in one process); every
 * rank's fused mix pull yields a partial (C_out, n) mix and the partials are summed onto `root` in RANK ORDER
 * (float32, deterministic) through peer memory over NVLink: a non-root rank stores its partial into the root's
 * mailbox and raises a flag, the root's gather kernel waits for the flags, adds and acknowledges -- one small
 * kernel per rank and pull on the bank's stream, no host synchronisation, no library collective.
 *   1. every rank: pgx_comm_create -> a PGX_COMM_HANDLE_BYTES blob describing its mailbox
 *   2. the host's plumbing all-gathers the blobs (torch.distributed / MPI / a pipe: any byte transport)
 *   3. every rank: pgx_comm_connect(all blobs in rank order) maps the peers (CUDA IPC across processes)
 * Every rank must issue the same sequence of reduces.  A rank that waits 20 s for a peer gives up and the next
 * pgx_comm_check / pgx_bank_wait reports it. */
#define PGX_COMM_HANDLE_BYTES 128
typedef struct pgx_comm pgx_comm;
PGX_API int pgx_comm_create(pgx_comm** out, int32_t device, int32_t rank, int32_t world, int32_t root,
                            int32_t max_floats, void* handle_out);
PGX_API int pgx_comm_connect(pgx_comm* comm, const void* handles);
PGX_API int pgx_comm_check(pgx_comm* comm);
PGX_API int pgx_comm_destroy(pgx_comm* comm);
/* Sum the ranks' partials part_dev[0..n) onto the root's y_dev (device pointers; y_dev may alias part_dev and is
 * ignored on the other ranks), enqueued on cuda_stream.  n <= max_floats. */
PGX_API int pgx_mix_reduce(pgx_comm* comm, const float* part_dev, float* y_dev, int32_t n, void* cuda_stream);
/* Mix pulls of this bank that carry PGX_PULL_REDUCE are reduced through `comm` (NULL detaches). */
PGX_API int pgx_bank_attach_comm(pgx_bank* bank, pgx_comm* comm);

/* ---- measurement: per-kernel device time, CUDA events on the launching stream ---- */
typedef struct pgx_profile {
  double ms_r2c, ms_mac, ms_c2r; /* summed durations of K1 / K3 (past-partition pass) / K2 launches */
  int64_t steps;                 /* block steps covered */
  double ms_fold, ms_now;        /* fold of split partials; K3 over the present slot (mix mode) */
  int64_t n_mac;                 /* K3 past-pass launches timed */
  double ms_mac_union;           /* time during which at least one of them was running (launches of consecutive
                                    blocks overlap on two streams) */
  double ms_conv1;               /* fused K1+K2 launches of single-partition (P = 1) conv pulls */
} pgx_profile;
/* begin: every following block step records events around each of its three kernels.
 * end: synchronise, sum the durations into *out, stop recording. */
PGX_API int pgx_bank_profile_begin(pgx_bank* bank);
PGX_API int pgx_bank_profile_end(pgx_bank* bank, pgx_profile* out);

/* ---- device-resident sources: the inputs of the path rendered in HBM -------------
 * SINE: n_voices constant-parameter SinePE streams (sine_pe.py:135-175): float64 phase from the sample
 *       index, float32 out.  freq / gain (= amplitude) / phase (radians) are [n_voices]; unison is ignored.
 * BLIT: n_voices voices of `unison` band-limited sawtooth oscillators each: BlitSawPE (unison = 1,
 *       blit_saw_pe.py:152-264) and SuperSawPE (super_saw_pe.py:282-305).  freq / gain (oscillator amplitude) /
 *       phase (initial phase in [0,1)) are [n_voices*unison], m_fixed [n_voices*unison] harmonics (0 = auto)
 *       or NULL, amp [n_voices] voice amplitude.  Stateful: a pull whose start is not the previous pull's end
 *       re-initialises phase and integrator (blit_saw_pe.py:183-186).
 * Output is planar [n_voices][channels][n] (channels are copies, as np.tile in the reference) or, with
 * PGX_PULL_MIX, the MixPE sum over voices [channels][n] in float32 input order (mix_pe.py:92-94, bit-exact).
 */
#define PGX_OSC_SINE 0
#define PGX_OSC_BLIT 1
typedef struct pgx_osc pgx_osc;
typedef struct pgx_osc_config {
  int32_t device, kind, n_voices, unison, channels, sample_rate, max_pull, reserved;
  double leak; /* BLIT leaky-integrator coefficient (blit_saw_pe.py:75, default 0.999) */
} pgx_osc_config;
PGX_API int pgx_osc_create(pgx_osc** out, const pgx_osc_config* cfg, const double* freq, const double* gain,
                   const double* phase, const int32_t* m_fixed, const double* amp);
PGX_API int pgx_osc_destroy(pgx_osc* osc);
PGX_API int pgx_osc_reset(pgx_osc* osc); /* on_start / on_stop: the next pull starts from the initial state */
/* Enqueue one pull on cuda_stream (NULL = the handle's own); *out_dev points at the handle's output buffer,
 * valid until the next render on this handle.  Not synchronised. */
PGX_API int pgx_osc_render_device(pgx_osc* osc, int64_t start, int32_t n, int32_t flags, void* cuda_stream,
                          const float** out_dev);
/* Same pull delivered to host memory (D2H inside, returns when y is complete). */
PGX_API int pgx_osc_render(pgx_osc* osc, int64_t start, int32_t n, int32_t flags, float* y);
PGX_API int pgx_osc_launches(pgx_osc* osc, int64_t* launches);
/* Speculative pulls: a consumer that has just delivered [s, s+n) may render [s+n, s+2n) right away, behind its own
 * work on the same stream, so that the block is ready when the next contiguous pull arrives (the reference renders a
 * block when it is asked for it, blit_saw_pe.py:150-262; the samples are the same either way).  Such a pull carries
 * PGX_OSC_SNAPSHOT in `flags`: the oscillator state it starts from is kept, and pgx_osc_rollback restores it (enqueued
 * on cuda_stream) when the next pull turns out to be a different one.  Any regular pull commits the speculation. */
#define PGX_OSC_SNAPSHOT 64
PGX_API int pgx_osc_rollback(pgx_osc* osc, void* cuda_stream);
/* Modulated SinePE: any of frequency / amplitude / phase is a PE (sine_pe.py:134-142,188-232: the stateful branch --
 * phase[i] = cumsum(2 pi f[i] / sr)[i] + carried phase (+ phase_mod[i]), float64, np.cumsum's left-to-right order).
 * freq / amp / phase: float32 [n_voices][n], what the parameter PEs rendered for this pull (widened to float64 on the
 * device), or NULL for a parameter that is the constant given to pgx_osc_create.  ctl_flags & PGX_CTL_HOST: the vectors
 * are host memory (staged by the call, which then returns once they have been consumed); otherwise device pointers
 * whose producer was enqueued on cuda_stream.  The phase carries over from pull to pull (the reference's base class
 * only allows contiguous pulls of a stateful PE); pgx_osc_reset starts over.
 * PGX_OSC_BLIT handles (BlitSawPE / SuperSawPE with a PE-valued frequency and / or amplitude, blit_saw_pe.py:161-262,
 * super_saw_pe.py:223-246,287-303): `freq` is the voice's frequency control, each oscillator using float32(freq * ratio)
 * with ratio = its `freq` entry at creation (GainPE(frequency_pe, ratio); pass the detune ratios -- 1 for a BlitSawPE --
 * when the frequency will be a PE); `amp` see PGX_CTL_AMP_OSC; the third vector (`phase`) is the harmonic-count control of a
 * PE-valued `m` (blit_saw_pe.py:175-177: max(int32(m), 1) per sample), or NULL.  Output as pgx_osc_render_device
 * (*out_dev, may be NULL) and / or, when y_host is not NULL, copied to host memory before the call returns. */
#define PGX_CTL_HOST 1
#define PGX_CTL_AMP_OSC 2 /* BLIT handles: the amplitude control replaces the OSCILLATOR amplitude (BlitSawPE: saw * 2 * amp,
                             blit_saw_pe.py:252-256); without it, it replaces the VOICE amplitude that multiplies the
                             float64 sum of the oscillators (SuperSawPE, super_saw_pe.py:287-303) */
PGX_API int pgx_osc_render_modulated(pgx_osc* osc, int32_t n, int32_t flags, const float* freq, const float* amp,
                                     const float* phase, int32_t ctl_flags, void* cuda_stream, const float** out_dev,
                                     float* y_host);

/* ---- host helper: the per-pull direction -> HRTF entry search of a moving source ------------------------
 * SpatialHRTF.hrtf_filename_for (spatial_pe.py:395-426) for n directions at once: az := min(180, |az|), least squared
 * distance in (elevation, azimuth) over the n_tab table entries, FIRST minimum in table order; plain float64 host
 * arithmetic in the reference's order (no device involved).  256 moving sources re-select their filters before every
 * pull: this is the host cost of that loop. */
PGX_API int pgx_nearest_direction(const double* tab_elev, const double* tab_az, int32_t n_tab, const double* azimuth,
                                  const double* elevation, int32_t n, int32_t* out_index);

/* ---- MixPE: replaces the float32 left-to-right sum of mix_pe.py:92-94 ------ */
/*
 * out[e] = ((in_0[e] + in_1[e]) + in_2[e]) + ...  in float32, in input order, for
 * e in [0, n_elems).  inputs: host, [n_inputs][n_elems] dense.  Bit-exact with
 * numpy's sequential "+=".
 */
PGX_API int pgx_mix_sum(int32_t device, const float* inputs, int32_t n_inputs, int64_t n_elems, float* out);
PGX_API int pgx_mix_sum_device(int32_t device, const float* inputs_dev, int32_t n_inputs, int64_t n_elems,
                       float* out_dev, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PGX_H_ */
