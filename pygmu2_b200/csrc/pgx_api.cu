// pgx_api.cu -- host engine + C ABI of libpgx.so (see include/pgx.h).
//
// A bank owns, on one GPU, for N lock-stepped audio streams:
//   hist  [N*c_x][2][B]       float   previous block / open block (time domain)
//   fdl   [N*c_x][R][B]       float2  frequency-domain delay line: packed spectra of the last windows,
//                                     a ring of R = P+1 rows (R = 1 when P = 1): P live rows + one spare
//                                     (R = P+T-1 on a time-tiled bank: T-1 spares)
//   Hd    [F*c_f][2R][B]      float2  filter partition spectra, reversed + doubled, scaled 1/B
//   ypast [2][<=8][n_out][B]  float2  sums over the past partitions of the open block (by block parity)
//   ypart [n_split][n_out][B] float2  split partials of the background pass in flight
//   ynow  [n_split][c_out][B] float2  present-slot partials (mix mode)
//   ytile [2T][n_split][n_out][B] float2  time-tiled banks: result sets of the tiled passes (by block mod 2T)
// and advances them one "block step" at a time.  A pull of n samples is cut at block boundaries; a
// partially filled block is transformed with zeros in the not-yet-known positions (causality makes the
// emitted samples exact) and re-transformed when more samples arrive, so any (start, duration) pull
// pattern is zero-latency like the reference (convolve_pe.py:250-342), while the block grid stays
// aligned for the partitioned filter.
//
// Schedule of block step i (block t), three CUDA streams joined by events:
//   ingest stream     : K1(i)  ingest + R2C of the open block -> ring row slot(t)
//   background stream : K3 over the P-1 *past* partitions of block t+1, launched as soon as the step that
//                       completes block t has its K1; it needs only committed rows, overlaps everything
//                       else and is reused by every partial pull of block t+1
//   critical stream   : [mix mode: K3 over the open slot of every stream] -> K2(i): past sums + present
//                       term, C2R, emit.  This is the caller's stream: y is complete when it drains.
// The spare ring row lets K1 of block t+1 run while the background pass of block t is still reading, so
// in a back-to-back queue of pulls the step time is the accumulate kernel alone.  Hazards and the event
// that orders each one are listed at run_step().
//
// Time tiling (conv pulls of banks whose pass is HBM-bound, bank->tile = T > 1): the background pass runs once per T
// blocks and yields the past sums of blocks t0 .. t0+T-1 from one read of the committed rows (issue_tile); K2 of block
// t0+kappa adds the kappa rows committed since (C2RArgs.n_recent).  Coverage is dropped by anything that changes what a
// past sum would be (reset, filter map, filter reload, a mix pull) and re-established on demand by the next pull.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "host_util.h"
#include "kernels.h"

namespace {
thread_local std::string g_err;
}

int pgx_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define fail pgx_fail

namespace {

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// The addressed elements of a (streams, chans, n) box must tile exactly one dense block.
bool layout_dense(const pgx_layout& l, int64_t S, int64_t C, int64_t n) {
  int64_t dims[3] = {S, C, n};
  int64_t str[3] = {l.stream, l.chan, l.samp};
  int64_t span = 1;
  for (int i = 0; i < 3; ++i) {
    if (dims[i] > 1) {
      if (str[i] <= 0) return false;
      span += (dims[i] - 1) * str[i];
    }
  }
  return span == S * C * n;
}

// K2 folds up to this many split partials itself; beyond that a fold kernel runs right after the
// background pass, off the critical path.
constexpr int kFoldAbove = 8;
constexpr int kRing = 8;  // event ring depth (steps / blocks in flight)

}  // namespace

struct pgx_bank {
  pgx_bank_config cfg{};
  int c_x = 1, P = 1, R = 1, B = 0;
  cudaStream_t stream = nullptr;   // default critical stream
  cudaStream_t s_in = nullptr;     // ingest stream (K1)
  cudaStream_t s_bg = nullptr;     // background stream (past-partition pass) of even blocks ...
  cudaStream_t s_bg2 = nullptr;    // ... and of odd blocks: the passes of consecutive blocks share nothing they
                                   // write, so the tail wave of one overlaps the first wave of the next
  cudaEvent_t ev_bgjoin = nullptr;
  bool two_bg = true;
  cudaEvent_t ev_call = nullptr;
  cudaEvent_t ev_k1[kRing] = {}, ev_k2[kRing] = {}, ev_mac[kRing] = {};
  float* hist = nullptr;
  float2* fdl = nullptr;
  float2* Hd = nullptr;
  float2* ypast[2] = {nullptr, nullptr};
  float2* ypart[2] = {nullptr, nullptr};  // by block parity, like ypast
  float2* ynow = nullptr;
  float2* tw = nullptr;
  int32_t* fmap = nullptr;         // map in use (own buffer or the caller's device array)
  // the bank's own maps: a ring of kMapSlots device rows + pinned staging rows, so that re-selecting the
  // filters between pulls (a moving HRTF source) neither drains the queue nor races with pulls in flight
  static constexpr int kMapSlots = 8;
  int32_t* fmap_own = nullptr;     // kMapSlots x [N] device
  int32_t* fmap_pinned = nullptr;  // kMapSlots x [N] pinned staging
  cudaEvent_t fmap_ev[kMapSlots] = {};        // H2D of the row done (copy-in stream)
  cudaEvent_t fmap_ret_crit[kMapSlots] = {};  // row retired: every reader was enqueued before these
  cudaEvent_t fmap_ret_bg[kMapSlots] = {};
  int fmap_cur = 0;
  int64_t fmap_sets = 0;
  bool fmap_pending = false;       // the next pull must order itself after fmap_ev[fmap_cur]
  cudaStream_t last_crit = nullptr;
  // host-buffer pulls: a ring of staging slots so that the H2D of pull i+1 and the D2H of pull i-1 overlap
  // the kernels of pull i (copy engines on their own streams)
  static constexpr int kSlots = PGX_SUBMIT_DEPTH;
  int n_slots = 3;                 // pulls in flight on this bank: 3, kSlots on a time-tiled bank (its passes cover `tile` blocks)
  float* x_stage[kSlots] = {};
  float* y_stage[kSlots] = {};
  int16_t* xpcm_stage[kSlots] = {};  // PCM16 staging, allocated on first use
  int16_t* ypcm_stage[kSlots] = {};
  // small pulls (a single PE graph: a few KB per pull) bounce through pinned host buffers owned by the bank, so that
  // the copies are truly asynchronous whatever memory the caller's arrays live in (a cudaMemcpyAsync from / to
  // pageable memory is staged by the driver and, device-to-host, blocks the calling thread)
  static constexpr size_t kBounceMax = 256 * 1024;
  static constexpr size_t kZeroCopyMax = 32 * 1024;  // graph replays of pulls this small read / write the bounce buffers in place
  char* hx_bounce[kSlots] = {};
  char* hy_bounce[kSlots] = {};
  void* y_user[kSlots] = {};       // pending copy-back: hy_bounce[slot] -> y_user[slot] once ev_done[slot] has fired
  size_t y_user_bytes[kSlots] = {};
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[kSlots] = {}, ev_y[kSlots] = {}, ev_done[kSlots] = {};
  int64_t next_ticket = 0;
  size_t hist_bytes = 0, fdl_bytes = 0, Hd_bytes = 0, ypart_bytes = 0, ysum_bytes = 0, ynow_bytes = 0, xs_bytes = 0,
         ys_bytes = 0;
  // ring / schedule state
  int head = 0, fill = 0, half = 0;
  int64_t step = 0;                // steps issued since the last full reset
  int64_t block = 0;               // index of the open block since the last full reset
  int64_t past_block = -1;         // block whose past sum was last issued ...
  int past_mode = -1;              // ... for this mode (0 conv, 1 mix)
  int64_t last_k2_of_par[2] = {-1, -1};  // last step whose K2 read ypast[par]
  bool prev_on_crit = false;       // the previous step ran ingest + output as one kernel on the critical stream
  pgx::MacPlan plan_conv{}, plan_mix{}, plan_now{};
  // time tiling of the conv pass (see issue_tile): one pass over the committed rows yields the past sums of `tile`
  // consecutive blocks; result sets are indexed by (block mod 2*tile)
  int tile = 1;
  int n_spare = 1;                 // ring rows beyond the P partitions (R = P + n_spare)
  pgx::MacPlan plan_tile{};
  float2* ytile = nullptr;         // [2*tile][n_split][n_out][B]
  size_t tile_set_elems = 0, ytile_bytes = 0;
  int64_t tile_base = -1;          // first block covered by the newest tiled pass (-1: none valid)
  static constexpr int kTilePasses = 4;
  struct TilePass { int64_t base = -1; cudaEvent_t ev = nullptr; } tpass[kTilePasses];
  int64_t n_tpass = 0;
  int64_t last_k2_of_set[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  // waits a stream has already made need not be repeated (streams are in-order): the newest tiled pass (by issue
  // count) the ingest stream / the critical stream of the last pull has waited for
  int64_t sin_waited_tpass = -1, crit_waited_tpass = -1;
  cudaStream_t crit_waited_stream = nullptr;
  int64_t last_past_issue_block = -1;   // newest block a per-block pass (issue_past) was issued for
  int sm_count = 148;
  float wet = 1.0f, dry = 0.0f;    // fused output stage: y = dry * x + wet * conv
  // two-level partitioning (cfg.tail_block > 0): this bank convolves with the first tail_B taps at block B; `tail`
  // convolves with the rest at block tail_B, one big block at a time, and its output is the addend of the next
  // tail_B output samples (h = [head | tail]: the tail's contribution to y[t .. t+tail_B) only needs x[.. t))
  pgx_bank* tail = nullptr;
  int tail_B = 0, tail_fill = 0, tail_cur = 0, tail_mode = -1;   // tail_mode: -1 undecided, 0 conv, 1 mix
  int full_filter_len = 0;
  float* xacc = nullptr;           // [N][c_in][tail_B] the big block being filled
  float* ytail[2] = {nullptr, nullptr};  // [N or 1][c_out][tail_B] tail contribution of the current / next big block
  size_t xacc_bytes = 0, ytail_bytes = 0;
  const float* addend = nullptr;   // set by the two-level driver for the next run_pull1 of this (head) bank
  pgx_layout addend_l{};
  bool serial = false;
  bool use_conv1 = true;           // P = 1 conv pulls: K1 and K2 fused into one kernel (PGX_CONV1=0 disables)
  int fft16 = 0;                   // B = 4096 fused step: radix-16 kernel variant (PGX_FFT16; see k_fft16.cu)
  int fused_max_p = 16;            // ... and conv pulls of banks with up to this many partitions (PGX_FUSED_MAXP)
  bool use_mix1 = true;            // P = 1 mixes of mono transforms: K1 + present-slot accumulate fused (PGX_MIX1=0)
  int mix1_rows = 0;               // partial rows per channel written by k_mix1 (= CTAs)
  unsigned int* mix1_ticket = nullptr;  // last-CTA ticket counter (k_mix1<LAST>)
  int64_t launches = 0, steps = 0;
  pgx_comm* comm = nullptr;        // cross-GPU mix reduce for pulls that carry PGX_PULL_REDUCE (not owned)
  // CUDA-graph replay of a whole small host pull (see graph_pull): one instantiated graph per staging slot and
  // shape; kernel-node arguments are refreshed before every launch, the topology never changes
  struct PullGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t n_h2d = nullptr, n_a = nullptr, n_mac = nullptr, n_fold = nullptr, n_k2 = nullptr, n_d2h = nullptr;
    const void* f_a = nullptr; const void* f_mac = nullptr; const void* f_k2 = nullptr;
    int shape = -1;                // 1 one fused kernel (k_conv1 / k_mix1<LAST>), 3 K1 -> K2 with the next block's past pass beside it
    bool fold = false;
    size_t xb = 0, yb = 0;
  };
  PullGraph pgraph[kSlots][2][2];  // [slot][mix][x already on the device]
  bool use_graph = true;           // PGX_GRAPH=0 disables
  bool graph_fuse = false;         // graph shape 3: K1 and K2 as ONE k_conv1 launch that also adds the folded past sum
                                   // (PGX_GRAPH_FUSE=1).  Off by default: measured slower in a pull loop -- the next
                                   // block's past pass then waits for the whole fused kernel instead of K1 alone and the
                                   // replays serialise on it (C2 through the PE API 29.5 vs 26.9 us per pull)
  bool after_graph = false;        // the previous step was a graph replay: the event ring must be re-armed before the
                                   // multi-stream schedule continues
  cudaEvent_t ev_join[4] = {};
  int64_t graph_pulls = 0;
  // per-kernel CUDA-event timing
  bool profiling = false;
  struct ProfSpan { int kind; cudaEvent_t a, b; };  // kind 0 = K1, 1 = K3 (past pass), 2 = K2, 3 = fold, 4 = K3 (present slot, mix), 5 = fused K1+K2 (P = 1)
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_pool_used = 0;
  int64_t prof_steps = 0;
};

namespace {

void free_bank(pgx_bank* b) {
  if (!b) return;
  free_bank(b->tail);
  b->tail = nullptr;
  cudaSetDevice(b->cfg.device);
  for (cudaStream_t s : {b->s_h2d, b->stream, b->s_in, b->s_bg, b->s_bg2, b->s_d2h})
    if (s) cudaStreamSynchronize(s);
  cudaFree(b->mix1_ticket);
  cudaFree(b->ytile);
  for (auto& tp : b->tpass)
    if (tp.ev) cudaEventDestroy(tp.ev);
  cudaFree(b->xacc);
  cudaFree(b->ytail[0]);
  cudaFree(b->ytail[1]);
  cudaFree(b->hist);
  cudaFree(b->fdl);
  cudaFree(b->Hd);
  cudaFree(b->ypast[0]);
  cudaFree(b->ypast[1]);
  cudaFree(b->ypart[0]);
  cudaFree(b->ypart[1]);
  cudaFree(b->ynow);
  cudaFree(b->tw);
  cudaFree(b->fmap_own);
  for (int i = 0; i < pgx_bank::kSlots; ++i) {
    cudaFree(b->x_stage[i]);
    cudaFree(b->y_stage[i]);
    cudaFree(b->xpcm_stage[i]);
    cudaFree(b->ypcm_stage[i]);
    if (b->hx_bounce[i]) cudaFreeHost(b->hx_bounce[i]);
    if (b->hy_bounce[i]) cudaFreeHost(b->hy_bounce[i]);
    for (cudaEvent_t e : {b->ev_h2d[i], b->ev_y[i], b->ev_done[i]})
      if (e) cudaEventDestroy(e);
  }
  if (b->fmap_pinned) cudaFreeHost(b->fmap_pinned);
  for (cudaEvent_t e : b->prof_pool) cudaEventDestroy(e);
  for (int i = 0; i < kRing; ++i)
    for (cudaEvent_t e : {b->ev_k1[i], b->ev_k2[i], b->ev_mac[i]})
      if (e) cudaEventDestroy(e);
  for (int i = 0; i < pgx_bank::kMapSlots; ++i)
    for (cudaEvent_t e : {b->fmap_ev[i], b->fmap_ret_crit[i], b->fmap_ret_bg[i]})
      if (e) cudaEventDestroy(e);
  for (int i = 0; i < pgx_bank::kSlots; ++i)
    for (int m = 0; m < 2; ++m)
      for (int d = 0; d < 2; ++d) {
        if (b->pgraph[i][m][d].exec) cudaGraphExecDestroy(b->pgraph[i][m][d].exec);
        if (b->pgraph[i][m][d].graph) cudaGraphDestroy(b->pgraph[i][m][d].graph);
      }
  for (cudaEvent_t e : b->ev_join)
    if (e) cudaEventDestroy(e);
  if (b->ev_call) cudaEventDestroy(b->ev_call);
  if (b->ev_bgjoin) cudaEventDestroy(b->ev_bgjoin);
  for (cudaStream_t s : {b->stream, b->s_in, b->s_bg, b->s_bg2, b->s_h2d, b->s_d2h})
    if (s) cudaStreamDestroy(s);
  delete b;
}

cudaEvent_t prof_event(pgx_bank* b) {
  if (b->prof_pool_used == b->prof_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    b->prof_pool.push_back(e);
  }
  return b->prof_pool[b->prof_pool_used++];
}

struct ProfScope {  // records a CUDA-event pair around one launch when profiling is on
  pgx_bank* b;
  cudaStream_t st;
  int kind;
  cudaEvent_t e0 = nullptr;
  ProfScope(pgx_bank* b_, cudaStream_t st_, int kind_) : b(b_), st(st_), kind(kind_) {
    if (b->profiling) {
      e0 = prof_event(b);
      cudaEventRecord(e0, st);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEvent_t e1 = prof_event(b);
      cudaEventRecord(e1, st);
      b->prof_spans.push_back({kind, e0, e1});
    }
  }
};

void fill_mac_common(pgx_bank* b, pgx::MacArgs& m, bool mix, int head) {
  const pgx_bank_config& c = b->cfg;
  m.fdl = reinterpret_cast<const float4*>(b->fdl);
  m.Hd = reinterpret_cast<const float4*>(b->Hd);
  m.fmap = b->fmap;
  m.N = c.n_streams; m.c_x = b->c_x; m.c_out = c.c_out; m.c_f = c.filter_channels; m.R = b->R; m.W4 = b->B / 2;
  m.q0 = b->R - 1 - head;
  m.mix = mix ? 1 : 0;  // the caller overrides this with its plan's layout
  m.n_out = mix ? c.c_out : c.n_streams * c.c_out;
}

// Background pass for block `blk` whose ring slot is `head`: sum over its P-1 past partitions, i.e. every
// ring row except slot head (the open block) and slot head+1 (the spare row, free for the next block's K1).
// Ordered after `after` (an event on the ingest stream by which every committed row is written) and after
// the last K2 that read the ypast buffer it overwrites.
void issue_past(pgx_bank* b, bool mix, int64_t blk, int head, cudaEvent_t after) {
  const int par = (int)(blk & 1);
  cudaStream_t sbg = (par && b->two_bg) ? b->s_bg2 : b->s_bg;
  cudaStreamWaitEvent(sbg, after, 0);
  if (b->last_k2_of_par[par] >= 0) cudaStreamWaitEvent(sbg, b->ev_k2[b->last_k2_of_par[par] % kRing], 0);
  pgx::MacArgs m{};
  fill_mac_common(b, m, mix, head);
  const pgx::MacPlan& pl = mix ? b->plan_mix : b->plan_conv;
  const bool fold = pl.n_partials > kFoldAbove;
  m.mix = pl.layout;
  m.yspec = reinterpret_cast<float4*>(fold ? b->ypart[par] : b->ypast[par]);
  m.Pt = b->R - 2;  // = P - 1
  m.jfix = -1;
  if (head + 1 < b->R) { m.off = 0; m.skip = head; m.nskip = 2; }
  else                 { m.off = 1; m.skip = b->R; m.nskip = 0; }  // open slot R-1, spare slot 0
  m.n_terms = pl.layout == 1 ? b->cfg.n_streams * m.Pt : m.Pt;
  m.n_split = pl.n_split; m.terms_per_split = pl.terms_per_split; m.n_otiles = pl.n_otiles; m.st = pl.st;
  m.variant = pl.variant; m.persistent_ctas = pl.persistent_ctas;
  {
    ProfScope ps(b, sbg, 1);
    pgx::launch_fdl_mac(m, sbg);
    b->launches += 1;
  }
  if (fold) {
    ProfScope ps(b, sbg, 3);
    pgx::launch_reduce_partials(reinterpret_cast<const float4*>(b->ypart[par]),
                                reinterpret_cast<float4*>(b->ypast[par]), pl.n_partials, m.n_out, b->B / 2, sbg);
    b->launches += 1;
  }
  cudaEventRecord(b->ev_mac[blk % kRing], sbg);
  if (blk > b->last_past_issue_block) b->last_past_issue_block = blk;
  b->past_block = blk;
  b->past_mode = mix ? 1 : 0;
}

// Time-tiled background pass: the past sums of blocks blk0 .. blk0+tile-1 (slot of blk0 = head0) from the rows
// committed so far, i.e. every ring row except the open slot head0 and the spare slot head0+1 -- the same rows as
// issue_past(blk0), read ONCE for `tile` outputs.  Block blk0+kappa then still lacks its kappa youngest committed
// rows and its present term: the output stage adds them (C2RArgs.n_recent).  Passes run on one background stream
// (they are `tile` times rarer than per-block passes), so they complete in issue order.
void issue_tile(pgx_bank* b, int64_t blk0, int head0, cudaEvent_t after) {
  const int T = b->tile, nsets = 2 * T;
  cudaStream_t sbg = b->s_bg;
  cudaStreamWaitEvent(sbg, after, 0);
  for (int kp = 0; kp < T; ++kp) {  // WAR: the sets being overwritten were read by the output stages of blocks 2*tile earlier
    int64_t lk = b->last_k2_of_set[(blk0 + kp) % nsets];
    if (lk < 0) continue;
    if (lk < b->step - kRing) lk = b->step - kRing + 1;   // its event was reused: the oldest one still valid is later
    cudaStreamWaitEvent(sbg, b->ev_k2[lk % kRing], 0);
  }
  pgx::MacArgs m{};
  fill_mac_common(b, m, false, head0);
  const pgx::MacPlan& pl = b->plan_tile;
  m.mix = 0;
  m.yspec = reinterpret_cast<float4*>(b->ytile);
  m.Pt = b->P - 1;   // the committed rows: all but the open slot head0 and the n_spare slots after it (mod R)
  m.jfix = -1;
  if (head0 + b->n_spare < b->R) { m.off = 0; m.skip = head0; m.nskip = 1 + b->n_spare; }
  else                           { m.off = head0 + b->n_spare + 1 - b->R; m.skip = b->R; m.nskip = 0; }
  m.n_terms = m.Pt;
  m.n_split = pl.n_split; m.terms_per_split = pl.terms_per_split; m.n_otiles = pl.n_otiles; m.st = pl.st;
  m.variant = pl.variant; m.persistent_ctas = pl.persistent_ctas;
  m.tile = T; m.tile_u = pl.tile_u; m.tile_stages = pl.tile_stages; m.P = b->P; m.head = head0;
  m.tile_set0 = (int)(blk0 % nsets); m.tile_nsets = nsets;
  m.tile_stride = (int64_t)(b->tile_set_elems / 2);   // in float4
  {
    ProfScope ps(b, sbg, 1);
    pgx::launch_fdl_mac(m, sbg);
    b->launches += 1;
  }
  pgx_bank::TilePass& tp = b->tpass[b->n_tpass % pgx_bank::kTilePasses];
  // the entry this one replaces leaves the ring K1's WAR scan looks at: if the ingest stream has not waited for that
  // pass yet (several passes per block: coverage dropped again and again inside ragged pulls) it does so now, before
  // the event is reused -- in the regular cadence that pass was waited for long ago and nothing is enqueued
  const int64_t evicted = b->n_tpass - pgx_bank::kTilePasses;
  if (evicted >= 0 && tp.base >= 0 && evicted > b->sin_waited_tpass) {
    cudaStreamWaitEvent(b->s_in, tp.ev, 0);
    if (!b->serial) b->sin_waited_tpass = evicted;
  }
  cudaEventRecord(tp.ev, sbg);
  tp.base = blk0;
  b->n_tpass += 1;
  b->tile_base = blk0;
}

// the tiled pass that covers block blk (valid coverage is the caller's business): its issue count, or -1
int64_t tile_pass_of(pgx_bank* b, int64_t blk) {
  for (int64_t k = b->n_tpass - 1; k >= 0 && k >= b->n_tpass - pgx_bank::kTilePasses; --k) {
    const pgx_bank::TilePass& tp = b->tpass[k % pgx_bank::kTilePasses];
    if (tp.base >= 0 && blk >= tp.base && blk < tp.base + b->tile) return k;
  }
  return -1;
}

// One block step.  Buffers, accessors and the ordering of every cross-stream hazard:
//   K1(i)   [ingest]     R x, hist[prev];  W hist[cur], fdl[slot(t)]
//   PAST(t) [background] R fdl[all but slot(t), slot(t+1)], Hd, fmap;  W ypart, ypast[t&1]
//   NOW(i)  [critical]   R fdl[slot(t)], Hd, fmap;  W ynow                      (mix mode only)
//   K2(i)   [critical]   R ypast[t&1], ynow, fdl[slot(t)], Hd, fmap;  W y
//   RAW  PAST(t+1) <- K1 of the step completing block t ............ wait ev_k1       (issue_past `after`)
//   RAW  K2(i), NOW(i) <- K1(i) ................................... critical waits ev_k1[i]
//   RAW  K2(i) <- PAST(t) ......................................... critical waits ev_mac[t]
//   WAR  K1(i) rewrites slot(t) / hist[cur] of an open block that K2(i-1), NOW(i-1) read (also every step
//        when R = 1: a single row) ................................ ingest waits ev_k2[i-1]
//   WAR  K1 of a new block t overwrites slot(t) = block t-R: read by PAST(t-2) (oldest term) and, as
//        present term, by K2/NOW of block t-R ..................... ingest waits ev_mac[t-2] and ev_k2[i-2]
//   WAR  PAST(t+1) overwrites ypast[(t+1)&1] read by K2 of block t-1 . background waits that K2's event
//   same-stream order covers hist halves (K1 only), ypart (background only), ynow (critical only).
// Time-tiled banks (conv pulls) replace PAST(t) by TILE(t0), one per `tile` blocks, on one background stream:
//   TILE(t0) [background] R fdl[all but slot(t0) and the n_spare slots after it], Hd, fmap;  W ytile[set(t0 .. t0+tile-1)]
//   K2(i)    [critical]   R ytile[set(t)], fdl[slot(t), slot(t-1) .. slot(t-(t-t0))], Hd, fmap;  W y
//   RAW  TILE(t0) <- K1 of the step completing block t0-1 ........ wait ev_k1                (issue_tile `after`)
//   RAW  K2(i) <- TILE(t0) covering block t ...................... critical waits the pass's event (once per pass and stream)
//   WAR  K1 of a new block t overwrites slot(t) = block t-R, which TILE(t0) reads while t0-P+1 <= t-R, i.e. for
//        every pass with t0 <= t-n_spare-1 ....................... ingest waits that pass's event (once per pass)
//   WAR  K1 of a new block t vs the output stages that read slot(t) R blocks ago: a step tile+2 back stands in for
//        them (a nearer one would wait for the pass that is still running and stall the next pass behind this K1)
//   WAR  TILE(t0) overwrites the sets of blocks t0-2*tile .. t0-tile-1 ... background waits the last K2 of each set
//   Coverage (tile_base) is dropped by reset / filter map / filter reload / a mix pull; the next conv step issues a
//   pass on demand for its own block, like the per-block schedule does.
int run_step(pgx_bank* b, const float* x_dev, const pgx_layout& xl, float* y_dev, const pgx_layout& yl, int pos,
             int take, bool mix, cudaStream_t crit) {
  struct StreamSwap {  // PGX_DEBUG_SERIAL=1: run all three roles on the critical stream
    pgx_bank* b; cudaStream_t in, bg;
    cudaStream_t bg2;
    StreamSwap(pgx_bank* b_, cudaStream_t crit_) : b(b_), in(b_->s_in), bg(b_->s_bg), bg2(b_->s_bg2) {
      if (b->serial) b->s_in = b->s_bg = b->s_bg2 = crit_;
    }
    ~StreamSwap() { b->s_in = in; b->s_bg = bg; b->s_bg2 = bg2; }
  } swap_guard(b, crit);
  const pgx_bank_config& c = b->cfg;
  const int B = b->B, P = b->P, R = b->R;
  const int64_t i = b->step, t = b->block;
  const bool completes = (b->fill + take == B);
  if (b->after_graph) {
    // the steps before this one were graph replays on the bank's stream: whatever the schedule below waits for
    // (ev_k1 / ev_k2 / ev_mac of earlier steps and blocks) is complete once that stream reaches this point
    for (int e = 0; e < kRing; ++e) {
      cudaEventRecord(b->ev_k1[e], b->stream);
      cudaEventRecord(b->ev_k2[e], b->stream);
      cudaEventRecord(b->ev_mac[e], b->stream);
    }
    b->after_graph = false;
  }

  // one kernel per step: single-partition banks, and banks with a few partitions (the past sum is added in-kernel)
  const bool fused1 = (!mix && b->use_conv1 && (R == 1 || P <= b->fused_max_p));
  pgx::R2CArgs r{};
  r.x = x_dev; r.xs = xl.stream; r.xc = xl.chan; r.xi = xl.samp; r.x_off = pos;
  r.hist = b->hist; r.fdl = b->fdl; r.tw = b->tw;
  r.n_fft = c.n_streams * b->c_x; r.c_in = c.c_in; r.c_x = b->c_x; r.B = B; r.R = R;
  r.slot = b->head; r.half = b->half; r.fill = b->fill; r.take = take;
  r.mixdown = (c.flags & PGX_FLAG_MIXDOWN_INPUT) ? 1 : 0;
  const bool whole = (b->fill == 0 && take == B);
  auto vec_ok = [&](const void* p, const pgx_layout& l) {  // float2 access at p[s*stream + c*chan + pos + even]
    return l.samp == 1 && (l.stream % 2) == 0 && (l.chan % 2) == 0 && (pos % 2) == 0 &&
           (reinterpret_cast<uintptr_t>(p) % 8) == 0;
  };
  r.fast = (whole && !r.mixdown && vec_ok(x_dev, xl)) ? 1 : 0;

  // single-partition mixes of mono transforms: K1 and the present-slot accumulate are one kernel, on the critical stream
  const bool mix1 = (mix && R == 1 && b->use_mix1 && b->mix1_rows > 0);
  // ---- ingest stream: K1 (single-partition conv pulls run K1 and K2 as one kernel on the critical stream)
  if (!fused1 && !mix1) {
    if (b->fill > 0 || R == 1 || b->prev_on_crit) {
      // same ring row (and open half) as the previous step -- or the previous step was a fused kernel on the
      // critical stream (k_conv1 / k_mix1), which WROTE hist and its ring row there: this K1 and the past pass
      // ordered after it must not start before that kernel is done (conv -> mix switch with pulls queued)
      if (i >= 1) cudaStreamWaitEvent(b->s_in, b->ev_k2[(i - 1) % kRing], 0);
    }
    if (!(b->fill > 0 || R == 1)) {
      // (a step far enough back stands in for the output stages that read this ring row R blocks ago.  On a tiled bank
      // the output stages of the last `tile` blocks wait for a pass that may still be running, and the next pass cannot
      // start before this K1: look further back than they reach -- R >= 35 there, so tile + 2 steps back is still
      // younger than every reader of the row)
      const int back = b->tile > 1 ? b->tile + 2 : 2;
      if (i >= back) cudaStreamWaitEvent(b->s_in, b->ev_k2[(i - back) % kRing], 0);
      // (per-block passes: every block had one issued.  A tiled bank only issues them for mix pulls)
      if (t >= 2 && (b->tile == 1 || b->last_past_issue_block >= t - 2))
        cudaStreamWaitEvent(b->s_in, b->ev_mac[(t - 2) % kRing], 0);
      // a tiled pass with first block blk0 reads the rows of blocks blk0-P+1 .. blk0-1; this K1 overwrites the row of
      // block t-R = t-P-n_spare: every tiled pass with blk0 <= t-n_spare-1 must be done (they complete in issue order)
      for (int64_t kk = b->n_tpass - 1; kk >= 0 && kk >= b->n_tpass - pgx_bank::kTilePasses; --kk) {
        const pgx_bank::TilePass& tp = b->tpass[kk % pgx_bank::kTilePasses];
        if (tp.base >= 0 && tp.base <= t - b->n_spare - 1) {
          if (kk > b->sin_waited_tpass || b->serial) {   // (else: the ingest stream is already behind that pass)
            cudaStreamWaitEvent(b->s_in, tp.ev, 0);
            if (!b->serial) b->sin_waited_tpass = kk;
          }
          break;
        }
      }
    }
    {
      ProfScope ps(b, b->s_in, 0);
      pgx::launch_r2c_ingest(r, b->s_in);
    }
    cudaEventRecord(b->ev_k1[i % kRing], b->s_in);
    b->launches += 1;
  } else {
    if (i >= 1) cudaStreamWaitEvent(crit, b->ev_k2[(i - 1) % kRing], 0);  // previous step (no-op on one stream)
    if (R > 1 && b->fill == 0 && t >= 2) cudaStreamWaitEvent(crit, b->ev_mac[(t - 2) % kRing], 0);  // a mix pass
  }

  // ---- background stream: past sum of the open block, if it is not in flight / valid already
  int n_split_past = 0;
  const int par = (int)(t & 1);
  const bool tiled = (b->tile > 1 && !mix && P > 1 && !fused1);
  int n_recent = 0;
  const float2* tile_set = nullptr;
  if (tiled) {
    if (!(b->tile_base >= 0 && t >= b->tile_base && t < b->tile_base + b->tile))
      issue_tile(b, t, b->head, b->ev_k1[i % kRing]);    // not covered (first pull, after a reset / map change): now
    const int64_t kk = tile_pass_of(b, t);
    if (kk >= 0 && !(crit == b->crit_waited_stream && kk == b->crit_waited_tpass)) {
      cudaStreamWaitEvent(crit, b->tpass[kk % pgx_bank::kTilePasses].ev, 0);
      b->crit_waited_stream = crit;
      b->crit_waited_tpass = kk;
    }
    n_recent = (int)(t - b->tile_base);
    n_split_past = b->plan_tile.n_split;
    const int set = (int)(t % (2 * b->tile));
    tile_set = b->ytile + (size_t)set * b->tile_set_elems;
    b->last_k2_of_set[set] = i;
  } else if (P > 1 && !fused1) {
    if (b->tile > 1) b->tile_base = -1;   // a mix pull on a tiled bank: conv coverage is recomputed when conv pulls return
    if (b->past_block != t || b->past_mode != (mix ? 1 : 0)) issue_past(b, mix, t, b->head, b->ev_k1[i % kRing]);
    cudaStreamWaitEvent(crit, b->ev_mac[t % kRing], 0);
    const int ns = (mix ? b->plan_mix : b->plan_conv).n_partials;
    n_split_past = ns > kFoldAbove ? 1 : ns;
  }

  // ---- critical stream: [NOW] + K2
  if (!fused1 && !mix1) cudaStreamWaitEvent(crit, b->ev_k1[i % kRing], 0);
  pgx::C2RArgs k{};
  k.yspec = tiled ? tile_set : b->ypast[par]; k.n_split = n_split_past;
  k.n_recent = n_recent;
  k.n_out = mix ? c.c_out : c.n_streams * c.c_out;
  bool step_done = false;  // the whole step ran inside k_mix1<LAST>: no K2 launch
  if (mix1) {
    k.Hd = b->Hd; k.fmap = b->fmap; k.c_f = c.filter_channels; k.R = R; k.c_out = c.c_out; k.head = b->head;
    k.y = y_dev; k.ys = 0; k.yc = yl.chan; k.yi = yl.samp; k.y_off = pos;
    k.B = B; k.fill = b->fill; k.take = take; k.tw = b->tw; k.wet = b->wet; k.dry = 0.0f; k.xdry = nullptr;
    k.add = nullptr; k.fast = 0;
    {
      pgx_layout ye = yl;
      ye.stream = 0;
      k.fast = (whole && vec_ok(y_dev, ye)) ? 1 : 0;
    }
    // few CTAs: the last one to finish folds their rows and does the inverse transforms itself
    const bool last = (b->mix1_rows <= 64 && c.c_out <= pgx::mix1_sources_per_cta(B, c.n_streams));
    {
      ProfScope ps(b, crit, last ? 5 : 4);
      pgx::launch_mix1(r, k, b->ynow, last ? b->mix1_ticket : nullptr, crit);
    }
    b->launches += 1;
    cudaEventRecord(b->ev_k1[i % kRing], crit);
    step_done = last;
    k.ynow = b->ynow; k.n_split_now = b->mix1_rows;
    if (b->mix1_rows > 64) {  // many CTAs: fold their rows with the wide kernel instead of inside K2
      float2* folded = b->ynow + (size_t)b->mix1_rows * c.c_out * B;
      ProfScope ps(b, crit, 3);
      pgx::launch_reduce_partials(reinterpret_cast<const float4*>(b->ynow), reinterpret_cast<float4*>(folded),
                                  b->mix1_rows, c.c_out, B / 2, crit);
      b->launches += 1;
      k.ynow = folded; k.n_split_now = 1;
    }
    k.fdl = nullptr;
  } else if (mix) {  // present term of every stream: K3 restricted to the open slot
    pgx::MacArgs m{};
    fill_mac_common(b, m, true, b->head);
    m.mix = b->plan_now.layout;  // always the flattened layout: one term per stream
    m.yspec = reinterpret_cast<float4*>(b->ynow);
    m.Pt = 1; m.off = 0; m.skip = R; m.nskip = 0; m.jfix = b->head;
    m.n_terms = c.n_streams;
    m.n_split = b->plan_now.n_split; m.terms_per_split = b->plan_now.terms_per_split;
    m.n_otiles = b->plan_now.n_otiles; m.st = b->plan_now.st;
    m.variant = b->plan_now.variant; m.persistent_ctas = b->plan_now.persistent_ctas;
    {
      ProfScope ps(b, crit, 4);
      pgx::launch_fdl_mac(m, crit);
    }
    b->launches += 1;
    k.ynow = b->ynow; k.n_split_now = m.n_split;
    k.fdl = nullptr;
  } else {
    k.ynow = nullptr; k.n_split_now = 0;
    k.fdl = b->fdl;
  }
  k.Hd = b->Hd; k.fmap = b->fmap; k.c_x = b->c_x; k.c_f = c.filter_channels; k.R = R; k.head = b->head;
  k.y = y_dev; k.ys = mix ? 0 : yl.stream; k.yc = yl.chan; k.yi = yl.samp; k.y_off = pos;
  k.c_out = c.c_out; k.B = B; k.fill = b->fill; k.take = take; k.tw = b->tw;
  k.wet = b->wet; k.dry = b->dry;
  k.fft16 = b->fft16;
  k.xdry = (!mix && b->dry != 0.0f) ? x_dev : nullptr;
  k.xs = xl.stream; k.xc = xl.chan; k.xi = xl.samp; k.x_off = pos;
  k.add = b->addend; k.as = mix ? 0 : b->addend_l.stream; k.ac = b->addend_l.chan; k.ai = b->addend_l.samp;
  {
    pgx_layout ye = yl;
    if (mix) ye.stream = 0;
    k.fast = (whole && vec_ok(y_dev, ye) && (!k.xdry || vec_ok(x_dev, xl)) &&
              (!k.add || vec_ok(b->addend, b->addend_l))) ? 1 : 0;
  }

  // the block commits with this step: its row is final once K1 has run, so the next block's past pass
  // can start now, overlapping this step's K2 and the next step's K1
  if (completes && tiled) {
    if (!(t + 1 >= b->tile_base && t + 1 < b->tile_base + b->tile))
      issue_tile(b, t + 1, (b->head + 1) % R, b->ev_k1[i % kRing]);
  } else if (completes && P > 1 && !fused1) issue_past(b, mix, t + 1, (b->head + 1) % R, b->ev_k1[i % kRing]);
  if (fused1 && P > 1) {  // past partitions summed inside the fused kernel: every ring row but the open and the spare
    k.n_past = R - 2;
    k.q0 = R - 1 - b->head;
    if (b->head + 1 < R) { k.p_off = 0; k.p_skip = b->head; k.p_nskip = 2; }
    else                 { k.p_off = 1; k.p_skip = R; k.p_nskip = 0; }
  }

  if (!step_done) {
    ProfScope ps(b, crit, fused1 ? 5 : 2);
    if (fused1) pgx::launch_conv1(r, k, crit);
    else pgx::launch_c2r_emit(k, crit);
    b->launches += 1;
  }
  if (fused1) cudaEventRecord(b->ev_k1[i % kRing], crit);
  cudaEventRecord(b->ev_k2[i % kRing], crit);
  b->last_k2_of_par[par] = i;
  b->prev_on_crit = fused1 || mix1;
  b->steps += 1;
  if (b->profiling) b->prof_steps += 1;
  b->step += 1;
  b->fill += take;
  if (completes) {  // block complete: commit the row, advance the ring
    b->head = (b->head + 1) % R;
    b->half ^= 1;
    b->fill = 0;
    b->block += 1;
  }
  return PGX_OK;
}

// One pull of one level on device buffers, enqueued on crit (no synchronisation).
int run_pull1(pgx_bank* b, const float* x_dev, const pgx_layout& xl, float* y_dev, const pgx_layout& yl, int n,
             bool mix, bool input_resident, cudaStream_t crit) {
  if (!input_resident) {
    // x is produced by work queued earlier on the caller's stream: the ingest stream must see it
    cudaEventRecord(b->ev_call, crit);
    cudaStreamWaitEvent(b->s_in, b->ev_call, 0);
  }
  b->last_crit = crit;
  if (b->fmap_pending) {  // a re-selected filter map is still being copied in: its readers wait for it
    cudaStreamWaitEvent(crit, b->fmap_ev[b->fmap_cur], 0);
    cudaStreamWaitEvent(b->s_bg, b->fmap_ev[b->fmap_cur], 0);
    cudaStreamWaitEvent(b->s_bg2, b->fmap_ev[b->fmap_cur], 0);
    b->fmap_pending = false;
  }
  int pos = 0;
  while (pos < n) {
    const int take = (b->B - b->fill < n - pos) ? (b->B - b->fill) : (n - pos);
    const int rc = run_step(b, x_dev, xl, y_dev, yl, pos, take, mix, crit);
    if (rc != PGX_OK) return rc;
    pos += take;
  }
  if (crit != b->stream) cudaEventRecord(b->fmap_ret_crit[b->fmap_cur], crit);
  PGX_CUDA(cudaGetLastError());
  return PGX_OK;
}

// One pull on device buffers, enqueued on crit (no synchronisation).  Two-level banks cut the pull at big-block
// boundaries: each piece goes through the head level with the tail level's contribution as addend and is
// gathered into the big block; a completed big block is pushed through the tail level at once.
int run_pull(pgx_bank* b, const float* x_dev, const pgx_layout& xl, float* y_dev, const pgx_layout& yl, int n,
             bool mix, bool input_resident, cudaStream_t crit) {
  if (mix && b->dry != 0.0f)
    return fail(PGX_ERR_INVALID, "a fused-mix pull has no dry path (dry = %g): set dry = 0 or pull per stream", (double)b->dry);
  if (!b->tail) return run_pull1(b, x_dev, xl, y_dev, yl, n, mix, input_resident, crit);
  if (b->tail_mode < 0) b->tail_mode = mix ? 1 : 0;
  if (b->tail_mode != (mix ? 1 : 0))
    return fail(PGX_ERR_INVALID, "a two-level bank cannot switch between per-stream and mixed pulls without a reset");
  const pgx_bank_config& c = b->cfg;
  const int TB = b->tail_B;
  const pgx_layout xa{(int64_t)c.c_in * TB, TB, 1};
  const pgx_layout ya{mix ? 0 : (int64_t)c.c_out * TB, TB, 1};
  int pos = 0;
  while (pos < n) {
    const int take = (TB - b->tail_fill < n - pos) ? (TB - b->tail_fill) : (n - pos);
    if (!input_resident && pos == 0) {  // x may come from work queued on crit: the head's ingest is ordered by run_pull1
    }
    pgx::launch_copy_block(x_dev + (int64_t)pos * xl.samp, xl.stream, xl.chan, xl.samp, b->xacc + b->tail_fill,
                           xa.stream, xa.chan, 1, c.n_streams, c.c_in, take, crit);
    b->launches += 1;
    b->addend = b->ytail[b->tail_cur] + b->tail_fill;
    b->addend_l = ya;
    int rc = run_pull1(b, x_dev + (int64_t)pos * xl.samp, xl, y_dev + (int64_t)pos * yl.samp, yl, take, mix,
                       input_resident, crit);
    b->addend = nullptr;
    if (rc != PGX_OK) return rc;
    b->tail_fill += take;
    if (b->tail_fill == TB) {  // the big block is complete: its tail contribution lands TB samples later
      rc = run_pull1(b->tail, b->xacc, xa, b->ytail[b->tail_cur ^ 1], ya, TB, mix, false, crit);
      if (rc != PGX_OK) return rc;
      b->tail_cur ^= 1;
      b->tail_fill = 0;
    }
    pos += take;
  }
  return PGX_OK;
}

// State changes outside run_pull (reset, filter reload / re-selection) are rare and synchronous: drain the
// bank's streams (a caller-owned critical stream is the caller's to drain) and drop the cached past sum.
void quiesce(pgx_bank* b) {
  if (b->tail) quiesce(b->tail);
  cudaStreamSynchronize(b->s_h2d);
  cudaStreamSynchronize(b->s_in);
  cudaStreamSynchronize(b->s_bg);
  cudaStreamSynchronize(b->s_bg2);
  cudaStreamSynchronize(b->stream);
  cudaStreamSynchronize(b->s_d2h);
  b->past_block = -1;
  b->past_mode = -1;
  b->tile_base = -1;
}

int check_pull_args(pgx_bank* b, const void* x, const void* y, int n) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (!x || !y) return fail(PGX_ERR_INVALID, "x / y must not be NULL");
  if (n < 1 || n > b->cfg.max_pull)
    return fail(PGX_ERR_INVALID, "pull of %d samples outside [1, max_pull=%d]", n, b->cfg.max_pull);
  return PGX_OK;
}

int prep_filters(pgx_bank* b, const float* h_host, int first_row, int n_rows) {
  const pgx_bank_config& c = b->cfg;
  float* h_dev = nullptr;
  const size_t hb = (size_t)n_rows * c.filter_len * sizeof(float);
  PGX_CUDA(cudaMalloc(&h_dev, hb));
  cudaError_t e = cudaMemcpyAsync(h_dev, h_host, hb, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) {
    pgx::FilterPrepArgs fp{};
    fp.h = h_dev;
    fp.Hd = b->Hd + (size_t)first_row * 2 * b->R * b->B;
    fp.tw = b->tw; fp.n_rows = n_rows; fp.L = c.filter_len; fp.B = b->B; fp.P = b->P; fp.R = b->R;
    pgx::launch_filter_prep(fp, b->stream);
    b->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(h_dev);
  if (e != cudaSuccess) return fail(PGX_ERR_CUDA, "filter preparation: %s", cudaGetErrorString(e));
  return PGX_OK;
}

}  // namespace

// Table entries in the outer loop, a chunk of queries in the inner one: the inner loop is a branch-free compare /
// select over contiguous doubles that the host compiler vectorises (AVX2 clone picked at load time where the CPU
// has it).  "d < best" strictly, entries in table order: the FIRST minimum wins, as in spatial_pe.py:413-416.
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
static void nearest_scan(const double* tab_elev, const double* tab_az, int32_t n_tab, const double* azimuth,
                         const double* elevation, int32_t n, int32_t* out_index) {
  constexpr int kChunk = 128;
  double az[kChunk], best[kChunk];
  int32_t bi[kChunk];
  for (int32_t q0 = 0; q0 < n; q0 += kChunk) {
    const int m = (n - q0 < kChunk) ? (n - q0) : kChunk;
    for (int q = 0; q < m; ++q) {
      az[q] = fmin(180.0, fabs(azimuth[q0 + q]));
      best[q] = INFINITY;
      bi[q] = 0;
    }
    const double* el = elevation + q0;
    for (int32_t i = 0; i < n_tab; ++i) {
      const double te = tab_elev[i], ta = tab_az[i];
      for (int q = 0; q < m; ++q) {
        const double de = te - el[q], da = ta - az[q];
        const double d = de * de + da * da;
        const bool lt = d < best[q];
        best[q] = lt ? d : best[q];
        bi[q] = lt ? i : bi[q];
      }
    }
    for (int q = 0; q < m; ++q) out_index[q0 + q] = bi[q];
  }
}

extern "C" {

int pgx_abi_version(void) { return PGX_ABI_VERSION; }

int pgx_struct_size(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(pgx_layout);
    case 1: return (int)sizeof(pgx_bank_config);
    case 2: return (int)sizeof(pgx_bank_info);
    case 3: return (int)sizeof(pgx_profile);
    case 4: return (int)sizeof(pgx_osc_config);
    default: return -1;
  }
}

const char* pgx_last_error(void) { return g_err.c_str(); }

int pgx_device_count(int* count) {
  if (!count) return fail(PGX_ERR_INVALID, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(PGX_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return PGX_OK;
}

int pgx_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail(PGX_ERR_INVALID, "pgx_host_alloc: bad arguments");
  PGX_CUDA(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
  return PGX_OK;
}

int pgx_host_alloc_flags(void** ptr, int64_t bytes, int32_t flags) {
  if (!ptr || bytes <= 0) return fail(PGX_ERR_INVALID, "pgx_host_alloc_flags: bad arguments");
  PGX_CUDA(cudaHostAlloc(ptr, (size_t)bytes, (flags & PGX_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return PGX_OK;
}

int pgx_host_free(void* ptr) {
  if (ptr) PGX_CUDA(cudaFreeHost(ptr));
  return PGX_OK;
}

int pgx_device_alloc(int32_t device, int64_t bytes, void** ptr) {
  if (!ptr || bytes <= 0) return fail(PGX_ERR_INVALID, "pgx_device_alloc: bad arguments");
  PGX_CUDA(cudaSetDevice(device));
  PGX_CUDA(cudaMalloc(ptr, (size_t)bytes));
  return PGX_OK;
}

int pgx_device_free(int32_t device, void* ptr) {
  if (!ptr) return PGX_OK;
  PGX_CUDA(cudaSetDevice(device));
  PGX_CUDA(cudaFree(ptr));
  return PGX_OK;
}

int pgx_device_upload(int32_t device, void* dst_dev, const void* src_host, int64_t bytes) {
  if (!dst_dev || (!src_host && bytes > 0) || bytes < 0) return fail(PGX_ERR_INVALID, "pgx_device_upload: bad arguments");
  PGX_CUDA(cudaSetDevice(device));
  if (src_host) PGX_CUDA(cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice));
  // a pageable-memory copy may return before its DMA has landed, and the banks read from non-blocking streams
  // that are not ordered with the legacy stream: make the data visible to every stream before returning
  PGX_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  return PGX_OK;
}

int pgx_device_zero(int32_t device, void* dst_dev, int64_t bytes) {
  if (!dst_dev || bytes < 0) return fail(PGX_ERR_INVALID, "pgx_device_zero: bad arguments");
  PGX_CUDA(cudaSetDevice(device));
  PGX_CUDA(cudaMemset(dst_dev, 0, (size_t)bytes));
  PGX_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // cudaMemset is asynchronous with respect to the host
  return PGX_OK;
}

static int create_single(pgx_bank** out, const pgx_bank_config* cfg, const float* h, const int32_t* filter_of_stream,
                         bool allow_tile = true) {
  if (!out || !cfg || !h) return fail(PGX_ERR_INVALID, "pgx_bank_create: NULL argument");
  *out = nullptr;
  const pgx_bank_config& c = *cfg;
  if (c.n_streams < 1) return fail(PGX_ERR_INVALID, "n_streams must be >= 1, got %d", c.n_streams);
  if (c.c_in < 1 || c.c_out < 1) return fail(PGX_ERR_INVALID, "channel counts must be >= 1");
  if (c.filter_len < 1) return fail(PGX_ERR_INVALID, "ConvolvePE filter must be non-empty");
  if (c.n_filters < 1) return fail(PGX_ERR_INVALID, "n_filters must be >= 1");
  if (!is_pow2(c.block) || c.block < 16 || c.block > 8192)
    return fail(PGX_ERR_INVALID, "block must be a power of two in [16, 8192], got %d", c.block);
  if (c.max_pull < 1) return fail(PGX_ERR_INVALID, "max_pull must be >= 1");
  const bool mixdown = (c.flags & PGX_FLAG_MIXDOWN_INPUT) != 0;
  const int c_x = mixdown ? 1 : c.c_in;
  // channel rules of convolve_pe.py:207-223: mono filter -> every source channel; multi-channel filter
  // -> fan-out of a mono source, or one filter channel per source channel; anything else is rejected.
  if (c.filter_channels == 1) {
    if (c.c_out != c_x)
      return fail(PGX_ERR_INVALID, "mono filter: c_out (%d) must equal source channels (%d)", c.c_out, c_x);
  } else {
    if (c.filter_channels != c.c_out)
      return fail(PGX_ERR_INVALID, "filter_channels (%d) must be 1 or c_out (%d)", c.filter_channels, c.c_out);
    if (c_x != 1 && c_x != c.filter_channels)
      return fail(PGX_ERR_INVALID,
                  "ConvolvePE filter channels (%d) must match src channels (%d), or be mono, or be "
                  "multi-channel with a mono source.",
                  c.filter_channels, c_x);
  }
  if (filter_of_stream)
    for (int s = 0; s < c.n_streams; ++s)
      if (filter_of_stream[s] < 0 || filter_of_stream[s] >= c.n_filters)
        return fail(PGX_ERR_INVALID, "filter_of_stream[%d]=%d outside [0,%d)", s, filter_of_stream[s], c.n_filters);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(PGX_ERR_NO_DEVICE, "no CUDA device: libpgx has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) return fail(PGX_ERR_INVALID, "device %d outside [0,%d)", c.device, ndev);
  PGX_CUDA(cudaSetDevice(c.device));

  pgx_bank* b = new (std::nothrow) pgx_bank();
  if (!b) return fail(PGX_ERR_NOMEM, "out of host memory");
  b->cfg = c;
  b->c_x = c_x;
  b->B = c.block;
  b->P = (c.filter_len + c.block - 1) / c.block;
  const size_t n_fft = (size_t)c.n_streams * c_x;
  {
    // time tiling: for banks that stream a long delay line (the HBM-bound class); PGX_TILE = 1 (off) | 2 | 4,
    // PGX_TILE_MIN = smallest n_streams * c_x * B that qualifies
    int want = 4;
    long tile_min = 1L << 17;
    if (const char* e = getenv("PGX_TILE")) want = atoi(e);
    if (const char* e = getenv("PGX_TILE_MIN")) tile_min = atol(e);
    if (allow_tile && (want == 2 || want == 4) && b->P >= 32 && b->B >= 256 && c.tail_block == 0 &&
        (long)n_fft * b->B >= tile_min)
      b->tile = want;
  }
  // ring rows: the P partitions of the open block's sum plus spare rows.  One spare lets the next block's ingest run
  // while the open block's pass reads the other rows; a pass that covers `tile` blocks is still reading when the
  // ingest is tile-1 blocks further, so a tiled bank keeps tile-1 spares.  Spare rows meet all-zero filter rows
  // (partitions >= P of the filter table are never written) wherever a kernel walks them.
  b->n_spare = b->P > 1 ? (b->tile > 2 ? b->tile - 1 : 1) : 0;
  b->R = b->P + b->n_spare;
  const int B = b->B, P = b->P, R = b->R;
  const size_t h_rows = (size_t)c.n_filters * c.filter_channels;
  const size_t n_out_max = (size_t)c.n_streams * c.c_out;

  b->hist_bytes = n_fft * 2 * B * sizeof(float);
  b->fdl_bytes = n_fft * R * B * sizeof(float2);
  b->Hd_bytes = h_rows * 2 * R * B * sizeof(float2);
  {
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess && prop.multiProcessorCount > 0)
      b->sm_count = prop.multiProcessorCount;
  }
  size_t y_conv = 0, y_mix = 0;
  if (P > 1) {  // background pass over the P-1 committed partitions
    b->plan_conv = pgx::mac_plan(c.n_streams, c.c_out, B / 2, R - 2, false, c.n_filters == 1, b->sm_count);
    b->plan_mix = pgx::mac_plan(c.n_streams, c.c_out, B / 2, c.n_streams * (R - 2), true, c.n_filters == 1, b->sm_count);
    y_conv = (size_t)b->plan_conv.n_partials * c.n_streams * c.c_out;
    y_mix = (size_t)b->plan_mix.n_partials * c.c_out;
  }
  b->plan_now = pgx::mac_plan(c.n_streams, c.c_out, B / 2, c.n_streams, true, c.n_filters == 1, b->sm_count);
  if (b->tile > 1) {
    b->plan_tile = pgx::mac_plan_tiled(c.n_streams, c.c_out, B / 2, P - 1, c.n_filters == 1, b->sm_count, b->tile);
    b->tile_set_elems = (size_t)b->plan_tile.n_split * n_out_max * B;
    b->ytile_bytes = (size_t)2 * b->tile * b->tile_set_elems * sizeof(float2);
  }
  b->ypart_bytes = (y_conv > y_mix ? y_conv : y_mix) * B * sizeof(float2);
  if (b->ypart_bytes == 0) b->ypart_bytes = sizeof(float2);
  b->ysum_bytes = n_out_max * B * sizeof(float2) * kFoldAbove;
  {
    const int g = pgx::mix1_sources_per_cta(B, c.n_streams);
    b->mix1_rows = (g > 0 && P == 1 && c_x == 1) ? (c.n_streams + g - 1) / g : 0;
    const int rows = b->plan_now.n_split > b->mix1_rows ? b->plan_now.n_split : b->mix1_rows;
    b->ynow_bytes = (size_t)(rows + 1) * c.c_out * B * sizeof(float2);  // + one folded row block
  }
  b->xs_bytes = (size_t)c.n_streams * c.c_in * c.max_pull * sizeof(float);
  b->ys_bytes = (size_t)c.n_streams * c.c_out * c.max_pull * sizeof(float);

  int rc = PGX_OK;
  auto guard = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == PGX_OK)
      rc = fail(e == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  };
  {
    int lo = 0, hi = 0;  // the latency-critical kernels (K1, K2) outrank the background pass for SM slots
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    guard(cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, hi), "cudaStreamCreate");
    guard(cudaStreamCreateWithPriority(&b->s_in, cudaStreamNonBlocking, hi), "cudaStreamCreate(in)");
    guard(cudaStreamCreateWithPriority(&b->s_bg, cudaStreamNonBlocking, lo), "cudaStreamCreate(bg)");
    guard(cudaStreamCreateWithPriority(&b->s_bg2, cudaStreamNonBlocking, lo), "cudaStreamCreate(bg2)");
    guard(cudaEventCreateWithFlags(&b->ev_bgjoin, cudaEventDisableTiming), "cudaEventCreate");
    if (const char* e = getenv("PGX_BG_STREAMS")) b->two_bg = (e[0] != '1');
    if (const char* e = getenv("PGX_CONV1")) b->use_conv1 = (e[0] != '0');
    if (const char* e = getenv("PGX_FUSED_MAXP")) b->fused_max_p = atoi(e);
    b->fft16 = pgx::conv1_r16_default();
    if (const char* e = getenv("PGX_MIX1")) b->use_mix1 = (e[0] != '0');
    if (const char* e = getenv("PGX_GRAPH")) b->use_graph = (e[0] != '0');
    if (const char* e = getenv("PGX_GRAPH_FUSE")) b->graph_fuse = (e[0] == '1');
    for (cudaEvent_t& e : b->ev_join) guard(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate(join)");
    guard(cudaStreamCreateWithPriority(&b->s_h2d, cudaStreamNonBlocking, hi), "cudaStreamCreate(h2d)");
    guard(cudaStreamCreateWithPriority(&b->s_d2h, cudaStreamNonBlocking, hi), "cudaStreamCreate(d2h)");
    if (const char* e = getenv("PGX_DEBUG_SERIAL")) {  // debugging aid: no overlap, one stream for everything
      if (e[0] == '1') b->serial = true;
    }
  }
  guard(cudaEventCreateWithFlags(&b->ev_call, cudaEventDisableTiming), "cudaEventCreate");
  for (int i = 0; i < kRing; ++i) {
    guard(cudaEventCreateWithFlags(&b->ev_k1[i], cudaEventDisableTiming), "cudaEventCreate");
    guard(cudaEventCreateWithFlags(&b->ev_k2[i], cudaEventDisableTiming), "cudaEventCreate");
    guard(cudaEventCreateWithFlags(&b->ev_mac[i], cudaEventDisableTiming), "cudaEventCreate");
  }
  for (int i = 0; i < pgx_bank::kMapSlots; ++i)
  {
    guard(cudaEventCreateWithFlags(&b->fmap_ev[i], cudaEventDisableTiming), "cudaEventCreate(fmap)");
    guard(cudaEventCreateWithFlags(&b->fmap_ret_crit[i], cudaEventDisableTiming), "cudaEventCreate(fmap)");
    guard(cudaEventCreateWithFlags(&b->fmap_ret_bg[i], cudaEventDisableTiming), "cudaEventCreate(fmap)");
  }
  guard(cudaMalloc(&b->hist, b->hist_bytes), "cudaMalloc(hist)");
  guard(cudaMalloc(&b->fdl, b->fdl_bytes), "cudaMalloc(fdl)");
  guard(cudaMalloc(&b->Hd, b->Hd_bytes), "cudaMalloc(Hd)");
  guard(cudaMalloc(&b->ypast[0], b->ysum_bytes), "cudaMalloc(ypast0)");
  guard(cudaMalloc(&b->ypast[1], b->ysum_bytes), "cudaMalloc(ypast1)");
  guard(cudaMalloc(&b->ypart[0], b->ypart_bytes), "cudaMalloc(ypart)");
  guard(cudaMalloc(&b->ypart[1], b->ypart_bytes), "cudaMalloc(ypart)");
  guard(cudaMalloc(&b->ynow, b->ynow_bytes), "cudaMalloc(ynow)");
  if (b->tile > 1) {
    guard(cudaMalloc(&b->ytile, b->ytile_bytes), "cudaMalloc(ytile)");
    for (auto& tp : b->tpass) guard(cudaEventCreateWithFlags(&tp.ev, cudaEventDisableTiming), "cudaEventCreate(tile)");
  }
  guard(cudaMalloc(&b->mix1_ticket, sizeof(unsigned int)), "cudaMalloc(ticket)");
  guard(cudaMemset(b->mix1_ticket, 0, sizeof(unsigned int)), "memset ticket");
  guard(cudaMalloc(&b->tw, (size_t)2 * B * sizeof(float2)), "cudaMalloc(tw)");
  guard(cudaMalloc(&b->fmap_own, (size_t)pgx_bank::kMapSlots * c.n_streams * sizeof(int32_t)), "cudaMalloc(fmap)");
  b->n_slots = b->tile > 1 ? pgx_bank::kSlots : 3;
  for (int i = 0; i < pgx_bank::kSlots; ++i) {
    if (i < b->n_slots) {
      guard(cudaMalloc(&b->x_stage[i], b->xs_bytes), "cudaMalloc(x_stage)");
      guard(cudaMalloc(&b->y_stage[i], b->ys_bytes), "cudaMalloc(y_stage)");
    }
    guard(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming), "cudaEventCreate");
    guard(cudaEventCreateWithFlags(&b->ev_y[i], cudaEventDisableTiming), "cudaEventCreate");
    guard(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming), "cudaEventCreate");
  }
  guard(cudaHostAlloc(&b->fmap_pinned, (size_t)pgx_bank::kMapSlots * c.n_streams * sizeof(int32_t), cudaHostAllocDefault),
        "cudaHostAlloc");
  if (rc != PGX_OK) {
    free_bank(b);
    return rc;
  }
  b->fmap = b->fmap_own;

  // twiddles in double, stored float: tw[k] = exp(-2*pi*i*k/2B)
  {
    std::vector<float2> tw((size_t)2 * B);
    const double w = -2.0 * M_PI / (2.0 * B);
    for (int k = 0; k < 2 * B; ++k) tw[k] = make_float2((float)cos(w * k), (float)sin(w * k));
    guard(cudaMemcpyAsync(b->tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, b->stream), "H2D tw");
    guard(cudaStreamSynchronize(b->stream), "sync tw");
  }
  for (int s = 0; s < c.n_streams; ++s) b->fmap_pinned[s] = filter_of_stream ? filter_of_stream[s] : (s % c.n_filters);
  guard(cudaMemcpyAsync(b->fmap_own, b->fmap_pinned, (size_t)c.n_streams * sizeof(int32_t), cudaMemcpyHostToDevice, b->stream),
        "H2D fmap");
  guard(cudaMemsetAsync(b->Hd, 0, b->Hd_bytes, b->stream), "memset Hd");
  guard(cudaMemsetAsync(b->hist, 0, b->hist_bytes, b->stream), "memset hist");
  guard(cudaMemsetAsync(b->fdl, 0, b->fdl_bytes, b->stream), "memset fdl");
  // filter spectra (replaces the one-time np.fft.rfft(h, n=nfft), convolve_pe.py:236-239)
  if (rc == PGX_OK) rc = prep_filters(b, h, 0, (int)h_rows);
  guard(cudaStreamSynchronize(b->stream), "sync init");
  if (rc != PGX_OK) {
    free_bank(b);
    return rc;
  }
  b->full_filter_len = c.filter_len;
  *out = b;
  return PGX_OK;
}

int pgx_bank_create(pgx_bank** out, const pgx_bank_config* cfg, const float* h, const int32_t* filter_of_stream) {
  if (!out || !cfg || !h) return fail(PGX_ERR_INVALID, "pgx_bank_create: NULL argument");
  *out = nullptr;
  const int TB = cfg->tail_block;
  if (TB <= 0 || cfg->filter_len <= TB) {  // single level (also when the whole filter fits the head)
    pgx_bank_config c1 = *cfg;
    c1.tail_block = 0;
    return create_single(out, &c1, h, filter_of_stream);
  }
  if (!is_pow2(TB) || TB > 8192 || TB < cfg->block)
    return fail(PGX_ERR_INVALID, "tail_block must be a power of two in [block, 8192], got %d", TB);
  if (cfg->n_filters < 1 || cfg->filter_channels < 1) return fail(PGX_ERR_INVALID, "bad filter set");
  // h = [head (TB taps) | tail]: two filter sets, two uniformly partitioned banks
  const size_t rows = (size_t)cfg->n_filters * cfg->filter_channels;
  const int L = cfg->filter_len, Lt = L - TB;
  std::vector<float> hh(rows * TB), ht(rows * Lt);
  for (size_t r = 0; r < rows; ++r) {
    memcpy(&hh[r * TB], h + r * L, (size_t)TB * sizeof(float));
    memcpy(&ht[r * Lt], h + r * L + TB, (size_t)Lt * sizeof(float));
  }
  pgx_bank_config ch = *cfg, ct = *cfg;
  ch.filter_len = TB; ch.tail_block = 0;
  ct.filter_len = Lt; ct.block = TB; ct.max_pull = TB; ct.tail_block = 0;
  pgx_bank *head = nullptr, *tail = nullptr;
  // (the two levels keep the per-block pass: their schedule interleaves the levels)
  int rc = create_single(&head, &ch, hh.data(), filter_of_stream, false);
  if (rc != PGX_OK) return rc;
  rc = create_single(&tail, &ct, ht.data(), filter_of_stream, false);
  if (rc != PGX_OK) {
    free_bank(head);
    return rc;
  }
  head->tail = tail;
  head->tail_B = TB;
  head->full_filter_len = L;
  head->xacc_bytes = (size_t)cfg->n_streams * cfg->c_in * TB * sizeof(float);
  head->ytail_bytes = (size_t)cfg->n_streams * cfg->c_out * TB * sizeof(float);
  cudaError_t e = cudaMalloc(&head->xacc, head->xacc_bytes);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMalloc(&head->ytail[i], head->ytail_bytes);
    if (e == cudaSuccess) e = cudaMemset(head->ytail[i], 0, head->ytail_bytes);
  }
  if (e == cudaSuccess) e = cudaMemset(head->xacc, 0, head->xacc_bytes);
  if (e != cudaSuccess) {
    free_bank(head);
    return fail(e == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "two-level buffers: %s", cudaGetErrorString(e));
  }
  *out = head;
  return PGX_OK;
}

int pgx_bank_destroy(pgx_bank* bank) {
  free_bank(bank);
  return PGX_OK;
}

int pgx_bank_get_info(pgx_bank* b, pgx_bank_info* info) {
  if (!b || !info) return fail(PGX_ERR_INVALID, "NULL argument");
  const pgx_bank_config& c = b->cfg;
  info->n_streams = c.n_streams; info->c_in = c.c_in; info->c_x = b->c_x; info->c_out = c.c_out;
  info->filter_len = b->full_filter_len; info->filter_channels = c.filter_channels; info->n_filters = c.n_filters;
  info->block = b->B; info->partitions = b->P; info->max_pull = c.max_pull; info->device = c.device;
  info->head = b->head; info->fill = b->fill;
  info->state_bytes = (int64_t)(b->hist_bytes + b->fdl_bytes + b->Hd_bytes + 2 * b->ypart_bytes + 2 * b->ysum_bytes +
                                b->ynow_bytes + b->ytile_bytes + b->n_slots * (b->xs_bytes + b->ys_bytes));
  info->block_steps = b->steps;
  info->tail_block = b->tail ? b->tail_B : 0;
  info->tail_partitions = b->tail ? b->tail->P : 0;
  if (b->tail) {
    const pgx_bank* t = b->tail;
    info->state_bytes += (int64_t)(t->hist_bytes + t->fdl_bytes + t->Hd_bytes + 2 * t->ypart_bytes + 2 * t->ysum_bytes +
                                   t->ynow_bytes + b->xacc_bytes + 2 * b->ytail_bytes);
  }
  info->mac_grid = b->plan_conv.grid; info->mac_split = b->plan_conv.n_split;
  info->mac_stream_tile = b->plan_conv.st; info->mac_occupancy = b->plan_conv.occupancy;
  info->kernel_launches = b->launches + (b->tail ? b->tail->launches : 0);
  info->graph_pulls = b->graph_pulls;
  info->mac_tile = b->tile;
  info->submit_depth = b->n_slots;
  if (b->tile > 1) {
    info->mac_grid = b->plan_tile.grid; info->mac_split = b->plan_tile.n_split;
    info->mac_stream_tile = b->plan_tile.st; info->mac_occupancy = b->plan_tile.occupancy;
  }
  return PGX_OK;
}

int pgx_bank_reset(pgx_bank* b, const int32_t* stream_ids, int32_t k) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  quiesce(b);
  if (b->tail) {
    const bool all = (k <= 0 || !stream_ids);
    if (!all && b->tail_mode == 1)
      return fail(PGX_ERR_INVALID, "per-stream reset of a two-level bank is not available in mixed mode: the tail "
                                   "contribution in flight is already summed over the streams");
    const int rc = pgx_bank_reset(b->tail, stream_ids, k);
    if (rc != PGX_OK) return rc;
    if (all) {
      PGX_CUDA(cudaMemset(b->xacc, 0, b->xacc_bytes));
      PGX_CUDA(cudaMemset(b->ytail[0], 0, b->ytail_bytes));
      PGX_CUDA(cudaMemset(b->ytail[1], 0, b->ytail_bytes));
      b->tail_fill = 0;
      b->tail_cur = 0;
      b->tail_mode = -1;
    } else {
      const size_t xr = (size_t)b->cfg.c_in * b->tail_B * sizeof(float), yr = (size_t)b->cfg.c_out * b->tail_B * sizeof(float);
      for (int i = 0; i < k; ++i) {
        const int s = stream_ids[i];
        if (s < 0 || s >= b->cfg.n_streams) return fail(PGX_ERR_INVALID, "stream id %d outside [0,%d)", s, b->cfg.n_streams);
        PGX_CUDA(cudaMemset(reinterpret_cast<char*>(b->xacc) + s * xr, 0, xr));
        PGX_CUDA(cudaMemset(reinterpret_cast<char*>(b->ytail[0]) + s * yr, 0, yr));
        PGX_CUDA(cudaMemset(reinterpret_cast<char*>(b->ytail[1]) + s * yr, 0, yr));
      }
    }
  }
  if (k <= 0 || !stream_ids) {
    PGX_CUDA(cudaMemsetAsync(b->hist, 0, b->hist_bytes, b->stream));
    PGX_CUDA(cudaMemsetAsync(b->fdl, 0, b->fdl_bytes, b->stream));
    b->head = b->fill = b->half = 0;
    b->step = b->block = 0;
    b->last_k2_of_par[0] = b->last_k2_of_par[1] = -1;
    for (auto& v : b->last_k2_of_set) v = -1;
    for (auto& tp : b->tpass) tp.base = -1;
    b->prev_on_crit = false;
    PGX_CUDA(cudaStreamSynchronize(b->stream));
    return PGX_OK;
  }
  const size_t hs = (size_t)b->c_x * 2 * b->B * sizeof(float);
  const size_t fs = (size_t)b->c_x * b->R * b->B * sizeof(float2);
  for (int i = 0; i < k; ++i) {
    const int s = stream_ids[i];
    if (s < 0 || s >= b->cfg.n_streams) return fail(PGX_ERR_INVALID, "stream id %d outside [0,%d)", s, b->cfg.n_streams);
    PGX_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(b->hist) + s * hs, 0, hs, b->stream));
    PGX_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(b->fdl) + s * fs, 0, fs, b->stream));
  }
  PGX_CUDA(cudaStreamSynchronize(b->stream));
  return PGX_OK;
}

int pgx_bank_load_filter(pgx_bank* b, int32_t filter_index, const float* h) {
  if (!b || !h) return fail(PGX_ERR_INVALID, "NULL argument");
  const pgx_bank_config& c = b->cfg;
  if (filter_index < 0 || filter_index >= c.n_filters)
    return fail(PGX_ERR_INVALID, "filter_index %d outside [0,%d)", filter_index, c.n_filters);
  if (b->tail) return fail(PGX_ERR_INVALID, "a two-level bank keeps its filters: the tail contribution in flight was "
                                          "computed with them (create a new bank, or use tail_block = 0)");
  PGX_CUDA(cudaSetDevice(c.device));
  quiesce(b);
  return prep_filters(b, h, filter_index * c.filter_channels, c.filter_channels);
}

int pgx_bank_set_filter_map(pgx_bank* b, const int32_t* filter_of_stream) {
  if (!b || !filter_of_stream) return fail(PGX_ERR_INVALID, "NULL argument");
  if (b->tail) return fail(PGX_ERR_INVALID, "a two-level bank keeps its filter map (use tail_block = 0 for moving sources)");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  const int N = b->cfg.n_streams;
  for (int s = 0; s < N; ++s)
    if (filter_of_stream[s] < 0 || filter_of_stream[s] >= b->cfg.n_filters)
      return fail(PGX_ERR_INVALID, "filter_of_stream[%d]=%d outside [0,%d)", s, filter_of_stream[s], b->cfg.n_filters);
  // retire the row in use: everything that reads it has been enqueued on these two streams by now
  const int cur = b->fmap_cur, nxt = (cur + 1) % pgx_bank::kMapSlots;
  // (a caller-owned critical stream may be gone by now: run_pull recorded the event on it after each pull)
  if (!b->last_crit || b->last_crit == b->stream) PGX_CUDA(cudaEventRecord(b->fmap_ret_crit[cur], b->stream));
  PGX_CUDA(cudaEventRecord(b->ev_bgjoin, b->s_bg2));
  PGX_CUDA(cudaStreamWaitEvent(b->s_bg, b->ev_bgjoin, 0));
  PGX_CUDA(cudaEventRecord(b->fmap_ret_bg[cur], b->s_bg));
  if (b->fmap_sets + 1 >= pgx_bank::kMapSlots) {  // row nxt was used before: its readers and its copy are done?
    PGX_CUDA(cudaEventSynchronize(b->fmap_ret_crit[nxt]));
    PGX_CUDA(cudaEventSynchronize(b->fmap_ret_bg[nxt]));
    PGX_CUDA(cudaEventSynchronize(b->fmap_ev[nxt]));
  }
  int32_t* stage = b->fmap_pinned + (size_t)nxt * N;
  memcpy(stage, filter_of_stream, (size_t)N * sizeof(int32_t));
  PGX_CUDA(cudaMemcpyAsync(b->fmap_own + (size_t)nxt * N, stage, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice,
                           b->s_h2d));
  PGX_CUDA(cudaEventRecord(b->fmap_ev[nxt], b->s_h2d));
  b->fmap_cur = nxt;
  b->fmap_sets += 1;
  b->fmap = b->fmap_own + (size_t)nxt * N;
  b->fmap_pending = true;
  b->past_block = -1;  // a cached past sum was computed with the previous map
  b->tile_base = -1;
  return PGX_OK;
}

int pgx_bank_set_output_gains(pgx_bank* b, float wet, float dry) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (dry != 0.0f && ((b->cfg.flags & PGX_FLAG_MIXDOWN_INPUT) || b->cfg.c_in != b->cfg.c_out))
    return fail(PGX_ERR_INVALID, "a dry path needs as many source channels (%d) as output channels (%d) and no mix-down",
                b->cfg.c_in, b->cfg.c_out);
  b->wet = wet;  // host-side state read at enqueue time: applies to the pulls submitted from now on
  b->dry = dry;
  return PGX_OK;
}

int pgx_bank_use_filter_map_device(pgx_bank* b, const int32_t* fmap_dev) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (b->tail) return fail(PGX_ERR_INVALID, "a two-level bank keeps its filter map");
  b->fmap = fmap_dev ? const_cast<int32_t*>(fmap_dev) : b->fmap_own + (size_t)b->fmap_cur * b->cfg.n_streams;
  b->past_block = -1;  // a cached past sum was computed with the previous map
  b->tile_base = -1;
  return PGX_OK;
}

// ---- CUDA-graph replay of a small host pull ------------------------------------------------------------------
// A single-stream (or small) bank pulled one block at a time is a latency chain: H2D copy, two to four small kernels
// on three streams joined by events, D2H copy -- ~20 driver calls and several cross-stream hand-overs for a few
// microseconds of work (the reference's loop being replaced: renderer.py:297-327 / audio_renderer.py:214-236, one
// pull per callback).  For whole-block pulls in the steady state the SAME work is one instantiated graph
//     [H2D] -> A -> [K2] -> D2H           A = K1, or the fused k_conv1 / k_mix1<LAST> (then there is no K2)
//               \-> MAC(t+1) [-> fold]    the next block's past-partition pass, beside K2 as in the streamed schedule
// replayed with cudaGraphLaunch after its kernel-node arguments (ring position, block parity, x / filter-map
// pointers, gains) have been refreshed with cudaGraphExecKernelNodeSetParams: same kernels, same arguments, same
// buffers as run_step, one launch call.  Successive replays are ordered by the bank's stream, which is what the
// cross-block hazards of run_step need (every earlier step is complete).
struct StepArgs {
  pgx::R2CArgs r{};
  pgx::C2RArgs k{};
  pgx::MacArgs m{};
  const float4* fold_in = nullptr; float4* fold_out = nullptr; int fold_n = 0; int64_t fold_cols = 0;
  float2* ynow = nullptr; unsigned int* ticket = nullptr;
  pgx::FilterPrepArgs fp{};
};

static bool vec_ok2(const void* p, const pgx_layout& l, int pos) {
  return l.samp == 1 && (l.stream % 2) == 0 && (l.chan % 2) == 0 && (pos % 2) == 0 && (reinterpret_cast<uintptr_t>(p) % 8) == 0;
}

// Which graph shape a whole-block pull of this bank is (0 = none: the streamed schedule handles it).
static int graph_shape(const pgx_bank* b, bool mix) {
  const int R = b->R, P = b->P;
  if (!mix) {
    if (b->use_conv1 && (R == 1 || P <= b->fused_max_p)) return 1;
    if (P > 1 && b->plan_conv.variant == 0) return 3;
    return 0;
  }
  const bool mix1 = (R == 1 && b->use_mix1 && b->mix1_rows > 0);
  if (mix1 && b->mix1_rows <= 64 && b->cfg.c_out <= pgx::mix1_sources_per_cta(b->B, b->cfg.n_streams)) return 1;
  return 0;
}

// The arguments run_step would hand to its kernels for the whole-block step that is about to run.
static void graph_args(pgx_bank* b, int shape, bool mix, const float* x_dev, const pgx_layout& xl, float* y_dev,
                       const pgx_layout& yl, StepArgs* a) {
  const pgx_bank_config& c = b->cfg;
  const int B = b->B, R = b->R;
  const int64_t t = b->block;
  const int par = (int)(t & 1);
  pgx::R2CArgs& r = a->r;
  r.x = x_dev; r.xs = xl.stream; r.xc = xl.chan; r.xi = xl.samp; r.x_off = 0;
  r.hist = b->hist; r.fdl = b->fdl; r.tw = b->tw;
  r.n_fft = c.n_streams * b->c_x; r.c_in = c.c_in; r.c_x = b->c_x; r.B = B; r.R = R;
  r.slot = b->head; r.half = b->half; r.fill = 0; r.take = B;
  r.mixdown = (c.flags & PGX_FLAG_MIXDOWN_INPUT) ? 1 : 0;
  r.fast = (!r.mixdown && vec_ok2(x_dev, xl, 0)) ? 1 : 0;
  pgx::C2RArgs& k = a->k;
  k.yspec = b->ypast[par]; k.n_split = 0;
  k.n_out = mix ? c.c_out : c.n_streams * c.c_out;
  k.Hd = b->Hd; k.fmap = b->fmap; k.c_f = c.filter_channels; k.R = R; k.c_out = c.c_out; k.head = b->head;
  k.y = y_dev; k.ys = mix ? 0 : yl.stream; k.yc = yl.chan; k.yi = yl.samp; k.y_off = 0;
  k.B = B; k.fill = 0; k.take = B; k.tw = b->tw; k.wet = b->wet; k.fft16 = b->fft16;
  k.add = nullptr;
  pgx_layout ye = yl;
  if (mix) ye.stream = 0;
  if (mix) {  // k_mix1<LAST>: exactly the fields run_step's mix1 branch sets
    k.dry = 0.0f; k.xdry = nullptr;
    k.fast = vec_ok2(y_dev, ye, 0) ? 1 : 0;
    k.ynow = b->ynow; k.n_split_now = b->mix1_rows; k.fdl = nullptr;
    a->ynow = b->ynow; a->ticket = b->mix1_ticket;
    return;
  }
  k.c_x = b->c_x; k.dry = b->dry;
  k.xdry = (b->dry != 0.0f) ? x_dev : nullptr;
  k.xs = xl.stream; k.xc = xl.chan; k.xi = xl.samp; k.x_off = 0;
  k.fast = (vec_ok2(y_dev, ye, 0) && (!k.xdry || vec_ok2(x_dev, xl, 0))) ? 1 : 0;
  k.ynow = nullptr; k.n_split_now = 0; k.fdl = b->fdl;
  if (shape == 1) {
    if (b->P > 1) {  // past partitions summed inside the fused kernel
      k.n_past = R - 2;
      k.q0 = R - 1 - b->head;
      if (b->head + 1 < R) { k.p_off = 0; k.p_skip = b->head; k.p_nskip = 2; }
      else                 { k.p_off = 1; k.p_skip = R; k.p_nskip = 0; }
    }
    return;
  }
  // shape 3: K2 sums the past pass of THIS block (issued by the previous step) ...
  const pgx::MacPlan& pl = b->plan_conv;
  k.n_split = pl.n_partials > kFoldAbove ? 1 : pl.n_partials;
  // ... and the past pass of the NEXT block runs beside it
  const int nhead = (b->head + 1) % R, npar = par ^ 1;
  pgx::MacArgs& m = a->m;
  fill_mac_common(b, m, false, nhead);
  const bool fold = pl.n_partials > kFoldAbove;
  m.mix = pl.layout;
  m.yspec = reinterpret_cast<float4*>(fold ? b->ypart[npar] : b->ypast[npar]);
  m.Pt = R - 2; m.jfix = -1;
  if (nhead + 1 < R) { m.off = 0; m.skip = nhead; m.nskip = 2; }
  else               { m.off = 1; m.skip = R; m.nskip = 0; }
  m.n_terms = m.Pt;
  m.n_split = pl.n_split; m.terms_per_split = pl.terms_per_split; m.n_otiles = pl.n_otiles; m.st = pl.st;
  m.variant = pl.variant; m.persistent_ctas = pl.persistent_ctas;
  a->fold_in = reinterpret_cast<const float4*>(b->ypart[npar]);
  a->fold_out = reinterpret_cast<float4*>(b->ypast[npar]);
  a->fold_n = pl.n_partials;
  a->fold_cols = (int64_t)m.n_out * (B / 2);
}

static cudaKernelNodeParams knode(const pgx::LaunchDesc& d, void** params) {
  cudaKernelNodeParams p{};
  p.func = const_cast<void*>(d.func);
  p.gridDim = d.grid; p.blockDim = d.block; p.sharedMemBytes = d.smem;
  p.kernelParams = params; p.extra = nullptr;
  return p;
}

// One whole-block host pull as a graph replay.  Returns PGX_OK and *done = true when the pull was enqueued this way;
// *done = false (nothing enqueued) when the pull is not graph material.
static int graph_pull(pgx_bank* b, const float* x, const pgx_layout& xl, int slot, const pgx_layout& yd, int32_t n, bool mix,
                      bool x_device, size_t xb, size_t yb, bool* done) {
  *done = false;
  const pgx_bank_config& c = b->cfg;
  if (!b->use_graph || b->tail || b->profiling || b->serial || b->fill != 0 || n != b->B) return PGX_OK;
  if (b->tile > 1 && !mix) return PGX_OK;   // time-tiled conv banks are throughput banks: the streamed schedule
  if ((size_t)c.n_streams * b->c_x * b->B > (size_t)(1 << 18)) return PGX_OK;   // big banks: the streamed schedule overlaps better
  const int shape = graph_shape(b, mix);
  if (shape == 0) return PGX_OK;
  if (shape == 3 && !(b->past_block == b->block && b->past_mode == 0)) return PGX_OK;  // this block's past sum must exist
  if (b->block < 2) return PGX_OK;                                    // steady state only
  // tiny pulls skip the copy engines altogether: the kernels read x from / write y to the pinned bounce buffers
  // directly (zero-copy over PCIe: a few KB of independent, coalesced accesses), which takes two dependent copies
  // and their hand-overs off the latency chain
  const bool zc_x = !x_device && xb <= pgx_bank::kZeroCopyMax, zc_y = yb <= pgx_bank::kZeroCopyMax;
  const float* x_dev = x_device ? x : (zc_x ? reinterpret_cast<const float*>(b->hx_bounce[slot]) : b->x_stage[slot]);
  float* y_dev = zc_y ? reinterpret_cast<float*>(b->hy_bounce[slot]) : b->y_stage[slot];
  StepArgs a;
  graph_args(b, shape, mix, x_dev, xl, y_dev, yd, &a);
  pgx::LaunchDesc dA, dMac, dFold, dK2;
  bool ok = true;
  // shape 3 on a bank whose output stage sums at most kFoldAbove past rows: ingest, transforms, present term, past rows
  // and emit as ONE fused launch (k_conv1<.., PAST> with n_past = 0 and the folded rows), K2 drops off the chain
  const bool fuse3 = shape == 3 && b->graph_fuse && b->use_conv1;
  if (shape == 1) ok = mix ? pgx::describe_mix1(a.r, true, &dA) : pgx::describe_conv1(a.r, a.k, &dA);
  else if (fuse3) ok = pgx::describe_conv1(a.r, a.k, &dA) && pgx::describe_fdl_mac(a.m, &dMac);
  else ok = pgx::describe_r2c_ingest(a.r, &dA) && pgx::describe_fdl_mac(a.m, &dMac) && pgx::describe_c2r_emit(a.k, &dK2);
  const bool fold = shape == 3 && b->plan_conv.n_partials > kFoldAbove;
  if (fold) ok = ok && pgx::describe_reduce_partials(a.m.n_out, b->B / 2, &dFold);
  if (!ok) return PGX_OK;
  void* pA_fused_conv[] = {&a.r, &a.k};
  void* pA_mix1[] = {&a.r, &a.k, &a.ynow, &a.ticket};
  void* pA_k1[] = {&a.r, &a.fp};
  void** pA = (shape == 3 && !fuse3) ? pA_k1 : (mix ? pA_mix1 : pA_fused_conv);
  void* pMac[] = {&a.m};
  void* pFold[] = {&a.fold_in, &a.fold_out, &a.fold_n, &a.fold_cols};
  void* pK2[] = {&a.k};

  pgx_bank::PullGraph& g = b->pgraph[slot][mix ? 1 : 0][x_device ? 1 : 0];
  const bool stale = g.exec && (g.shape != shape || g.f_a != dA.func || g.f_mac != dMac.func || g.f_k2 != dK2.func ||
                                g.fold != fold || g.xb != xb || g.yb != yb);
  if (stale) {
    cudaGraphExecDestroy(g.exec);
    cudaGraphDestroy(g.graph);
    g = pgx_bank::PullGraph{};
  }
  if (!g.exec) {
    // (a failure half-way must not leave a half-built graph behind: the next pull would build on top of it)
    cudaError_t ge = cudaSuccess;
    auto step = [&](cudaError_t e) { if (ge == cudaSuccess) ge = e; return ge == cudaSuccess; };
    if (g.graph) { cudaGraphDestroy(g.graph); g = pgx_bank::PullGraph{}; }
    cudaGraphNode_t dep = nullptr, last = nullptr;
    cudaKernelNodeParams kp{};
    if (step(cudaGraphCreate(&g.graph, 0)) && !x_device && !zc_x &&
        step(cudaGraphAddMemcpyNode1D(&g.n_h2d, g.graph, nullptr, 0, b->x_stage[slot], b->hx_bounce[slot], xb,
                                      cudaMemcpyHostToDevice)))
      dep = g.n_h2d;
    kp = knode(dA, pA);
    if (step(ge) && step(cudaGraphAddKernelNode(&g.n_a, g.graph, dep ? &dep : nullptr, dep ? 1 : 0, &kp))) last = g.n_a;
    if (shape == 3 && step(ge)) {
      if (!fuse3) {
        kp = knode(dK2, pK2);
        if (step(cudaGraphAddKernelNode(&g.n_k2, g.graph, &g.n_a, 1, &kp))) last = g.n_k2;
      }
      kp = knode(dMac, pMac);
      step(cudaGraphAddKernelNode(&g.n_mac, g.graph, &g.n_a, 1, &kp));
      if (fold && step(ge)) {
        kp = knode(dFold, pFold);
        step(cudaGraphAddKernelNode(&g.n_fold, g.graph, &g.n_mac, 1, &kp));
      }
    }
    if (!zc_y && step(ge))
      step(cudaGraphAddMemcpyNode1D(&g.n_d2h, g.graph, &last, 1, b->hy_bounce[slot], b->y_stage[slot], yb,
                                    cudaMemcpyDeviceToHost));
    if (step(ge)) step(cudaGraphInstantiate(&g.exec, g.graph, 0));
    if (ge != cudaSuccess) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      if (g.graph) cudaGraphDestroy(g.graph);
      g = pgx_bank::PullGraph{};
      return fail(PGX_ERR_CUDA, "graph replay: %s", cudaGetErrorString(ge));
    }
    g.shape = shape; g.f_a = dA.func; g.f_mac = dMac.func; g.f_k2 = dK2.func; g.fold = fold; g.xb = xb; g.yb = yb;
  } else {  // refresh the arguments that move from step to step
    cudaKernelNodeParams kp = knode(dA, pA);
    PGX_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.n_a, &kp));
    if (shape == 3) {
      if (!fuse3) {
        kp = knode(dK2, pK2);
        PGX_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.n_k2, &kp));
      }
      kp = knode(dMac, pMac);
      PGX_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.n_mac, &kp));
      if (fold) {
        kp = knode(dFold, pFold);
        PGX_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.n_fold, &kp));
      }
    }
  }
  // entering from the streamed schedule: everything it left on the other streams comes first
  if (!b->after_graph) {
    int e = 0;
    for (cudaStream_t s : {b->s_in, b->s_bg, b->s_bg2, b->s_h2d}) {
      PGX_CUDA(cudaEventRecord(b->ev_join[e], s));
      PGX_CUDA(cudaStreamWaitEvent(b->stream, b->ev_join[e], 0));
      ++e;
    }
    // (a caller-owned critical stream may be gone by now: its last output stage recorded ev_k2)
    if (b->step >= 1) PGX_CUDA(cudaStreamWaitEvent(b->stream, b->ev_k2[(b->step - 1) % kRing], 0));
  }
  if (b->fmap_pending) {
    PGX_CUDA(cudaStreamWaitEvent(b->stream, b->fmap_ev[b->fmap_cur], 0));
    b->fmap_pending = false;
  }
  PGX_CUDA(cudaGraphLaunch(g.exec, b->stream));
  // bookkeeping of run_step for a completed block
  const int par = (int)(b->block & 1);
  b->last_crit = b->stream;
  b->last_k2_of_par[par] = b->step;
  b->prev_on_crit = (shape == 1);
  b->launches += shape == 1 ? 1 : (fold ? 4 : 3) - (fuse3 ? 1 : 0);
  b->steps += 1;
  b->step += 1;
  if (shape == 3) { b->past_block = b->block + 1; b->past_mode = 0; }
  b->head = (b->head + 1) % b->R;
  b->half ^= 1;
  b->block += 1;
  b->after_graph = true;
  b->graph_pulls += 1;
  *done = true;
  return PGX_OK;
}

// Host-buffer pull, asynchronous: stage x into the next slot (H2D on the copy-in stream), enqueue the block
// steps, copy y back on the copy-out stream.  Returns a ticket; y is complete after submit_wait(ticket).
static int submit_host(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n, bool mix,
                       int64_t* ticket, bool x_device = false, bool x_pcm = false, bool y_pcm = false,
                       bool reduce = false) {
  int rc = check_pull_args(b, x, y, n);
  if (rc != PGX_OK) return rc;
  const pgx_bank_config& c = b->cfg;
  if (reduce) {
    if (!mix) return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE needs PGX_PULL_MIX: only a mix is summed over the ranks");
    if (!b->comm) return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE: no communicator attached (pgx_bank_attach_comm)");
    if (c.c_out * n > pgx_comm_max_floats(b->comm))
      return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE: %d floats exceed the communicator's %d", c.c_out * n,
                  pgx_comm_max_floats(b->comm));
  }
  const bool deliver = !reduce || pgx_comm_is_root(b->comm);  // only the root of a reduce copies y back
  // host x is staged with one copy, so it has to be one dense block; device-resident x is read in place
  if (!x_device && !layout_dense(xl, c.n_streams, c.c_in, n))
    return fail(PGX_ERR_INVALID, "x layout does not tile a dense block");
  pgx_layout yd = yl;
  if (mix) yd.stream = 0;
  if (!layout_dense(yd, mix ? 1 : c.n_streams, c.c_out, n))
    return fail(PGX_ERR_INVALID, "y layout does not tile a dense block");
  PGX_CUDA(cudaSetDevice(c.device));
  const size_t xb = (size_t)c.n_streams * c.c_in * n * sizeof(float);
  const size_t yb = (size_t)(mix ? 1 : c.n_streams) * c.c_out * n * sizeof(float);
  const int64_t tk = b->next_ticket;
  const int slot = (int)(tk % b->n_slots);
  // the slot's previous pull is over once its D2H has completed (its K1s read x_stage before that)
  if (tk >= b->n_slots) PGX_CUDA(cudaEventSynchronize(b->ev_done[slot]));
  if (b->y_user[slot]) {  // that pull was never waited for: its result still has to reach the caller's array
    memcpy(b->y_user[slot], b->hy_bounce[slot], b->y_user_bytes[slot]);
    b->y_user[slot] = nullptr;
  }
  const size_t xb_host = x_pcm ? xb / 2 : xb, yb_host = y_pcm ? yb / 2 : yb;
  const bool bounce_x = !x_device && xb_host <= pgx_bank::kBounceMax;
  const bool bounce_y = deliver && yb_host <= pgx_bank::kBounceMax;
  if (bounce_x && !b->hx_bounce[slot]) {
    size_t cap = b->xs_bytes < pgx_bank::kBounceMax ? b->xs_bytes : pgx_bank::kBounceMax;
    PGX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->hx_bounce[slot]), cap, cudaHostAllocDefault));
  }
  if (bounce_y && !b->hy_bounce[slot]) {
    size_t cap = b->ys_bytes < pgx_bank::kBounceMax ? b->ys_bytes : pgx_bank::kBounceMax;
    PGX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->hy_bounce[slot]), cap, cudaHostAllocDefault));
  }
  const void* x_src = x;
  if (bounce_x) {
    memcpy(b->hx_bounce[slot], x, xb_host);
    x_src = b->hx_bounce[slot];
  }
  void* y_dst = bounce_y ? static_cast<void*>(b->hy_bounce[slot]) : static_cast<void*>(y);
  if (x_pcm && !b->xpcm_stage[slot]) PGX_CUDA(cudaMalloc(&b->xpcm_stage[slot], b->xs_bytes / 2));
  if (y_pcm && !b->ypcm_stage[slot]) PGX_CUDA(cudaMalloc(&b->ypcm_stage[slot], b->ys_bytes / 2));
  if ((bounce_x || x_device) && bounce_y && !x_pcm && !y_pcm && !reduce) {  // small whole-block pull: one graph replay
    bool done = false;
    rc = graph_pull(b, x, xl, slot, yd, n, mix, x_device, xb, yb, &done);
    if (rc != PGX_OK) return rc;
    if (done) {
      PGX_CUDA(cudaEventRecord(b->ev_done[slot], b->stream));
      b->y_user[slot] = y;
      b->y_user_bytes[slot] = yb_host;
      b->next_ticket = tk + 1;
      if (ticket) *ticket = tk;
      return PGX_OK;
    }
  }
  if (x_device) {  // produced by work already queued on the bank's stream: run_pull orders the ingest after it
    rc = run_pull(b, x, xl, b->y_stage[slot], yd, n, mix, false, b->stream);
  } else {
    if (x_pcm) {   // half the H2D bytes; int16 / 32768 on the device, on the copy-in stream
      PGX_CUDA(cudaMemcpyAsync(b->xpcm_stage[slot], x_src, xb / 2, cudaMemcpyHostToDevice, b->s_h2d));
      pgx::launch_pcm16_to_f32(b->xpcm_stage[slot], b->x_stage[slot], (int64_t)(xb / sizeof(float)), b->s_h2d);
      b->launches += 1;
    } else
    PGX_CUDA(cudaMemcpyAsync(b->x_stage[slot], x_src, xb, cudaMemcpyHostToDevice, b->s_h2d));
    PGX_CUDA(cudaEventRecord(b->ev_h2d[slot], b->s_h2d));
    PGX_CUDA(cudaStreamWaitEvent(b->s_in, b->ev_h2d[slot], 0));
    // kernels on the bank's stream that read x themselves (fused single-partition steps, the dry path of the output
    // stage) need the copy too; on a tiled conv bank without dry gain only K1 (ingest stream) reads x
    if (!(b->tile > 1 && !mix && b->dry == 0.0f && !b->serial)) PGX_CUDA(cudaStreamWaitEvent(b->stream, b->ev_h2d[slot], 0));
    rc = run_pull(b, b->x_stage[slot], xl, b->y_stage[slot], yd, n, mix, true, b->stream);
  }
  if (rc != PGX_OK) {  // part of the pull may be enqueued and reading x_stage[slot]: drain before the slot is reused
    const std::string msg = g_err;
    quiesce(b);
    g_err = msg;
    return rc;
  }
  if (reduce) {    // sum of the ranks' partial mixes onto the root, behind this pull's output stage
    rc = pgx_comm_enqueue(b->comm, b->y_stage[slot], b->y_stage[slot], c.c_out * n, b->stream);
    if (rc != PGX_OK) return rc;
    b->launches += 1;
  }
  if (y_pcm && deliver) {  // clip(lrint(y * 32768)) on the device, half the D2H bytes
    pgx::launch_f32_to_pcm16(b->y_stage[slot], b->ypcm_stage[slot], (int64_t)(yb / sizeof(float)), b->stream);
    b->launches += 1;
  }
  if (reduce || y_pcm || b->tail || b->step < 1) {
    PGX_CUDA(cudaEventRecord(b->ev_y[slot], b->stream));
    PGX_CUDA(cudaStreamWaitEvent(b->s_d2h, b->ev_y[slot], 0));
  } else {  // the pull's last operation on the bank's stream is its last output stage, which recorded ev_k2
    PGX_CUDA(cudaStreamWaitEvent(b->s_d2h, b->ev_k2[(b->step - 1) % kRing], 0));
  }
  if (deliver) {
    if (y_pcm) PGX_CUDA(cudaMemcpyAsync(y_dst, b->ypcm_stage[slot], yb / 2, cudaMemcpyDeviceToHost, b->s_d2h));
    else
    PGX_CUDA(cudaMemcpyAsync(y_dst, b->y_stage[slot], yb, cudaMemcpyDeviceToHost, b->s_d2h));
  }
  PGX_CUDA(cudaEventRecord(b->ev_done[slot], b->s_d2h));
  if (bounce_y) {
    b->y_user[slot] = y;
    b->y_user_bytes[slot] = yb_host;
  }
  b->next_ticket = tk + 1;
  if (ticket) *ticket = tk;
  return PGX_OK;
}

static int submit_wait(pgx_bank* b, int64_t ticket) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (ticket < 0 || ticket >= b->next_ticket) return fail(PGX_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
  if (ticket + b->n_slots < b->next_ticket) return PGX_OK;  // its slot was recycled: long complete
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  const int slot = (int)(ticket % b->n_slots);
  PGX_CUDA(cudaEventSynchronize(b->ev_done[slot]));
  if (b->y_user[slot]) {  // small result: out of the bank's pinned bounce buffer into the caller's array
    memcpy(b->y_user[slot], b->hy_bounce[slot], b->y_user_bytes[slot]);
    b->y_user[slot] = nullptr;
  }
  if (b->comm) return pgx_comm_check(b->comm);  // a reduce whose peer never arrived gave up instead of hanging
  return PGX_OK;
}

static int process_host(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n, bool mix) {
  int64_t tk = 0;
  const int rc = submit_host(b, x, xl, y, yl, n, mix, &tk);
  return rc != PGX_OK ? rc : submit_wait(b, tk);
}

int pgx_bank_submit(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n, int32_t flags,
                    int64_t* ticket) {
  if (!ticket) return fail(PGX_ERR_INVALID, "ticket is NULL");
  if ((flags & PGX_PULL_X_DEVICE) && (flags & PGX_PULL_X_PCM16))
    return fail(PGX_ERR_INVALID, "PGX_PULL_X_DEVICE and PGX_PULL_X_PCM16 exclude each other");
  return submit_host(b, x, xl, y, yl, n, (flags & PGX_PULL_MIX) != 0, ticket, (flags & PGX_PULL_X_DEVICE) != 0,
                     (flags & PGX_PULL_X_PCM16) != 0, (flags & PGX_PULL_Y_PCM16) != 0, (flags & PGX_PULL_REDUCE) != 0);
}

int pgx_bank_pull(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n, int32_t flags) {
  int64_t tk = 0;
  const int rc = pgx_bank_submit(b, x, xl, y, yl, n, flags, &tk);
  return rc != PGX_OK ? rc : submit_wait(b, tk);
}

void* pgx_bank_stream(pgx_bank* b) { return b ? static_cast<void*>(b->stream) : nullptr; }

int pgx_bank_wait(pgx_bank* b, int64_t ticket) { return submit_wait(b, ticket); }

int pgx_bank_process(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n) {
  return process_host(b, x, xl, y, yl, n, false);
}

int pgx_bank_process_mix(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n) {
  return process_host(b, x, xl, y, yl, n, true);
}

int pgx_bank_process_device(pgx_bank* b, const float* x_dev, pgx_layout xl, float* y_dev, pgx_layout yl, int32_t n,
                            int32_t flags, void* cuda_stream) {
  const bool mix = (flags & PGX_PULL_MIX) != 0;
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (!x_dev || !y_dev) return fail(PGX_ERR_INVALID, "x / y must not be NULL");
  if (n < 1) return fail(PGX_ERR_INVALID, "pull of %d samples", n);
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : b->stream;
  pgx_layout yd = yl;
  if (mix) yd.stream = 0;
  const bool reduce = (flags & PGX_PULL_REDUCE) != 0;
  if (reduce) {
    if (!mix) return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE needs PGX_PULL_MIX: only a mix is summed over the ranks");
    if (!b->comm) return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE: no communicator attached (pgx_bank_attach_comm)");
    if (!layout_dense(yd, 1, b->cfg.c_out, n)) return fail(PGX_ERR_INVALID, "PGX_PULL_REDUCE: y must be one dense block");
  }
  const int rc = run_pull(b, x_dev, xl, y_dev, yd, n, mix, (flags & PGX_PULL_INPUT_RESIDENT) != 0, st);
  if (rc != PGX_OK || !reduce) return rc;
  b->launches += 1;
  return pgx_comm_enqueue(b->comm, y_dev, y_dev, b->cfg.c_out * n, st);  // in place: the root's y_dev becomes the sum
}

int pgx_bank_attach_comm(pgx_bank* b, pgx_comm* comm) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (comm && pgx_comm_device(comm) != b->cfg.device)
    return fail(PGX_ERR_INVALID, "communicator lives on device %d, bank on device %d", pgx_comm_device(comm), b->cfg.device);
  b->comm = comm;
  return PGX_OK;
}

int pgx_bank_synchronize(pgx_bank* b) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (b->tail) {
    const int rc = pgx_bank_synchronize(b->tail);
    if (rc != PGX_OK) return rc;
  }
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  PGX_CUDA(cudaStreamSynchronize(b->s_h2d));
  PGX_CUDA(cudaStreamSynchronize(b->stream));
  PGX_CUDA(cudaStreamSynchronize(b->s_in));
  PGX_CUDA(cudaStreamSynchronize(b->s_bg));
  PGX_CUDA(cudaStreamSynchronize(b->s_bg2));
  PGX_CUDA(cudaStreamSynchronize(b->s_d2h));
  return PGX_OK;
}

int pgx_bank_profile_begin(pgx_bank* b) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (b->tail) pgx_bank_profile_begin(b->tail);
  b->profiling = true;
  b->prof_spans.clear();
  b->prof_pool_used = 0;
  b->prof_steps = 0;
  return PGX_OK;
}

int pgx_bank_profile_end(pgx_bank* b, pgx_profile* out) {
  if (!b || !out) return fail(PGX_ERR_INVALID, "NULL argument");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  b->profiling = false;
  out->ms_r2c = out->ms_mac = out->ms_c2r = out->ms_fold = out->ms_now = 0.0;
  out->n_mac = 0;
  out->ms_mac_union = 0.0;
  out->ms_conv1 = 0.0;
  out->steps = b->prof_steps;
  PGX_CUDA(cudaDeviceSynchronize());
  std::vector<std::pair<float, float>> mac_iv;  // K3 past-pass launches as [start, end) from the first event
  cudaEvent_t base = b->prof_spans.empty() ? nullptr : b->prof_spans.front().a;
  for (const pgx_bank::ProfSpan& sp : b->prof_spans) {
    float ms = 0.f;
    PGX_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
    if (sp.kind == 0) out->ms_r2c += ms;
    else if (sp.kind == 1) {
      out->ms_mac += ms;
      out->n_mac += 1;
      float t0 = 0.f;
      if (sp.a != base) PGX_CUDA(cudaEventElapsedTime(&t0, base, sp.a));
      mac_iv.emplace_back(t0, t0 + ms);
    }
    else if (sp.kind == 2) out->ms_c2r += ms;
    else if (sp.kind == 3) out->ms_fold += ms;
    else if (sp.kind == 5) out->ms_conv1 += ms;
    else out->ms_now += ms;
  }
  // launches of consecutive blocks overlap on the two background streams: the kernel's busy time is the
  // union of their intervals
  std::sort(mac_iv.begin(), mac_iv.end());
  float cur0 = 0.f, cur1 = -1.f;
  for (const auto& iv : mac_iv) {
    if (cur1 < cur0 || iv.first > cur1) {
      if (cur1 > cur0) out->ms_mac_union += cur1 - cur0;
      cur0 = iv.first;
      cur1 = iv.second;
    } else if (iv.second > cur1) {
      cur1 = iv.second;
    }
  }
  if (cur1 > cur0) out->ms_mac_union += cur1 - cur0;
  b->prof_spans.clear();
  b->prof_pool_used = 0;
  if (b->tail) {  // the tail level's kernels belong to the same steps
    pgx_profile t{};
    const int rc = pgx_bank_profile_end(b->tail, &t);
    if (rc != PGX_OK) return rc;
    out->ms_r2c += t.ms_r2c; out->ms_mac += t.ms_mac; out->ms_c2r += t.ms_c2r; out->ms_fold += t.ms_fold;
    out->ms_now += t.ms_now; out->ms_conv1 += t.ms_conv1; out->ms_mac_union += t.ms_mac_union; out->n_mac += t.n_mac;
  }
  return PGX_OK;
}

int pgx_nearest_direction(const double* tab_elev, const double* tab_az, int32_t n_tab, const double* azimuth,
                          const double* elevation, int32_t n, int32_t* out_index) {
  if (!tab_elev || !tab_az || !azimuth || !elevation || !out_index || n_tab < 1 || n < 0)
    return fail(PGX_ERR_INVALID, "pgx_nearest_direction: bad arguments");
  nearest_scan(tab_elev, tab_az, n_tab, azimuth, elevation, n, out_index);
  return PGX_OK;
}

int pgx_mix_sum_device(int32_t device, const float* in_dev, int32_t n_inputs, int64_t n_elems, float* out_dev,
                       void* cuda_stream) {
  if (!in_dev || !out_dev) return fail(PGX_ERR_INVALID, "NULL argument");
  if (n_inputs < 1 || n_elems < 0) return fail(PGX_ERR_INVALID, "bad sizes");
  PGX_CUDA(cudaSetDevice(device));
  if (n_elems == 0) return PGX_OK;
  pgx::launch_mix_sum(in_dev, n_inputs, n_elems, out_dev, static_cast<cudaStream_t>(cuda_stream));
  PGX_CUDA(cudaGetLastError());
  return PGX_OK;
}

int pgx_mix_sum(int32_t device, const float* inputs, int32_t n_inputs, int64_t n_elems, float* out) {
  if (!inputs || !out) return fail(PGX_ERR_INVALID, "NULL argument");
  if (n_inputs < 1 || n_elems < 0) return fail(PGX_ERR_INVALID, "bad sizes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(PGX_ERR_NO_DEVICE, "no CUDA device: libpgx has no CPU fallback");
  PGX_CUDA(cudaSetDevice(device));
  if (n_elems == 0) return PGX_OK;
  float *in_dev = nullptr, *out_dev = nullptr;
  const size_t ib = (size_t)n_inputs * n_elems * sizeof(float), ob = (size_t)n_elems * sizeof(float);
  PGX_CUDA(cudaMalloc(&in_dev, ib));
  cudaError_t e = cudaMalloc(&out_dev, ob);
  if (e != cudaSuccess) {
    cudaFree(in_dev);
    return fail(PGX_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  int rc = PGX_OK;
  e = cudaMemcpy(in_dev, inputs, ib, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    pgx::launch_mix_sum(in_dev, n_inputs, n_elems, out_dev, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, out_dev, ob, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) rc = fail(PGX_ERR_CUDA, "pgx_mix_sum: %s", cudaGetErrorString(e));
  cudaFree(in_dev);
  cudaFree(out_dev);
  return rc;
}

}  // extern "C"
