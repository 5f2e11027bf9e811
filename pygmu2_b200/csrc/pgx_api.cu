// pgx_api.cu -- host engine + C ABI of libpgx.so (see include/pgx.h).
//
// A bank owns, on one GPU, for N lock-stepped audio streams:
//   hist  [N*c_x][2][B]      float   previous block / open block (time domain)
//   fdl   [N*c_x][P][B]      float2  frequency-domain delay line: packed spectra of the last P windows
//   Hd    [F*c_f][2P][B]     float2  filter partition spectra, reversed + doubled, scaled 1/B
//   yspec [n_split][n_out][B] float2 split partial sums of the multiply-accumulate
// and advances them one "block step" at a time.  A pull of n samples is cut at block boundaries; a
// partially filled block is transformed with zeros in the not-yet-known positions (causality makes the
// emitted samples exact) and re-transformed when more samples arrive, so any (start, duration) pull
// pattern is zero-latency like the reference (convolve_pe.py:250-342), while the block grid stays
// aligned for the partitioned filter.
//
// Schedule of one block step (two CUDA streams):
//   critical stream : K1 ingest + R2C of the open block  ->  [mix mode: K3 over the open slot of every
//                     stream]  ->  K2: past partial sums + present term, C2R, emit
//   background      : K3 over the P-1 *past* partitions of the open block.  It depends only on rows
//                     committed before the block opened, so it is launched as soon as the previous
//                     block's K1 has run, overlaps K1/K2, and is reused by every partial pull of the block.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pgx.h"
#include "kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define PGX_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(e__ == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                 \
  } while (0)

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

}  // namespace

struct pgx_bank {
  pgx_bank_config cfg{};
  int c_x = 1, P = 1, B = 0;
  cudaStream_t stream = nullptr;  // default critical stream
  cudaStream_t bg = nullptr;      // background stream for the past-partition pass
  cudaEvent_t ev_fork = nullptr, ev_past = nullptr;
  bool ev_past_recorded = false;
  float* hist = nullptr;
  float2* fdl = nullptr;
  float2* Hd = nullptr;
  float2* ypast[2] = {nullptr, nullptr};  // past-partition partial sums, double buffered by block parity
  float2* ynow = nullptr;                 // present-slot partial sums (mix mode)
  float2* tw = nullptr;
  int32_t* fmap = nullptr;
  int32_t* fmap_pinned = nullptr;
  float* x_stage = nullptr;
  float* y_stage = nullptr;
  size_t hist_bytes = 0, fdl_bytes = 0, Hd_bytes = 0, yspec_bytes = 0, xs_bytes = 0, ys_bytes = 0;
  int head = 0, fill = 0, half = 0;
  int par = 0;               // which ypast buffer belongs to the open block
  bool past_valid = false;   // ypast[par] holds the past sum of the open block ...
  int past_mode = -1;        // ... for this mode (0 conv, 1 mix)
  pgx::MacPlan plan_conv{}, plan_mix{}, plan_now{};
  int sm_count = 148;
  int64_t launches = 0, steps = 0;
  bool profiling = false;
  struct ProfSpan { int kind; cudaEvent_t a, b; };  // kind 0 = K1, 1 = K3, 2 = K2
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_pool_used = 0;
  int64_t prof_steps = 0;
};

namespace {

void free_bank(pgx_bank* b) {
  if (!b) return;
  cudaSetDevice(b->cfg.device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->hist);
  cudaFree(b->fdl);
  cudaFree(b->Hd);
  cudaFree(b->ypast[0]);
  cudaFree(b->ypast[1]);
  cudaFree(b->ynow);
  cudaFree(b->tw);
  cudaFree(b->fmap);
  cudaFree(b->x_stage);
  cudaFree(b->y_stage);
  for (cudaEvent_t e : b->prof_pool) cudaEventDestroy(e);
  if (b->bg) cudaStreamSynchronize(b->bg);
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_past) cudaEventDestroy(b->ev_past);
  if (b->bg) cudaStreamDestroy(b->bg);
  if (b->fmap_pinned) cudaFreeHost(b->fmap_pinned);
  if (b->stream) cudaStreamDestroy(b->stream);
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

void fill_mac_common(pgx_bank* b, pgx::MacArgs& m, bool mix) {
  const pgx_bank_config& c = b->cfg;
  m.fdl = reinterpret_cast<const float4*>(b->fdl);
  m.Hd = reinterpret_cast<const float4*>(b->Hd);
  m.fmap = b->fmap;
  m.N = c.n_streams; m.c_x = b->c_x; m.c_out = c.c_out; m.c_f = c.filter_channels; m.P = b->P; m.W4 = b->B / 2;
  m.q0 = b->P - 1 - b->head;
  m.mix = mix ? 1 : 0;
  m.n_out = mix ? c.c_out : c.n_streams * c.c_out;
}

// Background pass: sum over the P-1 committed partitions of the open block into ypast[par].
// Forked from `after` (an event on the critical stream after which every committed row is written).
void launch_past(pgx_bank* b, bool mix, cudaStream_t crit) {
  cudaEventRecord(b->ev_fork, crit);
  cudaStreamWaitEvent(b->bg, b->ev_fork, 0);
  pgx::MacArgs m{};
  fill_mac_common(b, m, mix);
  const pgx::MacPlan& pl = mix ? b->plan_mix : b->plan_conv;
  m.yspec = reinterpret_cast<float4*>(b->ypast[b->par]);
  m.Pt = b->P - 1; m.skip = b->head; m.jfix = -1;
  m.n_terms = mix ? b->cfg.n_streams * (b->P - 1) : (b->P - 1);
  m.n_split = pl.n_split; m.terms_per_split = pl.terms_per_split; m.n_otiles = pl.n_otiles; m.st = pl.st;
  {
    ProfScope ps(b, b->bg, 1);
    pgx::launch_fdl_mac(m, b->bg);
  }
  cudaEventRecord(b->ev_past, b->bg);
  b->ev_past_recorded = true;
  b->launches += 1;
  b->past_valid = true;
  b->past_mode = mix ? 1 : 0;
}

// One pull on device buffers, enqueued on st (no synchronisation).
int run_pull(pgx_bank* b, const float* x_dev, const pgx_layout& xl, float* y_dev, const pgx_layout& yl, int n,
             bool mix, cudaStream_t st) {
  const pgx_bank_config& c = b->cfg;
  const int B = b->B, P = b->P;
  int pos = 0;
  while (pos < n) {
    const int take = (B - b->fill < n - pos) ? (B - b->fill) : (n - pos);

    pgx::R2CArgs r{};
    r.x = x_dev; r.xs = xl.stream; r.xc = xl.chan; r.xi = xl.samp; r.x_off = pos;
    r.hist = b->hist; r.fdl = b->fdl; r.tw = b->tw;
    r.n_fft = c.n_streams * b->c_x; r.c_in = c.c_in; r.c_x = b->c_x; r.B = B; r.P = P;
    r.slot = b->head; r.half = b->half; r.fill = b->fill; r.take = take;
    r.mixdown = (c.flags & PGX_FLAG_MIXDOWN_INPUT) ? 1 : 0;
    {
      ProfScope ps(b, st, 0);
      pgx::launch_r2c_ingest(r, st);
    }
    b->launches += 1;

    const bool completes = (b->fill + take == B);
    int n_split_past = 0;
    if (P > 1) {
      if (!b->past_valid || b->past_mode != (mix ? 1 : 0)) launch_past(b, mix, st);
      cudaStreamWaitEvent(st, b->ev_past, 0);
      n_split_past = (mix ? b->plan_mix : b->plan_conv).n_split;
    }

    pgx::C2RArgs k{};
    k.yspec = b->ypast[b->par]; k.n_split = n_split_past;
    k.n_out = mix ? c.c_out : c.n_streams * c.c_out;
    if (mix) {  // present term of every stream: K3 restricted to the open slot
      pgx::MacArgs m{};
      fill_mac_common(b, m, true);
      m.yspec = reinterpret_cast<float4*>(b->ynow);
      m.Pt = 1; m.skip = P; m.jfix = b->head;
      m.n_terms = c.n_streams;
      m.n_split = b->plan_now.n_split; m.terms_per_split = b->plan_now.terms_per_split;
      m.n_otiles = b->plan_now.n_otiles; m.st = b->plan_now.st;
      {
        ProfScope ps(b, st, 1);
        pgx::launch_fdl_mac(m, st);
      }
      b->launches += 1;
      k.ynow = b->ynow; k.n_split_now = m.n_split;
      k.fdl = nullptr;
    } else {
      k.ynow = nullptr; k.n_split_now = 0;
      k.fdl = b->fdl;
    }
    k.Hd = b->Hd; k.fmap = b->fmap; k.c_x = b->c_x; k.c_f = c.filter_channels; k.P = P; k.head = b->head;
    k.y = y_dev; k.ys = mix ? 0 : yl.stream; k.yc = yl.chan; k.yi = yl.samp; k.y_off = pos;
    k.c_out = c.c_out; k.B = B; k.fill = b->fill; k.take = take; k.tw = b->tw;
    if (completes && P > 1) {
      // the block commits with this step: its row is final once K1 has run, so the next block's past
      // pass can start now and overlap this step's K2 (and the next step's K1)
      b->head = (b->head + 1) % P;
      b->par ^= 1;
      launch_past(b, mix, st);
      b->head = (b->head + P - 1) % P;
      b->par ^= 1;
    }
    {
      ProfScope ps(b, st, 2);
      pgx::launch_c2r_emit(k, st);
    }
    b->launches += 1;
    b->steps += 1;
    if (b->profiling) b->prof_steps += 1;
    b->fill += take;
    pos += take;
    if (completes) {  // block complete: commit the row, advance the ring
      b->head = (b->head + 1) % P;
      b->half ^= 1;
      b->fill = 0;
      if (P > 1) b->par ^= 1;  // past sum of the new open block was launched above
    }
  }
  PGX_CUDA(cudaGetLastError());
  return PGX_OK;
}

// state changes outside run_pull invalidate the cached past sum and must not race the background pass
void quiesce_background(pgx_bank* b) {
  if (b->ev_past_recorded) cudaStreamWaitEvent(b->stream, b->ev_past, 0);
  b->past_valid = false;
}

int check_pull_args(pgx_bank* b, const void* x, const void* y, int n) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (!x || !y) return fail(PGX_ERR_INVALID, "x / y must not be NULL");
  if (n < 1 || n > b->cfg.max_pull)
    return fail(PGX_ERR_INVALID, "pull of %d samples outside [1, max_pull=%d]", n, b->cfg.max_pull);
  return PGX_OK;
}

}  // namespace

extern "C" {

int pgx_abi_version(void) { return PGX_ABI_VERSION; }

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

int pgx_host_free(void* ptr) {
  if (ptr) PGX_CUDA(cudaFreeHost(ptr));
  return PGX_OK;
}

int pgx_bank_create(pgx_bank** out, const pgx_bank_config* cfg, const float* h, const int32_t* filter_of_stream) {
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
  const int B = b->B, P = b->P;
  const size_t n_fft = (size_t)c.n_streams * c_x;
  const size_t h_rows = (size_t)c.n_filters * c.filter_channels;

  b->hist_bytes = n_fft * 2 * B * sizeof(float);
  b->fdl_bytes = n_fft * P * B * sizeof(float2);
  b->Hd_bytes = h_rows * 2 * P * B * sizeof(float2);
  {
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess && prop.multiProcessorCount > 0)
      b->sm_count = prop.multiProcessorCount;
  }
  size_t y_conv = 0, y_mix = 0;
  if (P > 1) {  // background pass over the P-1 committed partitions
    b->plan_conv = pgx::mac_plan(c.n_streams, c.c_out, B / 2, P - 1, false, c.n_filters == 1, b->sm_count);
    b->plan_mix = pgx::mac_plan(c.n_streams, c.c_out, B / 2, c.n_streams * (P - 1), true, c.n_filters == 1, b->sm_count);
    y_conv = (size_t)b->plan_conv.n_split * c.n_streams * c.c_out;
    y_mix = (size_t)b->plan_mix.n_split * c.c_out;
  }
  b->plan_now = pgx::mac_plan(c.n_streams, c.c_out, B / 2, c.n_streams, true, c.n_filters == 1, b->sm_count);
  b->yspec_bytes = (y_conv > y_mix ? y_conv : y_mix) * B * sizeof(float2);
  if (b->yspec_bytes == 0) b->yspec_bytes = sizeof(float2);
  const size_t ynow_bytes = (size_t)b->plan_now.n_split * c.c_out * B * sizeof(float2);
  b->xs_bytes = (size_t)c.n_streams * c.c_in * c.max_pull * sizeof(float);
  b->ys_bytes = (size_t)c.n_streams * c.c_out * c.max_pull * sizeof(float);

  int rc = PGX_OK;
  auto guard = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == PGX_OK)
      rc = fail(e == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  };
  {
    int lo = 0, hi = 0;  // the critical path (K1/K2) outranks the background pass when both want an SM
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    guard(cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, hi), "cudaStreamCreate");
    guard(cudaStreamCreateWithPriority(&b->bg, cudaStreamNonBlocking, lo), "cudaStreamCreate(bg)");
  }
  guard(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
  guard(cudaEventCreateWithFlags(&b->ev_past, cudaEventDisableTiming), "cudaEventCreate");
  guard(cudaMalloc(&b->hist, b->hist_bytes), "cudaMalloc(hist)");
  guard(cudaMalloc(&b->fdl, b->fdl_bytes), "cudaMalloc(fdl)");
  guard(cudaMalloc(&b->Hd, b->Hd_bytes), "cudaMalloc(Hd)");
  guard(cudaMalloc(&b->ypast[0], b->yspec_bytes), "cudaMalloc(ypast0)");
  guard(cudaMalloc(&b->ypast[1], b->yspec_bytes), "cudaMalloc(ypast1)");
  guard(cudaMalloc(&b->ynow, ynow_bytes), "cudaMalloc(ynow)");
  guard(cudaMalloc(&b->tw, (size_t)2 * B * sizeof(float2)), "cudaMalloc(tw)");
  guard(cudaMalloc(&b->fmap, (size_t)c.n_streams * sizeof(int32_t)), "cudaMalloc(fmap)");
  guard(cudaMalloc(&b->x_stage, b->xs_bytes), "cudaMalloc(x_stage)");
  guard(cudaMalloc(&b->y_stage, b->ys_bytes), "cudaMalloc(y_stage)");
  guard(cudaHostAlloc(&b->fmap_pinned, (size_t)c.n_streams * sizeof(int32_t), cudaHostAllocDefault), "cudaHostAlloc");
  if (rc != PGX_OK) {
    free_bank(b);
    return rc;
  }

  // twiddles in double, stored float: tw[k] = exp(-2*pi*i*k/2B)
  {
    std::vector<float2> tw((size_t)2 * B);
    const double w = -2.0 * M_PI / (2.0 * B);
    for (int k = 0; k < 2 * B; ++k) tw[k] = make_float2((float)cos(w * k), (float)sin(w * k));
    guard(cudaMemcpyAsync(b->tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, b->stream), "H2D tw");
    guard(cudaStreamSynchronize(b->stream), "sync tw");
  }
  for (int s = 0; s < c.n_streams; ++s) b->fmap_pinned[s] = filter_of_stream ? filter_of_stream[s] : (s % c.n_filters);
  guard(cudaMemcpyAsync(b->fmap, b->fmap_pinned, (size_t)c.n_streams * sizeof(int32_t), cudaMemcpyHostToDevice, b->stream),
        "H2D fmap");

  // filter spectra (replaces the one-time np.fft.rfft(h, n=nfft), convolve_pe.py:236-239)
  {
    float* h_dev = nullptr;
    const size_t hb = h_rows * (size_t)c.filter_len * sizeof(float);
    guard(cudaMalloc(&h_dev, hb), "cudaMalloc(h)");
    if (rc == PGX_OK) {
      guard(cudaMemcpyAsync(h_dev, h, hb, cudaMemcpyHostToDevice, b->stream), "H2D h");
      pgx::FilterPrepArgs fp{};
      fp.h = h_dev; fp.Hd = b->Hd; fp.tw = b->tw; fp.n_rows = (int)h_rows; fp.L = c.filter_len; fp.B = B; fp.P = P;
      pgx::launch_filter_prep(fp, b->stream);
      b->launches += 1;
      guard(cudaGetLastError(), "k_filter_prep launch");
      guard(cudaStreamSynchronize(b->stream), "k_filter_prep");
    }
    cudaFree(h_dev);
  }
  guard(cudaMemsetAsync(b->hist, 0, b->hist_bytes, b->stream), "memset hist");
  guard(cudaMemsetAsync(b->fdl, 0, b->fdl_bytes, b->stream), "memset fdl");
  guard(cudaStreamSynchronize(b->stream), "sync init");
  if (rc != PGX_OK) {
    free_bank(b);
    return rc;
  }
  *out = b;
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
  info->filter_len = c.filter_len; info->filter_channels = c.filter_channels; info->n_filters = c.n_filters;
  info->block = b->B; info->partitions = b->P; info->max_pull = c.max_pull; info->device = c.device;
  info->head = b->head; info->fill = b->fill;
  info->state_bytes = (int64_t)(b->hist_bytes + b->fdl_bytes + b->Hd_bytes + 2 * b->yspec_bytes + b->xs_bytes + b->ys_bytes);
  info->kernel_launches = b->launches;
  info->block_steps = b->steps;
  info->mac_grid = b->plan_conv.grid; info->mac_split = b->plan_conv.n_split;
  info->mac_stream_tile = b->plan_conv.st; info->mac_occupancy = b->plan_conv.occupancy;
  return PGX_OK;
}

int pgx_bank_reset(pgx_bank* b, const int32_t* stream_ids, int32_t k) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  quiesce_background(b);
  if (k <= 0 || !stream_ids) {
    PGX_CUDA(cudaMemsetAsync(b->hist, 0, b->hist_bytes, b->stream));
    PGX_CUDA(cudaMemsetAsync(b->fdl, 0, b->fdl_bytes, b->stream));
    b->head = b->fill = b->half = 0;
    b->par = 0;
    return PGX_OK;
  }
  const size_t hs = (size_t)b->c_x * 2 * b->B * sizeof(float);
  const size_t fs = (size_t)b->c_x * b->P * b->B * sizeof(float2);
  for (int i = 0; i < k; ++i) {
    const int s = stream_ids[i];
    if (s < 0 || s >= b->cfg.n_streams) return fail(PGX_ERR_INVALID, "stream id %d outside [0,%d)", s, b->cfg.n_streams);
    PGX_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(b->hist) + s * hs, 0, hs, b->stream));
    PGX_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(b->fdl) + s * fs, 0, fs, b->stream));
  }
  return PGX_OK;
}

int pgx_bank_load_filter(pgx_bank* b, int32_t filter_index, const float* h) {
  if (!b || !h) return fail(PGX_ERR_INVALID, "NULL argument");
  const pgx_bank_config& c = b->cfg;
  if (filter_index < 0 || filter_index >= c.n_filters)
    return fail(PGX_ERR_INVALID, "filter_index %d outside [0,%d)", filter_index, c.n_filters);
  PGX_CUDA(cudaSetDevice(c.device));
  quiesce_background(b);
  float* h_dev = nullptr;
  const size_t hb = (size_t)c.filter_channels * c.filter_len * sizeof(float);
  PGX_CUDA(cudaMalloc(&h_dev, hb));
  cudaError_t e = cudaMemcpyAsync(h_dev, h, hb, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) {
    pgx::FilterPrepArgs fp{};
    fp.h = h_dev;
    fp.Hd = b->Hd + (size_t)filter_index * c.filter_channels * 2 * b->P * b->B;
    fp.tw = b->tw; fp.n_rows = c.filter_channels; fp.L = c.filter_len; fp.B = b->B; fp.P = b->P;
    pgx::launch_filter_prep(fp, b->stream);
    b->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(h_dev);
  if (e != cudaSuccess) return fail(PGX_ERR_CUDA, "pgx_bank_load_filter: %s", cudaGetErrorString(e));
  return PGX_OK;
}

int pgx_bank_set_filter_map(pgx_bank* b, const int32_t* filter_of_stream) {
  if (!b || !filter_of_stream) return fail(PGX_ERR_INVALID, "NULL argument");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  for (int s = 0; s < b->cfg.n_streams; ++s)
    if (filter_of_stream[s] < 0 || filter_of_stream[s] >= b->cfg.n_filters)
      return fail(PGX_ERR_INVALID, "filter_of_stream[%d]=%d outside [0,%d)", s, filter_of_stream[s], b->cfg.n_filters);
  quiesce_background(b);
  PGX_CUDA(cudaStreamSynchronize(b->stream));  // the pinned staging copy may still be in flight
  memcpy(b->fmap_pinned, filter_of_stream, (size_t)b->cfg.n_streams * sizeof(int32_t));
  PGX_CUDA(cudaMemcpyAsync(b->fmap, b->fmap_pinned, (size_t)b->cfg.n_streams * sizeof(int32_t), cudaMemcpyHostToDevice,
                           b->stream));
  return PGX_OK;
}

static int process_host(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n, bool mix) {
  int rc = check_pull_args(b, x, y, n);
  if (rc != PGX_OK) return rc;
  const pgx_bank_config& c = b->cfg;
  if (!layout_dense(xl, c.n_streams, c.c_in, n)) return fail(PGX_ERR_INVALID, "x layout does not tile a dense block");
  pgx_layout yd = yl;
  if (mix) yd.stream = 0;
  if (!layout_dense(yd, mix ? 1 : c.n_streams, c.c_out, n))
    return fail(PGX_ERR_INVALID, "y layout does not tile a dense block");
  PGX_CUDA(cudaSetDevice(c.device));
  const size_t xb = (size_t)c.n_streams * c.c_in * n * sizeof(float);
  const size_t yb = (size_t)(mix ? 1 : c.n_streams) * c.c_out * n * sizeof(float);
  PGX_CUDA(cudaMemcpyAsync(b->x_stage, x, xb, cudaMemcpyHostToDevice, b->stream));
  rc = run_pull(b, b->x_stage, xl, b->y_stage, yd, n, mix, b->stream);
  if (rc != PGX_OK) return rc;
  PGX_CUDA(cudaMemcpyAsync(y, b->y_stage, yb, cudaMemcpyDeviceToHost, b->stream));
  PGX_CUDA(cudaStreamSynchronize(b->stream));
  return PGX_OK;
}

int pgx_bank_process(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n) {
  return process_host(b, x, xl, y, yl, n, false);
}

int pgx_bank_process_mix(pgx_bank* b, const float* x, pgx_layout xl, float* y, pgx_layout yl, int32_t n) {
  return process_host(b, x, xl, y, yl, n, true);
}

int pgx_bank_process_device(pgx_bank* b, const float* x_dev, pgx_layout xl, float* y_dev, pgx_layout yl, int32_t n,
                            int32_t mix, void* cuda_stream) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  if (!x_dev || !y_dev) return fail(PGX_ERR_INVALID, "x / y must not be NULL");
  if (n < 1) return fail(PGX_ERR_INVALID, "pull of %d samples", n);
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : b->stream;
  pgx_layout yd = yl;
  if (mix) yd.stream = 0;
  return run_pull(b, x_dev, xl, y_dev, yd, n, mix != 0, st);
}

int pgx_bank_synchronize(pgx_bank* b) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
  PGX_CUDA(cudaSetDevice(b->cfg.device));
  PGX_CUDA(cudaStreamSynchronize(b->stream));
  PGX_CUDA(cudaStreamSynchronize(b->bg));
  return PGX_OK;
}

int pgx_bank_profile_begin(pgx_bank* b) {
  if (!b) return fail(PGX_ERR_INVALID, "bank is NULL");
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
  out->ms_r2c = out->ms_mac = out->ms_c2r = 0.0;
  out->steps = b->prof_steps;
  PGX_CUDA(cudaDeviceSynchronize());
  for (const pgx_bank::ProfSpan& sp : b->prof_spans) {
    float ms = 0.f;
    PGX_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
    if (sp.kind == 0) out->ms_r2c += ms;
    else if (sp.kind == 1) out->ms_mac += ms;
    else out->ms_c2r += ms;
  }
  b->prof_spans.clear();
  b->prof_pool_used = 0;
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
