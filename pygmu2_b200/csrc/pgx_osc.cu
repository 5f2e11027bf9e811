// pgx_osc.cu -- C ABI of the device-resident sources (include/pgx.h "sources" section): constant-parameter
// SinePE banks and BlitSawPE / SuperSawPE voice banks that render straight into device memory, so the
// convolution path's inputs never exist on the host.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "host_util.h"
#include "kernels.h"

struct pgx_osc {
  pgx_osc_config cfg{};
  int n_osc = 0;
  cudaStream_t stream = nullptr;
  double *params = nullptr;                       // sine: [V][3]
  double *consts = nullptr, *gain = nullptr, *amp = nullptr, *phase_init = nullptr;
  double *st_phase = nullptr, *st_int = nullptr;  // blit state
  float *out = nullptr, *mix = nullptr;           // [V][C][max_pull], [C][max_pull]
  double *mod_state = nullptr, *mod_scratch = nullptr;  // modulated sine: accumulated phase [V], phases [V][max_pull]
  float *ctl[3] = {nullptr, nullptr, nullptr};    // staged control vectors (host-supplied), [V][max_pull] each
  bool mod_started = false;
  double *snap_phase = nullptr, *snap_int = nullptr;   // state before a speculative pull (PGX_OSC_SNAPSHOT)
  int64_t snap_last_end = INT64_MIN;
  bool snap_has_last = false, snap_valid = false;
  double* osc_freq = nullptr;                     // BLIT: per-oscillator frequency as given (the detune ratio when modulated)
  int32_t* m_fixed_dev = nullptr;                 // BLIT: fixed harmonic counts (0 = auto)
  int64_t last_end = INT64_MIN;
  bool has_last = false;
  int64_t launches = 0;
};

namespace {

void free_osc(pgx_osc* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : {(void*)h->params, (void*)h->consts, (void*)h->gain, (void*)h->amp, (void*)h->phase_init,
                  (void*)h->st_phase, (void*)h->st_int, (void*)h->out, (void*)h->mix, (void*)h->mod_state,
                  (void*)h->mod_scratch, (void*)h->ctl[0], (void*)h->ctl[1], (void*)h->ctl[2], (void*)h->osc_freq,
                  (void*)h->m_fixed_dev, (void*)h->snap_phase, (void*)h->snap_int})
    cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

template <typename T>
cudaError_t upload(T** dst, const T* src, size_t n, cudaStream_t st) {
  cudaError_t e = cudaMalloc(dst, n * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

int reset_state(pgx_osc* h, cudaStream_t st) {
  if (h->cfg.kind != PGX_OSC_BLIT) return PGX_OK;
  PGX_CUDA(cudaMemcpyAsync(h->st_phase, h->phase_init, (size_t)h->n_osc * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PGX_CUDA(cudaMemsetAsync(h->st_int, 0, (size_t)h->n_osc * sizeof(double), st));
  return PGX_OK;
}

// Enqueue one pull on `st`; *out_dev = planar [V][C][n] or, mixed, [C][n].
int render_on(pgx_osc* h, int64_t start, int32_t n, int32_t flags, cudaStream_t st, const float** out_dev) {
  const pgx_osc_config& c = h->cfg;
  if (n < 1 || n > c.max_pull) return pgx_fail(PGX_ERR_INVALID, "pull of %d samples outside [1, max_pull=%d]", n, c.max_pull);
  const int C = c.channels;
  if (flags & PGX_OSC_SNAPSHOT) {  // a speculative pull: remember the state it starts from (pgx_osc_rollback undoes it)
    if (c.kind == PGX_OSC_BLIT) {
      const size_t nb = (size_t)h->n_osc * sizeof(double);
      if (!h->snap_phase) {
        PGX_CUDA(cudaMalloc(&h->snap_phase, nb));
        PGX_CUDA(cudaMalloc(&h->snap_int, nb));
      }
      if (!h->has_last || start != h->last_end) {   // the pull itself would start a new run: nothing to preserve but that
        const int rc = reset_state(h, st);
        if (rc != PGX_OK) return rc;
        h->has_last = true;
        h->last_end = start;
      }
      // (the kernel itself copies the state it starts from into the snapshot arrays: no extra calls on this path)
    }
    h->snap_last_end = h->last_end;
    h->snap_has_last = h->has_last;
    h->snap_valid = true;
  } else {
    h->snap_valid = false;           // a regular pull commits whatever came before it
  }
  if (c.kind == PGX_OSC_SINE) {
    pgx::SineArgs a{};
    a.params = h->params; a.out = h->out; a.os = (int64_t)C * n; a.oc = n; a.oi = 1;
    a.start = start; a.n_streams = c.n_voices; a.channels = C; a.n = n; a.sample_rate = c.sample_rate;
    pgx::launch_sine_bank(a, st);
  } else {
    if (!h->has_last || start != h->last_end) {  // discontinuous pull: state := initial (blit_saw_pe.py:183-186)
      const int rc = reset_state(h, st);
      if (rc != PGX_OK) return rc;
    }
    pgx::BlitArgs a{};
    a.consts = h->consts; a.gain = h->gain; a.amp = h->amp;
    a.st_phase = h->st_phase; a.st_int = h->st_int; a.out = h->out;
    a.os = (int64_t)C * n; a.oc = n; a.oi = 1; a.channels = C;
    a.leak = c.leak; a.n_voices = c.n_voices; a.unison = c.unison; a.n = n; a.sample_rate = c.sample_rate;
    if (flags & PGX_OSC_SNAPSHOT) { a.snap_phase = h->snap_phase; a.snap_int = h->snap_int; }
    pgx::launch_blit_bank(a, st);
  }
  h->launches += 1;
  h->last_end = start + n;
  h->has_last = true;
  if (flags & PGX_PULL_MIX) {  // MixPE over the voices: float32, left to right (mix_pe.py:92-94), bit-exact
    pgx::launch_mix_sum(h->out, c.n_voices, (int64_t)C * n, h->mix, st);
    h->launches += 1;
    *out_dev = h->mix;
  } else {
    *out_dev = h->out;
  }
  PGX_CUDA(cudaGetLastError());
  return PGX_OK;
}

}  // namespace

extern "C" {

int pgx_osc_create(pgx_osc** out, const pgx_osc_config* cfg, const double* freq, const double* gain,
                   const double* phase, const int32_t* m_fixed, const double* amp) {
  if (!out || !cfg || !freq || !gain || !phase) return pgx_fail(PGX_ERR_INVALID, "pgx_osc_create: NULL argument");
  *out = nullptr;
  const pgx_osc_config& c = *cfg;
  if (c.kind != PGX_OSC_SINE && c.kind != PGX_OSC_BLIT) return pgx_fail(PGX_ERR_INVALID, "unknown oscillator kind %d", c.kind);
  if (c.n_voices < 1 || c.channels < 1 || c.max_pull < 1 || c.sample_rate < 1)
    return pgx_fail(PGX_ERR_INVALID, "n_voices, channels, max_pull and sample_rate must be >= 1");
  const int U = c.kind == PGX_OSC_BLIT ? c.unison : 1;
  if (U < 1 || U > 32) return pgx_fail(PGX_ERR_INVALID, "unison must be in [1, 32], got %d", U);
  if (c.kind == PGX_OSC_BLIT && !amp) return pgx_fail(PGX_ERR_INVALID, "voice amplitudes missing");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return pgx_fail(PGX_ERR_NO_DEVICE, "no CUDA device: libpgx has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) return pgx_fail(PGX_ERR_INVALID, "device %d outside [0,%d)", c.device, ndev);
  PGX_CUDA(cudaSetDevice(c.device));
  pgx_osc* h = new (std::nothrow) pgx_osc();
  if (!h) return pgx_fail(PGX_ERR_NOMEM, "out of host memory");
  h->cfg = c;
  h->cfg.unison = U;
  h->n_osc = c.n_voices * U;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  const size_t no = (size_t)h->n_osc;
  if (c.kind == PGX_OSC_SINE) {
    std::vector<double> p(3 * no);
    for (size_t s = 0; s < no; ++s) { p[3 * s] = freq[s]; p[3 * s + 1] = gain[s]; p[3 * s + 2] = phase[s]; }
    if (e == cudaSuccess) e = upload(&h->params, p.data(), p.size(), h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // p goes out of scope
  } else {
    // blit_saw_pe.py:167-177,189,198,218 in the reference's own float64 expressions
    std::vector<double> cst(4 * no);
    const double sr = (double)c.sample_rate;
    for (size_t o = 0; o < no; ++o) {
      const double f = freq[o], fm = f > 1.0 ? f : 1.0;
      const double P = sr / fm;
      double M;
      if (m_fixed && m_fixed[o] > 0) {
        M = (double)m_fixed[o];
      } else {
        int m = (int)floor(sr / (2.0 * fm));
        m = m - (1 - m % 2);
        if (m < 1) m = 1;
        M = (double)m;
      }
      cst[4 * o] = f / sr; cst[4 * o + 1] = P; cst[4 * o + 2] = 1.0 / P; cst[4 * o + 3] = M;
    }
    if (e == cudaSuccess) e = upload(&h->consts, cst.data(), cst.size(), h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // cst goes out of scope
    if (e == cudaSuccess) e = upload(&h->osc_freq, freq, no, h->stream);
    {
      std::vector<int32_t> mf(no, 0);
      if (m_fixed) for (size_t o = 0; o < no; ++o) mf[o] = m_fixed[o] > 0 ? m_fixed[o] : 0;
      if (e == cudaSuccess) e = upload(&h->m_fixed_dev, mf.data(), no, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // mf goes out of scope
    }
    if (e == cudaSuccess) e = upload(&h->gain, gain, no, h->stream);
    if (e == cudaSuccess) e = upload(&h->phase_init, phase, no, h->stream);
    if (e == cudaSuccess) e = upload(&h->amp, amp, (size_t)c.n_voices, h->stream);
    if (e == cudaSuccess) e = cudaMalloc(&h->st_phase, no * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->st_int, no * sizeof(double));
  }
  const size_t ob = (size_t)c.n_voices * c.channels * c.max_pull * sizeof(float);
  if (e == cudaSuccess) e = cudaMalloc(&h->out, ob);
  if (e == cudaSuccess) e = cudaMalloc(&h->mix, (size_t)c.channels * c.max_pull * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    free_osc(h);
    return pgx_fail(e == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "pgx_osc_create: %s", cudaGetErrorString(e));
  }
  if (reset_state(h, h->stream) != PGX_OK || cudaStreamSynchronize(h->stream) != cudaSuccess) {
    free_osc(h);
    return pgx_fail(PGX_ERR_CUDA, "pgx_osc_create: state initialisation failed");
  }
  *out = h;
  return PGX_OK;
}

int pgx_osc_destroy(pgx_osc* h) {
  free_osc(h);
  return PGX_OK;
}

int pgx_osc_render_modulated(pgx_osc* h, int32_t n, int32_t flags, const float* freq, const float* amp,
                             const float* phase, int32_t ctl_flags, void* cuda_stream, const float** out_dev,
                             float* y_host) {
  if (!h || (!out_dev && !y_host)) return pgx_fail(PGX_ERR_INVALID, "NULL argument");
  const pgx_osc_config& c = h->cfg;
  if (n < 1 || n > c.max_pull) return pgx_fail(PGX_ERR_INVALID, "pull of %d samples outside [1, max_pull=%d]", n, c.max_pull);
  PGX_CUDA(cudaSetDevice(c.device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  const size_t V = (size_t)c.n_voices;
  if (c.kind == PGX_OSC_SINE && !h->mod_state) {
    PGX_CUDA(cudaMalloc(&h->mod_state, V * sizeof(double)));
    PGX_CUDA(cudaMalloc(&h->mod_scratch, V * c.max_pull * sizeof(double)));
  }
  const float* ctl[3] = {freq, amp, phase};
  if (ctl_flags & PGX_CTL_HOST) {  // what the parameter PEs rendered on the host: staged, in stream order
    for (int i = 0; i < 3; ++i) {
      if (!ctl[i]) continue;
      if (!h->ctl[i]) PGX_CUDA(cudaMalloc(&h->ctl[i], V * c.max_pull * sizeof(float)));
      PGX_CUDA(cudaMemcpyAsync(h->ctl[i], ctl[i], V * n * sizeof(float), cudaMemcpyHostToDevice, st));
      ctl[i] = h->ctl[i];
    }
  }
  if (c.kind == PGX_OSC_SINE) {
    pgx::SineModArgs a{};
    a.params = h->params; a.freq = ctl[0]; a.amp = ctl[1]; a.phase = ctl[2];
    a.state = h->mod_state; a.scratch = h->mod_scratch; a.out = h->out;
    a.os = (int64_t)c.channels * n; a.oc = n; a.oi = 1;
    a.n_voices = c.n_voices; a.channels = c.channels; a.n = n; a.sample_rate = c.sample_rate; a.max_pull = c.max_pull;
    a.first = h->mod_started ? 0 : 1;
    pgx::launch_sine_mod(a, st);
    h->mod_started = true;
  } else {
    // stateful like the constant-parameter path; a modulated BLIT PE is only ever pulled contiguously (the
    // reference's renderer enforces it for stateful PEs) and pgx_osc_reset starts over (blit_saw_pe.py:137-142)
    if (!h->has_last) {
      const int rc = reset_state(h, st);
      if (rc != PGX_OK) return rc;
    }
    pgx::BlitModArgs a{};
    a.osc_freq = h->osc_freq; a.gain = h->gain; a.vamp = h->amp; a.m_fixed = h->m_fixed_dev;
    a.freq = ctl[0]; a.amp = ctl[1]; a.m_ctl = ctl[2];
    a.st_phase = h->st_phase; a.st_int = h->st_int; a.out = h->out;
    a.os = (int64_t)c.channels * n; a.oc = n; a.oi = 1; a.leak = c.leak;
    a.n_voices = c.n_voices; a.unison = c.unison; a.channels = c.channels; a.n = n; a.sample_rate = c.sample_rate;
    a.amp_per_osc = (ctl_flags & PGX_CTL_AMP_OSC) ? 1 : 0;
    pgx::launch_blit_mod(a, st);
    h->has_last = true;
    h->last_end = (h->last_end == INT64_MIN ? 0 : h->last_end) + n;
  }
  h->launches += 1;
  const float* res = h->out;
  if (flags & PGX_PULL_MIX) {
    pgx::launch_mix_sum(h->out, c.n_voices, (int64_t)c.channels * n, h->mix, st);
    h->launches += 1;
    res = h->mix;
  }
  if (out_dev) *out_dev = res;
  PGX_CUDA(cudaGetLastError());
  if (y_host) {
    const size_t rows = (flags & PGX_PULL_MIX) ? 1 : V;
    PGX_CUDA(cudaMemcpyAsync(y_host, res, rows * c.channels * n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (y_host || (ctl_flags & PGX_CTL_HOST)) PGX_CUDA(cudaStreamSynchronize(st));  // y complete / the control arrays may go away
  return PGX_OK;
}

int pgx_osc_rollback(pgx_osc* h, void* cuda_stream) {
  if (!h) return pgx_fail(PGX_ERR_INVALID, "osc is NULL");
  if (!h->snap_valid) return PGX_OK;   // nothing speculative outstanding
  PGX_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  if (h->cfg.kind == PGX_OSC_BLIT) {
    const size_t nb = (size_t)h->n_osc * sizeof(double);
    PGX_CUDA(cudaMemcpyAsync(h->st_phase, h->snap_phase, nb, cudaMemcpyDeviceToDevice, st));
    PGX_CUDA(cudaMemcpyAsync(h->st_int, h->snap_int, nb, cudaMemcpyDeviceToDevice, st));
  }
  h->last_end = h->snap_last_end;
  h->has_last = h->snap_has_last;
  h->snap_valid = false;
  return PGX_OK;
}

int pgx_osc_reset(pgx_osc* h) {
  if (!h) return pgx_fail(PGX_ERR_INVALID, "osc is NULL");
  h->snap_valid = false;
  h->mod_started = false;  // modulated sine: the next pull starts from the constant phase again (sine_pe.py:110-118)
  h->has_last = false;  // the next pull re-initialises the state on its own stream, in order
  return PGX_OK;
}

int pgx_osc_render_device(pgx_osc* h, int64_t start, int32_t n, int32_t flags, void* cuda_stream, const float** out_dev) {
  if (!h || !out_dev) return pgx_fail(PGX_ERR_INVALID, "NULL argument");
  PGX_CUDA(cudaSetDevice(h->cfg.device));
  return render_on(h, start, n, flags, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream, out_dev);
}

int pgx_osc_render(pgx_osc* h, int64_t start, int32_t n, int32_t flags, float* y) {
  if (!h || !y) return pgx_fail(PGX_ERR_INVALID, "NULL argument");
  PGX_CUDA(cudaSetDevice(h->cfg.device));
  const float* src = nullptr;
  const int rc = render_on(h, start, n, flags, h->stream, &src);
  if (rc != PGX_OK) return rc;
  const size_t rows = (flags & PGX_PULL_MIX) ? 1 : (size_t)h->cfg.n_voices;
  PGX_CUDA(cudaMemcpyAsync(y, src, rows * h->cfg.channels * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  PGX_CUDA(cudaStreamSynchronize(h->stream));
  return PGX_OK;
}

int pgx_osc_launches(pgx_osc* h, int64_t* launches) {
  if (!h || !launches) return pgx_fail(PGX_ERR_INVALID, "NULL argument");
  *launches = h->launches;
  return PGX_OK;
}

}  // extern "C"
