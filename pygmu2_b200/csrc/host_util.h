// host_util.h -- error plumbing shared by the host-side translation units of libpgx.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/pgx.h"

// Sets the calling thread's pgx_last_error() message (printf-style) and returns `code`.
int pgx_fail(int code, const char* fmt, ...);

#define PGX_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return pgx_fail(e__ == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                      cudaGetErrorString(e__), __FILE__, __LINE__);                                \
  } while (0)

// pgx_comm.cu: the cross-GPU mix reduce, enqueued on `st` (no host synchronisation); y_dev may alias part_dev
struct pgx_comm;
int pgx_comm_enqueue(pgx_comm* c, const float* part_dev, float* y_dev, int32_t n, cudaStream_t st);
int pgx_comm_is_root(const pgx_comm* c);
int pgx_comm_max_floats(const pgx_comm* c);
int pgx_comm_device(const pgx_comm* c);
