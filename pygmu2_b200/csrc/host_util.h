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
