// pgx_comm.cu -- the cross-GPU MixPE sum: per-rank partial mixes summed onto a root over NVLink peer memory.
//
// The only exchange step of the path: streams are sharded over the GPUs of one box, every rank's
// accumulate kernel produces its partial (C_out, n) mix, and the partials are added (reference mix_pe.py:92-94 --
// the same float32 "+=", here over ranks, in rank order).  The message is 2-8 KB per pull, i.e. pure latency, so
// instead of a library collective the ranks talk through each other's memory:
//
//   every rank owns a MAILBOX in its HBM (mapped into the peers with CUDA IPC, or plain peer access inside one
//   process):   flags [S][W] u32 | ack u32 | data [S][W][max_floats] f32        (S = 16 slots, W = world size)
//
//   pull q, non-root rank r : k_comm_push   -- (bounded wait: the root has consumed pull q-S) store the partial
//                             into the ROOT's data[q%S][r] over NVLink, fence, then release-store flags[q%S][r] = q+1
//   pull q, root            : k_comm_gather -- acquire-spin on flags[q%S][*] (bounded), sum own partial + the W-1
//                             slots in rank order, write y, release-store ack = q+1 into every peer's mailbox
//
// Both kernels are one small CTA enqueued on the bank's stream right behind the pull's output stage, so the
// exchange of pull q overlaps the background pass of pull q+1 on every rank; nothing synchronises with the host.
// Deterministic: the sum order is the rank order whatever the arrival order.  A spin that runs out (a peer died)
// raises the comm's error word, which the next pgx_bank_wait / pgx_comm_check reports -- it never hangs the GPU.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <new>

#include "host_util.h"
#include "kernels.h"

namespace {

constexpr int kSlots = 16;
constexpr int kMaxWorld = 16;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s

struct WireHandle {          // what pgx_comm_create hands out (PGX_COMM_HANDLE_BYTES = 128)
  cudaIpcMemHandle_t ipc;    // 64 bytes
  int64_t pid;
  uint64_t ptr;              // raw device pointer: usable by ranks living in the same process
  int32_t device, rank, world, max_floats;
  uint64_t bytes;
  char pad[128 - 64 - 8 - 8 - 16 - 8];
};
static_assert(sizeof(WireHandle) == 128, "wire handle is 128 bytes");

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want; false when the bounded wait ran out
__device__ __forceinline__ bool spin_ge(const uint32_t* p, uint32_t want) {
  if (ld_acquire_sys(p) >= want) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
    for (int i = 0; i < 256; ++i)
      if (ld_acquire_sys(p) >= want) return true;
    if (global_ns() - t0 > kSpinTimeoutNs) return false;
    __nanosleep(64);
  }
}

// non-root: partial -> the root's slot, then the flag
__global__ void __launch_bounds__(256) k_comm_push(const float* __restrict__ part, float* __restrict__ remote_slot,
                                                   uint32_t* remote_flag, const uint32_t* local_ack, uint32_t q,
                                                   int n, uint32_t* err) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    if (q >= (uint32_t)kSlots && !spin_ge(local_ack, q - kSlots + 1)) {  // slot q%S still holds pull q-S
      ok = 0;
      *err = 1u;
    }
  }
  __syncthreads();
  if (!ok) return;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(remote_slot)) & 15) == 0) {
    const float4* s = reinterpret_cast<const float4*>(part);
    float4* d = reinterpret_cast<float4*>(remote_slot);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) remote_slot[i] = part[i];
  }
  __threadfence_system();  // this thread's stores are performed system-wide before ...
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(remote_flag, q + 1);  // ... the flag becomes visible to the root
}

struct GatherArgs {
  const float* part;            // the root's own partial
  float* y;                     // reduced mix (may alias part)
  const float* data;            // root mailbox: slot base, [W][max_floats]
  const uint32_t* flags;        // root mailbox: flags of this slot, [W]
  uint32_t* peer_ack[kMaxWorld];  // ack word inside every peer's mailbox (NULL for the root itself)
  uint32_t q;
  int32_t n, world, root, max_floats;
  uint32_t* err;
};

__global__ void __launch_bounds__(256) k_comm_gather(const GatherArgs a) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < a.world && threadIdx.x != a.root) {
    if (!spin_ge(a.flags + threadIdx.x, a.q + 1)) {
      ok = 0;
      *a.err = 1u;
    }
  }
  __syncthreads();
  if (ok) {
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
      float acc = 0.f;
      for (int r = 0; r < a.world; ++r) {  // rank order: ((p0 + p1) + p2) + ... in float32 like MixPE's "+="
        const float v = (r == a.root) ? a.part[i] : __ldcg(a.data + (size_t)r * a.max_floats + i);
        acc = (r == 0) ? v : __fadd_rn(acc, v);
      }
      a.y[i] = acc;
    }
  }
  __syncthreads();  // every read of the slot is done: the peers may reuse it
  if (threadIdx.x < a.world && threadIdx.x != a.root) st_release_sys(a.peer_ack[threadIdx.x], a.q + 1);
}

}  // namespace

struct pgx_comm {
  int device = 0, rank = 0, world = 1, root = 0, max_floats = 0;
  char* local = nullptr;
  size_t bytes = 0, off_ack = 0, off_data = 0;
  char* peer[kMaxWorld] = {};
  bool ipc_opened[kMaxWorld] = {};
  bool connected = false;
  uint32_t* err_host = nullptr;  // mapped pinned
  uint32_t* err_dev = nullptr;
  uint32_t q = 0;                // pulls reduced so far
};

namespace {
uint32_t* flags_of(char* base) { return reinterpret_cast<uint32_t*>(base); }
uint32_t* ack_of(const pgx_comm* c, char* base) { return reinterpret_cast<uint32_t*>(base + c->off_ack); }
float* data_of(const pgx_comm* c, char* base) { return reinterpret_cast<float*>(base + c->off_data); }
}  // namespace

// Enqueue the reduce of one partial on `st` (used by the bank and by pgx_mix_reduce).
int pgx_comm_enqueue(pgx_comm* c, const float* part_dev, float* y_dev, int32_t n, cudaStream_t st) {
  if (!c->connected) return pgx_fail(PGX_ERR_INVALID, "pgx_comm: connect the communicator first");
  if (n < 1 || n > c->max_floats) return pgx_fail(PGX_ERR_INVALID, "mix reduce of %d floats outside [1, %d]", n, c->max_floats);
  const uint32_t q = c->q;
  const int slot = (int)(q % kSlots);
  if (c->rank != c->root) {
    char* rb = c->peer[c->root];
    k_comm_push<<<1, 256, 0, st>>>(part_dev, data_of(c, rb) + ((size_t)slot * c->world + c->rank) * c->max_floats,
                                   flags_of(rb) + slot * c->world + c->rank, ack_of(c, c->local), q, n, c->err_dev);
  } else {
    GatherArgs a{};
    a.part = part_dev; a.y = y_dev;
    a.data = data_of(c, c->local) + (size_t)slot * c->world * c->max_floats;
    a.flags = flags_of(c->local) + slot * c->world;
    for (int r = 0; r < c->world; ++r) a.peer_ack[r] = (r == c->rank) ? nullptr : ack_of(c, c->peer[r]);
    a.q = q; a.n = n; a.world = c->world; a.root = c->root; a.max_floats = c->max_floats; a.err = c->err_dev;
    k_comm_gather<<<1, 256, 0, st>>>(a);
  }
  c->q = q + 1;
  PGX_CUDA(cudaGetLastError());
  return PGX_OK;
}

int pgx_comm_is_root(const pgx_comm* c) { return c->rank == c->root; }
int pgx_comm_max_floats(const pgx_comm* c) { return c->max_floats; }
int pgx_comm_device(const pgx_comm* c) { return c->device; }

extern "C" {

int pgx_comm_create(pgx_comm** out, int32_t device, int32_t rank, int32_t world, int32_t root, int32_t max_floats,
                    void* handle_out) {
  if (!out || !handle_out) return pgx_fail(PGX_ERR_INVALID, "pgx_comm_create: NULL argument");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld) return pgx_fail(PGX_ERR_INVALID, "world size %d outside [1, %d]", world, kMaxWorld);
  if (rank < 0 || rank >= world || root < 0 || root >= world)
    return pgx_fail(PGX_ERR_INVALID, "rank %d / root %d outside [0, %d)", rank, root, world);
  if (max_floats < 1 || max_floats > (1 << 22)) return pgx_fail(PGX_ERR_INVALID, "max_floats %d outside [1, 2^22]", max_floats);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return pgx_fail(PGX_ERR_NO_DEVICE, "no CUDA device: libpgx has no CPU fallback");
  if (device < 0 || device >= ndev) return pgx_fail(PGX_ERR_INVALID, "device %d outside [0,%d)", device, ndev);
  PGX_CUDA(cudaSetDevice(device));
  pgx_comm* c = new (std::nothrow) pgx_comm();
  if (!c) return pgx_fail(PGX_ERR_NOMEM, "out of host memory");
  c->device = device; c->rank = rank; c->world = world; c->root = root;
  c->max_floats = (max_floats + 3) & ~3;  // slots stay 16-byte aligned
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  c->off_ack = up((size_t)kSlots * world * sizeof(uint32_t));
  c->off_data = c->off_ack + 256;
  c->bytes = c->off_data + (size_t)kSlots * world * c->max_floats * sizeof(float);
  cudaError_t e = cudaMalloc(&c->local, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->err_host), sizeof(uint32_t), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *c->err_host = 0u;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_dev), c->err_host, 0);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  WireHandle w{};
  if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&w.ipc, c->local);
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    if (c->err_host) cudaFreeHost(c->err_host);
    delete c;
    return pgx_fail(e == cudaErrorMemoryAllocation ? PGX_ERR_NOMEM : PGX_ERR_CUDA, "pgx_comm_create: %s", cudaGetErrorString(e));
  }
  w.pid = (int64_t)getpid();
  w.ptr = reinterpret_cast<uint64_t>(c->local);
  w.device = device; w.rank = rank; w.world = world; w.max_floats = c->max_floats; w.bytes = c->bytes;
  memcpy(handle_out, &w, sizeof w);
  c->peer[rank] = c->local;
  if (world == 1) c->connected = true;
  *out = c;
  return PGX_OK;
}

int pgx_comm_connect(pgx_comm* c, const void* handles) {
  if (!c || !handles) return pgx_fail(PGX_ERR_INVALID, "pgx_comm_connect: NULL argument");
  PGX_CUDA(cudaSetDevice(c->device));
  const WireHandle* w = static_cast<const WireHandle*>(handles);
  for (int r = 0; r < c->world; ++r) {
    if (w[r].rank != r || w[r].world != c->world || w[r].max_floats != c->max_floats || w[r].bytes != c->bytes)
      return pgx_fail(PGX_ERR_INVALID, "pgx_comm_connect: handle %d does not describe rank %d of this communicator", r, r);
    if (r == c->rank) continue;
    if (c->rank != c->root && r != c->root) continue;  // a non-root rank only ever touches the root's mailbox
    if (c->peer[r]) continue;
    if (w[r].pid == (int64_t)getpid()) {  // same process: plain peer access
      if (w[r].device != c->device) {
        int can = 0;
        PGX_CUDA(cudaDeviceCanAccessPeer(&can, c->device, w[r].device));
        if (!can) return pgx_fail(PGX_ERR_CUDA, "device %d cannot access device %d", c->device, w[r].device);
        cudaError_t e = cudaDeviceEnablePeerAccess(w[r].device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        PGX_CUDA(e);
      }
      c->peer[r] = reinterpret_cast<char*>(w[r].ptr);
    } else {
      void* p = nullptr;
      PGX_CUDA(cudaIpcOpenMemHandle(&p, w[r].ipc, cudaIpcMemLazyEnablePeerAccess));
      c->peer[r] = static_cast<char*>(p);
      c->ipc_opened[r] = true;
    }
  }
  c->connected = true;
  return PGX_OK;
}

int pgx_comm_check(pgx_comm* c) {
  if (!c) return pgx_fail(PGX_ERR_INVALID, "comm is NULL");
  if (c->err_host && *reinterpret_cast<volatile uint32_t*>(c->err_host))
    return pgx_fail(PGX_ERR_CUDA, "mix reduce: rank %d waited %d s for a peer that never arrived", c->rank,
                    (int)(kSpinTimeoutNs / 1000000000ull));
  return PGX_OK;
}

int pgx_comm_destroy(pgx_comm* c) {
  if (!c) return PGX_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->local);
  if (c->err_host) cudaFreeHost(c->err_host);
  delete c;
  return PGX_OK;
}

int pgx_mix_reduce(pgx_comm* c, const float* part_dev, float* y_dev, int32_t n, void* cuda_stream) {
  if (!c || !part_dev) return pgx_fail(PGX_ERR_INVALID, "pgx_mix_reduce: NULL argument");
  if (c->rank == c->root && !y_dev) return pgx_fail(PGX_ERR_INVALID, "pgx_mix_reduce: the root needs an output buffer");
  PGX_CUDA(cudaSetDevice(c->device));
  if (c->world == 1) {  // nothing to add: y = part
    if (y_dev != part_dev)
      PGX_CUDA(cudaMemcpyAsync(y_dev, part_dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(cuda_stream)));
    return PGX_OK;
  }
  return pgx_comm_enqueue(c, part_dev, y_dev, n, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
