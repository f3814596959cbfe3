// p2p.cu -- one-shot all-reduce (mean) of the flat parameter gradient over NVLink peer memory.
//
// The path's only exchange step is the mean of the parameter gradient over the ranks (the reference
// leaves it to Lightning DDP -> NCCL).  The buffers are small (one layer's slice: 0.4 - 0.9 MB) and the last one of a step
// cannot overlap with anything, so what matters is latency, not bandwidth: every rank reads the slice of every peer
// straight out of the peer's HBM (CUDA IPC mappings, NVLink 5 / NVSwitch), sums in rank order -- all ranks get the same
// bits -- and writes the mean to its own output buffer.  One kernel, two flag exchanges:
//   ready[e]: my input slice for call e is complete  -> every peer may read it
//   done[e] : I have finished reading call e          -> every peer may overwrite its input in the next step
// Flags live in each rank's IPC buffer and are written by the PEERS (remote stores, release at system scope) and polled
// locally (acquire at system scope).  The call counter `epoch` lives on the device and is advanced by the kernel itself,
// so a captured CUDA graph replays correctly; every rank issues the same sequence of calls.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/gcpnet_b200.h"
#include "common.h"

namespace {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_BLOCKS = 32, P2P_THREADS = 512;  // 16 Ki threads x 16 bytes: a layer's slice (0.4 - 0.9 MB) in 2 - 4 rounds

struct P2PCtx {
  int rank, world;
  float* data[P2P_MAX_RANKS];                 // input buffers of all ranks (data[rank] = local)
  unsigned long long* flags[P2P_MAX_RANKS];   // flag blocks of all ranks: [ready: world][done: world]
  unsigned long long* epoch;                  // local: calls completed so far
  unsigned int* arrive;                       // local: block counter of the running call
  unsigned long long* go;                     // local: epoch whose inputs are ready (block 0 -> other blocks)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // peer HBM: never through a stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// WU = unroll bound of the rank loop (world <= WU): all peers' loads of one element are in flight together -- the loop is
// latency-bound (2 - 3 us per remote load), so their number in flight is what counts
template <int WU>
__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_mean_kernel(const P2PCtx c, long long offset, long long count, float* __restrict__ out) {
  const int tid = threadIdx.x;
  __shared__ unsigned long long e_sh;
  if (tid == 0) e_sh = ld_acquire_sys(c.epoch) + 1;
  __syncthreads();
  const unsigned long long e = e_sh;
  // ---- ready: block 0 tells every peer that this rank's input is complete (the kernels that wrote it precede this one
  //      in stream order), then waits for all peers; the other blocks wait for block 0
  if (blockIdx.x == 0) {
    if (tid < c.world) {
      __threadfence_system();
      st_release_sys(c.flags[tid] + c.rank, e);                       // ready[rank] in peer tid's block
      while (ld_acquire_sys(c.flags[c.rank] + tid) < e) { }           // ready[tid] in my block
    }
    __syncthreads();
    if (tid == 0) st_release_sys(c.go, e);
  } else {
    if (tid == 0) while (ld_acquire_sys(c.go) < e) { }
    __syncthreads();
  }
  // ---- sum in rank order (same bits on every rank), mean, local store
  const float inv = 1.f / (float)c.world;
  const long long n4 = count >> 2;
  for (long long i = (long long)blockIdx.x * P2P_THREADS + tid; i < n4; i += (long long)gridDim.x * P2P_THREADS) {
    float4 v[WU];
#pragma unroll
    for (int r = 0; r < WU; ++r)
      v[r] = r < c.world ? ld_peer(reinterpret_cast<const float4*>(c.data[r] + offset) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < WU; ++r)
      if (r < c.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }  // rank order: same bits everywhere
    reinterpret_cast<float4*>(out + offset)[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * P2P_THREADS + tid; i < count; i += (long long)gridDim.x * P2P_THREADS) {
    float acc = 0.f;
    for (int r = 0; r < c.world; ++r) acc += *reinterpret_cast<volatile const float*>(c.data[r] + offset + i);
    out[offset + i] = acc * inv;
  }
  // ---- done: the last block to finish tells every peer that this rank has read everything, waits for all peers (so
  //      that nobody's input is overwritten while somebody still reads it), and closes the call
  __syncthreads();
  __shared__ unsigned int last;
  if (tid == 0) { __threadfence(); last = atomicAdd(c.arrive, 1u) == gridDim.x - 1 ? 1u : 0u; }
  __syncthreads();
  if (last) {
    if (tid < c.world) {
      st_release_sys(c.flags[tid] + c.world + c.rank, e);             // done[rank] in peer tid's block
      while (ld_acquire_sys(c.flags[c.rank] + c.world + tid) < e) { }
    }
    __syncthreads();
    if (tid == 0) { *c.arrive = 0u; st_release_sys(c.epoch, e); }
  }
}

struct P2PHost {
  P2PCtx ctx{};
  void* base = nullptr;          // local IPC allocation: [data: floats][flags: 2 * world u64][epoch][go][arrive]
  size_t data_bytes = 0;
  void* peers[P2P_MAX_RANKS] = {nullptr};
  float* out = nullptr;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" {

/* Layout of the IPC allocation of one rank: floats [0, n) = the gradient sink, then the flag block. */
int gcpnet_p2p_create(int rank, int world, int64_t num_floats, void** handle_out, void** data_out, unsigned char ipc_handle[64]) {
  if (!handle_out || !data_out || !ipc_handle) return fail("p2p_create: null argument");
  if (world < 2 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || num_floats <= 0) return fail("p2p_create: bad rank / world / size");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  P2PHost* h = new P2PHost();
  h->data_bytes = align256((size_t)num_floats * sizeof(float));
  const size_t total = h->data_bytes + align256((size_t)(2 * world + 8) * sizeof(unsigned long long));
  if (cudaMalloc(&h->base, total) != cudaSuccess) { delete h; return fail("p2p_create: cudaMalloc failed"); }
  CUDA_TRY(cudaMemset(h->base, 0, total));
  cudaIpcMemHandle_t ipc;
  if (cudaIpcGetMemHandle(&ipc, h->base) != cudaSuccess) { cudaFree(h->base); delete h; return fail("p2p_create: cudaIpcGetMemHandle failed"); }
  memcpy(ipc_handle, &ipc, 64);
  h->ctx.rank = rank; h->ctx.world = world;
  *handle_out = h; *data_out = h->base;
  return 0;
}

/* all_handles: world x 64 bytes (every rank's IPC handle, own included); out: the local buffer the means are written to. */
int gcpnet_p2p_connect(void* handle, const unsigned char* all_handles, float* out) {
  P2PHost* h = (P2PHost*)handle;
  if (!h || !all_handles || !out) return fail("p2p_connect: null argument");
  const int world = h->ctx.world, rank = h->ctx.rank;
  for (int r = 0; r < world; ++r) {
    void* p = h->base;
    if (r != rank) {
      cudaIpcMemHandle_t ipc;
      memcpy(&ipc, all_handles + 64 * r, 64);
      if (cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return fail("p2p_connect: cudaIpcOpenMemHandle failed (no peer access between the ranks' devices?)");
      }
      h->peers[r] = p;
    }
    h->ctx.data[r] = (float*)p;
    h->ctx.flags[r] = (unsigned long long*)((char*)p + h->data_bytes);
  }
  unsigned long long* mine = h->ctx.flags[rank];
  h->ctx.epoch = mine + 2 * world;
  h->ctx.go = mine + 2 * world + 1;
  h->ctx.arrive = (unsigned int*)(mine + 2 * world + 2);
  h->out = out;
  return 0;
}

/* out[offset .. offset+count) = mean over the ranks of their input slices; enqueued on `stream` (capturable). */
int gcpnet_p2p_allreduce_mean(void* handle, int64_t offset, int64_t count, void* stream) {
  P2PHost* h = (P2PHost*)handle;
  if (!h || !h->out) return fail("p2p_allreduce_mean: not connected");
  if (count <= 0) return 0;
  if (offset % 4 != 0) return fail("p2p_allreduce_mean: offset must be a multiple of 4 floats");
  if (h->ctx.world <= 2) p2p_allreduce_mean_kernel<2><<<P2P_BLOCKS, P2P_THREADS, 0, (cudaStream_t)stream>>>(h->ctx, offset, count, h->out);
  else if (h->ctx.world <= 4) p2p_allreduce_mean_kernel<4><<<P2P_BLOCKS, P2P_THREADS, 0, (cudaStream_t)stream>>>(h->ctx, offset, count, h->out);
  else if (h->ctx.world <= 8) p2p_allreduce_mean_kernel<8><<<P2P_BLOCKS, P2P_THREADS, 0, (cudaStream_t)stream>>>(h->ctx, offset, count, h->out);
  else p2p_allreduce_mean_kernel<P2P_MAX_RANKS><<<P2P_BLOCKS, P2P_THREADS, 0, (cudaStream_t)stream>>>(h->ctx, offset, count, h->out);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_p2p_destroy(void* handle) {
  P2PHost* h = (P2PHost*)handle;
  if (!h) return 0;
  for (int r = 0; r < h->ctx.world; ++r)
    if (h->peers[r]) cudaIpcCloseMemHandle(h->peers[r]);
  cudaFree(h->base);
  delete h;
  return 0;
}

}  // extern "C"
