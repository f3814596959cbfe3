// graph.cu -- per-batch graph preprocessing: destination/source CSR orders, mean frames, localize.
// Replaces the index side of torch_scatter.scatter (gcpnet.py:946; comp/__init__.py:316) and
// comp/__init__.py:220-269 (localize).  Entry points declared in include/gcpnet_b200.h.
#include <cuda_runtime.h>

#include <cstdio>
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <set>
#include <string>

#include "../../include/gcpnet_b200.h"
#include "common.h"

// ---- graph build -------------------------------------------------------------------------------
// An edge_index entry outside [0, N) is the reference's IndexError / device-side assert (fancy indexing h[row], scatter);
// here it would silently corrupt the CSR build and every gather after it, so the build kernels stop the context instead.
__device__ __noinline__ void bad_edge_index(long long e, long long r, long long c, int N) {
  printf("gcpnet_graph_build: edge %lld = (%lld, %lld) is outside [0, %d)\n", e, r, c, N);
  asm volatile("trap;");
}
__global__ void edge_keys_kernel(const int64_t* __restrict__ edge_index, int E, int N, int* __restrict__ row32,
                                 int* __restrict__ col32, int* __restrict__ iota) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = edge_index[e], c = edge_index[(size_t)E + e];
  if (r < 0 || r >= N || c < 0 || c >= N) bad_edge_index(e, r, c, N);
  row32[e] = (int)r;
  col32[e] = (int)c;
  iota[e] = e;
}
__global__ void gather_src_kernel(const int* __restrict__ row32, const int* __restrict__ perm, int E,
                                  int* __restrict__ src, int* __restrict__ iota) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= E) return;
  src[p] = row32[perm[p]];
  iota[p] = p;
}
// ptr[i] = first position whose key >= i  (keys sorted ascending), i in [0, N]
__global__ void segment_ptr_kernel(const int* __restrict__ keys, int E, int N, int* __restrict__ ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > N) return;
  int lo = 0, hi = E;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < i) lo = mid + 1; else hi = mid; }
  ptr[i] = lo;
}
// fbar[i] = mean of frames over the edges leaving node i (comp/__init__.py:316-323), 0 if none
__global__ void mean_frame_kernel(const float* __restrict__ frames, const int* __restrict__ perm,
                                  const int* __restrict__ src_pos, const int* __restrict__ src_ptr, int N,
                                  float* __restrict__ fbar) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 9) return;
  const int i = idx / 9, c = idx - 9 * i;
  const int a = src_ptr[i], b = src_ptr[i + 1];
  float acc = 0.f;
  for (int q = a; q < b; ++q) acc += __ldg(frames + (size_t)perm[src_pos[q]] * 9 + c);
  fbar[idx] = b > a ? acc / (float)(b - a) : 0.f;
}

// ---- small graphs: the whole build in ONE CTA --------------------------------------------------------------------
// At NMS-small sizes (E = 5 120) the radix-sort pipeline above is 13 launches of mostly idle kernels.  One CTA does the
// same job: counting sort by destination / by source with shared-memory cursors, then every node orders its own segment
// by edge id (destination view) / by sorted position (source view), which reproduces the stable order of the radix sort
// exactly -- the result does not depend on the order in which the atomics land.
constexpr int SMALL_MAX_NODES = 12287, SMALL_MAX_EDGES = 16384, SMALL_NT = 1024;  // beyond that the radix-sort pipeline wins (measured)
constexpr int SMALL_LONG_SEG = 48;    // segments longer than this are ranked by the whole CTA instead of one thread's insertion sort
constexpr int SMALL_MAX_LONG = SMALL_MAX_EDGES / SMALL_LONG_SEG + 1;
// order the entries of arr[a, b) ascending (distinct keys), whole CTA: rank by counting, through `scratch`
__device__ __forceinline__ void cta_rank_sort(int* arr, int* scratch, int a, int b, int tid) {
  for (int p = a + tid; p < b; p += SMALL_NT) {
    const int key = arr[p];
    int rank = 0;
    for (int q = a; q < b; ++q) rank += arr[q] < key;
    scratch[a + rank] = key;
  }
  __syncthreads();
  for (int p = a + tid; p < b; p += SMALL_NT) arr[p] = scratch[p];
  __syncthreads();
}
__global__ void __launch_bounds__(SMALL_NT) graph_build_small_kernel(const int64_t* __restrict__ edge_index, int E, int N,
                                                                     const float* __restrict__ frames, int* __restrict__ perm,
                                                                     int* __restrict__ src, int* __restrict__ dst, int* __restrict__ dst_ptr,
                                                                     int* __restrict__ src_pos, int* __restrict__ src_ptr, float* __restrict__ fbar,
                                                                     int* __restrict__ scratch) {
  extern __shared__ int sh[];
  int* cd = sh;            // [N + 1] counts -> cursors (destination)
  int* cs = sh + (N + 1);  // [N + 1] (source)
  __shared__ int carry[2];
  __shared__ int long_n, long_seg[SMALL_MAX_LONG];
  const int tid = threadIdx.x;
  for (int i = tid; i <= N; i += SMALL_NT) { cd[i] = 0; cs[i] = 0; }
  if (tid == 0) long_n = 0;
  __syncthreads();
  for (int e = tid; e < E; e += SMALL_NT) {
    const int64_t r = edge_index[e], c = edge_index[(size_t)E + e];
    if (r < 0 || r >= N || c < 0 || c >= N) bad_edge_index(e, r, c, N);
    atomicAdd(&cd[(int)c], 1);
    atomicAdd(&cs[(int)r], 1);
  }
  __syncthreads();
  // exclusive scans (one warp each, chunks of 32 with a running carry): ptr arrays to global, cursors stay in shared memory
  if (tid < 64) {
    int* c = tid < 32 ? cd : cs;
    int* ptr = tid < 32 ? dst_ptr : src_ptr;
    const int lane = tid & 31;
    int run = 0;
    for (int base = 0; base <= N; base += 32) {
      const int i = base + lane;
      const int v = i < N ? c[i] : 0;
      int x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      const int excl = run + x - v;
      if (i <= N) { ptr[i] = excl; c[i] = excl; }
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    (void)carry;
  }
  __syncthreads();
  for (int e = tid; e < E; e += SMALL_NT) perm[atomicAdd(&cd[(int)edge_index[(size_t)E + e]], 1)] = e;
  __syncthreads();
  // order every destination segment by edge id: short segments by one thread's insertion sort, hub nodes' long
  // segments (O(deg^2) in one thread otherwise) by the whole CTA
  for (int i = tid; i < N; i += SMALL_NT) {
    const int a = dst_ptr[i], b = dst_ptr[i + 1];
    if (b - a > SMALL_LONG_SEG) { long_seg[atomicAdd(&long_n, 1)] = i; continue; }
    for (int p = a + 1; p < b; ++p) {
      const int key = perm[p];
      int q = p - 1;
      while (q >= a && perm[q] > key) { perm[q + 1] = perm[q]; --q; }
      perm[q + 1] = key;
    }
  }
  __syncthreads();
  for (int j = 0; j < long_n; ++j) { const int i = long_seg[j]; cta_rank_sort(perm, scratch, dst_ptr[i], dst_ptr[i + 1], tid); }
  for (int i = tid; i < N; i += SMALL_NT) {
    const int a = dst_ptr[i], b = dst_ptr[i + 1];
    for (int p = a; p < b; ++p) { dst[p] = i; src[p] = (int)edge_index[perm[p]]; }
  }
  __syncthreads();
  if (tid == 0) long_n = 0;
  for (int p = tid; p < E; p += SMALL_NT) src_pos[atomicAdd(&cs[src[p]], 1)] = p;
  __syncthreads();
  for (int i = tid; i < N; i += SMALL_NT) {
    const int a = src_ptr[i], b = src_ptr[i + 1];
    if (b - a > SMALL_LONG_SEG) { long_seg[atomicAdd(&long_n, 1)] = i; continue; }
    for (int p = a + 1; p < b; ++p) {
      const int key = src_pos[p];
      int q = p - 1;
      while (q >= a && src_pos[q] > key) { src_pos[q + 1] = src_pos[q]; --q; }
      src_pos[q + 1] = key;
    }
  }
  __syncthreads();
  for (int j = 0; j < long_n; ++j) { const int i = long_seg[j]; cta_rank_sort(src_pos, scratch, src_ptr[i], src_ptr[i + 1], tid); }
  for (int idx = tid; idx < N * 9; idx += SMALL_NT) {  // mean frame over the edges leaving each node
    const int i = idx / 9, c = idx - 9 * i;
    const int a = src_ptr[i], b = src_ptr[i + 1];
    float acc = 0.f;
    for (int q = a; q < b; ++q) acc += __ldg(frames + (size_t)perm[src_pos[q]] * 9 + c);
    fbar[idx] = b > a ? acc / (float)(b - a) : 0.f;
  }
}


// ---- autoregressive views (gcpnet.py:1065-1116) ----------------------------------------------------------------------------
// gather-row ids: node i reads row 2i (node_rep) on edges with row < col and row 2i+1 (node_rep_regressive) on the others,
// at BOTH ends of the edge (each of the reference's two passes feeds one table to both ends)
__global__ void ar_keys_kernel(const int64_t* __restrict__ edge_index, int E, int N, int64_t* __restrict__ keys) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = edge_index[e], c = edge_index[(size_t)E + e];
  if (r < 0 || r >= N || c < 0 || c >= N) bad_edge_index(e, r, c, N);
  const int64_t flag = r < c ? 0 : 1;
  keys[e] = 2 * r + flag;
  keys[(size_t)E + e] = 2 * c + flag;
}
__global__ void ar_derive_kernel(const int* __restrict__ gsrc, const int* __restrict__ gdst, const int* __restrict__ vdst_ptr,
                                 const int* __restrict__ vsrc_ptr, int E, int N, int* __restrict__ src, int* __restrict__ dst,
                                 int* __restrict__ dst_ptr, int* __restrict__ src_ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) { src[i] = gsrc[i] >> 1; dst[i] = gdst[i] >> 1; }
  if (i <= N) { dst_ptr[i] = vdst_ptr[2 * i]; src_ptr[i] = vsrc_ptr[2 * i]; }
}

// ---- node mask (gcpnet.py:1202-1217; comp/__init__.py:294-300) ----------------------------------------------------------
// relabel[i] = number of unmasked nodes before i (torch_geometric.utils.subgraph relabels in subset order); one CTA
__global__ void __launch_bounds__(1024) mask_scan_kernel(const unsigned char* __restrict__ mask, int N, int* __restrict__ relabel) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + tid;
    const int v = (i < N && mask[i]) ? 1 : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      wsum[lane] = w;  // inclusive
    }
    __syncthreads();
    const int excl = carry + (warp > 0 ? wsum[warp - 1] : 0) + x - v;
    if (i < N) relabel[i] = excl;
    __syncthreads();
    if (tid == 0) carry += wsum[31];
    __syncthreads();
  }
  if (tid == 0) relabel[N] = carry;
}
__global__ void mask_frames_kernel(const int64_t* __restrict__ edge_index, int E, const float* __restrict__ frames,
                                   const unsigned char* __restrict__ mask, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const bool m = mask[edge_index[e]] && mask[edge_index[(size_t)E + e]];
#pragma unroll
  for (int i = 0; i < 9; ++i) out[(size_t)e * 9 + i] = m ? frames[(size_t)e * 9 + i] : 0.f;  // predicated, never inf * 0
}
__global__ void mask_mean_frames_kernel(const float* __restrict__ frames, const unsigned char* __restrict__ mask,
                                        const int* __restrict__ relabel, const int* __restrict__ perm, const int* __restrict__ dst,
                                        const int* __restrict__ src_pos, const int* __restrict__ src_ptr, int N,
                                        float* __restrict__ fbar_ff, float* __restrict__ fbar_pos) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 9) return;
  const int i = idx / 9, c = idx - 9 * i;
  const int a = src_ptr[i], b = src_ptr[i + 1];
  const bool mi = mask[i] != 0;
  float sum_pos = 0.f, sum_ff = 0.f;
  int cnt_ff = 0;
  for (int q = a; q < b; ++q) {
    const int p = src_pos[q], j = dst[p];
    if (!(mi && mask[j])) continue;  // the only edges whose frames are read (others may hold inf)
    const float f = __ldg(frames + (size_t)perm[p] * 9 + c);
    sum_pos += f;
    // feed-forward GCPs run on the subgraph of the mask with RELABELLED node ids, and the reference hands scalarize the
    // original [N] mask: the subgraph edge (i', j') contributes iff mask[i'] & mask[j']  (gcpnet.py:1232-1239)
    ++cnt_ff;
    if (mask[relabel[i]] && mask[relabel[j]]) sum_ff += f;
  }
  fbar_pos[idx] = b > a ? sum_pos / (float)(b - a) : 0.f;
  fbar_ff[idx] = cnt_ff > 0 ? sum_ff / (float)cnt_ff : 0.f;
}

// ---- centroids -------------------------------------------------------------------------------------------------------------
__global__ void centroid_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N, int G,
                                const unsigned char* __restrict__ mask, float* __restrict__ centroid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * 3) return;
  const int g = idx / 3, c = idx - 3 * g;
  int lo = 0, hi = N;  // batch_index is non-decreasing: [first, last) of graph g by bisection
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (batch[mid] < g) lo = mid + 1; else hi = mid; }
  const int first = lo;
  hi = N;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (batch[mid] <= g) lo = mid + 1; else hi = mid; }
  float acc = 0.f;
  int cnt = 0;
  for (int i = first; i < lo; ++i)
    if (mask == nullptr || mask[i]) { acc += pos[(size_t)i * 3 + c]; ++cnt; }
  centroid[idx] = cnt > 0 ? acc / (float)cnt : 0.f;
}
__global__ void shift_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N, const float* __restrict__ centroid,
                             const unsigned char* __restrict__ mask, float sign, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 3) return;
  const int i = idx / 3, c = idx - 3 * i;
  if (mask != nullptr && !mask[i]) { out[idx] = __int_as_float(0x7f800000); return; }  // comp/__init__.py:187-193,208-211
  out[idx] = pos[idx] + sign * centroid[batch[i] * 3 + c];
}

// frames = [x_diff; x_cross; x_vertical] (comp/__init__.py:220-269)
__global__ void localize_kernel(const float* __restrict__ pos, const int64_t* __restrict__ edge_index, int E,
                                int norm_x_diff, const unsigned char* __restrict__ mask, float* __restrict__ frames) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = edge_index[e], c = edge_index[(size_t)E + e];
  if (mask != nullptr && !(mask[r] && mask[c])) {  // comp/__init__.py:229-236,262-264: masked edges carry +inf frames
    float* f = frames + (size_t)e * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = __int_as_float(0x7f800000);
    return;
  }
  const float ax = pos[3 * r], ay = pos[3 * r + 1], az = pos[3 * r + 2];
  const float bx = pos[3 * c], by = pos[3 * c + 1], bz = pos[3 * c + 2];
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  if (norm_x_diff) {
    const float dn = sqrtf(dx * dx + dy * dy + dz * dz) + 1.f;
    dx /= dn; dy /= dn; dz /= dn;
    const float cn = sqrtf(cx * cx + cy * cy + cz * cz) + 1.f;
    cx /= cn; cy /= cn; cz /= cn;
  }
  float* f = frames + (size_t)e * 9;
  f[0] = dx; f[1] = dy; f[2] = dz; f[3] = cx; f[4] = cy; f[5] = cz;
  f[6] = dy * cz - dz * cy; f[7] = dz * cx - dx * cz; f[8] = dx * cy - dy * cx;
}

extern "C" {

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t gcpnet_graph_workspace_bytes(int64_t E, int64_t N) {
  (void)N;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, (int)E, 0, 32, (cudaStream_t)0);
  // row32, col32, iota, sorted-src keys + cub temp
  return 4 * align256((size_t)(E > 0 ? E : 1) * sizeof(int)) + align256(cub_bytes) + 256;
}

int gcpnet_graph_build(const int64_t* edge_index, int64_t E64, int64_t N64, const float* frames, int32_t* perm,
                       int32_t* src, int32_t* dst, int32_t* dst_ptr, int32_t* src_pos, int32_t* src_ptr, float* fbar,
                       void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (E64 < 0 || N64 <= 0 || E64 >= (1LL << 31) || N64 >= (1LL << 31)) return fail("graph_build: sizes out of range");
  const int E = (int)E64, N = (int)N64;
  if (workspace_bytes < gcpnet_graph_workspace_bytes(E64, N64)) return fail("graph_build: workspace too small");
  const int T = 256;
  if (E == 0) {
    CUDA_TRY(cudaMemsetAsync(dst_ptr, 0, (size_t)(N + 1) * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(src_ptr, 0, (size_t)(N + 1) * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(fbar, 0, (size_t)N * 9 * sizeof(float), st));
    return 0;
  }
  GcpTimedScope timed(T_GRAPH_BUILD, st);
  if (N <= SMALL_MAX_NODES && E <= SMALL_MAX_EDGES) {
    const int bytes = 2 * (N + 1) * (int)sizeof(int);
    {  // opt in to the large dynamic shared memory once per device
      static std::mutex mu;
      static std::set<int> done;
      int dev = 0;
      CUDA_TRY(cudaGetDevice(&dev));
      std::lock_guard<std::mutex> lock(mu);
      if (!done.count(dev)) {
        CUDA_TRY(cudaFuncSetAttribute(graph_build_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (SMALL_MAX_NODES + 1) * (int)sizeof(int)));
        done.insert(dev);
      }
    }
    graph_build_small_kernel<<<1, SMALL_NT, bytes, st>>>(edge_index, E, N, frames, perm, src, dst, dst_ptr, src_pos, src_ptr, fbar, (int*)workspace);
    gcp_note_launches(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  char* ws = (char*)workspace;
  const size_t seg = align256((size_t)E * sizeof(int));
  int* row32 = (int*)ws; int* col32 = (int*)(ws + seg); int* iota = (int*)(ws + 2 * seg); int* srckeys = (int*)(ws + 3 * seg);
  void* cub_tmp = ws + 4 * seg;
  size_t cub_bytes = workspace_bytes - 4 * seg;
  int bits = 1;
  while ((1LL << bits) < N64) ++bits;
  edge_keys_kernel<<<(E + T - 1) / T, T, 0, st>>>(edge_index, E, N, row32, col32, iota);
  gcp_note_launches(1);
  // stable sort by destination: positions keep the caller's relative order inside a segment
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, col32, dst, iota, perm, E, 0, bits, st));
  gcp_note_launches(1);  // cub launches >= 1 kernel per sort; counted as one
  gather_src_kernel<<<(E + T - 1) / T, T, 0, st>>>(row32, perm, E, src, iota);
  gcp_note_launches(1);
  segment_ptr_kernel<<<(N + 1 + T - 1) / T, T, 0, st>>>(dst, E, N, dst_ptr);
  gcp_note_launches(1);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, src, srckeys, iota, src_pos, E, 0, bits, st));
  gcp_note_launches(1);
  segment_ptr_kernel<<<(N + 1 + T - 1) / T, T, 0, st>>>(srckeys, E, N, src_ptr);
  gcp_note_launches(1);
  mean_frame_kernel<<<(N * 9 + T - 1) / T, T, 0, st>>>(frames, perm, src_pos, src_ptr, N, fbar);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_localize_masked(const float* pos, const int64_t* edge_index, int64_t E, int norm_x_diff, const uint8_t* node_mask,
                           float* frames, void* stream) {
  if (E <= 0) return 0;
  localize_kernel<<<(int)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos, edge_index, (int)E, norm_x_diff, node_mask, frames);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int gcpnet_localize(const float* pos, const int64_t* edge_index, int64_t E, int norm_x_diff, float* frames, void* stream) {
  return gcpnet_localize_masked(pos, edge_index, E, norm_x_diff, nullptr, frames, stream);
}

// ---- autoregressive views ------------------------------------------------------------------------------------------------
size_t gcpnet_graph_ar_workspace_bytes(int64_t E, int64_t N) {
  return align256((size_t)(E > 0 ? E : 1) * 2 * sizeof(int64_t)) + align256((size_t)2 * N * 9 * sizeof(float)) +
         gcpnet_graph_workspace_bytes(E, 2 * N);
}
int gcpnet_graph_build_autoregressive(const int64_t* edge_index, int64_t E64, int64_t N64, const float* frames, int32_t* perm,
                                      int32_t* src, int32_t* dst, int32_t* dst_ptr, int32_t* src_pos, int32_t* src_ptr, float* fbar,
                                      int32_t* gsrc, int32_t* gdst, int32_t* vdst_ptr, int32_t* vsrc_ptr, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (E64 < 0 || N64 <= 0 || E64 >= (1LL << 31) || N64 >= (1LL << 30)) return fail("graph_build_autoregressive: sizes out of range");
  if (workspace_bytes < gcpnet_graph_ar_workspace_bytes(E64, N64)) return fail("graph_build_autoregressive: workspace too small");
  const int E = (int)E64, N = (int)N64, T = 256;
  char* ws = (char*)workspace;
  int64_t* keys = (int64_t*)ws;
  const size_t o1 = align256((size_t)(E > 0 ? E : 1) * 2 * sizeof(int64_t));
  float* fbar2 = (float*)(ws + o1);
  const size_t o2 = o1 + align256((size_t)2 * N * 9 * sizeof(float));
  if (E > 0) {
    ar_keys_kernel<<<(E + T - 1) / T, T, 0, st>>>(edge_index, E, N, keys);
    gcp_note_launches(1);
  }
  // the ordinary build over the 2N gather rows: sorted by (destination, flag); its CSRs are the gather-row CSRs
  if (gcpnet_graph_build(keys, E64, 2 * N64, frames, perm, gsrc, gdst, vdst_ptr, src_pos, vsrc_ptr, fbar2, ws + o2, workspace_bytes - o2, stream))
    return 1;
  ar_derive_kernel<<<((E > N + 1 ? E : N + 1) + T - 1) / T, T, 0, st>>>(gsrc, gdst, vdst_ptr, vsrc_ptr, E, N, src, dst, dst_ptr, src_ptr);
  gcp_note_launches(1);
  if (E > 0) {
    mean_frame_kernel<<<(N * 9 + T - 1) / T, T, 0, st>>>(frames, perm, src_pos, src_ptr, N, fbar);
    gcp_note_launches(1);
  } else {
    CUDA_TRY(cudaMemsetAsync(fbar, 0, (size_t)N * 9 * sizeof(float), st));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- node mask ---------------------------------------------------------------------------------------------------------------
int gcpnet_graph_mask(const int64_t* edge_index, int64_t E64, int64_t N64, const float* frames, const uint8_t* node_mask,
                      const gcpnet_graph* graph, float* frames_eff, float* fbar_ff, float* fbar_pos, void* workspace,
                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!node_mask || !graph || !frames_eff || !fbar_ff || !fbar_pos) return fail("graph_mask: null argument");
  if (E64 < 0 || N64 <= 0 || E64 >= (1LL << 31) || N64 >= (1LL << 31)) return fail("graph_mask: sizes out of range");
  if (workspace_bytes < (size_t)(N64 + 1) * sizeof(int)) return fail("graph_mask: workspace too small");
  const int E = (int)E64, N = (int)N64, T = 256;
  int* relabel = (int*)workspace;
  mask_scan_kernel<<<1, 1024, 0, st>>>(node_mask, N, relabel);
  gcp_note_launches(1);
  if (E > 0) {
    mask_frames_kernel<<<(E + T - 1) / T, T, 0, st>>>(edge_index, E, frames, node_mask, frames_eff);
    gcp_note_launches(1);
  }
  mask_mean_frames_kernel<<<(N * 9 + T - 1) / T, T, 0, st>>>(frames, node_mask, relabel, graph->perm, graph->dst, graph->src_pos,
                                                             graph->src_ptr, N, fbar_ff, fbar_pos);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- centralize / decentralize (comp/__init__.py:170-217) ----------------------------------------------------------------
int gcpnet_centralize(const float* pos, const int64_t* batch_index, int64_t N, int64_t G, const uint8_t* node_mask, float* centroid,
                      float* centered, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 0 || G <= 0) return 0;
  centroid_kernel<<<(int)((G * 3 + 127) / 128), 128, 0, st>>>(pos, batch_index, (int)N, (int)G, node_mask, centroid);
  shift_kernel<<<(int)((N * 3 + 255) / 256), 256, 0, st>>>(pos, batch_index, (int)N, centroid, node_mask, -1.f, centered);
  gcp_note_launches(2);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int gcpnet_decentralize(const float* pos, const int64_t* batch_index, int64_t N, const float* centroid, const uint8_t* node_mask,
                        float* out, void* stream) {
  if (N <= 0) return 0;
  shift_kernel<<<(int)((N * 3 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos, batch_index, (int)N, centroid, node_mask, 1.f, out);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
