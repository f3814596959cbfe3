// api.cu -- CUDA kernels (sm_100a) and the extern "C" entry points declared in include/gcpnet_b200.h.
//
// Every entry point only enqueues kernels on the caller's stream: no allocation, no host
// synchronisation, so the whole layer forward+backward can be captured into a CUDA graph.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <string>

#include "layer_setup.h"

using namespace gcp;

static thread_local std::string g_last_error;
static int fail(const std::string& msg) { g_last_error = msg; return 1; }
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t err__ = (expr);                                                                \
    if (err__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
template <int TE, int NT>
__global__ void __launch_bounds__(NT) edge_fwd_kernel(const __grid_constant__ EdgeParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ntiles = (p.E + TE - 1) / TE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) edge_fwd_tile<TE, NT>(p, smem, tile);
}

template <int TE, int NT>
__global__ void __launch_bounds__(NT) edge_bwd_kernel(const __grid_constant__ EdgeParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ntiles = (p.E + TE - 1) / TE;
  float* prow = p.partial + (size_t)blockIdx.x * p.partial_stride;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    edge_bwd_tile<TE, NT>(p, smem, tile, prow, tile != (int)blockIdx.x);
}

template <int TE, int NT>
__global__ void __launch_bounds__(NT) node_fwd_kernel(const __grid_constant__ NodeParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ntiles = (p.N + TE - 1) / TE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) node_fwd_tile<TE, NT>(p, smem, tile);
}

template <int TE, int NT>
__global__ void __launch_bounds__(NT) node_bwd_kernel(const __grid_constant__ NodeParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ntiles = (p.N + TE - 1) / TE;
  float* prow = p.partial + (size_t)blockIdx.x * p.partial_stride;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    node_bwd_tile<TE, NT>(p, smem, tile, prow, tile != (int)blockIdx.x);
}

// aggregate only (GCPMessagePassing.forward): out[i] = mean/sum over the destination segment
__global__ void aggregate_kernel(const float* __restrict__ msg, const int* __restrict__ dst_ptr, int N, int W,
                                 int reduce_mean, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * W) return;
  const int i = (int)(idx / W), f = (int)(idx - (long long)i * W);
  const int a = dst_ptr[i], b = dst_ptr[i + 1];
  float acc = 0.f;
  for (int q = a; q < b; ++q) acc += __ldg(msg + (size_t)q * W + f);
  if (reduce_mean && b - a > 1) acc /= (float)(b - a);
  out[idx] = acc;
}

// cotangent of the layer input: direct path (already in g_h/g_chi) + gathered-by-destination +
// gathered-by-source per-edge cotangents, summed in a fixed order (deterministic, no atomics)
__global__ void node_cotangent_reduce_kernel(float* __restrict__ g_h, float* __restrict__ g_chi,
                                             const float* __restrict__ grow, const float* __restrict__ gcol,
                                             const int* __restrict__ dst_ptr, const int* __restrict__ src_ptr,
                                             const int* __restrict__ src_pos, int N, int s, int v3) {
  const int W = s + v3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * W) return;
  const int i = (int)(idx / W), f = (int)(idx - (long long)i * W);
  float* out = f < s ? g_h + (size_t)i * s + f : g_chi + (size_t)i * v3 + (f - s);
  float acc = *out;
  for (int q = dst_ptr[i]; q < dst_ptr[i + 1]; ++q) acc += __ldg(gcol + (size_t)q * W + f);
  for (int q = src_ptr[i]; q < src_ptr[i + 1]; ++q) acc += __ldg(grow + (size_t)__ldg(src_pos + q) * W + f);
  *out = acc;
}

// flat parameter gradient = fixed-order sum of the per-CTA partial rows
__global__ void partial_reduce_kernel(float* __restrict__ out, const float* __restrict__ pe, int ne, int ge,
                                      const float* __restrict__ pn, int nn, int gn) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ne + nn) return;
  float acc = 0.f;
  if (idx < ne) { for (int g = 0; g < ge; ++g) acc += __ldg(pe + (size_t)g * ne + idx); }
  else { const int j = idx - ne; for (int g = 0; g < gn; ++g) acc += __ldg(pn + (size_t)g * nn + j); }
  out[idx] = acc;
}

// ---- graph build -------------------------------------------------------------------------------
__global__ void edge_keys_kernel(const int64_t* __restrict__ edge_index, int E, int* __restrict__ row32,
                                 int* __restrict__ col32, int* __restrict__ iota) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  row32[e] = (int)edge_index[e];
  col32[e] = (int)edge_index[(size_t)E + e];
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

// frames = [x_diff; x_cross; x_vertical] (comp/__init__.py:220-269, no node mask)
__global__ void localize_kernel(const float* __restrict__ pos, const int64_t* __restrict__ edge_index, int E,
                                int norm_x_diff, float* __restrict__ frames) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = edge_index[e], c = edge_index[(size_t)E + e];
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

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
template <class K>
static int set_smem(K kernel, int bytes) {
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

static int launch_edge_fwd(const EdgeParams& p, int TE, int grid, cudaStream_t st) {
  const int bytes = p.sm.total * 4;
  if (TE == 64) {
    if (set_smem(edge_fwd_kernel<64, EDGE_NT>, bytes)) return 1;
    edge_fwd_kernel<64, EDGE_NT><<<grid, EDGE_NT, bytes, st>>>(p);
  } else {
    if (set_smem(edge_fwd_kernel<32, EDGE_NT>, bytes)) return 1;
    edge_fwd_kernel<32, EDGE_NT><<<grid, EDGE_NT, bytes, st>>>(p);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int launch_edge_bwd(const EdgeParams& p, int TE, int grid, cudaStream_t st) {
  const int bytes = p.sm.total * 4;
  if (TE != 32) return fail("edge backward tile must be 32");
  if (set_smem(edge_bwd_kernel<32, EDGE_NT>, bytes)) return 1;
  edge_bwd_kernel<32, EDGE_NT><<<grid, EDGE_NT, bytes, st>>>(p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int launch_node_fwd(const NodeParams& p, int TE, int grid, cudaStream_t st) {
  const int bytes = p.sm.total * 4;
  if (TE == 32) {
    if (set_smem(node_fwd_kernel<32, NODE_NT>, bytes)) return 1;
    node_fwd_kernel<32, NODE_NT><<<grid, NODE_NT, bytes, st>>>(p);
  } else {
    if (set_smem(node_fwd_kernel<16, NODE_NT>, bytes)) return 1;
    node_fwd_kernel<16, NODE_NT><<<grid, NODE_NT, bytes, st>>>(p);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int launch_node_bwd(const NodeParams& p, int TE, int grid, cudaStream_t st) {
  const int bytes = p.sm.total * 4;
  if (TE == 32) {
    if (set_smem(node_bwd_kernel<32, NODE_NT>, bytes)) return 1;
    node_bwd_kernel<32, NODE_NT><<<grid, NODE_NT, bytes, st>>>(p);
  } else {
    if (set_smem(node_bwd_kernel<16, NODE_NT>, bytes)) return 1;
    node_bwd_kernel<16, NODE_NT><<<grid, NODE_NT, bytes, st>>>(p);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int gcpnet_version(void) { return 100; }
const char* gcpnet_last_error(void) { return g_last_error.c_str(); }

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
  char* ws = (char*)workspace;
  const size_t seg = align256((size_t)E * sizeof(int));
  int* row32 = (int*)ws; int* col32 = (int*)(ws + seg); int* iota = (int*)(ws + 2 * seg); int* srckeys = (int*)(ws + 3 * seg);
  void* cub_tmp = ws + 4 * seg;
  size_t cub_bytes = workspace_bytes - 4 * seg;
  int bits = 1;
  while ((1LL << bits) < N64) ++bits;
  edge_keys_kernel<<<(E + T - 1) / T, T, 0, st>>>(edge_index, E, row32, col32, iota);
  // stable sort by destination: positions keep the caller's relative order inside a segment
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, col32, dst, iota, perm, E, 0, bits, st));
  gather_src_kernel<<<(E + T - 1) / T, T, 0, st>>>(row32, perm, E, src, iota);
  segment_ptr_kernel<<<(N + 1 + T - 1) / T, T, 0, st>>>(dst, E, N, dst_ptr);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, src, srckeys, iota, src_pos, E, 0, bits, st));
  segment_ptr_kernel<<<(N + 1 + T - 1) / T, T, 0, st>>>(srckeys, E, N, src_ptr);
  mean_frame_kernel<<<(N * 9 + T - 1) / T, T, 0, st>>>(frames, perm, src_pos, src_ptr, N, fbar);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_localize(const float* pos, const int64_t* edge_index, int64_t E, int norm_x_diff, float* frames, void* stream) {
  if (E <= 0) return 0;
  localize_kernel<<<(int)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos, edge_index, (int)E, norm_x_diff, frames);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_layer_plan(const gcpnet_layer* layer, int64_t N, int64_t E, gcpnet_plan* plan) {
  if (!layer || !plan) return fail("layer_plan: null argument");
  const std::string e = make_plan(*layer, N, E, plan);
  if (!e.empty()) return fail("layer_plan: " + e);
  return 0;
}

static int run_edge_forward(const gcpnet_layer& l, const gcpnet_graph& g, const gcpnet_forward_io& io, cudaStream_t st) {
  if (g.num_edges == 0) return 0;
  const LayerOps ops = layer_ops(l);
  EdgeSmem sm;
  const int TE = pick_edge_tile(l, ops, g.num_edges, false, &sm);
  if (!TE) return fail("edge forward: no tile plan fits");
  EdgeParams p = make_edge_params(l, g, ops, sm);
  p.h = io.h; p.chi = io.chi; p.e = io.e; p.xi = io.xi; p.frames = io.frames;
  p.msg = io.msg; p.saved = io.saved_edge;
  int grid = (int)((g.num_edges + TE - 1) / TE);
  if (grid > MAX_PERSISTENT_CTAS) grid = MAX_PERSISTENT_CTAS;
  return launch_edge_fwd(p, TE, grid, st);
}

int gcpnet_layer_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                         const gcpnet_forward_io* io, void* stream) {
  if (!layer || !graph || !plan || !io) return fail("layer_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gcpnet_layer& l = *layer;
  const std::string e = check_layer(l);
  if (!e.empty()) return fail("layer_forward: " + e);
  if (l.has_pos && (!io->pos || !io->out_pos)) return fail("layer_forward: node positions required");
  if (graph->num_nodes <= 0) return 0;
  if (run_edge_forward(l, *graph, *io, st)) return 1;
  const LayerOps ops = layer_ops(l);
  NodeSmem sm;
  const int TE = pick_node_tile(l, ops, graph->num_nodes, false, &sm);
  if (!TE) return fail("node forward: no tile plan fits");
  NodeParams p = make_node_params(l, *graph, ops, sm);
  p.h = io->h; p.chi = io->chi; p.msg = io->msg; p.pos = io->pos;
  p.out_h = io->out_h; p.out_chi = io->out_chi; p.out_pos = io->out_pos; p.saved = io->saved_node;
  return launch_node_fwd(p, TE, plan->node_grid_fwd, st);
}

int gcpnet_message_passing_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                                   const gcpnet_forward_io* io, float* aggregate, void* stream) {
  if (!layer || !graph || !plan || !io || !aggregate) return fail("message_passing_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const std::string e = check_layer(*layer);
  if (!e.empty()) return fail("message_passing_forward: " + e);
  if (graph->num_nodes <= 0) return 0;
  if (run_edge_forward(*layer, *graph, *io, st)) return 1;
  const int W = layer->s + 3 * layer->v;
  const long long tot = graph->num_nodes * W;
  aggregate_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(io->msg, graph->dst_ptr, (int)graph->num_nodes, W,
                                                            layer->reduce_mean, aggregate);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_layer_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                          const gcpnet_backward_io* io, void* stream) {
  if (!layer || !graph || !plan || !io) return fail("layer_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gcpnet_layer& l = *layer;
  const gcpnet_graph& g = *graph;
  const std::string e = check_layer(l);
  if (!e.empty()) return fail("layer_backward: " + e);
  if (!io->saved_edge && g.num_edges > 0) return fail("layer_backward: forward ran without saved activations");
  if (!io->saved_node) return fail("layer_backward: forward ran without saved activations");
  if (g.num_nodes <= 0) return 0;
  const LayerOps ops = layer_ops(l);
  const int W = l.s + 3 * l.v;
  // node update backward -> direct cotangent of (h, chi) in g_h/g_chi, cotangent of the aggregate in ws_agg
  NodeSmem nsm;
  const int TN = pick_node_tile(l, ops, g.num_nodes, true, &nsm);
  if (!TN) return fail("node backward: no tile plan fits");
  NodeParams np = make_node_params(l, g, ops, nsm);
  np.saved = const_cast<float*>(io->saved_node);
  np.g_out_h = io->g_out_h; np.g_out_chi = io->g_out_chi; np.g_out_pos = io->g_out_pos;
  np.g_x_h = io->g_h; np.g_x_chi = io->g_chi; np.g_agg = io->ws_agg;
  np.partial = io->ws_node_partial;
  if (launch_node_bwd(np, TN, plan->node_grid_bwd, st)) return 1;
  int edge_grid = 0;
  if (g.num_edges > 0) {
    EdgeSmem esm;
    const int TE = pick_edge_tile(l, ops, g.num_edges, true, &esm);
    if (!TE) return fail("edge backward: no tile plan fits");
    EdgeParams ep = make_edge_params(l, g, ops, esm);
    ep.h = io->h; ep.chi = io->chi; ep.e = io->e; ep.xi = io->xi; ep.frames = io->frames;
    ep.saved = const_cast<float*>(io->saved_edge);
    ep.gagg = io->ws_agg;
    ep.grow = io->ws_edge; ep.gcol = io->ws_edge + (size_t)g.num_edges * W;
    ep.ge = io->g_e; ep.gxi = io->g_xi;
    ep.partial = io->ws_edge_partial;
    edge_grid = plan->edge_grid_bwd;
    if (launch_edge_bwd(ep, TE, edge_grid, st)) return 1;
    const long long tot = g.num_nodes * W;
    node_cotangent_reduce_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
        io->g_h, io->g_chi, ep.grow, ep.gcol, g.dst_ptr, g.src_ptr, g.src_pos, (int)g.num_nodes, l.s, 3 * l.v);
  }
  const int np_tot = l.n_edge_params + l.n_node_params;
  partial_reduce_kernel<<<(np_tot + 255) / 256, 256, 0, st>>>(io->g_params, io->ws_edge_partial, l.n_edge_params, edge_grid,
                                                           io->ws_node_partial, l.n_node_params, plan->node_grid_bwd);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
