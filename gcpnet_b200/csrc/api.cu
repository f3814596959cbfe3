// api.cu -- CUDA kernels (sm_100a) of the layer and the extern "C" entry points declared in
// include/gcpnet_b200.h.
//
// Every entry point only enqueues kernels on the caller's stream: no allocation, no host
// synchronisation, so the whole layer forward+backward can be captured into a CUDA graph.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "layer_setup.h"
#include "node_wgrad.cuh"
#include "layernorm.cuh"

using namespace gcp;

static thread_local std::string g_last_error;
int gcp_fail(const std::string& msg) { g_last_error = msg; return 1; }
static std::atomic<unsigned long long> g_launches{0};
void gcp_note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static std::atomic<int> g_opt_tc{1};
static std::atomic<int> g_opt_post_fused{2};   // ordering of the backward's node-level finish (run_tc_edge_backward)
static std::atomic<int> g_opt_early_fork{1};   // node parameter-gradient work forks right after the node backward
static std::atomic<long long*> g_tc_dbg{nullptr};
static std::atomic<bool> g_profile{false};
static std::mutex g_profile_mu;
static std::vector<cudaEvent_t> g_profile_ev[T_COUNT];  // begin, end, begin, end, ...
bool gcp_profile_on() { return g_profile.load(std::memory_order_relaxed); }
void gcp_profile_mark(int which, bool begin, cudaStream_t st) {
  (void)begin;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  std::lock_guard<std::mutex> lock(g_profile_mu);
  g_profile_ev[which].push_back(ev);
}

// ------------------------------------------------------------------------------------------
// kernels: persistent, one CTA per SM, weights streamed through the shared-memory ring
// ------------------------------------------------------------------------------------------
template <int TE, int NT, int SLF>
__global__ void __launch_bounds__(NT, 1) edge_fwd_kernel(const __grid_constant__ EdgeParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.E + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = edge_pipe(p, smem, mine);
  wpipe_start<NT>(wp);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    edge_fwd_tile<TE, NT, SLF>(p, smem, tile, wp, tile == (int)blockIdx.x);
}

template <int TE, int NT, int SLF, int SLD>
__global__ void __launch_bounds__(NT, 1) edge_bwd_kernel(const __grid_constant__ EdgeParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.E + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = edge_pipe(p, smem, mine);
#if GCP_STAMPS
  if (p.dbg != nullptr && blockIdx.x == 0) wp.dbg = p.dbg + 512;
#endif
  wpipe_start<NT>(wp);
  float* prow = p.partial + (size_t)blockIdx.x * p.partial_stride;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    edge_bwd_tile<TE, NT, SLF, SLD>(p, smem, tile, wp, prow, tile != (int)blockIdx.x);
#if GCP_STAMPS
    wp.dbg = nullptr;  // first tile only
#endif
  }
}

template <int TE, int NT, int SLF>
__global__ void __launch_bounds__(NT, 1) node_fwd_kernel(const __grid_constant__ NodeParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.N + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = node_pipe(p, smem, mine);
  wpipe_start<NT>(wp);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    node_fwd_tile<TE, NT, SLF>(p, smem, tile, wp, tile == (int)blockIdx.x);
}

template <int TE, int NT, int SLF, int SLD>
__global__ void __launch_bounds__(NT, 1) node_bwd_kernel(const __grid_constant__ NodeParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.N + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = node_pipe(p, smem, mine);
  if (p.dbg != nullptr && blockIdx.x == 0) wp.dbg = p.dbg + 336;
  wpipe_start<NT>(wp);
  float* prow = p.partial + (size_t)blockIdx.x * p.partial_stride;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    node_bwd_tile<TE, NT, SLF, SLD>(p, smem, tile, wp, prow, tile != (int)blockIdx.x);
}

template <int TE, int NT, int SLF>
__global__ void __launch_bounds__(NT, 1) gcp2op_fwd_kernel(const __grid_constant__ Gcp2OpParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.M + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = gcp2op_pipe(p, smem, mine);
  wpipe_start<NT>(wp);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    gcp2op_fwd_tile<TE, NT, SLF>(p, smem, tile, wp, tile == (int)blockIdx.x);
}
template <int TE, int NT, int SLF, int SLD>
__global__ void __launch_bounds__(NT, 1) gcp2op_bwd_kernel(const __grid_constant__ Gcp2OpParams p) {
  extern __shared__ __align__(128) float smem[];
  const int ntiles = (p.M + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  WPipe wp = gcp2op_pipe(p, smem, mine);
  wpipe_start<NT>(wp);
  float* prow = p.partial + (size_t)blockIdx.x * p.partial_stride;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    gcp2op_bwd_tile<TE, NT, SLF, SLD>(p, smem, tile, wp, prow, tile != (int)blockIdx.x);
}

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  pack_gcp<256>(p.ops[blockIdx.x], p.blob, (int)blockIdx.y, (int)gridDim.y);
}

// aggregate only (GCPMessagePassing.forward): out[i] = mean/sum over the destination segment, from the sums the edge
// kernel's tiles left behind (segment_total, gcp_tile.cuh)
__global__ void aggregate_kernel(const float* __restrict__ agg, const int* __restrict__ dst_ptr, int N, int W, int edge_rows,
                                 int reduce_mean, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * W) return;
  const int i = (int)(idx / W), f = (int)(idx - (long long)i * W);
  const int a = dst_ptr[i], b = dst_ptr[i + 1];
  float acc = segment_total(agg, N, W, edge_rows, dst_ptr, i, f);
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

// Autoregressive layers: the same sums over the rows of the [2N] gather table (row 2i = node_rep[i], row 2i+1 =
// node_rep_regressive[i]); the direct cotangent of node i (node update, in dir_h / dir_chi) joins row 2i.
__global__ void ar_cotangent_reduce_kernel(float* __restrict__ g_hg, float* __restrict__ g_chig, const float* __restrict__ dir_h,
                                           const float* __restrict__ dir_chi, const float* __restrict__ grow, const float* __restrict__ gcol,
                                           const int* __restrict__ vdst_ptr, const int* __restrict__ vsrc_ptr, const int* __restrict__ vsrc_pos,
                                           int R, int s, int v3) {
  const int W = s + v3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * W) return;
  const int u = (int)(idx / W), f = (int)(idx - (long long)u * W);
  float acc = 0.f;
  if ((u & 1) == 0) acc = f < s ? dir_h[(size_t)(u >> 1) * s + f] : dir_chi[(size_t)(u >> 1) * v3 + (f - s)];
  for (int q = vdst_ptr[u]; q < vdst_ptr[u + 1]; ++q) acc += __ldg(gcol + (size_t)q * W + f);
  for (int q = vsrc_ptr[u]; q < vsrc_ptr[u + 1]; ++q) acc += __ldg(grow + (size_t)__ldg(vsrc_pos + q) * W + f);
  if (f < s) g_hg[(size_t)u * s + f] = acc; else g_chig[(size_t)u * v3 + (f - s)] = acc;
}

// flat parameter gradient = fixed-order sum of the per-CTA partial rows
// `skip`: ranges of the NODE part that the tiles did not produce (their operands were spilled for node_wgrad_kernel, which
// writes those gradients itself): neither read nor written here
struct SkipRanges { int n; int off[2 * MAX_MSG_LAYERS + 2], len[2 * MAX_MSG_LAYERS + 2]; };
__global__ void partial_reduce_kernel(float* __restrict__ out, const float* __restrict__ pe, int ne, int ge,
                                      const float* __restrict__ pn, int nn, int gn, const SkipRanges skip,
                                      const SkipRanges eskip = SkipRanges{}) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ne + nn) return;
  float acc = 0.f;
  if (idx < ne) {
    for (int r = 0; r < eskip.n; ++r)  // produced by the off-tile product over all edges (launch_edge_wgrad)
      if (idx >= eskip.off[r] && idx < eskip.off[r] + eskip.len[r]) return;
    for (int g = 0; g < ge; ++g) acc += __ldg(pe + (size_t)g * ne + idx);
  } else {
    const int j = idx - ne;
    for (int r = 0; r < skip.n; ++r)
      if (j >= skip.off[r] && j < skip.off[r] + skip.len[r]) return;
    for (int g = 0; g < gn; ++g) acc += __ldg(pn + (size_t)g * nn + j);
  }
  out[idx] = acc;
}
// the regions of the node parameter gradient that launch_node_wgrad() produces (scalar_out / vector_out_scale, weight + bias)
static SkipRanges node_wgrad_ranges(const gcpnet_layer& l, const LayerPlan& lp) {
  SkipRanges s{};
  const GcpOp* op[3] = {&lp.ops.ff0, &lp.ops.ff1, l.has_pos ? &lp.ops.pu : nullptr};
  auto add = [&](int off, int len) {  // merge with the previous range when contiguous (weight followed by its bias)
    if (s.n > 0 && s.off[s.n - 1] + s.len[s.n - 1] == off) { s.len[s.n - 1] += len; return; }
    if (s.n < 2 * MAX_MSG_LAYERS + 2) { s.off[s.n] = off; s.len[s.n] = len; ++s.n; }
  };
  if (l.pre_norm) {  // gcp_norm.0 belongs to the standalone normalisation in front of the layer (run_prenorm_backward)
    add(l.ln_grad_off[0] - l.n_edge_params, l.s); add(l.ln_grad_off[1] - l.n_edge_params, l.s);
  }
  for (int k = 0; k < 3; ++k) {
    if (op[k] == nullptr) continue;
    const GcpOp& o = *op[k];
    add(o.o_Ws, o.so * gcp_k(o)); add(o.o_bs, o.so);
    if (o.vo > 0 && gcp_gated(o)) { add(o.o_Wg, o.vo * o.so); add(o.o_bg, o.vo); }
  }
  return s;
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
// opt in to > 48 KB of dynamic shared memory once per (kernel, device, size) -- not on every launch
template <class K>
static int set_smem(K kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& cur = done[{(const void*)kernel, dev}];
  if (bytes > cur) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cur = bytes;
  }
  return 0;
}

template <class K, class P>
static int launch(K kernel, const P& p, int grid, int nt, int bytes, cudaStream_t st) {
  if (set_smem(kernel, bytes)) return 1;
  kernel<<<grid, nt, bytes, st>>>(p);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

#define DISPATCH_EDGE(KERNEL, ...)                                                               \
  do {                                                                                           \
    const int bytes = p.sm.total * 4;                                                            \
    if (tp.TE == 32 && tp.SLF == 1) return launch(KERNEL<32, 256, 1 __VA_ARGS__>, p, tp.grid, 256, bytes, st); \
    if (tp.TE == 32 && tp.SLF == 2) return launch(KERNEL<32, 256, 2 __VA_ARGS__>, p, tp.grid, 256, bytes, st); \
    if (tp.TE == 48 && tp.SLF == 1) return launch(KERNEL<48, 384, 1 __VA_ARGS__>, p, tp.grid, 384, bytes, st); \
    if (tp.TE == 48 && tp.SLF == 2) return launch(KERNEL<48, 384, 2 __VA_ARGS__>, p, tp.grid, 384, bytes, st); \
    if (tp.TE == 64 && tp.SLF == 1) return launch(KERNEL<64, 512, 1 __VA_ARGS__>, p, tp.grid, 512, bytes, st); \
    if (tp.TE == 64 && tp.SLF == 2) return launch(KERNEL<64, 512, 2 __VA_ARGS__>, p, tp.grid, 512, bytes, st); \
    return fail("no kernel instantiation for this edge tile plan");                              \
  } while (0)
#define COMMA_SLD , 2

static int launch_edge_fwd(const EdgeParams& p, const EdgeTilePlan& tp, cudaStream_t st) {
  GcpTimedScope timed(T_EDGE_FWD, st);
  DISPATCH_EDGE(edge_fwd_kernel);
}
static int launch_edge_bwd(const EdgeParams& p, const EdgeTilePlan& tp, cudaStream_t st) {
  GcpTimedScope timed(T_EDGE_BWD, st);
  DISPATCH_EDGE(edge_bwd_kernel, COMMA_SLD);
}
static int launch_node_fwd(const NodeParams& p, const NodeTilePlan& tp, cudaStream_t st) {
  GcpTimedScope timed(T_NODE_FWD, st);
  const int bytes = p.sm.total * 4;
  if (tp.SLF == 1) return launch(node_fwd_kernel<NODE_TE, NODE_NT, 1>, p, tp.grid, NODE_NT, bytes, st);
  if (tp.SLF == 2) return launch(node_fwd_kernel<NODE_TE, NODE_NT, 2>, p, tp.grid, NODE_NT, bytes, st);
  if (tp.SLF == 4) return launch(node_fwd_kernel<NODE_TE, NODE_NT, 4>, p, tp.grid, NODE_NT, bytes, st);
  return fail("no kernel instantiation for this node tile plan");
}
static int launch_node_bwd(const NodeParams& p, const NodeTilePlan& tp, cudaStream_t st) {
  GcpTimedScope timed(T_NODE_BWD, st);
  const int bytes = p.sm.total * 4;
  if (tp.SLF == 1) return launch(node_bwd_kernel<NODE_TE, NODE_NT, 1, NODE_SLD>, p, tp.grid, NODE_NT, bytes, st);
  if (tp.SLF == 2) return launch(node_bwd_kernel<NODE_TE, NODE_NT, 2, NODE_SLD>, p, tp.grid, NODE_NT, bytes, st);
  if (tp.SLF == 4) return launch(node_bwd_kernel<NODE_TE, NODE_NT, 4, NODE_SLD>, p, tp.grid, NODE_NT, bytes, st);
  return fail("no kernel instantiation for this node tile plan");
}
static int launch_pack(const LayerOps& ops, float* blob, cudaStream_t st, bool skip_messages = false, bool skip_node = false) {
  const PackParams pp = make_pack_params(ops, blob, skip_messages, skip_node);
  if (pp.n == 0) return 0;
  pack_kernel<<<dim3(pp.n, 16), 256, 0, st>>>(pp);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static bool tc_option() { return g_opt_tc.load(std::memory_order_relaxed) != 0; }
int gcp_tc_launch_edge_bwd(const tc::TcBwdParams& b, int grid, cudaStream_t st);           // tc_api.cu
int gcp_tc_launch_post(const tc::TcPostParams& p, int which, cudaStream_t st);             // tc_api.cu

// ---- side stream: parameter-gradient post-processing off the critical path ---------------------------------------------
// The next layer's backward only needs dh / dchi; reducing the partials and the chain rule to the reference's parameters
// can overlap with it.  The caller owns the side stream (gcpnet_set_side_stream) and joins it with gcpnet_join() before
// anything consumes the parameter gradients.  Events are created once, outside any stream capture.
// State is kept per device (one process normally drives one GPU; a process that drives several gets one side stream each).
struct SideState {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;       // pool, created once outside any stream capture
  size_t next = 0;
  std::vector<cudaEvent_t> pending;      // recorded on the side stream, not yet joined
};
static std::mutex g_side_mu;
static std::map<int, SideState> g_side_by_dev;
static SideState* side_state() {         // call with g_side_mu held
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  return &g_side_by_dev[dev];
}
static cudaEvent_t side_event(SideState& s) {
  if (s.events.empty()) return nullptr;
  cudaEvent_t e = s.events[s.next % s.events.size()];
  ++s.next;
  return e;
}
int gcp_tc_launch_finalize(const float* partial, int rows, int stride, float* G, const float* npartial, int nrows, int nstride, float* Gn,
                           const tc::TcFinalParams& fp, int which, cudaStream_t st);       // tc_api.cu
int gcp_tc_launch_pack(const tc::TcPackProg& prog, cudaStream_t st);                      // tc_api.cu
int gcp_tc_launch_edge_fwd(const tc::TcEdgeParams& p, int grid, cudaStream_t st);          // tc_api.cu
int gcp_tc_launch_node_pre(const tc::TcEdgeParams& p, float* P, float* Q, cudaStream_t st);  // tc_api.cu
static int launch_tc_pack(const LayerPlan& lp, float* packed, cudaStream_t st) {
  tc::TcPackProg prog = lp.tc.pack;
  prog.blob = packed + lp.v2_packed_floats;
  return gcp_tc_launch_pack(prog, st);
}
static void fill_tc_common(tc::TcEdgeParams& p, const gcpnet_graph& g, const LayerPlan& lp, const float* h, const float* chi, const float* e,
                           const float* xi, const float* frames, const float* packed) {
  p.N = (int)g.num_nodes; p.E = (int)g.num_edges;
  p.h = h; p.chi = chi; p.e = e; p.xi = xi; p.frames = frames;
  p.perm = g.perm; p.src = g.src; p.dst = g.dst; p.dst_ptr = g.dst_ptr;
  p.blob = packed + lp.v2_packed_floats;
  const float* pq = packed + lp.v2_packed_floats + tc::rup(lp.tc.blob_floats, 32);  // per-node products of message GCP 0
  p.P = pq; p.Q = pq + (size_t)p.N * 2 * p.pw;
}
static int launch_tc_edge_fwd(const gcpnet_graph& g, const LayerPlan& lp, const gcpnet_forward_io& io, float* saved, cudaStream_t st) {
  GcpTimedScope timed(T_EDGE_FWD, st);
  tc::TcEdgeParams p = lp.tc.proto;
  fill_tc_common(p, g, lp, io.h, io.chi, io.e, io.xi, io.frames, io.packed);
  p.agg = io.agg; p.saved = saved;
  p.dbg = g_tc_dbg.load(std::memory_order_relaxed);
  if (gcp_tc_launch_node_pre(p, const_cast<float*>(p.P), const_cast<float*>(p.Q), st)) return 1;
  return gcp_tc_launch_edge_fwd(p, lp.tc.grid, st);
}
// fork: work that only feeds the PARAMETER gradient runs on the side stream (if the caller gave one and nobody is timing
// kernels); returns the stream to use (st itself when there is no side stream)
static cudaStream_t fork_side(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState* s = side_state();
  if (s == nullptr || s->stream == nullptr || gcp_profile_on()) return st;
  cudaEvent_t ev = side_event(*s);
  if (ev == nullptr) return st;
  if (cudaEventRecord(ev, st) != cudaSuccess || cudaStreamWaitEvent(s->stream, ev, 0) != cudaSuccess) return st;
  return s->stream;
}
// the side stream `ps` additionally waits for what has been enqueued on `st` so far
static int side_wait_main(cudaStream_t ps, cudaStream_t st) {
  if (ps == st) return 0;
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState* s = side_state();
  cudaEvent_t ev = s ? side_event(*s) : nullptr;
  if (ev == nullptr) return fail("side stream: out of events");
  CUDA_TRY(cudaEventRecord(ev, st));
  CUDA_TRY(cudaStreamWaitEvent(ps, ev, 0));
  return 0;
}
static int side_done(cudaStream_t ps, cudaStream_t st) {
  if (ps == st) return 0;
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState* s = side_state();
  cudaEvent_t ev = s ? side_event(*s) : nullptr;
  if (ev == nullptr) return fail("side stream: out of events");
  CUDA_TRY(cudaEventRecord(ev, ps));
  s->pending.push_back(ev);
  return 0;
}

// weight gradients of the node GCPs' scalar_out / vector_out_scale from the rows the node backward spilled: one
// output-parallel kernel over all nodes; it writes exactly the regions of the flat gradient that partial_reduce_kernel
// skips (node_wgrad_ranges)
static int launch_node_wgrad(const gcpnet_layer& l, const gcpnet_graph& g, const LayerPlan& lp, const float* ws_node_partial,
                             const float* saved_node, float* g_node_params, cudaStream_t st) {
  const long long N = g.num_nodes;
  const NodeSpill sp = node_spill_layout(N, lp.ops, l.has_pos != 0);
  const float* spill = ws_node_partial + (size_t)node_spill_offset(lp.nb.grid, l.n_node_params);
  const NodeSavedLayout sv = node_saved_layout((int)N, l.s, l.v, l.ff0.so, l.ff0.vo, l.has_pos != 0, l.training != 0);
  const GcpOp* op[3] = {&lp.ops.ff0, &lp.ops.ff1, l.has_pos ? &lp.ops.pu : nullptr};
  const long long tsaved[3] = {sv.T0, sv.T1, sv.TP};
  NodeWgradParams p{};
  p.N = (int)N; p.slope = l.slope;
  int cta = 0;
  auto add = [&](const float* G, int ldg, int J, const float* Z, int ldz, int I, int act, float* outW, float* outb) {
    NodeWgradJob& j = p.job[p.njobs++];
    j.G = G; j.ldg = ldg; j.J = J; j.Z = Z; j.ldz = ldz; j.I = I; j.act = act; j.outW = outW; j.outb = outb;
    j.JB = (J + 15) / 16; j.IG = ((I + 7) / 8 + 3) / 4; j.cta0 = cta;
    cta += j.JB * j.IG;
  };
  for (int k = 0; k < 3; ++k) {
    if (op[k] == nullptr) continue;
    const GcpOp& o = *op[k];
    add(spill + sp.gT[k], sp.ldg[k], o.so, spill + sp.Z[k], sp.ldz[k], gcp_k(o), ACT_NONE, g_node_params + o.o_Ws, g_node_params + o.o_bs);
    if (o.vo > 0 && gcp_gated(o))
      add(spill + sp.GG[k], sp.ldgg[k], o.vo, saved_node + tsaved[k], o.so, o.so, o.act_v, g_node_params + o.o_Wg, g_node_params + o.o_bg);
  }
  node_wgrad_kernel<<<cta, 256, 0, st>>>(p);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// FFMA edge backward with spilled operands: scalar_out / vector_out_scale weight gradients of every message GCP as one
// output-parallel product over all edges, written straight into the flat gradient
static SkipRanges edge_wgrad_ranges(const LayerPlan& lp) {
  SkipRanges s{};
  for (int k = 0; k < lp.ops.L; ++k) {
    const GcpOp& o = lp.ops.msg[k];
    s.off[s.n] = o.o_Ws; s.len[s.n] = o.so * gcp_k(o) + o.so; ++s.n;  // weight followed by its bias
    if (gcp_gated(o)) { s.off[s.n] = o.o_Wg; s.len[s.n] = o.vo * o.so + o.vo; ++s.n; }
  }
  return s;
}
static int launch_edge_wgrad(const gcpnet_layer& l, const gcpnet_graph& g, const LayerPlan& lp, float* spill, const float* saved_edge,
                             float* g_edge_params, cudaStream_t st) {
  const long long E = g.num_edges;
  const EdgeSpill sp = edge_spill_layout(E, lp.ops);
  long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS], tot;
  edge_saved_offsets(l, E, offT, offG, offS, offV, &tot);
  NodeWgradParams p{};
  p.N = (int)E; p.slope = l.slope;
  p.chunk_rows = EDGE_WGRAD_CHUNK_ROWS; p.nchunks = sp.nchunks; p.out_total = sp.out_total; p.scratch = spill + sp.scratch;
  int cta = 0, out0 = 0;
  auto add = [&](const float* G, int ldg, int J, const float* Z, int ldz, int I, int act, float* outW, float* outb) {
    NodeWgradJob& j = p.job[p.njobs++];
    j.G = G; j.ldg = ldg; j.J = J; j.Z = Z; j.ldz = ldz; j.I = I; j.act = act; j.outW = outW; j.outb = outb;
    j.JB = (J + 15) / 16; j.IG = ((I + 7) / 8 + 3) / 4; j.cta0 = cta; j.out0 = out0;
    cta += j.JB * j.IG;
    out0 += J * I + J;
  };
  for (int k = 0; k < lp.ops.L; ++k) {
    const GcpOp& o = lp.ops.msg[k];
    add(spill + sp.gT[k], sp.ldg[k], o.so, spill + sp.Z[k], sp.ldz[k], gcp_k(o), ACT_NONE, g_edge_params + o.o_Ws, g_edge_params + o.o_bs);
    if (gcp_gated(o))  // vector_out_scale reads act_v(T): the saved pre-activations of the forward pass
      add(spill + sp.GG[k], sp.ldgg[k], o.vo, saved_edge + offT[k], o.so, o.so, o.act_v, g_edge_params + o.o_Wg, g_edge_params + o.o_bg);
  }
  if (out0 != sp.out_total) return fail("edge_wgrad: output layout mismatch");
  node_wgrad_kernel<<<dim3(cta, p.nchunks), 256, 0, st>>>(p);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  if (p.nchunks > 1) {
    wgrad_chunk_reduce_kernel<<<(p.out_total + 255) / 256, 256, 0, st>>>(p);
    gcp_note_launches(1);
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

// tensor-core edge backward + node-level finish of message GCP 0 + chain rule to the reference's parameters
static int run_tc_edge_backward(const gcpnet_layer& l, const gcpnet_graph& g, const LayerPlan& lp, const gcpnet_backward_io& io, const float* gagg,
                                cudaStream_t st) {
  const tc::TcPlan& T = lp.tc;
  float* Y = io.ws_edge;
  float* A = Y + T.y_floats;
  float* G = A + T.a_floats;
  float* Gn = G + T.bproto.partial_stride;
  float* npart = Gn + T.node_partial_stride;
  {
    GcpTimedScope timed(T_EDGE_BWD, st);
    tc::TcBwdParams b = T.bproto;
    fill_tc_common(b.f, g, lp, io.h, io.chi, io.e, io.xi, io.frames, io.packed);
    b.f.saved = const_cast<float*>(io.saved_edge);
    b.f.agg = nullptr; b.f.dbg = g_tc_dbg.load(std::memory_order_relaxed);
    b.gagg = gagg; b.dst_ptr = g.dst_ptr;
    b.ge = io.g_e; b.gxi = io.g_xi; b.Y = Y; b.partial = io.ws_edge_partial;
    if (gcp_tc_launch_edge_bwd(b, T.grid, st)) return 1;
  }
  tc::TcPostParams pp{};
  pp.N = (int)g.num_nodes; pp.s = l.s; pp.v = l.v; pp.pw = T.proto.pw; pp.rows = T.proto.rows;
  pp.Y = Y; pp.y_img_g = T.bproto.y_img_g; pp.y_img_v = T.bproto.y_img_v;
  pp.dst_ptr = g.dst_ptr; pp.src_ptr = g.src_ptr; pp.src_pos = g.src_pos;
  pp.h = io.h; pp.chi = io.chi; pp.blob = io.packed + lp.v2_packed_floats; pp.nt = T.proto.nt;
  pp.A = A; pp.g_h = io.g_h; pp.g_chi = io.g_chi;
  pp.npartial = npart; pp.npartial_stride = T.node_partial_stride; pp.nctas = T.node_partial_ctas;
  // Node-level finish: per-node sums of the per-edge cotangents, then dh / dchi (the next layer's backward waits for
  // these) on the caller's stream; everything that only feeds the PARAMETER gradient on the side stream (if the caller
  // gave one).  Measured at cfg2 with the main chain on a high-priority stream (GraphedStep): mode 2 is 0.7 % faster per
  // step than modes 0 / 1; with equal priorities side work that overlaps the sums slows the critical path and mode 0 wins.
  // option "post_fused": 0 = sums -> fork -> dh/dchi || side work;  1 = fused sums + dh/dchi -> fork;
  //                      2 = fork -> fused sums + dh/dchi || reduction of the edge partial rows, rest of the side work after the sums
  const int mode = g_opt_post_fused.load(std::memory_order_relaxed);
  cudaStream_t ps = st;
  if (mode == 2) ps = fork_side(st);
  {
    GcpTimedScope timed(T_COT_REDUCE, st);
    if (gcp_tc_launch_post(pp, mode == 0 ? 1 : 3, st)) return 1;
  }
  if (mode != 2) ps = fork_side(st);
  if (mode == 0) {
    GcpTimedScope timed(T_COT_REDUCE, st);
    if (gcp_tc_launch_post(pp, 2, st)) return 1;
  }
  tc::TcFinalParams fp{};
  fp.L = l.num_message_layers; fp.s = l.s; fp.v = l.v; fp.se = l.se; fp.ve = l.ve; fp.pw = T.proto.pw; fp.n_edge_params = l.n_edge_params;
  fp.G = G; fp.Gn = Gn; fp.out = io.g_params;
  for (int k = 0; k < fp.L; ++k) {
    const gcpnet_gcp2& d = l.message[k];
    const tc::TcGcp& tg = T.proto.g[k];
    tc::TcFinalGcp& f = fp.g[k];
    f.si = d.si; f.vi = d.vi; f.so = d.so; f.vo = d.vo; f.hd = d.hd; f.nslot = tg.nslot; f.zc0 = tg.zc0; f.kz = tg.kz;
    f.off_tg = T.bproto.off_tg[k]; f.off_v = T.bproto.off_v[k];
    for (int i = 0; i < 7; ++i) f.grad_off[i] = d.grad_off[i];
    f.Wd = d.vector_down; f.Ws = d.scalar_out_w; f.bs = d.scalar_out_b; f.Wu = d.vector_up; f.Wg = d.vector_out_scale_w;
  }
  {
    GcpTimedScope timed(T_PARTIAL_REDUCE, ps);
    if (mode == 2) {
      if (gcp_tc_launch_finalize(io.ws_edge_partial, T.grid, T.bproto.partial_stride, G, npart, T.node_partial_ctas, T.node_partial_stride, Gn, fp, 1, ps))
        return 1;
      if (side_wait_main(ps, st)) return 1;  // the per-node sums (main stream) feed the node-level products
    }
    if (gcp_tc_launch_post(pp, 4, ps)) return 1;
    if (gcp_tc_launch_finalize(io.ws_edge_partial, T.grid, T.bproto.partial_stride, G, npart, T.node_partial_ctas, T.node_partial_stride, Gn, fp, mode == 2 ? 2 : 3, ps))
      return 1;
  }
  return side_done(ps, st);
}

// consistency of the optional variants (autoregressive gather views, pre_norm workspace)
static const char* check_variant(const gcpnet_layer& l, const gcpnet_graph& g, const float* h_gather, const float* chi_gather, const float* prenorm) {
  const bool ar = g.num_gather_rows > 0;
  if (ar != (l.autoregressive == 1)) return "autoregressive layers need the views of gcpnet_graph_build_autoregressive (and only they)";
  // autoregressive == 2: aggregate_with_row (gcpnet.py:946): the graph was built on the flipped edge_index and the gather
  // ids swap the two ends back (gsrc = dst, gdst = src)
  if ((g.gsrc != nullptr && !ar) != (l.autoregressive == 2)) return "gather ids without gather rows need aggregate_with_row (autoregressive = 2)";
  if (ar && (!g.gdst || !g.vdst_ptr || !g.vsrc_ptr || !g.vsrc_pos || g.num_gather_rows != 2 * g.num_nodes)) return "incomplete autoregressive graph views";
  if (ar && (!h_gather || !chi_gather)) return "autoregressive layers need the [2N] gather table";
  if (ar && l.pre_norm) return "pre_norm with an autoregressive gather table is not covered";
  if (l.pre_norm && !prenorm) return "pre_norm layers need the prenorm workspace";
  if (l.enable_e3 && g.node_mask != nullptr) return "enable_e3_equivariance with a node mask is not covered";
  return nullptr;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int gcpnet_version(void) { return 200; }
const char* gcpnet_last_error(void) { return g_last_error.c_str(); }
void gcpnet_profile_enable(int on) { g_profile.store(on != 0); }
void gcpnet_debug_stamps(long long* device_buffer) { g_tc_dbg.store(device_buffer); }
int gcpnet_set_side_stream(void* stream) {
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState* s = side_state();
  if (s == nullptr) return fail("set_side_stream: no current device");
  s->stream = (cudaStream_t)stream;
  if (s->stream != nullptr && s->events.empty()) {
    for (int i = 0; i < 256; ++i) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      s->events.push_back(e);
    }
  }
  return 0;
}
int gcpnet_join(void* stream) {
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState* s = side_state();
  if (s == nullptr) return 0;
  for (cudaEvent_t e : s->pending) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, e, 0));
  s->pending.clear();
  return 0;
}
int gcpnet_set_option(const char* name, int value) {
  if (name && std::string(name) == "tc") return g_opt_tc.exchange(value);
  if (name && std::string(name) == "post_fused") return g_opt_post_fused.exchange(value);
  if (name && std::string(name) == "early_fork") return g_opt_early_fork.exchange(value);
  return -1;
}
int gcpnet_profile_read(int which, double* total_ms, int64_t* launches) {
  if (which < 0 || which >= T_COUNT || !total_ms || !launches) return fail("profile_read: bad argument");
  std::lock_guard<std::mutex> lock(g_profile_mu);
  std::vector<cudaEvent_t>& v = g_profile_ev[which];
  double tot = 0.0; int64_t n = 0;
  for (size_t i = 0; i + 1 < v.size(); i += 2) {
    CUDA_TRY(cudaEventSynchronize(v[i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
    tot += ms; ++n;
  }
  for (cudaEvent_t e : v) cudaEventDestroy(e);
  v.clear();
  *total_ms = tot; *launches = n;
  return 0;
}
uint64_t gcpnet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int gcpnet_layer_plan(const gcpnet_layer* layer, int64_t N, int64_t E, gcpnet_plan* plan) {
  if (!layer || !plan) return fail("layer_plan: null argument");
  LayerPlan lp;
  const std::string e = make_layer_plan(*layer, N, E, &lp, plan, tc_option());
  if (!e.empty()) return fail("layer_plan: " + e);
  return 0;
}

static int run_edge_forward(const gcpnet_layer& l, const gcpnet_graph& g, const LayerPlan& lp, const gcpnet_forward_io& io,
                            cudaStream_t st) {
  if (g.num_edges == 0) return 0;
  if (lp.tc.ok) {  // tensor-core path (the plan decided; workspaces are sized for it)
    if (!io.packed_ready && launch_tc_pack(lp, io.packed, st)) return 1;
    return launch_tc_edge_fwd(g, lp, io, io.saved_edge, st);
  }
  EdgeParams p = make_edge_params(l, g, lp.ops, lp.ef, false, io.packed);
  p.h = io.h_gather ? io.h_gather : io.h; p.chi = io.chi_gather ? io.chi_gather : io.chi;
  p.e = io.e; p.xi = io.xi; p.frames = io.frames;
  p.agg = io.agg; p.saved = io.saved_edge;
  return launch_edge_fwd(p, lp.ef, st);
}

int gcpnet_layer_pack(const gcpnet_layer* layer, const gcpnet_plan* plan, int64_t N, int64_t E, float* packed, void* stream) {
  if (!layer || !plan || !packed) return fail("layer_pack: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  LayerPlan lp;
  const std::string e = make_layer_plan(*layer, N, E, &lp, nullptr, plan->tc_edge_path != 0);
  if (!e.empty()) return fail("layer_pack: " + e);
  if (launch_pack(lp.ops, packed, st, lp.tc.ok)) return 1;
  if (lp.tc.ok && E > 0 && launch_tc_pack(lp, packed, st)) return 1;
  return 0;
}

int gcpnet_layer_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                         const gcpnet_forward_io* io, void* stream) {
  if (!layer || !graph || !plan || !io) return fail("layer_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gcpnet_layer& l = *layer;
  if (l.has_pos && (!io->pos || !io->out_pos)) return fail("layer_forward: node positions required");
  if (!io->packed) return fail("layer_forward: packed-weight workspace required");
  LayerPlan lp;
  const std::string e = make_layer_plan(l, graph->num_nodes, graph->num_edges, &lp, nullptr, plan->tc_edge_path != 0);
  if (!e.empty()) return fail("layer_forward: " + e);
  if (graph->num_nodes <= 0) return 0;
  if (const char* m = check_variant(l, *graph, io->h_gather, io->chi_gather, io->prenorm)) return fail(std::string("layer_forward: ") + m);
  if (!io->packed_ready && launch_pack(lp.ops, io->packed, st, lp.tc.ok)) return 1;
  gcpnet_forward_io f = *io;
  if (l.pre_norm) {  // gcp_norm.0 on the layer input (gcpnet.py:1188-1189); everything below reads the normalised copy
    const long long N = graph->num_nodes;
    float* hn = io->prenorm; float* chin = io->prenorm + N * l.s;
    gcp_layernorm_fwd_kernel<<<(int)((N * 32 + 255) / 256), 256, 0, st>>>(io->h, io->chi, (int)N, l.s, l.v, l.ln0_w, l.ln0_b, l.ln_eps, l.vn_eps, hn, chin);
    gcp_note_launches(1);
    CUDA_TRY(cudaGetLastError());
    f.h = hn; f.chi = chin;
  }
  if (run_edge_forward(l, *graph, lp, f, st)) return 1;
  NodeParams p = make_node_params(l, *graph, lp.ops, lp.nf, false, f.packed);
  p.h = f.h; p.chi = f.chi; p.agg = f.agg; p.pos = f.pos; p.frames = f.frames;
  p.edge_rows = edge_tile_rows(lp);
  p.out_h = f.out_h; p.out_chi = f.out_chi; p.out_pos = f.out_pos; p.saved = f.saved_node;
  return launch_node_fwd(p, lp.nf, st);
}

int gcpnet_message_passing_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                                   const gcpnet_forward_io* io, float* aggregate, void* stream) {
  if (!layer || !graph || !plan || !io || !aggregate) return fail("message_passing_forward: null argument");
  if (const char* m = check_variant(*layer, *graph, io->h_gather, io->chi_gather, io->prenorm)) return fail(std::string("message_passing_forward: ") + m);
  cudaStream_t st = (cudaStream_t)stream;
  if (!io->packed) return fail("message_passing_forward: packed-weight workspace required");
  LayerPlan lp;
  const std::string e = make_layer_plan(*layer, graph->num_nodes, graph->num_edges, &lp, nullptr, plan->tc_edge_path != 0);
  if (!e.empty()) return fail("message_passing_forward: " + e);
  if (graph->num_nodes <= 0) return 0;
  if (launch_pack(lp.ops, io->packed, st, lp.tc.ok, true)) return 1;
  if (run_edge_forward(*layer, *graph, lp, *io, st)) return 1;
  const int W = layer->s + 3 * layer->v;
  const long long tot = graph->num_nodes * W;
  aggregate_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(io->agg, graph->dst_ptr, (int)graph->num_nodes, W,
                                                            edge_tile_rows(lp), layer->reduce_mean, aggregate);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// FFMA-tile edge backward + per-node sums of the gathered cotangents (accumulated into g_h / g_chi)
static int run_ffma_edge_backward(const gcpnet_layer& l, const gcpnet_graph& g, const LayerPlan& lp, const gcpnet_backward_io& io,
                                  const float* gagg, cudaStream_t st, int* edge_grid) {
  *edge_grid = 0;
  const int W = l.s + 3 * l.v;
  if (g.num_edges <= 0) {
    if (g.num_gather_rows > 0) {  // no edges: the gather-table cotangent is the direct part on the even rows
      const long long tot = g.num_gather_rows * W;
      ar_cotangent_reduce_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
          io.g_h_gather, io.g_chi_gather, io.g_h, io.g_chi, nullptr, nullptr, g.vdst_ptr, g.vsrc_ptr, g.vsrc_pos, (int)g.num_gather_rows, l.s, 3 * l.v);
      gcp_note_launches(1);
      CUDA_TRY(cudaGetLastError());
    }
    return 0;
  }
  EdgeParams ep = make_edge_params(l, g, lp.ops, lp.eb, true, io.packed);
  ep.h = io.h_gather ? io.h_gather : io.h; ep.chi = io.chi_gather ? io.chi_gather : io.chi;
  ep.e = io.e; ep.xi = io.xi; ep.frames = io.frames;
  ep.saved = const_cast<float*>(io.saved_edge);
  ep.gagg = gagg;
  ep.grow = io.ws_edge; ep.gcol = io.ws_edge + (size_t)g.num_edges * W;
  ep.ge = io.g_e; ep.gxi = io.g_xi;
  ep.partial = io.ws_edge_partial;
  if (io.ws_edge_spill != nullptr) {
    const EdgeSpill sp = edge_spill_layout(g.num_edges, lp.ops);
    ep.spill = io.ws_edge_spill;
    for (int k = 0; k < lp.ops.L; ++k) { ep.sp_gT[k] = sp.gT[k]; ep.sp_Z[k] = sp.Z[k]; ep.sp_GG[k] = sp.GG[k]; }
  }
  ep.dbg = g_tc_dbg.load(std::memory_order_relaxed);
  *edge_grid = lp.eb.grid;
  if (launch_edge_bwd(ep, lp.eb, st)) return 1;
  GcpTimedScope timed(T_COT_REDUCE, st);
  if (g.num_gather_rows == 0 && g.gsrc != nullptr) {
    // aggregate_with_row: the "row" features were gathered by DESTINATION of the flipped graph and vice versa
    const long long tot = g.num_nodes * W;
    node_cotangent_reduce_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
        io.g_h, io.g_chi, ep.gcol, ep.grow, g.dst_ptr, g.src_ptr, g.src_pos, (int)g.num_nodes, l.s, 3 * l.v);
  } else if (g.gsrc != nullptr) {
    const long long tot = g.num_gather_rows * W;
    ar_cotangent_reduce_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
        io.g_h_gather, io.g_chi_gather, io.g_h, io.g_chi, ep.grow, ep.gcol, g.vdst_ptr, g.vsrc_ptr, g.vsrc_pos, (int)g.num_gather_rows, l.s, 3 * l.v);
  } else {
    const long long tot = g.num_nodes * W;
    node_cotangent_reduce_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
        io.g_h, io.g_chi, ep.grow, ep.gcol, g.dst_ptr, g.src_ptr, g.src_pos, (int)g.num_nodes, l.s, 3 * l.v);
  }
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int layer_backward_body(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                               const gcpnet_backward_io* io, void* stream);
int gcpnet_layer_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                          const gcpnet_backward_io* io, void* stream) {
  if (!layer || !graph || !plan || !io) return fail("layer_backward: null argument");
  const gcpnet_layer& l = *layer;
  if (graph->num_nodes > 0)
    if (const char* m = check_variant(l, *graph, io->h_gather, io->chi_gather, io->prenorm)) return fail(std::string("layer_backward: ") + m);
  if (graph->gsrc != nullptr && (!io->g_h_gather || !io->g_chi_gather)) return fail("layer_backward: autoregressive layers need g_h_gather / g_chi_gather");
  if (!l.pre_norm || graph->num_nodes <= 0) return layer_backward_body(layer, graph, plan, io, stream);
  // pre_norm: the layer proper ran on the normalised input; its input cotangent lands in the workspace, then the
  // standalone gcp_norm.0 backward produces the caller's g_h / g_chi and the gcp_norm.0 parameter gradients
  if (!io->ws_prenorm) return fail("layer_backward: pre_norm layers need ws_prenorm");
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = graph->num_nodes, W = l.s + 3 * l.v;
  gcpnet_backward_io b = *io;
  b.h = io->prenorm; b.chi = io->prenorm + N * l.s;
  b.g_h = io->ws_prenorm; b.g_chi = io->ws_prenorm + N * l.s;
  if (layer_backward_body(layer, graph, plan, &b, stream)) return 1;
  float* stats = io->ws_prenorm + N * W;
  float* partial = stats + 2 * N;
  gcp_layernorm_bwd_kernel<<<(int)((N * 32 + 255) / 256), 256, 0, st>>>(io->h, io->chi, (int)N, l.s, l.v, l.ln0_w, l.ln_eps, l.vn_eps, b.g_h, b.g_chi,
                                                                       io->g_h, io->g_chi, stats);
  gcp_layernorm_wgrad_kernel<<<LN_PARTS, 128, 0, st>>>(io->h, b.g_h, stats, (int)N, l.s, partial);
  gcp_layernorm_wreduce_kernel<<<(2 * l.s + 127) / 128, 128, 0, st>>>(partial, LN_PARTS, l.s, io->g_params + l.ln_grad_off[0], io->g_params + l.ln_grad_off[1]);
  gcp_note_launches(3);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int layer_backward_body(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                               const gcpnet_backward_io* io, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const gcpnet_layer& l = *layer;
  const gcpnet_graph& g = *graph;
  if (!io->saved_edge && g.num_edges > 0) return fail("layer_backward: forward ran without saved activations");
  if (!io->saved_node) return fail("layer_backward: forward ran without saved activations");
  if (!io->packed) return fail("layer_backward: packed weights of the forward call required");
  LayerPlan lp;
  const std::string e = make_layer_plan(l, g.num_nodes, g.num_edges, &lp, nullptr, plan->tc_edge_path != 0);
  if (!e.empty()) return fail("layer_backward: " + e);
  if (g.num_nodes <= 0) return 0;
  const int W = l.s + 3 * l.v;
  // node update backward -> direct cotangent of (h, chi) in g_h/g_chi, cotangent of the aggregate in ws_agg
  NodeParams np = make_node_params(l, g, lp.ops, lp.nb, true, io->packed);
  np.dbg = g_tc_dbg.load(std::memory_order_relaxed);
  np.saved = const_cast<float*>(io->saved_node);
  np.h = io->h; np.chi = io->chi;  // node mask: masked-out nodes fed the layer input to the position GCP
  np.frames = io->frames;
  np.g_out_h = io->g_out_h; np.g_out_chi = io->g_out_chi; np.g_out_pos = io->g_out_pos;
  np.g_x_h = io->g_h; np.g_x_chi = io->g_chi; np.g_agg = io->ws_agg;
  np.partial = io->ws_node_partial;
  {
    const NodeSpill sp = node_spill_layout(g.num_nodes, lp.ops, l.has_pos != 0);
    np.spill = io->ws_node_partial + (size_t)node_spill_offset(lp.nb.grid, l.n_node_params);
    for (int k = 0; k < 3; ++k) { np.sp_gT[k] = sp.gT[k]; np.sp_Z[k] = sp.Z[k]; np.sp_GG[k] = sp.GG[k]; }
  }
  if (launch_node_bwd(np, lp.nb, st)) return 1;
  if (g.num_edges > 0 && lp.tc.ok) {
    {  // node parameter gradients: ready as soon as the node backward is done -> overlap with the edge backward
      const cudaStream_t ps = g_opt_early_fork.load(std::memory_order_relaxed) ? fork_side(st) : st;
      GcpTimedScope timed(T_PARTIAL_REDUCE, ps);
      partial_reduce_kernel<<<(l.n_node_params + 255) / 256, 256, 0, ps>>>(io->g_params + l.n_edge_params, nullptr, 0, 0, io->ws_node_partial,
                                                                            l.n_node_params, lp.nb.grid, node_wgrad_ranges(l, lp));
      gcp_note_launches(1);
      CUDA_TRY(cudaGetLastError());
      if (launch_node_wgrad(l, g, lp, io->ws_node_partial, io->saved_node, io->g_params + l.n_edge_params, ps)) return 1;
      if (side_done(ps, st)) return 1;
    }
    return run_tc_edge_backward(l, g, lp, *io, io->ws_agg, st);
  }
  int edge_grid = 0;
  if (run_ffma_edge_backward(l, g, lp, *io, io->ws_agg, st, &edge_grid)) return 1;
  const int np_tot = l.n_edge_params + l.n_node_params;
  const bool spilled = io->ws_edge_spill != nullptr && g.num_edges > 0;
  // everything below only feeds the parameter gradient: side stream when the caller gave one
  cudaStream_t ps = fork_side(st);
  {
    GcpTimedScope timed(T_PARTIAL_REDUCE, ps);
    partial_reduce_kernel<<<(np_tot + 255) / 256, 256, 0, ps>>>(io->g_params, io->ws_edge_partial, l.n_edge_params, edge_grid,
                                                             io->ws_node_partial, l.n_node_params, lp.nb.grid, node_wgrad_ranges(l, lp),
                                                             spilled ? edge_wgrad_ranges(lp) : SkipRanges{});
    gcp_note_launches(1);
    CUDA_TRY(cudaGetLastError());
    if (launch_node_wgrad(l, g, lp, io->ws_node_partial, io->saved_node, io->g_params + l.n_edge_params, ps)) return 1;
    if (spilled && launch_edge_wgrad(l, g, lp, io->ws_edge_spill, io->saved_edge, io->g_params, ps)) return 1;
  }
  if (ps != st && side_done(ps, st)) return 1;
  return 0;
}

// ---- GCP2 / GCPLayerNorm on their own ---------------------------------------------------------------------------------------
int gcpnet_gcp2_plan_query(const gcpnet_gcp2* op, int64_t M, gcpnet_gcp2_plan* plan) {
  if (!op || !plan) return fail("gcp2_plan: null argument");
  const Gcp2OpPlan P = plan_gcp2_op(*op, M);
  if (!P.error.empty()) return fail("gcp2_plan: " + P.error);
  gcpnet_gcp2_plan q{};
  q.tile = GCP2OP_TE; q.grid = P.grid; q.smem_fwd_bytes = P.smf.total * 4; q.smem_bwd_bytes = P.smb.total * 4;
  q.n_params = P.n_params; q.packed_floats = P.packed_floats;
  q.saved_floats = M * (long long)(op->so + op->vo); q.partial_floats = (long long)P.grid * P.n_params;
  *plan = q;
  return 0;
}
static Gcp2OpParams gcp2op_params(const Gcp2OpPlan& P, bool backward, int64_t M, const float* s_in, const float* v_in, const float* frames,
                                  int e3, float slope, const float* blob) {
  Gcp2OpParams p{};
  p.M = (int)M; p.e3 = e3; p.slope = slope; p.s_in = s_in; p.v_in = v_in; p.frames = frames; p.blob = blob;
  p.op = P.op; p.sm = backward ? P.smb : P.smf; p.seq = backward ? P.bwd : P.fwd;
  return p;
}
int gcpnet_gcp2_forward(const gcpnet_gcp2* op, int64_t M, const float* s_in, const float* v_in, const float* frames, int e3, float slope,
                        float* s_out, float* v_out, float* saved, float* packed, void* stream) {
  if (!op || !s_in || !v_in || !frames || !s_out || (!v_out && op->vo > 0) || !packed) return fail("gcp2_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const Gcp2OpPlan P = plan_gcp2_op(*op, M);
  if (!P.error.empty()) return fail("gcp2_forward: " + P.error);
  if (M <= 0) return 0;
  PackParams pp{};
  pp.n = 1; pp.blob = packed; pp.ops[0] = P.op;
  pack_kernel<<<dim3(1, 16), 256, 0, st>>>(pp);
  gcp_note_launches(1);
  Gcp2OpParams p = gcp2op_params(P, false, M, s_in, v_in, frames, e3, slope, packed);
  p.s_out = s_out; p.v_out = v_out; p.saved = saved;
  const int bytes = p.sm.total * 4;
  if (P.slf == 1) return launch(gcp2op_fwd_kernel<GCP2OP_TE, GCP2OP_NT, 1>, p, P.grid, GCP2OP_NT, bytes, st);
  return launch(gcp2op_fwd_kernel<GCP2OP_TE, GCP2OP_NT, 2>, p, P.grid, GCP2OP_NT, bytes, st);
}
int gcpnet_gcp2_backward(const gcpnet_gcp2* op, int64_t M, const float* s_in, const float* v_in, const float* frames, int e3, float slope,
                         const float* saved, const float* packed, const float* g_s_out, const float* g_v_out, float* g_s_in,
                         float* g_v_in, float* g_params, float* ws_partial, void* stream) {
  if (!op || !s_in || !v_in || !frames || !saved || !packed || !g_s_out || (!g_v_out && op->vo > 0) || !g_s_in || !g_v_in || !g_params || !ws_partial)
    return fail("gcp2_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const Gcp2OpPlan P = plan_gcp2_op(*op, M);
  if (!P.error.empty()) return fail("gcp2_backward: " + P.error);
  if (M <= 0) { CUDA_TRY(cudaMemsetAsync(g_params, 0, (size_t)P.n_params * sizeof(float), st)); return 0; }
  Gcp2OpParams p = gcp2op_params(P, true, M, s_in, v_in, frames, e3, slope, packed);
  p.saved = const_cast<float*>(saved); p.gs_out = g_s_out; p.gv_out = g_v_out; p.gs_in = g_s_in; p.gv_in = g_v_in;
  p.partial = ws_partial; p.partial_stride = P.n_params;
  const int bytes = p.sm.total * 4;
  int rc;
  if (P.slf == 1) rc = launch(gcp2op_bwd_kernel<GCP2OP_TE, GCP2OP_NT, 1, EDGE_SLD>, p, P.grid, GCP2OP_NT, bytes, st);
  else rc = launch(gcp2op_bwd_kernel<GCP2OP_TE, GCP2OP_NT, 2, EDGE_SLD>, p, P.grid, GCP2OP_NT, bytes, st);
  if (rc) return rc;
  partial_reduce_kernel<<<(P.n_params + 255) / 256, 256, 0, st>>>(g_params, ws_partial, P.n_params, P.grid, nullptr, 0, 0, SkipRanges{});
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int gcpnet_layernorm_forward(const float* h, const float* chi, int64_t N, int32_t s, int32_t v, const float* w, const float* b,
                             float* out_h, float* out_chi, void* stream) {
  if (N <= 0) return 0;
  if ((s > 0 && (!h || !w || !b || !out_h)) || (v > 0 && (!chi || !out_chi))) return fail("layernorm_forward: null argument");
  gcp_layernorm_fwd_kernel<<<(int)((N * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, chi, (int)N, s, v, w, b, 1e-5f, 1e-8f, out_h, out_chi);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int gcpnet_layernorm_backward(const float* h, const float* chi, int64_t N, int32_t s, int32_t v, const float* w, const float* g_out_h,
                              const float* g_out_chi, float* g_h, float* g_chi, float* g_w, float* g_b, float* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (s > 0 && (!g_w || !g_b)) return fail("layernorm_backward: null argument");
  if (N <= 0) {
    if (s > 0) { CUDA_TRY(cudaMemsetAsync(g_w, 0, (size_t)s * sizeof(float), st)); CUDA_TRY(cudaMemsetAsync(g_b, 0, (size_t)s * sizeof(float), st)); }
    return 0;
  }
  if (!workspace) return fail("layernorm_backward: workspace required");
  float* stats = workspace;
  float* partial = workspace + 2 * N;
  gcp_layernorm_bwd_kernel<<<(int)((N * 32 + 255) / 256), 256, 0, st>>>(h, chi, (int)N, s, v, w, 1e-5f, 1e-8f, g_out_h, g_out_chi, g_h, g_chi, stats);
  gcp_note_launches(1);
  if (s > 0) {
    gcp_layernorm_wgrad_kernel<<<LN_PARTS, 128, 0, st>>>(h, g_out_h, stats, (int)N, s, partial);
    gcp_layernorm_wreduce_kernel<<<(2 * s + 127) / 128, 128, 0, st>>>(partial, LN_PARTS, s, g_w, g_b);
    gcp_note_launches(2);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcpnet_message_passing_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                                    const gcpnet_backward_io* io, const float* g_aggregate, void* stream) {
  if (!layer || !graph || !plan || !io || !g_aggregate) return fail("message_passing_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gcpnet_layer& l = *layer;
  const gcpnet_graph& g = *graph;
  if (!io->saved_edge && g.num_edges > 0) return fail("message_passing_backward: forward ran without saved activations");
  if (!io->packed) return fail("message_passing_backward: packed weights of the forward call required");
  LayerPlan lp;
  const std::string e = make_layer_plan(l, g.num_nodes, g.num_edges, &lp, nullptr, plan->tc_edge_path != 0);
  if (!e.empty()) return fail("message_passing_backward: " + e);
  if (g.num_nodes <= 0) return 0;
  // the edge backward ACCUMULATES the gathered cotangents into g_h / g_chi (the layer's node backward writes the direct
  // part first): standalone, the direct part is zero
  CUDA_TRY(cudaMemsetAsync(io->g_h, 0, (size_t)g.num_nodes * l.s * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(io->g_chi, 0, (size_t)g.num_nodes * 3 * l.v * sizeof(float), st));
  if (g.num_edges == 0) {
    CUDA_TRY(cudaMemsetAsync(io->g_params, 0, (size_t)l.n_edge_params * sizeof(float), st));
    return 0;
  }
  if (lp.tc.ok) return run_tc_edge_backward(l, g, lp, *io, g_aggregate, st);
  int edge_grid = 0;
  if (run_ffma_edge_backward(l, g, lp, *io, g_aggregate, st, &edge_grid)) return 1;
  GcpTimedScope timed(T_PARTIAL_REDUCE, st);
  const bool spilled = io->ws_edge_spill != nullptr;
  partial_reduce_kernel<<<(l.n_edge_params + 255) / 256, 256, 0, st>>>(io->g_params, io->ws_edge_partial, l.n_edge_params, edge_grid,
                                                                        nullptr, 0, 0, SkipRanges{}, spilled ? edge_wgrad_ranges(lp) : SkipRanges{});
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  if (spilled && launch_edge_wgrad(l, g, lp, io->ws_edge_spill, io->saved_edge, io->g_params, st)) return 1;
  return 0;
}

}  // extern "C"
