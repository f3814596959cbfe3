// emul.cu -- CPU emulation of the CUDA tile routines (TEST INFRASTRUCTURE, never shipped or
// loaded by gcpnet_b200).  The kernels' bodies are __host__ __device__ "phase" code
// (gcpnet_b200/csrc/gcp_tile.cuh); here each CTA is run on the host, its NT thread bodies
// executed one after another per phase (forward or reverse order).  It lets the non-GPU test
// suite check indexing, tiling and the hand-derived backward against the oracle; what it cannot
// check (barrier placement, bank conflicts, launch configuration) is covered by the -m gpu tests.
// All pointers are HOST pointers.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../gcpnet_b200/csrc/layer_setup.h"

using namespace gcp;

static std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }

static float* alloc_smem(int floats) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, (size_t)floats * 4 + 64)) return nullptr;
  // poison with NaN so that reads of never-written shared memory show up in the results
  float* f = (float*)p;
  for (int i = 0; i < floats + 16; ++i) f[i] = __builtin_nanf("");
  return f;
}

template <int TE>
static void run_edge_fwd(const EdgeParams& p, int grid) {
  const int ntiles = (p.E + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    float* sm = alloc_smem(p.sm.total);
    for (int tile = cta; tile < ntiles; tile += grid) edge_fwd_tile<TE, EDGE_NT>(p, sm, tile);
    free(sm);
  }
}
template <int TE>
static void run_edge_bwd(const EdgeParams& p, int grid) {
  const int ntiles = (p.E + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    float* sm = alloc_smem(p.sm.total);
    float* prow = p.partial + (size_t)cta * p.partial_stride;
    for (int tile = cta; tile < ntiles; tile += grid) edge_bwd_tile<TE, EDGE_NT>(p, sm, tile, prow, tile != cta);
    free(sm);
  }
}
template <int TE>
static void run_node_fwd(const NodeParams& p, int grid) {
  const int ntiles = (p.N + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    float* sm = alloc_smem(p.sm.total);
    for (int tile = cta; tile < ntiles; tile += grid) node_fwd_tile<TE, NODE_NT>(p, sm, tile);
    free(sm);
  }
}
template <int TE>
static void run_node_bwd(const NodeParams& p, int grid) {
  const int ntiles = (p.N + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    float* sm = alloc_smem(p.sm.total);
    float* prow = p.partial + (size_t)cta * p.partial_stride;
    for (int tile = cta; tile < ntiles; tile += grid) node_bwd_tile<TE, NODE_NT>(p, sm, tile, prow, tile != cta);
    free(sm);
  }
}

extern "C" {

const char* emul_last_error(void) { return g_err.c_str(); }
void emul_set_reverse(int r) { gcp::g_emul_reverse = r; }

int emul_graph_build(const int64_t* edge_index, int64_t E, int64_t N, const float* frames, int32_t* perm, int32_t* src,
                     int32_t* dst, int32_t* dst_ptr, int32_t* src_pos, int32_t* src_ptr, float* fbar) {
  std::vector<int> idx(E);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return edge_index[E + a] < edge_index[E + b]; });
  for (int64_t p = 0; p < E; ++p) { perm[p] = idx[p]; src[p] = (int)edge_index[idx[p]]; dst[p] = (int)edge_index[E + idx[p]]; }
  for (int64_t i = 0; i <= N; ++i) dst_ptr[i] = (int)(std::lower_bound(dst, dst + E, (int)i) - dst);
  std::vector<int> pos(E);
  std::iota(pos.begin(), pos.end(), 0);
  std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return src[a] < src[b]; });
  std::vector<int> keys(E);
  for (int64_t q = 0; q < E; ++q) { src_pos[q] = pos[q]; keys[q] = src[pos[q]]; }
  for (int64_t i = 0; i <= N; ++i) src_ptr[i] = (int)(std::lower_bound(keys.begin(), keys.end(), (int)i) - keys.begin());
  for (int64_t i = 0; i < N; ++i)
    for (int c = 0; c < 9; ++c) {
      float acc = 0.f;
      for (int q = src_ptr[i]; q < src_ptr[i + 1]; ++q) acc += frames[(size_t)perm[src_pos[q]] * 9 + c];
      fbar[i * 9 + c] = src_ptr[i + 1] > src_ptr[i] ? acc / (float)(src_ptr[i + 1] - src_ptr[i]) : 0.f;
    }
  return 0;
}

int emul_layer_plan(const gcpnet_layer* layer, int64_t N, int64_t E, gcpnet_plan* plan) {
  const std::string e = make_plan(*layer, N, E, plan);
  return e.empty() ? 0 : fail(e);
}

// force_edge_tile / force_node_tile: 0 = planner's choice, else the tile size to emulate
int emul_layer_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                       const gcpnet_forward_io* io, int force_edge_tile, int force_node_tile, int mp_only, float* aggregate) {
  const gcpnet_layer& l = *layer; const gcpnet_graph& g = *graph;
  const std::string e = check_layer(l);
  if (!e.empty()) return fail(e);
  const LayerOps ops = layer_ops(l);
  if (g.num_edges > 0) {
    EdgeSmem sm; int TE = pick_edge_tile(l, ops, g.num_edges, false, &sm);
    if (force_edge_tile) { TE = force_edge_tile; sm = edge_plan_smem(TE, l.s, l.v, l.se, l.ve, ops.msg, l.num_message_layers, false, edge_wc_cap(l)); }
    EdgeParams p = make_edge_params(l, g, ops, sm);
    p.h = io->h; p.chi = io->chi; p.e = io->e; p.xi = io->xi; p.frames = io->frames; p.msg = io->msg; p.saved = io->saved_edge;
    int grid = (int)((g.num_edges + TE - 1) / TE); if (grid > 3) grid = 3;  // exercise the persistent loop
    if (TE == 64) run_edge_fwd<64>(p, grid); else if (TE == 32) run_edge_fwd<32>(p, grid); else return fail("bad edge tile");
  }
  const int W = l.s + 3 * l.v;
  if (mp_only) {
    for (int64_t i = 0; i < g.num_nodes; ++i)
      for (int f = 0; f < W; ++f) {
        float acc = 0.f;
        for (int q = g.dst_ptr[i]; q < g.dst_ptr[i + 1]; ++q) acc += io->msg[(size_t)q * W + f];
        if (l.reduce_mean && g.dst_ptr[i + 1] - g.dst_ptr[i] > 1) acc /= (float)(g.dst_ptr[i + 1] - g.dst_ptr[i]);
        aggregate[i * W + f] = acc;
      }
    return 0;
  }
  NodeSmem sm; int TN = pick_node_tile(l, ops, g.num_nodes, false, &sm);
  if (force_node_tile) { TN = force_node_tile; sm = node_plan_smem(TN, l.s, l.v, l.ff0.so, l.ff0.vo, ops.ff0, ops.ff1, l.has_pos ? &ops.pu : nullptr, false, node_wc_cap(l)); }
  NodeParams p = make_node_params(l, g, ops, sm);
  p.h = io->h; p.chi = io->chi; p.msg = io->msg; p.pos = io->pos;
  p.out_h = io->out_h; p.out_chi = io->out_chi; p.out_pos = io->out_pos; p.saved = io->saved_node;
  int grid = (int)((g.num_nodes + TN - 1) / TN); if (grid > 2) grid = 2;
  if (TN == 32) run_node_fwd<32>(p, grid); else if (TN == 16) run_node_fwd<16>(p, grid); else return fail("bad node tile");
  (void)plan;
  return 0;
}

int emul_layer_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                        const gcpnet_backward_io* io, int force_node_tile, int edge_grid, int node_grid) {
  const gcpnet_layer& l = *layer; const gcpnet_graph& g = *graph;
  const std::string e = check_layer(l);
  if (!e.empty()) return fail(e);
  const LayerOps ops = layer_ops(l);
  const int W = l.s + 3 * l.v;
  NodeSmem nsm; int TN = pick_node_tile(l, ops, g.num_nodes, true, &nsm);
  if (force_node_tile) { TN = force_node_tile; nsm = node_plan_smem(TN, l.s, l.v, l.ff0.so, l.ff0.vo, ops.ff0, ops.ff1, l.has_pos ? &ops.pu : nullptr, true, node_wc_cap(l)); }
  NodeParams np = make_node_params(l, g, ops, nsm);
  np.saved = const_cast<float*>(io->saved_node);
  np.g_out_h = io->g_out_h; np.g_out_chi = io->g_out_chi; np.g_out_pos = io->g_out_pos;
  np.g_x_h = io->g_h; np.g_x_chi = io->g_chi; np.g_agg = io->ws_agg; np.partial = io->ws_node_partial;
  const int ntn = (int)((g.num_nodes + TN - 1) / TN);
  if (node_grid > ntn) node_grid = ntn;
  if (TN == 32) run_node_bwd<32>(np, node_grid); else if (TN == 16) run_node_bwd<16>(np, node_grid); else return fail("bad node tile");
  if (g.num_edges > 0) {
    EdgeSmem esm; const int TE = pick_edge_tile(l, ops, g.num_edges, true, &esm);
    EdgeParams ep = make_edge_params(l, g, ops, esm);
    ep.h = io->h; ep.chi = io->chi; ep.e = io->e; ep.xi = io->xi; ep.frames = io->frames;
    ep.saved = const_cast<float*>(io->saved_edge); ep.gagg = io->ws_agg;
    ep.grow = io->ws_edge; ep.gcol = io->ws_edge + (size_t)g.num_edges * W; ep.ge = io->g_e; ep.gxi = io->g_xi;
    ep.partial = io->ws_edge_partial;
    const int nte = (int)((g.num_edges + TE - 1) / TE);
    if (edge_grid > nte) edge_grid = nte;
    if (TE != 32) return fail("bad edge bwd tile");
    run_edge_bwd<32>(ep, edge_grid);
    for (int64_t i = 0; i < g.num_nodes; ++i)
      for (int f = 0; f < W; ++f) {
        float* out = f < l.s ? io->g_h + i * l.s + f : io->g_chi + i * 3 * l.v + (f - l.s);
        float acc = *out;
        for (int q = g.dst_ptr[i]; q < g.dst_ptr[i + 1]; ++q) acc += ep.gcol[(size_t)q * W + f];
        for (int q = g.src_ptr[i]; q < g.src_ptr[i + 1]; ++q) acc += ep.grow[(size_t)g.src_pos[q] * W + f];
        *out = acc;
      }
  } else edge_grid = 0;
  for (int i = 0; i < l.n_edge_params + l.n_node_params; ++i) {
    float acc = 0.f;
    if (i < l.n_edge_params) for (int c = 0; c < edge_grid; ++c) acc += io->ws_edge_partial[(size_t)c * l.n_edge_params + i];
    else for (int c = 0; c < node_grid; ++c) acc += io->ws_node_partial[(size_t)c * l.n_node_params + (i - l.n_edge_params)];
    io->g_params[i] = acc;
  }
  (void)plan;
  return 0;
}

}  // extern "C"
