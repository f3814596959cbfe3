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

template <int TE, int SLF>
static void run_edge_fwd(const EdgeParams& p, int grid) {
  constexpr int NT = 8 * TE;
  const int ntiles = (p.E + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    const int mine = cta < ntiles ? (ntiles - cta + grid - 1) / grid : 0;
    if (!mine) continue;
    float* sm = alloc_smem(p.sm.total);
    WPipe wp = edge_pipe(p, sm, mine);
    wpipe_start<NT>(wp);
    for (int tile = cta; tile < ntiles; tile += grid) edge_fwd_tile<TE, NT, SLF>(p, sm, tile, wp, tile == cta);
    free(sm);
  }
}
template <int TE, int SLF>
static void run_edge_bwd(const EdgeParams& p, int grid) {
  constexpr int NT = 8 * TE;
  const int ntiles = (p.E + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    const int mine = cta < ntiles ? (ntiles - cta + grid - 1) / grid : 0;
    if (!mine) continue;
    float* sm = alloc_smem(p.sm.total);
    WPipe wp = edge_pipe(p, sm, mine);
    wpipe_start<NT>(wp);
    float* prow = p.partial + (size_t)cta * p.partial_stride;
    for (int tile = cta; tile < ntiles; tile += grid) edge_bwd_tile<TE, NT, SLF, EDGE_SLD>(p, sm, tile, wp, prow, tile != cta);
    free(sm);
  }
}
template <int SLF>
static void run_node_fwd(const NodeParams& p, int grid) {
  constexpr int TE = NODE_TE, NT = NODE_NT;
  const int ntiles = (p.N + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    const int mine = cta < ntiles ? (ntiles - cta + grid - 1) / grid : 0;
    if (!mine) continue;
    float* sm = alloc_smem(p.sm.total);
    WPipe wp = node_pipe(p, sm, mine);
    wpipe_start<NT>(wp);
    for (int tile = cta; tile < ntiles; tile += grid) node_fwd_tile<TE, NT, SLF>(p, sm, tile, wp, tile == cta);
    free(sm);
  }
}
template <int SLF>
static void run_node_bwd(const NodeParams& p, int grid) {
  constexpr int TE = NODE_TE, NT = NODE_NT;
  const int ntiles = (p.N + TE - 1) / TE;
  for (int cta = 0; cta < grid; ++cta) {
    const int mine = cta < ntiles ? (ntiles - cta + grid - 1) / grid : 0;
    if (!mine) continue;
    float* sm = alloc_smem(p.sm.total);
    WPipe wp = node_pipe(p, sm, mine);
    wpipe_start<NT>(wp);
    float* prow = p.partial + (size_t)cta * p.partial_stride;
    for (int tile = cta; tile < ntiles; tile += grid) node_bwd_tile<TE, NT, SLF, NODE_SLD>(p, sm, tile, wp, prow, tile != cta);
    free(sm);
  }
}

template <class F1, class F2>
static int by_slf(int slf, F1 one, F2 two) { if (slf == 1) { one(); return 0; } if (slf == 2) { two(); return 0; } return 1; }

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
  LayerPlan lp;
  const std::string e = make_layer_plan(*layer, N, E, &lp, plan, false);  // the emulation covers the FFMA tile path
  return e.empty() ? 0 : fail(e);
}

// force_edge_tile: 0 = planner's choice, else the edge tile size to emulate (32 / 48 / 64)
int emul_layer_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                       const gcpnet_forward_io* io, int force_edge_tile, int force_node_tile, int mp_only, float* aggregate) {
  const gcpnet_layer& l = *layer; const gcpnet_graph& g = *graph;
  LayerPlan lp;
  const std::string e = make_layer_plan(l, g.num_nodes, g.num_edges, &lp, nullptr, false);
  if (!e.empty()) return fail(e);
  if (force_edge_tile && !pick_edge_tile(l, lp.ops, g.num_edges, false, force_edge_tile, &lp.ef)) return fail("forced edge tile does not fit");
  (void)force_node_tile;
  {  // pack kernel
    const PackParams pp = make_pack_params(lp.ops, io->packed);
    for (int b = 0; b < pp.n; ++b) pack_gcp<256>(pp.ops[b], pp.blob);
  }
  if (g.num_edges > 0) {
    EdgeParams p = make_edge_params(l, g, lp.ops, lp.ef, false, io->packed);
    p.h = io->h_gather ? io->h_gather : io->h; p.chi = io->chi_gather ? io->chi_gather : io->chi;
    p.e = io->e; p.xi = io->xi; p.frames = io->frames; p.agg = io->agg; p.saved = io->saved_edge;
    int grid = lp.ef.grid; if (grid > 3) grid = 3;  // exercise the persistent loop
    int bad = 1;
    if (lp.ef.TE == 32) bad = by_slf(lp.ef.SLF, [&] { run_edge_fwd<32, 1>(p, grid); }, [&] { run_edge_fwd<32, 2>(p, grid); });
    if (lp.ef.TE == 48) bad = by_slf(lp.ef.SLF, [&] { run_edge_fwd<48, 1>(p, grid); }, [&] { run_edge_fwd<48, 2>(p, grid); });
    if (lp.ef.TE == 64) bad = by_slf(lp.ef.SLF, [&] { run_edge_fwd<64, 1>(p, grid); }, [&] { run_edge_fwd<64, 2>(p, grid); });
    if (bad) return fail("bad edge tile");
  }
  const int W = l.s + 3 * l.v;
  if (mp_only) {
    for (int64_t i = 0; i < g.num_nodes; ++i)
      for (int f = 0; f < W; ++f) {
        float acc = segment_total(io->agg, g.num_nodes, W, lp.ef.TE, g.dst_ptr, (int)i, f);
        if (l.reduce_mean && g.dst_ptr[i + 1] - g.dst_ptr[i] > 1) acc /= (float)(g.dst_ptr[i + 1] - g.dst_ptr[i]);
        aggregate[i * W + f] = acc;
      }
    return 0;
  }
  NodeParams p = make_node_params(l, g, lp.ops, lp.nf, false, io->packed);
  p.h = io->h; p.chi = io->chi; p.agg = io->agg; p.pos = io->pos; p.frames = io->frames;
  p.edge_rows = lp.ef.TE;
  p.out_h = io->out_h; p.out_chi = io->out_chi; p.out_pos = io->out_pos; p.saved = io->saved_node;
  int grid = lp.nf.grid; if (grid > 2) grid = 2;
  if (lp.nf.SLF == 1) run_node_fwd<1>(p, grid); else if (lp.nf.SLF == 2) run_node_fwd<2>(p, grid);
  else if (lp.nf.SLF == 4) run_node_fwd<4>(p, grid); else return fail("bad node tile");
  (void)plan;
  return 0;
}

// floats of the spill area of the off-tile weight-gradient product (0 on error)
long long emul_edge_spill_floats(const gcpnet_layer* layer, long long N, long long E) {
  LayerPlan lp;
  if (!make_layer_plan(*layer, N, E, &lp, nullptr, false).empty()) return 0;
  return edge_spill_layout(E, lp.ops).total;
}

int emul_layer_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                        const gcpnet_backward_io* io, int force_node_tile, int edge_grid, int node_grid) {
  const gcpnet_layer& l = *layer; const gcpnet_graph& g = *graph;
  LayerPlan lp;
  const std::string e = make_layer_plan(l, g.num_nodes, g.num_edges, &lp, nullptr, false);
  if (!e.empty()) return fail(e);
  (void)force_node_tile;
  const int W = l.s + 3 * l.v;
  NodeParams np = make_node_params(l, g, lp.ops, lp.nb, true, io->packed);
  np.saved = const_cast<float*>(io->saved_node);
  np.h = io->h; np.chi = io->chi; np.frames = io->frames;
  np.g_out_h = io->g_out_h; np.g_out_chi = io->g_out_chi; np.g_out_pos = io->g_out_pos;
  np.g_x_h = io->g_h; np.g_x_chi = io->g_chi; np.g_agg = io->ws_agg; np.partial = io->ws_node_partial;
  const int ntn = (int)((g.num_nodes + lp.nb.TE - 1) / lp.nb.TE);
  if (node_grid > ntn) node_grid = ntn;
  if (lp.nb.SLF == 1) run_node_bwd<1>(np, node_grid); else if (lp.nb.SLF == 2) run_node_bwd<2>(np, node_grid);
  else if (lp.nb.SLF == 4) run_node_bwd<4>(np, node_grid); else return fail("bad node tile");
  if (g.num_edges > 0) {
    EdgeParams ep = make_edge_params(l, g, lp.ops, lp.eb, true, io->packed);
    ep.h = io->h_gather ? io->h_gather : io->h; ep.chi = io->chi_gather ? io->chi_gather : io->chi;
    ep.e = io->e; ep.xi = io->xi; ep.frames = io->frames;
    ep.saved = const_cast<float*>(io->saved_edge); ep.gagg = io->ws_agg;
    ep.grow = io->ws_edge; ep.gcol = io->ws_edge + (size_t)g.num_edges * W; ep.ge = io->g_e; ep.gxi = io->g_xi;
    ep.partial = io->ws_edge_partial;
    if (io->ws_edge_spill != nullptr) {  // off-tile weight gradients: the tiles spill gT / Z / gg rows (layer_setup.h: EdgeSpill)
      const EdgeSpill sp = edge_spill_layout(g.num_edges, lp.ops);
      ep.spill = io->ws_edge_spill;
      for (int k = 0; k < lp.ops.L; ++k) { ep.sp_gT[k] = sp.gT[k]; ep.sp_Z[k] = sp.Z[k]; ep.sp_GG[k] = sp.GG[k]; }
    }
    const int nte = (int)((g.num_edges + lp.eb.TE - 1) / lp.eb.TE);
    if (edge_grid > nte) edge_grid = nte;
    int bad = 1;
    if (lp.eb.TE == 32) bad = by_slf(lp.eb.SLF, [&] { run_edge_bwd<32, 1>(ep, edge_grid); }, [&] { run_edge_bwd<32, 2>(ep, edge_grid); });
    if (lp.eb.TE == 48) bad = by_slf(lp.eb.SLF, [&] { run_edge_bwd<48, 1>(ep, edge_grid); }, [&] { run_edge_bwd<48, 2>(ep, edge_grid); });
    if (lp.eb.TE == 64) bad = by_slf(lp.eb.SLF, [&] { run_edge_bwd<64, 1>(ep, edge_grid); }, [&] { run_edge_bwd<64, 2>(ep, edge_grid); });
    if (bad) return fail("bad edge bwd tile");
    if (g.gsrc != nullptr) {  // autoregressive: sums over the rows of the [2N] gather table (ar_cotangent_reduce_kernel)
      for (int64_t u = 0; u < g.num_gather_rows; ++u)
        for (int f = 0; f < W; ++f) {
          float acc = 0.f;
          if ((u & 1) == 0) acc = f < l.s ? io->g_h[(u >> 1) * l.s + f] : io->g_chi[(u >> 1) * 3 * l.v + (f - l.s)];
          for (int q = g.vdst_ptr[u]; q < g.vdst_ptr[u + 1]; ++q) acc += ep.gcol[(size_t)q * W + f];
          for (int q = g.vsrc_ptr[u]; q < g.vsrc_ptr[u + 1]; ++q) acc += ep.grow[(size_t)g.vsrc_pos[q] * W + f];
          if (f < l.s) io->g_h_gather[u * l.s + f] = acc; else io->g_chi_gather[u * 3 * l.v + (f - l.s)] = acc;
        }
    } else
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
  if (io->ws_edge_spill != nullptr && g.num_edges > 0) {
    // host restatement of launch_edge_wgrad (api.cu): dW[j][i] = sum_e G[e][j] * act(Z[e][i]), db[j] = sum_e G[e][j] over the
    // spilled rows, written over the (unwritten) partial sums of those parameters
    const long long E = g.num_edges;
    const EdgeSpill sp = edge_spill_layout(E, lp.ops);
    long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS], tot;
    edge_saved_offsets(l, E, offT, offG, offS, offV, &tot);
    auto product = [&](const float* G, int ldg, int J, const float* Z, int ldz, int I, int act, float* outW, float* outb) {
      for (int j = 0; j < J; ++j) {
        double b = 0.0;
        for (long long n = 0; n < E; ++n) b += G[n * ldg + j];
        outb[j] = (float)b;
        for (int i = 0; i < I; ++i) {
          double a = 0.0;
          for (long long n = 0; n < E; ++n) a += (double)G[n * ldg + j] * (double)act_fwd(act, Z[n * ldz + i], l.slope);
          outW[(size_t)j * I + i] = (float)a;
        }
      }
    };
    for (int k = 0; k < lp.ops.L; ++k) {
      const GcpOp& o = lp.ops.msg[k];
      const float* spill = io->ws_edge_spill;
      product(spill + sp.gT[k], sp.ldg[k], o.so, spill + sp.Z[k], sp.ldz[k], gcp_k(o), ACT_NONE, io->g_params + o.o_Ws, io->g_params + o.o_bs);
      if (gcp_gated(o))
        product(spill + sp.GG[k], sp.ldgg[k], o.vo, io->saved_edge + offT[k], o.so, o.so, o.act_v, io->g_params + o.o_Wg, io->g_params + o.o_bg);
    }
  }
  if (l.pre_norm)  // gcp_norm.0 is applied (and differentiated) in front of the layer: the tiles never write these rows
    for (int j = 0; j < 2 * l.s; ++j) io->g_params[l.ln_grad_off[0] + j] = 0.f;
  (void)plan;
  return 0;
}

}  // extern "C"
