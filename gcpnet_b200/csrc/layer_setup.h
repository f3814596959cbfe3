// layer_setup.h -- host-side translation of the C-ABI descriptors (include/gcpnet_b200.h) into the
// kernel parameter blocks, plus tile / shared-memory planning.  Shared by the CUDA launchers
// (api.cu) and by the CPU emulation used in the non-GPU tests (tests/emul/emul.cu).
#pragma once
#include <string>

#include "../../include/gcpnet_b200.h"
#include "edge_kernels.cuh"
#include "node_kernels.cuh"

namespace gcp {

constexpr int EDGE_NT = 128;
constexpr int NODE_NT = 128;
constexpr int SMEM_LIMIT_BYTES = 227 * 1024;
constexpr int MAX_PERSISTENT_CTAS = 148 * 2;

inline GcpOp to_op(const gcpnet_gcp2& d, int grad_base) {
  GcpOp o{};
  o.si = d.si; o.vi = d.vi; o.so = d.so; o.vo = d.vo; o.hd = d.hd;
  o.act_s = d.act_s; o.act_v = d.act_v; o.vres = d.vector_residual;
  o.Wd = d.vector_down; o.Wdf = d.vector_down_frames; o.Ws = d.scalar_out_w; o.bs = d.scalar_out_b;
  o.Wu = d.vector_up; o.Wg = d.vector_out_scale_w; o.bg = d.vector_out_scale_b;
  o.o_Wd = d.grad_off[0] - grad_base; o.o_Wdf = d.grad_off[1] - grad_base; o.o_Ws = d.grad_off[2] - grad_base;
  o.o_bs = d.grad_off[3] - grad_base; o.o_Wu = d.grad_off[4] - grad_base; o.o_Wg = d.grad_off[5] - grad_base;
  o.o_bg = d.grad_off[6] - grad_base;
  return o;
}

inline std::string check_gcp2(const gcpnet_gcp2& d, const char* name) {
  auto err = [&](const std::string& m) { return std::string(name) + ": " + m; };
  if (d.si <= 0 || d.vi <= 0 || d.so <= 0 || d.vo < 0) return err("dims must be positive");
  if (d.hd <= 0 || d.hd > 16) return err("hidden vector dim must be in [1,16] (bottleneck too small for this build)");
  if (d.vector_residual && d.vi != d.vo) return err("vector_residual needs vi == vo");
  if (d.act_s < 0 || d.act_s > 5 || d.act_v < 0 || d.act_v > 5) return err("unknown nonlinearity");
  return "";
}

inline std::string check_layer(const gcpnet_layer& l) {
  if (l.num_message_layers < 1 || l.num_message_layers > GCPNET_MAX_MESSAGE_LAYERS) return "num_message_layers out of range";
  if (l.s <= 0 || l.v <= 0 || l.se <= 0 || l.ve <= 0) return "layer dims must be positive";
  for (int k = 0; k < l.num_message_layers; ++k) {
    const gcpnet_gcp2& g = l.message[k];
    std::string e = check_gcp2(g, "message_fusion");
    if (!e.empty()) return e;
    if (g.so != l.s || g.vo != l.v) return "message GCP output dims must equal node dims";
    if (k == 0 && (g.si != 2 * l.s + l.se || g.vi != 2 * l.v + l.ve)) return "message_fusion.0 input dims mismatch";
    if (k > 0 && (g.si != l.s || g.vi != l.v)) return "message_fusion.k input dims mismatch";
  }
  std::string e = check_gcp2(l.ff0, "feedforward_network.0");
  if (!e.empty()) return e;
  e = check_gcp2(l.ff1, "feedforward_network.1");
  if (!e.empty()) return e;
  if (l.ff0.si != l.s || l.ff0.vi != l.v || l.ff1.so != l.s || l.ff1.vo != l.v || l.ff0.so != l.ff1.si || l.ff0.vo != l.ff1.vi)
    return "feed-forward dims mismatch";
  if (l.has_pos) {
    e = check_gcp2(l.pos_update, "node_position_update_network.0");
    if (!e.empty()) return e;
    if (l.pos_update.si != l.s || l.pos_update.vi != l.v || l.pos_update.so != l.s || l.pos_update.vo != 1)
      return "position-update GCP dims mismatch";
  }
  if (l.training && !(l.p_drop >= 0.f && l.p_drop < 1.f)) return "dropout probability must be in [0,1)";
  return "";
}

struct LayerOps {
  GcpOp msg[MAX_MSG_LAYERS];
  GcpOp ff0, ff1, pu;
};
inline LayerOps layer_ops(const gcpnet_layer& l) {
  LayerOps o{};
  for (int k = 0; k < l.num_message_layers; ++k) o.msg[k] = to_op(l.message[k], 0);
  o.ff0 = to_op(l.ff0, l.n_edge_params);
  o.ff1 = to_op(l.ff1, l.n_edge_params);
  if (l.has_pos) o.pu = to_op(l.pos_update, l.n_edge_params);
  return o;
}

inline int edge_wc_cap(const gcpnet_layer& l) {
  // k-major data-gradient GEMM stages [K = s][64] in one chunk; n-major needs >= 64 x (8 + 4)
  int cap = 12288;
  const int need = E_OGD * E_NRD * round_up(l.s, 4);
  if (need > cap) cap = need;
  return cap;
}
inline int node_wc_cap(const gcpnet_layer& l) {
  int cap = 12288;
  int so = l.ff0.so > l.s ? l.ff0.so : l.s;
  const int need = N_OGD * N_NRD * round_up(so, 4);
  if (need > cap) cap = need;
  return cap;
}

inline void edge_saved_offsets(const gcpnet_layer& l, long long E, long long* offT, long long* offG, long long* offS,
                               long long* offV, long long* total) {
  long long off = 0;
  for (int k = 0; k < l.num_message_layers; ++k) {
    offT[k] = off; off += E * l.s;
    offG[k] = off; off += E * l.v;
    if (k < l.num_message_layers - 1) {
      offS[k] = off; off += E * l.s;
      offV[k] = off; off += E * 3 * l.v;
    } else { offS[k] = 0; offV[k] = 0; }
  }
  *total = off;
}

// Pick the largest edge tile whose shared-memory plan fits; small problems get the small tile so
// that the tile count covers the 148 SMs.
inline int pick_edge_tile(const gcpnet_layer& l, const LayerOps& ops, long long E, bool backward, EdgeSmem* out) {
  const int cands[2] = {64, 32};
  for (int ci = 0; ci < 2; ++ci) {
    const int TE = cands[ci];
    if (TE == 64 && (backward || E < 64LL * 148 * 2)) continue;
    EdgeSmem m = edge_plan_smem(TE, l.s, l.v, l.se, l.ve, ops.msg, l.num_message_layers, backward, edge_wc_cap(l));
    if ((long long)m.total * 4 <= SMEM_LIMIT_BYTES) { *out = m; return TE; }
  }
  return 0;
}
inline int pick_node_tile(const gcpnet_layer& l, const LayerOps& ops, long long N, bool backward, NodeSmem* out) {
  const int cands[2] = {32, 16};
  for (int ci = 0; ci < 2; ++ci) {
    const int TE = cands[ci];
    if (TE == 32 && N < 32LL * 148 * 2) continue;
    NodeSmem m = node_plan_smem(TE, l.s, l.v, l.ff0.so, l.ff0.vo, ops.ff0, ops.ff1, l.has_pos ? &ops.pu : nullptr, backward, node_wc_cap(l));
    if ((long long)m.total * 4 <= SMEM_LIMIT_BYTES) { *out = m; return TE; }
  }
  return 0;
}

inline std::string make_plan(const gcpnet_layer& l, long long N, long long E, gcpnet_plan* plan) {
  std::string e = check_layer(l);
  if (!e.empty()) return e;
  const LayerOps ops = layer_ops(l);
  EdgeSmem ef, eb; NodeSmem nf, nb;
  const int tef = pick_edge_tile(l, ops, E, false, &ef);
  const int teb = pick_edge_tile(l, ops, E, true, &eb);
  const int tnf = pick_node_tile(l, ops, N, false, &nf);
  const int tnb = pick_node_tile(l, ops, N, true, &nb);
  if (!tef || !teb || !tnf || !tnb) return "feature dims too large for the shared-memory tile plan of this build";
  gcpnet_plan p{};
  // forward and backward may use different tile sizes; saved activations are stored per sorted edge row
  p.edge_tile = tef; p.node_tile = tnf;
  auto tiles = [](long long n, int t) { return (int)((n + t - 1) / t); };
  auto grid = [&](long long n, int t) { int g = tiles(n, t); return g > MAX_PERSISTENT_CTAS ? MAX_PERSISTENT_CTAS : (g < 1 ? 1 : g); };
  p.edge_grid_fwd = grid(E, tef); p.edge_grid_bwd = grid(E, teb);
  p.node_grid_fwd = grid(N, tnf); p.node_grid_bwd = grid(N, tnb);
  p.edge_smem_fwd_bytes = ef.total * 4; p.edge_smem_bwd_bytes = eb.total * 4;
  p.node_smem_fwd_bytes = nf.total * 4; p.node_smem_bwd_bytes = nb.total * 4;
  const long long W = l.s + 3 * l.v;
  p.msg_floats = E * W;
  long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS], tot;
  edge_saved_offsets(l, E, offT, offG, offS, offV, &tot);
  p.saved_edge_floats = tot;
  p.saved_node_floats = node_saved_layout((int)N, l.s, l.v, l.ff0.so, l.ff0.vo, l.has_pos != 0, l.training != 0).total;
  p.edge_partial_floats = (long long)p.edge_grid_bwd * l.n_edge_params;
  p.node_partial_floats = (long long)p.node_grid_bwd * l.n_node_params;
  p.edge_cotangent_floats = 2 * E * W;
  p.agg_cotangent_floats = N * W;
  *plan = p;
  return "";
}

inline int edge_tile_bwd(const gcpnet_layer& l, long long E, EdgeSmem* m) { return pick_edge_tile(l, layer_ops(l), E, true, m); }

inline EdgeParams make_edge_params(const gcpnet_layer& l, const gcpnet_graph& g, const LayerOps& ops, const EdgeSmem& sm) {
  EdgeParams p{};
  p.N = (int)g.num_nodes; p.E = (int)g.num_edges; p.L = l.num_message_layers;
  p.s = l.s; p.v = l.v; p.se = l.se; p.ve = l.ve;
  p.residual = l.residual_messages; p.e3 = l.enable_e3; p.reduce_mean = l.reduce_mean; p.slope = l.slope;
  p.perm = g.perm; p.src = g.src; p.dst = g.dst; p.dst_ptr = g.dst_ptr;
  for (int k = 0; k < p.L; ++k) p.ops[k] = ops.msg[k];
  long long tot;
  edge_saved_offsets(l, g.num_edges, p.offT, p.offG, p.offS, p.offV, &tot);
  p.partial_stride = l.n_edge_params;
  p.sm = sm;
  return p;
}

inline NodeParams make_node_params(const gcpnet_layer& l, const gcpnet_graph& g, const LayerOps& ops, const NodeSmem& sm) {
  NodeParams p{};
  p.N = (int)g.num_nodes; p.s = l.s; p.v = l.v; p.hs = l.ff0.so; p.hv = l.ff0.vo;
  p.has_pos = l.has_pos; p.reduce_mean = l.reduce_mean; p.train = l.training;
  p.slope = l.slope; p.ln_eps = l.ln_eps; p.vn_eps = l.vn_eps; p.pos_weight = l.pos_weight; p.p_drop = l.training ? l.p_drop : 0.f;
  p.seed = l.seed; p.rng_ctr = (const long long*)l.rng_counter;
  p.fbar = g.fbar; p.dst_ptr = g.dst_ptr;
  p.ln0_w = l.ln0_w; p.ln0_b = l.ln0_b; p.ln1_w = l.ln1_w; p.ln1_b = l.ln1_b;
  p.ff0 = ops.ff0; p.ff1 = ops.ff1; p.pu = ops.pu;
  p.sv = node_saved_layout(p.N, l.s, l.v, p.hs, p.hv, l.has_pos != 0, l.training != 0);
  p.partial_stride = l.n_node_params;
  p.o_ln0w = l.ln_grad_off[0] - l.n_edge_params; p.o_ln0b = l.ln_grad_off[1] - l.n_edge_params;
  p.o_ln1w = l.ln_grad_off[2] - l.n_edge_params; p.o_ln1b = l.ln_grad_off[3] - l.n_edge_params;
  p.sm = sm;
  return p;
}

}  // namespace gcp
