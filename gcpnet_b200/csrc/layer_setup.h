// layer_setup.h -- host-side translation of the C-ABI descriptors (include/gcpnet_b200.h) into the
// kernel parameter blocks: packed-weight layout, chunk sequences of the shared-memory weight ring,
// tile / shared-memory planning.  Shared by the CUDA launchers (api.cu) and by the CPU emulation used
// in the non-GPU tests (tests/emul/emul.cu).
#pragma once
#include <string>

#include "../../include/gcpnet_b200.h"
#include "edge_kernels.cuh"
#include "node_kernels.cuh"
#include "gcp2_op.cuh"
#include "pack.cuh"
#include "tc_setup.h"

namespace gcp {

constexpr int SMEM_LIMIT_BYTES = 227 * 1024;
constexpr int NUM_SMS = 148;                 // B200
constexpr int EDGE_CAP_FLOATS = 8192;        // largest weight chunk of the edge kernels (32 KB ring slot)
constexpr int NODE_CAP_FLOATS = 16384;       // node kernels: wide feed-forward layers (64 KB ring slot)
constexpr int NODE_TE = 16, NODE_NT = 512;   // 32 threads per node
constexpr int EDGE_SLD = 2, NODE_SLD = 2;
constexpr int LN_PARTS = 64;                // node chunks of the standalone GCPLayerNorm backward (weight / bias partials)

inline GcpOp to_op(const gcpnet_gcp2& d, int grad_base) {
  GcpOp o{};
  o.si = d.si; o.vi = d.vi; o.so = d.so; o.vo = d.vo; o.hd = d.hd;
  o.act_s = d.act_s; o.act_v = d.act_v; o.vres = d.vector_residual;
  o.Wd = d.vector_down; o.Wdf = d.vector_down_frames; o.Ws = d.scalar_out_w; o.bs = d.scalar_out_b;
  o.Wu = d.vector_up; o.Wg = d.vector_out_scale_w; o.bg = d.vector_out_scale_b;
  o.o_Wd = d.grad_off[0] - grad_base; o.o_Wdf = d.grad_off[1] - grad_base; o.o_Ws = d.grad_off[2] - grad_base;
  o.o_bs = d.grad_off[3] - grad_base; o.o_Wu = d.grad_off[4] - grad_base; o.o_Wg = d.grad_off[5] - grad_base;
  o.o_bg = d.grad_off[6] - grad_base;
  o.flags = d.flags;
  return o;
}

inline std::string check_gcp2(const gcpnet_gcp2& d, const char* name) {
  auto err = [&](const std::string& m) { return std::string(name) + ": " + m; };
  if (d.si <= 0 || d.vi <= 0 || d.so <= 0 || d.vo < 0) return err("dims must be positive (vo may be 0: scalar-only output, gcpnet.py:443-446)");
  if (d.so % 4 != 0) return err("scalar output dims must be multiples of 4 in this build");
  if (d.hd <= 0 || d.hd > 32) return err("hidden vector dim must be in [1,32] (bottleneck too small for this build)");
  if (d.vector_residual && d.vi != d.vo) return err("vector_residual needs vi == vo");
  if (d.act_s < 0 || d.act_s > 5 || d.act_v < 0 || d.act_v > 5) return err("unknown nonlinearity");
  if (d.flags & ~(GCPNET_GCP2_NO_FRAMES | GCPNET_GCP2_NO_GATE)) return err("unknown flags");
  if ((d.flags & GCPNET_GCP2_NO_GATE) && d.vo > 0 && d.act_v != 0)
    return err("vector_gate=False with a vector nonlinearity (norm gating, gcpnet.py:349-350) is not built");
  return "";
}

inline std::string check_layer(const gcpnet_layer& l) {
  if (l.num_message_layers < 1 || l.num_message_layers > GCPNET_MAX_MESSAGE_LAYERS) return "num_message_layers out of range";
  if (l.s <= 0 || l.v <= 0 || l.se <= 0 || l.ve <= 0) return "layer dims must be positive";
  if (l.s % 4 != 0) return "node scalar dim must be a multiple of 4 in this build";
  for (int k = 0; k < l.num_message_layers; ++k) {
    const gcpnet_gcp2& g = l.message[k];
    std::string e = check_gcp2(g, "message_fusion");
    if (!e.empty()) return e;
    if (g.so != l.s || g.vo != l.v || g.vo <= 0) return "message GCP output dims must equal node dims";
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

// ---- packed-weight layout ------------------------------------------------------------------------
// Fills op.w (chunk sizes and strides) and assigns blob offsets starting at *cursor.
inline std::string plan_gcp_pack(GcpOp& op, int cap_floats, int* cursor) {
  GcpW& W = op.w;
  W.cols = hd_cols(op.hd); W.hdp = W.cols - 4;
  W.NP = round_up(op.so, 16);
  const int K = gcp_k(op);
  int kc_max = 0;
  for (int kc = 16; kc <= 128; kc += 16)
    if (W.NP * ld_vec(kc) + W.NP <= cap_floats) kc_max = kc;
  if (kc_max == 0) return "scalar_out is too wide for a weight-ring slot of this build";
  W.nWS = (K + kc_max - 1) / kc_max;
  if (W.nWS > MAX_WS_CHUNKS) return "scalar_out has too many input columns for this build";
  W.kc = round_up((K + W.nWS - 1) / W.nWS, 16);
  W.ldk = ld_vec(W.kc);
  W.ldg = ld_vec(op.so);
  W.o_bg = round_up(op.vo, 4) * W.ldg;  // zero rows up to a multiple of 4: read by the 4-deep k loop of the gT GEMM
  W.o_wu = W.o_bg + round_up(op.vo, 4);
  auto take = [&](int floats) { WChunk c; c.off = *cursor; c.floats = round_up(floats, 4); *cursor += round_up(c.floats, 32); return c; };
  W.S = take(op.vi * W.cols);
  W.ws_floats = round_up(W.NP * W.ldk + W.NP, 4);
  W.ws_stride = round_up(W.ws_floats, 32);
  W.ws_off = *cursor; *cursor += W.nWS * W.ws_stride;
  W.G = take(W.o_wu + op.vo * W.hdp > 32 ? W.o_wu + op.vo * W.hdp : 32);  // never an empty chunk (vo = 0: 32 zero floats)
  if (W.G.floats > cap_floats) return "vector_out_scale is too wide for a weight-ring slot of this build";
  return "";
}

inline void seq_push(WSeq& q, const WChunk& c) { q.c[q.n++] = c; if (c.floats > q.slot_floats) q.slot_floats = c.floats; }
inline WChunk ws_chunk(const GcpOp& op, int c) { WChunk k; k.off = op.w.ws_off + c * op.w.ws_stride; k.floats = op.w.ws_floats; return k; }
inline void seq_fwd(WSeq& q, const GcpOp& op) {
  seq_push(q, op.w.S);
  for (int c = 0; c < op.w.nWS; ++c) seq_push(q, ws_chunk(op, c));
  seq_push(q, op.w.G);
}
inline void seq_bwd(WSeq& q, const GcpOp& op) {
  seq_push(q, op.w.S);
  seq_push(q, op.w.G);
  for (int c = 0; c < op.w.nWS; ++c) seq_push(q, ws_chunk(op, c));
}

struct LayerOps {
  GcpOp msg[MAX_MSG_LAYERS];
  GcpOp ff0, ff1, pu;
  int L, has_pos;
  int packed_floats;
  WSeq edge_fwd, edge_bwd, node_fwd, node_bwd;
  std::string error;
};
inline LayerOps layer_ops(const gcpnet_layer& l, int edge_cap = EDGE_CAP_FLOATS, int node_cap = NODE_CAP_FLOATS) {
  LayerOps o{};
  o.L = l.num_message_layers; o.has_pos = l.has_pos;
  int cursor = 0;
  auto plan = [&](GcpOp& op, int cap) { if (o.error.empty()) o.error = plan_gcp_pack(op, cap, &cursor); };
  for (int k = 0; k < o.L; ++k) { o.msg[k] = to_op(l.message[k], 0); plan(o.msg[k], edge_cap); }
  o.ff0 = to_op(l.ff0, l.n_edge_params); plan(o.ff0, node_cap);
  o.ff1 = to_op(l.ff1, l.n_edge_params); plan(o.ff1, node_cap);
  if (l.has_pos) { o.pu = to_op(l.pos_update, l.n_edge_params); plan(o.pu, node_cap); }
  o.packed_floats = cursor;
  if (!o.error.empty()) return o;
  int need = 0;
  for (int k = 0; k < o.L; ++k) need += 2 + o.msg[k].w.nWS;
  int need_n = 6 + o.ff0.w.nWS + o.ff1.w.nWS + (l.has_pos ? o.pu.w.nWS : 0);
  if (need > MAX_WSEQ || need_n > MAX_WSEQ) { o.error = "layer needs too many weight chunks for this build"; return o; }
  for (int k = 0; k < o.L; ++k) seq_fwd(o.edge_fwd, o.msg[k]);
  for (int k = o.L - 1; k >= 0; --k) seq_bwd(o.edge_bwd, o.msg[k]);
  seq_fwd(o.node_fwd, o.ff0); seq_fwd(o.node_fwd, o.ff1);
  if (l.has_pos) { seq_fwd(o.node_fwd, o.pu); seq_bwd(o.node_bwd, o.pu); }
  seq_bwd(o.node_bwd, o.ff1); seq_bwd(o.node_bwd, o.ff0);
  for (WSeq* q : {&o.edge_fwd, &o.edge_bwd, &o.node_fwd, &o.node_bwd}) q->slot_floats = round_up(q->slot_floats, 32);
  return o;
}

inline PackParams make_pack_params(const LayerOps& o, float* blob, bool skip_messages = false, bool skip_node = false) {
  PackParams p{};
  p.blob = blob;
  for (int k = 0; k < o.L && !skip_messages; ++k) p.ops[p.n++] = o.msg[k];  // the tensor-core path packs its own message tiles
  if (skip_node) return p;  // GCPMessagePassing alone: no feed-forward / position-update weights
  p.ops[p.n++] = o.ff0; p.ops[p.n++] = o.ff1;
  if (o.has_pos) p.ops[p.n++] = o.pu;
  return p;
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

// ---- tile planning ---------------------------------------------------------------------------------
struct EdgeTilePlan { int TE, NT, SLF, nslot, grid; EdgeSmem sm; };
struct NodeTilePlan { int TE, NT, SLF, nslot, grid; NodeSmem sm; };

inline int edge_slf(const gcpnet_layer& l) { return (round_up(l.s, 16) / 16 + 3) / 4; }  // WN = 4 warps across the columns

// Edge tiles: 8 threads per edge, TE in {32, 48, 64}.  Prefer the smallest tile whose tile count fits one
// wave of one-CTA-per-SM (small graphs are latency bound: more CTAs in flight beat fatter tiles); large
// graphs take the fattest tile that fits shared memory and loop persistently.
inline bool pick_edge_tile(const gcpnet_layer& l, LayerOps& ops, long long E, bool backward, int force_te, EdgeTilePlan* out) {
  WSeq& q = backward ? ops.edge_bwd : ops.edge_fwd;
  const int slf = edge_slf(l);
  if (slf > 2) return false;
  const int cands[3] = {32, 48, 64};
  bool have = false;
  for (int ci = 0; ci < 3; ++ci) {
    const int TE = cands[ci];
    if (force_te && TE != force_te) continue;
    EdgeTilePlan p{};
    bool fits = false;
    for (int nslot = 3; nslot >= 1 && !fits; --nslot) {
      const EdgeSmem m = edge_plan_smem(TE, l.s, l.v, l.se, l.ve, ops.msg, l.num_message_layers, backward, nslot, q.slot_floats);
      if ((long long)m.total * 4 <= SMEM_LIMIT_BYTES) { p.sm = m; p.nslot = nslot; fits = true; }
    }
    if (!fits) break;  // larger tiles need even more shared memory
    p.TE = TE; p.NT = 8 * TE; p.SLF = slf;
    const long long tiles = (E + TE - 1) / TE;
    p.grid = (int)(tiles < 1 ? 1 : (tiles > NUM_SMS ? NUM_SMS : tiles));
    *out = p; have = true;
    if (tiles <= NUM_SMS) break;  // one wave: stop at the smallest such tile
  }
  if (have) q.nslot = out->nslot;
  return have;
}

inline int node_slf(const gcpnet_layer& l) {
  int np = round_up(l.ff0.so, 16);
  if (round_up(l.s, 16) > np) np = round_up(l.s, 16);
  const int wn = (NODE_NT / 32) / (NODE_TE / 16);
  const int need = (np / 16 + wn - 1) / wn;
  return need <= 1 ? 1 : (need <= 2 ? 2 : (need <= 4 ? 4 : 0));
}
inline bool pick_node_tile(const gcpnet_layer& l, LayerOps& ops, long long N, bool backward, NodeTilePlan* out) {
  WSeq& q = backward ? ops.node_bwd : ops.node_fwd;
  const int slf = node_slf(l);
  if (slf == 0) return false;
  for (int nslot = 3; nslot >= 1; --nslot) {  // 1 slot = no prefetch overlap, last resort for very wide layers
    NodeSmem m = node_plan_smem(NODE_TE, NODE_NT, l.s, l.v, l.ff0.so, l.ff0.vo, ops.ff0, ops.ff1, l.has_pos ? &ops.pu : nullptr,
                                backward, nslot, q.slot_floats);
    if ((long long)m.total * 4 > SMEM_LIMIT_BYTES) continue;
    NodeTilePlan p{};
    p.TE = NODE_TE; p.NT = NODE_NT; p.SLF = slf; p.nslot = nslot; p.sm = m;
    const long long tiles = (N + NODE_TE - 1) / NODE_TE;
    p.grid = (int)(tiles < 1 ? 1 : (tiles > NUM_SMS ? NUM_SMS : tiles));
    q.nslot = nslot;
    *out = p;
    return true;
  }
  return false;
}

// ---- a GCP2 on its own (gcp2_op.cuh) ----------------------------------------------------------------------------------
inline int gcp2_n_params(const gcpnet_gcp2& d) {
  const int nfs = (d.flags & GCPNET_GCP2_NO_FRAMES) ? 0 : 9, gate = (d.flags & GCPNET_GCP2_NO_GATE) ? 0 : 1;
  return d.hd * d.vi + (nfs / 3) * d.vi + d.so * (d.si + d.hd + nfs) + d.so + d.vo * d.hd + gate * (d.vo * d.so + d.vo);
}
inline Gcp2OpPlan plan_gcp2_op(const gcpnet_gcp2& d, long long M) {
  Gcp2OpPlan P{};
  P.error = check_gcp2(d, "gcp2");
  if (!P.error.empty()) return P;
  P.n_params = gcp2_n_params(d);
  for (int i = 0; i < 7; ++i)
    if (d.grad_off[i] < 0 || d.grad_off[i] >= P.n_params + (d.vo == 0 ? 1 : 0)) { P.error = "gcp2: grad_off outside the module's flat gradient"; return P; }
  P.slf = (round_up(d.so, 16) / 16 + 3) / 4;
  if (P.slf > 2) { P.error = "gcp2: scalar output dim too wide for this build"; return P; }
  const int caps[3] = {EDGE_CAP_FLOATS, 6144, 4096};
  for (int ci = 0; ci < 3; ++ci) {
    GcpOp op = to_op(d, 0);
    int cursor = 0;
    P.error = plan_gcp_pack(op, caps[ci], &cursor);
    if (!P.error.empty()) continue;
    WSeq f{}, b{};
    seq_fwd(f, op); seq_bwd(b, op);
    f.slot_floats = round_up(f.slot_floats, 32); b.slot_floats = round_up(b.slot_floats, 32);
    bool ok = false;
    for (int nslot = 3; nslot >= 1 && !ok; --nslot) {
      const EdgeSmem mf = edge_plan_smem(GCP2OP_TE, d.so, d.vo, 0, 0, &op, 1, false, nslot, f.slot_floats);
      const EdgeSmem mb = edge_plan_smem(GCP2OP_TE, d.so, d.vo, 0, 0, &op, 1, true, nslot, b.slot_floats);
      if ((long long)mf.total * 4 <= SMEM_LIMIT_BYTES && (long long)mb.total * 4 <= SMEM_LIMIT_BYTES) {
        f.nslot = nslot; b.nslot = nslot; P.smf = mf; P.smb = mb; ok = true;
      }
    }
    if (!ok) { P.error = "gcp2: feature dims too large for the shared-memory tile plan of this build"; continue; }
    P.op = op; P.fwd = f; P.bwd = b; P.packed_floats = round_up(cursor, 32); P.error.clear();
    const long long tiles = (M + GCP2OP_TE - 1) / GCP2OP_TE;
    P.grid = (int)(tiles < 1 ? 1 : (tiles > NUM_SMS ? NUM_SMS : tiles));
    return P;
  }
  return P;
}

struct LayerPlan {
  LayerOps ops;
  EdgeTilePlan ef, eb;
  NodeTilePlan nf, nb;
  tc::TcPlan tc;      // tensor-core edge path (tc_edge.cuh); tc.ok == false -> FFMA tiles only
  int v2_packed_floats;
};
// rows per tile of the edge forward kernel this plan launches (layout of the segment sums, segment_total in gcp_tile.cuh)
inline int edge_tile_rows(const LayerPlan& lp) { return lp.tc.ok ? lp.tc.proto.rows : lp.ef.TE; }

// Spill area of the FFMA edge backward (ws_edge_spill): per message GCP dense row matrices gT [E][so -> 4], Z [E][K -> 4],
// gg [E][vo -> 4] in sorted edge order (EdgeParams::spill, node_wgrad.cuh).  Beyond EDGE_SPILL_MAX_FLOATS the tiles keep
// forming the products themselves.
constexpr long long EDGE_SPILL_MAX_FLOATS = 1ll << 31;  // 8 GB
constexpr int EDGE_WGRAD_CHUNK_ROWS = 4096;            // rows (edges) per CTA of the product kernel
struct EdgeSpill {
  long long gT[MAX_MSG_LAYERS], Z[MAX_MSG_LAYERS], GG[MAX_MSG_LAYERS], total;
  int ldg[MAX_MSG_LAYERS], ldz[MAX_MSG_LAYERS], ldgg[MAX_MSG_LAYERS];
  long long scratch; int nchunks, out_total;  // partial blocks of the row chunks: [nchunks][out_total] behind the operand rows
};
inline EdgeSpill edge_spill_layout(long long E, const LayerOps& ops) {
  EdgeSpill sp{};
  long long off = 0;
  for (int k = 0; k < ops.L; ++k) {
    const GcpOp& op = ops.msg[k];
    sp.ldg[k] = round_up(op.so, 4); sp.ldz[k] = round_up(gcp_k(op), 4); sp.ldgg[k] = round_up(op.vo, 4);
    sp.gT[k] = off; off += E * sp.ldg[k];
    sp.Z[k] = off; off += E * sp.ldz[k];
    sp.GG[k] = off; off += E * sp.ldgg[k];
    sp.out_total += op.so * gcp_k(op) + op.so + (gcp_gated(op) ? op.vo * op.so + op.vo : 0);
  }
  sp.nchunks = (int)((E + EDGE_WGRAD_CHUNK_ROWS - 1) / EDGE_WGRAD_CHUNK_ROWS);
  if (sp.nchunks < 1) sp.nchunks = 1;
  sp.scratch = off;
  if (sp.nchunks > 1) off += (long long)sp.nchunks * sp.out_total;
  sp.total = off;
  return sp;
}

// Spill area of the node backward (behind the per-CTA partial rows in ws_node_partial): per GCP dense row matrices
// gT [N][so -> 4], Z [N][K -> 4], gg [N][vo -> 4]  (BwdBufs::sp_*, node_wgrad.cuh).
struct NodeSpill { long long gT[3], Z[3], GG[3], total; int ldg[3], ldz[3], ldgg[3]; };
inline long long node_spill_offset(int grid, int n_node_params) { return ((long long)grid * n_node_params + 3) / 4 * 4; }  // 16-byte aligned rows
inline NodeSpill node_spill_layout(long long N, const LayerOps& ops, bool has_pos) {
  NodeSpill sp{};
  const GcpOp* op[3] = {&ops.ff0, &ops.ff1, has_pos ? &ops.pu : nullptr};
  long long off = 0;
  for (int k = 0; k < 3; ++k) {
    if (op[k] == nullptr) continue;
    sp.ldg[k] = round_up(op[k]->so, 4); sp.ldz[k] = round_up(gcp_k(*op[k]), 4); sp.ldgg[k] = round_up(op[k]->vo, 4);
    sp.gT[k] = off; off += N * sp.ldg[k];
    sp.Z[k] = off; off += N * sp.ldz[k];
    sp.GG[k] = off; off += N * sp.ldgg[k];
  }
  sp.total = off;
  return sp;
}

inline std::string make_layer_plan(const gcpnet_layer& l, long long N, long long E, LayerPlan* lp, gcpnet_plan* plan,
                                   bool tc_allowed = true) {
  std::string e = check_layer(l);
  if (!e.empty()) return e;
  // The packed layout (hence the ring-slot size) is shared by forward and backward: take the largest
  // chunk caps for which BOTH directions fit shared memory.
  const int edge_caps[3] = {EDGE_CAP_FLOATS, 6144, 4096};
  const int node_caps[4] = {NODE_CAP_FLOATS, 10240, 8192, 6144};
  bool okE = false, okN = false;
  int ecap = 0, ncap = 0;
  std::string last;
  for (int i = 0; i < 3 && !okE; ++i) {
    LayerOps o = layer_ops(l, edge_caps[i], node_caps[0]);
    if (!o.error.empty()) { last = o.error; continue; }
    EdgeTilePlan a, b;
    if (pick_edge_tile(l, o, E, false, 0, &a) && pick_edge_tile(l, o, E, true, 0, &b)) { okE = true; ecap = edge_caps[i]; }
  }
  for (int i = 0; i < 4 && !okN; ++i) {
    LayerOps o = layer_ops(l, edge_caps[0], node_caps[i]);
    if (!o.error.empty()) { last = o.error; continue; }
    NodeTilePlan a, b;
    if (pick_node_tile(l, o, N, false, &a) && pick_node_tile(l, o, N, true, &b)) { okN = true; ncap = node_caps[i]; }
  }
  if (!okE || !okN) return last.empty() ? "feature dims too large for the shared-memory tile plan of this build" : last;
  lp->ops = layer_ops(l, ecap, ncap);
  if (!lp->ops.error.empty()) return lp->ops.error;
  okE = pick_edge_tile(l, lp->ops, E, false, 0, &lp->ef) && pick_edge_tile(l, lp->ops, E, true, 0, &lp->eb);
  okN = pick_node_tile(l, lp->ops, N, false, &lp->nf) && pick_node_tile(l, lp->ops, N, true, &lp->nb);
  if (!okE || !okN) return "feature dims too large for the shared-memory tile plan of this build";
  lp->tc = tc::make_tc_plan(l, N, E);
  if (!tc_allowed) lp->tc.ok = false;
  lp->v2_packed_floats = round_up(lp->ops.packed_floats, 32);
  if (plan) {
    gcpnet_plan p{};
    p.edge_tile = lp->ef.TE; p.node_tile = lp->nf.TE;
    p.edge_grid_fwd = lp->ef.grid; p.edge_grid_bwd = lp->eb.grid;
    p.node_grid_fwd = lp->nf.grid; p.node_grid_bwd = lp->nb.grid;
    p.edge_smem_fwd_bytes = lp->ef.sm.total * 4; p.edge_smem_bwd_bytes = lp->eb.sm.total * 4;
    p.node_smem_fwd_bytes = lp->nf.sm.total * 4; p.node_smem_bwd_bytes = lp->nb.sm.total * 4;
    const long long W = l.s + 3 * l.v;
    {  // per-destination sums + two carry rows per edge tile (segment_total, gcp_tile.cuh)
      const long long rows = edge_tile_rows(*lp);
      p.agg_floats = N * W + 2 * ((E + rows - 1) / rows) * W;
    }
    long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS], tot;
    edge_saved_offsets(l, E, offT, offG, offS, offV, &tot);
    p.saved_edge_floats = tot;
    if (lp->tc.ok) p.saved_edge_floats = lp->tc.saved_floats;
    p.saved_node_floats = node_saved_layout((int)N, l.s, l.v, l.ff0.so, l.ff0.vo, l.has_pos != 0, l.training != 0).total;
    p.edge_partial_floats = (long long)p.edge_grid_bwd * l.n_edge_params;
    if (lp->tc.ok) p.edge_partial_floats = lp->tc.partial_floats;
    {
      const long long sp = edge_spill_layout(E, lp->ops).total;
      p.edge_spill_floats = (!lp->tc.ok && E > 0 && sp <= EDGE_SPILL_MAX_FLOATS) ? sp : 0;
    }
    p.node_partial_floats = node_spill_offset(p.node_grid_bwd, l.n_node_params) + node_spill_layout(N, lp->ops, l.has_pos != 0).total;
    p.edge_cotangent_floats = 2 * E * W;
    if (lp->tc.ok)  // [Y | A | G | Gn | node partials]
      p.edge_cotangent_floats = lp->tc.y_floats + lp->tc.a_floats + lp->tc.bproto.partial_stride + lp->tc.node_partial_stride +
                                (long long)lp->tc.node_partial_ctas * lp->tc.node_partial_stride;
    p.agg_cotangent_floats = N * W;
    p.packed_floats = lp->v2_packed_floats + (lp->tc.ok ? tc::rup(lp->tc.blob_floats, 32) + lp->tc.pq_floats : 0);
    p.tc_edge_path = lp->tc.ok ? 1 : 0;
    if (l.pre_norm) {  // [normalised input] ; backward: [cotangent of it | per-node (mean, rstd) | LN_PARTS x 2s partials]
      p.prenorm_floats = N * W;
      p.prenorm_ws_floats = N * W + 2 * N + (long long)LN_PARTS * 2 * l.s;
    }
    *plan = p;
  }
  return "";
}

inline EdgeParams make_edge_params(const gcpnet_layer& l, const gcpnet_graph& g, const LayerOps& ops, const EdgeTilePlan& tp,
                                   bool backward, const float* blob) {
  EdgeParams p{};
  p.N = (int)g.num_nodes; p.E = (int)g.num_edges; p.L = l.num_message_layers;
  p.s = l.s; p.v = l.v; p.se = l.se; p.ve = l.ve;
  p.residual = l.residual_messages; p.e3 = l.enable_e3; p.reduce_mean = l.reduce_mean; p.slope = l.slope;
  p.perm = g.perm; p.src = g.src; p.dst = g.dst; p.dst_ptr = g.dst_ptr;
  p.gsrc = g.gsrc != nullptr ? g.gsrc : g.src; p.gdst = g.gdst != nullptr ? g.gdst : g.dst;
  p.blob = blob;
  p.attn_w = l.attn_w; p.attn_b = l.attn_b; p.o_attn_w = l.attn_grad_off[0]; p.o_attn_b = l.attn_grad_off[1];
  for (int k = 0; k < p.L; ++k) p.ops[k] = ops.msg[k];
  long long tot;
  edge_saved_offsets(l, g.num_edges, p.offT, p.offG, p.offS, p.offV, &tot);
  p.partial_stride = l.n_edge_params;
  p.sm = tp.sm;
  p.seq = backward ? ops.edge_bwd : ops.edge_fwd;
  p.seq.nslot = tp.nslot;
  return p;
}

inline NodeParams make_node_params(const gcpnet_layer& l, const gcpnet_graph& g, const LayerOps& ops, const NodeTilePlan& tp,
                                   bool backward, const float* blob) {
  NodeParams p{};
  p.N = (int)g.num_nodes; p.s = l.s; p.v = l.v; p.hs = l.ff0.so; p.hv = l.ff0.vo;
  p.has_pos = l.has_pos; p.reduce_mean = l.reduce_mean; p.train = l.training;
  p.slope = l.slope; p.ln_eps = l.ln_eps; p.vn_eps = l.vn_eps; p.pos_weight = l.pos_weight; p.p_drop = l.training ? l.p_drop : 0.f;
  p.seed = l.seed; p.rng_ctr = (const long long*)l.rng_counter;
  p.fbar = g.fbar; p.dst_ptr = g.dst_ptr;
  p.fbar_pos = g.fbar_pos; p.mask = g.node_mask; p.pre_norm = l.pre_norm;
  p.e3 = l.enable_e3; p.src_ptr = g.src_ptr; p.src_pos = g.src_pos; p.perm = g.perm;  // p.frames: set by the caller (io)
  p.ln0_w = l.ln0_w; p.ln0_b = l.ln0_b; p.ln1_w = l.ln1_w; p.ln1_b = l.ln1_b;
  p.blob = blob;
  p.ff0 = ops.ff0; p.ff1 = ops.ff1; p.pu = ops.pu;
  p.sv = node_saved_layout(p.N, l.s, l.v, p.hs, p.hv, l.has_pos != 0, l.training != 0);
  p.partial_stride = l.n_node_params;
  p.o_ln0w = l.ln_grad_off[0] - l.n_edge_params; p.o_ln0b = l.ln_grad_off[1] - l.n_edge_params;
  p.o_ln1w = l.ln_grad_off[2] - l.n_edge_params; p.o_ln1b = l.ln_grad_off[3] - l.n_edge_params;
  p.sm = tp.sm;
  p.seq = backward ? ops.node_bwd : ops.node_fwd;
  p.seq.nslot = tp.nslot;
  return p;
}

}  // namespace gcp
