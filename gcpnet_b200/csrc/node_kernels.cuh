// node_kernels.cuh -- the node update of one GCPNet layer, forward and backward.
//
// Reference: GCPInteractions.forward, src/models/components/gcpnet.py:1160-1262 (pre_norm=False,
// two feed-forward GCPs, no node mask):
//   m   = mean/sum over the destination segment of the edge messages         (:946)
//   x1  = x + Dropout0(m) ;  x1n = GCPLayerNorm0(x1)                         (:1220-1226)
//   f   = FF1(FF0(x1n))   both GCP2 with node_inputs=True                    (:1232-1239)
//   x2  = x1n + Dropout1(f) ; out = GCPLayerNorm1(x2)                        (:1242-1246)
//   pos = pos + clamp(w * P(out).vector[:,0,:], +-100)                       (:1118-1158, :1258)
// For node inputs, scalarize() averages frames*D over the edges LEAVING the node
// (comp/__init__.py:286,316-323); D is the same for all of them, so q = mean_frame (x) D and the
// per-node mean frame `fbar` is computed once per graph batch (graph_kernels.cuh).
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

struct NodeSmem {
  int XS, ldxs, XV, ldxv, X2S, ldx2s, X2V, ldx2v, ZB, ldzb, VB, ldvb, T0, ldt0, T1, ldt1, SG0, ldsg0, SG1, ldsg1;
  int HD, ldhd, F, WC, wc_cap, WS;
  // backward only
  int GXS, ldgxs, GXV, ldgxv, GS1, ldgs1, GV1, ldgv1, GS0, ldgs0, GV0, ldgv0, GU, ldgu, GG, ldgg, GNQ, ldnq, GHD, ldghd, YA, ldya;
  int total;
};

struct NodeSavedLayout {  // offsets (floats) into the per-layer saved-activation buffer, each block is [N][width]
  long long X1, X2, T0, SG0, VB, T1, SG1, TP, SGP, UPD, M0, M1, total;
};

struct NodeParams {
  int N, s, v, hs, hv;          // hidden feed-forward dims (hs, hv) = (4s, 2v)   (gcpnet.py:1014)
  int has_pos, reduce_mean, train;
  float slope, ln_eps, vn_eps, pos_weight, p_drop;
  unsigned long long seed;
  const long long* rng_ctr;     // device counter, advanced by the host side once per training forward
  const float *h, *chi, *msg, *fbar, *pos;
  const int* dst_ptr;
  const float *ln0_w, *ln0_b, *ln1_w, *ln1_b;
  GcpOp ff0, ff1, pu;
  float *out_h, *out_chi, *out_pos;
  float* saved;                 // nullptr: inference
  NodeSavedLayout sv;
  // backward
  const float *g_out_h, *g_out_chi, *g_out_pos;
  float *g_x_h, *g_x_chi, *g_agg;
  float* partial; int partial_stride;
  int o_ln0w, o_ln0b, o_ln1w, o_ln1b;
  NodeSmem sm;
};

constexpr int N_OGM = 32, N_NRM = 4, N_OGG = 16, N_OGD = 8, N_NRD = 4;

// counter-based uniform in [0,1): splitmix64 of (seed, counter, element)
GCP_HD float rng_uniform(unsigned long long seed, unsigned long long ctr, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1) + 0xBF58476D1CE4E5B9ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// GCPLayerNorm forward on the rows of a tile, in place (comp/__init__.py:138-167)
template <int TE, int NT>
GCP_HD void tile_layernorm_fwd(float* S, int lds, float* V, int ldv, int s, int v, const float* w, const float* bb,
                               float ln_eps, float vn_eps, int tid) {
  for (int e = tid; e < TE; e += NT) {
    float* sp = S + e * lds;
    float mean = 0.f;
    for (int j = 0; j < s; ++j) mean += sp[j];
    mean /= (float)s;
    float var = 0.f;
    for (int j = 0; j < s; ++j) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
    const float rstd = 1.f / sqrtf(var / (float)s + ln_eps);
    for (int j = 0; j < s; ++j) sp[j] = fmaf((sp[j] - mean) * rstd, GCP_LDG(w + j), GCP_LDG(bb + j));
    float* vp = V + e * ldv;
    float m = 0.f;
    for (int c = 0; c < v; ++c) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
    }
    const float inv = 1.f / sqrtf(m / (float)v);
    for (int c = 0; c < 3 * v; ++c) vp[c] *= inv;
  }
}

// GCPLayerNorm backward, in two phases.
// (1) tile_layernorm_stats: per-row mean / rstd of the scalar input -> STAT[e*ldst + {0,1}]
// (2) tile_layernorm_param_grads: d/d(weight)[j] = sum_e gy*xhat, d/d(bias)[j] = sum_e gy  -> this CTA's partial row
// (3) tile_layernorm_bwd: G (cotangent of the output) is replaced by the cotangent of the input.
template <int TE, int NT>
GCP_HD void tile_layernorm_stats(const float* S, int lds, int s, float ln_eps, float* STAT, int ldst, int tid) {
  for (int e = tid; e < TE; e += NT) {
    const float* sp = S + e * lds;
    float mean = 0.f;
    for (int j = 0; j < s; ++j) mean += sp[j];
    mean /= (float)s;
    float var = 0.f;
    for (int j = 0; j < s; ++j) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
    STAT[e * ldst] = mean;
    STAT[e * ldst + 1] = 1.f / sqrtf(var / (float)s + ln_eps);
  }
}
template <int TE, int NT>
GCP_HD void tile_layernorm_param_grads(const float* S, int lds, const float* GS, int ldgs, const float* STAT, int ldst,
                                       int s, float* pw, float* pb, bool accumulate, int tid) {
  for (int j = tid; j < s; j += NT) {
    float gw = 0.f, gb = 0.f;
    for (int e = 0; e < TE; ++e) {
      const float gy = GS[e * ldgs + j];
      gw = fmaf(gy, (S[e * lds + j] - STAT[e * ldst]) * STAT[e * ldst + 1], gw);
      gb += gy;
    }
    pw[j] = (accumulate ? pw[j] : 0.f) + gw;
    pb[j] = (accumulate ? pb[j] : 0.f) + gb;
  }
}
template <int TE, int NT>
GCP_HD void tile_layernorm_bwd(const float* S, int lds, const float* V, int ldv, float* GS, int ldgs, float* GV, int ldgv,
                               const float* STAT, int ldst, int s, int v, const float* w, float vn_eps, int tid) {
  for (int e = tid; e < TE; e += NT) {
    const float* sp = S + e * lds;
    float* gp = GS + e * ldgs;
    const float mean = STAT[e * ldst], rstd = STAT[e * ldst + 1];
    float m1 = 0.f, m2 = 0.f;
    for (int j = 0; j < s; ++j) {
      const float xhat = (sp[j] - mean) * rstd;
      const float gxh = gp[j] * GCP_LDG(w + j);
      m1 += gxh; m2 = fmaf(gxh, xhat, m2);
    }
    m1 /= (float)s; m2 /= (float)s;
    for (int j = 0; j < s; ++j) {
      const float xhat = (sp[j] - mean) * rstd;
      gp[j] = rstd * (gp[j] * GCP_LDG(w + j) - m1 - xhat * m2);
    }
    // vectors: y = V / r, r = sqrt(mean_c max(|V_c|^2, eps))
    const float* vp = V + e * ldv;
    float* gv = GV + e * ldgv;
    float m = 0.f, dot = 0.f;
    for (int c = 0; c < v; ++c) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
      dot = fmaf(gv[3 * c], vp[3 * c], fmaf(gv[3 * c + 1], vp[3 * c + 1], fmaf(gv[3 * c + 2], vp[3 * c + 2], dot)));
    }
    const float r = sqrtf(m / (float)v);
    const float coef = dot / ((float)v * r * r * r);
    for (int c = 0; c < v; ++c) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      const float ind = n2 > vn_eps ? 1.f : 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) gv[3 * c + x] = gv[3 * c + x] / r - coef * ind * vp[3 * c + x];
    }
  }
}

GCP_HD TileBufs node_bufs(const NodeParams& p, float* sm, int which) {  // 0: FF0, 1: FF1, 2: position update
  const NodeSmem& L = p.sm;
  TileBufs b;
  if (which == 1) { b.Z = sm + L.ZB; b.ldz = L.ldzb; b.V = sm + L.VB; b.ldv = L.ldvb; }
  else { b.Z = sm + L.XS; b.ldz = L.ldxs; b.V = sm + L.XV; b.ldv = L.ldxv; }
  if (which == 0) { b.T = sm + L.T0; b.ldt = L.ldt0; b.SG = sm + L.SG0; b.ldsg = L.ldsg0; }
  else { b.T = sm + L.T1; b.ldt = L.ldt1; b.SG = sm + L.SG1; b.ldsg = L.ldsg1; }
  b.HD = sm + L.HD; b.ldhd = L.ldhd; b.F = sm + L.F; b.WC = sm + L.WC; b.wc_cap = L.wc_cap; b.WS = sm + L.WS;
  return b;
}

template <int TE, int NT>
GCP_HDN void node_fwd_tile(const NodeParams& p, float* sm, int tile) {
  const NodeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.N - row0) < TE ? (p.N - row0) : TE;
  const int s = p.s, v = p.v, v3 = 3 * p.v, W = s + v3, hs = p.hs, hv3 = 3 * p.hv;
  float* XS = sm + L.XS; float* XV = sm + L.XV;
  const float keep_scale = 1.f / (1.f - p.p_drop);
  const unsigned long long ctr = (p.train && p.rng_ctr != nullptr) ? (unsigned long long)GCP_LDG(p.rng_ctr) : 0ull;
  auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
  // x1 = x + Dropout0(aggregate(messages))
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * W; item += NT) {
    const int e = item / W, f = item - e * W;
    float x1 = 0.f;
    if (e < nrows) {
      const int i = row0 + e;
      const int a = p.dst_ptr[i], bnd = p.dst_ptr[i + 1];
      float acc = 0.f;
      for (int q = a; q < bnd; ++q) acc += GCP_LDG(p.msg + (size_t)q * W + f);
      if (p.reduce_mean && bnd - a > 1) acc /= (float)(bnd - a);
      if (p.train) {
        // scalar channels: elementwise; vector channels: one draw per (node, channel) shared by xyz (comp:113)
        const int ch = f < s ? f : s + (f - s) / 3;
        const float mk = rng_uniform(p.seed, ctr, ((unsigned long long)i * (s + v) + ch) * 2ull) >= p.p_drop ? keep_scale : 0.f;
        acc *= mk;
        if (p.saved != nullptr && (f < s || (f - s) % 3 == 0)) p.saved[p.sv.M0 + (size_t)i * (s + v) + ch] = mk;
      }
      x1 = (f < s ? GCP_LDG(p.h + (size_t)i * s + f) : GCP_LDG(p.chi + (size_t)i * v3 + (f - s))) + acc;
      if (p.saved != nullptr) p.saved[p.sv.X1 + (size_t)i * W + f] = x1;
    }
    if (f < s) XS[e * L.ldxs + f] = x1; else XV[e * L.ldxv + (f - s)] = x1;
  }
  tile_load_rows<TE, NT>(sm + L.F, LDF, p.fbar, 9, rr, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln0_w, p.ln0_b, p.ln_eps, p.vn_eps, tid);
  GCP_PHASE_END
  // FF0: (s, v) -> (hs, hv)
  {
    const TileBufs b = node_bufs(p, sm, 0);
    gcp2_fwd_tile<TE, NT, N_OGM, N_NRM, N_OGG>(p.ff0, b, 0, p.slope);
    float* ZB = sm + L.ZB; float* VB = sm + L.VB;
    GCP_PHASE_BEGIN(NT)
    if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.sv.T0, row0, hs, b.T, b.ldt, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.sv.SG0, row0, p.hv, b.SG, b.ldsg, nrows, tid);
    }
    for (int item = tid; item < TE * hs; item += NT) {
      const int e = item / hs, j = item - e * hs;
      ZB[e * L.ldzb + j] = act_fwd(p.ff0.act_s, b.T[e * b.ldt + j], p.slope);
    }
    for (int item = tid; item < TE * p.hv; item += NT) {
      const int e = item / p.hv, o = item - e * p.hv;
      const float sg = b.SG[e * b.ldsg + o];
#pragma unroll
      for (int x = 0; x < 3; ++x) VB[e * L.ldvb + 3 * o + x] = gcp2_vec_up(p.ff0, b, e, o, x) * sg;
    }
    GCP_PHASE_END
    if (p.saved != nullptr) {
      GCP_PHASE_BEGIN(NT)
      tile_store_rows<TE, NT>(p.saved + p.sv.VB, row0, hv3, VB, L.ldvb, nrows, tid);
      GCP_PHASE_END
    }
  }
  // FF1: (hs, hv) -> (s, v), then x2 = x1n + Dropout1(f)
  {
    const TileBufs b = node_bufs(p, sm, 1);
    gcp2_fwd_tile<TE, NT, N_OGM, N_NRM, N_OGG>(p.ff1, b, 0, p.slope);
    GCP_PHASE_BEGIN(NT)
    if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.sv.T1, row0, s, b.T, b.ldt, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.sv.SG1, row0, v, b.SG, b.ldsg, nrows, tid);
    }
    for (int item = tid; item < TE * (s + v); item += NT) {
      const int e = item / (s + v), ch = item - e * (s + v);
      const int i = row0 + e;
      float mk = 1.f;
      if (p.train && e < nrows) {
        mk = rng_uniform(p.seed, ctr, ((unsigned long long)i * (s + v) + ch) * 2ull + 1ull) >= p.p_drop ? keep_scale : 0.f;
        if (p.saved != nullptr) p.saved[p.sv.M1 + (size_t)i * (s + v) + ch] = mk;
      }
      if (ch < s) {
        XS[e * L.ldxs + ch] += mk * act_fwd(p.ff1.act_s, b.T[e * b.ldt + ch], p.slope);
      } else {
        const int o = ch - s;
        const float sg = b.SG[e * b.ldsg + o] * mk;
#pragma unroll
        for (int x = 0; x < 3; ++x) XV[e * L.ldxv + 3 * o + x] += gcp2_vec_up(p.ff1, b, e, o, x) * sg;
      }
    }
    GCP_PHASE_END
  }
  GCP_PHASE_BEGIN(NT)
  if (p.saved != nullptr) {
    for (int item = tid; item < nrows * W; item += NT) {
      const int e = item / W, f = item - e * W;
      p.saved[p.sv.X2 + (size_t)(row0 + e) * W + f] = f < s ? XS[e * L.ldxs + f] : XV[e * L.ldxv + (f - s)];
    }
  }
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln1_w, p.ln1_b, p.ln_eps, p.vn_eps, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_store_rows<TE, NT>(p.out_h, row0, s, XS, L.ldxs, nrows, tid);
  tile_store_rows<TE, NT>(p.out_chi, row0, v3, XV, L.ldxv, nrows, tid);
  GCP_PHASE_END
  if (p.has_pos) {
    const TileBufs b = node_bufs(p, sm, 2);
    gcp2_fwd_tile<TE, NT, N_OGM, N_NRM, N_OGG>(p.pu, b, 0, p.slope);
    GCP_PHASE_BEGIN(NT)
    if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.sv.TP, row0, s, b.T, b.ldt, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.sv.SGP, row0, 1, b.SG, b.ldsg, nrows, tid);
    }
    for (int item = tid; item < nrows * 3; item += NT) {
      const int e = item / 3, x = item - 3 * e;
      const float raw = gcp2_vec_up(p.pu, b, e, 0, x) * b.SG[e * b.ldsg] * p.pos_weight;
      if (p.saved != nullptr) p.saved[p.sv.UPD + (size_t)(row0 + e) * 3 + x] = raw;
      const float upd = raw < -100.f ? -100.f : (raw > 100.f ? 100.f : raw);
      p.out_pos[(size_t)(row0 + e) * 3 + x] = GCP_LDG(p.pos + (size_t)(row0 + e) * 3 + x) + upd;
    }
    GCP_PHASE_END
  }
}

template <int TE, int NT>
GCP_HDN void node_bwd_tile(const NodeParams& p, float* sm, int tile, float* prow, bool accumulate) {
  const NodeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.N - row0) < TE ? (p.N - row0) : TE;
  const int s = p.s, v = p.v, v3 = 3 * p.v, W = s + v3, hs = p.hs, hv = p.hv, hv3 = 3 * p.hv;
  float* XS = sm + L.XS; float* XV = sm + L.XV; float* X2S = sm + L.X2S; float* X2V = sm + L.X2V;
  float* GXS = sm + L.GXS; float* GXV = sm + L.GXV;
  const int ldgxs = L.ldgxs, ldgxv = L.ldgxv;
  auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
  BwdBufs g;
  g.GU = sm + L.GU; g.ldgu = L.ldgu; g.GG = sm + L.GG; g.ldgg = L.ldgg; g.GNQ = sm + L.GNQ; g.ldnq = L.ldnq;
  g.GHD = sm + L.GHD; g.ldghd = L.ldghd;
  float* STAT = sm + L.GNQ; const int ldst = L.ldnq;  // GNQ is only live inside gcp2_bwd_tile
  // load x2 (raw copy + a copy that becomes out = LN1(x2)), the output cotangents, the mean frames
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * W; item += NT) {
    const int e = item / W, f = item - e * W;
    float x2 = 0.f, gx = 0.f;
    if (e < nrows) {
      x2 = p.saved[p.sv.X2 + (size_t)(row0 + e) * W + f];
      gx = f < s ? GCP_LDG(p.g_out_h + (size_t)(row0 + e) * s + f) : GCP_LDG(p.g_out_chi + (size_t)(row0 + e) * v3 + (f - s));
    }
    if (f < s) { X2S[e * L.ldx2s + f] = x2; XS[e * L.ldxs + f] = x2; GXS[e * ldgxs + f] = gx; }
    else { X2V[e * L.ldx2v + (f - s)] = x2; XV[e * L.ldxv + (f - s)] = x2; GXV[e * ldgxv + (f - s)] = gx; }
  }
  tile_load_rows<TE, NT>(sm + L.F, LDF, p.fbar, 9, rr, tid);
  GCP_PHASE_END
  // ---- position update backward: only the vector output of P carries a cotangent (gcpnet.py:1129-1137,1156)
  if (p.has_pos) {
    GCP_PHASE_BEGIN(NT)
    tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln1_w, p.ln1_b, p.ln_eps, p.vn_eps, tid);
    GCP_PHASE_END
    const TileBufs b = node_bufs(p, sm, 2);
    g.GS = sm + L.GS1; g.ldgs = L.ldgs1; g.GV = sm + L.GV1; g.ldgv = L.ldgv1;
    GCP_PHASE_BEGIN(NT)
    tile_load_rows<TE, NT>(b.T, b.ldt, p.saved + p.sv.TP, s, rr, tid);
    tile_load_rows<TE, NT>(b.SG, b.ldsg, p.saved + p.sv.SGP, 1, rr, tid);
    for (int item = tid; item < TE * s; item += NT) { const int e = item / s; g.GS[e * g.ldgs + (item - e * s)] = 0.f; }
    for (int item = tid; item < TE * 3; item += NT) {
      const int e = item / 3, x = item - 3 * e;
      float gv = 0.f;
      if (e < nrows) {
        const float raw = p.saved[p.sv.UPD + (size_t)(row0 + e) * 3 + x];
        gv = (raw >= -100.f && raw <= 100.f) ? GCP_LDG(p.g_out_pos + (size_t)(row0 + e) * 3 + x) * p.pos_weight : 0.f;
      }
      g.GV[e * g.ldgv + x] = gv;
    }
    GCP_PHASE_END
    gcp2_bwd_tile<TE, NT, N_OGM, N_NRM, N_OGD, N_NRD>(
        p.pu, b, g, 0, p.slope, prow, accumulate,
        [=](int e, int i, float val) { GXS[e * ldgxs + i] += val; },
        [=](int e, int c3, float val) { GXV[e * ldgxv + c3] += val; });
  }
  // ---- LayerNorm1 backward (input x2)
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_stats<TE, NT>(X2S, L.ldx2s, s, p.ln_eps, STAT, ldst, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_param_grads<TE, NT>(X2S, L.ldx2s, GXS, ldgxs, STAT, ldst, s, prow + p.o_ln1w, prow + p.o_ln1b, accumulate, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_bwd<TE, NT>(X2S, L.ldx2s, X2V, L.ldx2v, GXS, ldgxs, GXV, ldgxv, STAT, ldst, s, v, p.ln1_w, p.vn_eps, tid);
  GCP_PHASE_END
  // ---- x2 = x1n + Dropout1(f): cotangent of f, reload x1 (raw copy + copy that becomes x1n), FF inputs
  const TileBufs b1 = node_bufs(p, sm, 1);
  const TileBufs b0 = node_bufs(p, sm, 0);
  float* GS1 = sm + L.GS1; float* GV1 = sm + L.GV1; float* GS0 = sm + L.GS0; float* GV0 = sm + L.GV0;
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * W; item += NT) {
    const int e = item / W, f = item - e * W;
    float mk = 1.f, x1 = 0.f;
    if (e < nrows) {
      const int ch = f < s ? f : s + (f - s) / 3;
      if (p.train) mk = p.saved[p.sv.M1 + (size_t)(row0 + e) * (s + v) + ch];
      x1 = p.saved[p.sv.X1 + (size_t)(row0 + e) * W + f];
    }
    if (f < s) { GS1[e * L.ldgs1 + f] = GXS[e * ldgxs + f] * mk; X2S[e * L.ldx2s + f] = x1; XS[e * L.ldxs + f] = x1; }
    else { GV1[e * L.ldgv1 + (f - s)] = GXV[e * ldgxv + (f - s)] * mk; X2V[e * L.ldx2v + (f - s)] = x1; XV[e * L.ldxv + (f - s)] = x1; }
  }
  tile_load_rows<TE, NT>(b0.T, b0.ldt, p.saved + p.sv.T0, hs, rr, tid);
  tile_load_rows<TE, NT>(b0.SG, b0.ldsg, p.saved + p.sv.SG0, hv, rr, tid);
  tile_load_rows<TE, NT>(b1.V, b1.ldv, p.saved + p.sv.VB, hv3, rr, tid);
  tile_load_rows<TE, NT>(b1.T, b1.ldt, p.saved + p.sv.T1, s, rr, tid);
  tile_load_rows<TE, NT>(b1.SG, b1.ldsg, p.saved + p.sv.SG1, v, rr, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln0_w, p.ln0_b, p.ln_eps, p.vn_eps, tid);
  for (int item = tid; item < TE * hs; item += NT) {
    const int e = item / hs, j = item - e * hs;
    b1.Z[e * b1.ldz + j] = act_fwd(p.ff0.act_s, b0.T[e * b0.ldt + j], p.slope);
  }
  GCP_PHASE_END
  // ---- FF1 backward: cotangents (GS1, GV1) -> cotangents of FF0's outputs (GS0, GV0)
  {
    g.GS = GS1; g.ldgs = L.ldgs1; g.GV = GV1; g.ldgv = L.ldgv1;
    const int ldgs0 = L.ldgs0, ldgv0 = L.ldgv0;
    gcp2_bwd_tile<TE, NT, N_OGM, N_NRM, N_OGD, N_NRD>(
        p.ff1, b1, g, 0, p.slope, prow, accumulate,
        [=](int e, int i, float val) { GS0[e * ldgs0 + i] = val; },
        [=](int e, int c3, float val) { GV0[e * ldgv0 + c3] = val; });
  }
  // ---- FF0 backward: cotangents (GS0, GV0) -> accumulated into the cotangent of x1n
  {
    g.GS = GS0; g.ldgs = L.ldgs0; g.GV = GV0; g.ldgv = L.ldgv0;
    gcp2_bwd_tile<TE, NT, N_OGM, N_NRM, N_OGD, N_NRD>(
        p.ff0, b0, g, 0, p.slope, prow, accumulate,
        [=](int e, int i, float val) { GXS[e * ldgxs + i] += val; },
        [=](int e, int c3, float val) { GXV[e * ldgxv + c3] += val; });
  }
  // ---- LayerNorm0 backward (input x1, kept raw in X2S/X2V)
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_stats<TE, NT>(X2S, L.ldx2s, s, p.ln_eps, STAT, ldst, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_param_grads<TE, NT>(X2S, L.ldx2s, GXS, ldgxs, STAT, ldst, s, prow + p.o_ln0w, prow + p.o_ln0b, accumulate, tid);
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  tile_layernorm_bwd<TE, NT>(X2S, L.ldx2s, X2V, L.ldx2v, GXS, ldgxs, GXV, ldgxv, STAT, ldst, s, v, p.ln0_w, p.vn_eps, tid);
  GCP_PHASE_END
  // ---- x1 = x + Dropout0(m): direct cotangent of the layer input, cotangent of the aggregate
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < nrows * W; item += NT) {
    const int e = item / W, f = item - e * W;
    const int i = row0 + e;
    const float gx = f < s ? GXS[e * ldgxs + f] : GXV[e * ldgxv + (f - s)];
    float mk = 1.f;
    if (p.train) mk = p.saved[p.sv.M0 + (size_t)i * (s + v) + (f < s ? f : s + (f - s) / 3)];
    if (f < s) p.g_x_h[(size_t)i * s + f] = gx; else p.g_x_chi[(size_t)i * v3 + (f - s)] = gx;
    p.g_agg[(size_t)i * W + f] = gx * mk;
  }
  GCP_PHASE_END
}

// host-side planning -------------------------------------------------------------------------
inline NodeSavedLayout node_saved_layout(int N, int s, int v, int hs, int hv, bool has_pos, bool train) {
  NodeSavedLayout l{};
  long long off = 0;
  auto take = [&](long long w) { const long long o = off; off += (long long)N * w; return o; };
  l.X1 = take(s + 3 * v); l.X2 = take(s + 3 * v); l.T0 = take(hs); l.SG0 = take(hv); l.VB = take(3 * hv);
  l.T1 = take(s); l.SG1 = take(v);
  l.TP = has_pos ? take(s) : 0; l.SGP = has_pos ? take(1) : 0; l.UPD = has_pos ? take(3) : 0;
  l.M0 = train ? take(s + v) : 0; l.M1 = train ? take(s + v) : 0;
  l.total = off;
  return l;
}

inline NodeSmem node_plan_smem(int TE, int s, int v, int hs, int hv, const GcpOp& ff0, const GcpOp& ff1, const GcpOp* pu,
                               bool backward, int wc_cap) {
  NodeSmem m{};
  int off = 0;
  auto take = [&](int floats) { const int o = off; off += round_up(floats, 4) + 8; return o; };
  auto mx = [](int a, int b) { return a > b ? a : b; };
  int kx = gcp_k(ff0); if (pu) kx = mx(kx, gcp_k(*pu));
  int hdc = mx(hd_cols(ff0.hd), hd_cols(ff1.hd)); if (pu) hdc = mx(hdc, hd_cols(pu->hd));
  int small = mx(gcp2_small_floats(ff0.vi, ff0.vo, ff0.hd), gcp2_small_floats(ff1.vi, ff1.vo, ff1.hd));
  if (pu) small = mx(small, gcp2_small_floats(pu->vi, pu->vo, pu->hd));
  m.ldxs = ld_vec(kx); m.XS = take(TE * m.ldxs);
  m.ldxv = ld_scal(3 * v); m.XV = take(TE * m.ldxv);
  m.ldzb = ld_vec(gcp_k(ff1)); m.ZB = take(TE * m.ldzb);
  m.ldvb = ld_scal(3 * hv); m.VB = take(TE * m.ldvb);
  m.ldt0 = ld_vec(hs); m.T0 = take(TE * m.ldt0);
  m.ldt1 = ld_vec(s); m.T1 = take(TE * m.ldt1);
  m.ldsg0 = ld_scal(hv); m.SG0 = take(TE * m.ldsg0);
  m.ldsg1 = ld_scal(v); m.SG1 = take(TE * m.ldsg1);
  m.ldhd = ld_vec(3 * hdc); m.HD = take(TE * m.ldhd);
  m.F = take(TE * LDF);
  m.wc_cap = wc_cap; m.WC = take(wc_cap);
  m.WS = take(small);
  if (backward) {
    m.ldx2s = ld_vec(s); m.X2S = take(TE * m.ldx2s);
    m.ldx2v = ld_scal(3 * v); m.X2V = take(TE * m.ldx2v);
    m.ldgxs = ld_vec(s); m.GXS = take(TE * m.ldgxs);
    m.ldgxv = ld_scal(3 * v); m.GXV = take(TE * m.ldgxv);
    m.ldgs1 = ld_vec(s); m.GS1 = take(TE * m.ldgs1);
    m.ldgv1 = ld_scal(3 * v); m.GV1 = take(TE * m.ldgv1);
    m.ldgs0 = ld_vec(hs); m.GS0 = take(TE * m.ldgs0);
    m.ldgv0 = ld_scal(3 * hv); m.GV0 = take(TE * m.ldgv0);
    m.ldgu = ld_scal(3 * hv); m.GU = take(TE * m.ldgu);
    m.ldgg = ld_vec(hv); m.GG = take(TE * m.ldgg);
    int nq = mx(ff0.hd, ff1.hd) + 9; if (pu) nq = mx(nq, pu->hd + 9);
    m.ldnq = ld_scal(nq); m.GNQ = take(TE * m.ldnq);
    m.ldghd = m.ldhd; m.GHD = take(TE * m.ldghd);
    m.ldya = 0; m.YA = 0;
  }
  m.total = off;
  return m;
}

}  // namespace gcp
