// node_kernels.cuh -- the node update of one GCPNet layer, forward and backward (v2).
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
// per-node mean frame `fbar` is computed once per graph batch (graph.cu).
//
// One CTA = TE nodes x PARTS = NT/TE threads per node.  Row-wise reductions (LayerNorm statistics)
// are two-level: PARTS partial sums per row in shared memory, combined by every thread of the row.
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

struct NodeSmem {
  int XS, ldxs, XV, ldxv, X2S, ldx2s, X2V, ldx2v, ZB, ldzb, VB, ldvb, T0, ldt0, T1, ldt1, SG0, ldsg0, SG1, ldsg1;
  int HD, ldhd, F, WSM, RED, RING, MBAR;
  // backward only
  int GXS, ldgxs, GXV, ldgxv, GS1, ldgs1, GV1, ldgv1, GS0, ldgs0, GV0, ldgv0, GU, ldgu, GG, ldgg, GNQ, ldnq, GHD, ldghd;
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
  const float *h, *chi, *agg, *fbar, *pos;
  const float* fbar_pos;        // node mask: mean frames of the position-update GCP (nullptr: fbar)
  const unsigned char* mask;    // node mask (nullptr: every node takes part): masked-out rows keep the layer input
  int pre_norm;                 // gcp_norm.1 after the first residual, no normalisation at the end (gcpnet.py:1223-1224,1245)
  int e3;                       // enable_e3_equivariance: |x_cross projections| per outgoing edge (NodeE3)
  const int *src_ptr, *src_pos, *perm;
  const float* frames;
  const int* dst_ptr;
  int edge_rows;                // rows per edge tile of the kernel that produced `agg` (segment_total)
  const float *ln0_w, *ln0_b, *ln1_w, *ln1_b;
  const float* blob;
  float *out_h, *out_chi, *out_pos;
  float* saved;                 // nullptr: inference
  NodeSavedLayout sv;
  // backward
  const float *g_out_h, *g_out_chi, *g_out_pos;
  float *g_x_h, *g_x_chi, *g_agg;
  float* partial; int partial_stride;
  int o_ln0w, o_ln0b, o_ln1w, o_ln1b;
  NodeSmem sm;
  GcpOp ff0, ff1, pu;
  WSeq seq;
  float* spill;                  // backward: operands of the large weight-gradient products (nullptr: per-CTA partials)
  long long sp_gT[3], sp_Z[3], sp_GG[3];   // float offsets in spill, per GCP (0: FF0, 1: FF1, 2: position update)
  long long* dbg;                // optional clock64 stamps of CTA 0 (development aid), entries [320 + i]
};

// counter-based uniform in [0,1): splitmix64 of (seed, counter, element)
GCP_HD float rng_uniform(unsigned long long seed, unsigned long long ctr, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1) + 0xBF58476D1CE4E5B9ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// ---- row statistics with PARTS threads per row -------------------------------------------------
// RED layout: [TE][PARTS][4] floats
constexpr int RED_W = 6;

#if GCP_DEVICE_CODE
template <int PARTS>
__device__ __forceinline__ float part_sum(float x) {  // butterfly over PARTS neighbouring lanes: same bits in every lane
#pragma unroll
  for (int o = PARTS / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
#endif

// GCPLayerNorm forward, in place on (S, V) (comp/__init__.py:138-167).  3 phases.
template <int TE, int NT>
GCP_HDN_NOINLINE void tile_layernorm_fwd(float* S, int lds, float* V, int ldv, int s, int v, const float* w, const float* bb,
                                float ln_eps, float vn_eps, float* RED) {
  constexpr int PARTS = NT / TE;
#if GCP_DEVICE_CODE
  // Device: the PARTS threads of a row are neighbouring lanes of one warp, row statistics by shuffles (one barrier).
  static_assert(PARTS <= 32 && (PARTS & (PARTS - 1)) == 0, "row parts must tile a warp");
  {
    const int tid = (int)threadIdx.x, e = tid / PARTS, part = tid % PARTS;
    float* sp = S + e * lds;
    float* vp = V + e * ldv;
    float sum = 0.f, m = 0.f;
    for (int j = part; j < s; j += PARTS) sum += sp[j];
    for (int c = part; c < v; c += PARTS) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
    }
    sum = part_sum<PARTS>(sum); m = part_sum<PARTS>(m);
    const float mean = sum / (float)s;
    float var = 0.f;
    for (int j = part; j < s; j += PARTS) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
    var = part_sum<PARTS>(var);
    const float rstd = 1.f / sqrtf(var / (float)s + ln_eps);
    for (int j = part; j < s; j += PARTS) sp[j] = fmaf((sp[j] - mean) * rstd, GCP_LDG(w + j), GCP_LDG(bb + j));
    const float inv = 1.f / sqrtf(m / (float)v);
    for (int c = part; c < 3 * v; c += PARTS) vp[c] *= inv;
    __syncthreads();
    return;
  }
#endif
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  const float* sp = S + e * lds;
  float sum = 0.f;
  for (int j = part; j < s; j += PARTS) sum += sp[j];
  const float* vp = V + e * ldv;
  float m = 0.f;
  for (int c = part; c < v; c += PARTS) {
    const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
    m += n2 > vn_eps ? n2 : vn_eps;
  }
  RED[(e * PARTS + part) * RED_W + 0] = sum;
  RED[(e * PARTS + part) * RED_W + 1] = m;
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  float sum = 0.f;
  for (int q = 0; q < PARTS; ++q) sum += RED[(e * PARTS + q) * RED_W + 0];
  const float mean = sum / (float)s;
  const float* sp = S + e * lds;
  float var = 0.f;
  for (int j = part; j < s; j += PARTS) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
  RED[(e * PARTS + part) * RED_W + 2] = var;
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  float sum = 0.f, var = 0.f, m = 0.f;
  for (int q = 0; q < PARTS; ++q) {
    sum += RED[(e * PARTS + q) * RED_W + 0]; m += RED[(e * PARTS + q) * RED_W + 1]; var += RED[(e * PARTS + q) * RED_W + 2];
  }
  const float mean = sum / (float)s;
  const float rstd = 1.f / sqrtf(var / (float)s + ln_eps);
  float* sp = S + e * lds;
  for (int j = part; j < s; j += PARTS) sp[j] = fmaf((sp[j] - mean) * rstd, GCP_LDG(w + j), GCP_LDG(bb + j));
  const float inv = 1.f / sqrtf(m / (float)v);
  float* vp = V + e * ldv;
  for (int c = part; c < 3 * v; c += PARTS) vp[c] *= inv;
  GCP_PHASE_END
}

// GCPLayerNorm backward: (S, V) = the layer-norm INPUT, (GS, GV) = cotangent of the output, replaced by the
// cotangent of the input; weight/bias gradients go to this CTA's partial row.  4 phases.
template <int TE, int NT>
GCP_HDN_NOINLINE void tile_layernorm_bwd(const float* S, int lds, const float* V, int ldv, float* GS, int ldgs, float* GV, int ldgv,
                                int s, int v, const float* w, float ln_eps, float vn_eps, float* RED,
                                float* pw, float* pb, bool accumulate) {
  constexpr int PARTS = NT / TE;
#if GCP_DEVICE_CODE
  {
    const int tid = (int)threadIdx.x, e = tid / PARTS, part = tid % PARTS;
    const float* sp = S + e * lds;
    const float* vp = V + e * ldv;
    float* gp = GS + e * ldgs;
    float* gv = GV + e * ldgv;
    float sum = 0.f, m = 0.f, dot = 0.f;
    for (int j = part; j < s; j += PARTS) sum += sp[j];
    for (int c = part; c < v; c += PARTS) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
      dot = fmaf(gv[3 * c], vp[3 * c], fmaf(gv[3 * c + 1], vp[3 * c + 1], fmaf(gv[3 * c + 2], vp[3 * c + 2], dot)));
    }
    sum = part_sum<PARTS>(sum); m = part_sum<PARTS>(m); dot = part_sum<PARTS>(dot);
    const float mean = sum / (float)s;
    float var = 0.f;
    for (int j = part; j < s; j += PARTS) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
    var = part_sum<PARTS>(var);
    const float rstd = 1.f / sqrtf(var / (float)s + ln_eps);
    if (part == 0) { RED[2 * e] = mean; RED[2 * e + 1] = rstd; }
    // vectors: y = V / r, r = sqrt(mean_c max(|V_c|^2, eps))   (row-local, in place)
    {
      const float rr = sqrtf(m / (float)v);
      const float coef = dot / ((float)v * rr * rr * rr);
      for (int c = part; c < v; c += PARTS) {
        const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
        const float ind = n2 > vn_eps ? 1.f : 0.f;
#pragma unroll
        for (int x = 0; x < 3; ++x) gv[3 * c + x] = gv[3 * c + x] / rr - coef * ind * vp[3 * c + x];
      }
    }
    float m1 = 0.f, m2 = 0.f;
    for (int j = part; j < s; j += PARTS) {
      const float gxh = gp[j] * GCP_LDG(w + j);
      m1 += gxh; m2 = fmaf(gxh, (sp[j] - mean) * rstd, m2);
    }
    m1 = part_sum<PARTS>(m1) / (float)s; m2 = part_sum<PARTS>(m2) / (float)s;
    __syncthreads();
    // scalar_norm.weight / .bias gradients: thread per column, fixed row order (reads GS before it is replaced)
    for (int j = tid; j < s; j += NT) {
      float gw = 0.f, gb = 0.f;
#pragma unroll 4
      for (int r = 0; r < TE; ++r) {
        const float gy = GS[r * ldgs + j];
        gw = fmaf(gy, (S[r * lds + j] - RED[2 * r]) * RED[2 * r + 1], gw);
        gb += gy;
      }
      pw[j] = (accumulate ? pw[j] : 0.f) + gw;
      pb[j] = (accumulate ? pb[j] : 0.f) + gb;
    }
    __syncthreads();
    for (int j = part; j < s; j += PARTS) {
      const float xhat = (sp[j] - mean) * rstd;
      gp[j] = rstd * (gp[j] * GCP_LDG(w + j) - m1 - xhat * m2);
    }
    __syncthreads();
    return;
  }
#endif
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  const float* sp = S + e * lds;
  float sum = 0.f;
  for (int j = part; j < s; j += PARTS) sum += sp[j];
  const float* vp = V + e * ldv;
  const float* gv = GV + e * ldgv;
  float m = 0.f, dot = 0.f;
  for (int c = part; c < v; c += PARTS) {
    const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
    m += n2 > vn_eps ? n2 : vn_eps;
    dot = fmaf(gv[3 * c], vp[3 * c], fmaf(gv[3 * c + 1], vp[3 * c + 1], fmaf(gv[3 * c + 2], vp[3 * c + 2], dot)));
  }
  float* r = RED + (e * PARTS + part) * RED_W;
  r[0] = sum; r[1] = m; r[2] = dot;
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  float sum = 0.f;
  for (int q = 0; q < PARTS; ++q) sum += RED[(e * PARTS + q) * RED_W + 0];
  const float mean = sum / (float)s;
  const float* sp = S + e * lds;
  float var = 0.f;
  for (int j = part; j < s; j += PARTS) { const float d = sp[j] - mean; var = fmaf(d, d, var); }
  RED[(e * PARTS + part) * RED_W + 3] = var;
  GCP_PHASE_END
  // row statistics are now complete in RED: mean (col 0), var (col 3), m (col 1), dot (col 2)
  GCP_PHASE_BEGIN(NT)
  // scalar_norm.weight / .bias gradients: thread per column, fixed row order
  for (int j = tid; j < s; j += NT) {
    float gw = 0.f, gb = 0.f;
    for (int e = 0; e < TE; ++e) {
      float sum = 0.f, var = 0.f;
      for (int q = 0; q < PARTS; ++q) { sum += RED[(e * PARTS + q) * RED_W + 0]; var += RED[(e * PARTS + q) * RED_W + 3]; }
      const float mean = sum / (float)s, rstd = 1.f / sqrtf(var / (float)s + ln_eps);
      const float gy = GS[e * ldgs + j];
      gw = fmaf(gy, (S[e * lds + j] - mean) * rstd, gw);
      gb += gy;
    }
    pw[j] = (accumulate ? pw[j] : 0.f) + gw;
    pb[j] = (accumulate ? pb[j] : 0.f) + gb;
  }
  GCP_PHASE_END
  // vector part (consumes m, dot) and the row sums m1 = mean_j(gy*w), m2 = mean_j(gy*w*xhat) of the scalar part
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  float sum = 0.f, var = 0.f, m = 0.f, dot = 0.f;
  for (int q = 0; q < PARTS; ++q) {
    const float* r = RED + (e * PARTS + q) * RED_W;
    sum += r[0]; m += r[1]; dot += r[2]; var += r[3];
  }
  // vectors: y = V / r, r = sqrt(mean_c max(|V_c|^2, eps))
  {
    const float* vp = V + e * ldv;
    float* gv = GV + e * ldgv;
    const float rr = sqrtf(m / (float)v);
    const float coef = dot / ((float)v * rr * rr * rr);
    for (int c = part; c < v; c += PARTS) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      const float ind = n2 > vn_eps ? 1.f : 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) gv[3 * c + x] = gv[3 * c + x] / rr - coef * ind * vp[3 * c + x];
    }
  }
  // scalars: full-row sums recomputed by every thread of the row (s is small; avoids another reduction level)
  {
    const float mean = sum / (float)s, rstd = 1.f / sqrtf(var / (float)s + ln_eps);
    const float* sp = S + e * lds;
    const float* gp = GS + e * ldgs;
    float m1 = 0.f, m2 = 0.f;
    for (int j = 0; j < s; ++j) {
      const float xhat = (sp[j] - mean) * rstd;
      const float gxh = gp[j] * GCP_LDG(w + j);
      m1 += gxh; m2 = fmaf(gxh, xhat, m2);
    }
    m1 /= (float)s; m2 /= (float)s;
    // stash the results (own columns 4/5: columns 1/2 are still being read by the other threads of the row);
    // the in-place update of GS happens in the next phase (other threads still read GS here)
    float* r = RED + (e * PARTS + part) * RED_W;
    r[4] = m1; r[5] = m2;
  }
  GCP_PHASE_END
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE;
  float sum = 0.f, var = 0.f;
  for (int q = 0; q < PARTS; ++q) { sum += RED[(e * PARTS + q) * RED_W + 0]; var += RED[(e * PARTS + q) * RED_W + 3]; }
  const float mean = sum / (float)s, rstd = 1.f / sqrtf(var / (float)s + ln_eps);
  const float m1 = RED[(e * PARTS + part) * RED_W + 4], m2 = RED[(e * PARTS + part) * RED_W + 5];
  const float* sp = S + e * lds;
  float* gp = GS + e * ldgs;
  for (int j = part; j < s; j += PARTS) {
    const float xhat = (sp[j] - mean) * rstd;
    gp[j] = rstd * (gp[j] * GCP_LDG(w + j) - m1 - xhat * m2);
  }
  GCP_PHASE_END
}

GCP_HD TileBufs node_bufs(const NodeParams& p, float* sm, int which) {  // 0: FF0, 1: FF1, 2: position update
  const NodeSmem& L = p.sm;
  TileBufs b;
  if (which == 1) { b.Z = sm + L.ZB; b.ldz = L.ldzb; b.V = sm + L.VB; b.ldv = L.ldvb; }
  else { b.Z = sm + L.XS; b.ldz = L.ldxs; b.V = sm + L.XV; b.ldv = L.ldxv; }
  if (which == 0) { b.T = sm + L.T0; b.ldt = L.ldt0; b.SG = sm + L.SG0; b.ldsg = L.ldsg0; }
  else { b.T = sm + L.T1; b.ldt = L.ldt1; b.SG = sm + L.SG1; b.ldsg = L.ldsg1; }
  b.HD = sm + L.HD; b.ldhd = L.ldhd; b.F = sm + L.F; b.WSM = sm + L.WSM;
  return b;
}

GCP_HD WPipe node_pipe(const NodeParams& p, float* sm, int ntiles_mine) {
  WPipe w;
  w.slots = sm + p.sm.RING;
  w.mbar = reinterpret_cast<unsigned long long*>(sm + p.sm.MBAR);
  w.blob = p.blob; w.seq = &p.seq; w.head = 0; w.total = ntiles_mine * p.seq.n;
  return w;
}

template <int TE, int NT, int SLF>
GCP_HDN void node_fwd_tile(const NodeParams& p, float* sm, int tile, WPipe& wp, bool first_tile) {
  const NodeSmem& L = p.sm;
  constexpr int PARTS = NT / TE;
  const int row0 = tile * TE;
  const int nrows = (p.N - row0) < TE ? (p.N - row0) : TE;
  const NodeE3 n3{p.src_ptr, p.src_pos, p.perm, p.frames, row0, nrows};
  const int s = p.s, v = p.v, v3 = 3 * p.v, W = s + v3, hs = p.hs, hv = p.hv, hv3 = 3 * p.hv;
  float* XS = sm + L.XS; float* XV = sm + L.XV; float* RED = sm + L.RED;
  const float keep_scale = 1.f / (1.f - p.p_drop);
  const unsigned long long ctr = (p.train && p.rng_ctr != nullptr) ? (unsigned long long)GCP_LDG(p.rng_ctr) : 0ull;
  // x1 = x + Dropout0(aggregate(messages)): warp per node, lane per feature (coalesced message rows)
  GCP_PHASE_BEGIN(NT)
  if (!first_tile && p.has_pos) wpipe_refill(wp, wp.head - 1, tid);  // G chunk of the previous tile's position GCP
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32) {
    if (e < nrows) {
      const int i = row0 + e;
      const int a = p.dst_ptr[i], bnd = p.dst_ptr[i + 1];
      for (int f = lane; f < W; f += 32) {
        float acc = segment_total(p.agg, p.N, W, p.edge_rows, p.dst_ptr, i, f);  // summed by the edge kernel's tiles
        if (p.reduce_mean && bnd - a > 1) acc = acc / (float)(bnd - a);
        if (p.train) {
          // scalar channels: elementwise; vector channels: one draw per (node, channel) shared by xyz (comp:113)
          const int ch = f < s ? f : s + (f - s) / 3;
          const float mk = rng_uniform(p.seed, ctr, ((unsigned long long)i * (s + v) + ch) * 2ull) >= p.p_drop ? keep_scale : 0.f;
          acc *= mk;
          if (p.saved != nullptr && (f < s || (f - s) % 3 == 0)) p.saved[p.sv.M0 + (size_t)i * (s + v) + ch] = mk;
        }
        const float x1 = (f < s ? GCP_LDG(p.h + (size_t)i * s + f) : GCP_LDG(p.chi + (size_t)i * v3 + (f - s))) + acc;
        if (p.saved != nullptr) p.saved[p.sv.X1 + (size_t)i * W + f] = x1;
        if (f < s) XS[e * L.ldxs + f] = x1; else XV[e * L.ldxv + (f - s)] = x1;
      }
      if (lane < 9) sm[L.F + e * LDF + lane] = GCP_LDG(p.fbar + (size_t)i * 9 + lane);
    } else {
      for (int f = lane; f < W; f += 32) { if (f < s) XS[e * L.ldxs + f] = 0.f; else XV[e * L.ldxv + (f - s)] = 0.f; }
      if (lane < 9) sm[L.F + e * LDF + lane] = 0.f;
    }
  }
  GCP_PHASE_END
  // normalisation after the first residual: gcp_norm.0, or gcp_norm.1 when pre_norm (gcpnet.py:1223-1226)
  const float* lnA_w = p.pre_norm ? p.ln1_w : p.ln0_w;
  const float* lnA_b = p.pre_norm ? p.ln1_b : p.ln0_b;
  tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, lnA_w, lnA_b, p.ln_eps, p.vn_eps, RED);
  // FF0: (s, v) -> (hs, hv)
  {
    TileBufs b = node_bufs(p, sm, 0);
    if (p.e3) b.e3n = &n3;
    const float* gch = gcp2_fwd_tile_call<TE, NT, SLF>(p.ff0, b, wp, p.e3, p.slope, false);
    const float* wu = gch + p.ff0.w.o_wu;
    float* ZB = sm + L.ZB; float* VB = sm + L.VB;
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    for (int e = tid >> 5; e < TE; e += NT / 32) {
      const size_t q = (size_t)(row0 + e);
      const bool sv = p.saved != nullptr && e < nrows;
      for (int j = lane; j < hs; j += 32) {
        const float t = b.T[e * b.ldt + j];
        ZB[e * L.ldzb + j] = act_fwd(p.ff0.act_s, t, p.slope);
        if (sv) p.saved[p.sv.T0 + q * hs + j] = t;
      }
      if (sv) for (int o = lane; o < hv; o += 32) p.saved[p.sv.SG0 + q * hv + o] = b.SG[e * b.ldsg + o];
    }
    {
      const int e = tid % TE;
      const size_t q = (size_t)(row0 + e);
      for (int o = tid / TE; o < hv; o += PARTS) {
        const float sg = b.SG[e * b.ldsg + o];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float val = gcp2_vec_up(p.ff0, b, wu, e, o, x) * sg;
          VB[e * L.ldvb + 3 * o + x] = val;
          if (p.saved != nullptr && e < nrows) p.saved[p.sv.VB + q * hv3 + 3 * o + x] = val;
        }
      }
    }
    GCP_PHASE_END
    wp.head++;
  }
  // FF1: (hs, hv) -> (s, v), then x2 = x1n + Dropout1(f)
  {
    TileBufs b = node_bufs(p, sm, 1);
    if (p.e3) b.e3n = &n3;
    const float* gch = gcp2_fwd_tile_call<TE, NT, SLF>(p.ff1, b, wp, p.e3, p.slope, true);
    const float* wu = gch + p.ff1.w.o_wu;
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    for (int e = tid >> 5; e < TE; e += NT / 32) {
      const int i = row0 + e;
      const bool live = e < nrows;
      for (int j = lane; j < s; j += 32) {
        const float t = b.T[e * b.ldt + j];
        float mk = 1.f;
        if (p.train && live) {
          mk = rng_uniform(p.seed, ctr, ((unsigned long long)i * (s + v) + j) * 2ull + 1ull) >= p.p_drop ? keep_scale : 0.f;
          if (p.saved != nullptr) p.saved[p.sv.M1 + (size_t)i * (s + v) + j] = mk;
        }
        const float x2 = XS[e * L.ldxs + j] + mk * act_fwd(p.ff1.act_s, t, p.slope);
        XS[e * L.ldxs + j] = x2;
        if (p.saved != nullptr && live) { p.saved[p.sv.T1 + (size_t)i * s + j] = t; p.saved[p.sv.X2 + (size_t)i * W + j] = x2; }
      }
      if (p.saved != nullptr && live) for (int o = lane; o < v; o += 32) p.saved[p.sv.SG1 + (size_t)i * v + o] = b.SG[e * b.ldsg + o];
    }
    {
      const int e = tid % TE;
      const int i = row0 + e;
      const bool live = e < nrows;
      for (int o = tid / TE; o < v; o += PARTS) {
        float mk = 1.f;
        if (p.train && live) {
          mk = rng_uniform(p.seed, ctr, ((unsigned long long)i * (s + v) + s + o) * 2ull + 1ull) >= p.p_drop ? keep_scale : 0.f;
          if (p.saved != nullptr) p.saved[p.sv.M1 + (size_t)i * (s + v) + s + o] = mk;
        }
        const float sg = b.SG[e * b.ldsg + o] * mk;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float x2 = XV[e * L.ldxv + 3 * o + x] + gcp2_vec_up(p.ff1, b, wu, e, o, x) * sg;
          XV[e * L.ldxv + 3 * o + x] = x2;
          if (p.saved != nullptr && live) p.saved[p.sv.X2 + (size_t)i * W + s + 3 * o + x] = x2;
        }
      }
    }
    GCP_PHASE_END
    wp.head++;
  }
  if (!p.pre_norm) tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln1_w, p.ln1_b, p.ln_eps, p.vn_eps, RED);
  const bool pos_frames = p.has_pos && p.fbar_pos != nullptr && p.fbar_pos != p.fbar;
  if (p.mask != nullptr || pos_frames) {
    // node mask: masked-out nodes keep the layer input (gcpnet.py:1249-1251); the position GCP sees all nodes with its own
    // mean frames (derive_x_update runs on the full graph, gcpnet.py:1258)
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    for (int e = tid >> 5; e < nrows; e += NT / 32) {
      const size_t i = (size_t)(row0 + e);
      if (p.mask != nullptr && !p.mask[i])
        for (int f = lane; f < W; f += 32) {
          if (f < s) XS[e * L.ldxs + f] = GCP_LDG(p.h + i * s + f); else XV[e * L.ldxv + (f - s)] = GCP_LDG(p.chi + i * v3 + (f - s));
        }
      if (pos_frames && lane < 9) sm[L.F + e * LDF + lane] = GCP_LDG(p.fbar_pos + i * 9 + lane);
    }
    GCP_PHASE_END
  }
  GCP_PHASE_BEGIN(NT)
  wpipe_refill(wp, wp.head - 1, tid);  // G chunk of FF1 (the LayerNorm phases do not touch the ring)
  tile_store_rows<TE, NT>(p.out_h, row0, s, XS, L.ldxs, nrows, tid);
  tile_store_rows<TE, NT>(p.out_chi, row0, v3, XV, L.ldxv, nrows, tid);
  GCP_PHASE_END
  if (p.has_pos) {
    TileBufs b = node_bufs(p, sm, 2);
    if (p.e3) b.e3n = &n3;
    const float* gch = gcp2_fwd_tile_call<TE, NT, SLF>(p.pu, b, wp, p.e3, p.slope, false);
    const float* wu = gch + p.pu.w.o_wu;
    GCP_PHASE_BEGIN(NT)
    if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.sv.TP, row0, s, b.T, b.ldt, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.sv.SGP, row0, 1, b.SG, b.ldsg, nrows, tid);
    }
    for (int item = tid; item < nrows * 3; item += NT) {
      const int e = item / 3, x = item - 3 * e;
      const float raw = gcp2_vec_up(p.pu, b, wu, e, 0, x) * b.SG[e * b.ldsg] * p.pos_weight;
      if (p.saved != nullptr) p.saved[p.sv.UPD + (size_t)(row0 + e) * 3 + x] = raw;
      const float upd = raw < -100.f ? -100.f : (raw > 100.f ? 100.f : raw);
      p.out_pos[(size_t)(row0 + e) * 3 + x] = GCP_LDG(p.pos + (size_t)(row0 + e) * 3 + x) + upd;
    }
    GCP_PHASE_END
    wp.head++;
  }
}

#if GCP_DEVICE_CODE && GCP_STAMPS
#define GCP_NSTAMP(i) do { if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[320 + (i)] = clock64(); } while (0)
#else
#define GCP_NSTAMP(i) do { } while (0)
#endif
template <int TE, int NT, int SLF, int SLD>
GCP_HDN void node_bwd_tile(const NodeParams& p, float* sm, int tile, WPipe& wp, float* prow, bool accumulate) {
  const NodeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.N - row0) < TE ? (p.N - row0) : TE;
  const int s = p.s, v = p.v, v3 = 3 * p.v, W = s + v3, hs = p.hs, hv = p.hv, hv3 = 3 * p.hv;
  float* XS = sm + L.XS; float* XV = sm + L.XV; float* X2S = sm + L.X2S; float* X2V = sm + L.X2V;
  float* GXS = sm + L.GXS; float* GXV = sm + L.GXV; float* RED = sm + L.RED;
  const int ldgxs = L.ldgxs, ldgxv = L.ldgxv;
  auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
  const NodeE3 n3{p.src_ptr, p.src_pos, p.perm, p.frames, row0, nrows};
  BwdBufs g;
  g.GU = sm + L.GU; g.ldgu = L.ldgu; g.GG = sm + L.GG; g.ldgg = L.ldgg; g.GNQ = sm + L.GNQ; g.ldnq = L.ldnq;
  g.GHD = sm + L.GHD; g.ldghd = L.ldghd;
  auto set_spill = [&](int which) {
    if (p.spill == nullptr) return;
    g.sp_gT = p.spill + p.sp_gT[which]; g.sp_Z = p.spill + p.sp_Z[which]; g.sp_GG = p.spill + p.sp_GG[which];
    g.sp_row0 = row0; g.sp_nrows = nrows;
  };
  const bool pos_frames = p.has_pos && p.fbar_pos != nullptr && p.fbar_pos != p.fbar;
  GCP_NSTAMP(0);
  // load x2 (raw copy + a copy that becomes out = LN1(x2)), the output cotangents, the mean frames
  GCP_PHASE_BEGIN(NT)
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32) {
    const bool live = e < nrows;
    const size_t i = (size_t)(row0 + e);
    for (int f = lane; f < W; f += 32) {
      float x2 = 0.f, gx = 0.f;
      if (live) {
        x2 = p.saved[p.sv.X2 + i * W + f];
        gx = f < s ? GCP_LDG(p.g_out_h + i * s + f) : GCP_LDG(p.g_out_chi + i * v3 + (f - s));
      }
      if (f < s) { X2S[e * L.ldx2s + f] = x2; XS[e * L.ldxs + f] = x2; GXS[e * ldgxs + f] = gx; }
      else { X2V[e * L.ldx2v + (f - s)] = x2; XV[e * L.ldxv + (f - s)] = x2; GXV[e * ldgxv + (f - s)] = gx; }
    }
    if (lane < 9) sm[L.F + e * LDF + lane] = live ? GCP_LDG((pos_frames ? p.fbar_pos : p.fbar) + i * 9 + lane) : 0.f;
  }
  GCP_PHASE_END
  GCP_NSTAMP(1);
  // ---- position update backward: only the vector output of P carries a cotangent (gcpnet.py:1129-1137,1156)
  if (p.has_pos) {
    if (!p.pre_norm) tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, p.ln1_w, p.ln1_b, p.ln_eps, p.vn_eps, RED);
    if (p.mask != nullptr) {  // the position GCP of a masked-out node read the layer input
      GCP_PHASE_BEGIN(NT)
      const int lane = tid & 31;
      for (int e = tid >> 5; e < nrows; e += NT / 32) {
        const size_t i = (size_t)(row0 + e);
        if (!p.mask[i])
          for (int f = lane; f < W; f += 32) {
            if (f < s) XS[e * L.ldxs + f] = GCP_LDG(p.h + i * s + f); else XV[e * L.ldxv + (f - s)] = GCP_LDG(p.chi + i * v3 + (f - s));
          }
      }
      GCP_PHASE_END
    }
    TileBufs b = node_bufs(p, sm, 2);
    if (p.e3) b.e3n = &n3;
    g.GS = sm + L.GS1; g.ldgs = L.ldgs1; g.GV = sm + L.GV1; g.ldgv = L.ldgv1;
    GCP_PHASE_BEGIN(NT)
    tile_load_rows<TE, NT>(b.T, b.ldt, p.saved + p.sv.TP, s, rr, tid);
    tile_load_rows<TE, NT>(b.SG, b.ldsg, p.saved + p.sv.SGP, 1, rr, tid);
    const int lane = tid & 31;
    for (int e = tid >> 5; e < TE; e += NT / 32) {
      for (int f = lane; f < s; f += 32) g.GS[e * g.ldgs + f] = 0.f;
      if (lane < 3) {
        float gv = 0.f;
        if (e < nrows) {
          const float raw = p.saved[p.sv.UPD + (size_t)(row0 + e) * 3 + lane];
          gv = (raw >= -100.f && raw <= 100.f) ? GCP_LDG(p.g_out_pos + (size_t)(row0 + e) * 3 + lane) * p.pos_weight : 0.f;
        }
        g.GV[e * g.ldgv + lane] = gv;
      }
    }
    GCP_PHASE_END
    set_spill(2);
    gcp2_bwd_tile_call<TE, NT, SLF, SLD>(
        p.pu, b, g, wp, p.e3, p.slope, prow, accumulate, false, EmitTile{GXS, ldgxs, true}, EmitTile{GXV, ldgxv, true});
  }
  GCP_NSTAMP(2);
  if (p.mask != nullptr || pos_frames) {
    // node mask: for a masked-out node the output IS the layer input -> its cotangent leaves here and nothing flows into
    // the update chain (zero cotangents give zero data and weight gradients row by row); frames of the feed-forward GCPs
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    for (int e = tid >> 5; e < nrows; e += NT / 32) {
      const size_t i = (size_t)(row0 + e);
      if (p.mask != nullptr && !p.mask[i])
        for (int f = lane; f < W; f += 32) {
          if (f < s) { p.g_x_h[i * s + f] = GXS[e * ldgxs + f]; GXS[e * ldgxs + f] = 0.f; }
          else { p.g_x_chi[i * v3 + (f - s)] = GXV[e * ldgxv + (f - s)]; GXV[e * ldgxv + (f - s)] = 0.f; }
          p.g_agg[i * W + f] = 0.f;
        }
      if (pos_frames && lane < 9) sm[L.F + e * LDF + lane] = GCP_LDG(p.fbar + i * 9 + lane);
    }
    GCP_PHASE_END
  }
  // ---- LayerNorm1 backward (input x2); pre_norm: there is no normalisation at the end
  const float* lnA_w = p.pre_norm ? p.ln1_w : p.ln0_w;
  const float* lnA_b = p.pre_norm ? p.ln1_b : p.ln0_b;
  const int o_lnAw = p.pre_norm ? p.o_ln1w : p.o_ln0w, o_lnAb = p.pre_norm ? p.o_ln1b : p.o_ln0b;
  if (!p.pre_norm)
    tile_layernorm_bwd<TE, NT>(X2S, L.ldx2s, X2V, L.ldx2v, GXS, ldgxs, GXV, ldgxv, s, v, p.ln1_w, p.ln_eps, p.vn_eps, RED,
                               prow + p.o_ln1w, prow + p.o_ln1b, accumulate);
  GCP_NSTAMP(3);
  // ---- x2 = x1n + Dropout1(f): cotangent of f, reload x1 (raw copy + copy that becomes x1n), FF inputs
  TileBufs b1 = node_bufs(p, sm, 1);
  TileBufs b0 = node_bufs(p, sm, 0);
  if (p.e3) { b1.e3n = &n3; b0.e3n = &n3; }
  float* GS1 = sm + L.GS1; float* GV1 = sm + L.GV1; float* GS0 = sm + L.GS0; float* GV0 = sm + L.GV0;
  GCP_PHASE_BEGIN(NT)
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32) {
    const bool live = e < nrows;
    const size_t i = (size_t)(row0 + e);
    for (int f = lane; f < W; f += 32) {
      float mk = 1.f, x1 = 0.f;
      if (live) {
        const int ch = f < s ? f : s + (f - s) / 3;
        if (p.train) mk = p.saved[p.sv.M1 + i * (s + v) + ch];
        x1 = p.saved[p.sv.X1 + i * W + f];
      }
      if (f < s) { GS1[e * L.ldgs1 + f] = GXS[e * ldgxs + f] * mk; X2S[e * L.ldx2s + f] = x1; XS[e * L.ldxs + f] = x1; }
      else { GV1[e * L.ldgv1 + (f - s)] = GXV[e * ldgxv + (f - s)] * mk; X2V[e * L.ldx2v + (f - s)] = x1; XV[e * L.ldxv + (f - s)] = x1; }
    }
  }
  tile_load_rows<TE, NT>(b0.T, b0.ldt, p.saved + p.sv.T0, hs, rr, tid);
  tile_load_rows<TE, NT>(b0.SG, b0.ldsg, p.saved + p.sv.SG0, hv, rr, tid);
  tile_load_rows<TE, NT>(b1.V, b1.ldv, p.saved + p.sv.VB, hv3, rr, tid);
  tile_load_rows<TE, NT>(b1.T, b1.ldt, p.saved + p.sv.T1, s, rr, tid);
  tile_load_rows<TE, NT>(b1.SG, b1.ldsg, p.saved + p.sv.SG1, v, rr, tid);
  GCP_PHASE_END
  GCP_NSTAMP(4);
  tile_layernorm_fwd<TE, NT>(XS, L.ldxs, XV, L.ldxv, s, v, lnA_w, lnA_b, p.ln_eps, p.vn_eps, RED);
  GCP_PHASE_BEGIN(NT)
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32)
    for (int j = lane; j < hs; j += 32) b1.Z[e * b1.ldz + j] = act_fwd(p.ff0.act_s, b0.T[e * b0.ldt + j], p.slope);
  GCP_PHASE_END
  GCP_NSTAMP(5);
  // ---- FF1 backward: cotangents (GS1, GV1) -> cotangents of FF0's outputs (GS0, GV0)
  {
    g.GS = GS1; g.ldgs = L.ldgs1; g.GV = GV1; g.ldgv = L.ldgv1;
    const int ldgs0 = L.ldgs0, ldgv0 = L.ldgv0;
    set_spill(1);
    gcp2_bwd_tile_call<TE, NT, SLF, SLD>(
        p.ff1, b1, g, wp, p.e3, p.slope, prow, accumulate, false, EmitTile{GS0, ldgs0, false}, EmitTile{GV0, ldgv0, false});
  }
  GCP_NSTAMP(6);
  // ---- FF0 backward: cotangents (GS0, GV0) -> accumulated into the cotangent of x1n
  {
    g.GS = GS0; g.ldgs = L.ldgs0; g.GV = GV0; g.ldgv = L.ldgv0;
    set_spill(0);
    gcp2_bwd_tile_call<TE, NT, SLF, SLD>(
        p.ff0, b0, g, wp, p.e3, p.slope, prow, accumulate, false, EmitTile{GXS, ldgxs, true}, EmitTile{GXV, ldgxv, true});
  }
  GCP_NSTAMP(7);
  // ---- backward of the normalisation after the first residual (input x1, kept raw in X2S/X2V)
  tile_layernorm_bwd<TE, NT>(X2S, L.ldx2s, X2V, L.ldx2v, GXS, ldgxs, GXV, ldgxv, s, v, lnA_w, p.ln_eps, p.vn_eps, RED,
                             prow + o_lnAw, prow + o_lnAb, accumulate);
  GCP_NSTAMP(8);
  // ---- x1 = x + Dropout0(m): direct cotangent of the layer input, cotangent of the aggregate
  GCP_PHASE_BEGIN(NT)
  const int lane = tid & 31;
  for (int e = tid >> 5; e < nrows; e += NT / 32) {
    const size_t i = (size_t)(row0 + e);
    if (p.mask != nullptr && !p.mask[i]) continue;  // masked-out node: cotangents written above
    for (int f = lane; f < W; f += 32) {
      const float gx = f < s ? GXS[e * ldgxs + f] : GXV[e * ldgxv + (f - s)];
      float mk = 1.f;
      if (p.train) mk = p.saved[p.sv.M0 + i * (s + v) + (f < s ? f : s + (f - s) / 3)];
      if (f < s) p.g_x_h[i * s + f] = gx; else p.g_x_chi[i * v3 + (f - s)] = gx;
      p.g_agg[i * W + f] = gx * mk;
    }
  }
  GCP_PHASE_END
  GCP_NSTAMP(9);
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

inline NodeSmem node_plan_smem(int TE, int NT, int s, int v, int hs, int hv, const GcpOp& ff0, const GcpOp& ff1, const GcpOp* pu,
                               bool backward, int nslot, int slot_floats) {
  NodeSmem m{};
  int off = 0;
  auto take = [&](int floats) { const int o = off; off += round_up(floats, 4) + 8; return o; };
  auto mx = [](int a, int b) { return a > b ? a : b; };
  int kx = gcp_kpad(ff0); if (pu) kx = mx(kx, gcp_kpad(*pu));
  int hdc = mx(ff0.w.cols, ff1.w.cols); if (pu) hdc = mx(hdc, pu->w.cols);
  auto small_of = [](const GcpOp& o) { return o.vi * o.w.cols + o.vo * o.w.hdp; };
  int small = mx(small_of(ff0), small_of(ff1));
  if (pu) small = mx(small, small_of(*pu));
  m.ldxs = ld_vec(kx); m.XS = take(TE * m.ldxs);
  m.ldxv = ld_scal(3 * v); m.XV = take(TE * m.ldxv);
  m.ldzb = ld_vec(gcp_kpad(ff1)); m.ZB = take(TE * m.ldzb);
  m.ldvb = ld_scal(3 * hv); m.VB = take(TE * m.ldvb);
  m.ldt0 = ld_vec(hs); m.T0 = take(TE * m.ldt0);
  m.ldt1 = ld_vec(s); m.T1 = take(TE * m.ldt1);
  m.ldsg0 = ld_scal(hv); m.SG0 = take(TE * m.ldsg0);
  m.ldsg1 = ld_scal(v); m.SG1 = take(TE * m.ldsg1);
  m.ldhd = ld_vec(3 * hdc); m.HD = take(TE * m.ldhd);
  m.F = take(TE * LDF);
  m.WSM = take(small);
  m.RED = take(NT * RED_W);
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
  }
  off = round_up(off, 32);
  m.RING = off; off += nslot * slot_floats;
  m.MBAR = off; off += 2 * MAX_WSLOTS;
  m.total = off;
  return m;
}

}  // namespace gcp
