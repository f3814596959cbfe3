// edge_kernels.cuh -- the fused edge-message path of one GCPNet layer, forward and backward.
//
// Reference: GCPMessagePassing.message / .aggregate, src/models/components/gcpnet.py:898-947.
// One CTA processes a tile of TE edges taken in DESTINATION-SORTED order (CSR by `col`):
//   gather [h_row | e | h_col], [chi_row | xi | chi_col] and the edge frame into shared memory,
//   run the whole residual stack of L GCP2 modules on-chip, emit the final message rows.
// The per-destination reduction is deterministic: the node kernel sums the (contiguous) message
// rows of each destination segment in a fixed order -- no atomics anywhere.
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

struct EdgeSmem {  // offsets in floats into dynamic shared memory
  int ZA, ldza, VA, ldva, Z, ldz, V, ldv, HD, ldhd, F, T, ldt, SG, ldsg, WC, wc_cap, WS;
  // backward only
  int GS, ldgs, GV, ldgv, GU, ldgu, GG, ldgg, GNQ, ldnq, GHD, ldghd;
  int total;  // floats
};

struct EdgeParams {
  int N, E, L;
  int s, v, se, ve;
  int residual, e3, reduce_mean;
  float slope;
  const float *h, *chi, *e, *xi, *frames;  // h[N][s] chi[N][3v] e[E][se] xi[E][3ve] frames[E][9] (caller's edge order)
  const int *perm, *src, *dst;             // sorted position p -> original edge id / source node / destination node
  const int* dst_ptr;                      // [N+1] CSR row pointer of the destination-sorted order
  GcpOp ops[MAX_MSG_LAYERS];
  float* msg;                              // [E][s+3v] final messages, sorted order
  float* saved;                            // activations kept for backward (nullptr: inference)
  long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS];
  // backward
  const float* gagg;                       // [N][s+3v] cotangent of the aggregated messages
  float *grow, *gcol;                      // [E][s+3v] per-edge cotangents of (h,chi)[row] / (h,chi)[col], sorted order
  float *ge, *gxi;                         // [E][se], [E][3ve] caller's edge order
  float* partial;                          // [grid][partial_stride] per-CTA weight-gradient partials
  int partial_stride;
  EdgeSmem sm;
};

// micro-tile grids (see tile_gemm_*): scalar_out GEMM 8 x (NT/8) threads, 8 outputs per thread
constexpr int E_OGM = 8, E_NRM = 8, E_OGG = 16, E_OGD = 8, E_NRD = 8;

GCP_HD TileBufs edge_bufs(const EdgeParams& p, float* sm, int k) {
  const EdgeSmem& L = p.sm;
  TileBufs b;
  if (k == 0) { b.Z = sm + L.ZA; b.ldz = L.ldza; b.V = sm + L.VA; b.ldv = L.ldva; }
  else { b.Z = sm + L.Z; b.ldz = L.ldz; b.V = sm + L.V; b.ldv = L.ldv; }
  b.HD = sm + L.HD; b.ldhd = L.ldhd; b.F = sm + L.F; b.T = sm + L.T; b.ldt = L.ldt;
  b.SG = sm + L.SG; b.ldsg = L.ldsg; b.WC = sm + L.WC; b.wc_cap = L.wc_cap; b.WS = sm + L.WS;
  return b;
}

// gather the G0 inputs of a tile (gcpnet.py:911-917: [s_row | e | s_col], [v_row | xi | v_col])
template <int TE, int NT>
GCP_HD void edge_gather_inputs(const EdgeParams& p, float* sm, int row0, int nrows, int tid) {
  const EdgeSmem& L = p.sm;
  float* ZA = sm + L.ZA; float* VA = sm + L.VA; float* F = sm + L.F;
  const int s = p.s, v3 = 3 * p.v, se = p.se, ve3 = 3 * p.ve;
  const int *src = p.src + row0, *dst = p.dst + row0, *perm = p.perm + row0;
  auto rs = [=](int e) -> long long { return e < nrows ? src[e] : -1; };
  auto rd = [=](int e) -> long long { return e < nrows ? dst[e] : -1; };
  auto rp = [=](int e) -> long long { return e < nrows ? perm[e] : -1; };
  tile_load_rows<TE, NT>(ZA, L.ldza, p.h, s, rs, tid);
  tile_load_rows<TE, NT>(ZA + s, L.ldza, p.e, se, rp, tid);
  tile_load_rows<TE, NT>(ZA + s + se, L.ldza, p.h, s, rd, tid);
  tile_load_rows<TE, NT>(VA, L.ldva, p.chi, v3, rs, tid);
  tile_load_rows<TE, NT>(VA + v3, L.ldva, p.xi, ve3, rp, tid);
  tile_load_rows<TE, NT>(VA + v3 + ve3, L.ldva, p.chi, v3, rd, tid);
  tile_load_rows<TE, NT>(F, LDF, p.frames, 9, rp, tid);
}

template <int TE, int NT>
GCP_HDN void edge_fwd_tile(const EdgeParams& p, float* sm, int tile) {
  const EdgeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.E - row0) < TE ? (p.E - row0) : TE;
  const int s = p.s, v3 = 3 * p.v;
  float* Zs = sm + L.Z; float* Vs = sm + L.V;  // running message state (scalars in Z[:, :s])
  GCP_PHASE_BEGIN(NT)
  edge_gather_inputs<TE, NT>(p, sm, row0, nrows, tid);
  GCP_PHASE_END
  for (int k = 0; k < p.L; ++k) {
    const GcpOp& op = p.ops[k];
    const TileBufs b = edge_bufs(p, sm, k);
    gcp2_fwd_tile<TE, NT, E_OGM, E_NRM, E_OGG>(op, b, p.e3, p.slope);
    const bool add = (k > 0) && p.residual;
    GCP_PHASE_BEGIN(NT)
    // keep pre-activations and gates for backward
    if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.offT[k], row0, s, b.T, b.ldt, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.offG[k], row0, p.v, b.SG, b.ldsg, nrows, tid);
    }
    // scalar state update  r_s <- (r_s +) act_s(T)          (gcpnet.py:920-924)
    for (int item = tid; item < TE * s; item += NT) {
      const int e = item / s, j = item - e * s;
      const float val = act_fwd(op.act_s, b.T[e * b.ldt + j], p.slope);
      float* z = Zs + e * L.ldz + j;
      *z = add ? *z + val : val;
    }
    // vector state update  r_V <- (r_V +) U * sigmoid(gate)  (gcpnet.py:385-387)
    for (int item = tid; item < TE * p.v; item += NT) {
      const int e = item / p.v, o = item - e * p.v;
      const float sg = b.SG[e * b.ldsg + o];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float val = gcp2_vec_up(op, b, e, o, x) * sg;
        float* vp = Vs + e * L.ldv + 3 * o + x;
        *vp = add ? *vp + val : val;
      }
    }
    GCP_PHASE_END
    GCP_PHASE_BEGIN(NT)
    if (k == p.L - 1) {
      // final message rows, flattened [scalars | vectors] (comp/__init__.py:61-63)
      for (int item = tid; item < nrows * (s + v3); item += NT) {
        const int e = item / (s + v3), f = item - e * (s + v3);
        p.msg[(size_t)(row0 + e) * (s + v3) + f] = f < s ? Zs[e * L.ldz + f] : Vs[e * L.ldv + (f - s)];
      }
    } else if (p.saved != nullptr) {
      tile_store_rows<TE, NT>(p.saved + p.offS[k], row0, s, Zs, L.ldz, nrows, tid);
      tile_store_rows<TE, NT>(p.saved + p.offV[k], row0, v3, Vs, L.ldv, nrows, tid);
    }
    GCP_PHASE_END
  }
}

template <int TE, int NT>
GCP_HDN void edge_bwd_tile(const EdgeParams& p, float* sm, int tile, float* prow, bool accumulate) {
  const EdgeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.E - row0) < TE ? (p.E - row0) : TE;
  const int s = p.s, v3 = 3 * p.v, se = p.se, ve3 = 3 * p.ve, W = s + v3;
  BwdBufs g;
  g.GS = sm + L.GS; g.ldgs = L.ldgs; g.GV = sm + L.GV; g.ldgv = L.ldgv; g.GU = sm + L.GU; g.ldgu = L.ldgu;
  g.GG = sm + L.GG; g.ldgg = L.ldgg; g.GNQ = sm + L.GNQ; g.ldnq = L.ldnq; g.GHD = sm + L.GHD; g.ldghd = L.ldghd;
  // cotangent of the final message: gagg[dst] (/ in-degree for the mean reduce, gcpnet.py:946)
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * W; item += NT) {
    const int e = item / W, f = item - e * W;
    float val = 0.f;
    if (e < nrows) {
      const int d = p.dst[row0 + e];
      val = GCP_LDG(p.gagg + (size_t)d * W + f);
      if (p.reduce_mean) {
        const int deg = p.dst_ptr[d + 1] - p.dst_ptr[d];
        val /= (float)(deg > 1 ? deg : 1);
      }
    }
    if (f < s) g.GS[e * g.ldgs + f] = val; else g.GV[e * g.ldgv + (f - s)] = val;
  }
  tile_load_rows<TE, NT>(sm + L.F, LDF, p.frames, 9,
                         [=](int e) -> long long { return e < nrows ? p.perm[row0 + e] : -1; }, tid);
  GCP_PHASE_END
  for (int k = p.L - 1; k >= 0; --k) {
    const GcpOp& op = p.ops[k];
    const TileBufs b = edge_bufs(p, sm, k);
    auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
    GCP_PHASE_BEGIN(NT)
    if (k == 0) {
      edge_gather_inputs<TE, NT>(p, sm, row0, nrows, tid);
    } else {
      tile_load_rows<TE, NT>(b.Z, b.ldz, p.saved + p.offS[k - 1], s, rr, tid);
      tile_load_rows<TE, NT>(b.V, b.ldv, p.saved + p.offV[k - 1], v3, rr, tid);
    }
    tile_load_rows<TE, NT>(b.T, b.ldt, p.saved + p.offT[k], s, rr, tid);
    tile_load_rows<TE, NT>(b.SG, b.ldsg, p.saved + p.offG[k], p.v, rr, tid);
    GCP_PHASE_END
    if (k > 0) {
      const bool add = p.residual != 0;
      float* GS = g.GS; const int ldgs = g.ldgs; float* GV = g.GV; const int ldgv = g.ldgv;
      gcp2_bwd_tile<TE, NT, E_OGM, E_NRM, E_OGD, E_NRD>(
          op, b, g, p.e3, p.slope, prow, accumulate,
          [=](int e, int i, float val) { float* d = GS + e * ldgs + i; *d = add ? *d + val : val; },
          [=](int e, int c3, float val) { float* d = GV + e * ldgv + c3; *d = add ? *d + val : val; });
    } else {
      float* grow = p.grow; float* gcol = p.gcol; float* ge = p.ge; float* gxi = p.gxi;
      const int* perm = p.perm;
      gcp2_bwd_tile<TE, NT, E_OGM, E_NRM, E_OGD, E_NRD>(
          op, b, g, p.e3, p.slope, prow, accumulate,
          [=](int e, int i, float val) {
            if (e >= nrows) return;
            if (i < s) grow[(size_t)(row0 + e) * W + i] = val;
            else if (i < s + se) ge[(size_t)perm[row0 + e] * se + (i - s)] = val;
            else gcol[(size_t)(row0 + e) * W + (i - s - se)] = val;
          },
          [=](int e, int c3, float val) {
            if (e >= nrows) return;
            if (c3 < v3) grow[(size_t)(row0 + e) * W + s + c3] = val;
            else if (c3 < v3 + ve3) gxi[(size_t)perm[row0 + e] * ve3 + (c3 - v3)] = val;
            else gcol[(size_t)(row0 + e) * W + s + (c3 - v3 - ve3)] = val;
          });
    }
  }
}

// host-side smem planning (shared by the launcher and the emulation) ---------------------------
inline EdgeSmem edge_plan_smem(int TE, int s, int v, int se, int ve, const GcpOp* ops, int L, bool backward, int wc_cap) {
  EdgeSmem m{};
  int off = 0;
  auto take = [&](int floats) { const int o = off; off += round_up(floats, 4) + 8; return o; };  // 8 floats of slack
  int maxK = 0, maxHdCols = 0, maxSmall = 0;
  for (int k = 0; k < L; ++k) {
    if (k > 0) maxK = gcp_k(ops[k]) > maxK ? gcp_k(ops[k]) : maxK;
    const int c = hd_cols(ops[k].hd);
    maxHdCols = c > maxHdCols ? c : maxHdCols;
    const int sm = gcp2_small_floats(ops[k].vi, ops[k].vo, ops[k].hd);
    maxSmall = sm > maxSmall ? sm : maxSmall;
  }
  if (maxK == 0) maxK = s;
  m.ldza = ld_vec(gcp_k(ops[0])); m.ZA = take(TE * m.ldza);
  m.ldva = ld_scal(3 * ops[0].vi); m.VA = take(TE * m.ldva);
  m.ldz = ld_vec(maxK); m.Z = take(TE * m.ldz);
  m.ldv = ld_scal(3 * v); m.V = take(TE * m.ldv);
  m.ldhd = ld_vec(3 * maxHdCols); m.HD = take(TE * m.ldhd);
  m.F = take(TE * LDF);
  m.ldt = ld_vec(s); m.T = take(TE * m.ldt);
  m.ldsg = ld_scal(v); m.SG = take(TE * m.ldsg);
  m.wc_cap = wc_cap; m.WC = take(wc_cap);
  m.WS = take(maxSmall);
  if (backward) {
    m.ldgs = ld_vec(s); m.GS = take(TE * m.ldgs);
    m.ldgv = ld_scal(3 * v); m.GV = take(TE * m.ldgv);
    m.ldgu = ld_scal(3 * v); m.GU = take(TE * m.ldgu);
    m.ldgg = ld_vec(v); m.GG = take(TE * m.ldgg);
    int maxnq = 0;
    for (int k = 0; k < L; ++k) maxnq = (ops[k].hd + 9) > maxnq ? (ops[k].hd + 9) : maxnq;
    m.ldnq = ld_scal(maxnq); m.GNQ = take(TE * m.ldnq);
    m.ldghd = m.ldhd; m.GHD = take(TE * m.ldghd);
  }
  (void)se; (void)ve;
  m.total = off;
  return m;
}

}  // namespace gcp
