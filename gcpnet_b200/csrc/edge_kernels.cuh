// edge_kernels.cuh -- the fused edge-message path of one GCPNet layer, forward and backward (v2).
//
// Reference: GCPMessagePassing.message / .aggregate, src/models/components/gcpnet.py:898-947.
// One CTA processes a tile of TE edges taken in DESTINATION-SORTED order (CSR by `col`):
//   gather [h_row | e | h_col], [chi_row | xi | chi_col] and the edge frame into shared memory,
//   run the whole residual stack of L GCP2 modules on-chip (weights streamed through the shared-
//   memory ring by cp.async.bulk), emit the final message rows.
// The per-destination reduction is deterministic: the node kernel sums the (contiguous) message
// rows of each destination segment in a fixed order -- no atomics anywhere.
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

struct EdgeSmem {  // offsets in floats into dynamic shared memory
  int ZA, ldza, VA, ldva, Z, ldz, V, ldv, HD, ldhd, F, T, ldt, SG, ldsg, WSM, RING, MBAR;
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
  const int *gsrc, *gdst;                  // rows of (h, chi) gathered for the two ends (= src, dst; autoregressive layers:
                                           // rows of the [2N] gather table, gcpnet.py:1065-1116)
  const int* dst_ptr;                      // [N+1] CSR row pointer of the destination-sorted order
  const float* blob;                       // packed weights of the layer (pack.cuh)
  const float *attn_w, *attn_b;            // scalar message attention (gcpnet.py:931-934): m_s *= sigmoid(attn_w . m_s + attn_b); nullptr: off
  int o_attn_w, o_attn_b;                  // their gradient offsets inside a per-CTA partial row
  float* agg;                              // [N][s+3v] per-destination sums + [tiles][2][s+3v] carries (segment_total, gcp_tile.cuh)
  float* saved;                            // activations kept for backward (nullptr: inference)
  long long offT[MAX_MSG_LAYERS], offG[MAX_MSG_LAYERS], offS[MAX_MSG_LAYERS], offV[MAX_MSG_LAYERS];
  // backward
  const float* gagg;                       // [N][s+3v] cotangent of the aggregated messages
  float *grow, *gcol;                      // [E][s+3v] per-edge cotangents of (h,chi)[row] / (h,chi)[col], sorted order
  float *ge, *gxi;                         // [E][se], [E][3ve] caller's edge order
  float* partial;                          // [grid][partial_stride] per-CTA weight-gradient partials
  int partial_stride;
  // off-tile weight gradients of scalar_out / vector_out_scale: the tiles spill their rows of gT [so -> 4], Z [K -> 4] and
  // gg [vo -> 4] per message GCP (dense rows, sorted edge order) and edge_wgrad (node_wgrad.cuh) forms the products over ALL
  // edges -- instead of read-modify-writing 12 k partial sums per tile and GCP.  nullptr: the tiles form them.
  float* spill;
  long long sp_gT[MAX_MSG_LAYERS], sp_Z[MAX_MSG_LAYERS], sp_GG[MAX_MSG_LAYERS];
  EdgeSmem sm;
  GcpOp ops[MAX_MSG_LAYERS];
  WSeq seq;                                // chunk order of this kernel (forward or backward)
  long long* dbg;                          // development aid (GCP_STAMPS builds): phase stamps of CTA 0's first tile
};

GCP_HD TileBufs edge_bufs(const EdgeParams& p, float* sm, int k) {
  const EdgeSmem& L = p.sm;
  TileBufs b;
  if (k == 0) { b.Z = sm + L.ZA; b.ldz = L.ldza; b.V = sm + L.VA; b.ldv = L.ldva; }
  else { b.Z = sm + L.Z; b.ldz = L.ldz; b.V = sm + L.V; b.ldv = L.ldv; }
  b.HD = sm + L.HD; b.ldhd = L.ldhd; b.F = sm + L.F; b.T = sm + L.T; b.ldt = L.ldt;
  b.SG = sm + L.SG; b.ldsg = L.ldsg; b.WSM = sm + L.WSM;
  return b;
}

GCP_HD WPipe edge_pipe(const EdgeParams& p, float* sm, int ntiles_mine) {
  WPipe w;
  w.slots = sm + p.sm.RING;
  w.mbar = reinterpret_cast<unsigned long long*>(sm + p.sm.MBAR);
  w.blob = p.blob; w.seq = &p.seq; w.head = 0; w.total = ntiles_mine * p.seq.n;
  return w;
}

// gather the G0 inputs of a tile (gcpnet.py:911-917: [s_row | e | s_col], [v_row | xi | v_col]) and its frames
template <int TE, int NT>
GCP_HD void edge_gather_inputs(const EdgeParams& p, float* sm, int row0, int nrows, int tid) {
  const EdgeSmem& L = p.sm;
  float* ZA = sm + L.ZA; float* VA = sm + L.VA; float* F = sm + L.F;
  const int s = p.s, v3 = 3 * p.v, se = p.se, ve3 = 3 * p.ve;
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32) {
    float* zp = ZA + e * L.ldza; float* vp = VA + e * L.ldva; float* fp = F + e * LDF;
    if (e < nrows) {
      const int q = row0 + e;
      const size_t rs = (size_t)p.gsrc[q], rd = (size_t)p.gdst[q], rp = (size_t)p.perm[q];
      for (int f = lane; f < s; f += 32) { zp[f] = GCP_LDG(p.h + rs * s + f); zp[s + se + f] = GCP_LDG(p.h + rd * s + f); }
      for (int f = lane; f < se; f += 32) zp[s + f] = GCP_LDG(p.e + rp * se + f);
      for (int f = lane; f < v3; f += 32) { vp[f] = GCP_LDG(p.chi + rs * v3 + f); vp[v3 + ve3 + f] = GCP_LDG(p.chi + rd * v3 + f); }
      for (int f = lane; f < ve3; f += 32) vp[v3 + f] = GCP_LDG(p.xi + rp * ve3 + f);
      if (lane < 9) fp[lane] = GCP_LDG(p.frames + rp * 9 + lane);
    } else {
      for (int f = lane; f < 2 * s + se; f += 32) zp[f] = 0.f;
      for (int f = lane; f < 2 * v3 + ve3; f += 32) vp[f] = 0.f;
      if (lane < 9) fp[lane] = 0.f;
    }
  }
}

template <int TE, int NT, int SLF>
GCP_HDN void edge_fwd_tile(const EdgeParams& p, float* sm, int tile, WPipe& wp, bool first_tile) {
  const EdgeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.E - row0) < TE ? (p.E - row0) : TE;
  const int s = p.s, v = p.v, v3 = 3 * p.v;
  float* Zs = sm + L.Z; float* Vs = sm + L.V;  // running message state (scalars in Z[:, :s])
  GCP_PHASE_BEGIN(NT)
  if (!first_tile) wpipe_refill(wp, wp.head - 1, tid);  // G chunk of the previous tile's last GCP
  edge_gather_inputs<TE, NT>(p, sm, row0, nrows, tid);
  GCP_PHASE_END
  for (int k = 0; k < p.L; ++k) {
    const GcpOp& op = p.ops[k];
    const TileBufs b = edge_bufs(p, sm, k);
    const float* gch = gcp2_fwd_tile<TE, NT, SLF>(op, b, wp, p.e3, p.slope, k > 0);
    const float* wu = gch + op.w.o_wu;
    const bool add = (k > 0) && p.residual;
    const bool last = k == p.L - 1;
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    float* sT = p.saved != nullptr ? p.saved + p.offT[k] : nullptr;
    float* sG = p.saved != nullptr ? p.saved + p.offG[k] : nullptr;
    float* sS = (p.saved != nullptr && !last) ? p.saved + p.offS[k] : nullptr;
    float* sV = (p.saved != nullptr && !last) ? p.saved + p.offV[k] : nullptr;
    // scalar state update  r_s <- (r_s +) act_s(T)   (gcpnet.py:920-924); keep pre-activations for backward
    for (int e = tid >> 5; e < TE; e += NT / 32) {
      const float* tp = b.T + e * b.ldt;
      float* zp = Zs + e * L.ldz;
      const size_t q = (size_t)(row0 + e);
      for (int j = lane; j < s; j += 32) {
        const float t = tp[j];
        const float val = act_fwd(op.act_s, t, p.slope);
        const float r = add ? zp[j] + val : val;
        zp[j] = r;
        if (e < nrows) {
          if (sT) sT[q * s + j] = t;
          if (sS) sS[q * s + j] = r;
        }
      }
      if (sG && e < nrows) for (int o = lane; o < v; o += 32) sG[q * v + o] = b.SG[e * b.ldsg + o];
    }
    // vector state update  r_V <- (r_V +) U * sigmoid(gate)  (gcpnet.py:385-387): thread = (e, channel group)
    {
      const int e = tid % TE;
      const size_t q = (size_t)(row0 + e);
      for (int o = tid / TE; o < v; o += NT / TE) {
        const float sg = b.SG[e * b.ldsg + o];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float val = gcp2_vec_up(op, b, wu, e, o, x) * sg;
          float* vp = Vs + e * L.ldv + 3 * o + x;
          const float r = add ? *vp + val : val;
          *vp = r;
          if (e < nrows) {
            if (sV) sV[q * v3 + 3 * o + x] = r;
          }
        }
      }
    }
    GCP_PHASE_END
    wp.head++;  // G chunk released; the next phase (next GCP or next tile) refills its slot
  }
  // ---- scalar message attention (gcpnet.py:931-934): one thread per row
  if (p.attn_w != nullptr) {
    GCP_PHASE_BEGIN(NT)
    if (tid < TE) {
      float* zp = Zs + tid * L.ldz;
      float a = GCP_LDG(p.attn_b);
      for (int j = 0; j < s; ++j) a = fmaf(GCP_LDG(p.attn_w + j), zp[j], a);
      a = sigmoidf_(a);
      for (int j = 0; j < s; ++j) zp[j] *= a;
    }
    GCP_PHASE_END
  }
  // ---- aggregate (gcpnet.py:938-947) straight out of shared memory: the thread of a segment's FIRST row in the tile sums
  //      that destination's rows (fixed order), per column part
  GCP_PHASE_BEGIN(NT)
  const int e = tid % TE, part = tid / TE, W = s + v3;
  if (e < nrows) {
    const int q = row0 + e, d = p.dst[q];
    if (e == 0 || p.dst[q - 1] != d) {
      int e1 = e + 1;
      while (e1 < nrows && p.dst[row0 + e1] == d) ++e1;
      const int a = p.dst_ptr[d], b = p.dst_ptr[d + 1];
      float* out = (a >= row0 && b <= row0 + TE) ? p.agg + (size_t)d * W
                                                 : p.agg + (size_t)p.N * W + ((size_t)tile * 2 + (a < row0 ? 0 : 1)) * W;
      for (int f = part; f < W; f += NT / TE) {
        float acc = 0.f;
        for (int r = e; r < e1; ++r) acc += f < s ? Zs[r * L.ldz + f] : Vs[r * L.ldv + (f - s)];
        out[f] = acc;
      }
    }
  }
  GCP_PHASE_END
}

template <int TE, int NT, int SLF, int SLD>
GCP_HDN void edge_bwd_tile(const EdgeParams& p, float* sm, int tile, WPipe& wp, float* prow, bool accumulate) {
  const EdgeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.E - row0) < TE ? (p.E - row0) : TE;
  const int s = p.s, v = p.v, v3 = 3 * p.v, se = p.se, ve3 = 3 * p.ve, W = s + v3;
  BwdBufs g;
  g.GS = sm + L.GS; g.ldgs = L.ldgs; g.GV = sm + L.GV; g.ldgv = L.ldgv; g.GU = sm + L.GU; g.ldgu = L.ldgu;
  g.GG = sm + L.GG; g.ldgg = L.ldgg; g.GNQ = sm + L.GNQ; g.ldnq = L.ldnq; g.GHD = sm + L.GHD; g.ldghd = L.ldghd;
  for (int k = p.L - 1; k >= 0; --k) {
    const GcpOp& op = p.ops[k];
    const TileBufs b = edge_bufs(p, sm, k);
    if (p.spill != nullptr) {
      g.sp_gT = p.spill + p.sp_gT[k]; g.sp_Z = p.spill + p.sp_Z[k]; g.sp_GG = p.spill + p.sp_GG[k];
      g.sp_row0 = row0; g.sp_nrows = nrows;
    }
    GCP_PHASE_BEGIN(NT)
    const int lane = tid & 31;
    if (k == p.L - 1) {
      // cotangent of the final message: gagg[dst] (/ in-degree for the mean reduce, gcpnet.py:946); frames
      for (int e = tid >> 5; e < TE; e += NT / 32) {
        float* gs = g.GS + e * g.ldgs; float* gv = g.GV + e * g.ldgv; float* fp = sm + L.F + e * LDF;
        if (e < nrows) {
          const int d = p.dst[row0 + e];
          float scale = 1.f;
          if (p.reduce_mean) { const int deg = p.dst_ptr[d + 1] - p.dst_ptr[d]; scale = 1.f / (float)(deg > 1 ? deg : 1); }
          const float* gp = p.gagg + (size_t)d * W;
          for (int f = lane; f < s; f += 32) gs[f] = GCP_LDG(gp + f) * scale;
          for (int f = lane; f < v3; f += 32) gv[f] = GCP_LDG(gp + s + f) * scale;
          if (lane < 9) fp[lane] = GCP_LDG(p.frames + (size_t)p.perm[row0 + e] * 9 + lane);
        } else {
          for (int f = lane; f < s; f += 32) gs[f] = 0.f;
          for (int f = lane; f < v3; f += 32) gv[f] = 0.f;
          if (lane < 9) fp[lane] = 0.f;
        }
      }
    }
    // forward inputs of GCP k and its saved pre-activations / gates
    auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
    if (k == 0) {
      edge_gather_inputs<TE, NT>(p, sm, row0, nrows, tid);
    } else {
      tile_load_rows<TE, NT>(b.Z, b.ldz, p.saved + p.offS[k - 1], s, rr, tid);
      tile_load_rows<TE, NT>(b.V, b.ldv, p.saved + p.offV[k - 1], v3, rr, tid);
    }
    tile_load_rows<TE, NT>(b.T, b.ldt, p.saved + p.offT[k], s, rr, tid);
    tile_load_rows<TE, NT>(b.SG, b.ldsg, p.saved + p.offG[k], v, rr, tid);
    GCP_PHASE_END
    if (k == p.L - 1 && p.attn_w != nullptr) {
      // ---- scalar message attention backward.  The un-gated final scalars are r = (S_{L-2} +) act_s(T_{L-1}): both tiles
      //      were just loaded.  a = sigmoid(w . r + b); the message was r * a:
      //        g_r = g_m * a + (g_m . r) a (1 - a) w ;  g_w += (g_m . r) a (1 - a) r ;  g_b += (g_m . r) a (1 - a)
      const bool add = k > 0 && p.residual != 0;
      const int act = op.act_s;
      const float slope = p.slope;
      GCP_PHASE_BEGIN(NT)
      if (tid < TE) {
        const float* zp = b.Z + tid * b.ldz; const float* tp = b.T + tid * b.ldt;
        float* gs = g.GS + tid * g.ldgs;
        float a = GCP_LDG(p.attn_b), ga = 0.f;
        for (int j = 0; j < s; ++j) {
          const float r = (add ? zp[j] : 0.f) + act_fwd(act, tp[j], slope);
          a = fmaf(GCP_LDG(p.attn_w + j), r, a);
          ga = fmaf(gs[j], r, ga);
        }
        a = sigmoidf_(a);
        const float da = ga * a * (1.f - a);
        for (int j = 0; j < s; ++j) gs[j] = fmaf(gs[j], a, da * GCP_LDG(p.attn_w + j));
        g.GG[tid * g.ldgg] = da;  // (GG is scratch until the gate backward of this GCP)
      }
      GCP_PHASE_END
      GCP_PHASE_BEGIN(NT)
      for (int j = tid; j <= s; j += NT) {  // j == s: the bias
        float sum = 0.f;
        for (int e = 0; e < TE; ++e) {
          const float da = g.GG[e * g.ldgg];
          sum += j == s ? da : da * ((add ? b.Z[e * b.ldz + j] : 0.f) + act_fwd(act, b.T[e * b.ldt + j], slope));
        }
        float* dst = prow + (j == s ? p.o_attn_b : p.o_attn_w + j);
        *dst = (accumulate ? *dst : 0.f) + sum;
      }
      GCP_PHASE_END
    }
    // (gcp2_bwd_tile refills every chunk it releases, so nothing is pending on the ring here)
    const bool refill_first = false;
    if (k > 0) {
      const bool add = p.residual != 0;
      float* GS = g.GS; const int ldgs = g.ldgs; float* GV = g.GV; const int ldgv = g.ldgv;
      gcp2_bwd_tile<TE, NT, SLF, SLD>(
          op, b, g, wp, p.e3, p.slope, prow, accumulate, refill_first, EmitTile{GS, ldgs, add}, EmitTile{GV, ldgv, add});
    } else {
      float* grow = p.grow; float* gcol = p.gcol; float* ge = p.ge; float* gxi = p.gxi;
      const int* perm = p.perm;
      gcp2_bwd_tile<TE, NT, SLF, SLD>(
          op, b, g, wp, p.e3, p.slope, prow, accumulate, refill_first,
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
inline EdgeSmem edge_plan_smem(int TE, int s, int v, int se, int ve, const GcpOp* ops, int L, bool backward,
                               int nslot, int slot_floats) {
  EdgeSmem m{};
  int off = 0;
  auto take = [&](int floats) { const int o = off; off += round_up(floats, 4) + 8; return o; };  // 8 floats of slack
  int maxKpad = 0, maxHdCols = 0, maxSmall = 0;
  for (int k = 0; k < L; ++k) {
    if (k > 0) maxKpad = gcp_kpad(ops[k]) > maxKpad ? gcp_kpad(ops[k]) : maxKpad;
    const int c = ops[k].w.cols;
    maxHdCols = c > maxHdCols ? c : maxHdCols;
    const int small = ops[k].vi * ops[k].w.cols + ops[k].vo * ops[k].w.hdp;
    maxSmall = small > maxSmall ? small : maxSmall;
  }
  if (maxKpad == 0) maxKpad = round_up(s, 16);
  m.ldza = ld_vec(gcp_kpad(ops[0])); m.ZA = take(TE * m.ldza);
  m.ldva = ld_scal(3 * ops[0].vi); m.VA = take(TE * m.ldva);
  m.ldz = ld_vec(maxKpad); m.Z = take(TE * m.ldz);
  m.ldv = ld_scal(3 * v); m.V = take(TE * m.ldv);
  m.ldhd = ld_vec(3 * maxHdCols); m.HD = take(TE * m.ldhd);
  m.F = take(TE * LDF);
  m.ldt = ld_vec(s); m.T = take(TE * m.ldt);
  m.ldsg = ld_scal(v); m.SG = take(TE * m.ldsg);
  m.WSM = take(maxSmall);
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
  off = round_up(off, 32);  // ring slots 128-byte aligned
  m.RING = off; off += nslot * slot_floats;
  m.MBAR = off; off += 2 * MAX_WSLOTS;
  (void)se; (void)ve;
  m.total = off;
  return m;
}

}  // namespace gcp
