// gcp2_op.cuh -- one GCP2 perceptron on its own (GCP2.forward, src/models/components/gcpnet.py:393-468), forward and
// backward: the reference applies it outside the interaction layers in GCPEmbedding (edge embedding with node_inputs=False,
// node embedding with node_inputs=True, gcpnet.py:735-823) and in the task heads.  Entities are rows: every row brings its
// scalars, its vector channels and ONE 3x3 frame -- the edge's frame for edge entities, the mean frame over the node's
// outgoing edges for node entities (node-side scalarize is linear in the frames).  Same tile
// routines as the edge kernels (gcp_tile.cuh): persistent CTAs, weights streamed through the shared-memory ring.
#pragma once
#include <string>

#include "edge_kernels.cuh"

namespace gcp {

struct Gcp2OpParams {
  int M;                                   // rows (edges or nodes)
  int e3;
  float slope;
  const float *s_in, *v_in, *frames;       // [M][si], [M][3 vi], [M][9]
  float *s_out, *v_out;                    // [M][so], [M][3 vo]
  float* saved;                            // [M][so] pre-activations, then [M][vo] gates (nullptr: inference)
  const float* blob;
  // backward
  const float *gs_out, *gv_out;
  float *gs_in, *gv_in;
  float* partial; int partial_stride;
  EdgeSmem sm;
  GcpOp op;
  WSeq seq;
};

GCP_HD TileBufs gcp2op_bufs(const Gcp2OpParams& p, float* sm) {
  const EdgeSmem& L = p.sm;
  TileBufs b;
  b.Z = sm + L.ZA; b.ldz = L.ldza; b.V = sm + L.VA; b.ldv = L.ldva;
  b.HD = sm + L.HD; b.ldhd = L.ldhd; b.F = sm + L.F; b.T = sm + L.T; b.ldt = L.ldt;
  b.SG = sm + L.SG; b.ldsg = L.ldsg; b.WSM = sm + L.WSM;
  return b;
}
GCP_HD WPipe gcp2op_pipe(const Gcp2OpParams& p, float* sm, int ntiles_mine) {
  WPipe w;
  w.slots = sm + p.sm.RING;
  w.mbar = reinterpret_cast<unsigned long long*>(sm + p.sm.MBAR);
  w.blob = p.blob; w.seq = &p.seq; w.head = 0; w.total = ntiles_mine * p.seq.n;
  return w;
}

template <int TE, int NT>
GCP_HD void gcp2op_load_inputs(const Gcp2OpParams& p, const TileBufs& b, int row0, int nrows, int tid) {
  auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
  tile_load_rows<TE, NT>(b.Z, b.ldz, p.s_in, p.op.si, rr, tid);
  tile_load_rows<TE, NT>(b.V, b.ldv, p.v_in, 3 * p.op.vi, rr, tid);
  tile_load_rows<TE, NT>(b.F, LDF, p.frames, 9, rr, tid);
}

template <int TE, int NT, int SLF>
GCP_HDN void gcp2op_fwd_tile(const Gcp2OpParams& p, float* sm, int tile, WPipe& wp, bool first_tile) {
  const int row0 = tile * TE;
  const int nrows = (p.M - row0) < TE ? (p.M - row0) : TE;
  const GcpOp& op = p.op;
  const TileBufs b = gcp2op_bufs(p, sm);
  GCP_PHASE_BEGIN(NT)
  if (!first_tile) wpipe_refill(wp, wp.head - 1, tid);  // G chunk of the previous tile
  gcp2op_load_inputs<TE, NT>(p, b, row0, nrows, tid);
  GCP_PHASE_END
  const float* gch = gcp2_fwd_tile<TE, NT, SLF>(op, b, wp, p.e3, p.slope, false);
  const float* wu = gch + op.w.o_wu;
  GCP_PHASE_BEGIN(NT)
  const int lane = tid & 31;
  const int so = op.so, vo = op.vo;
  float* sT = p.saved;
  float* sG = p.saved != nullptr ? p.saved + (size_t)p.M * so : nullptr;
  for (int e = tid >> 5; e < nrows; e += NT / 32) {
    const size_t q = (size_t)(row0 + e);
    for (int j = lane; j < so; j += 32) {
      const float t = b.T[e * b.ldt + j];
      p.s_out[q * so + j] = act_fwd(op.act_s, t, p.slope);  // (gcpnet.py:465)
      if (sT) sT[q * so + j] = t;
    }
    if (sG) for (int o = lane; o < vo; o += 32) sG[q * vo + o] = b.SG[e * b.ldsg + o];
  }
  {  // V' = U * sigmoid(gate)   (gcpnet.py:385-387)
    const int e = tid % TE;
    const size_t q = (size_t)(row0 + e);
    if (e < nrows)
      for (int o = tid / TE; o < vo; o += NT / TE) {
        const float sg = b.SG[e * b.ldsg + o];
#pragma unroll
        for (int x = 0; x < 3; ++x) p.v_out[q * 3 * vo + 3 * o + x] = gcp2_vec_up(op, b, wu, e, o, x) * sg;
      }
  }
  GCP_PHASE_END
  wp.head++;
}

template <int TE, int NT, int SLF, int SLD>
GCP_HDN void gcp2op_bwd_tile(const Gcp2OpParams& p, float* sm, int tile, WPipe& wp, float* prow, bool accumulate) {
  const EdgeSmem& L = p.sm;
  const int row0 = tile * TE;
  const int nrows = (p.M - row0) < TE ? (p.M - row0) : TE;
  const GcpOp& op = p.op;
  const TileBufs b = gcp2op_bufs(p, sm);
  BwdBufs g;
  g.GS = sm + L.GS; g.ldgs = L.ldgs; g.GV = sm + L.GV; g.ldgv = L.ldgv; g.GU = sm + L.GU; g.ldgu = L.ldgu;
  g.GG = sm + L.GG; g.ldgg = L.ldgg; g.GNQ = sm + L.GNQ; g.ldnq = L.ldnq; g.GHD = sm + L.GHD; g.ldghd = L.ldghd;
  GCP_PHASE_BEGIN(NT)
  auto rr = [=](int e) -> long long { return e < nrows ? row0 + e : -1; };
  gcp2op_load_inputs<TE, NT>(p, b, row0, nrows, tid);
  tile_load_rows<TE, NT>(g.GS, g.ldgs, p.gs_out, op.so, rr, tid);
  if (op.vo > 0) tile_load_rows<TE, NT>(g.GV, g.ldgv, p.gv_out, 3 * op.vo, rr, tid);
  tile_load_rows<TE, NT>(b.T, b.ldt, p.saved, op.so, rr, tid);
  if (op.vo > 0) tile_load_rows<TE, NT>(b.SG, b.ldsg, p.saved + (size_t)p.M * op.so, op.vo, rr, tid);
  GCP_PHASE_END
  float* gs_in = p.gs_in; float* gv_in = p.gv_in;
  const int si = op.si, vi3 = 3 * op.vi;
  gcp2_bwd_tile<TE, NT, SLF, SLD>(
      op, b, g, wp, p.e3, p.slope, prow, accumulate, false,
      [=](int e, int i, float val) { if (e < nrows) gs_in[(size_t)(row0 + e) * si + i] = val; },
      [=](int e, int c3, float val) { if (e < nrows) gv_in[(size_t)(row0 + e) * vi3 + c3] = val; });
}

// ---- host-side planning --------------------------------------------------------------------------------------------------
constexpr int GCP2OP_TE = 32, GCP2OP_NT = 256;
struct Gcp2OpPlan {
  GcpOp op;
  WSeq fwd, bwd;
  EdgeSmem smf, smb;
  int slf, grid, packed_floats, n_params;
  std::string error;
};

}  // namespace gcp
