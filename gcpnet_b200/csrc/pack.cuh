// pack.cuh -- weight packing: nn.Linear-layout parameters -> the chunked, padded, transposed layout
// the tile code streams through shared memory (see GcpW in gcp_tile.cuh).  Runs once per layer call
// (parameters change every optimizer step); one CTA per GCP2 module.
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

constexpr int MAX_PACK_GCPS = MAX_MSG_LAYERS + 3;
struct PackParams {
  int n;
  float* blob;
  GcpOp ops[MAX_PACK_GCPS];
};

// value of float `idx` of the S chunk: WdT[vi][cols]
GCP_HD float pack_S(const GcpOp& op, int idx) {
  const int cols = op.w.cols, hdp = op.w.hdp;
  const int c = idx / cols, k = idx - c * cols;
  if (c >= op.vi) return 0.f;
  if (k < op.hd) return GCP_LDG(op.Wd + k * op.vi + c);
  if (k >= hdp && k < hdp + 3 && gcp_nfs(op) != 0) return GCP_LDG(op.Wdf + (k - hdp) * op.vi + c);
  return 0.f;
}
// G chunk: WG[vo][ldg] | bg[round_up(vo,4)] | WU[vo][hdp]
GCP_HD float pack_G(const GcpOp& op, int idx) {
  const GcpW& W = op.w;
  if (idx < W.o_bg) {
    const int o = idx / W.ldg, n = idx - o * W.ldg;
    return (o < op.vo && n < op.so && gcp_gated(op)) ? GCP_LDG(op.Wg + o * op.so + n) : 0.f;
  }
  if (idx < W.o_wu) { const int o = idx - W.o_bg; return (o < op.vo && gcp_gated(op)) ? GCP_LDG(op.bg + o) : 0.f; }
  const int r = idx - W.o_wu;
  const int o = r / W.hdp, k = r - o * W.hdp;
  return (o < op.vo && k < op.hd) ? GCP_LDG(op.Wu + o * op.hd + k) : 0.f;
}
// WS chunk c: W[NP][ldk], then bias[NP]
GCP_HD float pack_WS(const GcpOp& op, int c, int idx) {
  const GcpW& W = op.w;
  const int K = gcp_k(op);
  if (idx < W.NP * W.ldk) {
    const int n = idx / W.ldk, kk = idx - n * W.ldk;
    const int col = c * W.kc + kk;
    return (n < op.so && kk < W.kc && col < K) ? GCP_LDG(op.Ws + (size_t)n * K + col) : 0.f;
  }
  const int n = idx - W.NP * W.ldk;
  return n < op.so ? GCP_LDG(op.bs + n) : 0.f;
}

// `nslices` CTAs (or one host call with nslices = 1) pack one GCP; slice = which CTA this is
template <int NT>
GCP_HDN void pack_gcp(const GcpOp& op, float* blob, int slice = 0, int nslices = 1) {
  GCP_PHASE_BEGIN(NT)
  const int t0 = slice * NT + tid, step = nslices * NT;
  for (int i = t0; i < op.w.S.floats; i += step) blob[op.w.S.off + i] = pack_S(op, i);
  for (int i = t0; i < op.w.G.floats; i += step) blob[op.w.G.off + i] = pack_G(op, i);
  for (int c = 0; c < op.w.nWS; ++c)
    for (int i = t0; i < op.w.ws_floats; i += step) blob[op.w.ws_off + c * op.w.ws_stride + i] = pack_WS(op, c, i);
  GCP_PHASE_END
}

}  // namespace gcp
