// node_wgrad.cuh -- weight gradients of the node-update GCPs' two large Linear layers (scalar_out, vector_out_scale)
// as ONE output-parallel product over all nodes:  dW[j][i] = sum_n G[n][j] * fmap(Z[n][i]),  db[j] = sum_n G[n][j].
//
// Reference: what autograd derives for GCP2.scalar_out / vector_out_scale (gcpnet.py:441, :386) inside the feed-forward and
// position-update GCPs of GCPInteractions.forward (gcpnet.py:1232-1239, :1129-1137).  The node backward tiles spill their
// rows of G (cotangent of the Linear's output) and Z (its input); here a CTA owns a 16 x 32 block of dW, its 8 warps split
// the node rows, the products run on mma.sync.m16n8k8 (3xTF32, fp32 accumulation) and the 8 partial blocks are summed in
// fixed order -- deterministic, no atomics, and off the critical path (side stream).
#pragma once
#include "gcp_tile.cuh"

namespace gcp {

constexpr int NODE_WGRAD_MAX_JOBS = 2 * MAX_MSG_LAYERS;  // node update: <= 6; FFMA edge backward: two per message GCP
struct NodeWgradJob {
  const float* G; const float* Z;  // [N][ldg], [N][ldz]
  float* outW; float* outb;        // [J][I], [J]
  int ldg, J, ldz, I;
  int act;                         // activation applied to Z on the fly (vector_out_scale reads act_v(T))
  int cta0, JB, IG;                // first CTA of the job, 16-row blocks of J, 32-column groups of I
  int out0;                        // chunked mode: offset of this job's [J][I] | [J] block inside one scratch row
};
// Many rows (the edge path: rows = edges): grid.y row chunks of `chunk_rows` rows each write their partial blocks into
// scratch[chunk][out_total]; wgrad_chunk_reduce_kernel adds the chunks in order.  nchunks == 1: results go straight to outW / outb.
struct NodeWgradParams {
  int N, njobs;
  float slope;
  int chunk_rows, nchunks, out_total;
  float* scratch;
  NodeWgradJob job[NODE_WGRAD_MAX_JOBS];
};

__global__ void __launch_bounds__(256) node_wgrad_kernel(const __grid_constant__ NodeWgradParams p) {
#if GCP_DEVICE_CODE
  __shared__ float red[8][16][32];
  int jn = 0;
  while (jn + 1 < p.njobs && (int)blockIdx.x >= p.job[jn + 1].cta0) ++jn;
  const NodeWgradJob& jb = p.job[jn];
  const int local = (int)blockIdx.x - jb.cta0;
  const int ig = local / jb.JB, jblk = local - ig * jb.JB;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int J = jb.J, I = jb.I;
  const int nbeg = p.nchunks > 1 ? (int)blockIdx.y * p.chunk_rows : 0;
  const int N = p.nchunks > 1 ? (nbeg + p.chunk_rows < p.N ? nbeg + p.chunk_rows : p.N) : p.N;  // rows [nbeg, N)
  float* outW = p.nchunks > 1 ? p.scratch + (size_t)blockIdx.y * p.out_total + jb.out0 : jb.outW;
  float* outb = p.nchunks > 1 ? outW + (size_t)J * I : jb.outb;
  const int IB = (I + 7) >> 3;
  const int nb = IB - 4 * ig < 4 ? IB - 4 * ig : 4;
  const int j0 = 16 * jblk + g;
  const int ja = j0 < J ? j0 : J - 1, jc = j0 + 8 < J ? j0 + 8 : J - 1;  // clamped loads; results dropped below
  int ii[4];
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) { const int i0 = 8 * (4 * ig + bb) + g; ii[bb] = i0 < I ? i0 : I - 1; }
  float c[4][4];
#pragma unroll
  for (int bb = 0; bb < 4; ++bb)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[bb][q] = 0.f;
  const int act = jb.act;
  const float slope = p.slope;
  const bool do_bias = ig == 0 && jb.outb != nullptr;  // bias gradient = column sums of G: rides on the A fragments
  float ba = 0.f, bc = 0.f;
#pragma unroll 4
  for (int n0 = nbeg + 8 * warp; n0 < N; n0 += 64) {
    const int r0 = n0 + t, r1 = n0 + t + 4;
    const bool v0 = r0 < N, v1 = r1 < N;
    const float* g0 = jb.G + (size_t)(v0 ? r0 : 0) * jb.ldg;
    const float* g1 = jb.G + (size_t)(v1 ? r1 : 0) * jb.ldg;
    const float* z0 = jb.Z + (size_t)(v0 ? r0 : 0) * jb.ldz;
    const float* z1 = jb.Z + (size_t)(v1 ? r1 : 0) * jb.ldz;
    const float fa[4] = {v0 ? __ldg(g0 + ja) : 0.f, v0 ? __ldg(g0 + jc) : 0.f, v1 ? __ldg(g1 + ja) : 0.f, v1 ? __ldg(g1 + jc) : 0.f};
    float fb[4][2];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      fb[bb][0] = (v0 && bb < nb) ? __ldg(z0 + ii[bb]) : 0.f;
      fb[bb][1] = (v1 && bb < nb) ? __ldg(z1 + ii[bb]) : 0.f;
    }
    if (do_bias) { ba += fa[0] + fa[2]; bc += fa[1] + fa[3]; }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { ah[q] = __float_as_uint(fa[q]) & 0xffffe000u; al[q] = __float_as_uint(fa[q] - __uint_as_float(ah[q])); }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      if (bb < nb) {
        uint32_t bh[2], bl[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float x = act_fwd(act, fb[bb][q], slope);
          bh[q] = __float_as_uint(x) & 0xffffe000u; bl[q] = __float_as_uint(x - __uint_as_float(bh[q]));
        }
        wg_hmma(c[bb], ah, bh);
        wg_hmma(c[bb], al, bh);
        wg_hmma(c[bb], ah, bl);
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb)
#pragma unroll
    for (int q = 0; q < 4; ++q) red[warp][4 * bb + q][lane] = c[bb][q];
  __syncthreads();
  for (int idx = tid; idx < 512; idx += 256) {
    const int slot = idx >> 5, ln = idx & 31;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][slot][ln];
    const int bb = slot >> 2, q = slot & 3;
    const int j = 16 * jblk + (ln >> 2) + 8 * (q >> 1);
    const int i = 8 * (4 * ig + bb) + 2 * (ln & 3) + (q & 1);
    if (bb < nb && j < J && i < I) outW[(size_t)j * I + i] = s;
  }
  if (do_bias) {  // bias gradient of this block's 16 rows of J: per-thread column sums -> fixed-order sum over warps and row lanes
    __syncthreads();
    red[warp][0][lane] = ba;  // column j0 = 16 jblk + g      (lane = 4 g + t)
    red[warp][1][lane] = bc;  // column j0 + 8
    __syncthreads();
    if (tid < 16 && 16 * jblk + tid < J) {
      const int gg = tid & 7, half = tid >> 3;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) a += red[w][half][4 * gg + tt];
      outb[16 * jblk + tid] = a;
    }
  }
#endif
}

// chunked mode: out = sum over the row chunks, in chunk order (deterministic)
__global__ void __launch_bounds__(256) wgrad_chunk_reduce_kernel(const __grid_constant__ NodeWgradParams p) {
#if GCP_DEVICE_CODE
  const int idx = (int)(blockIdx.x * 256 + threadIdx.x);
  if (idx >= p.out_total) return;
  int jn = 0;
  while (jn + 1 < p.njobs && idx >= p.job[jn + 1].out0) ++jn;
  const NodeWgradJob& jb = p.job[jn];
  const int local = idx - jb.out0, nw = jb.J * jb.I;
  float s = 0.f;
  for (int c = 0; c < p.nchunks; ++c) s += __ldg(p.scratch + (size_t)c * p.out_total + idx);
  if (local < nw) jb.outW[local] = s;
  else if (jb.outb != nullptr) jb.outb[local - nw] = s;
#endif
}

}  // namespace gcp
