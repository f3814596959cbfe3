// tc_edge.cuh -- the fused edge-message forward on the 5th-generation tensor cores (tcgen05, 3xTF32).
//
// Reference: GCPMessagePassing.message, src/models/components/gcpnet.py:898-936, with GCP2.forward
// (:393-468) for every message GCP.  Same math as edge_kernels.cuh; different machine mapping:
//
//   * tile = 128 destination-sorted edges = the M dimension of every tcgen05.mma (cta_group::1);
//     thread (row r, part p): r = 32 * (warp % 4) + lane is the edge AND the TMEM lane the thread
//     can reach with tcgen05.ld/st; the CS = NT/128 threads of a row split its columns.
//   * every Linear of the GCP is a tensor-core GEMM with fp32 accumulation in TMEM:
//       vector_down (+ frame down-projection)   3 planes (x,y,z)  [128 x vi] . [16 x vi]^T
//       scalar_out                               [128 x K] . [so x K]^T     K = si + hd + 9 (padded)
//       vector_up                                3 planes          [128 x hd] . [vo x hd]^T
//       vector_out_scale (gate)                  [128 x so] . [vo x so]^T
//     evaluated as 3xTF32 (hi*hi + lo*hi + hi*lo): the hi operand is the fp32 tile in shared memory
//     ("slab" layout of umma.cuh, no swizzle), the lo part of A lives in TMEM (tcgen05.st by the row's
//     thread, consumed by the TS form of the MMA), the lo part of the weights comes pre-split from
//     the pack kernel.
//   * CUDA cores only do the row-local glue between GEMMs: safe-norm, frame scalarisation,
//     activation, residual, sigmoid gate -- each thread on its own row, straight out of TMEM.
//   * weights stream through two shared-memory rings filled by cp.async.bulk (one for the small
//     per-GCP matrices, one for the scalar_out tiles); saved activations (the inputs S, V of every
//     GCP, exactly the shared-memory images) leave through cp.async.bulk stores.
//   * message GCP 0 sees [h_row | e | h_col] (K = 2s + se + hd + 9): its scalar_out runs as K-segments
//     ([e | n | q], [h_row], [h_col]) over the two staging tiles, and its vector_down as three
//     channel segments (chi_row, xi, chi_col) accumulated into one TMEM tile.
#pragma once
#include "gcp_tile.cuh"
#include "umma.cuh"

namespace gcp {
namespace tc {

using namespace ::gcp::umma;

constexpr int TE = 128;          // edges per tile = UMMA M
constexpr int RP = 129;          // slab pitch in rows (129 * 16 B = 16 mod 128: column-group strides hit distinct banks)
constexpr int SLAB = RP * 4;     // floats per 4-column slab
constexpr int PW = 16;           // columns per (x,y,z) plane of the vector tiles (channels padded to 16)
constexpr int PLANE = (PW / 4) * SLAB;  // floats per plane
constexpr int DCOL = 13;         // vector_down accumulator: columns [0,hd) hidden channels, [13,16) frame-down vectors
constexpr int MAX_SEG = 4;
constexpr int MAX_RSEQ = 64;

struct TcChunk { int off, floats; };  // piece of the packed blob (floats; off multiple of 4)
struct TcSeg {                        // one K-segment of a scalar_out GEMM = one MMA batch
  int a_tile;                         // 0: Z tile ([e | n | q] for GCP 0, [S | n | q] otherwise), 1: T tile = h_row, 2: T tile = h_col
  int kc;                             // columns (multiple of 8), starting at column 0 of the tile
};
struct TcGcp {
  int si, vi, so, vo, hd, act_s, act_v, vres;
  int hdp, sop, vop;                  // hd -> 8, so -> 16, vo -> 16
  int zc0;                            // first Z-tile column of [n | q]  (se for GCP 0, si otherwise)
  int kz;                             // Z-tile columns read by the scalar_out GEMM: zc0 + hd + 9 -> 8
  int nseg; TcSeg seg[MAX_SEG];
  int nvseg, vkc[3];                  // vector_down channel segments (multiples of 8)
  // offsets (floats) inside the small chunk
  int o_wd_hi[3], o_wd_lo[3];         // vector_down B tiles [16][vkc] (slab pitch 16)
  int o_wu_hi, o_wu_lo;               // [vop][hdp]
  int o_wg_hi, o_wg_lo, gk;           // [vop][gk], gk = so -> 8
  int o_bs, o_bg;
};
struct RingDesc { int n, nslot, slot_floats, pad_; TcChunk c[MAX_RSEQ]; };

struct TcEdgeParams {
  int N, E, L;
  int s, v, se, ve;
  int residual, e3;
  float slope;
  const float *h, *chi, *e, *xi, *frames;
  const int *perm, *src, *dst;
  const float* blob;
  float* msg;                         // [E][s + 3v]
  float* saved;                       // per tile: (L-1) x [S image | V image]; nullptr = inference
  long long saved_tile_stride; int s_img, v_img;  // floats
  int ZBUF, TBUF, VBUF, HBUF, FBUF, RING_S, RING_W, BARS, smem_floats;   // shared-memory map (floats)
  int ZLO, TLO, VLO, HLO, HDACC, TACC, GACC, UACC, tmem_cols;            // TMEM column map
  TcGcp g[MAX_MSG_LAYERS];
  RingDesc ring_s, ring_w;
};



}  // namespace tc
}  // namespace gcp
