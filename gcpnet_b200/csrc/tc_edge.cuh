// tc_edge.cuh -- the fused edge-message forward on the 5th-generation tensor cores (tcgen05, 3xTF32):
// parameter structures and design notes (device code: tc_edge_dev.cuh, planning: tc_setup.h).
//
// Reference: GCPMessagePassing.message, src/models/components/gcpnet.py:898-936, with GCP2.forward
// (:393-468) for every message GCP.  Same math as edge_kernels.cuh; different machine mapping:
//
//   * tile = 128 destination-sorted edges = the M dimension of every tcgen05.mma (cta_group::1);
//     thread (row r, part p): r = 32 * (warp % 4) + lane is the edge AND the TMEM lane the thread
//     can reach with tcgen05.ld/st; the CS = NT/128 threads of a row split its columns.
//   * per GCP TWO tensor-core batches, fp32 accumulation in TMEM, 3xTF32 (hi*hi + lo*hi + hi*lo):
//       vector batch   3 planes (x,y,z):  [128 x vi] . [32 x vi]^T  ->  [ H | D | U ]
//                      rows 0..hd-1  vector_down           (hidden channels H, gcpnet.py:420)
//                      rows 13..15   vector_down_frames    (frame down-projection D, :426)
//                      rows 16..31   vector_up . vector_down   (U = H Wu^T, :364, composed by the pack kernel)
//       scalar batch   [128 x K] . [(so+16) x K]^T  ->  [ T | g ],   K = si + hd4 + 12 columns [ s | n | q(9) | 0 0 0 ]
//                      rows 0..so-1  scalar_out            (T, :441)
//                      rows so..     vector_out_scale . scalar_out   (gate pre-activation g, :386; the vector
//                                    nonlinearity is the identity in every shipped config, so the gate is linear in
//                                    the scalar_out INPUT and the pack kernel composes the two matrices)
//     The hi operand of A is the fp32 tile in shared memory ("slab" layout of umma.cuh, no swizzle), its lo
//     part lives in TMEM (tcgen05.st by the row's thread, consumed by the TS form of the MMA), the lo part
//     of the weights comes pre-split from the pack kernel.
//   * CUDA cores only do the row-local glue between the batches: safe-norm, frame scalarisation (epilogue A);
//     bias, activation, residual, sigmoid gate (epilogue B) -- each thread on its own row, out of TMEM.
//   * weights stream through two shared-memory rings filled by cp.async.bulk (small per-GCP tiles; scalar
//     batch tiles); saved activations (the inputs S, V of every GCP = the shared-memory images) leave
//     through cp.async.bulk stores.
//   * message GCP 0 sees [h_row | e | h_col]: its scalar batch runs as three K-segments ([e | n | q], h_row,
//     h_col) and its vector batch as three channel segments (chi_row, xi, chi_col) into the same accumulators.
#pragma once

namespace gcp {
namespace tc {

constexpr int TE = 128;          // edges per tile = UMMA M
constexpr int RP = 128;          // slab pitch in rows: every 8-row x 16-byte core matrix is one aligned 128-byte line
constexpr int SLAB = RP * 4;     // floats per 4-column slab
constexpr int PW = 16;           // columns per (x,y,z) plane of the vector tiles (channels padded to 16)
constexpr int PLANE = (PW / 4) * SLAB;  // floats per plane
constexpr int DCOL = 13;         // vector batch accumulator: columns [0,hd) hidden channels, [13,16) frame-down vectors,
constexpr int UCOL = 16;         //                           columns [16,32) ungated vector outputs
constexpr int VN = 32;           // N of the vector batch
constexpr int NSLOT = 12;        // Z-tile tail: hd -> 4 norm slots (at most 12), then 9 frame scalars, then 3 zero columns
constexpr int MAX_SEG = 4;
constexpr int MAX_RSEQ = 64;

struct TcChunk { int off, floats; };  // piece of the packed blob (floats; off multiple of 4)
struct TcSeg {                        // one K-segment of a scalar batch = one commit
  int a_tile;                         // 0: Z tile ([e | n | q] for GCP 0, [S | n | q] otherwise), 1: X tile = h_row, 2: X tile = h_col
  int kc;                             // columns (multiple of 8), starting at column 0 of the tile
};
struct TcGcp {
  int si, vi, so, vo, hd, act_s, vres;
  int sop;                            // so -> 16; the scalar batch has N = sop + 16
  int zc0, nslot;                     // first Z-tile column of the tail [n (nslot) | q (9) | 0 0 0]  (se for GCP 0, si otherwise)
  int nseg; TcSeg seg[MAX_SEG];
  int nvseg, vkc[3];                  // vector batch channel segments (multiples of 8)
  int o_wd_hi[3], o_wd_lo[3];         // offsets (floats) inside the small chunk: B tiles [32][vkc] (slab pitch 32)
  int o_bs, o_bg;                     // scalar_out bias [sop]; composed gate bias [16]
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
  long long* dbg;                     // optional [L][16] clock64 stamps of CTA 0's first tile (development aid)
  long long saved_tile_stride; int s_img, v_img;  // floats
  int ZBUF, XBUF, VBUF, FBUF, RING_S, RING_W, BARS, smem_floats;   // shared-memory map (floats)
  int ZLO, XLO, VLO, VACC, TACC, tmem_cols;                        // TMEM column map
  TcGcp g[12];                        // GCPNET_MAX_MESSAGE_LAYERS
  RingDesc ring_s, ring_w;
};



}  // namespace tc
}  // namespace gcp
