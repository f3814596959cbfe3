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
//   * message GCP 0 sees [h_row | e | h_col] and [chi_row | xi | chi_col]: the node-feature blocks are linear maps of
//     per-NODE data, evaluated once per node by a small pre-kernel (P, Q below) and added in the epilogues; at edge
//     level GCP 0 is an ordinary GCP over [e | n | q] and xi.
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
constexpr int MAX_RSEQ = 64;

struct TcChunk { int off, floats; };  // piece of the packed blob (floats; off multiple of 4)
struct TcGcp {
  int si, vi, so, vo, hd, act_s, vres;
  int sop;                            // so -> 16; the scalar batch has N = sop + 16
  int zc0, nslot, kz;                 // Z tile: tail [n (nslot) | q (9) | 0 0 0] starts at zc0 (se for GCP 0, si otherwise); kz = GEMM K
  int vkc;                            // K of the vector batch: channels of the V tile read (xi only for GCP 0)
  // offsets (floats) inside the small chunk
  int o_wv_hi, o_wv_lo;               // forward  B tile [32][vkc]   (slab pitch 32)
  int o_wvt_hi, o_wvt_lo;             // backward B tile [vkc][32]   (slab pitch vkc): data gradient of the vector batch
  int o_b;                            // composed bias [sop + 16]
};
struct RingDesc { int n, nslot, slot_floats, pad_; TcChunk c[MAX_RSEQ]; };

// Per-node pre-products of message GCP 0: the h_row / h_col column blocks of scalar_out and
// the chi_row / chi_col channel blocks of the vector batch do not depend on the edge, so they are evaluated once per
// NODE and gathered:  P[i] = [src: T-part (sop) g-part (16) | dst: ...],  Q[i] = [src: 3 planes x 32 | dst: 3 planes x 32].
struct TcNodeTiles { int ps, pd, qs, qd; };  // blob offsets of the B tiles [sop+16][s], [sop+16][s], [32][v8], [32][v8]

struct TcEdgeParams {
  int N, E, L;
  int rows;                           // edges per tile: 128, or fewer (multiple of 8) so that small graphs still give every SM a tile
  int s, v, se, ve;
  int residual, e3;
  float slope;
  const float *h, *chi, *e, *xi, *frames;
  const int *perm, *src, *dst;
  const float* blob;
  const float *P, *Q;                 // [N][2 * (sop + 16)], [N][2 * 96]
  int pw;                             // sop + 16
  float* agg;                         // [N][s + 3v] per-destination sums + [tiles][2][s + 3v] carries (segment_total)
  const int* dst_ptr;                 // [N + 1] CSR row pointer of the destination-sorted order
  float* saved;                       // per tile: (L-1) x [S image | V image]; nullptr = inference
  long long* dbg;                     // optional [L][16] clock64 stamps of CTA 0's first tile (development aid)
  long long saved_tile_stride; int s_img, v_img;  // floats
  int ZBUF, VBUF, FBUF, RING_S, RING_W, BARS, smem_floats;   // shared-memory map (floats)
  int ZLO, VLO, VACC, TACC, tmem_cols;                       // TMEM column map
  TcNodeTiles nt;
  TcGcp g[12];                        // GCPNET_MAX_MESSAGE_LAYERS
  RingDesc ring_s, ring_w;
};

// ---- backward -------------------------------------------------------------------------------------------------------
// Gradients are taken with respect to the COMPOSED matrices the forward uses (chain rule back to the reference's
// parameters happens once per layer in tc_finalize_kernel):
//   G_tg[k]  [pw][kz]   = sum_e [gT | gg][e] (x) Z[e]        gT = gS' . act'(T), gg = gate pre-activation cotangent
//                         (the Z tile's last padding column is set to 1, so column kz-1 of G_tg is the bias gradient)
//   G_v[k]   [32][16]   = sum_{e,xyz} [gH | gD | gU][e] (x) V_in[e]
// Data gradients: gZ = [gT | gg] . W_tg (one GEMM, N = kz), gV_in = [gH | gD | gU] . W_v (N = channels).
struct TcBwdParams {
  TcEdgeParams f;                     // forward view (recompute): tiles, rings are the BACKWARD ones
  const float* gagg;                  // [N][s + 3v] cotangent of the aggregated messages
  const int* dst_ptr;                 // [N+1] (mean reduce: 1 / in-degree)
  int reduce_mean;
  float *ge, *gxi;                    // [E][se], [E][3 ve]  caller's edge order
  float* Y;                           // per tile: [GTG image (pw/4 slabs) | GHDU image (3 planes x 8 slabs)] of message GCP 0
  int y_img_g, y_img_v;               // floats
  float* partial;                     // [grid][partial_stride]
  int partial_stride;
  int off_tg[12], off_v[12];          // offsets inside a partial row
  int GTG, GHDU;                      // shared-memory offsets (floats) of the cotangent tiles
  int GS, GV, GHDULO;                 // TMEM columns; GTG lo aliases f.ZLO, GZACC aliases f.TACC, GVACC aliases f.VLO
  int kzn[12];                        // N of the scalar data-gradient GEMM (kz -> 16)
  TcChunk wt_hi[12], wt_lo[12];       // (informational) transposed scalar tiles
};

struct TcPostParams {
  int N, s, v, pw, rows;
  const float* Y; int y_img_g, y_img_v;
  const int *dst_ptr, *src_ptr, *src_pos;
  const float *h, *chi, *blob;
  TcNodeTiles nt;
  float* A;                 // [N][2][pw + 96]
  float *g_h, *g_chi;       // accumulated into
  float* npartial; int npartial_stride, nctas;
};

// Chain rule from the composed matrices back to the reference's parameters (see TcBwdParams):
//   W_tg = [Ws ; Wg Ws],  b' = [bs ; Wg bs + bg],  W_v = [Wd ; Wdf (rows 13..15) ; Wu Wd (rows 16..)]
struct TcFinalGcp {
  int si, vi, so, vo, hd, nslot, zc0, kz, off_tg, off_v;
  int grad_off[7];  // vector_down, vector_down_frames, scalar_out_w, scalar_out_b, vector_up, vector_out_scale_w, vector_out_scale_b
  const float *Wd, *Ws, *bs, *Wu, *Wg;
};
struct TcFinalParams {
  int L, s, v, se, ve, pw, n_edge_params;
  const float* G;      // reduced edge-level composed gradients (partial row layout)
  const float* Gn;     // reduced node-level composed gradients of GCP 0: [src: pw x s | dst: pw x s | src: 32 x 16 | dst: 32 x 16]
  float* out;          // flat parameter gradient (message_fusion part)
  TcFinalGcp g[12];
};

}  // namespace tc
}  // namespace gcp
