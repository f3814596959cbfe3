// gcp_tile.cuh -- CTA-tile building blocks of the GCP2 perceptron (forward and backward), v2.
//
// Everything here works on a TILE of TE "entities" (edges in the message kernel, nodes in the
// node-update kernel) whose features sit in shared memory, entity-major: X[e][f], row stride ld.
// Reference semantics: GCP2.forward, src/models/components/gcpnet.py:393-468 (+ :353-391 for the
// vector gate), scalarize src/models/components/__init__.py:272-325, safe_norm :381-392.
//
// Design (B200):
//   * NT = 256..512 threads per CTA (8 threads per edge), one CTA per SM, persistent over tiles.
//   * Weights never go through registers on their way in: the host-side pack kernel (pack.cuh) lays
//     every GCP out as a few "chunks" in exactly the shared-memory layout the tile code wants, and the
//     CTA streams them through a small ring of shared-memory slots with cp.async.bulk (TMA 1-D bulk
//     copy, SASS UBLKCP) completing on mbarriers -- the next chunks land while the current GEMM runs.
//   * Dense Linear layers are register-tiled fp32 FFMA GEMMs: a warp owns a 16 x 16 output tile
//     (lane = 8 row groups x 4 column groups, 2 x 4 outputs per thread), operands are read with
//     conflict-free 128-bit shared loads (1 wavefront per instruction).
//   * All cooperative loops are "warp per row, lane per column" or use compile-time divisors: no
//     runtime integer division in any inner loop.
//
// Code style: a routine is a sequence of PHASES.  A phase is a parallel-for over the CTA's NT
// threads followed by a barrier; no per-thread value lives across a phase boundary (CTA-uniform
// values such as the pipeline position do).  The same source also compiles for the host, where a
// phase runs the NT thread bodies one after another (optionally in reverse order, to expose
// intra-phase hazards) -- that build is the CPU emulation used by the non-GPU tests (tests/emul);
// it is never part of the product path.
//
// Bank-conflict rules used for the strides (floats):
//   * arrays read with float4 along the feature axis by lanes that differ in the row:
//       ld % 8 == 4  -> 8 consecutive rows hit 8 disjoint 4-bank groups          (ld_vec())
//   * arrays read scalar by lanes that differ in the row: ld odd                  (ld_scal())
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

// Development aid (clock64 stage stamps, scripts/tc_bwd_stamps.py): compiled out unless built with -DGCP_STAMPS=1.
#ifndef GCP_STAMPS
#define GCP_STAMPS 0
#endif

namespace gcp {

#define GCP_HD __host__ __device__ __forceinline__
#define GCP_HDN __host__ __device__
#define GCP_HDN_NOINLINE __host__ __device__ __noinline__

inline int g_emul_reverse = 0;  // host emulation only: run thread bodies in reverse order when set
inline bool g_emul_first = false;  // host emulation only: true while the first-executed thread body of a phase runs
#if defined(__CUDA_ARCH__)
#define GCP_DEVICE_CODE 1
#define GCP_PHASE_BEGIN(NT) { const int tid = (int)threadIdx.x; (void)tid;
#define GCP_PHASE_END } __syncthreads();
#define GCP_LDG(p) __ldg(p)
#else
#define GCP_DEVICE_CODE 0
#define GCP_PHASE_BEGIN(NT) for (int tid_ = 0; tid_ < (NT); ++tid_) { const int tid = ::gcp::g_emul_reverse ? (NT) - 1 - tid_ : tid_; (void)tid; ::gcp::g_emul_first = (tid_ == 0);
#define GCP_PHASE_END }
#define GCP_LDG(p) (*(p))
#endif

constexpr int MAX_MSG_LAYERS = 12;
constexpr float SAFE_NORM_EPS = 1e-8f;  // comp/__init__.py:385

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKYRELU = 2, ACT_SILU = 3, ACT_SIGMOID = 4, ACT_SELU = 5 };

GCP_HD float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// src/models/__init__.py:41-57.  The shipped configs use relu / identity: those two stay inline, the transcendental
// ones live behind a call (inlining them at every call site of the unrolled tile loops bloats the kernels past the
// instruction cache: 30 % of the node-update kernel's SASS was activation bodies).
__host__ __device__ __noinline__ float act_fwd_slow(int a, float x, float slope);
__host__ __device__ __noinline__ float act_grad_slow(int a, float x, float slope);
GCP_HD float act_fwd(int a, float x, float slope) {
  if (a == ACT_NONE) return x;
  if (a == ACT_RELU) return x > 0.f ? x : 0.f;
  return act_fwd_slow(a, x, slope);
}
GCP_HD float act_grad(int a, float x, float slope) {
  if (a == ACT_NONE) return 1.f;
  if (a == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  return act_grad_slow(a, x, slope);
}
__host__ __device__ __noinline__ inline float act_fwd_slow(int a, float x, float slope) {
  switch (a) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_LEAKYRELU: return x > 0.f ? x : slope * x;
    case ACT_SILU: return x * sigmoidf_(x);
    case ACT_SIGMOID: return sigmoidf_(x);
    case ACT_SELU: {
      const float al = 1.6732632423543772848170429916717f, sc = 1.0507009873554804934193349852946f;
      return sc * (x > 0.f ? x : al * (expf(x) - 1.f));
    }
    default: return x;
  }
}
__host__ __device__ __noinline__ inline float act_grad_slow(int a, float x, float slope) {
  switch (a) {
    case ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case ACT_LEAKYRELU: return x > 0.f ? 1.f : slope;
    case ACT_SILU: { const float s = sigmoidf_(x); return s * (1.f + x * (1.f - s)); }
    case ACT_SIGMOID: { const float s = sigmoidf_(x); return s * (1.f - s); }
    case ACT_SELU: {
      const float al = 1.6732632423543772848170429916717f, sc = 1.0507009873554804934193349852946f;
      return sc * (x > 0.f ? 1.f : al * expf(x));
    }
    default: return 1.f;
  }
}

GCP_HD int round_up(int x, int m) { return (x + m - 1) / m * m; }
GCP_HD int ld_vec(int cols) { const int c4 = round_up(cols, 4); return (c4 % 8 == 4) ? c4 : c4 + 4; }
GCP_HD int ld_scal(int cols) { return cols | 1; }
GCP_HD int hd_cols(int hd) { return round_up(hd, 4) + 4; }          // H columns (padded) + 3 frame-down columns (+1 pad)
GCP_HD int ld_hd(int hd) { return ld_vec(3 * hd_cols(hd)); }

GCP_HD float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
GCP_HD void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------
// packed weights (written by pack.cuh) and the shared-memory weight ring
// ------------------------------------------------------------------------------------------
struct WChunk { int off, floats; };  // offset (floats) into the layer's packed blob; size multiple of 4
constexpr int MAX_WS_CHUNKS = 16;
constexpr int MAX_WSEQ = 96;
constexpr int MAX_WSLOTS = 3;

// Packed layout of one GCP2 (all row strides padded, all padding zero):
//   S  chunk: WdT[vi][cols]   WdT[c][k] = vector_down[k][c] (k < hd), WdT[c][hdp + cc] = vector_down_frames[cc][c]
//   WS chunk c (c < nWS): W[NP][ldk], W[n][kk] = scalar_out.weight[n][c*kc + kk], followed by scalar_out.bias
//             (NP floats; every chunk carries it so that all chunks have one size).  NP = round_up(so, 16):
//             padding rows are zero.  Chunk c starts at ws_off + c * ws_stride.
//   G  chunk: WG[round_up(vo,4)][ldg] = vector_out_scale.weight (zero rows beyond vo), then bias (round_up(vo,4)), then WU[vo][hdp] = vector_up.weight
struct GcpW {
  WChunk S, G;
  int ws_off, ws_floats, ws_stride;
  int nWS, kc, ldk, NP;
  int ldg, o_bg, o_wu;  // offsets inside the G chunk
  int cols, hdp;
};

// One GCP2 module (device view).  Raw weight pointers keep the nn.Linear layouts of the reference
// (gcpnet.py:303-322) and are only read by the pack kernel; the tile code reads the packed chunks.
// o_* = offsets (floats) of the matching gradient blocks inside a per-CTA partial-gradient row.
struct GcpOp {
  int si, vi, so, vo, hd;
  int act_s, act_v, vres;
  const float *Wd, *Wdf, *Ws, *bs, *Wu, *Wg, *bg;
  int o_Wd, o_Wdf, o_Ws, o_bs, o_Wu, o_Wg, o_bg;
  int flags;  // GCP2_NO_FRAMES | GCP2_NO_GATE (GCP-Baseline variants, gcpnet.py:302-322,344-350,424-437)
  GcpW w;
};
constexpr int GCP2_NO_FRAMES = 1, GCP2_NO_GATE = 2;
GCP_HD int gcp_nfs(const GcpOp& op) { return (op.flags & GCP2_NO_FRAMES) ? 0 : 9; }  // frame scalars read by scalar_out
GCP_HD bool gcp_gated(const GcpOp& op) { return (op.flags & GCP2_NO_GATE) == 0; }
GCP_HD int gcp_k(const GcpOp& op) { return op.si + op.hd + gcp_nfs(op); }
GCP_HD int gcp_kpad(const GcpOp& op) { return op.w.nWS * op.w.kc; }  // columns the GEMM reads from Z

struct WSeq {  // the order in which one kernel consumes chunks, per tile
  int n, nslot, slot_floats, pad_;
  WChunk c[MAX_WSEQ];
};

struct WPipe {  // CTA-uniform state of the ring
  float* slots;
  unsigned long long* mbar;
  const float* blob;
  const WSeq* seq;
  int head;   // position of the chunk the next wpipe_wait() returns
  int total;  // positions this CTA consumes over its whole life
  long long* dbg = nullptr;  // development aid: clock64 stamps of the backward phases (thread 0 of CTA 0), 16 per call
};

#if GCP_DEVICE_CODE
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#endif

// one thread: start the bulk copy of position `pos` into its slot
GCP_HD void wpipe_issue(const WPipe& w, int pos) {
  const WChunk ck = w.seq->c[pos % w.seq->n];
  const int slot = pos % w.seq->nslot;
  float* dst = w.slots + (size_t)slot * w.seq->slot_floats;
  const float* src = w.blob + ck.off;
#if GCP_DEVICE_CODE
  const uint32_t bar = smem_u32(&w.mbar[slot]);
  const uint32_t bytes = (uint32_t)ck.floats * 4u;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
#else
  memcpy(dst, src, (size_t)ck.floats * 4);
#endif
}
// every thread, before reading the head chunk
GCP_HD const float* wpipe_wait(const WPipe& w) {
  const int slot = w.head % w.seq->nslot;
#if GCP_DEVICE_CODE
  const uint32_t bar = smem_u32(&w.mbar[slot]);
  const uint32_t parity = (uint32_t)((w.head / w.seq->nslot) & 1);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
#endif
  return w.slots + (size_t)slot * w.seq->slot_floats;
}
// thread 0, at the START of the first phase after the barrier that ended the last read of the head
// chunk (the caller then advances w.head in CTA-uniform code): refill the slot just freed
GCP_HD void wpipe_refill(const WPipe& w, int released_pos, int tid) {
#if GCP_DEVICE_CODE
  const bool issuer = tid == 0;
#else
  // the emulated copy is instantaneous, so it has to happen before any thread body of this phase reads the
  // slot (on the device the readers block on the mbarrier instead): the first-executed body issues it
  const bool issuer = g_emul_first; (void)tid;
#endif
  if (issuer && released_pos + w.seq->nslot < w.total) wpipe_issue(w, released_pos + w.seq->nslot);
}
// kernel prologue (CTA-uniform): barrier init + first nslot copies.  NT-thread phase.
template <int NT>
GCP_HDN void wpipe_start(WPipe& w) {
  GCP_PHASE_BEGIN(NT)
  if (tid == 0) {
#if GCP_DEVICE_CODE
    for (int i = 0; i < w.seq->nslot; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&w.mbar[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    for (int i = 0; i < w.seq->nslot && i < w.total; ++i) wpipe_issue(w, i);
  }
  GCP_PHASE_END
  w.head = 0;
}

// Node entities with enable_e3_equivariance: scalarize takes |.| of the x_cross projections PER EDGE before averaging over
// the node's outgoing edges (comp/__init__.py:305-323), which the mean frame cannot express -> those three frame scalars
// walk the node's outgoing edges (source CSR) instead.
struct NodeE3 {
  const int *src_ptr, *src_pos, *perm;
  const float* frames;   // [E][9], caller's edge order
  int row0, nrows;       // node range of the tile
};

// Shared-memory views used by one GCP2 evaluation on a tile.
struct TileBufs {
  float* Z;   int ldz;   // [TE][ldz]  cols [0,si)=scalars in, [si,si+hd)=norms, [si+hd,si+hd+9)=frame scalars; pad cols zero
  float* V;   int ldv;   // [TE][ldv]  3*vi floats, (channel, xyz) xyz fastest, scalar stride
  float* HD;  int ldhd;  // [TE][ldhd] 3 x hd_cols: H[x][0..hd), frame-down D[x][0..3) at column hd_cols-4
  float* F;              // [TE][9]    frames (a, xyz)
  float* T;   int ldt;   // [TE][ldt]  pre-activation scalar_out
  float* SG;  int ldsg;  // [TE][ldsg] sigmoid gate per output vector channel
  float* WSM;            // persistent copy of the current GCP's small weights (backward): WdT | WU
  const NodeE3* e3n = nullptr;  // node entities with e3 frames (see NodeE3); nullptr: F holds everything
};
constexpr int LDF = 9;

#if GCP_DEVICE_CODE
// Register-fragment tensor-core MMA (m16n8k8, tf32 inputs, fp32 accumulation); used as 3xTF32 (hi*hi + lo*hi + hi*lo).
__device__ __forceinline__ void wg_hmma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// ------------------------------------------------------------------------------------------
// GEMM thread map: warp = 16 rows x 16 columns of outputs, lane = (eg = lane>>2, og = lane&3),
// thread rows {e0, e0+8}, 4 columns per 16-wide slice.
// ------------------------------------------------------------------------------------------
template <int TE, int NT>
struct GemmMap {
  static constexpr int NW = NT / 32, WE = TE / 16, WN = NW / WE;
  static_assert(NT % 32 == 0 && TE % 16 == 0 && WE >= 1 && WN >= 1 && WE * WN == NW, "bad GEMM thread grid");
  int e0, og, wn;
  GCP_HD GemmMap(int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    e0 = 16 * (warp % WE) + (lane >> 2);
    og = lane & 3;
    wn = warp / WE;
  }
  // Column (inside the 16-wide slice) of accumulator element j of this thread.  Node tiles (16 rows) run their GEMMs on
  // the register-fragment tensor-core MMA on the device (column pairs {2 og, 2 og + 1} of the two 8-column halves); the
  // FFMA loops (edge tiles; host emulation) own columns og + 4 j (n-major weights) resp. 4 og + j (k-major weights).
#if GCP_DEVICE_CODE
  static constexpr bool kMma = TE <= 16;
#else
  static constexpr bool kMma = false;
#endif
  GCP_HD int col_n(int j) const { return kMma ? 2 * og + (j & 1) + 8 * (j >> 1) : og + 4 * j; }
  GCP_HD int col_k(int j) const { return kMma ? 2 * og + (j & 1) + 8 * (j >> 1) : 4 * og + j; }
};

// acc[i][r][j] += sum_{kk<kc} xmap(X[e_r][k0+kk]) * Wc[n][kk],  n = 16*(wn + WN*i) + og + 4*j  (n-major weights)
template <int TE, int NT, int SL, class XMap>
GCP_HD void gemm_nmajor_chunk(float (&acc)[SL][2][4], const float* X, int ldx, int k0, int kc, const float* Wc, int ldk,
                              int nslices, const GemmMap<TE, NT>& m, XMap xmap) {
  constexpr int WN = GemmMap<TE, NT>::WN;
  const float* x0 = X + m.e0 * ldx + k0;
  const float* x1 = x0 + 8 * ldx;
#if GCP_DEVICE_CODE
  if (GemmMap<TE, NT>::kMma) {  // 3xTF32 on mma.sync m16n8k8: A = X rows {e0, e0 + 8}, B = n-major weights
    const int g = m.e0 & 7, t = m.og;
    for (int k8 = 0; k8 < kc; k8 += 8) {  // kc is a multiple of 4: the last step may be half empty
      const bool tail = k8 + 4 >= kc;
      const float fa[4] = {xmap(x0[k8 + t]), xmap(x1[k8 + t]), tail ? 0.f : xmap(x0[k8 + t + 4]), tail ? 0.f : xmap(x1[k8 + t + 4])};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { ah[q] = __float_as_uint(fa[q]) & 0xffffe000u; al[q] = __float_as_uint(fa[q] - __uint_as_float(ah[q])); }
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        const int sl = m.wn + WN * i;
        if (sl < nslices) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float* wp = Wc + (16 * sl + 8 * hh + g) * ldk + k8 + t;
            const float fb0 = wp[0], fb1 = tail ? 0.f : wp[4];
            uint32_t bh[2] = {__float_as_uint(fb0) & 0xffffe000u, __float_as_uint(fb1) & 0xffffe000u};
            uint32_t bl[2] = {__float_as_uint(fb0 - __uint_as_float(bh[0])), __float_as_uint(fb1 - __uint_as_float(bh[1]))};
            float c[4] = {acc[i][0][2 * hh], acc[i][0][2 * hh + 1], acc[i][1][2 * hh], acc[i][1][2 * hh + 1]};
            wg_hmma(c, ah, bh);
            wg_hmma(c, al, bh);
            wg_hmma(c, ah, bl);
            acc[i][0][2 * hh] = c[0]; acc[i][0][2 * hh + 1] = c[1]; acc[i][1][2 * hh] = c[2]; acc[i][1][2 * hh + 1] = c[3];
          }
        }
      }
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    const int sl = m.wn + WN * i;
    if (sl < nslices) {
      const float* w0 = Wc + (16 * sl + m.og) * ldk;
#pragma unroll 2
      for (int k4 = 0; k4 < kc; k4 += 4) {
        float4 a = ld4(x0 + k4), b = ld4(x1 + k4);
        a.x = xmap(a.x); a.y = xmap(a.y); a.z = xmap(a.z); a.w = xmap(a.w);
        b.x = xmap(b.x); b.y = xmap(b.y); b.z = xmap(b.z); b.w = xmap(b.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wv = ld4(w0 + 4 * j * ldk + k4);
          acc[i][0][j] = fmaf(a.x, wv.x, acc[i][0][j]); acc[i][0][j] = fmaf(a.y, wv.y, acc[i][0][j]);
          acc[i][0][j] = fmaf(a.z, wv.z, acc[i][0][j]); acc[i][0][j] = fmaf(a.w, wv.w, acc[i][0][j]);
          acc[i][1][j] = fmaf(b.x, wv.x, acc[i][1][j]); acc[i][1][j] = fmaf(b.y, wv.y, acc[i][1][j]);
          acc[i][1][j] = fmaf(b.z, wv.z, acc[i][1][j]); acc[i][1][j] = fmaf(b.w, wv.w, acc[i][1][j]);
        }
      }
    }
  }
}

// acc[i][r][c] = sum_{k<K} X[e_r][k] * Wc[k][16*(wn+WN*i) + 4*og + c]   (k-major weights: the data
// gradient through an nn.Linear whose packed weight is [K = out rows][ldw]).  K multiple of 4.
template <int TE, int NT, int SL>
GCP_HD void gemm_kmajor_chunk(float (&acc)[SL][2][4], const float* X, int ldx, int K, const float* Wc, int ldw,
                              int nslices, const GemmMap<TE, NT>& m) {
  constexpr int WN = GemmMap<TE, NT>::WN;
  const float* x0 = X + m.e0 * ldx;
  const float* x1 = x0 + 8 * ldx;
#if GCP_DEVICE_CODE
  if (GemmMap<TE, NT>::kMma) {  // 3xTF32 on mma.sync m16n8k8: B = k-major weights Wc[k][n]
    const int g = m.e0 & 7, t = m.og;
#pragma unroll
    for (int i = 0; i < SL; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[i][0][c] = 0.f; acc[i][1][c] = 0.f; }
    for (int k8 = 0; k8 < K; k8 += 8) {  // K is a multiple of 4; the last half step reads zero-padded rows / columns
      const bool tail = k8 + 4 >= K;
      const float fa[4] = {x0[k8 + t], x1[k8 + t], tail ? 0.f : x0[k8 + t + 4], tail ? 0.f : x1[k8 + t + 4]};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { ah[q] = __float_as_uint(fa[q]) & 0xffffe000u; al[q] = __float_as_uint(fa[q] - __uint_as_float(ah[q])); }
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        const int sl = m.wn + WN * i;
        if (sl < nslices) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float* wp = Wc + (k8 + t) * ldw + 16 * sl + 8 * hh + g;
            const float fb0 = wp[0], fb1 = tail ? 0.f : wp[4 * ldw];
            uint32_t bh[2] = {__float_as_uint(fb0) & 0xffffe000u, __float_as_uint(fb1) & 0xffffe000u};
            uint32_t bl[2] = {__float_as_uint(fb0 - __uint_as_float(bh[0])), __float_as_uint(fb1 - __uint_as_float(bh[1]))};
            float c[4] = {acc[i][0][2 * hh], acc[i][0][2 * hh + 1], acc[i][1][2 * hh], acc[i][1][2 * hh + 1]};
            wg_hmma(c, ah, bh);
            wg_hmma(c, al, bh);
            wg_hmma(c, ah, bl);
            acc[i][0][2 * hh] = c[0]; acc[i][0][2 * hh + 1] = c[1]; acc[i][1][2 * hh] = c[2]; acc[i][1][2 * hh + 1] = c[3];
          }
        }
      }
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    const int sl = m.wn + WN * i;
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[i][0][c] = 0.f; acc[i][1][c] = 0.f; }
    if (sl < nslices) {
      const float* w0 = Wc + 16 * sl + 4 * m.og;
#pragma unroll 2
      for (int k4 = 0; k4 < K; k4 += 4) {
        const float4 a = ld4(x0 + k4), b = ld4(x1 + k4);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 wv = ld4(w0 + (k4 + kk) * ldw);
          acc[i][0][0] = fmaf(av[kk], wv.x, acc[i][0][0]); acc[i][0][1] = fmaf(av[kk], wv.y, acc[i][0][1]);
          acc[i][0][2] = fmaf(av[kk], wv.z, acc[i][0][2]); acc[i][0][3] = fmaf(av[kk], wv.w, acc[i][0][3]);
          acc[i][1][0] = fmaf(bv[kk], wv.x, acc[i][1][0]); acc[i][1][1] = fmaf(bv[kk], wv.y, acc[i][1][1]);
          acc[i][1][2] = fmaf(bv[kk], wv.z, acc[i][1][2]); acc[i][1][3] = fmaf(bv[kk], wv.w, acc[i][1][3]);
        }
      }
    }
  }
}

struct XIdentity { GCP_HD float operator()(float x) const { return x; } };
struct XAct { int a; float slope; GCP_HD float operator()(float x) const { return act_fwd(a, x, slope); } };

// Weight gradient of an nn.Linear on a tile:  P[j][i] (+)= sum_e G[e][j] * fmap(Zin[e][i]),
// Pb[j] (+)= sum_e G[e][j].  P/Pb are this CTA's private rows in global memory (deterministic:
// every element is owned by exactly one thread; the previous partial is fetched BEFORE the
// reduction loop so its latency overlaps the math).  Rows e >= nrows of G must be zero.
// A thread owns a JW x IW block of P (JW in {1,4}, IW in {4,8}); G and Zin are read JW resp. IW columns
// at a time: J % JW == 0 or zero padding, and Zin carries readable columns up to round_up(I, IW).
#if GCP_DEVICE_CODE
// Device version: the products run on the register-fragment tensor-core MMA (mma.sync m16n8k8, tf32 inputs, fp32
// accumulation) as 3xTF32 (hi*hi + lo*hi + hi*lo, split in registers): the reduction index is the tile ROW, so the
// A / B fragments are read straight out of the row-major tiles.  One warp per 16 (j) x 8 (i) block of P.
template <int TE, int NT, int JW, int IW, class FMap>
__device__ __forceinline__ void tile_wgrad_ffma(const float* G, int ldg, int J, const float* Zin, int ldz, int I,
                                                float* P, float* Pb, bool accumulate, FMap fmap, int tid);
template <int TE, int NT, int JW, int IW, class FMap>
__device__ __forceinline__ void tile_wgrad(const float* G, int ldg, int J, const float* Zin, int ldz, int I,
                                           float* P, float* Pb, bool accumulate, FMap fmap, int tid) {
  static_assert(TE % 8 == 0, "tile rows");
  // The legacy tensor path only pays for short reductions (node tiles, 16 rows: the FFMA version is bound by its
  // per-output overhead there); on the wide edge tiles the FFMA blocks win (mma.sync tf32 runs at ~2x FFMA rate, 3xTF32).
  if (TE > 16) { tile_wgrad_ffma<TE, NT, JW, IW>(G, ldg, J, Zin, ldz, I, P, Pb, accumulate, fmap, tid); return; }
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int JB = (J + 15) >> 4, IB = (I + 7) >> 3;
  for (int blk = warp; blk < JB * IB; blk += NT / 32) {
    const int ib = blk / JB, jb = blk - ib * JB;
    const int j0 = 16 * jb + g, j1 = j0 + 8, i0 = 8 * ib + g;
    const int ja = j0 < J ? j0 : J - 1, jc = j1 < J ? j1 : J - 1, ii = i0 < I ? i0 : I - 1;  // clamped loads; results dropped below
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k0 = 0; k0 < TE; k0 += 8) {
      const float fa[4] = {G[(k0 + t) * ldg + ja], G[(k0 + t) * ldg + jc], G[(k0 + t + 4) * ldg + ja], G[(k0 + t + 4) * ldg + jc]};
      const float fb[2] = {fmap(Zin[(k0 + t) * ldz + ii]), fmap(Zin[(k0 + t + 4) * ldz + ii])};
      uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) { ah[q] = __float_as_uint(fa[q]) & 0xffffe000u; al[q] = __float_as_uint(fa[q] - __uint_as_float(ah[q])); }
#pragma unroll
      for (int q = 0; q < 2; ++q) { bh[q] = __float_as_uint(fb[q]) & 0xffffe000u; bl[q] = __float_as_uint(fb[q] - __uint_as_float(bh[q])); }
      wg_hmma(c, ah, bh);
      wg_hmma(c, al, bh);
      wg_hmma(c, ah, bl);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int j = 16 * jb + g + 8 * hh;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = 8 * ib + 2 * t + q;
        if (j < J && i < I) {
          float* dp = P + (size_t)j * I + i;
          *dp = (accumulate ? *dp : 0.f) + c[2 * hh + q];
        }
      }
    }
  }
  if (Pb != nullptr) {
    for (int j = tid; j < J; j += NT) {
      const float o = accumulate ? Pb[j] : 0.f;
      float s = 0.f;
      for (int e = 0; e < TE; ++e) s += G[e * ldg + j];
      Pb[j] = o + s;
    }
  }
}
#define GCP_WGRAD_NAME tile_wgrad_ffma
#else
#define GCP_WGRAD_NAME tile_wgrad
#endif
template <int TE, int NT, int JW, int IW, class FMap>
GCP_HD void GCP_WGRAD_NAME(const float* G, int ldg, int J, const float* Zin, int ldz, int I,
                       float* P, float* Pb, bool accumulate, FMap fmap, int tid) {
  static_assert((JW == 1 || JW == 4) && (IW == 4 || IW == 8), "tile shape");
  const int JT = (J + JW - 1) / JW, IT = (I + IW - 1) / IW;
  for (int tile = tid; tile < JT * IT; tile += NT) {
    const int it = tile / JT, jt = tile - it * JT;  // consecutive lanes: consecutive j (contiguous G columns)
    float acc[JW][IW], old[JW][IW];
#pragma unroll
    for (int a = 0; a < JW; ++a)
#pragma unroll
      for (int b = 0; b < IW; ++b) {
        acc[a][b] = 0.f;
        const int j = JW * jt + a, i = IW * it + b;
        old[a][b] = (accumulate && j < J && i < I) ? P[(size_t)j * I + i] : 0.f;
      }
    const float* gp = G + JW * jt;
    const float* zp = Zin + IW * it;
#pragma unroll 2
    for (int e = 0; e < TE; ++e) {
      float gv[JW], zv[IW];
      if (JW == 4) {
        const float4 g = ld4(gp + e * ldg);
        const float t4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int a = 0; a < JW; ++a) gv[a] = t4[a];
      } else {
        gv[0] = gp[e * ldg];
      }
      const float4 z0 = ld4(zp + e * ldz);
      zv[0] = fmap(z0.x); zv[1] = fmap(z0.y); zv[2] = fmap(z0.z); zv[3] = fmap(z0.w);
      if (IW == 8) {
        const float4 z1 = ld4(zp + e * ldz + 4);
        zv[IW - 4] = fmap(z1.x); zv[IW - 3] = fmap(z1.y); zv[IW - 2] = fmap(z1.z); zv[IW - 1] = fmap(z1.w);
      }
#pragma unroll
      for (int a = 0; a < JW; ++a)
#pragma unroll
        for (int b = 0; b < IW; ++b) acc[a][b] = fmaf(gv[a], zv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < JW; ++a) {
      const int j = JW * jt + a;
      if (j < J) {
#pragma unroll
        for (int b = 0; b < IW; ++b) {
          const int i = IW * it + b;
          if (i < I) P[(size_t)j * I + i] = old[a][b] + acc[a][b];
        }
      }
    }
  }
  if (Pb != nullptr) {
    for (int j = tid; j < J; j += NT) {
      const float o = accumulate ? Pb[j] : 0.f;
      float s = 0.f;
      for (int e = 0; e < TE; ++e) s += G[e * ldg + j];
      Pb[j] = o + s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// segment reduction fused into the edge kernels (GCPMessagePassing.aggregate, gcpnet.py:938-947)
// ------------------------------------------------------------------------------------------
// The edge kernels never write per-edge messages: edges are sorted by destination, so a tile sums the rows of every
// destination it holds in shared memory (fixed row order) and emits
//   buf[d][W]                   the segment's sum, when the whole segment lies inside ONE tile (t0 == t1 below), or
//   carry[tile][0 | 1][W]       the partial sum of the tile's first (started in an earlier tile) / last (continues in the
//                               next tile) segment,  carry = buf + N * W.
// A destination whose segment spans tiles t0 < ... < t1 is the last segment of t0 and the first of every later tile:
// consumers add those partials in tile order -- still a fixed order, still no atomics.
GCP_HD float segment_total(const float* buf, long long N, int W, int R, const int* dst_ptr, int i, int f) {
  const int a = dst_ptr[i], b = dst_ptr[i + 1];
  if (a == b) return 0.f;
  const int t0 = a / R, t1 = (b - 1) / R;
  if (t0 == t1) return GCP_LDG(buf + (size_t)i * W + f);
  const float* carry = buf + (size_t)N * W;
  float acc = GCP_LDG(carry + ((size_t)t0 * 2 + 1) * W + f);
  for (int t = t0 + 1; t <= t1; ++t) acc += GCP_LDG(carry + ((size_t)t * 2) * W + f);
  return acc;
}

// ------------------------------------------------------------------------------------------
// cooperative row copies: warp per row, lane per column (coalesced, no integer division)
// ------------------------------------------------------------------------------------------
// dst (smem, row stride ldd) <- rows of a global matrix with `len` contiguous floats per row,
// row index given by rowidx(e) (< 0: zero fill).
template <int TE, int NT, class RowIdx>
GCP_HD void tile_load_rows(float* dst, int ldd, const float* src, int len, RowIdx rowidx, int tid) {
  const int lane = tid & 31;
  for (int e = tid >> 5; e < TE; e += NT / 32) {
    const long long r = rowidx(e);
    const float* sp = src + (size_t)(r < 0 ? 0 : r) * len;
    float* dp = dst + e * ldd;
#if GCP_DEVICE_CODE
    if ((((unsigned)len | (unsigned)ldd) & 3u) == 0 && ((((size_t)src) | ((size_t)dst)) & 15u) == 0) {  // 16-byte path
      for (int f = 4 * lane; f < len; f += 128)
        *reinterpret_cast<float4*>(dp + f) = r >= 0 ? __ldg(reinterpret_cast<const float4*>(sp + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
#endif
    for (int f = lane; f < len; f += 32) dp[f] = r >= 0 ? GCP_LDG(sp + f) : 0.f;
  }
}
// global rows [row0 + e] (dense, len floats) <- smem rows, for e < nrows
template <int TE, int NT>
GCP_HD void tile_store_rows(float* dst, long long row0, int len, const float* src, int lds, int nrows, int tid) {
  const int lane = tid & 31;
  for (int e = tid >> 5; e < nrows; e += NT / 32) {
    float* dp = dst + (size_t)(row0 + e) * len;
    const float* sp = src + e * lds;
#if GCP_DEVICE_CODE
    if ((((unsigned)len | (unsigned)lds) & 3u) == 0 && ((((size_t)src) | ((size_t)dst)) & 15u) == 0) {  // 16-byte path
      for (int f = 4 * lane; f < len; f += 128) *reinterpret_cast<float4*>(dp + f) = *reinterpret_cast<const float4*>(sp + f);
      continue;
    }
#endif
    for (int f = lane; f < len; f += 32) dp[f] = sp[f];
  }
}

// ------------------------------------------------------------------------------------------
// GCP2 forward pieces
// ------------------------------------------------------------------------------------------
// HD[e][x][:] = sum_c V[e][c][x] * WdT[c][:]   (vector_down + vector_down_frames, gcpnet.py:420,426)
// item = (e, x, 4-column group); WdT read from `wdt` (the S chunk in its ring slot).
template <int TE, int NT>
GCP_HD void gcp2_vec_down(const GcpOp& op, const TileBufs& b, const float* wdt, int tid) {
  const int cols = op.w.cols, ng = cols >> 2;
  for (int item = tid; item < 3 * TE * ng; item += NT) {
    const int e = item % TE, r = item / TE;  // compile-time divisor
    const int x = r % 3, g = r / 3;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* vp = b.V + e * b.ldv + x;
    const float* w = wdt + 4 * g;
    for (int c = 0; c < op.vi; ++c) {
      const float v = vp[3 * c];
      const float4 wv = ld4(w + c * cols);
      acc.x = fmaf(v, wv.x, acc.x); acc.y = fmaf(v, wv.y, acc.y); acc.z = fmaf(v, wv.z, acc.z); acc.w = fmaf(v, wv.w, acc.w);
    }
    st4(b.HD + e * b.ldhd + x * cols + 4 * g, acc);
  }
}

// norms (safe_norm over xyz, gcpnet.py:421) and frame scalars (scalarize, comp:302-312) -> Z[:, si:K);
// columns [K, kpad) are zeroed (the GEMM reads them against zero weights).
template <int TE, int NT>
GCP_HD void gcp2_norm_scalarize(const GcpOp& op, const TileBufs& b, int e3, int tid) {
  const int cols = op.w.cols, hdp = op.w.hdp;
  const int nq = op.hd + gcp_nfs(op), kpad = gcp_kpad(op), K = op.si + nq;
  const int ncol = kpad - op.si;  // nq real columns + zero padding
  for (int item = tid; item < TE * ncol; item += NT) {
    const int e = item % TE, j = item / TE;
    const float* hp = b.HD + e * b.ldhd;
    float val = 0.f;
    if (j < op.hd) {
      const float a = hp[j], bb = hp[cols + j], c = hp[2 * cols + j];
      val = sqrtf(fmaf(a, a, fmaf(bb, bb, c * c)) + SAFE_NORM_EPS) + SAFE_NORM_EPS;
    } else if (j < nq) {
      const int q = j - op.hd;
      const int cc = q / 3, a = q - 3 * cc;  // q[3*cc + a]
      const float* f = b.F + e * LDF + 3 * a;
      val = f[0] * hp[hdp + cc];
      val = fmaf(f[1], hp[cols + hdp + cc], val);
      val = fmaf(f[2], hp[2 * cols + hdp + cc], val);
      if (e3 && a == 1) {
        if (b.e3n == nullptr) val = fabsf(val);  // comp:305-309 (edge entities: one frame per row)
        else {  // node entities: mean over the outgoing edges of |x_cross . D|
          const NodeE3& n3 = *b.e3n;
          val = 0.f;
          if (e < n3.nrows) {
            const int i = n3.row0 + e, q0 = n3.src_ptr[i], q1 = n3.src_ptr[i + 1];
            const float d0 = hp[hdp + cc], d1 = hp[cols + hdp + cc], d2 = hp[2 * cols + hdp + cc];
            for (int q = q0; q < q1; ++q) {
              const float* fe = n3.frames + (size_t)n3.perm[n3.src_pos[q]] * 9 + 3;
              val += fabsf(fmaf(GCP_LDG(fe + 2), d2, fmaf(GCP_LDG(fe + 1), d1, GCP_LDG(fe) * d0)));
            }
            if (q1 > q0) val /= (float)(q1 - q0);
          }
        }
      }
    }
    b.Z[e * b.ldz + op.si + j] = val;
  }
  (void)K;
}

// Ungated vector output U[e][o][x] = sum_k H[e][x][k] * Wu[o][k] (+ V_in[e][o][x] if vector_residual)
GCP_HD float gcp2_vec_up(const GcpOp& op, const TileBufs& b, const float* wu, int e, int o, int x) {
  const int cols = op.w.cols, hdp = op.w.hdp;
  const float* hp = b.HD + e * b.ldhd + x * cols;
  const float* w = wu + o * hdp;
  float u = 0.f;
  for (int k4 = 0; k4 < hdp; k4 += 4) {
    const float4 h = ld4(hp + k4), ww = ld4(w + k4);
    u = fmaf(h.x, ww.x, u); u = fmaf(h.y, ww.y, u); u = fmaf(h.z, ww.z, u); u = fmaf(h.w, ww.w, u);
  }
  if (op.vres) u += b.V[e * b.ldv + 3 * o + x];
  return u;
}

// GCP2 forward up to (T, SG).  On return the G chunk is still the ring's head (the caller reads
// WU from it in its update phase, then releases it with gcp2_fwd_finish()).
// Phases: [vec_down] [norm+scalarize] [GEMM chunk]* [gate].
template <int TE, int NT, int SL>
GCP_HDN const float* gcp2_fwd_tile(const GcpOp& op, const TileBufs& b, WPipe& wp, int e3, float slope, bool refill_first) {
  const GcpW& W = op.w;
  // ---- S chunk: vector_down / vector_down_frames
  GCP_PHASE_BEGIN(NT)
  if (refill_first) wpipe_refill(wp, wp.head - 1, tid);  // slot freed by the caller's previous phase
  const float* wdt = wpipe_wait(wp);
  gcp2_vec_down<TE, NT>(op, b, wdt, tid);
  GCP_PHASE_END
  wp.head++;
  GCP_PHASE_BEGIN(NT)
  wpipe_refill(wp, wp.head - 1, tid);
  gcp2_norm_scalarize<TE, NT>(op, b, e3, tid);
  GCP_PHASE_END
  // ---- scalar_out: T = Z * Ws^T + bs, K-chunked, accumulators stay in registers across chunks
  {
    float* T = b.T; const int ldt = b.ldt;
    const int nslices = W.NP >> 4;
#if GCP_DEVICE_CODE
    float acc[SL][2][4];
    const GemmMap<TE, NT> m((int)threadIdx.x);
    for (int c = 0; c < W.nWS; ++c) {
      if (c > 0) wpipe_refill(wp, wp.head - 1, (int)threadIdx.x);
      const float* wc = wpipe_wait(wp);
      if (c == 0) {
        const float* bias = wc + W.NP * W.ldk;
#pragma unroll
        for (int i = 0; i < SL; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_n(j);
            acc[i][0][j] = acc[i][1][j] = (n < W.NP) ? bias[n] : 0.f;
          }
      }
      gemm_nmajor_chunk<TE, NT, SL>(acc, b.Z, b.ldz, c * W.kc, W.kc, wc, W.ldk, nslices, m, XIdentity());
      if (c == W.nWS - 1) {
#pragma unroll
        for (int i = 0; i < SL; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_n(j);
            if (n < op.so) { T[m.e0 * ldt + n] = acc[i][0][j]; T[(m.e0 + 8) * ldt + n] = acc[i][1][j]; }
          }
      }
      __syncthreads();
      wp.head++;
    }
#else
    // host emulation: per-thread accumulators cannot live across phases -> park them in a scratch array
    static thread_local float accs[1024][SL][2][4];
    for (int c = 0; c < W.nWS; ++c) {
      GCP_PHASE_BEGIN(NT)
      if (c > 0) wpipe_refill(wp, wp.head - 1, tid);
      const float* wc = wpipe_wait(wp);
      const GemmMap<TE, NT> m(tid);
      float (&acc)[SL][2][4] = accs[tid];
      if (c == 0) {
        const float* bias = wc + W.NP * W.ldk;
        for (int i = 0; i < SL; ++i)
          for (int j = 0; j < 4; ++j) {
            const int n = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_n(j);
            acc[i][0][j] = acc[i][1][j] = (n < W.NP) ? bias[n] : 0.f;
          }
      }
      gemm_nmajor_chunk<TE, NT, SL>(acc, b.Z, b.ldz, c * W.kc, W.kc, wc, W.ldk, nslices, m, XIdentity());
      if (c == W.nWS - 1) {
        for (int i = 0; i < SL; ++i)
          for (int j = 0; j < 4; ++j) {
            const int n = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_n(j);
            if (n < op.so) { T[m.e0 * ldt + n] = acc[i][0][j]; T[(m.e0 + 8) * ldt + n] = acc[i][1][j]; }
          }
      }
      GCP_PHASE_END
      wp.head++;
    }
#endif
  }
  // ---- gate: SG = sigmoid(act_v(T) * Wg^T + bg)   (reads the PRE-activation scalars, gcpnet.py:386)
  const float* gch = nullptr;
  GCP_PHASE_BEGIN(NT)
  wpipe_refill(wp, wp.head - 1, tid);
  const float* g = wpipe_wait(wp);
  const XAct xa{op.act_v, slope};
  // thread = (e, o-group): e = tid % TE, o = tid / TE + (NT/TE) * j
  const int e = tid % TE;
  const float* tp = b.T + e * b.ldt;
  if (!gcp_gated(op)) {  // vector_gate=False, identity vector nonlinearity: V' = U (gcpnet.py:344-350)
    for (int o = tid / TE; o < op.vo; o += NT / TE) b.SG[e * b.ldsg + o] = 1.f;
  } else
  for (int o0 = tid / TE; o0 < op.vo; o0 += 2 * (NT / TE)) {
    const int o1 = o0 + NT / TE;
    const bool has1 = o1 < op.vo;
    const float* w0 = g + o0 * W.ldg;
    const float* w1 = g + (has1 ? o1 : o0) * W.ldg;
    float a0 = g[W.o_bg + o0], a1 = has1 ? g[W.o_bg + o1] : 0.f;
    for (int k4 = 0; k4 < op.so; k4 += 4) {
      float4 t = ld4(tp + k4);
      t.x = xa(t.x); t.y = xa(t.y); t.z = xa(t.z); t.w = xa(t.w);
      const float4 u = ld4(w0 + k4), v = ld4(w1 + k4);
      a0 = fmaf(t.x, u.x, a0); a0 = fmaf(t.y, u.y, a0); a0 = fmaf(t.z, u.z, a0); a0 = fmaf(t.w, u.w, a0);
      a1 = fmaf(t.x, v.x, a1); a1 = fmaf(t.y, v.y, a1); a1 = fmaf(t.z, v.z, a1); a1 = fmaf(t.w, v.w, a1);
    }
    b.SG[e * b.ldsg + o0] = sigmoidf_(a0);
    if (has1) b.SG[e * b.ldsg + o1] = sigmoidf_(a1);
  }
  GCP_PHASE_END
  gch = wp.slots + (size_t)(wp.head % wp.seq->nslot) * wp.seq->slot_floats;
  return gch;  // G chunk (WG | bg | WU), still held
}

// ------------------------------------------------------------------------------------------
// GCP2 backward on a tile
// ------------------------------------------------------------------------------------------
struct BwdBufs {
  float* GS;  int ldgs;   // [TE][ldgs]  in: grad wrt scalar output s' (post activation)
  float* GV;  int ldgv;   // [TE][ldgv]  in: grad wrt vector output V' (3*vo, scalar stride)
  float* GU;  int ldgu;   // [TE][ldgu]  scratch: grad wrt ungated U (3*vo)
  float* GG;  int ldgg;   // [TE][ldgg]  scratch: grad wrt gate pre-activation (vo), float4-read by wgrad
  float* GNQ; int ldnq;   // [TE][ldnq]  scratch: grad wrt [norms | frame scalars] (hd+9)
  float* GHD; int ldghd;  // [TE][ldghd] scratch: grad wrt HD (same layout as HD)
  // Optional spill of the operands of the two large weight-gradient products (scalar_out, vector_out_scale): when set,
  // the tile stores its rows of gT [so -> 4], Z [K -> 4] and gg [vo -> 4] (dense rows, node index = sp_row0 + e) and one
  // output-parallel kernel over ALL rows forms the products (node_wgrad.cuh) instead of a 16-row partial per CTA.
  float* sp_gT = nullptr; float* sp_Z = nullptr; float* sp_GG = nullptr;
  long long sp_row0 = 0; int sp_nrows = 0;
};

// Backward of one GCP2 on a tile.  On entry: b.Z[:, :si], b.V, b.F hold the forward inputs, b.T the
// saved pre-activations, b.SG the saved gates, g.GS / g.GV the output cotangents (rows >= nrows zero).
// The routine recomputes HD, norms and frame scalars, then produces
//   * scalar input cotangent: emitted through `emit_s(e, i, value)` (i < si)
//   * vector input cotangent: emitted through `emit_v(e, c3, value)` (c3 < 3*vi)
//   * weight-gradient partials into prow[op.o_*]  (accumulate: += instead of =)
// b.T is overwritten with the cotangent of the pre-activation.  Ring order: S, G, WS chunks.
// `refill_first`: the caller's previous phase released the chunk at head-1.
// accumulate-or-assign into a shared-memory tile: the emit functor of every caller that keeps the cotangent on chip
// (one functor TYPE -> one instantiation of gcp2_bwd_tile, called several times instead of inlined several times)
struct EmitTile {
  float* p; int ld; bool add;
  GCP_HD void operator()(int e, int i, float val) const { float* d = p + e * ld + i; *d = add ? *d + val : val; }
};
template <int TE, int NT, int SLF, int SLD, class EmitS, class EmitV>
GCP_HDN void gcp2_bwd_tile(const GcpOp& op, const TileBufs& b, const BwdBufs& g, WPipe& wp, int e3, float slope,
                           float* prow, bool accumulate, bool refill_first, EmitS emit_s, EmitV emit_v) {
  const GcpW& W = op.w;
  const int cols = W.cols, hdp = W.hdp;
  const int K = gcp_k(op);
  float* wdt_sm = b.WSM;                    // [vi][cols]
  float* wu_sm = b.WSM + op.vi * cols;      // [vo][hdp]
  constexpr int IW = 8;  // 4 x 8 blocks: 32 FMAs per 3 shared loads (4 x 4 blocks were shared-memory bound: r2 stage stamps, 23 k -> cycles per GCP)
  (void)hdp;
  // development aid: stamps 0..8 = phase boundaries, 9..14 = inside the first scalar_out chunk
#if GCP_DEVICE_CODE && GCP_STAMPS
  int dbg_i = 0;
#define GCP_BSTAMP() do { if (wp.dbg != nullptr && threadIdx.x == 0) wp.dbg[dbg_i] = clock64(); ++dbg_i; } while (0)
#define GCP_BSTAMP_AT(i) do { if (wp.dbg != nullptr && threadIdx.x == 0 && c == 0) wp.dbg[i] = clock64(); } while (0)
#else
#define GCP_BSTAMP() do { } while (0)
#define GCP_BSTAMP_AT(i) do { } while (0)
#endif
  GCP_BSTAMP();
  // ---- recompute: vector_down (S chunk; keep WdT for the final phases)
  GCP_PHASE_BEGIN(NT)
  if (refill_first) wpipe_refill(wp, wp.head - 1, tid);
  const float* wdt = wpipe_wait(wp);
  gcp2_vec_down<TE, NT>(op, b, wdt, tid);
  for (int i = tid; i < op.vi * cols; i += NT) wdt_sm[i] = wdt[i];
  GCP_PHASE_END
  GCP_BSTAMP();
  wp.head++;
  GCP_PHASE_BEGIN(NT)
  wpipe_refill(wp, wp.head - 1, tid);
  gcp2_norm_scalarize<TE, NT>(op, b, e3, tid);
  GCP_PHASE_END
  GCP_BSTAMP();
  float* T = b.T; const int ldt = b.ldt;
  // ---- gate backward (G chunk): gU = gV' * sg ; gsig = sum_x gV' * U ; gg = gsig * sg * (1 - sg)
  GCP_PHASE_BEGIN(NT)
  const float* gc = wpipe_wait(wp);
  const float* wu = gc + W.o_wu;
  for (int i = tid; i < op.vo * hdp; i += NT) wu_sm[i] = wu[i];
  const int e = tid % TE;
  for (int o = tid / TE; o < op.vo; o += NT / TE) {
    const float sg = b.SG[e * b.ldsg + o];
    float gsig = 0.f;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const float gv = g.GV[e * g.ldgv + 3 * o + x];
      gsig = fmaf(gv, gcp2_vec_up(op, b, wu, e, o, x), gsig);
      g.GU[e * g.ldgu + 3 * o + x] = gv * sg;
    }
    g.GG[e * g.ldgg + o] = gsig * sg * (1.f - sg);
  }
  // zero the float4 padding columns of GG so the 4-wide reads below see zeros
  for (int o = op.vo + tid / TE; o < round_up(op.vo, 4); o += NT / TE) g.GG[e * g.ldgg + o] = 0.f;
  GCP_PHASE_END
  GCP_BSTAMP();
  // ---- vector_out_scale / vector_up weight gradients (read ALL of T = pre-activations: own phase)
  GCP_PHASE_BEGIN(NT)
  if (!gcp_gated(op)) { }  // no vector_out_scale (sg = 1 -> gg = 0 above)
  else if (g.sp_GG != nullptr)
    tile_store_rows<TE, NT>(g.sp_GG, g.sp_row0, round_up(op.vo, 4), g.GG, g.ldgg, g.sp_nrows, tid);
  else
    tile_wgrad<TE, NT, 1, 4>(g.GG, g.ldgg, op.vo, T, ldt, op.so, prow + op.o_Wg, prow + op.o_bg, accumulate,
                             XAct{op.act_v, slope}, tid);
  // gWu[o][k] = sum_{e,x} gU[e][o][x] * H[e][x][k]
  for (int item = tid; item < op.vo * op.hd; item += NT) {
    const int o = item / op.hd, k = item - o * op.hd;
    float s = 0.f;
    for (int e = 0; e < TE; ++e) {
      const float* hp = b.HD + e * b.ldhd + k;
      const float* gu = g.GU + e * g.ldgu + 3 * o;
      s = fmaf(gu[0], hp[0], s); s = fmaf(gu[1], hp[cols], s); s = fmaf(gu[2], hp[2 * cols], s);
    }
    float* dst = prow + op.o_Wu + item;
    *dst = (accumulate ? *dst : 0.f) + s;
  }
  GCP_PHASE_END
  GCP_BSTAMP();
  // ---- cotangent of the pre-activation: gT = gS' * act_s'(T) + act_v'(T) * (gg * Wg)   (T <- gT in place;
  //      every thread reads only the T entries it overwrites)
  GCP_PHASE_BEGIN(NT)
  const float* gc = wp.slots + (size_t)(wp.head % wp.seq->nslot) * wp.seq->slot_floats;
  const GemmMap<TE, NT> m(tid);
  float acc[SLF][2][4];
  gemm_kmajor_chunk<TE, NT, SLF>(acc, g.GG, g.ldgg, round_up(op.vo, 4), gc, W.ldg, (op.so + 15) >> 4, m);
  const int as = op.act_s, av = op.act_v;
#pragma unroll
  for (int i = 0; i < SLF; ++i)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_k(c);
        const int e = m.e0 + 8 * r;
        if (n < op.so) {
          const float t = T[e * ldt + n];
          T[e * ldt + n] = fmaf(g.GS[e * g.ldgs + n], act_grad(as, t, slope), acc[i][r][c] * act_grad(av, t, slope));
        }
      }
  GCP_PHASE_END
  GCP_BSTAMP();
  wp.head++;  // G chunk released
  // ---- scalar_out: weight gradient + data gradient gz = gT * Ws (WS chunks, k-major view)
  for (int c = 0; c < W.nWS; ++c) {
    GCP_PHASE_BEGIN(NT)
    GCP_BSTAMP_AT(9);
    wpipe_refill(wp, wp.head - 1, tid);
    const float* wc = wpipe_wait(wp);
    GCP_BSTAMP_AT(10);
    if (c == 0) {
      if (g.sp_gT != nullptr) {
        tile_store_rows<TE, NT>(g.sp_gT, g.sp_row0, round_up(op.so, 4), T, ldt, g.sp_nrows, tid);
        tile_store_rows<TE, NT>(g.sp_Z, g.sp_row0, round_up(K, 4), b.Z, b.ldz, g.sp_nrows, tid);
      } else {
        tile_wgrad<TE, NT, 4, IW>(T, ldt, op.so, b.Z, b.ldz, K, prow + op.o_Ws, prow + op.o_bs, accumulate, XIdentity(), tid);
      }
    }
    GCP_BSTAMP_AT(11);
    const GemmMap<TE, NT> m(tid);
    float acc[SLD][2][4];
    gemm_kmajor_chunk<TE, NT, SLD>(acc, T, ldt, round_up(op.so, 4), wc, W.ldk, W.kc >> 4, m);
    GCP_BSTAMP_AT(12);
#pragma unroll
    for (int i = 0; i < SLD; ++i)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int il = 16 * (m.wn + GemmMap<TE, NT>::WN * i) + m.col_k(cc);
          const int col = c * W.kc + il;
          const int e = m.e0 + 8 * r;
          if (il < W.kc && col < K) {
            if (col < op.si) emit_s(e, col, acc[i][r][cc]);
            else g.GNQ[e * g.ldnq + (col - op.si)] = acc[i][r][cc];
          }
        }
    GCP_BSTAMP_AT(13);
    GCP_PHASE_END
    GCP_BSTAMP_AT(14);
    wp.head++;
  }
  GCP_BSTAMP();
  // ---- gHD: norms, vector_up and frame scalars back to the hidden vector channels
  GCP_PHASE_BEGIN(NT)
  wpipe_refill(wp, wp.head - 1, tid);
  for (int item = tid; item < 3 * TE * cols; item += NT) {
    const int e = item % TE, r = item / TE;  // compile-time divisors
    const int x = r % 3, kk = r / 3;         // kk < cols: hidden column
    const float* hp = b.HD + e * b.ldhd;
    const float* gnq = g.GNQ + e * g.ldnq;
    float acc = 0.f;
    if (kk < op.hd) {
      // n = sqrt(sum_x H^2 + eps) + eps  ->  dn/dH[x] = H[x] / (n - eps)
      const float a = hp[kk], bb = hp[cols + kk], c = hp[2 * cols + kk];
      const float root = sqrtf(fmaf(a, a, fmaf(bb, bb, c * c)) + SAFE_NORM_EPS);
      acc = gnq[kk] * hp[x * cols + kk] / root;
      for (int o = 0; o < op.vo; ++o) acc = fmaf(g.GU[e * g.ldgu + 3 * o + x], wu_sm[o * hdp + kk], acc);
    } else if (kk >= hdp && kk < hdp + 3 && gcp_nfs(op) != 0) {
      // frame scalars: q[3cc+a] = sum_x F[a][x] D[x][cc]  (|.| on a==1 when e3)
      const int cc = kk - hdp;
      for (int a = 0; a < 3; ++a) {
        float gq = gnq[op.hd + 3 * cc + a];
        if (e3 && a == 1 && b.e3n != nullptr) {
          // node entities: d/dD[x] of mean_edges |F1 . D| = mean_edges sign(F1 . D) F1[x]
          const NodeE3& n3 = *b.e3n;
          if (e < n3.nrows) {
            const int i = n3.row0 + e, q0 = n3.src_ptr[i], q1 = n3.src_ptr[i + 1];
            const float d0 = hp[hdp + cc], d1 = hp[cols + hdp + cc], d2 = hp[2 * cols + hdp + cc];
            float sx = 0.f;
            for (int q = q0; q < q1; ++q) {
              const float* fe = n3.frames + (size_t)n3.perm[n3.src_pos[q]] * 9 + 3;
              const float pr = fmaf(GCP_LDG(fe + 2), d2, fmaf(GCP_LDG(fe + 1), d1, GCP_LDG(fe) * d0));
              sx += pr > 0.f ? GCP_LDG(fe + x) : (pr < 0.f ? -GCP_LDG(fe + x) : 0.f);
            }
            if (q1 > q0) acc = fmaf(sx / (float)(q1 - q0), gq, acc);
          }
          continue;
        }
        if (e3 && a == 1) {
          const float* f = b.F + e * LDF + 3;
          float q = f[0] * hp[hdp + cc];
          q = fmaf(f[1], hp[cols + hdp + cc], q);
          q = fmaf(f[2], hp[2 * cols + hdp + cc], q);
          gq = q > 0.f ? gq : (q < 0.f ? -gq : 0.f);
        }
        acc = fmaf(b.F[e * LDF + 3 * a + x], gq, acc);
      }
    }
    g.GHD[e * g.ldghd + x * cols + kk] = acc;
  }
  GCP_PHASE_END
  GCP_BSTAMP();
  // ---- vector_down / vector_down_frames weight gradients; vector input cotangent
  GCP_PHASE_BEGIN(NT)
  // gWdT[c][k] = sum_{e,x} gHD[e][x][k] * V[e][c][x]
  for (int item = tid; item < op.vi * (op.hd + gcp_nfs(op) / 3); item += NT) {
    const int kk = item / op.vi, c = item - kk * op.vi;  // kk < hd: Wd row kk ; else Wdf row kk-hd
    const int col = kk < op.hd ? kk : hdp + (kk - op.hd);
    float s = 0.f;
    for (int e = 0; e < TE; ++e) {
      const float* ghp = g.GHD + e * g.ldghd + col;
      const float* vp = b.V + e * b.ldv + 3 * c;
      s = fmaf(ghp[0], vp[0], s); s = fmaf(ghp[cols], vp[1], s); s = fmaf(ghp[2 * cols], vp[2], s);
    }
    float* dst = kk < op.hd ? prow + op.o_Wd + kk * op.vi + c : prow + op.o_Wdf + (kk - op.hd) * op.vi + c;
    *dst = (accumulate ? *dst : 0.f) + s;
  }
  // gV[e][c][x] = sum_k gHD[e][x][k] * WdT[c][k] (+ gU[e][c][x] if vector_residual)
  {
    const int e = tid % TE;
    for (int c = tid / TE; c < op.vi; c += NT / TE) {
      const float* w = wdt_sm + c * cols;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float* ghp = g.GHD + e * g.ldghd + x * cols;
        float acc = 0.f;
        for (int k4 = 0; k4 < cols; k4 += 4) {
          const float4 gv = ld4(ghp + k4), wv = ld4(w + k4);
          acc = fmaf(gv.x, wv.x, acc); acc = fmaf(gv.y, wv.y, acc); acc = fmaf(gv.z, wv.z, acc); acc = fmaf(gv.w, wv.w, acc);
        }
        if (op.vres && op.vo > 0) acc += g.GU[e * g.ldgu + 3 * c + x];
        emit_v(e, 3 * c + x, acc);
      }
    }
  }
  GCP_PHASE_END
  GCP_BSTAMP();
#if GCP_DEVICE_CODE && GCP_STAMPS
  if (wp.dbg != nullptr) wp.dbg += 16;
#endif
}

// Out-of-line entry points for kernels that run each GCP ONCE per tile (node update): three inlined copies of these
// routines do not fit the instruction cache, one shared copy does.  The edge kernels loop over their GCPs and inline.
template <int TE, int NT, int SL>
GCP_HDN_NOINLINE const float* gcp2_fwd_tile_call(const GcpOp& op, const TileBufs& b, WPipe& wp, int e3, float slope, bool refill_first) {
  return gcp2_fwd_tile<TE, NT, SL>(op, b, wp, e3, slope, refill_first);
}
template <int TE, int NT, int SLF, int SLD>
GCP_HDN_NOINLINE void gcp2_bwd_tile_call(const GcpOp& op, const TileBufs& b, const BwdBufs& g, WPipe& wp, int e3, float slope,
                                         float* prow, bool accumulate, bool refill_first, EmitTile emit_s, EmitTile emit_v) {
  gcp2_bwd_tile<TE, NT, SLF, SLD>(op, b, g, wp, e3, slope, prow, accumulate, refill_first, emit_s, emit_v);
}

}  // namespace gcp
