// gcp_tile.cuh -- CTA-tile building blocks of the GCP2 perceptron (forward and backward).
//
// Everything here works on a TILE of TE "entities" (edges in the message kernel, nodes in the
// node-update kernel) whose features sit in shared memory, entity-major: X[e][f], row stride ld.
// Reference semantics: GCP2.forward, src/models/components/gcpnet.py:393-468 (+ :353-391 for the
// vector gate), scalarize src/models/components/__init__.py:272-325, safe_norm :381-392.
//
// Code style: a routine is a sequence of PHASES.  A phase is a parallel-for over the CTA's NT
// threads followed by a barrier; no value lives in a register across a phase boundary.  On the
// device a phase is `{ tid = threadIdx.x; ... } __syncthreads();`.  The same source also compiles
// for the host, where a phase runs the NT thread bodies one after another (optionally in reverse
// order, to expose intra-phase hazards) -- that build is the CPU emulation used by the non-GPU
// tests (tests/emul); it is never part of the product path.
//
// Bank-conflict rules used for the strides (floats):
//   * arrays read with float4 along the feature axis by lanes that differ in the row:
//       ld % 8 == 4  -> 8 consecutive rows hit 8 disjoint 4-bank groups          (ld_vec())
//   * arrays read scalar by lanes that differ in the row: ld odd                  (ld_scal())
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gcp {

#define GCP_HD __host__ __device__ __forceinline__
#define GCP_HDN __host__ __device__

inline int g_emul_reverse = 0;  // host emulation only: run thread bodies in reverse order when set
#if defined(__CUDA_ARCH__)
#define GCP_PHASE_BEGIN(NT) { const int tid = (int)threadIdx.x; (void)tid;
#define GCP_PHASE_END } __syncthreads();
#define GCP_LDG(p) __ldg(p)
#else
#define GCP_PHASE_BEGIN(NT) for (int tid_ = 0; tid_ < (NT); ++tid_) { const int tid = ::gcp::g_emul_reverse ? (NT) - 1 - tid_ : tid_; (void)tid;
#define GCP_PHASE_END }
#define GCP_LDG(p) (*(p))
#endif

constexpr int MAX_MSG_LAYERS = 12;
constexpr float SAFE_NORM_EPS = 1e-8f;  // comp/__init__.py:385

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKYRELU = 2, ACT_SILU = 3, ACT_SIGMOID = 4, ACT_SELU = 5 };

GCP_HD float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// src/models/__init__.py:41-57
GCP_HD float act_fwd(int a, float x, float slope) {
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
GCP_HD float act_grad(int a, float x, float slope) {
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

// One GCP2 module (device view).  Weight pointers use the nn.Linear layouts of the reference
// (gcpnet.py:303-322): Wd[hd][vi], Wdf[3][vi], Ws[so][si+hd+9], bs[so], Wu[vo][hd], Wg[vo][so], bg[vo].
// o_* = offsets (floats) of the matching gradient blocks inside a per-CTA partial-gradient row.
struct GcpOp {
  int si, vi, so, vo, hd;
  int act_s, act_v, vres;
  const float *Wd, *Wdf, *Ws, *bs, *Wu, *Wg, *bg;
  int o_Wd, o_Wdf, o_Ws, o_bs, o_Wu, o_Wg, o_bg;
  int pad_;
};
GCP_HD int gcp_k(const GcpOp& op) { return op.si + op.hd + 9; }

// Shared-memory views used by one GCP2 evaluation on a tile.
struct TileBufs {
  float* Z;   int ldz;   // [TE][ldz]  cols [0,si)=scalars in, [si,si+hd)=norms, [si+hd,si+hd+9)=frame scalars; pad cols zero
  float* V;   int ldv;   // [TE][ldv]  3*vi floats, (channel, xyz) xyz fastest, scalar stride
  float* HD;  int ldhd;  // [TE][ldhd] 3 x hd_cols: H[x][0..hd), frame-down D[x][0..3) at column hd_cols-4
  float* F;              // [TE][9]    frames (a, xyz)
  float* T;   int ldt;   // [TE][ldt]  pre-activation scalar_out
  float* SG;  int ldsg;  // [TE][ldsg] sigmoid gate per output vector channel
  float* WC;  int wc_cap;  // weight chunk staging
  float* WS;             // small weights: WdT[vi][hd_cols], Wu[vo][hdp], biases
};
constexpr int LDF = 9;

// ------------------------------------------------------------------------------------------
// dense tile GEMMs, fp32 FFMA, register micro-tiles, weights staged through shared memory
// ------------------------------------------------------------------------------------------
struct XIdentity { GCP_HD float operator()(float x) const { return x; } };
struct XAct { int a; float slope; GCP_HD float operator()(float x) const { return act_fwd(a, x, slope); } };

// Y[e][n] = bias[n] + sum_k xmap(X[e][k]) * W[n][k]   (W global, nn.Linear layout [N][K]).
// Thread (eg, og) owns rows e = eg + EG*i (i<ER) and outputs n = n0 + og + OG*j (j<NR).
// K is split into chunks that fit WC; partial sums between chunks go through Yacc (smem, [TE][ldy]).
template <int TE, int NT, int OG, int NR, class XMap, class Epi>
GCP_HDN void tile_gemm_nmajor(const float* X, int ldx, int K, const float* W, int N, const float* bias,
                              float* Wc, int wc_cap, float* Yacc, int ldy, XMap xmap, Epi epi) {
  constexpr int EG = NT / OG;
  constexpr int ER = TE / EG;
  static_assert(EG * OG == NT && ER * EG == TE && ER >= 1, "bad gemm thread grid");
  static_assert(NT % 32 == 0, "NT must be a multiple of the warp size");
  constexpr int NCH = OG * NR;
  const int K4 = round_up(K, 4);
  const int kcmax = ((wc_cap / NCH) - 4) & ~3;
  const int nkc = (K4 + kcmax - 1) / kcmax;
  for (int kci = 0; kci < nkc; ++kci) {
    const int k0 = kci * kcmax;
    const int kc = (K4 - k0) < kcmax ? (K4 - k0) : kcmax;
    const int ldw = ld_vec(kc);
    for (int n0 = 0; n0 < N; n0 += NCH) {
      GCP_PHASE_BEGIN(NT)
      for (int n = tid >> 5; n < NCH; n += NT / 32) {
        const int gn = n0 + n;
        const float* wrow = W + (size_t)gn * K + k0;
        for (int kk = tid & 31; kk < kc; kk += 32)
          Wc[n * ldw + kk] = (gn < N && k0 + kk < K) ? GCP_LDG(wrow + kk) : 0.f;
      }
      GCP_PHASE_END
      GCP_PHASE_BEGIN(NT)
      const int og = tid % OG, eg = tid / OG;
      float acc[ER][NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int n = n0 + og + OG * j;
#pragma unroll
        for (int i = 0; i < ER; ++i) {
          if (kci == 0) acc[i][j] = (bias != nullptr && n < N) ? GCP_LDG(bias + n) : 0.f;
          else acc[i][j] = (n < N) ? Yacc[(eg + EG * i) * ldy + n] : 0.f;
        }
      }
      const float* xbase = X + eg * ldx + k0;
      const float* wbase = Wc + og * ldw;
#pragma unroll 2
      for (int k4 = 0; k4 < kc; k4 += 4) {
        float4 xv[ER];
#pragma unroll
        for (int i = 0; i < ER; ++i) {
          xv[i] = ld4(xbase + i * EG * ldx + k4);
          xv[i].x = xmap(xv[i].x); xv[i].y = xmap(xv[i].y); xv[i].z = xmap(xv[i].z); xv[i].w = xmap(xv[i].w);
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          const float4 wv = ld4(wbase + j * OG * ldw + k4);
#pragma unroll
          for (int i = 0; i < ER; ++i) {
            acc[i][j] = fmaf(xv[i].x, wv.x, acc[i][j]);
            acc[i][j] = fmaf(xv[i].y, wv.y, acc[i][j]);
            acc[i][j] = fmaf(xv[i].z, wv.z, acc[i][j]);
            acc[i][j] = fmaf(xv[i].w, wv.w, acc[i][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int n = n0 + og + OG * j;
        if (n < N) {
#pragma unroll
          for (int i = 0; i < ER; ++i) {
            const int e = eg + EG * i;
            if (kci == nkc - 1) epi(e, n, acc[i][j]);
            else Yacc[e * ldy + n] = acc[i][j];
          }
        }
      }
      GCP_PHASE_END
    }
  }
}

// Y[e][n] = sum_k X[e][k] * W[k][n]   (W global, row-major [K][N]: the transpose-free form of the
// data-gradient through an nn.Linear whose weight is [K=out][N=in]).  NR multiple of 4.
// Thread (eg, og) owns rows e = eg + EG*i and outputs n = n0 + 4*(og + OG*j4) + c.
template <int TE, int NT, int OG, int NR, class Epi>
GCP_HDN void tile_gemm_kmajor(const float* X, int ldx, int K, const float* W, int N,
                              float* Wc, int wc_cap, float* Yacc, int ldy, Epi epi) {
  constexpr int EG = NT / OG;
  constexpr int ER = TE / EG;
  static_assert(EG * OG == NT && ER * EG == TE && ER >= 1, "bad gemm thread grid");
  static_assert(NR % 4 == 0, "NR must be a multiple of 4");
  constexpr int NR4 = NR / 4;
  constexpr int NCH = OG * NR;
  const int K4 = round_up(K, 4);
  const int kcmax = (wc_cap / NCH) & ~3;
  const int nkc = (K4 + kcmax - 1) / kcmax;
  for (int kci = 0; kci < nkc; ++kci) {
    const int k0 = kci * kcmax;
    const int kc = (K4 - k0) < kcmax ? (K4 - k0) : kcmax;
    for (int n0 = 0; n0 < N; n0 += NCH) {
      GCP_PHASE_BEGIN(NT)
      for (int idx = tid; idx < kc * NCH; idx += NT) {
        const int kk = idx / NCH, n = idx - kk * NCH;
        Wc[idx] = (k0 + kk < K && n0 + n < N) ? GCP_LDG(W + (size_t)(k0 + kk) * N + n0 + n) : 0.f;
      }
      GCP_PHASE_END
      GCP_PHASE_BEGIN(NT)
      const int og = tid % OG, eg = tid / OG;
      float acc[ER][NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int n = n0 + 4 * (og + OG * (j >> 2)) + (j & 3);
#pragma unroll
        for (int i = 0; i < ER; ++i)
          acc[i][j] = (kci > 0 && n < N) ? Yacc[(eg + EG * i) * ldy + n] : 0.f;
      }
      const float* xbase = X + eg * ldx + k0;
#pragma unroll 1
      for (int k4 = 0; k4 < kc; k4 += 4) {
        float4 xv[ER];
#pragma unroll
        for (int i = 0; i < ER; ++i) xv[i] = ld4(xbase + i * EG * ldx + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int j4 = 0; j4 < NR4; ++j4) {
            const float4 wv = ld4(Wc + (k4 + kk) * NCH + 4 * (og + OG * j4));
#pragma unroll
            for (int i = 0; i < ER; ++i) {
              const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
              acc[i][4 * j4 + 0] = fmaf(x, wv.x, acc[i][4 * j4 + 0]);
              acc[i][4 * j4 + 1] = fmaf(x, wv.y, acc[i][4 * j4 + 1]);
              acc[i][4 * j4 + 2] = fmaf(x, wv.z, acc[i][4 * j4 + 2]);
              acc[i][4 * j4 + 3] = fmaf(x, wv.w, acc[i][4 * j4 + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int n = n0 + 4 * (og + OG * (j >> 2)) + (j & 3);
        if (n < N) {
#pragma unroll
          for (int i = 0; i < ER; ++i) {
            const int e = eg + EG * i;
            if (kci == nkc - 1) epi(e, n, acc[i][j]);
            else Yacc[e * ldy + n] = acc[i][j];
          }
        }
      }
      GCP_PHASE_END
    }
  }
}

// Weight gradient of an nn.Linear on a tile:  P[j][i] (+)= sum_e G[e][j] * fmap(Zin[e][i]),
// Pb[j] (+)= sum_e G[e][j].  P/Pb are this CTA's private rows in global memory (deterministic:
// every element is owned by exactly one thread).  Rows e >= nrows of G must be zero.
// G and Zin are read 4 resp. 8 columns at a time: their arrays carry >= 8 floats of slack.
template <int TE, int NT, class FMap>
GCP_HDN void tile_wgrad(const float* G, int ldg, int J, const float* Zin, int ldz, int I,
                        float* P, float* Pb, bool accumulate, FMap fmap) {
  const int JT = (J + 3) / 4, IT = (I + 7) / 8;
  GCP_PHASE_BEGIN(NT)
  for (int tile = tid; tile < JT * IT; tile += NT) {
    const int jt = tile % JT, it = tile / JT;
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    const float* gp = G + 4 * jt;
    const float* zp = Zin + 8 * it;
#pragma unroll 2
    for (int e = 0; e < TE; ++e) {
      const float4 g = ld4(gp + e * ldg);
      float4 z0 = ld4(zp + e * ldz), z1 = ld4(zp + e * ldz + 4);
      const float gv[4] = {g.x, g.y, g.z, g.w};
      const float zv[8] = {fmap(z0.x), fmap(z0.y), fmap(z0.z), fmap(z0.w), fmap(z1.x), fmap(z1.y), fmap(z1.z), fmap(z1.w)};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(gv[a], zv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = 4 * jt + a;
      if (j < J) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int i = 8 * it + b;
          if (i < I) {
            float* dst = P + (size_t)j * I + i;
            *dst = (accumulate ? *dst : 0.f) + acc[a][b];
          }
        }
      }
    }
  }
  if (Pb != nullptr) {
    for (int j = tid; j < J; j += NT) {
      float s = 0.f;
      for (int e = 0; e < TE; ++e) s += G[e * ldg + j];
      Pb[j] = (accumulate ? Pb[j] : 0.f) + s;
    }
  }
  GCP_PHASE_END
}

// ------------------------------------------------------------------------------------------
// GCP2 forward on a tile
// ------------------------------------------------------------------------------------------
// Small-weight staging: WS = [ WdT: vi x hd_cols | Wu: vo x hdp ], zero padded.
//   WdT[c][k] = Wd[k][c] (k<hd), WdT[c][hd_cols-4+cc] = Wdf[cc][c] (cc<3)
template <int NT>
GCP_HDN void gcp2_stage_small(const GcpOp& op, float* WS) {
  const int cols = hd_cols(op.hd), hdp = cols - 4;
  GCP_PHASE_BEGIN(NT)
  for (int idx = tid; idx < op.vi * cols; idx += NT) {
    const int c = idx / cols, k = idx - c * cols;
    float w = 0.f;
    if (k < op.hd) w = GCP_LDG(op.Wd + k * op.vi + c);
    else if (k >= hdp && k < hdp + 3) w = GCP_LDG(op.Wdf + (k - hdp) * op.vi + c);
    WS[idx] = w;
  }
  float* WU = WS + op.vi * cols;
  for (int idx = tid; idx < op.vo * hdp; idx += NT) {
    const int o = idx / hdp, k = idx - o * hdp;
    WU[idx] = (k < op.hd) ? GCP_LDG(op.Wu + o * op.hd + k) : 0.f;
  }
  GCP_PHASE_END
}
GCP_HD int gcp2_small_floats(int vi, int vo, int hd) { return vi * hd_cols(hd) + vo * (hd_cols(hd) - 4); }

// HD[e][x][:] = sum_c V[e][c][x] * WdT[c][:]     (vector_down + vector_down_frames, gcpnet.py:420,426)
template <int TE, int NT, int COLS>
GCP_HDN void gcp2_vec_down_impl(const GcpOp& op, const TileBufs& b) {
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < 3 * TE; item += NT) {
    const int x = item / TE, e = item - x * TE;
    float acc[COLS];
#pragma unroll
    for (int k = 0; k < COLS; ++k) acc[k] = 0.f;
    const float* vp = b.V + e * b.ldv + x;
    for (int c = 0; c < op.vi; ++c) {
      const float v = vp[3 * c];
      const float* w = b.WS + c * COLS;
#pragma unroll
      for (int k4 = 0; k4 < COLS; k4 += 4) {
        const float4 wv = ld4(w + k4);
        acc[k4 + 0] = fmaf(v, wv.x, acc[k4 + 0]);
        acc[k4 + 1] = fmaf(v, wv.y, acc[k4 + 1]);
        acc[k4 + 2] = fmaf(v, wv.z, acc[k4 + 2]);
        acc[k4 + 3] = fmaf(v, wv.w, acc[k4 + 3]);
      }
    }
    float* hp = b.HD + e * b.ldhd + x * COLS;
#pragma unroll
    for (int k4 = 0; k4 < COLS; k4 += 4) st4(hp + k4, make_float4(acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]));
  }
  GCP_PHASE_END
}
template <int TE, int NT>
GCP_HDN void gcp2_vec_down(const GcpOp& op, const TileBufs& b) {
  switch (hd_cols(op.hd)) {
    case 8: gcp2_vec_down_impl<TE, NT, 8>(op, b); break;
    case 12: gcp2_vec_down_impl<TE, NT, 12>(op, b); break;
    case 16: gcp2_vec_down_impl<TE, NT, 16>(op, b); break;
    default: gcp2_vec_down_impl<TE, NT, 20>(op, b); break;  // hd <= 16 (checked on the host)
  }
}

// norms (safe_norm over xyz, gcpnet.py:421) and frame scalars (scalarize, comp:302-312) -> Z[:, si:K)
template <int TE, int NT>
GCP_HDN void gcp2_norm_scalarize(const GcpOp& op, const TileBufs& b, int e3) {
  const int cols = hd_cols(op.hd), hdp = cols - 4;
  const int nq = op.hd + 9;
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * nq; item += NT) {
    const int j = item / TE, e = item - j * TE;
    const float* hp = b.HD + e * b.ldhd;
    float val;
    if (j < op.hd) {
      const float a = hp[j], bb = hp[cols + j], c = hp[2 * cols + j];
      val = sqrtf(fmaf(a, a, fmaf(bb, bb, c * c)) + SAFE_NORM_EPS) + SAFE_NORM_EPS;
    } else {
      const int cc = (j - op.hd) / 3, a = (j - op.hd) - 3 * cc;  // q[3*cc + a]
      const float* f = b.F + e * LDF + 3 * a;
      val = f[0] * hp[hdp + cc];
      val = fmaf(f[1], hp[cols + hdp + cc], val);
      val = fmaf(f[2], hp[2 * cols + hdp + cc], val);
      if (e3 && a == 1) val = fabsf(val);  // comp:305-309
    }
    b.Z[e * b.ldz + op.si + j] = val;
  }
  {  // zero the float4 padding columns [K, round_up(K,4)) (K differs between GCPs sharing Z)
    const int K = op.si + nq, padn = round_up(K, 4) - K;
    for (int item = tid; item < TE * padn; item += NT) {
      const int e = item / padn;
      b.Z[e * b.ldz + K + (item - e * padn)] = 0.f;
    }
  }
  GCP_PHASE_END
}

// Full GCP2 forward up to (T, SG): caller then reads act_s(T) and gcp2_vec_out().
// OGm/NRm: micro-tile grid of the scalar_out GEMM; the gate GEMM uses (OGg, 1).
template <int TE, int NT, int OGm, int NRm, int OGg>
GCP_HDN void gcp2_fwd_tile(const GcpOp& op, const TileBufs& b, int e3, float slope) {
  gcp2_stage_small<NT>(op, b.WS);
  gcp2_vec_down<TE, NT>(op, b);
  gcp2_norm_scalarize<TE, NT>(op, b, e3);
  float* T = b.T; const int ldt = b.ldt;
  tile_gemm_nmajor<TE, NT, OGm, NRm>(b.Z, b.ldz, gcp_k(op), op.Ws, op.so, op.bs, b.WC, b.wc_cap, T, ldt, XIdentity(),
                                     [=](int e, int n, float v) { T[e * ldt + n] = v; });
  if (op.vo > 0) {
    float* SG = b.SG; const int ldsg = b.ldsg;
    // gate reads the PRE-activation scalars through act_v (gcpnet.py:386)
    tile_gemm_nmajor<TE, NT, OGg, 1>(T, ldt, op.so, op.Wg, op.vo, op.bg, b.WC, b.wc_cap, SG, ldsg, XAct{op.act_v, slope},
                                     [=](int e, int n, float v) { SG[e * ldsg + n] = sigmoidf_(v); });
  }
}

// Ungated vector output U[e][o][x] = sum_k H[e][x][k] * Wu[o][k] (+ V_in[e][o][x] if vector_residual)
GCP_HD float gcp2_vec_up(const GcpOp& op, const TileBufs& b, int e, int o, int x) {
  const int cols = hd_cols(op.hd), hdp = cols - 4;
  const float* hp = b.HD + e * b.ldhd + x * cols;
  const float* wu = b.WS + op.vi * cols + o * hdp;
  float u = 0.f;
  for (int k4 = 0; k4 < hdp; k4 += 4) {
    const float4 h = ld4(hp + k4), w = ld4(wu + k4);
    u = fmaf(h.x, w.x, u); u = fmaf(h.y, w.y, u); u = fmaf(h.z, w.z, u); u = fmaf(h.w, w.w, u);
  }
  if (op.vres) u += b.V[e * b.ldv + 3 * o + x];
  return u;
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
};

// Backward of one GCP2 on a tile.  On entry: b.Z[:, :si], b.V, b.F hold the forward inputs, b.T the
// saved pre-activations, b.SG the saved gates, g.GS / g.GV the output cotangents (rows >= nrows zero).
// The routine recomputes HD, norms and frame scalars, then produces
//   * scalar input cotangent: emitted column by column through `emit_s(e, i, value)` (i < si)
//   * vector input cotangent: emitted through `emit_v(e, c3, value)` (c3 < 3*vi)
//   * weight-gradient partials into prow[op.o_*]  (accumulate: += instead of =)
// b.T is overwritten with the cotangent of the pre-activation.
template <int TE, int NT, int OGm, int NRm, int OGd, int NRd, class EmitS, class EmitV>
GCP_HDN void gcp2_bwd_tile(const GcpOp& op, const TileBufs& b, const BwdBufs& g, int e3, float slope,
                           float* prow, bool accumulate, EmitS emit_s, EmitV emit_v) {
  const int cols = hd_cols(op.hd), hdp = cols - 4;
  const int K = gcp_k(op);
  gcp2_stage_small<NT>(op, b.WS);
  gcp2_vec_down<TE, NT>(op, b);
  gcp2_norm_scalarize<TE, NT>(op, b, e3);
  float* T = b.T; const int ldt = b.ldt;
  if (op.vo > 0) {
    // gate backward: gU = gV' * sg ; gsig = sum_x gV' * U ; gg = gsig * sg * (1 - sg)
    GCP_PHASE_BEGIN(NT)
    for (int item = tid; item < TE * op.vo; item += NT) {
      const int e = item / op.vo, o = item - e * op.vo;
      const float sg = b.SG[e * b.ldsg + o];
      float gsig = 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float gv = g.GV[e * g.ldgv + 3 * o + x];
        gsig = fmaf(gv, gcp2_vec_up(op, b, e, o, x), gsig);
        g.GU[e * g.ldgu + 3 * o + x] = gv * sg;
      }
      g.GG[e * g.ldgg + o] = gsig * sg * (1.f - sg);
    }
    // zero the float4 padding columns of GG so the wgrad's 4-wide reads see zeros
    for (int item = tid; item < TE * (round_up(op.vo, 4) - op.vo); item += NT) {
      const int padn = round_up(op.vo, 4) - op.vo;
      const int e = item / padn, o = op.vo + item - e * padn;
      g.GG[e * g.ldgg + o] = 0.f;
    }
    GCP_PHASE_END
    // vector_out_scale weight gradient (input of that Linear is act_v(T), T still pre-activation here)
    tile_wgrad<TE, NT>(g.GG, g.ldgg, op.vo, T, ldt, op.so, prow + op.o_Wg, prow + op.o_bg, accumulate, XAct{op.act_v, slope});
    // vector_up weight gradient: gWu[o][k] = sum_{e,x} gU[e][o][x] * H[e][x][k]
    GCP_PHASE_BEGIN(NT)
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
  }
  // cotangent of the pre-activation: gT = gS' * act_s'(T) + act_v'(T) * (Wg^T gg)
  {
    const int as = op.act_s, av = op.act_v;
    float* GS = g.GS; const int ldgs = g.ldgs;
    if (op.vo > 0) {
      // T <- gT in place through the small k-major GEMM (K = vo)
      // (each thread reads T[e][n] only for the (e, n) it then overwrites)
      const int vo4 = round_up(op.vo, 4);
      (void)vo4;
      tile_gemm_kmajor<TE, NT, OGd, NRd>(g.GG, g.ldgg, op.vo, op.Wg, op.so, b.WC, b.wc_cap, nullptr, 0,
                                         [=](int e, int n, float v) {
                                           const float t = T[e * ldt + n];
                                           T[e * ldt + n] = fmaf(GS[e * ldgs + n], act_grad(as, t, slope), v * act_grad(av, t, slope));
                                         });
    } else {
      GCP_PHASE_BEGIN(NT)
      for (int item = tid; item < TE * op.so; item += NT) {
        const int e = item / op.so, n = item - e * op.so;
        const float t = T[e * ldt + n];
        T[e * ldt + n] = GS[e * ldgs + n] * act_grad(as, t, slope);
      }
      GCP_PHASE_END
    }
    // zero T's float4 padding columns for the wgrad reads
    GCP_PHASE_BEGIN(NT)
    const int padn = round_up(op.so, 4) - op.so;
    for (int item = tid; item < TE * padn; item += NT) {
      const int e = item / padn, n = op.so + item - e * padn;
      T[e * ldt + n] = 0.f;
    }
    GCP_PHASE_END
  }
  // scalar_out weight gradient
  tile_wgrad<TE, NT>(T, ldt, op.so, b.Z, b.ldz, K, prow + op.o_Ws, prow + op.o_bs, accumulate, XIdentity());
  // data gradient through scalar_out: gz = gT * Ws ; columns < si go to the caller, the rest to GNQ
  {
    float* GNQ = g.GNQ; const int ldnq = g.ldnq; const int si = op.si;
    tile_gemm_kmajor<TE, NT, OGd, NRd>(T, ldt, op.so, op.Ws, K, b.WC, b.wc_cap, nullptr, 0,
                                       [=](int e, int n, float v) {
                                         if (n < si) emit_s(e, n, v);
                                         else GNQ[e * ldnq + (n - si)] = v;
                                       });
  }
  // gHD: norms, vector_up and frame scalars back to the hidden vector channels
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < 3 * TE; item += NT) {
    const int x = item / TE, e = item - x * TE;
    const float* hp = b.HD + e * b.ldhd;
    float* ghp = g.GHD + e * g.ldghd + x * cols;
    const float* gnq = g.GNQ + e * g.ldnq;
    for (int k = 0; k < hdp; ++k) {
      float acc = 0.f;
      if (k < op.hd) {
        // n = sqrt(sum_x H^2 + eps) + eps  ->  dn/dH[x] = H[x] / (n - eps)
        const float a = hp[k], bb = hp[cols + k], c = hp[2 * cols + k];
        const float root = sqrtf(fmaf(a, a, fmaf(bb, bb, c * c)) + SAFE_NORM_EPS);
        acc = gnq[k] * hp[x * cols + k] / root;
        for (int o = 0; o < op.vo; ++o)
          acc = fmaf(g.GU[e * g.ldgu + 3 * o + x], b.WS[op.vi * cols + o * hdp + k], acc);
      }
      ghp[k] = acc;
    }
    // frame scalars: q[3cc+a] = sum_x F[a][x] D[x][cc]  (|.| on a==1 when e3)
    for (int cc = 0; cc < 3; ++cc) {
      float acc = 0.f;
      for (int a = 0; a < 3; ++a) {
        float gq = gnq[op.hd + 3 * cc + a];
        if (e3 && a == 1) {
          const float* f = b.F + e * LDF + 3;
          float q = f[0] * hp[hdp + cc];
          q = fmaf(f[1], hp[cols + hdp + cc], q);
          q = fmaf(f[2], hp[2 * cols + hdp + cc], q);
          gq = q > 0.f ? gq : (q < 0.f ? -gq : 0.f);
        }
        acc = fmaf(b.F[e * LDF + 3 * a + x], gq, acc);
      }
      ghp[hdp + cc] = acc;
    }
    ghp[hdp + 3] = 0.f;
  }
  GCP_PHASE_END
  // vector_down / vector_down_frames weight gradients: gWdT[c][k] = sum_{e,x} gHD[e][x][k] * V[e][c][x]
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < op.vi * (op.hd + 3); item += NT) {
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
  GCP_PHASE_END
  // vector input cotangent: gV[e][c][x] = sum_k gHD[e][x][k] * WdT[c][k] (+ gU[e][c][x] if vector_residual)
  GCP_PHASE_BEGIN(NT)
  for (int item = tid; item < TE * op.vi; item += NT) {
    const int e = item / op.vi, c = item - e * op.vi;
    const float* w = b.WS + c * cols;
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
  GCP_PHASE_END
}

// cooperative row copy helpers ----------------------------------------------------------------
// dst (smem, row stride ldd) <- rows of a global matrix with `len` contiguous floats per row,
// row index given by idx(e) (< 0: zero fill).
template <int TE, int NT, class RowIdx>
GCP_HD void tile_load_rows(float* dst, int ldd, const float* src, int len, RowIdx rowidx, int tid) {
  for (int item = tid; item < TE * len; item += NT) {
    const int e = item / len, f = item - e * len;
    const long long r = rowidx(e);
    dst[e * ldd + f] = r >= 0 ? GCP_LDG(src + (size_t)r * len + f) : 0.f;
  }
}
// global rows [row0 + e] (dense, len floats) <- smem rows, for e < nrows
template <int TE, int NT>
GCP_HD void tile_store_rows(float* dst, long long row0, int len, const float* src, int lds, int nrows, int tid) {
  for (int item = tid; item < nrows * len; item += NT) {
    const int e = item / len, f = item - e * len;
    dst[(size_t)(row0 + e) * len + f] = src[e * lds + f];
  }
}

}  // namespace gcp
