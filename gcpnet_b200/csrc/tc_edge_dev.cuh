// tc_edge_dev.cuh -- device code of the tensor-core edge kernels (see tc_edge.cuh for the design notes and
// the parameter structures).  Included only by tc_api.cu.
#pragma once
#include "gcp_tile.cuh"
#include "umma.cuh"
#include "tc_edge.cuh"
#include "tc_setup.h"

// Development aid: clock64 stamps of CTA 0's stages (scripts/tc_bwd_stamps.py).  Compiled out of the shipped library
// (they cost 4-8 % of the executed instructions); build with GCPNET_NVCC_FLAGS=-DGCP_STAMPS=1 to get them.
#ifndef GCP_STAMPS
#define GCP_STAMPS 0
#endif

namespace gcp {
namespace tc {

using namespace ::gcp::umma;

// ------------------------------------------------------------------------------------------------
// pack kernel: nn.Linear weights -> B tiles in slab layout, hi / lo split, composed rows
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_pack_kernel(const __grid_constant__ TcPackProg prog) {
  const TcPackItem& it = prog.it[blockIdx.x];
  float* blob = prog.blob;
  if (it.kind == PK_BIAS) {  // [bs (R) | composed gate bias (16)]
    for (int i = blockIdx.y * 256 + threadIdx.x; i < it.R + 16; i += 256 * gridDim.y) {
      float val = 0.f;
      if (i < it.R) { if (i < it.nreal) val = __ldg(it.w + i); }
      else if (i - it.R < it.vo) {
        const int o = i - it.R;
        val = __ldg(it.w3 + o);
        for (int j = 0; j < it.nreal; ++j) val = fmaf(__ldg(it.w2 + (size_t)o * it.nreal + j), __ldg(it.w + j), val);
      }
      blob[it.dst_hi + i] = val;
    }
    return;
  }
  const int total = it.R * it.C;
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < total; idx += 256 * gridDim.y) {
    const int n = idx % it.R, kk = idx / it.R;
    float val = 0.f;
    for (int rg = 0; rg < it.nrange; ++rg) {
      if (kk >= it.tc0[rg] && kk < it.tc0[rg] + it.len[rg]) {
        const int col = it.rc0[rg] + kk - it.tc0[rg];
        if (it.kind == PK_SCALAR) {
          // rows [0, so): scalar_out.weight; rows [sop, sop + vo): vector_out_scale.weight . scalar_out.weight
          if (n < it.nreal) val = __ldg(it.w + (size_t)n * it.ldw + col);
          else if (n >= it.R - 16 && n - (it.R - 16) < it.vo) {
            const int o = n - (it.R - 16);
            for (int j = 0; j < it.nreal; ++j) val = fmaf(__ldg(it.w2 + (size_t)o * it.nreal + j), __ldg(it.w + (size_t)j * it.ldw + col), val);
          }
        } else {  // PK_VECTOR: rows [0, hd) vector_down; 13..15 vector_down_frames; [16, 16 + vo) vector_up . vector_down
          if (n < it.hd) val = __ldg(it.w + (size_t)n * it.ldw + col);
          else if (n >= DCOL && n < DCOL + 3) val = __ldg(it.w2 + (size_t)(n - DCOL) * it.ldw + col);
          else if (n >= UCOL && n - UCOL < it.vo) {
            const int o = n - UCOL;
            for (int j = 0; j < it.hd; ++j) val = fmaf(__ldg(it.w3 + (size_t)o * it.hd + j), __ldg(it.w + (size_t)j * it.ldw + col), val);
          }
        }
      }
    }
    const int off = it.tr ? ((n >> 2) * it.Rt + kk) * 4 + (n & 3) : ((kk >> 2) * it.R + n) * 4 + (kk & 3);
    blob[it.dst_hi + off] = val;
    if (it.dst_lo >= 0) blob[it.dst_lo + off] = umma::tf32_lo(val);
  }
  if (it.tr) {  // zero rows [C, Rt) of the transposed tile
    for (int idx = blockIdx.y * 256 + threadIdx.x; idx < (it.Rt - it.C) * it.R; idx += 256 * gridDim.y) {
      const int n = idx % it.R, kk = it.C + idx / it.R;
      const int off = ((n >> 2) * it.Rt + kk) * 4 + (n & 3);
      blob[it.dst_hi + off] = 0.f;
      if (it.dst_lo >= 0) blob[it.dst_lo + off] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-node pre-products of message GCP 0:  P[i] = [src | dst] x (T-part, g-part),  Q[i] = [src | dst] x 3 planes x 32
// ------------------------------------------------------------------------------------------------
constexpr int PRE_NODES = 4;  // nodes per thread: every weight load feeds four accumulators (the kernel is L1-bandwidth bound)
__global__ void __launch_bounds__(256) tc_node_pre_kernel(const float* __restrict__ h, const float* __restrict__ chi,
                                                          const float* __restrict__ blob, TcNodeTiles nt, int N, int s, int v, int pw,
                                                          float* __restrict__ P, float* __restrict__ Q) {
  const int per = 2 * pw + 192;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int groups = (N + PRE_NODES - 1) / PRE_NODES;
  if (idx >= (long long)groups * per) return;
  const int grp = (int)(idx / per), o = (int)(idx - (long long)grp * per);
  const int i0 = grp * PRE_NODES;
  int row[PRE_NODES];
#pragma unroll
  for (int r = 0; r < PRE_NODES; ++r) row[r] = i0 + r < N ? i0 + r : N - 1;  // clamped loads; stores are guarded
  float acc[PRE_NODES];
#pragma unroll
  for (int r = 0; r < PRE_NODES; ++r) acc[r] = 0.f;
  if (o < 2 * pw) {
    const int side = o / pw, c = o - side * pw;
    const float* B = blob + (side ? nt.pd : nt.ps);  // [pw][s] slab, pitch pw
    // s % 16 == 0 on this path: 16-byte loads, four columns per step; ONE accumulator per output in column order (the
    // summation order is part of the pinned numerics: ReLU units at ~0 flip with it, tests/test_gpu_parity.py)
#pragma unroll 2
    for (int j4 = 0; j4 < (s >> 2); ++j4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(B + (j4 * pw + c) * 4));
#pragma unroll
      for (int r = 0; r < PRE_NODES; ++r) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(h + (size_t)row[r] * s) + j4);
        acc[r] = fmaf(b.x, x.x, acc[r]); acc[r] = fmaf(b.y, x.y, acc[r]); acc[r] = fmaf(b.z, x.z, acc[r]); acc[r] = fmaf(b.w, x.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < PRE_NODES; ++r)
      if (i0 + r < N) P[(size_t)(i0 + r) * 2 * pw + o] = acc[r];
  } else {
    const int o2 = o - 2 * pw;
    const int side = o2 / 96, x = (o2 - side * 96) / 32, c = o2 & 31;
    const float* B = blob + (side ? nt.qd : nt.qs);  // [32][v8] slab, pitch 32
    int ch = 0;
    for (; ch + 4 <= v; ch += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(B + ((ch >> 2) * 32 + c) * 4));
#pragma unroll
      for (int r = 0; r < PRE_NODES; ++r) {
        const float* cp = chi + (size_t)row[r] * 3 * v + x;
        acc[r] = fmaf(b.x, __ldg(cp + 3 * ch), acc[r]); acc[r] = fmaf(b.y, __ldg(cp + 3 * ch + 3), acc[r]);
        acc[r] = fmaf(b.z, __ldg(cp + 3 * ch + 6), acc[r]); acc[r] = fmaf(b.w, __ldg(cp + 3 * ch + 9), acc[r]);
      }
    }
    for (; ch < v; ++ch) {
      const float b = __ldg(B + ((ch >> 2) * 32 + c) * 4 + (ch & 3));
#pragma unroll
      for (int r = 0; r < PRE_NODES; ++r) acc[r] = fmaf(b, __ldg(chi + (size_t)row[r] * 3 * v + x + 3 * ch), acc[r]);
    }
#pragma unroll
    for (int r = 0; r < PRE_NODES; ++r)
      if (i0 + r < N) Q[(size_t)(i0 + r) * 192 + o2] = acc[r];
  }
}

// ------------------------------------------------------------------------------------------------
// rings, bulk stores
// ------------------------------------------------------------------------------------------------
struct Ring {           // CTA-uniform state; only lane 0 of warp 0 issues copies
  float* slots; unsigned long long* bar; const RingDesc* d; const float* blob;
  int head, issued, total;
};
__device__ __forceinline__ void ring_fill(Ring& r) {  // converged warp 0: top the ring up (slots < head are free)
  const int head = uniform(r.head);
  int issued = uniform(r.issued);
  while (issued < r.total && issued < head + r.d->nslot) {
    const TcChunk ck = r.d->c[issued % r.d->n];
    const int slot = issued % r.d->nslot;
    if (elect_one()) {
      const uint32_t bar = smem_addr(&r.bar[slot]);
      const uint32_t bytes = (uint32_t)ck.floats * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(r.slots + (size_t)slot * r.d->slot_floats)), "l"(r.blob + ck.off), "r"(bytes), "r"(bar) : "memory");
    }
    ++issued;
  }
  r.issued = issued;
}
__device__ __forceinline__ const float* ring_wait(const Ring& r, int pos) {
  const int slot = pos % r.d->nslot;
  mbar_wait(&r.bar[slot], (uint32_t)((pos / r.d->nslot) & 1));
  return r.slots + (size_t)slot * r.d->slot_floats;
}
__device__ __forceinline__ void bulk_store(float* gdst, const float* ssrc, int floats) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"((uint32_t)floats * 4u) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Tile images in global memory are COMPACT: of every 4-column slab (RP rows x 16 bytes in shared memory) only the first
// `rows` rows -- the rows the tile really has -- are stored / loaded (small graphs run 32..48-row tiles: 3 x less traffic
// than whole slabs).  One elected thread issues `nslab` bulk copies; stores form one bulk group.
__device__ __forceinline__ void image_store(float* gdst, const float* ssrc, int nslab, int rows) {
  const uint32_t bytes = (uint32_t)rows * 16u;
  for (int j = 0; j < nslab; ++j)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst + (size_t)j * rows * 4), "r"(smem_addr(ssrc + (size_t)j * SLAB)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void image_load(float* sdst, const float* gsrc, int nslab, int rows, unsigned long long* bar) {
  const uint32_t b = smem_addr(bar), bytes = (uint32_t)rows * 16u;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * (uint32_t)nslab) : "memory");
  for (int j = 0; j < nslab; ++j)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(sdst + (size_t)j * SLAB)), "l"(gsrc + (size_t)j * rows * 4), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// row-owner stores: hi part to the slab tile, lo part (x - tf32(x); the tensor core truncates it) to TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lo_part(float x) { return x - __uint_as_float(__float_as_uint(x) & TF32_MASK); }
__device__ __forceinline__ void put4(float* tile, uint32_t tlo, int r, int c, float a, float b, float c2, float d) {
  *reinterpret_cast<float4*>(tile + slab_off(RP, r, c)) = make_float4(a, b, c2, d);
  tmem_st4(tlo + (uint32_t)c, lo_part(a), lo_part(b), lo_part(c2), lo_part(d));
}
__device__ __forceinline__ float4 get4(const float* tile, int r, int c) {
  return *reinterpret_cast<const float4*>(tile + slab_off(RP, r, c));
}

struct Who {      // where a thread sits
  int tid, r, part;
  uint32_t tl;    // TMEM base + this warp's lane window
};

// operands written by all threads -> visible to the tensor core; then CTA barrier
__device__ __forceinline__ void publish_and_sync() {
  wait_st();
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
}

// 3xTF32 product of an A tile (hi in shared memory, lo in TMEM) with a B tile (hi / lo in shared memory):
// D[128][N] (+)= A[:, 0:kc] . B[N][kc]^T.  Called by a whole warp with warp-uniform arguments (descriptor
// arithmetic stays on the uniform datapath); one elected lane issues.
__device__ __forceinline__ void mma3(uint32_t d_tmem, const float* a_hi, uint32_t a_lo, const float* b_hi, const float* b_lo,
                                     int b_rows, int kc, uint32_t idesc, bool& acc) {
  const uint64_t ad0 = desc_kmajor(a_hi, RP, 0, 0);
  const uint64_t bh0 = make_desc(smem_addr(b_hi), (uint32_t)b_rows * 16u, 128u);
  const uint64_t bl0 = make_desc(smem_addr(b_lo), (uint32_t)b_rows * 16u, 128u);
  const uint32_t astep = 2u * SLAB * 4u, bstep = 2u * (uint32_t)b_rows * 16u;
  for (int ks = 0; ks < (kc >> 3); ++ks) {
    const uint64_t ad = desc_advance(ad0, ks * astep), bh = desc_advance(bh0, ks * bstep), bl = desc_advance(bl0, ks * bstep);
    if (elect_one()) {
      mma_tf32(d_tmem, ad, bh, idesc, acc);
      mma_tf32(d_tmem, ad, bl, idesc, true);
      mma_tf32_ts(d_tmem, a_lo + (uint32_t)(8 * ks), bh, idesc, true);
    }
    acc = true;
  }
}

// The same for three planes at once (independent accumulators, plane strides given): consecutive MMAs of one
// accumulator chain are issued three instructions apart.
__device__ __forceinline__ void mma3_planes(uint32_t d_tmem, uint32_t d_stride, const float* a_hi, int a_stride, uint32_t a_lo, uint32_t a_lo_stride,
                                            const float* b_hi, const float* b_lo, int b_rows, int kc, uint32_t idesc, bool& acc) {
  const uint64_t ad0 = desc_kmajor(a_hi, RP, 0, 0);
  const uint64_t bh0 = make_desc(smem_addr(b_hi), (uint32_t)b_rows * 16u, 128u);
  const uint64_t bl0 = make_desc(smem_addr(b_lo), (uint32_t)b_rows * 16u, 128u);
  const uint32_t astep = 2u * SLAB * 4u, bstep = 2u * (uint32_t)b_rows * 16u, pstep = (uint32_t)a_stride * 4u;
  for (int ks = 0; ks < (kc >> 3); ++ks) {
    const uint64_t bh = desc_advance(bh0, ks * bstep), bl = desc_advance(bl0, ks * bstep);
    if (elect_one()) {
#pragma unroll
      for (int x = 0; x < 3; ++x) mma_tf32(d_tmem + x * d_stride, desc_advance(ad0, ks * astep + x * pstep), bh, idesc, acc);
#pragma unroll
      for (int x = 0; x < 3; ++x) mma_tf32(d_tmem + x * d_stride, desc_advance(ad0, ks * astep + x * pstep), bl, idesc, true);
#pragma unroll
      for (int x = 0; x < 3; ++x) mma_tf32_ts(d_tmem + x * d_stride, a_lo + x * a_lo_stride + (uint32_t)(8 * ks), bh, idesc, true);
    }
    acc = true;
  }
}

// ---- epilogue A: vector batch accumulator -> norms + frame scalars into the Z-tile tail ------------------------------
// VACC: 3 planes x 32 columns; [0, hd) hidden channels H, [13, 16) frame-down vectors D, [16, 32) ungated outputs U.
// Tail columns (from zc0): nslot norm slots | 9 frame scalars | 3 zeros.  Items: norm groups 0..2, frame scalars.
// GCP 0: the accumulator only holds the xi part; the chi_row / chi_col parts come from the per-node products qs / qd.
template <int CS>
__device__ __forceinline__ void epilogue_a(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w, const float* qs, const float* qd,
                                           bool pad_one = false) {
  float* Z = sm + p.ZBUF;
  const uint32_t zlo = w.tl + (uint32_t)p.ZLO;
#pragma unroll
  for (int item = 0; item < 4; ++item) {
    if ((item % CS) != w.part) continue;  // warp-uniform
    if (item < 3) {
      // norms n_j = sqrt(sum_x H_xj^2 + eps) + eps  (safe_norm, comp/__init__.py:381-392), channels 4*item .. +3
      float nv[4] = {0.f, 0.f, 0.f, 0.f};
      if (4 * item < g.hd) {
        float h[3][4];
#pragma unroll
        for (int x = 0; x < 3; ++x) tmem_ld4(w.tl + (uint32_t)(p.VACC + VN * x + 4 * item), h[x]);
        wait_ld();
        if (qs != nullptr) {
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(qs + 32 * x + 4 * item)), b = __ldg(reinterpret_cast<const float4*>(qd + 32 * x + 4 * item));
            h[x][0] += a.x + b.x; h[x][1] += a.y + b.y; h[x][2] += a.z + b.z; h[x][3] += a.w + b.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (4 * item + i < g.hd)
            nv[i] = sqrtf(fmaf(h[0][i], h[0][i], fmaf(h[1][i], h[1][i], h[2][i] * h[2][i])) + SAFE_NORM_EPS) + SAFE_NORM_EPS;
      }
      if (4 * item < g.nslot) put4(Z, zlo, w.r, g.zc0 + 4 * item, nv[0], nv[1], nv[2], nv[3]);
    } else {
      // frame scalars q[3c + a] = sum_x F[a][x] * D[x][c]  (scalarize, comp/__init__.py:302-312)
      float d[3][4];
#pragma unroll
      for (int x = 0; x < 3; ++x) tmem_ld4(w.tl + (uint32_t)(p.VACC + VN * x + 12), d[x]);  // columns 12..15: D at 13..15
      wait_ld();
      if (qs != nullptr) {
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(qs + 32 * x + 12)), b = __ldg(reinterpret_cast<const float4*>(qd + 32 * x + 12));
          d[x][0] += a.x + b.x; d[x][1] += a.y + b.y; d[x][2] += a.z + b.z; d[x][3] += a.w + b.w;
        }
      }
      const float* F = sm + p.FBUF + w.r * 9;
      float q[12];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int cc = t / 3, a = t - 3 * cc;
        float v = F[3 * a] * d[0][1 + cc];
        v = fmaf(F[3 * a + 1], d[1][1 + cc], v);
        v = fmaf(F[3 * a + 2], d[2][1 + cc], v);
        if (p.e3 && a == 1) v = fabsf(v);
        q[t] = v;
      }
      q[9] = q[10] = 0.f;
      q[11] = pad_one ? 1.f : 0.f;  // backward: a column of ones turns the weight-gradient GEMM's last column into the bias gradient
#pragma unroll
      for (int j = 0; j < 3; ++j) put4(Z, zlo, w.r, g.zc0 + g.nslot + 4 * j, q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
    }
  }
}

// ---- epilogue B: scalar batch accumulator [T | g] + vector batch U -> new scalar state (Z tile), new vector state (V tile)
template <int CS>
__device__ __forceinline__ void epilogue_b(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w, const float* smc,
                                           bool add, bool last, long long q, bool live, const float* ps, const float* pd,
                                           const float* qs, const float* qd) {
  float* Z = sm + p.ZBUF;
  float* V = sm + p.VBUF;
  const float* bs = smc + g.o_b;
  const float* bg = smc + g.o_b + g.sop;
  // scalars: S' = (S +) act_s(T + b)   (gcpnet.py:441,465; residual stack :920-924)
  for (int cg = w.part; 16 * cg < g.so; cg += CS) {
    float t[16];
    tmem_ld16(w.tl + (uint32_t)(p.TACC + 16 * cg), t);
    wait_ld();
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const int c = 16 * cg + 4 * i4;
      if (c < g.so) {  // so % 4 == 0
        const float4 old = add ? get4(Z, w.r, c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b4 = *reinterpret_cast<const float4*>(bs + c);
        float sv[4] = {t[4 * i4] + b4.x, t[4 * i4 + 1] + b4.y, t[4 * i4 + 2] + b4.z, t[4 * i4 + 3] + b4.w};
        if (ps != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(ps + c)), b = __ldg(reinterpret_cast<const float4*>(pd + c));
          sv[0] += a.x + b.x; sv[1] += a.y + b.y; sv[2] += a.z + b.z; sv[3] += a.w + b.w;
        }
        if (g.act_s == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sv[i] = fmaxf(sv[i], 0.f);
        } else if (g.act_s != ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sv[i] = act_fwd(g.act_s, sv[i], p.slope);
        }
        sv[0] += old.x; sv[1] += old.y; sv[2] += old.z; sv[3] += old.w;
        put4(Z, w.tl + (uint32_t)p.ZLO, w.r, c, sv[0], sv[1], sv[2], sv[3]);
      }
    }
  }
  // vectors: V' = (V +) (U (+ V_in)) * sigmoid(g + b')   (gcpnet.py:364-367,385-387)
  for (int gi = w.part; gi < (PW >> 2); gi += CS) {
    float gt[4], u[3][4];
    tmem_ld4(w.tl + (uint32_t)(p.TACC + g.sop + 4 * gi), gt);
#pragma unroll
    for (int x = 0; x < 3; ++x) tmem_ld4(w.tl + (uint32_t)(p.VACC + VN * x + UCOL + 4 * gi), u[x]);
    wait_ld();
    if (ps != nullptr) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ps + g.sop + 4 * gi)), b = __ldg(reinterpret_cast<const float4*>(pd + g.sop + 4 * gi));
      gt[0] += a.x + b.x; gt[1] += a.y + b.y; gt[2] += a.z + b.z; gt[3] += a.w + b.w;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(qs + 32 * x + UCOL + 4 * gi)), d = __ldg(reinterpret_cast<const float4*>(qd + 32 * x + UCOL + 4 * gi));
        u[x][0] += c.x + d.x; u[x][1] += c.y + d.y; u[x][2] += c.z + d.z; u[x][3] += c.w + d.w;
      }
    }
    float4 old[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) old[x] = (add || g.vres) ? get4(V + x * PLANE, w.r, 4 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    float nv[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = 4 * gi + i;
      const float sg = o < g.vo ? sigmoidf_(gt[i] + bg[o]) : 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float ov = i == 0 ? old[x].x : (i == 1 ? old[x].y : (i == 2 ? old[x].z : old[x].w));
        float uu = u[x][i];
        if (g.vres) uu += ov;
        nv[x][i] = (add ? ov : 0.f) + uu * sg;
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) put4(V + x * PLANE, w.tl + (uint32_t)(p.VLO + PW * x), w.r, 4 * gi, nv[x][0], nv[x][1], nv[x][2], nv[x][3]);
  }
}

// ---- aggregate (gcpnet.py:938-947) straight out of the Z / V tiles: the threads of a destination's rows in the tile share
//      its 4-column groups; each sums its group over the segment's rows, starting at its own row and wrapping around (a fixed
//      order per output element; the threads of a warp read different rows -> no bank conflicts).  Complete segments go to
//      agg[dst], the tile's open ends to its two carry rows (segment_total, gcp_tile.cuh).  The per-edge messages never go
//      to HBM.
template <int CS>
__device__ __forceinline__ void segment_sums(const TcEdgeParams& p, const float* Z, const float* V, const Who& w, int tile,
                                             int dst, bool live) {
  if (!live) return;
  const long long row0 = (long long)tile * p.rows;
  const long long a = p.dst_ptr[dst], b = p.dst_ptr[dst + 1];
  const int r0 = a > row0 ? (int)(a - row0) : 0;                                  // rows [r0, r1) of the tile belong to dst
  const int r1 = (int)((b < row0 + p.rows ? b : row0 + p.rows) - row0);
  const int n = r1 - r0, j = w.r - r0, W = p.s + 3 * p.v;
  float* out = (a >= row0 && b <= row0 + p.rows) ? p.agg + (size_t)dst * W
                                                 : p.agg + (size_t)p.N * W + ((size_t)tile * 2 + (a < row0 ? 0 : 1)) * W;
  const int gs = p.s >> 2, gv = p.v >> 2;  // 4-column groups of the scalars / of the vector channels
  for (int g = j * CS + w.part; g < gs + gv; g += n * CS) {
    if (g < gs) {
      float4 acc = get4(Z, w.r, 4 * g);
      for (int i = 1, r = w.r + 1; i < n; ++i, ++r) {
        if (r == r1) r = r0;
        const float4 t = get4(Z, r, 4 * g);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      *reinterpret_cast<float4*>(out + 4 * g) = acc;
    } else {
      const int gi = g - gs;
      float4 acc[3];
#pragma unroll
      for (int x = 0; x < 3; ++x) acc[x] = get4(V + x * PLANE, w.r, 4 * gi);
      for (int i = 1, r = w.r + 1; i < n; ++i, ++r) {
        if (r == r1) r = r0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float4 t = get4(V + x * PLANE, r, 4 * gi);
          acc[x].x += t.x; acc[x].y += t.y; acc[x].z += t.z; acc[x].w += t.w;
        }
      }
      float* mp = out + p.s + 12 * gi;  // [channel][xyz], xyz fastest; v % 4 == 0
      *reinterpret_cast<float4*>(mp) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[0].y);
      *reinterpret_cast<float4*>(mp + 4) = make_float4(acc[1].y, acc[2].y, acc[0].z, acc[1].z);
      *reinterpret_cast<float4*>(mp + 8) = make_float4(acc[2].z, acc[0].w, acc[1].w, acc[2].w);
    }
  }
}

// gather `len` contiguous floats (len % 4 == 0) of a global row into columns [0, len) of a slab tile, zero up to `cols`
template <int CS>
__device__ __forceinline__ void gather_row(float* tile, uint32_t tlo, const Who& w, const float* src, int len, int cols, bool live) {
  for (int c = 4 * w.part; c < cols; c += 4 * CS) {
    const float4 v = (live && c < len) ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    put4(tile, tlo, w.r, c, v.x, v.y, v.z, v.w);
  }
}
// gather a [channels][xyz] global row (nch % 4 == 0) into three 16-column planes, zero up to `cols` columns
template <int CS>
__device__ __forceinline__ void gather_planes(float* planes, uint32_t tlo, const Who& w, const float* src, int nch, int cols, bool live) {
  for (int gi = w.part; 4 * gi < cols; gi += CS) {
    float f[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) f[j] = 0.f;
    if (live && 4 * gi < nch) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + 12 * gi + 4 * j));
        f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) put4(planes + x * PLANE, tlo + (uint32_t)(PW * x), w.r, 4 * gi, f[x], f[3 + x], f[6 + x], f[9 + x]);
  }
}

// state every tile kernel carries around
struct TileCtx {
  float* sm;
  Who w;
  int uwarp;
  uint32_t tbase;
  unsigned long long* mma_bar;
  uint32_t mma_n;   // commits on mma_bar so far (CTA-uniform)
};
__device__ __forceinline__ void wait_mma(TileCtx& c) {
  mbar_wait(c.mma_bar, c.mma_n & 1u);
  ++c.mma_n;
  __syncwarp();
  fence_after_sync();
}

// Forward of GCP k on the tile: the V tile (and for k > 0 the Z tile's first s columns) hold the inputs; on return they hold
// the outputs.  `issue_extra` runs in the elected lane right after the vector batch is committed (bulk stores).
template <int CS, class Extra>
__device__ __forceinline__ void gcp_forward_tile(const TcEdgeParams& p, int k, TileCtx& c, Ring& rs, Ring& rw, bool add, bool write_state,
                                                 bool last, long long q, bool live, int src, int dst, int orig, Extra issue_extra,
                                                 const float** smc_out, bool pad_one = false) {
  const TcGcp& g = p.g[k];
  float* sm = c.sm;
  const Who& w = c.w;
  float* Z = sm + p.ZBUF; float* V = sm + p.VBUF;
  const uint32_t tbase = c.tbase;
#if GCP_STAMPS
  auto stamp = [&](int i) { if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 0) p.dbg[k * 16 + i] = clock64(); };
#else
  auto stamp = [](int) {};
#endif
  const float* ps = nullptr; const float* pd = nullptr; const float* qs = nullptr; const float* qd = nullptr;
  stamp(0);
  if (k == 0) {
    // edge vectors -> V tile; frames; per-node products of the two endpoints
    gather_planes<CS>(V, w.tl + (uint32_t)p.VLO, w, p.xi + (size_t)orig * 3 * p.ve, p.ve, g.vkc, live);
    if (w.part == CS - 1) {
      float* F = sm + p.FBUF + w.r * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) F[i] = live ? __ldg(p.frames + (size_t)orig * 9 + i) : 0.f;
    }
    ps = p.P + (size_t)src * 2 * p.pw; pd = p.P + (size_t)dst * 2 * p.pw + p.pw;
    qs = p.Q + (size_t)src * 192; qd = p.Q + (size_t)dst * 192 + 96;
  }
  publish_and_sync();
  stamp(1);
  const float* smc = ring_wait(rs, uniform(rs.head));  // small chunk of this GCP (every thread reads the bias from it)
  *smc_out = smc;
  if (c.uwarp == 0) {
    if (elect_one()) bulk_store_wait_read();
    ring_fill(rs);  // every thread is past the previous GCP's last read of its small chunk (barrier above)
    bool acc = false;
    mma3_planes(tbase + (uint32_t)p.VACC, VN, V, PLANE, tbase + (uint32_t)p.VLO, PW, smc + g.o_wv_hi, smc + g.o_wv_lo, VN, g.vkc,
                make_idesc(128, VN, 0, 0), acc);
    if (elect_one()) {
      commit(c.mma_bar);
      issue_extra();
    }
    __syncwarp();
  }
  stamp(2);
  wait_mma(c);
  stamp(3);
  // ---------------- epilogue A (+ GCP 0: edge scalars into the Z tile)
  epilogue_a<CS>(p, g, sm, w, qs, qd, pad_one);
  if (k == 0) gather_row<CS>(Z, w.tl + (uint32_t)p.ZLO, w, p.e + (size_t)orig * p.se, p.se, p.se, live);
  stamp(4);
  publish_and_sync();
  stamp(5);
  // ---------------- scalar batch
  if (c.uwarp == 0) {
    if (elect_one()) bulk_store_wait_read();
    ring_fill(rw);
    const int wh = uniform(rw.head);
    const float* bh = ring_wait(rw, wh);
    const float* bl = ring_wait(rw, wh + 1);
    bool acc = false;
    mma3(tbase + (uint32_t)p.TACC, Z, tbase + (uint32_t)p.ZLO, bh, bl, g.sop + 16, g.kz, make_idesc(128, g.sop + 16, 0, 0), acc);
    if (elect_one()) commit(c.mma_bar);
    __syncwarp();
  }
  stamp(6);
  wait_mma(c);
  stamp(7);
  rw.head += 2;
  if (c.uwarp == 0) ring_fill(rw);  // the two slots just released: prefetch while the epilogue runs
  // ---------------- epilogue B
  if (write_state) epilogue_b<CS>(p, g, sm, w, smc, add, last, q, live, ps, pd, qs, qd);
  stamp(8);
}

template <int CS>
__global__ void __launch_bounds__(128 * CS, 1) tc_edge_fwd_kernel(const __grid_constant__ TcEdgeParams p) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_slot;
  const int ntiles = (p.E + p.rows - 1) / p.rows;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  TileCtx c;
  c.sm = sm;
  c.w.tid = (int)threadIdx.x;
  const int warp = c.w.tid >> 5, lane = c.w.tid & 31;
  c.w.r = 32 * (warp & 3) + lane;
  c.w.part = warp >> 2;
  c.uwarp = uniform(warp);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + p.BARS);
  c.mma_bar = bars;  // [0]; ring barriers follow
  c.mma_n = 0;
  Ring rs{sm + p.RING_S, bars + 1, &p.ring_s, p.blob, 0, 0, mine * p.ring_s.n};
  Ring rw{sm + p.RING_W, bars + 1 + p.ring_s.nslot, &p.ring_w, p.blob, 0, 0, mine * p.ring_w.n};
  if (c.uwarp == 0) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    if (elect_one()) {
      mbar_init(c.mma_bar, 1);
      for (int i = 0; i < p.ring_s.nslot; ++i) mbar_init(&rs.bar[i], 1);
      for (int i = 0; i < p.ring_w.nslot; ++i) mbar_init(&rw.bar[i], 1);
      mbar_fence_init();
      fence_async_smem();
    }
    __syncwarp();
    ring_fill(rs);
    ring_fill(rw);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  c.tbase = uniform(tmem_slot);
  c.w.tl = c.tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  float* Z = sm + p.ZBUF; float* V = sm + p.VBUF;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long q = (long long)tile * p.rows + c.w.r;
    const bool live = c.w.r < p.rows && q < p.E;
    const int src = live ? p.src[q] : 0, dst = live ? p.dst[q] : 0, orig = live ? p.perm[q] : 0;
    float* saved_t = p.saved ? p.saved + (size_t)tile * p.saved_tile_stride : nullptr;
    // the gathers below overwrite tile columns other threads of the row still read in the previous tile's epilogue B
    if (tile != (int)blockIdx.x) { wait_st(); __syncthreads(); }
    for (int k = 0; k < p.L; ++k) {
      const float* smc;
      gcp_forward_tile<CS>(p, k, c, rs, rw, k > 0 && p.residual, true, k == p.L - 1, q, live, src, dst, orig,
                           [&]() {
                             if (k > 0 && saved_t) {  // inputs of this GCP = outputs of the previous one (published by the barrier)
                               image_store(saved_t + (size_t)(k - 1) * (p.s_img + p.v_img), Z, p.s >> 2, p.rows);
                               image_store(saved_t + (size_t)(k - 1) * (p.s_img + p.v_img) + p.s_img, V, 3 * (PW >> 2), p.rows);
                             }
                           },
                           &smc);
      rs.head += 1;
    }
    __syncthreads();  // the last GCP's epilogue B wrote the final messages into the Z / V tiles
    segment_sums<CS>(p, Z, V, c.w, tile, dst, live);
  }
  if (c.uwarp == 0 && elect_one()) bulk_store_wait_all();
  fence_before_sync();
  __syncthreads();
  if (c.uwarp == 0) tmem_dealloc(c.tbase, (uint32_t)p.tmem_cols);
}

// =====================================================================================================================
// backward
// =====================================================================================================================

constexpr int GPLANE = (VN / 4) * SLAB;  // floats per plane of the [gH | gD | gU] tile (32 columns)

// Legacy-path tensor-core MMA (m16n8k8, tf32) for the weight-gradient products: their reduction runs over the tile's
// ROWS, an operand orientation tcgen05 only accepts in the 128B-swizzled MN-major layout; the register-fragment MMA reads
// the slab tiles directly.  3xTF32 with the split done in registers.
__device__ __forceinline__ void hmma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// c[b][16 x 8 block at (m0, n0 + 8 b)] += sum over the tile rows e < rows of A[e][m0 + .] * B[e][n0 + 8 b + .]   (A, B slab tiles)
// NB column blocks share the A fragments; the three 3xTF32 terms keep separate accumulators (short dependency chains).
// Straight-line body: explicit 32-bit shared addresses with immediate offsets, no branch between the fragment loads and
// the MMAs (blocks beyond `nb` recompute block 0 and are dropped by the caller) -- with a per-block branch the compiler
// re-derives the shared window and re-converges the warp before every mma.sync, which made the product latency-bound
// (5.7 k cycles for 5 row steps, r2 stage stamps).
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <int NB>
__device__ __forceinline__ void wgrad_blocks(float (&c)[NB][4], const float* A, const float* B, int m0, int n0, int nb, int lane, int rows) {
  const int g = lane >> 2, t = lane & 3;
  const uint32_t a0 = smem_addr(A + slab_off(RP, t, m0 + g)), a1 = smem_addr(A + slab_off(RP, t, m0 + g + 8));
  uint32_t bb0[NB];
#pragma unroll
  for (int bb = 0; bb < NB; ++bb) bb0[bb] = smem_addr(B + slab_off(RP, t, n0 + 8 * (bb < nb ? bb : 0) + g));
  float c1[NB][4], c2[NB][4];
#pragma unroll
  for (int bb = 0; bb < NB; ++bb)
#pragma unroll
    for (int q = 0; q < 4; ++q) { c1[bb][q] = 0.f; c2[bb][q] = 0.f; }
#pragma unroll 1
  for (int k0 = 0; k0 < rows; k0 += 8) {  // 8 tile rows per step = 128 bytes inside a slab
    const uint32_t ko = (uint32_t)k0 * 16u;
    const float fa[4] = {lds_f32(a0 + ko), lds_f32(a1 + ko), lds_f32(a0 + ko + 64u), lds_f32(a1 + ko + 64u)};
    float fb[NB][2];
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) { fb[bb][0] = lds_f32(bb0[bb] + ko); fb[bb][1] = lds_f32(bb0[bb] + ko + 64u); }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { ah[q] = __float_as_uint(fa[q]) & TF32_MASK; al[q] = __float_as_uint(fa[q] - __uint_as_float(ah[q])); }
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      const uint32_t bh0 = __float_as_uint(fb[bb][0]) & TF32_MASK, bh1 = __float_as_uint(fb[bb][1]) & TF32_MASK;
      const uint32_t bl0 = __float_as_uint(fb[bb][0] - __uint_as_float(bh0)), bl1 = __float_as_uint(fb[bb][1] - __uint_as_float(bh1));
      hmma_tf32(c[bb], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
      hmma_tf32(c1[bb], al[0], al[1], al[2], al[3], bh0, bh1);
      hmma_tf32(c2[bb], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
    }
  }
#pragma unroll
  for (int bb = 0; bb < NB; ++bb)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[bb][q] += c1[bb][q] + c2[bb][q];
}
// previous partial of a 16 x 8 block (multi-tile CTAs accumulate)
__device__ __forceinline__ void wgrad_fetch(float (&o)[4], const float* G, int ld, int m0, int n0, int mrows, int ncols, bool accumulate, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int m = m0 + g + 8 * hh, n = n0 + 2 * t;
    float2 v = make_float2(0.f, 0.f);
    if (accumulate && m < mrows && n < ncols) v = *reinterpret_cast<const float2*>(G + (size_t)m * ld + n);
    o[2 * hh] = v.x; o[2 * hh + 1] = v.y;
  }
}
// fragment (+ the previous partial `o`) -> this CTA's partial block G[ld columns] at (m0, n0); rows >= mrows / columns >= ncols are dropped
__device__ __forceinline__ void wgrad_store(const float (&c)[4], const float (&o)[4], float* G, int ld, int m0, int n0, int mrows, int ncols, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int m = m0 + g + 8 * hh, n = n0 + 2 * t;
    if (m < mrows && n < ncols)  // ncols even
      *reinterpret_cast<float2*>(G + (size_t)m * ld + n) = make_float2(c[2 * hh] + o[2 * hh], c[2 * hh + 1] + o[2 * hh + 1]);
  }
}

template <int CS>
__global__ void __launch_bounds__(128 * CS, 1) tc_edge_bwd_kernel(const __grid_constant__ TcBwdParams b) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_slot;
  const TcEdgeParams& p = b.f;
  const int ntiles = (p.E + p.rows - 1) / p.rows;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  TileCtx c;
  c.sm = sm;
  c.w.tid = (int)threadIdx.x;
  const int warp = c.w.tid >> 5, lane = c.w.tid & 31;
  c.w.r = 32 * (warp & 3) + lane;
  c.w.part = warp >> 2;
  c.uwarp = uniform(warp);
  const Who& w = c.w;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + p.BARS);
  c.mma_bar = bars;
  unsigned long long* ld_bar = bars + 1;   // V image of the GCP about to run
  unsigned long long* ldz_bar = bars + 2;  // S image (issued one GCP ahead, as soon as the Z tile's last reader is done)
  c.mma_n = 0;
  uint32_t ld_n = 0, ldz_n = 0;
  Ring rs{sm + p.RING_S, bars + 3, &p.ring_s, p.blob, 0, 0, mine * p.ring_s.n};
  Ring rw{sm + p.RING_W, bars + 3 + p.ring_s.nslot, &p.ring_w, p.blob, 0, 0, mine * p.ring_w.n};
  if (c.uwarp == 0) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    if (elect_one()) {
      mbar_init(c.mma_bar, 1);
      mbar_init(ld_bar, 1);
      mbar_init(ldz_bar, 1);
      for (int i = 0; i < p.ring_s.nslot; ++i) mbar_init(&rs.bar[i], 1);
      for (int i = 0; i < p.ring_w.nslot; ++i) mbar_init(&rw.bar[i], 1);
      mbar_fence_init();
      fence_async_smem();
    }
    __syncwarp();
    ring_fill(rs);
    ring_fill(rw);
  }
  float* Z = sm + p.ZBUF; float* V = sm + p.VBUF; float* GTG = sm + b.GTG; float* GHDU = sm + b.GHDU;
  // [gH | gD | gU] tile: columns hd..12 are never written below and meet zero weights -- they only have to be finite
  for (int i = c.w.tid; i < 3 * GPLANE; i += 128 * CS) GHDU[i] = 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  c.tbase = uniform(tmem_slot);
  c.w.tl = c.tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t tbase = c.tbase;
  for (int cc = 4 * w.part; cc < 3 * VN; cc += 4 * CS) tmem_st4(w.tl + (uint32_t)(b.GHDULO + cc), 0.f, 0.f, 0.f, 0.f);
  const int W = p.s + 3 * p.v;
  float* prow = b.partial + (size_t)blockIdx.x * b.partial_stride;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const bool accumulate = tile != (int)blockIdx.x;
    const long long q = (long long)tile * p.rows + w.r;
    const bool live = w.r < p.rows && q < p.E;
    const int src = live ? p.src[q] : 0, dst = live ? p.dst[q] : 0, orig = live ? p.perm[q] : 0;
    const float* saved_t = p.saved + (size_t)tile * p.saved_tile_stride;
    // ---- cotangent of the final message: gagg[dst] (/ in-degree for the mean reduce, gcpnet.py:946) -> GS, GV (TMEM)
    {
      float scale = 0.f;
      if (live) {
        scale = 1.f;
        if (b.reduce_mean) { const int deg = b.dst_ptr[dst + 1] - b.dst_ptr[dst]; scale = 1.f / (float)(deg > 1 ? deg : 1); }
      }
      const float* gp = b.gagg + (size_t)dst * W;
      for (int cc = 4 * w.part; cc < p.s; cc += 4 * CS) {
        const float4 v = live ? __ldg(reinterpret_cast<const float4*>(gp + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_st4(w.tl + (uint32_t)(b.GS + cc), v.x * scale, v.y * scale, v.z * scale, v.w * scale);
      }
      for (int gi = w.part; gi < (PW >> 2); gi += CS) {
        float f[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) f[j] = 0.f;
        if (live && 4 * gi < p.v) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(gp + p.s + 12 * gi + 4 * j));
            f[4 * j] = v.x * scale; f[4 * j + 1] = v.y * scale; f[4 * j + 2] = v.z * scale; f[4 * j + 3] = v.w * scale;
          }
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) tmem_st4(w.tl + (uint32_t)(b.GV + PW * x + 4 * gi), f[x], f[3 + x], f[6 + x], f[9 + x]);
      }
    }
    if (w.part == CS - 1) {  // frames of the tile's edges (the forward loads them with GCP 0, which comes LAST here)
      float* F = sm + p.FBUF + w.r * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) F[i] = live ? __ldg(p.frames + (size_t)orig * 9 + i) : 0.f;
    }
    for (int k = p.L - 1; k >= 0; --k) {
      const TcGcp& g = p.g[k];
      const bool res = k > 0 && p.residual;
#if GCP_STAMPS
      auto bstamp = [&](int i) { if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 0) p.dbg[(12 + k) * 16 + i] = clock64(); };
#else
      auto bstamp = [](int) {};
#endif
      bstamp(0);
      // every reader of the Z / V / cotangent tiles of the previous GCP (weight-gradient products) is done
      wait_st();
      __syncthreads();
      // ---- inputs of GCP k: saved images (k > 0) -> Z[:, :s], V; their lo parts -> TMEM
      if (k > 0) {
        if (c.uwarp == 0 && elect_one()) {
          const float* sp = saved_t + (size_t)(k - 1) * (p.s_img + p.v_img);
          image_load(V, sp + p.s_img, 3 * (PW >> 2), p.rows, ld_bar);
          if (k == p.L - 1) image_load(Z, sp, p.s >> 2, p.rows, ldz_bar);  // first GCP of the tile: nobody prefetched its S image
        }
        mbar_wait(ldz_bar, ldz_n & 1u);  // prefetched one GCP ahead: normally complete
        ++ldz_n;
        __syncwarp();
        for (int cc = 4 * w.part; cc < p.s; cc += 4 * CS) {  // ... so its lo split overlaps with the V image still in flight
          const float4 v = get4(Z, w.r, cc);
          tmem_st4(w.tl + (uint32_t)(p.ZLO + cc), lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
        }
        mbar_wait(ld_bar, ld_n & 1u);
        ++ld_n;
        __syncwarp();
        for (int gi = w.part; gi < (PW >> 2); gi += CS)
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const float4 v = get4(V + x * PLANE, w.r, 4 * gi);
            tmem_st4(w.tl + (uint32_t)(p.VLO + PW * x + 4 * gi), lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
          }
      }
      bstamp(1);
      // ---- recompute the forward of GCP k up to [H | D | U] and [T | g]
      const float* smc;
      gcp_forward_tile<CS>(p, k, c, rs, rw, false, false, false, q, live, src, dst, orig, [] {}, &smc, true);
      const float* ps = nullptr; const float* pd = nullptr; const float* qs = nullptr; const float* qd = nullptr;
      if (k == 0) {
        ps = p.P + (size_t)src * 2 * p.pw; pd = p.P + (size_t)dst * 2 * p.pw + p.pw;
        qs = p.Q + (size_t)src * 192; qd = p.Q + (size_t)dst * 192 + 96;
      }
      const float* bs = smc + g.o_b;
      const float* bg = smc + g.o_b + g.sop;
      bstamp(2);
      // ---- epilogue 1: cotangents of the two batches' outputs: [gT | gg] (GTG tile), gU (GHDU tile columns 16..31)
      for (int cg = w.part; 16 * cg < g.so; cg += CS) {
        float t[16], gs[16];
        tmem_ld16(w.tl + (uint32_t)(p.TACC + 16 * cg), t);
        tmem_ld16(w.tl + (uint32_t)(b.GS + 16 * cg), gs);
        wait_ld();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int cc = 16 * cg + 4 * i4;
          if (cc < g.so) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + cc);
            float tv[4] = {t[4 * i4] + b4.x, t[4 * i4 + 1] + b4.y, t[4 * i4 + 2] + b4.z, t[4 * i4 + 3] + b4.w};
            if (ps != nullptr) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(ps + cc)), d4 = __ldg(reinterpret_cast<const float4*>(pd + cc));
              tv[0] += a.x + d4.x; tv[1] += a.y + d4.y; tv[2] += a.z + d4.z; tv[3] += a.w + d4.w;
            }
            float gt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) gt[i] = gs[4 * i4 + i] * act_grad(g.act_s, tv[i], p.slope);
            put4(GTG, w.tl + (uint32_t)p.ZLO, w.r, cc, gt[0], gt[1], gt[2], gt[3]);
          }
        }
      }
      for (int gi = w.part; gi < (PW >> 2); gi += CS) {
        float gt[4], u[3][4], gv[3][4];
        tmem_ld4(w.tl + (uint32_t)(p.TACC + g.sop + 4 * gi), gt);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          tmem_ld4(w.tl + (uint32_t)(p.VACC + VN * x + UCOL + 4 * gi), u[x]);
          tmem_ld4(w.tl + (uint32_t)(b.GV + PW * x + 4 * gi), gv[x]);
        }
        wait_ld();
        if (ps != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(ps + g.sop + 4 * gi)), d4 = __ldg(reinterpret_cast<const float4*>(pd + g.sop + 4 * gi));
          gt[0] += a.x + d4.x; gt[1] += a.y + d4.y; gt[2] += a.z + d4.z; gt[3] += a.w + d4.w;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const float4 e4 = __ldg(reinterpret_cast<const float4*>(qs + 32 * x + UCOL + 4 * gi)), f4 = __ldg(reinterpret_cast<const float4*>(qd + 32 * x + UCOL + 4 * gi));
            u[x][0] += e4.x + f4.x; u[x][1] += e4.y + f4.y; u[x][2] += e4.z + f4.z; u[x][3] += e4.w + f4.w;
          }
        }
        float4 vin[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) vin[x] = g.vres ? get4(V + x * PLANE, w.r, 4 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
        float gu[3][4], gg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = 4 * gi + i;
          const float sg = o < g.vo ? sigmoidf_(gt[i] + bg[o]) : 0.f;
          float gsig = 0.f;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const float vi = i == 0 ? vin[x].x : (i == 1 ? vin[x].y : (i == 2 ? vin[x].z : vin[x].w));
            gsig = fmaf(gv[x][i], u[x][i] + vi, gsig);
            gu[x][i] = gv[x][i] * sg;
          }
          gg[i] = gsig * sg * (1.f - sg);
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) put4(GHDU + x * GPLANE, w.tl + (uint32_t)(b.GHDULO + VN * x), w.r, UCOL + 4 * gi, gu[x][0], gu[x][1], gu[x][2], gu[x][3]);
        put4(GTG, w.tl + (uint32_t)p.ZLO, w.r, g.sop + 4 * gi, gg[0], gg[1], gg[2], gg[3]);
      }
      publish_and_sync();
      bstamp(3);
      // ---- scalar data gradient: gZ = [gT | gg] . W_tg   (accumulator aliases [T | g])
      if (c.uwarp == 0) {
        ring_fill(rw);
        const int wh = uniform(rw.head);
        const float* bh = ring_wait(rw, wh);
        const float* bl = ring_wait(rw, wh + 1);
        bool acc = false;
        mma3(tbase + (uint32_t)p.TACC, GTG, tbase + (uint32_t)p.ZLO, bh, bl, b.kzn[k], p.pw, make_idesc(128, b.kzn[k], 0, 0), acc);
        if (elect_one()) commit(c.mma_bar);
        __syncwarp();
      }
      bstamp(4);
      // ---- meanwhile: weight-gradient product G_tg += [gT | gg]^T . Z  (CUDA cores' tensor path, all warps)
      {
        constexpr int NB = 4;
        const int mt = p.pw >> 4, nt = (g.kz + 7) >> 3, ng = (nt + NB - 1) / NB;
        // warp 0 has just issued the data-gradient batch: when the other warps cover all blocks in one round it takes none
        const int nw = 4 * CS, skip0 = (mt * ng <= nw - 1) ? 1 : 0;
#if GCP_STAMPS
        if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 32) p.dbg[(12 + k) * 16 + 12] = clock64();
#endif
        for (int pr = warp - skip0; pr >= 0 && pr < mt * ng; pr += nw) {
          const int m0 = 16 * (pr % mt), nb0 = NB * (pr / mt);
          const int nb = nt - nb0 < NB ? nt - nb0 : NB;
          float cf[NB][4], old[NB][4];
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
#pragma unroll
            for (int q = 0; q < 4; ++q) cf[bb][q] = 0.f;
          wgrad_blocks<NB>(cf, GTG, Z, m0, 8 * nb0, nb, lane, p.rows);
          // multi-tile CTAs accumulate: ALL previous partials are fetched before the first store (one L2 round trip, not one per block)
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            wgrad_fetch(old[bb], prow + b.off_tg[k], g.kz, m0, 8 * (nb0 + bb), p.pw, bb < nb ? g.kz : 0, accumulate, lane);
#if GCP_STAMPS
          if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 32) p.dbg[(12 + k) * 16 + 13] = clock64();
#endif
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            if (bb < nb) wgrad_store(cf[bb], old[bb], prow + b.off_tg[k], g.kz, m0, 8 * (nb0 + bb), p.pw, g.kz, lane);
        }
#if GCP_STAMPS
        if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 32) p.dbg[(12 + k) * 16 + 14] = clock64();
#endif
      }
      bstamp(5);
      wait_mma(c);
      bstamp(6);
      rw.head += 2;
      if (c.uwarp == 0) ring_fill(rw);
      // ---- epilogue 3: gS (residual + gZ[:, :s]) -> GS; tail of gZ -> [gH | gD] (GHDU tile columns 0..15)
      if (k > 0) {
        for (int cg = w.part; 16 * cg < p.s; cg += CS) {
          float gz[16], gs[16];
          tmem_ld16(w.tl + (uint32_t)(p.TACC + 16 * cg), gz);
          if (res) tmem_ld16(w.tl + (uint32_t)(b.GS + 16 * cg), gs);
          wait_ld();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = gz[4 * i4 + i] + (res ? gs[4 * i4 + i] : 0.f);
            tmem_st4(w.tl + (uint32_t)(b.GS + 16 * cg + 4 * i4), o[0], o[1], o[2], o[3]);
          }
        }
      } else {
        for (int cg = w.part; 16 * cg < p.se; cg += CS) {  // cotangent of the edge scalars
          float gz[16];
          tmem_ld16(w.tl + (uint32_t)(p.TACC + 16 * cg), gz);
          wait_ld();
          if (live) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4)
              if (16 * cg + 4 * i4 < p.se)
                *reinterpret_cast<float4*>(b.ge + (size_t)orig * p.se + 16 * cg + 4 * i4) = make_float4(gz[4 * i4], gz[4 * i4 + 1], gz[4 * i4 + 2], gz[4 * i4 + 3]);
          }
        }
      }
#pragma unroll
      for (int item = 0; item < 4; ++item) {
        if ((item % CS) != w.part) continue;
        if (item < 3) {
          if (4 * item < g.nslot) {
            float gn[4], h[3][4];
            tmem_ld4(w.tl + (uint32_t)(p.TACC + g.zc0 + 4 * item), gn);
#pragma unroll
            for (int x = 0; x < 3; ++x) tmem_ld4(w.tl + (uint32_t)(p.VACC + VN * x + 4 * item), h[x]);
            wait_ld();
            if (qs != nullptr) {
#pragma unroll
              for (int x = 0; x < 3; ++x) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(qs + 32 * x + 4 * item)), d4 = __ldg(reinterpret_cast<const float4*>(qd + 32 * x + 4 * item));
                h[x][0] += a.x + d4.x; h[x][1] += a.y + d4.y; h[x][2] += a.z + d4.z; h[x][3] += a.w + d4.w;
              }
            }
            float gh[3][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // n = sqrt(sum_x H^2 + eps) + eps  ->  dn/dH[x] = H[x] / sqrt(sum_x H^2 + eps)
              const float root = sqrtf(fmaf(h[0][i], h[0][i], fmaf(h[1][i], h[1][i], h[2][i] * h[2][i])) + SAFE_NORM_EPS);
              const float f = (4 * item + i < g.hd) ? gn[i] / root : 0.f;
#pragma unroll
              for (int x = 0; x < 3; ++x) gh[x][i] = f * h[x][i];
            }
#pragma unroll
            for (int x = 0; x < 3; ++x) put4(GHDU + x * GPLANE, w.tl + (uint32_t)(b.GHDULO + VN * x), w.r, 4 * item, gh[x][0], gh[x][1], gh[x][2], gh[x][3]);
          }
        } else {
          float gq[12];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float t4[4];
            tmem_ld4(w.tl + (uint32_t)(p.TACC + g.zc0 + g.nslot + 4 * j), t4);
            gq[4 * j] = t4[0]; gq[4 * j + 1] = t4[1]; gq[4 * j + 2] = t4[2]; gq[4 * j + 3] = t4[3];
          }
          wait_ld();
          const float* F = sm + p.FBUF + w.r * 9;
          // q[3c + a] = sum_x F[a][x] D[x][c]  ->  gD[x][c] = sum_a F[a][x] gq[3c + a]
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            float gd[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) gd[cc] = fmaf(F[x], gq[3 * cc], fmaf(F[3 + x], gq[3 * cc + 1], F[6 + x] * gq[3 * cc + 2]));
            put4(GHDU + x * GPLANE, w.tl + (uint32_t)(b.GHDULO + VN * x), w.r, 12, 0.f, gd[0], gd[1], gd[2]);
          }
        }
      }
      publish_and_sync();
      bstamp(7);
      // ---- vector data gradient: gV_in = [gH | gD | gU] . W_v   (accumulator aliases the V lo region)
      if (c.uwarp == 0) {
        bool acc = false;
        mma3_planes(tbase + (uint32_t)p.VLO, PW, GHDU, GPLANE, tbase + (uint32_t)b.GHDULO, VN, smc + g.o_wvt_hi, smc + g.o_wvt_lo, 16, VN,
                    make_idesc(128, 16, 0, 0), acc);
        if (elect_one()) {
          commit(c.mma_bar);
          if (k > 1)  // S image of the next GCP: the Z tile's readers (scalar batches, weight-gradient product) are done
            image_load(Z, saved_t + (size_t)(k - 2) * (p.s_img + p.v_img), p.s >> 2, p.rows, ldz_bar);
          if (k == 0) {  // per-edge cotangents of message GCP 0's per-node products, for the node-level finish
            float* yp = b.Y + (size_t)tile * (b.y_img_g + b.y_img_v);
            image_store(yp, GTG, p.pw >> 2, p.rows);
            image_store(yp + b.y_img_g, GHDU, 3 * (VN >> 2), p.rows);
          }
        }
        __syncwarp();
      }
      bstamp(8);
      // ---- meanwhile: G_v += sum_xyz [gH | gD | gU]^T . V_in
      if (warp >= 4 * CS - 4) {
        const int pr = warp - (4 * CS - 4);  // 2 x 2 blocks of 16 x 8
        const int m0 = 16 * (pr & 1), n0 = 8 * (pr >> 1);
        float cf[1][4] = {{0.f, 0.f, 0.f, 0.f}}, old[4];
#pragma unroll 1
        for (int x = 0; x < 3; ++x) wgrad_blocks<1>(cf, GHDU + x * GPLANE, V + x * PLANE, m0, n0, 1, lane, p.rows);
        wgrad_fetch(old, prow + b.off_v[k], 16, m0, n0, VN, 16, accumulate, lane);
        wgrad_store(cf[0], old, prow + b.off_v[k], 16, m0, n0, VN, 16, lane);
      }
      bstamp(9);
      wait_mma(c);
      bstamp(10);
      // ---- epilogue 4: gV (residual (+ gU with vector_residual) + gV_in) -> GV; GCP 0: cotangent of the edge vectors
      for (int gi = w.part; gi < (PW >> 2); gi += CS) {
        float a[3][4], gv[3][4];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          tmem_ld4(w.tl + (uint32_t)(p.VLO + PW * x + 4 * gi), a[x]);
          if (res) tmem_ld4(w.tl + (uint32_t)(b.GV + PW * x + 4 * gi), gv[x]);
        }
        wait_ld();
        if (k > 0) {
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            float o[4];
            const float4 gu = g.vres ? get4(GHDU + x * GPLANE, w.r, UCOL + 4 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
            o[0] = a[x][0] + gu.x; o[1] = a[x][1] + gu.y; o[2] = a[x][2] + gu.z; o[3] = a[x][3] + gu.w;
            if (res) { o[0] += gv[x][0]; o[1] += gv[x][1]; o[2] += gv[x][2]; o[3] += gv[x][3]; }
            tmem_st4(w.tl + (uint32_t)(b.GV + PW * x + 4 * gi), o[0], o[1], o[2], o[3]);
          }
        } else if (live && 4 * gi < p.ve) {
          float f[12];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int x = 0; x < 3; ++x) f[3 * i + x] = a[x][i];
          float* gp = b.gxi + (size_t)orig * 3 * p.ve + 12 * gi;
#pragma unroll
          for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(gp + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
      bstamp(11);
      rs.head += 1;
    }
    if (c.uwarp == 0 && elect_one()) bulk_store_wait_read();  // the Y images of this tile have left shared memory
  }
  if (c.uwarp == 0 && elect_one()) bulk_store_wait_all();
  wait_st();
  fence_before_sync();
  __syncthreads();
  if (c.uwarp == 0) tmem_dealloc(c.tbase, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// node-level finish of message GCP 0 (backward): Y[e] = [gT | gg | (gH | gD | gU) x 3] are the
// cotangents of the gathered per-node products; sum them per node over outgoing / incoming edges (fixed order, no
// atomics), push them through the node tiles (data gradient) and against h / chi (weight gradient).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float y_at(const TcPostParams& p, int q, int c) {
  const int tile = q / p.rows, r = q - tile * p.rows;
  const float* yp = p.Y + (size_t)tile * (p.y_img_g + p.y_img_v);
  if (c < p.pw) return __ldg(yp + ((c >> 2) * p.rows + r) * 4 + (c & 3));  // compact images: slab pitch = rows
  const int c2 = c - p.pw, x = c2 >> 5, cc = c2 & 31;
  return __ldg(yp + p.y_img_g + ((x * (VN >> 2) + (cc >> 2)) * p.rows + r) * 4 + (cc & 3));
}
// sum of column c of the Y rows of node i's outgoing (side 0) / incoming (side 1) edges, in CSR order; four edges are
// fetched at a time so that their (dependent index -> value) loads overlap
__device__ __forceinline__ float y_node_sum(const TcPostParams& p, int i, int side, int c) {
  const int* ptr = side == 0 ? p.src_ptr : p.dst_ptr;
  const int e0 = __ldg(ptr + i), e1 = __ldg(ptr + i + 1);
  float acc = 0.f;
  for (int j = e0; j < e1; j += 4) {
    int q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = j + u < e1 ? (side == 0 ? __ldg(p.src_pos + j + u) : j + u) : -1;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = q[u] >= 0 ? y_at(p, q[u], c) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) if (q[u] >= 0) acc += v[u];
  }
  return acc;
}
// Four consecutive columns at once (one 16-byte load per edge instead of four 4-byte ones: the images keep 4-column groups
// contiguous); every column is still summed over the same edges in the same order -> same bits as y_node_sum.
__device__ __forceinline__ const float4* y_at4(const TcPostParams& p, int q, int c4) {  // c4 = column / 4
  const int tile = q / p.rows, r = q - tile * p.rows;
  const float* yp = p.Y + (size_t)tile * (p.y_img_g + p.y_img_v);
  const int g4 = p.pw >> 2;
  const int slab = c4 < g4 ? c4 : c4 - g4;  // [gT|gg] image: pw/4 slabs; [gH|gD|gU] image: 3 planes x 8 slabs, contiguous
  return reinterpret_cast<const float4*>(yp + (c4 < g4 ? 0 : p.y_img_g) + ((size_t)slab * p.rows + r) * 4);
}
__device__ __forceinline__ float4 y_node_sum4(const TcPostParams& p, int i, int side, int c4) {
  const int* ptr = side == 0 ? p.src_ptr : p.dst_ptr;
  const int e0 = __ldg(ptr + i), e1 = __ldg(ptr + i + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = e0; j < e1; j += 4) {
    int q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = j + u < e1 ? (side == 0 ? __ldg(p.src_pos + j + u) : j + u) : -1;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = q[u] >= 0 ? __ldg(y_at4(p, q[u], c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (q[u] >= 0) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  return acc;
}
__global__ void __launch_bounds__(256) tc_post_sum_kernel(const TcPostParams p) {
  const int per = p.pw + 96;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)p.N * 2 * per) return;
  const int i = (int)(idx / (2 * per)), rem = (int)(idx - (long long)i * 2 * per);
  const int side = rem / per, c = rem - side * per;
  p.A[idx] = y_node_sum(p, i, side, c);
}
__global__ void __launch_bounds__(256) tc_post_data_kernel(const TcPostParams p) {
  const int W = p.s + 3 * p.v, per = p.pw + 96;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)p.N * W) return;
  const int i = (int)(idx / W), o = (int)(idx - (long long)i * W);
  const float* As = p.A + (size_t)i * 2 * per;
  const float* Ad = As + per;
  float acc = 0.f;
  if (o < p.s) {
    const float* Bs = p.blob + p.nt.ps; const float* Bd = p.blob + p.nt.pd;  // [pw][s] slab, pitch pw
    const float* bs0 = Bs + (o >> 2) * p.pw * 4 + (o & 3);
    const float* bd0 = Bd + (o >> 2) * p.pw * 4 + (o & 3);
    float acc2 = 0.f;
#pragma unroll 2
    for (int c = 0; c < p.pw; c += 4) {  // pw % 4 == 0; A rows are 16-byte aligned
      const float4 a = *reinterpret_cast<const float4*>(As + c), d = *reinterpret_cast<const float4*>(Ad + c);
      acc = fmaf(a.x, __ldg(bs0 + 4 * c), fmaf(a.y, __ldg(bs0 + 4 * c + 4), fmaf(a.z, __ldg(bs0 + 4 * c + 8), fmaf(a.w, __ldg(bs0 + 4 * c + 12), acc))));
      acc2 = fmaf(d.x, __ldg(bd0 + 4 * c), fmaf(d.y, __ldg(bd0 + 4 * c + 4), fmaf(d.z, __ldg(bd0 + 4 * c + 8), fmaf(d.w, __ldg(bd0 + 4 * c + 12), acc2))));
    }
    acc += acc2;
    p.g_h[(size_t)i * p.s + o] += acc;
  } else {
    const int ch = (o - p.s) / 3, x = (o - p.s) - 3 * ch;
    const float* Bs = p.blob + p.nt.qs; const float* Bd = p.blob + p.nt.qd;  // [32][v8] slab, pitch 32
    for (int c = 0; c < VN; ++c) {
      const int off = ((ch >> 2) * VN + c) * 4 + (ch & 3);
      acc = fmaf(As[p.pw + 32 * x + c], __ldg(Bs + off), fmaf(Ad[p.pw + 32 * x + c], __ldg(Bd + off), acc));
    }
    p.g_chi[(size_t)i * 3 * p.v + (o - p.s)] += acc;
  }
}
// Fused version of the two kernels above (the layer below waits for dh / dchi): the CTA first forms the per-node sums of
// the nodes its 256 outputs touch (shared memory + the global copy the weight-gradient kernel reads; a node shared with
// the neighbouring CTA is summed by both, same bits), then pushes them through the node tiles.  Launch condition:
// s + 3v >= 48 (at most POST_ROWS nodes per CTA) and pw + 96 <= POST_PER.
constexpr int POST_ROWS = 8, POST_PER = 240;
__global__ void __launch_bounds__(256) tc_post_fused_kernel(const TcPostParams p) {
  __shared__ __align__(16) float As_sm[POST_ROWS][2 * POST_PER];
  const int W = p.s + 3 * p.v, per = p.pw + 96;
  const long long idx0 = (long long)blockIdx.x * 256;
  const int i_first = (int)(idx0 / W);
  const long long last = idx0 + 255 < (long long)p.N * W - 1 ? idx0 + 255 : (long long)p.N * W - 1;
  const int nrow = (int)(last / W) - i_first + 1;
  const int per4 = per >> 2;  // per = pw + 96, a multiple of 4
  for (int t = threadIdx.x; t < nrow * 2 * per4; t += 256) {
    const int r = t / (2 * per4), rem4 = t - r * 2 * per4;
    const int i = i_first + r, side = rem4 / per4, c4 = rem4 - side * per4;
    const float4 acc = y_node_sum4(p, i, side, c4);
    *reinterpret_cast<float4*>(&As_sm[r][side * per + 4 * c4]) = acc;
    *reinterpret_cast<float4*>(p.A + (size_t)i * 2 * per + side * per + 4 * c4) = acc;
  }
  __syncthreads();
  const long long idx = idx0 + threadIdx.x;
  if (idx >= (long long)p.N * W) return;
  const int i = (int)(idx / W), o = (int)(idx - (long long)i * W);
  const float* As = As_sm[i - i_first];
  const float* Ad = As + per;
  float acc = 0.f;
  if (o < p.s) {
    const float* Bs = p.blob + p.nt.ps; const float* Bd = p.blob + p.nt.pd;  // [pw][s] slab, pitch pw
    const float* bs0 = Bs + (o >> 2) * p.pw * 4 + (o & 3);
    const float* bd0 = Bd + (o >> 2) * p.pw * 4 + (o & 3);
    float acc2 = 0.f;
#pragma unroll 4
    for (int c = 0; c < p.pw; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(As + c);
      acc = fmaf(a.x, __ldg(bs0 + 4 * c), fmaf(a.y, __ldg(bs0 + 4 * c + 4), fmaf(a.z, __ldg(bs0 + 4 * c + 8), fmaf(a.w, __ldg(bs0 + 4 * c + 12), acc))));
    }
#pragma unroll 4
    for (int c = 0; c < p.pw; c += 4) {
      const float4 d = *reinterpret_cast<const float4*>(Ad + c);  // per % 4 == 0 (pw = sop + 16)
      acc2 = fmaf(d.x, __ldg(bd0 + 4 * c), fmaf(d.y, __ldg(bd0 + 4 * c + 4), fmaf(d.z, __ldg(bd0 + 4 * c + 8), fmaf(d.w, __ldg(bd0 + 4 * c + 12), acc2))));
    }
    acc += acc2;
    p.g_h[(size_t)i * p.s + o] += acc;
  } else {
    const int ch = (o - p.s) / 3, x = (o - p.s) - 3 * ch;
    const float* Bs = p.blob + p.nt.qs; const float* Bd = p.blob + p.nt.qd;  // [32][v8] slab, pitch 32
#pragma unroll 4
    for (int c = 0; c < VN; ++c) {
      const int off = ((ch >> 2) * VN + c) * 4 + (ch & 3);
      acc = fmaf(As[p.pw + 32 * x + c], __ldg(Bs + off), fmaf(Ad[p.pw + 32 * x + c], __ldg(Bd + off), acc));
    }
    p.g_chi[(size_t)i * 3 * p.v + (o - p.s)] += acc;
  }
}
// partial rows (one per node chunk = blockIdx.y) [src: pw x s | dst: pw x s | src: 32 x 16 | dst: 32 x 16]; thread = one output
__global__ void __launch_bounds__(256) tc_post_wgrad_kernel(const TcPostParams p) {
  const int per = p.pw + 96, chunk = (p.N + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * chunk, i1 = min(p.N, i0 + chunk);
  float* out = p.npartial + (size_t)blockIdx.y * p.npartial_stride;
  const int ns = p.pw * p.s, nv = VN * 16;
  const int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= 2 * ns + 2 * nv) return;
  float acc = 0.f;
  if (o < 2 * ns) {
    const int side = o / ns, r = o - side * ns, n = r / p.s, j = r - n * p.s;
    const float* ap = p.A + (size_t)side * per + n;
    const float* hp = p.h + j;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) acc = fmaf(__ldg(ap + (size_t)i * 2 * per), __ldg(hp + (size_t)i * p.s), acc);
  } else {
    const int o2 = o - 2 * ns, side = o2 / nv, r = o2 - side * nv, jj = r / 16, ch = r - jj * 16;
    if (ch < p.v)
      for (int i = i0; i < i1; ++i)
#pragma unroll
        for (int x = 0; x < 3; ++x) acc = fmaf(__ldg(p.A + ((size_t)i * 2 + side) * per + p.pw + 32 * x + jj), __ldg(p.chi + (size_t)i * 3 * p.v + 3 * ch + x), acc);
  }
  out[o] = acc;
}
// out[idx] = sum over rows of partial[row][idx]
__global__ void __launch_bounds__(256) tc_reduce_kernel(const float* __restrict__ partial, int rows, int stride, int n, float* __restrict__ out) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += __ldg(partial + (size_t)r * stride + idx);
  out[idx] = acc;
}

__device__ __forceinline__ float fin_gtg(const TcFinalParams& p, int k, int n, int r) {  // composed scalar gradient at (row n, reference column r)
  const TcFinalGcp& g = p.g[k];
  if (k == 0) {
    if (r < p.s) return p.Gn[(size_t)n * p.s + r];
    if (r < p.s + p.se) return p.G[g.off_tg + (size_t)n * g.kz + (r - p.s)];
    if (r < 2 * p.s + p.se) return p.Gn[(size_t)p.pw * p.s + (size_t)n * p.s + (r - p.s - p.se)];
  } else if (r < g.si) return p.G[g.off_tg + (size_t)n * g.kz + r];
  const int t = r - g.si;  // n | q
  const int col = t < g.hd ? g.zc0 + t : g.zc0 + g.nslot + (t - g.hd);
  return p.G[g.off_tg + (size_t)n * g.kz + col];
}
__device__ __forceinline__ float fin_gb(const TcFinalParams& p, int k, int n) { return p.G[p.g[k].off_tg + (size_t)n * p.g[k].kz + p.g[k].kz - 1]; }
__device__ __forceinline__ float fin_gv(const TcFinalParams& p, int k, int j, int c) {  // composed vector gradient at (row j, reference channel c)
  const TcFinalGcp& g = p.g[k];
  if (k == 0) {
    const float* gn = p.Gn + 2 * (size_t)p.pw * p.s;
    if (c < p.v) return gn[j * 16 + c];
    if (c < p.v + p.ve) return p.G[g.off_v + j * 16 + (c - p.v)];
    return gn[VN * 16 + j * 16 + (c - p.v - p.ve)];
  }
  return p.G[g.off_v + j * 16 + c];
}
__global__ void __launch_bounds__(256) tc_finalize_kernel(const __grid_constant__ TcFinalParams p) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.n_edge_params) return;
  // which GCP / tensor
  int k = 0, t = 0, base = 0;
  bool found = false;
  for (int kk = 0; kk < p.L && !found; ++kk) {
    const TcFinalGcp& g = p.g[kk];
    const int K = g.si + g.hd + 9;
    const int sizes[7] = {g.hd * g.vi, 3 * g.vi, g.so * K, g.so, g.vo * g.hd, g.vo * g.so, g.vo};
    for (int tt = 0; tt < 7; ++tt)
      if (idx >= g.grad_off[tt] && idx < g.grad_off[tt] + sizes[tt]) { k = kk; t = tt; base = g.grad_off[tt]; found = true; break; }
  }
  if (!found) return;
  const TcFinalGcp& g = p.g[k];
  const int i = idx - base, K = g.si + g.hd + 9, so = g.so;
  float val = 0.f;
  if (t == 2) {          // scalar_out.weight[n][r] = G(n, r) + sum_o Wg[o][n] G(so + o, r)
    const int n = i / K, r = i - n * K;
    val = fin_gtg(p, k, n, r);
    for (int o = 0; o < g.vo; ++o) val = fmaf(__ldg(g.Wg + o * so + n), fin_gtg(p, k, so + o, r), val);
  } else if (t == 3) {   // scalar_out.bias
    val = fin_gb(p, k, i);
    for (int o = 0; o < g.vo; ++o) val = fmaf(__ldg(g.Wg + o * so + i), fin_gb(p, k, so + o), val);
  } else if (t == 5) {   // vector_out_scale.weight: tc_finalize_wg_kernel
    return;
  } else if (t == 6) {   // vector_out_scale.bias
    val = fin_gb(p, k, so + i);
  } else if (t == 0) {   // vector_down.weight[j][c] = Gv(j, c) + sum_o Wu[o][j] Gv(16 + o, c)
    const int j = i / g.vi, c = i - j * g.vi;
    val = fin_gv(p, k, j, c);
    for (int o = 0; o < g.vo; ++o) val = fmaf(__ldg(g.Wu + o * g.hd + j), fin_gv(p, k, UCOL + o, c), val);
  } else if (t == 1) {   // vector_down_frames.weight[cc][c]
    const int cc = i / g.vi, c = i - cc * g.vi;
    val = fin_gv(p, k, DCOL + cc, c);
  } else {               // vector_up.weight[o][j] = sum_c Gv(16 + o, c) Wd[j][c]
    const int o = i / g.hd, j = i - o * g.hd;
    for (int c = 0; c < g.vi; ++c) val = fmaf(fin_gv(p, k, UCOL + o, c), __ldg(g.Wd + j * g.vi + c), val);
  }
  p.out[idx] = val;
}

// vector_out_scale.weight[o][n] = sum_r G(so + o, r) Ws[n][r] + Gb(so + o) bs[n]: one warp per element, lanes split r
__global__ void __launch_bounds__(256) tc_finalize_wg_kernel(const __grid_constant__ TcFinalParams p) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int per = p.g[0].vo * p.g[0].so;  // same (vo, so) for every message GCP
  if (wid >= p.L * per) return;
  const int k = wid / per, i = wid - k * per;
  const TcFinalGcp& g = p.g[k];
  const int K = g.si + g.hd + 9, so = g.so, o = i / so, n = i - o * so;
  float val = 0.f;
  for (int r = lane; r < K; r += 32) val = fmaf(fin_gtg(p, k, so + o, r), __ldg(g.Ws + (size_t)n * K + r), val);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
  if (lane == 0) p.out[g.grad_off[5] + i] = val + fin_gb(p, k, so + o) * __ldg(g.bs + n);
}

}  // namespace tc
}  // namespace gcp
