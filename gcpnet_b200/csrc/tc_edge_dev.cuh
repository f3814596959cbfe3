// tc_edge_dev.cuh -- device code of the tensor-core edge kernels (see tc_edge.cuh for the design notes and
// the parameter structures).  Included only by tc_api.cu.
#pragma once
#include "gcp_tile.cuh"
#include "umma.cuh"
#include "tc_edge.cuh"
#include "tc_setup.h"

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
    const int off = ((kk >> 2) * it.R + n) * 4 + (kk & 3);
    blob[it.dst_hi + off] = val;
    blob[it.dst_lo + off] = umma::tf32_lo(val);
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
template <int CS>
__device__ __forceinline__ void epilogue_a(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w) {
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
      q[9] = q[10] = q[11] = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) put4(Z, zlo, w.r, g.zc0 + g.nslot + 4 * j, q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
    }
  }
}

// ---- epilogue B: scalar batch accumulator [T | g] + vector batch U -> new scalar state (Z tile), new vector state (V tile)
template <int CS>
__device__ __forceinline__ void epilogue_b(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w, const float* smc,
                                           bool add, bool last, long long q, bool live) {
  float* Z = sm + p.ZBUF;
  float* V = sm + p.VBUF;
  const float* bs = smc + g.o_bs;
  const float* bg = smc + g.o_bg;
  const int W = p.s + 3 * p.v;
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
        if (g.act_s == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sv[i] = fmaxf(sv[i], 0.f);
        } else if (g.act_s != ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sv[i] = act_fwd(g.act_s, sv[i], p.slope);
        }
        sv[0] += old.x; sv[1] += old.y; sv[2] += old.z; sv[3] += old.w;
        put4(Z, w.tl + (uint32_t)p.ZLO, w.r, c, sv[0], sv[1], sv[2], sv[3]);
        if (last && live) *reinterpret_cast<float4*>(p.msg + q * W + c) = make_float4(sv[0], sv[1], sv[2], sv[3]);
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
    if (last && live && 4 * gi < g.vo) {
      float* mp = p.msg + q * W + p.s + 12 * gi;  // [channel][xyz], xyz fastest; vo % 4 == 0
      float f[12];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int x = 0; x < 3; ++x) f[3 * i + x] = nv[x][i];
#pragma unroll
      for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(mp + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
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

template <int CS>
__global__ void __launch_bounds__(128 * CS, 1) tc_edge_fwd_kernel(const __grid_constant__ TcEdgeParams p) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_slot;
  const int ntiles = (p.E + TE - 1) / TE;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (mine == 0) return;
  Who w;
  w.tid = (int)threadIdx.x;
  const int warp = w.tid >> 5, lane = w.tid & 31;
  w.r = 32 * (warp & 3) + lane;
  w.part = warp >> 2;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + p.BARS);
  unsigned long long* mma_bar = bars;  // [0]; ring barriers follow
  Ring rs{sm + p.RING_S, bars + 1, &p.ring_s, p.blob, 0, 0, mine * p.ring_s.n};
  Ring rw{sm + p.RING_W, bars + 1 + p.ring_s.nslot, &p.ring_w, p.blob, 0, 0, mine * p.ring_w.n};
  const int uwarp = uniform(warp);
  if (uwarp == 0) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    if (elect_one()) {
      mbar_init(mma_bar, 1);
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
  const uint32_t tbase = uniform(tmem_slot);
  w.tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t mma_n = 0;  // commits on mma_bar so far (CTA-uniform)
  float* Z = sm + p.ZBUF; float* X = sm + p.XBUF; float* V = sm + p.VBUF;
  const uint32_t idesc_v = make_idesc(128, VN, 0, 0);

  auto stamp = [&](int k, int i) { if (p.dbg != nullptr && blockIdx.x == 0 && w.tid == 0) p.dbg[k * 16 + i] = clock64(); };
  auto wait_mma = [&]() {
    mbar_wait(mma_bar, mma_n & 1u);
    ++mma_n;
    __syncwarp();
    fence_after_sync();
  };

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long q = (long long)tile * TE + w.r;
    const bool live = q < p.E;
    const int src = live ? p.src[q] : 0, dst = live ? p.dst[q] : 0, orig = live ? p.perm[q] : 0;
    float* saved_t = p.saved ? p.saved + (size_t)tile * p.saved_tile_stride : nullptr;
    // the gathers below overwrite tile columns other threads of the row still read in the previous tile's epilogue B
    if (tile != (int)blockIdx.x) { wait_st(); __syncthreads(); }

    for (int k = 0; k < p.L; ++k) {
      const TcGcp& g = p.g[k];
      const bool last = k == p.L - 1;
      const bool add = k > 0 && p.residual;
      stamp(k, 0);
      // ---------------- vector batch operands (GCP 0: chi_row planes -> Z tile, xi -> V tile, chi_col -> X tile; frames)
      if (k == 0) {
        gather_planes<CS>(Z, w.tl + (uint32_t)p.ZLO, w, p.chi + (size_t)src * 3 * p.v, p.v, g.vkc[0], live);
        gather_planes<CS>(V, w.tl + (uint32_t)p.VLO, w, p.xi + (size_t)orig * 3 * p.ve, p.ve, g.vkc[1], live);
        gather_planes<CS>(X, w.tl + (uint32_t)p.XLO, w, p.chi + (size_t)dst * 3 * p.v, p.v, g.vkc[2], live);
        if (w.part == CS - 1) {
          float* F = sm + p.FBUF + w.r * 9;
#pragma unroll
          for (int i = 0; i < 9; ++i) F[i] = live ? __ldg(p.frames + (size_t)orig * 9 + i) : 0.f;
        }
      }
      publish_and_sync();
      stamp(k, 1);
      const float* smc = ring_wait(rs, uniform(rs.head));  // small chunk of this GCP (every thread: biases are read in epilogue B)
      if (uwarp == 0) {
        if (elect_one()) bulk_store_wait_read();
        ring_fill(rs);  // every thread is past the previous GCP's epilogue B (barrier above): that slot is free
        {
          bool acc = false;
          const uint32_t d = tbase + (uint32_t)p.VACC;
          if (k == 0) {
            mma3_planes(d, VN, Z, PLANE, tbase + (uint32_t)p.ZLO, PW, smc + g.o_wd_hi[0], smc + g.o_wd_lo[0], VN, g.vkc[0], idesc_v, acc);
            mma3_planes(d, VN, V, PLANE, tbase + (uint32_t)p.VLO, PW, smc + g.o_wd_hi[1], smc + g.o_wd_lo[1], VN, g.vkc[1], idesc_v, acc);
            mma3_planes(d, VN, X, PLANE, tbase + (uint32_t)p.XLO, PW, smc + g.o_wd_hi[2], smc + g.o_wd_lo[2], VN, g.vkc[2], idesc_v, acc);
          } else {
            mma3_planes(d, VN, V, PLANE, tbase + (uint32_t)p.VLO, PW, smc + g.o_wd_hi[0], smc + g.o_wd_lo[0], VN, g.vkc[0], idesc_v, acc);
          }
        }
        if (elect_one()) {
          commit(mma_bar);
          if (k > 0 && saved_t) {  // inputs of this GCP = outputs of the previous one
            bulk_store(saved_t + (size_t)(k - 1) * (p.s_img + p.v_img), Z, p.s_img);
            bulk_store(saved_t + (size_t)(k - 1) * (p.s_img + p.v_img) + p.s_img, V, p.v_img);
          }
        }
        __syncwarp();
      }
      stamp(k, 2);
      wait_mma();
      stamp(k, 3);
      // ---------------- epilogue A (+ GCP 0: edge scalars into the Z tile)
      epilogue_a<CS>(p, g, sm, w);
      if (k == 0) gather_row<CS>(Z, w.tl + (uint32_t)p.ZLO, w, p.e + (size_t)orig * p.se, p.se, p.se, live);
      stamp(k, 4);
      publish_and_sync();
      stamp(k, 5);
      // ---------------- scalar batch: one commit per K-segment
      const uint32_t idesc_s = make_idesc(128, g.sop + 16, 0, 0);
      bool tacc = false;
      for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const TcSeg sg = g.seg[sgi];
        if (sg.a_tile != 0) {  // GCP 0: the X tile takes h_row, then h_col
          gather_row<CS>(X, w.tl + (uint32_t)p.XLO, w, p.h + (size_t)(sg.a_tile == 1 ? src : dst) * p.s, p.s, sg.kc, live);
          publish_and_sync();
        }
        if (uwarp == 0) {
          if (elect_one()) bulk_store_wait_read();
          ring_fill(rw);
          const int wh = uniform(rw.head);
          const float* bh = ring_wait(rw, wh);
          const float* bl = ring_wait(rw, wh + 1);
          mma3(tbase + (uint32_t)p.TACC, sg.a_tile == 0 ? Z : X, tbase + (uint32_t)(sg.a_tile == 0 ? p.ZLO : p.XLO), bh, bl, g.sop + 16, sg.kc,
               idesc_s, tacc);
          if (elect_one()) commit(mma_bar);
          __syncwarp();
        }
        tacc = true;
        stamp(k, 6);
        wait_mma();
        stamp(k, 7);
        rw.head += 2;
      }
      // ---------------- epilogue B
      epilogue_b<CS>(p, g, sm, w, smc, add, last, q, live);
      stamp(k, 8);
      rs.head += 1;
    }
  }
  if (uwarp == 0 && elect_one()) bulk_store_wait_all();
  fence_before_sync();
  __syncthreads();
  if (uwarp == 0) tmem_dealloc(tbase, (uint32_t)p.tmem_cols);
}

}  // namespace tc
}  // namespace gcp
