// tc_edge_dev.cuh -- device code of the tensor-core edge kernels (see tc_edge.cuh for the design notes and
// the parameter structures).  Included only by tc_api.cu.
#pragma once
#include "tc_edge.cuh"

namespace gcp {
namespace tc {

struct Ring {           // CTA-uniform state; only thread 0 issues copies
  float* slots; unsigned long long* bar; const RingDesc* d; const float* blob;
  int head, issued, total;
};
__device__ __forceinline__ void ring_fill(Ring& r) {  // thread 0: top the ring up (slots < head are free)
  while (r.issued < r.total && r.issued < r.head + r.d->nslot) {
    const TcChunk ck = r.d->c[r.issued % r.d->n];
    const int slot = r.issued % r.d->nslot;
    const uint32_t bar = smem_addr(&r.bar[slot]);
    const uint32_t bytes = (uint32_t)ck.floats * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(r.slots + (size_t)slot * r.d->slot_floats)), "l"(r.blob + ck.off), "r"(bytes), "r"(bar) : "memory");
    ++r.issued;
  }
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

__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(a)) : "memory");
}

// row-owner stores: hi part to the slab tile, lo part to TMEM
__device__ __forceinline__ void put4(float* tile, uint32_t tlo, int r, int c, float a, float b, float c2, float d) {
  *reinterpret_cast<float4*>(tile + slab_off(RP, r, c)) = make_float4(a, b, c2, d);
  tmem_st4(tlo + (uint32_t)c, tf32_lo(a), tf32_lo(b), tf32_lo(c2), tf32_lo(d));
}
__device__ __forceinline__ void put1(float* tile, uint32_t tlo, int r, int c, float a) {
  tile[slab_off(RP, r, c)] = a;
  tmem_st1(tlo + (uint32_t)c, tf32_lo(a));
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
// D[128][N] (+)= A[:, 0:kc] . B[N][kc]^T.  One thread.
__device__ __forceinline__ void mma3(uint32_t d_tmem, const float* a_hi, uint32_t a_lo, const float* b_hi, const float* b_lo,
                                     int b_rows, int kc, uint32_t idesc, bool& acc) {
  const uint64_t ad0 = desc_kmajor(a_hi, RP, 0, 0);
  const uint64_t bh0 = make_desc(smem_addr(b_hi), (uint32_t)b_rows * 16u, 128u);
  const uint64_t bl0 = make_desc(smem_addr(b_lo), (uint32_t)b_rows * 16u, 128u);
  const uint32_t astep = 2u * SLAB * 4u, bstep = 2u * (uint32_t)b_rows * 16u;
  for (int ks = 0; ks < (kc >> 3); ++ks) {
    const uint64_t ad = desc_advance(ad0, ks * astep), bh = desc_advance(bh0, ks * bstep), bl = desc_advance(bl0, ks * bstep);
    mma_tf32(d_tmem, ad, bh, idesc, acc);
    mma_tf32(d_tmem, ad, bl, idesc, true);
    mma_tf32_ts(d_tmem, a_lo + (uint32_t)(8 * ks), bh, idesc, true);
    acc = true;
  }
}

// ---- epilogue A: vector_down accumulator -> H tile (vector_up operand); norms + frame scalars -> Z tile ----------------
// HDACC: 3 planes x 16 columns, [0, hd) hidden channels, [13, 16) frame-down vectors (gcpnet.py:420,426).
template <int CS>
__device__ __forceinline__ void epilogue_a(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w) {
  float hdv[3][16];
#pragma unroll
  for (int x = 0; x < 3; ++x) tmem_ld16(w.tl + (uint32_t)(p.HDACC + 16 * x), hdv[x]);
  wait_ld();
  float* H = sm + p.HBUF;
  float* Z = sm + p.ZBUF;
  const uint32_t zlo = w.tl + (uint32_t)p.ZLO;
  // H planes (hdp columns each); parts take the 4-column groups round robin
#pragma unroll
  for (int x = 0; x < 3; ++x)
#pragma unroll
    for (int gi = 0; gi < 4; ++gi)
      if (4 * gi < g.hdp && ((x * 4 + gi) % CS) == w.part) {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (4 * gi + i < g.hd) ? hdv[x][4 * gi + i] : 0.f;
        put4(H + x * PLANE, w.tl + (uint32_t)(p.HLO + 16 * x), w.r, 4 * gi, t[0], t[1], t[2], t[3]);
      }
  // norms n_j = sqrt(sum_x H^2 + eps) + eps  (comp/__init__.py:381-392) -> Z[:, zc0 + j]
#pragma unroll
  for (int j = 0; j < DCOL; ++j)
    if (j < g.hd && (j % CS) == w.part) {
      const float a = hdv[0][j], b = hdv[1][j], c = hdv[2][j];
      put1(Z, zlo, w.r, g.zc0 + j, sqrtf(fmaf(a, a, fmaf(b, b, c * c)) + SAFE_NORM_EPS) + SAFE_NORM_EPS);
    }
  // frame scalars q[3c + a] = sum_x F[a][x] * D[x][c]  (scalarize, comp/__init__.py:302-312)
  const float* F = sm + p.FBUF + w.r * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t)
    if (((t + 1) % CS) == w.part) {
      const int cc = t / 3, a = t - 3 * cc;
      float q = F[3 * a] * hdv[0][DCOL + cc];
      q = fmaf(F[3 * a + 1], hdv[1][DCOL + cc], q);
      q = fmaf(F[3 * a + 2], hdv[2][DCOL + cc], q);
      if (p.e3 && a == 1) q = fabsf(q);
      put1(Z, zlo, w.r, g.zc0 + g.hd + t, q);
    }
  // zero padding up to the GEMM's K
  for (int c = g.zc0 + g.hd + 9 + w.part; c < g.kz; c += CS) put1(Z, zlo, w.r, c, 0.f);
}

// ---- epilogue B: scalar_out accumulator -> gate operand act_v(T) (T tile), new scalar state (Z tile), messages ---------
template <int CS>
__device__ __forceinline__ void epilogue_b(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w, const float* smc,
                                           bool add, bool last, long long q, bool live) {
  float* Z = sm + p.ZBUF;
  float* T = sm + p.TBUF;
  const float* bs = smc + g.o_bs;
  const int ngrp = g.sop >> 4;
  for (int cg = w.part; cg < ngrp; cg += CS) {
    float t[16];
    tmem_ld16(w.tl + (uint32_t)(p.TACC + 16 * cg), t);
    wait_ld();
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const int c = 16 * cg + 4 * i4;
      if (c < g.gk || c < p.s) {  // CTA-uniform
        float av[4], sv[4];
        const float4 old = (add && c < p.s) ? get4(Z, w.r, c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float o4[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = c + i;
          const float tv = n < g.so ? t[4 * i4 + i] + bs[n] : 0.f;
          av[i] = n < g.so ? act_fwd(g.act_v, tv, p.slope) : 0.f;
          sv[i] = n < g.so ? o4[i] + act_fwd(g.act_s, tv, p.slope) : 0.f;
        }
        if (c < g.gk) put4(T, w.tl + (uint32_t)p.TLO, w.r, c, av[0], av[1], av[2], av[3]);
        if (c < p.s) {
          put4(Z, w.tl + (uint32_t)p.ZLO, w.r, c, sv[0], sv[1], sv[2], sv[3]);
          if (last && live) *reinterpret_cast<float4*>(p.msg + q * (p.s + 3 * p.v) + c) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        }
      }
    }
  }
}

// ---- epilogue C: gate + vector_up accumulators -> new vector state (V tile), messages ---------------------------------
template <int CS>
__device__ __forceinline__ void epilogue_c(const TcEdgeParams& p, const TcGcp& g, float* sm, const Who& w, const float* smc,
                                           bool add, bool last, long long q, bool live) {
  float* V = sm + p.VBUF;
  const float* bg = smc + g.o_bg;
  const int ngrp = (g.vo + 3) >> 2;
  for (int gi = w.part; gi < (PW >> 2); gi += CS) {
    float gt[4], u[3][4];
    tmem_ld4(w.tl + (uint32_t)(p.GACC + 4 * gi), gt);
#pragma unroll
    for (int x = 0; x < 3; ++x) tmem_ld4(w.tl + (uint32_t)(p.UACC + 16 * x + 4 * gi), u[x]);
    wait_ld();
    float nv[3][4];
    float4 old[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) old[x] = (add || g.vres) ? get4(V + x * PLANE, w.r, 4 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = 4 * gi + i;
      const float sg = (gi < ngrp && o < g.vo) ? sigmoidf_(gt[i] + bg[o]) : 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const float ov = i == 0 ? old[x].x : (i == 1 ? old[x].y : (i == 2 ? old[x].z : old[x].w));
        float uu = u[x][i];
        if (g.vres) uu += ov;                       // U + V_in before gating (gcpnet.py:364-367)
        nv[x][i] = (add ? ov : 0.f) + uu * sg;      // residual message stack (gcpnet.py:920-924)
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) put4(V + x * PLANE, w.tl + (uint32_t)(p.VLO + 16 * x), w.r, 4 * gi, nv[x][0], nv[x][1], nv[x][2], nv[x][3]);
    if (last && live && gi < ngrp) {
      float* mp = p.msg + q * (p.s + 3 * p.v) + p.s + 12 * gi;  // [channel][xyz], xyz fastest
      float f[12];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int x = 0; x < 3; ++x) f[3 * i + x] = nv[x][i];
      if (4 * gi + 4 <= g.vo) {
#pragma unroll
        for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(mp + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      } else {
        for (int j = 0; j < 3 * (g.vo - 4 * gi); ++j) mp[j] = f[j];
      }
    }
  }
}

// gather `len` contiguous floats (len % 4 == 0) of a global row into columns [c0, c0+len) of a slab tile (+ lo part)
template <int CS>
__device__ __forceinline__ void gather_row(float* tile, uint32_t tlo, const Who& w, int c0, const float* src, int len, bool live) {
  for (int c = 4 * w.part; c < len; c += 4 * CS) {
    const float4 v = live ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    put4(tile, tlo, w.r, c0 + c, v.x, v.y, v.z, v.w);
  }
}
// gather a [channels][xyz] global row into three 16-column planes (channels beyond nch -> zero, up to `cols` columns)
template <int CS>
__device__ __forceinline__ void gather_planes(float* planes, uint32_t tlo, const Who& w, const float* src, int nch, int cols, bool live) {
  for (int gi = w.part; 4 * gi < cols; gi += CS) {
    float f[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) f[j] = 0.f;
    if (live) {
      if (4 * gi + 4 <= nch) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + 12 * gi + 4 * j));
          f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
        }
      } else {
        for (int j = 0; j < 12; ++j) if (4 * gi + j / 3 < nch) f[j] = __ldg(src + 12 * gi + j);
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) put4(planes + x * PLANE, tlo + (uint32_t)(16 * x), w.r, 4 * gi, f[x], f[3 + x], f[6 + x], f[9 + x]);
  }
}

__global__ void __launch_bounds__(256) tc_pack_kernel(const __grid_constant__ TcPackProg prog) {
  const TcPackItem& it = prog.it[blockIdx.x];
  float* blob = prog.blob;
  if (it.kind == 2) {
    for (int i = threadIdx.x; i < it.R; i += 256) blob[it.dst_hi + i] = i < it.nreal ? __ldg(it.w + i) : 0.f;
    return;
  }
  const int total = it.R * it.C;
  for (int idx = threadIdx.x; idx < total; idx += 256) {
    const int n = idx % it.R, kk = idx / it.R;
    float val = 0.f;
    for (int rg = 0; rg < it.nrange; ++rg) {
      if (kk >= it.tc0[rg] && kk < it.tc0[rg] + it.len[rg]) {
        const int col = it.rc0[rg] + kk - it.tc0[rg];
        if (it.kind == 0) { if (n < it.nreal) val = __ldg(it.w + (size_t)n * it.ldw + col); }
        else if (n < it.hd) val = __ldg(it.w + (size_t)n * it.ldw + col);
        else if (n >= DCOL && n < DCOL + 3) val = __ldg(it.w2 + (size_t)(n - DCOL) * it.ldw + col);
      }
    }
    const int off = ((kk >> 2) * it.R + n) * 4 + (kk & 3);
    blob[it.dst_hi + off] = val;
    blob[it.dst_lo + off] = umma::tf32_lo(val);
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
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  if (w.tid == 0) {
    mbar_init(mma_bar, 1);
    for (int i = 0; i < p.ring_s.nslot; ++i) mbar_init(&rs.bar[i], 1);
    for (int i = 0; i < p.ring_w.nslot; ++i) mbar_init(&rw.bar[i], 1);
    mbar_fence_init();
    fence_async_smem();
    ring_fill(rs);
    ring_fill(rw);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  w.tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t mma_n = 0;  // commits on mma_bar so far (CTA-uniform)
  const int W = p.s + 3 * p.v;
  float* Z = sm + p.ZBUF; float* T = sm + p.TBUF; float* V = sm + p.VBUF; float* H = sm + p.HBUF;
  const uint32_t idesc_dn = make_idesc(128, 16, 0, 0);

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

    for (int k = 0; k < p.L; ++k) {
      const TcGcp& g = p.g[k];
      const bool last = k == p.L - 1;
      const bool add = k > 0 && p.residual;
      // ---------------- vector_down operands
      if (k == 0) {
        // chi_row planes -> Z tile, chi_col planes -> T tile, xi planes -> V tile; frames
        gather_planes<CS>(Z, w.tl + (uint32_t)p.ZLO, w, p.chi + (size_t)src * 3 * p.v, p.v, g.vkc[0], live);
        gather_planes<CS>(V, w.tl + (uint32_t)p.VLO, w, p.xi + (size_t)orig * 3 * p.ve, p.ve, g.vkc[1], live);
        gather_planes<CS>(T, w.tl + (uint32_t)p.TLO, w, p.chi + (size_t)dst * 3 * p.v, p.v, g.vkc[2], live);
        if (w.part == 0) {
          float* F = sm + p.FBUF + w.r * 9;
#pragma unroll
          for (int i = 0; i < 9; ++i) F[i] = live ? __ldg(p.frames + (size_t)orig * 9 + i) : 0.f;
        }
      }
      publish_and_sync();
      const float* smc = ring_wait(rs, rs.head);  // small chunk of this GCP (all threads: biases are read in the epilogues)
      if (w.tid == 0) {
        bulk_store_wait_read();
        ring_fill(rs);  // every thread is past the previous GCP's epilogue C (barrier above): its small chunk's slot is free
        bool acc;
        for (int x = 0; x < 3; ++x) {
          acc = false;
          if (k == 0) {
            mma3(tbase + (uint32_t)(p.HDACC + 16 * x), Z + x * PLANE, tbase + (uint32_t)(p.ZLO + 16 * x), smc + g.o_wd_hi[0], smc + g.o_wd_lo[0], 16, g.vkc[0], idesc_dn, acc);
            mma3(tbase + (uint32_t)(p.HDACC + 16 * x), V + x * PLANE, tbase + (uint32_t)(p.VLO + 16 * x), smc + g.o_wd_hi[1], smc + g.o_wd_lo[1], 16, g.vkc[1], idesc_dn, acc);
            mma3(tbase + (uint32_t)(p.HDACC + 16 * x), T + x * PLANE, tbase + (uint32_t)(p.TLO + 16 * x), smc + g.o_wd_hi[2], smc + g.o_wd_lo[2], 16, g.vkc[2], idesc_dn, acc);
          } else {
            mma3(tbase + (uint32_t)(p.HDACC + 16 * x), V + x * PLANE, tbase + (uint32_t)(p.VLO + 16 * x), smc + g.o_wd_hi[0], smc + g.o_wd_lo[0], 16, g.vkc[0], idesc_dn, acc);
          }
        }
        commit(mma_bar);
        if (k > 0 && saved_t) bulk_store(saved_t + (size_t)(k - 1) * (p.s_img + p.v_img) + p.s_img, V, p.v_img);  // V_{k-1}
      }
      wait_mma();
      // ---------------- epilogue A (+ the gathers GCP 0's first two scalar_out segments need)
      epilogue_a<CS>(p, g, sm, w);
      if (k == 0) {
        gather_row<CS>(Z, w.tl + (uint32_t)p.ZLO, w, 0, p.e + (size_t)orig * p.se, p.se, live);
        gather_row<CS>(T, w.tl + (uint32_t)p.TLO, w, 0, p.h + (size_t)src * p.s, p.s, live);
        for (int c = p.s + 4 * w.part; c < g.seg[1].kc; c += 4 * CS) put4(T, w.tl + (uint32_t)p.TLO, w.r, c, 0.f, 0.f, 0.f, 0.f);
      }
      publish_and_sync();
      // ---------------- scalar_out (one batch per K-segment) + vector_up (with the first batch)
      const uint32_t idesc_s = make_idesc(128, g.sop, 0, 0), idesc_v = make_idesc(128, g.vop, 0, 0);
      bool tacc = false;
      for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const TcSeg sg = g.seg[sgi];
        if (sg.a_tile == 2) {  // GCP 0: the T tile now takes h_col
          gather_row<CS>(T, w.tl + (uint32_t)p.TLO, w, 0, p.h + (size_t)dst * p.s, p.s, live);
          publish_and_sync();
        }
        if (w.tid == 0) {
          bulk_store_wait_read();
          ring_fill(rw);
          const float* bh = ring_wait(rw, rw.head);
          const float* bl = ring_wait(rw, rw.head + 1);
          mma3(tbase + (uint32_t)p.TACC, sg.a_tile == 0 ? Z : T, tbase + (uint32_t)(sg.a_tile == 0 ? p.ZLO : p.TLO), bh, bl, g.sop, sg.kc, idesc_s, tacc);
          if (sgi == 0) {
            for (int x = 0; x < 3; ++x) {
              bool acc = false;
              mma3(tbase + (uint32_t)(p.UACC + 16 * x), H + x * PLANE, tbase + (uint32_t)(p.HLO + 16 * x), smc + g.o_wu_hi, smc + g.o_wu_lo, g.vop, g.hdp, idesc_v, acc);
            }
          }
          commit(mma_bar);
        }
        tacc = true;
        wait_mma();
        rw.head += 2;
      }
      // ---------------- epilogue B
      epilogue_b<CS>(p, g, sm, w, smc, add, last, q, live);
      publish_and_sync();
      // ---------------- gate
      if (w.tid == 0) {
        bulk_store_wait_read();
        ring_fill(rw);
        bool acc = false;
        mma3(tbase + (uint32_t)p.GACC, T, tbase + (uint32_t)p.TLO, smc + g.o_wg_hi, smc + g.o_wg_lo, g.vop, g.gk, idesc_v, acc);
        commit(mma_bar);
        if (!last && saved_t) bulk_store(saved_t + (size_t)k * (p.s_img + p.v_img), Z, p.s_img);  // S_k
      }
      wait_mma();
      // ---------------- epilogue C
      epilogue_c<CS>(p, g, sm, w, smc, add, last, q, live);
      rs.head += 1;
      // (next tile: every thread rewrites only its own row of Z / T / V; all tensor-core readers of this tile have
      //  completed and thread 0 has waited for the bulk-store reads inside the issue sections above)
    }
  }
  if (w.tid == 0) bulk_store_wait_all();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, (uint32_t)p.tmem_cols);
}


}  // namespace tc
}  // namespace gcp
