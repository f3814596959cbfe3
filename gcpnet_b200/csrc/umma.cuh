// umma.cuh -- thin inline-PTX layer over the Blackwell (sm_100a) 5th-generation tensor cores:
// tcgen05.mma kind::tf32 with shared-memory operands and TMEM accumulators, TMEM allocation,
// tcgen05.ld / tcgen05.st, commit -> mbarrier, and the fences that go with them.
//
// Operand layout used throughout this repo ("slab layout", no swizzle): a tile X[R rows][C cols] of
// fp32 lives in shared memory as   X[c / 4][r][c % 4]   -- for every group of four columns one slab
// of R x 16 bytes.  Properties:
//   * a thread that owns a row writes its four consecutive columns with ONE 128-bit store, and the
//     32 threads of a warp (32 consecutive rows) cover 512 contiguous bytes -> conflict free;
//   * as a K-major operand (rows = M or N, columns = K) it is the canonical INTERLEAVE layout with
//     core matrices of 8 rows x 16 B:  SBO (next 8 rows) = 128 B,  LBO (next 4 columns) = R * 16 B;
//   * the SAME bytes are a valid MN-major operand with rows = K and columns = M or N (what the
//     weight-gradient GEMMs need: the reduction runs over the tile's rows): core matrix = 8 rows
//     x 4 columns,  SBO (next 4 columns) = R * 16 B,  LBO (next 8 rows) = 128 B.
// Descriptor bit layout: cute/arch/mma_sm100_desc.hpp (CUTLASS 4.x), restated here.
//
// 3xTF32: the tensor core reads the top 19 bits of each fp32 word (tf32 = truncation), so the "hi"
// operand is the fp32 tile itself; the "lo" tile holds rna_tf32(x - trunc_tf32(x)).  A product is
// evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (error ~2^-21 per product).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcp {
namespace umma {

constexpr uint32_t TF32_MASK = 0xffffe000u;

__host__ __device__ __forceinline__ float tf32_trunc(float x) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__float_as_uint(x) & TF32_MASK);
#else
  union { float f; uint32_t u; } v; v.f = x; v.u &= TF32_MASK; return v.f;
#endif
}
// residual of the truncation, itself rounded to tf32 (round to nearest)
__host__ __device__ __forceinline__ float tf32_lo(float x) {
  const float r = x - tf32_trunc(x);
#if defined(__CUDA_ARCH__)
  uint32_t o;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(r));
  return __uint_as_float(o);
#else
  union { float f; uint32_t u; } v; v.f = r;
  v.u = (v.u + 0x1000u) & TF32_MASK;  // round half away (matches cvt.rna for finite values)
  return v.f;
#endif
}

// offset (floats) of element (r, c) of a slab-layout tile with R rows
__host__ __device__ __forceinline__ int slab_off(int R, int r, int c) { return ((c >> 2) * R + r) * 4 + (c & 3); }

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptors (SWIZZLE_NONE, version 1) ---------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// K-major view of a slab tile with R rows, starting at row r0 and column c0 (multiple of 4):
// rows = the M (or N) index, columns = K.
__device__ __forceinline__ uint64_t desc_kmajor(const float* tile, int R, int r0, int c0) {
  return make_desc(smem_addr(tile + slab_off(R, r0, c0)), (uint32_t)R * 16u, 128u);
}
// MN-major view: columns = the M (or N) index (starting at c0, multiple of 4), rows = K (starting at r0).
__device__ __forceinline__ uint64_t desc_mnmajor(const float* tile, int R, int r0, int c0) {
  return make_desc(smem_addr(tile + slab_off(R, r0, c0)), 128u, (uint32_t)R * 16u);
}

// ---- instruction descriptor: kind::tf32, fp32 accumulate ----------------------------------------
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 step.  Issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// same with the A operand in TMEM (lanes = rows of A, one fp32 column per k): D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// descriptor of the same tile `bytes` further on (bytes multiple of 16, no carry out of the 14-bit address field)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_at(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }

// warp-collective: thread i of the warp reads columns [col, col+N) of TMEM lane (32*(warp%4) + i).
// `taddr` must carry the warp's lane base (32*(warp%4)) in its lane field.
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
  v[0] = __uint_as_float(a); v[1] = __uint_as_float(b); v[2] = __uint_as_float(c); v[3] = __uint_as_float(d);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// warp-collective: thread i writes columns [col, col+4) of its TMEM lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}

// one lane of a converged warp (warp-uniform control flow around it keeps descriptor math on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// value of lane 0, provably warp-uniform for the compiler
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- mbarrier helpers (shared::cta) ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
#endif  // __CUDACC__

}  // namespace umma
}  // namespace gcp
