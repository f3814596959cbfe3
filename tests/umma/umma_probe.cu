// umma_probe.cu -- bring-up / known-answer tests of gcpnet_b200/csrc/umma.cuh on a real B200
// (test infrastructure; built and run by tests/test_gpu_umma.py and scripts).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu
//   ./umma_probe <case>      cases: kmajor, kmajor3x, mnmajor64, mnmajor128, n16, timing
//
// Each case stages slab-layout operand tiles (hi and lo parts) in shared memory, issues the
// tcgen05.mma sequence from one thread, commits to an mbarrier, reads the accumulator back with
// tcgen05.ld and compares with a CPU evaluation.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "../../gcpnet_b200/csrc/umma.cuh"

using namespace gcp::umma;

struct ProbeParams {
  int M, N, ksteps, passes;  // passes: 1 = hi*hi only, 3 = 3xTF32
  int a_mn, b_mn;            // operand major-ness
  int Ra, Rb;                // rows of the A / B slab tiles
  int a_step, b_step;        // floats to advance per k-step
  int a_floats, b_floats;    // tile sizes
  int ncols;                 // accumulator columns to read back
  int reps;                  // timing: repeat the whole MMA batch
  int ts;                    // 1: the lo pass of A comes from TMEM (columns 64..) instead of shared memory
  int swap_mn;               // MN-major descriptors: swap the LBO / SBO roles (bring-up experiment)
  const float *a_hi, *a_lo, *b_hi, *b_lo;
  float* d;                  // [128 lanes][ncols]
  long long* cycles;         // [reps]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeParams p) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  float* a_hi = smem;
  float* a_lo = a_hi + p.a_floats;
  float* b_hi = a_lo + p.a_floats;
  float* b_lo = b_hi + p.b_floats;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < p.a_floats; i += 128) { a_hi[i] = p.a_hi[i]; a_lo[i] = p.a_lo[i]; }
  for (int i = tid; i < p.b_floats; i += 128) { b_hi[i] = p.b_hi[i]; b_lo[i] = p.b_lo[i]; }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = make_idesc(p.M, p.N, p.a_mn, p.b_mn);
  if (p.ts) {  // thread = row: copy the row of A_lo into TMEM columns [64, 64 + 8*ksteps)
    for (int c = 0; c < 8 * p.ksteps; c += 4) {
      const float* src = a_lo + slab_off(p.Ra, tid, c);
      tmem_st4(tmem_at(tbase, 32 * warp, 64 + c), src[0], src[1], src[2], src[3]);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  for (int rep = 0; rep < p.reps; ++rep) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      bool acc = false;
      if (p.ts == 2) {  // tight issue loop: descriptors advanced by one add, lo pass of A from TMEM
        const uint64_t ad0 = make_desc(smem_addr(a_hi), (uint32_t)p.Ra * 16u, 128u);
        const uint64_t bh0 = make_desc(smem_addr(b_hi), (uint32_t)p.Rb * 16u, 128u);
        const uint64_t bl0 = make_desc(smem_addr(b_lo), (uint32_t)p.Rb * 16u, 128u);
        const uint32_t astep = (uint32_t)p.a_step * 4u, bstep = (uint32_t)p.b_step * 4u;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t ad = desc_advance(ad0, ks * astep), bh = desc_advance(bh0, ks * bstep), bl = desc_advance(bl0, ks * bstep);
          mma_tf32(tbase, ad, bh, idesc, acc);
          mma_tf32(tbase, ad, bl, idesc, true);
          mma_tf32_ts(tbase, tmem_at(tbase, 0, 64 + 8 * ks), bh, idesc, true);
          acc = true;
        }
      } else
      for (int pass = 0; pass < p.passes; ++pass) {
        const float* A = pass == 1 ? a_lo : a_hi;
        const float* B = pass == 2 ? b_lo : b_hi;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          if (p.ts && pass == 1) {
            const uint64_t bd = make_desc(smem_addr(B + ks * p.b_step), (uint32_t)p.Rb * 16u, 128u);
            mma_tf32_ts(tbase, tmem_at(tbase, 0, 64 + 8 * ks), bd, idesc, acc);
            continue;
          }
          const uint64_t ad = p.a_mn ? (p.swap_mn ? make_desc(smem_addr(A + ks * p.a_step), (uint32_t)p.Ra * 16u, 128u)
                                                  : make_desc(smem_addr(A + ks * p.a_step), 128u, (uint32_t)p.Ra * 16u))
                                     : make_desc(smem_addr(A + ks * p.a_step), (uint32_t)p.Ra * 16u, 128u);
          const uint64_t bd = p.b_mn ? (p.swap_mn ? make_desc(smem_addr(B + ks * p.b_step), (uint32_t)p.Rb * 16u, 128u)
                                                  : make_desc(smem_addr(B + ks * p.b_step), 128u, (uint32_t)p.Rb * 16u))
                                     : make_desc(smem_addr(B + ks * p.b_step), (uint32_t)p.Rb * 16u, 128u);
          mma_tf32(tbase, ad, bd, idesc, acc);
          acc = true;
        }
      }
      commit(&bar);
    }
    mbar_wait(&bar, (uint32_t)(rep & 1));
    fence_after_sync();
    if (tid == 0) p.cycles[rep] = clock64() - t0;
    __syncthreads();
  }
  // read back: warp w owns lanes 32w .. 32w+31
  for (int c0 = 0; c0 < p.ncols; c0 += 8) {
    float v[8];
    tmem_ld8(tmem_at(tbase, 32 * warp, c0), v);
    wait_ld();
    for (int j = 0; j < 8; ++j) p.d[(size_t)tid * p.ncols + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

// logical X[R][C] -> slab layout hi / lo
static void to_slab(const std::vector<float>& X, int R, int C, std::vector<float>& hi, std::vector<float>& lo) {
  hi.assign((size_t)R * C, 0.f); lo.assign((size_t)R * C, 0.f);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      const float x = X[(size_t)r * C + c];
      hi[slab_off(R, r, c)] = x;            // the tensor core truncates by itself
      lo[slab_off(R, r, c)] = tf32_lo(x);
    }
}

static int run_case(const std::string& name) {
  srand(1234);
  ProbeParams p{};
  p.reps = 1;
  std::vector<float> A, B;  // logical operands
  int Mrows = 0, Ncols = 0, K = 0;
  bool mn = false;
  if (name == "kmajor" || name == "kmajor3x" || name == "n16" || name == "timing" || name == "ts" || name == "tsfast") {
    p.ts = name == "ts" ? 1 : (name == "tsfast" ? 2 : 0);
    Mrows = 128; Ncols = name == "n16" ? 16 : 64; K = name == "kmajor" ? 16 : 80;
    p.passes = name == "kmajor" ? 1 : 3;
    if (name == "timing" || name == "tsfast") p.reps = 6;
    A.resize((size_t)Mrows * K); B.resize((size_t)Ncols * K);
    for (auto& x : A) x = frand();
    for (auto& x : B) x = frand();
    p.M = 128; p.N = Ncols; p.ksteps = K / 8; p.a_mn = 0; p.b_mn = 0; p.Ra = Mrows; p.Rb = Ncols;
    p.a_step = 2 * Mrows * 4; p.b_step = 2 * Ncols * 4;
  } else if (name.rfind("mnmajor", 0) == 0) {
    p.swap_mn = name.find("swap") != std::string::npos;
    // D[j][i] = sum_e G[e][j] * Z[e][i];  G[128 edges][J], Z[128 edges][80]
    mn = true;
    const int J = name.find("64") != std::string::npos ? 64 : 128;
    Mrows = J; Ncols = 80; K = 128;
    p.passes = 3;
    A.resize((size_t)K * J); B.resize((size_t)K * Ncols);  // stored [edge][feature]
    for (auto& x : A) x = frand();
    for (auto& x : B) x = frand();
    p.M = J; p.N = Ncols; p.ksteps = K / 8; p.a_mn = 1; p.b_mn = 1; p.Ra = K; p.Rb = K;
    p.a_step = 8 * 4; p.b_step = 8 * 4;  // 8 rows further down inside every slab
  } else {
    printf("unknown case %s\n", name.c_str());
    return 2;
  }
  std::vector<float> a_hi, a_lo, b_hi, b_lo;
  if (!mn) { to_slab(A, Mrows, K, a_hi, a_lo); to_slab(B, Ncols, K, b_hi, b_lo); }
  else { to_slab(A, K, Mrows, a_hi, a_lo); to_slab(B, K, Ncols, b_hi, b_lo); }
  p.a_floats = (int)a_hi.size(); p.b_floats = (int)b_hi.size();
  p.ncols = Ncols;
  float *da_hi, *da_lo, *db_hi, *db_lo, *dd;
  long long* dcyc;
  CK(cudaMalloc(&da_hi, a_hi.size() * 4)); CK(cudaMalloc(&da_lo, a_hi.size() * 4));
  CK(cudaMalloc(&db_hi, b_hi.size() * 4)); CK(cudaMalloc(&db_lo, b_hi.size() * 4));
  CK(cudaMalloc(&dd, 128 * Ncols * 4)); CK(cudaMalloc(&dcyc, 64 * 8));
  CK(cudaMemcpy(da_hi, a_hi.data(), a_hi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(da_lo, a_lo.data(), a_lo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db_hi, b_hi.data(), b_hi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db_lo, b_lo.data(), b_lo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, 128 * Ncols * 4));
  p.a_hi = da_hi; p.a_lo = da_lo; p.b_hi = db_hi; p.b_lo = db_lo; p.d = dd; p.cycles = dcyc;
  const int bytes = (2 * p.a_floats + 2 * p.b_floats) * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  probe_kernel<<<1, 128, bytes>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)128 * Ncols);
  std::vector<long long> cyc(64);
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc.data(), dcyc, 64 * 8, cudaMemcpyDeviceToHost));
  // reference
  double max_err_exact = 0, max_err_trunc = 0, max_ref = 0;
  for (int m = 0; m < Mrows; ++m) {
    const int lane = (p.M == 64) ? (m % 16) + 32 * (m / 16) : m;
    for (int n = 0; n < Ncols; ++n) {
      double exact = 0, trunc = 0;
      for (int k = 0; k < K; ++k) {
        const float a = mn ? A[(size_t)k * Mrows + m] : A[(size_t)m * K + k];
        const float b = mn ? B[(size_t)k * Ncols + n] : B[(size_t)n * K + k];
        exact += (double)a * (double)b;
        trunc += (double)tf32_trunc(a) * (double)tf32_trunc(b);
      }
      const double got = D[(size_t)lane * Ncols + n];
      max_err_exact = fmax(max_err_exact, fabs(got - exact));
      max_err_trunc = fmax(max_err_trunc, fabs(got - trunc));
      max_ref = fmax(max_ref, fabs(exact));
    }
  }
  if (getenv("PROBE_DUMP")) {
    for (int m = 0; m < 4; ++m) {
      const int lane = (p.M == 64) ? (m % 16) + 32 * (m / 16) : m;
      for (int n = 0; n < 6; ++n) {
        double exact = 0;
        for (int k = 0; k < K; ++k)
          exact += (double)(mn ? A[(size_t)k * Mrows + m] : A[(size_t)m * K + k]) * (double)(mn ? B[(size_t)k * Ncols + n] : B[(size_t)n * K + k]);
        printf("  d[%d][%d] got % .5f want % .5f |", m, n, D[(size_t)lane * Ncols + n], exact);
      }
      printf("\n");
    }
  }
  const double rel_exact = max_err_exact / max_ref, rel_trunc = max_err_trunc / max_ref;
  const double bar = p.passes == 3 ? rel_exact : rel_trunc;
  const double tol = p.passes == 3 ? 2e-6 : 2e-6;
  printf("case %-10s M=%d N=%d K=%d passes=%d : rel err vs exact %.3e, vs tf32-truncated %.3e -> %s\n", name.c_str(), p.M, p.N, K,
         p.passes, rel_exact, rel_trunc, bar < tol ? "PASS" : "FAIL");
  if (p.reps > 1) {
    printf("  cycles per batch of %d MMAs (issue -> commit -> mbarrier wake-up):", p.passes * p.ksteps);
    for (int i = 0; i < p.reps; ++i) printf(" %lld", cyc[i]);
    printf("\n");
  }
  return bar < tol ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: umma_probe <case>\n"); return 2; }
  return run_case(argv[1]);
}
