// tc_api.cu -- translation unit of the tensor-core (tcgen05) kernels: instantiations + launchers.
#include <cuda_runtime.h>

#include <map>
#include <mutex>

#include "common.h"
#include "tc_setup.h"
#include "tc_edge_dev.cuh"

using namespace gcp;

constexpr int TC_CS = 4;  // threads per tile row

static int tc_set_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& cur = done[{kernel, dev}];
  if (bytes > cur) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cur = bytes;
  }
  return 0;
}

int gcp_tc_launch_pack(const tc::TcPackProg& prog, cudaStream_t st) {
  tc::tc_pack_kernel<<<dim3(prog.n, 32), 256, 0, st>>>(prog);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcp_tc_launch_node_pre(const tc::TcEdgeParams& p, float* P, float* Q, cudaStream_t st) {
  const long long total = (long long)((p.N + tc::PRE_NODES - 1) / tc::PRE_NODES) * (2 * p.pw + 192);  // PRE_NODES nodes per thread
  if (total <= 0) return 0;
  tc::tc_node_pre_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(p.h, p.chi, p.blob, p.nt, p.N, p.s, p.v, p.pw, P, Q);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcp_tc_launch_edge_fwd(const tc::TcEdgeParams& p, int grid, cudaStream_t st) {
  const int bytes = p.smem_floats * 4;
  if (tc_set_smem((const void*)tc::tc_edge_fwd_kernel<TC_CS>, bytes)) return 1;
  tc::tc_edge_fwd_kernel<TC_CS><<<grid, 128 * TC_CS, bytes, st>>>(p);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcp_tc_launch_edge_bwd(const tc::TcBwdParams& b, int grid, cudaStream_t st) {
  const int bytes = b.f.smem_floats * 4;
  if (tc_set_smem((const void*)tc::tc_edge_bwd_kernel<TC_CS>, bytes)) return 1;
  tc::tc_edge_bwd_kernel<TC_CS><<<grid, 128 * TC_CS, bytes, st>>>(b);
  gcp_note_launches(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// node-level finish of message GCP 0 + reduction of the partials + chain rule to the reference's parameters
// which: 1 = per-node sums, 2 = data gradient (dh, dchi), 4 = node-level weight-gradient partials
int gcp_tc_launch_post(const tc::TcPostParams& p, int which, cudaStream_t st) {
  const long long n1 = (long long)p.N * 2 * (p.pw + 96), n2 = (long long)p.N * (p.s + 3 * p.v);
  int n = 0;
  if ((which & 3) == 3 && p.s + 3 * p.v >= 48 && p.pw + 96 <= tc::POST_PER) {  // fused: sums in shared memory
    tc::tc_post_fused_kernel<<<(int)((n2 + 255) / 256), 256, 0, st>>>(p); ++n;
  } else {
    if (which & 1) { tc::tc_post_sum_kernel<<<(int)((n1 + 255) / 256), 256, 0, st>>>(p); ++n; }
    if (which & 2) { tc::tc_post_data_kernel<<<(int)((n2 + 255) / 256), 256, 0, st>>>(p); ++n; }
  }
  if (which & 4) { tc::tc_post_wgrad_kernel<<<dim3((p.npartial_stride + 255) / 256, p.nctas), 256, 0, st>>>(p); ++n; }
  gcp_note_launches(n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
// which: 1 = reduce the per-CTA partial rows of the edge backward (needs only that kernel), 2 = reduce the node-level
// partial rows (tc_post_wgrad_kernel) and apply the chain rule
int gcp_tc_launch_finalize(const float* partial, int rows, int stride, float* G, const float* npartial, int nrows, int nstride, float* Gn,
                           const tc::TcFinalParams& fp, int which, cudaStream_t st) {
  int n = 0;
  if (which & 1) { tc::tc_reduce_kernel<<<(stride + 255) / 256, 256, 0, st>>>(partial, rows, stride, stride, G); ++n; }
  if (which & 2) {
    tc::tc_reduce_kernel<<<(nstride + 255) / 256, 256, 0, st>>>(npartial, nrows, nstride, nstride, Gn);
    tc::tc_finalize_kernel<<<(fp.n_edge_params + 255) / 256, 256, 0, st>>>(fp);
    tc::tc_finalize_wg_kernel<<<(fp.L * fp.g[0].vo * fp.g[0].so * 32 + 255) / 256, 256, 0, st>>>(fp);
    n += 3;
  }
  gcp_note_launches(n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
