// Throughput probe: legacy mma.sync.m16n8k8 tf32 vs FFMA, per SM, as a function of resident warps (dev tool).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void hmma_kernel(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  float c[4][4] = {};
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ffma_kernel(float* out, long long* cyc, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  float c[16] = {};
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = fmaf(a, b, c[j]);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    long long h = 0;
    hmma_kernel<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    hmma_kernel<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sm = (double)h / ((double)iters * 4 * warps);
    long long f = 0;
    ffma_kernel<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    ffma_kernel<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    cudaMemcpy(&f, cyc, 8, cudaMemcpyDeviceToHost);
    const double ffma_per_clk = (double)iters * 16 * 32 * warps / (double)f;
    printf("warps/SM %2d: mma.sync tf32 m16n8k8 %.2f cycles per instruction per SM (%.0f MAC/clk/SM); FFMA %.1f lanes/clk/SM\n",
           warps, per_sm, 1024.0 / per_sm, ffma_per_clk);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
