// layernorm.cuh -- GCPLayerNorm (comp/__init__.py:138-167) as standalone kernels: the pre_norm layers apply gcp_norm.0 to the
// layer input before the message passing (gcpnet.py:1188-1189), and GCPEmbedding normalises its inputs / outputs
// (gcpnet.py:727-733,815-817).  Scalars: nn.LayerNorm(s) (eps 1e-5, affine); vectors: v / sqrt(mean_c max(|v_c|^2, eps)).
// One warp per node; the weight / bias gradients are fixed-order sums (node chunks -> partial rows -> reduce): deterministic.
#pragma once
#include <cuda_runtime.h>

namespace gcp {

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// out = GCPLayerNorm(h, chi).  s == 0 or v == 0: that part is absent (scalar-only / vector-free inputs).
__global__ void __launch_bounds__(256) gcp_layernorm_fwd_kernel(const float* __restrict__ h, const float* __restrict__ chi, int N, int s, int v,
                                                                const float* __restrict__ w, const float* __restrict__ b, float ln_eps,
                                                                float vn_eps, float* __restrict__ out_h, float* __restrict__ out_chi) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= N) return;
  if (s > 0) {
    const float* x = h + (size_t)i * s;
    float sum = 0.f;
    for (int j = lane; j < s; j += 32) sum += x[j];
    const float mean = warp_sum(sum) / (float)s;
    float var = 0.f;
    for (int j = lane; j < s; j += 32) { const float d = x[j] - mean; var = fmaf(d, d, var); }
    const float rstd = 1.f / sqrtf(warp_sum(var) / (float)s + ln_eps);
    for (int j = lane; j < s; j += 32) out_h[(size_t)i * s + j] = fmaf((x[j] - mean) * rstd, __ldg(w + j), __ldg(b + j));
  }
  if (v > 0) {
    const float* vp = chi + (size_t)i * 3 * v;
    float m = 0.f;
    for (int c = lane; c < v; c += 32) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
    }
    const float inv = 1.f / sqrtf(warp_sum(m) / (float)v);
    for (int c = lane; c < 3 * v; c += 32) out_chi[(size_t)i * 3 * v + c] = vp[c] * inv;
  }
}

// data gradient + per-node statistics (mean, rstd) for the weight-gradient pass
__global__ void __launch_bounds__(256) gcp_layernorm_bwd_kernel(const float* __restrict__ h, const float* __restrict__ chi, int N, int s, int v,
                                                                const float* __restrict__ w, float ln_eps, float vn_eps,
                                                                const float* __restrict__ gy_h, const float* __restrict__ gy_chi,
                                                                float* __restrict__ g_h, float* __restrict__ g_chi, float* __restrict__ stats) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= N) return;
  if (s > 0) {
    const float* x = h + (size_t)i * s;
    const float* gy = gy_h + (size_t)i * s;
    float sum = 0.f;
    for (int j = lane; j < s; j += 32) sum += x[j];
    const float mean = warp_sum(sum) / (float)s;
    float var = 0.f;
    for (int j = lane; j < s; j += 32) { const float d = x[j] - mean; var = fmaf(d, d, var); }
    const float rstd = 1.f / sqrtf(warp_sum(var) / (float)s + ln_eps);
    float m1 = 0.f, m2 = 0.f;
    for (int j = lane; j < s; j += 32) {
      const float gxh = gy[j] * __ldg(w + j);
      m1 += gxh; m2 = fmaf(gxh, (x[j] - mean) * rstd, m2);
    }
    m1 = warp_sum(m1) / (float)s; m2 = warp_sum(m2) / (float)s;
    for (int j = lane; j < s; j += 32) {
      const float xhat = (x[j] - mean) * rstd;
      g_h[(size_t)i * s + j] = rstd * (gy[j] * __ldg(w + j) - m1 - xhat * m2);
    }
    if (lane == 0) { stats[2 * (size_t)i] = mean; stats[2 * (size_t)i + 1] = rstd; }
  }
  if (v > 0) {
    const float* vp = chi + (size_t)i * 3 * v;
    const float* gv = gy_chi + (size_t)i * 3 * v;
    float m = 0.f, dot = 0.f;
    for (int c = lane; c < v; c += 32) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      m += n2 > vn_eps ? n2 : vn_eps;
      dot = fmaf(gv[3 * c], vp[3 * c], fmaf(gv[3 * c + 1], vp[3 * c + 1], fmaf(gv[3 * c + 2], vp[3 * c + 2], dot)));
    }
    m = warp_sum(m); dot = warp_sum(dot);
    const float rr = sqrtf(m / (float)v);
    const float coef = dot / ((float)v * rr * rr * rr);
    for (int c = lane; c < v; c += 32) {
      const float n2 = fmaf(vp[3 * c], vp[3 * c], fmaf(vp[3 * c + 1], vp[3 * c + 1], vp[3 * c + 2] * vp[3 * c + 2]));
      const float ind = n2 > vn_eps ? 1.f : 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x) g_chi[(size_t)i * 3 * v + 3 * c + x] = gv[3 * c + x] / rr - coef * ind * vp[3 * c + x];
    }
  }
}

// partial[part][j] = sum over the part's nodes of gy * xhat, partial[part][s + j] = sum of gy   (blockIdx.x = part)
__global__ void __launch_bounds__(128) gcp_layernorm_wgrad_kernel(const float* __restrict__ h, const float* __restrict__ gy_h,
                                                                  const float* __restrict__ stats, int N, int s, float* __restrict__ partial) {
  const int parts = gridDim.x, chunk = (N + parts - 1) / parts;
  const int i0 = blockIdx.x * chunk, i1 = min(N, i0 + chunk);
  for (int j = threadIdx.x; j < s; j += blockDim.x) {
    float gw = 0.f, gb = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float gy = __ldg(gy_h + (size_t)i * s + j);
      gw = fmaf(gy, (__ldg(h + (size_t)i * s + j) - __ldg(stats + 2 * (size_t)i)) * __ldg(stats + 2 * (size_t)i + 1), gw);
      gb += gy;
    }
    partial[(size_t)blockIdx.x * 2 * s + j] = gw;
    partial[(size_t)blockIdx.x * 2 * s + s + j] = gb;
  }
}
__global__ void gcp_layernorm_wreduce_kernel(const float* __restrict__ partial, int parts, int s, float* __restrict__ g_w, float* __restrict__ g_b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * s) return;
  float acc = 0.f;
  for (int p = 0; p < parts; ++p) acc += partial[(size_t)p * 2 * s + j];
  if (j < s) g_w[j] = acc; else g_b[j - s] = acc;
}

}  // namespace gcp
