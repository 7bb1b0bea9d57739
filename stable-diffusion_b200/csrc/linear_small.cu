// Small-M linear layer on CUDA cores with fp32 activations: out[m, j] = act(x[m, :] . W[j, :] + b[j]).
// Used for the timestep-embedding MLP and the 22 ResBlock emb_layers (openaimodel.py:506-511, 217-223, 723-724,
// 267), where M = number of samples (<= 128) and the work is a weight-streaming GEMV: HBM/L2-bound on W
// (fp16, read once), fp32 accumulate, no tensor-core tile would be more than 1/64 full.
#include "../../include/sdb200.h"
#include "host.h"
#include <cuda_fp16.h>

namespace sdb {

constexpr int LS_ROWS = 8;

__global__ void __launch_bounds__(128)
    linear_small_kernel(const float* __restrict__ x, int M, int K, const __half* __restrict__ W, int N,
                        const float* __restrict__ bias, int act, float* __restrict__ out32,
                        __half* __restrict__ out16) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + warp;
  if (j >= N) return;
  const __half2* w2 = reinterpret_cast<const __half2*>(W + static_cast<size_t>(j) * K);
  const int K2 = K >> 1;
  for (int m0 = 0; m0 < M; m0 += LS_ROWS) {
    float acc[LS_ROWS];
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r) acc[r] = 0.f;
    for (int k2 = lane; k2 < K2; k2 += 32) {
      float2 w = __half22float2(w2[k2]);
#pragma unroll
      for (int r = 0; r < LS_ROWS; ++r) {
        if (m0 + r < M) {
          float2 xv = *reinterpret_cast<const float2*>(x + static_cast<size_t>(m0 + r) * K + 2 * k2);
          acc[r] = fmaf(w.x, xv.x, acc[r]);
          acc[r] = fmaf(w.y, xv.y, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < LS_ROWS; ++r) {
        if (m0 + r < M) {
          float v = acc[r] + (bias ? bias[j] : 0.f);
          if (act == SDB_ACT_SILU) v = v / (1.0f + __expf(-v));
          size_t o = static_cast<size_t>(m0 + r) * N + j;
          if (out32) out32[o] = v;
          if (out16) out16[o] = __float2half_rn(v);
        }
      }
    }
  }
}

__global__ void timestep_embedding_f32_kernel(const float* __restrict__ t, int n, int dim, float max_period,
                                              float* __restrict__ out) {
  int half_dim = dim / 2;
  int total = n * half_dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int j = i % half_dim, r = i / half_dim;
    float freq = expf(-logf(max_period) * static_cast<float>(j) / static_cast<float>(half_dim));
    float a = t[r] * freq;
    out[static_cast<size_t>(r) * dim + j] = cosf(a);
    out[static_cast<size_t>(r) * dim + half_dim + j] = sinf(a);
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_linear_small(const float* x, int32_t m, int32_t k, const void* w_f16, int32_t n, const float* bias,
                                int32_t act, float* out_f32, void* out_f16, sdb_stream_t stream) {
  SDB_CHECK(x && w_f16 && (out_f32 || out_f16), "sdb_linear_small: null pointer");
  SDB_CHECK(k % 2 == 0 && m > 0 && n > 0, "sdb_linear_small: bad sizes");
  SDB_CHECK(act == SDB_ACT_NONE || act == SDB_ACT_SILU, "sdb_linear_small: unsupported activation");
  linear_small_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, m, k, static_cast<const __half*>(w_f16), n, bias, act, out_f32, static_cast<__half*>(out_f16));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_timestep_embedding_f32(const float* t, int32_t n, int32_t dim, float max_period, float* out,
                                          sdb_stream_t stream) {
  SDB_CHECK(t && out && dim % 2 == 0, "sdb_timestep_embedding_f32: bad arguments");
  int total = n * dim / 2;
  timestep_embedding_f32_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, n, dim,
                                                                                                    max_period, out);
  SDB_LAUNCH_CHECK();
  return 0;
}
