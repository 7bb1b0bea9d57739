// Small-M linear layer on CUDA cores with fp32 activations: out[m, j] = act(x[m, :] . W[j, :] + b[j]).
// Used for the timestep-embedding MLP and the 22 ResBlock emb_layers (openaimodel.py:506-511, 217-223, 723-724,
// 267), where M = number of samples (<= 128) and the work is a weight-streaming GEMV: HBM/L2-bound on W
// (fp16, read once), fp32 accumulate, no tensor-core tile would be more than 1/64 full.
#include "../../include/sdb200.h"
#include "host.h"
#include <cuda_fp16.h>

namespace sdb {

constexpr int LS_ROWS = 8;      // samples per pass (activations of one pass live in shared memory)
constexpr int LS_WARPS = 8;
constexpr int LS_JPW = 4;       // output features per warp: their weight rows stream with 16-byte loads, all in flight

// Weight-streaming GEMV: a warp owns LS_JPW output features; each lane reads 16-byte pieces (8 fp16) of their weight
// rows, every load of a k-sweep issued before the first use (HBM-latency bound otherwise: the 22 emb_layers are 52 MB
// of weights for 2 x 1280 activations). Activations are staged once per block in shared memory as fp32.
__global__ void __launch_bounds__(32 * LS_WARPS)
    linear_small_kernel(const float* __restrict__ x, int M, int K, const __half* __restrict__ W, int N,
                        const float* __restrict__ bias, int act, float* __restrict__ out32,
                        __half* __restrict__ out16) {
  extern __shared__ float ls_x[];   // [min(M, LS_ROWS)][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = (blockIdx.x * LS_WARPS + warp) * LS_JPW;
  const int K8 = K >> 3;
  for (int m0 = 0; m0 < M; m0 += LS_ROWS) {
    const int mb = min(LS_ROWS, M - m0);
    __syncthreads();
    for (int i = threadIdx.x; i < mb * (K >> 2); i += blockDim.x)
      reinterpret_cast<float4*>(ls_x)[i] = reinterpret_cast<const float4*>(x + static_cast<size_t>(m0) * K)[i];
    __syncthreads();
    float acc[LS_JPW][LS_ROWS];
#pragma unroll
    for (int jj = 0; jj < LS_JPW; ++jj)
#pragma unroll
      for (int r = 0; r < LS_ROWS; ++r) acc[jj][r] = 0.f;
    for (int k8 = lane; k8 < K8; k8 += 32 * 2) {
      uint4 w[LS_JPW][2];
#pragma unroll
      for (int jj = 0; jj < LS_JPW; ++jj)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = k8 + u * 32;
          w[jj][u] = (j0 + jj < N && kk < K8)
                         ? __ldg(reinterpret_cast<const uint4*>(W + static_cast<size_t>(j0 + jj) * K) + kk)
                         : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kk = k8 + u * 32;
        if (kk >= K8) break;
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r) {
          if (r >= mb) break;
          const float4 xa = *reinterpret_cast<const float4*>(ls_x + static_cast<size_t>(r) * K + kk * 8);
          const float4 xb = *reinterpret_cast<const float4*>(ls_x + static_cast<size_t>(r) * K + kk * 8 + 4);
#pragma unroll
          for (int jj = 0; jj < LS_JPW; ++jj) {
            const __half2* h = reinterpret_cast<const __half2*>(&w[jj][u]);
            const float2 w0 = __half22float2(h[0]), w1 = __half22float2(h[1]), w2 = __half22float2(h[2]),
                         w3 = __half22float2(h[3]);
            float a = acc[jj][r];
            a = fmaf(w0.x, xa.x, a);
            a = fmaf(w0.y, xa.y, a);
            a = fmaf(w1.x, xa.z, a);
            a = fmaf(w1.y, xa.w, a);
            a = fmaf(w2.x, xb.x, a);
            a = fmaf(w2.y, xb.y, a);
            a = fmaf(w3.x, xb.z, a);
            a = fmaf(w3.y, xb.w, a);
            acc[jj][r] = a;
          }
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < LS_JPW; ++jj)
#pragma unroll
      for (int r = 0; r < LS_ROWS; ++r) {
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[jj][r] += __shfl_xor_sync(0xffffffffu, acc[jj][r], o);
      }
    if (lane == 0) {
#pragma unroll
      for (int jj = 0; jj < LS_JPW; ++jj) {
        const int j = j0 + jj;
        if (j >= N) break;
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r) {
          if (r >= mb) break;
          float v = acc[jj][r] + (bias ? bias[j] : 0.f);
          if (act == SDB_ACT_SILU) v = v / (1.0f + __expf(-v));
          const size_t o = static_cast<size_t>(m0 + r) * N + j;
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
  SDB_REC(sdb_linear_small(x, m, k, w_f16, n, bias, act, out_f32, out_f16, s_));
  SDB_CHECK(x && w_f16 && (out_f32 || out_f16), "sdb_linear_small: null pointer");
  SDB_CHECK(k % 8 == 0 && m > 0 && n > 0, "sdb_linear_small: k must be a multiple of 8 (got %d)", k);
  SDB_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_f16) & 15) == 0,
            "sdb_linear_small: x and w must be 16-byte aligned");
  SDB_CHECK(act == SDB_ACT_NONE || act == SDB_ACT_SILU, "sdb_linear_small: unsupported activation");
  const size_t smem = static_cast<size_t>(m < LS_ROWS ? m : LS_ROWS) * k * sizeof(float);
  SDB_CHECK(smem <= 200 * 1024, "sdb_linear_small: k=%d too large", k);
  static bool configured = false;
  if (!configured) {
    SDB_CUDA(cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const int per_block = LS_WARPS * LS_JPW;
  linear_small_kernel<<<(n + per_block - 1) / per_block, 32 * LS_WARPS, smem, static_cast<cudaStream_t>(stream)>>>(
      x, m, k, static_cast<const __half*>(w_f16), n, bias, act, out_f32, static_cast<__half*>(out_f16));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_timestep_embedding_f32(const float* t, int32_t n, int32_t dim, float max_period, float* out,
                                          sdb_stream_t stream) {
  SDB_REC(sdb_timestep_embedding_f32(t, n, dim, max_period, out, s_));
  SDB_CHECK(t && out && dim % 2 == 0, "sdb_timestep_embedding_f32: bad arguments");
  int total = n * dim / 2;
  timestep_embedding_f32_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, n, dim,
                                                                                                    max_period, out);
  SDB_LAUNCH_CHECK();
  return 0;
}
