// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   warp 0      : TMA producer (one elected lane) — A tile [128 rows x 64 ch] via a 4-D NHWC tensor map
//                 (3x3 taps are shifted boxes; out-of-image reads are zero-filled by TMA = conv padding),
//                 B tile [BN x 64] from the K-major weight matrix.
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (M=128, N=BN, K=16 per instruction),
//                 tcgen05.commit releases smem stages and finally signals the epilogue.
//   warps 2..5  : epilogue — tcgen05.ld accumulator rows (one row per thread), fused
//                 alpha/bias/FiLM/residual/activation, fp16 and/or fp32 stores.
//
// Replaces cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in the reference
// (ldm/modules/diffusionmodules/openaimodel.py:204,230,241,519,685; ldm/modules/attention.py:40-60,161-168,233-248).
#include "../../include/sdb200.h"
#include "host.h"
#include "ptx.cuh"

namespace sdb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;

struct GemmArgs {
  int M, N;
  int taps, chunks0, chunks_tot;  // 64-channel chunks in source 0 / both sources
  int H, W, NB;
  int TW, TH, TN, tiles_x, tiles_y;
  int k_iters, iters_per_split;
  float alpha;
  const float* bias;
  const float* film;
  int ldf;
  int rows_per_sample;
  const float* residual;
  int ldr;
  __half* out_f16;
  float* out_f32;
  int ldo;
  float* ws;
  int act;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ bool map_row(const GemmArgs& p, int m_tile, int r, int& out_row) {
  if (p.taps == 1) {
    out_row = m_tile * BM + r;
    return out_row < p.M;
  }
  int tx = m_tile % p.tiles_x;
  int t2 = m_tile / p.tiles_x;
  int ty = t2 % p.tiles_y;
  int tn = t2 / p.tiles_y;
  int x = r % p.TW;
  int y = (r / p.TW) % p.TH;
  int nl = r / (p.TW * p.TH);
  int gx = tx * p.TW + x, gy = ty * p.TH + y, gn = tn * p.TN + nl;
  out_row = (gn * p.H + gy) * p.W + gx;
  return gy < p.H && gn < p.NB;
}

// v[32] holds alpha-scaled-to-be accumulators for columns [col0, col0+32) of row `row`.
__device__ __forceinline__ void epilogue_store(const GemmArgs& p, int row, int col0, int ncols, float (&v)[32]) {
  int nvalid = min(32, ncols - col0);
  if (nvalid <= 0) return;
  const float* film = p.film ? p.film + static_cast<size_t>(row / p.rows_per_sample) * p.ldf : nullptr;
  const float* res = p.residual ? p.residual + static_cast<size_t>(row) * p.ldr : nullptr;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nvalid) {
      float x = v[j] * p.alpha;
      if (p.bias) x += __ldg(p.bias + col0 + j);
      if (film) x += __ldg(film + col0 + j);
      if (res) x += res[col0 + j];
      if (p.act == SDB_ACT_QUICK_GELU) x = x * sigmoidf_(1.702f * x);
      else if (p.act == SDB_ACT_SILU) x = x * sigmoidf_(x);
      v[j] = x;
    }
  }
  size_t o = static_cast<size_t>(row) * p.ldo + col0;
  bool vec = (nvalid == 32) && ((p.ldo & 7) == 0) && ((col0 & 7) == 0);
  if (p.out_f32) {
    float* d = p.out_f32 + o;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < nvalid; ++j) d[j] = v[j];
    }
  }
  if (p.out_f16) {
    __half* d = p.out_f16 + o;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        __half2 h0 = __floats2half2_rn(v[j], v[j + 1]);
        __half2 h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
        __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]);
        __half2 h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(d + j) = u;
      }
    } else {
      for (int j = 0; j < nvalid; ++j) d[j] = __float2half_rn(v[j]);
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int split = blockIdx.z;
  const int it_begin = split * p.iters_per_split;
  const int it_end = min(p.k_iters, it_begin + p.iters_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      int x0 = 0, y0 = 0, n0 = 0;
      if (p.taps == 1) {
        x0 = m_tile * BM;
      } else {
        int tx = m_tile % p.tiles_x;
        int t2 = m_tile / p.tiles_x;
        x0 = tx * p.TW;
        y0 = (t2 % p.tiles_y) * p.TH;
        n0 = (t2 / p.tiles_y) * p.TN;
      }
      for (int it = it_begin; it < it_end; ++it) {
        int i = it - it_begin;
        int s = i % STAGES;
        uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        int tap = it / p.chunks_tot;
        int cc = it - tap * p.chunks_tot;
        int dx = 0, dy = 0;
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        uint8_t* a_s = smem + s * STAGE_BYTES;
        uint8_t* b_s = a_s + A_BYTES;
        if (cc < p.chunks0)
          tma_load_4d(a_s, &tmA0, &full_bar[s], cc * BK, x0 + dx, y0 + dy, n0);
        else
          tma_load_4d(a_s, &tmA1, &full_bar[s], (cc - p.chunks0) * BK, x0 + dx, y0 + dy, n0);
        tma_load_2d(b_s, &tmB, &full_bar[s], it * BK, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      for (int it = it_begin; it < it_end; ++it) {
        int i = it - it_begin;
        int s = i % STAGES;
        uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        uint64_t da = umma_desc_k128(a_addr);
        uint64_t db = umma_desc_k128(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 32 bytes (16 fp16) along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
          umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    // epilogue warps 2..5 : TMEM lane group = warp % 4
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    int out_row;
    bool valid = map_row(p, m_tile, r, out_row);
    if (it_end > it_begin) {
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_d + (static_cast<uint32_t>(lg * 32) << 16);
    if (p.ws) {
      // split-K: raw fp32 partial tile
      float* dst = p.ws + (static_cast<size_t>(split) * p.M + out_row) * p.N + n_tile * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(taddr + c * 32, rr);
        tmem_ld_wait();
        if (valid) {
          int col0 = n_tile * BN + c * 32;
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) dst[c * 32 + j] = (it_end > it_begin) ? __uint_as_float(rr[j]) : 0.f;
        }
      }
    } else if (p.act == SDB_ACT_GEGLU) {
      constexpr int HALF = BN / 2;
#pragma unroll 1
      for (int c = 0; c < HALF / 32; ++c) {
        uint32_t xr[32], gr[32];
        tmem_ld32(taddr + c * 32, xr);
        tmem_ld32(taddr + HALF + c * 32, gr);
        tmem_ld_wait();
        if (valid) {
          float v[32];
          int colx = n_tile * BN + c * 32;  // accumulator column of x; gate at +HALF
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(xr[j]) * p.alpha;
            float g = __uint_as_float(gr[j]) * p.alpha;
            if (p.bias) {
              x += __ldg(p.bias + colx + j);
              g += __ldg(p.bias + colx + HALF + j);
            }
            v[j] = x * gelu_erf(g);
          }
          int ocol = n_tile * HALF + c * 32;
          size_t o = static_cast<size_t>(out_row) * p.ldo + ocol;
          if (p.out_f16) {
            __half* d = p.out_f16 + o;
#pragma unroll
            for (int j = 0; j < 32; j += 2) *reinterpret_cast<__half2*>(d + j) = __floats2half2_rn(v[j], v[j + 1]);
          }
          if (p.out_f32) {
            float* d = p.out_f32 + o;
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = v[j];
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(taddr + c * 32, rr);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          epilogue_store(p, out_row, n_tile * BN + c * 32, p.N, v);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, TMEM_COLS);
  }
}

// split-K second pass: sum partials and apply the fused epilogue
__global__ void splitk_epilogue_kernel(const GemmArgs p, int splits) {
  size_t total = static_cast<size_t>(p.M) * p.N;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int row = static_cast<int>(idx / p.N);
    int col = static_cast<int>(idx - static_cast<size_t>(row) * p.N);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += p.ws[static_cast<size_t>(s) * total + idx];
    float x = acc * p.alpha;
    if (p.bias) x += p.bias[col];
    if (p.film) x += p.film[static_cast<size_t>(row / p.rows_per_sample) * p.ldf + col];
    if (p.residual) x += p.residual[static_cast<size_t>(row) * p.ldr + col];
    if (p.act == SDB_ACT_QUICK_GELU) x = x * sigmoidf_(1.702f * x);
    else if (p.act == SDB_ACT_SILU) x = x * sigmoidf_(x);
    size_t o = static_cast<size_t>(row) * p.ldo + col;
    if (p.out_f32) p.out_f32[o] = x;
    if (p.out_f16) p.out_f16[o] = __float2half_rn(x);
  }
}

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN <= 64 ? 4 : BN <= 128 ? 3 : 4;  // <=128: 2 CTAs/SM co-reside (97 KB each)
  static constexpr int SMEM = STAGES * (A_BYTES + BN * BK * 2) + 1024;
};

template <int BN>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmArgs& p,
                       dim3 grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, Cfg::STAGES>;
  static bool configured = false;
  if (!configured) {
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM, st>>>(a0, a1, b, p);
  SDB_LAUNCH_CHECK();
  return 0;
}

static int pow2_divisor(int v, int cap) {
  int p = 1;
  while (p * 2 <= cap && v % (p * 2) == 0) p *= 2;
  return p;
}

// Tile model used to pick (block_n, split-K): every 64-wide K step of a 128 x bn tile moves (128 + bn) * 128 B
// from L2 (the binding resource for 1-CTA tiles), the epilogue drains bn columns, plus a fixed launch/prologue cost;
// CTAs run in waves of sm_count. Split-K adds an fp32 partial round trip and a second kernel.
static double tile_cost(int n, long m_tiles, int k_iters, int bn, int splits, long M) {
  long nt = (n + bn - 1) / bn;
  long tiles = m_tiles * nt * splits;
  long waves = (tiles + sm_count() - 1) / sm_count();
  double per_tile = static_cast<double>((k_iters + splits - 1) / splits) * (128.0 + bn) + 2.0 * bn + 300.0;
  double t = waves * per_tile;
  if (splits > 1) t += 1500.0 + static_cast<double>(splits) * M * n / (sm_count() * 64.0);
  return t;
}

static void pick_tiles(int n, long m_tiles, int k_iters, long M, bool geglu, int fixed_bn, int fixed_splits,
                       long ws_floats, int* bn_out, int* splits_out) {
  const int cands[] = {256, 160, 128, 64, 32};
  double best_t = 1e300;
  int best_bn = 128, best_s = 1;
  for (int bn : cands) {
    if (fixed_bn > 0 && bn != fixed_bn) continue;
    if (geglu && bn != 128) continue;
    if (fixed_bn <= 0 && bn > 32 && ((n + bn - 1) / bn) * bn - n >= bn / 2 && n > 32) continue;  // mostly padding
    int smax = 1;
    if (fixed_splits == -1 && !geglu) {
      smax = k_iters / 4;
      if (smax > 32) smax = 32;
      if (smax < 1) smax = 1;
    }
    for (int s = 1; s <= smax; ++s) {
      if (s > 1 && static_cast<long>(s) * M * n > ws_floats) break;
      int sp = fixed_splits > 1 ? fixed_splits : s;
      double t = tile_cost(n, m_tiles, k_iters, bn, sp, M);
      if (t < best_t) {
        best_t = t;
        best_bn = bn;
        best_s = sp;
      }
    }
  }
  *bn_out = best_bn;
  *splits_out = best_s;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_gemm(const sdb_gemm_desc* d, sdb_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SDB_CHECK(d && d->a0 && d->b, "sdb_gemm: null operand");
  SDB_CHECK(d->taps == 1 || d->taps == 9, "sdb_gemm: taps must be 1 or 9 (got %d)", d->taps);
  SDB_CHECK(d->c0 > 0 && d->c0 % 64 == 0 && d->c1 >= 0 && d->c1 % 64 == 0,
            "sdb_gemm: channel counts must be multiples of 64 (c0=%d c1=%d)", d->c0, d->c1);
  SDB_CHECK((d->c1 == 0) == (d->a1 == nullptr), "sdb_gemm: a1/c1 mismatch");
  SDB_CHECK(d->nb > 0 && d->h > 0 && d->w > 0 && d->n > 0, "sdb_gemm: bad dims");
  SDB_CHECK(d->out_f16 || d->out_f32, "sdb_gemm: no output");

  GemmArgs p{};
  const long M = static_cast<long>(d->nb) * d->h * d->w;
  SDB_CHECK(M < (1L << 31), "sdb_gemm: M too large");
  p.M = static_cast<int>(M);
  p.N = d->n;
  p.taps = d->taps;
  p.chunks0 = d->c0 / 64;
  p.chunks_tot = (d->c0 + d->c1) / 64;
  p.k_iters = d->taps * p.chunks_tot;
  p.H = d->h;
  p.W = d->w;
  p.NB = d->nb;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.bias = d->bias;
  p.film = d->film;
  p.ldf = d->ldf > 0 ? d->ldf : d->n;
  p.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : d->h * d->w;
  p.residual = d->residual;
  p.ldr = d->ldr > 0 ? d->ldr : d->n;
  p.out_f16 = static_cast<__half*>(d->out_f16);
  p.out_f32 = d->out_f32;
  p.act = d->act;
  const bool geglu = d->act == SDB_ACT_GEGLU;
  const int n_out = geglu ? d->n / 2 : d->n;
  p.ldo = d->ldo > 0 ? d->ldo : n_out;
  if (geglu) {
    SDB_CHECK(d->n % 128 == 0, "sdb_gemm: GEGLU needs n %% 128 == 0");
    SDB_CHECK(!d->film && !d->residual && d->splits <= 1, "sdb_gemm: GEGLU epilogue excludes film/residual/split-K");
  }

  long m_tiles;
  if (d->taps == 1) {
    p.TW = 128;
    p.TH = 1;
    p.TN = 1;
    p.tiles_x = static_cast<int>((M + 127) / 128);
    p.tiles_y = 1;
    m_tiles = p.tiles_x;
  } else {
    p.TW = pow2_divisor(d->w, 128);
    p.TH = pow2_divisor(d->h, 128 / p.TW);
    p.TN = 128 / (p.TW * p.TH);
    p.tiles_x = d->w / p.TW;
    p.tiles_y = (d->h + p.TH - 1) / p.TH;
    m_tiles = static_cast<long>(p.tiles_x) * p.tiles_y * ((d->nb + p.TN - 1) / p.TN);
  }

  int bn = 128, splits = 1;
  {
    int fixed_splits = d->splits;
    if (fixed_splits == -1 && d->workspace == nullptr) fixed_splits = 0;
    if (fixed_splits > p.k_iters) fixed_splits = p.k_iters;
    pick_tiles(d->n, m_tiles, p.k_iters, M, geglu, d->block_n, fixed_splits, d->workspace_floats, &bn, &splits);
  }
  SDB_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256, "sdb_gemm: unsupported block_n %d", bn);
  SDB_CHECK(!geglu || bn == 128, "sdb_gemm: GEGLU requires block_n 128");
  p.iters_per_split = (p.k_iters + splits - 1) / splits;
  splits = (p.k_iters + p.iters_per_split - 1) / p.iters_per_split;
  if (splits > 1) {
    SDB_CHECK(d->workspace != nullptr, "sdb_gemm: split-K needs a workspace");
    SDB_CHECK(d->workspace_floats <= 0 || static_cast<long>(splits) * M * d->n <= d->workspace_floats,
              "sdb_gemm: split-K workspace too small");
    p.ws = d->workspace;
  }

  // tensor maps
  CUtensorMap tA0, tA1, tB;
  const int ctot = d->c0 + d->c1;
  auto make_a = [&](CUtensorMap* tm, const void* base, int c) -> int {
    if (d->taps == 1) {
      uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(M), 1, 1};
      uint64_t str[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(c) * 2 * M,
                         static_cast<uint64_t>(c) * 2 * M};
      uint32_t box[4] = {64, 128, 1, 1};
      return make_tmap_f16(tm, base, 4, dims, str, box);
    }
    uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(d->w), static_cast<uint64_t>(d->h),
                        static_cast<uint64_t>(d->nb)};
    uint64_t str[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(c) * 2 * d->w,
                       static_cast<uint64_t>(c) * 2 * d->w * d->h};
    uint32_t box[4] = {64, static_cast<uint32_t>(p.TW), static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
    return make_tmap_f16(tm, base, 4, dims, str, box);
  };
  if (make_a(&tA0, d->a0, d->c0)) return 1;
  if (d->a1) {
    if (make_a(&tA1, d->a1, d->c1)) return 1;
  } else {
    tA1 = tA0;
  }
  {
    uint64_t K = static_cast<uint64_t>(d->taps) * ctot;
    uint64_t dims[2] = {K, static_cast<uint64_t>(d->n)};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    if (make_tmap_f16(&tB, d->b, 2, dims, str, box)) return 1;
  }

  dim3 grid(static_cast<unsigned>(m_tiles), static_cast<unsigned>((d->n + bn - 1) / bn), static_cast<unsigned>(splits));
  int rc;
  switch (bn) {
    case 32: rc = launch_gemm<32>(tA0, tA1, tB, p, grid, st); break;
    case 64: rc = launch_gemm<64>(tA0, tA1, tB, p, grid, st); break;
    case 128: rc = launch_gemm<128>(tA0, tA1, tB, p, grid, st); break;
    case 160: rc = launch_gemm<160>(tA0, tA1, tB, p, grid, st); break;
    case 256: rc = launch_gemm<256>(tA0, tA1, tB, p, grid, st); break;
    default: SDB_CHECK(false, "sdb_gemm: unsupported block_n %d", bn);
  }
  if (rc) return rc;
  if (splits > 1) {
    size_t total = static_cast<size_t>(p.M) * p.N;
    int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(sm_count()) * 8));
    splitk_epilogue_kernel<<<blocks, 256, 0, st>>>(p, splits);
    SDB_LAUNCH_CHECK();
  }
  return 0;
}
