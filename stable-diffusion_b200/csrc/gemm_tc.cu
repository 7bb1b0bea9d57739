// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a — persistent, warp-specialised.
//
//   grid        : min(#tiles, #SMs) persistent CTAs, static round-robin tile schedule (M fastest, so CTAs running
//                 side by side read the same weight tile from L2)
//   warp 0      : TMA producer (one elected lane) — A tile [128 rows x 64 ch] via 4-D NHWC tensor maps (3x3 taps are
//                 shifted boxes; out-of-image reads are zero-filled by TMA = conv padding; up to four A sources are
//                 concatenated along K: UNet skip concat, or hi/lo fp16 splits of one fp32 activation),
//                 B tile [BN x 64] from the K-major weight matrix. STAGES-deep mbarrier ring that runs across tiles.
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (M=128, N=BN, K=16), two accumulator buffers in
//                 TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//   warps 2..9  : epilogue (two warps per TMEM lane group, alternating 32-column chunks) — tcgen05.ld accumulator
//                 rows (one row per thread), fused alpha/bias/FiLM/residual/activation in registers, 16-byte stores
//                 into a swizzled staging tile, TMA store (cp.async.bulk.tensor) of each 32x32 block to the NHWC
//                 output: fp16 (optionally a hi+lo pair) and/or fp32, or raw fp32 split-K partials. Outputs whose row
//                 pitch TMA cannot address (N = 3, 4 ...) take a scalar transposed path.
//
// Replaces cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in the reference
// (ldm/modules/diffusionmodules/openaimodel.py:204,230,241,519,685; ldm/modules/attention.py:40-60,161-168,233-248).
#include "../../include/sdb200.h"
#include "host.h"
#include "ptx.cuh"

#include <algorithm>

namespace sdb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int MAX_SRC = 4;
constexpr int EPI_WARPS = 8;
constexpr int STG_WARP_BYTES = 8192;  // per epilogue warp: 2 x 4 KB fp32 tiles, or 2 x (2 KB hi + 2 KB lo) fp16 tiles
constexpr int STAGING_BYTES = EPI_WARPS * STG_WARP_BYTES;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int STAT_SLOTS = 4;   // fused GroupNorm statistics are spread over 4 accumulator copies (m_tile & 3)

struct TmapPack {
  CUtensorMap a[MAX_SRC];
  CUtensorMap b;
  CUtensorMap o32;   // fp32 output (or the split-K workspace), 5-D [C, W, H, NB, S], box {32, bw, bh, bn, 1}, SWIZZLE_128B
  CUtensorMap o16;   // fp16 output, same geometry, SWIZZLE_64B
  CUtensorMap o16lo; // fp16 low half
  CUtensorMap ows;   // split-K fp32 partial planes [C, W, H, NB, splits]
};

struct GemmArgs {
  int M, N;
  int taps, nsrc;
  int cb[MAX_SRC + 1];  // cumulative 64-channel chunk boundaries of the A sources; cb[nsrc] = chunks per tap
  int H, W, NB;
  int TW, TH, TN, tiles_x, tiles_y;
  int k_iters, iters_per_split, splits;
  int m_tiles, n_tiles;
  float alpha;
  const float* bias;
  const float* film;
  int ldf;
  int rows_per_sample;
  const float* residual;
  int ldr;
  __half* out_f16;
  __half* out_f16_lo;
  float* out_f32;
  int ldo;
  float* ws;
  int act;
  double* stats;    // optional per-(sample, channel) {sum, sum of squares} of the fp32 output (GroupNorm statistics)
  int stats_halves;       // 1: the 128 rows of a tile belong to one sample; 2: rows 0-63 / 64-127 to two samples
  int n_samples;
  int b_static;     // B is a weight matrix: safe to prefetch before griddepcontrol.wait
  int fast;         // outputs go through the TMA-store epilogue
  int bw, bh;       // store box: bw x bh x (32 / (bw*bh)) output pixels per epilogue warp
  int cstride, cshift;   // 3x3 conv: input pixel = cstride * o + tap - 1 + cshift (per axis)
  int film_table;   // 1: rows 0-63 / 64-127 of every tile belong to one sample each, so bias + FiLM fold into a
                    // per-tile shared-memory column table; 0: FiLM is read per row from global memory
  unsigned long long* trace;  // debug: per-CTA phase timestamps (sdb_debug_trace), NULL in production
};

// Exact-erf GELU (attention.py:44, F.gelu default) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far
// below the fp16 rounding of the GEGLU output): one MUFU.RCP + one MUFU.EX2 + 7 FMA instead of erff's branchy ~30.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SDB_ACT_QUICK_GELU) return x * sigmoidf_(1.702f * x);
  if (act == SDB_ACT_SILU) return x * sigmoidf_(x);
  return x;
}

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

// One element of the fused epilogue (after alpha/bias which are column-only).
__device__ __forceinline__ void store_elem(const GemmArgs& p, float x, int orow, int sample, int col, bool finish) {
  if (finish) {
    if (p.film) x += p.film[static_cast<size_t>(sample) * p.ldf + col];
    if (p.residual) x += p.residual[static_cast<size_t>(orow) * p.ldr + col];
    x = apply_act(x, p.act);
  }
  size_t o = static_cast<size_t>(orow) * p.ldo + col;
  if (p.out_f32) p.out_f32[o] = x;
  if (p.out_f16) {
    __half h = __float2half_rn(x);
    p.out_f16[o] = h;
    if (p.out_f16_lo) p.out_f16_lo[o] = __float2half_rn(x - __half2float(h));
  }
}

// Drain one 32x32 fp32 chunk that sits in the warp's padded staging tile (row = TMEM lane, col = chunk column) to
// global memory with row-contiguous accesses: 16 lanes x 2 columns per row, two rows per instruction.
//   mode 0: fused epilogue (alpha, bias, FiLM, residual, activation)   mode 1: raw split-K partial   mode 2: values
//   already final (GEGLU computed in row layout)
__device__ __forceinline__ void drain_chunk(const GemmArgs& p, const float* stage, int lane, int my_row, int my_sample,
                                            bool my_valid, int ocol0, int ncols, int mode, int split) {
  const int l16 = lane & 15, rsel = lane >> 4;
  const int col = ocol0 + 2 * l16;
  const bool c0 = col < ncols, c1 = col + 1 < ncols;
  float b0 = 0.f, b1 = 0.f;
  if (mode == 0 && p.bias) {
    if (c0) b0 = __ldg(p.bias + col);
    if (c1) b1 = __ldg(p.bias + col + 1);
  }
  const bool vec = c1 && ((p.ldo & 1) == 0) && (!p.residual || (p.ldr & 1) == 0) && (!p.film || (p.ldf & 1) == 0);
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int r = it * 2 + rsel;
    const int orow = __shfl_sync(0xffffffffu, my_row, r);
    const int sample = __shfl_sync(0xffffffffu, my_sample, r);
    const bool valid = __shfl_sync(0xffffffffu, my_valid ? 1 : 0, r) != 0;
    if (!valid || !c0) continue;
    float x0 = stage[r * 33 + 2 * l16];
    float x1 = stage[r * 33 + 2 * l16 + 1];
    if (mode == 1) {
      float* dst = p.ws + (static_cast<size_t>(split) * p.M + orow) * p.N + col;
      if (c1 && (p.N & 1) == 0) {
        *reinterpret_cast<float2*>(dst) = make_float2(x0, x1);
      } else {
        dst[0] = x0;
        if (c1) dst[1] = x1;
      }
      continue;
    }
    if (mode == 0) {
      x0 = x0 * p.alpha + b0;
      x1 = x1 * p.alpha + b1;
    }
    if (!vec) {
      store_elem(p, x0, orow, sample, col, mode == 0);
      if (c1) store_elem(p, x1, orow, sample, col + 1, mode == 0);
      continue;
    }
    if (mode == 0) {
      if (p.film) {
        float2 f = *reinterpret_cast<const float2*>(p.film + static_cast<size_t>(sample) * p.ldf + col);
        x0 += f.x;
        x1 += f.y;
      }
      if (p.residual) {
        float2 rv = *reinterpret_cast<const float2*>(p.residual + static_cast<size_t>(orow) * p.ldr + col);
        x0 += rv.x;
        x1 += rv.y;
      }
      x0 = apply_act(x0, p.act);
      x1 = apply_act(x1, p.act);
    }
    const size_t o = static_cast<size_t>(orow) * p.ldo + col;
    if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + o) = make_float2(x0, x1);
    if (p.out_f16) {
      __half2 h = __floats2half2_rn(x0, x1);
      *reinterpret_cast<__half2*>(p.out_f16 + o) = h;
      if (p.out_f16_lo) {
        float2 hf = __half22float2(h);
        *reinterpret_cast<__half2*>(p.out_f16_lo + o) = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
      }
    }
  }
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug trace: word w of this CTA's record (8 words per CTA after an 8-word launch header)
#define SDB_TR(w, val)                                                          \
  do {                                                                          \
    if (p.trace) p.trace[8 + blockIdx.x * 8 + (w)] = (val);                     \
  } while (0)

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN <= 32 ? 6 : BN <= 64 ? 6 : BN <= 128 ? 4 : BN <= 160 ? 4 : 3;
  static constexpr int TMEM_COLS = BN <= 32 ? 64 : BN <= 64 ? 128 : BN <= 128 ? 256 : 512;  // two accumulators
  static constexpr int SMEM = STAGES * (A_BYTES + BN * BK * 2) + STAGING_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TmapPack tm, const GemmArgs p) {
  constexpr int STAGES = GemmCfg<BN>::STAGES;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = GemmCfg<BN>::TMEM_COLS;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offset (keeps the shared address space visible to the compiler: LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t acc_full[2];
  __shared__ uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  // fused GroupNorm statistics: [lane group][column][sum, sum of squares]; one writer per slot per tile and a
  // fixed-order fold at the flush (deterministic; the cross-CTA combine uses fp64 atomics)
  __shared__ float colsum[4 * BN * 2];
  // per-tile column constants of the fused epilogue: bias[col] (+ FiLM[sample of the row half][col]), so the chunk
  // loop reads them from shared memory instead of paying a global-load latency per chunk
  __shared__ __align__(16) float coltab[2 * BN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_total = p.m_tiles * p.n_tiles * p.splits;
  pdl_launch_dependents();   // the next kernel may start its prologue while this one runs
  const long long clk0 = clock64();
  if (p.trace && threadIdx.x == 0) {
    SDB_TR(0, gtimer());
    if (blockIdx.x == 0) {
      p.trace[0] = gridDim.x;
      p.trace[1] = BN;
      p.trace[2] = p.splits;
      p.trace[3] = p.k_iters;
      p.trace[4] = p.M;
      p.trace[5] = p.N;
      p.trace[6] = p.taps;
      p.trace[7] = tiles_total;
    }
  }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nsrc; ++i) tma_prefetch_desc(&tm.a[i]);
    tma_prefetch_desc(&tm.b);
    if (p.fast) {
      if (p.ws) tma_prefetch_desc(&tm.ows);
      if (p.out_f32) tma_prefetch_desc(&tm.o32);
      if (p.out_f16) tma_prefetch_desc(&tm.o16);
      if (p.out_f16_lo) tma_prefetch_desc(&tm.o16lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPI_WARPS);
    }
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
  if (threadIdx.x == 0) SDB_TR(2, clock64() - clk0);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t i = 0;  // ring position, continues across tiles
      // PDL: the weight (B) tiles do not depend on the previous kernel — start streaming them for the first ring
      // pass of the first tile before waiting for the producer of the activations
      int npre = 0;
      if (p.b_static && blockIdx.x < tiles_total) {
        const int rest0 = blockIdx.x / p.m_tiles;
        const int n_tile0 = rest0 % p.n_tiles;
        const int split0 = rest0 / p.n_tiles;
        const int b0 = split0 * p.iters_per_split;
        const int e0 = min(p.k_iters, b0 + p.iters_per_split);
        npre = min(STAGES, e0 - b0);
        for (int j = 0; j < npre; ++j) {
          mbar_arrive_expect_tx(&full_bar[j], STAGE_BYTES);
          tma_load_2d(smem + j * STAGE_BYTES + A_BYTES, &tm.b, &full_bar[j], (b0 + j) * BK, n_tile0 * BN);
        }
      }
      pdl_wait();
      SDB_TR(3, clock64() - clk0);
      for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        const int m_tile = tile % p.m_tiles;
        const int rest = tile / p.m_tiles;
        const int n_tile = rest % p.n_tiles;
        const int split = rest / p.n_tiles;
        const int it_begin = split * p.iters_per_split;
        const int it_end = min(p.k_iters, it_begin + p.iters_per_split);
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
        const int cpt = p.cb[p.nsrc];
        for (int it = it_begin; it < it_end; ++it, ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          const bool prefetched = static_cast<int>(i) < npre;   // B already in flight, barrier already armed
          if (!prefetched) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          }
          const int tap = it / cpt;
          const int cc = it - tap * cpt;
          int dx = 0, dy = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          int src = 0;
          while (src + 1 < p.nsrc && cc >= p.cb[src + 1]) ++src;
          uint8_t* a_s = smem + s * STAGE_BYTES;
          tma_load_4d(a_s, &tm.a[src], &full_bar[s], (cc - p.cb[src]) * BK, x0 * p.cstride + dx + p.cshift,
                      y0 * p.cstride + dy + p.cshift, n0);
          if (!prefetched) tma_load_2d(a_s + A_BYTES, &tm.b, &full_bar[s], it * BK, n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      uint32_t i = 0, local = 0;
      for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++local) {
        const int split = (tile / p.m_tiles) / p.n_tiles;
        const int it_begin = split * p.iters_per_split;
        const int it_end = min(p.k_iters, it_begin + p.iters_per_split);
        const uint32_t ab = local & 1;
        mbar_wait(&acc_empty[ab], ((local >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_d + ab * BN;
        for (int it = it_begin; it < it_end; ++it, ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (i == 0) SDB_TR(4, clock64() - clk0);
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma_desc_k128(a_addr);
          const uint64_t db = umma_desc_k128(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 32 bytes (16 fp16) along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
            umma_f16(d_addr, da + 2 * k, db + 2 * k, idesc, (it > it_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[ab]);
      }
      SDB_TR(5, clock64() - clk0);
    }
  } else {
    // epilogue warps 2..9 : TMEM lane group = warp % 4; the two warps of a lane group alternate chunks.
    // Latency plan: everything the fused epilogue reads from global memory is requested BEFORE the accumulator is
    // ready - bias (+ FiLM) of the tile's columns go to a shared-memory table, the residual rows of a chunk are
    // prefetched into registers one chunk ahead (the first one while the main loop still runs) - so the chunk loop is
    // TMEM load -> FMAs -> staging -> TMA store with no exposed L2 round trip.
    pdl_wait();   // residual / FiLM reads and all output writes come after the previous kernel has completed
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int par = ew >> 2;
    const int et = threadIdx.x - 64;
    uint8_t* stg = staging + ew * STG_WARP_BYTES;
    float* stage = reinterpret_cast<float*>(stg);  // scalar path: [32][33] floats
    const bool geglu = (p.act == SDB_ACT_GEGLU) && !p.ws;
    constexpr int HALF = BN / 2;
    const int n_chunks = geglu ? HALF / 32 : BN / 32;
    const int n_lim = geglu ? p.N / 2 : p.N;   // output columns that exist
    const bool st32 = p.ws || p.out_f32;   // an fp32 tile is staged in some phase (output or split-K partial)
    const bool st16 = p.out_f16 != nullptr;
    // staging buffers per chunk parity: fp32 tiles 4 KB each; fp16 hi 2 KB + lo 2 KB each (fp32+fp16 together: single)
    const bool dbl = !(st32 && st16);
    const bool split_fast = p.ws && p.fast;   // raw fp32 partial planes; finished by splitk_epilogue_kernel
    const bool use_tab = p.fast && !split_fast;
    const bool pre_res = use_tab && !geglu && p.residual != nullptr;
    uint32_t flip = 0;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++local) {
      const int m_tile = tile % p.m_tiles;
      const int rest = tile / p.m_tiles;
      const int n_tile = rest % p.n_tiles;
      const int split = rest / p.n_tiles;
      const uint32_t ab = local & 1;
      int my_row;
      const int r_tile = lg * 32 + lane;
      const bool my_valid = map_row(p, m_tile, r_tile, my_row);
      const int my_sample = my_valid ? my_row / p.rows_per_sample : 0;
      // store-box origin of this warp's 32 rows
      int sx, sy, sn;
      if (p.taps == 1) {
        sx = m_tile * BM + lg * 32;
        sy = 0;
        sn = 0;
      } else {
        const int tx = m_tile % p.tiles_x;
        const int t2 = m_tile / p.tiles_x;
        const int r0 = lg * 32;
        sx = tx * p.TW + r0 % p.TW;
        sy = (t2 % p.tiles_y) * p.TH + (r0 / p.TW) % p.TH;
        sn = (t2 / p.tiles_y) * p.TN + r0 / (p.TW * p.TH);
      }
      // ---- column table of this tile: bias (+ FiLM of the sample each row half belongs to)
      if (use_tab) {
        if (local > 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // readers of the previous tile's table
        for (int i = et; i < 2 * BN; i += 32 * EPI_WARPS) {
          const int hsel = i / BN, cl = i - hsel * BN;
          const int col = n_tile * BN + cl;
          float t = 0.f;
          if (col < p.N) {
            if (p.bias) t = __ldg(p.bias + col);
            if (p.film && p.film_table) {
              int prow;
              if (map_row(p, m_tile, hsel * 64, prow))
                t += __ldg(p.film + static_cast<size_t>(prow / p.rows_per_sample) * p.ldf + col);
            }
          }
          coltab[i] = t;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }
      // residual rows of one 32-column chunk -> registers (row per thread, 128 contiguous bytes)
      float4 rcur[8];
      auto load_res = [&](int c, float4 (&r)[8]) {
        const int oc = n_tile * BN + c * 32;
        if (my_valid && oc < n_lim) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(my_row) * p.ldr + oc);
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = rp[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (pre_res && par < n_chunks) load_res(par, rcur);

      mbar_wait(&acc_full[ab], (local >> 1) & 1);
      tc_fence_after();
      if (local == 0 && threadIdx.x == 64) SDB_TR(6, clock64() - clk0);
      const uint32_t taddr = tmem_d + ab * BN + (static_cast<uint32_t>(lg * 32) << 16);
      int last_c = -1;
      for (int c = par; c < n_chunks; c += 2) last_c = c;
      if (last_c < 0) {  // BN = 32: the odd-parity warps own no chunk but still take part in the hand-off
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
      }
      // ---- fused epilogue + staging + TMA store of one 32x32 chunk held in registers (row per thread)
      auto emit = [&](float (&v)[32], int ocol0, bool fuse, bool raw_partial, const float4 (&res)[8]) {
        if (fuse) {
          // alpha * acc + (bias [+ FiLM]) from the tile's column table; all pointers are 16-byte aligned on this path
          const float4* tp = reinterpret_cast<const float4*>(coltab + (lg >= 2 ? BN : 0) + (ocol0 - n_tile * BN));
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = tp[q];
            v[4 * q] = fmaf(v[4 * q], p.alpha, t.x);
            v[4 * q + 1] = fmaf(v[4 * q + 1], p.alpha, t.y);
            v[4 * q + 2] = fmaf(v[4 * q + 2], p.alpha, t.z);
            v[4 * q + 3] = fmaf(v[4 * q + 3], p.alpha, t.w);
          }
          if (p.film && !p.film_table) {   // rows of a tile half span several samples: FiLM per row from global
            const float4* fp = reinterpret_cast<const float4*>(p.film + static_cast<size_t>(my_sample) * p.ldf + ocol0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 t = __ldg(fp + q);
              v[4 * q] += t.x;
              v[4 * q + 1] += t.y;
              v[4 * q + 2] += t.z;
              v[4 * q + 3] += t.w;
            }
          }
          if (p.residual) {   // prefetched (zeros for rows outside the problem)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              v[4 * q] += res[q].x;
              v[4 * q + 1] += res[q].y;
              v[4 * q + 2] += res[q].z;
              v[4 * q + 3] += res[q].w;
            }
          }
          if (p.act != SDB_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          }
        }
        const bool w32 = raw_partial || p.out_f32;
        const bool w16 = !raw_partial && p.out_f16;
        const bool w16lo = w16 && p.out_f16_lo;
        // staging buffer for this chunk; make sure the TMA store that last read it has finished reading
        const uint32_t bsel = dbl ? (flip & 1) : 0;
        if (lane == 0) {
          if (dbl) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        __syncwarp();
        uint8_t* s32 = stg + bsel * 4096;
        uint8_t* s16 = st32 ? stg + 4096 : stg + bsel * 4096;
        uint8_t* s16l = s16 + 2048;
        if (w32) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(s32 + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (w16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __half2 h[4];
            uint4 u, ul;
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
            u.x = *reinterpret_cast<uint32_t*>(&h[0]);
            u.y = *reinterpret_cast<uint32_t*>(&h[1]);
            u.z = *reinterpret_cast<uint32_t*>(&h[2]);
            u.w = *reinterpret_cast<uint32_t*>(&h[3]);
            const uint32_t off = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(s16 + off) = u;
            if (w16lo) {
              __half2 l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 hf = __half22float2(h[e]);
                l[e] = __floats2half2_rn(v[8 * q + 2 * e] - hf.x, v[8 * q + 2 * e + 1] - hf.y);
              }
              ul.x = *reinterpret_cast<uint32_t*>(&l[0]);
              ul.y = *reinterpret_cast<uint32_t*>(&l[1]);
              ul.z = *reinterpret_cast<uint32_t*>(&l[2]);
              ul.w = *reinterpret_cast<uint32_t*>(&l[3]);
              *reinterpret_cast<uint4*>(s16l + off) = ul;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (raw_partial) {
            tma_store_5d(&tm.ows, s32, ocol0, sx, sy, sn, split);
          } else {
            if (w32) tma_store_5d(&tm.o32, s32, ocol0, sx, sy, sn, 0);
            if (w16) tma_store_5d(&tm.o16, s16, ocol0, sx, sy, sn, 0);
            if (w16lo) tma_store_5d(&tm.o16lo, s16l, ocol0, sx, sy, sn, 0);
          }
          tma_store_commit();
        }
        if (!raw_partial && p.stats && w32) {
          // GroupNorm statistics of the value just produced: lane c folds column c of the staged 32x32 tile into the
          // CTA's shared column sums (flushed once per tile with one fp64 global atomic per column)
          const uint32_t vmask = __ballot_sync(0xffffffffu, my_valid);
          float cs = 0.f, cq = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            if ((vmask >> r) & 1u) {
              float x = *reinterpret_cast<const float*>(s32 + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4);
              cs += x;
              cq = fmaf(x, x, cq);
            }
          }
          const int cl = ocol0 - n_tile * BN + lane;
          colsum[(lg * BN + cl) * 2] = cs;
          colsum[(lg * BN + cl) * 2 + 1] = cq;
        }
        ++flip;
      };

#pragma unroll 1
      for (int c = par; c < n_chunks; c += 2) {
        float v[32];
        float4 rnxt[8];
        int ocol0;
        const bool has_next = pre_res && (c + 2 < n_chunks);
        if (geglu) {
          uint32_t xr[32], gr[32];
          tmem_ld32(taddr + c * 32, xr);
          tmem_ld32(taddr + HALF + c * 32, gr);
          tmem_ld_wait();
          const float* tb = coltab + c * 32;   // bias of the value half; gate half at +HALF
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = fmaf(__uint_as_float(xr[j]), p.alpha, tb[j]);
            const float g = fmaf(__uint_as_float(gr[j]), p.alpha, tb[HALF + j]);
            v[j] = x * gelu_erf(g);
          }
          ocol0 = n_tile * HALF + c * 32;
        } else {
          uint32_t rr[32];
          tmem_ld32(taddr + c * 32, rr);
          if (has_next) load_res(c + 2, rnxt);   // next chunk's residual rows: in flight across this chunk's work
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          ocol0 = n_tile * BN + c * 32;
        }
        if (c == last_c) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        if (!p.fast) {
          // scalar transposed path (row pitch not TMA-addressable); split-K partials are finished by
          // splitk_epilogue_kernel on this path
#pragma unroll
          for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
          __syncwarp();
          drain_chunk(p, stage, lane, my_row, my_sample, my_valid, ocol0, geglu ? p.N / 2 : p.N, geglu ? 2 : (p.ws ? 1 : 0),
                      split);
          __syncwarp();
          continue;
        }
        if (ocol0 < n_lim) emit(v, ocol0, !geglu && !split_fast, split_fast, rcur);   // (columns past N: nothing to write)
        if (has_next) {
#pragma unroll
          for (int q = 0; q < 8; ++q) rcur[q] = rnxt[q];
        }
      }
      // ---- per-tile flush of the fused GroupNorm column sums (see emit)
      if (p.stats && !p.ws) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // all smem column sums of this tile are in
        const int halves = p.stats_halves;
        for (int i = et; i < halves * BN; i += 32 * EPI_WARPS) {
          const int hsel = i / BN, cl = i - hsel * BN;
          const int col = n_tile * BN + cl;
          const int r_probe = hsel * 64;   // first tile row of this half
          int prow;
          const bool pv = map_row(p, m_tile, r_probe, prow);
          if (pv && col < p.N) {
            const int sample = prow / p.rows_per_sample;
            double* dst = p.stats + ((static_cast<size_t>(m_tile & (STAT_SLOTS - 1)) * p.n_samples + sample) * p.N + col) * 2;
            float a, b;
            if (halves == 2) {
              a = colsum[((2 * hsel) * BN + cl) * 2] + colsum[((2 * hsel + 1) * BN + cl) * 2];
              b = colsum[((2 * hsel) * BN + cl) * 2 + 1] + colsum[((2 * hsel + 1) * BN + cl) * 2 + 1];
            } else {
              a = (colsum[cl * 2] + colsum[(BN + cl) * 2]) + (colsum[(2 * BN + cl) * 2] + colsum[(3 * BN + cl) * 2]);
              b = (colsum[cl * 2 + 1] + colsum[(BN + cl) * 2 + 1]) +
                  (colsum[(2 * BN + cl) * 2 + 1] + colsum[(3 * BN + cl) * 2 + 1]);
            }
            atomicAdd(dst, static_cast<double>(a));
            atomicAdd(dst + 1, static_cast<double>(b));
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // slots consumed before the next tile rewrites them
      }
    }
    // smem may be released once the bulk stores have READ it; their global writes complete with the grid
    if (lane == 0) tma_store_wait_read<0>();
    tc_fence_before();
    if (threadIdx.x == 64) SDB_TR(7, clock64() - clk0);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, TMEM_COLS);
  }
  if (threadIdx.x == 0) SDB_TR(1, gtimer());
}

// split-K second pass: sum the fp32 partial planes and apply the fused epilogue. Block = 32 rows x 128 columns
// (thread: 4 adjacent columns of rows ty, ty+8, ty+16, ty+24), so the GroupNorm column sums of the result fold
// through shared memory into one fp64 atomic per column per block.
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const GemmArgs p, int splits) {
  __shared__ float red[8][128][2];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + tx * 4;
  const size_t plane = static_cast<size_t>(p.M) * p.N;
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < p.N) {
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = blockIdx.y * 32 + ty + 8 * i;
      if (row >= p.M) continue;
      const float* src = p.ws + static_cast<size_t>(row) * p.N + col;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sp = 0; sp < splits; ++sp) {
        float4 t = __ldcg(reinterpret_cast<const float4*>(src + sp * plane));
        acc.x += t.x;
        acc.y += t.y;
        acc.z += t.z;
        acc.w += t.w;
      }
      float x[4] = {acc.x * p.alpha + bv.x, acc.y * p.alpha + bv.y, acc.z * p.alpha + bv.z, acc.w * p.alpha + bv.w};
      const int sample = row / p.rows_per_sample;
      if (p.film) {
        float4 f = *reinterpret_cast<const float4*>(p.film + static_cast<size_t>(sample) * p.ldf + col);
        x[0] += f.x; x[1] += f.y; x[2] += f.z; x[3] += f.w;
      }
      if (p.residual) {
        float4 r = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(row) * p.ldr + col);
        x[0] += r.x; x[1] += r.y; x[2] += r.z; x[3] += r.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j] = apply_act(x[j], p.act);
        cs[j] += x[j];
        cq[j] = fmaf(x[j], x[j], cq[j]);
      }
      const size_t o = static_cast<size_t>(row) * p.ldo + col;
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(x[0], x[1], x[2], x[3]);
      if (p.out_f16) {
        __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(p.out_f16 + o) = u;
        if (p.out_f16_lo) {
          float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          __half2 l0 = __floats2half2_rn(x[0] - f0.x, x[1] - f0.y), l1 = __floats2half2_rn(x[2] - f1.x, x[3] - f1.y);
          uint2 w;
          w.x = *reinterpret_cast<uint32_t*>(&l0);
          w.y = *reinterpret_cast<uint32_t*>(&l1);
          *reinterpret_cast<uint2*>(p.out_f16_lo + o) = w;
        }
      }
    }
  }
  if (p.stats) {   // rows of one block belong to one sample (rows_per_sample % 32 == 0, checked on the host)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[ty][tx * 4 + j][0] = cs[j];
      red[ty][tx * 4 + j][1] = cq[j];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int c = blockIdx.x * 128 + threadIdx.x;
      if (c < p.N && blockIdx.y * 32 < p.M) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a += red[k][threadIdx.x][0];
          b += red[k][threadIdx.x][1];
        }
        const int sample = (blockIdx.y * 32) / p.rows_per_sample;
        double* dst = p.stats + ((static_cast<size_t>(blockIdx.y & (STAT_SLOTS - 1)) * p.n_samples + sample) * p.N + c) * 2;
        atomicAdd(dst, static_cast<double>(a));
        atomicAdd(dst + 1, static_cast<double>(b));
      }
    }
  }
}

// scalar variant for outputs whose width is not a multiple of 4 (no fused statistics)
__global__ void __launch_bounds__(256) splitk_epilogue_scalar_kernel(const GemmArgs p, int splits) {
  const size_t total = static_cast<size_t>(p.M) * p.N;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(idx / p.N);
    const int col = static_cast<int>(idx - static_cast<size_t>(row) * p.N);
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += p.ws[static_cast<size_t>(sp) * total + idx];
    float x = acc * p.alpha + (p.bias ? p.bias[col] : 0.f);
    store_elem(p, x, row, row / p.rows_per_sample, col, true);
  }
}

template <int BN>
static int launch_gemm(const TmapPack& tm, const GemmArgs& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  const int grid = std::min(tiles, sm_count());
  SDB_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM, st, tm, p));
  SDB_LAUNCH_CHECK();
  return 0;
}

static int pow2_divisor(int v, int cap) {
  int p = 1;
  while (p * 2 <= cap && v % (p * 2) == 0) p *= 2;
  return p;
}

// Tile model used to pick (block_n, split-K). Per CTA the 64-wide K step of a 128 x bn tile moves (128 + bn) * 128 B
// from L2 (the binding resource for 1-CTA tiles); the epilogue drains bn columns and, in the persistent kernel,
// overlaps the next tile's main loop, so a tile costs max(mainloop, epilogue) plus a fixed hand-off. CTAs run in
// waves of sm_count. Split-K adds an fp32 partial round trip and a second kernel.
static double tile_cost(int n, long m_tiles, int k_iters, int bn, int splits, long M) {
  long nt = (n + bn - 1) / bn;
  long tiles = m_tiles * nt * splits;
  long waves = (tiles + sm_count() - 1) / sm_count();
  double mainloop = static_cast<double>((k_iters + splits - 1) / splits) * (128.0 + bn) * 1.6;
  double epi = 10.0 * bn + 200.0;
  double first = mainloop + epi;                       // first tile of a CTA cannot overlap
  double steady = std::max(mainloop, epi) + 150.0;
  double t = first + (waves - 1) * steady + 800.0;
  if (splits > 1) t += 2500.0 + 6.0 * static_cast<double>(splits) * M * n / (sm_count() * 256.0);
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
  const void* srcs[MAX_SRC] = {d->a0, d->a1, d->a2, d->a3};
  const int chans[MAX_SRC] = {d->c0, d->c1, d->c2, d->c3};
  int nsrc = 0, ctot = 0;
  for (int i = 0; i < MAX_SRC; ++i) {
    if (srcs[i] == nullptr) break;
    SDB_CHECK(chans[i] > 0 && chans[i] % 64 == 0, "sdb_gemm: channel count of source %d must be a positive multiple of 64 (got %d)",
              i, chans[i]);
    ++nsrc;
    ctot += chans[i];
  }
  for (int i = nsrc; i < MAX_SRC; ++i)
    SDB_CHECK(srcs[i] == nullptr && chans[i] == 0, "sdb_gemm: A sources must be contiguous (a%d/c%d)", i, i);
  SDB_CHECK(d->nb > 0 && d->h > 0 && d->w > 0 && d->n > 0, "sdb_gemm: bad dims");
  SDB_CHECK(d->out_f16 || d->out_f32, "sdb_gemm: no output");
  SDB_CHECK(!d->out_f16_lo || d->out_f16, "sdb_gemm: out_f16_lo needs out_f16");

  GemmArgs p{};
  const int cstride = d->conv_stride > 1 ? d->conv_stride : 1;
  SDB_CHECK(cstride == 1 || (cstride == 2 && d->taps == 9), "sdb_gemm: conv_stride must be 1 or 2 (3x3 convs only)");
  SDB_CHECK(d->conv_shift == 0 || d->taps == 9, "sdb_gemm: conv_shift applies to 3x3 convs only");
  const int in_h = d->in_h > 0 ? d->in_h : d->h, in_w = d->in_w > 0 ? d->in_w : d->w;
  SDB_CHECK((in_h == d->h && in_w == d->w) || d->taps == 9, "sdb_gemm: in_h / in_w apply to 3x3 convs only");
  p.cstride = cstride;
  p.cshift = d->conv_shift;
  const long M = static_cast<long>(d->nb) * d->h * d->w;
  SDB_CHECK(M < (1L << 31), "sdb_gemm: M too large");
  p.M = static_cast<int>(M);
  p.N = d->n;
  p.taps = d->taps;
  p.nsrc = nsrc;
  p.cb[0] = 0;
  for (int i = 0; i < nsrc; ++i) p.cb[i + 1] = p.cb[i] + chans[i] / 64;
  for (int i = nsrc; i < MAX_SRC; ++i) p.cb[i + 1] = p.cb[nsrc];
  p.k_iters = d->taps * p.cb[nsrc];
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
  p.out_f16_lo = static_cast<__half*>(d->out_f16_lo);
  p.out_f32 = d->out_f32;
  p.act = d->act;
  p.b_static = d->b_dynamic ? 0 : 1;
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
  p.splits = splits;
  p.m_tiles = static_cast<int>(m_tiles);
  p.n_tiles = (d->n + bn - 1) / bn;
  SDB_CHECK(m_tiles * p.n_tiles * splits < (1L << 30), "sdb_gemm: too many tiles");

  // tensor maps
  TmapPack tm;
  for (int i = 0; i < nsrc; ++i) {
    const int c = chans[i];
    if (d->taps == 1) {
      uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(M), 1, 1};
      uint64_t str[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(c) * 2 * M,
                         static_cast<uint64_t>(c) * 2 * M};
      uint32_t box[4] = {64, 128, 1, 1};
      if (make_tmap_f16(&tm.a[i], srcs[i], 4, dims, str, box)) return 1;
    } else {
      // the INPUT image; a stride-2 conv traverses it with element strides {1, 2, 2, 1}: a box of 2*TW x 2*TH input
      // positions delivers the TW x TH pixels one tap of the output tile needs
      uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(in_w), static_cast<uint64_t>(in_h),
                          static_cast<uint64_t>(d->nb)};
      uint64_t str[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(c) * 2 * in_w,
                         static_cast<uint64_t>(c) * 2 * in_w * in_h};
      uint32_t box[4] = {64, static_cast<uint32_t>(p.TW * cstride), static_cast<uint32_t>(p.TH * cstride),
                         static_cast<uint32_t>(p.TN)};
      uint32_t est[4] = {1, static_cast<uint32_t>(cstride), static_cast<uint32_t>(cstride), 1};
      SDB_CHECK(box[1] <= 256 && box[2] <= 256, "sdb_gemm: strided box too large");
      if (make_tmap(&tm.a[i], srcs[i], 2, 128, 4, dims, str, box, est)) return 1;
    }
  }
  for (int i = nsrc; i < MAX_SRC; ++i) tm.a[i] = tm.a[0];
  {
    uint64_t K = static_cast<uint64_t>(d->taps) * ctot;
    uint64_t dims[2] = {K, static_cast<uint64_t>(d->n)};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    if (make_tmap_f16(&tm.b, d->b, 2, dims, str, box)) return 1;
  }
  // output maps for the TMA-store epilogue (fast path); otherwise the scalar transposed path is used
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool fast = (n_out % 32 == 0) && (p.ldo % 8 == 0) && (!p.bias || al16(p.bias)) &&
              (!p.film || (al16(p.film) && p.ldf % 4 == 0)) && (!p.residual || (al16(p.residual) && p.ldr % 4 == 0)) &&
              (!p.out_f32 || al16(p.out_f32)) && (!p.out_f16 || al16(p.out_f16)) &&
              (!p.out_f16_lo || al16(p.out_f16_lo)) && (!p.ws || (d->n % 32 == 0 && al16(p.ws)));
  p.fast = fast ? 1 : 0;
  p.bw = d->taps == 1 ? 32 : std::min(p.TW, 32);
  p.bh = d->taps == 1 ? 1 : std::min(p.TH, 32 / p.bw);
  tm.o32 = tm.b;
  tm.o16 = tm.b;
  tm.o16lo = tm.b;
  tm.ows = tm.b;
  if (fast) {
    const uint32_t bnn = static_cast<uint32_t>(32 / (p.bw * p.bh));
    auto make_out = [&](CUtensorMap* m, const void* base, int elem, int ncols, long ld, int nsplit) -> int {
      uint64_t e = static_cast<uint64_t>(elem);
      uint64_t pitch = static_cast<uint64_t>(ld) * e;
      uint64_t dims[5], str[4];
      uint32_t box[5] = {32, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), bnn, 1};
      dims[0] = static_cast<uint64_t>(ncols);
      if (d->taps == 1) {
        dims[1] = static_cast<uint64_t>(M);
        dims[2] = 1;
        dims[3] = 1;
        str[0] = pitch;
        str[1] = pitch * M;
        str[2] = pitch * M;
      } else {
        dims[1] = static_cast<uint64_t>(d->w);
        dims[2] = static_cast<uint64_t>(d->h);
        dims[3] = static_cast<uint64_t>(d->nb);
        str[0] = pitch;
        str[1] = pitch * d->w;
        str[2] = pitch * d->w * d->h;
      }
      dims[4] = static_cast<uint64_t>(nsplit);
      str[3] = pitch * M;
      return make_tmap(m, base, elem, elem == 4 ? 128 : 64, 5, dims, str, box);
    };
    if (p.ws && make_out(&tm.ows, p.ws, 4, d->n, d->n, splits)) return 1;
    if (p.out_f32 && make_out(&tm.o32, p.out_f32, 4, n_out, p.ldo, 1)) return 1;
    if (p.out_f16 && make_out(&tm.o16, p.out_f16, 2, n_out, p.ldo, 1)) return 1;
    if (p.out_f16_lo && make_out(&tm.o16lo, p.out_f16_lo, 2, n_out, p.ldo, 1)) return 1;
  }
  // fused GroupNorm statistics: per-(sample, channel) sums of the fp32 output, accumulated by the epilogue
  p.stats = nullptr;
  p.stats_halves = 1;
  p.n_samples = static_cast<int>((M + p.rows_per_sample - 1) / p.rows_per_sample);
  if (d->stats_out) {
    SDB_CHECK(fast && p.out_f32 && !geglu, "sdb_gemm: stats_out needs the TMA-store epilogue with an fp32 output");
    SDB_CHECK(p.rows_per_sample % 64 == 0, "sdb_gemm: stats_out needs rows_per_sample %% 64 == 0 (got %d)",
              p.rows_per_sample);
    p.stats = static_cast<double*>(d->stats_out);
    p.stats_halves = (p.rows_per_sample % 128 == 0) ? 1 : 2;
    if (!d->stats_prezeroed)
      SDB_CUDA(cudaMemsetAsync(p.stats, 0, static_cast<size_t>(STAT_SLOTS) * p.n_samples * d->n * 2 * sizeof(double), st));
  }

  // bias + FiLM fold into a per-tile column table when each 64-row half of every tile belongs to one sample
  p.film_table = 0;
  if (p.film) {
    if (d->taps == 1) p.film_table = (p.rows_per_sample % 64 == 0) ? 1 : 0;
    else p.film_table = (p.TW * p.TH >= 64 && p.rows_per_sample == d->h * d->w) ? 1 : 0;
  }
  p.trace = trace_slot(8 + 8 * 160);
  int rc;
  switch (bn) {
    case 32: rc = launch_gemm<32>(tm, p, st); break;
    case 64: rc = launch_gemm<64>(tm, p, st); break;
    case 128: rc = launch_gemm<128>(tm, p, st); break;
    case 160: rc = launch_gemm<160>(tm, p, st); break;
    default: rc = launch_gemm<256>(tm, p, st); break;
  }
  if (rc) return rc;
  if (splits > 1) {
    const bool vec = (p.N % 4 == 0) && (p.ldo % 4 == 0) && (!p.residual || p.ldr % 4 == 0) && (!p.film || p.ldf % 4 == 0);
    if (vec) {
      dim3 grid((p.N + 127) / 128, (p.M + 31) / 32);
      SDB_CUDA(launch_pdl(splitk_epilogue_kernel, grid, dim3(256), 0, st, p, splits));
    } else {
      SDB_CHECK(!p.stats, "sdb_gemm: stats_out with split-K needs n %% 4 == 0");
      size_t total = static_cast<size_t>(p.M) * p.N;
      int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(sm_count()) * 8));
      splitk_epilogue_scalar_kernel<<<blocks, 256, 0, st>>>(p, splits);
    }
    SDB_LAUNCH_CHECK();
  }
  return 0;
}
