// Flash-style fused attention on tcgen05: S = Q K^T and O += P V on the tensor cores with S, O in TMEM,
// online softmax in registers (one query row per thread), lazy rescale of O (only when the running max
// moves by more than 2^8), P staged through swizzled shared memory as the A operand of the second MMA.
//
//   warp 0     : TMA producer (Q once; K / V^T tiles double-buffered)
//   warp 1     : TMEM alloc + tcgen05.mma issue: S(j+1) is issued before PV(j) so softmax(j+1) overlaps PV(j)
//   warps 2..5 : softmax / correction / epilogue (128 query rows)
//
// Replaces the materialised einsum('b i d, b j d -> b i j') * scale -> softmax -> einsum('b i j, b j d -> b i d')
// of ldm/modules/attention.py:178-192 (and CLIP self-attention with a causal mask).
#include "../../include/sdb200.h"
#include "host.h"
#include "ptx.cuh"

namespace sdb {

constexpr int AQ = 128;   // query rows per CTA
constexpr int AKV = 64;   // kv rows per iteration
// K / V^T ring depth. The next K tile can only be requested once the PV MMA that last read the slot has finished, so
// with 2 slots the ~1 us TMA round trip sat on the per-iteration critical path (155 us for N=4096, d=40); 3 slots take
// it off (DPAD 192 keeps 2: shared memory).
template <int DPAD>
struct AttnCfg {
  static constexpr int ST = DPAD <= 128 ? 3 : 2;
  static constexpr int SMEM = AQ * DPAD * 2 + ST * (2 * AKV * DPAD * 2) + 2 * AQ * AKV * 2 + 1024;
};

struct AttnArgs {
  int nq, nkv, d, heads;
  int ldo;
  long long o_batch_stride;
  __half* out;
  float scale_log2;
  int causal;
};

__device__ __forceinline__ float fast_exp2(float x) {  // one MUFU.EX2, flush-to-zero (exp2(-inf) = 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DPAD>
__global__ void __launch_bounds__(192)
    attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  constexpr int PANELS = DPAD / 64;
  constexpr int Q_BYTES = AQ * DPAD * 2;
  constexpr int K_BYTES = AKV * DPAD * 2;
  constexpr int V_BYTES = DPAD * AKV * 2;
  constexpr int P_BYTES = AQ * AKV * 2;
  constexpr int TMEM_COLS = (128 + DPAD) <= 256 ? 256 : 512;
  constexpr int ST = AttnCfg<DPAD>::ST;
  constexpr uint32_t O_COL = 128;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offset (keeps the shared address space visible to the compiler: STS, not generic ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + Q_BYTES;        // ST stages
  uint8_t* v_s = k_s + ST * K_BYTES;   // ST stages
  uint8_t* p_s = v_s + ST * V_BYTES;   // 2 buffers

  __shared__ uint64_t q_full, k_full[ST], v_full[ST], kv_empty[ST], s_full[2], p_full[2], pv_done[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int q0 = blockIdx.x * AQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  int nkv_eff = p.nkv;
  if (p.causal) nkv_eff = min(p.nkv, q0 + AQ);  // kv blocks entirely above the diagonal are skipped
  const int n_iter = (nkv_eff + AKV - 1) / AKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
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
  const uint32_t tmem = tmem_base_smem;

  pdl_wait();   // q / k / v come from the previous kernels; everything above overlapped their tail
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&q_full, Q_BYTES);
      for (int pn = 0; pn < PANELS; ++pn)
        tma_load_3d(q_s + pn * (AQ * 128), &tmQ, &q_full, head * DPAD + pn * 64, q0, b);
      for (int j = 0; j < n_iter; ++j) {
        int s = j % ST;
        uint32_t ph = (j / ST) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], K_BYTES);
        for (int pn = 0; pn < PANELS; ++pn)
          tma_load_3d(k_s + s * K_BYTES + pn * (AKV * 128), &tmK, &k_full[s], head * DPAD + pn * 64, j * AKV, b);
        mbar_arrive_expect_tx(&v_full[s], V_BYTES);
        tma_load_3d(v_s + s * V_BYTES, &tmV, &v_full[s], j * AKV, head * DPAD, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AQ, AKV);
      constexpr uint32_t idesc_o = umma_idesc_f16(AQ, DPAD);
      const uint32_t q_addr = smem_u32(q_s);
      auto issue_s = [&](int j) {
        const int s = j & 1;            // S accumulator buffer
        const int ks = j % ST;          // K/V ring slot
        mbar_wait(&k_full[ks], (j / ST) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(k_s + ks * K_BYTES);
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk) {
          uint64_t da = umma_desc_k128(q_addr + (kk >> 2) * (AQ * 128) + (kk & 3) * 32);
          uint64_t db = umma_desc_k128(k_addr + (kk >> 2) * (AKV * 128) + (kk & 3) * 32);
          umma_f16(tmem + s * AKV, da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[s]);
      };
      mbar_wait(&q_full, 0);
      if (n_iter > 0) issue_s(0);
      for (int j = 0; j < n_iter; ++j) {
        int s = j & 1;
        uint32_t ph = (j >> 1) & 1;
        if (j + 1 < n_iter) issue_s(j + 1);
        const int ks = j % ST;
        mbar_wait(&p_full[s], ph);
        mbar_wait(&v_full[ks], (j / ST) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(p_s + s * P_BYTES);
        const uint32_t v_addr = smem_u32(v_s + ks * V_BYTES);
#pragma unroll
        for (int kk = 0; kk < AKV / 16; ++kk) {
          uint64_t da = umma_desc_k128(p_addr + kk * 32);
          uint64_t db = umma_desc_k128(v_addr + kk * 32);
          umma_f16(tmem + O_COL, da, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[ks]);
        umma_commit(&pv_done[s]);
      }
    }
  } else {
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
    float m_used = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < n_iter; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      float t[AKV];  // raw scores q.k (unscaled); the softmax scale is folded into one FFMA per element below
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem + lane_addr + s * AKV, r0);
        tmem_ld32(tmem + lane_addr + s * AKV + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          t[c] = __uint_as_float(r0[c]);
          t[32 + c] = __uint_as_float(r1[c]);
        }
      }
      const int kv0 = j * AKV;
      // only the ragged last block and (causal) diagonal blocks need per-element masking
      if ((kv0 + AKV > p.nkv) || (p.causal && kv0 + AKV - 1 > q0)) {
#pragma unroll
        for (int c = 0; c < AKV; ++c) {
          int kv = kv0 + c;
          if ((kv >= p.nkv) || (p.causal && kv > qi)) t[c] = -INFINITY;
        }
      }
      // 8 independent partial maxima / sums: a serial 64-deep FMNMX / FADD chain would expose ~4 cycles per element
      // with only two softmax warps per scheduler to hide it
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = t[i];
#pragma unroll
      for (int c = 8; c < AKV; ++c) mx[c & 7] = fmaxf(mx[c & 7], t[c]);
      const float m_raw = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_blk = m_raw * p.scale_log2;  // scale > 0
      if (j == 0) {
        m_used = m_blk;
      } else {
        float m_new = fmaxf(m_used, m_blk);
        bool need = m_new > m_used + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once PV(j-1) has landed
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          float alpha = need ? exp2f(m_used - m_new) : 1.0f;
#pragma unroll 1
          for (int c = 0; c < DPAD / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem + lane_addr + O_COL + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem + lane_addr + O_COL + c * 32, o);
          }
          tmem_st_wait();
          l *= alpha;
          if (need) m_used = m_new;
        }
      }
      float sm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_used;
#pragma unroll
      for (int c = 0; c < AKV; ++c) {
        t[c] = fast_exp2(fmaf(t[c], p.scale_log2, neg_m));
        sm[c & 7] += t[c];
      }
      l += ((sm[0] + sm[1]) + (sm[2] + sm[3])) + ((sm[4] + sm[5]) + (sm[6] + sm[7]));
      if (j >= 2) mbar_wait(&pv_done[s], ((j - 2) >> 1) & 1);  // P buffer s was read by PV(j-2)
      uint8_t* prow = p_s + s * P_BYTES + r * 128;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        __half2 h0 = __floats2half2_rn(t[c16 * 8 + 0], t[c16 * 8 + 1]);
        __half2 h1 = __floats2half2_rn(t[c16 * 8 + 2], t[c16 * 8 + 3]);
        __half2 h2 = __floats2half2_rn(t[c16 * 8 + 4], t[c16 * 8 + 5]);
        __half2 h3 = __floats2half2_rn(t[c16 * 8 + 6], t[c16 * 8 + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(prow + ((c16 ^ (r & 7)) << 4)) = u;
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_full[s]);
    }
    // epilogue: O / l
    if (n_iter > 0) {
      mbar_wait(&pv_done[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
      tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    __half* orow = p.out + static_cast<size_t>(b) * p.o_batch_stride + static_cast<size_t>(qi) * p.ldo + head * p.d;
#pragma unroll 1
    for (int c = 0; c < DPAD / 32; ++c) {
      if (c * 32 >= p.d) break;
      uint32_t o[32];
      tmem_ld32(tmem + lane_addr + O_COL + c * 32, o);
      tmem_ld_wait();
      if (qi < p.nq && n_iter > 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          int col = c * 32 + i;
          if (col < p.d)
            *reinterpret_cast<__half2*>(orow + col) =
                __floats2half2_rn(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int DPAD>
static int launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& p, dim3 grid,
                       cudaStream_t st) {
  constexpr int SMEM = AttnCfg<DPAD>::SMEM;
  auto kern = attention_tc_kernel<DPAD>;
  static bool configured = false;
  if (!configured) {
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  SDB_CUDA(launch_pdl(kern, grid, dim3(192), SMEM, st, q, k, v, p));
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention(const sdb_attn_desc* d, sdb_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SDB_CHECK(d && d->q && d->k && d->vt && d->out, "sdb_attention: null pointer");
  SDB_CHECK(d->dpad == 64 || d->dpad == 128 || d->dpad == 192, "sdb_attention: dpad must be 64/128/192 (got %d)",
            d->dpad);
  SDB_CHECK(d->d > 0 && d->d <= d->dpad && d->d % 2 == 0, "sdb_attention: bad head dim %d", d->d);
  SDB_CHECK(d->batch > 0 && d->heads > 0 && d->nq > 0 && d->nkv > 0, "sdb_attention: bad sizes");
  SDB_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldvt % 8 == 0, "sdb_attention: leading dims must be multiples of 8");
  SDB_CHECK(d->ldo % 2 == 0 && d->o_batch_stride % 2 == 0, "sdb_attention: output stride must be even");

  CUtensorMap tq, tk, tv;
  const uint64_t hd = static_cast<uint64_t>(d->heads) * d->dpad;
  {
    uint64_t dims[3] = {hd, static_cast<uint64_t>(d->nq), static_cast<uint64_t>(d->batch)};
    uint64_t str[2] = {static_cast<uint64_t>(d->ldq) * 2, static_cast<uint64_t>(d->q_batch_stride) * 2};
    uint32_t box[3] = {64, AQ, 1};
    if (make_tmap_f16(&tq, d->q, 3, dims, str, box)) return 1;
  }
  {
    uint64_t dims[3] = {hd, static_cast<uint64_t>(d->nkv), static_cast<uint64_t>(d->batch)};
    uint64_t str[2] = {static_cast<uint64_t>(d->ldk) * 2, static_cast<uint64_t>(d->k_batch_stride) * 2};
    uint32_t box[3] = {64, AKV, 1};
    if (make_tmap_f16(&tk, d->k, 3, dims, str, box)) return 1;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d->nkv), hd, static_cast<uint64_t>(d->batch)};
    uint64_t str[2] = {static_cast<uint64_t>(d->ldvt) * 2, static_cast<uint64_t>(d->vt_batch_stride) * 2};
    uint32_t box[3] = {AKV, static_cast<uint32_t>(d->dpad), 1};
    if (make_tmap_f16(&tv, d->vt, 3, dims, str, box)) return 1;
  }
  AttnArgs p{};
  p.nq = d->nq;
  p.nkv = d->nkv;
  p.d = d->d;
  p.heads = d->heads;
  p.ldo = d->ldo;
  p.o_batch_stride = d->o_batch_stride;
  p.out = static_cast<__half*>(d->out);
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.causal = d->causal;
  dim3 grid((d->nq + AQ - 1) / AQ, d->heads, d->batch);
  switch (d->dpad) {
    case 64: return launch_attn<64>(tq, tk, tv, p, grid, st);
    case 128: return launch_attn<128>(tq, tk, tv, p, grid, st);
    default: return launch_attn<192>(tq, tk, tv, p, grid, st);
  }
}
