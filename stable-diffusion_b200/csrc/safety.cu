// Post-processing kernels of the txt2img script around the decoded image (scripts/txt2img.py:69-95, 261-264, 314-327):
// the CLIP image preprocessing of the safety checker (PIL bicubic resize with its 8-bit fixed-point arithmetic ->
// normalise -> NCHW), ViT patch extraction, the concept-embedding decision of diffusers'
// StableDiffusionSafetyChecker, blanking of flagged images, and the DWT-DCT invisible watermark of the
// `invisible-watermark` package (EmbedMaxDct, method 'dwtDct'). Streaming byte / float work: HBM-bound.
#include "../../include/sdb200.h"
#include "host.h"
#include <cuda_fp16.h>

namespace sdb {

static inline int sgrid(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  size_t cap = static_cast<size_t>(sm_count()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}
#define SGRID_STRIDE(i, n)                                                                \
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < (n); \
       i += static_cast<size_t>(gridDim.x) * blockDim.x)

__device__ __forceinline__ uint8_t clip8(int v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// One pass of PIL's ImagingResample for 8-bit images (Resample.c: ImagingResampleHorizontal_8bpc / Vertical_8bpc):
// out = clip8((2^21 + sum_k in[xmin + k] * coef[k]) >> 22) with the fixed-point coefficient table the host precomputes
// exactly as precompute_coeffs + normalize_coeffs_8bpc do. `axis_stride` / `line_stride` address the resampled axis and
// the other pixel axis of an interleaved [B, H, W, 3] image. The source is either uint8 or fp32 in [0, 1] (converted
// as numpy_to_pil does: (x * 255).round().astype(uint8), txt2img.py:78 via diffusers' numpy_to_pil).
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src8, const float* __restrict__ src32, int n_img,
                                   int lines, int in_size, int out_size, int ksize, const int* __restrict__ bounds,
                                   const int* __restrict__ coefs, long in_axis_stride, long in_line_stride,
                                   long in_img_stride, long out_axis_stride, long out_line_stride, long out_img_stride,
                                   uint8_t* __restrict__ out) {
  const size_t total = static_cast<size_t>(n_img) * lines * out_size * 3;
  SGRID_STRIDE(i, total) {
    const int ch = static_cast<int>(i % 3);
    size_t r = i / 3;
    const int xo = static_cast<int>(r % out_size);
    r /= out_size;
    const int line = static_cast<int>(r % lines);
    const int img = static_cast<int>(r / lines);
    const int xmin = bounds[2 * xo], xcnt = bounds[2 * xo + 1];
    const int* k = coefs + static_cast<size_t>(xo) * ksize;
    int ss = 1 << 21;
    const size_t base = static_cast<size_t>(img) * in_img_stride + static_cast<size_t>(line) * in_line_stride + ch;
    for (int x = 0; x < xcnt; ++x) {
      const size_t a = base + static_cast<size_t>(xmin + x) * in_axis_stride;
      const int px = src8 ? static_cast<int>(src8[a]) : static_cast<int>(clip8(static_cast<int>(rintf(src32[a] * 255.0f))));
      ss += px * k[x];
    }
    out[static_cast<size_t>(img) * out_img_stride + static_cast<size_t>(line) * out_line_stride +
        static_cast<size_t>(xo) * out_axis_stride + ch] = clip8(ss >> 22);
  }
}

// center crop + rescale 1/255 + normalise (CLIPFeatureExtractor) : uint8 [B, H, W, 3] -> fp32 NCHW [B, 3, S, S]
__global__ void clip_normalize_kernel(const uint8_t* __restrict__ img, int nb, int h, int w, int s, int top, int left,
                                      float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
  const size_t total = static_cast<size_t>(nb) * 3 * s * s;
  SGRID_STRIDE(i, total) {
    const int x = static_cast<int>(i % s);
    size_t r = i / s;
    const int y = static_cast<int>(r % s);
    r /= s;
    const int c = static_cast<int>(r % 3);
    const int b = static_cast<int>(r / 3);
    const float v = static_cast<float>(img[((static_cast<size_t>(b) * h + top + y) * w + left + x) * 3 + c]) * (1.0f / 255.0f);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[i] = (v - mean) / sd;
  }
}

// ViT patch extraction (CLIPVisionEmbeddings.patch_embedding = Conv2d(3, h, P, stride P, bias=False) as a GEMM):
// NCHW fp32 [B, 3, S, S] -> fp16 [B * (S/P)^2, kpad], k = (c * P + py) * P + px (the conv weight's own flattening)
__global__ void patchify_kernel(const float* __restrict__ x, int nb, int s, int p, int kpad, __half* __restrict__ out) {
  const int g = s / p;
  const size_t total = static_cast<size_t>(nb) * g * g * kpad;
  SGRID_STRIDE(i, total) {
    const int k = static_cast<int>(i % kpad);
    size_t r = i / kpad;
    const int gx = static_cast<int>(r % g);
    r /= g;
    const int gy = static_cast<int>(r % g);
    const int b = static_cast<int>(r / g);
    float v = 0.f;
    if (k < 3 * p * p) {
      const int px = k % p, py = (k / p) % p, c = k / (p * p);
      v = x[((static_cast<size_t>(b) * 3 + c) * s + gy * p + py) * s + gx * p + px];
    }
    out[i] = __float2half_rn(v);
  }
}

// StableDiffusionSafetyChecker.forward after the vision tower (diffusers safety_checker.py): cosine distances of the
// image embedding to 3 "special care" and 17 concept embeddings, scores rounded to 3 decimals, + 0.01 adjustment when
// a special-care concept fires, flagged when any concept score > 0. One block per image.
__global__ void safety_scores_kernel(const float* __restrict__ emb, int dim, const float* __restrict__ special,
                                     const float* __restrict__ special_w, int n_special, const float* __restrict__ concept,
                                     const float* __restrict__ concept_w, int n_concept, float* __restrict__ scores,
                                     int* __restrict__ flagged) {
  __shared__ float red[32];
  __shared__ float cosv[64];
  const int b = blockIdx.x;
  const float* e = emb + static_cast<size_t>(b) * dim;
  auto block_sum = [&](float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    return t;
  };
  float ee = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) ee = fmaf(e[i], e[i], ee);
  const float en = sqrtf(block_sum(ee));
  for (int c = 0; c < n_special + n_concept; ++c) {
    const float* t = c < n_special ? special + static_cast<size_t>(c) * dim : concept + static_cast<size_t>(c - n_special) * dim;
    float dot = 0.f, tt = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
      dot = fmaf(e[i], t[i], dot);
      tt = fmaf(t[i], t[i], tt);
    }
    const float d = block_sum(dot);
    const float tn = sqrtf(block_sum(tt));
    if (threadIdx.x == 0) cosv[c] = d / (en * tn);   // cosine_distance(): normalised dot product
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float adjustment = 0.f;
    for (int c = 0; c < n_special; ++c) {
      const float sc = rintf((cosv[c] - special_w[c] + adjustment) * 1000.f) / 1000.f;   // (the library's loop carries it)
      scores[static_cast<size_t>(b) * (n_special + n_concept) + c] = sc;
      if (sc > 0.f) adjustment = 0.01f;
    }
    int bad = 0;
    for (int c = 0; c < n_concept; ++c) {
      const float sc = rintf((cosv[n_special + c] - concept_w[c] + adjustment) * 1000.f) / 1000.f;
      scores[static_cast<size_t>(b) * (n_special + n_concept) + n_special + c] = sc;
      if (sc > 0.f) bad = 1;
    }
    flagged[b] = bad;
  }
}

// images[idx] = 0 for flagged images (safety_checker.py: `images[idx] = np.zeros(images[idx].shape)`)
__global__ void blank_flagged_kernel(float* __restrict__ img, size_t per_image, int nb, const int* __restrict__ flagged) {
  const size_t total = per_image * nb;
  SGRID_STRIDE(i, total) {
    if (flagged[i / per_image]) img[i] = 0.f;
  }
}

// ---- invisible watermark (imwatermark EmbedMaxDct, method 'dwtDct', scales [0, 36, 36], block 4) ----
// cv2's 8-bit BGR <-> YUV in its own 14-bit fixed point (checked exhaustively against cv2 in tests/test_safety_cpu.py)
__device__ __forceinline__ int rs14(int x) { return (x + (1 << 13)) >> 14; }
__global__ void wm_rgb2yuv_kernel(const uint8_t* __restrict__ rgb, size_t npix, uint8_t* __restrict__ yuv) {
  SGRID_STRIDE(i, npix) {
    const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    const int y = rs14(b * 1868 + g * 9617 + r * 4899);
    yuv[3 * i] = clip8(y);
    yuv[3 * i + 1] = clip8(rs14((b - y) * 8061 + (128 << 14)));
    yuv[3 * i + 2] = clip8(rs14((r - y) * 14369 + (128 << 14)));
  }
}
__global__ void wm_yuv2rgb_kernel(const uint8_t* __restrict__ yuv, size_t npix, uint8_t* __restrict__ rgb) {
  SGRID_STRIDE(i, npix) {
    const int y = yuv[3 * i], u = yuv[3 * i + 1] - 128, v = yuv[3 * i + 2] - 128;
    rgb[3 * i] = clip8(y + rs14(v * 18678));
    rgb[3 * i + 1] = clip8(y + rs14(u * -6472 + v * -9519));
    rgb[3 * i + 2] = clip8(y + rs14(u * 33292));
  }
}
// One thread per 8x8 pixel block of the U plane = one 4x4 block of the Haar approximation band: quantise its
// largest-magnitude coefficient (excluding the first) to carry one watermark bit; the inverse transform moves the four
// pixels of that coefficient's 2x2 cell by delta / 2 and is written back with C truncation (float64 -> uint8).
__global__ void wm_embed_kernel(uint8_t* __restrict__ yuv, int nb, int h, int w, const uint8_t* __restrict__ bits,
                                int n_bits, double scale) {
  const int r4 = (h / 4 * 4) / 8, c4 = (w / 4 * 4) / 8;
  const size_t total = static_cast<size_t>(nb) * r4 * c4;
  SGRID_STRIDE(t, total) {
    const int j = static_cast<int>(t % c4);
    const int i = static_cast<int>((t / c4) % r4);
    const int b = static_cast<int>(t / (static_cast<size_t>(c4) * r4));
    uint8_t* base = yuv + ((static_cast<size_t>(b) * h + 8 * i) * w + 8 * j) * 3 + 1;
    const size_t rowp = static_cast<size_t>(w) * 3;
    int best = 1;
    double bestv = 0.0, bestabs = -1.0;
    for (int k = 1; k < 16; ++k) {
      const int cy = k >> 2, cx = k & 3;
      const uint8_t* q = base + (2 * cy) * rowp + (2 * cx) * 3;
      const double ca = (static_cast<double>(q[0]) + q[3] + q[rowp] + q[rowp + 3]) * 0.5;
      if (fabs(ca) > bestabs) {   // np.argmax: first occurrence of the maximum
        bestabs = fabs(ca);
        bestv = ca;
        best = k;
      }
    }
    const int num = i * c4 + j;
    const double bit = static_cast<double>(bits[num % n_bits]);
    double qv = (floor(bestabs / scale) + 0.25 + 0.5 * bit) * scale;
    if (bestv < 0.0) qv = -qv;
    const double half_delta = (qv - bestv) * 0.5;
    uint8_t* q = base + (2 * (best >> 2)) * rowp + (2 * (best & 3)) * 3;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        uint8_t* px = q + dy * rowp + dx * 3;
        const double nv = static_cast<double>(*px) + half_delta;
        *px = static_cast<uint8_t>(static_cast<long long>(nv) & 0xff);
      }
  }
}

}  // namespace sdb

using namespace sdb;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int sdb_resample_u8(const void* src_u8, const float* src_f32, int32_t n_img, int32_t lines, int32_t in_size,
                               int32_t out_size, int32_t ksize, const int32_t* bounds, const int32_t* coefs,
                               int32_t vertical, int32_t other_size_in, void* out_u8, sdb_stream_t stream) {
  SDB_REC(sdb_resample_u8(src_u8, src_f32, n_img, lines, in_size, out_size, ksize, bounds, coefs, vertical, other_size_in, out_u8, s_));
  SDB_CHECK((src_u8 != nullptr) != (src_f32 != nullptr), "sdb_resample_u8: exactly one source");
  SDB_CHECK(bounds && coefs && out_u8 && n_img > 0 && lines > 0 && in_size > 0 && out_size > 0 && ksize > 0,
            "sdb_resample_u8: bad arguments");
  // interleaved [B, H, W, 3] images. horizontal: axis = x (stride 3), lines = rows (stride W*3); vertical: axis = y
  // (stride W*3), lines = columns (stride 3); other_size_in = the image extent along the non-resampled axis
  long in_axis, in_line, in_img, out_axis, out_line, out_img;
  if (!vertical) {
    in_axis = 3;
    in_line = 3L * in_size;
    in_img = in_line * lines;
    out_axis = 3;
    out_line = 3L * out_size;
    out_img = out_line * lines;
  } else {
    in_axis = 3L * other_size_in;
    in_line = 3;
    in_img = in_axis * in_size;
    out_axis = 3L * other_size_in;
    out_line = 3;
    out_img = out_axis * out_size;
  }
  const size_t total = static_cast<size_t>(n_img) * lines * out_size * 3;
  resample_u8_kernel<<<sgrid(total), 256, 0, ST>>>(static_cast<const uint8_t*>(src_u8), src_f32, n_img, lines, in_size,
                                                   out_size, ksize, bounds, coefs, in_axis, in_line, in_img, out_axis,
                                                   out_line, out_img, static_cast<uint8_t*>(out_u8));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_clip_normalize(const void* img_u8, int32_t nb, int32_t h, int32_t w, int32_t size, float m0, float m1,
                                  float m2, float s0, float s1, float s2, float* out_nchw, sdb_stream_t stream) {
  SDB_REC(sdb_clip_normalize(img_u8, nb, h, w, size, m0, m1, m2, s0, s1, s2, out_nchw, s_));
  SDB_CHECK(img_u8 && out_nchw && h >= size && w >= size, "sdb_clip_normalize: bad arguments");
  const int top = (h - size) / 2, left = (w - size) / 2;   // center crop (CLIPFeatureExtractor.center_crop)
  const size_t total = static_cast<size_t>(nb) * 3 * size * size;
  clip_normalize_kernel<<<sgrid(total), 256, 0, ST>>>(static_cast<const uint8_t*>(img_u8), nb, h, w, size, top, left, m0, m1,
                                                      m2, s0, s1, s2, out_nchw);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_patchify(const float* x_nchw, int32_t nb, int32_t size, int32_t patch, int32_t kpad, void* out_f16,
                            sdb_stream_t stream) {
  SDB_REC(sdb_patchify(x_nchw, nb, size, patch, kpad, out_f16, s_));
  SDB_CHECK(x_nchw && out_f16 && patch > 0 && size % patch == 0 && kpad >= 3 * patch * patch && kpad % 64 == 0,
            "sdb_patchify: bad arguments (size %d patch %d kpad %d)", size, patch, kpad);
  const size_t total = static_cast<size_t>(nb) * (size / patch) * (size / patch) * kpad;
  patchify_kernel<<<sgrid(total), 256, 0, ST>>>(x_nchw, nb, size, patch, kpad, static_cast<__half*>(out_f16));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_safety_scores(const float* image_embeds, int32_t nb, int32_t dim, const float* special_embeds,
                                 const float* special_weights, int32_t n_special, const float* concept_embeds,
                                 const float* concept_weights, int32_t n_concept, float* scores, int32_t* flagged,
                                 sdb_stream_t stream) {
  SDB_REC(sdb_safety_scores(image_embeds, nb, dim, special_embeds, special_weights, n_special, concept_embeds, concept_weights, n_concept, scores, flagged, s_));
  SDB_CHECK(image_embeds && special_embeds && special_weights && concept_embeds && concept_weights && scores && flagged,
            "sdb_safety_scores: null pointer");
  SDB_CHECK(nb > 0 && dim > 0 && n_special >= 0 && n_concept > 0 && n_special + n_concept <= 64, "sdb_safety_scores: bad sizes");
  safety_scores_kernel<<<nb, 256, 0, ST>>>(image_embeds, dim, special_embeds, special_weights, n_special, concept_embeds,
                                           concept_weights, n_concept, scores, flagged);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_blank_flagged(float* images, int64_t per_image, int32_t nb, const int32_t* flagged, sdb_stream_t stream) {
  SDB_REC(sdb_blank_flagged(images, per_image, nb, flagged, s_));
  SDB_CHECK(images && flagged && per_image > 0 && nb > 0, "sdb_blank_flagged: bad arguments");
  blank_flagged_kernel<<<sgrid(static_cast<size_t>(per_image) * nb), 256, 0, ST>>>(images, static_cast<size_t>(per_image), nb,
                                                                                  flagged);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_watermark_dwtdct(const void* rgb_u8, int32_t nb, int32_t h, int32_t w, const void* bits_u8, int32_t n_bits,
                                    float scale, void* yuv_scratch_u8, void* out_rgb_u8, sdb_stream_t stream) {
  SDB_REC(sdb_watermark_dwtdct(rgb_u8, nb, h, w, bits_u8, n_bits, scale, yuv_scratch_u8, out_rgb_u8, s_));
  SDB_CHECK(rgb_u8 && bits_u8 && yuv_scratch_u8 && out_rgb_u8 && nb > 0 && n_bits > 0 && scale > 0.f,
            "sdb_watermark_dwtdct: bad arguments");
  SDB_CHECK(static_cast<long>(h) * w >= 256 * 256, "sdb_watermark_dwtdct: image too small (the encoder needs >= 256x256 pixels)");
  const size_t npix = static_cast<size_t>(nb) * h * w;
  wm_rgb2yuv_kernel<<<sgrid(npix), 256, 0, ST>>>(static_cast<const uint8_t*>(rgb_u8), npix, static_cast<uint8_t*>(yuv_scratch_u8));
  SDB_LAUNCH_CHECK();
  const size_t blocks = static_cast<size_t>(nb) * ((h / 4 * 4) / 8) * ((w / 4 * 4) / 8);
  wm_embed_kernel<<<sgrid(blocks), 256, 0, ST>>>(static_cast<uint8_t*>(yuv_scratch_u8), nb, h, w,
                                                 static_cast<const uint8_t*>(bits_u8), n_bits, static_cast<double>(scale));
  SDB_LAUNCH_CHECK();
  wm_yuv2rgb_kernel<<<sgrid(npix), 256, 0, ST>>>(static_cast<const uint8_t*>(yuv_scratch_u8), npix, static_cast<uint8_t*>(out_rgb_u8));
  SDB_LAUNCH_CHECK();
  return 0;
}
