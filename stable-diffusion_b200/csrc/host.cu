#include "host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace sdb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is resolved through the runtime so the library has no link-time libcuda dependency
// (it must load, and export its symbols, on a box without a driver).
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = encode_fn();
  SDB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  SDB_CHECK(rank >= 1 && rank <= 5, "TMA rank %d", rank);
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;   // traversal stride: box[i] tensor positions yield box[i] / es[i] elements
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  SDB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    SDB_CHECK((gstr[i] & 15) == 0, "TMA stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i]);
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDB_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled failed (%d): rank %d elem %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u]",
            (int)r, rank, elem_bytes, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
            (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
            (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
            rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  return 0;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SDB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static unsigned long long* g_trace = nullptr;
static long long g_trace_words = 0, g_trace_used = 0;
void set_trace(void* buf, long long n_words) {
  g_trace = static_cast<unsigned long long*>(buf);
  g_trace_words = buf ? n_words : 0;
  g_trace_used = 0;
}
long long trace_used() { return g_trace_used; }
unsigned long long* trace_slot(int n_words) {
  if (!g_trace || g_trace_used + n_words > g_trace_words) return nullptr;
  unsigned long long* p = g_trace + g_trace_used;
  g_trace_used += n_words;
  return p;
}

static long long g_launches = 0;
void count_launch() { ++g_launches; }
long long launch_count() { return g_launches; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace sdb
