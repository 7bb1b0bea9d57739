// Host-side helpers shared by the launchers: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

namespace sdb {

// plan recording (plan.cu): while a plan is being recorded on this thread every launching entry point also appends a
// closure that repeats the call on another stream
bool plan_recording();
void plan_record(std::function<int(cudaStream_t)> fn);
#define SDB_REC(call_with_stream_s_)                                                        \
  do {                                                                                      \
    if (::sdb::plan_recording()) ::sdb::plan_record([=](cudaStream_t s_) { return call_with_stream_s_; }); \
  } while (0)

// Last error string for the C-ABI (sdb_last_error). Thread-local: one engine per process/device.
void set_error(const char* fmt, ...);
const char* last_error();

#define SDB_CHECK(cond, ...)              \
  do {                                    \
    if (!(cond)) {                        \
      ::sdb::set_error(__VA_ARGS__);      \
      return 1;                           \
    }                                     \
  } while (0)

#define SDB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::sdb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

// fp16 tiled tensor map with SWIZZLE_128B and zero OOB fill. rank<=4; dims/strides innermost first;
// strides_bytes[i] is the byte stride of dim i+1 (dim 0 is contiguous). Returns 0 on success.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
// general form: elem_bytes 2 (fp16) or 4 (fp32); swizzle_bytes 0, 64 or 128; rank <= 5
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr);

int sm_count();

// debug phase tracing (sdb_debug_trace): next n-word slot of the trace buffer, or nullptr when tracing is off / full
unsigned long long* trace_slot(int n_words);
void set_trace(void* buf, long long n_words);
long long trace_used();

// kernels launched through the C ABI since load (bench.py's gpu_launches)
void count_launch();
long long launch_count();

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may begin while its predecessor in the
// stream drains; it must execute griddepcontrol.wait before touching memory the predecessor produces.
// SDB_PDL=0 in the environment disables the attribute (plain stream order).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define SDB_LAUNCH_CHECK()              \
  do {                                  \
    ::sdb::count_launch();              \
    SDB_CUDA(cudaGetLastError());       \
  } while (0)

}  // namespace sdb
