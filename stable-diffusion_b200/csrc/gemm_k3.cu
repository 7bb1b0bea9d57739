// gemm_tc_kernel instantiations of epilogue KIND 3 (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace sdb {
int launch_gemm_kind3(int bn, int cg, const TmapPack& tm, const GemmArgs& p, cudaStream_t st) {
  return launch_gemm_kind<3>(bn, cg, tm, p, st);
}
}  // namespace sdb
