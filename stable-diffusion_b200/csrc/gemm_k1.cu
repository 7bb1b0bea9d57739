// gemm_tc_kernel instantiations of epilogue KIND 1 (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace sdb {
int launch_gemm_kind1(int bn, int cg, const TmapPack& tm, const GemmArgs& p, cudaStream_t st) {
  return launch_gemm_kind<1>(bn, cg, tm, p, st);
}
}  // namespace sdb
