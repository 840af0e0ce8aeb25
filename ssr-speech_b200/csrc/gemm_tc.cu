// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace ssrb {
bool gemm_tc_supported(const GemmArgs&) { return false; }
size_t gemm_tc_workspace_bytes(int, int) { return 0; }
int gemm_tc(const GemmArgs&, void*, size_t, cudaStream_t) { set_error("tcgen05 GEMM not built"); return 1; }
}
