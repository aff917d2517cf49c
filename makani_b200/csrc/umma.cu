// tcgen05 / TMA path -- placeholder until the kernels land (phase B).  umma_plan_init() reports "unavailable" so that
// every TF32 request fails loudly (b200sht_* returns B200SHT_ERR_UNSUPPORTED) instead of silently using another path.
#include "common.cuh"
namespace b200sht {
int umma_plan_init(Plan* pl) { pl->umma_state = nullptr; return -1; }
void umma_plan_destroy(Plan*) {}
int umma_available() { return 0; }
int legendre_analysis_umma(const Plan*, const float*, float*, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return B200SHT_ERR_UNSUPPORTED; }
int legendre_synthesis_umma(const Plan*, const float*, float*, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return B200SHT_ERR_UNSUPPORTED; }
int mix_forward_umma(const Plan*, int, const float*, const void*, const void*, float*, int, int, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return B200SHT_ERR_UNSUPPORTED; }
int mix_backward_umma(const Plan*, int, const float*, const void*, const float*, float*, void*, void*, int, int, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return B200SHT_ERR_UNSUPPORTED; }
}  // namespace b200sht
