// placeholder until the tcgen05 path lands
#include "conv_impl.h"
namespace b200 {
bool conv_umma_supported(const UmmaConvArgs&) { return false; }
int conv_umma_run(const UmmaConvArgs&, cudaStream_t) { return B200SEG_ERR_INVALID; }
bool wgrad_umma_supported(const UmmaWgradArgs&) { return false; }
int wgrad_umma_run(const UmmaWgradArgs&, cudaStream_t) { return B200SEG_ERR_INVALID; }
}
