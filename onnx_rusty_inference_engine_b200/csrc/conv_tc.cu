// tcgen05 3xTF32 implicit-GEMM convolution -- placeholder while the kernel is brought up.
#include "internal.h"

namespace b200 {
struct TcWeights { int dummy; };
int tc_supported(const ConvArgs&) { return B200_EUNSUPPORTED; }
int tc_prepare_weights(const float*, int, int, cudaStream_t, std::shared_ptr<TcWeights>*) {
  B200_FAIL(B200_EUNSUPPORTED, "tcgen05 path not built");
}
int launch_conv_tc(const ConvArgs&, const TcWeights&, cudaStream_t) { B200_FAIL(B200_EUNSUPPORTED, "tcgen05 path not built"); }
}  // namespace b200
