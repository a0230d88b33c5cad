// emd_umma.cu -- tcgen05/TMEM implicit-GEMM convolution (placeholder until the kernel lands:
// reports "unsupported", so the 16-bit modes run on the CUDA-core kernel).
#include "emd_kernels.h"
namespace emd {
bool umma_supported(const ConvParams&, int) { return false; }
cudaError_t launch_conv_umma(const ConvParams&, int, int, cudaStream_t) { return cudaErrorNotSupported; }
size_t umma_pack_weights(const float*, int, int, int, int, void*) { return 0; }
}  // namespace emd
