// quantize_reference.cu — instantiates vbq_quantize_kernel for kModeReference (see quantize_kernel.cuh).
#include "quantize_kernel.cuh"

int vbq_launch_quantize_reference(const QArgs &a, int dev, int sms, cudaStream_t st) {
    return launch_quantize_mode<kModeReference>(a, dev, sms, st);
}
