// quantize_strict.cu — instantiates vbq_quantize_kernel for kModeStrict (see quantize_kernel.cuh).
#include "quantize_kernel.cuh"

int vbq_launch_quantize_strict(const QArgs &a, int dev, int sms, cudaStream_t st) {
    return launch_quantize_mode<kModeStrict>(a, dev, sms, st);
}
