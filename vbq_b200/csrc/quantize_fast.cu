// quantize_fast.cu — instantiates vbq_quantize_kernel for kModeFast (see quantize_kernel.cuh).
#include "quantize_kernel.cuh"

int vbq_launch_quantize_fast(const QArgs &a, int dev, int sms, cudaStream_t st) {
    return launch_quantize_mode<kModeFast>(a, dev, sms, st);
}
