// rm_probe.cu — live measurement of the FP32 (non-tensor) pipe peak, the roofline denominator of the
// raymarch path.  MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only; SURVEY.md §6/§8d asks
// for a dependent-free FFMA microbenchmark on all SMs of the same box.  Each thread runs 16
// independent FFMA chains so neither latency (4 cycles) nor register-bank conflicts bound the rate;
// FLOPs = 2 per FFMA.
#include <cuda_runtime.h>

#include "rm_types.h"

namespace rm {

template <int ITERS>
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

// Returns cudaError_t as int.  tflops = best of `reps` timed launches.
int probe_fp32_peak(int n_sms, void* stream_, float* scratch, double* tflops) {
    cudaStream_t stream = (cudaStream_t)stream_;
    constexpr int ITERS = 2048;
    const int blocks = n_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return (int)e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return (int)e;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, stream);
        ffma_peak_kernel<ITERS><<<blocks, threads, 0, stream>>>(scratch, 0.999f, 1e-4f);
        cudaEventRecord(e1, stream);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16.0 * 8.0 * (double)ITERS * (double)blocks * (double)threads;
        double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;  // first launch is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    *tflops = best;
    return (int)e;
}

}  // namespace rm
