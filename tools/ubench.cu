// ubench.cu — pipe-rate microbenchmarks that size the raymarch search loop on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu ; run under gpurun.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[16];
    u64 y[8];
    double z[8];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(x[2 * i], x[2 * i + 1]); y[i] = *reinterpret_cast<u64*>(&v); z[i] = x[i]; }
    float2 ab = make_float2(a, b);
    u64 A = *reinterpret_cast<u64*>(&ab);
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(y[i]) : "l"(A));
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(A));
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = fma(z[i], (double)a, (double)b);
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + a;
        } else if (MODE == 6) {  // broadcast LDS.128 + 4 FADD consuming it
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = sh[(it + i) & 255];
                x[4 * i] += v.x; x[4 * i + 1] += v.y; x[4 * i + 2] += v.z; x[4 * i + 3] += v.w;
            }
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fminf(x[i], fminf(a, x[(i + 1) & 15]));
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&y[i]); s += v.x + v.y + (float)z[i]; }
    if (s == 1234.5f) out[0] = s;
}
template <int MODE>
void run(const char* name, double ops_per_thread_iter, float* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 0.999f, 1e-4f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 0.999f, 1e-4f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = ops_per_thread_iter * ITERS * blocks * 256.0 / 32.0;
    double per_smsp_clk = warp_instr / (ms * 1e-3) / (148.0 * 4) / 1.965e9;
    printf("%-34s %8.3f ms  %7.3f warp-instr/clk/SMSP (at 1.965 GHz)  %8.2f Tlane-op/s\n", name, ms, per_smsp_clk, warp_instr * 32 / (ms * 1e-3) / 1e12);
}
int main() {
    float* d;
    cudaMalloc(&d, 4);
    run<0>("FFMA (16 chains)", 16, d);
    run<1>("FFMA2 f32x2 (8 chains)", 8, d);
    run<2>("FADD2 f32x2 (8 chains)", 8, d);
    run<5>("FADD (16 chains)", 16, d);
    run<3>("MUFU.SQRT (16 chains)", 16, d);
    run<4>("DFMA (8 chains)", 8, d);
    run<6>("LDS.128 bcast + 4 FADD", 5, d);
    run<7>("FMNMX x2 (16 chains)", 32, d);
    return 0;
}
