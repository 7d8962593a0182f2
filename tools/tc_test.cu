// tc_test.cu — stand-alone check of the tcgen05 building blocks used by the tensor-core sphere search
// (split-TF32 distance matrix, no-swizzle K-major operands, TMEM accumulators) + epilogue microbenchmarks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tc_test tools/tc_test.cu ; run under gpurun.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define DEV __device__ __forceinline__
DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbar_init(uint32_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
DEV void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
DEV void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
DEV void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
DEV uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {  // no-swizzle K-major smem matrix descriptor
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {  // kind::tf32, fp32 accumulate, A and B K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
DEV void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
DEV void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
DEV void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DEV float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

constexpr int M = 128, N = 128, K = 16;
constexpr uint32_t kSBO = 128, kLBO_A = (M / 8) * 128, kLBO_B = (N / 8) * 128;

// ---- 1. correctness: D[128 x 128] = A[128 x 16] * B[128 x 16]^T through TMEM ----
__global__ void __launch_bounds__(128) mma_check(const float* __restrict__ gA, const float* __restrict__ gB, float* __restrict__ gD) {
    __shared__ __align__(128) float sA[M * K];
    __shared__ __align__(128) float sB[N * K];
    __shared__ __align__(8) unsigned long long barFull, barMma;
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&barFull), 1);
        mbar_init(smem_u32(&barMma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmemBase;
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&barFull), (unsigned)(sizeof(sA) + sizeof(sB)));
        bulk_g2s(smem_u32(sA), gA, sizeof(sA), smem_u32(&barFull));
        bulk_g2s(smem_u32(sB), gB, sizeof(sB), smem_u32(&barFull));
        mbar_wait(smem_u32(&barFull), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t id = idesc_tf32(M, N);
#pragma unroll
        for (int k = 0; k < K / 8; ++k) {  // one instruction covers K = 8 tf32 = two 16-byte K chunks
            const uint64_t da = umma_desc(smem_u32(sA) + 2u * k * kLBO_A, kLBO_A, kSBO);
            const uint64_t db = umma_desc(smem_u32(sB) + 2u * k * kLBO_B, kLBO_B, kSBO);
            umma_tf32(tm, da, db, id, k > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&barMma));
    }
    __syncwarp();
    mbar_wait(smem_u32(&barMma), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) gD[row * N + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u));
}

// ---- 2. epilogue microbenchmarks: TMEM read rate and the min-reduction on top of it ----
template <int MODE>  // 0: tcgen05.ld only   1: ld + FMNMX3 reduce   2: FMNMX3 only (registers)
__global__ void __launch_bounds__(256) epi_bench(float* out, long long* cycles, int iters) {
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmemBase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    float acc = 3.0e38f + lane;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(lane + j));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 128; c += 32) {  // 128 columns of this warp's 32 lanes = one block's share
            if (MODE != 2) {
                tmem_ld32(tm + (uint32_t)c, v);
                tmem_wait_ld();
            }
            if (MODE != 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) acc = min3(acc, __uint_as_float(v[j]) + (MODE == 2 ? acc * 0.f : 0.f), __uint_as_float(v[j + 1]));
            } else {
                acc = fminf(acc, __uint_as_float(v[it & 31]));
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(256u));
}

static float tf32_rna(float x) {  // cvt.rna.tf32.f32: round to nearest (ties away), low 13 mantissa bits cleared
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u = (u + 0x1000u) & ~0x1FFFu;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}
static size_t canon(int r, int k, int rows) { return (size_t)(k / 4) * (rows / 8) * 32 + (size_t)(r / 8) * 32 + (size_t)(r % 8) * 4 + (k % 4); }

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

int main() {
    std::vector<float> A(M * K, 0.f), B(N * K, 0.f), D(M * N, 0.f);
    std::vector<double> P(M * 3), T(N * 3), TT(N);
    uint32_t s = 12345;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (double)(s >> 8) / 16777216.0;
    };
    for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) P[i * 3 + c] = (float)(-4.0 + 8.0 * rnd());
    for (int j = 0; j < N; ++j) {
        double tt = 0;
        for (int c = 0; c < 3; ++c) {
            T[j * 3 + c] = (float)(-2.5 + 5.0 * rnd());
            tt += T[j * 3 + c] * T[j * 3 + c];
        }
        TT[j] = (float)tt;
    }
    for (int i = 0; i < M; ++i) {
        for (int c = 0; c < 3; ++c) {
            const float p2 = 2.f * (float)P[i * 3 + c];
            const float hi = tf32_rna(p2), lo = tf32_rna(p2 - hi);
            A[canon(i, c, M)] = hi;
            A[canon(i, 3 + c, M)] = lo;
            A[canon(i, 6 + c, M)] = hi;
            A[canon(i, 9 + c, M)] = lo;
        }
        A[canon(i, 12, M)] = 1.f;
        A[canon(i, 13, M)] = 1.f;
        A[canon(i, 14, M)] = 1.f;
    }
    for (int j = 0; j < N; ++j) {
        for (int c = 0; c < 3; ++c) {
            const float t = (float)T[j * 3 + c];
            const float hi = tf32_rna(t), lo = tf32_rna(t - hi);
            B[canon(j, c, N)] = hi;
            B[canon(j, 3 + c, N)] = hi;
            B[canon(j, 6 + c, N)] = lo;
            B[canon(j, 9 + c, N)] = lo;
        }
        const float tt = (float)TT[j];
        const float h = tf32_rna(tt), m = tf32_rna(tt - h), l = tf32_rna((tt - h) - m);
        B[canon(j, 12, N)] = h;
        B[canon(j, 13, N)] = m;
        B[canon(j, 14, N)] = l;
    }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    mma_check<<<1, 128>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxErr = 0, maxRef = 0;
    int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double ref = TT[j];
            for (int c = 0; c < 3; ++c) ref += 2.0 * P[i * 3 + c] * T[j * 3 + c];
            double err = std::fabs((double)D[i * N + j] - ref);
            if (!(err < 1e-3)) {
                if (bad < 5) printf("  mismatch D[%d][%d] = %g, ref %g\n", i, j, D[i * N + j], ref);
                ++bad;
            }
            if (err > maxErr) maxErr = err;
            if (std::fabs(ref) > maxRef) maxRef = std::fabs(ref);
        }
    printf("mma_check: max |D - ref| = %.3e (max |ref| %.1f), %d of %d beyond 1e-3\n", maxErr, maxRef, bad, M * N);

    // microbenchmarks: one CTA of 8 warps per SM and two per SM
    float* dOut;
    long long* dCyc;
    CK(cudaMalloc(&dOut, 296 * 256 * 4));
    CK(cudaMalloc(&dCyc, 296 * 8));
    const int iters = 2000;
    for (int grid : {148, 296}) {
        for (int mode = 0; mode < 3; ++mode) {
            if (mode == 0) epi_bench<0><<<grid, 256>>>(dOut, dCyc, iters);
            if (mode == 1) epi_bench<1><<<grid, 256>>>(dOut, dCyc, iters);
            if (mode == 2) epi_bench<2><<<grid, 256>>>(dOut, dCyc, iters);
            CK(cudaDeviceSynchronize());
            std::vector<long long> cyc(grid);
            CK(cudaMemcpy(cyc.data(), dCyc, grid * 8, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (long long c : cyc) avg += (double)c;
            avg /= grid;
            // per iteration every CTA consumes 128 rows x 256 columns = 32768 accumulators
            printf("epi_bench mode %d grid %d: %.1f cycles per 128x256 tile per CTA (%.1f accumulators/clk/CTA)\n", mode, grid, avg / iters,
                   32768.0 * iters / avg);
        }
    }
    return bad ? 2 : 0;
}
