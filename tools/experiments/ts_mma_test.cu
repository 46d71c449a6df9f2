// Experiment (not part of the product): does tcgen05.mma take its A operand from TMEM in the layout
//   lane = row, 32-bit column j = K elements (2j | 2j+1 << 16) ?
// D[128 x 128] = A[128 x 64] * B[128 x 64]^T, A written to TMEM with tcgen05.st by the row-owner threads, B in shared memory
// (SWIZZLE_128B, K-major).  Prints the max abs error against a host reference for both half-word orders.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(128, 1) k_test(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int swap_halves) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* sB = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B: row n (128 rows), 64 K elements = 128 B per row, 16-byte chunk c stored at chunk c ^ (n & 7)
    for (int i = tid; i < 128 * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        *(uint4*)(sB + n * 128 + ((c ^ (n & 7)) << 4)) = *(const uint4*)(B + n * 64 + c * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // A: thread tid owns row tid: 64 K elements = 32 packed words -> TMEM columns 128..159 of lane tid
    uint32_t w[32];
    for (int j = 0; j < 32; j++) {
        const uint32_t lo = __bfloat16_as_ushort(A[tid * 64 + 2 * j]), hi = __bfloat16_as_ushort(A[tid * 64 + 2 * j + 1]);
        w[j] = swap_halves ? (hi | (lo << 16)) : (lo | (hi << 16));
    }
    const uint32_t a_addr = tmem_base + 128 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(a_addr),
        "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
        "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]),
        "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]),
        "r"(w[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t dB = make_desc(smem_u32(sB));
        for (int k = 0; k < 4; k++) {
            const uint32_t a_k = tmem_base + 128 + k * 8;   // 16 K elements = 8 columns
            const uint64_t b_k = dB + (uint64_t)((k * 32) >> 4);
            const uint32_t acc = k ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base),
                "r"(a_k), "l"(b_k), "r"(IDESC), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem_base + c0 + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[tid * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

// ---- throughput: cycles per tcgen05.mma (M=128, K=16) issued back to back by one thread, operands resident
template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int iters) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (16384 + 32768) / 16; i += 128) *(uint4*)(sm + i * 16) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        const uint64_t dA = make_desc(smem_u32(sm)), dB = make_desc(smem_u32(sm) + 16384);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const int k = i & 3;
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            if (A_TMEM) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base),
                             "r"(tmem_base + 256 + k * 8), "l"(dB + adv), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
                             "l"(dA + adv), "l"(dB + adv), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}
template <int N, bool A_TMEM>
static void run_rate(long long* d_out, int grid) {
    const int iters = 4000;
    cudaFuncSetAttribute(k_rate<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_rate<N, A_TMEM><<<grid, 128, 65536>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("RATE N=%d A=%s grid=%d: %s, %.1f cycles per MMA (floor %d)\n", N, A_TMEM ? "tmem" : "smem", grid, cudaGetErrorString(e), (double)c / iters, 128 * N / 256);
}

int main() {
    {
        long long* d_out; cudaMalloc(&d_out, 8);
        for (int grid : {1, 148}) {
            run_rate<64, false>(d_out, grid); run_rate<64, true>(d_out, grid);
            run_rate<128, false>(d_out, grid); run_rate<128, true>(d_out, grid);
            run_rate<256, false>(d_out, grid); run_rate<256, true>(d_out, grid);
        }
    }
    std::vector<__nv_bfloat16> A(128 * 64), B(128 * 64);
    std::vector<float> Af(128 * 64), Bf(128 * 64), ref(128 * 128), D(128 * 128);
    srand(1);
    for (int i = 0; i < 128 * 64; i++) {
        A[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Af[i] = __bfloat162float(A[i]);
        B[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Bf[i] = __bfloat162float(B[i]);
    }
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 128; n++) {
            double s = 0;
            for (int k = 0; k < 64; k++) s += (double)Af[m * 64 + k] * Bf[n * 64 + k];
            ref[m * 128 + n] = (float)s;
        }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int sw = 0; sw < 2; sw++) {
        cudaMemset(dD, 0, D.size() * 4);
        k_test<<<1, 128, 32768>>>(dA, dB, dD, sw);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < 128 * 128; i++) worst = fmax(worst, fabs((double)D[i] - ref[i]));
        printf("TSMMA swap_halves=%d: %s, max abs err %.3g (D[0]=%g ref[0]=%g, D[129]=%g ref[129]=%g)\n", sw, cudaGetErrorString(e), worst, D[0], ref[0], D[129], ref[129]);
    }
    return 0;
}
