// gemm_tc.cu — the dense contractions of the forward on the 5th-gen tensor cores (tcgen05).
//
//   Cout[M,N] = act(A[M,K] · W[N,K]^T + bias) (+ Res)        fp32 in, fp32 out, fp32 accumulate
//
// Precision: the north_star bound is 1e-3 absolute on fp32 logits, which a single bf16 pass
// (2^-9 operand rounding over ~9 chained contractions) does not meet.  Each fp32 operand is split
// into two bf16 terms x = hi + lo (lo = bf16(x - hi)); three tcgen05.mma passes accumulate
// hi·hi + hi·lo + lo·hi in the fp32 TMEM accumulator (the dropped lo·lo term is 2^-18 relative).
// Weights are split once at load time; activations are split on the fly while being staged.
//
// Tile: M=128 rows (4 positions x 32 read tokens) x BN columns, K in blocks of 64 (one 128-byte
// swizzle atom of bf16).  Shared-memory operands use the canonical K-major SWIZZLE_128B layout
// (8-row x 128-byte atoms, SBO = 1024 B; 16-byte chunk c of row r lives at chunk c ^ (r & 7)).
// One thread issues the MMAs; tcgen05.commit signals an mbarrier; the four warps read their 32
// TMEM lanes (= one position each) with tcgen05.ld for the fused bias/ReLU/residual epilogue.
// Several CTAs are resident per SM (64 KB smem, BN TMEM columns each), so one CTA's staging
// overlaps another's MMAs without an intra-CTA pipeline.
#include <cuda_bf16.h>

#include "common.cuh"
#include "forward.h"

namespace hb {

namespace {

constexpr int BM = 128, BK = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);        // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)0 << 16;                          // LBO: unused for swizzled K-major (single atom along K)
    d |= (uint64_t)(1024u >> 4) << 32;               // SBO = 8 rows * 128 B, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
    return d;
}

// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace

template <int BN, int ACT, int RES>
__global__ void __launch_bounds__(128) k_gemm_tc(const float* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ Whi,
                                                 const __nv_bfloat16* __restrict__ Wlo, const float* __restrict__ bias,
                                                 float* Cout, int ldc, const float* Res, int K) {
    extern __shared__ uint8_t smem_dyn[];
    // SWIZZLE_128B operands need 1024-byte alignment; the dynamic segment starts after the static one
    uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t* sAhi = smem;                       // 128 rows x 128 B
    uint8_t* sAlo = sAhi + BM * 128;
    uint8_t* sBhi = sAlo + BM * 128;            // BN rows x 128 B
    uint8_t* sBlo = sBhi + BN * 128;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    constexpr uint32_t idesc = make_idesc(BN);

    uint32_t parity = 0;
    const int nkb = K / BK;
    for (int kb = 0; kb < nkb; kb++) {
        const int k0 = kb * BK;
        // ---- stage A: 128 rows x 8 chunks of 8 floats -> bf16 hi / lo, swizzled
#pragma unroll
        for (int it = 0; it < (BM * 8) / 128; it++) {
            const int item = it * 128 + tid;
            const int r = item >> 3, c = item & 7;
            const float4 v0 = *(const float4*)(A + (size_t)(m0 + r) * lda + k0 + c * 8);
            const float4 v1 = *(const float4*)(A + (size_t)(m0 + r) * lda + k0 + c * 8 + 4);
            const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; e++) split_bf16(x[e], h[e], l[e]);
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
            *(uint4*)(sAhi + off) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
            *(uint4*)(sAlo + off) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
        }
        // ---- stage W (already split): BN rows x 8 chunks of 16 B
#pragma unroll
        for (int it = 0; it < (BN * 8) / 128; it++) {
            const int item = it * 128 + tid;
            const int r = item >> 3, c = item & 7;
            const size_t g = (size_t)(n0 + r) * K + k0 + c * 8;
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
            *(uint4*)(sBhi + off) = *(const uint4*)(Whi + g);
            *(uint4*)(sBlo + off) = *(const uint4*)(Wlo + g);
        }
        // generic-proxy writes -> visible to the tensor core (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dAh = make_desc(smem_u32(sAhi)), dAl = make_desc(smem_u32(sAlo));
            const uint64_t dBh = make_desc(smem_u32(sBhi)), dBl = make_desc(smem_u32(sBlo));
#pragma unroll
            for (int k = 0; k < BK / 16; k++) {
                const uint64_t adv = (uint64_t)((k * 32) >> 4);  // +32 bytes along K inside the swizzle atom
                mma_bf16(tmem_d, dAh + adv, dBh + adv, idesc, (kb | k) ? 1u : 0u);
                mma_bf16(tmem_d, dAh + adv, dBl + adv, idesc, 1u);
                mma_bf16(tmem_d, dAl + adv, dBh + adv, idesc, 1u);
            }
            // arrives on the mbarrier once all MMAs issued so far have completed (implies the before_thread_sync fence)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // the operands may only be overwritten (and the accumulator read) after the MMAs finished
        mbar_wait(&mbar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }

    // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 (= rows m0+32w.. = one position)
    const int row = m0 + warp * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* out = Cout + (size_t)row * ldc + n0 + c0;
        const float* res = RES ? Res + (size_t)row * ldc + n0 + c0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bv = *(const float4*)(bias + n0 + c0 + j);
            float4 o = make_float4(__uint_as_float(v[j]) + bv.x, __uint_as_float(v[j + 1]) + bv.y,
                                   __uint_as_float(v[j + 2]) + bv.z, __uint_as_float(v[j + 3]) + bv.w);
            if (ACT) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (RES) { const float4 r = *(const float4*)(res + j); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
            *(float4*)(out + j) = o;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN));
    }
}

// split fp32 weights into bf16 hi / lo (run once at model load)
__global__ void k_split_bf16(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_bf16(w[i], hi[i], lo[i]);
}

cudaError_t split_weights(const float* w, size_t n, void** hi, void** lo) {
    cudaError_t e = cudaMalloc(hi, n * 2);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(lo, n * 2);
    if (e != cudaSuccess) return e;
    k_split_bf16<<<(unsigned)((n + 255) / 256), 256>>>(w, (__nv_bfloat16*)*hi, (__nv_bfloat16*)*lo, n);
    return cudaGetLastError();
}

template <int BN, int ACT, int RES>
static cudaError_t launch_one(const float* A, int lda, const void* Whi, const void* Wlo, const float* bias, float* Cout, int ldc,
                              const float* Res, size_t M, int N, int K, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = (size_t)(2 * BM + 2 * BN) * 128 + 1024;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, ACT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((unsigned)(M / BM), (unsigned)(N / BN));
    k_gemm_tc<BN, ACT, RES><<<grid, 128, smem, st>>>(A, lda, (const __nv_bfloat16*)Whi, (const __nv_bfloat16*)Wlo, bias, Cout, ldc, Res, K);
    return cudaGetLastError();
}

// M % 128 == 0, K % 64 == 0, N % 64 == 0
cudaError_t gemm_tc(int act, int res, const float* A, int lda, const void* Whi, const void* Wlo, const float* bias, float* Cout,
                    int ldc, const float* Res, size_t M, int N, int K, cudaStream_t st) {
    const bool wide = (N % 128) == 0;
#define HB_GEMM_CASE(a, r)                                                                                         \
    if (act == a && res == r)                                                                                      \
        return wide ? launch_one<128, a, r>(A, lda, Whi, Wlo, bias, Cout, ldc, Res, M, N, K, st)                    \
                    : launch_one<64, a, r>(A, lda, Whi, Wlo, bias, Cout, ldc, Res, M, N, K, st);
    HB_GEMM_CASE(0, 0)
    HB_GEMM_CASE(1, 0)
    HB_GEMM_CASE(0, 1)
#undef HB_GEMM_CASE
    return cudaErrorInvalidValue;
}

}  // namespace hb
