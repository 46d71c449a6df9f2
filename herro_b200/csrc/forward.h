// forward.h — device-resident weights of the forward stage (see herro_b200/weights.py for the
// blob format and tensor names).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "common.cuh"

namespace hb {

constexpr int MAX_LAYERS = 8;

struct SplitW {  // bf16 hi / lo split of an fp32 weight matrix [N,K] (gemm_tc.cu)
    const void *hi = nullptr, *lo = nullptr;
};

struct FwdLayer {
    const float *ln1_g, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
    SplitW s_qkv, s_o, s_1, s_2;
    SplitW s_qkvp;               // Wqkv with rows regrouped per head [H][q(32) | k(32) | v(32)][C] (fused QKV+attention kernel)
    const float* bqkvp = nullptr;  // bqkv in the same order; nullptr: fused kernel unavailable
};

struct FwdWeights {
    int stem_k, C, H, layers, F, D;
    const float* stem_tab;  // [K][12][C]  = sum_e stem_w[c][e][j] * emb[t][e]   (embedding folded into the conv)
    const float* stem_wq;   // [K][C]      = stem_w[c][6][j]                     (quality channel)
    const float* stem_b;    // [C]
    const float* read_pos;  // [31][C]
    FwdLayer layer[MAX_LAYERS];
    const float *lnf_g, *lnf_b, *wc, *bc, *wb, *bb, *wi, *bi;
    SplitW s_c;
    SplitW s_stem;          // W' of the tensor-core stem, [C][stem_kp]
    int stem_kblocks = 0;   // 0: tensor-core stem unavailable (C != 128 or too many taps)
    int num_sms;
    // debugging aids / A-B parity tests, read from the environment ONCE in hb_create (HERRO_B200_NO_FUSE_{LN,FFN,ATTN})
    int no_fuse_ln = 0, no_fuse_ffn = 0, no_fuse_attn = 0, no_fuse_oproj = 0;
};

// gemm_tc.cu
enum { GEMM_OUT_F32 = 0, GEMM_OUT_F32_RELU = 1, GEMM_OUT_F32_RES = 2, GEMM_OUT_SPLIT_RELU = 3,
       GEMM_OUT_F32_RES_LN = 4 /* N == 128: out = acc+bias+res (fp32) and LayerNorm(out) as split bf16 */ };
struct GemmArgs {
    const __nv_bfloat16 *Ahi, *Alo;  // activations, split bf16, row stride lda (elements)
    size_t lda;
    const __nv_bfloat16 *Whi, *Wlo;  // weights [N,K], split bf16
    uint32_t K;
    const float* bias;               // [N]
    float* out;                      // fp32 output (modes F32*), row stride ldc
    const float* res;                // residual (mode F32_RES), same layout as out
    size_t ldc;
    __nv_bfloat16 *out_hi, *out_lo;  // split bf16 output (mode SPLIT_RELU), row stride ldo
    size_t ldo;
    const float *ln_g, *ln_b;        // LayerNorm affine (mode F32_RES_LN)
    uint32_t m_tiles, n_chunks, k_blocks;  // M/128, N/128, K/64
    int mode;
};
struct FfnArgs {  // k_ffn_ws: fused FFN for C == 128, F == 512
    const __nv_bfloat16 *Hhi, *Hlo;        // LayerNorm(X), split bf16, [T][128]
    const __nv_bfloat16 *W1hi, *W1lo;      // [F][128]
    const __nv_bfloat16 *W2hi, *W2lo;      // [128][F]
    const float *b1, *b2;                  // [F], [128]
    float* X;                              // residual stream [T][128], updated in place
    const float *ln_g, *ln_b;              // LayerNorm that follows
    __nv_bfloat16 *out_hi, *out_lo;        // LayerNorm(X_new), split bf16, [T][128]
    uint32_t F, m_tiles;
    // fused attention out-projection (k_ffn_ws<true>): Hhi/Hlo then hold the attention output O, and the kernel first computes
    // X += O · Wo^T + bo, H = LayerNorm(X; ln2) on chip.  Wohi == nullptr: unfused (H is read as given).
    const __nv_bfloat16 *Wohi = nullptr, *Wolo = nullptr;  // [128][128]
    const float *bo = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    int store_x = 1;  // 0: do not write the updated residual stream back (last layer: only LayerNorm(X) is consumed)
    // 1: X is tile-blocked, [tile of 128 rows][32 chunks of 4 columns][128 rows][4 floats], so that the row-owner threads of the
    // epilogues (one TMEM lane = one row each) read and write it with coalesced 16-byte accesses (a warp = 32 consecutive rows
    // of one chunk = 512 contiguous bytes) and no shared-memory transposes.  Only in the fully fused graph, where k_stem_tc and
    // k_ffn_ws<true> are the only kernels that touch X (forward.cu).
    int x_blocked = 0;
};
cudaError_t ffn_tc(const FfnArgs& a, int num_sms, cudaStream_t st);
struct QkvAttnArgs {  // k_qkv_attn_ws: QKV projection + read-axis attention for C == 128, 4 heads
    const __nv_bfloat16 *Hhi, *Hlo;      // LayerNorm(X), split bf16, [T][128]
    const __nv_bfloat16 *Whi, *Wlo;      // head-grouped Wqkv [4*96][128]
    const float* bias;                   // head-grouped bqkv [4*96]
    __nv_bfloat16 *out_hi, *out_lo;      // attention output, split bf16, [T][128] (may alias Hhi/Hlo: tile-local in-place)
    uint32_t m_tiles;
};
cudaError_t qkv_attn_tc(const QkvAttnArgs& a, int num_sms, cudaStream_t st);
struct StemArgs {  // k_stem_tc: the stem as a contraction over taps x 16 features (C == 128 only)
    const __nv_bfloat16 *Whi, *Wlo;  // W' [128][Kp], split bf16
    uint32_t Kp;                     // k_blocks * 64
    uint32_t k_blocks, taps;
    const float *bias, *read_pos;    // [128], [31][128]
    float* X;                        // [positions*32][128]
    uint32_t n0, npos;               // work-list range
    const float *ln_g = nullptr, *ln_b = nullptr;          // optional: LayerNorm(X) of every row, emitted as split bf16
    __nv_bfloat16 *out_hi = nullptr, *out_lo = nullptr;    // [positions*32][128]; nullptr: X only
    int x_blocked = 0;                                     // X in the tile-blocked layout (FfnArgs::x_blocked)
};
cudaError_t stem_tc(const BatchView& b, const StemArgs& a, int num_sms, cudaStream_t st);
cudaError_t split_weights(const float* w, size_t n, void** hi, void** lo);
cudaError_t gemm_tc(const GemmArgs& a, int num_sms, cudaStream_t st);
// forward.cu (fp32 SIMT contraction: the self test's reference)
void gemm_simt(int act, int res, const float* A, int lda, const float* Wt, const float* bias, float* Cout, int ldc,
               const float* Res, size_t M, int N, int K, cudaStream_t st);

// Per-kernel-class CUDA-event timing on the launching stream (off during replays).
struct KTimer {
    bool on = false;
    cudaStream_t st = nullptr;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int cls; size_t e0, e1; };
    std::vector<Rec> recs;
    uint64_t launches[16] = {0};
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void begin(int cls) {
        launches[cls]++;
        if (!on) return;
        recs.push_back(Rec{cls, used, used + 1});
        cudaEventRecord(get(), st);
        get();
    }
    void end() {
        if (!on) return;
        cudaEventRecord(pool[recs.back().e1], st);
    }
    // after the stream is synchronised: add elapsed ms per class, reset
    void collect(double* ms, uint64_t* n) {
        for (auto& r : recs) { float t = 0; cudaEventElapsedTime(&t, pool[r.e0], pool[r.e1]); ms[r.cls] += t; }
        for (int i = 0; i < 16; i++) { n[i] += launches[i]; launches[i] = 0; }
        recs.clear();
        used = 0;
    }
    void discard() { for (auto& l : launches) l = 0; recs.clear(); used = 0; }
    void destroy() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};
enum { K_TOKENIZE = 0, K_PASS1, K_SCORES, K_PASS2A, K_SCAN, K_PILEUP, K_LISTS, K_STEM, K_LAYERNORM, K_GEMM, K_ATTENTION,
       K_HEADS, K_CONSENSUS, K_FFN, K_QKV_ATTN };

size_t fwd_workspace_bytes(const FwdWeights& wt, uint32_t chunk_pos);
int launch_forward_chunk(const BatchView& b, const FwdWeights& wt, uint32_t n0, uint32_t npos, uint8_t* ws,
                         float* logits, float* info, cudaStream_t st, KTimer& kt);
uint64_t forward_flops_per_pos(const FwdWeights& wt, uint64_t* gemm_flops);
// algorithmic FLOPs per supported position attributed to the kernel class that executes them on the active code path
void forward_class_flops_per_pos(const FwdWeights& wt, uint64_t (&out)[16]);

// features.cu
cudaError_t features_configure(uint32_t W);
int launch_features_a(const BatchView& b, cudaStream_t st, KTimer& kt);
int launch_pileup(const BatchView& b, cudaStream_t st, KTimer& kt, bool v1);
// windowing_dev.cu: extract_windows on the device (parse, boundaries, op-slot scan); returns the number of kernels launched
int launch_windowing(const BatchView& b, cudaStream_t st);
// k_scan_u32 of features.cu
void launch_scan_u32(const uint32_t* in, uint64_t* out, uint32_t n, uint32_t* counters, int total_slot, uint64_t cap, int overflow_slot,
                     cudaStream_t st);
// pileup.cu
cudaError_t pileup_configure();
void launch_pileup_v2(const BatchView& b, cudaStream_t st);
int launch_features_c1(const BatchView& b, cudaStream_t st, KTimer& kt);
int launch_features_c2(const BatchView& b, cudaStream_t st, KTimer& kt);
int launch_consensus(const BatchView& b, cudaStream_t st, KTimer& kt);

}  // namespace hb
