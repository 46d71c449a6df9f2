// forward.h — device-resident weights of the forward stage (see herro_b200/weights.py for the
// blob format and tensor names).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace hb {

constexpr int MAX_LAYERS = 8;

struct FwdLayer {
    const float *ln1_g, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
};

struct FwdWeights {
    int stem_k, C, H, layers, F, D;
    const float* stem_tab;  // [K][12][C]  = sum_e stem_w[c][e][j] * emb[t][e]   (embedding folded into the conv)
    const float* stem_wq;   // [K][C]      = stem_w[c][6][j]                     (quality channel)
    const float* stem_b;    // [C]
    const float* read_pos;  // [31][C]
    FwdLayer layer[MAX_LAYERS];
    const float *lnf_g, *lnf_b, *wc, *bc, *wb, *bb, *wi, *bi;
};

size_t fwd_workspace_floats(const FwdWeights& wt, uint32_t chunk_pos);
int launch_forward_chunk(const BatchView& b, const FwdWeights& wt, uint32_t n0, uint32_t npos, float* ws,
                         float* logits, float* info, cudaStream_t st);

// features.cu
cudaError_t features_configure(uint32_t W);
int launch_features_a(const BatchView& b, cudaStream_t st);
int launch_pileup(const BatchView& b, cudaStream_t st);
int launch_features_c1(const BatchView& b, cudaStream_t st);
int launch_features_c2(const BatchView& b, cudaStream_t st);
int launch_consensus(const BatchView& b, cudaStream_t st);

}  // namespace hb
