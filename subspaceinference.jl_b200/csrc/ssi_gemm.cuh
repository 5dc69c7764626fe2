// ssi_gemm.cuh -- the strided FP32 SIMT GEMM shared by the gradient (ssi_grad.cu) and training (ssi_train.cu) paths.
#pragma once
#include "ssi_common.cuh"

struct gemm_t {
    // C(o, j) = sum_k A(o, k) B(k, j) for batch b = blockIdx.z; all strides in elements
    const float* A; long long a_so, a_sk, a_sb;
    const float* B; long long b_sk, b_sj, b_sb;
    float* C;       long long c_so, c_sj, c_sb;
    int O, J;
    long long K;            // contraction length (per batch when split > 0: the last batch may be shorter)
    long long split;        // > 0: batch b covers k in [b*split, min(K, (b+1)*split))
    int epi;                // 0 plain, 1 bias + activation, 2 multiply by act'(Hprev(o, j))
    int act;
    const float* bias; long long bias_sb;
    const float* Hprev; long long h_sb;       // indexed like C
};

// a_kfast: A is contiguous along k (else along o); b_jfast: B is contiguous along j (else along k)
int ssi_launch_gemm(ssi_ctx* ctx, const gemm_t& g, int batches, bool a_kfast, bool b_jfast);
// delta_L(o, j) = -(pred - y) * coef * act_L'(pred) and per-(sample, chunk of 256 datapoints) squared-error partials
int ssi_launch_delta_out(ssi_ctx* ctx, const float* pred, long long pred_sb, const float* Y, float* delta, long long delta_sb,
                         long long N, int O, int act, float coef, int n_chunks, int g, double* partials);
// gb(o) = sum_j delta(o, j), fixed order
int ssi_launch_rowsum(ssi_ctx* ctx, const float* delta, long long d_sb, int O, long long N, int g, float* out, long long out_sb);
