// ssi_basis.cu — BASIS path of the batched log-posterior: Chain(Dense(in, H, act), Dense(H, O)) with a narrow
// output (O <= 2) and a small subspace (M <= 20), e.g. the UCI-style 13-50-1 network with 4096 chains.
//
// Replaces the body of the reference's density(z) (src/space_inference.jl:90-95) for that shape:
//   * first Dense layer: the pre-activation is affine in z,
//         (W1_swa + sum_m z_m P1_m) x + (b1_swa + sum_m z_m pb_m) = B_M(x) + sum_m z_m B_m(x),
//     so the M+1 basis pre-activations B_m = [P | W_swa]_m applied to the dataset are computed ONCE per
//     (data, subspace) in exact FP32 (ssi_build_first_layer_bases) and a sample costs M FMAs per hidden unit
//     instead of `in` (K1 and the first GEMM of K2 fused away);
//   * second Dense layer + logpdf(MvNormal): a dot product and a squared error in registers, reduced with
//     warp shuffles to one FP64 partial per (sample, datapoint tile)  (K3).
//
// Mapping.  A CTA owns one tile of 64 datapoints (lane l holds datapoints 2l, 2l+1) and a block of samples;
// each warp walks its share of the samples ST at a time.  The z of a warp's current samples are warp-uniform:
// they are loaded once per pass into registers as pairs of adjacent samples, so the inner products are packed FFMA2 with the
// basis value as the scalar operand, and the basis tile is the only per-lane operand.
// The basis tile lives in shared memory as [j][m][64] (conflict-free 8-byte reads); it is laid out per tile
// in global memory once, so staging it is a contiguous copy.  Bound: FP32 FMA issue.
#include "ssi_common.cuh"

#include <algorithm>

#define B1_TI 64
#define B1_WARPS 8
#define B1_THREADS (B1_WARPS * 32)

struct b1_params {
    int N, H, O, n_tiles, S, s_per_cta, act_out, w2n;
    long long n, w2_off, b2_off;
    const float* tiles;     // [tile][H][M+1][B1_TI]
    const float* PW;        // [P | W_swa], n x (M+1) column-major
    const float* Z;         // [S][M] (device)
    const float* W2;        // [S/ST][H+1][ST][OT]: second-layer weights (j < H) and bias (j = H) of every sample, k_b1_project_w2
    const float* Y;         // O x N column-major
    double* partials;       // [S][n_tiles]
};

template <int ACT>
__device__ __forceinline__ float b1_act(float v, int act_rt) {
    if (ACT == SSI_ACT_RELU) return fmaxf(v, 0.0f);
    if (ACT == SSI_ACT_TANH) return tanhf(v);
    if (ACT == SSI_ACT_IDENTITY) return v;
    return ssi_act(v, act_rt);
}

template <int M, int ST, int OT, int ACT>
__global__ void __launch_bounds__(B1_THREADS, 2)
k_logpost_basis1h(const b1_params p, const int act_hidden) {
    extern __shared__ float4 b1_smem4[];
    const int H = p.H;
    float* sB = reinterpret_cast<float*>(b1_smem4);           // [H][M+1][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int s_begin = blockIdx.y * p.s_per_cta;
    const int s_end = min(p.S, s_begin + p.s_per_cta);

    {   // stage the basis tile (contiguous)
        const float4* src = reinterpret_cast<const float4*>(p.tiles + (size_t)tile * H * (M + 1) * B1_TI);
        float4* dst = reinterpret_cast<float4*>(sB);
        const int n4 = H * (M + 1) * B1_TI / 4;
        for (int e = tid; e < n4; e += B1_THREADS) dst[e] = __ldg(src + e);
    }
    __syncthreads();

    // this lane's datapoints and targets
    const long long i0 = (long long)tile * B1_TI + 2 * lane;
    float y[2][OT];
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int o = 0; o < OT; ++o) y[d][o] = (i0 + d < p.N && o < p.O) ? p.Y[o + (i0 + d) * p.O] : 0.0f;

    const float* bl = sB + 2 * lane;

    for (int s0 = s_begin + warp * ST; s0 < s_end; s0 += B1_WARPS * ST) {
        // second-layer weights and bias of this warp's ST samples, (W_swa + P z)[second layer] (src/space_inference.jl:91),
        // projected once per call by k_b1_project_w2: warp-uniform 16-byte loads that hit L1
        const float* w2s = p.W2 + (size_t)(s0 / ST) * (H + 1) * ST * OT;

        // warp-uniform z of the ST samples (clamped: a partial block recomputes the last sample and drops it),
        // kept as pairs of adjacent samples: the inner products are packed FFMA2 (two samples per instruction; a
        // three-register scalar FFMA issues at half rate on sm_100, the packed form does not lose that factor)
        float2 z2[ST / 2][M];
#pragma unroll
        for (int t = 0; t < ST; t += 2) {
            const int sa = min(s0 + t, p.S - 1), sb = min(s0 + t + 1, p.S - 1);
#pragma unroll
            for (int m = 0; m < M; ++m) z2[t / 2][m] = make_float2(__ldg(p.Z + (long long)sa * M + m), __ldg(p.Z + (long long)sb * M + m));
        }

        float2 pred[ST / 2][2][OT];      // [sample pair][datapoint][output] = (sample t, sample t+1)
#pragma unroll
        for (int t = 0; t < ST / 2; ++t)
#pragma unroll
            for (int o = 0; o < OT; ++o) pred[t][0][o] = pred[t][1][o] = make_float2(0.0f, 0.0f);

#pragma unroll 2
        for (int j = 0; j < H; ++j) {
            float2 b[M + 1];
#pragma unroll
            for (int m = 0; m <= M; ++m) b[m] = *reinterpret_cast<const float2*>(bl + (j * (M + 1) + m) * B1_TI);
            float w[ST * OT];       // [sample][output]
            if (ST * OT % 4 == 0) {
#pragma unroll
                for (int q = 0; q < ST * OT; q += 4) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w2s + j * ST * OT + q));
                    w[q] = w4.x; w[q + 1] = w4.y; w[q + 2] = w4.z; w[q + 3] = w4.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < ST * OT; q += 2) {
                    const float2 w2v = __ldg(reinterpret_cast<const float2*>(w2s + j * ST * OT + q));
                    w[q] = w2v.x; w[q + 1] = w2v.y;
                }
            }
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                float2 bd[M + 1];           // this datapoint's basis values, duplicated for the sample pair
#pragma unroll
                for (int m = 0; m <= M; ++m) { const float v = d ? b[m].y : b[m].x; bd[m] = make_float2(v, v); }
#pragma unroll
                for (int t = 0; t < ST / 2; ++t) {
                    float2 x = bd[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) x = __ffma2_rn(bd[m], z2[t][m], x);
                    x.x = b1_act<ACT>(x.x, act_hidden);
                    x.y = b1_act<ACT>(x.y, act_hidden);
#pragma unroll
                    for (int o = 0; o < OT; ++o)
                        pred[t][d][o] = __ffma2_rn(x, make_float2(w[(2 * t) * OT + o], w[(2 * t + 1) * OT + o]), pred[t][d][o]);
                }
            }
        }

        // output bias + activation, squared error, one FP64 partial per (sample, tile)   (:94)
#pragma unroll
        for (int t = 0; t < ST; ++t) {
            double sse = 0.0;
#pragma unroll
            for (int o = 0; o < OT; ++o) {
                const float b2 = __ldg(w2s + (H * ST + t) * OT + o);
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    if (o < p.O && i0 + d < p.N) {
                        const float pv = (t & 1) ? pred[t / 2][d][o].y : pred[t / 2][d][o].x;
                        const float df = ssi_act(pv + b2, p.act_out) - y[d][o];
                        sse += (double)df * (double)df;
                    }
                }
            }
            sse = ssi_warp_sum(sse);
            if (lane == 0 && s0 + t < s_end) p.partials[(long long)(s0 + t) * p.n_tiles + tile] = sse;
        }
    }
}

// ---------------------------------------------------------------------------------------
// value + gradient for the same shape with a scalar output (O = 1): forward-mode tangents in registers.
//   d pred / d z_m = pb2_m + sum_j [ P2_m[j] h_j + w2_j act'(h_j) B_m(x)[j] ]
// is accumulated next to pred in the same pass over the hidden units (2M more packed FMAs per unit and hidden unit), then
//   d lp / d z_m = sum_i e_i * d pred_i / d z_m,   e_i = -(pred_i - y_i)/sigma_m^2 * act_out'(pred_i),
// is reduced over the tile's datapoints with warp shuffles: one float partial per (tile, sample, m), summed in fixed
// order by k_grad_finish.  This is what l_pi_grad (src/space_inference.jl:107) costs on this path: ~2.5x a density
// evaluation instead of ForwardDiff's M+1 forward passes.  Same mapping as k_logpost_basis1h with ST = 4 samples per warp
// pass (the tangents need M x 4 more accumulators).
#define B1G_ST 4
template <int M, int ACT>
__global__ void __launch_bounds__(B1_THREADS, 2)
k_logpost_basis1h_grad(const b1_params p, const int act_hidden, const float coef, float* __restrict__ gpart, const long long slab) {
    constexpr int ST = B1G_ST;
    extern __shared__ float4 b1_smem4[];
    const int H = p.H;
    float* sB = reinterpret_cast<float*>(b1_smem4);           // [H][M+1][64]
    float* sP2 = sB + (size_t)H * (M + 1) * B1_TI;            // [M+1][w2n]: second-layer rows of [P | W_swa] (j), bias at j = H
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int s_begin = blockIdx.y * p.s_per_cta;
    const int s_end = min(p.S, s_begin + p.s_per_cta);
    {
        const float4* src = reinterpret_cast<const float4*>(p.tiles + (size_t)tile * H * (M + 1) * B1_TI);
        float4* dst = reinterpret_cast<float4*>(sB);
        const int n4 = H * (M + 1) * B1_TI / 4;
        for (int e = tid; e < n4; e += B1_THREADS) dst[e] = __ldg(src + e);
        for (int e = tid; e < (M + 1) * p.w2n; e += B1_THREADS) {
            const int m = e / p.w2n, j = e - m * p.w2n;
            float v = 0.0f;
            if (j <= H) v = p.PW[(j < H ? p.w2_off + j : p.b2_off) + (long long)m * p.n];
            sP2[e] = v;
        }
    }
    __syncthreads();
    const long long i0 = (long long)tile * B1_TI + 2 * lane;
    const float y0 = i0 < p.N ? p.Y[i0] : 0.0f, y1 = i0 + 1 < p.N ? p.Y[i0 + 1] : 0.0f;
    const float* bl = sB + 2 * lane;

    for (int s0 = s_begin + warp * ST; s0 < s_end; s0 += B1_WARPS * ST) {
        const float* w2s = p.W2 + (size_t)(s0 / ST) * (H + 1) * ST;
        float2 z2[ST / 2][M];
#pragma unroll
        for (int t = 0; t < ST; t += 2) {
            const int sa = min(s0 + t, p.S - 1), sb = min(s0 + t + 1, p.S - 1);
#pragma unroll
            for (int m = 0; m < M; ++m) z2[t / 2][m] = make_float2(__ldg(p.Z + (long long)sa * M + m), __ldg(p.Z + (long long)sb * M + m));
        }
        float2 pred[ST / 2][2];          // [sample pair][datapoint]
        float2 dp[ST / 2][2][M];         // d pred / d z_m (without the bias row, added at the end)
#pragma unroll
        for (int t = 0; t < ST / 2; ++t)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                pred[t][d] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int m = 0; m < M; ++m) dp[t][d][m] = make_float2(0.0f, 0.0f);
            }

#pragma unroll 1
        for (int j = 0; j < H; ++j) {
            float2 b[M + 1];
#pragma unroll
            for (int m = 0; m <= M; ++m) b[m] = *reinterpret_cast<const float2*>(bl + (j * (M + 1) + m) * B1_TI);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w2s + j * ST));
            const float2 wp[2] = {make_float2(w4.x, w4.y), make_float2(w4.z, w4.w)};
            float p2[M];
#pragma unroll
            for (int m = 0; m < M; ++m) p2[m] = sP2[m * p.w2n + j];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
#pragma unroll
                for (int t = 0; t < ST / 2; ++t) {
                    const float bM = d ? b[M].y : b[M].x;
                    float2 x = make_float2(bM, bM);
#pragma unroll
                    for (int m = 0; m < M; ++m) { const float bv = d ? b[m].y : b[m].x; x = __ffma2_rn(make_float2(bv, bv), z2[t][m], x); }
                    float2 h, c;
                    h.x = b1_act<ACT>(x.x, act_hidden);
                    h.y = b1_act<ACT>(x.y, act_hidden);
                    const int a = ACT < 0 ? act_hidden : ACT;
                    c.x = wp[t].x * act_deriv_from_output(h.x, a);       // w2_j act'(h_j)
                    c.y = wp[t].y * act_deriv_from_output(h.y, a);
                    pred[t][d] = __ffma2_rn(h, wp[t], pred[t][d]);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float bv = d ? b[m].y : b[m].x;
                        dp[t][d][m] = __ffma2_rn(h, make_float2(p2[m], p2[m]), dp[t][d][m]);
                        dp[t][d][m] = __ffma2_rn(c, make_float2(bv, bv), dp[t][d][m]);
                    }
                }
            }
        }
        // residuals, squared error, e_i, and the reduction of e_i * d pred_i / d z_m over the tile's datapoints
#pragma unroll
        for (int t = 0; t < ST; ++t) {
            const float b2 = __ldg(w2s + H * ST + t);
            double sse = 0.0;
            float g[M];
#pragma unroll
            for (int m = 0; m < M; ++m) g[m] = 0.0f;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                if (i0 + d < p.N) {
                    const float pv = ssi_act(((t & 1) ? pred[t / 2][d].y : pred[t / 2][d].x) + b2, p.act_out);
                    const float df = pv - (d ? y1 : y0);
                    sse += (double)df * (double)df;
                    const float e = -df * coef * act_deriv_from_output(pv, p.act_out);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float dpm = ((t & 1) ? dp[t / 2][d][m].y : dp[t / 2][d][m].x) + sP2[m * p.w2n + H];   // + pb2_m
                        g[m] = fmaf(e, dpm, g[m]);
                    }
                }
            }
            sse = ssi_warp_sum(sse);
#pragma unroll
            for (int m = 0; m < M; ++m) g[m] = ssi_warp_sum(g[m]);
            if (lane == 0 && s0 + t < s_end) {
                p.partials[(long long)(s0 + t) * p.n_tiles + tile] = sse;
                float* dst = gpart + (long long)tile * slab + (long long)(s0 + t) * M;
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m] = g[m];
            }
        }
    }
}

// W2[s / ST][j][s % ST][o] = (W_swa + P z_s)[second layer]: weight (o, j) for j < H, bias o for j = H; zero for o >= O and for the
// samples that pad the last block of ST  (K1 for the second layer, once per call)
__global__ void __launch_bounds__(256)
k_b1_project_w2(const float* __restrict__ PW, const float* __restrict__ Z, long long n, int M, int S, int H, int O, int OT, int ST,
                long long w2_off, long long b2_off, float* __restrict__ out, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int o = (int)(e % OT);
    const int t = (int)((e / OT) % ST);
    const int j = (int)((e / OT / ST) % (H + 1));
    const long long blk = e / OT / ST / (H + 1);
    const long long s = blk * ST + t;
    float v = 0.0f;
    if (o < O && s < S) {
        const long long k = j < H ? w2_off + o + (long long)j * O : b2_off + o;
        v = PW[k + (long long)M * n];
        for (int m = 0; m < M; ++m) v = fmaf(PW[k + (long long)m * n], Z[s * M + m], v);
    }
    out[e] = v;
}

// tiles[tile][j][m][ii] = bases[m][tile*64 + ii][j]  (zero beyond N): one-time re-layout so that a CTA's
// basis tile is one contiguous block
__global__ void __launch_bounds__(256)
k_b1_tile_layout(const float* __restrict__ bases, long long N, int ld, int H, int M1, float* __restrict__ tiles, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int ii = (int)(e % B1_TI);
    const int m = (int)((e / B1_TI) % M1);
    const int j = (int)((e / B1_TI / M1) % H);
    const long long tile = e / B1_TI / M1 / H;
    const long long i = tile * B1_TI + ii;
    tiles[e] = i < N ? bases[((long long)m * N + i) * ld + j] : 0.0f;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct ssi_b1_state {
    bool ready = false;
    float* tiles = nullptr;
    int n_tiles = 0;
};

void ssi_b1_invalidate(ssi_ctx* ctx) {
    if (ctx->b1) ctx->b1->ready = false;
}

void ssi_b1_destroy(ssi_ctx* ctx) {
    if (!ctx->b1) return;
    cudaFree(ctx->b1->tiles);
    delete ctx->b1;
    ctx->b1 = nullptr;
}

static int b1_st(int M) { return M <= 6 ? 8 : (M <= 12 ? 4 : 2); }   // samples per warp pass: ST * M uniform registers

static size_t b1_smem_bytes(const ssi_ctx* ctx) {
    const ssi_model_t& m = ctx->model;
    const int H = m.dims[1], M = ctx->M, OT = m.dims[2] <= 1 ? 1 : 2;
    (void)OT;
    return sizeof(float) * ((size_t)H * (M + 1) * B1_TI);
}

static bool b1_m_supported(int M) { return (M >= 1 && M <= 8) || M == 10 || M == 12 || M == 16 || M == 20; }

bool ssi_b1_supported(const ssi_ctx* ctx) {
    if (!ctx->has_model || !ctx->has_sub) return false;
    const ssi_model_t& m = ctx->model;
    if (m.L != 2 || m.dims[2] > 2 || !b1_m_supported(ctx->M)) return false;
    if (ctx->N >= (1ll << 31) - B1_TI) return false;
    // two CTAs per SM: the basis tile must leave room for both
    return b1_smem_bytes(ctx) <= 110 * 1024;
}

int ssi_b1_prepare(ssi_ctx* ctx) {
    if (!ssi_b1_supported(ctx)) return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "basis path: needs Chain(Dense, Dense) with O <= 2 and a supported M");
    if (!ctx->b1) ctx->b1 = new ssi_b1_state();
    ssi_b1_state* s = ctx->b1;
    if (s->ready) return SSI_OK;
    const ssi_model_t& m = ctx->model;
    const int H = m.dims[1], M1 = ctx->M + 1;
    const int64_t N = ctx->N;
    cudaFree(s->tiles);
    s->tiles = nullptr;
    s->n_tiles = (int)((N + B1_TI - 1) / B1_TI);
    const long long total = (long long)s->n_tiles * H * M1 * B1_TI;
    SSI_CUDA(ctx, cudaMalloc(&s->tiles, sizeof(float) * (size_t)total));
    // bases [m][N][H] in scratch, exact FP32 (bias parts included)
    SSI_TRY(ssi_reserve(ctx, ctx->bH0, sizeof(float) * (size_t)M1 * N * H));
    SSI_TRY(ssi_build_first_layer_bases(ctx, (float*)ctx->bH0.p, H));
    k_b1_tile_layout<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->bH0.p, N, H, H, M1, s->tiles, total);
    SSI_LAUNCH_CHECK(ctx);
    s->ready = true;
    return SSI_OK;
}

typedef void (*b1_kernel_t)(const b1_params, const int);

template <int M, int OT>
static b1_kernel_t b1_pick_act(int act) {
    constexpr int ST = M <= 6 ? 8 : (M <= 12 ? 4 : 2);
    switch (act) {
        case SSI_ACT_RELU: return k_logpost_basis1h<M, ST, OT, SSI_ACT_RELU>;
        case SSI_ACT_TANH: return k_logpost_basis1h<M, ST, OT, SSI_ACT_TANH>;
        default:           return k_logpost_basis1h<M, ST, OT, -1>;
    }
}
template <int OT>
static b1_kernel_t b1_pick(int M, int act) {
    switch (M) {
        case 1: return b1_pick_act<1, OT>(act);
        case 2: return b1_pick_act<2, OT>(act);
        case 3: return b1_pick_act<3, OT>(act);
        case 4: return b1_pick_act<4, OT>(act);
        case 5: return b1_pick_act<5, OT>(act);
        case 6: return b1_pick_act<6, OT>(act);
        case 7: return b1_pick_act<7, OT>(act);
        case 8: return b1_pick_act<8, OT>(act);
        case 10: return b1_pick_act<10, OT>(act);
        case 12: return b1_pick_act<12, OT>(act);
        case 16: return b1_pick_act<16, OT>(act);
        case 20: return b1_pick_act<20, OT>(act);
        default: return nullptr;
    }
}

int ssi_reduce_partials(ssi_ctx* ctx, const double* partials, int64_t B, int parts, double* d_out);

// second-layer weights of the S samples of one launch, in blocks of ST samples; returns the device pointer (ctx scratch)
static int b1_project_w2(ssi_ctx* ctx, const float* dZs, int S, int ST, int OT, const float** out) {
    const ssi_model_t& m = ctx->model;
    const int H = m.dims[1];
    const long long blocks = (S + ST - 1) / ST;
    const long long total = blocks * (H + 1) * ST * OT;
    SSI_TRY(ssi_reserve(ctx, ctx->bW, sizeof(float) * (size_t)total));
    k_b1_project_w2<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->dP, dZs, m.n, ctx->M, S, H, m.dims[2], OT, ST,
                                                                             m.w_off[1], m.b_off[1], (float*)ctx->bW.p, total);
    SSI_LAUNCH_CHECK(ctx);
    *out = (const float*)ctx->bW.p;
    return SSI_OK;
}

typedef void (*b1g_kernel_t)(const b1_params, const int, const float, float*, const long long);
template <int M>
static b1g_kernel_t b1g_pick_act(int act) {
    switch (act) {
        case SSI_ACT_RELU: return k_logpost_basis1h_grad<M, SSI_ACT_RELU>;
        case SSI_ACT_TANH: return k_logpost_basis1h_grad<M, SSI_ACT_TANH>;
        default:           return k_logpost_basis1h_grad<M, -1>;
    }
}
static b1g_kernel_t b1g_pick(int M, int act) {
    switch (M) {
        case 1: return b1g_pick_act<1>(act);
        case 2: return b1g_pick_act<2>(act);
        case 3: return b1g_pick_act<3>(act);
        case 4: return b1g_pick_act<4>(act);
        case 5: return b1g_pick_act<5>(act);
        case 6: return b1g_pick_act<6>(act);
        case 7: return b1g_pick_act<7>(act);
        case 8: return b1g_pick_act<8>(act);
        default: return nullptr;
    }
}

static size_t b1g_smem_bytes(const ssi_ctx* ctx) {
    const ssi_model_t& m = ctx->model;
    const int H = m.dims[1], M = ctx->M;
    const int w2n = (H + 1 + 3) / 4 * 4;
    return sizeof(float) * ((size_t)H * (M + 1) * B1_TI + (size_t)(M + 1) * w2n);
}

// the fast value+gradient path: one hidden layer, scalar output, M <= 8 (tangent accumulators live in registers)
bool ssi_b1_grad_supported(const ssi_ctx* ctx) {
    return ssi_b1_supported(ctx) && ctx->model.dims[2] == 1 && ctx->M <= 8 && b1g_smem_bytes(ctx) <= 110 * 1024;
}

int ssi_b1_grad_sse(ssi_ctx* ctx, const float* dZ, int64_t B, double coef, double* d_sse, float** d_gpart, int* n_parts) {
    SSI_TRY(ssi_b1_prepare(ctx));
    ssi_b1_state* s = ctx->b1;
    const ssi_model_t& m = ctx->model;
    const int M = ctx->M, H = m.dims[1];
    b1g_kernel_t kern = b1g_pick(M, m.act[0]);
    if (!kern) return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "basis gradient path: M=%d is not instantiated", M);
    const size_t smem = b1g_smem_bytes(ctx);
    SSI_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSI_TRY(ssi_reserve(ctx, ctx->bPartials, sizeof(double) * (size_t)B * s->n_tiles));
    SSI_TRY(ssi_reserve(ctx, ctx->bGradP, sizeof(float) * (size_t)s->n_tiles * M * B));
    double* partials = (double*)ctx->bPartials.p;
    float* gpart = (float*)ctx->bGradP.p;
    const long long slab = (long long)M * B;

    const int ST = B1G_ST;
    // one launch per 2^20 samples (int indices inside the kernel); the warp-uniform z are read from global memory
    const int64_t per = 1 << 20;
    for (int64_t b0 = 0; b0 < B; b0 += per) {
        const int S = (int)std::min<int64_t>(per, B - b0);
        b1_params p{};
        p.N = (int)ctx->N; p.H = H; p.O = 1; p.n_tiles = s->n_tiles; p.S = S; p.act_out = m.act[1];
        p.w2n = (H + 1 + 3) / 4 * 4;
        int spc = 256;
        while (spc > B1_WARPS * ST && (long long)s->n_tiles * ((S + spc - 1) / spc) < 8ll * ctx->sm_count) spc >>= 1;
        p.s_per_cta = std::max(spc, B1_WARPS * ST);
        p.n = m.n; p.w2_off = m.w_off[1]; p.b2_off = m.b_off[1];
        p.tiles = s->tiles; p.PW = ctx->dP; p.Z = dZ + b0 * M; p.Y = ctx->dY;
        p.partials = partials + b0 * s->n_tiles;
        SSI_TRY(b1_project_w2(ctx, p.Z, S, ST, 1, &p.W2));
        dim3 grid(s->n_tiles, (S + p.s_per_cta - 1) / p.s_per_cta);
        ssi_kt_begin(ctx);
        kern<<<grid, B1_THREADS, smem, ctx->stream>>>(p, m.act[0], (float)coef, gpart + b0 * M, slab);
        SSI_LAUNCH_CHECK(ctx);
        ssi_kt_end(ctx);
    }
    *d_gpart = gpart;
    *n_parts = s->n_tiles;
    return ssi_reduce_partials(ctx, partials, B, s->n_tiles, d_sse);
}

int ssi_b1_sse(ssi_ctx* ctx, const float* dZ, int64_t B, double* d_sse) {
    SSI_TRY(ssi_b1_prepare(ctx));
    ssi_b1_state* s = ctx->b1;
    const ssi_model_t& m = ctx->model;
    const int M = ctx->M, H = m.dims[1], O = m.dims[2], OT = O <= 1 ? 1 : 2;
    b1_kernel_t kern = OT == 1 ? b1_pick<1>(M, m.act[0]) : b1_pick<2>(M, m.act[0]);
    if (!kern) return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "basis path: M=%d is not instantiated", M);
    const size_t smem = b1_smem_bytes(ctx);
    SSI_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSI_TRY(ssi_reserve(ctx, ctx->bPartials, sizeof(double) * (size_t)B * s->n_tiles));
    double* partials = (double*)ctx->bPartials.p;

    const int ST = b1_st(M);
    // one launch per 2^20 samples (int indices inside the kernel); the warp-uniform z are read from global memory
    const int64_t per = 1 << 20;
    for (int64_t b0 = 0; b0 < B; b0 += per) {
        const int S = (int)std::min<int64_t>(per, B - b0);
        b1_params p{};
        p.N = (int)ctx->N; p.H = H; p.O = O; p.n_tiles = s->n_tiles; p.S = S; p.act_out = m.act[1];
        p.w2n = ((H + 1) * OT + 3) / 4 * 4;
        // samples per CTA: enough to amortise staging the basis tile, few enough to fill the machine several times over
        int spc = 512;
        while (spc > B1_WARPS * ST && (long long)s->n_tiles * ((S + spc - 1) / spc) < 8ll * ctx->sm_count) spc >>= 1;
        p.s_per_cta = std::max(spc, B1_WARPS * ST);
        p.n = m.n; p.w2_off = m.w_off[1]; p.b2_off = m.b_off[1];
        p.tiles = s->tiles; p.PW = ctx->dP; p.Z = dZ + b0 * M; p.Y = ctx->dY;
        p.partials = partials + b0 * s->n_tiles;
        SSI_TRY(b1_project_w2(ctx, p.Z, S, ST, OT, &p.W2));
        dim3 grid(s->n_tiles, (S + p.s_per_cta - 1) / p.s_per_cta);
        ssi_kt_begin(ctx);
        kern<<<grid, B1_THREADS, smem, ctx->stream>>>(p, m.act[0]);
        SSI_LAUNCH_CHECK(ctx);
        ssi_kt_end(ctx);
    }
    return ssi_reduce_partials(ctx, partials, B, s->n_tiles, d_sse);
}
