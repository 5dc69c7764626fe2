// ssi_decoder.cu — a non-linear subspace operator in place of P.
//
// Replaces, for the density of auto_inference (src/space_inference.jl:246-251; decoder trained by
// auto_encoder_subspace, src/subspace_construction.jl:125-141),
//     new_W = W_swa + decoder(z)
// where `decoder` is a Flux Chain of Dense layers  z (Mz) -> ... -> h (Hd) -> Dense(Hd, n, act_out).
// The last layer is affine up to its activation:  decoder(z) = act_out(P' h(z) + c)  with  P' = its n x Hd weight
// matrix (column-major, exactly the layout of P) and c its bias.  With act_out = identity (the usual decoder head)
//     new_W = (W_swa + c) + P' h(z)
// is the reference's affine subspace again, of dimension Hd, evaluated at z' = h(z): the whole hot path (projection fused into
// the GEMM operands, first layer affine in z', tensor cores, M-space prior) applies unchanged, and this file adds only
// the tiny per-sample map h and its transpose-Jacobian for gradients.  A non-linear head needs the weights
// materialised per sample: that case runs on the LAYERED path (W = W_swa + act_out(c + P' z')).
// The sampler keeps working in the user's z (Mz): proposals, traces and the z prior never see z'.
#include "ssi_common.cuh"

#include <algorithm>
#include <cmath>

struct ssi_decoder_t {
    int L = 0;                                 // Dense layers of the decoder, the head included
    int dims[SSI_MAX_LAYERS + 1] = {0};
    int act[SSI_MAX_LAYERS] = {0};
    int64_t w_off[SSI_MAX_LAYERS] = {0}, b_off[SSI_MAX_LAYERS] = {0};    // offsets into theta (Flux.destructure order)
    float* theta = nullptr;                    // parameters of the layers before the head, on the device
    int64_t n_hidden_params = 0;
};

struct dec_desc_t {
    int L;                                     // hidden layers (the head excluded)
    int dims[SSI_MAX_LAYERS + 1];
    int act[SSI_MAX_LAYERS];
    long long w_off[SSI_MAX_LAYERS], b_off[SSI_MAX_LAYERS];
};

// z' = h(z): one thread per sample, widths <= SSI_MAX_M
__global__ void __launch_bounds__(128)
k_dec_forward(const dec_desc_t d, const float* __restrict__ theta, const float* __restrict__ Z, long long B, float* __restrict__ Zin) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float h[2][SSI_MAX_M];
    for (int i = 0; i < d.dims[0]; ++i) h[0][i] = Z[i + b * d.dims[0]];
    int cur = 0;
    for (int l = 0; l < d.L; ++l) {
        const int in = d.dims[l], out = d.dims[l + 1];
        const float* W = theta + d.w_off[l];
        const float* bias = theta + d.b_off[l];
        for (int o = 0; o < out; ++o) {
            float v = bias[o];
            for (int i = 0; i < in; ++i) v = fmaf(W[o + (long long)i * out], h[cur][i], v);
            h[cur ^ 1][o] = ssi_act(v, d.act[l]);
        }
        cur ^= 1;
    }
    const int Hd = d.dims[d.L];
    for (int o = 0; o < Hd; ++o) Zin[o + b * Hd] = h[cur][o];
}

// g = J_h(z)' g'  (reverse pass through the hidden layers; activations recomputed and kept per thread)
__global__ void __launch_bounds__(64)
k_dec_backward(const dec_desc_t d, const float* __restrict__ theta, const float* __restrict__ Z, const double* __restrict__ Gin, long long B,
               double* __restrict__ Gout) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float h[SSI_MAX_LAYERS + 1][SSI_MAX_M];
    for (int i = 0; i < d.dims[0]; ++i) h[0][i] = Z[i + b * d.dims[0]];
    for (int l = 0; l < d.L; ++l) {
        const int in = d.dims[l], out = d.dims[l + 1];
        const float* W = theta + d.w_off[l];
        const float* bias = theta + d.b_off[l];
        for (int o = 0; o < out; ++o) {
            float v = bias[o];
            for (int i = 0; i < in; ++i) v = fmaf(W[o + (long long)i * out], h[l][i], v);
            h[l + 1][o] = ssi_act(v, d.act[l]);
        }
    }
    double g[2][SSI_MAX_M];
    const int Hd = d.dims[d.L];
    for (int o = 0; o < Hd; ++o) g[0][o] = Gin[o + b * Hd];
    int cur = 0;
    for (int l = d.L - 1; l >= 0; --l) {
        const int in = d.dims[l], out = d.dims[l + 1];
        const float* W = theta + d.w_off[l];
        for (int i = 0; i < in; ++i) g[cur ^ 1][i] = 0.0;
        for (int o = 0; o < out; ++o) {
            const double delta = g[cur][o] * (double)act_deriv_from_output(h[l + 1][o], d.act[l]);
            for (int i = 0; i < in; ++i) g[cur ^ 1][i] += (double)W[o + (long long)i * out] * delta;
        }
        cur ^= 1;
    }
    for (int i = 0; i < d.dims[0]; ++i) Gout[i + b * d.dims[0]] = g[cur][i];
}

// lp = [LL] ll + [PRIOR_W] pw + [PRIOR_Z] logN(z; 0, sigma_z^2 I) with z the USER's z; terms rewritten accordingly;
// grad (Mz x B, optional) += d prior_z / dz
__global__ void k_dec_combine(const double* __restrict__ terms_in /* 3 x B from the inner evaluation */, const float* __restrict__ Z, int Mz, long long B,
                              uint32_t mask, double c_z, double inv2sz2, double* __restrict__ lp, double* __restrict__ terms_out,
                              double* __restrict__ grad) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double z2 = 0.0;
    for (int i = 0; i < Mz; ++i) { const double zi = (double)Z[i + b * Mz]; z2 += zi * zi; }
    const double ll = terms_in[3 * b], pw = terms_in[3 * b + 1], pz = c_z - z2 * inv2sz2;
    double v = 0.0;
    if (mask & SSI_TERM_LL) v += ll;
    if (mask & SSI_TERM_PRIOR_W) v += pw;
    if (mask & SSI_TERM_PRIOR_Z) v += pz;
    lp[b] = v;
    if (terms_out) { terms_out[3 * b] = ll; terms_out[3 * b + 1] = pw; terms_out[3 * b + 2] = pz; }
    if (grad && (mask & SSI_TERM_PRIOR_Z))
        for (int i = 0; i < Mz; ++i) grad[i + b * Mz] -= 2.0 * inv2sz2 * (double)Z[i + b * Mz];
}

void ssi_dec_destroy(ssi_ctx* ctx) {
    if (!ctx->dec) return;
    cudaFree(ctx->dec->theta);
    delete ctx->dec;
    ctx->dec = nullptr;
    ctx->dec_active = false;
    cudaFree(ctx->dWbase);
    ctx->dWbase = nullptr;
    ctx->dec_out_act = SSI_ACT_IDENTITY;
}

static dec_desc_t dec_desc(const ssi_decoder_t* dc) {
    dec_desc_t d{};
    d.L = dc->L - 1;
    for (int l = 0; l <= d.L; ++l) d.dims[l] = dc->dims[l];
    for (int l = 0; l < d.L; ++l) { d.act[l] = dc->act[l]; d.w_off[l] = dc->w_off[l]; d.b_off[l] = dc->b_off[l]; }
    return d;
}

int ssi_install_subspace_host(ssi_ctx* ctx, const float* W_swa, const float* P, int64_t n, int M);

int ssi_set_decoder_impl(ssi_ctx* ctx, const float* W_swa, int n_layers, const int32_t* dims, const int32_t* act, const float* theta) {
    if (!ctx->has_model) return ssi_fail(ctx, SSI_ERR_STATE, "ssi_set_model must be called before ssi_set_decoder");
    if (n_layers < 1 || n_layers > SSI_MAX_LAYERS || !dims || !act || !theta || !W_swa)
        return ssi_fail(ctx, SSI_ERR_ARG, "decoder: n_layers must be in [1,%d] and W_swa, dims, act, theta non-NULL", SSI_MAX_LAYERS);
    const int64_t n = ctx->model.n;
    if (dims[n_layers] != n) return ssi_fail(ctx, SSI_ERR_ARG, "decoder output width %d does not match the model's %lld parameters", dims[n_layers], (long long)n);
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l] < 1 || dims[l] > SSI_MAX_M)
            return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "decoder widths before the head must be in [1,%d] (dims[%d] = %d)", SSI_MAX_M, l, dims[l]);
        if (act[l] < SSI_ACT_IDENTITY || act[l] > SSI_ACT_SIGMOID) return ssi_fail(ctx, SSI_ERR_ARG, "decoder act[%d] is not a supported activation", l);
    }
    SSI_TRY(ssi_use_device(ctx));
    SSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ssi_dec_destroy(ctx);
    ssi_decoder_t* dc = new ssi_decoder_t();
    dc->L = n_layers;
    int64_t off = 0;
    for (int l = 0; l <= n_layers; ++l) dc->dims[l] = dims[l];
    for (int l = 0; l < n_layers; ++l) {
        dc->act[l] = act[l];
        dc->w_off[l] = off; off += (int64_t)dims[l] * dims[l + 1];
        dc->b_off[l] = off; off += dims[l + 1];
    }
    const int head = n_layers - 1, Hd = dims[head];
    dc->n_hidden_params = dc->w_off[head];
    ctx->dec = dc;
    if (dc->n_hidden_params > 0) {
        SSI_CUDA(ctx, cudaMalloc(&dc->theta, sizeof(float) * (size_t)dc->n_hidden_params));
        SSI_CUDA(ctx, cudaMemcpyAsync(dc->theta, theta, sizeof(float) * (size_t)dc->n_hidden_params, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the inner affine subspace: P' = head weights (n x Hd column-major), offset = W_swa + c (identity head) or c alone
    const float* Phead = theta + dc->w_off[head];
    const float* chead = theta + dc->b_off[head];
    std::vector<float> w0((size_t)n);
    const bool affine = act[head] == SSI_ACT_IDENTITY;
    for (int64_t i = 0; i < n; ++i) w0[(size_t)i] = affine ? W_swa[i] + chead[i] : chead[i];
    SSI_TRY(ssi_install_subspace_host(ctx, w0.data(), Phead, n, Hd));
    if (!affine) {
        SSI_CUDA(ctx, cudaMalloc(&ctx->dWbase, sizeof(float) * (size_t)n));
        SSI_CUDA(ctx, cudaMemcpyAsync(ctx->dWbase, W_swa, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        ctx->dec_out_act = act[head];
    }
    SSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->Mz = dims[0];
    ctx->dec_active = true;
    return SSI_OK;
}

// z' = h(z) for B samples into the context's scratch; *dZin = the mapped points (Hd x B)
static int dec_map(ssi_ctx* ctx, const float* dZ, int64_t B, const float** dZin) {
    const ssi_decoder_t* dc = ctx->dec;
    if (dc->L == 1) { *dZin = dZ; return SSI_OK; }            // a single Dense(Mz, n): z' = z
    SSI_TRY(ssi_reserve(ctx, ctx->bDecZ, sizeof(float) * (size_t)ctx->M * B));
    k_dec_forward<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(dec_desc(dc), dc->theta, dZ, B, (float*)ctx->bDecZ.p);
    SSI_LAUNCH_CHECK(ctx);
    *dZin = (const float*)ctx->bDecZ.p;
    return SSI_OK;
}

int ssi_dec_project(ssi_ctx* ctx, const float* dZ, int64_t B, const float** dZin) { return dec_map(ctx, dZ, B, dZin); }

static int dec_check_mask(ssi_ctx* ctx, uint32_t mask) {
    if ((mask & ~(SSI_TERM_LL | SSI_TERM_PRIOR_W | SSI_TERM_PRIOR_Z)) || mask == 0)
        return ssi_fail(ctx, SSI_ERR_ARG, "prior_mask must be a non-empty OR of SSI_TERM_*");
    if ((mask & SSI_TERM_PRIOR_W) && ctx->dec_out_act != SSI_ACT_IDENTITY)
        return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "the weight prior is evaluated in the subspace and needs a decoder whose last layer has the identity activation");
    return SSI_OK;
}

int ssi_dec_logpost(ssi_ctx* ctx, const float* dZ, int64_t B, double sigma_m, double sigma_p, double sigma_z, uint32_t mask,
                    double* d_lp, double* d_terms) {
    SSI_TRY(dec_check_mask(ctx, mask));
    if (!(sigma_z > 0)) return ssi_fail(ctx, SSI_ERR_ARG, "sigma_m, sigma_p, sigma_z must be positive");
    const float* dZin = nullptr;
    SSI_TRY(dec_map(ctx, dZ, B, &dZin));
    SSI_TRY(ssi_reserve(ctx, ctx->bDecT, sizeof(double) * 3 * (size_t)B));
    double* t_in = (double*)ctx->bDecT.p;
    const uint32_t inner = SSI_TERM_LL | ((mask & SSI_TERM_PRIOR_W) ? SSI_TERM_PRIOR_W : 0u);
    ctx->dec_active = false;                       // the inner evaluation is the reference's affine path at z'
    const int rc = ssi_logpost_device(ctx, dZin, B, sigma_m, sigma_p, sigma_z, inner, d_lp, t_in);
    ctx->dec_active = true;
    if (rc != SSI_OK) return rc;
    const double LOG_2PI = 1.8378770664093454835606594728112;
    const double c_z = -0.5 * (double)ctx->Mz * LOG_2PI - (double)ctx->Mz * std::log(sigma_z);
    k_dec_combine<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(t_in, dZ, ctx->Mz, B, mask, c_z, 1.0 / (2.0 * sigma_z * sigma_z), d_lp, d_terms, nullptr);
    SSI_LAUNCH_CHECK(ctx);
    return SSI_OK;
}

int ssi_dec_logpost_grad(ssi_ctx* ctx, const float* dZ, int64_t B, double sigma_m, double sigma_p, double sigma_z, uint32_t mask,
                         double* d_lp, double* d_grad) {
    SSI_TRY(dec_check_mask(ctx, mask));
    if (ctx->dec_out_act != SSI_ACT_IDENTITY)
        return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "gradients need a decoder whose last layer has the identity activation");
    if (!(mask & SSI_TERM_LL)) return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "gradients through a decoder need the likelihood term in prior_mask");
    const ssi_decoder_t* dc = ctx->dec;
    const float* dZin = nullptr;
    SSI_TRY(dec_map(ctx, dZ, B, &dZin));
    SSI_TRY(ssi_reserve(ctx, ctx->bDecT, sizeof(double) * (3 + (size_t)ctx->M) * (size_t)B));
    double* t_in = (double*)ctx->bDecT.p;
    double* g_in = t_in + 3 * (size_t)B;
    const uint32_t inner = SSI_TERM_LL | ((mask & SSI_TERM_PRIOR_W) ? SSI_TERM_PRIOR_W : 0u);
    ctx->dec_active = false;
    int rc = ssi_logpost_grad_device(ctx, dZin, B, sigma_m, sigma_p, sigma_z, inner, d_lp, g_in);
    // the inner terms (ll, pw) for the combination: the value pass is cheap next to the gradient
    if (rc == SSI_OK && !(mask == inner)) rc = ssi_logpost_device(ctx, dZin, B, sigma_m, sigma_p, sigma_z, inner, d_lp, t_in);
    ctx->dec_active = true;
    if (rc != SSI_OK) return rc;
    if (dc->L == 1) {
        SSI_CUDA(ctx, cudaMemcpyAsync(d_grad, g_in, sizeof(double) * (size_t)ctx->Mz * B, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        k_dec_backward<<<(unsigned)((B + 63) / 64), 64, 0, ctx->stream>>>(dec_desc(dc), dc->theta, dZ, g_in, B, d_grad);
        SSI_LAUNCH_CHECK(ctx);
    }
    if (!(mask == inner)) {
        const double LOG_2PI = 1.8378770664093454835606594728112;
        const double c_z = -0.5 * (double)ctx->Mz * LOG_2PI - (double)ctx->Mz * std::log(sigma_z);
        k_dec_combine<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(t_in, dZ, ctx->Mz, B, mask, c_z, 1.0 / (2.0 * sigma_z * sigma_z), d_lp, nullptr, d_grad);
        SSI_LAUNCH_CHECK(ctx);
    }
    return SSI_OK;
}
