// ssi_gram_tc.cu — K6 on the tensor cores: G = A'A for a tall-skinny n x K FP32 matrix (K <= 128).
//
// Replaces the factorisation work inside psvd(A) (src/subspace_construction.jl:63) for the shapes the
// construction path produces (K = number of deviation columns, n ~ 1e7): one pass over A at HBM speed.
//
//   * A is column-major n x K == K rows of length n.  A TMA box {32 rows-of-A, 128 columns-of-A} lands in a RAW ring of
//     six 16 KB slots as a [128][32] FP32 tile with 128-byte swizzle (columns >= K are zero-filled by TMA).  The ring
//     is deep because a tile gathers one 128-byte segment from each of the K columns (K different DRAM pages).
//   * six converter warps turn a raw tile into the operand of tcgen05.mma kind::f16: two FP16 planes H = fp16(s x) and
//     L = fp16(s x - H), [128][32] halves each = 64-byte rows, K-major SWIZZLE_64B, H and L back to back (16 KB per
//     operand slot, six slots).  s = 2^j puts the largest |A| (tracked by k_swa_push, one atomicMax per block) at 2^14,
//     so H cannot overflow and H + L carries 22 bits of every value that matters.  Round 1 split into TF32 planes in
//     place (FP32 containers): 109 KB of shared-memory traffic per 16 KB of A, 80 % of the crossbar, 52 % of the HBM
//     peak; 16-bit planes make it 71 KB and halve the tensor-core time.
//   * one thread issues, per 16 rows of A,  [HH | HL] += H [H ; L]'  (the B operand spans the H tile and the L tile
//     behind it, so H is fetched once; FP32 accumulators in TMEM).  G = HH + HL + HL' drops only L L', 2^-24 relative.
//     (BF16 planes were tried: their L L' term is 2^-18 of H H' and all-positive on the diagonal, so it must be kept,
//     and accumulated next to H H' in FP32 a third of it was truncated away -- 3e-7 of lambda_0, measured.  Loading A
//     with plain 16-byte loads straight into the operand layout -- no raw ring -- was tried too: 1.5 ms, bound by the
//     latency of its own loads.)
//   * FP32 accumulation in the tensor core truncates, so accumulators are drained every `chunk` tiles (1024 rows) into
//     per-CTA FP64 partials (each thread owns its elements: plain load-add-store, no atomics) and reset; a final
//     kernel sums the per-CTA partials in fixed order, symmetrises and undoes the scale.
//
// Roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-7 converters (one tile each),
// warps 8-11 drain.  TMEM: 2 x (HH | HL) x 128 columns = 512.
#include "ssi_common.cuh"
#include "ssi_ptx.cuh"

#include <cuda_fp16.h>
#include <algorithm>

#define GT_ROWS 32                   // rows of A per tile (= 128 bytes of FP32 = one swizzle row)
#define GT_RAW_BYTES (128 * 128)     // [128 columns of A][32 rows] FP32
#define GT_OP_BYTES (2 * 128 * 64)   // H | L: [128 columns of A][32 rows] FP16 each
#define GT_CONV 6                    // converter warps: the split costs ~2000 cycles of latency per tile and warp, four warps were
                                     // the bottleneck (970 cycles per tile at 56 % of the HBM peak)
#define GT_RAW GT_CONV               // raw and operand slots: one of each per converter warp (tile t -> slot t % 6 = its converter),
#define GT_OPS GT_CONV               // so that a warp only ever waits on barriers it cycles itself and can never run two phases
                                     // ahead (an mbarrier parity wait cannot tell phase p from phase p + 2)
#define GT_THREADS (64 + 32 * GT_CONV + 128)
#define GT_OFF_OP (GT_RAW * GT_RAW_BYTES)
#define GT_OFF_BAR (GT_OFF_OP + GT_OPS * GT_OP_BYTES)
#define GT_NBAR (2 * GT_RAW + 2 * GT_OPS + 4)
#define GT_SMEM_TOTAL (GT_OFF_BAR + GT_NBAR * 8 + 16)

struct gram_tc_params {
    const float* amax;         // device: largest |A| (from k_swa_push), defines the power-of-two scale of the FP16 planes
    long long n;
    int K, NP;                 // NP = K rounded up to 16 (MMA N)
    long long tiles_total;     // ceil(n / 32)
    long long tiles_per_cta;
    int chunk;                 // tiles per FP32 accumulation chunk
    double* partial;           // [grid][2][NP][128]  (pass, column b, row a)
};

__global__ void __launch_bounds__(GT_THREADS, 1)
k_gram_tc(const __grid_constant__ CUtensorMap tmA, const gram_tc_params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // barriers: raw_full[R] raw_empty[R] op_full[O] op_empty[O] tfull[2] tempty[2]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + GT_NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_rfull = smem_u32(s_bar), bar_rempty = bar_rfull + 8 * GT_RAW;
    const uint32_t bar_ofull = bar_rempty + 8 * GT_RAW, bar_oempty = bar_ofull + 8 * GT_OPS;
    const uint32_t bar_tfull = bar_oempty + 8 * GT_OPS, bar_tempty = bar_tfull + 16;

    if (threadIdx.x == 0) {
        if (smem_base & 1023u) { printf("ssi_gram_tc: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < GT_RAW; ++s) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, 32); }
        for (int s = 0; s < GT_OPS; ++s) { mbar_init(bar_ofull + 8 * s, 32); mbar_init(bar_oempty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 128); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const long long t_begin = (long long)blockIdx.x * p.tiles_per_cta;
    const long long t_end = min(p.tiles_total, t_begin + p.tiles_per_cta);
    const long long my_tiles = max(0ll, t_end - t_begin);
    const long long n_chunks = (my_tiles + p.chunk - 1) / p.chunk;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t_begin; t < t_end; ++t) {
                mbar_wait(bar_rempty + 8 * stage, phase ^ 1);          // the converter that read this raw slot is done with it
                mbar_expect_tx(bar_rfull + 8 * stage, GT_RAW_BYTES);
                // A is streamed exactly once: do not let it evict the FP64 partials from L2
                tma_load_2d_hint(smem_base + stage * GT_RAW_BYTES, &tmA, bar_rfull + 8 * stage, (int)(t * GT_ROWS), 0, TC_EVICT_FIRST);
                if (++stage == GT_RAW) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // one MMA per k-step: the B operand spans the H tile (128 rows; rows >= K are zero) and the first NP rows of the
            // L tile behind it, so D = H [H ; L]^T = [HH (columns 0..127) | HL (columns 128..128+NP)] and the A operand (H)
            // is fetched from shared memory once instead of twice
            const uint32_t idesc = umma_idesc_f16(128 + p.NP, UMMA_FMT_F16, UMMA_FMT_F16);
            int stage = 0;
            uint32_t phase = 0;
            long long t = 0;
            for (long long c = 0; c < n_chunks; ++c) {
                const uint32_t ab = (uint32_t)(c & 1), aphase = (uint32_t)((c >> 1) & 1);
                mbar_wait(bar_tempty + 8 * ab, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_acc = tmem_base + ab * 256;
                const long long c_end = min(my_tiles, (c + 1) * (long long)p.chunk);
                for (bool first = true; t < c_end; ++t, first = false) {
                    mbar_wait(bar_ofull + 8 * stage, phase);            // H and L tiles are ready
                    tc_fence_after();
                    const uint64_t dh = umma_desc_sw64(smem_base + GT_OFF_OP + stage * GT_OP_BYTES);
#pragma unroll
                    for (int k = 0; k < GT_ROWS / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 32 >> 4);    // 16 FP16 = 32 bytes per k-step inside the 64-byte row
                        umma_bf16(d_acc, dh + ko, dh + ko, idesc, !(first && k == 0));
                    }
                    umma_commit(bar_oempty + 8 * stage);
                    if (++stage == GT_OPS) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * ab);
            }
        }
    } else if (warp < 2 + GT_CONV) {
        // ================= converters: raw FP32 tile -> FP16 planes H | L in the K-major SWIZZLE_64B operand layout =================
        const int cw = warp - 2;                  // this warp takes tiles cw, cw + GT_CONV, ...
        const float amax = *p.amax;
        const float sc = (amax > 0.0f && amax < 3.0e38f) ? scalbnf(1.0f, max(-100, min(100, 14 - ilogbf(amax)))) : 1.0f;
        // lane -> (column c0 + lane / 4 of an 8-column group, 8-row chunk lane % 4): two adjacent 16-byte chunks of the raw
        // row (rows 8j .. 8j+7 of A) become one 16-byte chunk of each plane
        const int cl = lane >> 2, j = lane & 3;
        for (long long t = cw; t < my_tiles; t += GT_CONV) {
            const int rs = (int)(t % GT_RAW), os = (int)(t % GT_OPS);
            mbar_wait(bar_rfull + 8 * rs, (uint32_t)((t / GT_RAW) & 1));
            mbar_wait(bar_oempty + 8 * os, (uint32_t)(((t / GT_OPS) & 1) ^ 1));       // the MMAs that read this operand slot have retired
            const uint8_t* raw = smem + (size_t)rs * GT_RAW_BYTES;
            uint8_t* opH = smem + GT_OFF_OP + (size_t)os * GT_OP_BYTES;
            uint8_t* opL = opH + 128 * 64;
#pragma unroll 4
            for (int g8 = 0; g8 < 16; ++g8) {
                const int c = g8 * 8 + cl;                    // column of A = row of the operand; c & 7 == cl
                const float4 x0 = *reinterpret_cast<const float4*>(raw + c * 128 + (((2 * j) ^ cl) << 4));
                const float4 x1 = *reinterpret_cast<const float4*>(raw + c * 128 + (((2 * j + 1) ^ cl) << 4));
                const float v0 = x0.x * sc, v1 = x0.y * sc, v2 = x0.z * sc, v3 = x0.w * sc;
                const float v4 = x1.x * sc, v5 = x1.y * sc, v6 = x1.z * sc, v7 = x1.w * sc;
                const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                const __half2 h45 = __floats2half2_rn(v4, v5), h67 = __floats2half2_rn(v6, v7);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23), f45 = __half22float2(h45), f67 = __half22float2(h67);
                const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
                const __half2 l45 = __floats2half2_rn(v4 - f45.x, v5 - f45.y), l67 = __floats2half2_rn(v6 - f67.x, v7 - f67.y);
                const uint32_t off = (uint32_t)c * 64 + (uint32_t)((j ^ ((c >> 1) & 3)) << 4);      // SWIZZLE_64B: chunk ^ bits [7, 9) of the address
                *reinterpret_cast<uint4*>(opH + off) = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23),
                                                                  *reinterpret_cast<const uint32_t*>(&h45), *reinterpret_cast<const uint32_t*>(&h67));
                *reinterpret_cast<uint4*>(opL + off) = make_uint4(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23),
                                                                  *reinterpret_cast<const uint32_t*>(&l45), *reinterpret_cast<const uint32_t*>(&l67));
            }
            fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bar_ofull + 8 * os);
            mbar_arrive(bar_rempty + 8 * rs);
        }
    } else {
        // ================= drain: TMEM chunk sums -> FP64 per-CTA partials =================
        // The partials (grid x 2 x NP x 128 doubles, ~34 MB at K = 100) are read-modify-written once per chunk;
        // they are tagged evict_last so that they stay in the 126 MB L2 while A streams through (evict_first).
        const int q = warp & 3;                    // TMEM lane quarter
        const int a = q * 32 + lane;               // row of the accumulator = column a of A
        const uint64_t pol = l2_policy_evict_last();
        double* part = p.partial + (long long)blockIdx.x * 2 * p.NP * 128;
        for (long long c = 0; c < n_chunks; ++c) {
            const uint32_t ab = (uint32_t)(c & 1), aphase = (uint32_t)((c >> 1) & 1);
            mbar_wait(bar_tfull + 8 * ab, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                // HH is symmetric: this warp's rows [32q, 32q+32) only need the column blocks that reach the diagonal
#pragma unroll 1
                for (int b0 = (pass == 0 ? q * 32 : 0); b0 < p.NP; b0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + pass * 128 + b0, v);
                    tmem_ld_wait();
                    double* dst = part + ((long long)pass * p.NP + b0) * 128 + a;
                    if (c == 0) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) st_f64_hint(dst + jj * 128, (double)__uint_as_float(v[jj]), pol);
                    } else {
                        double old[16];
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) old[jj] = ld_f64_hint(dst + jj * 128, pol);
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) st_f64_hint(dst + jj * 128, old[jj] + (double)__uint_as_float(v[jj]), pol);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * ab);
        }
        if (n_chunks == 0) {                       // a CTA without rows still owns a (zero) partial
            for (int e = 0; e < 2 * p.NP; ++e) part[(long long)e * 128 + a] = 0.0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// stage 1: S[e] = sum over CTAs of partial[c][e], fixed order, fully coalesced (e over 2 x NP x 128)
__global__ void __launch_bounds__(128)
k_gram_tc_sum(const double* __restrict__ partial, int grid, int elems, int NP, double* __restrict__ S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    // HH entries (pass 0) left of a warp's diagonal block are never written by the drain warps (and never read by stage 2)
    const int row = e & 127, col = e >> 7;
    if (col < NP && col < (row & ~31)) { S[e] = 0.0; return; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
    for (; c + 4 <= grid; c += 4) {
        s0 += partial[(long long)c * elems + e];
        s1 += partial[(long long)(c + 1) * elems + e];
        s2 += partial[(long long)(c + 2) * elems + e];
        s3 += partial[(long long)(c + 3) * elems + e];
    }
    for (; c < grid; ++c) s0 += partial[(long long)c * elems + e];
    S[e] = (s0 + s1) + (s2 + s3);
}
// stage 2: G[a,b] = HH[min,max] + HL[a][b] + HL[b][a]   (S is [pass][column][row 128])
__global__ void k_gram_tc_sym(const double* __restrict__ S, int K, int NP, const float* __restrict__ amax, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * K) return;
    const int a = e % K, b = e / K;
    const int lo = min(a, b), hi = max(a, b);
    const float am = *amax;
    const int j = (am > 0.0f && am < 3.0e38f) ? max(-100, min(100, 14 - ilogbf(am))) : 0;      // the planes carried 2^j A
    G[a + (long long)b * K] = scalbn(S[(long long)hi * 128 + lo] + S[((long long)NP + b) * 128 + a] + S[((long long)NP + a) * 128 + b], -2 * j);
}

bool ssi_gram_tc_usable(const ssi_ctx* ctx, const float* dA, int64_t n, int64_t ld, int K) {
    return K >= 1 && K <= 128 && (ld % 4) == 0 && n >= (1 << 16) && ((reinterpret_cast<uintptr_t>(dA) & 15) == 0) &&
           n < (1ll << 31) && ctx->opt_gram_fp64 <= 0;
}

int ssi_gram_tc_device(ssi_ctx* ctx, const float* dA, int64_t n, int64_t ld, int K, double* dG, const float* d_amax) {
    PFN_ssi_encodeTiled encode = nullptr;
    SSI_TRY(ssi_tensormap_encoder(ctx, &encode));
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {GT_ROWS, 128};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dA), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ssi_fail(ctx, SSI_ERR_CUDA, "cuTensorMapEncodeTiled (gram) failed with CUresult %d", (int)r);

    gram_tc_params p{};
    p.amax = d_amax;
    p.n = n;
    p.K = K;
    p.NP = (K + 15) / 16 * 16;
    p.tiles_total = (n + GT_ROWS - 1) / GT_ROWS;
    const int grid = (int)std::min<long long>(ctx->sm_count, p.tiles_total);
    p.tiles_per_cta = (p.tiles_total + grid - 1) / grid;
    p.chunk = ctx->opt_gram_chunk > 0 ? ctx->opt_gram_chunk : 32;       // 1024 rows per FP32 chunk
    const int elems = 2 * p.NP * 128;
    SSI_TRY(ssi_reserve(ctx, ctx->bGram, sizeof(double) * (size_t)(grid + 1) * elems));
    p.partial = (double*)ctx->bGram.p;
    double* dS = p.partial + (size_t)grid * elems;
    SSI_CUDA(ctx, cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_TOTAL));
    ssi_kt_begin(ctx);
    k_gram_tc<<<grid, GT_THREADS, GT_SMEM_TOTAL, ctx->stream>>>(map, p);
    SSI_LAUNCH_CHECK(ctx);
    ssi_kt_end(ctx);
    k_gram_tc_sum<<<(elems + 127) / 128, 128, 0, ctx->stream>>>(p.partial, grid, elems, p.NP, dS);
    SSI_LAUNCH_CHECK(ctx);
    k_gram_tc_sym<<<(K * K + 255) / 256, 256, 0, ctx->stream>>>(dS, K, p.NP, d_amax, dG);
    SSI_LAUNCH_CHECK(ctx);
    return SSI_OK;
}
