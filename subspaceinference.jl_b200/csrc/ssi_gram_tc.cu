// ssi_gram_tc.cu — K6 on the tensor cores: G = A'A for a tall-skinny n x K FP32 matrix (K <= 128).
//
// Replaces the factorisation work inside psvd(A) (src/subspace_construction.jl:63) for the shapes the
// construction path produces (K = number of deviation columns, n ~ 1e7): one pass over A at HBM speed.
//
//   * A is column-major n x K == K rows of length n.  A TMA box {32 rows-of-A, 128 columns-of-A} lands in
//     shared memory as a [128][32] FP32 tile with 128-byte swizzle, which IS the K-major operand layout
//     of tcgen05.mma kind::tf32 (contraction index = row of A); columns >= K are zero-filled by TMA.
//   * converter warps split the tile in place: H = x with the low 13 mantissa bits cleared (exactly what
//     TF32 keeps), L = x - H (again cleared to TF32), written to a second tile.  The transform is
//     element-wise at identical addresses, so it is independent of the swizzle.
//   * one thread issues, per 8 rows of A,  HH += H H'  and  HL += H L'  (FP32 accumulators in TMEM);
//     G = HH + HL + HL' drops only the L L' term (2^-22 relative).
//   * FP32 accumulation in the tensor core truncates, so accumulators are drained every `chunk` tiles
//     into per-CTA FP64 partials (each thread owns its elements: plain load-add-store, no atomics) and
//     reset; a final kernel sums the per-CTA partials in fixed order and symmetrises.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-5 converters (one tile each),
// warps 6-9 drain.  TMEM: 2 x (HH | HL) x 128 columns = 512.  Shared memory: 6 slots of (H | L) tile pairs.
#include "ssi_common.cuh"
#include "ssi_ptx.cuh"

#include <algorithm>

#define GT_ROWS 32                   // rows of A per tile (= 128 bytes of FP32 = one swizzle row)
#define GT_TILE_BYTES (128 * 128)    // [128 columns of A][32 rows] FP32
// One ring of (H | L) slot pairs.  TMA lands the FP32 tile in the first half of a slot; ONE converter warp per tile (four
// tiles are being split at any time) overwrites it with H in place and writes L behind it; the slot returns to the producer
// when its MMAs have retired.  Deep, because a tile gathers one 128-byte segment from each of the K columns (K different
// DRAM pages).  (An earlier version had all four converter warps share each tile and separate raw / operand rings: the same
// 1.2 ms at n = 10M, K = 100 -- neither the converters' latency chain nor the FP32 chunk length is what holds the kernel at
// 52 % of the HBM peak; tensor pipe, shared memory and DRAM all sit near half.)
#define GT_SLOTS 6
#define GT_THREADS 320
#define GT_OFF_BAR (GT_SLOTS * 2 * GT_TILE_BYTES)
#define GT_NBAR (3 * GT_SLOTS + 4)
#define GT_SMEM_TOTAL (GT_OFF_BAR + GT_NBAR * 8 + 16)

struct gram_tc_params {
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
    // barriers: raw_full[S] op_full[S] op_empty[S] tfull[2] tempty[2]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + GT_NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_rfull = smem_u32(s_bar), bar_ofull = bar_rfull + 8 * GT_SLOTS, bar_oempty = bar_ofull + 8 * GT_SLOTS;
    const uint32_t bar_tfull = bar_oempty + 8 * GT_SLOTS, bar_tempty = bar_tfull + 16;

    if (threadIdx.x == 0) {
        if (smem_base & 1023u) { printf("ssi_gram_tc: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < GT_SLOTS; ++s) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_ofull + 8 * s, 32); mbar_init(bar_oempty + 8 * s, 1); }
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
                mbar_wait(bar_oempty + 8 * stage, phase ^ 1);          // the MMAs that read this slot have retired
                mbar_expect_tx(bar_rfull + 8 * stage, GT_TILE_BYTES);
                // A is streamed exactly once: do not let it evict the FP64 partials from L2.  (Tried and dropped, no
                // gain: issuing several row blocks back to back, or by 32-column groups with the row blocks innermost.)
                tma_load_2d_hint(smem_base + stage * 2 * GT_TILE_BYTES, &tmA, bar_rfull + 8 * stage, (int)(t * GT_ROWS), 0, TC_EVICT_FIRST);
                if (++stage == GT_SLOTS) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // one MMA per k-step: the B operand spans the H tile (128 rows; rows >= K are TMA zero fill) and the first NP
            // rows of the L tile behind it, so D = H [H ; L]^T = [HH (columns 0..127) | HL (columns 128..128+NP)] and the
            // A operand (H) is fetched from shared memory once instead of twice
            const uint32_t idesc = umma_idesc_tf32(128 + p.NP);
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
                    const uint64_t dh = umma_desc_sw128(smem_base + stage * 2 * GT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < GT_ROWS / 8; ++k) {
                        const uint64_t ko = (uint64_t)(k * 32 >> 4);    // 8 TF32 = 32 bytes per k-step
                        umma_tf32(d_acc, dh + ko, dh + ko, idesc, !(first && k == 0));
                    }
                    umma_commit(bar_oempty + 8 * stage);
                    if (++stage == GT_SLOTS) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * ab);
            }
        }
    } else if (warp < 6) {
        // ================= converters: split the FP32 tile into TF32 hi / lo, one warp per tile =================
        const int cw = warp - 2;                  // 0..3: this warp takes tiles cw, cw + 4, ...
        for (long long t = cw; t < my_tiles; t += 4) {
            const int slot = (int)(t % GT_SLOTS);
            const uint32_t ph = (uint32_t)((t / GT_SLOTS) & 1);
            mbar_wait(bar_rfull + 8 * slot, ph);
            float4* H = reinterpret_cast<float4*>(smem + (size_t)slot * 2 * GT_TILE_BYTES);
            float4* L = H + GT_TILE_BYTES / 16;
            // the split is element-wise at identical offsets, so it is independent of the 128-byte swizzle
#pragma unroll 1
            for (int r0 = 0; r0 < GT_TILE_BYTES / 16 / 32; r0 += 8) {
                float4 x[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) x[r] = H[lane + (r0 + r) * 32];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(x[r].x) & 0xFFFFE000u);
                    h.y = __uint_as_float(__float_as_uint(x[r].y) & 0xFFFFE000u);
                    h.z = __uint_as_float(__float_as_uint(x[r].z) & 0xFFFFE000u);
                    h.w = __uint_as_float(__float_as_uint(x[r].w) & 0xFFFFE000u);
                    l.x = __uint_as_float(__float_as_uint(x[r].x - h.x) & 0xFFFFE000u);
                    l.y = __uint_as_float(__float_as_uint(x[r].y - h.y) & 0xFFFFE000u);
                    l.z = __uint_as_float(__float_as_uint(x[r].z - h.z) & 0xFFFFE000u);
                    l.w = __uint_as_float(__float_as_uint(x[r].w - h.w) & 0xFFFFE000u);
                    H[lane + (r0 + r) * 32] = h;
                    L[lane + (r0 + r) * 32] = l;
                }
            }
            fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bar_ofull + 8 * slot);
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
                        for (int j = 0; j < 16; ++j) st_f64_hint(dst + j * 128, (double)__uint_as_float(v[j]), pol);
                    } else {
                        double old[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) old[j] = ld_f64_hint(dst + j * 128, pol);
#pragma unroll
                        for (int j = 0; j < 16; ++j) st_f64_hint(dst + j * 128, old[j] + (double)__uint_as_float(v[j]), pol);
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
__global__ void k_gram_tc_sym(const double* __restrict__ S, int K, int NP, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * K) return;
    const int a = e % K, b = e / K;
    const int lo = min(a, b), hi = max(a, b);
    G[a + (long long)b * K] = S[(long long)hi * 128 + lo] + S[((long long)NP + b) * 128 + a] + S[((long long)NP + a) * 128 + b];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool ssi_gram_tc_usable(const ssi_ctx* ctx, const float* dA, int64_t n, int64_t ld, int K) {
    return K >= 1 && K <= 128 && (ld % 4) == 0 && n >= (1 << 16) && ((reinterpret_cast<uintptr_t>(dA) & 15) == 0) &&
           n < (1ll << 31) && ctx->opt_gram_fp64 <= 0;
}

int ssi_gram_tc_device(ssi_ctx* ctx, const float* dA, int64_t n, int64_t ld, int K, double* dG) {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SSI_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return ssi_fail(ctx, SSI_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (PFN_encodeTiled)fn;
    }
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
    k_gram_tc_sym<<<(K * K + 255) / 256, 256, 0, ctx->stream>>>(dS, K, p.NP, dG);
    SSI_LAUNCH_CHECK(ctx);
    return SSI_OK;
}
