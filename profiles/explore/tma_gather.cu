// Explore: how fast can TMA gather {R rows x 100 columns} boxes of a column-major n x 100 FP32 matrix (the access pattern of
// the Gram and P kernels), as a function of the contiguous segment R*4 bytes and of the bytes in flight per SM?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather tma_gather.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(a), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(x), "r"(y) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(128, 1) k_gather(const __grid_constant__ CUtensorMap tm, long long tiles_total, int rows, int tile_bytes, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long per = (tiles_total + gridDim.x - 1) / gridDim.x;
    const long long t0 = blockIdx.x * per, t1 = min(tiles_total, t0 + per);
    float acc = 0.f;
    if (threadIdx.x == 0) {
        // prologue
        long long issued = t0;
        for (int s = 0; s < STAGES && issued < t1; ++s, ++issued) {
            mbar_expect_tx(smem_u32(&bar[s]), tile_bytes);
            tma_load_2d(smem_u32(smem) + s * tile_bytes, &tm, smem_u32(&bar[s]), (int)(issued * rows), 0);
        }
        int s = 0; uint32_t ph = 0;
        for (long long t = t0; t < t1; ++t) {
            mbar_wait(smem_u32(&bar[s]), ph);
            acc += *reinterpret_cast<volatile float*>(smem + s * tile_bytes);
            if (issued < t1) {
                mbar_expect_tx(smem_u32(&bar[s]), tile_bytes);
                tma_load_2d(smem_u32(smem) + s * tile_bytes, &tm, smem_u32(&bar[s]), (int)(issued * rows), 0);
                ++issued;
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (acc == 12345.f) *sink = acc;
    }
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int STAGES>
void run(PFN_encode enc, float* dA, long long n, int K, int rows, float* sink) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    cuuint32_t box[2] = {(cuuint32_t)rows, (cuuint32_t)K};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     rows == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int tile_bytes = rows * K * 4;
    const int smem = STAGES * tile_bytes + 1024;
    if (smem > 227 * 1024) return;
    cudaFuncSetAttribute(k_gather<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const long long tiles = n / rows;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        k_gather<STAGES><<<148, 128, smem>>>(map, tiles, rows, tile_bytes, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("rows %4d (%5d B segments) stages %2d in-flight %6.1f KB/SM : %.3f ms  %.0f GB/s %s\n", rows, rows * 4, STAGES, STAGES * tile_bytes / 1024.0,
           best, (double)tiles * tile_bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const long long n = 10020864; const int K = 100;
    float* dA; cudaMalloc(&dA, n * K * 4); cudaMemset(dA, 0, n * K * 4);
    float* sink; cudaMalloc(&sink, 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encode enc = (PFN_encode)fn;
    for (int rows : {32, 64, 128, 256}) {
        run<2>(enc, dA, n, K, rows, sink);
        run<3>(enc, dA, n, K, rows, sink);
        run<4>(enc, dA, n, K, rows, sink);
        run<6>(enc, dA, n, K, rows, sink);
        run<8>(enc, dA, n, K, rows, sink);
        run<12>(enc, dA, n, K, rows, sink);
        run<16>(enc, dA, n, K, rows, sink);
    }
    return 0;
}
