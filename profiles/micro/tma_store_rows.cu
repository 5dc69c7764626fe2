// Micro-benchmark: TMA store throughput as a function of the contiguous bytes per (scattered) row.
//   mode 0: 2-D box {64 bf16, 32 rows}            rows of 128 B, row stride = sample stride (what k_tc_basis_mma stores today)
//   mode 1: 3-D box {64 bf16, 2, 32 rows}         two adjacent 128-B segments per row = 256 B contiguous per sample row
//   mode 2: 3-D box {64 bf16, 4, 32 rows}         512 B contiguous per sample row
// Global layout: [samples 128][rows R][seg S][64] bf16, a CTA walks (row r, sample group of 32) tiles; bytes per store = 32*S*128.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_rows tma_store_rows.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DIMS>
__global__ void __launch_bounds__(128, 1) k_store(const __grid_constant__ CUtensorMap tm, int rows, int seg_groups, int iters_unused) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // fill staging once (content irrelevant)
    for (int i = threadIdx.x; i < 4 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    // tiles: (row r, seg group sg, sample group g of 4): each is one TMA store of a [32 samples] box
    const long long total = (long long)rows * seg_groups * 4;
    int buf = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int g = (int)(t & 3);
        const long long rest = t >> 2;
        const int sg = (int)(rest % seg_groups);
        const int r = (int)(rest / seg_groups);
        const uint32_t src = smem_u32(smem) + buf * 16384;
        if (DIMS == 3) {
            // coordinates: {elem 0, row index (r*seg_total + sg*S) folded by the map as dim1 = segments within the row block, dim2 = sample}
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         :: "l"(&tm), "r"(0), "r"(r * seg_groups + sg), "r"(g * 32), "r"(src) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                         :: "l"(&tm), "r"(0), "r"(0), "r"(r * seg_groups + sg), "r"(g * 32), "r"(src) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        buf = (buf + 1) & 3;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int samples = 128, rows = 4096, segs = 16;        // per sample: rows x segs x 128 B = 8 MB; total 1 GB
    const size_t bytes = (size_t)samples * rows * segs * 128;
    void* buf;
    CHECK(cudaMalloc(&buf, bytes));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encodeTiled encode = (PFN_encodeTiled)fn;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int S = 1; S <= 4; S *= 2) {
        // view: dim0 = 64 elems, dim1 = S adjacent segments, dim2 = (rows*segs/S) segment groups, dim3 = samples
        CUtensorMap tm;
        CUresult r;
        if (S == 1) {
            cuuint64_t dims[3] = {64, (cuuint64_t)rows * segs, (cuuint64_t)samples};
            cuuint64_t strides[2] = {128, (cuuint64_t)rows * segs * 128};
            cuuint32_t box[3] = {64, 1, 32};
            cuuint32_t es[3] = {1, 1, 1};
            r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[4] = {64, (cuuint64_t)S, (cuuint64_t)rows * segs / S, (cuuint64_t)samples};
            cuuint64_t strides[3] = {128, (cuuint64_t)S * 128, (cuuint64_t)rows * segs * 128};
            cuuint32_t box[4] = {64, (cuuint32_t)S, 1, 32};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode failed S=%d: %d\n", S, (int)r); continue; }
        const int seg_groups_total = segs / S;     // per row
        for (int rep = 0; rep < 3; ++rep) {
            CHECK(cudaEventRecord(e0));
            if (S == 1) {
                CHECK(cudaFuncSetAttribute(k_store<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
                k_store<3><<<148, 128, 66560>>>(tm, rows, seg_groups_total, 0);
            } else {
                CHECK(cudaFuncSetAttribute(k_store<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
                k_store<4><<<148, 128, 66560>>>(tm, rows, seg_groups_total, 0);
            }
            CHECK(cudaEventRecord(e1));
            CHECK(cudaEventSynchronize(e1));
            CHECK(cudaGetLastError());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("contiguous %4d B per sample row: %.3f ms  %.2f TB/s\n", S * 128, ms, bytes / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
