// Micro-probe (dev tool, not part of the library): how fast can one thread feed shared memory with TMA?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ../lib/probe_tma probe_tma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "kernels_tc.cuh"
using namespace april::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// mode 1: one barrier for all bytes; lanes of warp 0 issue ops in parallel (lane i issues ops i, i+32, ...)
__global__ void __launch_bounds__(128, 1) probe_par(const __grid_constant__ CUtensorMap map, int box_rows, int nops, int same_data, int krange, int lanes, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 200 * 1024;
    const uint32_t bytes = box_rows * 128;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int slots = (200 * 1024) / bytes;
        long long t0 = clock64();
        if (threadIdx.x == 0) mbar_expect_tx(bar, bytes * nops);
        __syncwarp();
        int row = same_data ? 0 : (int)blockIdx.x * box_rows;
        for (int i = threadIdx.x; i < nops; i += lanes) if ((int)threadIdx.x < lanes) tma_load_2d(base + (i % slots) * bytes, &map, bar, (i % krange) * 64, row);
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; out[blockIdx.x * 4 + 2] = t2 - t0; }
    }
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, int box_rows, int nops, int same_data, int krange, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + 200 * 1024;
    if (threadIdx.x == 0) { for (int i = 0; i < nops; i++) mbar_init(bars + 8 * i, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map);
        const uint32_t bytes = box_rows * 128;
        const int slots = (200 * 1024) / bytes;
        long long t0 = clock64();
        for (int i = 0; i < nops; i++) {
            mbar_expect_tx(bars + 8 * i, bytes);
            int row = same_data ? 0 : (int)blockIdx.x * box_rows;
            tma_load_2d(base + (i % slots) * bytes, &map, bars + 8 * i, (i % krange) * 64, row);
        }
        long long t1 = clock64();
        mbar_wait(bars + 0, 0);
        long long t2 = clock64();
        for (int i = 1; i < nops; i++) mbar_wait(bars + 8 * i, 0);
        long long t3 = clock64();
        out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; out[blockIdx.x * 4 + 2] = t3 - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const uint64_t rows = 148 * 256, cols = 4096;
    __half* d; CK(cudaMalloc(&d, rows * cols * 2)); CK(cudaMemset(d, 0, rows * cols * 2));
    long long* out; CK(cudaMalloc(&out, 148 * 4 * 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 2048));
    CK(cudaFuncSetAttribute(probe_par, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 2048));
    for (int promo = 1; promo < 2; promo++)
    for (int box_rows : { 16, 64, 128, 256 }) {
        CUtensorMap m; cuuint64_t dims[2] = { cols, rows }; cuuint64_t strides[1] = { cols * 2 }; cuuint32_t box[2] = { 64, (cuuint32_t)box_rows }; cuuint32_t es[2] = { 1, 1 };
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        for (int grid : { 64 }) for (int same : { 1 }) for (int lanes : { 1, 2, 4, 8, 32 }) {
            int nops = 32;
            for (int rep = 0; rep < 3; rep++) { CK(cudaMemset(out, 0, 148 * 32)); probe_par<<<grid, 128, 200 * 1024 + 2048>>>(m, box_rows, nops, same, 32, lanes, out); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
            std::vector<long long> h(grid * 4); CK(cudaMemcpy(h.data(), out, grid * 32, cudaMemcpyDeviceToHost));
            double a = 0, c = 0, mx = 0; for (int i = 0; i < grid; i++) { a += h[i * 4]; c += h[i * 4 + 2]; mx = std::max(mx, (double)h[i * 4 + 2]); }
            a /= grid; c /= grid;
            printf("PAR box_rows=%3d grid=%3d same=%d lanes=%2d : issue %6.0f cyc  all-done avg %6.0f max %6.0f cyc  => %.1f B/clk/SM (avg)\n", box_rows, grid, same, lanes, a, c, mx, nops * box_rows * 128.0 / c);
        }
        for (int grid : { 1 }) for (int same : { 0 }) {
            int nops = 32;
            for (int rep = 0; rep < 3; rep++) { probe<<<grid, 128, 200 * 1024 + 2048>>>(m, box_rows, nops, same, 32, out); CK(cudaDeviceSynchronize()); }
            std::vector<long long> h(grid * 4); CK(cudaMemcpy(h.data(), out, grid * 32, cudaMemcpyDeviceToHost));
            double a = 0, b = 0, c = 0; for (int i = 0; i < grid; i++) { a += h[i * 4]; b += h[i * 4 + 1]; c += h[i * 4 + 2]; }
            a /= grid; b /= grid; c /= grid;
            printf("promo256=%d box_rows=%3d grid=%3d same=%d : issue %6.0f cyc (%.0f/op)  first-done %6.0f  all-done %6.0f cyc  => %.1f B/clk/SM\n", promo, box_rows, grid, same, a, a / nops, b, c, nops * box_rows * 128.0 / c);
        }
    }
    return 0;
}
