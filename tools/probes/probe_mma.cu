// Micro-probe (dev tool, not part of the library): issue-to-completion cost of back-to-back tcgen05.mma (kind::f16,
// K = 16, both operands K-major SWIZZLE_128B tiles in shared memory) as a function of M, N, the number of
// accumulators rotated over and how the K slices walk through shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ../lib/probe_mma probe_mma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_tc.cuh"
using namespace april::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// walk: 0 = same K slice every time; 1 = 4 slices of one k-block, then the next k-block (16 A k-blocks, 2 B k-blocks)
__global__ void __launch_bounds__(128, 1) probe(int M, int N, int nacc, int walk, int iters, int nwarps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a0 = base, b0 = base + 16 * 8192, bar = b0 + 2 * 32768, slot = bar + 8 + 64;   // barriers at bar + 16 w (slot sits between them)
    for (uint32_t i = threadIdx.x; i < (16 * 8192 + 2 * 32768) / 2; i += blockDim.x) reinterpret_cast<__half*>(bp)[i] = __float2half((float)((i * 2654435761u) >> 24) * (1.f / 256.f) - 0.5f);
    if (threadIdx.x == 0) { for (int w = 0; w < 4; w++) mbar_init(bar + 16 * w, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(bp + (slot - base));
    const uint32_t idesc = make_idesc_f16(M, N);
    const int wid = threadIdx.x >> 5;
    if (wid < nwarps) {
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const int kbA = walk ? (i & 15) : 0, kbB = walk ? (i & 1) : 0;
            const uint64_t ad = make_sw128_desc(a0 + kbA * 8192), bd = make_sw128_desc(b0 + kbB * 32768);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; k++) umma_f16(tmem + (wid * nacc + (i * 4 + k) % nacc) * N, ad + (walk ? 2 * k : 0), bd + (walk ? 2 * k : 0), idesc, 1);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(bar + 16 * wid);
        __syncwarp();
        mbar_wait(bar + 16 * wid, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }   // warp 0's view; all warps run the same count
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const size_t smem = 1024 + 16 * 8192 + 2 * 32768 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* out; CK(cudaMalloc(&out, 148 * 2 * sizeof(long long)));
    const int iters = 512;
    printf("grid  M    N  nacc walk | issue clk/mma   total clk/mma   (floor N/2 @M=128)\n");
    for (int grid : { 1 })
        for (int walk : { 1 })
            for (int M : { 64, 128 })
                for (int N : { 32, 64, 128, 256 })
                    for (int nwarps : { 1, 2, 4 }) {
                        const int nacc = 1;
                        if (nwarps * nacc * N > 512) continue;
                        probe<<<grid, 128, smem>>>(M, N, nacc, walk, iters, nwarps, out);
                        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
                        probe<<<grid, 128, smem>>>(M, N, nacc, walk, iters, nwarps, out);
                        CK(cudaDeviceSynchronize());
                        std::vector<long long> h((size_t)grid * 2);
                        CK(cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                        double a = 0, b = 0; for (int c = 0; c < grid; c++) { a += h[c * 2]; b += h[c * 2 + 1]; }
                        printf("%4d %4d %4d %4d %4d | %10.1f %14.1f   (per warp; %d warps issue concurrently)\n", grid, M, N, nacc, walk, a / grid / (iters * 4), b / grid / (iters * 4), nwarps);
                    }
    return 0;
}
