// Microbenchmark: issue rate of tcgen05.mma (M = 128, K = 16, bf16, both operands in shared memory, no-swizzle K-major
// with 8-row groups packed back to back) as a function of N, the leading-dimension byte offset of A and the row shift of
// the A window. Answers: is the shifted-window convolution bound by instruction issue or by operand fetch?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I q-learning_b200/csrc -o tools/microbench/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include "qnet_conv.cuh"
using namespace qlc;
using namespace qlc::qnet;

template <int N, int STYLE, int M = 128>
__global__ void __launch_bounds__(128, 1) rate_kernel(uint32_t lbo_a, uint32_t shift_rows, int iters, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t tmem_u = __reduce_max_sync(0xFFFFFFFFu, tmem_base);
        constexpr uint32_t IDESC = instr_desc_bf16(M, N);
        const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((lbo_a >> 4) << 16);
        const uint32_t b_lo0 = ((smem_u32(smem) + 96 * 1024) >> 4) | ((uint32_t)(N * 16 >> 4) << 16);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (STYLE == 0) {
                #pragma unroll
                for (int j = 0; j < 16; ++j)
                    tc_mma_bf16_elect_lo(tmem_u + (uint32_t)((j & 3) * N) % 256u, a_lo0 + shift_rows * (uint32_t)(j & 3) + (uint32_t)(j >> 2) * 2u * (lbo_a >> 4), b_lo0 + (uint32_t)((j & 3) * 2 * N), IDESC, 1u);
            } else {
                if (threadIdx.x == 0) {
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t a_lo = a_lo0 + shift_rows * (uint32_t)(j & 3) + (uint32_t)(j >> 2) * 2u * (lbo_a >> 4), b_lo = b_lo0 + (uint32_t)((j & 3) * 2 * N);
                        tc_mma_bf16(tmem_u + (uint32_t)((j & 3) * N) % 256u, ((uint64_t)0x4008 << 32) | a_lo, ((uint64_t)0x4008 << 32) | b_lo, IDESC, 1u);
                    }
                }
                __syncwarp();
            }
        }
        const long long t1 = clock64();
        tc_commit_elect(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

template <int N, int STYLE, int M = 128>
static void run(uint32_t lbo, uint32_t shift, const char* what) {
    long long* d; cudaMalloc(&d, 16);
    auto k = rate_kernel<N, STYLE, M>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 200;
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 160 * 1024>>>(lbo, shift, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d style=%d lbo=%5u shift=%2u : issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  %s %s\n", M, N, STYLE, lbo, shift, (double)h[0] / (iters * 16), (double)h[1] / (iters * 16), what,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

// TMEM read rate: NW warps (one per lane quadrant) each issue tcgen05.ld 32x32b.x32 (4 KB per instruction) back to back
__global__ void __launch_bounds__(128, 1) tmem_read_kernel(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    uint32_t accs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t0 = clock64();
    if ((int)warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int c = 0; c < 512; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((warp * 32u) << 16) + (uint32_t)c, v);
                #pragma unroll
                for (int i = 0; i < 32; ++i) accs[i & 7] ^= v[i];
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0; for (int i = 0; i < 8; ++i) acc ^= accs[i];
    if (acc == 0x12345678u) sink[tid] = (float)acc;
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}
static void run_tmem(int nwarps) {
    long long* d; float* sink; cudaMalloc(&d, 16); cudaMalloc(&sink, 4096);
    const int iters = 50;
    for (int rep = 0; rep < 2; ++rep) tmem_read_kernel<<<148, 128>>>(nwarps, iters, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 16);
    printf("tmem read, %d warps: %6.1f cyc per 32x32b.x32 ld (+32 LOP) per warp -> %.0f B/clk per SM  %s\n", nwarps, per, nwarps * 4096.0 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}

int main() {
    run_tmem(1); run_tmem(2); run_tmem(4);
    run<32, 0>(7056, 1, "conv1-like"); run<32, 0>(7168, 0, "aligned"); run<32, 0>(7168, 8, "8-row shift"); run<32, 1>(7056, 1, "conv1-like, single thread");
    run<64, 0>(3200, 1, "conv2-like"); run<64, 0>(3200, 0, "aligned"); run<64, 0>(3888, 1, "conv3-like"); run<64, 1>(3200, 1, "single thread");
    run<128, 0>(3200, 1, ""); run<128, 0>(3200, 0, "aligned"); run<256, 0>(3200, 1, ""); run<256, 0>(3200, 0, "aligned");
    run<16, 0>(3200, 1, ""); run<8, 0>(3200, 0, "aligned");
    // role swap (weights as A with M = cout = 64, pixels as B with N = 256 / 128): does an M = 64 MMA cost half an M = 128 one?
    run<256, 0, 64>(1024, 0, "M=64: weights as A, 256 pixels as B"); run<128, 0, 64>(1024, 0, "M=64, N=128"); run<64, 0, 64>(1024, 0, "M=64, N=64");
    return 0;
}
