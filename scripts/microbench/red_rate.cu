// Microbenchmark (exploration, not product): how fast can a B200 increment random uint32 cells?
//   global RED into a table of S bytes (L2-resident or not), locality variants, shared-memory ATOMS.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_rate red_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode 0: fully random cell; 1: all lanes of a warp inside one 128 B line; 2: lanes inside one 32 B sector (8 cells);
// 3: consecutive lanes -> consecutive cells (coalesced); 4: random but each thread does 2 adjacent-cell ops
template <int MODE>
__global__ void red_kernel(uint32_t* table, uint32_t mask, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = tid >> 5;
    uint32_t s = mix(tid + 12345u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        uint32_t a;
        if (MODE == 0) { s = mix(s + i); a = s & mask; }
        else if (MODE == 1) { uint32_t w = mix(warp * 1315423911u + i); a = ((w << 5) | (mix(s + i) & 31u)) & mask; s = mix(s); }
        else if (MODE == 2) { uint32_t w = mix(warp * 1315423911u + i); a = ((w << 3) | (lane & 7u)) & mask; }
        else { uint32_t w = mix(warp * 1315423911u + i); a = ((w << 5) | lane) & mask; }
        atomicAdd(table + a, 1u);
    }
}

__global__ void smem_kernel(uint32_t* out, int iters, int cells_log2) {
    extern __shared__ uint32_t sm[];
    const uint32_t cells = 1u << cells_log2;
    for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 999u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) { s = mix(s + i); atomicAdd(&sm[s & (cells - 1)], 1u); }
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) acc += sm[i];
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}

// ATOMS whose return value is consumed (rank / overflow test)
__global__ void smem_ret_kernel(uint32_t* out, int iters, int cells_log2) {
    extern __shared__ uint32_t sm[];
    const uint32_t cells = 1u << cells_log2;
    for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 999u), acc = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) { s = mix(s + i); acc += atomicAdd(&sm[s & (cells - 1)], 1u); }
    __syncthreads();
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}

// plain (non-atomic) smem increments, one warp owns a private table: is the rate limit the atomic or the LSU?
__global__ void smem_plain_kernel(uint32_t* out, int iters, int cells_log2) {
    extern __shared__ uint32_t sm[];
    const uint32_t cells = 1u << cells_log2;
    for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 999u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) { s = mix(s + i); uint32_t a = s & (cells - 1); sm[a] = sm[a] + 1; }
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) acc += sm[i];
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}

template <typename F> float time_ms(F f, int reps = 3) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < reps; ++r) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
    uint32_t* table; cudaMalloc(&table, (size_t)1 << 30); cudaMemset(table, 0, (size_t)1 << 30);
    const int iters = 256;
    for (int blk : {256, 1024}) for (int bps : {4, 8}) {
        if (blk * bps > 2048) continue;
        const int grid = 148 * bps;
        const double ops = (double)grid * blk * iters;
        for (int lg = 22; lg <= 28; lg += 2) {       // cells: 4 Mi (16 MB) .. 256 Mi (1 GiB)
            const uint32_t mask = (1u << lg) - 1u;
            float m0 = time_ms([&] { red_kernel<0><<<grid, blk>>>(table, mask, iters); });
            float m1 = time_ms([&] { red_kernel<1><<<grid, blk>>>(table, mask, iters); });
            float m2 = time_ms([&] { red_kernel<2><<<grid, blk>>>(table, mask, iters); });
            float m3 = time_ms([&] { red_kernel<3><<<grid, blk>>>(table, mask, iters); });
            printf("global RED blk=%4d x%d/SM table=%5d MB : random %7.1f G/s | warp-in-128B-line %7.1f | warp-in-32B-sector %7.1f | coalesced %7.1f\n",
                   blk, bps, (4 << lg) >> 20, ops / m0 / 1e6, ops / m1 / 1e6, ops / m2 / 1e6, ops / m3 / 1e6);
        }
    }
    uint32_t* out; cudaMalloc(&out, 4);
    for (int lg : {13, 15}) for (int blk : {256, 1024}) {
        const size_t bytes = (size_t)4 << lg;
        cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        cudaFuncSetAttribute(smem_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        const int grid = 148 * (lg == 13 ? 4 : 1) * (blk == 256 && lg == 13 ? 1 : 1);
        const int it = 4096;
        const double ops = (double)grid * blk * it;
        float m = time_ms([&] { smem_kernel<<<grid, blk, bytes>>>(out, it, lg); });
        float p = time_ms([&] { smem_plain_kernel<<<grid, blk, bytes>>>(out, it, lg); });
        cudaFuncSetAttribute(smem_ret_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        float r = time_ms([&] { smem_ret_kernel<<<grid, blk, bytes>>>(out, it, lg); });
        printf("smem cells=2^%d blk=%4d grid=%d : ATOMS %7.1f G/s | ATOMS+return %7.1f G/s | plain LDS+STS %7.1f G/s\n", lg, blk, grid, ops / m / 1e6, ops / r / 1e6, ops / p / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
