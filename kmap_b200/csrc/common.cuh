// Shared device helpers for libkmap_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kmap_b200.h"

#define KMAP_PAD_WORDS 4          // zero padding (valid words) after the last position
#define KMAP_EMPTY_SLOT 0xFFFFFFFFu

void kmap_set_error(const char* fmt, ...);
int kmap_check_launch(const char* what);

#define KMAP_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) { kmap_set_error("%s: %s", __func__, msg); return KMAP_ERR_BAD_ARG; } \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned int grid_for(int64_t n_threads, int block) {
    int64_t g = (n_threads + block - 1) / block;
    return (unsigned int)(g < 1 ? 1 : g);
}

// the exchange step of a sharded count (comm.cu): NCCL communicator + the stream the all-reduces are issued on
// scatter: instead of leaving every merged table on every rank (all-reduce), rank r of `world` ends up owning cells
// [r * n / world, (r + 1) * n / world) of every table (reduce-scatter by key range; the other cells hold partial sums):
// half the exchange volume, and everything behind the merge (reductions to the lower levels, compaction) runs on 1 / world
// of the cells.  Needs world | 4^kmin.
struct KmapMerge { void* comm; cudaStream_t stream; int scatter; int rank; int world; };
#define KMAP_COMM_CTAS 16         // CTAs the collective may use = SMs the counting kernels leave free while it runs
// Measured (profiles/r02_merge_sweep_8gpu.txt): with 2 ranks the exchange hides behind the per-bucket count and 16 CTAs are
// enough; from 4 ranks on a rank's count is shorter than the exchange, which then wants 32 CTAs and fewer, larger calls.
int kmap_comm_ctas(int world);    // ... KMAP_COMM_CTAS for < 4 ranks, 32 from 4 ranks on, unless the environment says otherwise (KMAP_COMM_CTAS)
int kmap_merge_chunks(int world); // key ranges the level-kmax table is merged in while it is counted: 4 (< 4 ranks) or 2 (KMAP_MERGE_CHUNKS)
int kmap_comm_world(void* comm);  // number of ranks of a communicator (1 on failure)
int kmap_allreduce_u32_on(uint32_t* buf, int64_t n, void* comm, cudaStream_t s);
// merge one table the way `m` says: all-reduce, or reduce-scatter in place (rank r's block stays where it is in `buf`)
int kmap_merge_table_on(uint32_t* buf, int64_t n, const KmapMerge* m);
// would that merge take the peer-memory exchange (peer.cu: the buffer lies in the table area of the rank's peer region)?
bool kmap_merge_on_peer_memory(const uint32_t* buf, int64_t n, const KmapMerge* m);

// dense tables of several levels: t[k] = uint32[4^k] (only the levels a kernel uses are set); passed by value
struct KmapTableSet { uint32_t* t[16]; };

// ---- 2-bit arithmetic -----------------------------------------------------------------------------------
// mask of the low 2k bits
__host__ __device__ __forceinline__ uint32_t lowmask32(int k) { return k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u); }
__host__ __device__ __forceinline__ uint64_t lowmask64(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// number of non-zero 2-bit groups of x among the groups selected by `low` (taichi_core.py:63-72 as one popcount)
__device__ __forceinline__ uint32_t nz_groups32(uint32_t x, uint32_t low) {
    return __popc((x | (x >> 1)) & 0x55555555u & low);
}
__device__ __forceinline__ uint32_t nz_groups64(uint64_t x, uint64_t low) {
    return __popcll((x | (x >> 1)) & 0x5555555555555555ull & low);
}

// reverse complement of a k-mer hash (taichi_core.py:181-206): complement, then reverse the 2-bit groups
__device__ __forceinline__ uint32_t revcom32(uint32_t h, int k) {
    uint32_t r = __brev(~h);                                   // bit reversal also swaps the bits inside a group
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);   // ... swap them back
    return r >> (32 - 2 * k);
}
__device__ __forceinline__ uint64_t revcom64(uint64_t h, int k) {
    uint64_t r = __brevll(~h);
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return r >> (64 - 2 * k);
}

// ---- packed-sequence windows ----------------------------------------------------------------------------
// 16 bases starting at position p, first base in the most significant bits
__device__ __forceinline__ uint32_t window16(const uint32_t* __restrict__ packed, int64_t p) {
    const int64_t w = p >> 4;
    const uint32_t s = ((uint32_t)p & 15u) * 2u;
    return __funnelshift_l(__ldg(packed + w + 1), __ldg(packed + w), s);
}
// hash of the k-mer (k <= 16) starting at p
__device__ __forceinline__ uint32_t window_hash(const uint32_t* __restrict__ packed, int64_t p, int k) {
    return window16(packed, p) >> (32 - 2 * k);
}
// 32 bases starting at position p, first base in the most significant bits (windows of 17..31 bases: 64-bit hashes)
__device__ __forceinline__ uint64_t window32(const uint32_t* __restrict__ packed, int64_t p) {
    const int64_t w = p >> 4;
    const uint32_t s = ((uint32_t)p & 15u) * 2u;
    const uint32_t a = __ldg(packed + w), b = __ldg(packed + w + 1), c = __ldg(packed + w + 2);
    return ((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s);
}
// validity bits of positions p .. p+31 (bit i = position p+i)
__device__ __forceinline__ uint32_t valid32(const uint32_t* __restrict__ valid, int64_t p) {
    const int64_t w = p >> 5;
    return __funnelshift_r(__ldg(valid + w), __ldg(valid + w + 1), (uint32_t)p & 31u);
}
__device__ __forceinline__ bool window_ok(const uint32_t* __restrict__ valid, int64_t p, int k) {
    const uint32_t km = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    return (valid32(valid, p) & km) == km;
}

// cheap integer mixer for the per-read hash sets
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; x *= 0x297a2d39u; x ^= x >> 15;
    return x;
}

// splitmix64 finaliser; the synthetic generator is counter-based on it (kmap_b200/synth.py is the numpy twin)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
