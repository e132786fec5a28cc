// Hamming-ball count aggregation of find_motif (motif_discovery.py:666-673).
//
// The reference scans the merged (kh, cnt) list once per candidate and strand.  On the dense forward table F the
// same number is   sum over y in B(c,d) u B(rc c,d) of F[y] * (1 + [y == rc y])   (merged entries are
// (min(h, rc h), F[h] + F[rc h]), palindromes doubled, and the union of the two balls is closed under reverse
// complement -- DESIGN.md section 4.4), so only |B| <= 2 * 578 257 cells (k=14, d=5) are gathered per candidate
// instead of 8 B x 1.3e8 list entries.
#include "common.cuh"

namespace {

constexpr int HB_BLOCK = 256;
constexpr int HB_MAXK = 16;

struct BallShape {
    int k, d;
    unsigned long long start[HB_MAXK + 2];       // start[j] = #members with fewer than j substitutions
    unsigned long long binom[HB_MAXK + 1][HB_MAXK + 1];
    unsigned int pow3[HB_MAXK + 1];
};

// member number t of B(c, d): pick j = #substituted positions, unrank the position set (lexicographic
// combinatorial number system) and the base-3 substitution digits
__device__ __forceinline__ uint32_t ball_member_at(const BallShape& s, uint32_t c, unsigned long long t) {
    int j = 0;
    while (j < s.d && t >= s.start[j + 1]) ++j;
    t -= s.start[j];
    unsigned long long ci = t / s.pow3[j];
    unsigned int si = (unsigned int)(t % s.pow3[j]);
    uint32_t y = c;
    int x = 0;
    for (int i = 0; i < j; ++i) {
        // choose position x (0 = first base) such that the remaining j-1-i positions fit behind it
        while (true) {
            const unsigned long long cnt = s.binom[s.k - 1 - x][j - 1 - i];
            if (ci < cnt) break;
            ci -= cnt; ++x;
        }
        const unsigned int digit = si % 3u; si /= 3u;
        const int shift = 2 * (s.k - 1 - x);
        const uint32_t b = (c >> shift) & 3u;
        y = (y & ~(3u << shift)) | (((b + 1u + digit) & 3u) << shift);
        ++x;
    }
    return y;
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v) {
    __shared__ unsigned long long ws[HB_BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < HB_BLOCK / 32 ? ws[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xFFFFFFFFu, t, o);
    }
    __syncthreads();
    return t;   // valid in thread 0
}

__global__ void __launch_bounds__(HB_BLOCK) hamball_sum_kernel(const uint32_t* __restrict__ F, BallShape shape,
                                                               const uint32_t* __restrict__ cand, int revcom,
                                                               unsigned long long* __restrict__ sums) {
    const int k = shape.k, d = shape.d;
    const uint32_t low = lowmask32(k);
    const uint32_t c = __ldg(cand + blockIdx.y) & low;
    const uint32_t centre = blockIdx.z == 0 ? c : revcom32(c, k);
    const unsigned long long total = shape.start[d + 1];
    unsigned long long acc = 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * HB_BLOCK + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * HB_BLOCK) {
        const uint32_t y = ball_member_at(shape, centre, t);
        if (blockIdx.z == 1 && (int)nz_groups32(y ^ c, low) <= d) continue;   // already counted in the forward ball
        unsigned long long f = __ldg(F + y);
        if (revcom && f && revcom32(y, k) == y) f *= 2;                       // palindromes are doubled by merge_revcom
        acc += f;
    }
    acc = block_sum_u64(acc);
    if (threadIdx.x == 0 && acc) atomicAdd(sums + blockIdx.y, acc);
}

constexpr int HBL_MAXM = 16;
struct CandList { uint32_t c[HBL_MAXM]; uint32_t rc[HBL_MAXM]; };

// the reference formulation: one pass over the merged list for all candidates
__global__ void __launch_bounds__(HB_BLOCK) hamball_sum_list_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ cnt,
                                                                    int64_t n, int k, const uint32_t* __restrict__ cand, int m, int d,
                                                                    int revcom, unsigned long long* __restrict__ sums) {
    __shared__ uint32_t sc[HBL_MAXM], src[HBL_MAXM];
    const uint32_t low = lowmask32(k);
    if (threadIdx.x < m) { const uint32_t c = cand[threadIdx.x] & low; sc[threadIdx.x] = c; src[threadIdx.x] = revcom32(c, k); }
    __syncthreads();
    unsigned long long acc[HBL_MAXM];
#pragma unroll
    for (int i = 0; i < HBL_MAXM; ++i) acc[i] = 0;
    for (int64_t j = (int64_t)blockIdx.x * HB_BLOCK + threadIdx.x; j < n; j += (int64_t)gridDim.x * HB_BLOCK) {
        const uint32_t h = __ldg(kh + j);
        const unsigned long long w = (unsigned long long)(long long)__ldg(cnt + j);
#pragma unroll
        for (int i = 0; i < HBL_MAXM; ++i) {
            if (i < m) {
                uint32_t dist = nz_groups32(h ^ sc[i], low);
                if (revcom) { const uint32_t rd = nz_groups32(h ^ src[i], low); dist = rd < dist ? rd : dist; }
                if ((int)dist <= d) acc[i] += w;
            }
        }
    }
    for (int i = 0; i < m; ++i) {
        const unsigned long long t = block_sum_u64(acc[i]);
        if (threadIdx.x == 0 && t) atomicAdd(sums + i, t);
    }
}

}  // namespace

extern "C" {

int kmap_hamball_sum(const uint32_t* table, int k, const uint32_t* cand, int m, int d, int revcom, uint64_t* sums, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 15 && m >= 0 && d >= 0, "bad argument");
    if (m == 0) return KMAP_OK;
    KMAP_REQUIRE(table && cand && sums, "null pointer");
    if (d > k) d = k;
    cudaStream_t s = as_stream(stream);
    BallShape sh;
    sh.k = k; sh.d = d;
    for (int n = 0; n <= HB_MAXK; ++n)
        for (int r = 0; r <= HB_MAXK; ++r)
            sh.binom[n][r] = r == 0 ? 1 : (n == 0 ? 0 : sh.binom[n - 1][r - 1] + (r <= n - 1 ? sh.binom[n - 1][r] : 0));
    sh.pow3[0] = 1;
    for (int j = 1; j <= HB_MAXK; ++j) sh.pow3[j] = sh.pow3[j - 1] * 3u;
    sh.start[0] = 0;
    for (int j = 0; j <= d; ++j) sh.start[j + 1] = sh.start[j] + sh.binom[k][j] * sh.pow3[j];
    for (int j = d + 1; j <= HB_MAXK; ++j) sh.start[j + 1] = sh.start[d + 1];
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)m * 8, s);
    if (e != cudaSuccess) { kmap_set_error("hamball_sum: %s", cudaGetErrorString(e)); return (int)e; }
    unsigned long long blocks = (sh.start[d + 1] + HB_BLOCK - 1) / HB_BLOCK;
    if (blocks > 148 * 4) blocks = 148 * 4;
    dim3 grid((unsigned int)blocks, (unsigned int)m, revcom ? 2 : 1);
    hamball_sum_kernel<<<grid, HB_BLOCK, 0, s>>>(table, sh, cand, revcom, reinterpret_cast<unsigned long long*>(sums));
    return kmap_check_launch("hamball_sum");
}

int kmap_hamball_sum_list(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, const uint32_t* cand, int m, int d,
                          int revcom, uint64_t* sums, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 15 && m >= 0 && m <= HBL_MAXM && d >= 0 && n >= 0, "bad argument (m <= 16)");
    if (m == 0) return KMAP_OK;
    KMAP_REQUIRE(cand && sums, "null pointer");
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)m * 8, s);
    if (e != cudaSuccess) { kmap_set_error("hamball_sum_list: %s", cudaGetErrorString(e)); return (int)e; }
    if (n == 0) return KMAP_OK;
    int64_t blocks = (n + HB_BLOCK - 1) / HB_BLOCK;
    if (blocks > 148 * 8) blocks = 148 * 8;
    hamball_sum_list_kernel<<<(unsigned int)blocks, HB_BLOCK, 0, s>>>(kh, cnt, n, k, cand, m, d, revcom,
                                                                      reinterpret_cast<unsigned long long*>(sums));
    return kmap_check_launch("hamball_sum_list");
}

}  // extern "C"
