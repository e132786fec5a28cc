// Integer parts of the consumers of the occurrence scan, computed from the scan results that are
// already in HBM instead of parsing final.motif_occurence.csv back:
//   get_motif_co_occurence_mat (motif_discovery.py:1189-1254)  reads per motif, reads shared by every pair of motifs, and for
//       every pair the per-read difference of the MEDIAN listed positions, in read order;
//   the `sum()` of find_motif (motif_discovery.py:648)          total of a count list.
// The float parts (np.median over the per-pair lists, the kernel density of get_motif_pos_density, :1256-1343) stay on the
// host, as the reference computes them.  A read with more than 20 listed positions in some cell is left to the host too:
// the reference keeps a RANDOM 20 of them (:1467-1469, numpy's global RNG), so its medians depend on that pick.
#include "common.cuh"

namespace {

constexpr int CO_MAX = 31;                         // motifs per call (bit 31 of the presence word flags a read for the host)
struct ScanPtrs { const int64_t* off[CO_MAX]; const int32_t* pos[CO_MAX]; };

// twice the median of the ascending positions pos[lo .. hi) (np.median: the middle one, or the mean of the middle two)
__device__ __forceinline__ int median2(const int32_t* __restrict__ pos, int64_t lo, int64_t hi) {
    const int64_t c = hi - lo, h = c >> 1;
    return (c & 1) ? 2 * __ldg(pos + lo + h) : __ldg(pos + lo + h - 1) + __ldg(pos + lo + h);
}

__global__ void __launch_bounds__(256) cooc_reads_kernel(ScanPtrs sp, int m, int64_t n_seq, uint32_t* __restrict__ present,
                                                         unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_cnt[CO_MAX * CO_MAX];
    for (int i = threadIdx.x; i < m * m; i += 256) s_cnt[i] = 0;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n_seq; r += (int64_t)gridDim.x * 256) {
        uint32_t mask = 0;
        bool over = false;
        for (int i = 0; i < m; ++i) {
            const int64_t c = __ldg(sp.off[i] + r + 1) - __ldg(sp.off[i] + r);
            if (c > 0) mask |= 1u << i;
            over |= c > 20;
        }
        present[r] = over ? (mask | 0x80000000u) : mask;
        if (over || mask == 0) continue;
        for (uint32_t a = mask; a; a &= a - 1) {
            const int i = __ffs(a) - 1;
            atomicAdd(&s_cnt[i * m + i], 1u);
            for (uint32_t b = a & (a - 1); b; b &= b - 1) atomicAdd(&s_cnt[i * m + (__ffs(b) - 1)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m * m; i += 256)
        if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long)s_cnt[i]);
}

__global__ void __launch_bounds__(256) cooc_pair_flags_kernel(const uint32_t* __restrict__ present, int64_t n_seq, uint32_t both,
                                                              uint32_t* __restrict__ flags) {
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n_seq; r += (int64_t)gridDim.x * 256) {
        const uint32_t p = __ldg(present + r);
        flags[r] = ((p & both) == both && !(p >> 31)) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) cooc_pair_fill_kernel(const int64_t* __restrict__ off_i, const int32_t* __restrict__ pos_i,
                                                             const int64_t* __restrict__ off_j, const int32_t* __restrict__ pos_j,
                                                             const uint32_t* __restrict__ flags, const int64_t* __restrict__ at, int64_t n_seq,
                                                             int64_t* __restrict__ read_out, int32_t* __restrict__ diff2_out) {
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n_seq; r += (int64_t)gridDim.x * 256) {
        if (!__ldg(flags + r)) continue;
        const int64_t o = __ldg(at + r);
        read_out[o] = r;
        diff2_out[o] = median2(pos_j, __ldg(off_j + r), __ldg(off_j + r + 1)) - median2(pos_i, __ldg(off_i + r), __ldg(off_i + r + 1));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) sum_counts_kernel(const T* __restrict__ cnt, int64_t n, unsigned long long* __restrict__ out) {
    long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += (long long)__ldg(cnt + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __shared__ long long warp_acc[8];
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += warp_acc[w];
        atomicAdd(out, (unsigned long long)t);
    }
}

unsigned int grid_cap(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    return (unsigned int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

int kmap_sum_counts_i32(const int32_t* cnt, int64_t n, int64_t* sum_out, void* stream) {
    KMAP_REQUIRE(n >= 0 && sum_out && (cnt || n == 0), "bad argument");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(sum_out, 0, 8, s);
    if (n) sum_counts_kernel<int32_t><<<grid_cap(n), 256, 0, s>>>(cnt, n, reinterpret_cast<unsigned long long*>(sum_out));
    return kmap_check_launch("sum_counts");
}

int kmap_sum_counts_i64(const int64_t* cnt, int64_t n, int64_t* sum_out, void* stream) {
    KMAP_REQUIRE(n >= 0 && sum_out && (cnt || n == 0), "bad argument");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(sum_out, 0, 8, s);
    if (n) sum_counts_kernel<long long><<<grid_cap(n), 256, 0, s>>>(reinterpret_cast<const long long*>(cnt), n, reinterpret_cast<unsigned long long*>(sum_out));
    return kmap_check_launch("sum_counts");
}

int kmap_cooc_reads(const int64_t* const* offsets_host, const int32_t* const* positions_host, int m, int64_t n_seq, uint32_t* present_out,
                    int64_t* counts_out, void* stream) {
    KMAP_REQUIRE(m >= 1 && m <= CO_MAX && n_seq >= 0 && offsets_host && positions_host && counts_out, "bad argument (at most 31 motifs per call)");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(counts_out, 0, (size_t)m * m * 8, s);
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(present_out, "null pointer");
    ScanPtrs sp;
    for (int i = 0; i < CO_MAX; ++i) { sp.off[i] = i < m ? offsets_host[i] : nullptr; sp.pos[i] = i < m ? positions_host[i] : nullptr; }
    cooc_reads_kernel<<<grid_cap(n_seq), 256, 0, s>>>(sp, m, n_seq, present_out, reinterpret_cast<unsigned long long*>(counts_out));
    return kmap_check_launch("cooc_reads");
}

int kmap_cooc_pair_flags(const uint32_t* present, int64_t n_seq, int i, int j, uint32_t* flags_out, void* stream) {
    KMAP_REQUIRE(n_seq >= 0 && i >= 0 && j >= 0 && i < CO_MAX && j < CO_MAX && i != j, "bad argument");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(present && flags_out, "null pointer");
    cooc_pair_flags_kernel<<<grid_cap(n_seq), 256, 0, as_stream(stream)>>>(present, n_seq, (1u << i) | (1u << j), flags_out);
    return kmap_check_launch("cooc_pair_flags");
}

int kmap_cooc_pair_fill(const int64_t* off_i, const int32_t* pos_i, const int64_t* off_j, const int32_t* pos_j, const uint32_t* flags,
                        const int64_t* flag_offsets, int64_t n_seq, int64_t* read_out, int32_t* diff2_out, void* stream) {
    KMAP_REQUIRE(n_seq >= 0, "negative size");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(off_i && off_j && flags && flag_offsets && read_out && diff2_out, "null pointer");
    cooc_pair_fill_kernel<<<grid_cap(n_seq), 256, 0, as_stream(stream)>>>(off_i, pos_i, off_j, pos_j, flags, flag_offsets, n_seq, read_out, diff2_out);
    return kmap_check_launch("cooc_pair_fill");
}

}  // extern "C"
