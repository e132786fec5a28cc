// k-mer counting into a dense uint32[4^k] table in HBM: comp_kmer_hash_taichi + count_uniq_hash
// (kmer_count.py:449-491) fused, with remove_duplicate_hash_per_seq (kmer_count.py:743-760) fused in for the
// default (non-repetitive) mode.  Hashes are never materialised.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// plain count: one thread = 32 consecutive positions (one validity word).  The thread keeps three packed words
// and two validity words in registers; every window is a funnel shift with a compile-time amount.
// ------------------------------------------------------------------------------------------------------------
constexpr int COUNT_BLOCK = 256;

__global__ void __launch_bounds__(COUNT_BLOCK) count_dense_kernel(const uint32_t* __restrict__ packed,
                                                                  const uint32_t* __restrict__ valid,
                                                                  int64_t n_words, int k, uint32_t* __restrict__ table) {
    const int64_t t = (int64_t)blockIdx.x * COUNT_BLOCK + threadIdx.x;
    if (t >= n_words) return;
    const uint32_t v0 = __ldg(valid + t), v1 = __ldg(valid + t + 1);
    if (v0 == 0) return;                       // nothing valid starts here
    const uint2 w01 = __ldg(reinterpret_cast<const uint2*>(packed + 2 * t));
    const uint32_t w2 = __ldg(packed + 2 * t + 2);
    const uint32_t km = (1u << k) - 1u;
    const int sh = 32 - 2 * k;
    // run-length aggregation inside the thread: homopolymer / tandem runs hit one cell many times
    uint32_t prev = 0xFFFFFFFFu, run = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t vb = __funnelshift_r(v0, v1, i);
        if ((vb & km) == km) {
            const uint32_t x = (i < 16) ? __funnelshift_l(w01.y, w01.x, 2 * i) : __funnelshift_l(w2, w01.y, 2 * (i - 16));
            const uint32_t h = x >> sh;
            if (h == prev) { ++run; }
            else {
                if (run) atomicAdd(table + prev, run);
                prev = h; run = 1;
            }
        }
    }
    if (run) atomicAdd(table + prev, run);
}

// ------------------------------------------------------------------------------------------------------------
// de-duplicated count, short reads: one warp = one read.  Each round takes 32 windows; duplicates inside the
// round are removed with __match_any_sync, duplicates across rounds with a per-warp open-addressing set in
// shared memory that is filled warp-synchronously (plain loads/stores, no atomics: a lane that finds an empty
// slot writes its hash, the warp syncs, and whoever reads its own hash back has won the slot).
// ------------------------------------------------------------------------------------------------------------
constexpr int DD_WARPS = 8;                    // warps per block
constexpr int DD_SLOTS = 512;                  // slots per warp (2 KB)
constexpr int DD_WARP_MAX = 256;               // windows a warp handles on chip (load factor <= 0.5)
constexpr int DD_BLOCK_SLOTS = 16384;          // block-per-read set (64 KB)
constexpr int DD_BLOCK_MAX = DD_BLOCK_SLOTS / 2;

struct DedupWork {            // layout of the `work` scratch (uint32 words)
    uint32_t n_medium;        // reads with DD_WARP_MAX < windows <= DD_BLOCK_MAX
    uint32_t n_long;          // reads with more windows
    uint32_t pad[2];
    // followed by uint32 medium_ids[n_seq], uint32 long_ids[n_seq]  (read index; n_seq < 2^32)
};

__global__ void __launch_bounds__(DD_WARPS * 32) count_dedup_warp_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, int64_t n,
    const int64_t* __restrict__ borders, int64_t n_seq, int k, uint32_t* __restrict__ table,
    uint32_t* __restrict__ work) {
    __shared__ uint32_t sets[DD_WARPS][DD_SLOTS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* set = sets[wib];
    const int64_t warp0 = (int64_t)blockIdx.x * DD_WARPS + wib;
    const int64_t n_warps = (int64_t)gridDim.x * DD_WARPS;
    const uint32_t km = (1u << k) - 1u;
    const int sh = 32 - 2 * k;
    uint32_t* medium_ids = work + 4;
    uint32_t* long_ids = work + 4 + n_seq;

    for (int64_t r = warp0; r < n_seq; r += n_warps) {
        int64_t st = __ldg(borders + 2 * r), en = __ldg(borders + 2 * r + 1);
        if (st < 0) st = 0;
        if (en > n) en = n;
        const int64_t n_win = en - st - k + 1;          // windows that lie inside the read
        if (n_win <= 0) continue;
        if (n_win > DD_WARP_MAX) {
            if (lane == 0) {
                if (n_win <= DD_BLOCK_MAX) medium_ids[atomicAdd(&work[0], 1u)] = (uint32_t)r;
                else long_ids[atomicAdd(&work[1], 1u)] = (uint32_t)r;
            }
            continue;
        }
        // clear the set (16 words per lane)
        {
            uint4* s4 = reinterpret_cast<uint4*>(set);
            const uint4 e = make_uint4(KMAP_EMPTY_SLOT, KMAP_EMPTY_SLOT, KMAP_EMPTY_SLOT, KMAP_EMPTY_SLOT);
#pragma unroll
            for (int j = 0; j < DD_SLOTS / 4 / 32; ++j) s4[lane + 32 * j] = e;
        }
        __syncwarp();
        for (int64_t i0 = 0; i0 < n_win; i0 += 32) {
            const int64_t i = i0 + lane;
            bool ok = i < n_win;
            uint32_t h = 0x80000000u | (uint32_t)lane;   // distinct dummy keys for idle lanes
            if (ok) {
                const int64_t p = st + i;
                ok = (valid32(valid, p) & km) == km;
                if (ok) h = window16(packed, p) >> sh;
            }
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
            bool pending = ok && ((peers & ((1u << lane) - 1u)) == 0);   // lowest lane of each distinct hash
            bool fresh = false;
            uint32_t slot = mix32(h) & (DD_SLOTS - 1);
            while (__any_sync(0xFFFFFFFFu, pending)) {
                uint32_t cur = 0;
                if (pending) cur = set[slot];
                __syncwarp();
                if (pending) {
                    if (cur == h) pending = false;                       // seen in an earlier round
                    else if (cur == KMAP_EMPTY_SLOT) set[slot] = h;      // try to claim
                    else slot = (slot + 1) & (DD_SLOTS - 1);             // occupied by another hash
                }
                __syncwarp();
                if (pending && cur == KMAP_EMPTY_SLOT) {
                    if (set[slot] == h) { pending = false; fresh = true; }
                    else slot = (slot + 1) & (DD_SLOTS - 1);             // lost the slot to another hash
                }
                __syncwarp();
            }
            if (fresh) atomicAdd(table + h, 1u);
        }
        __syncwarp();
    }
}

// medium reads: one block = one read, 64 KB set, atomicCAS insertion
__global__ void __launch_bounds__(256) count_dedup_block_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, int64_t n,
    const int64_t* __restrict__ borders, int64_t n_seq, int k, uint32_t* __restrict__ table,
    const uint32_t* __restrict__ work) {
    extern __shared__ uint32_t bset[];
    const uint32_t n_medium = work[0];
    const uint32_t* medium_ids = work + 4;
    const uint32_t km = (1u << k) - 1u;
    const int sh = 32 - 2 * k;
    for (uint32_t q = blockIdx.x; q < n_medium; q += gridDim.x) {
        const int64_t r = medium_ids[q];
        int64_t st = borders[2 * r], en = borders[2 * r + 1];
        if (st < 0) st = 0;
        if (en > n) en = n;
        const int64_t n_win = en - st - k + 1;
        for (int i = threadIdx.x; i < DD_BLOCK_SLOTS; i += blockDim.x) bset[i] = KMAP_EMPTY_SLOT;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n_win; i += blockDim.x) {
            const int64_t p = st + i;
            if ((valid32(valid, p) & km) != km) continue;
            const uint32_t h = window16(packed, p) >> sh;
            uint32_t slot = mix32(h) & (DD_BLOCK_SLOTS - 1);
            while (true) {
                const uint32_t old = atomicCAS(&bset[slot], KMAP_EMPTY_SLOT, h);
                if (old == KMAP_EMPTY_SLOT) { atomicAdd(table + h, 1u); break; }
                if (old == h) break;
                slot = (slot + 1) & (DD_BLOCK_SLOTS - 1);
            }
        }
        __syncthreads();
    }
}

// long reads: one read per launch, a 4^k-bit "seen" bitmap in HBM; the first window to set a bit counts
__global__ void __launch_bounds__(256) count_dedup_bitmap_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, int64_t st, int64_t n_win, int k,
    uint32_t* __restrict__ table, uint32_t* __restrict__ bitmap) {
    const uint32_t km = (1u << k) - 1u;
    const int sh = 32 - 2 * k;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_win; i += stride) {
        const int64_t p = st + i;
        if ((valid32(valid, p) & km) != km) continue;
        const uint32_t h = window16(packed, p) >> sh;
        const uint32_t bit = 1u << (h & 31);
        const uint32_t old = atomicOr(bitmap + (h >> 5), bit);
        if (!(old & bit)) atomicAdd(table + h, 1u);
    }
}

// count_uniq_hash (kmer_count.py:476-491) on an already materialised hash array: invalid hashes (>= 4^k) are skipped
__global__ void __launch_bounds__(256) count_hashes_kernel(const uint32_t* __restrict__ hash, int64_t n, uint32_t n_cells,
                                                           uint32_t* __restrict__ table) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t h = __ldg(hash + i);
        if (h < n_cells) atomicAdd(table + h, 1u);
    }
}
// table[kh[i]] += cnt[i]: rebuilds a dense table from a (unique hash, count) list
__global__ void __launch_bounds__(256) scatter_counts_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ cnt, int64_t n,
                                                             uint32_t n_cells, uint32_t* __restrict__ table) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t h = __ldg(kh + i);
        if (h < n_cells) atomicAdd(table + h, (uint32_t)__ldg(cnt + i));
    }
}
// the in-place side effect of merge_revcom on the caller's count array (kmer_count.py:661):
// cnt[i] += cnt[index of rc(kh[i])] when that reverse complement is present (a palindrome adds itself)
__global__ void __launch_bounds__(256) list_add_rc_counts_kernel(const uint32_t* __restrict__ kh, int32_t* __restrict__ cnt, int64_t n,
                                                                 int k, const uint32_t* __restrict__ table) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) cnt[i] += (int32_t)__ldg(table + revcom32(__ldg(kh + i) & lowmask32(k), k));
}

static unsigned int strided_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return (unsigned int)(g < 1 ? 1 : g);
}

}  // namespace

// reads too long for the warp path, listed in `work` (medium ids at work+4, long ids at work+4+n_seq)
int kmap_count_long_reads(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                          int k, uint32_t* table, uint32_t* work, uint32_t* bitmap, const uint32_t counts[2], cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    int rc;
    if (counts[0]) {
        // (per-device attribute: set on every call, not once per process)
        cudaFuncSetAttribute(count_dedup_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_BLOCK_SLOTS * 4);
        const unsigned int g = counts[0] < 148u * 3u ? counts[0] : 148u * 3u;
        count_dedup_block_kernel<<<g, 256, DD_BLOCK_SLOTS * 4, s>>>(packed, valid, n, borders, n_seq, k, table, work);
        rc = kmap_check_launch("count_dedup_block");
        if (rc) return rc;
    }
    if (counts[1]) {
        if (!bitmap) { kmap_set_error("count_dense_dedup: %u reads need the bitmap scratch", counts[1]); return KMAP_ERR_NEED_SCRATCH; }
        // read the ids and borders of the long reads back (few, by construction)
        uint32_t* ids = new uint32_t[counts[1]];
        e = cudaMemcpyAsync(ids, work + 4 + n_seq, (size_t)counts[1] * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        const int64_t bitmap_words = ((int64_t)1 << (2 * k)) / 32 > 0 ? ((int64_t)1 << (2 * k)) / 32 : 1;
        for (uint32_t q = 0; q < counts[1] && e == cudaSuccess; ++q) {
            int64_t b[2];
            e = cudaMemcpyAsync(b, borders + 2 * (int64_t)ids[q], 16, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) break;
            int64_t st = b[0] < 0 ? 0 : b[0], en = b[1] > n ? n : b[1];
            const int64_t n_win = en - st - k + 1;
            if (n_win <= 0) continue;
            e = cudaMemsetAsync(bitmap, 0, (size_t)bitmap_words * 4, s);
            if (e != cudaSuccess) break;
            int64_t g = (n_win + 255) / 256;
            if (g > 148 * 16) g = 148 * 16;
            count_dedup_bitmap_kernel<<<(unsigned int)g, 256, 0, s>>>(packed, valid, st, n_win, k, table, bitmap);
            e = cudaGetLastError();
        }
        delete[] ids;
        if (e != cudaSuccess) { kmap_set_error("count_dense_dedup(long reads): %s", cudaGetErrorString(e)); return (int)e; }
    }
    return KMAP_OK;
}

extern "C" {

int kmap_count_hashes_u32(const uint32_t* hash, int64_t n, int k, uint32_t* table, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(hash && table, "null pointer");
    count_hashes_kernel<<<strided_grid(n), 256, 0, as_stream(stream)>>>(hash, n, 1u << (2 * k), table);
    return kmap_check_launch("count_hashes");
}
int kmap_scatter_counts(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, uint32_t* table, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && cnt && table, "null pointer");
    scatter_counts_kernel<<<strided_grid(n), 256, 0, as_stream(stream)>>>(kh, cnt, n, 1u << (2 * k), table);
    return kmap_check_launch("scatter_counts");
}
int kmap_list_add_rc_counts(const uint32_t* kh, int32_t* cnt, int64_t n, int k, const uint32_t* table, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && cnt && table, "null pointer");
    list_add_rc_counts_kernel<<<strided_grid(n), 256, 0, as_stream(stream)>>>(kh, cnt, n, k, table);
    return kmap_check_launch("list_add_rc_counts");
}

int kmap_count_dense(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && table, "null pointer");
    const int64_t n_words = (n + 31) / 32;
    count_dense_kernel<<<grid_for(n_words, COUNT_BLOCK), COUNT_BLOCK, 0, as_stream(stream)>>>(packed, valid, n_words, k, table);
    return kmap_check_launch("count_dense");
}

int64_t kmap_dedup_work_words(int64_t n_seq) { return 4 + 2 * n_seq; }

int kmap_count_dense_dedup(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders,
                           int64_t n_seq, int k, uint32_t* table, uint32_t* work, uint32_t* bitmap, void* stream) {
    KMAP_REQUIRE(n >= 0 && n_seq >= 0 && k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    KMAP_REQUIRE(n_seq < (int64_t)0xFFFFFFFFll, "too many reads for one call (shard the input)");
    if (n == 0 || n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && table && borders && work, "null pointer");
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(work, 0, 16, s);
    if (e != cudaSuccess) { kmap_set_error("count_dense_dedup: %s", cudaGetErrorString(e)); return (int)e; }
    int64_t blocks = (n_seq + DD_WARPS - 1) / DD_WARPS;
    const int64_t max_blocks = 148 * 8 * 4;     // persistent-ish: 8 blocks of 8 warps per SM, 4 waves
    if (blocks > max_blocks) blocks = max_blocks;
    count_dedup_warp_kernel<<<(unsigned int)blocks, DD_WARPS * 32, 0, s>>>(packed, valid, n, borders, n_seq, k, table, work);
    int rc = kmap_check_launch("count_dedup_warp");
    if (rc) return rc;
    uint32_t counts[2] = {0, 0};
    e = cudaMemcpyAsync(counts, work, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { kmap_set_error("count_dense_dedup: %s", cudaGetErrorString(e)); return (int)e; }
    if (counts[0] || counts[1]) return kmap_count_long_reads(packed, valid, n, borders, n_seq, k, table, work, bitmap, counts, s);
    return KMAP_OK;
}

}  // extern "C"
