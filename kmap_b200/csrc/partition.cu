// Level-k count by key partitioning: comp_kmer_hash_taichi + count_uniq_hash (kmer_count.py:449-491) for a table that
// does not fit in L2.  DESIGN.md section 4.5.
//
// Measured on B200: random RED.ADD into an L2-resident table sustains ~187 G updates/s (22 G/s when the table lives in
// HBM), shared-memory ATOMS ~2000 G/s.  So instead of one global atomic per window, every counted window's key is
// split into (bucket = key >> 16, suffix = key & 0xFFFF):
//   1. bucket_hist_kernel   one pass over the packed reads: windows per bucket        -> bucket offsets (scan)
//   2. partition_kernel     one pass: a CTA counting-sorts a tile of 32768 positions by bucket in shared memory,
//                           reserves room in every non-empty bucket with one global atomic and writes the 16-bit
//                           suffixes as runs (lanes of a warp store to consecutive addresses).  (Private per-CTA
//                           ranges without the atomics were measured: slower, 148 x 4096 open partial lines spill
//                           out of L2 and DRAM writes grow 2.5x.)
//   3. bucket_count_kernel  one CTA per bucket: 65536 cells as packed 16-bit counters in shared memory (128 KB),
//                           suffixes stream in with 128-bit loads, the table slice is written once, coalesced.
// The result is identical to kmap_count_dense (integer sums are order independent).
#include "common.cuh"
#include "tile.cuh"

namespace {

constexpr int PT_THREADS = KMAP_TILE_THREADS;
constexpr int PT_TILE = PT_THREADS * 32;          // positions (= staged entries) per tile
constexpr int PT_MAX_BUCKETS = 4096;              // k <= 14
constexpr int PT_MAX_PER = PT_MAX_BUCKETS / PT_THREADS;
constexpr int PT_WU = 8;                          // write-out entries in flight per thread

__device__ __forceinline__ void tile_hist(const TileWords& t, int sh, uint32_t* cnt) {
    if (t.fresh) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((t.fresh >> i) & 1u) atomicAdd(&cnt[key_at(t, i, sh) >> 16], 1u);
    }
}

// ---- 1. windows per bucket (+ the run-end corrections of the all-k count) -------------------------------------------------
// With TERMINAL, the pass also does what terminal_corrections_kernel (count_all.cu) does: "+1 at level v" for every
// window with exactly v valid bases (kmin <= v < k) in front of a run end.  Those are scattered global REDs; issued from
// this kernel they overlap its shared-memory-bound histogram work instead of costing passes of their own.
template <bool TERMINAL>
__global__ void __launch_bounds__(PT_THREADS) bucket_hist_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                 const uint32_t* __restrict__ hide, int64_t n_words, int64_t n_tiles,
                                                                 int k, int n_buckets, unsigned long long* __restrict__ hist,
                                                                 KmapTableSet tabs, int kmin) {
    __shared__ uint32_t cnt[PT_MAX_BUCKETS];
    __shared__ uint32_t* stab[16];
    if (TERMINAL && threadIdx.x < 16) stab[threadIdx.x] = tabs.t[threadIdx.x];
    for (int b = threadIdx.x; b < n_buckets; b += PT_THREADS) cnt[b] = 0;
    __syncthreads();
    const int sh = 32 - 2 * k;
    RawWords nxt = load_raw_words(packed, valid, hide, n_words, blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RawWords r = nxt;
        nxt = load_raw_words(packed, valid, hide, n_words, tile + gridDim.x);      // (past the end: zeros)
        tile_hist(cook(r, k), sh, cnt);
        if (TERMINAL) {
            run_end_corrections(packed, valid, hide, r, tile * PT_THREADS + threadIdx.x, kmin, k, stab);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_buckets; b += PT_THREADS)
        if (cnt[b]) atomicAdd(hist + b, (unsigned long long)cnt[b]);
}

// block-wide exclusive scan of one value per thread (PT_THREADS threads; two barriers inside)
template <typename T>
__device__ __forceinline__ T block_scan_excl(T v, T* warp_sums) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        const T s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : T(0);
        T si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T y = __shfl_up_sync(0xFFFFFFFFu, si, o);
            if (lane >= o) si += y;
        }
        warp_sums[lane] = si - s;
    }
    __syncthreads();
    return warp_sums[w] + incl - v;
}

// exclusive scan of the bucket histogram (<= 4096 values, one block): base[0..n_buckets], cursor[b] = base[b]
__global__ void __launch_bounds__(PT_THREADS) bucket_scan_kernel(const unsigned long long* __restrict__ hist, int n_buckets,
                                                                 unsigned long long* __restrict__ base, unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long warp_sums[32];
    const int per = (n_buckets + PT_THREADS - 1) / PT_THREADS;
    const int lo = threadIdx.x * per, hi = min(lo + per, n_buckets);
    unsigned long long mine = 0;
    for (int b = lo; b < hi; ++b) mine += hist[b];
    unsigned long long run = block_scan_excl<unsigned long long>(mine, warp_sums);
    if (threadIdx.x == PT_THREADS - 1) base[n_buckets] = run + mine;
    for (int b = lo; b < hi; ++b) { base[b] = run; cursor[b] = run; run += hist[b]; }
}

// ---- 2. partition -------------------------------------------------------------------------------------------------------
// Per tile: (a) histogram by bucket, (b) exclusive scan -> tile-local starts, (c) room in every non-empty bucket
// reserved with one global atomic (all CTAs append to the same moving tail of a bucket, so the partial sectors of
// neighbouring runs meet in L2), (d) scatter of the keys into bucket order in shared memory, (e) write-out: entry i of
// the sorted tile goes to gdelta[bucket] + i, so consecutive lanes write consecutive addresses inside a run.
// (Measured alternatives, 1e8 reads x 100 bp, k = 14: two resident CTAs of 512 threads with tiles of 16384 positions
// take 66.6 ms against 40.1 ms -- the per-tile costs (scan over 4096 buckets, 4096 cursor atomics, barriers) double and
// the runs get shorter; private per-CTA destination ranges without cursor atomics take 51 ms, see the file header.)
// The tile loop is software-pipelined: the raw words of the next tile are in flight during the whole current tile, and
// its histogram (fire-and-forget shared-memory atomics) is issued together with the latency-bound write-out.
template <int PER>       // buckets per thread in the scan step: n_buckets <= PER * PT_THREADS
__global__ void __launch_bounds__(PT_THREADS, 1) partition_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                  const uint32_t* __restrict__ hide, int64_t n_words, int64_t n_tiles,
                                                                  int k, int n_buckets, unsigned long long* __restrict__ cursor,
                                                                  uint16_t* __restrict__ suffixes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sorted = reinterpret_cast<uint32_t*>(smem_raw);                                  // PT_TILE keys, grouped by bucket
    unsigned long long* gdelta = reinterpret_cast<unsigned long long*>(sorted + PT_TILE);     // global start - tile start, per bucket
    uint32_t* cnt = reinterpret_cast<uint32_t*>(gdelta + PER * PT_THREADS);                   // histogram of the coming tile
    uint32_t* off = cnt + PER * PT_THREADS;                                                   // running tile offsets
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t tile_total;
    const int sh = 32 - 2 * k;
#pragma unroll
    for (int j = 0; j < PER; ++j) cnt[PER * threadIdx.x + j] = 0;
    __syncthreads();
    TileWords cur = load_tile_words(packed, valid, hide, n_words, blockIdx.x, k);
    RawWords raw = load_raw_words(packed, valid, hide, n_words, (int64_t)blockIdx.x + gridDim.x);
    tile_hist(cur, sh, cnt);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();                       // histogram of `cur` complete; write-out of the previous tile complete
        // (b) exclusive scan over the buckets; thread owns buckets [PER*tid, PER*tid + PER)
        uint32_t c[PER], s[PER];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { c[j] = cnt[PER * threadIdx.x + j]; cnt[PER * threadIdx.x + j] = 0; mine += c[j]; }
        uint32_t run = block_scan_excl<uint32_t>(mine, warp_sums);
        if (threadIdx.x == PT_THREADS - 1) tile_total = run + mine;
#pragma unroll
        for (int j = 0; j < PER; ++j) { s[j] = run; off[PER * threadIdx.x + j] = run; run += c[j]; }
        __syncthreads();
        // (c) reserve room in the global buckets (latency overlaps the shared-memory scatter below)
        unsigned long long g[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            g[j] = 0;
            if (c[j]) g[j] = atomicAdd(cursor + PER * threadIdx.x + j, (unsigned long long)c[j]);
        }
        // (d) scatter the keys into bucket order
        if (cur.fresh) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if ((cur.fresh >> i) & 1u) {
                    const uint32_t key = key_at(cur, i, sh);
                    sorted[atomicAdd(&off[key >> 16], 1u)] = key;
                }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) gdelta[PER * threadIdx.x + j] = g[j] - s[j];
        // the next tile: its words arrived long ago; fetch the one after it
        cur = cook(raw, k);
        raw = load_raw_words(packed, valid, hide, n_words, tile + 2 * (int64_t)gridDim.x);
        __syncthreads();
        const uint32_t total = tile_total;
        // (a) histogram of the next tile, then (e) write-out of this one
        tile_hist(cur, sh, cnt);
        for (uint32_t i0 = threadIdx.x; i0 < total; i0 += PT_WU * PT_THREADS) {
            uint32_t key[PT_WU];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) {
                const uint32_t i = i0 + u * PT_THREADS;
                key[u] = i < total ? sorted[i] : 0u;
            }
            unsigned long long d[PT_WU];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) d[u] = gdelta[key[u] >> 16];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) {
                const uint32_t i = i0 + u * PT_THREADS;
                if (i < total) suffixes[d[u] + i] = (uint16_t)key[u];
            }
        }
    }
}

// ---- 3. per-bucket count in shared memory --------------------------------------------------------------------------------
constexpr int BC_THREADS = 1024;
constexpr int BC_CELLS = 65536;
constexpr int BC_WORDS = BC_CELLS / 2;            // two 16-bit counters per word
constexpr int BC_UNROLL = 4;                      // 128-bit loads in flight per thread

// A half-word counter that reaches 0x8000 is folded into the global cell at once (the fold happens long before the
// half could carry into its neighbour: at most BC_THREADS increments are in flight).
__device__ __forceinline__ void bump(uint32_t* sm, uint32_t s, uint32_t* __restrict__ slice, int* spilled) {
    const uint32_t shift = (s & 1u) << 4;
    const uint32_t old = atomicAdd(&sm[s >> 1], 1u << shift);
    if (((old >> shift) & 0xFFFFu) == 0x7FFFu) {
        atomicSub(&sm[s >> 1], 0x8000u << shift);
        atomicAdd(slice + s, 0x8000u);
        *spilled = 1;
    }
}

__global__ void __launch_bounds__(BC_THREADS, 1) bucket_count_kernel(const uint16_t* __restrict__ suffixes,
                                                                     const unsigned long long* __restrict__ base, int n_buckets,
                                                                     uint32_t* __restrict__ table) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ int spilled;
    for (int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        for (int w = threadIdx.x; w < BC_WORDS; w += BC_THREADS) sm[w] = 0;
        if (threadIdx.x == 0) spilled = 0;
        __syncthreads();
        uint32_t* slice = table + (size_t)b * BC_CELLS;
        const unsigned long long lo = base[b], hi = base[b + 1];
        // head up to a 16-byte boundary, 8 suffixes per 128-bit load, tail
        unsigned long long a0 = (lo + 7ull) & ~7ull;
        if (a0 > hi) a0 = hi;
        const unsigned long long a1 = a0 + ((hi - a0) & ~7ull);
        int my_spill = 0;
        for (unsigned long long i = lo + threadIdx.x; i < a0; i += BC_THREADS) bump(sm, suffixes[i], slice, &my_spill);
        const uint4* v = reinterpret_cast<const uint4*>(suffixes + a0);
        const unsigned long long n_vec = (a1 - a0) >> 3;
        for (unsigned long long i = threadIdx.x; i < n_vec; i += BC_UNROLL * BC_THREADS) {
            uint4 q[BC_UNROLL];
#pragma unroll
            for (int u = 0; u < BC_UNROLL; ++u) {
                const unsigned long long iu = i + (unsigned long long)u * BC_THREADS;
                q[u] = iu < n_vec ? __ldcs(v + iu) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < BC_UNROLL; ++u) {
                if (i + (unsigned long long)u * BC_THREADS >= n_vec) break;
                const uint32_t ws[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bump(sm, ws[j] & 0xFFFFu, slice, &my_spill);
                    bump(sm, ws[j] >> 16, slice, &my_spill);
                }
            }
        }
        for (unsigned long long i = a1 + threadIdx.x; i < hi; i += BC_THREADS) bump(sm, suffixes[i], slice, &my_spill);
        if (my_spill) spilled = 1;
        __syncthreads();
        if (!spilled) {               // the slice was zero: plain coalesced stores
            uint4* out = reinterpret_cast<uint4*>(slice);
            for (int w = threadIdx.x; w < BC_WORDS / 2; w += BC_THREADS) {
                const uint2 p = reinterpret_cast<const uint2*>(sm)[w];
                out[w] = make_uint4(p.x & 0xFFFFu, p.x >> 16, p.y & 0xFFFFu, p.y >> 16);
            }
        } else {                      // some cells already hold folded 0x8000s: add on top
            for (int w = threadIdx.x; w < BC_WORDS; w += BC_THREADS) {
                const uint32_t p = sm[w];
                if (p & 0xFFFFu) atomicAdd(slice + 2 * w, p & 0xFFFFu);
                if (p >> 16) atomicAdd(slice + 2 * w + 1, p >> 16);
            }
        }
        __syncthreads();
    }
}

struct PartScratch {
    unsigned long long *hist, *base, *cursor;
    uint16_t* suffixes;
};

static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static PartScratch carve(void* scratch, int n_buckets) {
    PartScratch p;
    uint8_t* q = reinterpret_cast<uint8_t*>(scratch);
    p.hist = reinterpret_cast<unsigned long long*>(q);
    p.base = p.hist + n_buckets;
    p.cursor = p.base + n_buckets + 1;
    p.suffixes = reinterpret_cast<uint16_t*>(q + align_up((int64_t)(3 * n_buckets + 1) * 8, 256));
    return p;
}

}  // namespace

extern "C" int64_t kmap_partition_scratch_bytes(int64_t n, int k) {
    if (k < 9 || k > 14 || n < 0) return 0;
    const int n_buckets = 1 << (2 * (k - 8));
    return align_up((int64_t)(3 * n_buckets + 1) * 8, 256) + align_up(2 * n, 256) + 256;
}

// table[h] = number of counted windows with key h, for every h (the slice of every bucket is overwritten or, where
// folded counters were spilled, added to: the caller zeroes the table first).  hide may be NULL.
// terminal_tabs (may be NULL): also add the run-end corrections of levels kmin..k-1 to those tables (count_all.cu).
int kmap_count_partitioned(const uint32_t* packed, const uint32_t* valid, const uint32_t* hide, int64_t n, int k, uint32_t* table,
                           void* scratch, const KmapTableSet* terminal_tabs, int kmin, void* const* step_events, cudaStream_t s) {
    const int n_buckets = 1 << (2 * (k - 8));
    const PartScratch p = carve(scratch, n_buckets);
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_tiles = (n_words + PT_THREADS - 1) / PT_THREADS;
    cudaError_t e = cudaMemsetAsync(p.hist, 0, (size_t)n_buckets * 8, s);
    if (e != cudaSuccess) { kmap_set_error("count_partitioned: %s", cudaGetErrorString(e)); return (int)e; }
    const unsigned int g1 = (unsigned int)(n_tiles < 148 * 2 ? n_tiles : 148 * 2);
    if (terminal_tabs)
        bucket_hist_kernel<true><<<g1, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, p.hist, *terminal_tabs, kmin);
    else
        bucket_hist_kernel<false><<<g1, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, p.hist, KmapTableSet(), k);
    bucket_scan_kernel<<<1, PT_THREADS, 0, s>>>(p.hist, n_buckets, p.base, p.cursor);
    if (step_events && step_events[0]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(step_events[0]), s);
    const unsigned int g2 = (unsigned int)(n_tiles < 148 ? n_tiles : 148);
    static bool attr_set = false;
    const int smem1 = PT_TILE * 4 + 1 * PT_THREADS * 16, smem4 = PT_TILE * 4 + PT_MAX_PER * PT_THREADS * 16;
    if (!attr_set) {
        cudaFuncSetAttribute(partition_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
        cudaFuncSetAttribute(partition_kernel<PT_MAX_PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4);
        cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_WORDS * 4);
        attr_set = true;
    }
    if (n_buckets <= PT_THREADS)
        partition_kernel<1><<<g2, PT_THREADS, smem1, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, p.cursor, p.suffixes);
    else
        partition_kernel<PT_MAX_PER><<<g2, PT_THREADS, smem4, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, p.cursor, p.suffixes);
    if (step_events && step_events[1]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(step_events[1]), s);
    const unsigned int g3 = (unsigned int)(n_buckets < 148 ? n_buckets : 148);
    bucket_count_kernel<<<g3, BC_THREADS, BC_WORDS * 4, s>>>(p.suffixes, p.base, n_buckets, table);
    return kmap_check_launch("count_partitioned");
}

extern "C" int kmap_count_dense_partitioned(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table,
                                            void* scratch, int64_t scratch_bytes, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 9 && k <= 14, "the partitioned count covers 9 <= k <= 14");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && table && scratch, "null pointer");
    KMAP_REQUIRE(scratch_bytes >= kmap_partition_scratch_bytes(n, k), "scratch too small (kmap_partition_scratch_bytes)");
    return kmap_count_partitioned(packed, valid, nullptr, n, k, table, scratch, nullptr, k, nullptr, as_stream(stream));
}
