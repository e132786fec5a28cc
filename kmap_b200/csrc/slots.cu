// Level-k count by SLOTTED key partitioning (12 <= k <= 14): comp_kmer_hash_taichi + count_uniq_hash
// (kmer_count.py:449-491) for a table that does not fit in L2.  DESIGN.md section 4.5.
//
// partition.cu sorts every tile exactly (histogram, scan, scatter, write-out of runs): three shared-memory atomics per
// window over the three launches, plus a histogram pass of its own to size the buckets.  Here the buckets need no
// sizing: the key is split into (bucket = top 12 bits, suffix = the other 2k-12 bits) and every (bucket, tile) pair
// owns ONE 32-byte sector of the scratch: a 16-bit count followed by up to 15 suffixes,
//     slots[bucket][tile][16]  (uint16).
// A tile of 32768 positions puts ~7 windows into each of the 4096 buckets, so a sector overflows about once in a
// thousand; the rare window that finds its sector full is counted with one global RED instead (exact for any input,
// merely slower for inputs whose keys pile up in few buckets).
//   1. slot_partition_kernel  one pass, two barriers per tile: one shared-memory atomic per window appends its suffix to
//                             the bucket's sector in a 128 KB stage; the stage is written out with 128-bit stores, whole
//                             sectors only, neighbouring tiles next to each other.  It also does the run-end
//                             corrections of the all-k count (tile.cuh).
//   2. slot_count_kernel      one CTA per bucket: its sectors are one contiguous stream; 4^(k-6) cells as packed 16-bit
//                             counters in shared memory, slice += counters, coalesced.
#include "common.cuh"
#include "tile.cuh"

namespace {

constexpr int SL_THREADS = KMAP_TILE_THREADS;
constexpr int SL_BUCKETS = 4096;
constexpr int SL_CAP = 15;                         // suffixes per sector (halfword 0 is the count)

template <bool TERMINAL>
__global__ void __launch_bounds__(SL_THREADS, 1) slot_partition_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                       const uint32_t* __restrict__ hide, int64_t n_words, int64_t n_tiles,
                                                                       int k, uint4* __restrict__ slots, uint32_t* __restrict__ table,
                                                                       KmapTableSet tabs, int kmin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint16_t* stage = reinterpret_cast<uint16_t*>(smem_raw);                                  // [SL_BUCKETS][16]
    uint32_t* fill = reinterpret_cast<uint32_t*>(smem_raw + SL_BUCKETS * 32);                 // [SL_BUCKETS]
    __shared__ uint32_t* stab[16];
    if (TERMINAL && threadIdx.x < 16) stab[threadIdx.x] = tabs.t[threadIdx.x];
    for (int b = threadIdx.x; b < SL_BUCKETS; b += SL_THREADS) fill[b] = 0;
    __syncthreads();
    const int sh = 32 - 2 * k;
    const int sbits = 2 * k - 12;
    const uint32_t smask = (1u << sbits) - 1u;
    RawWords nxt = load_raw_words(packed, valid, hide, n_words, blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RawWords r = nxt;
        nxt = load_raw_words(packed, valid, hide, n_words, tile + gridDim.x);      // (past the end: zeros)
        const TileWords t = cook(r, k);
        if (t.fresh) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if ((t.fresh >> i) & 1u) {
                    const uint32_t key = key_at(t, i, sh);
                    const uint32_t b = key >> sbits;
                    const uint32_t slot = atomicAdd(&fill[b], 1u);
                    if (slot < SL_CAP) stage[b * 16 + 1 + slot] = (uint16_t)(key & smask);
                    else atomicAdd(table + key, 1u);                                // sector full: count it directly
                }
        }
        if (TERMINAL) run_end_corrections(r, load_raw_prev(packed, valid, hide, n_words, tile), kmin, k, stab);
        __syncthreads();
        // write-out: 8192 x 16 bytes; lanes 2i, 2i+1 carry the two halves of one sector
        const uint4* st4 = reinterpret_cast<const uint4*>(stage);
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = st4[threadIdx.x + SL_THREADS * j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t q = threadIdx.x + SL_THREADS * j;
            const uint32_t b = q >> 1;
            if ((q & 1u) == 0) {
                const uint32_t c = min(fill[b], (uint32_t)SL_CAP);
                fill[b] = 0;
                v[j].x = (v[j].x & 0xFFFF0000u) | c;
            }
            __stcs(slots + ((size_t)b * n_tiles + tile) * 2 + (q & 1u), v[j]);
        }
        __syncthreads();
    }
}

// ---- per-bucket count in shared memory -----------------------------------------------------------------------------------
constexpr int SC_THREADS = 1024;
constexpr int SC_MAX_WORDS = 32768;               // 65536 cells (k = 14), two 16-bit counters per word

// A half-word counter that reaches 0x8000 is folded into the global cell at once (long before the half could carry
// into its neighbour: at most SC_THREADS increments are in flight).
__device__ __forceinline__ void bump(uint32_t* sm, uint32_t s, uint32_t* __restrict__ slice) {
    const uint32_t shift = (s & 1u) << 4;
    const uint32_t old = atomicAdd(&sm[s >> 1], 1u << shift);
    if (((old >> shift) & 0xFFFFu) == 0x7FFFu) {
        atomicSub(&sm[s >> 1], 0x8000u << shift);
        atomicAdd(slice + s, 0x8000u);
    }
}

__device__ __forceinline__ void bump_pair(uint32_t* sm, uint32_t w, int i0, uint32_t c, uint32_t* __restrict__ slice) {
    if ((uint32_t)i0 <= c && i0 >= 1) bump(sm, w & 0xFFFFu, slice);          // halfword i0 holds suffix number i0 (1-based)
    if ((uint32_t)(i0 + 1) <= c) bump(sm, w >> 16, slice);
}

__global__ void __launch_bounds__(SC_THREADS, 1) slot_count_kernel(const uint4* __restrict__ slots, int64_t n_tiles, int k,
                                                                   uint32_t* __restrict__ table) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const int cells = 1 << (2 * k - 12);
    const int words = cells >> 1;
    for (int b = blockIdx.x; b < SL_BUCKETS; b += gridDim.x) {
        for (int w = threadIdx.x; w < words; w += SC_THREADS) sm[w] = 0;
        __syncthreads();
        uint32_t* slice = table + (size_t)b * cells;
        const uint4* src = slots + (size_t)b * n_tiles * 2;
        for (int64_t t0 = threadIdx.x; t0 < n_tiles; t0 += 2 * SC_THREADS) {
            const int64_t t1 = t0 + SC_THREADS;
            const uint4 a0 = __ldcs(src + 2 * t0), c0 = __ldcs(src + 2 * t0 + 1);
            uint4 a1 = make_uint4(0, 0, 0, 0), c1 = a1;
            if (t1 < n_tiles) { a1 = __ldcs(src + 2 * t1); c1 = __ldcs(src + 2 * t1 + 1); }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint4 a = u ? a1 : a0, c = u ? c1 : c0;
                const uint32_t n = a.x & 0xFFFFu;
                if (n == 0) continue;
                bump_pair(sm, a.x, 0, n, slice); bump_pair(sm, a.y, 2, n, slice);
                bump_pair(sm, a.z, 4, n, slice); bump_pair(sm, a.w, 6, n, slice);
                if (n >= 8) {
                    bump_pair(sm, c.x, 8, n, slice); bump_pair(sm, c.y, 10, n, slice);
                    bump_pair(sm, c.z, 12, n, slice); bump_pair(sm, c.w, 14, n, slice);
                }
            }
        }
        __syncthreads();
        // slice += counters (the slice already holds the windows that found their sector full, and folded 0x8000s)
        uint4* out = reinterpret_cast<uint4*>(slice);
        for (int w = threadIdx.x; w < words / 2; w += SC_THREADS) {
            const uint2 p = reinterpret_cast<const uint2*>(sm)[w];
            uint4 o = __ldcg(out + w);
            o.x += p.x & 0xFFFFu; o.y += p.x >> 16; o.z += p.y & 0xFFFFu; o.w += p.y >> 16;
            out[w] = o;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int64_t kmap_slot_scratch_bytes(int64_t n, int k) {
    if (k < 12 || k > 14 || n < 0) return 0;
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_tiles = (n_words + SL_THREADS - 1) / SL_THREADS;
    return (int64_t)SL_BUCKETS * n_tiles * 32 + 256;
}

// table[h] += number of counted windows with key h (the caller zeroes the table first).  hide may be NULL.
// terminal_tabs (may be NULL): also add the run-end corrections of levels kmin..k-1 to those tables (count_all.cu).
// step_events (may be NULL): [0] recorded after the partition pass.
int kmap_count_slotted(const uint32_t* packed, const uint32_t* valid, const uint32_t* hide, int64_t n, int k, uint32_t* table,
                       void* scratch, const KmapTableSet* terminal_tabs, int kmin, void* const* step_events, cudaStream_t s) {
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_tiles = (n_words + SL_THREADS - 1) / SL_THREADS;
    uint4* slots = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    static bool attr_set = false;
    const int smem_p = SL_BUCKETS * 32 + SL_BUCKETS * 4;
    const int smem_c = (1 << (2 * k - 12)) * 2;
    if (!attr_set) {
        cudaFuncSetAttribute(slot_partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
        cudaFuncSetAttribute(slot_partition_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
        cudaFuncSetAttribute(slot_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_MAX_WORDS * 4);
        attr_set = true;
    }
    const unsigned int g = (unsigned int)(n_tiles < 148 ? n_tiles : 148);
    if (terminal_tabs)
        slot_partition_kernel<true><<<g, SL_THREADS, smem_p, s>>>(packed, valid, hide, n_words, n_tiles, k, slots, table, *terminal_tabs, kmin);
    else
        slot_partition_kernel<false><<<g, SL_THREADS, smem_p, s>>>(packed, valid, hide, n_words, n_tiles, k, slots, table, KmapTableSet(), k);
    if (step_events && step_events[0]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(step_events[0]), s);
    slot_count_kernel<<<148 * (smem_c <= 32768 ? 2 : 1), SC_THREADS, smem_c, s>>>(slots, n_tiles, k, table);
    return kmap_check_launch("count_slotted");
}

extern "C" int kmap_count_dense_slotted(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table,
                                        void* scratch, int64_t scratch_bytes, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 12 && k <= 14, "the slotted count covers 12 <= k <= 14");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && table && scratch, "null pointer");
    KMAP_REQUIRE(scratch_bytes >= kmap_slot_scratch_bytes(n, k), "scratch too small (kmap_slot_scratch_bytes)");
    return kmap_count_slotted(packed, valid, nullptr, n, k, table, scratch, nullptr, k, nullptr, as_stream(stream));
}
