// Level-k count by key partitioning: comp_kmer_hash_taichi + count_uniq_hash (kmer_count.py:449-491) for a table that
// does not fit in L2.  DESIGN.md section 4.5.
//
// Measured on B200: random RED.ADD into an L2-resident table sustains ~187 G updates/s (22 G/s when the table lives in
// HBM), shared-memory ATOMS ~2000 G/s.  So instead of one global atomic per window, every counted window's key is
// split into (bucket = key >> 16, suffix = key & 0xFFFF):
//   1. bucket_hist_kernel   one pass over the packed reads: windows per (tile of 32768 positions, bucket), with running
//                           sums -> after a small scan, the exact place of every tile's run in every bucket
//   2. partition_kernel     one pass: a CTA counting-sorts a tile by bucket in shared memory (one shared-memory atomic
//                           per window) and writes the 16-bit suffixes as runs (lanes of a warp store to consecutive
//                           addresses) at the precomputed places: no histogram of its own, no global atomics.  (Private per-CTA
//                           ranges without the atomics were measured: slower, 148 x 4096 open partial lines spill
//                           out of L2 and DRAM writes grow 2.5x.)
//   3. bucket_count_kernel  one CTA per bucket: 65536 cells as packed 16-bit counters in shared memory (128 KB),
//                           suffixes stream in with 128-bit loads, the table slice is written once, coalesced.
// The result is identical to kmap_count_dense (integer sums are order independent).
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tile.cuh"

namespace {

constexpr int PT_THREADS = KMAP_TILE_THREADS;
constexpr int PT_TILE_WORDS = KMAP_TILE_WORDS;    // validity words per tile, two per thread (tile.cuh)
constexpr int PT_TILE = PT_TILE_WORDS * 32;       // positions (= staged entries at most) per tile: 65 472
constexpr int PT_SLOTS = 65536;                   // entries of the staging buffer (PT_TILE rounded up)
constexpr int PT_MAX_BUCKETS = 4096;              // k <= 14
constexpr int PT_MAX_PER = PT_MAX_BUCKETS / PT_THREADS;
constexpr int PT_MAX_X = PT_MAX_BUCKETS / 4;      // buckets of the routed level k-1 (one per thread at most)
constexpr int PT_MAX_ALL = PT_MAX_BUCKETS + PT_MAX_X;
constexpr int PT_XCAP = PT_TILE / 14 + 2;         // routed entries a tile can hold (each needs a run end: >= 14 positions apart)
#ifndef KMAP_PT_WU
#define KMAP_PT_WU 4
#endif
constexpr int PT_WU = KMAP_PT_WU;                // write-out entries in flight per thread

__device__ __forceinline__ uint32_t base16_at(const TileWords& t, int i) {           // 16 bases from position i of the word
    return (i < 16) ? __funnelshift_l(t.w1, t.w0, 2 * i) : __funnelshift_l(t.w2, t.w1, 2 * (i - 16));
}

// extra_base: index of the first extra bucket (= n_buckets).  A routed correction is the (k-1)-mer at a position with
// exactly k-1 valid bases: bucket extra_base + (its key >> 16), suffix = its low 16 bits.
__device__ __forceinline__ void tile_hist(const TileWords& t, int sh, uint32_t* cnt, int extra_base) {
    if (t.fresh) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((t.fresh >> i) & 1u) atomicAdd(&cnt[base16_at(t, i) >> (sh + 16)], 1u);
    }
    if (t.corr) {                                   // about one per read
        uint32_t c = t.corr;
        do {
            const int i = __ffs(c) - 1;
            c &= c - 1;
            atomicAdd(&cnt[extra_base + (base16_at(t, i) >> (sh + 18))], 1u);
        } while (c);
    }
}

// ---- 1. windows per (tile, bucket) (+ the run-end corrections of the all-k count) ------------------------------------------
// CTA h walks the contiguous tile range [n_tiles*h/H, n_tiles*(h+1)/H).  For every tile it histograms the windows by
// bucket in shared memory and writes the row counts[tile][*] (uint16) together with off[tile][*] = the number of windows
// the EARLIER tiles of this range put into each bucket (a thread owns PER buckets; their running sums live in its
// registers).  chunk_total[h][*] is what the whole range holds.  After the scan below, the first entry tile t writes into
// bucket b is chunk_base[h(t)][b] + off[t][b]: exact, so the partition pass needs neither a histogram of its own nor
// cursor atomics, and the runs of consecutive tiles are adjacent in every bucket (partial sectors meet in L2).
// With TERMINAL, the pass also does what terminal_corrections_kernel (count_all.cu) does: "+1 at level v" for every
// window with exactly v valid bases (kmin <= v < k) in front of a run end.  Those are scattered global REDs; issued from
// this kernel they overlap its shared-memory-bound histogram work instead of costing passes of their own.
constexpr int PT_HGRID = 296;                     // CTAs of the histogram pass

__host__ __device__ __forceinline__ int64_t chunk_first_tile(int64_t n_tiles, int h) { return n_tiles * h / PT_HGRID; }

template <bool TERMINAL, int PER>
__global__ void __launch_bounds__(PT_THREADS) bucket_hist_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                 const uint32_t* __restrict__ hide, int64_t n_words, int64_t n_tiles,
                                                                 int k, int n_buckets, int n_extra, uint16_t* __restrict__ counts,
                                                                 uint32_t* __restrict__ off, uint32_t* __restrict__ chunk_total,
                                                                 KmapTableSet tabs, int kmin, int kcorr, int single_from) {
    __shared__ __align__(16) uint32_t cnt2[2][PT_MAX_ALL];          // double buffer: one barrier per tile
    __shared__ uint32_t* stab[16];
    if (TERMINAL && threadIdx.x < 16) stab[threadIdx.x] = tabs.t[threadIdx.x];
    for (int b = threadIdx.x; b < 2 * PT_MAX_ALL; b += PT_THREADS) (&cnt2[0][0])[b] = 0;
    __syncthreads();
    const int sh = 32 - 2 * k;
    const int n_all = n_buckets + n_extra;                          // row length of counts / off / chunk_total
    const bool route = n_extra > 0;
    const int64_t t0 = chunk_first_tile(n_tiles, blockIdx.x), t1 = chunk_first_tile(n_tiles, blockIdx.x + 1);
    const int b0 = PER * threadIdx.x;                              // this thread's buckets: b0 .. b0 + PER - 1
    const int bx = n_buckets + threadIdx.x;                        // ... and, when routing, extra bucket bx
    const bool has_x = (int)threadIdx.x < n_extra;
    uint32_t run[PER];
    uint32_t runx = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) run[j] = 0;
    RawPair nxt = load_raw_pair(packed, valid, hide, n_words, t0 < t1 ? t0 : n_tiles);
    RawPrev nxt_prev;
    nxt_prev.vp = nxt_prev.hp = nxt_prev.wp = 0;
    if (TERMINAL) nxt_prev = load_raw_prev(packed, valid, hide, n_words, t0 < t1 ? t0 : n_tiles);
    for (int64_t tile = t0; tile < t1; ++tile) {
        uint32_t* cnt = cnt2[(tile - t0) & 1];
        const RawPair r = nxt;
        const RawPrev rp = nxt_prev;
        nxt = load_raw_pair(packed, valid, hide, n_words, tile + 1 < t1 ? tile + 1 : n_tiles);       // (past the end: zeros)
        if (TERMINAL) nxt_prev = load_raw_prev(packed, valid, hide, n_words, tile + 1 < t1 ? tile + 1 : n_tiles);
        const RawWords ra = first_word(r), rb = second_word(r);
        tile_hist(cook(ra, k, route), sh, cnt, n_buckets);
        tile_hist(cook(rb, k, route), sh, cnt, n_buckets);
        if (TERMINAL) {                                             // levels kmin .. kcorr-1
            run_end_corrections(ra, rp, kmin, kcorr, stab, single_from);
            RawPrev rq;
            rq.vp = r.v0; rq.hp = r.h0; rq.wp = r.w1;
            run_end_corrections(rb, rq, kmin, kcorr, stab, single_from);
        }
        __syncthreads();       // this tile's counts are complete; the other buffer was zeroed before the previous barrier
        if (b0 < n_buckets) {
            uint32_t c[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) { c[j] = cnt[b0 + j]; cnt[b0 + j] = 0; }
            uint16_t* crow = counts + (size_t)tile * n_all + b0;
            uint32_t* orow = off + (size_t)tile * n_all + b0;
            if (PER == 4) {
                __stcs(reinterpret_cast<uint2*>(crow), make_uint2(c[0] | (c[1 % PER] << 16), c[2 % PER] | (c[3 % PER] << 16)));
                __stcs(reinterpret_cast<uint4*>(orow), make_uint4(run[0], run[1 % PER], run[2 % PER], run[3 % PER]));
            } else {
#pragma unroll
                for (int j = 0; j < PER; ++j) { crow[j] = (uint16_t)c[j]; orow[j] = run[j]; }
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) run[j] += c[j];
        }
        if (has_x) {
            const uint32_t cx = cnt[bx];
            cnt[bx] = 0;
            counts[(size_t)tile * n_all + bx] = (uint16_t)cx;
            off[(size_t)tile * n_all + bx] = runx;
            runx += cx;
        }
    }
    if (b0 < n_buckets) {
#pragma unroll
        for (int j = 0; j < PER; ++j) chunk_total[(size_t)blockIdx.x * n_all + b0 + j] = run[j];
    }
    if (has_x) chunk_total[(size_t)blockIdx.x * n_all + bx] = runx;
}

// block-wide exclusive scan of one value per thread (two barriers inside)
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

// base[b] = first entry of bucket b (base[n_buckets] = total), chunk_base[h][b] = base[b] + what the ranges before h put
// into bucket b.  Three small launches: column sums (one thread per bucket), a one-block scan of the totals, column
// prefixes.
__global__ void __launch_bounds__(256) bucket_total_kernel(const uint32_t* __restrict__ chunk_total, int n_buckets,
                                                           unsigned long long* __restrict__ total) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b >= n_buckets) return;
    unsigned long long t = 0;
    for (int h = 0; h < PT_HGRID; ++h) t += chunk_total[(size_t)h * n_buckets + b];
    total[b] = t;
}

__global__ void __launch_bounds__(PT_THREADS) bucket_scan_kernel(const unsigned long long* __restrict__ total, int n_buckets,
                                                                 unsigned long long* __restrict__ base) {
    __shared__ unsigned long long warp_sums[32];
    const int per = (n_buckets + PT_THREADS - 1) / PT_THREADS;       // <= PT_MAX_PER + 1
    const int lo = threadIdx.x * per, hi = min(lo + per, n_buckets);
    unsigned long long mine = 0;
    for (int b = lo; b < hi; ++b) mine += total[b];
    unsigned long long run = block_scan_excl<unsigned long long>(mine, warp_sums);
    if (threadIdx.x == PT_THREADS - 1) base[n_buckets] = run + mine;
    for (int b = lo; b < hi; ++b) { base[b] = run; run += total[b]; }
}

__global__ void __launch_bounds__(256) chunk_base_kernel(const uint32_t* __restrict__ chunk_total, int n_buckets,
                                                         const unsigned long long* __restrict__ base,
                                                         unsigned long long* __restrict__ chunk_base) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b >= n_buckets) return;
    unsigned long long r = base[b];
    for (int h = 0; h < PT_HGRID; ++h) {
        chunk_base[(size_t)h * n_buckets + b] = r;
        r += chunk_total[(size_t)h * n_buckets + b];
    }
}

// ---- 2. partition -------------------------------------------------------------------------------------------------------
// Per tile of 65 472 positions (two validity words per thread): (b) exclusive scan of the tile's bucket counts (read from
// pass 1) -> tile-local starts, (d) counting-sort scatter of the 16-bit key suffixes into bucket order in shared memory
// (one shared-memory atomic per window), (e) write-out: entry i of the sorted tile goes to gd[run of i] + i, so consecutive
// lanes write consecutive addresses inside a run.  The staging buffer holds ONLY the suffixes (2 bytes per entry, which is
// what lets a tile be 64 K positions: a (tile, bucket) run is then ~14 suffixes = 28 bytes, about one 32-byte sector, where
// the 32 K-position tiles of the first version wrote runs of ~7 = 7.4 sectors per store request, and the write-out was 19
// of the kernel's 32.5 ms); the run an entry belongs to is recovered from a bitmap of run starts: run(i) = F[i / 32] +
// popc(heads[i / 32] & bits 0..i%32), F[w] = the run that contains entry 32w - 1, and gd[] is indexed by run (non-empty
// buckets in bucket order).  Tiles are dealt by an atomic ticket, so the tiles in flight are neighbours and so are their
// runs in every bucket.  The raw words and the count / offset rows of the next tile are requested a tile ahead.
// Routed corrections (n_extra > 0): the windows with exactly k-1 valid bases go to buckets n_buckets .. n_buckets + n_extra - 1
// (bucket = n_buckets + the top bits of the (k-1)-mer, suffix = its low 16 bits).  In shared memory they are laid out from
// the END of the staging buffer downwards, the ordinary entries from the start upwards (together at most one entry per
// position), so that neither layout needs the other's total; they are few (about one per read), so each carries its bucket
// in a side array instead of taking part in the run bitmap.
template <int PER>
struct TileRow { uint32_t c[PER]; uint32_t o[PER]; uint32_t cx, ox; };        // cx, ox: the thread's extra bucket (routed level k-1)

template <int PER>
__device__ __forceinline__ TileRow<PER> load_tile_row(const uint16_t* __restrict__ counts, const uint32_t* __restrict__ off,
                                                     int64_t n_tiles, int64_t tile, int n_buckets, int n_extra) {
    TileRow<PER> r;
#pragma unroll
    for (int j = 0; j < PER; ++j) r.c[j] = r.o[j] = 0;
    r.cx = r.ox = 0;
    const int b0 = PER * threadIdx.x;
    const int n_all = n_buckets + n_extra;
    if (tile < n_tiles && b0 < n_buckets) {
        const uint16_t* crow = counts + (size_t)tile * n_all + b0;
        const uint32_t* orow = off + (size_t)tile * n_all + b0;
        if (PER == 4) {
            const uint2 c = __ldcs(reinterpret_cast<const uint2*>(crow));
            const uint4 o = __ldcs(reinterpret_cast<const uint4*>(orow));
            r.c[0] = c.x & 0xFFFFu; r.c[1 % PER] = c.x >> 16; r.c[2 % PER] = c.y & 0xFFFFu; r.c[3 % PER] = c.y >> 16;
            r.o[0] = o.x; r.o[1 % PER] = o.y; r.o[2 % PER] = o.z; r.o[3 % PER] = o.w;
        } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) { r.c[j] = __ldcs(crow + j); r.o[j] = __ldcs(orow + j); }
        }
    }
    if (tile < n_tiles && (int)threadIdx.x < n_extra) {
        r.cx = __ldcs(counts + (size_t)tile * n_all + n_buckets + threadIdx.x);
        r.ox = __ldcs(off + (size_t)tile * n_all + n_buckets + threadIdx.x);
    }
    return r;
}

// shared memory of partition_kernel (dynamic): staging buffer, per-bucket running offsets, per-run destinations, run-start
// bitmap (double buffered) + F, bucket of every routed entry
constexpr int PT_SMEM = PT_SLOTS * 2 + PT_MAX_ALL * 4 + PT_MAX_ALL * 8 + 2 * (PT_SLOTS / 32) * 4 + (PT_SLOTS / 32 + 2) * 2 + PT_XCAP * 2 + 64;

__device__ __forceinline__ void scatter_word(const TileWords& t, int sh, int n_buckets, uint32_t* off, uint16_t* sorted, uint16_t* xbucket) {
    if (t.fresh) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((t.fresh >> i) & 1u) {
                const uint32_t key = base16_at(t, i) >> sh;
                sorted[atomicAdd(&off[key >> 16], 1u)] = (uint16_t)key;
            }
    }
    if (t.corr) {
        uint32_t c = t.corr;
        do {
            const int i = __ffs(c) - 1;
            c &= c - 1;
            const uint32_t key = base16_at(t, i) >> (sh + 2);             // the (k-1)-mer
            const uint32_t pos = atomicAdd(&off[n_buckets + (key >> 16)], 1u);
            sorted[pos] = (uint16_t)key;
            xbucket[PT_SLOTS - 1 - pos] = (uint16_t)(key >> 16);
        } while (c);
    }
}

template <int PER>       // buckets per thread in the scan step: n_buckets <= PER * PT_THREADS
__global__ void __launch_bounds__(PT_THREADS, 1) partition_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                  const uint32_t* __restrict__ hide, int64_t n_words, int64_t n_tiles,
                                                                  int k, int n_buckets, int n_extra, const uint16_t* __restrict__ counts,
                                                                  const uint32_t* __restrict__ off_rows,
                                                                  const unsigned long long* __restrict__ chunk_base,
                                                                  uint16_t* __restrict__ suffixes, unsigned long long* __restrict__ ticket) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint16_t* sorted = reinterpret_cast<uint16_t*>(smem_raw);                                  // PT_SLOTS suffixes, grouped by bucket
    unsigned long long* gd = reinterpret_cast<unsigned long long*>(sorted + PT_SLOTS);        // global start - tile start: per RUN, then per extra bucket
    uint32_t* off = reinterpret_cast<uint32_t*>(gd + PT_MAX_ALL);                             // running tile offsets per bucket
    uint32_t* heads2 = off + PT_MAX_ALL;                                                      // 2 x run-start bitmap
    uint16_t* F = reinterpret_cast<uint16_t*>(heads2 + 2 * (PT_SLOTS / 32));                  // run that contains entry 32w - 1
    uint16_t* xbucket = F + (PT_SLOTS / 32 + 2);                                              // bucket of routed entry (from the end)
    __shared__ unsigned long long warp_sums[32];
    __shared__ uint32_t tile_total, tile_total_x;
    const int sh = 32 - 2 * k;
    const int b0 = PER * threadIdx.x;
    const int n_all = n_buckets + n_extra;
    const bool route = n_extra > 0;
    const bool has_x = (int)threadIdx.x < n_extra;
    const int bx = n_buckets + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // Tiles are handed out in order by an atomic ticket: the tiles in flight are then always the ~148 most recent ones,
    // so a run written into a bucket gets its neighbours (the runs of the adjacent tiles) within a tile time and the
    // partial lines complete in L2.  (A static round-robin lets the CTAs drift apart: 4x the DRAM writes, measured.)
    __shared__ long long s_ticket[2];
    if (threadIdx.x == 0) {
        s_ticket[0] = (long long)atomicAdd(ticket, 1ull);
        s_ticket[1] = (long long)atomicAdd(ticket, 1ull);
    }
    for (int w = threadIdx.x; w < 2 * (PT_SLOTS / 32); w += PT_THREADS) heads2[w] = 0;
    __syncthreads();
    int64_t tile = s_ticket[0], tile_next = s_ticket[1];
    __syncthreads();
    RawPair raw = load_raw_pair(packed, valid, hide, n_words, tile);
    TileRow<PER> row = load_tile_row<PER>(counts, off_rows, n_tiles, tile, n_buckets, n_extra);
    int h = -1;                                        // range of pass 1 that holds the current tile
    unsigned long long cb[PER];
    unsigned long long cbx = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) cb[j] = 0;
    int parity = 0;
    while (tile < n_tiles) {
        if (threadIdx.x == 0) s_ticket[0] = (long long)atomicAdd(ticket, 1ull);       // the tile after the next one
        const TileWords cur_a = cook(first_word(raw), k, route), cur_b = cook(second_word(raw), k, route);
        const TileRow<PER> tr = row;
        raw = load_raw_pair(packed, valid, hide, n_words, tile_next);
        row = load_tile_row<PER>(counts, off_rows, n_tiles, tile_next, n_buckets, n_extra);
        int hh = (int)((tile * PT_HGRID) / n_tiles);
        while (hh + 1 < PT_HGRID && chunk_first_tile(n_tiles, hh + 1) <= tile) ++hh;
        while (hh > 0 && chunk_first_tile(n_tiles, hh) > tile) --hh;
        if (hh != h) {                                 // (block-uniform) a new range: fetch its bucket bases
            h = hh;
#pragma unroll
            for (int j = 0; j < PER; ++j) cb[j] = b0 + j < n_buckets ? __ldg(chunk_base + (size_t)h * n_all + b0 + j) : 0ull;
            cbx = has_x ? __ldg(chunk_base + (size_t)h * n_all + bx) : 0ull;
        }
        // (b) exclusive scan over the buckets; thread owns buckets [PER*tid, PER*tid + PER) and extra bucket n_buckets + tid.
        // One 64-bit scan carries three running sums: ordinary entries (bits 0..19), routed entries (20..39), non-empty
        // ordinary buckets = runs (40..59)
        uint32_t mine = 0, mine_runs = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { mine += tr.c[j]; mine_runs += tr.c[j] ? 1u : 0u; }
        const unsigned long long both = (unsigned long long)mine | ((unsigned long long)tr.cx << 20) | ((unsigned long long)mine_runs << 40);
        const unsigned long long excl = block_scan_excl<unsigned long long>(both, warp_sums);   // (its barriers also fence the previous write-out)
        uint32_t run = (uint32_t)excl & 0xFFFFFu;
        const uint32_t runx = (uint32_t)(excl >> 20) & 0xFFFFFu;
        uint32_t rid = (uint32_t)(excl >> 40);
        const int64_t tile_after = s_ticket[0];                      // (thread 0 writes it again only after two more barriers)
        if (threadIdx.x == PT_THREADS - 1) { tile_total = run + mine; tile_total_x = runx + tr.cx; }
        uint32_t* heads = heads2 + parity * (PT_SLOTS / 32);         // (zeroed while the previous tile was scattered)
        if (threadIdx.x == 0) F[0] = 0xFFFFu;                        // "run -1" precedes entry 0
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (b0 + j < n_buckets) off[b0 + j] = run;
            const uint32_t c = tr.c[j];
            if (c) {
                const uint32_t end = run + c;
                gd[rid] = cb[j] + tr.o[j] - run;
                atomicOr(&heads[run >> 5], 1u << (run & 31u));
                for (uint32_t w = (run >> 5) + 1; w <= (end >> 5); ++w) F[w] = (uint16_t)rid;      // entry 32w - 1 lies in this run
                ++rid;
                run = end;
            }
        }
        if (has_x) {
            const uint32_t start = (uint32_t)PT_SLOTS - runx - tr.cx;  // this bucket's entries: sorted[start .. start + cx)
            off[bx] = start;
            gd[PT_MAX_BUCKETS + threadIdx.x] = cbx + tr.ox - start;
        }
        __syncthreads();
        {   // the bitmap of the next tile: nobody reads it any more (its last readers wrote out the previous tile)
            uint32_t* hn = heads2 + (parity ^ 1) * (PT_SLOTS / 32);
            hn[2 * threadIdx.x] = 0; hn[2 * threadIdx.x + 1] = 0;
        }
        // (d) scatter the suffixes into bucket order
        scatter_word(cur_a, sh, n_buckets, off, sorted, xbucket);
        scatter_word(cur_b, sh, n_buckets, off, sorted, xbucket);
        __syncthreads();
        const uint32_t total = tile_total;
        // (e) write-out.  Measured (skipping phases): loads + scan 3.6 ms, scatter 10.0 ms, write-out 15.2 ms of 28.8 ms at 1e8
        // reads.  The write-out costs ~4.7 cycles per (store instruction, 128-byte line) pair -- 32 consecutive entries touch
        // ~3.3 runs -- so its cost follows the number of runs, not bytes or instructions: a variant in which a lane took two
        // neighbouring entries (32-bit store when aligned, else two 16-bit stores) issued fewer instructions but visited every
        // line from up to three of them and took 38 ms.
        for (uint32_t i0 = threadIdx.x; i0 < total; i0 += PT_WU * PT_THREADS) {
            uint32_t sfx[PT_WU], rr[PT_WU];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) {
                const uint32_t i = i0 + u * PT_THREADS;              // (i >> 5 is the same for the lanes of a warp)
                const uint32_t w = (i >> 5) & (PT_SLOTS / 32 - 1);   // (an index beyond `total` stays inside the arrays)
                sfx[u] = sorted[i & (PT_SLOTS - 1)];
                rr[u] = ((uint32_t)F[w] + __popc(heads[w] & (0xFFFFFFFFu >> (31 - lane)))) & 0xFFFFu;
            }
            unsigned long long d[PT_WU];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) d[u] = gd[rr[u] & (PT_MAX_BUCKETS - 1)];
#pragma unroll
            for (int u = 0; u < PT_WU; ++u) {
                const uint32_t i = i0 + u * PT_THREADS;
                if (i < total) suffixes[d[u] + i] = (uint16_t)sfx[u];
            }
        }
        if (route) {
            for (uint32_t i = (uint32_t)PT_SLOTS - tile_total_x + threadIdx.x; i < (uint32_t)PT_SLOTS; i += PT_THREADS)
                suffixes[gd[PT_MAX_BUCKETS + xbucket[PT_SLOTS - 1 - i]] + i] = sorted[i];
        }
        tile = tile_next;
        tile_next = tile_after;
        parity ^= 1;
    }
}

// ---- 3. per-bucket count in shared memory --------------------------------------------------------------------------------
constexpr int BC_THREADS = 1024;
constexpr int BC_CELLS = 65536;
constexpr int BC_WORDS = BC_CELLS / 2;            // two 16-bit counters per word
constexpr int BC_QUEUES = 8;                      // per-bucket count launches of one call (ranges of a sharded count)
#ifndef KMAP_BC_UNROLL
#define KMAP_BC_UNROLL 4
#endif
constexpr int BC_UNROLL = KMAP_BC_UNROLL;         // 128-bit loads in flight per thread

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

// One segment [lo, hi) of a bucket's suffixes -> the bucket's slice of the table.  shared_slice: other CTAs add to the same
// slice right now (a bucket cut into segments), so the counters leave through atomics; else this CTA is the only writer:
// plain stores into a zeroed slice, or (add_mode: routed level k-1, whose slice already carries the -1 corrections of the
// per-read scan) a plain read-modify-write.
__device__ __forceinline__ void count_segment(const uint16_t* __restrict__ suffixes, unsigned long long lo, unsigned long long hi,
                                              uint32_t* __restrict__ slice, bool add_mode, bool shared_slice, uint32_t* sm, int* spilled) {
    for (int w = threadIdx.x; w < BC_WORDS; w += BC_THREADS) sm[w] = 0;
    if (threadIdx.x == 0) *spilled = 0;
    __syncthreads();
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
    if (my_spill) *spilled = 1;
    __syncthreads();
    const bool atomics = shared_slice || *spilled;           // (folded 0x8000s already sit in some cells: add on top)
    if (atomics) {
        for (int w = threadIdx.x; w < BC_WORDS; w += BC_THREADS) {
            const uint32_t p = sm[w];
            if (p & 0xFFFFu) atomicAdd(slice + 2 * w, p & 0xFFFFu);
            if (p >> 16) atomicAdd(slice + 2 * w + 1, p >> 16);
        }
    } else if (add_mode) {
        uint4* out = reinterpret_cast<uint4*>(slice);
        for (int w = threadIdx.x; w < BC_WORDS / 2; w += BC_THREADS) {
            const uint2 p = reinterpret_cast<const uint2*>(sm)[w];
            if (p.x | p.y) {
                uint4 v4 = out[w];
                v4.x += p.x & 0xFFFFu; v4.y += p.x >> 16; v4.z += p.y & 0xFFFFu; v4.w += p.y >> 16;
                out[w] = v4;
            }
        }
    } else {
        uint4* out = reinterpret_cast<uint4*>(slice);
        for (int w = threadIdx.x; w < BC_WORDS / 2; w += BC_THREADS) {
            const uint2 p = reinterpret_cast<const uint2*>(sm)[w];
            out[w] = make_uint4(p.x & 0xFFFFu, p.x >> 16, p.y & 0xFFFFu, p.y >> 16);
        }
    }
    __syncthreads();
}

// Buckets n_buckets .. n_all-1 hold the routed run-end corrections of level k-1: their counters are ADDED to the slice of
// `lower` (the level k-1 table).  A bucket with more than seg_len suffixes (a heavy hitter: the cells of a planted motif
// collect millions of windows, and same-address shared-memory atomics serialise) is cut into segments: this kernel counts
// the first one and queues the others for bucket_segments_kernel, so that one bucket never holds up a whole launch
// (measured: the bucket of the planted 14-mer of cfg3 took 16x the time of an average one).
// queue[0] = items pushed, queue[1] = items taken, queue[2..] = items (bucket | segment << 13).
__global__ void __launch_bounds__(BC_THREADS, 1) bucket_count_kernel(const uint16_t* __restrict__ suffixes,
                                                                     const unsigned long long* __restrict__ base, int n_buckets,
                                                                     int b_lo, int b_hi, uint32_t* __restrict__ table, uint32_t* __restrict__ lower,
                                                                     unsigned long long seg_len, uint32_t* __restrict__ queue) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ int spilled;
    for (int b = b_lo + blockIdx.x; b < b_hi; b += gridDim.x) {
        const bool add_mode = b >= n_buckets;
        uint32_t* slice = add_mode ? lower + (size_t)(b - n_buckets) * BC_CELLS : table + (size_t)b * BC_CELLS;
        const unsigned long long lo = base[b], hi = base[b + 1];
        const unsigned long long n_seg = hi > lo ? (hi - lo + seg_len - 1) / seg_len : 1;
        if (n_seg > 1 && threadIdx.x == 0) {
            const uint32_t at = atomicAdd(&queue[0], (uint32_t)(n_seg - 1));
            for (uint32_t j = 1; j < (uint32_t)n_seg; ++j) queue[2 + at + j - 1] = (uint32_t)b | (j << 13);
        }
        count_segment(suffixes, lo, n_seg > 1 ? lo + seg_len : hi, slice, add_mode, n_seg > 1, sm, &spilled);
    }
}

__global__ void __launch_bounds__(BC_THREADS, 1) bucket_segments_kernel(const uint16_t* __restrict__ suffixes,
                                                                        const unsigned long long* __restrict__ base, int n_buckets,
                                                                        uint32_t* __restrict__ table, uint32_t* __restrict__ lower,
                                                                        unsigned long long seg_len, uint32_t* __restrict__ queue) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ int spilled;
    __shared__ uint32_t item;
    const uint32_t n_items = queue[0];                      // (complete: the kernel that pushes them has finished)
    for (;;) {
        if (threadIdx.x == 0) item = atomicAdd(&queue[1], 1u);
        __syncthreads();
        const uint32_t it = item;
        __syncthreads();
        if (it >= n_items) return;
        const uint32_t e = queue[2 + it];
        const int b = (int)(e & 0x1FFFu);
        const unsigned long long j = e >> 13;
        const bool add_mode = b >= n_buckets;
        uint32_t* slice = add_mode ? lower + (size_t)(b - n_buckets) * BC_CELLS : table + (size_t)b * BC_CELLS;
        const unsigned long long lo = base[b] + j * seg_len, end = base[b + 1];
        count_segment(suffixes, lo, lo + seg_len < end ? lo + seg_len : end, slice, add_mode, true, sm, &spilled);
    }
}

// One level of the folded run-end corrections: T_v += C_v and, unless v is the lowest level, C_(v-1) += the 4:1 reduction of C_v
// over the FIRST base (cells h, q + h, 2q + h, 3q + h with q = 4^(v-1)).  One thread = 4 consecutive cells of every quarter.
__global__ void __launch_bounds__(256) fold_corrections_kernel(const uint32_t* __restrict__ cv, uint32_t* __restrict__ tv,
                                                               uint32_t* __restrict__ clow, int v) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    auto add4 = [](uint4 a, const uint4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; };
    if (!clow) {
        const int64_t n4 = (((int64_t)1 << (2 * v)) + 3) / 4;             // (4^v >= 4 for v >= 1)
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
            if (((int64_t)1 << (2 * v)) < 4) { if (i == 0) tv[0] += cv[0]; continue; }
            reinterpret_cast<uint4*>(tv)[i] = add4(reinterpret_cast<const uint4*>(tv)[i], reinterpret_cast<const uint4*>(cv)[i]);
        }
        return;
    }
    const int64_t q = (int64_t)1 << (2 * (v - 1));
    if (q < 4) {                                                          // v = 1: four cells
        if (blockIdx.x == 0 && threadIdx.x == 0) { uint32_t sum = 0; for (int b = 0; b < 4; ++b) { tv[b] += cv[b]; sum += cv[b]; } clow[0] += sum; }
        return;
    }
    const int64_t q4 = q / 4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < q4; i += stride) {
        uint4 sum = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint4 c = reinterpret_cast<const uint4*>(cv)[b * q4 + i];
            reinterpret_cast<uint4*>(tv)[b * q4 + i] = add4(reinterpret_cast<const uint4*>(tv)[b * q4 + i], c);
            sum = add4(sum, c);
        }
        reinterpret_cast<uint4*>(clow)[i] = add4(reinterpret_cast<const uint4*>(clow)[i], sum);
    }
}

// KMAP_FOLD_RUN_ENDS (read once): the folded run-end corrections of kmap_count_partitioned -- 0 = scattered REDs per level from
// the histogram pass (round 1), 1 = folded on a single GPU, 2 (default) = also under the one-exchange sharded count
static int kmap_fold_run_ends() {
    static const int mode = [] { const char* e = getenv("KMAP_FOLD_RUN_ENDS"); return e ? atoi(e) : 2; }();
    return mode;
}

struct PartScratch {
    uint32_t* chunk_total;            // [PT_HGRID][n_all]
    unsigned long long* base;         // [n_all + 1]
    unsigned long long* chunk_base;   // [PT_HGRID][n_all]
    uint16_t* counts;                 // [n_tiles][n_all]
    uint32_t* off;                    // [n_tiles][n_all]
    uint16_t* suffixes;               // one per counted window (+ one per routed correction: never the same position)
    unsigned long long* ticket;       // tile dispenser of the partition pass
    unsigned long long* total;        // [n_all]
    uint32_t* queues;                 // BC_QUEUES x (2 + n_all) words: segment queues of the per-bucket count launches
    uint32_t* corr;                   // folded run-end corrections: level v (< k) at cell offset corr_offset(v)
};
// (levels back to back, every one starting at a multiple of 4 cells: 128-bit accesses)
static int64_t corr_offset(int v) { return ((((int64_t)1 << (2 * v)) - 1) / 3 + 4 * v + 3) & ~(int64_t)3; }

static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// sized for n_buckets + n_buckets / 4 rows entries whether or not the corrections of level k-1 are routed
static int64_t carve(void* scratch, int n_buckets, int64_t n, PartScratch* p) {
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_tiles = (n_words + PT_TILE_WORDS - 1) / PT_TILE_WORDS;
    const int64_t n_all = n_buckets + n_buckets / 4;
    uint8_t* q = reinterpret_cast<uint8_t*>(scratch);
    int64_t o = 0;
    auto take = [&](int64_t bytes) { uint8_t* r = q ? q + o : nullptr; o += align_up(bytes, 256); return r; };
    uint8_t* a0 = take((int64_t)PT_HGRID * n_all * 4);
    uint8_t* a1 = take((int64_t)(n_all + 1) * 8);
    uint8_t* a2 = take((int64_t)PT_HGRID * n_all * 8);
    uint8_t* a3 = take(n_tiles * n_all * 2);
    uint8_t* a4 = take(n_tiles * n_all * 4);
    uint8_t* a5 = take(2 * n + 64);
    uint8_t* a6 = take(8);
    uint8_t* a7 = take((int64_t)n_all * 8);
    uint8_t* a8 = take((int64_t)BC_QUEUES * (2 + n_all) * 4);
    uint8_t* a9 = take((((int64_t)n_buckets << 16) / 3 + 128) * 4);
    if (p) {
        p->corr = reinterpret_cast<uint32_t*>(a9);
        p->chunk_total = reinterpret_cast<uint32_t*>(a0);
        p->base = reinterpret_cast<unsigned long long*>(a1);
        p->chunk_base = reinterpret_cast<unsigned long long*>(a2);
        p->counts = reinterpret_cast<uint16_t*>(a3);
        p->off = reinterpret_cast<uint32_t*>(a4);
        p->suffixes = reinterpret_cast<uint16_t*>(a5);
        p->ticket = reinterpret_cast<unsigned long long*>(a6);
        p->total = reinterpret_cast<unsigned long long*>(a7);
        p->queues = reinterpret_cast<uint32_t*>(a8);
    }
    return o + 256;
}

}  // namespace

extern "C" int64_t kmap_partition_scratch_bytes(int64_t n, int k) {
    if (k < 9 || k > 14 || n < 0) return 0;
    return carve(nullptr, 1 << (2 * (k - 8)), n, nullptr);
}

// table[h] = number of counted windows with key h, for every h (the slice of every bucket is overwritten or, where
// folded counters were spilled, added to: the caller zeroes the table first).  hide may be NULL.
// terminal_tabs (may be NULL): also add the run-end corrections of levels kmin..k-1 to those tables (count_all.cu): scattered
// REDs issued from the histogram pass for the levels whose tables are L2 resident; when the level k-1 table is beyond L2
// (k = 14: 256 MB) its corrections are ROUTED -- they travel through the partition as n_buckets / 4 extra buckets and are
// added to the table by the per-bucket count (a scattered RED into a DRAM-resident table costs a sector read-modify-write:
// 22 G/s against 187 G/s in L2, profiles/r01_red_rate_microbench.txt; measured 4.7 ms of the 13.6 ms histogram pass).
int kmap_count_partitioned(const uint32_t* packed, const uint32_t* valid, const uint32_t* hide, int64_t n, int k, uint32_t* table,
                           void* scratch, const KmapTableSet* terminal_tabs, int kmin, void* const* step_events, cudaStream_t s,
                           const KmapMerge* merge) {
    const int n_buckets = 1 << (2 * (k - 8));
    const bool route = terminal_tabs && k - 1 >= kmin && k - 1 > 12;
    const int n_extra = route ? n_buckets / 4 : 0;
    const int n_all = n_buckets + n_extra;
    const int kcorr = route ? k - 1 : k;                // the fused REDs cover levels kmin .. kcorr-1
    PartScratch p;
    carve(scratch, n_buckets, n, &p);
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_tiles = (n_words + PT_TILE_WORDS - 1) / PT_TILE_WORDS;
    const KmapTableSet none = KmapTableSet();
    const KmapTableSet& tt = terminal_tabs ? *terminal_tabs : none;
    const int km = terminal_tabs ? kmin : k;
    const bool local = n > 0;                           // (an empty shard of a sharded count: collectives only)
    // Peer-memory exchange (peer.cu) with every level in one buffer, largest first (engine.alloc_tables): ONE exchange of the
    // whole buffer after the count -- 7 % more cells than levels k and k-1 alone, and no exchange kernels waiting for SMs
    // while the partition pass holds them all.
    int64_t span_all = (int64_t)n_buckets * 65536;
    bool one_exchange = false;
    if (merge && merge->comm && !merge->scatter && terminal_tabs && kmap_merge_on_peer_memory(table, span_all, merge)) {
        int u = k - 1;
        while (u >= kmin && tt.t[u] == table + span_all) { span_all += (int64_t)1 << (2 * u); --u; }
        one_exchange = u < kmin && kmap_merge_on_peer_memory(table, span_all, merge);
    }
    // Folded run-end corrections (single GPU): a run of r valid bases owes "+1 at level v" to the v-mer in front of its end for
    // every v <= min(r, k-1): five scattered REDs per read for k = 8..14.  The v-mer is the (v+1)-mer minus its FIRST base, so
    // C_v = (4:1 reduction of C_(v+1) over the first base) + the runs of exactly v bases: each run is entered ONCE, at the level
    // of its length (the routed level k-1 entries are exactly the runs of k-1 or more bases), into correction tables of their
    // own, and fold_corrections_kernel brings them down level by level and adds them to the count tables: streaming passes
    // over 4^(k-1) cells instead of (k-1-kmin) REDs per read.
    // (KMAP_FOLD_RUN_ENDS: 1 = single GPU, 2 = also the sharded count whose tables are merged by ONE exchange after the count --
    //  the corrections must be complete before they are exchanged, which the early merges of the NCCL path do not wait for)
    const bool fold = terminal_tabs && local && k - 1 >= kmin && ((!merge || !merge->comm) ? kmap_fold_run_ends() >= 1 : (kmap_fold_run_ends() >= 2 && one_exchange));
    KmapTableSet ctabs = KmapTableSet();
    if (fold) {
        for (int v = kmin; v < k; ++v) ctabs.t[v] = p.corr + corr_offset(v);
        cudaMemsetAsync(ctabs.t[kmin], 0, (size_t)(corr_offset(k) - corr_offset(kmin)) * 4, s);
    }
    const KmapTableSet& ht = fold ? ctabs : tt;        // where the histogram pass sends its run-end updates
    const int single_from = fold ? (route ? k - 1 : (1 << 20)) : 0;   // (not routed: no run is long enough to be skipped)
    if (local) {
        if (n_buckets <= PT_THREADS) {
            if (terminal_tabs) bucket_hist_kernel<true, 1><<<PT_HGRID, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_total, ht, km, kcorr, single_from);
            else bucket_hist_kernel<false, 1><<<PT_HGRID, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_total, ht, km, kcorr, single_from);
        } else {
            if (terminal_tabs) bucket_hist_kernel<true, PT_MAX_PER><<<PT_HGRID, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_total, ht, km, kcorr, single_from);
            else bucket_hist_kernel<false, PT_MAX_PER><<<PT_HGRID, PT_THREADS, 0, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_total, ht, km, kcorr, single_from);
        }
        bucket_total_kernel<<<(n_all + 255) / 256, 256, 0, s>>>(p.chunk_total, n_all, p.total);
        bucket_scan_kernel<<<1, PT_THREADS, 0, s>>>(p.total, n_all, p.base);
        chunk_base_kernel<<<(n_all + 255) / 256, 256, 0, s>>>(p.chunk_total, n_all, p.base, p.chunk_base);
        if (step_events && step_events[0]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(step_events[0]), s);
    }
    if (merge && merge->comm && terminal_tabs && !one_exchange) {
        // the corrections of levels kmin .. kcorr-1 are complete (per-read scan + the REDs of the histogram pass): merge them
        // over the ranks while the partition pass runs.  Neighbouring tables (engine.alloc_tables lays them out largest first
        // in one buffer, identically on every rank) go in one call.
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return kmap_check_launch("count_partitioned(event)");
        cudaEventRecord(ev, s);                        // (the partition pass is queued behind it)
        int v = kcorr - 1;
        bool first = true;
        while (v >= kmin) {
            uint32_t* lo = tt.t[v];
            int64_t cells = (int64_t)1 << (2 * v);
            int u = v - 1;
            while (!merge->scatter && u >= kmin && tt.t[u] == lo + cells) { cells += (int64_t)1 << (2 * u); --u; }   // (scattered: table by table)
            if (first) { cudaStreamWaitEvent(merge->stream, ev, 0); first = false; }
            const int rcm = kmap_merge_table_on(lo, cells, merge);
            if (rcm) { cudaEventDestroy(ev); return rcm; }
            v = u;
        }
        cudaEventDestroy(ev);
    }
    if (local) {
        cudaMemsetAsync(p.ticket, 0, 8, s);
        cudaMemsetAsync(p.queues, 0, (size_t)BC_QUEUES * (2 + n_all) * 4, s);
        const unsigned int g2 = (unsigned int)(n_tiles < 148 ? n_tiles : 148);
        const int smem1 = PT_SMEM, smem4 = PT_SMEM;
        // (function attributes are per device: set on every call -- it is a host-side table write -- rather than once per process)
        cudaFuncSetAttribute(partition_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
        cudaFuncSetAttribute(partition_kernel<PT_MAX_PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4);
        cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_WORDS * 4);
        cudaFuncSetAttribute(bucket_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_WORDS * 4);
        if (n_buckets <= PT_THREADS)
            partition_kernel<1><<<g2, PT_THREADS, smem1, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_base, p.suffixes, p.ticket);
        else
            partition_kernel<PT_MAX_PER><<<g2, PT_THREADS, smem4, s>>>(packed, valid, hide, n_words, n_tiles, k, n_buckets, n_extra, p.counts, p.off, p.chunk_base, p.suffixes, p.ticket);
        if (step_events && step_events[1]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(step_events[1]), s);
    }
    // A CTA of the per-bucket count fills an SM.  With an exchange running beside it, KMAP_COMM_CTAS SMs are left to the
    // collective's kernels (comm.cu caps them at that many CTAs) and the grid is one CTA per bucket instead of one persistent
    // CTA per SM, so that SMs change hands at bucket granularity (measured at 2 GPUs with persistent CTAs: the collective got
    // its SMs only between launches and the counting launches lost theirs to it: 6.9 ms instead of 4.4).
    const int64_t cells_k = (int64_t)n_buckets * BC_CELLS;
    const bool peer_path = kmap_merge_on_peer_memory(table, cells_k, merge);     // exchange AFTER the count (see below): nothing runs beside it
    const bool beside = merge && merge->comm && !peer_path;
    const int bc_grid = beside ? n_all : 148;
    // segments: 1.25 x the average bucket (no bucket of a uniform input is cut), at least 2^17 suffixes, a multiple of 8
    unsigned long long seg_len = (unsigned long long)(n / n_buckets + 1) * 5 / 4;
    if (seg_len < (1ull << 17)) seg_len = 1ull << 17;
    seg_len = (seg_len + 7ull) & ~7ull;
    int n_launch = 0;
    auto count_range = [&](int b_lo, int b_hi) {
        const int nb = b_hi - b_lo;
        if (nb <= 0 || !local) return;
        uint32_t* q = p.queues + (size_t)(n_launch++ % BC_QUEUES) * (2 + n_all);
        bucket_count_kernel<<<(unsigned int)(nb < bc_grid ? nb : bc_grid), BC_THREADS, BC_WORDS * 4, s>>>(p.suffixes, p.base, n_buckets, b_lo, b_hi, table,
                                                                                                        route ? ht.t[k - 1] : nullptr, seg_len, q);
        bucket_segments_kernel<<<beside ? 148 - kmap_comm_ctas(merge->world) : 148, BC_THREADS, BC_WORDS * 4, s>>>(p.suffixes, p.base, n_buckets, table,
                                                                                                       route ? ht.t[k - 1] : nullptr, seg_len, q);
    };
    auto fold_chain = [&]() {                           // brings the folded run-end corrections down the levels and into the tables
        for (int v = k - 1; fold && v >= kmin; --v) {
            const int64_t q4 = v > kmin ? ((int64_t)1 << (2 * (v - 1))) / 4 : ((int64_t)1 << (2 * v)) / 4;      // uint4 groups per launch
            int64_t g = (q4 + 255) / 256;
            if (g > 148 * 16) g = 148 * 16;
            fold_corrections_kernel<<<(unsigned int)(g < 1 ? 1 : g), 256, 0, s>>>(ctabs.t[v], tt.t[v], v > kmin ? ctabs.t[v - 1] : nullptr, v);
        }
    };
    if (!merge || !merge->comm) {
        count_range(0, n_all);
        fold_chain();
        return kmap_check_launch("count_partitioned");
    }
    // Sharded input: the slices of the table are final as soon as their buckets are counted, so the all-reduce over the
    // ranks runs range by range on the merge stream while the next range is being counted (the routed buckets first: they
    // complete the level k-1 corrections, which are merged as a whole).
    cudaEvent_t ev;
    int rc = KMAP_OK;
    const bool trace = getenv("KMAP_MERGE_TRACE") != nullptr;                          // (debugging aid: timeline of the overlap)
    cudaEvent_t t0 = nullptr, tc[8], tm[8];
    int nt = 0;
    if (trace) { cudaEventCreate(&t0); cudaEventRecord(t0, s); }
    auto merge_after = [&](uint32_t* buf, int64_t cells) {
        if (rc) return;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { rc = kmap_check_launch("count_partitioned(event)"); return; }
        cudaEventRecord(ev, s);
        cudaStreamWaitEvent(merge->stream, ev, 0);
        cudaEventDestroy(ev);                          // (released once the wait has been satisfied)
        if (trace) { cudaEventCreate(&tc[nt]); cudaEventRecord(tc[nt], s); }
        rc = kmap_merge_table_on(buf, cells, merge);
        if (trace) { cudaEventCreate(&tm[nt]); cudaEventRecord(tm[nt], merge->stream); ++nt; }
    };
    if (peer_path) {
        // The peer-memory exchange (peer.cu) is short next to the count and its kernels want the SMs the count fills (a CTA of
        // the per-bucket count leaves room for one small CTA beside it: an exchange issued range by range while the count runs
        // took 3 x as long as alone, 2 GPUs: 5.5 ms for this phase against 3.7 + 0.9 back to back).  So: count everything, then
        // ONE exchange -- of the level-k table and the routed level k-1 together when they are neighbours in memory.
        count_range(0, n_all);
        fold_chain();
        const int64_t cells_r = route ? (int64_t)1 << (2 * (k - 1)) : 0;
        if (one_exchange) {
            merge_after(table, span_all);
        } else if (route && !merge->scatter && tt.t[k - 1] == table + cells_k && kmap_merge_on_peer_memory(table, cells_k + cells_r, merge)) {
            merge_after(table, cells_k + cells_r);
        } else {
            if (route) merge_after(tt.t[k - 1], cells_r);
            merge_after(table, cells_k);
        }
    } else {
        if (route) {
            count_range(n_buckets, n_all);
            merge_after(tt.t[k - 1], (int64_t)1 << (2 * (k - 1)));
        }
        // (scattered: a rank's block of the table must stay one contiguous key range, so the table is merged in one piece)
        const int n_chunks = (n_buckets >= 1024 && !merge->scatter) ? kmap_merge_chunks(merge->world) : 1;
        for (int c = 0; c < n_chunks; ++c) {
            const int b_lo = n_buckets * c / n_chunks, b_hi = n_buckets * (c + 1) / n_chunks;
            count_range(b_lo, b_hi);
            merge_after(table + (size_t)b_lo * BC_CELLS, (int64_t)(b_hi - b_lo) * BC_CELLS);
        }
    }
    if (trace) {
        cudaStreamSynchronize(s); cudaStreamSynchronize(merge->stream);
        for (int i = 0; i < nt; ++i) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, t0, tc[i]); cudaEventElapsedTime(&b, t0, tm[i]);
            fprintf(stderr, "[merge trace] range %d: counted at %.3f ms, merged at %.3f ms\n", i, a, b);
            cudaEventDestroy(tc[i]); cudaEventDestroy(tm[i]);
        }
        cudaEventDestroy(t0);
    }
    if (rc) return rc;
    return kmap_check_launch("count_partitioned");
}

extern "C" int kmap_count_dense_partitioned(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table,
                                            void* scratch, int64_t scratch_bytes, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 9 && k <= 14, "the partitioned count covers 9 <= k <= 14");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && table && scratch, "null pointer");
    KMAP_REQUIRE(scratch_bytes >= kmap_partition_scratch_bytes(n, k), "scratch too small (kmap_partition_scratch_bytes)");
    return kmap_count_partitioned(packed, valid, nullptr, n, k, table, scratch, nullptr, k, nullptr, as_stream(stream), nullptr);
}
