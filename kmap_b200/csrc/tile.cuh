// Tile helpers shared by the level-k partition kernels (partition.cu): a CTA walks the packed reads in tiles
// of blockDim.x validity words, one word (32 window positions) per thread.
#pragma once
#include "common.cuh"

#define KMAP_TILE_THREADS 1024

// windows of k valid bases starting at bits 0..31 of V = v1:v0 (log-step run-length test, k <= 16)
__device__ __forceinline__ uint32_t window_mask(uint32_t v0, uint32_t v1, int k) {
    const uint64_t V = ((uint64_t)v1 << 32) | v0;
    const uint64_t r2 = V & (V >> 1);             // runs >= 2
    const uint64_t r4 = r2 & (r2 >> 2);           // >= 4
    const uint64_t r8 = r4 & (r4 >> 4);           // >= 8
    uint64_t m = ~0ull;
    int off = 0;
    if (k & 16) { m &= r8 & (r8 >> 8); off += 16; }
    if (k & 8) { m &= r8 >> off; off += 8; }
    if (k & 4) { m &= r4 >> off; off += 4; }
    if (k & 2) { m &= r2 >> off; off += 2; }
    if (k & 1) { m &= V >> off; }
    return (uint32_t)m;
}

struct TileWords { uint32_t fresh, w0, w1, w2, corr; };       // corr: windows of exactly k-1 valid bases (routed run-end corrections)
struct RawWords { uint32_t v0, v1, h, w0, w1, w2; };

// this thread's 32 positions of tile `tile`: validity (+ one word of look-ahead), hidden windows, the three packed
// words covering them.  Nothing is consumed here, so the loads of the NEXT tile can be in flight during a whole tile.
__device__ __forceinline__ RawWords load_raw_words(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                   const uint32_t* __restrict__ hide, int64_t n_words, int64_t tile) {
    RawWords r;
    r.v0 = r.v1 = r.h = r.w0 = r.w1 = r.w2 = 0;
    const int64_t w = tile * blockDim.x + threadIdx.x;                    // validity word index
    if (w < n_words) {                                            // (the arrays carry KMAP_PAD_WORDS zero words of padding)
        r.v0 = __ldcs(valid + w); r.v1 = __ldcs(valid + w + 1);
        if (hide) r.h = __ldcs(hide + w);
        const uint2 p = __ldcs(reinterpret_cast<const uint2*>(packed + 2 * w));
        r.w0 = p.x; r.w1 = p.y; r.w2 = __ldcs(packed + 2 * w + 2);
    }
    return r;
}
// route: also flag the windows that have exactly k-1 valid bases in front of a run end -- the "+1 at level k-1" corrections,
// when they travel through the partition as extra buckets instead of being scattered REDs (partition.cu)
__device__ __forceinline__ TileWords cook(const RawWords& r, int k, bool route = false) {
    TileWords t;
    const uint32_t full = window_mask(r.v0, r.v1, k);
    t.fresh = full & ~r.h;
    t.corr = route ? (window_mask(r.v0, r.v1, k - 1) & ~full & ~r.h) : 0u;
    t.w0 = r.w0; t.w1 = r.w1; t.w2 = r.w2;
    return t;
}
__device__ __forceinline__ TileWords load_tile_words(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                     const uint32_t* __restrict__ hide, int64_t n_words, int64_t tile, int k) {
    return cook(load_raw_words(packed, valid, hide, n_words, tile), k);
}

__device__ __forceinline__ uint32_t key_at(const TileWords& t, int i, int sh) {          // i is a compile-time constant
    const uint32_t x = (i < 16) ? __funnelshift_l(t.w1, t.w0, 2 * i) : __funnelshift_l(t.w2, t.w1, 2 * (i - 16));
    return x >> sh;
}


// ---- two validity words (64 window positions) per thread: the tiles of the partition kernels --------------------------------
// A tile is KMAP_TILE_WORDS = 2046 validity words (65 472 positions): thread t owns words 2t and 2t+1 (thread 1023 idles), so
// that no (tile, bucket) count can exceed 65 535 (the rows of pass 1 hold them as uint16) and every thread's first word index
// is even (128-bit loads of the packed words).
#define KMAP_TILE_WORDS 2046
struct RawPair { uint32_t v0, v1, v2, h0, h1, w0, w1, w2, w3, w4; };

__device__ __forceinline__ RawPair load_raw_pair(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                 const uint32_t* __restrict__ hide, int64_t n_words, int64_t tile) {
    RawPair r;
    r.v0 = r.v1 = r.v2 = r.h0 = r.h1 = r.w0 = r.w1 = r.w2 = r.w3 = r.w4 = 0;
    const int64_t w = tile * KMAP_TILE_WORDS + 2 * threadIdx.x;           // even
    if (2 * threadIdx.x < KMAP_TILE_WORDS && w < n_words) {               // (the arrays carry KMAP_PAD_WORDS zero words of padding)
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(valid + w));
        r.v0 = v.x; r.v1 = v.y; r.v2 = __ldcs(valid + w + 2);
        if (hide) { const uint2 h = __ldcs(reinterpret_cast<const uint2*>(hide + w)); r.h0 = h.x; r.h1 = h.y; }
        const uint4 p = __ldcs(reinterpret_cast<const uint4*>(packed + 2 * w));
        r.w0 = p.x; r.w1 = p.y; r.w2 = p.z; r.w3 = p.w; r.w4 = __ldcs(packed + 2 * w + 4);
    }
    return r;
}
__device__ __forceinline__ RawWords first_word(const RawPair& r) { RawWords a; a.v0 = r.v0; a.v1 = r.v1; a.h = r.h0; a.w0 = r.w0; a.w1 = r.w1; a.w2 = r.w2; return a; }
__device__ __forceinline__ RawWords second_word(const RawPair& r) { RawWords a; a.v0 = r.v1; a.v1 = r.v2; a.h = r.h1; a.w0 = r.w2; a.w1 = r.w3; a.w2 = r.w4; return a; }


// what run_end_corrections needs beyond RawWords: the validity / hidden bits of the 32 positions before this thread's
// word and the packed word before its first one.  Loaded together with the tile (no dependent loads later).
struct RawPrev { uint32_t vp, hp, wp; };

__device__ __forceinline__ RawPrev load_raw_prev(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                 const uint32_t* __restrict__ hide, int64_t n_words, int64_t tile) {
    RawPrev r;
    r.vp = r.hp = r.wp = 0;
    const int64_t w = tile * KMAP_TILE_WORDS + 2 * threadIdx.x;             // the first word of this thread's pair
    if (2 * threadIdx.x < KMAP_TILE_WORDS && w < n_words && w > 0) {
        r.vp = __ldcs(valid + w - 1);
        if (hide) r.hp = __ldcs(hide + w - 1);
        r.wp = __ldcs(packed + 2 * w - 1);
    }
    return r;
}

// "+1 at level v" for every window with exactly v valid bases (kmin <= v < k) in front of a run end inside this thread's
// word (bit j set, bit j+1 clear): such a window cannot be extended, so the 4:1 table reductions of the all-k count
// (count_all.cu) do not bring it down from the level above.  Windows hidden in `hide` belong to reads that are counted
// by the direct per-k kernels.  stab[v] = table of level v.  Everything comes from registers: the only memory operations
// are the REDs.
// single_from > 0 (the folded form, partition.cu): a run contributes ONE update, at the level of its length capped at k-1 --
// the levels below are brought down by 4:1 reductions over the FIRST base of the correction tables -- and a run of
// single_from or more bases contributes none here (it travels through the partition as a routed entry).
__device__ __forceinline__ void run_end_corrections(const RawWords& r, const RawPrev& q, int kmin, int k, uint32_t* const* stab,
                                                    int single_from = 0) {
    uint32_t ends = r.v0 & ~((r.v0 >> 1) | (r.v1 << 31));                  // bit j: position j valid, j+1 not
    if (ends == 0) return;
    const uint64_t W = ((uint64_t)r.v0 << 32) | q.vp;                      // position j of this word = bit 32 + j
    const uint64_t H = ((uint64_t)r.h << 32) | q.hp;
    do {
        const int j = __ffs(ends) - 1;
        ends &= ends - 1;
        const uint64_t inv = ~(W << (31 - j));                             // bit 63 = position j, going down = going back
        const int back = inv ? __clzll(inv) : 64;                          // valid bases ending at position j (>= 1)
        const int vmax = min(back, k - 1);
        if (vmax < kmin) continue;
        if (single_from > 0 && back >= single_from) continue;
        // the windows of kmin..vmax bases that end at j start at j-v+1 >= -14: 32 bases from offset o = 16 + j - vmax + 1 of
        // the 64 bases [wp w0 w1 w2] cover them all
        const int o = 17 + j - vmax;                                       // 2 .. 48
        const uint32_t a = o < 16 ? q.wp : (o < 32 ? r.w0 : r.w1);
        const uint32_t b = o < 16 ? r.w0 : (o < 32 ? r.w1 : r.w2);
        const uint32_t c = o < 16 ? r.w1 : (o < 32 ? r.w2 : 0u);
        const uint32_t hi = __funnelshift_l(b, a, 2 * (o & 15)), lo = __funnelshift_l(c, b, 2 * (o & 15));
        const int vlow = single_from > 0 ? vmax : kmin;
        for (int v = vmax; v >= vlow; --v) {
            if ((H >> (33 + j - v)) & 1ull) continue;
            const uint32_t x = __funnelshift_l(lo, hi, 2 * (vmax - v));
            atomicAdd(stab[v] + (x >> (32 - 2 * v)), 1u);
        }
    } while (ends);
}
