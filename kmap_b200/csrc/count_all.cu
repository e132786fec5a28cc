// Counting every k of a range in one go (the first-round counts of scan_motif, motif_discovery.py:262-273 +
// 627-636, for all k in [kmin, kmax]) with ~1/(kmax-kmin+1) of the atomic traffic of k independent passes.
//
// Identity used.  Let fresh_k(i) = 1 iff the window of k bases at position i is valid and is
// the first occurrence of its k-mer inside its read (repetitive mode: iff it is valid).  The reference's table is
//     T_k[h] = sum_i fresh_k(i) [w_k(i) = h].
// A (k+1)-window's k-prefix is the k-window at the same position, so
//     T_k[h] = sum_b T_{k+1}[4h + b] + sum_i (fresh_k(i) - fresh_{k+1}(i)) [w_k(i) = h].
// With vlen(i) = number of consecutive valid bases from i and dd(i) = the longest prefix (in bases, 0 if none >= kmin)
// that the window at i shares with an EARLIER window of the same read, fresh_k(i) = [k <= vlen(i)] [k > dd(i)], hence
// the correction term is +1 at level k = vlen(i) (a window that cannot be extended: one per valid run and level) and
// -1 at level k = dd(i) (a repeated k-mer whose extension is new: rare).  So only T_kmax needs one atomic per
// window; every smaller table is a 4:1 streaming reduction of the next one plus a few corrections.
//
// T_kmax itself (1 GiB at k = 14) is filled in key-range passes so that the slice being updated stays L2 resident.
#include "common.cuh"

namespace {

constexpr int AK_WARPS = 4;
constexpr int AK_WARP_MAX = 256;          // windows (at level kmin) a warp de-duplicates on chip
constexpr int AK_BLOCK_MAX = 8192;        // beyond this a read goes to the bitmap path

typedef KmapTableSet TableSet;            // t[k] = dense table of level k (only kmin..kmax are set)

__device__ __forceinline__ int run_length(uint32_t vb) { return vb == 0xFFFFFFFFu ? 32 : __ffs(~vb) - 1; }

// ---- A: per-read duplicate analysis at level kmin, corrections for kmin..kmax-1, duplicate mask for level kmax -------
// A warp works on R = 32 / G reads at a time, G lanes per read (G = 8 when every read of the batch has at most 128
// windows at level kmin, else 16): the cost of the scan is instructions per ROUND (one window of every lane), not per
// window, so four 100-bp reads side by side in one warp take about the rounds one read used to take with a warp of its
// own (206 -> ~75 warp instructions per read).  Lane j of a group owns the C = ceil(n_pos / G) consecutive windows
// j*C .. j*C+C-1 of its read (C <= 16): it fetches its stretch of the packed read and of the validity mask once (two
// funnel shifts each) and rolls through its windows with compile-time shifts.  Windows are visited in rounds
// (round t = window t of every lane); "earlier" below means an earlier (round, lane) pair of the SAME read -- any
// fixed total order gives the same tables.
// Repeats inside a read are rare, so the common path is a filter, not a set: every warp owns a bitmap of AK_BM_BITS
// bits in shared memory, indexed by a hash of (group, kmin-mer).  One ATOMS.OR per window sets the bit and returns
// whether it was already set; set bits are cleared again (plain stores of zero words) when the reads are done, so
// there are no epochs and no probing.  A window that found its bit set MAY repeat an earlier one (or collide with
// another window: ~0.3 % of the windows of a 100-bp read): a cheap exact test on the kmin-mers of the group decides
// whether anything repeats at all, and only then the warp compares the window with every earlier window of the read
// -- all still in registers -- and gets the exact repeat depth.  When several lanes of one round share a bit, the
// lane that won the atomic need not be the lowest one; every flagged lane therefore also broadcasts its window to the
// HIGHER lanes of its group and round, which is how an unflagged winner learns about the earlier windows it repeats.
#ifndef KMAP_AK_BM_BITS
#define KMAP_AK_BM_BITS 65536
#endif
constexpr int AK_BM_BITS = KMAP_AK_BM_BITS;      // per warp; 16 index bits are hashed, the low ones are used
constexpr int AK_BM_WORDS = AK_BM_BITS / 32;
constexpr int AK_BM_LOG_WORDS = AK_BM_BITS == 65536 ? 11 : AK_BM_BITS == 32768 ? 10 : AK_BM_BITS == 16384 ? 9 : 8;
static_assert((1 << AK_BM_LOG_WORDS) == AK_BM_WORDS, "AK_BM_BITS must be 8192, 16384, 32768 or 65536");
#ifndef KMAP_AK_BLOCKS
#define KMAP_AK_BLOCKS 6
#endif
#ifndef KMAP_AK_PIPE
#define KMAP_AK_PIPE 0
#endif
constexpr int AK_BLOCKS = KMAP_AK_BLOCKS;  // resident blocks per SM the scan is compiled for (register cap)
constexpr int AK_MAXC = 12;               // windows per lane (even)

__device__ __forceinline__ int common_prefix(uint32_t a, uint32_t b) {       // equal leading bases of two 16-base words
    const uint32_t diff = a ^ b;
    return diff ? (__clz(diff) >> 1) : 16;
}

struct DedupStaged { uint32_t st_lo, st_hi; int L; uint32_t va, vb, w0, w1, w2; };

struct DedupCtx {
    const uint32_t* __restrict__ packed;
    const uint32_t* __restrict__ valid;
    uint32_t* __restrict__ dupmask;
    uint32_t* __restrict__ work;
    uint32_t* const* stab;
    uint32_t* bm;
    uint32_t dummy_off;                // byte offset (from bm) of this lane's dummy word
    int64_t n_seq, n;
    const int64_t* __restrict__ borders;
    int kmin, kmax, lane;
};

// raw words of this lane's stretch of read `jj` of the batch (no use yet: the loads overlap the work on the previous reads)
template <int GS>
__device__ __forceinline__ void dedup_stage(const DedupCtx& c, uint32_t stlo_lane, uint32_t sthi_lane, int L_lane, int n_in, int pass,
                                            DedupStaged& g) {
    constexpr int G = 1 << GS, R = 32 >> GS;
    const int rr = pass * R + (c.lane >> GS);                     // read of the batch this lane's group works on
    const int src = rr & 31;
    g.st_lo = __shfl_sync(0xFFFFFFFFu, stlo_lane, src);
    g.st_hi = __shfl_sync(0xFFFFFFFFu, sthi_lane, src);
    const int len = __shfl_sync(0xFFFFFFFFu, L_lane, src);
    g.L = (rr < n_in && pass * R < 32) ? len : 0;                 // (words that are not loaded keep stale values: never looked at)
    const int n_pos = g.L - c.kmin + 1;
    if (n_pos <= 0 || n_pos > AK_WARP_MAX) return;
    const int C = (n_pos + G - 1) >> GS;
    const int base = (c.lane & (G - 1)) * C;
    if (base >= n_pos) return;
    // word indices fit 32 bits (the host side checks n < 2^36)
    const uint32_t bv = (g.st_lo & 31u) + (uint32_t)base;
    const uint32_t* vd = c.valid + (__funnelshift_r(g.st_lo, g.st_hi, 5) + (bv >> 5));
    g.va = __ldg(vd); g.vb = __ldg(vd + 1);
    const uint32_t bp = (g.st_lo & 15u) + (uint32_t)base;
    const uint32_t* pk = c.packed + (__funnelshift_r(g.st_lo, g.st_hi, 4) + (bp >> 4));
    g.w0 = __ldg(pk); g.w1 = __ldg(pk + 1); g.w2 = __ldg(pk + 2);
}

// windows of k (<= 15) valid bases starting at bits 0..15 of v (bit i = position i; log-step run-length test)
__device__ __forceinline__ uint32_t run_mask(uint32_t v, int k) {
    const uint32_t r2 = v & (v >> 1), r4 = r2 & (r2 >> 2), r8 = r4 & (r4 >> 4);
    uint32_t m = 0xFFFFFFFFu;
    int off = 0;
    if (k & 8) { m &= r8; off = 8; }
    if (k & 4) { m &= r4 >> off; off += 4; }
    if (k & 2) { m &= r2 >> off; off += 2; }
    if (k & 1) { m &= v >> off; }
    return m;
}

// everything that happens to the R reads of one pass.  Three phases:
//   filter  every window of every lane sets its bit (rounds fully unrolled and independent of each other: the atomics of
//           a pass are all in flight together, nothing is voted on per round); a lane remembers which of its windows
//           found their bit set;
//   exact   only for the rounds in which some lane was flagged (about one per pass: bitmap collisions and real repeats):
//           the flagged windows are compared with the windows of their read, which are rebuilt from the two register
//           words of every lane;
//   clear   one store per window.
template <int GS>
__device__ __forceinline__ void dedup_process(const DedupCtx& c, const DedupStaged& g, int64_t batch, int n_in, int pass) {
    constexpr int G = 1 << GS, R = 32 >> GS;
    const int lane = c.lane, kmin = c.kmin, kmax = c.kmax;
    const int grp = lane >> GS, j = lane & (G - 1);
    const int rr = pass * R + grp;
    const int64_t st = (int64_t)(((uint64_t)g.st_hi << 32) | g.st_lo);
    const int n_pos = g.L - kmin + 1;
    if (__any_sync(0xFFFFFFFFu, n_pos > AK_WARP_MAX)) {
        // too long for the on-chip path: hide the read from the masked count and queue it for the direct kernels
        if (n_pos > AK_WARP_MAX) {
            const int64_t en_raw = __ldg(c.borders + 2 * (batch + rr) + 1);      // (g.L is clamped; rare path: fetch the end again)
            const int64_t en_grp = en_raw > c.n ? c.n : en_raw;
            const int64_t r = batch + rr;
            const int64_t L64 = en_grp - st;
            for (int64_t w = (st >> 5) + j; w <= ((en_grp - 1) >> 5); w += G) {
                const int64_t lo = w << 5;
                uint32_t bits = 0xFFFFFFFFu;
                if (lo < st) bits &= 0xFFFFFFFFu << (st - lo);
                if (lo + 32 > en_grp) bits &= 0xFFFFFFFFu >> (lo + 32 - en_grp);
                atomicOr(c.dupmask + w, bits);
            }
            if (j == 0) {
                if (L64 - kmin + 1 <= AK_BLOCK_MAX) (c.work + 4)[atomicAdd(&c.work[0], 1u)] = (uint32_t)r;          // medium reads
                else (c.work + 4 + c.n_seq)[atomicAdd(&c.work[1], 1u)] = (uint32_t)r;                               // long reads
            }
        }
    }
    // this lane's stretch of its read, in 32-bit arithmetic relative to the read start
    const int C = (n_pos > 0 && n_pos <= AK_WARP_MAX) ? (n_pos + G - 1) >> GS : 0;      // windows per lane of this group, 0..16
    const int Cmax = __reduce_max_sync(0xFFFFFFFFu, C);
    if (Cmax == 0) return;
    const int base = j * C;
    uint32_t vbits = 0, hi = 0, lo = 0;
    if (base < n_pos) {                                         // (C > 0 here)
        vbits = __funnelshift_r(g.va, g.vb, ((int)(g.st_lo & 31u) + base) & 31);
        const int room = g.L - base;                            // positions of the read from `base` on
        if (room < 32) vbits &= (1u << room) - 1u;              // never look past the read end
        const int sh = ((int)(g.st_lo & 15u) + base) & 15;
        hi = __funnelshift_l(g.w1, g.w0, 2 * sh);               // bases base .. base+15
        lo = __funnelshift_l(g.w2, g.w1, 2 * sh);               // bases base+16 .. base+31
    }
    // bit t: window base+t belongs to this lane and has kmin valid bases (a window that would leave the read has fewer
    // than kmin valid bits left: `room` above already excludes it)
    const uint32_t okm = run_mask(vbits, kmin) & ((1u << C) - 1u);
    const int key_shift = 32 - 2 * kmin;
    const uint32_t salt = (uint32_t)grp * 0x3C6EF372u;          // reads that share the warp's bitmap use different slots
    uint8_t* const bm8 = reinterpret_cast<uint8_t*>(c.bm);

    // ---- filter -------------------------------------------------------------------------------------------------------
    // (a lane without a window at round t sends its atomic to a private dummy word instead: no branch, no operand select)
    uint32_t woff2[AK_MAXC / 2];                                // byte offsets of the touched words, two per register
    uint32_t flagged = 0;                                       // bit t: window t found its bit already set
#pragma unroll
    for (int t = 0; t < AK_MAXC; ++t) {
        if ((t & 3) == 0 && t >= Cmax) break;                   // (warp-uniform, tested every fourth round)
        const uint32_t x = __funnelshift_l(lo, hi, 2 * t);      // 16 bases from window base+t
        const uint32_t h = (x >> key_shift) * 0x9E3779B1u + salt;
        uint32_t wo = (h >> (32 - AK_BM_LOG_WORDS - 2)) & (uint32_t)((AK_BM_WORDS - 1) << 2);     // byte offset of the word
        wo = ((okm >> t) & 1u) ? wo : c.dummy_off;
        if (t & 1) woff2[t >> 1] |= wo << 16; else woff2[t >> 1] = wo;
        const uint32_t bit = __funnelshift_l(0u, 1u, h);        // 1 << (h & 31): M is odd, so these are the low key bits mixed
        const uint32_t old = atomicOr(reinterpret_cast<uint32_t*>(bm8 + wo), bit);
        if (old & bit) flagged |= 1u << t;
    }
    flagged &= okm;

    // ---- exact: the rounds with a flagged window -------------------------------------------------------------------------
    uint32_t rounds = __reduce_or_sync(0xFFFFFFFFu, flagged);
    while (rounds) {
        const int t = __ffs(rounds) - 1;
        rounds &= rounds - 1;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, (flagged >> t) & 1u);
        const uint32_t x = __funnelshift_l(lo, hi, 2 * t);
        // depth dd of the longest repeat with an earlier window of the read (order: round, then lane)
        const int vlen = ((okm >> t) & 1u) ? run_length(vbits >> t) : 0;     // windows shorter than kmin never match anything
        int dd = 0;
        do {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t xi = __shfl_sync(0xFFFFFFFFu, x, b);
            const bool same_read = (b >> GS) == grp;
            // quick exact test: does any other window of this read -- an earlier round, or this round -- start with the same
            // kmin bases?  (validity ignored: conservative).  Most flagged windows are bitmap collisions and stop here.
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < AK_MAXC; ++u)
                if ((((__funnelshift_l(lo, hi, 2 * u) ^ xi) >> key_shift) == 0u)) m |= 1u << u;
            m &= ((1u << C) - 1u) & ((2u << t) - 1u);           // this lane's own windows of the rounds up to t ...
            if (lane == b) m &= ~(1u << t);                      // ... except the flagged window itself
            const bool hit = m != 0u;
            if (!__any_sync(0xFFFFFFFFu, hit && same_read)) continue;
            const int vi = min(__shfl_sync(0xFFFFFFFFu, vlen, b), kmax);
            int best = 0;
            if (same_read) {
                for (int u = 0; u < t; ++u) {
                    const int lu = ((okm >> u) & 1u) ? run_length(vbits >> u) : 0;
                    best = max(best, min(common_prefix(__funnelshift_l(lo, hi, 2 * u), xi), min(lu, vi)));
                }
                const int same = min(common_prefix(x, xi), min(vlen, vi));
                if (lane < b) best = max(best, same);
                else if (lane > b && same >= kmin) dd = max(dd, same);      // b precedes this lane in the round
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));   // (stays inside the group)
            if (lane == b) dd = max(dd, best);
        } while (todo);
        // fresh_k(i) = [k <= vlen][k > dd]: the window is hidden from level kmax when it repeats at every level;
        // otherwise level dd loses one count (this cancels either the count its extension brings up from level
        // dd+1, or -- when dd == vlen -- the +1 the run-end corrections add for a window that ends its run)
        if (dd >= kmin) {
            if (dd >= kmax) {
                const int64_t p = st + base + t;
                atomicOr(c.dupmask + (p >> 5), 1u << (p & 31));
            } else {
                atomicAdd(c.stab[dd] + (x >> (32 - 2 * dd)), 0xFFFFFFFFu);
            }
        }
    }

    // ---- clear: give the bitmap back (every bit set above lives in the word of one of this lane's windows) ---------------
    __syncwarp();
#pragma unroll
    for (int t = 0; t < AK_MAXC; ++t) {
        if ((t & 3) == 0 && t >= Cmax) break;
        *reinterpret_cast<uint32_t*>(bm8 + ((t & 1) ? woff2[t >> 1] >> 16 : woff2[t >> 1] & 0xFFFFu)) = 0u;
    }
    __syncwarp();
}

template <int GS>
__device__ __forceinline__ void dedup_batch(const DedupCtx& c, uint32_t stlo_lane, uint32_t sthi_lane, int L_lane,
                                            int64_t batch, int n_in) {
    constexpr int R = 32 >> GS;
    // two staging register sets alternate, so that the software pipeline needs no register moves
    DedupStaged sa, sb;
    sa.st_lo = sa.st_hi = 0; sa.L = 0; sa.va = sa.vb = sa.w0 = sa.w1 = sa.w2 = 0;
    sb = sa;
    const int n_pass = (n_in + R - 1) / R;
#if !KMAP_AK_PIPE
#pragma unroll 1
    for (int p = 0; p < n_pass; ++p) {
        dedup_stage<GS>(c, stlo_lane, sthi_lane, L_lane, n_in, p, sa);
        dedup_process<GS>(c, sa, batch, n_in, p);
    }
    return;
#endif
    dedup_stage<GS>(c, stlo_lane, sthi_lane, L_lane, n_in, 0, sa);
#pragma unroll 1
    for (int p = 0; p < n_pass; p += 2) {
        dedup_stage<GS>(c, stlo_lane, sthi_lane, L_lane, n_in, p + 1, sb);
        dedup_process<GS>(c, sa, batch, n_in, p);
        dedup_stage<GS>(c, stlo_lane, sthi_lane, L_lane, n_in, p + 2, sa);
        dedup_process<GS>(c, sb, batch, n_in, p + 1);
    }
}

__global__ void __launch_bounds__(AK_WARPS * 32, AK_BLOCKS) dedup_scan_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, int64_t n, const int64_t* __restrict__ borders,
    int64_t n_seq, int kmin, int kmax, TableSet tabs, uint32_t* __restrict__ dupmask, uint32_t* __restrict__ work) {
    __shared__ __align__(16) uint32_t bm_all[AK_WARPS][AK_BM_WORDS + 32];    // + one dummy word per lane
    __shared__ uint32_t* stab[16];
    if (threadIdx.x < 16) stab[threadIdx.x] = tabs.t[threadIdx.x];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    {
        uint4* b4 = reinterpret_cast<uint4*>(bm_all[wib]);
#pragma unroll
        for (int j = 0; j < AK_BM_WORDS / 4 / 32; ++j) b4[lane + 32 * j] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int64_t warp0 = (int64_t)blockIdx.x * AK_WARPS + wib;
    const int64_t n_warps = (int64_t)gridDim.x * AK_WARPS;
    DedupCtx c;
    c.packed = packed; c.valid = valid; c.dupmask = dupmask; c.work = work; c.stab = stab; c.bm = bm_all[wib];
    c.dummy_off = (uint32_t)(AK_BM_WORDS + lane) * 4u;
    c.n_seq = n_seq; c.n = n; c.borders = borders;
    c.kmin = kmin; c.kmax = kmax; c.lane = lane;

    // A warp takes 32 consecutive reads at a time: lane j fetches and clamps the borders of read j (one coalesced 16-byte
    // load per read, the 64-bit arithmetic done once per lane instead of once per lane and read), then walks them R at a
    // time, getting (start, length) of each by shuffles.  The words of the next R reads are requested before the current
    // ones are processed (nothing consumes them until the next iteration), so their DRAM latency overlaps the work.
    const longlong2* borders2 = reinterpret_cast<const longlong2*>(borders);
    for (int64_t batch = warp0 * 32; batch < n_seq; batch += n_warps * 32) {
        const int64_t r_lane = batch + lane;
        const longlong2 be = r_lane < n_seq ? __ldg(borders2 + r_lane) : make_longlong2(0, 0);
        const int64_t st_lane = be.x < 0 ? 0 : be.x, en_lane = be.y > n ? n : be.y;
        const int64_t len_lane = en_lane - st_lane;
        const int L_lane = (int)(len_lane < 0 ? 0 : (len_lane > 0x3FFFFFFF ? 0x3FFFFFFF : len_lane));
        const uint32_t stlo_lane = (uint32_t)st_lane, sthi_lane = (uint32_t)((uint64_t)st_lane >> 32);
        const int n_in = (int)(n_seq - batch < 32 ? n_seq - batch : 32);
        const int np_lane = L_lane - kmin + 1;
        const int np_max = __reduce_max_sync(0xFFFFFFFFu, np_lane <= AK_WARP_MAX ? np_lane : 0);
        if (np_max <= 4 * AK_MAXC) dedup_batch<2>(c, stlo_lane, sthi_lane, L_lane, batch, n_in);
        else if (np_max <= 8 * AK_MAXC) dedup_batch<3>(c, stlo_lane, sthi_lane, L_lane, batch, n_in);
        else if (np_max <= 16 * AK_MAXC) dedup_batch<4>(c, stlo_lane, sthi_lane, L_lane, batch, n_in);
        else dedup_batch<5>(c, stlo_lane, sthi_lane, L_lane, batch, n_in);
    }
}

// ---- "+1 at level v" for the windows that end a valid run (every mode) ----------------------------------------------------
// A window with exactly v valid bases in front of its run end (kmin <= v < kmax) cannot be extended to the next level, so
// the 4:1 reduction does not bring it down from above: it is added here.  One thread = one validity word; only run ends
// are visited (bit j set, bit j+1 clear: about one per read).  One launch handles ONE level, and for a table beyond L2
// one key-prefix slice of it, so that the scattered updates always land in an L2-resident range (random RED into a
// DRAM-resident table runs at 1/8 of the L2 rate, profiles/r01_red_rate_microbench.txt); re-reading the validity bits
// per launch is cheap next to that.  Reads that dedup_scan_kernel routed to the direct per-k kernels are hidden as a
// whole in `hide` (a window that ends a run is never hidden for any other reason).
__global__ void __launch_bounds__(256) terminal_corrections_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                   const uint32_t* __restrict__ hide, int64_t n_groups, int v_lo,
                                                                   int v_hi, int prefix_bases, uint32_t prefix, TableSet tabs) {
    __shared__ uint32_t* stab[16];
    if (threadIdx.x < 16) stab[threadIdx.x] = tabs.t[threadIdx.x];
    __syncthreads();
    // one thread = 4 validity words per step (the arrays are padded: KMAP_PAD_WORDS), grid-stride
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * 256) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(valid) + g);
        if ((q.x | q.y | q.z | q.w) == 0) continue;
        const uint32_t vw[6] = {g > 0 ? __ldcs(valid + 4 * g - 1) : 0u, q.x, q.y, q.z, q.w, __ldcs(valid + 4 * g + 4)};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t v0 = vw[c + 1];
            uint32_t ends = v0 & ~((v0 >> 1) | (vw[c + 2] << 31));
            if (ends == 0) continue;
            const int64_t t = 4 * g + c;
            const uint64_t W = ((uint64_t)v0 << 32) | vw[c];            // position j of this word = bit 32 + j
            uint64_t H = 0;
            if (hide) H = ((uint64_t)__ldg(hide + t) << 32) | (t > 0 ? __ldg(hide + t - 1) : 0u);
            while (ends) {
                const int j = __ffs(ends) - 1;
                ends &= ends - 1;
                const uint64_t inv = ~(W << (31 - j));                   // bit 63 = position j, going down = going back
                const int back = inv ? __clzll(inv) : 64;                // valid bases ending at position j (>= 1)
                const int vmax = min(back, v_hi);
                if (vmax < v_lo) continue;
                // the windows of v_lo..vmax bases that end at j start at j-v+1: one 32-base fetch covers them all
                const int64_t p0 = t * 32 + j - vmax + 1;
                const uint32_t hi = window16(packed, p0), lo = window16(packed, p0 + 16);
                for (int v = vmax; v >= v_lo; --v) {
                    if ((H >> (33 + j - v)) & 1ull) continue;
                    const uint32_t x = __funnelshift_l(lo, hi, 2 * (vmax - v));
                    if (prefix_bases && (x >> (32 - 2 * prefix_bases)) != prefix) continue;
                    atomicAdd(stab[v] + (x >> (32 - 2 * v)), 1u);
                }
            }
        }
    }
}

// ---- B: level-kmax count restricted to the keys that start with a given prefix of PB bases -----------------------------
// One thread = 32 consecutive window positions.  Instead of hashing all 32 windows and filtering, the thread first
// builds the set of positions whose first PB bases equal the pass prefix with a few bit operations on the packed words
// (2-bit groups compared in place, position i at bit 62-2i of a 64-bit mask), then visits only those (32 / 4^PB on
// average).  A pass therefore costs ~1/6 of a full count and issues atomics only for its own key range, whose slice of
// the table (4^(k-PB) cells) stays L2 resident.
__device__ __forceinline__ uint32_t eq_groups(uint32_t w, uint32_t rep) {      // bit 2g set iff 2-bit group g of w == rep's
    const uint32_t x = w ^ rep;
    return ~(x | (x >> 1)) & 0x55555555u;
}

template <int PB>
__global__ void __launch_bounds__(256) count_prefix_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                           const uint32_t* __restrict__ hide_mask, int64_t n_groups, int k,
                                                           uint32_t* __restrict__ table, uint32_t prefix) {
    // one thread = 4 validity words = 128 window positions; everything it needs arrives in six 128-bit / 32-bit
    // streaming loads issued together (the inputs are read once per pass: evict-first keeps the table slice in L2)
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n_groups) return;
    const uint4 v4 = __ldcs(reinterpret_cast<const uint4*>(valid) + g);
    const uint32_t v_next = __ldcs(valid + 4 * g + 4);
    const uint4 pa = __ldcs(reinterpret_cast<const uint4*>(packed) + 2 * g);
    const uint4 pb = __ldcs(reinterpret_cast<const uint4*>(packed) + 2 * g + 1);
    const uint32_t p_next = __ldcs(packed + 8 * g + 8);
    uint4 h4 = make_uint4(0, 0, 0, 0);
    if (hide_mask) h4 = __ldcs(reinterpret_cast<const uint4*>(hide_mask) + g);
    const uint32_t vw[5] = {v4.x, v4.y, v4.z, v4.w, v_next};
    const uint32_t pw[9] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, p_next};
    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w};
    const uint32_t km = (1u << k) - 1u;
    const int sh = 32 - 2 * k;
    uint32_t reps[PB > 0 ? PB : 1];
#pragma unroll
    for (int j = 0; j < PB; ++j) reps[j] = ((prefix >> (2 * (PB - 1 - j))) & 3u) * 0x55555555u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t v0 = vw[q];
        if (v0 == 0) continue;
        const uint32_t w0 = pw[2 * q], w1 = pw[2 * q + 1], w2 = pw[2 * q + 2];
        // candidate positions: base j of the window equals base j of the prefix, j = 0..PB-1
        uint64_t sel = 0x5555555555555555ull;
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const uint64_t e01 = ((uint64_t)eq_groups(w0, reps[j]) << 32) | eq_groups(w1, reps[j]);
            sel &= j == 0 ? e01 : ((e01 << (2 * j)) | (eq_groups(w2, reps[j]) >> (32 - 2 * j)));
        }
        const uint32_t v1 = vw[q + 1], hide = hw[q];
        while (sel) {
            const int b = __clzll(sel);
            sel &= ~(0x8000000000000000ull >> b);
            const int i = b >> 1;                                    // window position inside this word
            const uint32_t vb = __funnelshift_r(v0, v1, i);
            if ((vb & km) != km || ((hide >> i) & 1u)) continue;
            const uint32_t x = (i < 16) ? __funnelshift_l(w1, w0, 2 * i) : __funnelshift_l(w2, w1, 2 * (i - 16));
            atomicAdd(table + (x >> sh), 1u);
        }
    }
}

// ---- derive: T_k[h] += T_{k+1}[4h] + .. + T_{k+1}[4h+3] -----------------------------------------------------------------
__global__ void __launch_bounds__(256) derive_table_kernel(const uint4* __restrict__ upper, uint32_t* __restrict__ lower, int64_t n_cells) {
    int64_t h = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; h < n_cells; h += stride) {
        const uint4 v = __ldg(upper + h);
        lower[h] += v.x + v.y + v.z + v.w;
    }
}

}  // namespace

// implemented in count.cu: direct per-k handling of reads too long for the warp path, driven by the lists in `work`
int kmap_count_long_reads(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                          int k, uint32_t* table, uint32_t* work, uint32_t* bitmap, const uint32_t counts[2], cudaStream_t s);
// implemented in partition.cu: level-k count through key partitioning + shared-memory counters
int kmap_count_partitioned(const uint32_t* packed, const uint32_t* valid, const uint32_t* hide, int64_t n, int k, uint32_t* table,
                           void* scratch, const KmapTableSet* terminal_tabs, int kmin, void* const* step_events, cudaStream_t s,
                           const KmapMerge* merge);

// ---- the sharded form: this rank's reads, tables merged over the ranks -------------------------------------------------------
// Every pre-derive buffer is linear in the reads (T_k = fold(T_k+1) + corrections_k, and the fold is linear), so the merged
// tables are obtained by all-reducing the level-kmax table and the correction buffers of the lower levels and deriving
// AFTERWARDS.  That lets the exchange start early: the corrections of the L2-resident levels are final after the histogram
// pass (merged while the partition pass runs), the routed level and the slices of the level-kmax table are merged range by
// range while the per-bucket count is still going (partition.cu).  Reads beyond the on-chip paths (counted by the direct
// kernels after the reductions) are rare: whether ANY rank has some is learned from an 8-byte all-reduce of the queue
// counters; if so their tables are built apart, merged and added.
namespace {

__global__ void __launch_bounds__(256) add_tables_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) dst[i] += src[i];
}

struct OneShotEvent {       // record on one stream, make another wait, release
    // keep != NULL: only record on `from` and hand the event over (the caller makes `to` wait later and destroys it)
    static int chain(cudaStream_t from, cudaStream_t to, cudaEvent_t* keep) {
        (void)to;
        if (cudaEventCreateWithFlags(keep, cudaEventDisableTiming) != cudaSuccess) { *keep = nullptr; return kmap_check_launch("count_all_k(event)"); }
        cudaEventRecord(*keep, from);
        return KMAP_OK;
    }
    static int chain(cudaStream_t from, cudaStream_t to) {
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return kmap_check_launch("count_all_k(event)");
        cudaEventRecord(ev, from);
        cudaStreamWaitEvent(to, ev, 0);
        cudaEventDestroy(ev);
        return KMAP_OK;
    }
};

int count_all_impl(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                   int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                   uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                   void* const* phase_events, void* stream, const KmapMerge* merge) {
    KMAP_REQUIRE(n >= 0 && n_seq >= 0 && kmin >= 1 && kmin <= kmax && kmax <= 15, "need 1 <= kmin <= kmax <= 15");
    KMAP_REQUIRE(n_seq < (int64_t)0xFFFFFFFFll, "too many reads for one call (shard the input)");
    KMAP_REQUIRE(n < ((int64_t)1 << 36), "too many positions for one call (shard the input)");
    KMAP_REQUIRE(tables_host, "null pointer");
    KMAP_REQUIRE(scheme >= KMAP_KMAX_PREFIX_PASSES && scheme <= KMAP_KMAX_SORTED, "unknown scheme");
    cudaStream_t s = as_stream(stream);
    cudaEvent_t zeroed = nullptr;
    TableSet tabs;
    for (int k = 0; k < 16; ++k) tabs.t[k] = nullptr;
    // The level-kmax table of the partitioned count is first touched by the per-bucket count, long after the per-read scan:
    // with an exchange stream at hand its zero fill (1 GiB at k = 14) runs there, beside the scan, which leaves HBM idle.
    // (single GPU: a stream created for the call and released when its fill is done)
    bool zero_aside = dedup && scheme != KMAP_KMAX_PREFIX_PASSES && part_scratch && kmax >= 12 && kmax <= 14 && n > 0 && n_seq > 0;
    cudaStream_t aside = merge && merge->stream && merge->stream != s ? merge->stream : nullptr;
    const bool own_aside = zero_aside && !aside;
    if (own_aside && cudaStreamCreateWithFlags(&aside, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); zero_aside = false; }
    for (int k = kmin; k <= kmax; ++k) {
        tabs.t[k] = tables_host[k - kmin];
        KMAP_REQUIRE(tabs.t[k], "null table");
        cudaError_t e = cudaSuccess;
        if (k == kmax && zero_aside) {
            int rcz = OneShotEvent::chain(s, aside);                      // (whatever the caller did with the table before is done)
            if (!rcz) e = cudaMemsetAsync(tabs.t[k], 0, ((size_t)1 << (2 * k)) * 4, aside);
            if (!rcz) rcz = OneShotEvent::chain(aside, s, &zeroed);       // (waited for in front of the level-kmax count)
            if (own_aside) cudaStreamDestroy(aside);                      // (the stream goes away when the fill has completed)
            if (rcz) return rcz;
        } else {
            e = cudaMemsetAsync(tabs.t[k], 0, ((size_t)1 << (2 * k)) * 4, s);
        }
        if (e != cudaSuccess) { kmap_set_error("count_all_k: %s", cudaGetErrorString(e)); return (int)e; }
    }
    // an empty shard launches nothing of its own but still takes part in every collective of the sharded form
    const bool local = n > 0 && (!dedup || n_seq > 0);
    if (!local && !merge) return KMAP_OK;
    KMAP_REQUIRE(!local || (packed && valid), "null pointer");
    if (dedup) KMAP_REQUIRE(dupmask && work && (borders || !local), "de-duplication needs borders, dupmask and work scratch");
    // optional instrumentation (bench.py): 6 caller-owned cudaEvent_t recorded after zeroing, after the per-read scan,
    // after the level-kmax passes and after the table reductions; [4], [5] inside the partitioned level-kmax count: after
    // the bucket histogram (+ run-end corrections) and after the partition pass
    auto mark = [&](int i) { if (phase_events && phase_events[i]) cudaEventRecord(reinterpret_cast<cudaEvent_t>(phase_events[i]), s); };
    mark(0);
    const int64_t n_words = (n + 31) / 32;
    cudaError_t e = cudaSuccess;
    int rc = KMAP_OK;
    if (dedup) {
        e = cudaMemsetAsync(work, 0, 16, s);
        if (e == cudaSuccess && local) e = cudaMemsetAsync(dupmask, 0, (size_t)kmap_valid_words(n) * 4, s);
        if (e != cudaSuccess) { kmap_set_error("count_all_k: %s", cudaGetErrorString(e)); return (int)e; }
        if (local) {
            int64_t blocks = (n_seq + 32 * AK_WARPS - 1) / (32 * AK_WARPS);          // a warp takes 32 reads at a time
            if (blocks > 148 * AK_BLOCKS * 8) blocks = 148 * AK_BLOCKS * 8;           // AK_BLOCKS blocks of 4 warps (32 KB of marks each) per SM
            dedup_scan_kernel<<<(unsigned int)blocks, AK_WARPS * 32, 0, s>>>(packed, valid, n, borders, n_seq, kmin, kmax, tabs, dupmask, work);
        }
        if (merge) {               // work[2..3] = the queue counters of all ranks: does ANY rank hold reads for the direct kernels?
            e = cudaMemcpyAsync(work + 2, work, 8, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) { kmap_set_error("count_all_k: %s", cudaGetErrorString(e)); return (int)e; }
            if ((rc = OneShotEvent::chain(s, merge->stream))) return rc;
            if ((rc = kmap_allreduce_u32_on(work + 2, 2, merge->comm, merge->stream))) return rc;
        }
    }
    const bool use_partition = scheme != KMAP_KMAX_PREFIX_PASSES && part_scratch && kmax >= 12 && kmax <= 14;
    if (!use_partition && local) {          // (the partitioned count does these corrections inside its histogram pass)
        const int64_t n_groups = (n_words + 3) / 4;
        int64_t tb = (n_groups + 255) / 256;
        if (tb > 148 * 16) tb = 148 * 16;
        // launches: consecutive levels whose tables together stay L2 resident share one pass; a level beyond that is
        // done in key-prefix slices of 4^12 cells
        const size_t budget = (size_t)96 << 20;
        int v = kmin;
        while (v < kmax) {
            const int pb = v > 12 ? v - 12 : 0;
            if (pb > 0) {
                for (uint32_t prefix = 0; prefix < (1u << (2 * pb)); ++prefix)
                    terminal_corrections_kernel<<<(unsigned int)tb, 256, 0, s>>>(packed, valid, dedup ? dupmask : nullptr, n_groups, v, v,
                                                                                 pb, prefix, tabs);
                ++v;
                continue;
            }
            int hi = v;
            size_t bytes = (size_t)4 << (2 * v);
            while (hi + 1 < kmax && hi + 1 <= 12 && bytes + ((size_t)4 << (2 * (hi + 1))) <= budget) { ++hi; bytes += (size_t)4 << (2 * hi); }
            terminal_corrections_kernel<<<(unsigned int)tb, 256, 0, s>>>(packed, valid, dedup ? dupmask : nullptr, n_groups, v, hi, 0, 0u, tabs);
            v = hi + 1;
        }
    }
    rc = kmap_check_launch("count_all_k(scan)");
    if (rc) return rc;
    mark(1);
    const uint32_t* hide = dedup ? dupmask : nullptr;
    if (zeroed) {                                                                      // the zero fill of the level-kmax table
        cudaStreamWaitEvent(s, zeroed, 0);
        cudaEventDestroy(zeroed);
        zeroed = nullptr;
    }
    if (use_partition) {
        // level kmax through key partitioning + shared-memory counters (partition.cu).  Run-end corrections: fused into the
        // histogram pass, except that a level whose table is beyond L2 -- level 13 under k = 14 -- travels through the
        // partition itself as extra buckets.  With `merge`, the tables are all-reduced from inside, as they become final.
        KMAP_REQUIRE(part_scratch_bytes >= kmap_partition_scratch_bytes(n, kmax), "partition scratch too small");
        rc = kmap_count_partitioned(packed, valid, hide, local ? n : 0, kmax, tabs.t[kmax], part_scratch, kmin < kmax ? &tabs : nullptr, kmin,
                                    phase_events ? phase_events + 4 : nullptr, s, merge);
        if (rc) return rc;
    } else {
        // level kmax in key-prefix passes: 4^PB passes, each updating a 4^(kmax-PB)-cell slice that stays in L2
        int PB = 0;
        if (n_partitions <= 0) { while (PB < 3 && PB < kmax && (((size_t)4 << (2 * kmax)) >> (2 * PB)) > ((size_t)96 << 20)) ++PB; }
        else { while (PB < 3 && PB < kmax && (1 << (2 * PB)) < n_partitions) ++PB; }
        const int64_t n_groups = (n_words + 3) / 4;
        const unsigned int gB = grid_for(n_groups, 256);
        for (uint32_t prefix = 0; local && prefix < (1u << (2 * PB)); ++prefix) {
            switch (PB) {
                case 0: count_prefix_kernel<0><<<gB, 256, 0, s>>>(packed, valid, hide, n_groups, kmax, tabs.t[kmax], prefix); break;
                case 1: count_prefix_kernel<1><<<gB, 256, 0, s>>>(packed, valid, hide, n_groups, kmax, tabs.t[kmax], prefix); break;
                case 2: count_prefix_kernel<2><<<gB, 256, 0, s>>>(packed, valid, hide, n_groups, kmax, tabs.t[kmax], prefix); break;
                default: count_prefix_kernel<3><<<gB, 256, 0, s>>>(packed, valid, hide, n_groups, kmax, tabs.t[kmax], prefix); break;
            }
        }
        rc = kmap_check_launch("count_all_k(count)");
        if (rc) return rc;
        if (merge) {               // no early slices on this path: every pre-derive buffer after the count
            if ((rc = OneShotEvent::chain(s, merge->stream))) return rc;
            for (int k = kmax; k >= kmin && !rc; --k) rc = kmap_merge_table_on(tabs.t[k], (int64_t)1 << (2 * k), merge);
            if (rc) return rc;
        }
    }
    if (merge && (rc = OneShotEvent::chain(merge->stream, s))) return rc;        // the reductions read merged buffers
    mark(2);
    const bool scattered = merge && merge->scatter && merge->world > 1;      // this rank owns (and reduces) one key range of every level
    for (int k = kmax - 1; k >= kmin; --k) {
        int64_t cells = (int64_t)1 << (2 * k), lo = 0;
        if (scattered) { cells /= merge->world; lo = cells * merge->rank; }
        int64_t g = (cells + 255) / 256;
        if (g > 148 * 32) g = 148 * 32;
        derive_table_kernel<<<(unsigned int)g, 256, 0, s>>>(reinterpret_cast<const uint4*>(tabs.t[k + 1]) + lo, tabs.t[k] + lo, cells);
    }
    rc = kmap_check_launch("count_all_k(derive)");
    if (rc) return rc;
    mark(3);
    if (dedup) {
        uint32_t counts[4] = {0, 0, 0, 0};
        e = cudaMemcpyAsync(counts, work, 16, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { kmap_set_error("count_all_k: %s", cudaGetErrorString(e)); return (int)e; }
        const bool mine = counts[0] || counts[1];
        if (!merge) {
            for (int k = kmin; mine && k <= kmax; ++k) {
                rc = kmap_count_long_reads(packed, valid, n, borders, n_seq, k, tabs.t[k], work, bitmap, counts, s);
                if (rc) return rc;
            }
        } else if (counts[2] || counts[3]) {
            // some rank holds long reads: their per-level counts go into tables of their own, which are merged and added
            size_t total = 0;
            for (int k = kmin; k <= kmax; ++k) total += (size_t)1 << (2 * k);
            uint32_t* delta = nullptr;
            e = cudaMallocAsync(reinterpret_cast<void**>(&delta), total * 4, s);
            if (e == cudaSuccess) e = cudaMemsetAsync(delta, 0, total * 4, s);
            if (e != cudaSuccess) { kmap_set_error("count_all_k(long reads of a sharded input): %s", cudaGetErrorString(e)); return (int)e; }
            size_t o = 0;
            for (int k = kmin; k <= kmax; ++k) {
                if (mine) {
                    rc = kmap_count_long_reads(packed, valid, n, borders, n_seq, k, delta + o, work, bitmap, counts, s);
                    if (rc) break;
                }
                o += (size_t)1 << (2 * k);
            }
            if (!rc) rc = OneShotEvent::chain(s, merge->stream);
            if (!rc) rc = kmap_allreduce_u32_on(delta, (int64_t)total, merge->comm, merge->stream);
            if (!rc) rc = OneShotEvent::chain(merge->stream, s);
            o = 0;
            for (int k = kmin; k <= kmax && !rc; ++k) {
                const int64_t cells = (int64_t)1 << (2 * k);
                int64_t g = (cells + 255) / 256;
                if (g > 148 * 32) g = 148 * 32;
                add_tables_kernel<<<(unsigned int)g, 256, 0, s>>>(tabs.t[k], delta + o, cells);
                o += (size_t)cells;
            }
            cudaFreeAsync(delta, s);
            if (rc) return rc;
            rc = kmap_check_launch("count_all_k(long reads)");
            if (rc) return rc;
        }
    }
    return KMAP_OK;
}

}  // namespace

extern "C" int kmap_count_all_k(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                                int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                                uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                                void* const* phase_events, void* stream) {
    return count_all_impl(packed, valid, n, borders, n_seq, kmin, kmax, dedup, tables_host, dupmask, work, bitmap, n_partitions, scheme,
                          part_scratch, part_scratch_bytes, phase_events, stream, nullptr);
}

extern "C" int kmap_count_all_k_sharded(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                                        int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                                        uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                                        void* const* phase_events, void* stream, void* comm, void* comm_stream) {
    KMAP_REQUIRE(comm && comm_stream, "the sharded count needs a communicator and a stream for the exchange");
    const KmapMerge merge = {comm, as_stream(comm_stream), 0, 0, kmap_comm_world(comm)};
    return count_all_impl(packed, valid, n, borders, n_seq, kmin, kmax, dedup, tables_host, dupmask, work, bitmap, n_partitions, scheme,
                          part_scratch, part_scratch_bytes, phase_events, stream, &merge);
}

// The same with the merged tables left SCATTERED over the ranks by key range: rank r owns cells [r * 4^k / world,
// (r + 1) * 4^k / world) of every level k (the other cells of its buffers hold partial sums).  Reduce-scatter instead of
// all-reduce halves the exchange and the reductions to the lower levels run on the owned ranges only; the compaction of a
// range is kmap_compact_merge_range, the ranges of ranks 0, 1, .. concatenate to the reference's list.  world must divide 4^kmin.
extern "C" int kmap_count_all_k_scattered(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                                          int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                                          uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                                          void* const* phase_events, void* stream, void* comm, void* comm_stream, int rank, int world) {
    KMAP_REQUIRE(comm && comm_stream, "the sharded count needs a communicator and a stream for the exchange");
    KMAP_REQUIRE(world >= 1 && rank >= 0 && rank < world && kmin >= 1 && (((int64_t)1 << (2 * kmin)) % world) == 0,
                 "the number of ranks must divide 4^kmin");
    const KmapMerge merge = {comm, as_stream(comm_stream), 1, rank, world};
    return count_all_impl(packed, valid, n, borders, n_seq, kmin, kmax, dedup, tables_host, dupmask, work, bitmap, n_partitions, scheme,
                          part_scratch, part_scratch_bytes, phase_events, stream, &merge);
}
