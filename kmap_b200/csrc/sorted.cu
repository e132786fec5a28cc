// The sort / run-length counting path: comp_kmer_hash_taichi + remove_duplicate_hash_per_seq + count_uniq_hash +
// merge_revcom (kmer_count.py:449-491, 643-685, 743-760) for k-mers whose dense 4^k table is out of reach
// (16 <= k <= 31: uint64 hashes, int64 counts, kmer_count.py:351-365), and the list-based Hamming-ball kernels of
// find_motif / ex_hamball for those hashes (motif_discovery.py:666-673, 959-986).
//
// It follows the reference's own data flow -- one hash per position, duplicates inside a read turned into the invalid
// hash, np.unique -- with every step on the device:
//   window_keys_kernel          one 64-bit key per position from the packed reads (all-ones = invalid: the window leaves
//                               the array or touches a 255), coalesced 8-byte stores
//   dedup_keys_warp_kernel      one warp = one read (<= 256 windows): first occurrences win a slot of a per-warp
//                               shared-memory set, later ones become invalid; longer reads go to the multi-pass block
//                               kernel (dedup_keys_block_kernel), any length
//   LSD radix sort              8 bits per pass, ceil(2k / 8) passes: per-tile digit histogram -> per-digit scan over the
//                               tiles -> stable scatter (ranks inside a tile from __match_any_sync).  The first pass drops
//                               the invalid keys, so the later ones only move real k-mers.
//   run-length encoding         heads of runs -> unique keys (ascending, like np.unique) and run lengths
//   merge_revcom                partner = binary search of rc(h) in the sorted list; survivors keep the list order, value
//                               min(h, rc h), a palindrome is its own partner (count doubled) -- SURVEY Q6
#include "common.cuh"
#include <cstdlib>
#include "tile.cuh"

namespace {

constexpr unsigned long long KEY_EMPTY = ~0ull;
typedef unsigned long long u64;

__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- keys -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_keys_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                          int64_t n, int k, u64* __restrict__ keys) {
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    u64 key = KEY_EMPTY;
    if (window_ok(valid, p, k)) key = window32(packed, p) >> (64 - 2 * k);      // (the zero padding makes windows past n invalid)
    keys[p] = key;
}

// ---- per-read de-duplication of a key / hash array ---------------------------------------------------------------------
constexpr int DW_WARPS = 8;
constexpr int DW_SLOTS = 512;                    // per warp
constexpr int DW_MAX = 256;                      // positions a warp handles (load factor <= 0.5)
constexpr int DW_PRE = 4;                        // rounds of 32 keys fetched one read ahead

// reads of at most DW_MAX positions; the others are listed in long_ids (count in *n_long)
template <typename H>
__global__ void __launch_bounds__(DW_WARPS * 32) dedup_keys_warp_kernel(H* __restrict__ keys, int64_t n, const int64_t* __restrict__ borders,
                                                                        int64_t n_seq, uint32_t* __restrict__ n_long,
                                                                        uint32_t* __restrict__ long_ids) {
    __shared__ H sets[DW_WARPS][DW_SLOTS];
    const H empty = (H)~(H)0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    H* set = sets[wib];
    const int64_t n_warps = (int64_t)gridDim.x * DW_WARPS;
    // Software pipeline over this warp's reads: the borders are requested two reads ahead and the first DW_PRE rounds of
    // keys one read ahead, so that the two dependent memory latencies (borders -> keys) overlap the work on earlier reads.
    // (Another read's keys are never written by this warp, so the early loads see what they should.)
    struct Staged { int64_t st; int64_t len; H h[DW_PRE]; };
    auto load_borders = [&](int64_t r) {
        return r < n_seq ? __ldg(reinterpret_cast<const longlong2*>(borders) + r) : make_longlong2(0, 0);
    };
    auto stage = [&](const longlong2& b, Staged& g) {
        const int64_t st = b.x < 0 ? 0 : b.x, en = b.y > n ? n : b.y;
        g.st = st;
        g.len = en - st;
#pragma unroll
        for (int q = 0; q < DW_PRE; ++q) {
            const int64_t i = q * 32 + lane;
            g.h[q] = (g.len <= DW_MAX && i < g.len) ? keys[st + i] : empty;
        }
    };
    const int64_t r0 = (int64_t)blockIdx.x * DW_WARPS + wib;
    Staged cur, nxt;
    stage(load_borders(r0), cur);
    longlong2 raw1 = load_borders(r0 + n_warps);
    for (int64_t r = r0; r < n_seq; r += n_warps) {
        const longlong2 raw2 = load_borders(r + 2 * n_warps);
        stage(raw1, nxt);
        raw1 = raw2;
        const Staged g = cur;
        cur = nxt;
        const int64_t st = g.st, len = g.len;
        if (len <= 1) continue;
        if (len > DW_MAX) {
            if (lane == 0) long_ids[atomicAdd(n_long, 1u)] = (uint32_t)r;
            continue;
        }
        // slots = the power of two >= 2 * len (load factor <= 0.5), at least 64: short reads clear and probe a small set
        uint32_t slots = 64;
        while (slots < 2u * (uint32_t)len) slots <<= 1;
        const uint32_t slot_mask = slots - 1u;
        for (uint32_t j = lane; j < slots; j += 32) set[j] = empty;
        __syncwarp();
#pragma unroll 1
        for (int q = 0; q * 32 < len; ++q) {
            const int64_t i = q * 32 + lane;
            H h = empty;
            if (q < DW_PRE) {
#pragma unroll
                for (int u = 0; u < DW_PRE; ++u) if (u == q) h = g.h[u];
            } else if (i < len) {
                h = keys[st + i];
            }
            const bool live = h != empty;
            // one lane per distinct key of this round: the lowest one (it is the first occurrence inside the round)
            const H probe = live ? h : (H)(((H)1 << (8 * sizeof(H) - 1)) | (H)lane);     // distinct non-keys for idle lanes (keys have < 64 bits)
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, probe);
            const bool leader = live && (peers & ((1u << lane) - 1u)) == 0;
            bool dup = live && !leader;
            // warp-uniform probing loop (predicated body: per-lane breaks out of a divergent loop cost more than the probes)
            bool pending = leader;
            uint32_t slot = (uint32_t)mix64((u64)h) & slot_mask;
            while (__any_sync(0xFFFFFFFFu, pending)) {
                const H seen = pending ? reinterpret_cast<volatile H*>(set)[slot] : h;      // only leaders of distinct keys write in this round
                H old = seen;
                if (pending && seen == empty) old = atomicCAS(&set[slot], empty, h);
                const bool claimed = pending && seen == empty && old == empty;
                const bool found = pending && old == h;                  // inserted in an earlier round
                dup = dup || found;
                pending = pending && !claimed && !found;
                slot = (slot + 1) & slot_mask;
            }
            if (dup) keys[st + i] = empty;
            __syncwarp();
        }
        __syncwarp();
    }
}

// any length: one block per listed read (or per read when ids == NULL); the set lives in shared memory and the read is
// taken in n_pass passes, pass p owning the keys with mix(h) % n_pass == p.  Keeps the FIRST occurrence
// (kmer_count.py:755-759: np.unique(return_index)).
constexpr int DB_SLOTS = 4096;
template <typename H>
__global__ void __launch_bounds__(256) dedup_keys_block_kernel(H* __restrict__ keys, int64_t n, const int64_t* __restrict__ borders,
                                                               int64_t n_seq, const uint32_t* __restrict__ ids,
                                                               const uint32_t* __restrict__ n_ids) {
    __shared__ H skeys[DB_SLOTS];
    __shared__ uint32_t firsts[DB_SLOTS];
    const H empty = (H)~(H)0;
    const int64_t n_reads = ids ? (int64_t)*n_ids : n_seq;
    for (int64_t q = blockIdx.x; q < n_reads; q += gridDim.x) {
        const int64_t r = ids ? (int64_t)ids[q] : q;
        int64_t st = borders[2 * r], en = borders[2 * r + 1];
        if (st < 0) st = 0;
        if (en > n) en = n;
        const int64_t len = en - st;
        if (len <= 1) continue;
        const uint32_t n_pass = (uint32_t)((len + DB_SLOTS / 2 - 1) / (DB_SLOTS / 2));
        for (uint32_t pass = 0; pass < n_pass; ++pass) {
            for (int i = threadIdx.x; i < DB_SLOTS; i += blockDim.x) { skeys[i] = empty; firsts[i] = 0xFFFFFFFFu; }
            __syncthreads();
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                const H h = keys[st + i];
                if (h == empty) continue;
                const u64 m = mix64((u64)h);
                if (n_pass > 1 && (uint32_t)(m >> 32) % n_pass != pass) continue;
                uint32_t slot = (uint32_t)m & (DB_SLOTS - 1);
                while (true) {
                    const H old = atomicCAS(&skeys[slot], empty, h);
                    if (old == empty || old == h) { atomicMin(&firsts[slot], (uint32_t)i); break; }
                    slot = (slot + 1) & (DB_SLOTS - 1);
                }
            }
            __syncthreads();
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                const H h = keys[st + i];
                if (h == empty) continue;
                const u64 m = mix64((u64)h);
                if (n_pass > 1 && (uint32_t)(m >> 32) % n_pass != pass) continue;
                uint32_t slot = (uint32_t)m & (DB_SLOTS - 1);
                while (skeys[slot] != h) slot = (slot + 1) & (DB_SLOTS - 1);
                if (firsts[slot] != (uint32_t)i) keys[st + i] = empty;
            }
            __syncthreads();
        }
    }
}

// ---- LSD radix sort of 64-bit keys -----------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;        // 4096 keys per tile
constexpr int RS_WARP_KEYS = RS_TILE / (RS_THREADS / 32);   // 512 consecutive keys per warp

// ---- onesweep: one histogram read for all passes, then ONE kernel per 8-bit digit ----------------------------------------------
// A histogram -> scan -> scatter pass reads the keys twice per digit and runs two small scans in between (that was round 1's
// form: 2.72 ms for the 4.1e7 keys that take 2.15 ms here).  Onesweep (Adinets & Merrill): (a) one kernel reads the keys once and histograms every digit position at the same time;
// (b) per digit position one kernel does count + inter-tile prefix + scatter: a tile publishes its per-digit counts in a
// status word (flag | value in one 64-bit word, so the word itself is the message and no fence is needed), then looks back
// over its predecessors until it meets one that has published an inclusive prefix (decoupled look-back).  Tiles are
// numbered by an atomic ticket, so every predecessor of a running tile is running or done and the look-back cannot
// deadlock.  Traffic: 8 B/key once + 16 B/key per digit, against 24 B/key per digit.
// Stable scatter inside a tile: warp w owns keys [w * 512, (w + 1) * 512) in 16 rounds of 32 consecutive keys; inside a round
// the rank among equal digits comes from __match_any_sync, across rounds from a per-warp counter, across warps from the tile's
// digit scan.  The keys are first put in digit order in shared memory, so that the write-out stores runs of consecutive
// addresses (one run per digit present in the tile) instead of one scattered 8-byte store per key.
constexpr u64 OS_LOCAL = 1ull << 62, OS_PREFIX = 2ull << 62, OS_VALUE = (1ull << 62) - 1ull;
constexpr int OS_MAX_PASSES = 8;

__global__ void __launch_bounds__(256) onesweep_hist_kernel(const u64* __restrict__ in, int64_t n, int passes, u64* __restrict__ hist) {
    __shared__ uint32_t sh[OS_MAX_PASSES][256];
    for (int p = 0; p < passes; ++p) sh[p][threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += 4 * stride) {
        u64 key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) key[u] = i0 + u * stride < n ? __ldg(in + i0 + u * stride) : KEY_EMPTY;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (key[u] == KEY_EMPTY) continue;
            for (int p = 0; p < passes; ++p) atomicAdd(&sh[p][(uint32_t)(key[u] >> (8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p)
        if (sh[p][threadIdx.x]) atomicAdd(hist + p * 256 + threadIdx.x, (u64)sh[p][threadIdx.x]);
}

// hist[p][d] -> first output index of digit d in pass p; *n_valid = number of non-empty keys
__global__ void __launch_bounds__(256) onesweep_bases_kernel(u64* __restrict__ hist, int passes, u64* __restrict__ n_valid) {
    __shared__ u64 v[256];
    for (int p = 0; p < passes; ++p) {
        v[threadIdx.x] = hist[p * 256 + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 run = 0;
            for (int d = 0; d < 256; ++d) { const u64 t = v[d]; v[d] = run; run += t; }
            if (p == 0) *n_valid = run;
        }
        __syncthreads();
        hist[p * 256 + threadIdx.x] = v[threadIdx.x];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS, 3) onesweep_pass_kernel(const u64* __restrict__ in, u64* __restrict__ out, int64_t n_upper,
                                                                   const u64* __restrict__ n_dev, int shift, int drop_empty,
                                                                   const u64* __restrict__ digit_base, u64* __restrict__ status,
                                                                   unsigned int* __restrict__ ticket) {
    __shared__ u64 skeys[RS_TILE];                         // the tile in digit order
    __shared__ uint32_t wcnt[RS_THREADS / 32][256];        // per (warp, digit): count, then start inside the tile
    __shared__ u64 gdelta[256];                            // global index of the digit's run - its start inside the tile
    __shared__ uint32_t scan_ws[RS_THREADS / 32];
    __shared__ uint32_t tile_total;
    __shared__ unsigned int s_tile;
    const int64_t n = n_dev ? (int64_t)*n_dev : n_upper;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int q = 0; q < RS_THREADS / 32; ++q) wcnt[q][threadIdx.x] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile * RS_TILE >= n) return;                       // (block-uniform; nobody looks back at a tile beyond the keys)
    const int64_t base = tile * RS_TILE + (int64_t)w * RS_WARP_KEYS;
    u64 key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];          // while the counters are being filled: count before this round | leader lane << 16 | rank among peers << 21
    uint32_t live = 0;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        key[r] = i < n ? __ldcs(in + i) : KEY_EMPTY;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool ok = i < n && !(drop_empty && key[r] == KEY_EMPTY);
        const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? d : (256u + (uint32_t)lane));
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        uint32_t old = 0;
        if (ok && (uint32_t)lane == leader) old = atomicAdd(&wcnt[w][d], (uint32_t)__popc(peers));
        rank[r] = old | (leader << 16) | ((uint32_t)__popc(peers & ((1u << lane) - 1u)) << 21);
        if (ok) live |= 1u << r;
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t before = __shfl_sync(0xFFFFFFFFu, rank[r] & 0xFFFFu, (rank[r] >> 16) & 31u);
        rank[r] = before + (rank[r] >> 21);
    }
    __syncthreads();
    {
        // thread d: digit d's keys of the warps in order; an exclusive scan over the digits gives the tile layout, the
        // look-back over the earlier tiles gives the digit's place in the output
        const uint32_t d = threadIdx.x;
        uint32_t cnt_w[RS_THREADS / 32];
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < RS_THREADS / 32; ++q) { cnt_w[q] = wcnt[q][d]; mine += cnt_w[q]; }
        volatile u64* st = status + (size_t)tile * 256 + d;
        *st = (tile == 0 ? OS_PREFIX : OS_LOCAL) | (u64)mine;
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) scan_ws[w] = incl;
        __syncthreads();
        uint32_t pre = 0, all = 0;
#pragma unroll
        for (int q = 0; q < RS_THREADS / 32; ++q) { if (q < w) pre += scan_ws[q]; all += scan_ws[q]; }
        uint32_t start = pre + incl - mine;                 // first slot of digit d in the tile
        if (threadIdx.x == 0) tile_total = all;
        u64 before = 0;                                     // keys with digit d in the earlier tiles
        if (tile > 0) {
            // Look-back, eight predecessors per step with their status words requested together: the walk is a chain of
            // dependent L2 round trips otherwise, and the time a tile spends in it sets how far back the next tiles must walk.
            int64_t t = tile - 1;
            bool done = false;
            while (!done) {
                u64 sv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int64_t tt = t - u;
                    sv[u] = tt >= 0 ? *(const volatile u64*)(status + (size_t)tt * 256 + d) : OS_PREFIX;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (done) break;
                    u64 x = sv[u];
                    while ((x >> 62) == 0) x = *(const volatile u64*)(status + (size_t)(t - u) * 256 + d);    // not published yet
                    before += x & OS_VALUE;
                    if ((x >> 62) == 2) done = true;
                }
                t -= 8;
            }
            *st = OS_PREFIX | (before + (u64)mine);
        }
        gdelta[d] = digit_base[d] + before - start;
#pragma unroll
        for (int q = 0; q < RS_THREADS / 32; ++q) { wcnt[q][d] = start; start += cnt_w[q]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
        if ((live >> r) & 1u) skeys[wcnt[w][(uint32_t)(key[r] >> shift) & 255u] + rank[r]] = key[r];
    __syncthreads();
    const uint32_t total = tile_total;
    for (uint32_t i = threadIdx.x; i < total; i += RS_THREADS) {
        const u64 kk = skeys[i];
        out[gdelta[(uint32_t)(kk >> shift) & 255u] + i] = kk;
    }
}

// ---- run-length encoding of the sorted keys ------------------------------------------------------------------------------------
constexpr int RL_BLOCK = 256;
constexpr int RL_ITEMS = 8;
constexpr int RL_TILE = RL_BLOCK * RL_ITEMS;

__device__ __forceinline__ uint32_t rl_block_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[RL_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    uint32_t pre = 0, all = 0;
    for (int q = 0; q < RL_BLOCK / 32; ++q) { if (q < w) pre += warp_sums[q]; all += warp_sums[q]; }
    *total = all;
    __syncthreads();
    return pre + incl - v;
}

__device__ __forceinline__ bool is_head(const u64* __restrict__ keys, int64_t i) { return i == 0 || __ldg(keys + i) != __ldg(keys + i - 1); }

__global__ void __launch_bounds__(RL_BLOCK) head_count_kernel(const u64* __restrict__ keys, int64_t n, uint64_t* __restrict__ tile_counts) {
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) if (base + j < n) c += is_head(keys, base + j);
    uint32_t total;
    rl_block_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(RL_BLOCK) head_write_kernel(const u64* __restrict__ keys, int64_t n, const uint64_t* __restrict__ tile_offsets,
                                                              u64* __restrict__ kh_out, long long* __restrict__ pos_out) {
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    uint32_t heads = 0, c = 0;
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) if (base + j < n && is_head(keys, base + j)) { heads |= 1u << j; ++c; }
    uint32_t total;
    uint64_t o = tile_offsets[blockIdx.x] + rl_block_scan(c, &total);
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j)
        if ((heads >> j) & 1u) { kh_out[o] = keys[base + j]; pos_out[o] = base + j; ++o; }
}

__global__ void __launch_bounds__(256) run_length_kernel(const long long* __restrict__ pos, int64_t n_uniq, long long n, long long* __restrict__ cnt) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= n_uniq) return;
    cnt[r] = (r + 1 < n_uniq ? pos[r + 1] : n) - pos[r];
}

__global__ void __launch_bounds__(1024) scan_tiles64_kernel(uint64_t* __restrict__ v, int64_t n_tiles) {
    __shared__ uint64_t partial[1024];
    const int64_t chunk = (n_tiles + 1023) / 1024;
    const int64_t lo = (int64_t)threadIdx.x * chunk;
    const int64_t hi = lo + chunk < n_tiles ? lo + chunk : n_tiles;
    uint64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += v[i];
    partial[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const uint64_t t = partial[i]; partial[i] = run; run += t; }
        v[n_tiles] = run;
    }
    __syncthreads();
    uint64_t run = partial[threadIdx.x];
    for (int64_t i = lo; i < hi; ++i) { const uint64_t t = v[i]; v[i] = run; run += t; }
}

// ---- merge_revcom on a sorted unique list ----------------------------------------------------------------------------------------
// index of x in the ascending array kh[0..n), or -1
__device__ __forceinline__ int64_t find_key(const u64* __restrict__ kh, int64_t n, u64 x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(kh + mid) < x) lo = mid + 1; else hi = mid;
    }
    return (lo < n && __ldg(kh + lo) == x) ? lo : -1;
}

// cnt_out[index of kh[i] in uniq] += cnt[i]: sums the per-rank (hash, count) lists of a sharded count into the list of the
// distinct hashes of all ranks (count_uniq_hash over the whole input = the per-shard results added up, kmer_count.py:476-491)
__global__ void __launch_bounds__(256) list_add_counts_kernel(const u64* __restrict__ uniq, int64_t n_uniq, const u64* __restrict__ kh,
                                                              const long long* __restrict__ cnt, int64_t n, unsigned long long* __restrict__ cnt_out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (; i < n; i += stride) {
        const int64_t j = find_key(uniq, n_uniq, __ldg(kh + i));
        if (j >= 0) atomicAdd(cnt_out + j, (unsigned long long)__ldg(cnt + i));
    }
}

// partner[i] = index of rc(kh[i]) in the list or -1; tile_counts = survivors per tile
// (entry i is dropped iff its partner is present and kh[i] > rc(kh[i]), kmer_count.py:668-676)
__global__ void __launch_bounds__(RL_BLOCK) merge_plan_kernel(const u64* __restrict__ kh, int64_t n, int k, bool keep_higher,
                                                              long long* __restrict__ partner, uint64_t* __restrict__ tile_counts) {
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    uint32_t c = 0;
    for (int j = 0; j < RL_ITEMS; ++j) {
        const int64_t i = base + j;
        if (i < n) {
            const u64 h = __ldg(kh + i), rc = revcom64(h, k);
            const int64_t p = find_key(kh, n, rc);
            partner[i] = p;
            c += !(p >= 0 && (keep_higher ? h < rc : h > rc));
        }
    }
    uint32_t total;
    rl_block_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// survivors in list order: (min(h, rc h), cnt[h] + cnt[rc h]); summed (may be NULL) gets cnt[i] + cnt[partner] for EVERY i,
// which is what the reference leaves in the caller's count array (kmer_count.py:661)
__global__ void __launch_bounds__(RL_BLOCK) merge_write_kernel(const u64* __restrict__ kh, const long long* __restrict__ cnt, int64_t n, int k,
                                                               bool keep_higher, const long long* __restrict__ partner, const uint64_t* __restrict__ tile_offsets,
                                                               u64* __restrict__ kh_out, long long* __restrict__ cnt_out,
                                                               long long* __restrict__ summed) {
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    u64 val[RL_ITEMS];
    long long sum[RL_ITEMS];
    uint32_t keep = 0, c = 0;
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        const int64_t i = base + j;
        val[j] = 0; sum[j] = 0;
        if (i < n) {
            const u64 h = __ldg(kh + i), rc = revcom64(h, k);
            const long long p = partner[i];
            sum[j] = cnt[i] + (p >= 0 ? cnt[p] : 0ll);
            val[j] = (h < rc) != keep_higher ? h : rc;
            if (!(p >= 0 && (keep_higher ? h < rc : h > rc))) { keep |= 1u << j; ++c; }
        }
    }
    uint32_t total;
    uint64_t o = tile_offsets[blockIdx.x] + rl_block_scan(c, &total);
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        if (summed && base + j < n) summed[base + j] = sum[j];
        if ((keep >> j) & 1u) { kh_out[o] = val[j]; cnt_out[o] = sum[j]; ++o; }
    }
}

// ---- Hamming-ball sums and extraction on (uint64 kh, int64 cnt) lists --------------------------------------------------------------
constexpr int HL_BLOCK = 256;
constexpr int HL_MAXM = 16;

__global__ void __launch_bounds__(HL_BLOCK) hamball_sum_list64_kernel(const u64* __restrict__ kh, const long long* __restrict__ cnt, int64_t n,
                                                                      int k, const u64* __restrict__ cand, int m, int d, int revcom,
                                                                      u64* __restrict__ sums) {
    __shared__ u64 sc[HL_MAXM], src[HL_MAXM];
    __shared__ u64 ws[HL_BLOCK / 32];
    const u64 low = lowmask64(k);
    if (threadIdx.x < m) { const u64 c = cand[threadIdx.x] & low; sc[threadIdx.x] = c; src[threadIdx.x] = revcom64(c, k); }
    __syncthreads();
    u64 acc[HL_MAXM];
#pragma unroll
    for (int i = 0; i < HL_MAXM; ++i) acc[i] = 0;
    for (int64_t j = (int64_t)blockIdx.x * HL_BLOCK + threadIdx.x; j < n; j += (int64_t)gridDim.x * HL_BLOCK) {
        const u64 h = __ldg(kh + j);
        const u64 wgt = (u64)__ldg(cnt + j);
#pragma unroll
        for (int i = 0; i < HL_MAXM; ++i) {
            if (i < m) {
                uint32_t dist = nz_groups64(h ^ sc[i], low);
                if (revcom) { const uint32_t rd = nz_groups64(h ^ src[i], low); dist = rd < dist ? rd : dist; }
                if ((int)dist <= d) acc[i] += wgt;
            }
        }
    }
    for (int i = 0; i < m; ++i) {
        u64 v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 t = 0;
            for (int q = 0; q < HL_BLOCK / 32; ++q) t += ws[q];
            if (t) atomicAdd(sums + i, t);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ bool ball_member64(u64 h, u64 conseq, u64 rc_conseq, int k, int d, int revcom, u64 low, u64* value) {
    uint32_t dist = nz_groups64(h ^ conseq, low);
    bool flip = false;
    if (revcom) {
        const uint32_t rd = nz_groups64(h ^ rc_conseq, low);
        flip = rd < dist;                           // ties stay forward (motif_discovery.py:965)
        dist = rd < dist ? rd : dist;
    }
    if ((int)dist > d) return false;
    *value = flip ? revcom64(h & low, k) : h;
    return true;
}

__global__ void __launch_bounds__(RL_BLOCK) ball64_count_kernel(const u64* __restrict__ kh, int64_t n, u64 conseq, u64 rc_conseq, int k, int d,
                                                                int revcom, uint64_t* __restrict__ tile_counts) {
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    const u64 low = lowmask64(k);
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        u64 v;
        if (base + j < n) c += ball_member64(__ldg(kh + base + j), conseq, rc_conseq, k, d, revcom, low, &v);
    }
    uint32_t total;
    rl_block_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(RL_BLOCK) ball64_write_kernel(const u64* __restrict__ kh, const long long* __restrict__ cnt, int64_t n, u64 conseq,
                                                                u64 rc_conseq, int k, int d, int revcom, const uint64_t* __restrict__ tile_offsets,
                                                                u64* __restrict__ kh_out, long long* __restrict__ cnt_out, u64* __restrict__ cnt_mat) {
    __shared__ u64 smat[4 * 32];
    if (threadIdx.x < 128) smat[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    const u64 low = lowmask64(k);
    u64 vals[RL_ITEMS];
    long long cnts[RL_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        u64 v;
        if (base + j < n && ball_member64(__ldg(kh + base + j), conseq, rc_conseq, k, d, revcom, low, &v)) {
            vals[c] = v; cnts[c] = __ldg(cnt + base + j); ++c;
        }
    }
    uint32_t total;
    const uint64_t o = tile_offsets[blockIdx.x] + rl_block_scan(c, &total);
    for (uint32_t j = 0; j < c; ++j) {
        if (kh_out) { kh_out[o + j] = vals[j]; cnt_out[o + j] = cnts[j]; }
        for (int pos = 0; pos < k; ++pos) {
            const uint32_t b = (uint32_t)(vals[j] >> (2 * (k - 1 - pos))) & 3u;
            atomicAdd(&smat[b * 32 + pos], (u64)cnts[j]);
        }
    }
    __syncthreads();
    if (total && threadIdx.x < 128) {
        const int b = threadIdx.x >> 5, pos = threadIdx.x & 31;
        if (pos < k && smat[threadIdx.x]) atomicAdd(&cnt_mat[b * k + pos], smat[threadIdx.x]);
    }
}

int read_u64(const void* dev, int64_t* host, cudaStream_t s, const char* what) {
    uint64_t t = 0;
    cudaError_t e = cudaMemcpyAsync(&t, dev, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { kmap_set_error("%s: %s", what, cudaGetErrorString(e)); return (int)e; }
    *host = (int64_t)t;
    return KMAP_OK;
}

int64_t rs_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }
int64_t rl_tiles(int64_t n) { return (n + RL_TILE - 1) / RL_TILE; }

template <typename H>
int dedup_keys(H* keys, int64_t n, const int64_t* borders, int64_t n_seq, uint32_t* work, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(work, 0, 16, s);
    if (e != cudaSuccess) { kmap_set_error("dedup_keys: %s", cudaGetErrorString(e)); return (int)e; }
    int64_t blocks = (n_seq + DW_WARPS - 1) / DW_WARPS;
    if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;
    dedup_keys_warp_kernel<H><<<(unsigned int)blocks, DW_WARPS * 32, 0, s>>>(keys, n, borders, n_seq, work, work + 4);
    dedup_keys_block_kernel<H><<<148 * 4, 256, 0, s>>>(keys, n, borders, n_seq, work + 4, work);
    return kmap_check_launch("dedup_keys");
}

}  // namespace

extern "C" {

int kmap_window_keys_u64(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint64_t* keys, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 31, "the 64-bit key path covers 1 <= k <= 31");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && keys, "null pointer");
    window_keys_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(packed, valid, n, k, reinterpret_cast<u64*>(keys));
    return kmap_check_launch("window_keys");
}

int64_t kmap_dedup_keys_work_words(int64_t n_seq) { return 4 + (n_seq < 0 ? 0 : n_seq); }

int kmap_dedup_hash_per_read_u64(uint64_t* hash, int64_t n, const int64_t* borders, int64_t n_seq, uint32_t* work, void* stream) {
    KMAP_REQUIRE(n >= 0 && n_seq >= 0 && n_seq < (int64_t)0xFFFFFFFFll, "bad size");
    if (n == 0 || n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(hash && borders && work, "null pointer");
    return dedup_keys<u64>(reinterpret_cast<u64*>(hash), n, borders, n_seq, work, as_stream(stream));
}

// scratch (uint64 words): the sort's histograms / counters / status words (see kmap_sort_keys_u64); the run-length step
// re-uses the front of it for its tile counts
int64_t kmap_sort_scratch_words(int64_t n) {
    if (n < 0) return 0;
    const int64_t t = rs_tiles(n);
    const int64_t sort_words = OS_MAX_PASSES * 256 + 2 + 256 * t;   // digit histograms, two counters, one status word per (tile, digit)
    const int64_t rl_words = rl_tiles(n) + 2;
    return (sort_words > rl_words ? sort_words : rl_words) + 2;
}

int kmap_sort_keys_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int key_bits, uint64_t* scratch, int64_t* n_valid_host,
                       int64_t* n_unique_host, void* stream) {
    KMAP_REQUIRE(n >= 0 && key_bits >= 1 && key_bits <= 64 && n_valid_host && n_unique_host, "bad argument");
    *n_valid_host = 0; *n_unique_host = 0;
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(keys && tmp && scratch, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t n_tiles = rs_tiles(n);
    u64* n_dev = nullptr;
    u64* a = reinterpret_cast<u64*>(keys);
    u64* b = reinterpret_cast<u64*>(tmp);
    const int passes = (key_bits + 7) / 8;
    {
        // scratch: [0..2047] per-pass digit histograms -> bases, [2048] n_valid (n_dev), [2049] ticket, then the status words
        u64* hist = reinterpret_cast<u64*>(scratch);
        n_dev = hist + OS_MAX_PASSES * 256;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(n_dev + 1);
        u64* status = n_dev + 2;
        cudaMemsetAsync(hist, 0, (size_t)(OS_MAX_PASSES * 256 + 2) * 8, s);
        int64_t hb = (n + 1023) / 1024;
        if (hb > 148 * 8) hb = 148 * 8;
        onesweep_hist_kernel<<<(unsigned int)hb, 256, 0, s>>>(a, n, passes, hist);
        onesweep_bases_kernel<<<1, 256, 0, s>>>(hist, passes, n_dev);
        for (int p = 0; p < passes; ++p) {
            cudaMemsetAsync(status, 0, (size_t)n_tiles * 256 * 8, s);
            cudaMemsetAsync(ticket, 0, 4, s);
            onesweep_pass_kernel<<<(unsigned int)n_tiles, RS_THREADS, 0, s>>>(a, b, n, p == 0 ? nullptr : n_dev, 8 * p, p == 0, hist + p * 256,
                                                                              status, ticket);
            u64* t = a; a = b; b = t;
        }
    }
    int rc = kmap_check_launch("sort_keys");
    if (rc) return rc;
    rc = read_u64(n_dev, n_valid_host, s, "sort_keys");
    if (rc) return rc;
    const int64_t nv = *n_valid_host;
    if (a != reinterpret_cast<u64*>(keys) && nv > 0) {        // odd number of passes: bring the result home
        cudaError_t e = cudaMemcpyAsync(keys, a, (size_t)nv * 8, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) { kmap_set_error("sort_keys: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if (nv == 0) return KMAP_OK;
    const int64_t tiles = rl_tiles(nv);
    head_count_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(reinterpret_cast<const u64*>(keys), nv, scratch);
    scan_tiles64_kernel<<<1, 1024, 0, s>>>(scratch, tiles);
    rc = kmap_check_launch("sort_keys(heads)");
    if (rc) return rc;
    return read_u64(scratch + tiles, n_unique_host, s, "sort_keys(heads)");
}

int kmap_rle_u64(const uint64_t* sorted_keys, int64_t n, uint64_t* scratch, int64_t* pos_scratch, uint64_t* kh_out, int64_t* cnt_out,
                 int64_t capacity, void* stream) {
    KMAP_REQUIRE(n >= 0, "bad argument");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(sorted_keys && scratch && pos_scratch && kh_out && cnt_out, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t tiles = rl_tiles(n);
    const u64* keys = reinterpret_cast<const u64*>(sorted_keys);
    head_count_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(keys, n, scratch);
    scan_tiles64_kernel<<<1, 1024, 0, s>>>(scratch, tiles);
    int rc = kmap_check_launch("rle(count)");
    if (rc) return rc;
    int64_t n_uniq = 0;
    rc = read_u64(scratch + tiles, &n_uniq, s, "rle");
    if (rc) return rc;
    if (capacity < n_uniq) { kmap_set_error("rle: capacity %lld < %lld", (long long)capacity, (long long)n_uniq); return KMAP_ERR_CAPACITY; }
    head_write_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(keys, n, scratch, reinterpret_cast<u64*>(kh_out), reinterpret_cast<long long*>(pos_scratch));
    run_length_kernel<<<grid_for(n_uniq, 256), 256, 0, s>>>(reinterpret_cast<const long long*>(pos_scratch), n_uniq, (long long)n,
                                                            reinterpret_cast<long long*>(cnt_out));
    return kmap_check_launch("rle(write)");
}

int kmap_list_add_counts_u64(const uint64_t* uniq, int64_t n_uniq, const uint64_t* kh, const int64_t* cnt, int64_t n, int64_t* cnt_out,
                             void* stream) {
    KMAP_REQUIRE(n >= 0 && n_uniq >= 0, "negative size");
    if (n == 0 || n_uniq == 0) return KMAP_OK;
    KMAP_REQUIRE(uniq && kh && cnt && cnt_out, "null pointer");
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    list_add_counts_kernel<<<(unsigned int)g, 256, 0, as_stream(stream)>>>(reinterpret_cast<const u64*>(uniq), n_uniq, reinterpret_cast<const u64*>(kh),
                                                                           reinterpret_cast<const long long*>(cnt), n,
                                                                           reinterpret_cast<unsigned long long*>(cnt_out));
    return kmap_check_launch("list_add_counts");
}

int64_t kmap_merge_sorted_scratch_words(int64_t n) { return n < 0 ? 0 : n + rl_tiles(n) + 2; }

int kmap_merge_revcom_sorted_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, int keep_higher, uint64_t* scratch, uint64_t* kh_out,
                                 int64_t* cnt_out, int64_t capacity, int64_t* n_out_host, int64_t* summed_cnt, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 31 && n_out_host, "bad argument (1 <= k <= 31)");
    *n_out_host = 0;
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && cnt && scratch, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t tiles = rl_tiles(n);
    long long* partner = reinterpret_cast<long long*>(scratch);
    uint64_t* tile_counts = scratch + n;
    const u64* khp = reinterpret_cast<const u64*>(kh);
    merge_plan_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(khp, n, k, keep_higher != 0, partner, tile_counts);
    scan_tiles64_kernel<<<1, 1024, 0, s>>>(tile_counts, tiles);
    int rc = kmap_check_launch("merge_revcom_sorted(plan)");
    if (rc) return rc;
    rc = read_u64(tile_counts + tiles, n_out_host, s, "merge_revcom_sorted");
    if (rc) return rc;
    if (capacity < *n_out_host || !kh_out || !cnt_out) {
        if (capacity == 0) return KMAP_OK;           // size query
        kmap_set_error("merge_revcom_sorted: capacity %lld < %lld", (long long)capacity, (long long)*n_out_host);
        return KMAP_ERR_CAPACITY;
    }
    merge_write_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(khp, reinterpret_cast<const long long*>(cnt), n, k, keep_higher != 0, partner, tile_counts,
                                                               reinterpret_cast<u64*>(kh_out), reinterpret_cast<long long*>(cnt_out),
                                                               reinterpret_cast<long long*>(summed_cnt));
    return kmap_check_launch("merge_revcom_sorted(write)");
}

int kmap_hamball_sum_list_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, const uint64_t* cand, int m, int d, int revcom,
                              uint64_t* sums, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 31 && m >= 0 && m <= HL_MAXM && d >= 0 && n >= 0, "bad argument (k <= 31, m <= 16)");
    if (m == 0) return KMAP_OK;
    KMAP_REQUIRE(cand && sums, "null pointer");
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)m * 8, s);
    if (e != cudaSuccess) { kmap_set_error("hamball_sum_list_u64: %s", cudaGetErrorString(e)); return (int)e; }
    if (n == 0) return KMAP_OK;
    int64_t blocks = (n + HL_BLOCK - 1) / HL_BLOCK;
    if (blocks > 148 * 8) blocks = 148 * 8;
    hamball_sum_list64_kernel<<<(unsigned int)blocks, HL_BLOCK, 0, s>>>(reinterpret_cast<const u64*>(kh), reinterpret_cast<const long long*>(cnt),
                                                                        n, k, reinterpret_cast<const u64*>(cand), m, d, revcom,
                                                                        reinterpret_cast<u64*>(sums));
    return kmap_check_launch("hamball_sum_list_u64");
}

int kmap_hamball_extract_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, uint64_t conseq, int d, int revcom, uint64_t* scratch,
                             uint64_t* kh_out, int64_t* cnt_out, int64_t capacity, int64_t* n_out_host, int64_t* cnt_mat, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 31 && n >= 0, "k out of range");
    KMAP_REQUIRE(scratch && n_out_host && cnt_mat, "null pointer");
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(cnt_mat, 0, (size_t)4 * k * 8, s);
    if (e != cudaSuccess) { kmap_set_error("hamball_extract_u64: %s", cudaGetErrorString(e)); return (int)e; }
    *n_out_host = 0;
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && cnt, "null pointer");
    uint64_t rc = 0, com = (~conseq) & lowmask64(k);
    for (int i = 0; i < k; ++i) { rc = (rc << 2) | (com & 3ull); com >>= 2; }
    const int64_t tiles = rl_tiles(n);
    const u64* khp = reinterpret_cast<const u64*>(kh);
    ball64_count_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(khp, n, conseq, rc, k, d, revcom, scratch);
    scan_tiles64_kernel<<<1, 1024, 0, s>>>(scratch, tiles);
    int r = kmap_check_launch("hamball_extract_u64(count)");
    if (r) return r;
    r = read_u64(scratch + tiles, n_out_host, s, "hamball_extract_u64");
    if (r) return r;
    const bool want_list = capacity > 0;
    if (want_list && (capacity < *n_out_host || !kh_out || !cnt_out)) {
        kmap_set_error("hamball_extract_u64: capacity %lld < %lld", (long long)capacity, (long long)*n_out_host);
        return KMAP_ERR_CAPACITY;
    }
    ball64_write_kernel<<<(unsigned int)tiles, RL_BLOCK, 0, s>>>(khp, reinterpret_cast<const long long*>(cnt), n, conseq, rc, k, d, revcom, scratch,
                                                                want_list ? reinterpret_cast<u64*>(kh_out) : nullptr,
                                                                reinterpret_cast<long long*>(cnt_out), reinterpret_cast<u64*>(cnt_mat));
    return kmap_check_launch("hamball_extract_u64(write)");
}

}  // extern "C"
