// Order-preserving stream compaction over the dense table and over (kh, cnt) lists:
//   * count_uniq_hash + merge_revcom (kmer_count.py:476-491, 643-685) regenerated from the forward table
//   * ex_hamball_kh_arr + cal_cnt_mat (motif_discovery.py:959-986) on a merged list
//   * exclusive scan utility
// All three use the same three-kernel shape (tile counts -> scan of tile counts -> ranked write) so that the
// output order is the input order, which the reference's results depend on.
#include "common.cuh"

namespace {

constexpr int CP_BLOCK = 256;
constexpr int CP_ITEMS = 8;                       // consecutive items per thread (2 x 128-bit loads)
constexpr int CP_TILE = CP_BLOCK * CP_ITEMS;      // 2048 items per block

// exclusive scan of one value per thread across the block; returns the thread's offset, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[CP_BLOCK / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < CP_BLOCK / 32 ? warp_sums[lane] : 0;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < CP_BLOCK / 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, si, o);
            if (lane >= o) si += y;
        }
        if (lane < CP_BLOCK / 32) warp_sums[lane] = si - s;
        if (lane == CP_BLOCK / 32 - 1) block_total = si;
    }
    __syncthreads();
    const uint32_t off = warp_sums[w] + incl - v;
    *total = block_total;
    __syncthreads();
    return off;
}

// ---- G[h] = F[rc(h)] as a tiled permutation ------------------------------------------------------------------------------
// The merge needs F[rc(h)] next to F[h]; for a table beyond L2 a gather costs one random 32-byte sector per cell
// (2.7e8 of them at k = 14: 9 of the 12.7 ms the merge used to take).  Write h = [A | M | B] with A, B of 3 bases:
// rc(h) = [rc B | rc M | rc A], so the 4096 cells that share the middle M are the reverse complements of the 4096 cells
// that share the middle rc(M), and both sets consist of 64 runs of 64 consecutive cells (256 B).  One block moves one
// such tile through shared memory: coalesced reads, a 64 x 64 transpose with the 3-base reverse complement applied to
// both indices, coalesced writes.
__device__ __forceinline__ uint32_t rc3(uint32_t x) {          // reverse complement of 3 bases (6 bits)
    x = ~x & 63u;
    return ((x & 3u) << 4) | (x & 12u) | (x >> 4);
}

__global__ void __launch_bounds__(256) revcom_permute_kernel(const uint32_t* __restrict__ F, uint32_t* __restrict__ G, int k) {
    __shared__ uint32_t S[64][65];
    const int km = k - 6;                                       // bases in the middle
    const uint32_t mid = blockIdx.x;
    const uint32_t rcmid = km ? revcom32(mid, km) : 0u;
    const int top = 2 * (k - 3);
    const uint32_t t = threadIdx.x, y = t & 63u;
#pragma unroll 4
    for (uint32_t it = 0; it < 16; ++it) {
        const uint32_t x = it * 4 + (t >> 6);
        S[x][y] = __ldcs(F + (((size_t)x << top) | ((size_t)rcmid << 6) | y));
    }
    __syncthreads();
    const uint32_t ry = rc3(y);
#pragma unroll 4
    for (uint32_t it = 0; it < 16; ++it) {
        const uint32_t a = it * 4 + (t >> 6);
        G[((size_t)a << top) | ((size_t)mid << 6) | y] = S[ry][rc3(a)];
    }
}

// ---- merged-entry rule on the forward table F (SURVEY Q6 recipe; kmer_count.py:656-683) ------------------
// h survives iff F[h] > 0 and not (rc(h) present and h > rc(h)); value = min(h, rc h); count = F[h] + F[rc h]
// (a palindrome is its own partner, so its count doubles).  frc = F[rc h].  keep_higher = merge_revcom's
// keep_lower_hash_flag=False (kmer_count.py:671, 682): the comparisons turn around, value = max(h, rc h).
__device__ __forceinline__ bool merged_entry(uint32_t h, uint32_t fh, uint32_t frc, int k, bool keep_higher, uint32_t* value, uint32_t* count) {
    if (fh == 0) return false;
    const uint32_t rc = revcom32(h, k);
    if (frc > 0 && (keep_higher ? h < rc : h > rc)) return false;
    *value = (h < rc) != keep_higher ? h : rc;
    *count = fh + frc;
    return true;
}

// this thread's CP_ITEMS consecutive cells of F and their partners (from G when given, else gathered from F)
__device__ __forceinline__ void load_items(const uint32_t* __restrict__ F, const uint32_t* __restrict__ G, int64_t n_cells, int k,
                                           int revcom, int64_t base, uint32_t* fh, uint32_t* frc) {
    if (base + CP_ITEMS <= n_cells) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(F + base)), b = __ldg(reinterpret_cast<const uint4*>(F + base) + 1);
        fh[0] = a.x; fh[1] = a.y; fh[2] = a.z; fh[3] = a.w; fh[4] = b.x; fh[5] = b.y; fh[6] = b.z; fh[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < CP_ITEMS; ++j) fh[j] = base + j < n_cells ? __ldg(F + base + j) : 0u;
    }
    if (!revcom) {
#pragma unroll
        for (int j = 0; j < CP_ITEMS; ++j) frc[j] = 0;
    } else if (G && base + CP_ITEMS <= n_cells) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(G + base)), b = __ldcs(reinterpret_cast<const uint4*>(G + base) + 1);
        frc[0] = a.x; frc[1] = a.y; frc[2] = a.z; frc[3] = a.w; frc[4] = b.x; frc[5] = b.y; frc[6] = b.z; frc[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < CP_ITEMS; ++j)
            frc[j] = (base + j < n_cells && fh[j]) ? __ldg(F + revcom32((uint32_t)(base + j), k)) : 0u;
    }
}

__device__ __forceinline__ bool item_entry(uint32_t h, uint32_t fh, uint32_t frc, int k, int revcom, uint32_t* value, uint32_t* count) {
    if (!revcom) { *value = h; *count = fh; return fh != 0; }
    return merged_entry(h, fh, frc, k, revcom == 2, value, count);
}

__global__ void __launch_bounds__(CP_BLOCK) table_tile_count_kernel(const uint32_t* __restrict__ F, const uint32_t* __restrict__ G,
                                                                    int64_t n_cells, int k, int revcom,
                                                                    uint64_t* __restrict__ tile_counts, int64_t tile0) {
    const int64_t base = (tile0 + blockIdx.x) * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    uint32_t fh[CP_ITEMS], frc[CP_ITEMS];
    load_items(F, G, n_cells, k, revcom, base, fh, frc);
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) {
        uint32_t v, cnt;
        c += item_entry((uint32_t)(base + j), fh[j], frc[j], k, revcom, &v, &cnt);
    }
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(CP_BLOCK) table_tile_write_kernel(const uint32_t* __restrict__ F, const uint32_t* __restrict__ G,
                                                                    int64_t n_cells, int k, int revcom,
                                                                    const uint64_t* __restrict__ tile_offsets,
                                                                    uint32_t* __restrict__ kh_out, int32_t* __restrict__ cnt_out, int64_t tile0) {
    __shared__ uint32_t svals[CP_TILE], scnts[CP_TILE];           // the tile's entries in order: coalesced write-out
    const int64_t base = (tile0 + blockIdx.x) * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    uint32_t fh[CP_ITEMS], frc[CP_ITEMS];
    load_items(F, G, n_cells, k, revcom, base, fh, frc);
    uint32_t vals[CP_ITEMS], cnts[CP_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) {
        uint32_t v, cnt;
        if (item_entry((uint32_t)(base + j), fh[j], frc[j], k, revcom, &v, &cnt)) { vals[c] = v; cnts[c] = cnt; ++c; }
    }
    uint32_t total;
    const uint32_t off = block_exclusive_scan(c, &total);
#pragma unroll
    for (uint32_t j = 0; j < CP_ITEMS; ++j)
        if (j < c) { svals[off + j] = vals[j]; scnts[off + j] = cnts[j]; }
    __syncthreads();
    const uint64_t o = tile_offsets[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < total; i += CP_BLOCK) {
        __stcs(kh_out + o + i, svals[i]);
        __stcs(reinterpret_cast<uint32_t*>(cnt_out) + o + i, scnts[i]);
    }
}

// exclusive scan of n_tiles uint64 values in place, total appended at [n_tiles]; one block
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint64_t* __restrict__ v, int64_t n_tiles) {
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

// ---- Hamming-ball extraction from a merged list (motif_discovery.py:959-975) ------------------------------
__device__ __forceinline__ bool ball_member(uint32_t h, uint32_t conseq, uint32_t rc_conseq, int k, int d, int revcom,
                                            uint32_t low, uint32_t* value) {
    uint32_t dist = nz_groups32(h ^ conseq, low);
    bool flip = false;
    if (revcom) {
        const uint32_t rd = nz_groups32(h ^ rc_conseq, low);
        flip = rd < dist;                           // ties stay forward
        dist = rd < dist ? rd : dist;
    }
    if ((int)dist > d) return false;
    *value = flip ? revcom32(h & low, k) : h;
    return true;
}

__global__ void __launch_bounds__(CP_BLOCK) ball_tile_count_kernel(const uint32_t* __restrict__ kh, int64_t n, uint32_t conseq,
                                                                   uint32_t rc_conseq, int k, int d, int revcom,
                                                                   uint64_t* __restrict__ tile_counts) {
    const int64_t base = (int64_t)blockIdx.x * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    const uint32_t low = lowmask32(k);
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) {
        uint32_t v;
        if (base + j < n) c += ball_member(__ldg(kh + base + j), conseq, rc_conseq, k, d, revcom, low, &v);
    }
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// writes the members (if kh_out != NULL) and accumulates the 4 x k count matrix (cal_cnt_mat)
__global__ void __launch_bounds__(CP_BLOCK) ball_tile_write_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ cnt,
                                                                   int64_t n, uint32_t conseq, uint32_t rc_conseq, int k, int d,
                                                                   int revcom, const uint64_t* __restrict__ tile_offsets,
                                                                   uint32_t* __restrict__ kh_out, int32_t* __restrict__ cnt_out,
                                                                   unsigned long long* __restrict__ cnt_mat) {
    __shared__ unsigned long long smat[4 * 16];
    if (threadIdx.x < 64) smat[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    const uint32_t low = lowmask32(k);
    uint32_t vals[CP_ITEMS];
    int32_t cnts[CP_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) {
        uint32_t v;
        if (base + j < n && ball_member(__ldg(kh + base + j), conseq, rc_conseq, k, d, revcom, low, &v)) {
            vals[c] = v; cnts[c] = __ldg(cnt + base + j); ++c;
        }
    }
    uint32_t total;
    const uint32_t off = block_exclusive_scan(c, &total);
    const uint64_t o = tile_offsets[blockIdx.x] + off;
    for (uint32_t j = 0; j < c; ++j) {
        if (kh_out) { kh_out[o + j] = vals[j]; cnt_out[o + j] = cnts[j]; }
        for (int pos = 0; pos < k; ++pos) {
            const uint32_t b = (vals[j] >> (2 * (k - 1 - pos))) & 3u;
            atomicAdd(&smat[b * 16 + pos], (unsigned long long)(long long)cnts[j]);
        }
    }
    __syncthreads();
    if (total && threadIdx.x < 64) {
        const int b = threadIdx.x >> 4, pos = threadIdx.x & 15;
        if (pos < k && smat[threadIdx.x]) atomicAdd(&cnt_mat[b * k + pos], smat[threadIdx.x]);
    }
}

// ---- exclusive scan of uint32 counts into int64 offsets -----------------------------------------------------
__global__ void __launch_bounds__(CP_BLOCK) u32_tile_sum_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                uint64_t* __restrict__ tile_counts) {
    const int64_t base = (int64_t)blockIdx.x * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    uint32_t c = 0;   // a tile holds 2048 items; per-read hit counts are far below 2^32 / 2048
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) if (base + j < n) c += __ldg(in + base + j);
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}
__global__ void __launch_bounds__(CP_BLOCK) u32_tile_write_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                  const uint64_t* __restrict__ tile_offsets, int64_t n_tiles,
                                                                  int64_t* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    uint32_t x[CP_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) { x[j] = base + j < n ? __ldg(in + base + j) : 0; c += x[j]; }
    uint32_t total;
    const uint32_t off = block_exclusive_scan(c, &total);
    uint64_t run = tile_offsets[blockIdx.x] + off;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) if (base + j < n) { out[base + j] = (int64_t)run; run += x[j]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int64_t)tile_offsets[n_tiles];
}

int sync_read_total(const uint64_t* dev, int64_t* host, cudaStream_t s, const char* what) {
    uint64_t t = 0;
    cudaError_t e = cudaMemcpyAsync(&t, dev, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { kmap_set_error("%s: %s", what, cudaGetErrorString(e)); return (int)e; }
    *host = (int64_t)t;
    return KMAP_OK;
}

}  // namespace

extern "C" {

int64_t kmap_list_scratch_words(int64_t n) { return (n + CP_TILE - 1) / CP_TILE + 2; }
// tile counts, and for a table beyond L2 (k >= 13) the permuted copy G[h] = F[rc h]
int64_t kmap_compact_scratch_words(int k) {
    const int64_t tiles = kmap_list_scratch_words((int64_t)1 << (2 * k));
    return k >= 13 ? ((tiles + 31) & ~(int64_t)31) + ((int64_t)1 << (2 * k - 1)) : tiles;
}

int kmap_compact_merge_range(const uint32_t* table, int k, int revcom, int64_t cell_lo, int64_t cell_hi, uint64_t* scratch, uint32_t* kh_out,
                             int32_t* cnt_out, int64_t capacity, int64_t* n_out_host, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    KMAP_REQUIRE(table && scratch && n_out_host, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t n_cells = (int64_t)1 << (2 * k);
    KMAP_REQUIRE(cell_lo >= 0 && cell_lo <= cell_hi && cell_hi <= n_cells && cell_lo % CP_TILE == 0 && (cell_hi % CP_TILE == 0 || cell_hi == n_cells),
                 "the cell range must be aligned to 2048 cells");
    const int64_t tile0 = cell_lo / CP_TILE;
    const int64_t n_tiles = (cell_hi - cell_lo + CP_TILE - 1) / CP_TILE;
    *n_out_host = 0;
    if (n_tiles == 0) return KMAP_OK;
    uint32_t* G = nullptr;
    if (revcom && k >= 13) {
        G = reinterpret_cast<uint32_t*>(scratch + ((kmap_list_scratch_words(n_cells) + 31) & ~(int64_t)31));
        revcom_permute_kernel<<<1u << (2 * (k - 6)), 256, 0, s>>>(table, G, k);
    }
    table_tile_count_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(table, G, n_cells, k, revcom, scratch, tile0);
    scan_tiles_kernel<<<1, 1024, 0, s>>>(scratch, n_tiles);
    int rc = kmap_check_launch("compact_merge(count)");
    if (rc) return rc;
    rc = sync_read_total(scratch + n_tiles, n_out_host, s, "compact_merge");
    if (rc) return rc;
    if (capacity < *n_out_host || !kh_out || !cnt_out) {
        if (capacity == 0) return KMAP_OK;           // size query
        kmap_set_error("compact_merge: capacity %lld < %lld", (long long)capacity, (long long)*n_out_host);
        return KMAP_ERR_CAPACITY;
    }
    if (*n_out_host == 0) return KMAP_OK;
    table_tile_write_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(table, G, n_cells, k, revcom, scratch, kh_out, cnt_out, tile0);
    return kmap_check_launch("compact_merge(write)");
}

int kmap_compact_merge(const uint32_t* table, int k, int revcom, uint64_t* scratch, uint32_t* kh_out, int32_t* cnt_out,
                       int64_t capacity, int64_t* n_out_host, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 15, "dense tables support 1 <= k <= 15");
    return kmap_compact_merge_range(table, k, revcom, 0, (int64_t)1 << (2 * k), scratch, kh_out, cnt_out, capacity, n_out_host, stream);
}

int kmap_hamball_extract(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, uint32_t conseq, int d, int revcom,
                         uint64_t* scratch, uint32_t* kh_out, int32_t* cnt_out, int64_t capacity, int64_t* n_out_host,
                         int64_t* cnt_mat, void* stream) {
    KMAP_REQUIRE(k >= 1 && k <= 15 && n >= 0, "k out of range");
    KMAP_REQUIRE(scratch && n_out_host && cnt_mat, "null pointer");
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(cnt_mat, 0, (size_t)4 * k * 8, s);
    if (e != cudaSuccess) { kmap_set_error("hamball_extract: %s", cudaGetErrorString(e)); return (int)e; }
    *n_out_host = 0;
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && cnt, "null pointer");
    // host-side scalar reverse complement of the consensus (kmer_count.py:626-640)
    uint32_t rc = 0, com = (~conseq) & lowmask32(k);
    for (int i = 0; i < k; ++i) { rc = (rc << 2) | (com & 3u); com >>= 2; }
    const int64_t n_tiles = (n + CP_TILE - 1) / CP_TILE;
    ball_tile_count_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(kh, n, conseq, rc, k, d, revcom, scratch);
    scan_tiles_kernel<<<1, 1024, 0, s>>>(scratch, n_tiles);
    int r = kmap_check_launch("hamball_extract(count)");
    if (r) return r;
    r = sync_read_total(scratch + n_tiles, n_out_host, s, "hamball_extract");
    if (r) return r;
    const bool want_list = capacity > 0;
    if (want_list && (capacity < *n_out_host || !kh_out || !cnt_out)) {
        kmap_set_error("hamball_extract: capacity %lld < %lld", (long long)capacity, (long long)*n_out_host);
        return KMAP_ERR_CAPACITY;
    }
    ball_tile_write_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(kh, cnt, n, conseq, rc, k, d, revcom, scratch,
                                                                     want_list ? kh_out : nullptr, cnt_out,
                                                                     reinterpret_cast<unsigned long long*>(cnt_mat));
    return kmap_check_launch("hamball_extract(write)");
}

int kmap_exclusive_scan_u32(const uint32_t* in, int64_t n, int64_t* out, uint64_t* scratch, void* stream) {
    KMAP_REQUIRE(n >= 0 && out && scratch, "bad argument");
    cudaStream_t s = as_stream(stream);
    if (n == 0) { cudaMemsetAsync(out, 0, 8, s); return KMAP_OK; }
    const int64_t n_tiles = (n + CP_TILE - 1) / CP_TILE;
    u32_tile_sum_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(in, n, scratch);
    scan_tiles_kernel<<<1, 1024, 0, s>>>(scratch, n_tiles);
    u32_tile_write_kernel<<<(unsigned int)n_tiles, CP_BLOCK, 0, s>>>(in, n, scratch, n_tiles, out);
    return kmap_check_launch("exclusive_scan_u32");
}

}  // extern "C"
