// mask_input (kmer_count.py:580-610) and get_motif_occurence (motif_discovery.py:1422-1477) on the packed
// representation.  Both compare every window of the sequence with a consensus hash; neither materialises the
// hash or distance arrays of the reference.
#include "common.cuh"

namespace {

constexpr int MK_BLOCK = 256;
constexpr int MK_MAXM = 16;

// ---- mask: pass 1, flag word per 32 positions ----------------------------------------------------------------
// flag bit i <=> some consensus j has dist(window_i, cons[j]) <= d[j], where an invalid window (touches 255 or
// leaves the array) compares like the all-ones hash, i.e. like T..T (kmer_count.py:592-598).
//
// Bit-sliced: one thread = 32 window positions.  The 48 bases it can see are split into a plane of high bits and a
// plane of low bits (bit p = base at position p); for consensus base i the positions whose base at offset i differs are
// ((HI >> i) ^ hi_i) | ((LO >> i) ^ lo_i) with hi_i / lo_i all-ones or zero, so every logic instruction works on 32
// windows at once.  The k mismatch planes are summed by a carry-save adder tree into a 5-bit bit-sliced count and
// compared with d bit-serially: ~100 logic instructions per consensus and word, no POPC, instead of ~8 per window.
__device__ __forceinline__ uint32_t even_bits32(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a | b));
}
__device__ __forceinline__ void half_add(uint32_t a, uint32_t b, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b;
    carry = a & b;
}

__global__ void __launch_bounds__(MK_BLOCK) mask_flag_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                             int64_t n, int64_t n_words, int k, const uint32_t* __restrict__ cons,
                                                             const int32_t* __restrict__ dmax, int m, uint32_t* __restrict__ flags) {
    __shared__ uint2 smask[MK_MAXM][16];       // (.x, .y) = (hi_i, lo_i) of consensus j, all-ones or zero
    __shared__ int sd[MK_MAXM];
    __shared__ int inv_hit;                    // does an invalid window fall inside some ball?
    const uint32_t low = lowmask32(k);
    if (threadIdx.x < MK_MAXM * 16) {
        const int j = threadIdx.x >> 4, i = threadIdx.x & 15;
        uint2 v = make_uint2(0, 0);
        if (j < m && i < k) {
            const uint32_t base = ((cons[j] & low) >> (2 * (k - 1 - i))) & 3u;     // base i of the consensus, first base first
            v.x = 0u - (base >> 1);
            v.y = 0u - (base & 1u);
        }
        smask[j][i] = v;
    }
    if (threadIdx.x == 0) {
        int hit = 0;
        for (int j = 0; j < m; ++j) {
            sd[j] = dmax[j];
            hit |= (int)nz_groups32(0xFFFFFFFFu ^ (cons[j] & low), low) <= dmax[j];
        }
        inv_hit = hit;
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * MK_BLOCK + threadIdx.x;
    if (t >= n_words) return;
    const uint32_t v0 = __ldg(valid + t), v1 = __ldg(valid + t + 1);
    const uint2 w01 = __ldg(reinterpret_cast<const uint2*>(packed + 2 * t));
    const uint32_t w2 = __ldg(packed + 2 * t + 2);
    // planes: after a bit reversal base p' of a word sits at bits (2p', 2p'+1) = (high bit, low bit)
    const uint32_t r0 = __brev(w01.x), r1 = __brev(w01.y), r2 = __brev(w2);
    const uint32_t hi_a = even_bits32(r0) | (even_bits32(r1) << 16), hi_b = even_bits32(r2);        // positions 0..31, 32..47
    const uint32_t lo_a = even_bits32(r0 >> 1) | (even_bits32(r1 >> 1) << 16), lo_b = even_bits32(r2 >> 1);
    // windows of k valid bases starting at bits 0..31 (log-step run-length test, k <= 16)
    uint32_t wm;
    {
        const uint64_t V = ((uint64_t)v1 << 32) | v0;
        const uint64_t q2 = V & (V >> 1), q4 = q2 & (q2 >> 2), q8 = q4 & (q4 >> 4);
        uint64_t mm = ~0ull;
        int off = 0;
        if (k & 16) { mm &= q8 & (q8 >> 8); off += 16; }
        if (k & 8) { mm &= q8 >> off; off += 8; }
        if (k & 4) { mm &= q4 >> off; off += 4; }
        if (k & 2) { mm &= q2 >> off; off += 2; }
        if (k & 1) { mm &= V >> off; }
        wm = (uint32_t)mm;
    }
    uint32_t hit = 0;
    if (wm) {
        for (int j = 0; j < m; ++j) {
            uint32_t x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint2 c = smask[j][i];                                   // (zero planes for i >= k: see below)
                x[i] = (__funnelshift_r(hi_a, hi_b, i) ^ c.x) | (__funnelshift_r(lo_a, lo_b, i) ^ c.y);
            }
            // offsets i >= k do not belong to the window
#pragma unroll
            for (int i = 0; i < 16; ++i) if (i >= k) x[i] = 0;
            // carry-save adder tree: 16 planes of weight 1 -> 5-bit count (b4 b3 b2 b1 b0)
            uint32_t s1[6], c2[8];
#pragma unroll
            for (int g = 0; g < 5; ++g) full_add(x[3 * g], x[3 * g + 1], x[3 * g + 2], s1[g], c2[g]);
            s1[5] = x[15];
            uint32_t t0, t1, b0;
            full_add(s1[0], s1[1], s1[2], t0, c2[5]);
            full_add(s1[3], s1[4], s1[5], t1, c2[6]);
            half_add(t0, t1, b0, c2[7]);
            uint32_t u0, u1, c4[4];
            full_add(c2[0], c2[1], c2[2], u0, c4[0]);
            full_add(c2[3], c2[4], c2[5], u1, c4[1]);
            uint32_t u2, b1;
            full_add(u0, u1, c2[6], u2, c4[2]);
            half_add(u2, c2[7], b1, c4[3]);
            uint32_t y0, c8a, c8b, b2, b3, b4;
            full_add(c4[0], c4[1], c4[2], y0, c8a);
            half_add(y0, c4[3], b2, c8b);
            half_add(c8a, c8b, b3, b4);
            // count <= d, bit-serial from the top
            const int d = sd[j];
            uint32_t le;
            if (d >= 16) {
                le = ~0u;
            } else if (d < 0) {
                le = 0u;
            } else {
                const uint32_t bits[4] = {b0, b1, b2, b3};
                uint32_t lt = 0, eq = ~b4;                                      // count < 16 required (d < 16)
#pragma unroll
                for (int q = 3; q >= 0; --q) {
                    if ((d >> q) & 1) { lt |= eq & ~bits[q]; eq &= bits[q]; }
                    else eq &= ~bits[q];
                }
                le = lt | eq;
            }
            hit |= le;
        }
    }
    uint32_t f = (hit & wm) | (inv_hit ? ~wm : 0u);
    // positions >= n do not exist
    const int64_t p0 = t * 32;
    if (p0 + 32 > n) f &= (p0 >= n) ? 0u : ((1u << (n - p0)) - 1u);
    flags[t] = f;
}

// ---- mask: pass 2, dilate the flags by k to the right and clear those validity bits ---------------------------
__global__ void __launch_bounds__(MK_BLOCK) mask_dilate_kernel(const uint32_t* __restrict__ flags, int64_t n_words, int k,
                                                               uint32_t* __restrict__ valid) {
    const int64_t t = (int64_t)blockIdx.x * MK_BLOCK + threadIdx.x;
    if (t >= n_words) return;
    const uint64_t cur = flags[t];
    const uint64_t prev = t > 0 ? flags[t - 1] : 0;
    uint64_t x = (cur << 32) | prev;           // bit 32+i = position 32t+i, bit i = position 32(t-1)+i
    // OR of x << s for s = 0 .. k-1  (k <= 32), by doubling
    int covered = 1;
    while (covered * 2 <= k) { x |= x << covered; covered *= 2; }
    x |= x << (k - covered);
    const uint32_t kill = (uint32_t)(x >> 32);
    if (kill) valid[t] &= ~kill;
}

// ---- per-read occurrence scan -----------------------------------------------------------------------------------
// One warp per read.  Window positions 0 .. n_pos-1 where n_pos = L-k+1, or -- reproducing the reference's negative
// slice hash_arr[0:L-k+1] for reads shorter than k (motif_discovery.py:1447) -- max(0, 2L-k+1) all-invalid windows.
__device__ __forceinline__ int64_t occurrence_n_pos(int64_t L, int k) {
    const int64_t m = L - k + 1;
    if (m > 0) return m;
    if (m == 0) return 0;
    return L + m > 0 ? L + m : 0;
}

// hash-width helpers: uint32 for k <= 16 (one 16-base fetch), uint64 for 17 <= k <= 31
template <typename H> struct HashOps;
template <> struct HashOps<uint32_t> {
    static __device__ __forceinline__ uint32_t low(int k) { return lowmask32(k); }
    static __device__ __forceinline__ uint32_t key(const uint32_t* __restrict__ packed, int64_t p, int k) { return window16(packed, p) >> (32 - 2 * k); }
    static __device__ __forceinline__ uint32_t dist(uint32_t x, uint32_t low) { return nz_groups32(x, low); }
    static __device__ __forceinline__ uint32_t rc(uint32_t h, int k) { return revcom32(h, k); }
};
template <> struct HashOps<uint64_t> {
    static __device__ __forceinline__ uint64_t low(int k) { return lowmask64(k); }
    static __device__ __forceinline__ uint64_t key(const uint32_t* __restrict__ packed, int64_t p, int k) { return window32(packed, p) >> (64 - 2 * k); }
    static __device__ __forceinline__ uint32_t dist(uint64_t x, uint64_t low) { return nz_groups64(x, low); }
    static __device__ __forceinline__ uint64_t rc(uint64_t h, int k) { return revcom64(h, k); }
};

template <bool FILL, typename H>
__global__ void __launch_bounds__(MK_BLOCK) occurrence_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                              const int64_t* __restrict__ borders, int64_t n_seq, int k, H conseq,
                                                              int d, int revcom, uint8_t* __restrict__ min_dist, uint32_t* __restrict__ n_hit,
                                                              const int64_t* __restrict__ offsets, int32_t* __restrict__ pos_out) {
    typedef HashOps<H> Ops;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * MK_BLOCK + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * MK_BLOCK) >> 5;
    const H low = Ops::low(k);
    const H c = conseq & low, rc = Ops::rc(c, k);
    const uint32_t km = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    uint32_t inv_d = Ops::dist((H)~(H)0 ^ c, low);
    if (revcom) { const uint32_t r = Ops::dist((H)~(H)0 ^ rc, low); inv_d = r < inv_d ? r : inv_d; }

    for (int64_t r = warp0; r < n_seq; r += n_warps) {
        const int64_t st = __ldg(borders + 2 * r), en = __ldg(borders + 2 * r + 1);
        const int64_t L = en - st;
        const int64_t n_pos = occurrence_n_pos(L, k);
        uint32_t best = 255, hits = 0;
        if (FILL) { best = min_dist[r]; if (best == 255) continue; }
        int64_t out = FILL ? offsets[r] : 0;
        for (int64_t i0 = 0; i0 < n_pos; i0 += 32) {
            const int64_t i = i0 + lane;
            uint32_t dist = 255;
            if (i < n_pos) {
                const int64_t p = st + i;
                const bool ok = (L >= k) && ((valid32(valid, p) & km) == km);
                if (ok) {
                    const H h = Ops::key(packed, p, k);
                    dist = Ops::dist(h ^ c, low);
                    if (revcom) { const uint32_t rd = Ops::dist(h ^ rc, low); dist = rd < dist ? rd : dist; }
                } else {
                    dist = inv_d;
                }
                if ((int)dist > d) dist = 255;
            }
            if (!FILL) {
                uint32_t wmin = dist;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, wmin, o); wmin = y < wmin ? y : wmin; }
                if (wmin < best) { best = wmin; hits = 0; }
                if (best != 255) hits += __popc(__ballot_sync(0xFFFFFFFFu, dist == best));
            } else {
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, dist == best);
                if (dist == best) pos_out[out + __popc(bal & ((1u << lane) - 1u))] = (int32_t)i;
                out += __popc(bal);
            }
        }
        if (!FILL && lane == 0) { min_dist[r] = (uint8_t)best; n_hit[r] = hits; }
    }
}

// ---- mask for 17 <= k <= 31 (64-bit hashes): one thread = one position, one flag word per warp ----------------------------------
// Same rule as mask_flag_kernel: flag <=> some consensus within d of the PRE-mask window, an invalid window comparing like
// the all-ones hash (T..T).  These k are beyond the stock motif_def_table's accepted range, so this is the plain version.
__global__ void __launch_bounds__(MK_BLOCK) mask_flag_wide_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid,
                                                                  int64_t n, int k, const uint64_t* __restrict__ cons,
                                                                  const int32_t* __restrict__ dmax, int m, uint32_t* __restrict__ flags) {
    __shared__ uint64_t sc[MK_MAXM];
    __shared__ int sd[MK_MAXM];
    const uint64_t low = lowmask64(k);
    if (threadIdx.x < m) { sc[threadIdx.x] = cons[threadIdx.x] & low; sd[threadIdx.x] = dmax[threadIdx.x]; }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * MK_BLOCK + threadIdx.x;            // (whole warps run past n: the ballot needs them)
    bool hit = false;
    if (p < n) {
        const uint64_t h = window_ok(valid, p, k) ? (window32(packed, p) >> (64 - 2 * k)) : low;
        for (int j = 0; j < m; ++j) hit = hit || (int)nz_groups64(h ^ sc[j], low) <= sd[j];
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31) == 0 && (p >> 5) < (n + 31) / 32) flags[p >> 5] = word;
}

}  // namespace

extern "C" {

int kmap_mask(const uint32_t* packed, const uint32_t* valid_pre, uint32_t* valid, int64_t n, int k, const uint32_t* cons,
              const int32_t* d, int m, uint32_t* flag_scratch, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 16 && m >= 0 && m <= MK_MAXM, "bad argument (k <= 16, m <= 16)");
    if (n == 0 || m == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid_pre && valid && cons && d && flag_scratch, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t n_words = (n + 31) / 32;
    mask_flag_kernel<<<grid_for(n_words, MK_BLOCK), MK_BLOCK, 0, s>>>(packed, valid_pre, n, n_words, k, cons, d, m, flag_scratch);
    mask_dilate_kernel<<<grid_for(n_words, MK_BLOCK), MK_BLOCK, 0, s>>>(flag_scratch, n_words, k, valid);
    return kmap_check_launch("mask");
}

static unsigned int occurrence_grid(int64_t n_seq) {
    int64_t blocks = (n_seq * 32 + MK_BLOCK - 1) / MK_BLOCK;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return (unsigned int)(blocks < 1 ? 1 : blocks);
}

int kmap_occurrence_count(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                          uint32_t conseq, int d, int revcom, uint8_t* min_dist, uint32_t* n_hit, void* stream) {
    KMAP_REQUIRE(n_seq >= 0 && k >= 1 && k <= 16, "bad argument (k <= 16)");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && borders && min_dist && n_hit, "null pointer");
    occurrence_kernel<false, uint32_t><<<occurrence_grid(n_seq), MK_BLOCK, 0, as_stream(stream)>>>(
        packed, valid, borders, n_seq, k, conseq, d, revcom, min_dist, n_hit, nullptr, nullptr);
    return kmap_check_launch("occurrence_count");
}

int kmap_occurrence_fill(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                         uint32_t conseq, int d, int revcom, const uint8_t* min_dist, const int64_t* offsets, int32_t* pos_out,
                         void* stream) {
    KMAP_REQUIRE(n_seq >= 0 && k >= 1 && k <= 16, "bad argument (k <= 16)");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && borders && min_dist && offsets && pos_out, "null pointer");
    occurrence_kernel<true, uint32_t><<<occurrence_grid(n_seq), MK_BLOCK, 0, as_stream(stream)>>>(
        packed, valid, borders, n_seq, k, conseq, d, revcom, const_cast<uint8_t*>(min_dist), nullptr, offsets, pos_out);
    return kmap_check_launch("occurrence_fill");
}

// ---- 64-bit hashes (17 <= k <= 31; any 1 <= k <= 31 is accepted) -------------------------------------------------------------------
int kmap_mask_u64(const uint32_t* packed, const uint32_t* valid_pre, uint32_t* valid, int64_t n, int k, const uint64_t* cons,
                  const int32_t* d, int m, uint32_t* flag_scratch, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 31 && m >= 0 && m <= MK_MAXM, "bad argument (k <= 31, m <= 16)");
    if (n == 0 || m == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid_pre && valid && cons && d && flag_scratch, "null pointer");
    cudaStream_t s = as_stream(stream);
    const int64_t n_words = (n + 31) / 32;
    mask_flag_wide_kernel<<<grid_for(n_words * 32, MK_BLOCK), MK_BLOCK, 0, s>>>(packed, valid_pre, n, k, cons, d, m, flag_scratch);
    mask_dilate_kernel<<<grid_for(n_words, MK_BLOCK), MK_BLOCK, 0, s>>>(flag_scratch, n_words, k, valid);
    return kmap_check_launch("mask_u64");
}

int kmap_occurrence_count_u64(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                              uint64_t conseq, int d, int revcom, uint8_t* min_dist, uint32_t* n_hit, void* stream) {
    KMAP_REQUIRE(n_seq >= 0 && k >= 1 && k <= 31, "bad argument (k <= 31)");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && borders && min_dist && n_hit, "null pointer");
    occurrence_kernel<false, uint64_t><<<occurrence_grid(n_seq), MK_BLOCK, 0, as_stream(stream)>>>(
        packed, valid, borders, n_seq, k, conseq, d, revcom, min_dist, n_hit, nullptr, nullptr);
    return kmap_check_launch("occurrence_count_u64");
}

int kmap_occurrence_fill_u64(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                             uint64_t conseq, int d, int revcom, const uint8_t* min_dist, const int64_t* offsets, int32_t* pos_out,
                             void* stream) {
    KMAP_REQUIRE(n_seq >= 0 && k >= 1 && k <= 31, "bad argument (k <= 31)");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(packed && valid && borders && min_dist && offsets && pos_out, "null pointer");
    occurrence_kernel<true, uint64_t><<<occurrence_grid(n_seq), MK_BLOCK, 0, as_stream(stream)>>>(
        packed, valid, borders, n_seq, k, conseq, d, revcom, const_cast<uint8_t*>(min_dist), nullptr, offsets, pos_out);
    return kmap_check_launch("occurrence_fill_u64");
}

}  // extern "C"
