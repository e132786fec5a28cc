// 1:1 replacements of the reference's integer Taichi kernels (taichi_core.py:3-224) on plain arrays.
// These keep the reference's array-in / array-out shape (one hash per position, one distance per hash); the
// fused path in count.cu / mask.cu never materialises those arrays.
#include "common.cuh"

namespace {

constexpr int HASH_BLOCK = 256;
constexpr int HASH_PER_THREAD = 4;
constexpr int HASH_TILE = HASH_BLOCK * HASH_PER_THREAD;

// One hash per position (taichi_core.py:3-31 / 34-61).  A block stages TILE + k - 1 bytes in shared memory
// (bytes past the end read as 255, which also makes windows that leave the array invalid), each thread rolls
// 4 consecutive hashes and stores them as one vector.
template <typename H>
__global__ void __launch_bounds__(HASH_BLOCK) kmer2hash_kernel(const uint8_t* __restrict__ seq, int64_t n, int k,
                                                               H* __restrict__ out) {
    __shared__ uint8_t tile[HASH_TILE + 32];
    const int64_t base = (int64_t)blockIdx.x * HASH_TILE;
    const int span = HASH_TILE + k;
    for (int i = threadIdx.x; i < span; i += HASH_BLOCK) {
        const int64_t p = base + i;
        tile[i] = p < n ? __ldg(seq + p) : (uint8_t)255;
    }
    __syncthreads();
    const int t0 = threadIdx.x * HASH_PER_THREAD;
    const H mask = (sizeof(H) == 4) ? (H)lowmask32(k) : (H)lowmask64(k);
    H h = 0;
    int bad = 0;  // number of missing bases inside the current window
    for (int j = 0; j < k; ++j) {
        const uint8_t b = tile[t0 + j];
        bad += (b == 255);
        h = (H)((h << 2) + b);
    }
    H res[HASH_PER_THREAD];
#pragma unroll
    for (int i = 0; i < HASH_PER_THREAD; ++i) {
        res[i] = bad ? (H)~(H)0 : (H)(h & mask);
        const uint8_t gone = tile[t0 + i], in = tile[t0 + i + k];
        bad += (in == 255) - (gone == 255);
        h = (H)((h << 2) + in);
    }
    const int64_t p0 = base + t0;
#pragma unroll
    for (int i = 0; i < HASH_PER_THREAD; ++i)
        if (p0 + i < n) out[p0 + i] = res[i];
}

enum DistMode { DIST_FULL = 0, DIST_HEAD = 1, DIST_TAIL = 2 };

// taichi_core.py:63-177: distance of every hash to one target over `groups` 2-bit groups, after an optional
// right shift (head variant).  Memory-bound: 4/8 B in, 1 B out per element.
template <typename H>
__global__ void ham_dist_kernel(const H* __restrict__ kh, int64_t n, H target, int shift, int groups,
                                uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    H x = (H)(kh[i] >> shift) ^ target;
    uint32_t d;
    if (sizeof(H) == 4) d = nz_groups32((uint32_t)x, lowmask32(groups));
    else d = nz_groups64((uint64_t)x, lowmask64(groups));
    out[i] = (uint8_t)d;
}

template <typename H>
__global__ void revcom_kernel(const H* __restrict__ in, int64_t n, int k, H* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sizeof(H) == 4) out[i] = (H)revcom32((uint32_t)(in[i] & lowmask32(k)), k);
    else out[i] = (H)revcom64((uint64_t)in[i] & lowmask64(k), k);
}

// sample_disp_kmer's labelling (motif_discovery.py:849-892) fused: for every unique k-mer the nearest consensus by
// head distance (first len(conseq) bases vs the consensus) or, in revcom mode, tail distance (last len(conseq) bases vs
// its reverse complement); a consensus farther than its own max distance counts as distance k; label = first nearest
// consensus, or n_conseq when even the nearest is beyond dmax_k; a k-mer labelled i that is strictly closer to the reverse
// complement of consensus i is reverse-complemented in place.  One thread = one k-mer; the reference scans the list
// 2 x n_conseq times and builds two n_conseq x n matrices on the host.
template <typename H>
__global__ void __launch_bounds__(256) label_kmers_kernel(H* __restrict__ kh, int64_t n, int k, const H* __restrict__ conseq,
                                                          const H* __restrict__ rc_conseq, const int32_t* __restrict__ conseq_len,
                                                          const int32_t* __restrict__ dmax, int n_conseq, int dmax_k, int revcom,
                                                          int32_t* __restrict__ label) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const H h = kh[i];
    int best = 0x7FFFFFFF, best_i = 0;
    bool best_flip = false;
    for (int c = 0; c < n_conseq; ++c) {
        const int len = __ldg(conseq_len + c);
        int d, rd = 0x7FFFFFFF;
        if (sizeof(H) == 4) {
            const uint32_t low = lowmask32(len);
            d = (int)nz_groups32((uint32_t)(h >> (2 * (k - len))) ^ (uint32_t)__ldg(conseq + c), low);
            if (revcom) rd = (int)nz_groups32((uint32_t)h ^ (uint32_t)__ldg(rc_conseq + c), low);
        } else {
            const uint64_t low = lowmask64(len);
            d = (int)nz_groups64((uint64_t)(h >> (2 * (k - len))) ^ (uint64_t)__ldg(conseq + c), low);
            if (revcom) rd = (int)nz_groups64((uint64_t)h ^ (uint64_t)__ldg(rc_conseq + c), low);
        }
        const bool flip = rd < d;
        int dist = flip ? rd : d;
        if (dist > __ldg(dmax + c)) dist = k;
        if (dist < best) { best = dist; best_i = c; best_flip = flip; }        // strict: np.argmin keeps the first minimum
    }
    if (best > dmax_k) { best_i = n_conseq; best_flip = false; }
    label[i] = best_i;
    if (best_flip) {
        if (sizeof(H) == 4) kh[i] = (H)revcom32((uint32_t)(h & lowmask32(k)), k);
        else kh[i] = (H)revcom64((uint64_t)h & lowmask64(k), k);
    }
}

// Top-kk candidates of a count array for find_motif's `np.argpartition(cnt, -top_k)[-top_k:]` (motif_discovery.py:657):
// every block emits its kk largest (value, index) pairs, ordered by (value descending, index ascending); the host merges
// blocks x kk candidates.  The caller asks for top_k + 1 and uses the result only when the boundary is strict (see
// motif_discovery.find_motif_on_device): which of several equal values numpy's introselect keeps is not reproducible.
constexpr int TOPK_MAX = 8;
constexpr int TOPK_BLOCK = 256;
template <typename T>
__global__ void __launch_bounds__(TOPK_BLOCK) topk_candidates_kernel(const T* __restrict__ cnt, int64_t n, int kk, T* __restrict__ out_val,
                                                                     long long* __restrict__ out_idx) {
    __shared__ T s_val[TOPK_BLOCK / 32];
    __shared__ long long s_idx[TOPK_BLOCK / 32];
    __shared__ int s_owner[TOPK_BLOCK / 32];
    __shared__ int winner;
    T val[TOPK_MAX];
    long long idx[TOPK_MAX];
#pragma unroll
    for (int j = 0; j < TOPK_MAX; ++j) { val[j] = 0; idx[j] = -1; }          // idx < 0: empty slot
    auto better = [](T av, long long ai, T bv, long long bi) {              // a before b?  (empty slots last)
        if (bi < 0) return ai >= 0;
        if (ai < 0) return false;
        return av > bv || (av == bv && ai < bi);
    };
    for (int64_t i = (int64_t)blockIdx.x * TOPK_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * TOPK_BLOCK) {
        const T v = __ldg(cnt + i);
        if (!better(v, i, val[TOPK_MAX - 1], idx[TOPK_MAX - 1])) continue;
        T cv = v;
        long long ci = i;
#pragma unroll
        for (int j = 0; j < TOPK_MAX; ++j) {                                 // insertion into the sorted list
            if (better(cv, ci, val[j], idx[j])) { const T tv = val[j]; const long long ti = idx[j]; val[j] = cv; idx[j] = ci; cv = tv; ci = ti; }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int head = 0;                                                            // next unused entry of this thread's list
    for (int r = 0; r < kk; ++r) {
        T hv = 0;
        long long hi = -1;
#pragma unroll
        for (int j = 0; j < TOPK_MAX; ++j) if (j == head) { hv = val[j]; hi = idx[j]; }
        T bv = hv;
        long long bi = hi;
        int owner = threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
            const long long oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            const int oo = __shfl_xor_sync(0xFFFFFFFFu, owner, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; owner = oo; }
        }
        if (lane == 0) { s_val[w] = bv; s_idx[w] = bi; s_owner[w] = owner; }
        __syncthreads();
        if (threadIdx.x == 0) {
            T gv = s_val[0];
            long long gi = s_idx[0];
            int go = s_owner[0];
            for (int q = 1; q < TOPK_BLOCK / 32; ++q)
                if (better(s_val[q], s_idx[q], gv, gi)) { gv = s_val[q]; gi = s_idx[q]; go = s_owner[q]; }
            out_val[(size_t)blockIdx.x * kk + r] = gv;
            out_idx[(size_t)blockIdx.x * kk + r] = gi;
            winner = gi >= 0 ? go : -1;
        }
        __syncthreads();
        if ((int)threadIdx.x == winner) ++head;
        __syncthreads();
    }
}

// remove_duplicate_hash_per_seq (kmer_count.py:743-760): one block per read; a shared-memory map
// hash -> smallest position (open addressing), filled in passes when the read has more distinct hashes than the
// map holds (pass p owns the hashes with mix(h) % n_pass == p).  Any position that is not the first occurrence
// of its hash is overwritten with the invalid hash.
constexpr int DEDUP_SLOTS = 4096;
__global__ void __launch_bounds__(256) dedup_hash_per_read_kernel(uint32_t* __restrict__ hash, int64_t n,
                                                                  const int64_t* __restrict__ borders, int64_t n_seq) {
    __shared__ uint32_t keys[DEDUP_SLOTS];
    __shared__ uint32_t firsts[DEDUP_SLOTS];
    for (int64_t r = blockIdx.x; r < n_seq; r += gridDim.x) {
        int64_t st = borders[2 * r], en = borders[2 * r + 1];
        if (st < 0) st = 0;
        if (en > n) en = n;
        const int64_t len = en - st;
        if (len <= 1) continue;
        const uint32_t n_pass = (uint32_t)((len + DEDUP_SLOTS / 2 - 1) / (DEDUP_SLOTS / 2));
        for (uint32_t pass = 0; pass < n_pass; ++pass) {
            for (int i = threadIdx.x; i < DEDUP_SLOTS; i += blockDim.x) { keys[i] = KMAP_EMPTY_SLOT; firsts[i] = 0xFFFFFFFFu; }
            __syncthreads();
            // phase 1: claim a slot per distinct hash, keep the minimum offset
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                const uint32_t h = hash[st + i];
                if (h == 0xFFFFFFFFu) continue;
                const uint32_t m = mix32(h);
                if (n_pass > 1 && (m >> 12) % n_pass != pass) continue;
                uint32_t slot = m & (DEDUP_SLOTS - 1);
                while (true) {
                    const uint32_t old = atomicCAS(&keys[slot], KMAP_EMPTY_SLOT, h);
                    if (old == KMAP_EMPTY_SLOT || old == h) { atomicMin(&firsts[slot], (uint32_t)i); break; }
                    slot = (slot + 1) & (DEDUP_SLOTS - 1);
                }
            }
            __syncthreads();
            // phase 2: everything but the first occurrence becomes invalid
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                const uint32_t h = hash[st + i];
                if (h == 0xFFFFFFFFu) continue;
                const uint32_t m = mix32(h);
                if (n_pass > 1 && (m >> 12) % n_pass != pass) continue;
                uint32_t slot = m & (DEDUP_SLOTS - 1);
                while (keys[slot] != h) slot = (slot + 1) & (DEDUP_SLOTS - 1);
                if (firsts[slot] != (uint32_t)i) hash[st + i] = 0xFFFFFFFFu;
            }
            __syncthreads();
        }
    }
}

template <typename H>
int launch_label(H* kh, int64_t n, int k, const H* conseq, const H* rc_conseq, const int32_t* conseq_len, const int32_t* dmax,
                 int n_conseq, int dmax_k, int revcom, int32_t* label, void* stream, int kmax) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= kmax && n_conseq >= 0, "bad argument");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && label && (n_conseq == 0 || (conseq && rc_conseq && conseq_len && dmax)), "null pointer");
    label_kmers_kernel<H><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(kh, n, k, conseq, rc_conseq, conseq_len, dmax, n_conseq, dmax_k,
                                                                          revcom, label);
    return kmap_check_launch("label_kmers");
}

template <typename T>
int launch_topk(const T* cnt, int64_t n, int kk, T* out_val, int64_t* out_idx, int n_blocks, void* stream) {
    KMAP_REQUIRE(n >= 0 && kk >= 1 && kk <= TOPK_MAX && n_blocks >= 1, "bad argument (kk <= 8)");
    KMAP_REQUIRE(out_val && out_idx && (cnt || n == 0), "null pointer");
    topk_candidates_kernel<T><<<(unsigned int)n_blocks, TOPK_BLOCK, 0, as_stream(stream)>>>(cnt, n, kk, out_val, reinterpret_cast<long long*>(out_idx));
    return kmap_check_launch("topk_candidates");
}

template <typename H>
int launch_hash(const uint8_t* seq, int64_t n, int k, H* out, void* stream, int kmax) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= kmax, "k out of range for this hash width");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(seq && out, "null pointer");
    kmer2hash_kernel<H><<<grid_for(n, HASH_TILE), HASH_BLOCK, 0, as_stream(stream)>>>(seq, n, k, out);
    return kmap_check_launch("kmer2hash");
}

template <typename H>
int launch_dist(const H* kh, int64_t n, H target, int k, int conseq_len, int mode, uint8_t* out, void* stream, int kmax) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= kmax, "k out of range for this hash width");
    KMAP_REQUIRE(conseq_len >= 0 && conseq_len <= k, "consensus longer than k");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(kh && out, "null pointer");
    const int shift = mode == DIST_HEAD ? 2 * (k - conseq_len) : 0;
    const int groups = mode == DIST_FULL ? k : conseq_len;
    ham_dist_kernel<H><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(kh, n, target, shift, groups, out);
    return kmap_check_launch("ham_dist");
}

}  // namespace

extern "C" {

int kmap_kmer2hash_u32(const uint8_t* seq, int64_t n, int k, uint32_t* out, void* stream) {
    return launch_hash<uint32_t>(seq, n, k, out, stream, 15);
}
int kmap_kmer2hash_u64(const uint8_t* seq, int64_t n, int k, uint64_t* out, void* stream) {
    return launch_hash<uint64_t>(seq, n, k, out, stream, 31);
}
int kmap_ham_dist_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, uint8_t* out, void* stream) {
    return launch_dist<uint32_t>(kh, n, target, k, k, DIST_FULL, out, stream, 16);
}
int kmap_ham_dist_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, uint8_t* out, void* stream) {
    return launch_dist<uint64_t>(kh, n, target, k, k, DIST_FULL, out, stream, 32);
}
int kmap_ham_dist_head_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, int c, uint8_t* out, void* stream) {
    return launch_dist<uint32_t>(kh, n, target, k, c, DIST_HEAD, out, stream, 16);
}
int kmap_ham_dist_head_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, int c, uint8_t* out, void* stream) {
    return launch_dist<uint64_t>(kh, n, target, k, c, DIST_HEAD, out, stream, 32);
}
int kmap_ham_dist_tail_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, int c, uint8_t* out, void* stream) {
    return launch_dist<uint32_t>(kh, n, target, k, c, DIST_TAIL, out, stream, 16);
}
int kmap_ham_dist_tail_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, int c, uint8_t* out, void* stream) {
    return launch_dist<uint64_t>(kh, n, target, k, c, DIST_TAIL, out, stream, 32);
}
int kmap_revcom_u32(const uint32_t* in, int64_t n, int k, uint32_t* out, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 16, "k out of range");
    if (n == 0) return KMAP_OK;
    revcom_kernel<uint32_t><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(in, n, k, out);
    return kmap_check_launch("revcom");
}
int kmap_revcom_u64(const uint64_t* in, int64_t n, int k, uint64_t* out, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 32, "k out of range");
    if (n == 0) return KMAP_OK;
    revcom_kernel<uint64_t><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(in, n, k, out);
    return kmap_check_launch("revcom");
}
int kmap_label_kmers_u32(uint32_t* kh, int64_t n, int k, const uint32_t* conseq, const uint32_t* rc_conseq, const int32_t* conseq_len,
                         const int32_t* dmax, int n_conseq, int dmax_k, int revcom, int32_t* label, void* stream) {
    return launch_label<uint32_t>(kh, n, k, conseq, rc_conseq, conseq_len, dmax, n_conseq, dmax_k, revcom, label, stream, 15);
}
int kmap_label_kmers_u64(uint64_t* kh, int64_t n, int k, const uint64_t* conseq, const uint64_t* rc_conseq, const int32_t* conseq_len,
                         const int32_t* dmax, int n_conseq, int dmax_k, int revcom, int32_t* label, void* stream) {
    return launch_label<uint64_t>(kh, n, k, conseq, rc_conseq, conseq_len, dmax, n_conseq, dmax_k, revcom, label, stream, 31);
}
int kmap_topk_candidates_i32(const int32_t* cnt, int64_t n, int kk, int32_t* out_val, int64_t* out_idx, int n_blocks, void* stream) {
    return launch_topk<int32_t>(cnt, n, kk, out_val, out_idx, n_blocks, stream);
}
int kmap_topk_candidates_i64(const int64_t* cnt, int64_t n, int kk, int64_t* out_val, int64_t* out_idx, int n_blocks, void* stream) {
    return launch_topk<long long>(reinterpret_cast<const long long*>(cnt), n, kk, reinterpret_cast<long long*>(out_val), out_idx, n_blocks, stream);
}
int kmap_dedup_hash_per_read_u32(uint32_t* hash, int64_t n, const int64_t* borders, int64_t n_seq, void* stream) {
    KMAP_REQUIRE(n >= 0 && n_seq >= 0, "negative size");
    if (n == 0 || n_seq == 0) return KMAP_OK;
    const unsigned int grid = (unsigned int)(n_seq < 148 * 16 ? n_seq : 148 * 16);
    dedup_hash_per_read_kernel<<<grid, 256, 0, as_stream(stream)>>>(hash, n, borders, n_seq);
    return kmap_check_launch("dedup_hash_per_read");
}

}  // extern "C"
