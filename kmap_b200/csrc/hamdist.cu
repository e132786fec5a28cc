// cal_samp_kmer_hamdist_mat (motif_discovery.py:759-808): all-pairs Hamming distance of packed k-mers as
// XOR + popcount tiles.  Output-write bound (1 B per pair): each thread owns 16 consecutive columns and emits one
// 128-bit store per row.  Pairs that share the label of a consensus shorter than k compare only the first
// head_len bases ((a >> s) ^ (b >> s) == (a ^ b) >> s, motif_discovery.py:790-800).
//
// 32-bit keys (k <= 16) take the bit-plane kernel: a key is split into the low bits and the high bits of its bases,
// 16 bits each, and a register holds the planes of TWO columns, so one XOR + one (XOR, OR) LOP3 give the mismatch
// flags of two pairs; what remains per pair is one POPC (its own pipe, 16 per clock and SM: 1e10 pairs = 2.15 ms,
// next to 1.55 ms for the 1 B/pair store at the HBM peak) and the byte packing on the FMA pipe.
#include "common.cuh"

namespace {

constexpr int HD_TX = 32, HD_TY = 8;
constexpr int HD_COLS = 16;                        // columns per thread
constexpr int HD_ROWS = 64;                        // rows per block
constexpr int HD_TILE_COLS = HD_TX * HD_COLS;      // 512 columns per block

template <typename H>
__device__ __forceinline__ uint32_t pair_dist(H x) {
    if (sizeof(H) == 4) return __popc(((uint32_t)x | ((uint32_t)x >> 1)) & 0x55555555u);
    return __popcll(((uint64_t)x | ((uint64_t)x >> 1)) & 0x5555555555555555ull);
}

template <typename H>
__global__ void __launch_bounds__(HD_TX * HD_TY) hamdist_kernel(const H* __restrict__ kh, const int32_t* __restrict__ labels,
                                                               int64_t n, int k, const int32_t* __restrict__ head_len, int n_labels,
                                                               int64_t row0, int64_t row1, uint8_t* __restrict__ out) {
    __shared__ H row_key[HD_ROWS];
    __shared__ int row_label[HD_ROWS];
    __shared__ int row_shift[HD_ROWS];
    const H low = (sizeof(H) == 4) ? (H)lowmask32(k) : (H)lowmask64(k);
    const int tid = threadIdx.y * HD_TX + threadIdx.x;
    const int64_t rbase = row0 + (int64_t)blockIdx.y * HD_ROWS;
    if (tid < HD_ROWS) {
        const int64_t i = rbase + tid;
        H a = 0; int l = -1, s = 0;
        if (i < row1) {
            a = kh[i] & low;
            l = labels ? labels[i] : -1;
            if (l >= 0 && l < n_labels) { const int hl = head_len[l]; if (hl < k) s = 2 * (k - hl); }
        }
        row_key[tid] = a; row_label[tid] = l; row_shift[tid] = s;
    }
    const int64_t j0 = (int64_t)blockIdx.x * HD_TILE_COLS + (int64_t)threadIdx.x * HD_COLS;
    H col[HD_COLS];
    int clab[HD_COLS];
#pragma unroll
    for (int c = 0; c < HD_COLS; ++c) {
        const int64_t j = j0 + c;
        col[c] = j < n ? (H)(kh[j] & low) : (H)0;
        clab[c] = (j < n && labels) ? labels[j] : -2;
    }
    __syncthreads();
    const bool vec_ok = (n % 16 == 0) && (j0 + HD_COLS <= n) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
    for (int r = threadIdx.y; r < HD_ROWS; r += HD_TY) {
        const int64_t i = rbase + r;
        if (i >= row1) break;
        const H a = row_key[r];
        const int s = row_shift[r];
        uint32_t packed_out[HD_COLS / 4];
        if (s == 0) {
#pragma unroll
            for (int q = 0; q < HD_COLS / 4; ++q) {
                uint32_t w = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) w |= pair_dist<H>(a ^ col[4 * q + b]) << (8 * b);
                packed_out[q] = w;
            }
        } else {
            const int l = row_label[r];
#pragma unroll
            for (int q = 0; q < HD_COLS / 4; ++q) {
                uint32_t w = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const H x = a ^ col[4 * q + b];
                    w |= pair_dist<H>(clab[4 * q + b] == l ? (H)(x >> s) : x) << (8 * b);
                }
                packed_out[q] = w;
            }
        }
        uint8_t* dst = out + (i - row0) * n + j0;
        if (vec_ok) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(packed_out[0], packed_out[1], packed_out[2], packed_out[3]);
        } else {
#pragma unroll
            for (int c = 0; c < HD_COLS; ++c)
                if (j0 + c < n) dst[c] = (uint8_t)(packed_out[c >> 2] >> (8 * (c & 3)));
        }
    }
}

// ---- bit-plane kernel for 32-bit keys ------------------------------------------------------------------------------------
// bits 0, 2, 4, .. of x gathered into the low 16 bits
__device__ __forceinline__ uint32_t even_bits(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

constexpr int HP_TX = 32, HP_TY = 8;
constexpr int HP_COLS = 16;                        // columns per thread (8 plane pairs)
constexpr int HP_CHUNK = 64;                       // rows staged in shared memory at a time
constexpr int HP_ROWS = 1024;                      // rows per block: the column set-up is paid once per 1024 rows
constexpr int HP_TILE_COLS = HP_TX * HP_COLS;

__global__ void __launch_bounds__(HP_TX * HP_TY) hamdist_planes_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ labels,
                                                                      int64_t n, int k, const int32_t* __restrict__ head_len,
                                                                      int n_labels, int64_t row0, int64_t row1,
                                                                      uint8_t* __restrict__ out) {
    // row planes (replicated in both halves), override mask and label: double-buffered chunks of HP_CHUNK rows
    __shared__ uint4 row_info[2][HP_CHUNK];
    const uint32_t low = lowmask32(k);
    const int tid = threadIdx.y * HP_TX + threadIdx.x;
    const int64_t rbase = row0 + (int64_t)blockIdx.y * HP_ROWS;
    const int64_t rend = min(rbase + (int64_t)HP_ROWS, row1);
    auto stage_rows = [&](int buf, int64_t r0) {
        if (tid < HP_CHUNK) {
            const int64_t i = r0 + tid;
            uint32_t a = 0, m = 0xFFFFu; int l = -1;
            if (i < rend) {
                a = __ldg(kh + i) & low;
                l = labels ? __ldg(labels + i) : -1;
                if (l >= 0 && l < n_labels) {
                    const int hl = __ldg(head_len + l);
                    if (hl < k) m = ((1u << k) - 1u) & ~((1u << (k - hl)) - 1u);     // keep the first hl of the k bases
                    else l = -1;                                                    // no override for this row
                } else {
                    l = -1;
                }
            }
            row_info[buf][tid] = make_uint4(even_bits(a) * 0x10001u, even_bits(a >> 1) * 0x10001u, m, (uint32_t)l);
        }
    };
    stage_rows(0, rbase);
    // this thread's 16 columns as 8 plane pairs
    const int64_t j0 = (int64_t)blockIdx.x * HP_TILE_COLS + (int64_t)threadIdx.x * HP_COLS;
    uint32_t qlo[HP_COLS / 2], qhi[HP_COLS / 2];
    int clab[HP_COLS];
    {
        uint32_t kc[HP_COLS];
        const bool full = j0 + HP_COLS <= n;
        if (full && (reinterpret_cast<uintptr_t>(kh) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < HP_COLS / 4; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(kh + j0) + q);
                kc[4 * q] = v.x; kc[4 * q + 1] = v.y; kc[4 * q + 2] = v.z; kc[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < HP_COLS; ++c) kc[c] = j0 + c < n ? __ldg(kh + j0 + c) : 0u;
        }
        if (labels && full && (reinterpret_cast<uintptr_t>(labels) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < HP_COLS / 4; ++q) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(labels + j0) + q);
                clab[4 * q] = v.x; clab[4 * q + 1] = v.y; clab[4 * q + 2] = v.z; clab[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < HP_COLS; ++c) clab[c] = (labels && j0 + c < n) ? __ldg(labels + j0 + c) : -2;
        }
#pragma unroll
        for (int p = 0; p < HP_COLS / 2; ++p) {
            const uint32_t b0 = kc[2 * p] & low, b1 = kc[2 * p + 1] & low;
            qlo[p] = even_bits(b0) | (even_bits(b1) << 16);
            qhi[p] = even_bits(b0 >> 1) | (even_bits(b1 >> 1) << 16);
        }
    }
    const bool vec_ok = (n % 16 == 0) && (j0 + HP_COLS <= n) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int buf = 0;
    for (int64_t r0 = rbase; r0 < rend; r0 += HP_CHUNK, buf ^= 1) {
        __syncthreads();                                   // row_info[buf] is staged; everybody is done with row_info[buf ^ 1]
        if (r0 + HP_CHUNK < rend) stage_rows(buf ^ 1, r0 + HP_CHUNK);
        if (j0 >= n) continue;
#pragma unroll 1
        for (int r = threadIdx.y; r < HP_CHUNK; r += HP_TY) {
            const int64_t i = r0 + r;
            if (i >= rend) break;
            const uint4 info = row_info[buf][r];
            const uint32_t alo = info.x, ahi = info.y;
            const int l = (int)info.w;
            uint32_t d[HP_COLS];
            if (l < 0) {
#pragma unroll
                for (int p = 0; p < HP_COLS / 2; ++p) {
                    const uint32_t f = (alo ^ qlo[p]) | (ahi ^ qhi[p]);
                    d[2 * p] = __popc(f & 0xFFFFu);
                    d[2 * p + 1] = __popc(f >> 16);
                }
            } else {
                const uint32_t m = info.z;
#pragma unroll
                for (int p = 0; p < HP_COLS / 2; ++p) {
                    const uint32_t f = (alo ^ qlo[p]) | (ahi ^ qhi[p]);
                    d[2 * p] = __popc(f & (clab[2 * p] == l ? m : 0xFFFFu));
                    d[2 * p + 1] = __popc((f >> 16) & (clab[2 * p + 1] == l ? m : 0xFFFFu));
                }
            }
            uint32_t w[HP_COLS / 4];
#pragma unroll
            for (int q = 0; q < HP_COLS / 4; ++q)
                w[q] = (d[4 * q + 3] << 24) + (d[4 * q + 2] << 16) + (d[4 * q + 1] << 8) + d[4 * q];
            uint8_t* dst = out + (i - row0) * n + j0;
            if (vec_ok) {
                __stcs(reinterpret_cast<uint4*>(dst), make_uint4(w[0], w[1], w[2], w[3]));
            } else {
#pragma unroll
                for (int c = 0; c < HP_COLS; ++c)
                    if (j0 + c < n) dst[c] = (uint8_t)d[c];
            }
        }
    }
}

template <typename H>
int launch_hamdist(const H* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels, int64_t row0,
                   int64_t row1, uint8_t* out, void* stream, int kmax) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= kmax, "k out of range for this hash width");
    KMAP_REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= n, "bad row range");
    KMAP_REQUIRE(n_labels == 0 || (labels && head_len), "labels/head_len missing");
    if (n == 0 || row0 == row1) return KMAP_OK;
    KMAP_REQUIRE(kh && out, "null pointer");
    dim3 grid((unsigned int)((n + HD_TILE_COLS - 1) / HD_TILE_COLS), (unsigned int)((row1 - row0 + HD_ROWS - 1) / HD_ROWS));
    hamdist_kernel<H><<<grid, dim3(HD_TX, HD_TY), 0, as_stream(stream)>>>(kh, n_labels ? labels : nullptr, n, k, head_len, n_labels,
                                                                        row0, row1, out);
    return kmap_check_launch("hamdist_matrix");
}

}  // namespace

extern "C" {
int kmap_hamdist_matrix_u32(const uint32_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels,
                            int64_t row0, int64_t row1, uint8_t* out, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 16, "k out of range for this hash width");
    KMAP_REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= n, "bad row range");
    KMAP_REQUIRE(n_labels == 0 || (labels && head_len), "labels/head_len missing");
    if (n == 0 || row0 == row1) return KMAP_OK;
    KMAP_REQUIRE(kh && out, "null pointer");
    dim3 grid((unsigned int)((n + HP_TILE_COLS - 1) / HP_TILE_COLS), (unsigned int)((row1 - row0 + HP_ROWS - 1) / HP_ROWS));
    hamdist_planes_kernel<<<grid, dim3(HP_TX, HP_TY), 0, as_stream(stream)>>>(kh, n_labels ? labels : nullptr, n, k, head_len, n_labels,
                                                                              row0, row1, out);
    return kmap_check_launch("hamdist_matrix");
}
int kmap_hamdist_matrix_u64(const uint64_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels,
                            int64_t row0, int64_t row1, uint8_t* out, void* stream) {
    return launch_hamdist<uint64_t>(kh, labels, n, k, head_len, n_labels, row0, row1, out, stream, 32);
}
}
