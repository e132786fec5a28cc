// cal_samp_kmer_hamdist_mat (motif_discovery.py:759-808): all-pairs Hamming distance of packed k-mers as
// XOR + popcount tiles.  Output-write bound (1 B per pair): each thread owns 16 consecutive columns and emits one
// 128-bit store per row.  Pairs that share the label of a consensus shorter than k compare only the first
// head_len bases ((a >> s) ^ (b >> s) == (a ^ b) >> s, motif_discovery.py:790-800).
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
    return launch_hamdist<uint32_t>(kh, labels, n, k, head_len, n_labels, row0, row1, out, stream, 16);
}
int kmap_hamdist_matrix_u64(const uint64_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels,
                            int64_t row0, int64_t row1, uint8_t* out, void* stream) {
    return launch_hamdist<uint64_t>(kh, labels, n, k, head_len, n_labels, row0, row1, out, stream, 32);
}
}
