// input.bin bytes (A0 C1 G2 T3, 255 = N / separator; kmer_count.py:244-263) <-> 2-bit packed + validity mask.
#include "common.cuh"

namespace {

// 4 bytes (each 0..3) held in a little-endian word -> 8 bits, first byte in the two MOST significant bits.
// y = b0 + b1<<8 + b2<<16 + b3<<24; the multiplier drops b0..b3 at bits 30,28,26,24 with no overlapping terms.
__device__ __forceinline__ uint32_t squeeze4(uint32_t x) { return ((x & 0x03030303u) * 0x40100401u) >> 24; }
// 4 bytes -> 4 validity bits (bit i set iff byte i < 4), first byte in bit 0
__device__ __forceinline__ uint32_t valid4(uint32_t x) {
    const uint32_t z = x & 0xFCFCFCFCu;
    const uint32_t nz = (z | ((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;  // 0x80 where the byte is non-zero
    const uint32_t ok = (~nz >> 7) & 0x01010101u;
    return ((ok * 0x01020408u) >> 24) & 0xFu;
}

// one thread = 32 positions = one validity word + two packed words; 2 x 128-bit loads when in range
__global__ void __launch_bounds__(256) pack2bit_kernel(const uint8_t* __restrict__ seq, int64_t n, int64_t n_valid_words,
                                                       uint32_t* __restrict__ packed, uint32_t* __restrict__ valid) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid_words) return;
    const int64_t p0 = t * 32;
    uint32_t w[8];
    if (p0 + 32 <= n && ((reinterpret_cast<uintptr_t>(seq) & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(seq + p0));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(seq + p0 + 16));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t x = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t p = p0 + 4 * j + b;
                const uint32_t v = p < n ? (uint32_t)__ldg(seq + p) : 255u;
                x |= v << (8 * b);
            }
            w[j] = x;
        }
    }
    uint32_t vbits = 0, hi = 0, lo = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t v4 = valid4(w[j]);
        vbits |= v4 << (4 * j);
        // zero the 2-bit code of invalid bases so the packed word is canonical
        uint32_t keep = v4 | (v4 << 7) | (v4 << 14) | (v4 << 21);          // bit 8i <- validity of byte i
        keep = (keep & 0x01010101u) * 3u;
        const uint32_t q = squeeze4(w[j] & keep);
        if (j < 4) hi |= q << (24 - 8 * j); else lo |= q << (24 - 8 * (j - 4));
    }
    valid[t] = vbits;
    packed[2 * t] = hi;
    packed[2 * t + 1] = lo;
}

__global__ void __launch_bounds__(256) apply_valid_kernel(uint8_t* __restrict__ seq, int64_t n,
                                                          const uint32_t* __restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 positions
    const int64_t p0 = i * 4;
    if (p0 >= n) return;
    const uint32_t v = (__ldg(valid + (p0 >> 5)) >> (p0 & 31)) & 0xFu;
    if (v == 0xFu) return;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (p0 + b < n && !((v >> b) & 1u)) seq[p0 + b] = 255;
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t value) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = value;
}

// dst[i] += src[i], four cells per thread
__global__ void add_u32_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int64_t n4, uint32_t* __restrict__ dst1,
                               const uint32_t* __restrict__ src1, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = i; j < n4; j += stride) {
        uint4 a = dst[j];
        const uint4 b = __ldcs(src + j);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[j] = a;
    }
    for (int64_t j = 4 * n4 + i; j < n; j += stride) dst1[j] += src1[j];
}

__global__ void rebase_borders_kernel(int64_t* __restrict__ borders, int64_t n_values, int64_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_values; i += stride) borders[i] -= offset;
}

// offsets[r] = sum of the strides of reads < r (int64[n_seq + 1]) -> rows [start, separator index] of the border matrix
__global__ void borders_from_offsets_kernel(const int64_t* __restrict__ offsets, int64_t n_seq, int64_t* __restrict__ borders) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < n_seq; r += stride) {
        const int64_t st = __ldg(offsets + r), nx = __ldg(offsets + r + 1);
        reinterpret_cast<longlong2*>(borders)[r] = make_longlong2(st, nx - 1);
    }
}

}  // namespace

extern "C" {

int kmap_borders_from_strides(const uint32_t* strides, int64_t n_seq, int64_t* offsets, int64_t* borders_out, uint64_t* scratch, void* stream) {
    KMAP_REQUIRE(n_seq >= 0, "negative size");
    if (n_seq == 0) return KMAP_OK;
    KMAP_REQUIRE(strides && offsets && borders_out && scratch, "null pointer");
    KMAP_REQUIRE(((uintptr_t)borders_out & 15) == 0, "the border matrix must be 16-byte aligned");
    int rc = kmap_exclusive_scan_u32(strides, n_seq, offsets, scratch, stream);
    if (rc) return rc;
    int64_t g = (n_seq + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    borders_from_offsets_kernel<<<(unsigned int)g, 256, 0, as_stream(stream)>>>(offsets, n_seq, borders_out);
    return kmap_check_launch("borders_from_strides");
}

int kmap_add_u32(uint32_t* dst, const uint32_t* src, int64_t n_words, void* stream) {
    KMAP_REQUIRE(n_words >= 0, "negative size");
    if (n_words == 0) return KMAP_OK;
    KMAP_REQUIRE(dst && src, "null pointer");
    KMAP_REQUIRE(((uintptr_t)dst & 15) == 0 && ((uintptr_t)src & 15) == 0, "tables must be 16-byte aligned");
    int64_t g = (n_words / 4 + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    add_u32_kernel<<<(unsigned int)g, 256, 0, as_stream(stream)>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src),
                                                                  n_words / 4, dst, src, n_words);
    return kmap_check_launch("add_u32");
}

int kmap_rebase_borders(int64_t* borders, int64_t n_seq, int64_t offset, void* stream) {
    KMAP_REQUIRE(n_seq >= 0, "negative size");
    if (n_seq == 0 || offset == 0) return KMAP_OK;
    KMAP_REQUIRE(borders, "null pointer");
    int64_t g = (2 * n_seq + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    rebase_borders_kernel<<<(unsigned int)g, 256, 0, as_stream(stream)>>>(borders, 2 * n_seq, offset);
    return kmap_check_launch("rebase_borders");
}

int64_t kmap_valid_words(int64_t n) { return (n + 31) / 32 + KMAP_PAD_WORDS; }
int64_t kmap_packed_words(int64_t n) { return 2 * kmap_valid_words(n); }

int kmap_pack2bit(const uint8_t* seq, int64_t n, uint32_t* packed, uint32_t* valid, void* stream) {
    KMAP_REQUIRE(n >= 0, "negative size");
    KMAP_REQUIRE(packed && valid && (seq || n == 0), "null pointer");
    const int64_t nv = kmap_valid_words(n);
    pack2bit_kernel<<<grid_for(nv, 256), 256, 0, as_stream(stream)>>>(seq, n, nv, packed, valid);
    return kmap_check_launch("pack2bit");
}

int kmap_apply_valid_to_seq(uint8_t* seq, int64_t n, const uint32_t* valid, void* stream) {
    KMAP_REQUIRE(n >= 0, "negative size");
    if (n == 0) return KMAP_OK;
    KMAP_REQUIRE(seq && valid, "null pointer");
    apply_valid_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(seq, n, valid);
    return kmap_check_launch("apply_valid_to_seq");
}

int kmap_fill_u32(uint32_t* p, int64_t n_words, uint32_t value, void* stream) {
    KMAP_REQUIRE(n_words >= 0, "negative size");
    if (n_words == 0) return KMAP_OK;
    KMAP_REQUIRE(p, "null pointer");
    if (value == 0) {
        cudaError_t e = cudaMemsetAsync(p, 0, (size_t)n_words * 4, as_stream(stream));
        if (e != cudaSuccess) { kmap_set_error("kmap_fill_u32: %s", cudaGetErrorString(e)); return (int)e; }
        return KMAP_OK;
    }
    int64_t g = (n_words + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    fill_u32_kernel<<<(unsigned int)g, 256, 0, as_stream(stream)>>>(p, n_words, value);
    return kmap_check_launch("fill_u32");
}

}  // extern "C"
