// Host side of the boundary: input.bin bytes (A0 C1 G2 T3, 255 = N / separator; kmer_count.py:244-263) -> the 2-bit packed
// + validity-mask form the device works on (same layout as pack2bit_kernel in pack.cu writes: per 32 positions one validity
// word, bit i = position i, and two packed words with the first base in the most significant bits).
//
// Why on the host: the end-to-end call is bound by the PCIe link (11.7 GB of one-byte-per-base input at ~55 GB/s = 213 of
// the 245 ms measured in round 1).  Packed, the same reads are 0.375 B/position: the host cores re-encode chunk i+1 into a
// pinned staging buffer while chunk i travels and is counted.  This is an encoder in front of the link, not a compute
// fallback: no k-mer is hashed or counted here.  No CUDA in this file (plain C++, AVX2 when the CPU has it).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
#include "../../include/kmap_b200.h"

namespace {

// 32 positions -> (valid word, hi packed word, lo packed word); p points at 32 readable bytes
inline void pack32_scalar(const uint8_t* p, uint32_t& valid, uint32_t& hi, uint32_t& lo) {
    uint32_t v = 0;
    uint64_t bits = 0;
    for (int i = 0; i < 32; ++i) {
        const uint32_t b = p[i];
        const uint32_t ok = b < 4u;
        v |= ok << i;
        bits = (bits << 2) | (ok ? b : 0u);
    }
    valid = v;
    hi = (uint32_t)(bits >> 32);
    lo = (uint32_t)bits;
}

// 32 bytes -> validity word + the two packed words (as one 64-bit value, hi word in the low half = packed[2w])
#define KMAP_PACK32(x, vword, both)                                                                                   \
    do {                                                                                                              \
        const __m256i ok_ = _mm256_cmpeq_epi8(_mm256_and_si256(x, fc), zero);          /* 0xFF where the byte is 0..3 */ \
        vword = (uint32_t)_mm256_movemask_epi8(ok_);                                                                  \
        const __m256i code_ = _mm256_and_si256(x, _mm256_and_si256(ok_, three));       /* invalid bases packed as 0 */   \
        const __m256i r_ = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code_, mul1), mul2), pick);     \
        both = (uint64_t)(uint32_t)_mm256_extract_epi32(r_, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(r_, 4) << 32); \
    } while (0)

// nt: 0 = plain stores, 1 = whole 32-byte vectors of results written with streaming stores (the words are next read by the
// copy engine, not by this core: no read-for-ownership of the staging lines, no cache pollution)
__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t* seq, int64_t w0, int64_t w1, uint32_t* packed, uint32_t* valid, int nt) {
    const __m256i fc = _mm256_set1_epi8((char)0xFC), three = _mm256_set1_epi8(3), zero = _mm256_setzero_si256();
    const __m256i mul1 = _mm256_set1_epi16(0x0104);            // maddubs: byte0 * 4 + byte1 (first base higher)
    const __m256i mul2 = _mm256_set1_epi32(0x00010010);        // madd: half0 * 16 + half1
    // byte 0 of the four 32-bit lanes of each 128-bit half, last lane first: the little-endian word then has q0 on top
    const __m256i pick = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    int64_t w = w0;
    if (nt && ((reinterpret_cast<uintptr_t>(valid) | reinterpret_cast<uintptr_t>(packed)) & 31) == 0) {
        for (; w < w1 && (w & 7); ++w) {
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(seq + 32 * w));
            uint64_t both;
            KMAP_PACK32(x, valid[w], both);
            memcpy(packed + 2 * w, &both, 8);
        }
        alignas(32) uint32_t vb[8];
        alignas(32) uint64_t pb[8];
        for (; w + 8 <= w1; w += 8) {
#pragma GCC unroll 8
            for (int j = 0; j < 8; ++j) {
                const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(seq + 32 * (w + j)));
                KMAP_PACK32(x, vb[j], pb[j]);
            }
            _mm256_stream_si256(reinterpret_cast<__m256i*>(valid + w), _mm256_load_si256(reinterpret_cast<const __m256i*>(vb)));
            _mm256_stream_si256(reinterpret_cast<__m256i*>(packed + 2 * w), _mm256_load_si256(reinterpret_cast<const __m256i*>(pb)));
            _mm256_stream_si256(reinterpret_cast<__m256i*>(packed + 2 * w + 8), _mm256_load_si256(reinterpret_cast<const __m256i*>(pb + 4)));
        }
        _mm_sfence();
    }
    for (; w < w1; ++w) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(seq + 32 * w));
        uint64_t both;
        KMAP_PACK32(x, valid[w], both);
        memcpy(packed + 2 * w, &both, 8);
    }
}

void pack_range_scalar(const uint8_t* seq, int64_t w0, int64_t w1, uint32_t* packed, uint32_t* valid) {
    for (int64_t w = w0; w < w1; ++w) pack32_scalar(seq + 32 * w, valid[w], packed[2 * w], packed[2 * w + 1]);
}

bool have_avx2() {
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
}

}  // namespace

extern "C" {

int kmap_host_threads(void) {
    const unsigned n = std::thread::hardware_concurrency();
    return (int)(n ? n : 1);
}

int kmap_host_pack2bit(const uint8_t* seq, int64_t n, uint32_t* packed, uint32_t* valid, int n_threads) {
    if (n < 0 || !packed || !valid || (!seq && n)) return KMAP_ERR_BAD_ARG;
    static const int nt = [] { const char* e = getenv("KMAP_HOST_PACK_NT"); return e ? atoi(e) : 1; }();
    const int64_t n_words = (n + 31) / 32 + 4;                 // kmap_valid_words(n): KMAP_PAD_WORDS zero words behind
    const int64_t full = n / 32;                               // words whose 32 bytes are all inside the input
    if (n_threads <= 0) n_threads = kmap_host_threads();
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, full / (1 << 15) + 1));
    auto work = [&](int64_t a, int64_t b) {
        if (have_avx2()) pack_range_avx2(seq, a, b, packed, valid, nt);
        else pack_range_scalar(seq, a, b, packed, valid);
    };
    if (n_threads == 1) {
        work(0, full);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, full * t / n_threads, full * (t + 1) / n_threads);
        for (auto& th : pool) th.join();
    }
    for (int64_t w = full; w < n_words; ++w) {                 // ragged tail + padding: positions behind n read as invalid
        uint8_t tmp[32];
        memset(tmp, 255, sizeof tmp);
        if (32 * w < n) memcpy(tmp, seq + 32 * w, (size_t)(n - 32 * w));
        pack32_scalar(tmp, valid[w], packed[2 * w], packed[2 * w + 1]);
    }
    return KMAP_OK;
}

// Border matrix of a chunk of reads in the back-to-back layout `kmap preproc` writes (kmer_count.py:335-343: st_0 = first,
// en_i = index of the separator behind read i, st_{i+1} = en_i + 1) -> one uint32 per read, en_i - st_i + 1 (bases +
// separator): 4 instead of 16 bytes per read over the link; the device rebuilds the matrix by a prefix sum
// (kmap_borders_from_strides).  Returns KMAP_ERR_BAD_ARG when the rows are not back to back from `first` (the caller then
// ships the matrix itself).
int kmap_host_border_strides(const int64_t* borders, int64_t n_seq, int64_t first, uint32_t* strides_out, int n_threads) {
    if (n_seq < 0 || (n_seq && (!borders || !strides_out))) return KMAP_ERR_BAD_ARG;
    if (n_seq == 0) return KMAP_OK;
    if (borders[0] != first) return KMAP_ERR_BAD_ARG;
    if (n_threads <= 0) n_threads = kmap_host_threads();
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_seq / (1 << 16) + 1));
    std::vector<int> bad(n_threads, 0);
    auto work = [&](int t) {
        const int64_t a = n_seq * t / n_threads, b = n_seq * (t + 1) / n_threads;
        int wrong = 0;
        for (int64_t r = a; r < b; ++r) {
            const int64_t st = borders[2 * r], en = borders[2 * r + 1];
            const int64_t stride = en - st + 1;
            wrong |= (stride < 1) | (stride > 0xFFFFFFFFll) | (r + 1 < n_seq && borders[2 * r + 2] != en + 1);
            strides_out[r] = (uint32_t)stride;
        }
        bad[t] = wrong;
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < n_threads; ++t)
        if (bad[t]) return KMAP_ERR_BAD_ARG;
    return KMAP_OK;
}

}  // extern "C"
