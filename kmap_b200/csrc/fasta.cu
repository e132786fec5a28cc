// preproc ingest: FASTA text -> input.bin / input.seqboarder.bin contents (kmer_count.py:244-263 dna2arr,
// 308-323 read_dnaseq_file, 326-347 convert_fasta_to_binary) on the device.
//
// The reference walks the file twice through Bio.SeqIO and encodes base by base in a Python loop.  Here the file bytes
// are parsed data-parallel.  What a byte becomes depends on the TYPE OF ITS LINE (header line = starts with '>', or
// sequence line), i.e. on the most recent line start, which can be arbitrarily far back.  So:
//   1. fasta_tile_summary_kernel   per tile of 4096 bytes: (type of the last line started in the tile or UNKNOWN, sequence
//                                  characters on lines of known type SEQUENCE, characters on the line that was already in
//                                  progress when the tile began, header lines started)
//   2. fasta_tile_scan_kernel      one block: scan of the summaries under the (associative, non-commutative) operator `comb`
//                                  -> per tile: type of the line in progress at its first byte, sequence characters and
//                                  records before it
//   3. fasta_emit_kernel           per tile again, now with everything known: every sequence character becomes its code
//                                  (upper-cased; A0 C1 G2 T3, anything else 255), every header line start except the very
//                                  first becomes the 255 separator that closes the previous record, and announces where its
//                                  own record starts.  The tile's outputs are contiguous: staged in shared memory, written
//                                  coalesced.
// Text conventions (what the reference's text-mode file iteration + "".join(line.split()) amount to for ASCII input):
// '\n' and '\r' both end a line; white space (9-13, 28-32) inside sequence lines is dropped; UTF-8 continuation bytes are
// dropped so that one non-ASCII character gives one 255 as it does in text mode; text before the first header line is
// ignored (the host passes the text from the first header on).
// The text may be fed in chunks: the 4-value state {sequence characters so far, records so far, type of the line in
// progress, last byte} carries over.
#include "common.cuh"

namespace {

constexpr int FA_THREADS = 512;
constexpr int FA_PER = 16;                          // bytes per thread: one 128-bit load
constexpr int FA_TILE = FA_THREADS * FA_PER;        // 8192 bytes
enum : int { FA_SEQ = 0, FA_HDR = 1, FA_UNK = 2 };

// Byte classes are computed four bytes per word with plain integer arithmetic (the __vcmp*4 intrinsics are emulated by ~10
// instructions each on this architecture -- measured: the kernels were bound by them), giving a 0x80 flag per byte, then
// squeezed to one bit per byte, so that everything after that is bit arithmetic on 16-bit masks.
__device__ __forceinline__ uint32_t zero80(uint32_t v) { return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u; }   // 0x80 per zero byte
__device__ __forceinline__ uint32_t eq80(uint32_t w, uint32_t c) { return zero80(w ^ c); }
__device__ __forceinline__ uint32_t bits80(uint32_t m) { return ((m >> 7) * 0x01020408u) >> 24; }              // 0x80 flags -> 4 bits
// ASCII bytes with lo <= c <= hi (constants replicated in every byte, < 128)
__device__ __forceinline__ uint32_t range80(uint32_t w, uint32_t lo, uint32_t hi) {
    const uint32_t a = w & 0x7F7F7F7Fu;
    const uint32_t x = a + (0x80808080u - lo), y = a + (0x7F7F7F7Fu - hi);
    return x & ~y & ~w & 0x80808080u;
}
__device__ __forceinline__ uint32_t eol80(uint32_t w) { return eq80(w, 0x0A0A0A0Au) | eq80(w, 0x0D0D0D0Du); }
// dropped from sequence lines: white space as str.split() sees it (9-13, 28-32) and UTF-8 continuation bytes (10xxxxxx)
__device__ __forceinline__ uint32_t dropped80(uint32_t w) {
    return range80(w, 0x09090909u, 0x0D0D0D0Du) | range80(w, 0x1C1C1C1Cu, 0x20202020u) | (w & ~(w << 1) & 0x80808080u);
}
// upper() + dna2arr (kmer_count.py:244-263, 316): A0 C1 G2 T3 in either case, anything else 255.
// (c >> 1) & 3 maps A C T G to 0 1 2 3; x ^ (x >> 1) swaps the last two.
__device__ __forceinline__ uint32_t code4(uint32_t w) {
    const uint32_t lower = w | 0x20202020u;
    const uint32_t ok = (eq80(lower & 0xFDFDFDFDu, 0x61616161u) | eq80(lower, 0x67676767u) | eq80(lower, 0x74747474u)) >> 7;   // 0x01 per a/c/g/t
    const uint32_t x = (w >> 1) & 0x03030303u;
    return (x ^ ((x >> 1) & 0x01010101u)) | ((ok ^ 0x01010101u) * 255u);
}

struct ThreadBytes {
    uint32_t w[4];        // the 16 bytes, byte j = (w[j >> 2] >> (8 * (j & 3))) & 255
    uint32_t ls, hdr, ch; // bit j: byte j starts a line / starts a header line / is a character if its line is a sequence line
    int n;                // bytes that exist (0..16)
};

__device__ __forceinline__ ThreadBytes load_thread_bytes(const uint8_t* __restrict__ text, int64_t n, int64_t tile, uint32_t carry_last_byte) {
    ThreadBytes t;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0;
    t.ls = t.hdr = t.ch = 0;
    const int64_t i0 = tile * FA_TILE + (int64_t)threadIdx.x * FA_PER;
    t.n = (int)(n - i0 < 0 ? 0 : (n - i0 > FA_PER ? FA_PER : n - i0));
    if (t.n == 0) return t;
    if (t.n == FA_PER) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + i0));
        t.w[0] = v.x; t.w[1] = v.y; t.w[2] = v.z; t.w[3] = v.w;
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < FA_PER; ++j)
            if (j < t.n) w[j >> 2] |= (uint32_t)__ldg(text + i0 + j) << (8 * (j & 3));
        t.w[0] = w[0]; t.w[1] = w[1]; t.w[2] = w[2]; t.w[3] = w[3];
    }
    const uint32_t prev = i0 > 0 ? (uint32_t)__ldg(text + i0 - 1) : carry_last_byte;
    const uint32_t exist = t.n == FA_PER ? 0xFFFFu : ((1u << t.n) - 1u);
    uint32_t eol = 0, gt = 0, drop = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        eol |= bits80(eol80(t.w[q])) << (4 * q);
        gt |= bits80(eq80(t.w[q], 0x3E3E3E3Eu)) << (4 * q);
        drop |= bits80(dropped80(t.w[q])) << (4 * q);
    }
    t.ls = ((eol << 1) | (prev == 10u || prev == 13u ? 1u : 0u)) & exist;
    t.hdr = t.ls & gt;
    t.ch = ~drop & exist;
    return t;
}

// what a thread's 16 bytes hold, by line type: characters on lines that begin inside the thread's bytes and are sequence
// lines (known), characters on the line that was in progress at the thread's first byte (pending), the type of the last
// line started (FA_UNK if none).  seq_fill = the bytes on sequence lines that begin inside the thread's bytes: every
// sequence-line start is extended upwards to the byte before the next line start by letting a carry ripple through the
// run of non-start bytes above it; before = the bytes before the first line start.
struct ThreadCounts { uint32_t known_seq, pending, seq_fill, before; int last; };
__device__ __forceinline__ ThreadCounts count_thread(const ThreadBytes& t) {
    ThreadCounts c;
    const uint32_t not_start = ~t.ls;
    const uint32_t seq_start = t.ls & ~t.hdr;
    c.seq_fill = seq_start | (((not_start + (seq_start << 1)) ^ not_start) & not_start);
    c.before = t.ls ? ((t.ls & (0u - t.ls)) - 1u) : 0xFFFFFFFFu;
    c.known_seq = __popc(t.ch & c.seq_fill);
    c.pending = __popc(t.ch & c.before);
    c.last = t.ls ? (((t.hdr >> (31 - __clz(t.ls))) & 1u) ? FA_HDR : FA_SEQ) : FA_UNK;
    return c;
}

// type of the line in progress at this thread's first byte: the last line started by an earlier thread of the tile, else
// `carry` (the tile's own incoming type; FA_UNK in pass 1).  *tile_last = the same for the byte after the tile.
// warp_last: FA_THREADS / 32 + 1 ints of shared memory (the cross-warp part is done once, by warp 0).
__device__ __forceinline__ int incoming_type(int last, int carry, int* warp_last, int* tile_last) {
    constexpr int NW = FA_THREADS / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o && incl == FA_UNK) incl = y;
    }
    if (lane == 31) warp_last[w] = incl;
    __syncthreads();
    if (w == 0) {
        // exclusive "last known" scan over the warps, seeded with the carry; slot NW = the whole tile
        int v = lane < NW ? warp_last[lane] : FA_UNK;
        int sc = v;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, sc, o);
            if (lane >= o && sc == FA_UNK) sc = y;
        }
        int ex = __shfl_up_sync(0xFFFFFFFFu, sc, 1);
        if (lane == 0 || ex == FA_UNK) ex = carry;
        if (sc == FA_UNK) sc = carry;
        __syncwarp();
        if (lane < NW) warp_last[lane] = ex;
        if (lane == NW - 1) warp_last[NW] = sc;
    }
    __syncthreads();
    int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0 || excl == FA_UNK) excl = warp_last[w];
    *tile_last = warp_last[NW];
    return excl;
}

// sum over the block of three small counts packed into one 64-bit value (16 bits each would do; 21 are used); the
// result is valid in thread 0 only.  ws: FA_THREADS / 32 words of shared memory.
__device__ __forceinline__ unsigned long long block_sum3(uint32_t a, uint32_t b, uint32_t c, unsigned long long* ws) {
    unsigned long long v = (unsigned long long)a | ((unsigned long long)b << 21) | ((unsigned long long)c << 42);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x == 0)
        for (int q = 0; q < FA_THREADS / 32; ++q) t += ws[q];
    return t;
}

// ---- 1. per-tile summaries ---------------------------------------------------------------------------------------
struct TileSummary { uint32_t last, known_seq, pending, n_hdr; };

__global__ void __launch_bounds__(FA_THREADS) fasta_tile_summary_kernel(const uint8_t* __restrict__ text, int64_t n, uint32_t carry_last_byte,
                                                                        TileSummary* __restrict__ summ) {
    __shared__ int warp_last[FA_THREADS / 32 + 1];
    __shared__ unsigned long long ws[FA_THREADS / 32];
    const ThreadBytes t = load_thread_bytes(text, n, blockIdx.x, carry_last_byte);
    const ThreadCounts c = count_thread(t);
    int tile_last;
    const int in_type = incoming_type(c.last, FA_UNK, warp_last, &tile_last);
    const unsigned long long sums = block_sum3(c.known_seq + (in_type == FA_SEQ ? c.pending : 0u), in_type == FA_UNK ? c.pending : 0u,
                                               __popc(t.hdr), ws);
    if (threadIdx.x == 0) {
        TileSummary s;
        s.last = (uint32_t)tile_last;
        s.known_seq = (uint32_t)(sums & 0x1FFFFFull); s.pending = (uint32_t)((sums >> 21) & 0x1FFFFFull); s.n_hdr = (uint32_t)(sums >> 42);
        summ[blockIdx.x] = s;
    }
}

// ---- 2. scan of the summaries -----------------------------------------------------------------------------------------
struct Span { int last; unsigned long long seq, pending, n_hdr; };
__device__ __forceinline__ Span span_identity() { Span s; s.last = FA_UNK; s.seq = s.pending = s.n_hdr = 0; return s; }
// the text of a followed by the text of b
__device__ __forceinline__ Span comb(const Span& a, const Span& b) {
    Span r;
    r.last = b.last != FA_UNK ? b.last : a.last;
    r.seq = a.seq + b.seq + (a.last == FA_SEQ ? b.pending : 0ull);
    r.pending = a.pending + (a.last == FA_UNK ? b.pending : 0ull);
    r.n_hdr = a.n_hdr + b.n_hdr;
    return r;
}
__device__ __forceinline__ Span span_of(const TileSummary& t) {
    Span s; s.last = (int)t.last; s.seq = t.known_seq; s.pending = t.pending; s.n_hdr = t.n_hdr; return s;
}
__device__ __forceinline__ Span shfl_up_span(const Span& s, int o) {
    Span r;
    r.last = __shfl_up_sync(0xFFFFFFFFu, s.last, o);
    r.seq = __shfl_up_sync(0xFFFFFFFFu, s.seq, o);
    r.pending = __shfl_up_sync(0xFFFFFFFFu, s.pending, o);
    r.n_hdr = __shfl_up_sync(0xFFFFFFFFu, s.n_hdr, o);
    return r;
}

struct TileStart { unsigned long long seq_and_type, n_hdr; };      // (sequence characters before the tile) << 2 | line type; records before

constexpr int FS_THREADS = 1024;
// totals[0..3] = sequence characters, records, type of the line in progress, last byte -- after this chunk (global counts)
__global__ void __launch_bounds__(FS_THREADS) fasta_tile_scan_kernel(const TileSummary* __restrict__ summ, int64_t n_tiles,
                                                                     const uint8_t* __restrict__ text, int64_t n,
                                                                     unsigned long long seq0, unsigned long long hdr0, int type0,
                                                                     uint32_t last_byte0, TileStart* __restrict__ starts,
                                                                     unsigned long long* __restrict__ totals) {
    __shared__ Span warp_tot[FS_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t per = (n_tiles + FS_THREADS - 1) / FS_THREADS;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    Span mine = span_identity();
    for (int64_t t = lo; t < hi; ++t) mine = comb(mine, span_of(summ[t]));
    Span incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Span y = shfl_up_span(incl, o);
        if (lane >= o) incl = comb(y, incl);
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    Span pre;                                       // everything before this thread's range, the carried state included
    pre.last = type0; pre.seq = seq0; pre.pending = 0; pre.n_hdr = hdr0;
    for (int q = 0; q < w; ++q) pre = comb(pre, warp_tot[q]);
    Span excl = shfl_up_span(incl, 1);
    if (lane == 0) excl = span_identity();
    pre = comb(pre, excl);
    for (int64_t t = lo; t < hi; ++t) {
        TileStart ts;
        ts.seq_and_type = (pre.seq << 2) | (unsigned long long)pre.last;
        ts.n_hdr = pre.n_hdr;
        starts[t] = ts;
        pre = comb(pre, span_of(summ[t]));
    }
    if (threadIdx.x == FS_THREADS - 1) {            // its range is the last one (possibly empty): pre = the whole chunk
        totals[0] = pre.seq; totals[1] = pre.n_hdr; totals[2] = (unsigned long long)pre.last;
        totals[3] = n > 0 ? (unsigned long long)text[n - 1] : (unsigned long long)last_byte0;
    }
}

// ---- 3. emit ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FA_THREADS) fasta_emit_kernel(const uint8_t* __restrict__ text, int64_t n, uint32_t carry_last_byte,
                                                                const TileStart* __restrict__ starts, uint8_t* __restrict__ seq_out,
                                                                long long seq_origin, long long* __restrict__ rec_start,
                                                                unsigned long long hdr0) {
    __shared__ int warp_last[FA_THREADS / 32 + 1];
    __shared__ uint32_t wsum[FA_THREADS / 32 + 1];
    __shared__ uint8_t stage[FA_TILE];
    constexpr int NW = FA_THREADS / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const ThreadBytes t = load_thread_bytes(text, n, blockIdx.x, carry_last_byte);
    const ThreadCounts c = count_thread(t);
    const TileStart ts = starts[blockIdx.x];
    int tile_last;
    const int in_type = incoming_type(c.last, (int)(ts.seq_and_type & 3ull), warp_last, &tile_last);
    const uint32_t n_seq = c.known_seq + (in_type == FA_SEQ ? c.pending : 0u);
    const uint32_t n_hdr = __popc(t.hdr);
    // exclusive scan of (outputs, headers) per thread; both fit 16 bits (a tile has 8192 bytes)
    const uint32_t v = (n_seq + n_hdr) | (n_hdr << 16);
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {                                               // the cross-warp part once: exclusive prefixes, slot NW = total
        const uint32_t mine = lane < NW ? wsum[lane] : 0u;
        uint32_t sc = mine;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, sc, o);
            if (lane >= o) sc += y;
        }
        __syncwarp();
        if (lane < NW) wsum[lane] = sc - mine;
        if (lane == NW - 1) wsum[NW] = sc;
    }
    __syncthreads();
    const uint32_t total = wsum[NW];
    const uint32_t excl = wsum[w] + incl - v;
    uint32_t pos = excl & 0xFFFFu;                              // index among the tile's outputs
    unsigned long long rec = ts.n_hdr + (excl >> 16);           // global index of the next record to start
    // global position of the tile's output number i: first + i, where (sequence characters + separators) written before the
    // tile = seq + max(records - 1, 0); the very first header of the file has no separator in front, its slot is position -1
    const long long first = (long long)(ts.seq_and_type >> 2) + (long long)ts.n_hdr - 1;
    // bytes that produce an output: characters on sequence lines, and header starts (the 255 that closes the previous
    // record, kmer_count.py:261-262 -- '>' encodes to 255 by itself)
    const uint32_t sel = ((c.seq_fill | (in_type == FA_SEQ ? c.before : 0u)) & t.ch) | t.hdr;
    uint32_t code[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) code[q] = code4(t.w[q]);
#pragma unroll
    for (int j = 0; j < FA_PER; ++j)
        if ((sel >> j) & 1u) stage[pos + __popc(sel & ((1u << j) - 1u))] = (uint8_t)(code[j >> 2] >> (8 * (j & 3)));
    for (uint32_t hb = t.hdr; hb; hb &= hb - 1) {
        const int j = __ffs(hb) - 1;
        rec_start[rec - hdr0] = first + (long long)(pos + __popc(sel & ((1u << j) - 1u))) + 1;
        ++rec;
    }
    __syncthreads();
    const uint32_t n_out = total & 0xFFFFu;
    for (uint32_t i = threadIdx.x; i < n_out; i += FA_THREADS) {
        const long long p = first + (long long)i;
        if (p >= seq_origin && p >= 0) seq_out[p - seq_origin] = stage[i];
    }
}

__global__ void fasta_final_separator_kernel(uint8_t* seq_out, long long index) { seq_out[index] = 255; }

__global__ void __launch_bounds__(256) borders_from_starts_kernel(const long long* __restrict__ rec_start, int64_t n_rec, long long total_len,
                                                                  long long* __restrict__ borders) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= n_rec) return;
    const long long st = rec_start[r];
    const long long nx = r + 1 < n_rec ? rec_start[r + 1] : total_len;
    borders[2 * r] = st;
    borders[2 * r + 1] = nx - 1;                                // index of the record's separator (kmer_count.py:339-340)
}

static int64_t fa_tiles(int64_t n) { return (n + FA_TILE - 1) / FA_TILE; }

}  // namespace

extern "C" {

int64_t kmap_fasta_scratch_words(int64_t n) { return n < 0 ? 0 : 4 + 4 * fa_tiles(n); }

int kmap_fasta_scan(const uint8_t* text, int64_t n, const int64_t* state_in_host, uint64_t* scratch, int64_t* state_out_host, void* stream) {
    KMAP_REQUIRE(n >= 0 && state_in_host && state_out_host && scratch, "bad argument");
    KMAP_REQUIRE(n == 0 || text, "null pointer");
    KMAP_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15u) == 0, "text must be 16-byte aligned");
    KMAP_REQUIRE(state_in_host[2] == FA_SEQ || state_in_host[2] == FA_HDR, "state: line type must be 0 or 1");
    cudaStream_t s = as_stream(stream);
    const int64_t n_tiles = fa_tiles(n);
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(scratch);
    TileSummary* summ = reinterpret_cast<TileSummary*>(scratch + 4);
    TileStart* starts = reinterpret_cast<TileStart*>(scratch + 4 + 2 * n_tiles);
    if (n_tiles) fasta_tile_summary_kernel<<<(unsigned int)n_tiles, FA_THREADS, 0, s>>>(text, n, (uint32_t)state_in_host[3], summ);
    fasta_tile_scan_kernel<<<1, FS_THREADS, 0, s>>>(summ, n_tiles, text, n, (unsigned long long)state_in_host[0],
                                                    (unsigned long long)state_in_host[1], (int)state_in_host[2],
                                                    (uint32_t)state_in_host[3], starts, totals);
    int rc = kmap_check_launch("fasta_scan");
    if (rc) return rc;
    unsigned long long h[4];
    cudaError_t e = cudaMemcpyAsync(h, totals, 32, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { kmap_set_error("fasta_scan: %s", cudaGetErrorString(e)); return (int)e; }
    for (int i = 0; i < 4; ++i) state_out_host[i] = (int64_t)h[i];
    return KMAP_OK;
}

int kmap_fasta_emit(const uint8_t* text, int64_t n, const int64_t* state_in_host, const uint64_t* scratch, uint8_t* seq_out,
                    int64_t seq_origin, int64_t* rec_start_out, int final_chunk, const int64_t* state_out_host, void* stream) {
    KMAP_REQUIRE(n >= 0 && state_in_host && state_out_host && scratch, "bad argument");
    cudaStream_t s = as_stream(stream);
    const int64_t n_tiles = fa_tiles(n);
    const int64_t new_records = state_out_host[1] - state_in_host[1];
    const int64_t outputs = (state_out_host[0] - state_in_host[0]) + new_records            // (the file's first header has no separator in front)
                            - (state_in_host[1] == 0 && state_out_host[1] > 0 ? 1 : 0) + (final_chunk && state_out_host[1] > 0 ? 1 : 0);
    KMAP_REQUIRE(outputs == 0 || seq_out, "null pointer (seq_out)");
    KMAP_REQUIRE(new_records == 0 || rec_start_out, "null pointer (rec_start_out)");
    const TileStart* starts = reinterpret_cast<const TileStart*>(scratch + 4 + 2 * n_tiles);
    if (n_tiles) {
        fasta_emit_kernel<<<(unsigned int)n_tiles, FA_THREADS, 0, s>>>(text, n, (uint32_t)state_in_host[3], starts, seq_out, (long long)seq_origin,
                                                                       reinterpret_cast<long long*>(rec_start_out),
                                                                       (unsigned long long)state_in_host[1]);
    }
    if (final_chunk && state_out_host[1] > 0) {     // the separator of the last record
        const long long last = state_out_host[0] + state_out_host[1] - 1;
        fasta_final_separator_kernel<<<1, 1, 0, s>>>(seq_out, last - (long long)seq_origin);
    }
    return kmap_check_launch("fasta_emit");
}

int kmap_borders_from_starts(const int64_t* rec_start, int64_t n_rec, int64_t total_len, int64_t* borders, void* stream) {
    KMAP_REQUIRE(n_rec >= 0, "bad argument");
    if (n_rec == 0) return KMAP_OK;
    KMAP_REQUIRE(rec_start && borders, "null pointer");
    borders_from_starts_kernel<<<grid_for(n_rec, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(rec_start), n_rec,
                                                                                   (long long)total_len, reinterpret_cast<long long*>(borders));
    return kmap_check_launch("borders_from_starts");
}

}  // extern "C"
