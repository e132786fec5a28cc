// Counter-based synthetic read generator (bench / tests).  kmap_b200/synth.py is the NumPy twin: every random
// decision is splitmix64 of (seed, read id, slot), so any read range can be produced on any GPU or on the host
// without a stateful stream and the two produce identical bytes (tests/test_synth.py).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t draw(uint64_t seed, uint64_t read, uint64_t slot) {
    return splitmix64(splitmix64(seed + read) + slot);
}
__device__ __forceinline__ float unit24(uint64_t u) { return (float)(u >> 40) * (1.0f / 16777216.0f); }

constexpr uint64_t SLOT_MOTIF = 1ull << 20;
constexpr uint64_t SLOT_N = 1ull << 21;

__global__ void __launch_bounds__(256) synth_reads_kernel(uint64_t seed, int64_t read0, int64_t n_reads, int L,
                                                          const uint8_t* __restrict__ motifs, const int32_t* __restrict__ motif_len,
                                                          const float* __restrict__ motif_cum_frac, int n_motifs, float mut_rate,
                                                          float n_rate, int64_t pos0, uint8_t* __restrict__ seq,
                                                          int64_t* __restrict__ borders) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = L + 1;
    if (idx >= n_reads * stride) return;
    const int64_t rl = idx / stride;
    const int pos = (int)(idx - rl * stride);
    const uint64_t r = (uint64_t)(read0 + rl);
    if (pos == L) {
        seq[idx] = 255;
        if (borders) { borders[2 * rl] = pos0 + rl * stride; borders[2 * rl + 1] = pos0 + rl * stride + L; }
        return;
    }
    uint32_t base = (uint32_t)(draw(seed, r, (uint64_t)(pos >> 5)) >> (2 * (pos & 31))) & 3u;
    if (n_motifs > 0) {
        const float u = unit24(draw(seed, r, SLOT_MOTIF));
        int mi = -1, off = 0;
        for (int i = 0; i < n_motifs; ++i) { if (u < motif_cum_frac[i]) { mi = i; break; } off += 32; }
        if (mi >= 0) {
            const int M = motif_len[mi];
            if (L >= M) {
                const int start = (int)(draw(seed, r, SLOT_MOTIF + 1) % (uint64_t)(L - M + 1));
                const int j = pos - start;
                if (j >= 0 && j < M && unit24(draw(seed, r, SLOT_MOTIF + 16 + (uint64_t)j)) >= mut_rate) base = motifs[off + j];
            }
        }
    }
    if (n_rate > 0.0f && unit24(draw(seed, r, SLOT_N + (uint64_t)pos)) < n_rate) base = 255;
    seq[idx] = (uint8_t)base;
}

}  // namespace

extern "C" int kmap_synth_reads(uint64_t seed, int64_t read0, int64_t n_reads, int L, const uint8_t* motifs, const int32_t* motif_len,
                                const float* motif_cum_frac, int n_motifs, float mut_rate, float n_rate, int64_t pos0, uint8_t* seq,
                                int64_t* borders, void* stream) {
    KMAP_REQUIRE(n_reads >= 0 && L >= 1 && n_motifs >= 0, "bad argument");
    if (n_reads == 0) return KMAP_OK;
    KMAP_REQUIRE(seq && (n_motifs == 0 || (motifs && motif_len && motif_cum_frac)), "null pointer");
    const int64_t total = n_reads * (int64_t)(L + 1);
    synth_reads_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(seed, read0, n_reads, L, motifs, motif_len, motif_cum_frac,
                                                                          n_motifs, mut_rate, n_rate, pos0, seq, borders);
    return kmap_check_launch("synth_reads");
}
