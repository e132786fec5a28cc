// cal_samp_kmer_hamdist_mat (motif_discovery.py:759-808) as an int8 GEMM on the 5th-generation tensor cores: the comparator
// the distance-matrix kernel of hamdist.cu (XOR + popcount) is benchmarked against (BASELINE.json north_star, config 5).
//
// Formulation.  A k-mer becomes a row of K = 64 int8 (4 per base, 16 bases; bases >= k are zero):
//     A[i][4b + c] = [base b of k-mer i has code c]          (one-hot)
//     B[j][4b + c] = [base b of k-mer j has code != c]       (its complement over the k bases)
// so that (A B^T)[i][j] = sum_b [base b differs] = the Hamming distance itself: the accumulator needs no "k - matches" in
// the epilogue.  Pairs that share the label of a consensus shorter than k compare only the head (md:790-800); those few
// are recomputed in the epilogue from the keys.
//
// Kernel (one persistent CTA per SM, 10 warps, hand-written tcgen05 / TMEM / mbarrier / cp.async.bulk PTX):
//   warp 0      producer: one cp.async.bulk per operand tile (the operands are pre-laid out in global memory in the
//               canonical K-major no-swizzle core-matrix order, so a 128-row A tile is 8 KB and a 256-row B tile 16 KB of
//               contiguous bytes) into a 4-stage shared-memory ring, completion on an mbarrier (expect_tx)
//   warp 1      allocates the 512 TMEM columns (two 128 x 256 int32 accumulators), then one lane issues, per output tile,
//               two tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32 each) and tcgen05.commit's the stage's
//               "empty" barrier and the accumulator's "full" barrier
//   warps 2..9  epilogue: tcgen05.ld 32x32b.x32 (warp w reads TMEM lanes 32 (w % 4) .. + 31, half of the columns), byte
//               packing, swizzled staging in shared memory, coalesced 16-byte streaming stores of 256-byte row segments
// The output (1 B per pair) is what bounds it, exactly as for the popcount kernel: the tensor pipe is idle > 95 % of the time.
#include "common.cuh"

namespace {

constexpr int MM_M = 128;                 // rows of an output tile (TMEM lanes)
constexpr int MM_N = 256;                 // columns of an output tile (TMEM columns of one accumulator)
constexpr int MM_KB = 64;                 // bytes of one operand row (K = 64 int8)
constexpr int MM_A_BYTES = MM_M * MM_KB;  // 8 KB
constexpr int MM_B_BYTES = MM_N * MM_KB;  // 16 KB
constexpr int MM_STAGES = 4;
constexpr int MM_EPI_WARPS = 8;
constexpr int MM_THREADS = 32 * (2 + MM_EPI_WARPS);
constexpr int MM_STAGE_BYTES = MM_A_BYTES + MM_B_BYTES;
constexpr int MM_OUT_BYTES = MM_M * MM_N;                     // 32 KB of output per tile
constexpr int MM_SMEM = MM_STAGES * MM_STAGE_BYTES + 2 * MM_OUT_BYTES + 2 * MM_N * 8 + 1024;
constexpr uint32_t MM_SPIN_LIMIT = 1u << 28;                  // a wait that long is a bug: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > MM_SPIN_LIMIT) __trap();
    }
}
// global -> shared bulk copy (TMA engine, no tensor map: the tile is contiguous), completion counted on the mbarrier
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor; canonical layout in
// 16-byte units ((8, n), 2) : ((1, SBO), LBO)): rows of a core matrix 16 B apart, LBO = distance of the two 16-byte K chunks
// of one MMA (K = 32 int8), SBO = distance of consecutive 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                                   // descriptor version of sm_100
    return d;                                                 // base offset 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor (InstrDescriptor): dense, no saturate, D = S32, A = B = signed int8, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t instr_desc_i8() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MM_N >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {      // arrives on the barrier when the MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand preparation ---------------------------------------------------------------------------------------------------
// One thread = one (k-mer, 16-byte K chunk).  Row i of an operand lives at ((i / 8) * 4 + kc) * 128 + (i % 8) * 16: 8-row
// core matrices of one K chunk are contiguous (128 B), the four K chunks of a row group follow each other (LBO = 128 B) and
// row groups are 512 B apart (SBO), so any 8-aligned block of rows is one contiguous piece of memory.
// side[i] = (key, label as a column, label that triggers the head override as a row or -3, shift of the override)
__global__ void __launch_bounds__(256) onehot_operands_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ labels, int64_t n,
                                                              int64_t n_pad, int k, const int32_t* __restrict__ head_len, int n_labels,
                                                              uint4* __restrict__ A, uint4* __restrict__ B, int4* __restrict__ side) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = t >> 2;
    const int kc = (int)(t & 3);
    if (i >= n_pad) return;
    const uint32_t key = i < n ? (__ldg(kh + i) & lowmask32(k)) : 0u;
    uint32_t a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int base = 4 * kc + q;                                 // base 0 = the most significant 2-bit group of the hash
        if (base < k && i < n) {
            const uint32_t code = (key >> (2 * (k - 1 - base))) & 3u;
            a[q] = 1u << (8 * code);
            b[q] = 0x01010101u ^ a[q];
        } else {
            a[q] = 0; b[q] = 0;
        }
    }
    const int64_t at = ((i >> 3) * 4 + kc) * 8 + (i & 7);            // in 16-byte units
    A[at] = make_uint4(a[0], a[1], a[2], a[3]);
    B[at] = make_uint4(b[0], b[1], b[2], b[3]);
    if (kc == 0) {
        int lab = -2, eff = -3, sh = 0;
        if (i < n) {
            lab = labels ? __ldg(labels + i) : -1;
            if (labels && lab >= 0 && lab < n_labels) {
                const int hl = __ldg(head_len + lab);
                if (hl < k) { eff = lab; sh = 2 * (k - hl); }
            }
        }
        side[i] = make_int4((int)key, lab, eff, sh);
    }
}

// ---- the GEMM ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MM_THREADS, 1) hamdist_mma_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                                                   const int4* __restrict__ side, int64_t n, int64_t row0, int64_t row1,
                                                                   int64_t rb0, int64_t n_rb, int64_t n_cb, uint8_t* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_mem = smem;                                               // MM_STAGES x (A tile, B tile)
    uint8_t* out_mem = smem + MM_STAGES * MM_STAGE_BYTES;                    // 2 x staged output tile
    int2* col_side = reinterpret_cast<int2*>(out_mem + 2 * MM_OUT_BYTES);    // 2 x MM_N x (key, label)
    __shared__ uint64_t full_bar[MM_STAGES], empty_bar[MM_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's contiguous range of the (row block, column block) tiles
    const int64_t n_tiles = n_rb * n_cb;
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], MM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                          // 512 columns: two accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t t = t_lo; t < t_hi; ++t, ++it) {
                const int s = it % MM_STAGES;
                if (it >= MM_STAGES) mbar_wait(&empty_bar[s], ((it / MM_STAGES) - 1) & 1);
                const int64_t rb = rb0 + t / n_cb, cb = t % n_cb;
                uint8_t* sa = stage_mem + (size_t)s * MM_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], MM_STAGE_BYTES);
                bulk_load(sa, A + rb * MM_A_BYTES, MM_A_BYTES, &full_bar[s]);
                bulk_load(sa + MM_A_BYTES, B + cb * MM_B_BYTES, MM_B_BYTES, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_i8();
            uint32_t it = 0;
            for (int64_t t = t_lo; t < t_hi; ++t, ++it) {
                const int s = it % MM_STAGES, a = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[a], ((it >> 1) - 1) & 1);       // the epilogue has drained this accumulator
                mbar_wait(&full_bar[s], (it / MM_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(stage_mem + (size_t)s * MM_STAGE_BYTES), sb = sa + MM_A_BYTES;
                const uint32_t d = tmem_base + (uint32_t)a * MM_N;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)                                 // K = 2 x 32: the second MMA starts two K chunks further
                    mma_i8(d, smem_desc(sa + kk * 256, 128, 512), smem_desc(sb + kk * 256, 128, 512), idesc, kk);
                mma_commit(&empty_bar[s]);                                     // the stage may be refilled once the MMAs have read it
                mma_commit(&acc_full[a]);
            }
        }
    } else {
        const int e = warp - 2;                         // epilogue warp 0..7
        const int q = warp & 3;                         // TMEM lane quarter this warp may access (warp id % 4)
        const int half = e >> 2;                        // which 128 of the 256 columns
        const int et = threadIdx.x - 64;                // 0..255 among the epilogue threads
        const int row_in_tile = 32 * q + lane;
        uint32_t it = 0;
        for (int64_t t = t_lo; t < t_hi; ++t, ++it) {
            const int a = it & 1;
            const int64_t rb = rb0 + t / n_cb, cb = t % n_cb;
            const int64_t gi = rb * MM_M + row_in_tile;                        // global row (k-mer index)
            const int4 mine = __ldg(side + gi);
            // column keys / labels of this tile (one column per epilogue thread); visible after the barrier below, which every
            // thread reaches only after its stores of two tiles ago
            {
                const int4 cs = __ldg(side + cb * MM_N + et);
                col_side[a * MM_N + et] = make_int2(cs.x, cs.y);
            }
            mbar_wait(&acc_full[a], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");                     // col_side[a] is complete
            uint8_t* stg = out_mem + (size_t)a * MM_OUT_BYTES + (size_t)row_in_tile * MM_N;
            const int2* cs = col_side + a * MM_N;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32];
                const int col = half * 128 + c0;
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * MM_N + col), v);
                if (mine.z >= 0) {                                             // a row of a short consensus: same-label pairs use the head
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int2 c = cs[col + j];
                        if (c.y == mine.z) v[j] = nz_groups32(((uint32_t)mine.x ^ (uint32_t)c.x) >> mine.w, 0xFFFFFFFFu);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {                                  // 16 columns = one 16-byte chunk of the row
                    uint32_t w[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const uint32_t* p = v + 16 * g + 4 * x;
                        w[x] = __byte_perm(__byte_perm(p[0], p[1], 0x0040), __byte_perm(p[2], p[3], 0x0040), 0x5410);
                    }
                    const int chunk = (col >> 4) + g;                          // 16-byte chunk of the 256-byte row, XOR-swizzled by row
                    *reinterpret_cast<uint4*>(stg + 16 * (chunk ^ (row_in_tile & 15))) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);                         // this warp's part of the accumulator has been read
            asm volatile("bar.sync 1, 256;" ::: "memory");                     // the staged tile is complete
            const uint8_t* tile_stg = out_mem + (size_t)a * MM_OUT_BYTES;
            const bool vec_ok = (n & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
            for (int j = 0; j < MM_OUT_BYTES / 16 / 256; ++j) {                // 8 chunks per thread: 16 lanes cover one 256-byte row segment
                const int c = et + 256 * j;
                const int r = c >> 4, ch = c & 15;
                const int64_t grow = rb * MM_M + r, gcol = cb * MM_N + 16 * ch;
                if (grow < row0 || grow >= row1 || gcol >= n) continue;
                const uint4 val = *reinterpret_cast<const uint4*>(tile_stg + (size_t)r * MM_N + 16 * (ch ^ (r & 15)));
                uint8_t* dst = out + (grow - row0) * n + gcol;
                if (vec_ok) {
                    __stcs(reinterpret_cast<uint4*>(dst), val);
                } else {
                    const uint32_t ws[4] = {val.x, val.y, val.z, val.w};
                    for (int b = 0; b < 16 && gcol + b < n; ++b) dst[b] = (uint8_t)(ws[b >> 2] >> (8 * (b & 3)));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int64_t pad_rows(int64_t n) { return (n + MM_N - 1) / MM_N * MM_N; }

}  // namespace

extern "C" {

int64_t kmap_hamdist_mma_scratch_bytes(int64_t n) {
    if (n < 0) return 0;
    const int64_t np = pad_rows(n);
    return 2 * np * MM_KB + np * 16 + 256;
}

int kmap_hamdist_matrix_onehot_mma(const uint32_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels,
                                   int64_t row0, int64_t row1, uint8_t* out, void* scratch, int64_t scratch_bytes, void* stream) {
    KMAP_REQUIRE(n >= 0 && k >= 1 && k <= 16, "k out of range for this hash width");
    KMAP_REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= n, "bad row range");
    KMAP_REQUIRE(n_labels == 0 || (labels && head_len), "labels/head_len missing");
    if (n == 0 || row0 == row1) return KMAP_OK;
    KMAP_REQUIRE(kh && out && scratch, "null pointer");
    KMAP_REQUIRE(scratch_bytes >= kmap_hamdist_mma_scratch_bytes(n) && ((uintptr_t)scratch & 255) == 0, "scratch too small or not 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int64_t np = pad_rows(n);
    uint8_t* A = reinterpret_cast<uint8_t*>(scratch);
    uint8_t* B = A + np * MM_KB;
    int4* side = reinterpret_cast<int4*>(B + np * MM_KB);
    onehot_operands_kernel<<<grid_for(4 * np, 256), 256, 0, s>>>(kh, n_labels ? labels : nullptr, n, np, k, head_len, n_labels,
                                                                 reinterpret_cast<uint4*>(A), reinterpret_cast<uint4*>(B), side);
    const int64_t rb0 = row0 / MM_M, rb1 = (row1 + MM_M - 1) / MM_M;
    const int64_t n_rb = rb1 - rb0, n_cb = np / MM_N;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = n_rb * n_cb;
    const unsigned int grid = (unsigned int)(n_tiles < sms ? n_tiles : sms);
    cudaFuncSetAttribute(hamdist_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    hamdist_mma_kernel<<<grid, MM_THREADS, MM_SMEM, s>>>(A, B, side, n, row0, row1, rb0, n_rb, n_cb, out);
    return kmap_check_launch("hamdist_matrix_onehot_mma");
}

}  // extern "C"
