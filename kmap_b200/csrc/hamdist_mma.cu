// cal_samp_kmer_hamdist_mat (motif_discovery.py:759-808) as an int8 GEMM on the 5th-generation tensor cores: the comparator
// the distance-matrix kernel of hamdist.cu (XOR + popcount) is benchmarked against (BASELINE.json north_star, config 5).
//
// Formulation.  A k-mer becomes a row of K = 64 int8 (4 per base, 16 bases; bases >= k are zero):
//     A[i][4b + c] = [base b of k-mer i has code c]          (one-hot)
//     B[j][4b + c] = [base b of k-mer j has code != c]       (its complement over the k bases)
// so that (A B^T)[i][j] = sum_b [base b differs] = the Hamming distance itself: the accumulator needs no "k - matches" in
// the epilogue.  Pairs that share the label L of a consensus shorter than k compare only the first hl_L bases (md:790-800).
// That rule is bilinear too: the tail mismatches of such a pair are subtracted again by extra K columns, four per tail base
// of every such label, with A'[i] = -[base b of i has code c][label_i = L] and B'[j] = [base b of j has code != c]
// [label_j = L] -- so the tensor cores apply the override and the epilogue never looks at a label.  K = 4 (k + sum of
// the tail lengths) rounded up to 32, at most 128; beyond that the epilogue recomputes those pairs from the keys.
//
// Kernel (one persistent CTA per SM, 18 warps, hand-written tcgen05 / TMEM / mbarrier / cp.async.bulk PTX):
//   warp 0      producer: one cp.async.bulk per operand tile (the operands are pre-laid out in global memory in the
//               canonical K-major no-swizzle core-matrix order, so a 128-row A tile is 8 KB and a 256-row B tile 16 KB of
//               contiguous bytes) into a 4-stage shared-memory ring, completion on an mbarrier (expect_tx)
//   warp 1      allocates the 512 TMEM columns (two 128 x 256 int32 accumulators), then one lane issues, per output tile,
//               two tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32 each) and tcgen05.commit's the stage's
//               "empty" barrier and the accumulator's "full" barrier
//   warps 2..17 epilogue, two groups of eight warps that alternate tiles (group g drains accumulator g): tcgen05.ld
//               32x32b.x32 (warp w reads TMEM lanes 32 (w % 4) .. + 31, half of the columns), byte packing, staging in shared memory in the 128-byte-swizzle pattern, and ONE thread hands the
//               tile to the TMA engine (cp.async.bulk.tensor.2d, two 128 x 128-byte boxes; the tensor map clips the ragged
//               edges) -- no store instruction of the SM touches the output.  When the output pitch is not a multiple of 16
//               bytes (no tensor map possible) the warps store the staged tile themselves, 16 bytes per lane.
// The output (1 B per pair) is what bounds it, exactly as for the popcount kernel: the tensor pipe is idle > 95 % of the time.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace {

constexpr int MM_M = 128;                 // rows of an output tile (TMEM lanes)
constexpr int MM_N = 256;                 // columns of an output tile (TMEM columns of one accumulator)
constexpr int MM_KB_MAX = 128;            // bytes of one operand row at most (K <= 128 int8; 64 without override columns)
constexpr int MM_MAX_SLOTS = MM_KB_MAX / 4;   // K slots of four bytes (one base position each)
constexpr int MM_STAGES = 4;              // at most
constexpr int MM_EPI_WARPS = 8;                                // per epilogue group (there are two, one per accumulator)
constexpr int MM_THREADS = 32 * (2 + 2 * MM_EPI_WARPS);
constexpr int MM_OUT_BYTES = MM_M * MM_N;                     // 32 KB of output per tile
constexpr int MM_SMEM_FIXED = 2 * MM_OUT_BYTES + 2 * MM_N * 8 + 1024;      // staging, column side data, alignment slack
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

// shared -> global tensor store of one box (TMA), tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src_smem, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src_smem)), "r"(x), "r"(y) : "memory");
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- operand preparation ---------------------------------------------------------------------------------------------------
// K slot sl (four bytes) of a row: sl < k = base sl of the k-mer; k <= sl < n_slots = tail base slot_base[sl] of the override
// label slot_lab[sl].  One thread = one (k-mer, 16-byte K chunk of four slots).  Row i of an operand lives at
// ((i / 8) * nkc + kc) * 128 + (i % 8) * 16: 8-row core matrices of one K chunk are contiguous (128 B), the nkc K chunks
// of a row group follow each other (LBO = 128 B) and row groups are nkc * 128 B apart (SBO), so any 8-aligned block of
// rows is one contiguous piece of memory.
// side[i] = (key, label as a column, label that triggers the head override as a row or -3, shift of the override)
struct SlotMap { int n_slots; signed char lab[MM_MAX_SLOTS]; signed char base[MM_MAX_SLOTS]; };

__global__ void __launch_bounds__(256) onehot_operands_kernel(const uint32_t* __restrict__ kh, const int32_t* __restrict__ labels, int64_t n,
                                                              int64_t n_pad, int k, const int32_t* __restrict__ head_len, int n_labels,
                                                              int nkc, SlotMap sm, uint4* __restrict__ A, uint4* __restrict__ B,
                                                              int4* __restrict__ side) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = t / nkc;
    const int kc = (int)(t - i * nkc);
    if (i >= n_pad) return;
    const uint32_t key = i < n ? (__ldg(kh + i) & lowmask32(k)) : 0u;
    int lab = -2, eff = -3, sh = 0;
    if (i < n) {
        lab = labels ? __ldg(labels + i) : -1;
        if (labels && lab >= 0 && lab < n_labels) {
            const int hl = __ldg(head_len + lab);
            if (hl < k) { eff = lab; sh = 2 * (k - hl); }
        }
    }
    uint32_t a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sl = 4 * kc + q;
        a[q] = 0; b[q] = 0;
        if (i >= n || sl >= sm.n_slots) continue;
        const int base = sl < k ? sl : (int)sm.base[sl];             // base 0 = the most significant 2-bit group of the hash
        const uint32_t code = (key >> (2 * (k - 1 - base))) & 3u;
        if (sl < k) {
            a[q] = 1u << (8 * code);
            b[q] = 0x01010101u ^ a[q];
        } else {                                                     // tail base of override label sm.lab[sl]: -(mismatch) for its pairs
            if (eff == (int)sm.lab[sl]) a[q] = 0xFFu << (8 * code);
            if (lab == (int)sm.lab[sl]) b[q] = 0x01010101u ^ (1u << (8 * code));
        }
    }
    const int64_t at = ((i >> 3) * nkc + kc) * 8 + (i & 7);          // in 16-byte units
    A[at] = make_uint4(a[0], a[1], a[2], a[3]);
    B[at] = make_uint4(b[0], b[1], b[2], b[3]);
    if (kc == 0) side[i] = make_int4((int)key, lab, eff, sh);
}

// ---- the GEMM ---------------------------------------------------------------------------------------------------------------
template <bool TMA_STORE>
__global__ void __launch_bounds__(MM_THREADS, 1) hamdist_mma_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                                                   const int4* __restrict__ side, int64_t n, int64_t row0, int64_t row1,
                                                                   int64_t rb0, int64_t n_rb, int64_t n_cb, uint8_t* __restrict__ out,
                                                                   const __grid_constant__ CUtensorMap out_map, int nkc, int n_stages,
                                                                   int epi_override) {
    const uint32_t a_bytes = (uint32_t)MM_M * 16u * (uint32_t)nkc, b_bytes = (uint32_t)MM_N * 16u * (uint32_t)nkc;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // (the swizzled staging boxes of the TMA store need 1024-byte alignment in the shared window: align by hand, the allocation has the slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_mem = smem;                                               // n_stages x (A tile, B tile)
    uint8_t* out_mem = smem + (size_t)n_stages * stage_bytes;               // 2 x staged output tile (a multiple of 1024 B in)
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
            int64_t rb = rb0 + t_lo / n_cb, cb = t_lo % n_cb;
            for (int64_t t = t_lo; t < t_hi; ++t, ++it, ++cb) {
                const int s = (int)(it % (uint32_t)n_stages);
                if (cb == n_cb) { cb = 0; ++rb; }
                if (it >= (uint32_t)n_stages) mbar_wait(&empty_bar[s], ((it / (uint32_t)n_stages) - 1) & 1);
                uint8_t* sa = stage_mem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], stage_bytes);
                bulk_load(sa, A + rb * a_bytes, a_bytes, &full_bar[s]);
                bulk_load(sa + a_bytes, B + cb * b_bytes, b_bytes, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_i8();
            uint32_t it = 0;
            for (int64_t t = t_lo; t < t_hi; ++t, ++it) {
                const int s = (int)(it % (uint32_t)n_stages), a = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[a], ((it >> 1) - 1) & 1);       // the epilogue has drained this accumulator
                mbar_wait(&full_bar[s], (it / (uint32_t)n_stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(stage_mem + (size_t)s * stage_bytes), sb = sa + a_bytes;
                const uint32_t d = tmem_base + (uint32_t)a * MM_N;
                const uint32_t sbo = 128u * (uint32_t)nkc;
                for (int kk = 0; kk < nkc / 2; ++kk)                           // K = nkc / 2 x 32: every MMA starts two K chunks further
                    mma_i8(d, smem_desc(sa + kk * 256, 128, sbo), smem_desc(sb + kk * 256, 128, sbo), idesc, kk);
                mma_commit(&empty_bar[s]);                                     // the stage may be refilled once the MMAs have read it
                mma_commit(&acc_full[a]);
            }
        }
    } else {
        // Two epilogue groups of eight warps: group g drains accumulator g (tiles g, g + 2, ..) into staging buffer g behind
        // named barrier 1 + g, so one group packs / stores while the other waits for its TMEM loads or its accumulator.
        const int e = warp - 2;                         // epilogue warp 0..15
        const int grp = e >> 3;                         // = accumulator, staging buffer
        const int q = warp & 3;                         // TMEM lane quarter this warp may access (warp id % 4)
        const int half = (e >> 2) & 1;                  // which 128 of the 256 columns
        const int et = (e & 7) * 32 + lane;             // 0..255 inside the group
        const int row_in_tile = 32 * q + lane;
        const int a = grp;
        const uint32_t bar_id = 1u + (uint32_t)grp;
        // this row's key / override label and this thread's column of the tile: fetched one tile ahead (a dependent global load
        // in front of every tile would add its whole latency to the tile)
        int4 mine = make_int4(0, -2, -3, 0), col_cur = make_int4(0, -2, -3, 0);
        const int64_t t_first = t_lo + grp;
        int64_t rb = rb0 + t_first / n_cb, cb = t_first % n_cb;    // (walked incrementally: a 64-bit division per tile costs as much as the tile)
        if (t_first < t_hi) {
            mine = __ldg(side + rb * MM_M + row_in_tile);
            col_cur = __ldg(side + cb * MM_N + et);
        }
        uint32_t use = 0;                               // how often this group has used its accumulator
        for (int64_t t = t_first; t < t_hi; t += 2, ++use) {
            int64_t rbn = rb, cbn = cb + 2;
            while (cbn >= n_cb) { cbn -= n_cb; ++rbn; }
            int4 mine_next = mine, col_next = col_cur;
            if (t + 2 < t_hi) {
                if (rbn != rb) mine_next = __ldg(side + rbn * MM_M + row_in_tile);
                col_next = __ldg(side + cbn * MM_N + et);
            }
            // column keys / labels of this tile (one column per thread of the group); visible after the barrier below, which every
            // thread reaches only after it has finished with the group's previous tile
            col_side[a * MM_N + et] = make_int2(col_cur.x, col_cur.y);
            mbar_wait(&acc_full[a], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the TMA store of the group's previous tile has finished READING the staging buffer
            if (TMA_STORE && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");         // col_side[a] is complete, staging[a] is free
            // staging: two half tiles of 128 rows x 128 bytes in the 128-byte-swizzle pattern (16-byte chunk c of row r sits at
            // chunk c ^ (r & 7)): what the tensor map expects, and conflict-free for the 16-byte stores of 8 consecutive rows
            uint8_t* stg = out_mem + (size_t)a * MM_OUT_BYTES + (size_t)half * (MM_OUT_BYTES / 2) + (size_t)row_in_tile * 128;
            const int2* cs = col_side + a * MM_N;
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * MM_N + half * 128);
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {                                   // 32 columns at a time
                uint32_t vv[32];
                tmem_ld32(t_row + 32 * ci, vv);
                tmem_ld_wait();
                const int col = half * 128 + 32 * ci;
                if (epi_override && mine.z >= 0) {                             // (only when the override columns did not fit into K)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int2 c = cs[col + j];
                        if (c.y == mine.z) vv[j] = nz_groups32(((uint32_t)mine.x ^ (uint32_t)c.x) >> mine.w, 0xFFFFFFFFu);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {                                  // 16 columns = one 16-byte chunk of the row
                    uint32_t w[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const uint32_t* p = vv + 16 * g + 4 * x;
                        w[x] = __byte_perm(__byte_perm(p[0], p[1], 0x0040), __byte_perm(p[2], p[3], 0x0040), 0x5410);
                    }
                    const int chunk = 2 * ci + g;                              // chunk of the 128-byte half row
                    *reinterpret_cast<uint4*>(stg + 16 * (chunk ^ (row_in_tile & 7))) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);                         // this warp's part of the accumulator has been read
            if (TMA_STORE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the staged bytes, visible to the TMA engine
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");         // the staged tile is complete
            const uint8_t* tile_stg = out_mem + (size_t)a * MM_OUT_BYTES;
            if (TMA_STORE && rb * MM_M >= row0) {      // (a block that starts above row0 would need a negative box coordinate)
                if (et == 0) {
                    const int32_t y = (int32_t)(rb * MM_M - row0), x = (int32_t)(cb * MM_N);
                    tma_store_2d(&out_map, tile_stg, x, y);
                    tma_store_2d(&out_map, tile_stg + MM_OUT_BYTES / 2, x + 128, y);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
#pragma unroll
                for (int j = 0; j < MM_OUT_BYTES / 16 / 256; ++j) {            // 8 chunks per thread: 8 lanes cover a 128-byte half row
                    const int c = et + 256 * j;
                    const int hf = c >> 10, r = (c >> 3) & 127, ch = c & 7;
                    const int64_t grow = rb * MM_M + r, gcol = cb * MM_N + 128 * hf + 16 * ch;
                    if (grow < row0 || grow >= row1 || gcol >= n) continue;
                    const uint4 val = *reinterpret_cast<const uint4*>(tile_stg + (size_t)hf * (MM_OUT_BYTES / 2) + (size_t)r * 128 + 16 * (ch ^ (r & 7)));
                    uint8_t* dst = out + (grow - row0) * n + gcol;
                    if (((n | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
                        __stcs(reinterpret_cast<uint4*>(dst), val);
                    } else {
                        const uint32_t ws[4] = {val.x, val.y, val.z, val.w};
                        for (int b = 0; b < 16 && gcol + b < n; ++b) dst[b] = (uint8_t)(ws[b >> 2] >> (8 * (b & 3)));
                    }
                }
            }
            mine = mine_next;
            col_cur = col_next;
            rb = rbn; cb = cbn;
        }
        if (TMA_STORE && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store has landed
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
    return 2 * np * MM_KB_MAX + np * 16 + 256;
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
    // K slots: bases 0..k-1, then the tail bases of every label whose consensus is shorter than k (the override columns).
    // head_len is small and lives on the device: one synchronising copy.
    SlotMap sm;
    memset(&sm, 0, sizeof sm);
    sm.n_slots = k;
    bool epi_override = false;
    if (n_labels > 0) {
        if (n_labels > 4096) {
            epi_override = true;
        } else {
            int32_t hl[4096];
            cudaError_t e = cudaMemcpyAsync(hl, head_len, (size_t)n_labels * 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { kmap_set_error("hamdist_matrix_onehot_mma: %s", cudaGetErrorString(e)); return (int)e; }
            int slots = k;
            for (int L = 0; L < n_labels; ++L)
                if (hl[L] < k) slots += k - (hl[L] < 0 ? 0 : hl[L]);
            if (slots > MM_MAX_SLOTS || n_labels > 127) {
                epi_override = true;                       // does not fit into K = 128: the epilogue recomputes those pairs
            } else {
                for (int L = 0; L < n_labels; ++L)
                    for (int bq = hl[L] < 0 ? 0 : hl[L]; bq < k; ++bq) { sm.lab[sm.n_slots] = (signed char)L; sm.base[sm.n_slots] = (signed char)bq; ++sm.n_slots; }
            }
        }
    }
    const int nkc = (4 * sm.n_slots + 31) / 32 * 2;        // 16-byte K chunks per row: K = 16 * nkc bytes, a multiple of 32
    const int stage_bytes = (MM_M + MM_N) * 16 * nkc;
    int n_stages = (227 * 1024 - MM_SMEM_FIXED - 2048) / stage_bytes;
    if (n_stages > MM_STAGES) n_stages = MM_STAGES;
    const int smem_bytes = n_stages * stage_bytes + MM_SMEM_FIXED;
    uint8_t* A = reinterpret_cast<uint8_t*>(scratch);
    uint8_t* B = A + np * MM_KB_MAX;
    int4* side = reinterpret_cast<int4*>(B + np * MM_KB_MAX);
    onehot_operands_kernel<<<grid_for((int64_t)nkc * np, 256), 256, 0, s>>>(kh, n_labels ? labels : nullptr, n, np, k, head_len, n_labels, nkc, sm,
                                                                        reinterpret_cast<uint4*>(A), reinterpret_cast<uint4*>(B), side);
    const int64_t rb0 = row0 / MM_M, rb1 = (row1 + MM_M - 1) / MM_M;
    const int64_t n_rb = rb1 - rb0, n_cb = np / MM_N;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = n_rb * n_cb;
    const unsigned int grid = (unsigned int)(n_tiles < sms ? n_tiles : sms);
    // output tensor map for the TMA stores: uint8 [row1 - row0][n], boxes of 128 rows x 128 bytes, 128-byte swizzle.  Needs a
    // 16-byte aligned base and pitch; otherwise the epilogue warps store the tile themselves.
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    bool tma = (n % 16 == 0) && ((uintptr_t)out % 16 == 0) && getenv("KMAP_HAMDIST_MMA_NO_TMA") == nullptr;
    if (tma) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            tma = false;
            cudaGetLastError();
        } else {
            const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)(row1 - row0)};
            const cuuint64_t strides[1] = {(cuuint64_t)n};
            const cuuint32_t box[2] = {128, 128}, estr[2] = {1, 1};
            const CUresult r = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, out, dims, strides, box, estr,
                                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) tma = false;
        }
    }
    if (tma) {
        cudaFuncSetAttribute(hamdist_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        hamdist_mma_kernel<true><<<grid, MM_THREADS, smem_bytes, s>>>(A, B, side, n, row0, row1, rb0, n_rb, n_cb, out, map, nkc, n_stages,
                                                                     (int)epi_override);
    } else {
        cudaFuncSetAttribute(hamdist_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        hamdist_mma_kernel<false><<<grid, MM_THREADS, smem_bytes, s>>>(A, B, side, n, row0, row1, rb0, n_rb, n_cb, out, map, nkc, n_stages,
                                                                      (int)epi_override);
    }
    return kmap_check_launch("hamdist_matrix_onehot_mma");
}

}  // extern "C"
