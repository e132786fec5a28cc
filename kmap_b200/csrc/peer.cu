// The exchange step of the sharded counting path over NVLink PEER MEMORY, in narrow form.
//
// What is exchanged are count tables (uint32 cells, sums modulo 2^32; the reference has no multi-process path -- reads are
// its independent units, kmer_count.py:755-759 -- so the merged table only has to equal the table of the whole input).  A
// rank's level-14 table of a 1e8-read input sharded over 8 GPUs holds about 4 windows per cell and the merged one about 32:
// almost every cell fits a signed byte, while a ring all-reduce ships 4 bytes per cell twice.  So the tables travel as ONE
// BYTE per cell, and a cell that does not fit (a planted motif, a repetitive input, a -1 correction that wrapped far) is
// marked with the escape byte 0x80 and fetched as the full word from the owner's table, which lives in peer-mapped memory
// too: lossless for any input, no side lists, no capacities, and a table full of large counts degrades to reading words.
//
// Every rank allocates one REGION (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the other ranks of the node):
//     [ table area: uint32 cells ][ staging area: 1 byte per cell ][ result area: 1 byte per cell ][ flags ]
// The tables that are to be merged this way are views of the table area (api.TableAllReduce.alloc_tables).  One exchange of
// the cells [c0, c0 + n) of the area, rank r owning the r-th of `world` equal key ranges of them -- every byte that crosses
// NVLink is a posted STORE (loads over the links reached half the rate, see below):
//   1. narrow_push_kernel   own table cells -> bytes, stored straight into the OWNER's staging area (slot = this rank)
//      barrier
//   2. reduce_staged_kernel owned range: the `world` staged slots are summed (local loads), the sums written to the own
//                           table as words and -- narrowed again -- stored into the result area of EVERY rank
//      barrier
//   3. widen_kernel         the other ranges: own result bytes -> own table words (escapes: the owner's table)   (local)
//      barrier              (nobody zeroes or rewrites a table a peer may still be reading)
// Per GPU and direction that is (world-1)/world bytes per cell twice instead of 4 (world-1)/world twice for the ring.  The
// barriers are flag words in peer memory (one store per peer, one spin on local memory), bounded: a rank that never shows
// up raises the region's status word, which the host checks (kmap_comm_peer_status), instead of hanging the GPU.
// KMAP_PEER_MODE=pull keeps the first version for comparison: step 1 narrows locally and step 2 LOADS the bytes of all
// ranks over NVLink (measured on 8 B200s, 2^28 cells: 1.29 ms against 2.65 ms for the NCCL all-reduce; its loads + stores ran
// at 324 GB/s per direction where NCCL's stores reach 709).
// What makes the buffers safe: the staging slots and the result bytes of a range are written only between the barriers that
// separate them from their readers; the words fetched for escapes lie in ranges their holders do not write in that step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace {

constexpr int PEER_MAX = 8;                 // one NVSwitch node
constexpr int64_t PEER_FLAG_BYTES = 4096;
constexpr unsigned long long PEER_TIMEOUT_NS = 120ull * 1000000000ull;

struct PeerPtrs { uint8_t* base[PEER_MAX]; };

struct PeerEx {
    int rank = 0, world = 1;
    PeerPtrs p;
    void* opened[PEER_MAX];                 // what cudaIpcCloseMemHandle wants back (NULL for the own region)
    int64_t table_cells = 0;                // capacity of the table area
    uint32_t epoch = 0;                     // barriers issued so far (the same sequence on every rank)
    int64_t stage_off() const { return table_cells * 4; }
    int64_t result_off() const { return table_cells * 5; }
    int64_t flags_off() const { return table_cells * 6; }
};

__device__ __forceinline__ uint32_t narrow4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    auto one = [](uint32_t v) -> uint32_t { return (v + 127u <= 254u) ? (v & 0xFFu) : 0x80u; };
    return one(a) | (one(b) << 8) | (one(c) << 16) | (one(d) << 24);
}
__device__ __forceinline__ uint32_t narrow16(const uint4 v) { return narrow4(v.x, v.y, v.z, v.w); }

// step 1: one thread = 16 cells (four 128-bit loads, one 128-bit store)
__global__ void __launch_bounds__(256) narrow_kernel(const uint4* __restrict__ table, uint4* __restrict__ narrow, int64_t n16) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < n16; g += stride) {
        const uint4 a = table[4 * g], b = table[4 * g + 1], c = table[4 * g + 2], d = table[4 * g + 3];
        narrow[g] = make_uint4(narrow16(a), narrow16(b), narrow16(c), narrow16(d));
    }
}

__device__ __forceinline__ uint4 ld_peer16(const uint4* p) {        // never from a stale line of this SM
    uint4 v;
    asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_peer4(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// non-zero iff some byte of x is the escape 0x80 (zero-byte test on x ^ 0x80808080)
__device__ __forceinline__ uint32_t has_escape(uint32_t x) {
    const uint32_t y = x ^ 0x80808080u;
    return (y - 0x01010101u) & ~y & 0x80808080u;
}

// adds the four cells packed in `x` (bytes of rank `src`'s narrow area, first cell `cell`) to acc[0..3]
__device__ __forceinline__ void add_bytes(uint32_t x, uint32_t acc[4], const uint8_t* src_base, int64_t cell) {
    if (has_escape(x)) {                                    // some byte is the escape: its word comes from the owner's table
        const uint32_t* wide = reinterpret_cast<const uint32_t*>(src_base) + cell;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = (x >> (8 * j)) & 0xFFu;
            acc[j] += b == 0x80u ? ld_peer4(wide + j) : (uint32_t)(int32_t)(int8_t)b;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += (uint32_t)((int32_t)(x << (24 - 8 * j)) >> 24);
    }
}

// step 2 of the PULL form: the owned range [cell_lo, cell_lo + 16 n16) of the table area, the bytes of every rank loaded over
// NVLink from its staging area (natural cell order).  push = 0: reduce-scatter only (scattered merge).
template <int W>
__global__ void __launch_bounds__(256) reduce_push_kernel(PeerPtrs p, uint8_t* own_base, int64_t narrow_off, int64_t result_off, int64_t cell_lo, int64_t n16, int push) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < n16; g += stride) {
        const int64_t cell = cell_lo + 16 * g;
        uint4 in[W];
#pragma unroll
        for (int r = 0; r < W; ++r) in[r] = ld_peer16(reinterpret_cast<const uint4*>(p.base[r] + narrow_off + cell));
        uint32_t acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0;
#pragma unroll
        for (int r = 0; r < W; ++r) {
            add_bytes(in[r].x, acc, p.base[r], cell);
            add_bytes(in[r].y, acc + 4, p.base[r], cell + 4);
            add_bytes(in[r].z, acc + 8, p.base[r], cell + 8);
            add_bytes(in[r].w, acc + 12, p.base[r], cell + 12);
        }
        uint4* own = reinterpret_cast<uint4*>(own_base) + cell / 4;
        own[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        own[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
        own[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
        own[3] = make_uint4(acc[12], acc[13], acc[14], acc[15]);
        if (push) {
            const uint4 e = make_uint4(narrow4(acc[0], acc[1], acc[2], acc[3]), narrow4(acc[4], acc[5], acc[6], acc[7]),
                                       narrow4(acc[8], acc[9], acc[10], acc[11]), narrow4(acc[12], acc[13], acc[14], acc[15]));
#pragma unroll
            for (int r = 0; r < W; ++r) *reinterpret_cast<uint4*>(p.base[r] + result_off + cell) = e;
        }
    }
    __threadfence_system();                 // the stores into the peers' memory are performed before this grid is done
}

// step 1: one thread = 16 cells of the own table (four 128-bit loads) -> 16 bytes stored into the staging area of the rank
// that owns them, slot `rank` (own16 = groups per owner; the slots of an exchange fill exactly its n bytes of the area)
__global__ void __launch_bounds__(256) narrow_push_kernel(PeerPtrs p, const uint4* __restrict__ table, int64_t stage_off, int64_t cell0,
                                                          int64_t groups, int64_t own16, int rank) {
    __shared__ uint8_t* dst_base[PEER_MAX];
    if (threadIdx.x < PEER_MAX) dst_base[threadIdx.x] = p.base[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += stride) {
        const uint4 a = table[4 * g], b = table[4 * g + 1], c = table[4 * g + 2], d = table[4 * g + 3];
        const int64_t o = g / own16, i = g - o * own16;
        *reinterpret_cast<uint4*>(dst_base[o] + stage_off + cell0 + (rank * own16 + i) * 16) =
            make_uint4(narrow16(a), narrow16(b), narrow16(c), narrow16(d));
    }
    __threadfence_system();
}

// step 2: the owned range: sum of the staged slots (local), words into the own table, narrowed sums into every result area
template <int W>
__global__ void __launch_bounds__(256) reduce_staged_kernel(PeerPtrs p, uint8_t* own_base, int rank, int64_t stage_off, int64_t result_off,
                                                            int64_t cell0, int64_t own16, int push) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < own16; i += stride) {
        const int64_t cell = cell0 + (rank * own16 + i) * 16;
        uint4 in[W];
#pragma unroll
        for (int r = 0; r < W; ++r) in[r] = *reinterpret_cast<const uint4*>(own_base + stage_off + cell0 + (r * own16 + i) * 16);
        uint32_t acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0;
#pragma unroll
        for (int r = 0; r < W; ++r) {
            add_bytes(in[r].x, acc, p.base[r], cell);
            add_bytes(in[r].y, acc + 4, p.base[r], cell + 4);
            add_bytes(in[r].z, acc + 8, p.base[r], cell + 8);
            add_bytes(in[r].w, acc + 12, p.base[r], cell + 12);
        }
        uint4* own = reinterpret_cast<uint4*>(own_base) + cell / 4;
        own[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        own[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
        own[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
        own[3] = make_uint4(acc[12], acc[13], acc[14], acc[15]);
        if (push) {
            const uint4 e = make_uint4(narrow4(acc[0], acc[1], acc[2], acc[3]), narrow4(acc[4], acc[5], acc[6], acc[7]),
                                       narrow4(acc[8], acc[9], acc[10], acc[11]), narrow4(acc[12], acc[13], acc[14], acc[15]));
#pragma unroll
            for (int r = 0; r < W; ++r) *reinterpret_cast<uint4*>(p.base[r] + result_off + cell) = e;
        }
    }
    __threadfence_system();
}

// step 3: every range but the owned one: own result bytes (the owners' sums) -> own table words
__global__ void __launch_bounds__(256) widen_kernel(PeerPtrs p, uint8_t* own_base, int world, int64_t narrow_off, int64_t cell0, int64_t groups,
                                                    int64_t own_lo16, int64_t own_n16) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    const int64_t todo = groups - own_n16;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < todo; i += stride) {
        const int64_t g = i < own_lo16 ? i : i + own_n16;
        const int64_t cell = cell0 + 16 * g;
        const uint4 x = *reinterpret_cast<const uint4*>(own_base + narrow_off + cell);
        uint32_t acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0;
        const uint32_t any = has_escape(x.x) | has_escape(x.y) | has_escape(x.z) | has_escape(x.w);
        const uint8_t* owner_base = own_base;
        if (any) {                                          // (rare) which rank owns group g: the ranges are groups * r / world
            int o = 0;
            while (o + 1 < world && groups * (o + 1) / world <= g) ++o;
            owner_base = p.base[o];
        }
        add_bytes(x.x, acc, owner_base, cell);
        add_bytes(x.y, acc + 4, owner_base, cell + 4);
        add_bytes(x.z, acc + 8, owner_base, cell + 8);
        add_bytes(x.w, acc + 12, owner_base, cell + 12);
        uint4* own = reinterpret_cast<uint4*>(own_base) + cell / 4;
        own[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        own[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
        own[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
        own[3] = make_uint4(acc[12], acc[13], acc[14], acc[15]);
    }
}

// barrier over the ranks: lane t tells rank t "rank `rank` has reached barrier `epoch`" and waits for rank t's word
__global__ void peer_barrier_kernel(PeerPtrs p, int rank, int world, int64_t flags_off, uint32_t epoch) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    uint32_t* theirs = reinterpret_cast<uint32_t*>(p.base[t] + flags_off) + 16 * rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.base[rank] + flags_off) + 16 * t;
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > PEER_TIMEOUT_NS) {                   // a peer is gone: say so instead of spinning for ever
            reinterpret_cast<uint32_t*>(p.base[rank] + flags_off)[16 * PEER_MAX] = 1u + (uint32_t)t;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

void barrier(PeerEx* x, cudaStream_t s) {
    ++x->epoch;
    peer_barrier_kernel<<<1, 32, 0, s>>>(x->p, x->rank, x->world, x->flags_off(), x->epoch);
}

unsigned int grid_cap(int64_t threads, int per_sm) {
    int64_t g = (threads + 255) / 256;
    if (g > 148 * per_sm) g = 148 * per_sm;
    return (unsigned int)(g < 1 ? 1 : g);
}

}  // namespace

// ---- used by comm.cu -----------------------------------------------------------------------------------------------
void* kmap_peer_new(int rank, int world, void* my_region, int64_t table_cells, const uint8_t* handles) {
    if (world < 2 || world > PEER_MAX || rank < 0 || rank >= world || !my_region || table_cells <= 0 || table_cells % 16) {
        kmap_set_error("comm_attach_peers: 2..8 ranks, a region and a multiple of 16 cells");
        return nullptr;
    }
    PeerEx* x = new PeerEx();
    x->rank = rank; x->world = world; x->table_cells = table_cells;
    for (int r = 0; r < PEER_MAX; ++r) { x->p.base[r] = nullptr; x->opened[r] = nullptr; }
    for (int r = 0; r < world; ++r) {
        if (r == rank) { x->p.base[r] = static_cast<uint8_t*>(my_region); continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            kmap_set_error("comm_attach_peers: cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (x->opened[q]) cudaIpcCloseMemHandle(x->opened[q]);
            delete x;
            return nullptr;
        }
        x->opened[r] = ptr;
        x->p.base[r] = static_cast<uint8_t*>(ptr);
    }
    return x;
}

void kmap_peer_delete(void* peer) {
    PeerEx* x = static_cast<PeerEx*>(peer);
    if (!x) return;
    for (int r = 0; r < x->world; ++r) if (x->opened[r]) cudaIpcCloseMemHandle(x->opened[r]);
    delete x;
}

// does [buf, buf + n) lie in the table area of the region, in a shape the exchange takes?
bool kmap_peer_covers(const void* peer, const uint32_t* buf, int64_t n, int scatter) {
    const PeerEx* x = static_cast<const PeerEx*>(peer);
    if (!x || n <= 0) return false;
    const uint32_t* lo = reinterpret_cast<const uint32_t*>(x->p.base[x->rank]);
    if (buf < lo || buf + n > lo + x->table_cells) return false;
    (void)scatter;
    return (buf - lo) % 16 == 0 && n % (16 * (int64_t)x->world) == 0;    // whole groups of 16 cells, the same number for every owner
}

#define PEER_FOR_WORLD(KERNEL, ...)                                        \
    switch (x->world) {                                                   \
        case 2: KERNEL<2><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        case 3: KERNEL<3><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        case 4: KERNEL<4><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        case 5: KERNEL<5><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        case 6: KERNEL<6><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        case 7: KERNEL<7><<<g2, 256, 0, s>>>(__VA_ARGS__); break;          \
        default: KERNEL<8><<<g2, 256, 0, s>>>(__VA_ARGS__); break;         \
    }

int kmap_peer_exchange(void* peer, uint32_t* buf, int64_t n, int scatter, cudaStream_t s) {
    PeerEx* x = static_cast<PeerEx*>(peer);
    const int64_t cell0 = buf - reinterpret_cast<const uint32_t*>(x->p.base[x->rank]);
    const int64_t groups = n / 16;
    const int64_t own16 = groups / x->world;                             // groups per owner (kmap_peer_covers: no remainder)
    const int64_t own_lo16 = own16 * x->rank;
    uint8_t* mine = x->p.base[x->rank];
    static const bool trace = getenv("KMAP_PEER_TRACE") != nullptr;       // (debugging aid: per-step times; synchronises the stream)
    static const bool pull = [] { const char* e = getenv("KMAP_PEER_MODE"); return e && !strcmp(e, "pull"); }();
    static const int per_sm = [] { const char* e = getenv("KMAP_PEER_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : 8; }();
    cudaEvent_t ev[7];
    int ne = 0;
    auto mark = [&]() { if (trace) { cudaEventCreate(&ev[ne]); cudaEventRecord(ev[ne], s); ++ne; } };
    mark();
    if (pull)
        narrow_kernel<<<grid_cap(groups, per_sm), 256, 0, s>>>(reinterpret_cast<const uint4*>(mine) + cell0 / 4,
                                                               reinterpret_cast<uint4*>(mine + x->stage_off() + cell0), groups);
    else
        narrow_push_kernel<<<grid_cap(groups, per_sm), 256, 0, s>>>(x->p, reinterpret_cast<const uint4*>(mine) + cell0 / 4, x->stage_off(), cell0,
                                                                    groups, own16, x->rank);
    mark();
    barrier(x, s);
    mark();
    const unsigned int g2 = grid_cap(own16, per_sm);
    const int push = scatter ? 0 : 1;
    if (pull) {
        PEER_FOR_WORLD(reduce_push_kernel, x->p, mine, x->stage_off(), x->result_off(), cell0 + 16 * own_lo16, own16, push)
    } else {
        PEER_FOR_WORLD(reduce_staged_kernel, x->p, mine, x->rank, x->stage_off(), x->result_off(), cell0, own16, push)
    }
    mark();
    barrier(x, s);
    mark();
    if (!scatter) {
        widen_kernel<<<grid_cap(groups - own16, per_sm), 256, 0, s>>>(x->p, mine, x->world, x->result_off(), cell0, groups, own_lo16, own16);
        mark();
        barrier(x, s);
        mark();
    }
    if (trace) {
        cudaStreamSynchronize(s);
        static const char* names[] = {"narrow", "barrier", "reduce", "barrier", "widen", "barrier"};
        char line[512];
        int o = snprintf(line, sizeof line, "[peer trace] rank %d, %lld cells (%s):", x->rank, (long long)n, pull ? "pull" : "push");
        for (int i = 0; i + 1 < ne; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); o += snprintf(line + o, sizeof line - o, " %s %.3f", names[i], ms); }
        fprintf(stderr, "%s ms\n", line);
        for (int i = 0; i < ne; ++i) cudaEventDestroy(ev[i]);
    }
    return kmap_check_launch("table_allreduce(peer memory)");
}

int kmap_peer_status_of(void* peer, int* status_out, cudaStream_t s) {
    PeerEx* x = static_cast<PeerEx*>(peer);
    uint32_t v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, x->p.base[x->rank] + x->flags_off() + 4 * 16 * PEER_MAX, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { kmap_set_error("comm_peer_status: %s", cudaGetErrorString(e)); return (int)e; }
    *status_out = (int)v;
    return KMAP_OK;
}

extern "C" {

int64_t kmap_peer_region_bytes(int64_t table_cells) {
    if (table_cells <= 0) return 0;
    table_cells = (table_cells + 15) / 16 * 16;
    return table_cells * 6 + PEER_FLAG_BYTES;
}

int kmap_peer_region_alloc(int64_t table_cells, void** region_out, uint8_t* handle_out) {
    KMAP_REQUIRE(table_cells > 0 && region_out && handle_out, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle is documented as 64 bytes");
    const int64_t bytes = kmap_peer_region_bytes(table_cells);
    void* ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(static_cast<uint8_t*>(ptr) + bytes - PEER_FLAG_BYTES, 0, PEER_FLAG_BYTES);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        kmap_set_error("peer_region_alloc: %s", cudaGetErrorString(e));
        cudaGetLastError();
        if (ptr) cudaFree(ptr);
        return (int)e;
    }
    memcpy(handle_out, &h, sizeof(h));
    *region_out = ptr;
    return KMAP_OK;
}

int kmap_peer_region_free(void* region) {
    if (!region) return KMAP_OK;
    const cudaError_t e = cudaFree(region);
    if (e != cudaSuccess) { kmap_set_error("peer_region_free: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return KMAP_OK;
}

}  // extern "C"
