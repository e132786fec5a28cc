// The exchange step of the sharded counting path: integer all-reduce of dense count tables over the ranks of one node
// (`ncclAllReduce(sum, ncclUint32, 4^k)` over NVLink / NVSwitch; the reference has no multi-process
// path -- reads are its independent units, kmer_count.py:755-759).  NCCL is bound at run time (dlopen of the libnccl the
// process already carries -- PyTorch's -- else the system one), so the library loads and exports every symbol on a box
// without NCCL; the calls then fail loudly.  The communicator is created from a 128-byte unique id that the host plumbing
// (torch.distributed) broadcasts; it is owned by the caller and passed to every call: the library keeps no communicator.
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include "common.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;      // (optional)
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

// (the symbol table of the NCCL library: process-wide by nature, immutable once resolved)
const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.handle) break; }       // already in the process
        if (!api.handle) for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(api.handle, "ncclCommInitRankConfig"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(dlsym(api.handle, "ncclCommCount"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(dlsym(api.handle, "ncclReduceScatter"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.ReduceScatter && api.GetErrorString;
    });
    return api;
}

// what kmap_comm_init hands out: the NCCL communicator and, once kmap_comm_attach_peers has been called, the peer-memory
// exchange of peer.cu that takes over the tables living in its region
struct KmapComm { ncclComm_t nccl; void* peer; };
inline KmapComm* as_comm(void* c) { return static_cast<KmapComm*>(c); }

int nccl_fail(const char* what, ncclResult_t r) {
    kmap_set_error("%s: %s", what, nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "NCCL error");
    return KMAP_ERR_COMM;
}

}  // namespace

// implemented in peer.cu
void* kmap_peer_new(int rank, int world, void* my_region, int64_t table_cells, const uint8_t* handles);
void kmap_peer_delete(void* peer);
bool kmap_peer_covers(const void* peer, const uint32_t* buf, int64_t n, int scatter);
int kmap_peer_exchange(void* peer, uint32_t* buf, int64_t n, int scatter, cudaStream_t s);
int kmap_peer_status_of(void* peer, int* status_out, cudaStream_t s);

// used by count_all.cu / partition.cu for the merges they overlap with counting
int kmap_allreduce_u32_on(uint32_t* buf, int64_t n, void* comm, cudaStream_t s) {
    if (n == 0) return KMAP_OK;
    if (as_comm(comm)->peer && kmap_peer_covers(as_comm(comm)->peer, buf, n, 0)) return kmap_peer_exchange(as_comm(comm)->peer, buf, n, 0, s);
    const NcclApi& api = nccl_api();
    if (!api.ok) { kmap_set_error("table_allreduce: no NCCL library in this process"); return KMAP_ERR_COMM; }
    const ncclResult_t r = api.AllReduce(buf, buf, (size_t)n, ncclUint32, ncclSum, as_comm(comm)->nccl, s);
    return r == ncclSuccess ? KMAP_OK : nccl_fail("table_allreduce", r);
}

int kmap_comm_ctas(int world) {
    static const int v = [] { const char* e = getenv("KMAP_COMM_CTAS"); const int x = e ? atoi(e) : 0; return x > 0 && x <= 128 ? x : 0; }();
    return v ? v : (world >= 4 ? 32 : KMAP_COMM_CTAS);
}
int kmap_merge_chunks(int world) {
    static const int v = [] { const char* e = getenv("KMAP_MERGE_CHUNKS"); const int x = e ? atoi(e) : 0; return x > 0 && x <= 8 ? x : 0; }();
    return v ? v : (world >= 4 ? 2 : 4);
}
int kmap_comm_world(void* comm) {
    const NcclApi& api = nccl_api();
    int n = 1;
    if (!api.ok || !api.CommCount || !comm || api.CommCount(as_comm(comm)->nccl, &n) != ncclSuccess) return 1;
    return n;
}

bool kmap_merge_on_peer_memory(const uint32_t* buf, int64_t n, const KmapMerge* m) {
    return m && m->comm && n > 0 && as_comm(m->comm)->peer && kmap_peer_covers(as_comm(m->comm)->peer, buf, n, m->scatter);
}

int kmap_merge_table_on(uint32_t* buf, int64_t n, const KmapMerge* m) {
    if (n == 0) return KMAP_OK;
    if (!m->scatter || m->world <= 1 || n % m->world != 0) return kmap_allreduce_u32_on(buf, n, m->comm, m->stream);
    if (as_comm(m->comm)->peer && kmap_peer_covers(as_comm(m->comm)->peer, buf, n, 1)) return kmap_peer_exchange(as_comm(m->comm)->peer, buf, n, 1, m->stream);
    const NcclApi& api = nccl_api();
    if (!api.ok) { kmap_set_error("table_reduce_scatter: no NCCL library in this process"); return KMAP_ERR_COMM; }
    const int64_t block = n / m->world;         // in place: the receive buffer is this rank's block of the send buffer
    const ncclResult_t r = api.ReduceScatter(buf, buf + (size_t)m->rank * block, (size_t)block, ncclUint32, ncclSum,
                                             as_comm(m->comm)->nccl, m->stream);
    return r == ncclSuccess ? KMAP_OK : nccl_fail("table_reduce_scatter", r);
}

extern "C" {

int kmap_comm_available(void) { return nccl_api().ok ? 1 : 0; }

int kmap_comm_unique_id(uint8_t* id_out) {
    KMAP_REQUIRE(id_out, "null pointer");
    const NcclApi& api = nccl_api();
    if (!api.ok) { kmap_set_error("comm_unique_id: no NCCL library in this process"); return KMAP_ERR_COMM; }
    ncclUniqueId id;
    const ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail("comm_unique_id", r);
    memcpy(id_out, id.internal, NCCL_UNIQUE_ID_BYTES);
    return KMAP_OK;
}

int kmap_comm_init(const uint8_t* id_in, int rank, int world, void** comm_out) {
    KMAP_REQUIRE(id_in && comm_out && world >= 1 && rank >= 0 && rank < world, "bad argument");
    const NcclApi& api = nccl_api();
    if (!api.ok) { kmap_set_error("comm_init: no NCCL library in this process"); return KMAP_ERR_COMM; }
    ncclUniqueId id;
    memcpy(id.internal, id_in, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    // The exchange runs BESIDE the counting kernels (count_all.cu / partition.cu leave KMAP_COMM_CTAS SMs free for it), so
    // the collective is capped at that many CTAs: one that wants more would wait for whole waves of the counting kernel.
    ncclResult_t r;
    if (api.CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.maxCTAs = kmap_comm_ctas(world);
        r = api.CommInitRankConfig(&comm, world, id, rank, &cfg);              // binds to the current CUDA device
    } else {
        r = api.CommInitRank(&comm, world, id, rank);
    }
    if (r != ncclSuccess) return nccl_fail("comm_init", r);
    *comm_out = new KmapComm{comm, nullptr};
    return KMAP_OK;
}

int kmap_comm_destroy(void* comm) {
    if (!comm) return KMAP_OK;
    KmapComm* c = as_comm(comm);
    kmap_peer_delete(c->peer);
    const NcclApi& api = nccl_api();
    const ncclResult_t r = api.ok ? api.CommDestroy(c->nccl) : ncclSuccess;
    delete c;
    return r == ncclSuccess ? KMAP_OK : nccl_fail("comm_destroy", r);
}

int kmap_comm_attach_peers(void* comm, int rank, int world, void* my_region, int64_t table_cells, const uint8_t* handles_host) {
    KMAP_REQUIRE(comm && my_region && handles_host, "null pointer");
    KmapComm* c = as_comm(comm);
    KMAP_REQUIRE(!c->peer, "peers are attached already");
    KMAP_REQUIRE(world == kmap_comm_world(comm), "world differs from the communicator's");
    c->peer = kmap_peer_new(rank, world, my_region, table_cells, handles_host);
    return c->peer ? KMAP_OK : KMAP_ERR_COMM;
}

int kmap_comm_detach_peers(void* comm) {
    if (!comm) return KMAP_OK;
    kmap_peer_delete(as_comm(comm)->peer);
    as_comm(comm)->peer = nullptr;
    return KMAP_OK;
}

int kmap_comm_peer_status(void* comm, int* status_out_host, void* stream) {
    KMAP_REQUIRE(comm && status_out_host, "null pointer");
    *status_out_host = 0;
    if (!as_comm(comm)->peer) return KMAP_OK;
    const int rc = kmap_peer_status_of(as_comm(comm)->peer, status_out_host, as_stream(stream));
    if (rc == KMAP_OK && *status_out_host) {
        kmap_set_error("peer-memory exchange: rank %d never reached a barrier (waited 120 s)", *status_out_host - 1);
        return KMAP_ERR_COMM;
    }
    return rc;
}

int kmap_table_allreduce(uint32_t* table, int64_t n_cells, void* comm, void* stream) {
    KMAP_REQUIRE(n_cells >= 0 && (table || n_cells == 0) && comm, "bad argument");
    return kmap_allreduce_u32_on(table, n_cells, comm, as_stream(stream));
}

int kmap_table_reduce_scatter(uint32_t* table, int64_t n_cells, int rank, int world, void* comm, void* stream) {
    KMAP_REQUIRE(n_cells >= 0 && (table || n_cells == 0) && comm && world >= 1 && rank >= 0 && rank < world, "bad argument");
    KMAP_REQUIRE(n_cells % world == 0, "the number of cells must be a multiple of the number of ranks");
    const KmapMerge m = {comm, as_stream(stream), 1, rank, world};
    return kmap_merge_table_on(table, n_cells, &m);
}

}  // extern "C"
