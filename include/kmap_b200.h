/*
 * kmap_b200.h -- C ABI of libkmap_b200.so: the sm_100a CUDA implementation of KMAP's scan_motif counting path.
 *
 * The reference (chengl7-lab/kmap v0.0.7) has no FFI of its own: its hot path is a set of Python functions that
 * call Taichi kernels on NumPy arrays.  Each entry point below names the reference function / kernel it replaces
 * (paths relative to the reference's src/kmap/).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value or a negative KMAP_ERR_* otherwise;
 *     kmap_last_error() gives a message for the calling thread.  Nothing throws.
 *   - all array pointers are DEVICE pointers owned by the caller (the Python side allocates them as torch
 *     tensors) unless the parameter name ends in _host.  `stream` is a cudaStream_t passed as void*.
 *     Calls are asynchronous on that stream except where a *_host output forces a synchronisation (documented).
 *   - the library keeps no global state.
 *   - hashes are the reference's: 2 bits per base (A0 C1 G2 T3), first base most significant
 *     (taichi_core.py:9-19); uint32 for k < 16, uint64 for 16 <= k < 32 (kmer_count.py:359-365);
 *     the invalid hash is all-ones (kmer_count.py:369-370).
 *
 * Packed sequence format (produced by kmap_pack2bit, consumed by the fused kernels)
 *   packed : uint32[kmap_packed_words(n)]  16 bases per word, base p in word p>>4 at bits [31-2(p&15), 30-2(p&15)]
 *            (first base in the most significant bits, so a funnel shift yields the reference hash directly);
 *            positions holding 255 are stored as 0.
 *   valid  : uint32[kmap_valid_words(n)]   1 bit per position, position p = bit (p&31) of word p>>5; 0 for 255
 *            (N, read separators, masked bases) and for the zero padding after position n-1.
 */
#ifndef KMAP_B200_H
#define KMAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMAP_OK 0
#define KMAP_ERR_BAD_ARG (-1)
#define KMAP_ERR_CAPACITY (-2)      /* output buffer too small; the required size was written to the *_host out-param */
#define KMAP_ERR_NEED_SCRATCH (-3)  /* a read longer than the on-chip paths needs the caller-provided bitmap scratch */
#define KMAP_ERR_COMM (-4)          /* the exchange step failed (no NCCL library in the process, or an NCCL error) */

const char* kmap_last_error(void);
int kmap_version(void);

/* sizes of the packed representation for n positions (includes the zero padding the kernels rely on) */
int64_t kmap_packed_words(int64_t n);
int64_t kmap_valid_words(int64_t n);

/* ------------------------------------------------------------------------------------------------------------
 * 1:1 primitives: drop-in for the ten integer Taichi kernels (taichi_core.py:3-224)
 * ---------------------------------------------------------------------------------------------------------- */
/* kmer2hash_kernel_uint32/64 (taichi_core.py:25-61) behind comp_kmer_hash_taichi (kmer_count.py:449-473):
 * one hash per position of seq (uint8, 255 = missing); invalid if the window leaves the array or touches 255. */
int kmap_kmer2hash_u32(const uint8_t* seq, int64_t n, int k, uint32_t* hash_out, void* stream);
int kmap_kmer2hash_u64(const uint8_t* seq, int64_t n, int k, uint64_t* hash_out, void* stream);
/* cal_ham_dist_kernel_uint32/64 (taichi_core.py:75-104) behind cal_hamming_dist (kmer_count.py:494-515) */
int kmap_ham_dist_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, uint8_t* dist_out, void* stream);
int kmap_ham_dist_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, uint8_t* dist_out, void* stream);
/* cal_partial_ham_dist_head/tail kernels (taichi_core.py:108-177) behind cal_hamming_dist_head/_tail
 * (kmer_count.py:518-577) */
int kmap_ham_dist_head_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, int conseq_len, uint8_t* dist_out, void* stream);
int kmap_ham_dist_head_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, int conseq_len, uint8_t* dist_out, void* stream);
int kmap_ham_dist_tail_u32(const uint32_t* kh, int64_t n, uint32_t target, int k, int conseq_len, uint8_t* dist_out, void* stream);
int kmap_ham_dist_tail_u64(const uint64_t* kh, int64_t n, uint64_t target, int k, int conseq_len, uint8_t* dist_out, void* stream);
/* revcom_hash_kernel_uint32/64 (taichi_core.py:181-224) behind get_revcom_hash_arr (kmer_count.py:613-623) */
int kmap_revcom_u32(const uint32_t* in, int64_t n, int k, uint32_t* out, void* stream);
int kmap_revcom_u64(const uint64_t* in, int64_t n, int k, uint64_t* out, void* stream);
/* candidates for find_motif's top-k selection (np.argpartition(cnt, -top_k)[-top_k:], motif_discovery.py:657): each of the
 * n_blocks blocks writes its kk (<= 8) largest (count, index) pairs, ordered by (count descending, index ascending), to
 * out_val / out_idx [n_blocks * kk] (index -1 = fewer than kk elements seen).  The host merges the candidates. */
int kmap_topk_candidates_i32(const int32_t* cnt, int64_t n, int kk, int32_t* out_val, int64_t* out_idx, int n_blocks, void* stream);
int kmap_topk_candidates_i64(const int64_t* cnt, int64_t n, int kk, int64_t* out_val, int64_t* out_idx, int n_blocks, void* stream);
/* the labelling step of sample_disp_kmer (motif_discovery.py:849-892) for all unique k-mers at once: label[i] = index of
 * the nearest consensus (head distance over the first conseq_len[c] bases to conseq[c], or -- revcom != 0 -- tail distance
 * over the last conseq_len[c] bases to rc_conseq[c]; a consensus farther than dmax[c] counts as distance k; ties keep
 * the first), or n_conseq when the nearest is farther than dmax_k; k-mers strictly closer to the reverse complement of
 * their consensus are reverse-complemented in place. */
int kmap_label_kmers_u32(uint32_t* kh, int64_t n, int k, const uint32_t* conseq, const uint32_t* rc_conseq, const int32_t* conseq_len,
                         const int32_t* dmax, int n_conseq, int dmax_k, int revcom, int32_t* label, void* stream);
int kmap_label_kmers_u64(uint64_t* kh, int64_t n, int k, const uint64_t* conseq, const uint64_t* rc_conseq, const int32_t* conseq_len,
                         const int32_t* dmax, int n_conseq, int dmax_k, int revcom, int32_t* label, void* stream);
/* remove_duplicate_hash_per_seq (kmer_count.py:743-760): in place on a hash array; borders = int64[n_seq][2]
 * ([start, end) per read).  Keeps the FIRST occurrence of every hash inside a read. */
int kmap_dedup_hash_per_read_u32(uint32_t* hash, int64_t n, const int64_t* borders, int64_t n_seq, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fused hot path on the packed representation
 * ---------------------------------------------------------------------------------------------------------- */
/* input.bin bytes (kmer_count.py:244-263, 326-347) -> packed + valid */
int kmap_pack2bit(const uint8_t* seq, int64_t n, uint32_t* packed, uint32_t* valid, void* stream);
/* write the masking state back: seq[i] = 255 wherever valid bit i is 0 (what mask_input leaves in seq_np_arr,
 * kmer_count.py:599-602) */
int kmap_apply_valid_to_seq(uint8_t* seq, int64_t n, const uint32_t* valid, void* stream);

/* comp_kmer_hash_taichi + count_uniq_hash (kmer_count.py:449-491) fused: table[h] += 1 for every valid window.
 * table = uint32[4^k], k <= 15.  The caller zeroes the table (kmap_fill_u32) when it wants a fresh count. */
int kmap_count_dense(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table, void* stream);
/* the same with remove_duplicate_hash_per_seq (kmer_count.py:743-760) fused in: every distinct k-mer of a read
 * counts once.  borders = int64[n_seq][2] as in input.seqboarder.bin.pkl.
 * work : uint32[kmap_dedup_work_words(n_seq)] scratch; bitmap : uint32[4^k/32] scratch or NULL (only needed when
 * a read has more than KMAP_DEDUP_BLOCK_MAX windows; KMAP_ERR_NEED_SCRATCH is returned if it is NULL then).
 * Synchronises the stream once (reads two counters back). */
int kmap_count_dense_dedup(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders,
                           int64_t n_seq, int k, uint32_t* table, uint32_t* work, uint32_t* bitmap, void* stream);
int64_t kmap_dedup_work_words(int64_t n_seq);

/* The first-round counts of scan_motif for EVERY k in [kmin, kmax] at once (motif_discovery.py:262-273 calling
 * 627-636 per k): fills tables_host[k - kmin] = uint32[4^k] for each k (tables_host is a HOST array of device
 * pointers; every table is zeroed first).  Only the level-kmax table is built with one atomic per window, in
 * n_partitions key-range passes (0 = choose so that a slice stays L2 resident); each smaller table is the 4:1
 * reduction of the next one plus the per-read corrections derived in csrc/count_all.cu.  Results are identical to
 * kmax-kmin+1 calls of kmap_count_dense[_dedup].  dupmask = uint32[kmap_valid_words(n)] scratch, work/bitmap as in
 * kmap_count_dense_dedup (dedup != 0 only).  phase_events = NULL, or a HOST array of 6 cudaEvent_t (entries may be
 * NULL) recorded on the stream after zeroing / after the per-read scan / after the level-kmax count / after the
 * reductions, and [4], [5] inside the partitioned level-kmax count: after the bucket histogram (KMAP_KMAX_SORTED
 * only), after the partition pass (instrumentation for bench.py).  Synchronises the stream once when dedup != 0. */
#define KMAP_KMAX_PREFIX_PASSES 0    /* global atomics in n_partitions key-prefix passes (no scratch) */
#define KMAP_KMAX_SORTED 1           /* csrc/partition.cu, scratch = kmap_partition_scratch_bytes(n, kmax) */
int kmap_count_all_k(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                     int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                     uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                     void* const* phase_events, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, reads sharded by contiguous ranges (the reference has no such path; reads are its
 * independent units, kmer_count.py:755-759).  The one exchange step is an integer all-reduce of the dense tables over
 * NVLink / NVSwitch (NCCL, bound at run time from the library the process already carries).  The communicator is made
 * from a 128-byte unique id: rank 0 calls kmap_comm_unique_id and the host plumbing (torch.distributed) broadcasts the
 * bytes; every rank then calls kmap_comm_init on ITS device.  The handle is owned by the caller (kmap_comm_destroy).
 * ---------------------------------------------------------------------------------------------------------- */
int kmap_comm_available(void);
int kmap_comm_unique_id(uint8_t* id_out_host);                                     /* 128 bytes, HOST */
int kmap_comm_init(const uint8_t* id_host, int rank, int world, void** comm_out_host);
int kmap_comm_destroy(void* comm);
/* table[i] = sum over the ranks of table[i], in place (uint32, modular): count_uniq_hash of the whole input from the
 * per-shard tables.  Asynchronous on `stream`. */
int kmap_table_allreduce(uint32_t* table, int64_t n_cells, void* comm, void* stream);
/* the same exchange as a reduce-scatter by key range, in place: afterwards cells [rank * n_cells / world, (rank + 1) * n_cells /
 * world) hold the sums over the ranks (the other cells partial sums); n_cells must be a multiple of world */
int kmap_table_reduce_scatter(uint32_t* table, int64_t n_cells, int rank, int world, void* comm, void* stream);
/* The same exchange over NVLink PEER MEMORY, one byte per cell (csrc/peer.cu): a rank's table of a sharded input holds small
 * counts, so the cells travel as signed bytes -- the owner of a key range loads the bytes of every rank over NVLink, sums,
 * and stores the narrowed sums into every rank's memory -- and a cell that does not fit a byte is marked and fetched as a
 * word from the owner's table: bit-identical to the all-reduce for any input at 1/8 of its link traffic.  Every rank
 * allocates a region (kmap_peer_region_alloc: [table area: table_cells uint32][2 bytes per cell][flags], returns the
 * 64-byte IPC handle), the host plumbing gathers the handles of all ranks, and kmap_comm_attach_peers maps the peers'
 * regions.  From then on kmap_table_allreduce / kmap_table_reduce_scatter / kmap_count_all_k_sharded / _scattered take
 * this path for every table (a multiple of 16 x world cells at a multiple of 16 cells) that lies INSIDE the table area of the
 * region, the same cells on every rank, and the NCCL path for any other buffer.  At most 8 ranks, one node.  A rank that
 * does not show up at a barrier for 120 s is reported by kmap_comm_peer_status (collective calls do not hang the GPU). */
int64_t kmap_peer_region_bytes(int64_t table_cells);
int kmap_peer_region_alloc(int64_t table_cells, void** region_out_host, uint8_t* handle_out_host);   /* 64 bytes, HOST */
int kmap_peer_region_free(void* region);
int kmap_comm_attach_peers(void* comm, int rank, int world, void* my_region, int64_t table_cells, const uint8_t* handles_host);
int kmap_comm_detach_peers(void* comm);
int kmap_comm_peer_status(void* comm, int* status_out_host, void* stream);
/* kmap_count_all_k on this rank's shard of the reads with the tables MERGED over the ranks of `comm` on return (every rank
 * gets the tables of the whole input).  The all-reduces are issued on `comm_stream` as the buffers become final -- the
 * corrections of the small levels during the partition pass, the slices of the level-kmax table while the per-bucket
 * count is still running -- and the 4:1 reductions run on the merged buffers.  Every rank must call it with the same
 * kmin, kmax, dedup, scheme and with tables laid out the same way (one buffer, largest level first); an empty shard
 * (n = 0) takes part in the collectives only.  `stream` waits for the exchange before the reductions. */
int kmap_count_all_k_sharded(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                             int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                             uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                             void* const* phase_events, void* stream, void* comm, void* comm_stream);
/* kmap_count_all_k_sharded with the merged tables left scattered over the ranks by key range (reduce-scatter instead of
 * all-reduce: half the exchange volume): rank r owns cells [r * 4^k / world, (r + 1) * 4^k / world) of every level k, the
 * reductions to the lower levels run on the owned ranges only.  world must divide 4^kmin.  Compact a range with
 * kmap_compact_merge_range: the lists of ranks 0, 1, .. concatenate to the reference's list. */
int kmap_count_all_k_scattered(const uint32_t* packed, const uint32_t* valid, int64_t n, const int64_t* borders, int64_t n_seq,
                               int kmin, int kmax, int dedup, uint32_t* const* tables_host, uint32_t* dupmask, uint32_t* work,
                               uint32_t* bitmap, int n_partitions, int scheme, void* part_scratch, int64_t part_scratch_bytes,
                               void* const* phase_events, void* stream, void* comm, void* comm_stream, int rank, int world);

/* kmap_count_dense for tables beyond L2 (9 <= k <= 14) without one global atomic per window: windows are partitioned
 * by the top bits of their key into 4^(k-8) buckets of 16-bit suffixes (scratch), then every bucket is counted in
 * shared memory and its 65536-cell slice of the table written once (csrc/partition.cu).  Same result as
 * kmap_count_dense on a zeroed table (the caller zeroes it).  scratch = kmap_partition_scratch_bytes(n, k) bytes.
 * kmap_count_all_k uses the same scheme for its level-kmax table with scheme = KMAP_KMAX_SORTED (12 <= kmax <= 14);
 * with part_scratch == NULL it falls back to n_partitions key-prefix passes of global atomics. */
int kmap_count_dense_partitioned(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint32_t* table,
                                 void* scratch, int64_t scratch_bytes, void* stream);
int64_t kmap_partition_scratch_bytes(int64_t n, int k);

int kmap_fill_u32(uint32_t* p, int64_t n_words, uint32_t value, void* stream);
/* dst[i] += src[i]: merges the tables of one chunk of reads into the running totals when the reads are streamed through
 * the device in chunks (kmap_b200/api.py; reads are independent units, kmer_count.py:755-759).  16-byte aligned. */
int kmap_add_u32(uint32_t* dst, const uint32_t* src, int64_t n_words, void* stream);
/* borders[i][0..1] -= offset, in place: the rows of input.seqboarder.bin.pkl that belong to a chunk of reads, made
 * relative to the first position of the chunk */
int kmap_rebase_borders(int64_t* borders, int64_t n_seq, int64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host side of the boundary (csrc/host_pack.cpp: plain C++, HOST pointers, no stream): the end-to-end call is bound by
 * the PCIe link when input.bin travels at one byte per base (kmer_count.py:326-347 layout), so the host cores re-encode
 * each chunk of reads into the packed form (0.375 B/position) in pinned staging memory while the previous chunk travels
 * and is counted (kmap_b200/api.py count_tables_streamed).  Encoders only: nothing is hashed or counted on the host.
 * ---------------------------------------------------------------------------------------------------------- */
int kmap_host_threads(void);
/* what kmap_pack2bit writes, from and to host memory; packed = uint32[kmap_packed_words(n)], valid = uint32[kmap_valid_words(n)];
 * n_threads <= 0: all hardware threads */
int kmap_host_pack2bit(const uint8_t* seq, int64_t n, uint32_t* packed, uint32_t* valid, int n_threads);
/* rows of input.seqboarder.bin.pkl for reads laid out back to back from position `first` (kmer_count.py:335-343: st_0 = first,
 * st_{i+1} = en_i + 1) -> strides_out[i] = en_i - st_i + 1 (uint32: 4 instead of 16 bytes per read over the link).
 * KMAP_ERR_BAD_ARG if the rows are not back to back (the caller then ships the matrix itself). */
int kmap_host_border_strides(const int64_t* borders, int64_t n_seq, int64_t first, uint32_t* strides_out, int n_threads);
/* device: the border matrix (relative to the first read) back from the strides: offsets = int64[n_seq + 1] (exclusive prefix
 * sums, kept), borders_out = int64[n_seq][2] = (offsets[i], offsets[i + 1] - 1); scratch = uint64[kmap_list_scratch_words(n_seq)] */
int kmap_borders_from_strides(const uint32_t* strides, int64_t n_seq, int64_t* offsets, int64_t* borders_out, uint64_t* scratch, void* stream);

/* count_uniq_hash (kmer_count.py:476-491) for callers that hold a materialised hash array: table[h] += 1 for every
 * h < 4^k (the invalid hash is skipped) */
int kmap_count_hashes_u32(const uint32_t* hash, int64_t n, int k, uint32_t* table, void* stream);
/* rebuild a dense table from a (unique hash, count) list: table[kh[i]] += cnt[i]; used by merge_revcom when it is
 * called on lists (kmer_count.py:643) */
int kmap_scatter_counts(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, uint32_t* table, void* stream);
/* the side effect of merge_revcom on its count argument (kmer_count.py:661): cnt[i] += table[rc(kh[i])] */
int kmap_list_add_rc_counts(const uint32_t* kh, int32_t* cnt, int64_t n, int k, const uint32_t* table, void* stream);

/* count_uniq_hash + merge_revcom (kmer_count.py:476-491, 643-685) from the dense forward table, in the
 * reference's exact output order (ascending forward hash of the surviving entries; value = min(h, rc h) when
 * revcom == 1, max(h, rc h) when revcom == 2 -- merge_revcom's keep_lower_hash_flag=False --; palindromes doubled).  scratch = uint64[kmap_compact_scratch_words(k)] (256-byte aligned; for k >= 13 it
 * also holds a permuted copy of the table, G[h] = F[rc h], so that the merge reads its partner cell without a gather).
 * Two-step: call with capacity 0 (out pointers may be NULL) to get *n_out_host, then with buffers.
 * Synchronises the stream. */
int kmap_compact_merge(const uint32_t* table, int k, int revcom, uint64_t* scratch, uint32_t* kh_out,
                       int32_t* cnt_out, int64_t capacity, int64_t* n_out_host, void* stream);
/* The same for the forward hashes in [cell_lo, cell_hi) only (multiples of 2048; cell_hi may be 4^k): the entries of the
 * merged list whose forward hash lies in the range, in list order -- the lists of consecutive ranges concatenate to the
 * whole list, so the ranks of a sharded count can each compact and ship one key range. */
int kmap_compact_merge_range(const uint32_t* table, int k, int revcom, int64_t cell_lo, int64_t cell_hi, uint64_t* scratch, uint32_t* kh_out,
                             int32_t* cnt_out, int64_t capacity, int64_t* n_out_host, void* stream);
int64_t kmap_compact_scratch_words(int k);

/* Hamming-ball count of find_motif (motif_discovery.py:666-673) for m candidate consensus hashes, evaluated by
 * neighbour enumeration over the dense forward table: sums[i] = sum of merged counts within distance d of
 * cand[i] (or of its reverse complement when revcom).  sums = uint64[m] (overwritten). */
int kmap_hamball_sum(const uint32_t* table, int k, const uint32_t* cand, int m, int d, int revcom, uint64_t* sums,
                     void* stream);
/* the same quantity computed the reference's way, as a scan of the merged (kh, cnt) list: used for lists that
 * did not come from a dense table (a loaded k{k}.pkl) */
int kmap_hamball_sum_list(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, const uint32_t* cand, int m,
                          int d, int revcom, uint64_t* sums, void* stream);

/* mask_input (kmer_count.py:580-610) on the packed representation: for each of the m (<= 16) consensus hashes flag
 * the positions whose PRE-MASK window (validity taken from valid_pre) is within d[i] (invalid windows compare as
 * T..T) and clear bits [i, min(i+k, n)) of valid.  valid_pre may alias valid; callers with more than 16 consensus
 * hashes pass a snapshot as valid_pre on every call.  flag_scratch = uint32[kmap_valid_words(n)]; cons/d are
 * device arrays. */
int kmap_mask(const uint32_t* packed, const uint32_t* valid_pre, uint32_t* valid, int64_t n, int k, const uint32_t* cons,
              const int32_t* d, int m, uint32_t* flag_scratch, void* stream);

/* ex_hamball_kh_arr (motif_discovery.py:959-975) on a merged (kh, cnt) list + cal_cnt_mat
 * (motif_discovery.py:978-986).  Ball members keep list order; members strictly closer to rc(conseq) are
 * reverse-complemented.  cnt_mat = int64[4*k] row-major [base][pos] (overwritten).
 * scratch = uint64[kmap_list_scratch_words(n)].  Two-step like kmap_compact_merge.  Synchronises the stream. */
int kmap_hamball_extract(const uint32_t* kh, const int32_t* cnt, int64_t n, int k, uint32_t conseq, int d,
                         int revcom, uint64_t* scratch, uint32_t* kh_out, int32_t* cnt_out, int64_t capacity,
                         int64_t* n_out_host, int64_t* cnt_mat, void* stream);
int64_t kmap_list_scratch_words(int64_t n);

/* get_motif_occurence (motif_discovery.py:1441-1465) for all reads at once.  Per read r: positions whose
 * min(fwd, rc) distance to conseq is <= d and equal to the read's minimum such distance.
 * step 1: min_dist[r] (255 = no hit) and n_hit[r];  step 2 (after the caller's exclusive scan of n_hit into
 * offsets, int64[n_seq]): pos_out[offsets[r] ...] = hit positions ascending. */
int kmap_occurrence_count(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq,
                          int k, uint32_t conseq, int d, int revcom, uint8_t* min_dist, uint32_t* n_hit, void* stream);
int kmap_occurrence_fill(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq,
                         int k, uint32_t conseq, int d, int revcom, const uint8_t* min_dist, const int64_t* offsets,
                         int32_t* pos_out, void* stream);

/* the same three for 64-bit hashes (17 <= k <= 31; any k <= 31 is accepted): plain one-thread-per-position flag kernel */
int kmap_mask_u64(const uint32_t* packed, const uint32_t* valid_pre, uint32_t* valid, int64_t n, int k, const uint64_t* cons,
                  const int32_t* d, int m, uint32_t* flag_scratch, void* stream);
int kmap_occurrence_count_u64(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                              uint64_t conseq, int d, int revcom, uint8_t* min_dist, uint32_t* n_hit, void* stream);
int kmap_occurrence_fill_u64(const uint32_t* packed, const uint32_t* valid, const int64_t* borders, int64_t n_seq, int k,
                             uint64_t conseq, int d, int revcom, const uint8_t* min_dist, const int64_t* offsets,
                             int32_t* pos_out, void* stream);

/* The rows of a *.motif_occurence.csv (gen_motif_occurence_file, motif_discovery.py:1409-1418; cell format :1472-1475)
 * from the results of the occurrence scan, formatted natively: for every read r in [r0, r1) with at least one hit, the
 * line "r;cell_0;...;cell_{m-1};seq_len[r]\n", cell_j = pos_host[j][offsets_host[j][r] .. offsets_host[j][r+1]) joined
 * by ','.  ALL pointers are HOST pointers (offsets_host / pos_host: m pointers each); no device work.  The file is
 * created (append == 0) or appended to.  Returns the number of rows written, or a negative KMAP_ERR_*.
 * (Cells with more than 20 positions need the reference's random pick, :1467-1469: the Python caller formats those rows.) */
int64_t kmap_write_occurrence_rows(const char* path_host, int append, int m, const int64_t* const* offsets_host,
                                   const int32_t* const* pos_host, const int64_t* seq_len_host, int64_t r0, int64_t r1);

/* cal_samp_kmer_hamdist_mat (motif_discovery.py:777-803): rows [row0, row1) of the n x n distance matrix of
 * kh[] at k bases; pairs whose labels are equal and have head_len[label] < k use only the first head_len bases.
 * head_len = int32[n_labels] (k for labels without override).  out = uint8[(row1-row0) * n] row-major. */
int kmap_hamdist_matrix_u32(const uint32_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len,
                            int n_labels, int64_t row0, int64_t row1, uint8_t* out, void* stream);
int kmap_hamdist_matrix_u64(const uint64_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len,
                            int n_labels, int64_t row0, int64_t row1, uint8_t* out, void* stream);
/* The same matrix (k <= 16) computed as an int8 one-hot GEMM on the tcgen05 tensor cores (csrc/hamdist_mma.cu): the
 * formulation BASELINE config 5 asks to be measured against XOR/popcount, and the faster one (2.1 against 2.6-2.9 ms for 1e10
 * pairs); bit-identical output.  The head override is applied by extra K columns when they fit into K = 128, else by the
 * epilogue.  scratch = kmap_hamdist_mma_scratch_bytes(n) bytes of device memory, 256-byte aligned (the one-hot operands).
 * Reads head_len back (n_labels ints): synchronises the stream once. */
int64_t kmap_hamdist_mma_scratch_bytes(int64_t n);
int kmap_hamdist_matrix_onehot_mma(const uint32_t* kh, const int32_t* labels, int64_t n, int k, const int32_t* head_len, int n_labels,
                                   int64_t row0, int64_t row1, uint8_t* out, void* scratch, int64_t scratch_bytes, void* stream);

/* exclusive prefix sum of uint32 counts into int64 offsets (out[n] = total); scratch = uint64[kmap_list_scratch_words(n)] */
int kmap_exclusive_scan_u32(const uint32_t* in, int64_t n, int64_t* out, uint64_t* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Integer parts of the consumers of the occurrence scan (csrc/consumers.cu), from the scan results in device memory
 * (offsets int64[n_seq + 1] and ascending positions int32[] per consensus, as kmap_occurrence_count / _fill + the scan
 * produce them) instead of final.motif_occurence.csv parsed back.
 * ---------------------------------------------------------------------------------------------------------- */
/* the `sum()` of find_motif (motif_discovery.py:648): *sum_out (device int64) = total of a count list */
int kmap_sum_counts_i32(const int32_t* cnt, int64_t n, int64_t* sum_out, void* stream);
int kmap_sum_counts_i64(const int64_t* cnt, int64_t n, int64_t* sum_out, void* stream);
/* get_motif_co_occurence_mat (motif_discovery.py:1189-1254), m <= 31 motifs.  offsets_host / positions_host: HOST arrays of m
 * device pointers.  present_out[r]: bit i = motif i is listed in read r, bit 31 = some cell of the read lists more than 20
 * positions (the reference keeps a random 20 of them, :1467-1469: such reads are left to the host and are in no count).
 * counts_out int64[m * m] (device): [i][i] = reads that list motif i, [i][j], i < j = reads that list both. */
int kmap_cooc_reads(const int64_t* const* offsets_host, const int32_t* const* positions_host, int m, int64_t n_seq, uint32_t* present_out,
                    int64_t* counts_out, void* stream);
/* flags_out[r] = 1 iff read r lists motifs i and j (and is not left to the host); scan it with kmap_exclusive_scan_u32, then
 * kmap_cooc_pair_fill writes, in read order, the read index and TWICE (median position of j - median position of i)
 * (:1227-1240: np.median of the listed positions; the factor two keeps the .5 of an even count in an integer). */
int kmap_cooc_pair_flags(const uint32_t* present, int64_t n_seq, int i, int j, uint32_t* flags_out, void* stream);
int kmap_cooc_pair_fill(const int64_t* off_i, const int32_t* pos_i, const int64_t* off_j, const int32_t* pos_j, const uint32_t* flags,
                        const int64_t* flag_offsets, int64_t n_seq, int64_t* read_out, int32_t* diff2_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * sort / run-length counting path (csrc/sorted.cu): the same pipeline for k-mers without a dense table
 * (16 <= k <= 31: uint64 hashes, int64 counts, kmer_count.py:351-365; any 1 <= k <= 31 is accepted so that the
 * two paths can be measured against each other).  It follows the reference's own flow: one hash per position ->
 * per-read de-duplication in place -> np.unique.
 * ---------------------------------------------------------------------------------------------------------- */
/* comp_kmer_hash_taichi (kmer_count.py:449-473) from the packed reads: keys[p] = hash of the window at p, all-ones when
 * the window leaves the array or touches a 255 */
int kmap_window_keys_u64(const uint32_t* packed, const uint32_t* valid, int64_t n, int k, uint64_t* keys, void* stream);
/* remove_duplicate_hash_per_seq (kmer_count.py:743-760) on a uint64 hash array, in place: later occurrences of a hash
 * inside a read become all-ones.  work = uint32[kmap_dedup_keys_work_words(n_seq)] scratch. */
int kmap_dedup_hash_per_read_u64(uint64_t* hash, int64_t n, const int64_t* borders, int64_t n_seq, uint32_t* work, void* stream);
int64_t kmap_dedup_keys_work_words(int64_t n_seq);
/* count_uniq_hash (kmer_count.py:476-491), step 1: LSD radix sort of the low key_bits bits, 8 bits per pass; all-ones
 * keys are dropped by the first pass.  On return keys[0 .. *n_valid_host) is ascending; *n_unique_host = number of
 * distinct keys.  tmp = uint64[n]; scratch = uint64[kmap_sort_scratch_words(n)].  Synchronises the stream. */
int kmap_sort_keys_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int key_bits, uint64_t* scratch, int64_t* n_valid_host,
                       int64_t* n_unique_host, void* stream);
int64_t kmap_sort_scratch_words(int64_t n);
/* step 2: run-length encoding of n sorted keys: kh_out = distinct keys ascending, cnt_out = their multiplicities
 * (capacity >= the number of distinct keys).  scratch as above, pos_scratch = int64[capacity].  Synchronises. */
int kmap_rle_u64(const uint64_t* sorted_keys, int64_t n, uint64_t* scratch, int64_t* pos_scratch, uint64_t* kh_out,
                 int64_t* cnt_out, int64_t capacity, void* stream);
/* count_uniq_hash (kmer_count.py:476-491) of a sharded input = the per-shard (hash, count) lists added up:
 * cnt_out[j] += cnt[i] where uniq[j] == kh[i] (uniq ascending and distinct: the sorted union of the shards' hashes; an
 * absent hash is skipped).  The caller zeroes cnt_out and calls this once per shard list. */
int kmap_list_add_counts_u64(const uint64_t* uniq, int64_t n_uniq, const uint64_t* kh, const int64_t* cnt, int64_t n, int64_t* cnt_out,
                             void* stream);
/* merge_revcom (kmer_count.py:643-685) on an ASCENDING unique list: survivors in list order, value min(h, rc h), count
 * cnt[h] + cnt[rc h] (a palindrome is its own partner: doubled); keep_higher != 0 is keep_lower_hash_flag=False
 * (kmer_count.py:671, 682: the higher hash of a pair survives, value max(h, rc h)).  Two-step like kmap_compact_merge (capacity 0 = size
 * query).  summed_cnt (may be NULL, must not alias cnt) = int64[n]: cnt[i] + cnt[partner of i] for every i, what the
 * reference leaves in the caller's count array (kmer_count.py:661).
 * scratch = uint64[kmap_merge_sorted_scratch_words(n)].  Synchronises. */
int kmap_merge_revcom_sorted_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, int keep_higher, uint64_t* scratch, uint64_t* kh_out,
                                 int64_t* cnt_out, int64_t capacity, int64_t* n_out_host, int64_t* summed_cnt, void* stream);
int64_t kmap_merge_sorted_scratch_words(int64_t n);
/* kmap_hamball_sum_list / kmap_hamball_extract for uint64 hashes and int64 counts (k <= 31) */
int kmap_hamball_sum_list_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, const uint64_t* cand, int m, int d,
                              int revcom, uint64_t* sums, void* stream);
int kmap_hamball_extract_u64(const uint64_t* kh, const int64_t* cnt, int64_t n, int k, uint64_t conseq, int d, int revcom,
                             uint64_t* scratch, uint64_t* kh_out, int64_t* cnt_out, int64_t capacity, int64_t* n_out_host,
                             int64_t* cnt_mat, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * preproc ingest (csrc/fasta.cu): FASTA text -> the contents of input.bin.pkl / input.seqboarder.bin.pkl
 * (kmer_count.py:244-263 dna2arr, 308-323 read_dnaseq_file, 326-347 convert_fasta_to_binary) on the device.
 * The text (device bytes, 16-byte aligned, starting at the first header line) may be fed in chunks; `state` is a HOST
 * int64[4] carried from chunk to chunk: {sequence characters so far, records so far, type of the line in progress
 * (0 sequence line, 1 header line), last byte}; before the first chunk {0, 0, 0, 10}.
 * kmap_fasta_scan : fills scratch = uint64[kmap_fasta_scratch_words(n)] and state_out_host.  Synchronises the stream.
 * kmap_fasta_emit : writes the encoded bytes of the chunk (upper-cased; A0 C1 G2 T3, anything else 255; one 255 after
 *                   every record) to seq_out[p - seq_origin] for their positions p in input.bin, and rec_start_out[j] =
 *                   first position of the j-th record that begins in this chunk.  A chunk produces positions
 *                   [in[0] + max(in[1]-1, 0), out[0] + out[1] - 1) plus, with final_chunk, the last separator.
 * kmap_borders_from_starts : borders[r] = {start_r, start_{r+1} - 1} with start_{n_rec} = total_len, the
 *                   [start, separator index] rows of input.seqboarder.bin (kmer_count.py:335-343).
 * ---------------------------------------------------------------------------------------------------------- */
int64_t kmap_fasta_scratch_words(int64_t n);
int kmap_fasta_scan(const uint8_t* text, int64_t n, const int64_t* state_in_host, uint64_t* scratch, int64_t* state_out_host,
                    void* stream);
int kmap_fasta_emit(const uint8_t* text, int64_t n, const int64_t* state_in_host, const uint64_t* scratch, uint8_t* seq_out,
                    int64_t seq_origin, int64_t* rec_start_out, int final_chunk, const int64_t* state_out_host, void* stream);
int kmap_borders_from_starts(const int64_t* rec_start, int64_t n_rec, int64_t total_len, int64_t* borders, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * synthetic reads (bench / tests): counter-based generator, identical to kmap_b200/synth.py
 * seq = uint8[n_reads*(L+1)] in the input.bin layout, borders = int64[n_reads][2] (may be NULL)
 * ---------------------------------------------------------------------------------------------------------- */
int kmap_synth_reads(uint64_t seed, int64_t read0, int64_t n_reads, int L, const uint8_t* motifs, const int32_t* motif_len,
                     const float* motif_cum_frac, int n_motifs, float mut_rate, float n_rate, int64_t pos0,
                     uint8_t* seq, int64_t* borders, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KMAP_B200_H */
