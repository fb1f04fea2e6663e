/* rg_b200.h — C-ABI of the B200-native RoarGraph hot path (librg_b200.so).
 *
 * Drop-in boundary for the reference's distance-bound paths (paths relative to /root/reference):
 *   - IndexBipartite::SearchRoarGraph        src/index_bipartite.cpp:2311-2420
 *   - IndexBipartite::SearchBipartiteGraph   src/index_bipartite.cpp:282-356
 *   - exact_knn (compute_groundtruth)        thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:126-248
 * The reference has no FFI of its own (single C++ process); these entry points are what its
 * host code would bind at the `#pragma omp parallel for` over queries
 * (tests/test_search_roargraph.cpp:203-209) and at the exact_knn call (compute_groundtruth.cpp:404).
 * INTEGRATION.md shows the reference-side binding.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * RG_E* code, with a human-readable message in rg_last_error() (thread-local).  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with RG_ECUDA.
 * An rg_index is not re-entrant: one host thread per index at a time (the reference shares one
 * index among OpenMP threads; here the whole query batch crosses the boundary in one call).
 */
#ifndef RG_B200_H
#define RG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* efanna2e::Metric values (include/efanna2e/distance.h:15).  COSINE is INNER_PRODUCT on rows the caller
 * has already normalised, exactly as in the reference (src/index.cpp:16-20, src/index_bipartite.cpp:2679-2684,
 * tests/test_search_roargraph.cpp:167-172). */
#define RG_METRIC_L2 0
#define RG_METRIC_INNER_PRODUCT 1
#define RG_METRIC_COSINE 4

#define RG_OK 0
#define RG_EINVAL (-1)     /* bad argument */
#define RG_ECUDA (-2)      /* CUDA runtime error / no device */
#define RG_ENOTENOUGH (-3) /* "not enough results: N, expected: k"  (src/index_bipartite.cpp:2408-2412) */
#define RG_ESTATE (-4)     /* graph not set, etc. */
#define RG_ENOMEM (-5)

typedef struct rg_index rg_index;
typedef struct rg_knn_base rg_knn_base;

const char *rg_last_error(void);
/* ABI version of this header. */
int rg_abi_version(void);
/* Make `device` the calling thread's current CUDA device (one process — or one host thread — per GPU).  Optional: without
 * it the library adopts whatever device is current on the calling thread (device 0 in a fresh thread).  An rg_index /
 * rg_knn_base remembers the device it was created on and switches to it on every call. */
int rg_init(int device);
/* Number of SMs etc. of the active device; returns <0 when no device is usable. */
int rg_device_sm_count(void);

/* ---- index lifetime ---------------------------------------------------------------------------
 * base: nd x dim float32, row-major, row stride == dim, dim % 8 == 0 (the reference pads every row to x8
 * floats with zeros, include/efanna2e/util.h:191-199 / :37-75, and constructs the index with the padded
 * dim, tests/test_search_roargraph.cpp:132,160).  base_on_device != 0 means `base` is already a device
 * pointer that outlives the index (no copy is made); otherwise it is a host pointer copied to HBM once.
 * Replaces: IndexBipartite ctor + LoadSearchNeededData (include/index_bipartite.h:27,62). */
int rg_index_create(rg_index **out, const float *base, uint64_t nd, uint32_t dim, int metric, int base_on_device);

/* The same from an fbin file (u32 n, u32 d, n*d f32; README.md:14, include/efanna2e/util.h:107-211), streamed: the file is read
 * in 64 MB slabs into two pinned staging buffers; each slab is handed to `hook` (optional; rows of the file's UNPADDED dim,
 * e.g. the reference's normalize<float> for COSINE, src/index_bipartite.cpp:2679-2684) and copied to HBM asynchronously
 * while the next slab is being read, the rows zero-padded to x8 floats by the copy itself (util.h:37-75).  Replaces
 * load_data + data_align + the pageable 8-20 GB host copy of LoadVectorData.  out_nd / out_dim (padded) may be NULL. */
typedef void (*rg_row_hook)(float *rows, uint64_t n, uint32_t dim, void *user);
int rg_index_create_from_fbin(rg_index **out, const char *path, int metric, rg_row_hook hook, void *user, uint64_t *out_nd,
                              uint32_t *out_dim);

/* Projection graph as host CSR (row_off[nd+1], nbrs[row_off[nd]]) + entry point, i.e. the content of
 * IndexBipartite::GetProjectionGraph() / LoadProjectionGraph (src/index_bipartite.cpp:2097-2117). */
int rg_index_set_graph(rg_index *idx, uint32_t ep, const uint64_t *row_off, const uint32_t *nbrs);

/* Raw bipartite graph over n_total = nd + nd_sq nodes as host CSR (IndexBipartite::Load, :2058-2071). */
int rg_index_set_bipartite(rg_index *idx, uint64_t n_total, const uint64_t *row_off, const uint32_t *nbrs);

int rg_index_destroy(rg_index *idx);

/* ---- BuildRoarGraph on the device ----------------------------------------------
 * knn_ids: nsq x knn_stride base ids, row i = the nearest base rows of training query i in ascending distance — the
 * content of the learn-base kNN file read by LoadLearnBaseKNN (src/index_bipartite.cpp:2622-2642); the first M_sq of
 * every row are used (LinkProjection :1063-1067).  knn_on_device != 0: device pointer.  M_pjbp / L_pjpq as in
 * tests/test_build_roargraph.cpp:52-60.  n_rounds = number of bulk-synchronous rounds of the connectivity-enhancement
 * phase (0 = automatic).  Replaces BuildRoarGraph (:143-233); the result is deterministic, and statistically — not
 * id-for-id — equivalent to the reference's thread-timing-dependent build.  Sets the index's graph and entry point. */
int rg_index_build_roargraph(rg_index *idx, const uint32_t *knn_ids, int knn_on_device, uint64_t nsq,
                             uint32_t knn_stride, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq, uint32_t n_rounds);
/* Export the index's projection graph (set or built) as host CSR: what SaveProjectionGraph writes (:2606-2619). */
int rg_index_graph_size(rg_index *idx, uint64_t *nnz, uint32_t *ep, uint32_t *max_degree);
int rg_index_get_graph(rg_index *idx, uint64_t *row_off, uint32_t *nbrs);

/* ---- SearchRoarGraph, whole batch -------------------------------------------------------------
 * queries: nq x dim (same padded dim as the base; normalised by the caller for COSINE).
 * Outputs (caller-allocated): ids[nq*k], dists[nq*k] = internal distances in (distance asc, id asc) order
 * (-IP for ip/cosine, squared L2 for l2), cmps[nq] = distance evaluations excluding the entry point,
 * hops[nq] = expansions — the pair SearchRoarGraph returns (:2354,2366,2397,2419).  cmps/hops may be NULL.
 * Requires k <= L_pq (tests/test_search_roargraph.cpp:192-195).
 * Host-buffer form: H2D of queries, kernel, D2H of results, synchronous. */
int rg_search_roargraph(rg_index *idx, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq, uint32_t *ids,
                        float *dists, uint32_t *cmps, uint32_t *hops);
/* Device-buffer form: all pointers are device pointers, work is enqueued on `stream` (a cudaStream_t, may be
 * NULL for the default stream) and the call returns without synchronising.  Query failures ("not enough
 * results") are latched in the index and reported by rg_index_check(). */
int rg_search_roargraph_dev(rg_index *idx, const float *d_queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                            uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops, void *stream);
/* Synchronise the index's work on `stream` and return RG_OK or the latched error of earlier *_dev calls. */
int rg_index_check(rg_index *idx, void *stream);

/* ---- SearchBipartiteGraph, whole batch ----------------------------------------------------------
 * init_ids: nq x n_init base ids used as entry points (the reference draws 10 per query from an unseeded
 * std::random_device, :287-294; pass them explicitly for reproducibility) or NULL to let the library draw
 * n_init ids per query from `seed` (splitmix64 counter; uniform over [0, nd)). */
int rg_search_bipartite(rg_index *idx, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                        const uint32_t *init_ids, uint32_t n_init, uint64_t seed, uint32_t *ids, float *dists,
                        uint32_t *cmps, uint32_t *hops);
int rg_search_bipartite_dev(rg_index *idx, const float *d_queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                            const uint32_t *d_init_ids, uint32_t n_init, uint64_t seed, uint32_t *d_ids,
                            float *d_dists, uint32_t *d_cmps, uint32_t *d_hops, void *stream);

/* ---- roofline accounting --------------------------------------------------------------------------
 * Totals of the LAST search call on this index (synchronises): distance evaluations (unique, as the
 * reference counts them), expansions, and the sum of out-degrees of the expanded nodes.  Algorithmic bytes
 * per SURVEY.md 8(d) = (cmps + nq) * dim * 4 + adj_words * 4. */
int rg_search_last_stats(rg_index *idx, uint64_t *total_cmps, uint64_t *total_hops, uint64_t *adj_words);
/* Launch geometry chosen for the last search (for reporting): grid, block, dynamic smem bytes, rows per batch. */
int rg_search_last_launch(rg_index *idx, uint32_t *grid, uint32_t *block, uint32_t *smem_bytes, uint32_t *rows_per_batch);

/* ---- exact_knn ---------------------------------------------------------------------------------------
 * Brute-force kNN of nq queries against nb base rows (row-major, row stride == dim, NO padding required,
 * as compute_groundtruth reads raw fbin).  Outputs ids[nq*K] (int32, like the reference's closest_points) and
 * dists[nq*K] = internal distances ascending: -IP for INNER_PRODUCT, squared L2 for L2, squared L2 of the
 * normalised rows for COSINE (compute_groundtruth.cpp:146-174).  Candidate selection may run on tensor cores in
 * bf16; every returned distance is recomputed in fp32.  id_offset is added to the ids (the reference's
 * `start_id` for base parts, compute_groundtruth.cpp:413). */
int rg_knn(const float *base, uint64_t nb, const float *queries, uint64_t nq, uint32_t dim, uint32_t K, int metric,
           int32_t id_offset, int32_t *ids, float *dists);
int rg_knn_dev(const float *d_base, uint64_t nb, const float *d_queries, uint64_t nq, uint32_t dim, uint32_t K,
               int metric, int32_t id_offset, int32_t *d_ids, float *d_dists, void *stream);

/* The base side of exact_knn prepared ONCE for any number of query batches — compute_groundtruth keeps a base part in
 * memory and walks the queries past it 1 024 at a time (compute_groundtruth.cpp:187-200, part loop :396-416): squared norms,
 * the normalised copy for COSINE, the bf16 operand copy.  base_on_device != 0: `base` is a device pointer that outlives the
 * object (no copy); otherwise host rows copied to HBM once.  rg_knn_base_search_dev scores one batch of device-resident
 * queries against it; results as rg_knn_dev.  Used by the base-sharded 10 M x 10 M run (BASELINE configs[3]): every GPU
 * keeps its shard prepared while the query batches stream past. */
int rg_knn_base_create(rg_knn_base **out, const float *base, uint64_t nb, uint32_t dim, int metric, int base_on_device);
int rg_knn_base_destroy(rg_knn_base *kb);
int rg_knn_base_search_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                           int32_t *d_ids, float *d_dists, void *stream);
/* Sharded base (G shards of one base, BASELINE configs[3]): the shards agree on ONE float per query before they filter.
 * rg_knn_base_seed_dev writes, per query, an upper bound of the distance of the `rank`-th best row among the shard's first
 * 65 536 rows (from the tensor-core scores alone; +inf when the shard is too small).  With rank = ceil(K / G) the MAXIMUM of
 * the shards' values (one all-reduce of nq floats) bounds the GLOBAL K-th best distance: every shard holds >= rank rows at or
 * below the maximum, G * rank >= K.  rg_knn_base_search_seeded_dev then scores the batch with that bound in force from row 0:
 * a shard re-ranks ~K/4 instead of ~K * log2(rows) survivors per query.  Its per-shard lists may hold fewer than K entries
 * (id -1, distance +inf behind them): every row of the global top-K is at or below the bound, so the MERGED top-K is exact.
 * bound_rows = the number of rows the bound is (about) the K-th best of (G x 65 536; 0 = unknown): with RG_KNN_WHOLE_SHARD=1
 * in the environment and the survivors expected over the whole shard fitting a candidate list, the shard is scored in ONE
 * filter + merge launch pair (measured slower than the default growing chunks; an option). */
int rg_knn_base_seed_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, uint32_t rank, float *d_tau, void *stream);
int rg_knn_base_search_seeded_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                                  const float *d_tau, uint64_t bound_rows, int32_t *d_ids, float *d_dists, void *stream);
/* The same with ONE re-sync of the bound inside the shard: at the first chunk boundary at or past `resync_rows` rows (after the
 * last chunk when the shard is shorter, so that every shard calls the hook the same number of times: once per internal query
 * batch — 151 552 queries for K <= 512 on a 148-SM part; the count depends only on nq, K and the SM count) the shard writes its `rank`-th best EXACT distance so far into d_bound[q] (+inf while its list is shorter)
 * and calls hook(user, q_off, n, stream), which must reduce d_bound[q_off, q_off + n) to the MAXIMUM over the shards on
 * `stream` (one all-reduce) and return RG_OK; the search then continues with tau = min(tau, d_bound).  With rank * G >= K the
 * maximum bounds the global K-th best distance exactly like the seed bound, but from re-ranked rows of a fifth of the base
 * instead of the seed rows' scores, so the rest of the shard passes ~K/G instead of ~2.5 K/G survivors per query. */
typedef int (*rg_knn_bound_hook)(void *user, uint64_t q_off, uint64_t n, void *stream);
int rg_knn_base_search_resync_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                                  const float *d_tau, uint64_t bound_rows, uint32_t rank, uint64_t resync_rows, float *d_bound,
                                  rg_knn_bound_hook hook, void *user, int32_t *d_ids, float *d_dists, void *stream);
/* Device milliseconds per phase of this thread's kNN calls since the last rg_knn_profile_reset(), measured only when the
 * environment has RG_KNN_TRACE=1 (events + a synchronisation per phase: a diagnostic pass, never the timed one).
 * out[0..6] = base preparation, query conversion, tau seed pass, tensor-core filter, re-rank/merge, chunks, overflow retries. */
int rg_knn_profile(double *out7);
int rg_knn_profile_reset(void);
/* on = 1 / 0: switch the per-phase timers on / off for this process whatever RG_KNN_TRACE says; -1: follow the environment. */
int rg_knn_trace(int on);

/* rg_knn / rg_knn_dev keep their device scratch (bf16 operand copies, candidate lists; a few GB at 10 M rows) in a
 * grow-only per-process workspace so that repeated calls do not pay cudaMalloc/cudaFree; this gives it back. */
int rg_knn_release_workspace(void);

/* Merge G per-shard top-K lists per query (each ascending) into one ascending top-K:
 * the std::sort merge of compute_groundtruth.cpp:424-444 / the host merge of per-shard search results.
 * in_dists/in_ids: [G][nq][K] device arrays (e.g. an ncclAllGather result); out: [nq][K]. */
int rg_topk_merge_dev(const float *d_in_dists, const uint32_t *d_in_ids, uint32_t G, uint64_t nq, uint32_t K,
                      float *d_out_dists, uint32_t *d_out_ids, void *stream);

/* The same merge for ONE all-gather per exchange (BASELINE.json north_star): rg_topk_pack_dev packs a shard's (distance, id)
 * block into one 8-byte word per entry — high word = the float's bits, low word = id + id_offset (the shard's first global
 * row, compute_groundtruth.cpp:413; the "no result" id 0xffffffff is kept) — and rg_topk_merge_packed_dev merges the
 * gathered [G][nq][K_in] words into the K_out best per query for the queries [q_lo, q_hi) of the batch (out arrays are
 * indexed by the query's position in the batch).  K_in may differ from K_out (a shard may return fewer than k). */
int rg_topk_pack_dev(const float *d_dists, const uint32_t *d_ids, uint64_t n, uint32_t id_offset, uint64_t *d_packed, void *stream);
int rg_topk_merge_packed_dev(const uint64_t *d_packed, uint32_t G, uint64_t nq, uint32_t K_in, uint32_t K_out, uint64_t q_lo,
                             uint64_t q_hi, float *d_out_dists, uint32_t *d_out_ids, void *stream);

/* ---- several GPUs of one box from ONE host process ------------------------------------------------------------------
 * The reference's CLIs are single processes; for BASELINE.json's multi-GPU configurations the drop-in stays one process
 * and gives every GPU one worker thread and one NCCL rank (ncclCommInitAll; libnccl.so.2 is bound at run time, only when a
 * group of more than one GPU is created).  The base rows are sharded by equal counts like compute_groundtruth's base parts
 * (compute_groundtruth.cpp:396-416): shard s = rows [s*nd/G, (s+1)*nd/G) on GPU dev_ids[s] (NULL = 0..n_gpus-1), ids made
 * global by the shard's first row (:413), per exchange step ONE all-gather of the packed per-shard top-k blocks and the
 * k-way merge (:424-444).  Python / torchrun jobs use one process per GPU and torch.distributed for the same exchange
 * (roargraph_b200/sharded.py). */
typedef struct rg_group rg_group;
typedef struct rg_sharded_index rg_sharded_index;
int rg_group_create(rg_group **out, int n_gpus, const int *dev_ids);
int rg_group_destroy(rg_group *g);
int rg_group_size(rg_group *g);
/* exact_knn with the base sharded over the group (BASELINE configs[3]); host buffers, contract of rg_knn: every GPU keeps
 * its base part prepared, uploads 1/G of each 151 552-query batch (the rows are all-gathered over NVLink), scores the
 * batch against its part, and merges + returns its 1/G slice of the batch's queries after the packed all-gather. */
int rg_group_knn(rg_group *g, const float *base, uint64_t nb, const float *queries, uint64_t nq, uint32_t dim, uint32_t K,
                 int metric, int32_t *ids, float *dists);
/* Base-sharded RoarGraph search (BASELINE configs[4]): one rg_index per shard, each with its OWN graph over its rows.
 * base: host rows, padded like rg_index_create's.  rg_sharded_index_shard hands out shard s's rg_index (e.g. for
 * rg_index_set_graph with LOCAL ids, rg_index_get_graph, rg_search_last_stats) and its first global row / row count.
 * rg_sharded_index_build_roargraph computes every shard's own learn-base kNN lists (the training queries against the
 * shard's rows, K = M_sq) and builds its graph on its GPU.  rg_sharded_search_roargraph: host buffers; every GPU uploads 1/G
 * of the batch, NVLink all-gather of the rows, per-shard SearchRoarGraph at L_pq, packed all-gather, merge; ids are global
 * rows, (distance asc, id asc) — the merge of the reference's per-shard results. */
int rg_sharded_index_create(rg_group *g, rg_sharded_index **out, const float *base, uint64_t nd, uint32_t dim, int metric);
int rg_sharded_index_destroy(rg_sharded_index *si);
int rg_sharded_index_shards(rg_sharded_index *si);
int rg_sharded_index_shard(rg_sharded_index *si, int shard, rg_index **index, uint64_t *row_lo, uint64_t *rows);
int rg_sharded_index_build_roargraph(rg_sharded_index *si, const float *train, uint64_t nsq, uint32_t M_sq, uint32_t M_pjbp,
                                     uint32_t L_pjpq);
int rg_sharded_search_roargraph(rg_sharded_index *si, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq, uint32_t *ids,
                                float *dists);

#ifdef __cplusplus
}
#endif
#endif /* RG_B200_H */
