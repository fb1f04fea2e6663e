// Internal declarations shared by the translation units of librg_b200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/rg_b200.h"

namespace rg {

constexpr uint32_t kExpandedBit = 0x80000000u;  // queue id word: bit 31 = "expanded" flag (neighbor.h:24 `flag`)
constexpr uint32_t kIdMask = 0x7fffffffu;
constexpr int kMaxRowsPerBatch = 32;
constexpr int kCandCap = 96;  // neighbour ids examined per pass (3 per lane)

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define RG_CUDA(call)                                          \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return rg::cuda_fail(_e, #call); \
    } while (0)

// Launch-time description of one beam-search batch (passed by value to the kernels).
struct SearchParams {
    const float *base;        // nd x dim
    const uint32_t *adj;      // fixed-stride adjacency: row u = [deg, ids...] , adj_stride words   (RoarGraph)
    const uint64_t *row_off;  // CSR (bipartite graph)
    const uint32_t *csr_nbrs;
    uint8_t *tags;            // n_slots x tag_stride visited tags (visited_list_pool.h:8-29, 8-bit epochs)
    uint32_t *slot_epoch;     // n_slots
    const float *queries;     // nq x dim
    const uint32_t *init_ids; // bipartite: nq x n_init (or nullptr -> seeded)
    uint32_t *out_ids;        // nq x k
    float *out_dists;         // nq x k
    uint32_t *out_cmps;       // nq or null
    uint32_t *out_hops;       // nq or null
    uint32_t *exp_ids;        // build mode: nq x exp_cap expanded node ids in expansion order
    float *exp_dists;         //             and their distances to the query node
    uint32_t *exp_cnt;        //             number recorded per query (<= exp_cap)
    uint32_t exp_cap, node_lo; // build mode: query i is base row node_lo + i
    unsigned long long *stats; // [0]=cmps [1]=hops [2]=adj_words [3]=first error word
    uint32_t *counter;        // work-queue ticket
    uint64_t tag_stride;      // bytes per slot (>= nd, 16-aligned)
    uint64_t seed;
    uint32_t nd, dim, adj_stride, ep;
    uint32_t nq, k, L, n_init;
    uint32_t rows_per_batch;  // R: rows staged per mbarrier round (multiple of 8, <= 32)
    uint32_t row_stride_smem; // floats; == dim padded so that (stride % 32) == 16 (bank-conflict-free float4 reads)
    uint32_t warp_smem_bytes; // bytes of dynamic smem per warp
    int metric_l2;            // 1: squared L2, 0: negative inner product
    const uint32_t *order;    // nq query indices in the order the tickets hand them out (nullptr: 0, 1, 2, ...)
    uint32_t cand_words;      // candidate-id scratch per warp (words)
    uint32_t lazy_queue;      // 1: long queue kept as sorted arrays + a pending block flushed every ~30 accepted candidates
    uint32_t drop_ep;         // 1: SearchRoarGraph — the entry point is not marked visited and may be evaluated again (9.1)
    uint32_t merge_insert;    // 1: the candidates of a row group enter the queue in one merge pass (RG_QUEUE_MERGE, default 1)
    // L2-resident exact visited set (SearchRoarGraph only; see rg_search.cu "visited table"): per slot 2^vt_bbits buckets
    // of 16 x 16-bit entries; 0 = use the tag arrays
    uint32_t *vtab;           // n_slots x vt_words
    uint32_t vt_words;        // 32-bit words per slot (8 per bucket)
    uint32_t vt_bbits, vt_rbits, vt_mask, vt_mul, vt_mulinv;
    uint32_t vt_smem;         // 1: the table lives in the warp's shared memory (vtab == nullptr, vt_words per warp)
};

struct SearchLaunch {
    uint32_t grid, block, smem_bytes, rows_per_batch, warps_per_cta, ctas_per_sm, warps_per_sm;
};

// Computes the launch geometry (and per-warp smem layout fields of `p`) for queue length L.
int plan_search(SearchParams &p, SearchLaunch &l, int sm_count, uint32_t max_slots);
int launch_roargraph(const SearchParams &p, const SearchLaunch &l, cudaStream_t s);
int launch_bipartite(const SearchParams &p, const SearchLaunch &l, cudaStream_t s);
int launch_build_search(const SearchParams &p, const SearchLaunch &l, cudaStream_t s);
// order[0..nq) <- the batch's query indices grouped by their nearest of kLandmarks base rows (scratch: 2 * nq + kLandmarks + 1 words)
constexpr uint32_t kLandmarks = 1024;
int order_queries(const float *base, uint64_t nd, uint32_t dim, int metric_l2, const float *queries, uint32_t nq, uint32_t *scratch,
                  cudaStream_t s);
uint32_t max_search_slots(int sm_count);

template <class T>
struct DevBuf {  // grow-only device scratch
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return RG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        RG_CUDA(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
        return RG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace rg

struct rg_index {
    int device = 0;      // the CUDA device the index lives on (the current device when it was created)
    int sm_count = 0;
    uint64_t nd = 0;
    uint32_t dim = 0;
    int metric = 0;
    const float *d_base = nullptr;
    bool own_base = false;
    // projection graph, fixed-stride rows [deg, ids...]
    uint32_t *d_adj = nullptr;
    uint32_t adj_stride = 0, ep = 0, max_deg = 0;
    bool has_graph = false;
    // bipartite graph, CSR
    uint64_t *d_row_off = nullptr;
    uint32_t *d_csr = nullptr;
    uint64_t n_total = 0;
    bool has_bipartite = false;
    // visited tags
    uint8_t *d_tags = nullptr;
    uint32_t *d_slot_epoch = nullptr;
    uint32_t n_slots = 0;
    uint64_t tag_stride = 0;
    uint32_t *d_vtab = nullptr;  // hashed visited tables (n_vt_slots x vt_words), all-ones = empty
    uint64_t vtab_words = 0;
    // per-call scratch
    uint32_t *d_counter = nullptr;
    unsigned long long *d_stats = nullptr;
    rg::DevBuf<float> q, dists;
    rg::DevBuf<uint32_t> ids, cmps, hops, init;
    rg::DevBuf<uint32_t> order_scratch;  // query ordering: keys[nq] | order[nq] | bins[kLandmarks + 1]
    rg::SearchLaunch last_launch{};
    uint32_t last_k = 0;  // k of the last search (for the "not enough results" text)
    cudaStream_t stream = nullptr;  // used by the host-buffer entry points
};

namespace rg {

int sm_count();             // SM count of the calling thread's current device (0 when there is none)
int ensure_device_ready();  // RG_OK when the current device is a usable sm_100 GPU
int enter_index(rg_index *idx);  // makes the index's device current on the calling thread
int ensure_tags(rg_index *idx, uint32_t slots);
int ensure_vtab(rg_index *idx, uint64_t words);
uint32_t affordable_slots(rg_index *idx);

// tensor-core candidate filter of the kNN (rg_knn_tc.cu)
void knn_release_workspace();
bool knn_tc_supported(uint32_t dim);
uint32_t knn_tc_pitch(uint32_t dim);
int knn_tc_convert(const float *x, uint64_t n, uint32_t dim, void *out_bf16, float *sq, cudaStream_t s);
int knn_tc_max_norm(const float *sq, uint64_t n, float *d_out, cudaStream_t s);
int knn_tc_filter(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t dim, uint64_t row_lo, uint64_t row_hi,
                  int metric_l2, const float *tau, const float *qn, const float *bn, float bmax, unsigned long long *cand,
                  uint32_t *cand_cnt, uint32_t cap, uint32_t *overflow, int sm_count, cudaStream_t s, uint32_t smem_reserve = 0);
uint64_t knn_tc_seed_rows(uint64_t nb, uint32_t K);
int knn_tc_seed(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t dim, uint64_t n_seed, uint32_t K,
                int metric_l2, const float *qn, const float *bn, float bmax, float *seed_buf, float *tau, int sm_count,
                cudaStream_t s, uint32_t smem_reserve = 0);

// device time per phase of the kNN calls of this thread since the last reset (filled only under RG_KNN_TRACE=1)
struct KnnProfile {
    double convert_base_ms = 0, convert_queries_ms = 0, seed_ms = 0, filter_ms = 0, merge_ms = 0;
    uint64_t chunks = 0, overflow_retries = 0;
};
const KnnProfile &knn_last_profile();
void knn_reset_profile();
void knn_set_trace(int on);

// BuildRoarGraph (rg_build.cu)
int build_roargraph_device(rg_index *idx, const uint32_t *d_knn, uint64_t nsq, uint32_t knn_stride, uint32_t M_sq,
                           uint32_t M, uint32_t L, uint32_t n_rounds, cudaStream_t s);
int adjacency_nnz_device(const uint32_t *d_adj, uint32_t stride, uint64_t nd, uint64_t *nnz);

// exact kNN (rg_knn.cu)
int knn_device(const float *d_base, uint64_t nb, const float *d_queries, uint64_t nq, uint32_t dim, uint32_t K,
               int metric, int32_t id_offset, int32_t *d_ids, float *d_dists, cudaStream_t s);
int knn_base_prepare(rg_knn_base *kb, const float *d_base, uint64_t nb, uint32_t dim, int metric, bool from_ws, cudaStream_t s);
// mid-shard re-sync of the shared bound (include/rg_b200.h, rg_knn_base_search_resync_dev)
struct KnnResync {
    uint32_t rank;      // every shard publishes its rank-th best exact distance so far, rank * G >= K
    uint64_t rows;      // at the first chunk boundary at or past this many rows
    float *d_bound;     // nq floats, caller-owned, device
    int (*hook)(void *user, uint64_t q_off, uint64_t n, void *stream);  // reduce d_bound[q_off, q_off + n) to the MAX over the shards
    void *user;
};
int knn_base_search(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset, int32_t *d_ids,
                    float *d_dists, cudaStream_t s, const float *d_tau_ext = nullptr, uint64_t bound_rows = 0,
                    const KnnResync *resync = nullptr);
int knn_base_seed(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, uint32_t rank, float *d_tau_out, cudaStream_t s);
void knn_base_free(rg_knn_base *kb);
// new object over device rows `d_base` (own_in: a device copy the object frees, or nullptr)
int knn_base_new(rg_knn_base **out, const float *d_base, float *own_in, uint64_t nb, uint32_t dim, int metric);
int topk_pack_device(const float *d_dists, const uint32_t *d_ids, uint64_t n, uint32_t id_offset, unsigned long long *d_packed,
                     cudaStream_t s);
int topk_merge_packed_device(const unsigned long long *d_packed, uint32_t G, uint64_t nq, uint32_t K_in, uint32_t K_out,
                             uint64_t q_lo, uint64_t q_hi, float *d_out_dists, uint32_t *d_out_ids, cudaStream_t s);
int topk_merge_device(const float *d_in_dists, const uint32_t *d_in_ids, uint32_t G, uint64_t nq, uint32_t K,
                      float *d_out_dists, uint32_t *d_out_ids, cudaStream_t s);

}  // namespace rg
