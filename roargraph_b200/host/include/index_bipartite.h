// efanna2e::IndexBipartite — host mirror of the reference's index object (include/index_bipartite.h:22-172) for the
// B200 build.  Same constructor, loaders, search entry points, parameters and error behaviour; the work happens in
// librg_b200.so (include/rg_b200.h), there is no CPU search path.
//
//   reference member (index_bipartite.h)                     here
//   IndexBipartite(dim, n, Metric, Index*)              :27  same; creates nothing on the device yet
//   LoadSearchNeededData / LoadVectorData               :62  fbin -> padded host rows (normalised for COSINE) -> HBM
//   LoadProjectionGraph                                 :105 index file -> CSR -> fixed-stride adjacency in HBM
//   Load / Save (bipartite graph)                       :29-30
//   InitVisitedListPool(num_threads)                    :133 accepted, no-op (visited tags live on the device)
//   SearchRoarGraph(query,k,qid,params,indices,dists)   :100 batch of one through SearchRoarGraphBatch
//   SearchBipartiteGraph(...)                           :137 batch of one through SearchBipartiteGraphBatch
//   LoadLearnBaseKNN                                    :109 reads the ids of the training queries' kNN file
//   BuildRoarGraph(n_sq, sq, n_bp, bp, params)          :38  device build (rg_index_build_roargraph)
//   BuildBipartite(n_sq, sq, n_bp, bp, params)          :36  host assembly of the raw bipartite graph from the kNN lists
//                                                            (qbaseNNbipartite, src/index_bipartite.cpp:235-280: pure id
//                                                            shuffling, no distances); uploaded on the first search
//   SaveProjectionGraph                                 :103 same file format
//   GetProjectionGraph                                  :113 host copy of the adjacency lists
//   need_normalize                                      :145
// New: SearchRoarGraphBatch / SearchBipartiteGraphBatch — the whole query batch crosses the boundary in one call,
// where the reference's CLI runs `#pragma omp parallel for` over SearchRoarGraph (tests/test_search_roargraph.cpp:203).
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "efanna2e/index.h"
#include "efanna2e/util.h"

struct rg_index;
struct rg_group;
struct rg_sharded_index;

namespace efanna2e {

class IndexBipartite : public Index {
  typedef std::vector<std::vector<uint32_t>> CompactGraph;

 public:
  explicit IndexBipartite(const size_t dimension, const size_t n, Metric m, Index *initializer);
  ~IndexBipartite() override;

  void Save(const char *filename) override;  // bipartite graph: u32 npts, {u32 deg, u32 ids[deg]} * npts
  void Load(const char *filename) override;
  void Search(const float *query, const float *x, size_t k, const Parameters &parameters, unsigned *indices,
              float *res_dists) override;
  void Build(size_t n, const float *data, const Parameters &parameters) override;

  void LoadVectorData(const char *base_file, const char *sampled_query_file);
  inline void LoadSearchNeededData(const char *base_file, const char *sampled_query_file) {
    LoadVectorData(base_file, sampled_query_file);
  }
  // --gpus N of the CLIs (call before any data is loaded): the base rows are sharded over the first N GPUs of the box by
  // equal counts (compute_groundtruth's base parts), every shard owns its OWN RoarGraph over its rows, and a search runs
  // every query on every shard and merges the per-shard top-k after one all-gather (include/rg_b200.h, rg_sharded_*).
  // Index files become one file per shard, "<path>.shard<s>of<N>", each a valid reference index over that shard's rows.
  void SetGpus(uint32_t n);
  uint32_t GetGpus() const { return gpus_; }
  static std::string ShardFile(const std::string &path, uint32_t shard, uint32_t n_shards);
  // Hand over base rows already in host memory (stride == GetDimension(), normalised for COSINE); copied to HBM.
  void SetBaseData(const float *base, size_t n);

  void LoadProjectionGraph(const char *filename);
  void SaveProjectionGraph(const char *filename);
  void LoadLearnBaseKNN(const char *filename);
  void InitVisitedListPool(uint32_t num_threads) { (void)num_threads; }

  void BuildRoarGraph(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data, const Parameters &parameters);
  // bp_data may be null when only the graph file is wanted (test_only_build_bipartite): the graph depends on the
  // training queries' kNN lists alone
  void BuildBipartite(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data, const Parameters &parameters);

  std::pair<uint32_t, uint32_t> SearchRoarGraph(const float *query, size_t k, size_t &qid, const Parameters &parameters,
                                                unsigned *indices, std::vector<float> &res_dists);
  std::pair<uint32_t, uint32_t> SearchBipartiteGraph(const float *query, size_t k, size_t &qid,
                                                     const Parameters &parameters, unsigned *indices,
                                                     std::vector<float> &dists);
  // queries: nq x GetDimension(); indices/dists: nq x k; cmps/hops: nq (may be null)
  void SearchRoarGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters, unsigned *indices,
                            float *dists, uint32_t *cmps, uint32_t *hops);
  // init_ids: nq x n_init entry points, or null to draw them on the host the way the reference does
  // (src/index_bipartite.cpp:287-294: one std::random_device word per query -> std::mt19937 ->
  // std::uniform_int_distribution<uint32_t>(0, nd - 1)); query i's word is (uint32_t)(seed + i), or an unseeded
  // std::random_device word when seed == kRandomDeviceSeed
  static constexpr uint64_t kRandomDeviceSeed = ~0ull;
  void DrawEntryPoints(size_t nq, uint32_t n_init, uint64_t seed, std::vector<uint32_t> &out) const;
  void SearchBipartiteGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters,
                                 const uint32_t *init_ids, uint32_t n_init, uint64_t seed, unsigned *indices, float *dists,
                                 uint32_t *cmps, uint32_t *hops);

  CompactGraph &GetProjectionGraph();
  CompactGraph &GetBipartiteGraph() { return bipartite_graph_; }
  uint32_t GetProjectionEp() const { return projection_ep_; }
  std::vector<std::vector<uint32_t>> &GetLearnBaseKNN() { return learn_base_knn_; }
  // totals of the last batch (for the CLI's extra roofline columns)
  void LastSearchStats(uint64_t &cmps, uint64_t &hops, uint64_t &adj_words) const;

  Index *initializer_;
  bool need_normalize = false;

 protected:
  void EnsureDeviceIndex();
  void UploadProjectionGraph();
  void UploadBipartiteGraph();
  [[noreturn]] void Fail(int rc) const;

  CompactGraph bipartite_graph_;
  CompactGraph projection_graph_;
  bool projection_host_valid_ = false;
  bool bipartite_uploaded_ = false;
  std::vector<std::vector<uint32_t>> learn_base_knn_;
  uint32_t projection_ep_ = 0;
  float *owned_base_ = nullptr;
  size_t dev_dim_ = 0;  // row stride on the device when it differs from GetDimension() (unpadded build input)
  rg_index *dev_ = nullptr;
  uint32_t gpus_ = 1;
  rg_group *group_ = nullptr;
  rg_sharded_index *sharded_ = nullptr;
  std::mutex call_mutex_;  // one caller at a time per rg_index
};

}  // namespace efanna2e
