// efanna2e::IndexBipartite over librg_b200.so.  File formats and error texts follow the reference
// (src/index_bipartite.cpp there): LoadProjectionGraph :2097-2117, SaveProjectionGraph :2606-2619, Save/Load :2045-2071,
// LoadLearnBaseKNN :2622-2642, LoadVectorData :2664-2695, "not enough results" :2408-2412.
#include "index_bipartite.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <stdexcept>

#include "rg_b200.h"

namespace efanna2e {

namespace {
int to_rg_metric(Metric m) {
  switch (m) {
    case L2: return RG_METRIC_L2;
    case COSINE: return RG_METRIC_COSINE;
    default: return RG_METRIC_INNER_PRODUCT;
  }
}

void read_adjacency(std::ifstream &in, uint32_t npts, std::vector<std::vector<uint32_t>> &g, double &deg_sum) {
  g.assign(npts, {});
  deg_sum = 0;
  for (uint32_t i = 0; i < npts; ++i) {
    uint32_t d = 0;
    in.read((char *)&d, 4);
    if (!in) throw std::runtime_error("index file truncated");
    g[i].resize(d);
    in.read((char *)g[i].data(), (std::streamsize)d * 4);
    deg_sum += d;
  }
}

void to_csr(const std::vector<std::vector<uint32_t>> &g, std::vector<uint64_t> &row_off, std::vector<uint32_t> &nbrs) {
  row_off.assign(g.size() + 1, 0);
  for (size_t i = 0; i < g.size(); ++i) row_off[i + 1] = row_off[i] + g[i].size();
  nbrs.resize(row_off.back());
  for (size_t i = 0; i < g.size(); ++i)
    if (!g[i].empty()) memcpy(nbrs.data() + row_off[i], g[i].data(), g[i].size() * 4);
}
}  // namespace

IndexBipartite::IndexBipartite(const size_t dimension, const size_t n, Metric m, Index *initializer)
    : Index(dimension, n, m), initializer_(initializer) {
  bipartite_ = true;
  if (m == COSINE) need_normalize = true;
}

IndexBipartite::~IndexBipartite() {
  if (dev_) rg_index_destroy(dev_);
  if (sharded_) rg_sharded_index_destroy(sharded_);
  if (group_) rg_group_destroy(group_);
  free(owned_base_);
}

void IndexBipartite::SetGpus(uint32_t n) {
  if (n == 0) throw std::invalid_argument("--gpus must be at least 1");
  if (dev_ || sharded_) throw std::runtime_error("SetGpus must be called before data is loaded");
  gpus_ = n;
}

std::string IndexBipartite::ShardFile(const std::string &path, uint32_t shard, uint32_t n_shards) {
  return path + ".shard" + std::to_string(shard) + "of" + std::to_string(n_shards);
}

void IndexBipartite::Fail(int rc) const {
  const std::string msg = rg_last_error();
  if (rc == RG_EINVAL) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

void IndexBipartite::EnsureDeviceIndex() {
  if (dev_ || sharded_) return;
  if (!data_bp_) throw std::runtime_error("base data not loaded");
  if (gpus_ > 1) {
    int rc = rg_group_create(&group_, (int)gpus_, nullptr);
    if (rc != RG_OK) Fail(rc);
    rc = rg_sharded_index_create(group_, &sharded_, data_bp_, nd_, (uint32_t)(dev_dim_ ? dev_dim_ : dimension_), to_rg_metric(metric_));
    if (rc != RG_OK) Fail(rc);
    return;
  }
  int rc = rg_index_create(&dev_, data_bp_, nd_, (uint32_t)(dev_dim_ ? dev_dim_ : dimension_), to_rg_metric(metric_), 0);
  if (rc != RG_OK) Fail(rc);
}

void IndexBipartite::SetBaseData(const float *base, size_t n) {
  if (dev_) {
    rg_index_destroy(dev_);
    dev_ = nullptr;
  }
  if (sharded_) {
    rg_sharded_index_destroy(sharded_);
    sharded_ = nullptr;
  }
  data_bp_ = base;
  nd_ = n;
  EnsureDeviceIndex();
}

// LoadVectorData (:2664-2695).  Single GPU: the base file streams straight into HBM (rg_index_create_from_fbin: 64 MB pinned
// slabs, the copy of slab i overlapping the read of slab i + 1, rows padded to x8 by the copy engine, COSINE rows normalised
// per slab by the reference's normalize on all host cores) — no 8-20 GB pageable host copy, GetBasePointSet() stays null.
// Sharded (--gpus > 1) or RG_HOST_BASE=1: the rows are loaded to host memory first, as the reference does.
static void normalize_hook(float *rows, uint64_t n, uint32_t dim, void *) { normalize_rows(rows, n, dim); }

void IndexBipartite::LoadVectorData(const char *base_file, const char *sampled_query_file) {
  uint32_t base_num = 0, base_dim = 0, sq_num = 0, q_dim = 0;
  load_meta<float>(base_file, base_num, base_dim);
  if (strlen(sampled_query_file) != 0) {
    load_meta<float>(sampled_query_file, sq_num, q_dim);
    if (base_dim != q_dim) throw std::runtime_error("base and query dimension mismatch");
  }
  const char *host_base = getenv("RG_HOST_BASE");
  if (gpus_ == 1 && !(host_base && host_base[0] == '1')) {
    if ((base_dim + 7u) / 8u * 8u != dimension_)
      throw std::runtime_error("base dimension (padded to x8) does not match the index dimension");
    if (need_normalize) std::cout << "Normalizing base data" << std::endl;
    if (dev_) {
      rg_index_destroy(dev_);
      dev_ = nullptr;
    }
    uint64_t nd = 0;
    uint32_t pdim = 0;
    int rc = rg_index_create_from_fbin(&dev_, base_file, to_rg_metric(metric_), need_normalize ? normalize_hook : nullptr, nullptr,
                                       &nd, &pdim);
    if (rc != RG_OK) {
      const std::string msg = rg_last_error();
      if (msg == "Data file size wrong!") {  // util.h:124
        std::cerr << "Read file incompleted! filename:" << base_file << std::endl;
        throw std::runtime_error(msg);
      }
      Fail(rc);
    }
    std::cout << "new_dim: " << pdim << std::endl;
    free(owned_base_);
    owned_base_ = nullptr;
    data_bp_ = nullptr;
    nd_ = nd;
    nd_sq_ = sq_num;
    return;
  }
  float *base = nullptr;
  load_data<float>(base_file, base_num, base_dim, base);
  unsigned padded = base_dim;
  base = data_align(base, base_num, padded);
  if (padded != dimension_) {
    free(base);
    throw std::runtime_error("base dimension (padded to x8) does not match the index dimension");
  }
  if (need_normalize) {
    std::cout << "Normalizing base data" << std::endl;
    normalize_rows(base, base_num, padded);
  }
  free(owned_base_);
  owned_base_ = base;
  nd_sq_ = sq_num;
  SetBaseData(base, base_num);
}

void IndexBipartite::UploadProjectionGraph() {
  EnsureDeviceIndex();
  std::vector<uint64_t> row_off;
  std::vector<uint32_t> nbrs;
  if (projection_graph_.size() != nd_) throw std::runtime_error("projection graph size does not match the base size");
  to_csr(projection_graph_, row_off, nbrs);
  int rc = rg_index_set_graph(dev_, projection_ep_, row_off.data(), nbrs.data());
  if (rc != RG_OK) Fail(rc);
  projection_host_valid_ = true;
}

void IndexBipartite::LoadProjectionGraph(const char *filename) {
  if (gpus_ > 1) {  // one index file per shard, local ids
    EnsureDeviceIndex();
    for (uint32_t s = 0; s < gpus_; ++s) {
      const std::string f = ShardFile(filename, s, gpus_);
      std::ifstream in(f, std::ios::binary);
      if (!in.is_open()) throw std::runtime_error("cannot open index file " + f);
      uint32_t ep = 0, npts = 0;
      in.read((char *)&ep, 4);
      in.read((char *)&npts, 4);
      CompactGraph g;
      double deg = 0;
      read_adjacency(in, npts, g, deg);
      rg_index *shard = nullptr;
      uint64_t lo = 0, rows = 0;
      int rc = rg_sharded_index_shard(sharded_, (int)s, &shard, &lo, &rows);
      if (rc != RG_OK) Fail(rc);
      if (npts != rows) throw std::runtime_error("index file " + f + " does not match the shard's rows");
      std::cout << "Projection graph shard " << s << " (rows " << lo << ".." << lo + rows << "), ep: " << ep
                << ", avg_degree: " << deg / npts << std::endl;
      std::vector<uint64_t> row_off;
      std::vector<uint32_t> nbrs;
      to_csr(g, row_off, nbrs);
      rc = rg_index_set_graph(shard, ep, row_off.data(), nbrs.data());
      if (rc != RG_OK) Fail(rc);
    }
    return;
  }
  // one read of the whole file, decoded straight into CSR (10 M small vectors cost more than the search sweep); the host
  // adjacency lists are materialised from the device copy only if GetProjectionGraph() is called
  std::ifstream in(filename, std::ios::binary | std::ios::ate);
  if (!in.is_open()) throw std::runtime_error(std::string("cannot open index file ") + filename);
  const size_t fsize = (size_t)in.tellg();
  if (fsize < 8 || fsize % 4) throw std::runtime_error("index file truncated");
  std::vector<uint32_t> words(fsize / 4);
  in.seekg(0, std::ios::beg);
  in.read((char *)words.data(), (std::streamsize)fsize);
  if (!in) throw std::runtime_error("index file truncated");
  projection_ep_ = words[0];
  const uint32_t npts = words[1];
  std::cout << "Projection graph, ep: " << projection_ep_ << std::endl;
  if (npts != nd_) throw std::runtime_error("projection graph size does not match the base size");
  std::vector<uint64_t> row_off((size_t)npts + 1, 0);
  std::vector<uint32_t> nbrs;
  nbrs.reserve(words.size() - 2 - npts);
  size_t pos = 2;
  for (uint32_t i = 0; i < npts; ++i) {
    if (pos >= words.size()) throw std::runtime_error("index file truncated");
    const uint32_t d = words[pos++];
    if (pos + d > words.size()) throw std::runtime_error("index file truncated");
    nbrs.insert(nbrs.end(), words.begin() + pos, words.begin() + pos + d);
    pos += d;
    row_off[i + 1] = nbrs.size();
  }
  std::cout << "Projection graph, avg_degree: " << (double)nbrs.size() / npts << std::endl;
  EnsureDeviceIndex();
  int rc = rg_index_set_graph(dev_, projection_ep_, row_off.data(), nbrs.data());
  if (rc != RG_OK) Fail(rc);
  projection_graph_.clear();
  projection_host_valid_ = false;
}

IndexBipartite::CompactGraph &IndexBipartite::GetProjectionGraph() {
  if (!projection_host_valid_ && dev_) {
    uint64_t nnz = 0;
    uint32_t ep = 0, md = 0;
    int rc = rg_index_graph_size(dev_, &nnz, &ep, &md);
    if (rc == RG_OK) {
      std::vector<uint64_t> row_off(nd_ + 1);
      std::vector<uint32_t> nbrs(nnz);
      rc = rg_index_get_graph(dev_, row_off.data(), nbrs.data());
      if (rc != RG_OK) Fail(rc);
      projection_graph_.assign(nd_, {});
      for (size_t i = 0; i < nd_; ++i)
        projection_graph_[i].assign(nbrs.begin() + row_off[i], nbrs.begin() + row_off[i + 1]);
      projection_ep_ = ep;
      projection_host_valid_ = true;
    }
  }
  return projection_graph_;
}

void IndexBipartite::SaveProjectionGraph(const char *filename) {
  if (gpus_ > 1) {  // one file per shard in the reference's format (:2606-2619), local ids
    if (!sharded_) throw std::runtime_error("index not built");
    for (uint32_t s = 0; s < gpus_; ++s) {
      rg_index *shard = nullptr;
      uint64_t lo = 0, rows = 0, nnz = 0;
      uint32_t ep = 0, md = 0;
      int rc = rg_sharded_index_shard(sharded_, (int)s, &shard, &lo, &rows);
      if (rc == RG_OK) rc = rg_index_graph_size(shard, &nnz, &ep, &md);
      if (rc != RG_OK) Fail(rc);
      std::vector<uint64_t> row_off(rows + 1);
      std::vector<uint32_t> nbrs(nnz);
      rc = rg_index_get_graph(shard, row_off.data(), nbrs.data());
      if (rc != RG_OK) Fail(rc);
      std::ofstream out(ShardFile(filename, s, gpus_), std::ios::binary | std::ios::out);
      if (!out.is_open()) throw std::runtime_error("cannot open file");
      const uint32_t npts = (uint32_t)rows;
      out.write((const char *)&ep, 4);
      out.write((const char *)&npts, 4);
      for (uint64_t i = 0; i < rows; ++i) {
        const uint32_t d = (uint32_t)(row_off[i + 1] - row_off[i]);
        out.write((const char *)&d, 4);
        out.write((const char *)(nbrs.data() + row_off[i]), (std::streamsize)d * 4);
      }
    }
    return;
  }
  CompactGraph &g = GetProjectionGraph();
  std::ofstream out(filename, std::ios::binary | std::ios::out);
  if (!out.is_open()) throw std::runtime_error("cannot open file");
  const uint32_t npts = (uint32_t)g.size();
  out.write((const char *)&projection_ep_, 4);
  out.write((const char *)&npts, 4);
  for (uint32_t i = 0; i < npts; ++i) {
    const uint32_t d = (uint32_t)g[i].size();
    out.write((const char *)&d, 4);
    out.write((const char *)g[i].data(), (std::streamsize)d * 4);
  }
}

void IndexBipartite::Save(const char *filename) {
  std::ofstream out(filename, std::ios::binary | std::ios::out);
  if (!out.is_open()) throw std::runtime_error("cannot open file");
  const uint32_t npts = (uint32_t)bipartite_graph_.size();
  out.write((const char *)&npts, 4);
  for (uint32_t i = 0; i < npts; ++i) {
    const uint32_t d = (uint32_t)bipartite_graph_[i].size();
    out.write((const char *)&d, 4);
    out.write((const char *)bipartite_graph_[i].data(), (std::streamsize)d * 4);
  }
}

void IndexBipartite::Load(const char *filename) {
  std::ifstream in(filename, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error(std::string("cannot open index file ") + filename);
  uint32_t npts = 0;
  in.read((char *)&npts, 4);
  double deg = 0;
  read_adjacency(in, npts, bipartite_graph_, deg);
  bipartite_uploaded_ = false;
  if (data_bp_ || dev_) UploadBipartiteGraph();
}

void IndexBipartite::UploadBipartiteGraph() {
  EnsureDeviceIndex();
  if (bipartite_graph_.size() < nd_) throw std::runtime_error("bipartite graph has fewer nodes than the base");
  std::vector<uint64_t> row_off;
  std::vector<uint32_t> nbrs;
  to_csr(bipartite_graph_, row_off, nbrs);
  int rc = rg_index_set_bipartite(dev_, bipartite_graph_.size(), row_off.data(), nbrs.data());
  if (rc != RG_OK) Fail(rc);
  bipartite_uploaded_ = true;
}

// BuildBipartite (:42-141) = qbaseNNbipartite (:235-280) as far as the saved graph goes: training query sq becomes node
// nd + sq; its edges are its first M_pjbp nearest base rows except every occurrence of the top-1 (:262-269), and the
// top-1 base row gets the reverse edge to nd + sq (:270-273).  The reference appends the reverse edges from OpenMP threads
// under per-node locks, so their order within a base row depends on thread timing; here it is ascending sq, which is what
// the reference produces with -T 1.  (CalculateProjectionep, :92, has no effect on the bipartite graph or its file.)
void IndexBipartite::BuildBipartite(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data,
                                    const Parameters &parameters) {
  std::cout << "start build bipartite index" << std::endl;
  (void)sq_data;
  const uint32_t M_pjbp = parameters.Get<uint32_t>("M_pjbp");
  if (learn_base_knn_.size() < n_sq) throw std::runtime_error("LoadLearnBaseKNN must be called before BuildBipartite");
  if (n_bp + n_sq > 0xffffffffull) throw std::invalid_argument("more than 2^32 nodes");
  if (need_normalize && bp_data) {
    std::cout << "normalizing base data" << std::endl;
    normalize_rows(const_cast<float *>(bp_data), n_bp, dimension_);
  }
  nd_sq_ = n_sq;
  // the graph is assembled on the host; the base goes to HBM with the first search (a build-only run needs no GPU)
  if (dev_) {
    rg_index_destroy(dev_);
    dev_ = nullptr;
  }
  data_bp_ = bp_data;
  nd_ = n_bp;
  const uint32_t u32_nd = (uint32_t)n_bp;
  bipartite_graph_.assign(n_bp + n_sq, {});
  for (size_t sq = 0; sq < n_sq; ++sq) {
    std::vector<uint32_t> &nn_base = learn_base_knn_[sq];
    if (nn_base.size() > M_pjbp) {
      nn_base.resize(M_pjbp);
      nn_base.shrink_to_fit();
    }
    if (nn_base.empty()) continue;
    const uint32_t cur_tgt = nn_base[0];
    if (cur_tgt >= u32_nd) throw std::runtime_error("learn-base kNN id out of range");
    std::vector<uint32_t> &row = bipartite_graph_[n_bp + sq];
    for (uint32_t id : nn_base) {
      if (id == cur_tgt) continue;
      if (id >= u32_nd) throw std::runtime_error("learn-base kNN id out of range");
      row.push_back(id);
    }
    bipartite_graph_[cur_tgt].push_back((uint32_t)sq + u32_nd);
  }
  std::cout << "100% of save bipartite graph finish" << std::endl;
  double deg_b = 0, deg_q = 0;
  size_t max_b = 0, min_b = (size_t)-1, max_q = 0, min_q = (size_t)-1;
  for (size_t i = 0; i < bipartite_graph_.size(); ++i) {
    const size_t d = bipartite_graph_[i].size();
    if (i < n_bp) {
      deg_b += d;
      max_b = std::max(max_b, d);
      min_b = std::min(min_b, d);
    } else {
      deg_q += d;
      max_q = std::max(max_q, d);
      min_q = std::min(min_q, d);
    }
  }
  std::cout << "Bipartite nd degree avg: " << (n_bp ? deg_b / n_bp : 0) << std::endl;
  std::cout << "Bipartite nd degree max: " << max_b << std::endl;
  std::cout << "Bipartite nd degree min: " << (n_bp ? min_b : 0) << std::endl;
  std::cout << "Bipartite nd_sq degree avg: " << (n_sq ? deg_q / n_sq : 0) << std::endl;
  std::cout << "Bipartite nd_sq degree max: " << max_q << std::endl;
  std::cout << "Bipartite nd_sq degree min: " << (n_sq ? min_q : 0) << std::endl;
  bipartite_uploaded_ = false;
  has_built = true;
}

void IndexBipartite::LoadLearnBaseKNN(const char *filename) {
  std::ifstream in(filename, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error("cannot open file");
  uint32_t npts = 0, k_dim = 0;
  in.read((char *)&npts, 4);
  in.read((char *)&k_dim, 4);
  std::cout << "learn base knn npts: " << npts << ", k_dim: " << k_dim << std::endl;
  learn_base_knn_.assign(npts, {});
  std::vector<uint32_t> block((size_t)npts * k_dim);
  in.read((char *)block.data(), (std::streamsize)block.size() * 4);
  if (!in) throw std::runtime_error("learn-base kNN file truncated");
  for (uint32_t i = 0; i < npts; ++i) learn_base_knn_[i].assign(block.begin() + (size_t)i * k_dim, block.begin() + (size_t)(i + 1) * k_dim);
}

void IndexBipartite::BuildRoarGraph(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data,
                                    const Parameters &parameters) {
  std::cout << "start build bipartite index" << std::endl;
  (void)sq_data;  // the reference never reads the training vectors either, only their kNN lists (:1063)
  const uint32_t M_sq = parameters.Get<uint32_t>("M_sq");
  const uint32_t M_pjbp = parameters.Get<uint32_t>("M_pjbp");
  const uint32_t L_pjpq = parameters.Get<uint32_t>("L_pjpq");
  if (learn_base_knn_.empty() && gpus_ == 1) throw std::runtime_error("LoadLearnBaseKNN must be called before BuildRoarGraph");
  if (need_normalize) {
    std::cout << "normalizing base data" << std::endl;
    normalize_rows(const_cast<float *>(bp_data), n_bp, dimension_);
  }
  nd_sq_ = n_sq;
  if (dimension_ % 8 != 0) {
    // the reference's build CLI hands over raw fbin rows (tests/test_build_roargraph.cpp:105-107, data_align commented
    // out); the device layout wants rows padded to x8 floats with zeros, which changes no distance
    const size_t pd = (dimension_ + 7) / 8 * 8;
    float *padded = (float *)calloc(n_bp * pd, sizeof(float));
    if (!padded) throw std::runtime_error("out of host memory padding the base rows");
    for (size_t i = 0; i < n_bp; ++i) memcpy(padded + i * pd, bp_data + i * dimension_, dimension_ * sizeof(float));
    free(owned_base_);
    owned_base_ = padded;
    dev_dim_ = pd;
    bp_data = padded;
  }
  SetBaseData(bp_data, n_bp);
  if (gpus_ > 1) {
    // every shard needs the training queries' neighbours INSIDE the shard: they are computed on the shard's GPU from the
    // training vectors (the learn-base kNN file holds the neighbours over the whole base, which is not what a shard needs)
    if (!sq_data) throw std::runtime_error("the sharded build (--gpus > 1) needs the training-query vectors");
    std::cout << "begin link projection on " << gpus_ << " shards" << std::endl;
    int rc = rg_sharded_index_build_roargraph(sharded_, sq_data, n_sq, M_sq, M_pjbp, L_pjpq);
    if (rc != RG_OK) Fail(rc);
    for (uint32_t s = 0; s < gpus_; ++s) {
      rg_index *shard = nullptr;
      uint64_t lo = 0, rows = 0, nnz = 0;
      uint32_t ep = 0, md = 0;
      rg_sharded_index_shard(sharded_, (int)s, &shard, &lo, &rows);
      rg_index_graph_size(shard, &nnz, &ep, &md);
      std::cout << "shard " << s << ": projection ep: " << ep << ", Projection degree avg: " << (double)nnz / (double)rows
                << ", Projection degree max: " << md << std::endl;
    }
    has_built = true;
    return;
  }
  const size_t nsq = learn_base_knn_.size();
  size_t stride = 0;
  for (auto &r : learn_base_knn_) stride = std::max(stride, r.size());
  std::vector<uint32_t> flat(nsq * stride, 0xffffffffu);
  for (size_t i = 0; i < nsq; ++i) memcpy(flat.data() + i * stride, learn_base_knn_[i].data(), learn_base_knn_[i].size() * 4);
  std::cout << "begin link projection" << std::endl;
  int rc = rg_index_build_roargraph(dev_, flat.data(), 0, nsq, (uint32_t)stride, std::min<uint32_t>(M_sq, (uint32_t)stride),
                                    M_pjbp, L_pjpq, parameters.Get<uint32_t>("build_rounds", 0u));
  if (rc != RG_OK) Fail(rc);
  projection_host_valid_ = false;
  uint64_t nnz = 0;
  uint32_t md = 0;
  rg_index_graph_size(dev_, &nnz, &projection_ep_, &md);
  std::cout << "projection ep: " << projection_ep_ << std::endl;
  std::cout << "total degree: " << nnz << std::endl;
  std::cout << "Projection degree avg: " << (double)nnz / (double)n_bp << std::endl;
  std::cout << "Projection degree max: " << md << std::endl;
  has_built = true;
}

void IndexBipartite::Build(size_t n, const float *data, const Parameters &parameters) {
  BuildRoarGraph(nd_sq_, nullptr, n, data, parameters);
}

void IndexBipartite::SearchRoarGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters,
                                          unsigned *indices, float *dists, uint32_t *cmps, uint32_t *hops) {
  const uint32_t L_pq = parameters.Get<uint32_t>("L_pq");
  // an rg_index takes one caller at a time; the reference's CLIs call SearchRoarGraph from an OpenMP loop on one shared
  // index (tests/test_search_roargraph.cpp:203), so concurrent callers queue here
  std::lock_guard<std::mutex> guard(call_mutex_);
  if (sharded_) {  // per-query cmps / hops are per-shard quantities: LastSearchStats sums them over the shards
    int rc = rg_sharded_search_roargraph(sharded_, queries, (uint32_t)nq, (uint32_t)k, L_pq, indices, dists);
    if (rc != RG_OK) Fail(rc);
    if (cmps) std::fill(cmps, cmps + nq, 0u);
    if (hops) std::fill(hops, hops + nq, 0u);
    return;
  }
  if (!dev_) throw std::runtime_error("index not loaded");
  int rc = rg_search_roargraph(dev_, queries, (uint32_t)nq, (uint32_t)k, L_pq, indices, dists, cmps, hops);
  if (rc != RG_OK) Fail(rc);
}

std::pair<uint32_t, uint32_t> IndexBipartite::SearchRoarGraph(const float *query, size_t k, size_t &qid,
                                                              const Parameters &parameters, unsigned *indices,
                                                              std::vector<float> &res_dists) {
  (void)qid;
  uint32_t cmps = 0, hops = 0;
  res_dists.resize(k);
  SearchRoarGraphBatch(query, 1, k, parameters, indices, res_dists.data(), &cmps, &hops);
  return {cmps, hops};
}

// Entry points the way the reference draws them (src/index_bipartite.cpp:287-294): ONE std::random_device word per query
// seeds a std::mt19937 whose first n_init draws through std::uniform_int_distribution<uint32_t>(0, nd - 1) are the ids.
// seed_word(i) = the word query i uses: an unseeded std::random_device like the reference, or (uint32_t)(seed + i) so
// that a run is reproducible — and comparable id for id with the reference compiled against a fixed random_device.
void IndexBipartite::DrawEntryPoints(size_t nq, uint32_t n_init, uint64_t seed, std::vector<uint32_t> &out) const {
  out.resize(nq * n_init);
  const uint32_t u32_nd = (uint32_t)nd_;
  for (size_t i = 0; i < nq; ++i) {
    uint32_t word;
    if (seed == kRandomDeviceSeed) {
      std::random_device rd;
      word = rd();
    } else {
      word = (uint32_t)(seed + i);
    }
    std::mt19937 gen(word);
    std::uniform_int_distribution<uint32_t> dis(0, u32_nd - 1);
    for (uint32_t j = 0; j < n_init; ++j) out[i * n_init + j] = dis(gen);
  }
}

void IndexBipartite::SearchBipartiteGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters,
                                               const uint32_t *init_ids, uint32_t n_init, uint64_t seed,
                                               unsigned *indices, float *dists, uint32_t *cmps, uint32_t *hops) {
  const uint32_t L_pq = parameters.Get<uint32_t>("L_pq");
  std::lock_guard<std::mutex> guard(call_mutex_);
  if (!bipartite_uploaded_ && !bipartite_graph_.empty() && (data_bp_ || dev_)) UploadBipartiteGraph();
  if (!dev_) throw std::runtime_error("index not loaded");
  std::vector<uint32_t> drawn;
  if (!init_ids) {
    DrawEntryPoints(nq, n_init, seed, drawn);
    init_ids = drawn.data();
  }
  int rc = rg_search_bipartite(dev_, queries, (uint32_t)nq, (uint32_t)k, L_pq, init_ids, n_init, 0, indices, dists, cmps, hops);
  if (rc != RG_OK) Fail(rc);
}

std::pair<uint32_t, uint32_t> IndexBipartite::SearchBipartiteGraph(const float *query, size_t k, size_t &qid,
                                                                   const Parameters &parameters, unsigned *indices,
                                                                   std::vector<float> &dists) {
  (void)qid;  // unused by the reference as well
  uint32_t cmps = 0, hops = 0;
  dists.resize(k);
  // the drop-in single-query entry draws from std::random_device exactly like the reference (:287-294)
  SearchBipartiteGraphBatch(query, 1, k, parameters, nullptr, 10, kRandomDeviceSeed, indices, dists.data(), &cmps, &hops);
  return {cmps, hops};
}

void IndexBipartite::Search(const float *query, const float *x, size_t k, const Parameters &parameters, unsigned *indices,
                            float *res_dists) {
  (void)x;  // the Index virtual has no call site in the reference either ("dead/legacy search")
  SearchRoarGraphBatch(query, 1, k, parameters, indices, res_dists, nullptr, nullptr);
}

void IndexBipartite::LastSearchStats(uint64_t &cmps, uint64_t &hops, uint64_t &adj_words) const {
  cmps = hops = adj_words = 0;
  if (sharded_) {
    for (uint32_t s = 0; s < gpus_; ++s) {
      rg_index *shard = nullptr;
      uint64_t c = 0, h = 0, a = 0;
      if (rg_sharded_index_shard(sharded_, (int)s, &shard, nullptr, nullptr) == RG_OK &&
          rg_search_last_stats(shard, &c, &h, &a) == RG_OK) {
        cmps += c;
        hops += h;
        adj_words += a;
      }
    }
    return;
  }
  if (dev_) rg_search_last_stats(dev_, &cmps, &hops, &adj_words);
}

}  // namespace efanna2e
