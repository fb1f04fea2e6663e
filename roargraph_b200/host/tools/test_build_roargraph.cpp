// test_build_roargraph — the reference's index-construction CLI (tests/test_build_roargraph.cpp there) on the B200
// library: same flags (--data_type --dist --base_data_path --sampled_query_data_path --projection_index_save_path
// --learn_base_nn_path --M_sq --M_pjbp --L_pjpq -T), same output file format (SaveProjectionGraph).
#include <chrono>
#include <iostream>
#include <thread>

#include "cli_args.h"
#include "index_bipartite.h"

int main(int argc, char **argv) {
  cli::Parser args("Arguments");
  args.add("data_type", "data type <int8/uint8/float>", true)
      .add("dist", "distance function <l2/ip/cosine>", true)
      .add("base_data_path", "Input data file in bin format", true)
      .add("sampled_query_data_path", "Sampled query file in bin format", true)
      .add("projection_index_save_path", "Path prefix for saving projetion index file components", true)
      .add("learn_base_nn_path", "Path of learn-base NN file", true)
      .add("M_sq", "Number of neighbors for sampled query points to build the bipartite graph", false, false, "32").numeric()
      .add("M_pjbp", "Number of neighbors for projection graph", false, false, "32").numeric()
      .add("L_pjpq", "Priority queue length for projection graph searching", false, false, "32").numeric()
      .add("gpus", "GPUs of this box the base rows are sharded over: one index per shard, files <path>.shard<s>of<N>", false, false, "1").numeric()
      .add("build_rounds", "bulk-synchronous rounds of the connectivity phase (0 = automatic)", false, false, "0").numeric()
      .add("num_threads", "Number of host threads (kept for compatibility; the build runs on the GPU)", false, false,
           std::to_string(std::max(1u, std::thread::hardware_concurrency())), 'T').numeric();
  try {
    if (args.parse(argc, argv)) return 0;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  const std::string dist = args.str("dist");
  std::cout << "sampled query: " << args.str("sampled_query_data_path") << std::endl;
  try {
    uint32_t base_num, base_dim, sq_num, sq_dim;
    efanna2e::load_meta<float>(args.str("base_data_path").c_str(), base_num, base_dim);
    efanna2e::load_meta<float>(args.str("sampled_query_data_path").c_str(), sq_num, sq_dim);
    efanna2e::Parameters parameters;
    parameters.Set<uint32_t>("M_sq", (uint32_t)args.u64("M_sq"));
    parameters.Set<uint32_t>("M_pjbp", (uint32_t)args.u64("M_pjbp"));
    parameters.Set<uint32_t>("L_pjpq", (uint32_t)args.u64("L_pjpq"));
    parameters.Set<uint32_t>("num_threads", (uint32_t)args.u64("num_threads"));
    parameters.Set<uint32_t>("build_rounds", (uint32_t)args.u64("build_rounds"));
    float *data_bp = nullptr;
    efanna2e::load_data<float>(args.str("base_data_path").c_str(), base_num, base_dim, data_bp);
    unsigned padded = base_dim;
    data_bp = efanna2e::data_align(data_bp, base_num, padded);
    efanna2e::Metric dist_metric = efanna2e::INNER_PRODUCT;
    if (dist == "l2") {
      dist_metric = efanna2e::L2;
      std::cout << "Using l2 as distance metric" << std::endl;
    } else if (dist == "ip") {
      std::cout << "Using inner product as distance metric" << std::endl;
    } else if (dist == "cosine") {
      dist_metric = efanna2e::COSINE;
      std::cout << "Using cosine as distance metric" << std::endl;
    } else {
      std::cout << "Unknown distance type: " << dist << std::endl;
      return -1;
    }
    efanna2e::IndexBipartite index(padded, base_num, dist_metric, nullptr);
    const uint32_t gpus = (uint32_t)args.u64("gpus");
    index.SetGpus(gpus);
    float *data_sq = nullptr;
    if (gpus > 1) {
      // a shard's graph is built from the training queries' neighbours INSIDE the shard, computed on its GPU from the
      // training vectors; the learn-base kNN file (neighbours over the whole base) is not read
      efanna2e::load_data<float>(args.str("sampled_query_data_path").c_str(), sq_num, sq_dim, data_sq);
      unsigned sq_padded = sq_dim;
      data_sq = efanna2e::data_align(data_sq, sq_num, sq_padded);
      if (sq_padded != padded) throw std::runtime_error("base and query dimension mismatch");
      if (dist_metric == efanna2e::COSINE) efanna2e::normalize_rows(data_sq, sq_num, sq_padded);
    } else {
      index.LoadLearnBaseKNN(args.str("learn_base_nn_path").c_str());
    }
    auto s = std::chrono::high_resolution_clock::now();
    index.BuildRoarGraph(sq_num, data_sq, base_num, data_bp, parameters);
    auto e = std::chrono::high_resolution_clock::now();
    std::cout << "indexing time: " << std::chrono::duration<double>(e - s).count() << "\n";
    index.SaveProjectionGraph(args.str("projection_index_save_path").c_str());
    free(data_bp);
    free(data_sq);
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  return 0;
}
