// test_only_build_bipartite — the reference's raw-bipartite-graph construction CLI (tests/test_only_build_bipartite.cpp
// there) on the B200 host mirror: same flags, same output file (IndexBipartite::Save: u32 npts, {u32 deg, u32 ids[deg]}).
// The graph is a pure re-arrangement of the training queries' kNN ids (qbaseNNbipartite, src/index_bipartite.cpp:235-280)
// — there is no distance work on this path, so it stays on the host and needs no GPU; the base / query files are opened
// for their headers only (the reference loads them to normalise and to compute an entry point that the saved file does
// not contain).  Consumer: test_search_bipartite.
#include <chrono>
#include <iostream>
#include <thread>

#include "cli_args.h"
#include "index_bipartite.h"

int main(int argc, char **argv) {
  cli::Parser args("Arguments");
  args.add("data_type", "data type <int8/uint8/float>", true)
      .add("dist", "distance function <l2/ip>", true)
      .add("base_data_path", "Input data file in bin format", true)
      .add("sampled_query_data_path", "Sampled query file in bin format", true)
      .add("bipartite_index_save_path", "Path prefix for saving bipartite index file components", true)
      .add("projection_index_save_path", "Path prefix for saving projetion index file components", true)
      .add("M_bp", "Number of neighbors for base points to build the bipartite graph", false, false, "32").numeric()
      .add("M_sq", "Number of neighbors for sampled query points to build the bipartite graph", false, false, "32").numeric()
      .add("M_pjbp", "Number of neighbors for projection graph", false, false, "32").numeric()
      .add("L_pjpq", "Priority queue length for projection graph searching", false, false, "32").numeric()
      .add("L_pq", "Priority queue length for searching", false, false, "32").numeric()
      .add("num_threads", "Number of host threads (kept for compatibility)", false, false,
           std::to_string(std::max(1u, std::thread::hardware_concurrency())), 'T').numeric()
      .add("learn_base_nn_path", "Path of learn-base NN file", true)
      .add("base_learn_nn_path", "Path of base-learn NN file", true);
  try {
    if (args.parse(argc, argv)) return 0;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  const std::string dist = args.str("dist"), out_file = args.str("bipartite_index_save_path");
  std::cout << "sampled query: " << args.str("sampled_query_data_path") << std::endl;
  try {
    uint32_t base_num, base_dim, sq_num, sq_dim;
    efanna2e::load_meta<float>(args.str("base_data_path").c_str(), base_num, base_dim);
    efanna2e::load_meta<float>(args.str("sampled_query_data_path").c_str(), sq_num, sq_dim);
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
    efanna2e::Parameters parameters;
    std::cout << "Index save path: " << args.str("projection_index_save_path") << std::endl;
    efanna2e::IndexBipartite index_bipartite(base_dim, (size_t)base_num + sq_num, dist_metric, nullptr);
    parameters.Set<uint32_t>("M_bp", (uint32_t)args.u64("M_bp"));
    parameters.Set<uint32_t>("M_sq", (uint32_t)args.u64("M_sq"));
    parameters.Set<uint32_t>("L_pq", (uint32_t)args.u64("L_pq"));
    parameters.Set<uint32_t>("M_pjbp", (uint32_t)args.u64("M_pjbp"));
    parameters.Set<uint32_t>("L_pjpq", (uint32_t)args.u64("L_pjpq"));
    parameters.Set<uint32_t>("num_threads", (uint32_t)args.u64("num_threads"));
    std::cout << "M_bp: " << args.u64("M_bp") << std::endl;
    index_bipartite.LoadLearnBaseKNN(args.str("learn_base_nn_path").c_str());
    auto s = std::chrono::high_resolution_clock::now();
    index_bipartite.BuildBipartite(sq_num, nullptr, base_num, nullptr, parameters);
    auto e = std::chrono::high_resolution_clock::now();
    std::cout << "indexing time: " << std::chrono::duration<double>(e - s).count() << "\n";
    index_bipartite.Save(out_file.c_str());
    std::cout << "Save index to " << out_file << std::endl;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  return 0;
}
