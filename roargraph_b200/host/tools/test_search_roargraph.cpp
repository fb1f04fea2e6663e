// test_search_roargraph — the reference's QPS/recall benchmark CLI (tests/test_search_roargraph.cpp there) re-hosted on
// the B200 library: same flags, same stdout table and CSV columns (L_pq,qps,avg_cmps,ms_per_query,recall,avg_hops),
// one batched device call per L_pq where the reference loops `#pragma omp parallel for` over SearchRoarGraph (:203-209).
// Extra CSV/stdout columns after the reference's: gather GB/s (algorithmic bytes of SURVEY.md 8(d) / wall time).
#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <thread>

#include "cli_args.h"
#include "index_bipartite.h"

// tests/test_search_roargraph.cpp:23-36 of the reference: |gt[:k] ∩ res[:k]| summed over queries / (k * q_num)
static float ComputeRecall(uint32_t q_num, uint32_t k, uint32_t gt_dim, const uint32_t *res, const uint32_t *gt) {
  uint64_t hit = 0;
  for (uint32_t q = 0; q < q_num; ++q)
    for (uint32_t i = 0; i < k; ++i) {
      const uint32_t want = gt[(size_t)q * gt_dim + i];
      for (uint32_t j = 0; j < k; ++j)
        if (res[(size_t)q * k + j] == want) { ++hit; break; }
    }
  return (float)hit / (float)((uint64_t)k * q_num);
}

int main(int argc, char **argv) {
  cli::Parser args("Arguments");
  args.add("data_type", "data type <int8/uint8/float>", true)
      .add("dist", "distance function <l2/ip/cosine>", true)
      .add("base_data_path", "Input data file in bin format", true)
      .add("query_path", "Query file in bin format", true)
      .add("gt_path", "Groundtruth file in bin format", true)
      .add("projection_index_save_path", "Path prefix for saving projetion index file components", true)
      .add("L_pq", "Priority queue length for searching", true, true).numeric()
      .add("k", "k nearest neighbors", true, false, "1").numeric()
      .add("evaluation_save_path", "Path prefix for saving evaluation results", false)
      .add("gpus", "GPUs of this box the base rows are sharded over (per-shard index files <path>.shard<s>of<N>)", false, false, "1").numeric()
      .add("num_threads", "Number of host threads (kept for compatibility; the search runs on the GPU)", false, false,
           std::to_string(std::max(1u, std::thread::hardware_concurrency())), 'T').numeric();
  try {
    if (args.parse(argc, argv)) return 0;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  const std::string base_data_file = args.str("base_data_path"), query_file = args.str("query_path"),
                    gt_file = args.str("gt_path"), index_file = args.str("projection_index_save_path"),
                    dist = args.str("dist"), eval_path = args.str("evaluation_save_path");
  const std::vector<uint32_t> L_vec = args.u32_list("L_pq");
  const uint32_t k = (uint32_t)args.u64("k"), num_threads = (uint32_t)args.u64("num_threads");

  try {
    uint32_t base_num, base_dim;
    efanna2e::load_meta<float>(base_data_file.c_str(), base_num, base_dim);
    efanna2e::Parameters parameters;
    parameters.Set<uint32_t>("num_threads", num_threads);
    uint32_t q_pts, q_dim;
    efanna2e::load_meta<float>(query_file.c_str(), q_pts, q_dim);
    float *query_data = nullptr;
    efanna2e::load_data<float>(query_file.c_str(), q_pts, q_dim, query_data);
    float *aligned_query_data = efanna2e::data_align(query_data, q_pts, q_dim);

    uint32_t gt_pts, gt_dim;
    uint32_t *gt_ids = nullptr;
    float *gt_dists = nullptr;
    efanna2e::load_gt_meta<uint32_t>(gt_file.c_str(), gt_pts, gt_dim);
    efanna2e::load_gt_data_with_dist<uint32_t, float>(gt_file.c_str(), gt_pts, gt_dim, gt_ids, gt_dists);
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
    const uint32_t gpus = (uint32_t)args.u64("gpus");
    if (!std::filesystem::exists(gpus > 1 ? efanna2e::IndexBipartite::ShardFile(index_file, 0, gpus) : index_file)) {
      std::cout << "projection index file does not exist." << std::endl;
      return -1;
    }
    efanna2e::IndexBipartite index(q_dim, base_num, dist_metric, nullptr);
    index.SetGpus(gpus);
    index.LoadSearchNeededData(base_data_file.c_str(), "");
    std::cout << "Load graph index: " << index_file << std::endl;
    index.LoadProjectionGraph(index_file.c_str());
    if (index.need_normalize) {
      std::cout << "Normalizing query data" << std::endl;
      for (uint32_t i = 0; i < q_pts; i++) efanna2e::normalize<float>(aligned_query_data + (size_t)i * q_dim, q_dim);
    }
    index.InitVisitedListPool(num_threads);

    std::cout << "k: " << k << std::endl;
    std::vector<uint32_t> res((size_t)q_pts * k, 0), cmps(q_pts, 0), hops(q_pts, 0);
    std::vector<float> res_dists((size_t)q_pts * k, 0.f);
    std::ofstream evaluation_out;
    if (!eval_path.empty()) evaluation_out.open(eval_path, std::ios::out);
    std::cout << "Using thread: " << num_threads << std::endl;
    std::cout << "L_pq" << "\t\tQPS" << "\t\t\tavg_visited" << "\tmean_latency" << "\trecall@" << k << "\tavg_hops"
              << "\tgather_GB/s" << std::endl;
    for (uint32_t L_pq : L_vec) {
      if (k > L_pq) {
        std::cout << "L_pq must greater or equal than k" << std::endl;
        exit(1);
      }
      parameters.Set<uint32_t>("L_pq", L_pq);
      // warm-up like the reference's 100 serial searches (:198-200), as one small batch
      index.SearchRoarGraphBatch(aligned_query_data, std::min<uint32_t>(100, q_pts), k, parameters, res.data(),
                                 res_dists.data(), nullptr, nullptr);
      auto start = std::chrono::high_resolution_clock::now();
      index.SearchRoarGraphBatch(aligned_query_data, q_pts, k, parameters, res.data(), res_dists.data(), cmps.data(),
                                 hops.data());
      auto end = std::chrono::high_resolution_clock::now();
      const double us = (double)std::chrono::duration_cast<std::chrono::microseconds>(end - start).count();
      const float ms = (float)(us / 1000.0);
      const float qps = (float)q_pts / (ms / 1000.0f);
      const float recall = ComputeRecall(q_pts, k, gt_dim, res.data(), gt_ids);
      double sum_cmps = 0, sum_hops = 0;
      for (uint32_t i = 0; i < q_pts; ++i) {
        sum_cmps += cmps[i];
        sum_hops += hops[i];
      }
      uint64_t t_cmps, t_hops, t_adj;
      index.LastSearchStats(t_cmps, t_hops, t_adj);
      const double gbs = ((double)(t_cmps + q_pts) * q_dim * 4.0 + (double)t_adj * 4.0) / (us * 1e-6) / 1e9;
      if (gpus > 1) {  // summed over the shards: every query runs on every shard
        sum_cmps = (double)t_cmps;
        sum_hops = (double)t_hops;
      }
      const float avg_cmps = (float)(sum_cmps / q_pts), avg_hops = (float)(sum_hops / q_pts);
      std::cout << L_pq << "\t\t" << qps << "\t\t" << avg_cmps << "\t\t" << (ms / q_pts) << "\t\t" << recall << "\t\t"
                << avg_hops << "\t\t" << gbs << std::endl;
      if (evaluation_out.is_open())
        evaluation_out << L_pq << "," << qps << "," << avg_cmps << "," << (ms / q_pts) << "," << recall << "," << avg_hops
                       << "," << gbs << std::endl;
    }
    free(aligned_query_data);
    delete[] gt_ids;
    delete[] gt_dists;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  return 0;
}
