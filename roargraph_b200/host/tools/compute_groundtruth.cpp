// compute_groundtruth — the brute-force kNN tool the reference vendors from DiskANN
// (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp there) on the B200 library: same flags
// (--data_type --dist_fn{l2,mips,cosine} --base_file --query_file --gt_file --K [--tags_file]) and the same output file
// (int32 n, int32 K, n*K uint32 ids, n*K float distances; +IP for mips, :438-439; squared L2 otherwise).
// The base is processed in parts of at most PARTSIZE rows like the reference (:32, :396-448); each part is one
// rg_knn call (ids offset by the part's first row, :413) and parts are merged per query by (distance, id).
// --gpus N (the one added flag, default 1): every part is sharded over N GPUs of the box by rg_group_knn — the same part
// loop one level down: every GPU keeps its rows prepared, the query batches stream past, ONE packed all-gather + merge per
// batch (BASELINE.json configs[3]).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cli_args.h"
#include "rg_b200.h"

// rows per base part (:32 of the reference); RG_GT_PARTSIZE overrides it (a base that does not fit one GPU beside the
// query batch, tests)
static uint64_t part_size() {
  static const uint64_t v = [] {
    const char *e = getenv("RG_GT_PARTSIZE");
    const uint64_t x = e ? strtoull(e, nullptr, 10) : 0;
    return x ? x : (uint64_t)20000000;
  }();
  return v;
}
#define PARTSIZE (part_size())

template <typename T>
static void load_bin_as_float(const std::string &path, std::vector<float> &out, uint64_t &npts, uint64_t &ndims,
                              uint64_t part, uint64_t &start_id) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error("Could not open file " + path);
  int32_t n32 = 0, d32 = 0;
  in.read((char *)&n32, 4);
  in.read((char *)&d32, 4);
  const uint64_t total = (uint64_t)n32;
  ndims = (uint64_t)d32;
  start_id = part * PARTSIZE;
  const uint64_t end_id = std::min(start_id + PARTSIZE, total);
  npts = end_id - start_id;
  std::cout << "Reading bin file " << path << " part " << part << ": " << npts << " x " << ndims << std::endl;
  out.resize(npts * ndims);
  in.seekg((std::streamoff)(8 + start_id * ndims * sizeof(T)), std::ios::beg);
  if (sizeof(T) == sizeof(float)) {
    in.read((char *)out.data(), (std::streamsize)(npts * ndims * sizeof(float)));
  } else {
    std::vector<T> tmp(1 << 20);
    uint64_t done = 0;
    while (done < npts * ndims) {
      const uint64_t n = std::min<uint64_t>(tmp.size(), npts * ndims - done);
      in.read((char *)tmp.data(), (std::streamsize)(n * sizeof(T)));
      for (uint64_t i = 0; i < n; ++i) out[done + i] = (float)tmp[i];
      done += n;
    }
  }
  if (!in) throw std::runtime_error("file " + path + " is truncated");
}

static uint64_t num_points(const std::string &path) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error("Could not open file " + path);
  int32_t n32 = 0;
  in.read((char *)&n32, 4);
  return (uint64_t)n32;
}

template <typename T>
static int aux_main(const std::string &base_file, const std::string &query_file, const std::string &gt_file, uint64_t K,
                    int metric, int gpus) {
  rg_group *group = nullptr;
  if (gpus > 1) {
    if (rg_group_create(&group, gpus, nullptr) != RG_OK) throw std::runtime_error(rg_last_error());
    std::cout << "base parts sharded over " << gpus << " GPUs" << std::endl;
  }
  std::vector<float> queries, base;
  uint64_t nq = 0, dim = 0, qd = 0, dummy = 0;
  load_bin_as_float<T>(query_file, queries, nq, qd, 0, dummy);
  const uint64_t total = num_points(base_file);
  const uint64_t parts = (total + PARTSIZE - 1) / PARTSIZE;
  std::vector<int32_t> ids(nq * K), part_ids;
  std::vector<float> dists(nq * K), part_d;
  auto t0 = std::chrono::high_resolution_clock::now();
  for (uint64_t p = 0; p < parts; ++p) {
    uint64_t np = 0, start = 0;
    load_bin_as_float<T>(base_file, base, np, dim, p, start);
    if (dim != qd) throw std::runtime_error("base and query dimensions differ");
    const uint64_t Kp = std::min<uint64_t>(K, np);
    std::vector<int32_t> &oi = p == 0 && parts == 1 ? ids : part_ids;
    std::vector<float> &od = p == 0 && parts == 1 ? dists : part_d;
    oi.assign(nq * Kp, -1);
    od.assign(nq * Kp, 0.f);
    int rc;
    if (group) {
      rc = rg_group_knn(group, base.data(), np, queries.data(), nq, (uint32_t)dim, (uint32_t)Kp, metric, oi.data(), od.data());
      if (rc == RG_OK && start)
        for (auto &v : oi) v += (int32_t)start;  // start_id of this part (:413); inside the part the shards are already offset
    } else {
      rc = rg_knn(base.data(), np, queries.data(), nq, (uint32_t)dim, (uint32_t)Kp, metric, (int32_t)start, oi.data(), od.data());
    }
    if (rc != RG_OK) throw std::runtime_error(rg_last_error());
    if (parts == 1) {
      if (Kp != K) throw std::runtime_error("K exceeds the number of base points");
      break;
    }
    // merge this part into the running top-K by (distance, id)
    std::vector<std::pair<float, int32_t>> tmp;
    for (uint64_t q = 0; q < nq; ++q) {
      tmp.clear();
      if (p > 0)
        for (uint64_t i = 0; i < K; ++i)
          if (ids[q * K + i] >= 0) tmp.emplace_back(dists[q * K + i], ids[q * K + i]);
      for (uint64_t i = 0; i < Kp; ++i) tmp.emplace_back(od[q * Kp + i], oi[q * Kp + i]);
      std::sort(tmp.begin(), tmp.end());
      for (uint64_t i = 0; i < K; ++i) {
        ids[q * K + i] = i < tmp.size() ? tmp[i].second : -1;
        dists[q * K + i] = i < tmp.size() ? tmp[i].first : 0.f;
      }
    }
  }
  if (group) rg_group_destroy(group);
  auto t1 = std::chrono::high_resolution_clock::now();
  const double sec = std::chrono::duration<double>(t1 - t0).count();
  std::cout << "kNN of " << nq << " queries x " << total << " points, K=" << K << ": " << sec << " s, "
            << (double)nq * (double)total / sec << " pairs/s (file IO of the base parts included)" << std::endl;
  if (metric == RG_METRIC_INNER_PRODUCT)
    for (auto &d : dists) d = -d;  // stored as +IP (:438-439)
  std::ofstream out(gt_file, std::ios::binary);
  if (!out.is_open()) throw std::runtime_error("Could not open file " + gt_file);
  const int32_t n32 = (int32_t)nq, k32 = (int32_t)K;
  out.write((const char *)&n32, 4);
  out.write((const char *)&k32, 4);
  out.write((const char *)ids.data(), (std::streamsize)(ids.size() * 4));
  out.write((const char *)dists.data(), (std::streamsize)(dists.size() * 4));
  std::cout << "Finished writing truthset" << std::endl;
  return 0;
}

int main(int argc, char **argv) {
  cli::Parser args("Arguments");
  args.add("data_type", "data type <int8/uint8/float>", true)
      .add("dist_fn", "distance function <l2/mips/cosine>", true)
      .add("base_file", "File containing the base vectors in binary format", true)
      .add("query_file", "File containing the query vectors in binary format", true)
      .add("gt_file", "File name for the writing ground truth in binary format", true)
      .add("K", "Number of ground truth nearest neighbors to compute", true).numeric()
      .add("tags_file", "File containing the tags in binary format", false)
      .add("gpus", "GPUs of this box the base parts are sharded over (one NCCL rank per GPU inside this process)", false, false, "1").numeric();
  try {
    if (args.parse(argc, argv)) return 0;
  } catch (const std::exception &ex) {
    std::cerr << ex.what() << '\n';
    return -1;
  }
  const std::string data_type = args.str("data_type"), dist_fn = args.str("dist_fn");
  if (data_type != "float" && data_type != "int8" && data_type != "uint8") {
    std::cout << "Unsupported type. float, int8 and uint8 types are supported." << std::endl;
    return -1;
  }
  int metric;
  if (dist_fn == "l2") metric = RG_METRIC_L2;
  else if (dist_fn == "mips") metric = RG_METRIC_INNER_PRODUCT;
  else if (dist_fn == "cosine") metric = RG_METRIC_COSINE;
  else {
    std::cerr << "Unsupported distance function. Use l2/mips/cosine." << std::endl;
    return -1;
  }
  if (!args.str("tags_file").empty()) {
    std::cerr << "--tags_file (DiskANN tag remapping) is not used by RoarGraph and is not supported." << std::endl;
    return -1;
  }
  try {
    const uint64_t K = args.u64("K");
    const int gpus = (int)args.u64("gpus");
    if (gpus < 1) throw std::runtime_error("--gpus must be at least 1");
    if (data_type == "float") return aux_main<float>(args.str("base_file"), args.str("query_file"), args.str("gt_file"), K, metric, gpus);
    if (data_type == "int8") return aux_main<int8_t>(args.str("base_file"), args.str("query_file"), args.str("gt_file"), K, metric, gpus);
    return aux_main<uint8_t>(args.str("base_file"), args.str("query_file"), args.str("gt_file"), K, metric, gpus);
  } catch (const std::exception &e) {
    std::cout << std::string(e.what()) << std::endl;
    std::cerr << "Compute GT failed." << std::endl;
    return -1;
  }
}
