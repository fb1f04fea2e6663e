// The C-ABI calls INTEGRATION.md shows (options A / B.1 / B.2 / B.3 and the --gpus form), with the same argument order and
// types, compiled and linked against librg_b200.so by tests/test_capi_symbols.py::test_integration_snippets_compile_and_link.
// Never run: it only has to compile as C++ and resolve every symbol.
#include <cstdint>
#include <vector>

#include "rg_b200.h"

static int hook(void *, uint64_t, uint64_t, void *) { return RG_OK; }

int integration_calls(const float *base, uint64_t nd, uint32_t dim, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                      const uint32_t *knn_ids, uint64_t sq_num, uint32_t K, const float *train, int n_gpus) {
    rg_index *dev = nullptr;
    int rc = rg_index_create(&dev, base, nd, dim, RG_METRIC_INNER_PRODUCT, /*base_on_device=*/0);
    std::vector<uint64_t> row_off(nd + 1, 0);
    std::vector<uint32_t> nbrs;
    rc |= rg_index_set_graph(dev, /*ep=*/0u, row_off.data(), nbrs.data());
    std::vector<uint32_t> res((size_t)nq * k), cmps(nq), hops(nq);
    std::vector<float> dists((size_t)nq * k);
    rc |= rg_search_roargraph(dev, queries, nq, k, L_pq, res.data(), dists.data(), cmps.data(), hops.data());
    if (rc == RG_ENOTENOUGH) (void)rg_last_error();
    // B.2
    std::vector<int32_t> kn((size_t)nq * K);
    std::vector<float> kd((size_t)nq * K);
    rc |= rg_knn(base, nd, queries, nq, dim, K, RG_METRIC_L2, /*id_offset=*/0, kn.data(), kd.data());
    // B.3
    rc |= rg_index_build_roargraph(dev, knn_ids, 0, sq_num, K, /*M_sq=*/100, /*M_pjbp=*/35, /*L_pjpq=*/500, /*n_rounds=*/0);
    uint64_t nnz = 0;
    uint32_t ep = 0, max_deg = 0;
    rc |= rg_index_graph_size(dev, &nnz, &ep, &max_deg);
    rc |= rg_index_get_graph(dev, row_off.data(), nbrs.data());
    rc |= rg_index_destroy(dev);
    // streaming loader
    uint64_t out_nd = 0;
    uint32_t out_dim = 0;
    rc |= rg_index_create_from_fbin(&dev, "base.fbin", RG_METRIC_COSINE, /*hook=*/nullptr, /*user=*/nullptr, &out_nd, &out_dim);
    // --gpus N
    rg_group *g = nullptr;
    rc |= rg_group_create(&g, n_gpus, /*dev_ids=*/nullptr);
    rg_sharded_index *si = nullptr;
    rc |= rg_sharded_index_create(g, &si, base, nd, dim, RG_METRIC_INNER_PRODUCT);
    rc |= rg_sharded_index_build_roargraph(si, train, sq_num, 100, 35, 500);
    rg_index *shard = nullptr;
    uint64_t lo = 0, rows = 0;
    rc |= rg_sharded_index_shard(si, 0, &shard, &lo, &rows);
    rc |= rg_sharded_search_roargraph(si, queries, nq, k, L_pq, res.data(), dists.data());
    rc |= rg_group_knn(g, base, nd, queries, nq, dim, K, RG_METRIC_L2, kn.data(), kd.data());
    rc |= rg_sharded_index_destroy(si);
    rc |= rg_group_destroy(g);
    // sharded base from a torchrun-style host: prepared shard, shared bound, optional re-sync, packed exchange
    rg_knn_base *kb = nullptr;
    rc |= rg_knn_base_create(&kb, base, nd, dim, RG_METRIC_L2, /*base_on_device=*/0);
    float *d_q = nullptr, *d_tau = nullptr, *d_bound = nullptr, *d_d = nullptr;
    int32_t *d_i = nullptr;
    uint64_t *d_packed = nullptr;
    rc |= rg_knn_base_seed_dev(kb, d_q, nq, K, /*rank=*/(K + 7) / 8, d_tau, /*stream=*/nullptr);
    rc |= rg_knn_base_search_seeded_dev(kb, d_q, nq, K, 0, d_tau, /*bound_rows=*/8 * 65536, d_i, d_d, nullptr);
    rc |= rg_knn_base_search_resync_dev(kb, d_q, nq, K, 0, d_tau, 8 * 65536, (K + 7) / 8, /*resync_rows=*/nd / 5, d_bound, hook, nullptr,
                                        d_i, d_d, nullptr);
    rc |= rg_topk_pack_dev(d_d, reinterpret_cast<const uint32_t *>(d_i), (uint64_t)nq * K, /*id_offset=*/0, d_packed, nullptr);
    rc |= rg_topk_merge_packed_dev(d_packed, /*G=*/8, nq, K, K, 0, nq, d_d, reinterpret_cast<uint32_t *>(d_i), nullptr);
    rc |= rg_knn_base_destroy(kb);
    return rc;
}
