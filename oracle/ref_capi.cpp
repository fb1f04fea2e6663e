// TEST INFRASTRUCTURE ONLY — never linked or loaded by the product path.
//
// extern "C" wrapper around the *reference's own* classes, compiled from the sources where they
// lie under /root/reference (src/index_bipartite.cpp, src/index.cpp, include/**) into
// oracle/_ref/libref_roargraph.so by oracle/Makefile. It drives the reference exactly the way its
// CLIs do:
//   build : tests/test_build_roargraph.cpp:105-136  (load_data -> LoadLearnBaseKNN -> BuildRoarGraph
//           -> SaveProjectionGraph)
//   search: tests/test_search_roargraph.cpp:160-209 (IndexBipartite ctor -> LoadSearchNeededData ->
//           LoadProjectionGraph -> InitVisitedListPool -> 100-query warm-up -> omp parallel for
//           schedule(dynamic,1) over SearchRoarGraph)
// so that tests/ and bench.py's cpu_baseline / --impl reference legs can compare against, and time,
// the unmodified reference CPU path.
#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "efanna2e/distance.h"
#include "efanna2e/neighbor.h"
#include "efanna2e/parameters.h"
#include "efanna2e/util.h"
#include "index_bipartite.h"

namespace {
thread_local std::string g_err;
struct Quiet {  // the reference prints progress to std::cout; silence it inside library calls
    std::streambuf *old;
    std::ostringstream sink;
    explicit Quiet(bool on) : old(nullptr) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
    ~Quiet() { if (old) std::cout.rdbuf(old); }
};
efanna2e::Metric to_metric(int m) {
    switch (m) {
        case 0: return efanna2e::L2;
        case 1: return efanna2e::INNER_PRODUCT;
        case 4: return efanna2e::COSINE;
        default: return efanna2e::L2;
    }
}
struct RefIndex {
    efanna2e::IndexBipartite *index;
    unsigned dim;   // padded (x8) dimension the index was constructed with
    unsigned nd;
    int threads_inited;
    bool own_base;
};
// data_bp_ is a protected member of efanna2e::Index (include/efanna2e/index.h:62): a derived class may point the
// unmodified reference at a base array that is already in memory instead of re-reading an fbin file.
struct MemIndex : efanna2e::IndexBipartite {
    using efanna2e::IndexBipartite::IndexBipartite;
    void set_base(const float *p) { data_bp_ = p; }
};
}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

int ref_num_procs() { return omp_get_num_procs(); }

// Reference distance operator on raw pointers (include/efanna2e/distance.h:39-89, :108-225).
float ref_distance(int metric, const float *a, const float *b, unsigned dim) {
    static efanna2e::DistanceL2 l2;
    static efanna2e::DistanceInnerProduct ip;
    return metric == 0 ? l2.compare(a, b, dim) : ip.compare(a, b, dim);
}

// tests/test_build_roargraph.cpp:82-136 without the argument parser.
int ref_build_roargraph(const char *base_fbin, const char *learn_base_nn, const char *out_index, int metric,
                        unsigned M_sq, unsigned M_pjbp, unsigned L_pjpq, unsigned num_threads, int quiet) {
    try {
        Quiet q(quiet != 0);
        uint32_t base_num, base_dim;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        float *data_bp = nullptr;
        efanna2e::Parameters parameters;
        efanna2e::load_data<float>(base_fbin, base_num, base_dim, data_bp);
        // NB the build CLI passes the *unpadded* base_dim to the ctor and load_data pads rows to x8
        // (tests/test_build_roargraph.cpp:105,117); identical when dim % 8 == 0.
        efanna2e::IndexBipartite index(base_dim, base_num, to_metric(metric), nullptr);
        parameters.Set<uint32_t>("M_sq", M_sq);
        parameters.Set<uint32_t>("M_pjbp", M_pjbp);
        parameters.Set<uint32_t>("L_pjpq", L_pjpq);
        parameters.Set<uint32_t>("num_threads", num_threads);
        index.LoadLearnBaseKNN(learn_base_nn);
        omp_set_num_threads(num_threads);
        // sq_num only sizes lock arrays; the CLI passes the sampled-query count. We read it from the kNN file.
        uint32_t sq_num = 0;
        {
            std::ifstream in(learn_base_nn, std::ios::binary);
            in.read((char *)&sq_num, 4);
        }
        index.BuildRoarGraph(sq_num, nullptr, base_num, data_bp, parameters);
        index.SaveProjectionGraph(out_index);
        free(data_bp);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// tests/test_search_roargraph.cpp:160-173.
void *ref_index_load(const char *base_fbin, const char *index_file, int metric, unsigned padded_dim,
                     unsigned num_threads, int quiet) {
    try {
        Quiet q(quiet != 0);
        uint32_t base_num, base_dim;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        auto *r = new RefIndex();
        r->index = new efanna2e::IndexBipartite(padded_dim, base_num, to_metric(metric), nullptr);
        r->index->LoadSearchNeededData(base_fbin, "");
        r->index->LoadProjectionGraph(index_file);
        r->index->InitVisitedListPool(num_threads);
        r->dim = padded_dim;
        r->nd = base_num;
        r->threads_inited = (int)num_threads;
        r->own_base = true;
        return r;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

// Same as ref_index_load, but the base rows (padded to x8, already normalised for COSINE — what LoadVectorData
// would have produced, src/index_bipartite.cpp:2679-2686) are handed over in memory; the caller keeps ownership.
void *ref_index_from_memory(const float *base, uint64_t nd, unsigned padded_dim, int metric, const char *index_file,
                            unsigned num_threads, int quiet) {
    try {
        Quiet q(quiet != 0);
        auto *r = new RefIndex();
        auto *mi = new MemIndex(padded_dim, nd, to_metric(metric), nullptr);
        mi->set_base(base);
        r->index = mi;
        r->index->LoadProjectionGraph(index_file);
        r->index->InitVisitedListPool(num_threads);
        r->dim = padded_dim;
        r->nd = (unsigned)nd;
        r->threads_inited = (int)num_threads;
        r->own_base = false;
        return r;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

int ref_index_need_normalize(void *h) { return ((RefIndex *)h)->index->need_normalize ? 1 : 0; }
const float *ref_index_base(void *h) { return ((RefIndex *)h)->index->GetBasePointSet(); }

// In-place row normalisation with the reference's own routine (include/efanna2e/util.h:214-225), as the
// search CLI applies to COSINE queries (tests/test_search_roargraph.cpp:167-172).
void ref_normalize_rows(float *rows, uint64_t n, unsigned dim) {
    for (uint64_t i = 0; i < n; ++i) efanna2e::normalize<float>(rows + i * dim, dim);
}

// The timed loop of tests/test_search_roargraph.cpp:196-209 for one L_pq. `queries` must already be padded to
// the index dimension (and normalised for COSINE). Returns wall seconds of the parallel region in *wall_s.
int ref_search_roargraph(void *h, const float *queries, uint64_t nq, unsigned k, unsigned L_pq, unsigned num_threads,
                         int warmup100, unsigned *ids, float *dists, unsigned *cmps, unsigned *hops,
                         double *wall_s) {
    RefIndex *r = (RefIndex *)h;
    try {
        efanna2e::Parameters parameters;
        parameters.Set<uint32_t>("num_threads", num_threads);
        parameters.Set<uint32_t>("L_pq", L_pq);
        omp_set_num_threads(num_threads);
        const size_t dim = r->dim;
        if (warmup100) {
            std::vector<unsigned> tmp(k);
            std::vector<float> tmpd(k);
            for (size_t i = 0; i < 100 && i < nq; ++i)
                r->index->SearchRoarGraph(queries + i * dim, k, i, parameters, tmp.data(), tmpd);
        }
        std::string err;
        auto start = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t i = 0; i < nq; ++i) {
            std::vector<float> d(k, 0.0f);
            try {
                auto ret = r->index->SearchRoarGraph(queries + i * dim, k, i, parameters, ids + i * k, d);
                if (cmps) cmps[i] = ret.first;
                if (hops) hops[i] = ret.second;
                if (dists) std::memcpy(dists + i * k, d.data(), k * sizeof(float));
            } catch (const std::exception &e) {
#pragma omp critical
                err = e.what();
            }
        }
        auto end = std::chrono::high_resolution_clock::now();
        if (wall_s) *wall_s = std::chrono::duration<double>(end - start).count();
        if (!err.empty()) {
            g_err = err;
            return -2;
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// ---- SearchBipartiteGraph (src/index_bipartite.cpp:282-356) ------------------------------------------------------------
// tests/test_search_bipartite.cpp:158-168: ctor(padded dim, base_num + sq_num) -> LoadSearchNeededData -> Load ->
// InitVisitedListPool.  `n_sq` only sizes the ctor argument (the search reads nd_ from the base file).
void *ref_bipartite_load(const char *base_fbin, const char *bipartite_index_file, int metric, unsigned padded_dim,
                         uint64_t n_sq, unsigned num_threads, int quiet) {
    try {
        Quiet q(quiet != 0);
        uint32_t base_num, base_dim;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        auto *r = new RefIndex();
        r->index = new efanna2e::IndexBipartite(padded_dim, base_num + n_sq, to_metric(metric), nullptr);
        r->index->LoadSearchNeededData(base_fbin, "");
        r->index->Load(bipartite_index_file);
        r->index->InitVisitedListPool(num_threads);
        r->dim = padded_dim;
        r->nd = base_num;
        r->threads_inited = (int)num_threads;
        r->own_base = true;
        return r;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

// 1 when this library was built with oracle/shims/fixed_random_device.h pre-included (libref_roargraph_detrand.so):
// the word std::random_device returns inside SearchBipartiteGraph is then the one set below, per query.
int ref_has_fixed_random_device() {
#ifdef RG_DETRAND
    return 1;
#else
    return 0;
#endif
}

// The 10 entry points the reference derives from one random_device word (:287-294): libstdc++'s own mt19937 and
// uniform_int_distribution, driven exactly as the reference drives them.
void ref_bipartite_entry_points(uint32_t seed_word, uint32_t nd, uint32_t *out10) {
    std::mt19937 gen(seed_word);
    std::uniform_int_distribution<uint32_t> dis(0, nd - 1);
    for (uint32_t i = 0; i < 10; ++i) out10[i] = dis(gen);
}

// The timed loop of tests/test_search_bipartite.cpp:196-209 for one L_pq.  With the fixed random_device, query i sees the
// seed word seed_base + i (whatever thread runs it) and init_ids_out[i*10..] receives the entry points that word yields;
// without it the entry points are the reference's own unseeded draw and init_ids_out is left untouched.
int ref_search_bipartite(void *h, const float *queries, uint64_t nq, unsigned k, unsigned L_pq, unsigned num_threads,
                         uint32_t seed_base, unsigned *ids, float *dists, unsigned *cmps, unsigned *hops,
                         uint32_t *init_ids_out, double *wall_s) {
    RefIndex *r = (RefIndex *)h;
    try {
        efanna2e::Parameters parameters;
        parameters.Set<uint32_t>("num_threads", num_threads);
        parameters.Set<uint32_t>("L_pq", L_pq);
        omp_set_num_threads(num_threads);
        const size_t dim = r->dim;
        std::string err;
        int bad_draws = 0;
        auto start = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t i = 0; i < nq; ++i) {
            std::vector<float> d(k, 0.0f);
            try {
#ifdef RG_DETRAND
                rg_detrand::next_seed = seed_base + (uint32_t)i;
                const uint64_t before = rg_detrand::calls;
                if (init_ids_out) ref_bipartite_entry_points(seed_base + (uint32_t)i, r->nd, init_ids_out + i * 10);
#endif
                auto ret = r->index->SearchBipartiteGraph(queries + i * dim, k, i, parameters, ids + i * k, d);
#ifdef RG_DETRAND
                if (rg_detrand::calls != before + 1) {
#pragma omp atomic
                    bad_draws++;
                }
#endif
                if (cmps) cmps[i] = ret.first;
                if (hops) hops[i] = ret.second;
                if (dists) std::memcpy(dists + i * k, d.data(), k * sizeof(float));
            } catch (const std::exception &e) {
#pragma omp critical
                err = e.what();
            }
        }
        auto end = std::chrono::high_resolution_clock::now();
        if (wall_s) *wall_s = std::chrono::duration<double>(end - start).count();
        if (!err.empty()) {
            g_err = err;
            return -2;
        }
        if (bad_draws) {
            g_err = "SearchBipartiteGraph did not draw exactly one random_device word per query";
            return -3;
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

void ref_index_free(void *h) {
    RefIndex *r = (RefIndex *)h;
    if (!r) return;
    // The reference never frees its base array or graph (IndexBipartite::~IndexBipartite is empty); release the
    // big base block ourselves so test processes do not accumulate it.
    if (r->own_base) free((void *)r->index->GetBasePointSet());
    delete r->index;
    delete r;
}

}  // extern "C"
