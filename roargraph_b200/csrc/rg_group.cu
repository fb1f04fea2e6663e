// rg_group.cu — the GPUs of one box driven from ONE host process: rg_group_* / rg_sharded_* of include/rg_b200.h.
//
// This is what "host code stays C++" means for the multi-GPU configurations of BASELINE.json (configs[3], configs[4]): the
// reference's CLIs are single processes, so the drop-in keeps one process, gives every GPU one worker thread and one NCCL
// rank (ncclCommInitAll), shards the base rows by equal counts exactly like compute_groundtruth's base parts
// (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:396-416, ids offset by the part's start :413) and merges the
// per-shard top-k lists after ONE all-gather of packed (distance, id) words per exchange step (the std::sort merge of
// :424-444).  Python / torchrun jobs (bench.py) use one process per GPU and torch.distributed for the same exchange.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the single-GPU entry points of librg_b200.so keep working on a
// machine without it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "rg_internal.h"
#include "rg_rendezvous.h"

namespace rg {
namespace {

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            err = std::string("libnccl.so.2 not found: ") + dlerror();
            return false;
        }
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !AllGather || !AllReduce || !GetErrorString) {
            err = "libnccl.so.2 lacks a required symbol";
            return false;
        }
        return true;
    }
};
Nccl g_nccl;

}  // namespace
}  // namespace rg

struct rg_group {
    int n = 0;
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> stream;
};

struct rg_sharded_index {
    rg_group *group = nullptr;
    uint64_t nd = 0;
    uint32_t dim = 0;
    int metric = 0;
    std::vector<rg_index *> shard;
    std::vector<uint64_t> row_lo;  // G + 1 entries
    // per-shard device scratch of the sharded search, grown on demand
    struct Scratch {
        rg::DevBuf<float> q, dists, out_d;
        rg::DevBuf<uint32_t> ids, out_i;
        rg::DevBuf<unsigned long long> packed, gathered;
    };
    std::vector<Scratch> scratch;
};

using namespace rg;

namespace rg {
namespace {

// fn(shard) on one host thread per GPU of the group, each with its device current; the first failure's code and message
// are handed back to the calling thread (rg_last_error is thread-local)
// kPeerFailed: this worker is fine but left because another one failed before a collective (Rendezvous); the caller gets the
// failing worker's code and message.  A worker that fails in the prologue below reports to `rv` (when the call uses one) so
// that its peers do not wait for it.
constexpr int kPeerFailed = -100;
int peer_failed() {
    set_error("another GPU of the group failed before the exchange");
    return kPeerFailed;
}

int run_on_all(rg_group *g, const std::function<int(int)> &fn, Rendezvous *rv = nullptr) {
    std::vector<int> rc(g->n, RG_OK);
    std::vector<std::string> msg(g->n);
    std::vector<std::thread> th;
    th.reserve(g->n);
    for (int s = 0; s < g->n; ++s)
        th.emplace_back([&, s]() {
            cudaError_t e = cudaSetDevice(g->dev[s]);
            if (e != cudaSuccess) {
                rc[s] = cuda_fail(e, "cudaSetDevice");
            } else {
                rc[s] = ensure_device_ready();
                if (rc[s] == RG_OK) rc[s] = fn(s);
                else if (rv) rv->agree(false);
            }
            if (e != cudaSuccess && rv) rv->agree(false);
            if (rc[s] != RG_OK) msg[s] = rg_last_error();
        });
    for (auto &t : th) t.join();
    for (int pass = 0; pass < 2; ++pass)  // a worker's own failure first, "a peer failed" only if nothing better is known
        for (int s = 0; s < g->n; ++s)
            if (rc[s] != RG_OK && (pass == 1 || rc[s] != kPeerFailed)) {
                set_error("GPU " + std::to_string(g->dev[s]) + ": " + msg[s]);
                return rc[s] == kPeerFailed ? RG_ECUDA : rc[s];
            }
    return RG_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
    set_error(std::string("NCCL error: ") + g_nccl.GetErrorString(r) + " in " + what);
    return RG_ECUDA;
}
#define RG_NCCL(call)                                              \
    do {                                                           \
        ncclResult_t r_ = (call);                                  \
        if (r_ != ncclSuccess) return nccl_fail(r_, #call);        \
    } while (0)

uint64_t shard_lo(uint64_t n, int s, int G) { return (uint64_t)s * n / (uint64_t)G; }

}  // namespace
}  // namespace rg

extern "C" {

int rg_group_create(rg_group **out, int n_gpus, const int *dev_ids) {
    if (!out || n_gpus < 1 || n_gpus > 32) { set_error("rg_group_create: n_gpus must be in [1, 32]"); return RG_EINVAL; }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (librg_b200 has no CPU fallback)");
        return RG_ECUDA;
    }
    rg_group *g = new rg_group();
    g->n = n_gpus;
    for (int s = 0; s < n_gpus; ++s) {
        const int d = dev_ids ? dev_ids[s] : s;
        if (d < 0 || d >= have) {
            delete g;
            set_error("rg_group_create: no such CUDA device");
            return RG_EINVAL;
        }
        g->dev.push_back(d);
    }
    g->comm.assign(n_gpus, nullptr);
    g->stream.assign(n_gpus, nullptr);
    if (n_gpus > 1) {
        std::string err;
        if (!g_nccl.load(err)) {
            delete g;
            set_error(err);
            return RG_ECUDA;
        }
        ncclResult_t r = g_nccl.CommInitAll(g->comm.data(), n_gpus, g->dev.data());
        if (r != ncclSuccess) {
            delete g;
            return nccl_fail(r, "ncclCommInitAll");
        }
    }
    int rc = run_on_all(g, [&](int s) {
        RG_CUDA(cudaStreamCreateWithFlags(&g->stream[s], cudaStreamNonBlocking));
        return RG_OK;
    });
    if (rc != RG_OK) {
        rg_group_destroy(g);
        return rc;
    }
    *out = g;
    return RG_OK;
}

int rg_group_destroy(rg_group *g) {
    if (!g) return RG_OK;
    for (int s = 0; s < g->n; ++s) {
        cudaSetDevice(g->dev[s]);
        if (g->stream[s]) cudaStreamDestroy(g->stream[s]);
        if (g->comm[s]) g_nccl.CommDestroy(g->comm[s]);
    }
    delete g;
    return RG_OK;
}

int rg_group_size(rg_group *g) { return g ? g->n : 0; }

int rg_sharded_index_create(rg_group *g, rg_sharded_index **out, const float *base, uint64_t nd, uint32_t dim, int metric) {
    if (!g || !out || !base || nd < (uint64_t)g->n) { set_error("rg_sharded_index_create: bad argument"); return RG_EINVAL; }
    rg_sharded_index *si = new rg_sharded_index();
    si->group = g;
    si->nd = nd;
    si->dim = dim;
    si->metric = metric;
    si->shard.assign(g->n, nullptr);
    si->scratch.resize(g->n);
    for (int s = 0; s <= g->n; ++s) si->row_lo.push_back(shard_lo(nd, s, g->n));
    int rc = run_on_all(g, [&](int s) {
        const uint64_t lo = si->row_lo[s], hi = si->row_lo[s + 1];
        return rg_index_create(&si->shard[s], base + lo * dim, hi - lo, dim, metric, 0);
    });
    if (rc != RG_OK) {
        rg_sharded_index_destroy(si);
        return rc;
    }
    *out = si;
    return RG_OK;
}

int rg_sharded_index_destroy(rg_sharded_index *si) {
    if (!si) return RG_OK;
    for (int s = 0; s < si->group->n; ++s) {
        cudaSetDevice(si->group->dev[s]);
        if (si->shard[s]) rg_index_destroy(si->shard[s]);
        auto &sc = si->scratch[s];
        sc.q.release();
        sc.dists.release();
        sc.out_d.release();
        sc.ids.release();
        sc.out_i.release();
        sc.packed.release();
        sc.gathered.release();
    }
    delete si;
    return RG_OK;
}

int rg_sharded_index_shards(rg_sharded_index *si) { return si ? si->group->n : 0; }

int rg_sharded_index_shard(rg_sharded_index *si, int shard, rg_index **index, uint64_t *row_lo, uint64_t *rows) {
    if (!si || shard < 0 || shard >= si->group->n) { set_error("rg_sharded_index_shard: no such shard"); return RG_EINVAL; }
    if (index) *index = si->shard[shard];
    if (row_lo) *row_lo = si->row_lo[shard];
    if (rows) *rows = si->row_lo[shard + 1] - si->row_lo[shard];
    return RG_OK;
}

int rg_sharded_index_build_roargraph(rg_sharded_index *si, const float *train, uint64_t nsq, uint32_t M_sq, uint32_t M_pjbp,
                                     uint32_t L_pjpq) {
    if (!si || !train || nsq == 0) { set_error("rg_sharded_index_build_roargraph: null/empty argument"); return RG_EINVAL; }
    rg_group *g = si->group;
    const int knn_metric = si->metric == RG_METRIC_L2 ? RG_METRIC_L2 : RG_METRIC_INNER_PRODUCT;  // COSINE rows are normalised
    return run_on_all(g, [&](int s) {
        // the shard's own learn-base kNN lists (what compute_groundtruth would write for this base part), then its own graph
        rg_index *idx = si->shard[s];
        const uint64_t rows = si->row_lo[s + 1] - si->row_lo[s];
        if (M_sq > rows) { set_error("rg_sharded_index_build_roargraph: M_sq exceeds the shard's rows"); return RG_EINVAL; }
        float *d_train = nullptr, *d_d = nullptr;
        int32_t *d_i = nullptr;
        cudaError_t e = cudaMalloc(&d_train, nsq * si->dim * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&d_d, nsq * M_sq * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&d_i, nsq * M_sq * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(d_train, train, nsq * si->dim * sizeof(float), cudaMemcpyHostToDevice);
        int rc = e == cudaSuccess ? RG_OK : cuda_fail(e, "training-query upload");
        if (rc == RG_OK) rc = knn_device(idx->d_base, rows, d_train, nsq, si->dim, M_sq, knn_metric, 0, d_i, d_d, g->stream[s]);
        if (d_train) cudaFree(d_train);
        if (d_d) cudaFree(d_d);
        knn_release_workspace();
        if (rc == RG_OK)
            rc = rg_index_build_roargraph(idx, reinterpret_cast<const uint32_t *>(d_i), 1, nsq, M_sq, M_sq, M_pjbp, L_pjpq, 0);
        if (d_i) cudaFree(d_i);
        return rc;
    });
}

int rg_sharded_search_roargraph(rg_sharded_index *si, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq, uint32_t *ids,
                                float *dists) {
    if (!si || !queries || !ids || !dists) { set_error("rg_sharded_search_roargraph: null buffer"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    rg_group *g = si->group;
    const int G = g->n;
    const uint32_t dim = si->dim;
    const uint64_t per = ((uint64_t)nq + G - 1) / G;  // query rows uploaded per GPU; the rest arrives over NVLink
    Rendezvous rv(G);
    return run_on_all(g, [&](int s) {
        auto &sc = si->scratch[s];
        cudaStream_t st = g->stream[s];
        int rc;
        if ((rc = sc.q.reserve((size_t)G * per * dim)) == RG_OK && (rc = sc.ids.reserve((size_t)nq * k)) == RG_OK &&
            (rc = sc.dists.reserve((size_t)nq * k)) == RG_OK && (rc = sc.packed.reserve((size_t)nq * k)) == RG_OK &&
            (rc = sc.gathered.reserve((size_t)G * nq * k)) == RG_OK && s == 0 &&
            (rc = sc.out_i.reserve((size_t)nq * k)) == RG_OK)
            rc = sc.out_d.reserve((size_t)nq * k);
        // nobody enters the collectives unless every GPU has its buffers (a worker that ran out of memory alone would leave
        // its peers blocked in the all-gather)
        if (!rv.agree(rc == RG_OK)) return rc != RG_OK ? rc : peer_failed();
        const uint64_t lo = std::min<uint64_t>((uint64_t)s * per, nq), hi = std::min<uint64_t>(lo + per, nq);
        float *mine = sc.q.p + (size_t)s * per * dim;
        if (hi > lo) RG_CUDA(cudaMemcpyAsync(mine, queries + lo * dim, (hi - lo) * dim * sizeof(float), cudaMemcpyHostToDevice, st));
        if (G > 1) RG_NCCL(g_nccl.AllGather(mine, sc.q.p, per * dim, ncclFloat, g->comm[s], st));
        rc = rg_search_roargraph_dev(si->shard[s], sc.q.p, nq, k, L_pq, sc.ids.p, sc.dists.p, nullptr, nullptr, st);
        if (rc != RG_OK) return rc;
        rc = topk_pack_device(sc.dists.p, sc.ids.p, (uint64_t)nq * k, (uint32_t)si->row_lo[s], sc.packed.p, st);
        if (rc != RG_OK) return rc;
        if (G > 1) RG_NCCL(g_nccl.AllGather(sc.packed.p, sc.gathered.p, (size_t)nq * k, ncclUint64, g->comm[s], st));  // the ONE exchange
        else RG_CUDA(cudaMemcpyAsync(sc.gathered.p, sc.packed.p, (size_t)nq * k * 8, cudaMemcpyDeviceToDevice, st));
        if (s == 0) {  // every rank holds every list; the first one merges and answers
            rc = topk_merge_packed_device(sc.gathered.p, G, nq, k, k, 0, nq, sc.out_d.p, sc.out_i.p, st);
            if (rc != RG_OK) return rc;
            RG_CUDA(cudaMemcpyAsync(ids, sc.out_i.p, (size_t)nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            RG_CUDA(cudaMemcpyAsync(dists, sc.out_d.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        return rg_index_check(si->shard[s], st);  // synchronises; "not enough results" on a shard surfaces here
    }, &rv);
}

int rg_group_knn(rg_group *g, const float *base, uint64_t nb, const float *queries, uint64_t nq, uint32_t dim, uint32_t K,
                 int metric, int32_t *ids, float *dists) {
    if (!g || !base || !queries || !ids || !dists || nb < (uint64_t)g->n || dim == 0) { set_error("rg_group_knn: bad argument"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    const int G = g->n;
    uint64_t batch = 151552;  // whole waves of the tensor-core filter (rg_knn.cu)
    if (const char *hb = getenv("RG_KNN_HOST_BATCH")) batch = std::max<uint64_t>(1, strtoull(hb, nullptr, 10));
    batch = std::min(batch, nq);
    const uint64_t per = (batch + G - 1) / G;
    // the same re-sync point on every shard: a fifth of the smallest shard, on the chunk grid (rg_knn.cu)
    // (RG_KNN_RESYNC=1 enables it: worth ~2 % on one rank's share of the 8-way 10 M x 10 M job, less than the all-reduce it
    // adds costs when the ranks are out of step, profiles/r2_knn_shard_schedules.txt)
    uint64_t resync_rows = (nb / G) / 5;
    if (const char *rr = getenv("RG_KNN_RESYNC_ROWS")) resync_rows = strtoull(rr, nullptr, 10);
    const char *re = getenv("RG_KNN_RESYNC");
    const bool resync_on = re && re[0] == '1';
    if (K > nb / G) { set_error("rg_group_knn: K exceeds a shard's rows"); return RG_EINVAL; }  // the smallest shard
    Rendezvous rv(G);
    return run_on_all(g, [&](int s) {
        cudaStream_t st = g->stream[s];
        const uint64_t lo = shard_lo(nb, s, G), hi = shard_lo(nb, s + 1, G);
        float *d_base = nullptr;
        cudaError_t e = cudaMalloc(&d_base, (hi - lo) * dim * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(d_base, base + lo * dim, (hi - lo) * dim * sizeof(float), cudaMemcpyHostToDevice);
        int rc = RG_OK;
        rg_knn_base *kb = nullptr;
        if (e != cudaSuccess) {
            if (d_base) cudaFree(d_base);
            rc = cuda_fail(e, "rg_group_knn: base shard upload");
        } else {
            rc = knn_base_new(&kb, d_base, d_base, hi - lo, dim, metric);  // the object owns d_base from here
        }
        DevBuf<float> q, dd, out_d, tau, bound;
        DevBuf<int32_t> di;
        DevBuf<uint32_t> out_i;
        DevBuf<unsigned long long> packed, gathered;
        if (rc == RG_OK && (rc = q.reserve((size_t)G * per * dim)) == RG_OK && (rc = dd.reserve(batch * K)) == RG_OK &&
            (rc = di.reserve(batch * K)) == RG_OK && (rc = packed.reserve(batch * K)) == RG_OK &&
            (rc = gathered.reserve((size_t)G * batch * K)) == RG_OK && (rc = out_d.reserve(batch * K)) == RG_OK &&
            (rc = out_i.reserve(batch * K)) == RG_OK && (rc = tau.reserve(batch)) == RG_OK)
            rc = bound.reserve(batch);
        // nobody enters the collectives unless every GPU holds its shard and its buffers
        if (!rv.agree(rc == RG_OK) && rc == RG_OK) rc = peer_failed();
        if (rc == RG_OK) {
            for (uint64_t q0 = 0; q0 < nq && rc == RG_OK; q0 += batch) {
                const uint64_t n = std::min(batch, nq - q0);
                const uint64_t pr = (n + G - 1) / G;  // rows of this batch uploaded by (and merged on) each GPU
                const uint64_t a = std::min<uint64_t>((uint64_t)s * pr, n), b = std::min<uint64_t>(a + pr, n);
                float *mine = q.p + (size_t)s * pr * dim;
                if (b > a) {
                    e = cudaMemcpyAsync(mine, queries + (q0 + a) * dim, (b - a) * dim * sizeof(float), cudaMemcpyHostToDevice, st);
                    if (e != cudaSuccess) { rc = cuda_fail(e, "rg_group_knn: query upload"); break; }
                }
                if (G > 1) {
                    ncclResult_t r = g_nccl.AllGather(mine, q.p, pr * dim, ncclFloat, g->comm[s], st);
                    if (r != ncclSuccess) { rc = nccl_fail(r, "ncclAllGather(queries)"); break; }
                }
                if (G > 1) {
                    // the shards agree on one float per query before they filter (include/rg_b200.h, rg_knn_base_seed_dev)
                    rc = knn_base_seed(kb, q.p, n, K, (K + G - 1) / G, tau.p, st);
                    if (rc != RG_OK) break;
                    ncclResult_t r = g_nccl.AllReduce(tau.p, tau.p, n, ncclFloat, ncclMax, g->comm[s], st);
                    if (r != ncclSuccess) { rc = nccl_fail(r, "ncclAllReduce(tau)"); break; }
                    // ... and once more inside the shard, from the rows re-ranked so far (rg_knn_base_search_resync_dev)
                    struct HookCtx { rg_group *g; int s; float *bound; ncclResult_t r; } hc{g, s, bound.p, ncclSuccess};
                    KnnResync rs{(K + G - 1) / G, resync_rows, bound.p,
                                 [](void *u, uint64_t q_off, uint64_t cnt, void *stream) -> int {
                                     HookCtx *c = static_cast<HookCtx *>(u);
                                     c->r = g_nccl.AllReduce(c->bound + q_off, c->bound + q_off, cnt, ncclFloat, ncclMax, c->g->comm[c->s],
                                                             (cudaStream_t)stream);
                                     return c->r == ncclSuccess ? RG_OK : RG_ECUDA;
                                 },
                                 &hc};
                    rc = knn_base_search(kb, q.p, n, K, 0, di.p, dd.p, st, tau.p, (uint64_t)G * 65536u, resync_on ? &rs : nullptr);
                    if (hc.r != ncclSuccess) { rc = nccl_fail(hc.r, "ncclAllReduce(bound)"); break; }
                } else {
                    rc = knn_base_search(kb, q.p, n, K, 0, di.p, dd.p, st);
                }
                if (rc != RG_OK) break;
                rc = topk_pack_device(dd.p, reinterpret_cast<const uint32_t *>(di.p), n * K, (uint32_t)lo, packed.p, st);
                if (rc != RG_OK) break;
                if (G > 1) {
                    ncclResult_t r = g_nccl.AllGather(packed.p, gathered.p, n * K, ncclUint64, g->comm[s], st);  // the ONE exchange
                    if (r != ncclSuccess) { rc = nccl_fail(r, "ncclAllGather(top-K)"); break; }
                } else {
                    cudaMemcpyAsync(gathered.p, packed.p, n * K * 8, cudaMemcpyDeviceToDevice, st);
                }
                // every GPU merges, and writes back, its own slice of the batch's queries
                rc = topk_merge_packed_device(gathered.p, G, n, K, K, a, b, out_d.p, out_i.p, st);
                if (rc != RG_OK) break;
                if (b > a) {
                    cudaMemcpyAsync(ids + (q0 + a) * K, out_i.p + a * K, (b - a) * K * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
                    cudaMemcpyAsync(dists + (q0 + a) * K, out_d.p + a * K, (b - a) * K * sizeof(float), cudaMemcpyDeviceToHost, st);
                }
                e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) rc = cuda_fail(e, "rg_group_knn");
            }
        }
        q.release();
        tau.release();
        bound.release();
        dd.release();
        di.release();
        packed.release();
        gathered.release();
        out_d.release();
        out_i.release();
        if (kb) knn_base_free(kb);
        knn_release_workspace();
        return rc;
    }, &rv);
}

}  // extern "C"
