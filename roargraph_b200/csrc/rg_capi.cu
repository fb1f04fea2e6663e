// rg_capi.cu — the extern "C" boundary of librg_b200.so (include/rg_b200.h): index object, uploads, launches.
// No CPU fallback anywhere: every compute entry point needs a CUDA device and fails with RG_ECUDA otherwise.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cstdlib>

#include "rg_internal.h"

namespace rg {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();  // clear the sticky-less error state
    return RG_ECUDA;
}

// Device bookkeeping.  The library never picks a device on its own: every entry point works on the calling thread's
// current CUDA device (rg_init just makes one current), and index / kNN-base objects remember theirs — so several host
// threads can each drive their own GPU of one process (include/rg_b200.h, rg_group_*).
constexpr int kMaxDev = 64;
static std::mutex g_dev_mu;
static int g_dev_sms[kMaxDev] = {};   // 0 = not probed yet, -1 = unusable
static bool g_l2_fetch_done[kMaxDev] = {};

static int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (librg_b200 has no CPU fallback)");
        return RG_ECUDA;
    }
    int dev = 0;
    RG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) { set_error("CUDA device index out of range"); return RG_ECUDA; }
    std::lock_guard<std::mutex> g(g_dev_mu);
    if (g_dev_sms[dev] > 0) return RG_OK;
    cudaDeviceProp prop;
    RG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        g_dev_sms[dev] = -1;
        set_error("librg_b200 is built for sm_100a (B200) only");
        return RG_ECUDA;
    }
    g_dev_sms[dev] = prop.multiProcessorCount;
    // The beam search gathers 800-byte / 2 KB rows at random: RG_L2_FETCH=32|64|128 sets the granularity at which L2
    // fetches from HBM (a hint; the driver's default over-fetches around short rows)
    if (!g_l2_fetch_done[dev]) {
        g_l2_fetch_done[dev] = true;
        if (const char *v = getenv("RG_L2_FETCH")) {
            const size_t gr = (size_t)strtoul(v, nullptr, 10);
            if (gr == 32 || gr == 64 || gr == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gr);
            cudaGetLastError();
        }
    }
    return RG_OK;
}

}  // namespace rg

using namespace rg;

namespace rg {
int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 0;
    std::lock_guard<std::mutex> g(g_dev_mu);
    return g_dev_sms[dev] > 0 ? g_dev_sms[dev] : 0;
}
int ensure_device_ready() { return ensure_device(); }
int enter_index(rg_index *idx) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    int dev = 0;
    RG_CUDA(cudaGetDevice(&dev));
    if (dev != idx->device) RG_CUDA(cudaSetDevice(idx->device));
    return RG_OK;
}
}

int rg::ensure_tags(rg_index *idx, uint32_t slots) {
    if (slots <= idx->n_slots) return RG_OK;
    if (idx->d_tags) cudaFree(idx->d_tags);
    if (idx->d_slot_epoch) cudaFree(idx->d_slot_epoch);
    idx->d_tags = nullptr;
    idx->d_slot_epoch = nullptr;
    idx->n_slots = 0;
    const size_t bytes = (size_t)slots * idx->tag_stride;
    RG_CUDA(cudaMalloc(&idx->d_tags, bytes));
    RG_CUDA(cudaMalloc(&idx->d_slot_epoch, slots * sizeof(uint32_t)));
    RG_CUDA(cudaMemset(idx->d_tags, 0, bytes));
    {
        // Stagger the 8-bit epochs: a slot clears its whole tag array when the epoch wraps (VisitedList::reset), and
        // slots that all started at 0 would all wrap inside the same launch (a 2-5x slower batch every ~90 batches of
        // 10 000 queries); spread over the 255 phases the clears cost every launch the same ~1 %.
        std::vector<uint32_t> e0(slots);
        for (uint32_t i = 0; i < slots; ++i) e0[i] = (uint32_t)(((uint64_t)i * 97u) % 255u);
        RG_CUDA(cudaMemcpy(idx->d_slot_epoch, e0.data(), slots * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    idx->n_slots = slots;
    return RG_OK;
}

int rg::ensure_vtab(rg_index *idx, uint64_t words) {
    if (words <= idx->vtab_words) return RG_OK;
    if (idx->d_vtab) cudaFree(idx->d_vtab);
    idx->d_vtab = nullptr;
    idx->vtab_words = 0;
    RG_CUDA(cudaMalloc(&idx->d_vtab, words * sizeof(uint32_t)));
    RG_CUDA(cudaMemset(idx->d_vtab, 0xff, words * sizeof(uint32_t)));
    idx->vtab_words = words;
    return RG_OK;
}

// Number of query slots (resident warps) whose tag arrays fit in a bounded share of free HBM.
uint32_t rg::affordable_slots(rg_index *idx) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 64;
    size_t budget = (free_b + (size_t)idx->n_slots * idx->tag_stride) / 3;  // at most a third of what is free
    size_t s = budget / idx->tag_stride;
    if (s < 4) s = 4;
    if (s > max_search_slots(idx->sm_count)) s = max_search_slots(idx->sm_count);
    return (uint32_t)s;
}

static int fill_params(rg_index *idx, SearchParams &p, uint32_t nq, uint32_t k, uint32_t L) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    if (k == 0 || k > L) { set_error("L_pq must greater or equal than k"); return RG_EINVAL; }  // CLI :192-195
    if (L >= (1u << 20)) { set_error("L_pq too large"); return RG_EINVAL; }
    memset(&p, 0, sizeof(p));
    p.base = idx->d_base;
    p.nd = (uint32_t)idx->nd;
    p.dim = idx->dim;
    p.nq = nq;
    p.k = k;
    p.L = L;
    idx->last_k = k;
    p.metric_l2 = idx->metric == RG_METRIC_L2 ? 1 : 0;
    p.tag_stride = idx->tag_stride;
    p.counter = idx->d_counter;
    p.stats = idx->d_stats;
    return RG_OK;
}

static int check_latched(rg_index *idx) {
    unsigned long long st[4];
    RG_CUDA(cudaMemcpy(st, idx->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[3] & 1ull) {
        // same text as the reference's runtime_error (src/index_bipartite.cpp:2408-2412)
        char buf[96];
        snprintf(buf, sizeof(buf), "not enough results: %u, expected: %u", (unsigned)(st[3] >> 32), idx->last_k);
        set_error(buf);
        return RG_ENOTENOUGH;
    }
    return RG_OK;
}

extern "C" {

const char *rg_last_error(void) { return g_err.c_str(); }
int rg_abi_version(void) { return 1; }

int rg_init(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (librg_b200 has no CPU fallback)");
        return RG_ECUDA;
    }
    if (device < 0 || device >= n) { set_error("rg_init: no such CUDA device"); return RG_EINVAL; }
    RG_CUDA(cudaSetDevice(device));
    return ensure_device();
}

int rg_device_sm_count(void) {
    int rc = ensure_device();
    return rc == RG_OK ? sm_count() : rc;
}

int rg_index_create(rg_index **out, const float *base, uint64_t nd, uint32_t dim, int metric, int base_on_device) {
    if (!out || !base || nd == 0 || dim == 0) { set_error("rg_index_create: null/empty argument"); return RG_EINVAL; }
    if (dim % 8 != 0) { set_error("rg_index_create: dim must be padded to a multiple of 8 (util.h data_align)"); return RG_EINVAL; }
    if (nd >= 0x7fffffffull) { set_error("rg_index_create: nd must be < 2^31"); return RG_EINVAL; }
    if (metric != RG_METRIC_L2 && metric != RG_METRIC_INNER_PRODUCT && metric != RG_METRIC_COSINE) {
        set_error("rg_index_create: unknown metric");
        return RG_EINVAL;
    }
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    rg_index *idx = new rg_index();
    cudaGetDevice(&idx->device);
    idx->sm_count = sm_count();
    idx->nd = nd;
    idx->dim = dim;
    idx->metric = metric;
    idx->tag_stride = (nd + 15) & ~15ull;
    if (base_on_device) {
        idx->d_base = base;
    } else {
        float *d = nullptr;
        cudaError_t e = cudaMalloc(&d, nd * dim * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(d, base, nd * dim * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (d) cudaFree(d);
            delete idx;
            return cuda_fail(e, "base upload");
        }
        idx->d_base = d;
        idx->own_base = true;
    }
    cudaError_t e = cudaMalloc(&idx->d_counter, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_stats, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(idx->d_stats, 0, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        rg_index_destroy(idx);
        return cuda_fail(e, "index scratch");
    }
    *out = idx;
    return RG_OK;
}

int rg_index_create_from_fbin(rg_index **out, const char *path, int metric, rg_row_hook hook, void *user, uint64_t *out_nd,
                              uint32_t *out_dim) {
    if (!out || !path) { set_error("rg_index_create_from_fbin: null argument"); return RG_EINVAL; }
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    FILE *f = fopen(path, "rb");
    if (!f) { set_error(std::string("rg_index_create_from_fbin: cannot open ") + path); return RG_EINVAL; }
    uint32_t hdr[2] = {0, 0};
    if (fread(hdr, 4, 2, f) != 2 || hdr[0] == 0 || hdr[1] == 0) {
        fclose(f);
        set_error("Data file size wrong!");  // util.h:124
        return RG_EINVAL;
    }
    const uint64_t nd = hdr[0];
    const uint32_t dim = hdr[1], pdim = (dim + 7u) & ~7u;  // rows padded to x8 floats with zeros (util.h:37-75, :191-199)
    fseek(f, 0, SEEK_END);
    const uint64_t fsize = (uint64_t)ftell(f);
    if (fsize != 8 + nd * dim * sizeof(float)) {
        fclose(f);
        set_error("Data file size wrong!");
        return RG_EINVAL;
    }
    fseek(f, 8, SEEK_SET);
    float *d_base = nullptr;
    cudaError_t e = cudaMalloc(&d_base, nd * pdim * sizeof(float));
    if (e == cudaSuccess && pdim != dim) e = cudaMemset(d_base, 0, nd * pdim * sizeof(float));
    const uint64_t slab_rows = std::max<uint64_t>(1, std::min<uint64_t>(nd, (64ull << 20) / ((uint64_t)dim * 4u)));
    float *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaMallocHost(&pin[b], slab_rows * dim * sizeof(float));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
    }
    bool short_read = false;
    int slab = 0;
    for (uint64_t r0 = 0; r0 < nd && e == cudaSuccess; r0 += slab_rows, slab ^= 1) {
        const uint64_t n = std::min(slab_rows, nd - r0);
        e = cudaEventSynchronize(done[slab]);  // the copy that last used this staging buffer has finished
        if (e != cudaSuccess) break;
        if (fread(pin[slab], sizeof(float) * dim, n, f) != n) { short_read = true; break; }
        if (hook) hook(pin[slab], n, dim, user);  // e.g. the reference's normalize<float> per row for COSINE
        // pitch conversion on the copy engine: file rows of dim floats -> device rows of pdim floats (padding stays zero)
        e = cudaMemcpy2DAsync(d_base + r0 * pdim, (size_t)pdim * 4, pin[slab], (size_t)dim * 4, (size_t)dim * 4, n,
                              cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(done[slab], st);
    }
    if (e == cudaSuccess && !short_read) e = cudaStreamSynchronize(st);
    fclose(f);
    for (int b = 0; b < 2; ++b) {
        if (pin[b]) cudaFreeHost(pin[b]);
        if (done[b]) cudaEventDestroy(done[b]);
    }
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess || short_read) {
        if (d_base) cudaFree(d_base);
        if (short_read) { set_error("Data file size wrong!"); return RG_EINVAL; }
        return cuda_fail(e, "rg_index_create_from_fbin");
    }
    rg_index *idx = nullptr;
    rc = rg_index_create(&idx, d_base, nd, pdim, metric, 1);
    if (rc != RG_OK) {
        cudaFree(d_base);
        return rc;
    }
    idx->own_base = true;  // the index frees what this function allocated
    *out = idx;
    if (out_nd) *out_nd = nd;
    if (out_dim) *out_dim = pdim;
    return RG_OK;
}

int rg_index_set_graph(rg_index *idx, uint32_t ep, const uint64_t *row_off, const uint32_t *nbrs) {
    if (!idx || !row_off || (!nbrs && row_off[idx->nd] != 0)) { set_error("rg_index_set_graph: null argument"); return RG_EINVAL; }
    if (ep >= idx->nd) { set_error("rg_index_set_graph: entry point out of range"); return RG_EINVAL; }
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    const uint64_t nd = idx->nd;
    // Pass 1: duplicate-free degree of every row.  The reference's visited array makes a repeated id inside one
    // adjacency list a no-op (src/index_bipartite.cpp:2378), so dropping repeats at upload changes nothing observable.
    std::vector<uint32_t> tmp;
    uint32_t max_deg = 0;
    for (uint64_t u = 0; u < nd; ++u) {
        if (row_off[u + 1] < row_off[u]) { set_error("rg_index_set_graph: row_off not monotone"); return RG_EINVAL; }
        const uint64_t deg = row_off[u + 1] - row_off[u];
        if (deg > max_deg) max_deg = (uint32_t)std::min<uint64_t>(deg, 0xffffffffu);
    }
    if (max_deg > 4096) { set_error("rg_index_set_graph: out-degree above 4096 is not supported by the fixed-stride layout"); return RG_EINVAL; }
    const uint32_t stride = (max_deg + 1 + 7) & ~7u;  // [deg, ids...] padded to whole 32-byte sectors
    if (idx->d_adj) cudaFree(idx->d_adj);
    idx->d_adj = nullptr;
    RG_CUDA(cudaMalloc(&idx->d_adj, (size_t)nd * stride * sizeof(uint32_t)));
    // CSR -> fixed-stride rows in slabs: the host threads fill one pinned slab while the previous one is on its way to HBM
    // (a 10 M-node graph is 2.9 GB of rows; built single-threaded in pageable memory it cost more than the search sweep)
    const uint64_t slab_rows = std::max<uint64_t>(1, std::min<uint64_t>(nd, (64ull << 20) / ((uint64_t)stride * 4u)));
    uint32_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaMallocHost(&pin[b], slab_rows * stride * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
    }
    const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<uint32_t> max_dedup(nt, 0);
    std::vector<int> bad(nt, 0);
    int slab = 0;
    for (uint64_t u0 = 0; u0 < nd && e == cudaSuccess; u0 += slab_rows, slab ^= 1) {
        const uint64_t n = std::min(slab_rows, nd - u0);
        e = cudaEventSynchronize(done[slab]);  // the copy that last used this slab has finished
        if (e != cudaSuccess) break;
        uint32_t *host = pin[slab];
        std::vector<std::thread> th;
        const uint64_t per = (n + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                const uint64_t a = std::min<uint64_t>((uint64_t)t * per, n), b = std::min<uint64_t>(a + per, n);
                for (uint64_t i = a; i < b; ++i) {
                    const uint64_t u = u0 + i;
                    uint32_t *row = host + (size_t)i * stride;
                    uint32_t d = 0;
                    for (uint64_t j = row_off[u]; j < row_off[u + 1]; ++j) {
                        const uint32_t v = nbrs[j];
                        if (v >= nd) { bad[t] = 1; continue; }
                        bool dup = false;
                        for (uint32_t q = 0; q < d; ++q)
                            if (row[1 + q] == v) { dup = true; break; }
                        if (!dup) row[1 + d++] = v;
                    }
                    row[0] = d;
                    for (uint32_t q = d + 1; q < stride; ++q) row[q] = 0xffffffffu;
                    max_dedup[t] = std::max(max_dedup[t], d);
                }
            });
        for (auto &x : th) x.join();
        e = cudaMemcpyAsync(idx->d_adj + (size_t)u0 * stride, host, n * stride * sizeof(uint32_t), cudaMemcpyHostToDevice, idx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(done[slab], idx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    for (int b = 0; b < 2; ++b) {
        if (pin[b]) cudaFreeHost(pin[b]);
        if (done[b]) cudaEventDestroy(done[b]);
    }
    if (e != cudaSuccess) return cuda_fail(e, "rg_index_set_graph: adjacency upload");
    for (unsigned t = 0; t < nt; ++t)
        if (bad[t]) {
            cudaFree(idx->d_adj);
            idx->d_adj = nullptr;
            idx->has_graph = false;
            set_error("rg_index_set_graph: neighbour id out of range");
            return RG_EINVAL;
        }
    idx->adj_stride = stride;
    idx->max_deg = max_deg;
    idx->ep = ep;
    idx->has_graph = true;
    return RG_OK;
}

int rg_index_set_bipartite(rg_index *idx, uint64_t n_total, const uint64_t *row_off, const uint32_t *nbrs) {
    if (!idx || !row_off || n_total < idx->nd) { set_error("rg_index_set_bipartite: bad argument"); return RG_EINVAL; }
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    const uint64_t nd = idx->nd;
    if (!nbrs && row_off[n_total] != 0) { set_error("rg_index_set_bipartite: null neighbour array"); return RG_EINVAL; }
    for (uint64_t u = 0; u < n_total; ++u) {
        if (row_off[u + 1] < row_off[u]) { set_error("rg_index_set_bipartite: row_off not monotone"); return RG_EINVAL; }
        for (uint64_t j = row_off[u]; j < row_off[u + 1]; ++j) {
            const uint32_t v = nbrs[j];
            if (v >= n_total) { set_error("rg_index_set_bipartite: neighbour id out of range"); return RG_EINVAL; }
            // The graph is bipartite (qbaseNNbipartite, :262-273): a query node (u >= nd) points at base nodes, which the
            // search evaluates against the base array and tags in a visited array of nd entries (:326-334); a base node
            // points at query nodes, whose lists are walked as second hops.  A base -> base edge would make a base id a
            // "query node" and its neighbours' neighbours — possibly query-node ids >= nd — rows of the base array.
            if (u >= nd && v >= nd) { set_error("rg_index_set_bipartite: query node points at a query node"); return RG_EINVAL; }
            if (u < nd && v < nd) { set_error("rg_index_set_bipartite: base node points at a base node"); return RG_EINVAL; }
        }
    }
    // Device layout: the search only ever walks base node -> query node -> base node (:324-341), so every base node u gets
    // ONE contiguous list: the distinct base ids reachable through its query nodes, in first-occurrence order.  An id that
    // sits in several of u's second-hop lists is evaluated once by the reference too (its visited array) and the order
    // inside a hop is unobservable, so cmps, hops and results are unchanged — while the kernel reads one
    // coalesced id list per hop instead of chasing row_off / ids per query node, and probes every visited tag once.
    std::vector<uint64_t> ro(nd + 1, 0);
    {
        // pass 1 (parallel over base nodes): distinct two-hop ids per node, counted
        const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::vector<uint32_t>> chunk_ids(nt);
        std::vector<std::vector<uint32_t>> chunk_cnt(nt);
        std::vector<std::thread> th;
        const uint64_t per = (nd + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                const uint64_t u0 = std::min<uint64_t>((uint64_t)t * per, nd), u1 = std::min<uint64_t>(u0 + per, nd);
                std::vector<uint32_t> &out = chunk_ids[t];
                std::vector<uint32_t> &cnt = chunk_cnt[t];
                cnt.assign(u1 - u0, 0);
                std::vector<uint32_t> seen;  // open-addressing set of the current node's ids
                for (uint64_t u = u0; u < u1; ++u) {
                    uint64_t raw = 0;
                    for (uint64_t j = row_off[u]; j < row_off[u + 1]; ++j) raw += row_off[nbrs[j] + 1] - row_off[nbrs[j]];
                    if (!raw) continue;
                    uint64_t cap = 16;
                    while (cap < 2 * raw) cap <<= 1;
                    seen.assign(cap, 0xffffffffu);
                    const size_t start = out.size();
                    for (uint64_t j = row_off[u]; j < row_off[u + 1]; ++j) {
                        const uint32_t qn = nbrs[j];
                        for (uint64_t i = row_off[qn]; i < row_off[qn + 1]; ++i) {
                            const uint32_t v = nbrs[i];
                            uint64_t h = ((uint64_t)v * 0x9E3779B97F4A7C15ull >> 20) & (cap - 1);
                            while (seen[h] != 0xffffffffu && seen[h] != v) h = (h + 1) & (cap - 1);
                            if (seen[h] == v) continue;
                            seen[h] = v;
                            out.push_back(v);
                        }
                    }
                    cnt[u - u0] = (uint32_t)(out.size() - start);
                }
            });
        for (auto &x : th) x.join();
        // pass 2: offsets + concatenation in node order
        uint64_t total = 0;
        for (unsigned t = 0; t < nt; ++t) {
            const uint64_t u0 = std::min<uint64_t>((uint64_t)t * per, nd);
            for (size_t i = 0; i < chunk_cnt[t].size(); ++i) {
                ro[u0 + i] = total;
                total += chunk_cnt[t][i];
            }
        }
        ro[nd] = total;
        if (idx->d_row_off) cudaFree(idx->d_row_off);
        if (idx->d_csr) cudaFree(idx->d_csr);
        idx->d_row_off = nullptr;
        idx->d_csr = nullptr;
        RG_CUDA(cudaMalloc(&idx->d_row_off, ro.size() * sizeof(uint64_t)));
        RG_CUDA(cudaMalloc(&idx->d_csr, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
        RG_CUDA(cudaMemcpy(idx->d_row_off, ro.data(), ro.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        uint64_t at = 0;
        for (unsigned t = 0; t < nt; ++t) {
            if (chunk_ids[t].empty()) continue;
            RG_CUDA(cudaMemcpy(idx->d_csr + at, chunk_ids[t].data(), chunk_ids[t].size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            at += chunk_ids[t].size();
        }
    }
    idx->n_total = n_total;
    idx->has_bipartite = true;
    return RG_OK;
}

int rg_index_destroy(rg_index *idx) {
    if (!idx) return RG_OK;
    enter_index(idx);
    if (idx->own_base && idx->d_base) cudaFree((void *)idx->d_base);
    if (idx->d_adj) cudaFree(idx->d_adj);
    if (idx->d_row_off) cudaFree(idx->d_row_off);
    if (idx->d_csr) cudaFree(idx->d_csr);
    if (idx->d_tags) cudaFree(idx->d_tags);
    if (idx->d_vtab) cudaFree(idx->d_vtab);
    if (idx->d_slot_epoch) cudaFree(idx->d_slot_epoch);
    if (idx->d_counter) cudaFree(idx->d_counter);
    if (idx->d_stats) cudaFree(idx->d_stats);
    idx->q.release();
    idx->dists.release();
    idx->ids.release();
    idx->cmps.release();
    idx->hops.release();
    idx->init.release();
    idx->order_scratch.release();
    if (idx->stream) cudaStreamDestroy(idx->stream);
    delete idx;
    return RG_OK;
}

int rg_index_build_roargraph(rg_index *idx, const uint32_t *knn_ids, int knn_on_device, uint64_t nsq,
                             uint32_t knn_stride, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq, uint32_t n_rounds) {
    if (!idx || !knn_ids || nsq == 0 || knn_stride == 0) { set_error("rg_index_build_roargraph: null/empty argument"); return RG_EINVAL; }
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    const uint32_t *d_knn = knn_ids;
    uint32_t *tmp = nullptr;
    if (!knn_on_device) {
        RG_CUDA(cudaMalloc(&tmp, nsq * knn_stride * sizeof(uint32_t)));
        cudaError_t e = cudaMemcpy(tmp, knn_ids, nsq * knn_stride * sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(tmp); return cuda_fail(e, "kNN upload"); }
        d_knn = tmp;
    }
    int rc = build_roargraph_device(idx, d_knn, nsq, knn_stride, M_sq, M_pjbp, L_pjpq, n_rounds, idx->stream);
    if (tmp) cudaFree(tmp);
    return rc;
}

int rg_index_graph_size(rg_index *idx, uint64_t *nnz, uint32_t *ep, uint32_t *max_degree) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    if (!idx->has_graph) { set_error("projection graph not set"); return RG_ESTATE; }
    uint64_t total = 0;
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    int rc = adjacency_nnz_device(idx->d_adj, idx->adj_stride, idx->nd, &total);
    if (rc != RG_OK) return rc;
    if (nnz) *nnz = total;
    if (ep) *ep = idx->ep;
    if (max_degree) *max_degree = idx->max_deg;
    return RG_OK;
}

int rg_index_get_graph(rg_index *idx, uint64_t *row_off, uint32_t *nbrs) {
    if (!idx || !row_off) { set_error("rg_index_get_graph: null argument"); return RG_EINVAL; }
    if (!idx->has_graph) { set_error("projection graph not set"); return RG_ESTATE; }
    const uint64_t nd = idx->nd;
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    const uint32_t stride = idx->adj_stride;
    // stream the fixed-stride rows back in slabs and compact them into CSR
    const uint64_t slab = std::max<uint64_t>(1, (64ull << 20) / (stride * 4ull));
    std::vector<uint32_t> host(slab * stride);
    row_off[0] = 0;
    for (uint64_t u0 = 0; u0 < nd; u0 += slab) {
        const uint64_t n = std::min(slab, nd - u0);
        RG_CUDA(cudaMemcpy(host.data(), idx->d_adj + u0 * stride, n * stride * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t *row = host.data() + i * stride;
            const uint64_t b = row_off[u0 + i];
            if (nbrs) memcpy(nbrs + b, row + 1, (size_t)row[0] * 4);
            row_off[u0 + i + 1] = b + row[0];
        }
    }
    return RG_OK;
}

static int search_dev(rg_index *idx, bool bipartite, const float *d_q, uint32_t nq, uint32_t k, uint32_t L,
                      const uint32_t *d_init, uint32_t n_init, uint64_t seed, uint32_t *d_ids, float *d_dists,
                      uint32_t *d_cmps, uint32_t *d_hops, cudaStream_t s) {
    SearchParams p;
    int rc = fill_params(idx, p, nq, k, L);
    if (rc != RG_OK) return rc;
    if ((rc = enter_index(idx)) != RG_OK) return rc;
    if (nq == 0) return RG_OK;
    if (!d_q || !d_ids || !d_dists) { set_error("search: null buffer"); return RG_EINVAL; }
    if (bipartite) {
        if (!idx->has_bipartite) { set_error("search: bipartite graph not set"); return RG_ESTATE; }
        if (n_init == 0 || n_init > (uint32_t)kCandCap) { set_error("search: n_init must be in [1, 96]"); return RG_EINVAL; }
        p.row_off = idx->d_row_off;
        p.csr_nbrs = idx->d_csr;
        p.init_ids = d_init;
        p.n_init = n_init;
        p.seed = seed;
    } else {
        if (!idx->has_graph) { set_error("search: projection graph not set"); return RG_ESTATE; }
        p.adj = idx->d_adj;
        p.adj_stride = idx->adj_stride;
        p.ep = idx->ep;
        const char *pf = getenv("RG_PREFETCH_ADJ");
        // doubles as the adjacency-prefetch switch; off by default: at 24 resident warps per SM the extra L2 fills cost
        // more DRAM bandwidth than the shorter dependent read saves (profiles/r1_tune_10M.csv)
        p.n_init = (pf && pf[0] == '1') ? 1u : (pf && pf[0] == '2') ? 2u : 0u;  // 1: on queue insertion, 2: the next-best node at hop start
        p.drop_ep = 1;  // SearchRoarGraph does not mark its entry point visited (src/index_bipartite.cpp:2349 is commented out)
    }
    p.queries = d_q;
    p.out_ids = d_ids;
    p.out_dists = d_dists;
    p.out_cmps = d_cmps;
    p.out_hops = d_hops;
    SearchLaunch l;
    // RG_VISITED=smem: the exact hashed visited set (same layout as the L2-resident table below) inside every warp's
    // shared-memory working set — no DRAM transaction and no DRAM latency for the visited checks, at the price of resident
    // warps.  Only where a table for ~1.25 x the expected evaluations fits (short queues); else the tag arrays.
    {
        const char *vm = getenv("RG_VISITED");
        if (!bipartite && vm && vm[0] == 's') {
            uint32_t bits = 1;
            while (bits < 31 && (1ull << bits) < idx->nd) ++bits;
            const uint64_t want_entries = (uint64_t)(48.0 * L + 256.0) * 5 / 4;
            uint32_t bbits = 0;
            while ((16ull << bbits) < want_entries) ++bbits;
            if (const char *fb = getenv("RG_VTAB_BBITS")) bbits = (uint32_t)strtoul(fb, nullptr, 10);
            if (bits > bbits + 14) bbits = bits - 14;
            if (bbits > bits) bbits = bits;
            const uint64_t words = 8ull << bbits;
            if (words * 4 <= (64u << 10)) {
                p.vt_smem = 1;
                p.vt_words = (uint32_t)words;
                p.vt_bbits = bbits;
                p.vt_rbits = bits - bbits;
                p.vt_mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
                p.vt_mul = 0x9E3779B1u;
                uint32_t inv = p.vt_mul;
                for (int it = 0; it < 5; ++it) inv *= 2u - p.vt_mul * inv;
                p.vt_mulinv = inv;
            }
        }
    }
    // cudaMemGetInfo (affordable_slots) costs milliseconds and serialises with the device: only ask when this launch wants
    // more slots than already have tag arrays
    rc = plan_search(p, l, idx->sm_count, 0xffffffffu);
    if (rc != RG_OK) return rc;
    if (l.grid * l.warps_per_cta > idx->n_slots) {
        const uint32_t cap = std::max(idx->n_slots, affordable_slots(idx));
        if (l.grid * l.warps_per_cta > cap) {
            rc = plan_search(p, l, idx->sm_count, cap);
            if (rc != RG_OK) return rc;
        }
    }
    rc = ensure_tags(idx, l.grid * l.warps_per_cta);
    if (rc != RG_OK) return rc;
    p.tags = idx->d_tags;
    p.slot_epoch = idx->d_slot_epoch;
    p.vtab = nullptr;
    if (!p.vt_smem) p.vt_words = 0;
    if (!bipartite && !p.vt_smem) {
        // Visited set of SearchRoarGraph: a per-slot hash table small enough for all slots' tables to stay in L2, so the
        // visited checks stop costing DRAM transactions (the byte-per-node tag array costs one random 32-byte DRAM read and
        // one write-back per examined neighbour: ~25 % of the kernel's DRAM transactions).  Exact: 16-bit entries hold the
        // low bits of a bijective hash of the id, the bucket index holds the rest.  Tables that would not fit the L2
        // budget (large L_pq) fall back to the tag arrays; so does a query whose bucket pair overflows, mid-flight.
        const char *vm = getenv("RG_VISITED");  // "hash" opts in; default = the tag arrays (the table's extra L2 round trips
        const bool force_tags = !vm || vm[0] != 'h', force_hash = vm && vm[0] == 'h';  // cost more than its DRAM savings, r1 tune)
        const char *mb = getenv("RG_VTAB_MB");
        const uint64_t budget = (mb ? strtoull(mb, nullptr, 10) : 64ull) << 20;
        uint32_t bits = 1;
        while (bits < 31 && (1ull << bits) < idx->nd) ++bits;
        const uint64_t want_entries = (uint64_t)(48.0 * L + 256.0) * 5 / 4;  // ~1.25 x the expected distance evaluations
        uint32_t bbits = 0;
        while ((16ull << bbits) < want_entries) ++bbits;
        if (const char *fb = getenv("RG_VTAB_BBITS")) bbits = (uint32_t)strtoul(fb, nullptr, 10);  // test hook: tiny tables overflow
        if (bits > bbits + 14) bbits = bits - 14;  // the remainder must fit 14 bits
        if (bbits > bits) bbits = bits;
        const uint64_t words = 8ull << bbits;
        const uint64_t slots = (uint64_t)l.grid * l.warps_per_cta;
        if (!force_tags && (force_hash || slots * words * 4 <= budget) && words * 4 <= (64u << 20)) {
            rc = ensure_vtab(idx, slots * words);
            if (rc != RG_OK) return rc;
            p.vtab = idx->d_vtab;
            p.vt_words = (uint32_t)words;
            p.vt_bbits = bbits;
            p.vt_rbits = bits - bbits;
            p.vt_mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
            p.vt_mul = 0x9E3779B1u;  // odd: id -> id * mul mod 2^bits is a bijection
            uint32_t inv = p.vt_mul;  // Newton iteration for the inverse modulo 2^32
            for (int it = 0; it < 5; ++it) inv *= 2u - p.vt_mul * inv;
            p.vt_mulinv = inv;
        }
    }
    idx->last_launch = l;
    // RG_QUERY_ORDER=1 (opt-in): hand the batch out grouped by region (rg_search.cu, order_queries).  Measured on the 10 M x 200
    // workload it does not pay: the search kernel itself is unchanged within noise (the ~2 000 resident queries touch far more
    // than the 126 MB L2 holds even when they are neighbours) and the landmark pass costs 1.3 ms per 10 000 queries
    // (1.02 M vs 1.18 M QPS on the 10 k batch, 1.14 M vs 1.24 M on the 100 k batch).
    {
        const char *oe = getenv("RG_QUERY_ORDER");
        const bool order_env = oe && oe[0] == '1';
        if (order_env && nq >= 4096 && idx->dim % 4 == 0) {
            if ((rc = idx->order_scratch.reserve(2 * (size_t)nq + kLandmarks + 1)) != RG_OK) return rc;
            rc = order_queries(idx->d_base, idx->nd, idx->dim, p.metric_l2, d_q, nq, idx->order_scratch.p, s);
            if (rc != RG_OK) return rc;
            p.order = idx->order_scratch.p + nq;
        }
    }
    return bipartite ? launch_bipartite(p, l, s) : launch_roargraph(p, l, s);
}

int rg_search_roargraph_dev(rg_index *idx, const float *d_queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                            uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops, void *stream) {
    return search_dev(idx, false, d_queries, nq, k, L_pq, nullptr, 0, 0, d_ids, d_dists, d_cmps, d_hops,
                      (cudaStream_t)stream);
}

int rg_search_bipartite_dev(rg_index *idx, const float *d_queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                            const uint32_t *d_init_ids, uint32_t n_init, uint64_t seed, uint32_t *d_ids,
                            float *d_dists, uint32_t *d_cmps, uint32_t *d_hops, void *stream) {
    return search_dev(idx, true, d_queries, nq, k, L_pq, d_init_ids, n_init, seed, d_ids, d_dists, d_cmps, d_hops,
                      (cudaStream_t)stream);
}

int rg_index_check(rg_index *idx, void *stream) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    RG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return check_latched(idx);
}

static int search_host(rg_index *idx, bool bipartite, const float *q, uint32_t nq, uint32_t k, uint32_t L,
                       const uint32_t *init_ids, uint32_t n_init, uint64_t seed, uint32_t *ids, float *dists,
                       uint32_t *cmps, uint32_t *hops) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    if (!q || !ids || !dists) { set_error("search: null buffer"); return RG_EINVAL; }
    int rc;
    if ((rc = enter_index(idx)) != RG_OK) return rc;
    if ((rc = idx->q.reserve((size_t)nq * idx->dim)) != RG_OK) return rc;
    if ((rc = idx->ids.reserve((size_t)nq * k)) != RG_OK) return rc;
    if ((rc = idx->dists.reserve((size_t)nq * k)) != RG_OK) return rc;
    if ((rc = idx->cmps.reserve(nq)) != RG_OK) return rc;
    if ((rc = idx->hops.reserve(nq)) != RG_OK) return rc;
    cudaStream_t s = idx->stream;
    static const bool trace = getenv("RG_TRACE_E2E") != nullptr;  // per-phase device times of the host-buffer call, to stderr
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace) {
        for (auto &e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], s);
    }
    RG_CUDA(cudaMemcpyAsync(idx->q.p, q, (size_t)nq * idx->dim * sizeof(float), cudaMemcpyHostToDevice, s));
    if (trace) cudaEventRecord(ev[1], s);
    const uint32_t *d_init = nullptr;
    if (bipartite && init_ids) {
        if ((rc = idx->init.reserve((size_t)nq * n_init)) != RG_OK) return rc;
        RG_CUDA(cudaMemcpyAsync(idx->init.p, init_ids, (size_t)nq * n_init * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        d_init = idx->init.p;
    }
    rc = search_dev(idx, bipartite, idx->q.p, nq, k, L, d_init, n_init, seed, idx->ids.p, idx->dists.p, idx->cmps.p,
                    idx->hops.p, s);
    if (rc != RG_OK) return rc;
    if (trace) cudaEventRecord(ev[2], s);
    RG_CUDA(cudaMemcpyAsync(ids, idx->ids.p, (size_t)nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    RG_CUDA(cudaMemcpyAsync(dists, idx->dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (cmps) RG_CUDA(cudaMemcpyAsync(cmps, idx->cmps.p, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (hops) RG_CUDA(cudaMemcpyAsync(hops, idx->hops.p, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (trace) cudaEventRecord(ev[3], s);
    RG_CUDA(cudaStreamSynchronize(s));
    if (trace) {
        float h2d = 0, kern = 0, d2h = 0;
        cudaEventElapsedTime(&h2d, ev[0], ev[1]);
        cudaEventElapsedTime(&kern, ev[1], ev[2]);
        cudaEventElapsedTime(&d2h, ev[2], ev[3]);
        fprintf(stderr, "[rg e2e] nq=%u h2d %.3f ms, search %.3f ms, d2h %.3f ms\n", nq, h2d, kern, d2h);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return check_latched(idx);
}

int rg_search_roargraph(rg_index *idx, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq, uint32_t *ids,
                        float *dists, uint32_t *cmps, uint32_t *hops) {
    return search_host(idx, false, queries, nq, k, L_pq, nullptr, 0, 0, ids, dists, cmps, hops);
}

int rg_search_bipartite(rg_index *idx, const float *queries, uint32_t nq, uint32_t k, uint32_t L_pq,
                        const uint32_t *init_ids, uint32_t n_init, uint64_t seed, uint32_t *ids, float *dists,
                        uint32_t *cmps, uint32_t *hops) {
    return search_host(idx, true, queries, nq, k, L_pq, init_ids, n_init, seed, ids, dists, cmps, hops);
}

int rg_search_last_stats(rg_index *idx, uint64_t *total_cmps, uint64_t *total_hops, uint64_t *adj_words) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    { int rc_dev_ = enter_index(idx); if (rc_dev_ != RG_OK) return rc_dev_; }
    RG_CUDA(cudaDeviceSynchronize());
    unsigned long long st[4];
    RG_CUDA(cudaMemcpy(st, idx->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (total_cmps) *total_cmps = st[0];
    if (total_hops) *total_hops = st[1];
    if (adj_words) *adj_words = st[2];
    return RG_OK;
}

int rg_search_last_launch(rg_index *idx, uint32_t *grid, uint32_t *block, uint32_t *smem_bytes, uint32_t *rows_per_batch) {
    if (!idx) { set_error("null index"); return RG_EINVAL; }
    if (grid) *grid = idx->last_launch.grid;
    if (block) *block = idx->last_launch.block;
    if (smem_bytes) *smem_bytes = idx->last_launch.smem_bytes;
    if (rows_per_batch) *rows_per_batch = idx->last_launch.rows_per_batch;
    return RG_OK;
}

int rg_knn_dev(const float *d_base, uint64_t nb, const float *d_queries, uint64_t nq, uint32_t dim, uint32_t K,
               int metric, int32_t id_offset, int32_t *d_ids, float *d_dists, void *stream) {
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    return knn_device(d_base, nb, d_queries, nq, dim, K, metric, id_offset, d_ids, d_dists, (cudaStream_t)stream);
}

int rg_knn(const float *base, uint64_t nb, const float *queries, uint64_t nq, uint32_t dim, uint32_t K, int metric,
           int32_t id_offset, int32_t *ids, float *dists) {
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    if (!base || !queries || !ids || !dists || nb == 0 || dim == 0) { set_error("rg_knn: null/empty argument"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    // The base goes to HBM once; the queries stream through in batches (the reference walks them 1 024 at a time against
    // each base part, compute_groundtruth.cpp:187), so device memory is bounded by base + one batch however many queries
    // there are (10 M x 200 queries + their K = 100 results would otherwise pin 16 GB beside the base).
    uint64_t batch = 1u << 20;
    if (const char *hb = getenv("RG_KNN_HOST_BATCH")) batch = std::max<uint64_t>(1, strtoull(hb, nullptr, 10));
    batch = std::min(batch, nq);
    float *d_b = nullptr, *d_q = nullptr, *d_d = nullptr;
    int32_t *d_i = nullptr;
    cudaError_t e = cudaMalloc(&d_b, nb * dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_q, batch * dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_d, batch * K * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_i, batch * K * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_b, base, nb * dim * sizeof(float), cudaMemcpyHostToDevice);
    rc = RG_OK;
    for (uint64_t q0 = 0; q0 < nq && e == cudaSuccess && rc == RG_OK; q0 += batch) {
        const uint64_t n = std::min(batch, nq - q0);
        e = cudaMemcpy(d_q, queries + q0 * dim, n * dim * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) break;
        rc = knn_device(d_b, nb, d_q, n, dim, K, metric, id_offset, d_i, d_d, nullptr);
        if (rc == RG_OK) e = cudaDeviceSynchronize();
        if (rc == RG_OK && e == cudaSuccess) e = cudaMemcpy(ids + q0 * K, d_i, n * K * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (rc == RG_OK && e == cudaSuccess) e = cudaMemcpy(dists + q0 * K, d_d, n * K * sizeof(float), cudaMemcpyDeviceToHost);
    }
    if (d_b) cudaFree(d_b);
    if (d_q) cudaFree(d_q);
    if (d_d) cudaFree(d_d);
    if (d_i) cudaFree(d_i);
    if (e != cudaSuccess) return cuda_fail(e, "rg_knn");
    return rc;
}

int rg_knn_base_create(rg_knn_base **out, const float *base, uint64_t nb, uint32_t dim, int metric, int base_on_device) {
    if (!out || !base || nb == 0 || dim == 0) { set_error("rg_knn_base_create: null/empty argument"); return RG_EINVAL; }
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    const float *d_base = base;
    float *own = nullptr;
    if (!base_on_device) {
        cudaError_t e = cudaMalloc(&own, nb * dim * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(own, base, nb * dim * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (own) cudaFree(own);
            return cuda_fail(e, "rg_knn_base_create: base upload");
        }
        d_base = own;
    }
    rg_knn_base *kb = nullptr;
    rc = knn_base_new(&kb, d_base, own, nb, dim, metric);
    if (rc != RG_OK) return rc;
    *out = kb;
    return RG_OK;
}

int rg_knn_base_destroy(rg_knn_base *kb) {
    knn_base_free(kb);
    return RG_OK;
}

int rg_knn_base_search_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                           int32_t *d_ids, float *d_dists, void *stream) {
    if (!kb) { set_error("rg_knn_base_search_dev: null base"); return RG_EINVAL; }
    return knn_base_search(kb, d_queries, nq, K, id_offset, d_ids, d_dists, (cudaStream_t)stream);
}

int rg_knn_base_seed_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, uint32_t rank, float *d_tau, void *stream) {
    if (!kb) { set_error("rg_knn_base_seed_dev: null base"); return RG_EINVAL; }
    return knn_base_seed(kb, d_queries, nq, K, rank, d_tau, (cudaStream_t)stream);
}

int rg_knn_base_search_seeded_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                                  const float *d_tau, uint64_t bound_rows, int32_t *d_ids, float *d_dists, void *stream) {
    if (!kb) { set_error("rg_knn_base_search_seeded_dev: null base"); return RG_EINVAL; }
    return knn_base_search(kb, d_queries, nq, K, id_offset, d_ids, d_dists, (cudaStream_t)stream, d_tau, bound_rows);
}

int rg_knn_base_search_resync_dev(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset,
                                  const float *d_tau, uint64_t bound_rows, uint32_t rank, uint64_t resync_rows, float *d_bound,
                                  rg_knn_bound_hook hook, void *user, int32_t *d_ids, float *d_dists, void *stream) {
    if (!kb) { set_error("rg_knn_base_search_resync_dev: null base"); return RG_EINVAL; }
    if (!d_tau) { set_error("rg_knn_base_search_resync_dev: needs the seed bound (rg_knn_base_seed_dev)"); return RG_EINVAL; }
    KnnResync rs{rank, resync_rows, d_bound, hook, user};
    return knn_base_search(kb, d_queries, nq, K, id_offset, d_ids, d_dists, (cudaStream_t)stream, d_tau, bound_rows, &rs);
}

int rg_knn_profile(double *out7) {
    if (!out7) { set_error("rg_knn_profile: null buffer"); return RG_EINVAL; }
    const KnnProfile &p = knn_last_profile();
    out7[0] = p.convert_base_ms;
    out7[1] = p.convert_queries_ms;
    out7[2] = p.seed_ms;
    out7[3] = p.filter_ms;
    out7[4] = p.merge_ms;
    out7[5] = (double)p.chunks;
    out7[6] = (double)p.overflow_retries;
    return RG_OK;
}

int rg_knn_profile_reset(void) {
    knn_reset_profile();
    return RG_OK;
}

int rg_knn_trace(int on) {
    knn_set_trace(on);
    return RG_OK;
}

int rg_knn_release_workspace(void) {
    knn_release_workspace();
    return RG_OK;
}

int rg_topk_merge_dev(const float *d_in_dists, const uint32_t *d_in_ids, uint32_t G, uint64_t nq, uint32_t K,
                      float *d_out_dists, uint32_t *d_out_ids, void *stream) {
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    return topk_merge_device(d_in_dists, d_in_ids, G, nq, K, d_out_dists, d_out_ids, (cudaStream_t)stream);
}

int rg_topk_pack_dev(const float *d_dists, const uint32_t *d_ids, uint64_t n, uint32_t id_offset, uint64_t *d_packed, void *stream) {
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    return topk_pack_device(d_dists, d_ids, n, id_offset, reinterpret_cast<unsigned long long *>(d_packed), (cudaStream_t)stream);
}

int rg_topk_merge_packed_dev(const uint64_t *d_packed, uint32_t G, uint64_t nq, uint32_t K_in, uint32_t K_out, uint64_t q_lo,
                             uint64_t q_hi, float *d_out_dists, uint32_t *d_out_ids, void *stream) {
    int rc = ensure_device();
    if (rc != RG_OK) return rc;
    return topk_merge_packed_device(reinterpret_cast<const unsigned long long *>(d_packed), G, nq, K_in, K_out, q_lo, q_hi,
                                    d_out_dists, d_out_ids, (cudaStream_t)stream);
}

}  // extern "C"
