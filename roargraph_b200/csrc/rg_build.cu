// rg_build.cu — BuildRoarGraph on sm_100a: a bulk-synchronous restatement of the reference's index construction
// (IndexBipartite::BuildRoarGraph -> LinkProjection, src/index_bipartite.cpp:143-233, :1043-1277; paths relative to
// /root/reference).  SURVEY.md 8(f)-1: the connectivity-enhancement search is the same beam-search kernel as the
// query path (rg_search.cu, mode 2) and is 85-94 % of the reference's build time.
//
// The reference mutates the graph from 64 OpenMP threads under per-node mutexes, so its output depends on thread
// timing; parity for this row is therefore statistical (degree distribution, recall at a given L_pq), not id-exact.
// Here every phase is a deterministic data-parallel pass:
//   1a  projection      per training query: pivot = its nearest base row; prune its remaining M_sq-1 neighbours
//                       around the pivot (PruneBiSearchBaseGetBase :1612-1694); the last query of a pivot wins (:1093).
//   1b  reverse edges   twice (ProjectionAddReverse inside the loop :1095 and over all nodes :1105-1108): every edge
//                       u->v offers u to v; lists above M_pjbp are re-pruned (PruneProjectionReverseCandidates :1527-1610).
//   3   connectivity    in rounds over contiguous node ranges: beam search of every base row on the supply graph at
//                       L_pjpq (SearchProjectionGraphInternal :1279-1350), prune the expanded set
//                       (PruneProjectionBaseSearchCandidates :1846-1940) -> supply[v]; reverse edges with cap 2*M_pjbp
//                       re-pruned to M_pjbp (SupplyAddReverse :1352-1392, PruneProjectionInternalReverseCandidates :1434-1525).
//   4   supply lists above M_pjbp re-pruned (:1224-1247); final row = projection ++ (supply \ projection) (:1251-1267).
// One CTA per node does candidate assembly, distances to the node, a bitonic sort by (distance, id) and the greedy
// occlusion scan with the selected rows staged in shared memory.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>

#include <cub/cub.cuh>

#include "rg_internal.h"

namespace rg {
namespace {

constexpr int kPruneThreads = 128;
constexpr uint32_t kMaxCand = 1024;
constexpr uint32_t kFullMask = 0xffffffffu;

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct PruneParams {
    const float *base;
    uint32_t dim, nd;
    int l2;
    int mode;  // 1 projection from kNN, 2 projection reverse, 3 supply from search, 4 supply reverse, 5 supply re-prune
    uint32_t node_lo, n_nodes, M;
    const uint32_t *knn;
    uint32_t knn_stride, M_sq;
    const int *owner;
    uint32_t *qlist;      // mode 1: the pruned list of every training query (nsq x proj_stride)
    int offers_by_query;  // mode 2: the offers' low word is a training-query index (the time of the push), not a node id
    uint32_t *proj, *supply;
    uint32_t proj_stride, supply_stride;
    const unsigned long long *in_keys;
    const uint64_t *in_off;
    const uint32_t *exp_ids;
    const float *exp_dists;
    const uint32_t *exp_cnt;
    uint32_t exp_cap;
};

// warp-cooperative internal distance between two rows (a in shared or global memory, b in shared memory)
__device__ __forceinline__ float warp_distance(const float *a, const float *b, uint32_t dim4, int l2, int lane) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    const float4 *b4 = reinterpret_cast<const float4 *>(b);
    float s = 0.f;
    for (uint32_t i = lane; i < dim4; i += 32) {
        const float4 x = a4[i], y = b4[i];
        if (l2) {
            const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
            s = fmaf(d0, d0, s);
            s = fmaf(d1, d1, s);
            s = fmaf(d2, d2, s);
            s = fmaf(d3, d3, s);
        } else {
            s = fmaf(x.x, y.x, s);
            s = fmaf(x.y, y.y, s);
            s = fmaf(x.z, y.z, s);
            s = fmaf(x.w, y.w, s);
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    return l2 ? s : -s;
}

__device__ __forceinline__ bool list_has(const uint32_t *list, uint32_t n, uint32_t id) {
    for (uint32_t i = 0; i < n; ++i)
        if (list[i] == id) return true;
    return false;
}

// Shared-memory layout (dynamic): keys[kMaxCand] u64 | raw ids[kMaxCand] | own[128] | skip[128] | result[128] |
// vrow[dim] | prow[dim] | sel[M * dim]
struct PruneSmem {
    unsigned long long *keys;
    uint32_t *raw, *own, *skip, *result;
    float *vrow, *prow, *sel;
};

// The occlusion prune shared by every phase (:1871-1890 and its siblings): raw[0..n) candidate ids of node v (keys already
// hold (distance to v, id) when have_dists) -> sorted by (distance, id), greedy scan "candidate j is dropped when a selected
// row is closer to it than v is", at most cap survivors, optionally padded back to cap with the nearest remaining
// candidates (`fill`).  result[0..returned) = the new list.  Called by the whole CTA.
__device__ uint32_t prune_candidates(const PruneParams &p, const PruneSmem &sm, uint32_t v, uint32_t n, uint32_t cap, bool fill,
                                     bool have_dists, uint32_t n_skip) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dim4 = p.dim >> 2;
    unsigned long long *keys = sm.keys;
    if (!have_dists) {
        for (uint32_t c = warp; c < n; c += kPruneThreads / 32) {
            const uint32_t id = sm.raw[c];
            if (id >= p.nd) {  // not a base row (e.g. a -1 filler in a short kNN list)
                if (lane == 0) keys[c] = ~0ull;
                continue;
            }
            const float d = warp_distance(p.base + (size_t)id * p.dim, sm.vrow, dim4, p.l2, lane);
            if (lane == 0) keys[c] = ((unsigned long long)f2ord(d) << 32) | id;
        }
    }
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = n + tid; i < m; i += kPruneThreads) keys[i] = ~0ull;
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < (m >> 1); i += kPruneThreads) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    // greedy occlusion scan over the sorted candidates
    uint32_t nres = 0;
    uint32_t prev = 0xffffffffu;
    for (uint32_t i = 0; i < n && nres < cap; ++i) {
        const unsigned long long key = keys[i];
        const uint32_t id = (uint32_t)key;
        if (id == v || id == prev || id >= p.nd) continue;
        prev = id;
        if (nres == 0 && n_skip && list_has(sm.skip, n_skip, id)) continue;
        const float pd = ord2f((uint32_t)(key >> 32));
        for (uint32_t j = tid; j < dim4; j += kPruneThreads)
            reinterpret_cast<float4 *>(sm.prow)[j] = __ldg(reinterpret_cast<const float4 *>(p.base + (size_t)id * p.dim) + j);
        __syncthreads();
        int occ = 0;
        for (uint32_t r = warp; r < nres; r += kPruneThreads / 32) {
            const float djk = warp_distance(sm.prow, sm.sel + (size_t)r * p.dim, dim4, p.l2, lane);
            if (djk < pd) {
                occ = 1;
                break;
            }
        }
        occ = __syncthreads_or(occ);
        if (!occ) {
            for (uint32_t j = tid; j < dim4; j += kPruneThreads)
                reinterpret_cast<float4 *>(sm.sel + (size_t)nres * p.dim)[j] = reinterpret_cast<const float4 *>(sm.prow)[j];
            if (tid == 0) sm.result[nres] = id;
            ++nres;
        }
        __syncthreads();
    }
    if (fill) {  // pad with the nearest remaining candidates (:1680-1688, :1603-1607); every thread computes the same count
        __shared__ uint32_t s_fill;
        if (tid == 0) {
            uint32_t pv = 0xffffffffu, nr = nres;
            for (uint32_t i = 0; i < n && nr < cap; ++i) {
                const uint32_t id = (uint32_t)keys[i];
                if (id == v || id == pv || id >= p.nd) continue;
                pv = id;
                if (!list_has(sm.result, nr, id)) sm.result[nr++] = id;
            }
            s_fill = nr;
        }
        __syncthreads();
        nres = s_fill;
    }
    __syncthreads();
    return nres;
}

__global__ void __launch_bounds__(kPruneThreads) prune_kernel(const PruneParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    PruneSmem sm;
    sm.keys = reinterpret_cast<unsigned long long *>(smem);
    sm.raw = reinterpret_cast<uint32_t *>(sm.keys + kMaxCand);
    sm.own = sm.raw + kMaxCand;
    sm.skip = sm.own + 128;
    sm.result = sm.skip + 128;
    sm.vrow = reinterpret_cast<float *>(sm.result + 128);
    sm.prow = sm.vrow + p.dim;
    sm.sel = sm.prow + p.dim;
    unsigned long long *keys = sm.keys;
    uint32_t *raw = sm.raw, *own = sm.own, *skip = sm.skip, *result = sm.result;
    __shared__ uint32_t s_n, s_pos_lo, s_more;

    const int tid = threadIdx.x;
    const uint32_t v = p.mode == 1 ? p.knn[(size_t)blockIdx.x * p.knn_stride] : p.node_lo + blockIdx.x;
    const uint32_t dim4 = p.dim >> 2;
    const uint32_t M = p.M;
    if (v >= p.nd) {  // a training query without a valid nearest base row
        if (p.mode == 1 && tid == 0) p.qlist[(size_t)blockIdx.x * p.proj_stride] = 0;
        return;
    }

    if (p.mode == 2 || p.mode == 4) {
        // Reverse edges (ProjectionAddReverse :1394-1432, SupplyAddReverse :1352-1392) replayed in the reference's order: the
        // offers of this round reach node v in ascending source id (the reference walks the sources in id order); a source not
        // yet in the list is appended while the list has room (cap = M_pjbp for the projection graph, 2 * M_pjbp for the
        // supply graph); the first one that does not fit triggers a prune of list + that source back to <= M_pjbp, after which
        // the list has room again.  (One prune over list + ALL offers — what a bulk-synchronous round suggests — leaves every
        // busy list at <= M_pjbp and loses 6 % of the final degree against the reference.)
        uint32_t *grow = (p.mode == 2) ? p.proj + (size_t)v * p.proj_stride : p.supply + (size_t)v * p.supply_stride;
        const uint32_t keep_cap = (p.mode == 2) ? M : 2 * M;
        const uint64_t ib = p.in_off[v], ie = p.in_off[v + 1];
        if (ib == ie) return;
        const uint32_t n_own = min(grow[0], 128u);
        for (uint32_t i = tid; i < n_own; i += kPruneThreads) raw[i] = grow[1 + i];
        for (uint32_t i = tid; i < dim4; i += kPruneThreads)
            reinterpret_cast<float4 *>(sm.vrow)[i] = __ldg(reinterpret_cast<const float4 *>(p.base + (size_t)v * p.dim) + i);
        if (tid == 0) {
            s_n = n_own;
            s_pos_lo = 0;
        }
        __syncthreads();
        const uint32_t n_off = (uint32_t)min((uint64_t)0xffffffffu, ie - ib);
        bool changed = false;
        for (;;) {
            if (tid == 0) {  // append phase: offers in order, until the list is full and one more distinct source is waiting
                uint32_t n = s_n, pos = s_pos_lo, more = 0;
                while (pos < n_off) {
                    uint32_t src = (uint32_t)p.in_keys[ib + pos];
                    if (p.offers_by_query) {
                        // phase 1 in the reference's order: training query `src` replaced its pivot's list and pushed the pivot
                        // to the members of its pruned list at time `src`; a push that precedes the LAST replacement of v's own
                        // list (v being the pivot of query owner[v]) was wiped by it (:1093)
                        if ((int)src <= p.owner[v]) {
                            ++pos;
                            continue;
                        }
                        src = p.knn[(size_t)src * p.knn_stride];
                    }
                    if (src == v || list_has(raw, n, src)) {
                        ++pos;
                        continue;
                    }
                    if (n < keep_cap) {
                        raw[n++] = src;
                        ++pos;
                        continue;
                    }
                    raw[n] = src;  // does not fit: prune list + this source
                    more = 1;
                    ++pos;
                    break;
                }
                s_n = n;
                s_pos_lo = pos;
                s_more = more;
            }
            __syncthreads();
            const uint32_t n = s_n;
            if (n != n_own) changed = true;
            if (!s_more) break;
            const uint32_t nres = prune_candidates(p, sm, v, n + 1, M, /*fill=*/p.mode == 2, false, 0);
            for (uint32_t i = tid; i < nres; i += kPruneThreads) raw[i] = result[i];
            if (tid == 0) s_n = nres;
            changed = true;
            __syncthreads();
        }
        if (changed) {
            const uint32_t n = s_n;
            for (uint32_t i = tid; i < n; i += kPruneThreads) grow[1 + i] = raw[i];
            __syncthreads();
            if (tid == 0) grow[0] = n;
        }
        return;
    }

    uint32_t n = 0;        // candidates in raw[]
    uint32_t n_skip = 0;   // skip list (first pick must not already be a projection neighbour, :1859-1862)
    bool have_dists = false, fill = false;
    uint32_t *target = nullptr;
    uint32_t target_cap = 0;

    if (p.mode == 1) {
        // one CTA per TRAINING QUERY: its pivot's pruned list (PruneBiSearchBaseGetBase :1612-1694) depends on the query's
        // kNN list alone, so every query's list is formed here; which one the pivot keeps (the last, :1093) and which pushes
        // survive is replayed per destination in mode 2
        n = min(p.M_sq, kMaxCand);
        for (uint32_t i = tid; i < n; i += kPruneThreads) raw[i] = p.knn[(size_t)blockIdx.x * p.knn_stride + i];
        fill = true;
        target = p.qlist + (size_t)blockIdx.x * p.proj_stride;
        target_cap = M;
    } else if (p.mode == 3) {
        const uint32_t i0 = blockIdx.x;
        n = min(p.exp_cnt[i0], min(p.exp_cap, kMaxCand));
        for (uint32_t i = tid; i < n; i += kPruneThreads) {
            const uint32_t id = p.exp_ids[(size_t)i0 * p.exp_cap + i];
            raw[i] = id;
            keys[i] = ((unsigned long long)f2ord(p.exp_dists[(size_t)i0 * p.exp_cap + i]) << 32) | id;
        }
        have_dists = true;
        const uint32_t *prow_ = p.proj + (size_t)v * p.proj_stride;
        n_skip = min(prow_[0], 128u);
        for (uint32_t i = tid; i < n_skip; i += kPruneThreads) skip[i] = prow_[1 + i];
        target = p.supply + (size_t)v * p.supply_stride;
        target_cap = M;
    } else {
        // mode 5: supply lists above M_pjbp re-pruned (:1224-1247)
        uint32_t *grow = p.supply + (size_t)v * p.supply_stride;
        const uint32_t n_own = min(grow[0], 128u);
        if (n_own <= M) return;
        const uint32_t *prow_ = p.proj + (size_t)v * p.proj_stride;
        n_skip = min(prow_[0], 128u);
        for (uint32_t i = tid; i < n_skip; i += kPruneThreads) skip[i] = prow_[1 + i];
        for (uint32_t i = tid; i < n_own; i += kPruneThreads) {
            own[i] = grow[1 + i];
            raw[i] = own[i];
        }
        n = n_own;
        target = grow;
        target_cap = M;
    }
    __syncthreads();
    // the node's own row
    for (uint32_t i = tid; i < dim4; i += kPruneThreads)
        reinterpret_cast<float4 *>(sm.vrow)[i] = __ldg(reinterpret_cast<const float4 *>(p.base + (size_t)v * p.dim) + i);
    __syncthreads();
    const uint32_t nres = prune_candidates(p, sm, v, n, target_cap, fill, have_dists, n_skip);
    if (tid == 0) target[0] = nres;
    for (uint32_t i = tid; i < nres; i += kPruneThreads) target[1 + i] = result[i];
}

__global__ void owner_kernel(const uint32_t *knn, uint32_t knn_stride, uint64_t nsq, uint32_t nd, int *owner) {
    const uint64_t sq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sq >= nsq) return;
    const uint32_t piv = knn[sq * knn_stride];
    if (piv < nd) atomicMax(owner + piv, (int)sq);  // "projection_graph_[cur_tgt] = pruned_list": the last writer wins
}

// projection_graph_[pivot] = the pruned list of the LAST training query that has this pivot (:1093)
__global__ void owner_copy_kernel(const int *owner, const uint32_t *qlist, uint32_t *proj, uint32_t stride, uint64_t nd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd * stride) return;
    const uint64_t v = i / stride;
    const int o = owner[v];
    if (o >= 0) proj[i] = qlist[(uint64_t)o * stride + (i % stride)];
}

// keys (dst << 32 | src) for every edge of nodes [lo, lo + n); unused slots = ~0
__global__ void emit_edges_kernel(const uint32_t *adj, uint32_t stride, uint32_t lo, uint32_t n, uint32_t slots,
                                  unsigned long long *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n * slots) return;
    const uint32_t u = lo + (uint32_t)(i / slots), j = (uint32_t)(i % slots);
    const uint32_t *row = adj + (size_t)u * stride;
    keys[i] = (j < row[0]) ? (((unsigned long long)row[1 + j] << 32) | u) : ~0ull;
}

__global__ void seg_offsets_kernel(const unsigned long long *keys, uint64_t n, uint32_t nd, uint64_t *off) {
    const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d > nd) return;
    const unsigned long long want = (unsigned long long)d << 32;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    off[d] = lo;
}

// CalculateProjectionep (:2004-2041): centroid, then the row closest to it in squared L2 (lowest id on ties).  The column
// sums are formed in double precision in a FIXED order (per-block partial sums over a contiguous row range, then the
// partials summed in block order), so the centroid — and with it the entry point — does not depend on scheduling; the
// reference adds the rows one after the other in fp32, which differs from the exact sum by more than this does.
__global__ void colsum_partial_kernel(const float *base, uint64_t nd, uint32_t dim, double *partial) {
    const uint64_t rows_per = (nd + gridDim.x - 1) / gridDim.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per, r1 = min(nd, r0 + rows_per);
    for (uint32_t dd = threadIdx.x; dd < dim; dd += blockDim.x) {
        double s = 0.0;
        for (uint64_t r = r0; r < r1; ++r) s += (double)base[r * dim + dd];
        partial[(uint64_t)blockIdx.x * dim + dd] = s;
    }
}
__global__ void colsum_final_kernel(const double *partial, uint32_t n_partial, uint32_t dim, double *sum) {
    const uint32_t dd = blockIdx.x * blockDim.x + threadIdx.x;
    if (dd >= dim) return;
    double s = 0.0;
    for (uint32_t b = 0; b < n_partial; ++b) s += partial[(uint64_t)b * dim + dd];
    sum[dd] = s;
}
__global__ void closest_to_center_kernel(const float *base, uint64_t nd, uint32_t dim, const double *sum,
                                         unsigned long long *best) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nd) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (uint32_t j = lane; j < dim; j += 32) {
        const float c = (float)(sum[j] / (double)nd);
        const float df = c - base[row * dim + j];
        s = fmaf(df, df, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    if (lane == 0) atomicMin(best, ((unsigned long long)f2ord(s) << 32) | (uint32_t)row);
}

__global__ void copy_rows_kernel(const uint32_t *src, uint32_t sstride, uint32_t *dst, uint32_t dstride, uint64_t nd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd * sstride) return;
    const uint64_t u = i / sstride;
    const uint32_t j = (uint32_t)(i % sstride);
    dst[u * dstride + j] = src[i];
}

// final row = projection ++ (supply \ projection), at most 2*M supply entries (:1251-1267)
__global__ void merge_rows_kernel(const uint32_t *proj, uint32_t pstride, const uint32_t *supply, uint32_t sstride,
                                  uint32_t *out, uint32_t ostride, uint64_t nd, uint32_t M, uint32_t *max_deg) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nd) return;
    const uint32_t *pr = proj + u * pstride, *sr = supply + u * sstride;
    uint32_t *o = out + u * ostride;
    uint32_t n = 0;
    const uint32_t np = pr[0];
    for (uint32_t i = 0; i < np; ++i) o[1 + n++] = pr[1 + i];
    uint32_t added = 0;
    for (uint32_t j = 0; j < sr[0] && added < 2 * M && n + 1 < ostride; ++j) {
        const uint32_t id = sr[1 + j];
        bool dup = false;
        for (uint32_t i = 0; i < n; ++i)
            if (o[1 + i] == id) { dup = true; break; }
        if (!dup) {
            o[1 + n++] = id;
            ++added;
        }
    }
    o[0] = n;
    for (uint32_t i = n + 1; i < ostride; ++i) o[i] = 0xffffffffu;
    atomicMax(max_deg, n);
}

struct Buf {
    void *p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { RG_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); return RG_OK; }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

size_t prune_smem(uint32_t dim, uint32_t M) {
    return kMaxCand * 8 + kMaxCand * 4 + 3 * 128 * 4 + (size_t)(2 + M) * dim * 4;
}

}  // namespace

// Sorted reverse-edge offers for the edges of nodes [lo, lo + n) of `adj`: keys (dst << 32 | src) ascending + per-dst offsets.
static int gather_offers(const uint32_t *adj, uint32_t stride, uint32_t lo, uint32_t n, uint32_t slots, uint32_t nd,
                         unsigned long long *keys_a, unsigned long long *keys_b, void *cub_tmp, size_t cub_bytes,
                         uint64_t *in_off, const unsigned long long **sorted, cudaStream_t s) {
    const uint64_t total = (uint64_t)n * slots;
    emit_edges_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(adj, stride, lo, n, slots, keys_a);
    int end_bit = 33;
    while (end_bit < 64 && (1ull << (end_bit - 32)) <= nd) ++end_bit;
    (void)end_bit;  // padding keys are all-ones: sort on the full 64 bits so they stay at the end
    RG_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, keys_a, keys_b, total, 0, 64, s));
    seg_offsets_kernel<<<(nd + 1 + 255) / 256, 256, 0, s>>>(keys_b, total, nd, in_off);
    *sorted = keys_b;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

__global__ void degree_sum_kernel(const uint32_t *adj, uint32_t stride, uint64_t nd, unsigned long long *total) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long d = u < nd ? adj[u * stride] : 0ull;
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0 && d) atomicAdd(total, d);
}

int adjacency_nnz_device(const uint32_t *d_adj, uint32_t stride, uint64_t nd, uint64_t *nnz) {
    unsigned long long *d_total = nullptr, h = 0;
    RG_CUDA(cudaMalloc(&d_total, 8));
    cudaMemset(d_total, 0, 8);
    degree_sum_kernel<<<(unsigned)((nd + 255) / 256), 256>>>(d_adj, stride, nd, d_total);
    cudaError_t e = cudaMemcpy(&h, d_total, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_total);
    if (e != cudaSuccess) return cuda_fail(e, "adjacency_nnz");
    *nnz = h;
    return RG_OK;
}

int build_roargraph_device(rg_index *idx, const uint32_t *d_knn, uint64_t nsq, uint32_t knn_stride, uint32_t M_sq,
                           uint32_t M, uint32_t L, uint32_t n_rounds, cudaStream_t s) {
    const uint32_t nd = (uint32_t)idx->nd;
    const uint32_t dim = idx->dim;
    const int l2 = idx->metric == RG_METRIC_L2 ? 1 : 0;
    if (M == 0 || M > 64) { set_error("rg_index_build_roargraph: M_pjbp must be in [1, 64]"); return RG_EINVAL; }
    if (M_sq == 0 || M_sq > knn_stride) { set_error("rg_index_build_roargraph: M_sq exceeds the kNN list length"); return RG_EINVAL; }
    if (L == 0 || L >= (1u << 20)) { set_error("rg_index_build_roargraph: bad L_pjpq"); return RG_EINVAL; }
    const uint32_t pstride = (M + 1 + 7) & ~7u, sstride = (2 * M + 1 + 7) & ~7u, ostride = sstride;
    const size_t smem = prune_smem(dim, M);
    if (smem > 227u * 1024u) { set_error("rg_index_build_roargraph: dim * M_pjbp too large for the prune kernel"); return RG_EINVAL; }
    RG_CUDA(cudaFuncSetAttribute(prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // rounds of the connectivity phase
    uint32_t batch;
    if (n_rounds == 0) {
        batch = std::max<uint32_t>(256u, std::min<uint32_t>(1u << 19, (nd + 47) / 48));
    } else {
        batch = (nd + n_rounds - 1) / n_rounds;
    }
    batch = std::min(batch, nd);
    // the connectivity search hands ALL its expanded nodes to the prune (full_retset, :1316); hops exceed L_pjpq by a few, and
    // the prune pool holds kMaxCand candidates: refuse a queue length the pool would silently truncate
    if (L + 64 > kMaxCand) { set_error("rg_index_build_roargraph: L_pjpq above 960 is not supported (prune pool of 1024 candidates)"); return RG_EINVAL; }
    const uint32_t exp_cap = L + 64;

    Buf b_proj, b_supply, b_owner, b_keys_a, b_keys_b, b_cub, b_off, b_sum, b_best, b_exp_ids, b_exp_d, b_exp_cnt, b_out, b_maxdeg;
    int rc;
    const uint64_t offer_slots = std::max<uint64_t>(std::max<uint64_t>((uint64_t)nd * M, (uint64_t)nsq * M), (uint64_t)batch * 2 * M);
    size_t cub_bytes = 0;
    RG_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                           offer_slots, 0, 64, s));
    if ((rc = b_proj.alloc((size_t)nd * pstride * 4)) || (rc = b_supply.alloc((size_t)nd * sstride * 4)) ||
        (rc = b_owner.alloc((size_t)nd * 4)) || (rc = b_keys_a.alloc(offer_slots * 8)) ||
        (rc = b_keys_b.alloc(offer_slots * 8)) || (rc = b_cub.alloc(cub_bytes)) || (rc = b_off.alloc(((size_t)nd + 1) * 8)) ||
        (rc = b_sum.alloc(dim * 8)) || (rc = b_best.alloc(8)) || (rc = b_exp_ids.alloc((size_t)batch * exp_cap * 4)) ||
        (rc = b_exp_d.alloc((size_t)batch * exp_cap * 4)) || (rc = b_exp_cnt.alloc((size_t)batch * 4)) ||
        (rc = b_maxdeg.alloc(4)))
        return rc;
    uint32_t *proj = b_proj.as<uint32_t>(), *supply = b_supply.as<uint32_t>();
    RG_CUDA(cudaMemsetAsync(proj, 0, (size_t)nd * pstride * 4, s));
    RG_CUDA(cudaMemsetAsync(supply, 0, (size_t)nd * sstride * 4, s));

    // entry point
    RG_CUDA(cudaMemsetAsync(b_best.p, 0xff, 8, s));
    {
        const uint32_t n_partial = std::min<uint32_t>(nd, 4096);
        Buf b_partial;
        if ((rc = b_partial.alloc((size_t)n_partial * dim * 8))) return rc;
        colsum_partial_kernel<<<n_partial, 256, 0, s>>>(idx->d_base, nd, dim, b_partial.as<double>());
        colsum_final_kernel<<<(dim + 255) / 256, 256, 0, s>>>(b_partial.as<double>(), n_partial, dim, b_sum.as<double>());
        RG_CUDA(cudaStreamSynchronize(s));  // b_partial goes out of scope
    }
    closest_to_center_kernel<<<(nd + 7) / 8, 256, 0, s>>>(idx->d_base, nd, dim, b_sum.as<double>(), b_best.as<unsigned long long>());
    unsigned long long h_best = 0;
    RG_CUDA(cudaMemcpyAsync(&h_best, b_best.p, 8, cudaMemcpyDeviceToHost, s));
    RG_CUDA(cudaStreamSynchronize(s));
    const uint32_t ep = (uint32_t)h_best;

    PruneParams pp{};
    pp.base = idx->d_base;
    pp.dim = dim;
    pp.nd = nd;
    pp.l2 = l2;
    pp.M = M;
    pp.knn = d_knn;
    pp.knn_stride = knn_stride;
    pp.M_sq = M_sq;
    pp.owner = b_owner.as<int>();
    pp.proj = proj;
    pp.supply = supply;
    pp.proj_stride = pstride;
    pp.supply_stride = sstride;
    pp.in_off = b_off.as<uint64_t>();

    static const bool trace = getenv("RG_BUILD_TRACE") != nullptr;  // per-phase degree totals (synchronises), to stderr
    auto t_last = std::chrono::steady_clock::now();
    auto t_lap = t_last;
    auto report = [&](const char *what, const uint32_t *adj, uint32_t stride) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const auto t_now = std::chrono::steady_clock::now();
        const double sec = std::chrono::duration<double>(t_now - t_last).count();
        uint64_t nnz = 0;
        if (adjacency_nnz_device(adj, stride, nd, &nnz) == RG_OK)
            fprintf(stderr, "[rg build] %-46s %7.2f s, total degree %llu, avg %.4f\n", what, sec, (unsigned long long)nnz,
                    (double)nnz / nd);
        t_last = std::chrono::steady_clock::now();
    };
    double t_search = 0, t_prune = 0, t_reverse = 0;  // connectivity phase split (trace only)
    unsigned long long search_cmps = 0, search_adj = 0;
    auto lap = [&](double &acc) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const auto t_now = std::chrono::steady_clock::now();
        acc += std::chrono::duration<double>(t_now - t_lap).count();
        t_lap = t_now;
    };
    // phase 1a: every training query's pruned list; the pivot keeps the list of its last query
    RG_CUDA(cudaMemsetAsync(b_owner.p, 0xff, (size_t)nd * 4, s));
    owner_kernel<<<(unsigned)((nsq + 255) / 256), 256, 0, s>>>(d_knn, knn_stride, nsq, nd, b_owner.as<int>());
    Buf b_qlist;
    if ((rc = b_qlist.alloc((size_t)nsq * pstride * 4))) return rc;
    pp.qlist = b_qlist.as<uint32_t>();
    pp.mode = 1;
    pp.node_lo = 0;
    pp.n_nodes = (uint32_t)nsq;
    prune_kernel<<<(unsigned)nsq, kPruneThreads, smem, s>>>(pp);
    RG_CUDA(cudaGetLastError());
    owner_copy_kernel<<<(unsigned)(((uint64_t)nd * pstride + 255) / 256), 256, 0, s>>>(b_owner.as<int>(), pp.qlist, proj, pstride, nd);
    report("projection, pivots pruned (1a)", proj, pstride);
    // phase 1b: reverse edges.  Round 1 replays the pushes the reference makes right after every training query's prune
    // (ProjectionAddReverse inside the loop, :1095): offers = (member of query sq's list) <- pivot(sq), in query order, those
    // older than the destination's own last replacement dropped.  Round 2 = the pass over all nodes (:1105-1108).
    for (int round = 0; round < 2; ++round) {
        const unsigned long long *sorted = nullptr;
        if (round == 0)
            rc = gather_offers(pp.qlist, pstride, 0, (uint32_t)nsq, M, nd, b_keys_a.as<unsigned long long>(),
                               b_keys_b.as<unsigned long long>(), b_cub.p, cub_bytes, b_off.as<uint64_t>(), &sorted, s);
        else
            rc = gather_offers(proj, pstride, 0, nd, M, nd, b_keys_a.as<unsigned long long>(), b_keys_b.as<unsigned long long>(),
                               b_cub.p, cub_bytes, b_off.as<uint64_t>(), &sorted, s);
        if (rc != RG_OK) return rc;
        pp.mode = 2;
        pp.offers_by_query = round == 0 ? 1 : 0;
        pp.node_lo = 0;
        pp.n_nodes = nd;
        pp.in_keys = sorted;
        prune_kernel<<<nd, kPruneThreads, smem, s>>>(pp);
        RG_CUDA(cudaGetLastError());
        report(round == 0 ? "projection + reverse edges, round 1 (1b)" : "projection + reverse edges, round 2 (1b)", proj, pstride);
    }
    pp.offers_by_query = 0;
    RG_CUDA(cudaStreamSynchronize(s));
    cudaFree(b_qlist.p);
    b_qlist.p = nullptr;
    // supply graph starts as the projection graph (:1179-1184)
    copy_rows_kernel<<<(unsigned)(((uint64_t)nd * pstride + 255) / 256), 256, 0, s>>>(proj, pstride, supply, sstride, nd);

    // phase 3
    SearchParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.base = idx->d_base;
    sp.nd = nd;
    sp.dim = dim;
    sp.k = 1;
    sp.L = L;
    sp.metric_l2 = l2;
    sp.tag_stride = idx->tag_stride;
    sp.counter = idx->d_counter;
    sp.stats = idx->d_stats;
    sp.adj = supply;
    sp.adj_stride = sstride;
    sp.ep = ep;
    sp.n_init = 0;
    sp.exp_ids = b_exp_ids.as<uint32_t>();
    sp.exp_dists = b_exp_d.as<float>();
    sp.exp_cnt = b_exp_cnt.as<uint32_t>();
    sp.exp_cap = exp_cap;
    pp.exp_ids = sp.exp_ids;
    pp.exp_dists = sp.exp_dists;
    pp.exp_cnt = sp.exp_cnt;
    pp.exp_cap = exp_cap;
    for (uint32_t lo = 0; lo < nd; lo += batch) {
        const uint32_t n = std::min(batch, nd - lo);
        sp.nq = n;
        sp.node_lo = lo;
        SearchLaunch l;
        rc = plan_search(sp, l, sm_count(), std::max(idx->n_slots, affordable_slots(idx)));
        if (rc != RG_OK) return rc;
        rc = ensure_tags(idx, l.grid * l.warps_per_cta);
        if (rc != RG_OK) return rc;
        sp.tags = idx->d_tags;
        sp.slot_epoch = idx->d_slot_epoch;
        if (trace) {
            cudaStreamSynchronize(s);
            t_lap = std::chrono::steady_clock::now();
        }
        rc = launch_build_search(sp, l, s);
        if (rc != RG_OK) return rc;
        lap(t_search);
        if (trace) {  // evaluations of this round's searches -> gather bytes of the build's beam search
            unsigned long long st[4] = {0, 0, 0, 0};
            cudaMemcpy(st, idx->d_stats, sizeof(st), cudaMemcpyDeviceToHost);
            search_cmps += st[0];
            search_adj += st[2];
        }
        pp.mode = 3;
        pp.node_lo = lo;
        pp.n_nodes = n;
        prune_kernel<<<n, kPruneThreads, smem, s>>>(pp);
        RG_CUDA(cudaGetLastError());
        lap(t_prune);
        const unsigned long long *sorted = nullptr;
        rc = gather_offers(supply, sstride, lo, n, 2 * M, nd, b_keys_a.as<unsigned long long>(), b_keys_b.as<unsigned long long>(),
                           b_cub.p, cub_bytes, b_off.as<uint64_t>(), &sorted, s);
        if (rc != RG_OK) return rc;
        pp.mode = 4;
        pp.node_lo = 0;
        pp.n_nodes = nd;
        pp.in_keys = sorted;
        prune_kernel<<<nd, kPruneThreads, smem, s>>>(pp);
        RG_CUDA(cudaGetLastError());
        lap(t_reverse);
    }
    if (trace)
        fprintf(stderr, "[rg build] connectivity phase: beam search %.2f s (%.1f evaluations per node, %.0f GB/s of gathered rows + "
                "adjacency), prune of the expanded sets %.2f s, reverse edges %.2f s\n",
                t_search, (double)search_cmps / nd, ((double)(search_cmps + nd) * dim * 4.0 + (double)search_adj * 4.0) / t_search / 1e9,
                t_prune, t_reverse);
    report("supply graph after connectivity enhancement (3)", supply, sstride);
    // phase 4
    pp.mode = 5;
    pp.node_lo = 0;
    pp.n_nodes = nd;
    prune_kernel<<<nd, kPruneThreads, smem, s>>>(pp);
    RG_CUDA(cudaGetLastError());
    report("supply graph after the degree check (4)", supply, sstride);
    if ((rc = b_out.alloc((size_t)nd * ostride * 4))) return rc;
    RG_CUDA(cudaMemsetAsync(b_maxdeg.p, 0, 4, s));
    merge_rows_kernel<<<(nd + 255) / 256, 256, 0, s>>>(proj, pstride, supply, sstride, b_out.as<uint32_t>(), ostride, nd, M,
                                                        b_maxdeg.as<uint32_t>());
    RG_CUDA(cudaGetLastError());
    uint32_t h_max = 0;
    RG_CUDA(cudaMemcpyAsync(&h_max, b_maxdeg.p, 4, cudaMemcpyDeviceToHost, s));
    RG_CUDA(cudaStreamSynchronize(s));
    if (idx->d_adj) cudaFree(idx->d_adj);
    idx->d_adj = b_out.as<uint32_t>();
    b_out.p = nullptr;  // ownership moves to the index
    idx->adj_stride = ostride;
    idx->max_deg = h_max;
    idx->ep = ep;
    idx->has_graph = true;
    return RG_OK;
}

}  // namespace rg
