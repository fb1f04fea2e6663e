// rg_search.cu — batched greedy beam search on sm_100a: one warp per query.
//
// Replaces the per-query loop of IndexBipartite::SearchRoarGraph (src/index_bipartite.cpp:2311-2420) and
// IndexBipartite::SearchBipartiteGraph (:282-356) of the reference (paths relative to /root/reference) for a
// whole query batch.  Results are bit-identical to the reference's CPU path: same ids, same fp32 distances,
// same cmps / hops.
//
// Design:
//   * persistent CTAs, each warp pulls query ids from an atomic ticket (the reference's schedule(dynamic,1));
//   * candidate queue (NeighborPriorityQueue, neighbor.h:138-223) = sorted (dist,id|flag) arrays in shared memory,
//     maintained by warp-cooperative 32-ary search + parallel shift;
//   * visited set (VisitedList, visited_list_pool.h:8-29) = exact per-warp tag array in HBM with 8-bit epochs;
//   * neighbour rows are gathered HBM -> shared memory with one cp.async.bulk (TMA 1-D bulk copy, UBLKCP) per row,
//     completion tracked per group of 8 rows on an mbarrier, so every row of a hop is in flight at once;
//   * distances reproduce the reference's AVX-512 rounding order exactly (distance.h:39-89,179-225 as compiled):
//     4 lanes per row, each owning 4 of the 16 virtual SIMD lanes; unfused mul+add in the 16-wide loop,
//     fma in the 8-wide tail, pairwise folds by shuffles.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "rg_internal.h"

namespace rg {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// One row HBM -> shared memory through the bulk-copy (TMA) engine; completes `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// One row through the bulk-copy engine with an L2 eviction policy (rows are streamed once: evict-first keeps them from
// pushing the visited tables out of L2).
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 32-byte bucket of the visited table, kept in L2 with evict-last priority (256-bit load / store, sm_100)
__device__ __forceinline__ void ld_bucket(const uint32_t *p, uint32_t (&v)[8]) {
    asm volatile("ld.relaxed.gpu.global.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p)
                 : "memory");
}
// The same bucket in the warp's shared memory (RG_VISITED=smem): two 128-bit shared loads / stores.
__device__ __forceinline__ void ld_bucket_smem(const uint32_t *p, uint32_t (&v)[8]) {
    const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <bool kSmem>
__device__ __forceinline__ void ld_bucket_any(const uint32_t *p, uint32_t (&v)[8]) {
    if (kSmem) ld_bucket_smem(p, v);
    else ld_bucket(p, v);
}
__device__ __forceinline__ void st_bucket_empty(uint32_t *p) {
    const uint32_t x = 0xffffffffu;
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(x) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ bool nbr_less(float d1, uint32_t i1, float d2, uint32_t i2) {  // neighbor.h:29-31
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

// Distance::compare as the reference computes it (distance.h:39-89 / :179-225 compiled with -Ofast -march=native;
// DESIGN.md documents the rounding order).  `t` = lane & 3 owns virtual SIMD lanes 4t..4t+3 of one row;
// the result is valid in the lane with t == 0.  dim % 8 == 0, so only the 16-wide loop and the 8-wide block exist.
template <bool kL2>
__device__ __forceinline__ float row_distance(const float *row, const float *q, int n16, bool has8, int t) {
    const float4 *r4 = reinterpret_cast<const float4 *>(row) + t;
    const float4 *q4 = reinterpret_cast<const float4 *>(q) + t;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for (int c = 0; c < n16; ++c) {
        const float4 a = r4[4 * c];
        const float4 b = q4[4 * c];
        if (kL2) {
            const float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y), d2 = __fsub_rn(a.z, b.z),
                        d3 = __fsub_rn(a.w, b.w);
            s0 = __fadd_rn(s0, __fmul_rn(d0, d0));
            s1 = __fadd_rn(s1, __fmul_rn(d1, d1));
            s2 = __fadd_rn(s2, __fmul_rn(d2, d2));
            s3 = __fadd_rn(s3, __fmul_rn(d3, d3));
        } else {
            s0 = __fadd_rn(s0, __fmul_rn(a.x, b.x));
            s1 = __fadd_rn(s1, __fmul_rn(a.y, b.y));
            s2 = __fadd_rn(s2, __fmul_rn(a.z, b.z));
            s3 = __fadd_rn(s3, __fmul_rn(a.w, b.w));
        }
    }
    // t[l] = s[l+8] + s[l]  (virtual lanes l+8 live two threads up)
    float t0 = __fadd_rn(__shfl_down_sync(kFull, s0, 2, 4), s0);
    float t1 = __fadd_rn(__shfl_down_sync(kFull, s1, 2, 4), s1);
    float t2 = __fadd_rn(__shfl_down_sync(kFull, s2, 2, 4), s2);
    float t3 = __fadd_rn(__shfl_down_sync(kFull, s3, 2, 4), s3);
    if (has8 && t < 2) {
        const float4 a = r4[4 * n16];
        const float4 b = q4[4 * n16];
        if (kL2) {
            const float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y), d2 = __fsub_rn(a.z, b.z),
                        d3 = __fsub_rn(a.w, b.w);
            t0 = __fmaf_rn(d0, d0, t0);
            t1 = __fmaf_rn(d1, d1, t1);
            t2 = __fmaf_rn(d2, d2, t2);
            t3 = __fmaf_rn(d3, d3, t3);
        } else {
            t0 = __fmaf_rn(a.x, b.x, t0);
            t1 = __fmaf_rn(a.y, b.y, t1);
            t2 = __fmaf_rn(a.z, b.z, t2);
            t3 = __fmaf_rn(a.w, b.w, t3);
        }
    }
    // u[l] = t[l+4] + t[l]
    const float u0 = __fadd_rn(__shfl_down_sync(kFull, t0, 1, 4), t0);
    const float u1 = __fadd_rn(__shfl_down_sync(kFull, t1, 1, 4), t1);
    const float u2 = __fadd_rn(__shfl_down_sync(kFull, t2, 1, 4), t2);
    const float u3 = __fadd_rn(__shfl_down_sync(kFull, t3, 1, 4), t3);
    const float res = __fadd_rn(__fadd_rn(u0, u1), __fadd_rn(u2, u3));  // two vhaddps
    return kL2 ? res : -res;
}

// Per-warp state of one query.
struct WarpCtx {
    float *q;          // dim
    float *rows;       // R x rstride
    float *qd;         // L   queue distances, ascending
    uint32_t *qi;      // L   queue ids | kExpandedBit
    uint32_t *cand;    // kCandCap
    uint32_t bar0;     // shared address of mbarrier[0] (4 barriers, one per group of 8 rows)
    uint32_t rows_u32; // shared address of rows
    uint32_t size, cur, phase;
    uint64_t row_policy;  // L2 eviction policy of the row copies (0 = none)
    int lane;
    // lazy queue (long queues): accepted candidates wait here, unsorted, and enter the sorted arrays kPendMax at a time
    float *pd;           // kPendSlots distances
    uint32_t *pi;        // kPendSlots ids | kExpandedBit
    uint32_t npend, npu; // entries waiting / of them not yet expanded
};

constexpr uint32_t kPendSlots = 32;
constexpr uint32_t kPendMax = 31;  // lane 31 of the sorted pending block stays a sentinel (32-entry search by shuffles)

// NeighborPriorityQueue::insert (neighbor.h:150-183), warp-cooperative.  All lanes pass the same (cd, cid).
// Returns true when the candidate entered the queue.
__device__ __forceinline__ bool queue_insert(WarpCtx &w, uint32_t L, float cd, uint32_t cid) {
    const int lane = w.lane;
    float *qd = w.qd;
    uint32_t *qi = w.qi;
    if (w.size == L) {  // :151 full and last < nbr  (an exact duplicate of `last` is caught below)
        const float ld = qd[L - 1];
        const uint32_t li = qi[L - 1] & kIdMask;
        if (!nbr_less(cd, cid, ld, li)) return false;
    }
    // lo = number of entries e with !(c < e), found by 32-ary search (entries are sorted)
    uint32_t lo = 0, len = w.size;
    while (len > 32) {
        const uint32_t step = (len + 31) >> 5;
        const uint32_t end = lo + len;
        const uint32_t idx = min(lo + (lane + 1) * step, end) - 1;
        const float ed = qd[idx];
        const uint32_t ei = qi[idx] & kIdMask;
        const bool le = !nbr_less(cd, cid, ed, ei);
        const uint32_t nle = __popc(__ballot_sync(kFull, le));
        if (nle == 32) {
            lo = end;
            len = 0;
        } else {
            const uint32_t new_lo = (nle == 0) ? lo : min(lo + nle * step, end);
            const uint32_t p_hi = min(lo + (nle + 1) * step, end) - 1;  // first probe known to be > c
            lo = new_lo;
            len = p_hi - new_lo;
        }
    }
    {
        const uint32_t idx = lo + lane;
        bool le = false;
        if ((uint32_t)lane < len) le = !nbr_less(cd, cid, qd[idx], qi[idx] & kIdMask);
        lo += __popc(__ballot_sync(kFull, le));
    }
    // :161 same id met on the search path == exact duplicate sits right below the insertion point
    if (lo > 0 && (qi[lo - 1] & kIdMask) == cid && qd[lo - 1] == cd) return false;
    const uint32_t newsize = min(w.size + 1, L);
    // :169 memmove of the tail one slot up (the last entry falls off when full)
    for (int hi = (int)newsize - 1; hi > (int)lo; hi -= 32) {
        const int j = hi - lane;
        const bool act = j > (int)lo;
        float td = 0.f;
        uint32_t ti = 0;
        if (act) {
            td = qd[j - 1];
            ti = qi[j - 1];
        }
        __syncwarp();
        if (act) {
            qd[j] = td;
            qi[j] = ti;
        }
        __syncwarp();
    }
    if (lane == 0) {
        qd[lo] = cd;
        qi[lo] = cid;  // flag = false (:174)
    }
    __syncwarp();
    w.size = newsize;
    if (lo < w.cur) w.cur = lo;  // :180
    return true;
}

// Up to 8 candidates of one row group (held by the lanes with (lane & 3) == 0, `pass` set) merged into the queue in ONE pass
// over the tail instead of one shift per candidate.  Sequential NeighborPriorityQueue::insert calls leave the queue as "the L
// smallest (distance, id) of everything seen, first unexpanded entry under the cursor" whatever their order,
// so the merge is observably identical: every candidate finds its upper bound among the current entries by its own binary
// search (all in parallel), exact duplicates are dropped as at neighbor.h:161, the accepted candidates are ranked among
// themselves, entry j moves up by the number of candidates that land at or before it, and whatever ends beyond L falls
// off.  Long queues pay O(L / 32) once per group instead of once per candidate.
// `unique_ids`: the group's ids are pairwise distinct (true for every visited-filtered adjacency list; false for the bipartite
// search's random entry points, which may repeat and then go through the one-by-one insert whose duplicate check drops them).
template <bool kPrefetch>
__device__ __forceinline__ void queue_insert_group(WarpCtx &w, const SearchParams &p, float d, uint32_t id, bool pass,
                                                   bool unique_ids) {
    const int lane = w.lane;
    const uint32_t L = p.L;
    uint32_t m = __ballot_sync(kFull, pass);
    if (!m) return;
    if (!p.merge_insert || !unique_ids || (m & (m - 1)) == 0) {  // one candidate, possible repeats, or merge off: plain insert
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(kFull, d, src);
            const uint32_t cid = __shfl_sync(kFull, id, src);
            const bool in = queue_insert(w, L, cd, cid);
            if (kPrefetch && in) {
                // the candidate sits in the top-L now and will very likely be expanded (hops ~= L):
                // pull its adjacency row towards L2 so the dependent read at expansion time is short.
                const char *a = reinterpret_cast<const char *>(p.adj + (size_t)cid * p.adj_stride);
                const uint32_t bytes = p.adj_stride * 4u;
                if ((uint32_t)lane * 32u < bytes) prefetch_l2(a + lane * 32u);
            }
        }
        return;
    }
    float *qd = w.qd;
    uint32_t *qi = w.qi;
    const uint32_t size = w.size;
    // upper bound of this lane's candidate among the entries (binary search per candidate lane, in parallel)
    uint32_t pos = 0xffffffffu;
    if (pass) {
        uint32_t lo = 0, hi = size;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (nbr_less(d, id, qd[mid], qi[mid] & kIdMask)) hi = mid;
            else lo = mid + 1;
        }
        const bool dup = lo > 0 && (qi[lo - 1] & kIdMask) == id && qd[lo - 1] == d;
        if (!dup) pos = lo;
    }
    const uint32_t acc = __ballot_sync(kFull, pos != 0xffffffffu);
    if (!acc) return;
    // positions of the (up to 8) accepted candidates in every lane; rank of this lane's candidate among them
    uint32_t posv[8];
    uint32_t rank = 0, minpos = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        posv[c] = __shfl_sync(kFull, pos, 4 * c);
        const float od = __shfl_sync(kFull, d, 4 * c);
        const uint32_t oi = __shfl_sync(kFull, id, 4 * c);
        if (posv[c] != 0xffffffffu && nbr_less(od, oi, d, id)) ++rank;  // ids within a group are distinct
        minpos = min(minpos, posv[c]);
    }
    const uint32_t fin = pos + rank;  // meaningful where pos is
    const uint32_t newsize = min(size + (uint32_t)__popc(acc), L);
    // entries [minpos, size) move up by the number of candidates landing at or before them, top chunk first: a chunk's
    // targets lie at or above its own slots, all of which have been read (earlier chunks) or are being read (this one)
    for (int top = (int)size - 1; top >= (int)minpos; top -= 32) {
        const int j = top - lane;
        const bool act = j >= (int)minpos;
        float td = 0.f;
        uint32_t ti = 0, nj = 0;
        if (act) {
            td = qd[j];
            ti = qi[j];
            uint32_t sh = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) sh += (posv[c] <= (uint32_t)j) ? 1u : 0u;
            nj = (uint32_t)j + sh;
        }
        __syncwarp();
        if (act && nj < L) {
            qd[nj] = td;
            qi[nj] = ti;
        }
        __syncwarp();
    }
    const bool placed = pos != 0xffffffffu && fin < L;
    if (placed) {
        qd[fin] = d;
        qi[fin] = id;  // flag = false (:174)
    }
    __syncwarp();
    w.size = newsize;
    if (minpos < w.cur) w.cur = minpos;  // the smallest accepted candidate has rank 0: it lands at minpos (< L, it beat the worst)
    if (kPrefetch) {
        uint32_t pm = __ballot_sync(kFull, placed);
        while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1;
            const uint32_t cid = __shfl_sync(kFull, id, src);
            const char *a = reinterpret_cast<const char *>(p.adj + (size_t)cid * p.adj_stride);
            const uint32_t bytes = p.adj_stride * 4u;
            if ((uint32_t)lane * 32u < bytes) prefetch_l2(a + lane * 32u);
        }
    }
}

// NeighborPriorityQueue::closest_unexpanded (neighbor.h:185-192): flag entry `cur`, advance to the next unflagged one.
__device__ __forceinline__ uint32_t queue_pop(WarpCtx &w, float *dist_out = nullptr) {
    const uint32_t u = w.qi[w.cur] & kIdMask;
    if (dist_out) *dist_out = w.qd[w.cur];
    __syncwarp();
    if (w.lane == 0) w.qi[w.cur] = u | kExpandedBit;
    __syncwarp();
    uint32_t c = w.cur + 1;
    for (;;) {
        const uint32_t j = c + w.lane;
        const bool stop = (j >= w.size) || !(w.qi[j] & kExpandedBit);
        const uint32_t m = __ballot_sync(kFull, stop);
        if (m) {
            c += __ffs(m) - 1;
            break;
        }
        c += 32;
    }
    w.cur = min(c, w.size);
    return u;
}

// ---- lazy queue (L_pq >= RG_LAZY_QUEUE_MIN_L) ------------------------------------------------------------------------------
// With a long queue the O(L / 32) shift of every insert pass is most of a hop.  The lazy form keeps NeighborPriorityQueue's
// sorted arrays ("main") plus up to kPendMax accepted candidates in an unsorted pending block, and pays the shift once per
// ~30 accepted candidates.  It stays observably identical to the reference's eager queue Q* (the L best (distance, id) of
// everything seen, neighbor.h:150-192):
//   * a candidate is accepted iff it beats main's last entry when main is full — a superset of what Q* accepts, because
//     Q* is the L best of main + pending and its worst is therefore no worse than main's; what Q* would have dropped falls
//     off the end at the next flush;
//   * closest_unexpanded = the smaller of main[cur] and the best unexpanded pending entry.  It is the best unexpanded entry
//     seen, so it is the one the reference expands next — provided it lies inside Q*, i.e. fewer than L entries are better:
//     its rank = (entries of main before it) + (pending entries before it, all already expanded); at rank >= L the reference
//     has no unexpanded entry left inside its queue and stops, and so does this;
//   * an expanded pending entry keeps its flag through the flush, entries never re-enter, and the only id that
//     can be evaluated twice — the entry point, which SearchRoarGraph does not mark visited (9.1) — is always rejected by the
//     reference (exact duplicate, or worse than a worst entry that only improves) and is dropped here.

// number of main entries e with !(c < e): the upper bound NeighborPriorityQueue::insert computes (neighbor.h:155-167)
__device__ __forceinline__ uint32_t queue_upper_bound(const WarpCtx &w, float cd, uint32_t cid) {
    const int lane = w.lane;
    const float *qd = w.qd;
    const uint32_t *qi = w.qi;
    uint32_t lo = 0, len = w.size;
    while (len > 32) {
        const uint32_t step = (len + 31) >> 5;
        const uint32_t end = lo + len;
        const uint32_t idx = min(lo + (lane + 1) * step, end) - 1;
        const bool le = !nbr_less(cd, cid, qd[idx], qi[idx] & kIdMask);
        const uint32_t nle = __popc(__ballot_sync(kFull, le));
        if (nle == 32) {
            lo = end;
            len = 0;
        } else {
            const uint32_t new_lo = (nle == 0) ? lo : min(lo + nle * step, end);
            const uint32_t p_hi = min(lo + (nle + 1) * step, end) - 1;
            lo = new_lo;
            len = p_hi - new_lo;
        }
    }
    bool le = false;
    if ((uint32_t)lane < len) le = !nbr_less(cd, cid, qd[lo + lane], qi[lo + lane] & kIdMask);
    return lo + __popc(__ballot_sync(kFull, le));
}

// pending block -> main: sort the block across the lanes, move every main entry up by the number of pending entries that
// precede it (top chunk first, in place), drop the pending entries in, cut at L, recompute the cursor
__device__ __forceinline__ void queue_flush(WarpCtx &w, uint32_t L) {
    const uint32_t np = w.npend;
    if (!np) return;
    const int lane = w.lane;
    float *qd = w.qd;
    uint32_t *qi = w.qi;
    float d = (uint32_t)lane < np ? w.pd[lane] : INFINITY;
    uint32_t iw = (uint32_t)lane < np ? w.pi[lane] : kIdMask;  // sentinel: (+inf, largest id) sorts behind everything
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float od = __shfl_xor_sync(kFull, d, j);
            const uint32_t oi = __shfl_xor_sync(kFull, iw, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            const bool other_less = nbr_less(od, oi & kIdMask, d, iw & kIdMask);
            const bool mine_less = nbr_less(d, iw & kIdMask, od, oi & kIdMask);
            const bool take = (lower == up) ? other_less : mine_less;  // keep the smaller / the larger of the pair
            if (take) {
                d = od;
                iw = oi;
            }
        }
    }
    // lane r < np now holds the r-th smallest pending entry; lanes >= np hold sentinels
    const uint32_t size = w.size;
    // pending entries preceding a key (the block is sorted: a 5-step search over the lanes)
    auto pend_before = [&](float ed, uint32_t ei) -> uint32_t {
        uint32_t c = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const float pdv = __shfl_sync(kFull, d, c + step - 1);
            const uint32_t piv = __shfl_sync(kFull, iw, c + step - 1);
            if (nbr_less(pdv, piv & kIdMask, ed, ei)) c += step;
        }
        return c;  // <= 31: lane 31 is a sentinel (np <= kPendMax)
    };
    // where the old cursor entry goes
    uint32_t cur_new = 0xffffffffu;
    if (w.cur < size) cur_new = w.cur + pend_before(qd[w.cur], qi[w.cur] & kIdMask);
    else (void)pend_before(0.f, 0u);  // keep the shuffles convergent
    // upper bound of every pending entry in main: one cooperative 32-ary search per entry would serialise, so each lane
    // runs its own binary search (np <= 31 searches in parallel)
    uint32_t ub = 0;
    {
        uint32_t lo = 0, hi = (uint32_t)lane < np ? size : 0u;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (nbr_less(d, iw & kIdMask, qd[mid], qi[mid] & kIdMask)) hi = mid;
            else lo = mid + 1;
        }
        ub = lo;
    }
    const uint32_t minpos = __shfl_sync(kFull, ub, 0);
    for (int top = (int)size - 1; top >= (int)minpos; top -= 32) {
        const int j = top - lane;
        const bool act = j >= (int)minpos;
        const float ed = act ? qd[j] : 0.f;
        const uint32_t ew = act ? qi[j] : 0u;
        const uint32_t nj = (uint32_t)j + pend_before(ed, ew & kIdMask);
        __syncwarp();
        if (act && nj < L) {
            qd[nj] = ed;
            qi[nj] = ew;
        }
        __syncwarp();
    }
    const uint32_t fin = ub + (uint32_t)lane;
    const bool placed = (uint32_t)lane < np && fin < L;
    if (placed) {
        qd[fin] = d;
        qi[fin] = iw;  // an entry expanded while it was pending keeps its flag
    }
    __syncwarp();
    const uint32_t newsize = min(size + np, L);
    // first unexpanded entry: the old cursor entry at its new place, or the first unexpanded newcomer
    uint32_t first = placed && !(iw & kExpandedBit) ? fin : 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(kFull, first, o));
    w.cur = min(min(cur_new, first), newsize);
    w.size = newsize;
    w.npend = 0;
    w.npu = 0;
}

// up to 8 accepted candidates of a row group (lanes with `pass`) join the pending block
__device__ __forceinline__ void pending_append(WarpCtx &w, const SearchParams &p, float d, uint32_t id, bool pass, bool drop_ep) {
    if (drop_ep && id == p.ep) pass = false;  // the entry point evaluated again: the reference's queue always rejects it
    uint32_t m = __ballot_sync(kFull, pass);
    if (!m) return;
    if (w.npend + (uint32_t)__popc(m) > kPendMax) {
        queue_flush(w, p.L);
        if (pass && w.size == p.L) pass = nbr_less(d, id, w.qd[p.L - 1], w.qi[p.L - 1] & kIdMask);
        m = __ballot_sync(kFull, pass);
        if (!m) return;
    }
    if (pass) {
        const uint32_t slot = w.npend + __popc(m & ((1u << w.lane) - 1u));
        w.pd[slot] = d;
        w.pi[slot] = id;
    }
    __syncwarp();
    w.npend += __popc(m);
    w.npu += __popc(m);
}

// closest_unexpanded over main + pending; 0xffffffff = no unexpanded entry inside the reference's top L
__device__ __forceinline__ uint32_t queue_pop_lazy(WarpCtx &w, uint32_t L, float *dist_out) {
    const int lane = w.lane;
    const bool hasA = w.cur < w.size;
    const float ad = hasA ? w.qd[w.cur] : INFINITY;
    const uint32_t ai = hasA ? (w.qi[w.cur] & kIdMask) : kIdMask;
    // best unexpanded pending entry
    const bool mine = (uint32_t)lane < w.npend && !(w.pi[lane] & kExpandedBit);
    const float md = mine ? w.pd[lane] : INFINITY;
    const uint32_t mi = mine ? w.pi[lane] : kIdMask;
    float bd = md;
    uint32_t bi = mi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(kFull, bd, o);
        const uint32_t oi = __shfl_xor_sync(kFull, bi, o);
        if (nbr_less(od, oi, bd, bi)) {
            bd = od;
            bi = oi;
        }
    }
    const bool hasB = w.npu != 0;
    if (!hasA && !hasB) return 0xffffffffu;
    const bool useB = hasB && (!hasA || nbr_less(bd, bi, ad, ai));
    const float ud = useB ? bd : ad;
    const uint32_t ui = useB ? bi : ai;
    if (w.size + w.npend > L) {
        // the queue may hold more than L entries: is this one inside the L best?  (everything better is already expanded)
        const bool before = (uint32_t)lane < w.npend && nbr_less(w.pd[lane], w.pi[lane] & kIdMask, ud, ui);
        const uint32_t cntp = __popc(__ballot_sync(kFull, before));
        const uint32_t posm = useB ? queue_upper_bound(w, ud, ui) : w.cur;
        if (posm + cntp >= L) return 0xffffffffu;
    }
    if (dist_out) *dist_out = ud;
    __syncwarp();
    if (useB) {
        if (mine && mi == ui) w.pi[lane] = ui | kExpandedBit;
        __syncwarp();
        --w.npu;
    } else {
        if (lane == 0) w.qi[w.cur] = ui | kExpandedBit;
        __syncwarp();
        uint32_t c = w.cur + 1;
        for (;;) {
            const uint32_t j = c + lane;
            const bool stop = (j >= w.size) || !(w.qi[j] & kExpandedBit);
            const uint32_t mm = __ballot_sync(kFull, stop);
            if (mm) {
                c += __ffs(mm) - 1;
                break;
            }
            c += 32;
        }
        w.cur = min(c, w.size);
    }
    return ui;
}

// Evaluate cand[0..n) against the query and insert them: gather rows with bulk copies, 8 rows per mbarrier group.
template <bool kL2>
__device__ __forceinline__ void evaluate_candidates(WarpCtx &w, const SearchParams &p, uint32_t n, bool prefetch_adj,
                                                    bool unique_ids) {
    const int lane = w.lane;
    const uint32_t R = p.rows_per_batch;
    const uint32_t row_bytes = p.dim * 4u;
    const uint32_t rstride = p.row_stride_smem;
    const int n16 = p.dim >> 4;
    const bool has8 = (p.dim & 8u) != 0;
    const int t = lane & 3;
    for (uint32_t b0 = 0; b0 < n; b0 += R) {
        const uint32_t nb = min(R, n - b0);
        const uint32_t ngroups = (nb + 7) >> 3;
        if ((uint32_t)lane < ngroups) mbar_expect_tx(w.bar0 + 8u * lane, min(8u, nb - 8u * lane) * row_bytes);
        if ((uint32_t)lane < nb) {
            const uint32_t id = w.cand[b0 + lane];
            if (w.row_policy) bulk_g2s_hint(w.rows_u32 + lane * rstride * 4u, p.base + (size_t)id * p.dim, row_bytes, w.bar0 + 8u * (lane >> 3), w.row_policy);
            else bulk_g2s(w.rows_u32 + lane * rstride * 4u, p.base + (size_t)id * p.dim, row_bytes, w.bar0 + 8u * (lane >> 3));
        }
        for (uint32_t g = 0; g < ngroups; ++g) {
            mbar_wait(w.bar0 + 8u * g, (w.phase >> g) & 1u);
            w.phase ^= (1u << g);
            const uint32_t r = 8u * g + (lane >> 2);
            const bool valid = r < nb;
            const float d = row_distance<kL2>(w.rows + (size_t)(valid ? r : 0) * rstride, w.q, n16, has8, t);
            const uint32_t id = valid ? w.cand[b0 + r] : 0u;
            bool pass = valid && t == 0;
            if (pass && w.size == p.L) pass = nbr_less(d, id, w.qd[p.L - 1], w.qi[p.L - 1] & kIdMask);
            if (p.lazy_queue && unique_ids) pending_append(w, p, d, id, pass, p.drop_ep != 0);
            else if (prefetch_adj) queue_insert_group<true>(w, p, d, id, pass, unique_ids);
            else queue_insert_group<false>(w, p, d, id, pass, unique_ids);
        }
        __syncwarp();
    }
}

// Same, with the row ring used as two halves of 8 rows (rows_per_batch == 16): the bulk copies of group g + 2 are issued
// as soon as group g has been read, so eight rows stay in flight while the warp computes distances and inserts — the
// single-buffered loop above has nothing in flight during that time.  Insertion order is unchanged (groups in order).
template <bool kL2>
__device__ __forceinline__ void evaluate_candidates_db(WarpCtx &w, const SearchParams &p, uint32_t n, bool prefetch_adj,
                                                       bool unique_ids) {
    const int lane = w.lane;
    const uint32_t row_bytes = p.dim * 4u;
    const uint32_t rstride = p.row_stride_smem;
    const int n16 = p.dim >> 4;
    const bool has8 = (p.dim & 8u) != 0;
    const int t = lane & 3;
    const uint32_t ngroups = (n + 7) >> 3;
    auto issue = [&](uint32_t g) {
        const uint32_t h = g & 1u;
        const uint32_t cnt = min(8u, n - 8u * g);
        if (lane == 0) mbar_expect_tx(w.bar0 + 8u * h, cnt * row_bytes);
        __syncwarp();
        if ((uint32_t)lane < cnt) {
            const uint32_t id = w.cand[8u * g + lane];
            if (w.row_policy) bulk_g2s_hint(w.rows_u32 + (h * 8u + lane) * rstride * 4u, p.base + (size_t)id * p.dim, row_bytes, w.bar0 + 8u * h, w.row_policy);
            else bulk_g2s(w.rows_u32 + (h * 8u + lane) * rstride * 4u, p.base + (size_t)id * p.dim, row_bytes, w.bar0 + 8u * h);
        }
    };
    issue(0);
    if (ngroups > 1) issue(1);
    for (uint32_t g = 0; g < ngroups; ++g) {
        const uint32_t h = g & 1u;
        mbar_wait(w.bar0 + 8u * h, (w.phase >> h) & 1u);
        w.phase ^= (1u << h);
        const uint32_t r = 8u * g + (lane >> 2);
        const bool valid = r < n;
        const float d = row_distance<kL2>(w.rows + (size_t)(h * 8u + (valid ? (lane >> 2) : 0)) * rstride, w.q, n16, has8, t);
        const uint32_t id = valid ? w.cand[r] : 0u;
        __syncwarp();  // every lane has read its row: the half may be refilled
        if (g + 2 < ngroups) issue(g + 2);
        bool pass = valid && t == 0;
        if (pass && w.size == p.L) pass = nbr_less(d, id, w.qd[p.L - 1], w.qi[p.L - 1] & kIdMask);
        if (p.lazy_queue && unique_ids) pending_append(w, p, d, id, pass, p.drop_ep != 0);
        else if (prefetch_adj) queue_insert_group<true>(w, p, d, id, pass, unique_ids);
        else queue_insert_group<false>(w, p, d, id, pass, unique_ids);
    }
    __syncwarp();
}

template <bool kL2>
__device__ __forceinline__ void evaluate(WarpCtx &w, const SearchParams &p, uint32_t n, bool prefetch_adj,
                                         bool unique_ids = true) {
    if (p.rows_per_batch == 16u) evaluate_candidates_db<kL2>(w, p, n, prefetch_adj, unique_ids);
    else evaluate_candidates<kL2>(w, p, n, prefetch_adj, unique_ids);
}

// Visited filter (visited_list_pool.h / src/index_bipartite.cpp:2378-2386): appends the not-yet-visited ids of
// list[0..cnt) to cand[ncand..] and tags them.  Up to kCandCap ids per call; lists are duplicate-free (deduplicated at upload).
__device__ __forceinline__ uint32_t filter_unvisited_ids(WarpCtx &w, uint8_t *tags, uint8_t epoch, const uint32_t (&id)[3],
                                                         uint32_t cnt, uint32_t ncand, uint32_t exclude) {
    const int lane = w.lane;
    uint8_t tg[3];
    bool valid[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) valid[j] = (uint32_t)(j * 32 + lane) < cnt;
#pragma unroll
    for (int j = 0; j < 3; ++j) tg[j] = valid[j] ? __ldcg(tags + id[j]) : epoch;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool fresh = valid[j] && tg[j] != epoch && id[j] != exclude;
        if (fresh) __stcg(tags + id[j], epoch);
        const uint32_t m = __ballot_sync(kFull, fresh);
        if (fresh) w.cand[ncand + __popc(m & ((1u << lane) - 1u))] = id[j];
        ncand += __popc(m);
    }
    __syncwarp();
    return ncand;
}
// ---- visited table --------------------------------------------------------------------------------------------------------
// Exact visited set of one query in 2^bbits buckets of 16 x 16-bit entries (one 32-byte sector per bucket).  With
// h = (id * mul) mod 2^bits (a bijection, mul odd) the bucket is the top bbits of h and the entry stores the low rbits
// (<= 14) — bucket index + entry identify the id, so membership is exact.  Bit 14 of an entry marks an id whose home
// bucket was full when it arrived and that lives in the next bucket instead; 0xffff = empty.  Lookup: found in the home
// bucket -> visited; home bucket has an empty entry -> not visited (entries are never removed, so nothing was ever displaced
// from a non-full bucket) -> insert; home full -> same test on the next bucket with the displaced pattern; both full ->
// the query migrates to the tag array (vt_migrate) and continues there.
// Returns 0 = already visited, 1 = fresh and recorded, 2 = bucket full.
__device__ __forceinline__ int vt_bucket_insert(uint32_t *bucket, uint32_t (&w)[8], uint32_t pat) {
    const uint32_t pat2 = pat | (pat << 16);
    uint32_t hit = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) hit |= __vcmpeq2(w[i], pat2);
    if (hit) return 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        for (;;) {
            const bool lo_empty = (w[i] & 0xffffu) == 0xffffu, hi_empty = (w[i] >> 16) == 0xffffu;
            if (!lo_empty && !hi_empty) break;
            const uint32_t desired = lo_empty ? ((w[i] & 0xffff0000u) | pat) : ((w[i] & 0x0000ffffu) | (pat << 16));
            const uint32_t old = atomicCAS(bucket + i, w[i], desired);
            if (old == w[i]) return 1;
            w[i] = old;  // another lane of this warp filled a half of this word (a different id: lists are duplicate-free)
        }
    }
    return 2;
}

// Visited filter on the table: same contract as filter_unvisited_ids; *overflow is set (warp-uniform) when some lane could
// not record a fresh id (ovf_ids[j] != 0xffffffff names them).
template <bool kSmem>
__device__ __forceinline__ uint32_t filter_unvisited_vt(WarpCtx &w, const SearchParams &p, uint32_t *vt, const uint32_t (&id)[3],
                                                        uint32_t cnt, uint32_t ncand, uint32_t (&ovf_ids)[3], bool *overflow) {
    const int lane = w.lane;
    const uint32_t nb_mask = (1u << p.vt_bbits) - 1u, rmask = (1u << p.vt_rbits) - 1u;
    bool valid[3];
    uint32_t bkt[3], rem[3];
    uint32_t wb[3][8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        valid[j] = (uint32_t)(j * 32 + lane) < cnt;
        const uint32_t h = (id[j] * p.vt_mul) & p.vt_mask;
        bkt[j] = h >> p.vt_rbits;
        rem[j] = h & rmask;
        ovf_ids[j] = 0xffffffffu;
        if (valid[j]) ld_bucket_any<kSmem>(vt + (size_t)bkt[j] * 8u, wb[j]);
    }
    bool any_ovf = false;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        bool fresh = false;
        if (valid[j]) {
            int r = vt_bucket_insert(vt + (size_t)bkt[j] * 8u, wb[j], rem[j]);
            if (r == 2) {
                const uint32_t nb = (bkt[j] + 1u) & nb_mask;
                uint32_t wn[8];
                ld_bucket_any<kSmem>(vt + (size_t)nb * 8u, wn);
                r = vt_bucket_insert(vt + (size_t)nb * 8u, wn, rem[j] | 0x4000u);
                if (r == 2) {
                    ovf_ids[j] = id[j];  // fresh (in neither bucket) but not recorded
                    any_ovf = true;
                }
            }
            fresh = r != 0;
        }
        const uint32_t m = __ballot_sync(kFull, fresh);
        if (fresh) w.cand[ncand + __popc(m & ((1u << lane) - 1u))] = id[j];
        ncand += __popc(m);
    }
    *overflow = __any_sync(kFull, any_ovf);
    __syncwarp();
    return ncand;
}

// The table is (nearly) full: mark every recorded id in the tag array and carry on with tags for the rest of the query.
template <bool kSmem>
__device__ __forceinline__ void vt_migrate(const SearchParams &p, const uint32_t *vt, uint8_t *tags, uint8_t epoch, int lane) {
    const uint32_t nbuckets = 1u << p.vt_bbits, nb_mask = nbuckets - 1u, rmask = (1u << p.vt_rbits) - 1u;
    for (uint32_t b = lane; b < nbuckets; b += 32) {
        uint32_t wv[8];
        ld_bucket_any<kSmem>(vt + (size_t)b * 8u, wv);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t e = (i & 1) ? (wv[i >> 1] >> 16) : (wv[i >> 1] & 0xffffu);
            if (e == 0xffffu) continue;
            const uint32_t home = (e & 0x4000u) ? ((b - 1u) & nb_mask) : b;
            const uint32_t h = (home << p.vt_rbits) | (e & rmask);
            const uint32_t id = (h * p.vt_mulinv) & p.vt_mask;
            __stcg(tags + id, epoch);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// kMode: 0 = SearchRoarGraph, 1 = SearchBipartiteGraph, 2 = SearchProjectionGraphInternal (the build's
// connectivity-enhancement search, src/index_bipartite.cpp:1279-1350: the query is base row node_lo + ticket, that node
// is never evaluated, the entry point IS marked visited, and the expanded nodes are the output).
// kMinBlocks: resident CTAs per SM the register allocation is capped for (4 -> 104 regs, 6 -> 80, 8 -> 64; no spills).
// kVt: the visited set is the hashed table (SearchRoarGraph only); compiled separately so that the tag-array kernels do not
// carry its registers.
template <bool kL2, int kMode, int kMinBlocks, bool kVt>
__global__ void __launch_bounds__(128, kMinBlocks) beam_search_kernel(const SearchParams p) {
    constexpr bool kBipartite = kMode == 1;
    constexpr bool kBuild = kMode == 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    uint8_t *ws = smem_raw + (size_t)warp * p.warp_smem_bytes;

    WarpCtx w;
    w.lane = lane;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ws);  // 4 x 8 B, padded to 64 B
    w.q = reinterpret_cast<float *>(ws + 64);
    w.rows = w.q + p.dim;
    w.qd = w.rows + (size_t)p.rows_per_batch * p.row_stride_smem;
    w.qi = reinterpret_cast<uint32_t *>(w.qd + p.L);
    w.cand = w.qi + p.L;
    w.pd = reinterpret_cast<float *>(w.cand + p.cand_words);  // kPendSlots floats + kPendSlots ids, only when lazy_queue
    w.pi = reinterpret_cast<uint32_t *>(w.pd + kPendSlots);
    w.bar0 = smem_u32(bars);
    w.rows_u32 = smem_u32(w.rows);
    w.phase = 0;
    if (lane < 4) mbar_init(w.bar0 + 8u * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    const uint32_t slot = blockIdx.x * warps_per_cta + warp;
    uint8_t *tags = p.tags + (size_t)slot * p.tag_stride;
    uint32_t epoch = p.slot_epoch[slot];
    constexpr bool use_vt = kVt && kMode == 0;  // hashed visited table (SearchRoarGraph), else the tag array
    // table of this warp: its slice of the L2-resident array, or (RG_VISITED=smem) the tail of its shared-memory working set
    const bool vt_smem = use_vt && p.vt_smem;
    uint32_t *vt = nullptr;
    if (use_vt) {
        if (vt_smem) vt = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(w.cand + p.cand_words + (p.lazy_queue ? 2u * kPendSlots : 0u)) + 31u) & ~(uintptr_t)31u);
        else vt = p.vtab + (size_t)slot * p.vt_words;
    }
    w.row_policy = (use_vt && !vt_smem) ? policy_evict_first() : 0ull;
    // VisitedList::reset (visited_list_pool.h:20-26): bump the tag, clear the array when it wraps
    auto next_epoch = [&]() {
        if (++epoch > 255u) {
            uint4 z = make_uint4(0, 0, 0, 0);
            for (uint64_t i = (uint64_t)lane * 16u; i < p.tag_stride; i += 512u) *reinterpret_cast<uint4 *>(tags + i) = z;
            epoch = 1;
            __threadfence_block();
        }
    };
    unsigned long long acc_cmps = 0, acc_hops = 0, acc_adj = 0;
    const bool prefetch_adj = !kBipartite && p.n_init == 1;  // n_init doubles as the prefetch switch for RoarGraph
    const bool prefetch_next = !kBipartite && p.n_init == 2;

    for (;;) {
        uint32_t qidx = 0;
        if (lane == 0) {
            qidx = atomicAdd(p.counter, 1u);
            // the tickets hand the batch out grouped by region (order_queries): warps that are resident together walk the same
            // part of the graph and find each other's rows in L2
            if (p.order && qidx < p.nq) qidx = __ldg(p.order + qidx);
        }
        qidx = __shfl_sync(kFull, qidx, 0);
        if (qidx >= p.nq) break;

        bool on_tags = !use_vt;  // a query starts on the table and migrates to the tags only if a bucket pair overflows
        if (use_vt) {
            if (vt_smem) {
                const uint4 e = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                for (uint32_t i = (uint32_t)lane * 4u; i < p.vt_words; i += 128u) *reinterpret_cast<uint4 *>(vt + i) = e;
            } else {
                for (uint32_t i = (uint32_t)lane * 8u; i < p.vt_words; i += 256u) st_bucket_empty(vt + i);
            }
            __syncwarp();
        } else {
            next_epoch();
        }
        // visited filter of up to 96 ids held 3 per lane: table or tags
        auto visit = [&](const uint32_t (&vid)[3], uint32_t vcnt, uint32_t exclude) -> uint32_t {
            if (!use_vt || on_tags) return filter_unvisited_ids(w, tags, (uint8_t)epoch, vid, vcnt, 0, exclude);
            uint32_t ovf_ids[3];
            bool ovf;
            const uint32_t n = vt_smem ? filter_unvisited_vt<true>(w, p, vt, vid, vcnt, 0, ovf_ids, &ovf)
                                       : filter_unvisited_vt<false>(w, p, vt, vid, vcnt, 0, ovf_ids, &ovf);
            if (ovf) {
                next_epoch();
                __syncwarp();
                if (vt_smem) vt_migrate<true>(p, vt, tags, (uint8_t)epoch, lane);
                else vt_migrate<false>(p, vt, tags, (uint8_t)epoch, lane);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (ovf_ids[j] != 0xffffffffu) __stcg(tags + ovf_ids[j], (uint8_t)epoch);
                __syncwarp();
                on_tags = true;
            }
            return n;
        };
        {
            const float4 *src = reinterpret_cast<const float4 *>(
                kBuild ? p.base + (size_t)(p.node_lo + qidx) * p.dim : p.queries + (size_t)qidx * p.dim);
            float4 *dst = reinterpret_cast<float4 *>(w.q);
            for (uint32_t i = lane; i < (p.dim >> 2); i += 32) dst[i] = __ldg(src + i);
        }
        w.size = 0;
        w.cur = 0;
        w.npend = 0;
        w.npu = 0;
        __syncwarp();

        uint32_t cmps = 0, hops = 0, adjw = 0;
        // entry points (:2323-2353 / :287-310)
        if (!kBipartite) {
            if (lane == 0) {
                w.cand[0] = p.ep;  // evaluated + inserted, NOT marked visited (:2349 is commented out)
                if (kBuild) __stcg(tags + p.ep, (uint8_t)epoch);  // :1307 the build's search does mark it
            }
            __syncwarp();
            evaluate<kL2>(w, p, 1, false, /*unique_ids=*/false);  // straight into the (empty) sorted arrays, lazy queue or not
        } else {
            for (uint32_t i = lane; i < p.n_init; i += 32) {
                uint32_t id;
                if (p.init_ids) id = p.init_ids[(size_t)qidx * p.n_init + i];
                else id = (uint32_t)(splitmix64(p.seed + (uint64_t)qidx * p.n_init + i) % p.nd);
                w.cand[i] = id;
                __stcg(tags + id, (uint8_t)epoch);  // :306 entry points are marked visited
            }
            __syncwarp();
            evaluate<kL2>(w, p, p.n_init, false, /*unique_ids=*/false);  // random entry points may repeat
        }

        while (w.cur < w.size || w.npu) {  // has_unexpanded_node
            float ud;
            uint32_t u;
            if (p.lazy_queue) {
                u = queue_pop_lazy(w, p.L, kBuild ? &ud : nullptr);
                if (u == 0xffffffffu) break;  // nothing unexpanded inside the L best
            } else {
                u = queue_pop(w, kBuild ? &ud : nullptr);
            }
            if (kBuild && lane == 0 && hops < p.exp_cap) {  // full_retset.push_back(cur_check_node) :1316
                p.exp_ids[(size_t)qidx * p.exp_cap + hops] = u;
                p.exp_dists[(size_t)qidx * p.exp_cap + hops] = ud;
            }
            ++hops;
            if (prefetch_next && w.cur < w.size) {
                // unless this hop finds something closer, the node now at `cur` is expanded next: pull its adjacency row
                // towards L2 while this hop's rows are gathered (3 x 128 B; wrong guesses cost 1 % of a hop's traffic)
                const char *a = reinterpret_cast<const char *>(p.adj + (size_t)(w.qi[w.cur] & kIdMask) * p.adj_stride);
                if ((uint32_t)lane * 128u < p.adj_stride * 4u) prefetch_l2(a + lane * 128u);
            }
            if (!kBipartite) {
                // the whole fixed-stride adjacency row in ONE round trip: words lane, lane + 32, lane + 64 (word 0 = degree);
                // a load of the degree followed by a dependent load of the ids would cost two DRAM latencies per hop
                const uint32_t *arow = p.adj + (size_t)u * p.adj_stride;
                uint32_t wv[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint32_t i = j * 32 + lane;
                    wv[j] = i < p.adj_stride ? __ldg(arow + i) : 0u;
                }
                const uint32_t deg = __shfl_sync(kFull, wv[0], 0);
                adjw += deg;
                const uint32_t exclude = kBuild ? p.node_lo + qidx : 0xffffffffu;
                if (deg < (uint32_t)kCandCap) {
                    uint32_t nid[3];  // neighbour i = word i + 1: one lane up, lane 31 takes word 0 of the next chunk
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t up = __shfl_sync(kFull, wv[j], (lane + 1) & 31);
                        const uint32_t nx = __shfl_sync(kFull, wv[j < 2 ? j + 1 : 2], 0);
                        nid[j] = lane == 31 ? nx : up;
                    }
                    const uint32_t n = visit(nid, deg, exclude);
                    cmps += n;
                    if (n) evaluate<kL2>(w, p, n, prefetch_adj);
                } else {
                    for (uint32_t off = 0; off < deg; off += kCandCap) {
                        const uint32_t c = min(deg - off, (uint32_t)kCandCap);
                        uint32_t nid[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) nid[j] = (uint32_t)(j * 32 + lane) < c ? __ldg(arow + 1 + off + j * 32 + lane) : 0u;
                        const uint32_t n = visit(nid, c, exclude);
                        cmps += n;
                        if (n) evaluate<kL2>(w, p, n, prefetch_adj);
                    }
                }
            } else {
                // Two-hop expansion through the query nodes (:324-341): row_off / csr_nbrs hold, per BASE node, the distinct base
                // ids reachable through its query nodes (rg_index_set_bipartite flattens base -> query -> base at upload; an id
                // that sits in several second-hop lists of the node is evaluated once by the reference too, and the order
                // inside a hop is unobservable).  The list streams through the visited filter 96 ids at a time
                // like a RoarGraph adjacency row; the ids of the next pass are loaded before this pass is filtered, and fresh
                // ids gather across passes so that whole 8-row groups go through the row ring.  hops++ precedes the
                // empty-adjacency skip (:320-323).
                const uint64_t b = __ldg(p.row_off + u), e = __ldg(p.row_off + u + 1);
                adjw += (uint32_t)(e - b);
                uint32_t ncand = 0;  // fresh ids waiting in cand[] (2 x kCandCap words in this mode)
                auto load_ids = [&](uint64_t off, uint32_t (&nid)[3]) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint64_t pos = off + j * 32 + lane;
                        nid[j] = pos < e ? __ldg(p.csr_nbrs + pos) : 0u;
                    }
                };
                uint32_t cur[3], nxt[3] = {0u, 0u, 0u};
                if (b < e) load_ids(b, cur);
                for (uint64_t off = b; off < e; off += kCandCap) {
                    if (off + kCandCap < e) load_ids(off + kCandCap, nxt);  // in flight while this pass is filtered / evaluated
                    const uint32_t cnt = (uint32_t)min((uint64_t)kCandCap, e - off);
                    ncand = filter_unvisited_ids(w, tags, (uint8_t)epoch, cur, cnt, ncand, 0xffffffffu);
                    if (ncand >= 64u) {
                        cmps += ncand;
                        evaluate<kL2>(w, p, ncand, false);
                        ncand = 0;
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) cur[j] = nxt[j];
                }
                if (ncand) {
                    cmps += ncand;
                    evaluate<kL2>(w, p, ncand, false);
                }
            }
        }

        if (p.lazy_queue) queue_flush(w, p.L);
        if (kBuild) {
            if (lane == 0) p.exp_cnt[qidx] = min(hops, p.exp_cap);
        } else {
            // :2408-2418
            if (w.size < p.k && lane == 0) atomicCAS(p.stats + 3, 0ull, ((unsigned long long)w.size << 32) | 1ull);
            for (uint32_t i = lane; i < p.k; i += 32) {
                const bool ok = i < w.size;
                p.out_ids[(size_t)qidx * p.k + i] = ok ? (w.qi[i] & kIdMask) : 0xffffffffu;
                p.out_dists[(size_t)qidx * p.k + i] = ok ? w.qd[i] : INFINITY;  // missing slots sort last in any merge
            }
        }
        if (lane == 0) {
            if (p.out_cmps) p.out_cmps[qidx] = cmps;
            if (p.out_hops) p.out_hops[qidx] = hops;
        }
        acc_cmps += cmps;
        acc_hops += hops;
        acc_adj += adjw;
        __syncwarp();
    }
    if (lane == 0) {
        p.slot_epoch[slot] = epoch;
        if (acc_hops) {
            atomicAdd(p.stats + 0, acc_cmps);
            atomicAdd(p.stats + 1, acc_hops);
            atomicAdd(p.stats + 2, acc_adj);
        }
    }
}

// Visited-tag maintenance before a launch: a slot whose 8-bit epoch could wrap during the coming launch gets its tag array
// cleared here by a whole grid (a 10 MB clear by the one warp that owns the slot takes tens of milliseconds under load and
// stalls the launch; the in-kernel clear remains as the fallback for a slot that still overruns).
constexpr uint64_t kClearChunk = 1ull << 20;
__global__ void tags_clear_kernel(uint8_t *tags, const uint32_t *slot_epoch, uint64_t tag_stride, uint32_t thr) {
    const uint32_t slot = blockIdx.y;
    if (slot_epoch[slot] < thr) return;
    const uint64_t off = (uint64_t)blockIdx.x * kClearChunk;
    if (off >= tag_stride) return;
    const uint64_t n16 = (min(kClearChunk, tag_stride - off)) >> 4;  // tag_stride is a multiple of 16
    uint4 *dst = reinterpret_cast<uint4 *>(tags + (size_t)slot * tag_stride + off);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = z;
}
__global__ void tags_reset_kernel(uint32_t *slot_epoch, uint32_t n_slots, uint32_t thr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots && slot_epoch[i] >= thr) slot_epoch[i] = 0;
}
int maintain_tags(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    const uint32_t slots = l.grid * l.warps_per_cta;
    const uint32_t per_slot = (p.nq + slots - 1) / slots;
    // a slot is cleared here when its epoch could wrap during this launch: the dynamic tickets give a slot about per_slot
    // queries, rarely more than 1.5x that (a slot that still overruns clears its own array inside the kernel)
    const uint32_t margin = std::min<uint32_t>(250u, per_slot + per_slot / 2u + 8u);
    const uint32_t thr = 255u - margin;
    dim3 grid((unsigned)((p.tag_stride + kClearChunk - 1) / kClearChunk), slots);
    tags_clear_kernel<<<grid, 256, 0, s>>>(p.tags, p.slot_epoch, p.tag_stride, thr);
    tags_reset_kernel<<<(slots + 255) / 256, 256, 0, s>>>(p.slot_epoch, slots, thr);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

// ---- query ordering ---------------------------------------------------------------------------------------------------------
// Queries are independent, so the order in which the tickets hand them out is unobservable; it decides, however, which rows
// the ~2 000 resident warps gather at the same time.  Grouping the batch by the nearest of kLandmarks evenly spaced base rows
// makes concurrently resident queries walk the same region of the graph, so that a row one warp has fetched is still in the
// 126 MB L2 when its neighbours ask for it.  One warp per query scores the landmarks (their rows stay in L2), a counting sort
// over the landmark ids produces the order; the whole pass costs a few tens of microseconds per 10 000 queries.
__global__ void landmark_assign_kernel(const float *__restrict__ base, uint64_t nd, uint32_t dim, int metric_l2,
                                       const float *__restrict__ queries, uint32_t nq, uint32_t *__restrict__ keys,
                                       uint32_t *__restrict__ bins) {
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    const float *qr = queries + (size_t)q * dim;
    const uint64_t step = nd / kLandmarks ? nd / kLandmarks : 1;
    float best = INFINITY;
    uint32_t best_l = 0;
    for (uint32_t l0 = 0; l0 < kLandmarks; l0 += 4) {  // four landmarks per pass: independent loads
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float *br[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) br[t] = base + min((uint64_t)(l0 + t) * step, nd - 1) * dim;
        for (uint32_t j = lane * 4; j < dim; j += 128) {
            const float4 qv = *reinterpret_cast<const float4 *>(qr + j);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(br[t] + j));
                if (metric_l2) {
                    const float d0 = qv.x - x.x, d1 = qv.y - x.y, d2 = qv.z - x.z, d3 = qv.w - x.w;
                    acc[t] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                } else {
                    acc[t] -= qv.x * x.x + qv.y * x.y + qv.z * x.z + qv.w * x.w;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(kFull, acc[t], o);
            if (acc[t] < best) {
                best = acc[t];
                best_l = l0 + t;
            }
        }
    }
    if (lane == 0) {
        keys[q] = best_l;
        atomicAdd(bins + best_l, 1u);
    }
}
__global__ void landmark_scan_kernel(uint32_t *bins) {  // exclusive scan of kLandmarks counters (one CTA of kLandmarks threads)
    __shared__ uint32_t s[kLandmarks];
    const uint32_t t = threadIdx.x;
    s[t] = bins[t];
    __syncthreads();
    for (uint32_t o = 1; o < kLandmarks; o <<= 1) {
        const uint32_t v = t >= o ? s[t - o] : 0u;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    bins[t] = s[t] - bins[t];
}
__global__ void landmark_scatter_kernel(const uint32_t *keys, uint32_t nq, uint32_t *bins, uint32_t *order) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) order[atomicAdd(bins + keys[q], 1u)] = q;
}

uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return (uint32_t)strtoul(v, nullptr, 10);
}

constexpr uint32_t kMaxSmemPerCta = 227u * 1024u;

uint32_t warp_smem_bytes(uint32_t dim, uint32_t rstride, uint32_t R, uint32_t L, uint32_t vt_bytes = 0, uint32_t cand_words = kCandCap) {
    // cand_words carries the pending block of the lazy queue too (plan_search adds 2 * kPendSlots words)
    uint32_t b = 64 + dim * 4 + R * rstride * 4 + L * 8 + cand_words * 4;
    if (vt_bytes) b += 32 + vt_bytes;  // the shared-memory visited table, 32-byte aligned behind the candidate ids
    return (b + 127u) & ~127u;
}

template <bool kL2, int kMode, int kMinBlocks, bool kVt = false>
int launch_v(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    auto kern = beam_search_kernel<kL2, kMode, kMinBlocks, kVt>;
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes));
    static const bool trace = getenv("RG_TRACE_SEARCH") != nullptr;  // diagnostic: device time of each phase, synchronises
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace) {
        for (auto &e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], s);
    }
    RG_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(uint32_t), s));
    RG_CUDA(cudaMemsetAsync(p.stats, 0, 4 * sizeof(unsigned long long), s));
    if (trace) cudaEventRecord(ev[1], s);
    int rc = maintain_tags(p, l, s);
    if (rc != RG_OK) return rc;
    if (trace) cudaEventRecord(ev[2], s);
    kern<<<l.grid, l.block, l.smem_bytes, s>>>(p);
    RG_CUDA(cudaGetLastError());
    if (trace) {
        cudaEventRecord(ev[3], s);
        cudaStreamSynchronize(s);
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        fprintf(stderr, "[rg search] memsets %.3f ms, tag maintenance %.3f ms, kernel %.3f ms\n", a, b, c);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return RG_OK;
}

template <bool kL2, int kMode>
int launch_t(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    if (kMode == 0 && (p.vtab || p.vt_smem)) return launch_v<kL2, 0, 4, true>(p, l, s);  // needs the 4-CTA register budget
    if (l.warps_per_sm > 24) return launch_v<kL2, kMode, 8>(p, l, s);  // 64 registers
    if (l.warps_per_sm > 16) return launch_v<kL2, kMode, 6>(p, l, s);  // 80
    return launch_v<kL2, kMode, 4>(p, l, s);                           // up to 127
}

}  // namespace

uint32_t max_search_slots(int sm_count) { return (uint32_t)sm_count * 32u; }

// Warps per CTA and CTAs per SM for a per-warp working set of `wb` bytes: 4 warps per CTA unless a finer CTA granularity
// keeps more warps resident (long queues: the per-warp working set grows with L_pq, and whole 4-warp CTAs then waste up to
// a third of the shared memory).  At most 32 warps per SM (64 registers per thread), at most 32 CTAs per SM.
// Returns the resident warps per SM (0: one warp does not fit).
static uint32_t resident_geometry(uint32_t wb, uint32_t env_warps, uint32_t &warps, uint32_t &ctas) {
    warps = 4;
    ctas = 0;
    uint32_t best_total = 0;
    for (uint32_t wc = 4; wc >= 1; wc >>= 1) {
        if (env_warps && wc != env_warps) continue;
        if (wc * wb > kMaxSmemPerCta) continue;
        uint32_t c = kMaxSmemPerCta / (wc * wb + 1024u);
        c = std::min(c, 32u / wc);
        const uint32_t total = c * wc;
        if (total > best_total + (best_total >> 3)) {  // a finer granularity must buy > 12 % more resident warps
            best_total = total;
            warps = wc;
            ctas = c;
        }
    }
    return best_total;
}

int plan_search(SearchParams &p, SearchLaunch &l, int sm_count, uint32_t max_slots) {
    const uint32_t dim = p.dim;
    p.row_stride_smem = dim + ((16u + 32u - (dim & 31u)) & 31u);  // stride % 32 == 16 -> conflict-free float4 reads
    const uint32_t env_warps = env_u32("RG_WARPS_PER_CTA", 0);
    const uint32_t vt_bytes = p.vt_smem ? p.vt_words * 4u : 0u;
    // candidate-id scratch per warp: one 96-id pass; the bipartite search (row_off set) gathers the fresh ids of several
    // passes before it evaluates them, so that whole 8-row groups go through the row ring
    p.cand_words = p.row_off ? 2u * (uint32_t)kCandCap : (uint32_t)kCandCap;
    // long queues: lazy form (pending block, one shift per ~30 accepted candidates); RG_LAZY_QUEUE_MIN_L moves the threshold
    // (0 = never)
    {
        const uint32_t min_l = env_u32("RG_LAZY_QUEUE_MIN_L", 256);
        p.lazy_queue = (min_l != 0 && p.L >= min_l && !p.vtab && !p.vt_smem) ? 1u : 0u;
    }
    const uint32_t cw = p.cand_words + (p.lazy_queue ? 2u * kPendSlots : 0u);
    uint32_t warps = 4, ctas = 0;
    uint32_t R = env_u32("RG_ROWS_PER_BATCH", 0);
    if (R == 0) {
        // Row ring: 8 rows single-buffered, or 2 x 8 double-buffered (eight rows stay in flight while the warp computes
        // distances and inserts).  The second ring costs resident warps; measured on the 10 M x 200 L_pq sweep
        // (profiles/r1_sweep_10M_ring.csv) it pays when the single-buffered geometry already saturates the memory system
        // (>= 20 warps per SM) or when it keeps at least 72 % of the resident warps.
        uint32_t w8c, c8, w16c, c16;
        const uint32_t w8 = resident_geometry(warp_smem_bytes(dim, p.row_stride_smem, 8, p.L, vt_bytes, cw), env_warps, w8c, c8);
        const uint32_t w16 = resident_geometry(warp_smem_bytes(dim, p.row_stride_smem, 16, p.L, vt_bytes, cw), env_warps, w16c, c16);
        R = (w16 > 0 && (w8 >= 20 || w16 * 100u >= w8 * 72u)) ? 16u : 8u;
    }
    if (R < 8) R = 8;
    if (R > (uint32_t)kMaxRowsPerBatch) R = kMaxRowsPerBatch;
    R &= ~7u;
    uint32_t wb = warp_smem_bytes(dim, p.row_stride_smem, R, p.L, vt_bytes, cw);
    while (wb > kMaxSmemPerCta && R > 8) {
        R -= 8;
        wb = warp_smem_bytes(dim, p.row_stride_smem, R, p.L, vt_bytes, cw);
    }
    if (wb > kMaxSmemPerCta) {
        set_error("L_pq / dim too large for one warp's shared-memory working set");
        return RG_EINVAL;
    }
    resident_geometry(wb, env_warps, warps, ctas);
    if (ctas < 1) {
        warps = 1;
        ctas = 1;
    }
    const uint32_t env_ctas = env_u32("RG_CTAS_PER_SM", 0);
    if (env_ctas && env_ctas < ctas) ctas = env_ctas;
    uint32_t grid = (uint32_t)sm_count * ctas;
    const uint32_t need = (p.nq + warps - 1) / warps;
    if (grid > need) grid = need;
    if (env_u32("RG_WAVE_FIT", 0) && (uint64_t)grid * warps < p.nq) {
        // opt-in: trim the grid so that the batch is a whole number of waves (10 000 queries on 2 072 slots = 4.83 waves ->
        // 2 000 slots, 5.0 waves): the last wave then ends together instead of leaving 17 % of the slots idle for a query's
        // duration.  Dynamic tickets already balance most of it; measured before it becomes the default.
        const uint32_t slots = grid * warps;
        const uint32_t waves = (p.nq + slots - 1) / slots;
        const uint32_t fit = (p.nq + waves - 1) / waves;
        grid = std::max(1u, std::min(grid, (fit + warps - 1) / warps));
    }
    if (grid * warps > max_slots) grid = max_slots / warps;
    if (grid < 1) grid = 1;
    p.rows_per_batch = R;
    p.merge_insert = env_u32("RG_QUEUE_MERGE", 1) ? 1u : 0u;
    p.warp_smem_bytes = wb;
    l.grid = grid;
    l.block = warps * 32;
    l.smem_bytes = warps * wb;
    l.rows_per_batch = R;
    l.warps_per_cta = warps;
    l.ctas_per_sm = ctas;
    l.warps_per_sm = ctas * warps;
    return RG_OK;
}

int order_queries(const float *base, uint64_t nd, uint32_t dim, int metric_l2, const float *queries, uint32_t nq, uint32_t *scratch,
                  cudaStream_t s) {
    uint32_t *keys = scratch, *order = scratch + nq, *bins = scratch + 2 * (size_t)nq;
    RG_CUDA(cudaMemsetAsync(bins, 0, (kLandmarks + 1) * sizeof(uint32_t), s));
    landmark_assign_kernel<<<(nq + 7) / 8, 256, 0, s>>>(base, nd, dim, metric_l2, queries, nq, keys, bins);
    landmark_scan_kernel<<<1, kLandmarks, 0, s>>>(bins);
    landmark_scatter_kernel<<<(nq + 255) / 256, 256, 0, s>>>(keys, nq, bins, order);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int launch_roargraph(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    return p.metric_l2 ? launch_t<true, 0>(p, l, s) : launch_t<false, 0>(p, l, s);
}
int launch_bipartite(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    return p.metric_l2 ? launch_t<true, 1>(p, l, s) : launch_t<false, 1>(p, l, s);
}
int launch_build_search(const SearchParams &p, const SearchLaunch &l, cudaStream_t s) {
    return p.metric_l2 ? launch_t<true, 2>(p, l, s) : launch_t<false, 2>(p, l, s);
}

}  // namespace rg
