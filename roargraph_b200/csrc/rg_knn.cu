// rg_knn.cu — brute-force training-query x base kNN on sm_100a (exact_knn of the reference's compute_groundtruth,
// thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:126-248, paths relative to /root/reference).
//
// The reference materialises a 1024 x npoints fp32 distance matrix per query batch with MKL SGEMM and re-streams it
// through a per-query max-heap (:190-234).  Here the distance tile never leaves the SM: a tiled fp32 GEMM computes
// 128 x 128 scores per CTA and its epilogue keeps only scores below the query's running K-th best (tau), appending
// them to a small per-query candidate list; a merge kernel folds the candidates into the sorted top-K and tightens
// tau.  The base is streamed in chunks of doubling size so that the expected number of survivors per chunk stays ~K.
// Internal distances as the reference defines them: IP -> -dot (:118-120), L2 -> -2 dot + |p|^2 + |q|^2 (:92-100),
// COSINE -> L2 on rows divided by their norm (:146-174).  Ties: earlier id wins (strict '>' at :214).
#include <cfloat>
#include <cstdio>
#include <mutex>
#include <vector>

#include "rg_internal.h"

namespace rg {
namespace {

constexpr int TQ = 128, TB = 128, TK = 16;
constexpr int KNN_THREADS = 256;

__device__ __forceinline__ uint32_t f2ord(float f) {  // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// squared norms, one warp per row (compute_l2sq, :68-75)
__global__ void row_sqnorm_kernel(const float *x, uint64_t n, uint32_t dim, float *out) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float *r = x + row * dim;
    float s = 0.f;
    for (uint32_t j = lane; j < dim; j += 32) s = fmaf(r[j], r[j], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

// rows / sqrt(norm)  (:146-168; epsilon when the norm is 0)
__global__ void row_normalize_kernel(const float *x, uint64_t n, uint32_t dim, const float *sq, float *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    float nm = sqrtf(sq[i / dim]);
    if (nm == 0.f) nm = FLT_EPSILON;
    out[i] = x[i] / nm;
}

// Tile GEMM + threshold filter.  scores(q, b) for q in [q0, q0+TQ), b in [b0, b0+TB) of the current base chunk.
// cand layout: per query `cap` packed (ordered-dist << 32 | id) words.
__global__ void __launch_bounds__(KNN_THREADS) knn_gemm_filter_kernel(
    const float *__restrict__ Q, const float *__restrict__ B, uint32_t dim, uint64_t nq, uint64_t chunk_lo,
    uint64_t chunk_hi, const float *__restrict__ qn, const float *__restrict__ bn, int metric_l2,
    const float *__restrict__ tau, unsigned long long *__restrict__ cand, uint32_t *__restrict__ cand_cnt, uint32_t cap,
    uint32_t *overflow) {
    __shared__ float sQ[TK][TQ + 4];
    __shared__ float sB[TK][TB + 4];
    const int tid = threadIdx.x;
    const uint64_t q0 = (uint64_t)blockIdx.y * TQ;
    const uint64_t b0 = chunk_lo + (uint64_t)blockIdx.x * TB;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each (strided by 16)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader mapping: 128 rows x 16 k per tile = 2048 floats = 8 per thread: row = tid >> 1, k-half = (tid & 1) * 8
    const int lrow = tid >> 1, lk = (tid & 1) * 8;
    const uint64_t qrow = q0 + lrow, brow = b0 + lrow;
    const bool qok = qrow < nq, bok = brow < chunk_hi;
    const float *qp = Q + qrow * dim, *bp = B + brow * dim;

    for (uint32_t k0 = 0; k0 < dim; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t kk = k0 + lk + j;
            sQ[lk + j][lrow] = (qok && kk < dim) ? __ldg(qp + kk) : 0.f;
            sB[lk + j][lrow] = (bok && kk < dim) ? __ldg(bp + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sQ[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = sB[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint64_t q = q0 + ty + 16 * i;
        if (q >= nq) continue;
        const float t = tau[q];
        const float qq = metric_l2 ? qn[q] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint64_t b = b0 + tx + 16 * j;
            if (b >= chunk_hi) continue;
            float s;
            if (metric_l2) s = (-2.0f * acc[i][j] + bn[b]) + qq;
            else s = -acc[i][j];
            if (s < t) {
                const uint32_t slot = atomicAdd(cand_cnt + 2 * q, 1u);  // sub-list 0 (see CandList)
                if (slot < cap) cand[q * cap + slot] = ((unsigned long long)f2ord(s) << 32) | (uint32_t)b;
                else *overflow = 1u;
            }
        }
    }
}

// Fold candidates into the running sorted top-K of each query (one CTA per query), update tau, reset the counter.
// top: nq x K packed words ascending (unused tail = ~0ull).
// With `rr.base != nullptr` the candidates carry only a base id (tensor-core filter, rg_knn_tc.cu) and their exact fp32
// internal distance is computed here first: one warp per candidate, query row staged in shared memory.
struct Rerank {
    const float *base, *queries;  // fp32 rows, row stride == dim
    const float *qn, *bn;         // squared norms (L2)
    uint32_t dim;
    int metric_l2;
    int vec;  // rows are 16-byte aligned with dim % 4 == 0: float4 loads (and that accumulation order) in both merge kernels
};

// A query's candidate list is `cap` words; the 2-CTA tensor-core filter fills it as two private sub-lists (one per
// column half of a tile: [0, cap/2) and [cap/2, cap), counters cnt[2q] and cnt[2q+1]) so that no atomic sits on its
// epilogue path; every other producer appends to sub-list 0 alone, which may then use the whole capacity.
struct CandList {
    uint32_t n0, n1, half_cap;
    __device__ CandList(const uint32_t *cand_cnt, uint64_t q, uint32_t cap) {
        n0 = cand_cnt[2 * q];
        n1 = cand_cnt[2 * q + 1];
        half_cap = cap >> 1;
        if (n1) {
            n0 = min(n0, half_cap);
            n1 = min(n1, half_cap);
        } else {
            n0 = min(n0, cap);
        }
    }
    __device__ uint32_t size() const { return n0 + n1; }
    __device__ uint32_t slot(uint32_t i) const { return i < n0 ? i : half_cap + (i - n0); }
};

__device__ __forceinline__ void bitonic_sort_smem(unsigned long long *sk, uint32_t m) {
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == up) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

// the same network for m <= 64, run by ONE warp (lane i owns comparator i): no block barrier per step
__device__ __forceinline__ void bitonic_sort_warp(unsigned long long *sk, uint32_t m) {
    const uint32_t i = threadIdx.x & 31;
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            if (i < (m >> 1)) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == up) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncwarp();
        }
    }
}

// number of entries of the ascending list a[0, n) below `key`
__device__ __forceinline__ uint32_t count_below(const unsigned long long *a, uint32_t n, unsigned long long key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// General path (any candidate count up to cap): one CTA of 256 threads per query of the work list.
__device__ void merge_query_big(uint64_t q, unsigned long long *__restrict__ top, uint32_t K, unsigned long long *__restrict__ cand,
                                uint32_t *__restrict__ cand_cnt, uint32_t cap, float *__restrict__ tau, uint32_t sort_n,
                                const Rerank &rr, unsigned long long *sk) {
    const CandList cl(cand_cnt, q, cap);
    const uint32_t n = cl.size();
    if (n == 0) return;
    unsigned long long *cq = cand + q * cap;
    if (rr.base) {
        float *sq = reinterpret_cast<float *>(sk + sort_n);  // dim floats after the sort buffer (16-byte aligned: sort_n * 8)
        for (uint32_t j = threadIdx.x; j < rr.dim; j += blockDim.x) sq[j] = rr.queries[q * rr.dim + j];
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const float qq = rr.metric_l2 ? rr.qn[q] : 0.f;
        // four candidates per warp pass: their row loads are independent, so the DRAM latencies overlap.  Every candidate's
        // partial sums are formed in exactly the order of knn_merge_fast_kernel (lane-strided float4 groups x, y, z, w when
        // `vec`, lane-strided scalars otherwise; xor-shuffle tree), so a distance does not depend on which kernel re-ranked it
        for (uint32_t i0 = warp * 4; i0 < n; i0 += nwarps * 4) {
            uint32_t b[4];
            const float *br[4];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                b[t] = (uint32_t)cq[cl.slot(min(i0 + t, n - 1))];
                br[t] = rr.base + (uint64_t)b[t] * rr.dim;
            }
            if (rr.vec) {
                for (uint32_t j = lane; j < rr.dim / 4; j += 32) {
                    const float4 qv = reinterpret_cast<const float4 *>(sq)[j];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float4 x = __ldg(reinterpret_cast<const float4 *>(br[t]) + j);
                        acc[t] = fmaf(qv.x, x.x, acc[t]);
                        acc[t] = fmaf(qv.y, x.y, acc[t]);
                        acc[t] = fmaf(qv.z, x.z, acc[t]);
                        acc[t] = fmaf(qv.w, x.w, acc[t]);
                    }
                }
            } else {
                for (uint32_t j = lane; j < rr.dim; j += 32) {
                    const float qv = sq[j];
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fmaf(qv, __ldg(br[t] + j), acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
                for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
            __syncwarp();
            if (lane < 4 && i0 + lane < n) {
                const float a = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
                const uint32_t bb = lane == 0 ? b[0] : lane == 1 ? b[1] : lane == 2 ? b[2] : b[3];
                const float d = rr.metric_l2 ? (-2.0f * a + rr.bn[bb]) + qq : -a;
                cq[cl.slot(i0 + lane)] = ((unsigned long long)f2ord(d) << 32) | bb;
            }
            __syncwarp();
        }
        __syncthreads();
    }
    // smallest power of two holding n + K
    uint32_t m = 1;
    while (m < n + K) m <<= 1;
    if (m > sort_n) m = sort_n;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        unsigned long long v = ~0ull;
        if (i < n) v = cq[cl.slot(i)];
        else if (i < n + K) v = top[q * K + (i - n)];
        sk[i] = v;
    }
    __syncthreads();
    bitonic_sort_smem(sk, m);
    for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) top[q * K + i] = sk[i];
    if (threadIdx.x == 0) {
        const unsigned long long kth = sk[K - 1];
        if (kth != ~0ull) tau[q] = fminf(tau[q], ord2f((uint32_t)(kth >> 32)));  // tau only tightens (it may have been seeded)
        cand_cnt[2 * q] = 0;
        cand_cnt[2 * q + 1] = 0;
    }
    __syncthreads();
}

// work_cnt == nullptr: one CTA per query (blockIdx.x); otherwise the CTAs walk the work list left by the fast kernel
__global__ void knn_merge_kernel(unsigned long long *__restrict__ top, uint32_t K, unsigned long long *__restrict__ cand,
                                 uint32_t *__restrict__ cand_cnt, uint32_t cap, float *__restrict__ tau, uint32_t sort_n,
                                 const Rerank rr, const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ work_cnt) {
    extern __shared__ unsigned long long sk[];
    if (!work_cnt) {
        merge_query_big(blockIdx.x, top, K, cand, cand_cnt, cap, tau, sort_n, rr, sk);
        return;
    }
    const uint32_t nw = *work_cnt;
    for (uint32_t w = blockIdx.x; w < nw; w += gridDim.x) merge_query_big(worklist[w], top, K, cand, cand_cnt, cap, tau, sort_n, rr, sk);
}

// Fast path of the tensor-core pipeline (candidates carry ids only): one CTA of 128 threads per query, at most `sn` - K
// candidates (more: the query goes to the work list of the general kernel).  Exact fp32 distances of eight candidates per
// warp pass (128-bit row loads when VEC), candidates that do not beat the current K-th best are dropped before the sort,
// and the sort buffer holds only `sn` words, so sixteen queries are resident per SM and the gather latencies overlap.
template <bool VEC, bool kMergeByRank = true>
__global__ void __launch_bounds__(128) knn_merge_fast_kernel(unsigned long long *__restrict__ top, uint32_t K,
                                                             const unsigned long long *__restrict__ cand,
                                                             uint32_t *__restrict__ cand_cnt, uint32_t cap, float *__restrict__ tau,
                                                             uint32_t sn, const Rerank rr, uint32_t *__restrict__ worklist,
                                                             uint32_t *__restrict__ work_cnt, uint64_t nq) {
    extern __shared__ unsigned long long sk[];
    __shared__ uint32_t s_kept;
    // grid-stride over the queries: one CTA per query when the grid is nq, or a small persistent grid (two CTAs per SM) when
    // this merge has to fit beside the other query lane's filter CTA (rg_knn.cu, two query lanes)
    for (uint64_t q = blockIdx.x; q < nq; q += gridDim.x) {
    __syncthreads();  // the previous query's sort buffer and s_kept are no longer read
    const CandList cl(cand_cnt, q, cap);
    const uint32_t n = cl.size();
    if (n == 0) continue;
    if (n + K > sn) {
        if (threadIdx.x == 0) worklist[atomicAdd(work_cnt, 1u)] = (uint32_t)q;
        continue;
    }
    float *sq = reinterpret_cast<float *>(sk + sn);
    if (VEC) {
        const float4 *q4 = reinterpret_cast<const float4 *>(rr.queries + q * rr.dim);
        for (uint32_t j = threadIdx.x; j < rr.dim / 4; j += blockDim.x) reinterpret_cast<float4 *>(sq)[j] = q4[j];
    } else {
        for (uint32_t j = threadIdx.x; j < rr.dim; j += blockDim.x) sq[j] = rr.queries[q * rr.dim + j];
    }
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    const unsigned long long kth = top[q * K + K - 1];  // (distance, id) of the current K-th best, ~0 while the list is short
    const unsigned long long *cq = cand + q * cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float qq = rr.metric_l2 ? rr.qn[q] : 0.f;
    for (uint32_t i0 = warp * 8; i0 < n; i0 += 32) {
        uint32_t myb = 0;
        if (lane < 8 && i0 + lane < n) myb = (uint32_t)cq[cl.slot(i0 + lane)];
        float acc[8];
        uint32_t b[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            b[t] = __shfl_sync(0xffffffffu, myb, t);  // tail slots repeat row 0 of the base: harmless, dropped below
            acc[t] = 0.f;
        }
        if (VEC) {
            for (uint32_t j = lane; j < rr.dim / 4; j += 32) {
                const float4 qv = reinterpret_cast<const float4 *>(sq)[j];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 x = __ldg(reinterpret_cast<const float4 *>(rr.base + (uint64_t)b[t] * rr.dim) + j);
                    acc[t] = fmaf(qv.x, x.x, acc[t]);
                    acc[t] = fmaf(qv.y, x.y, acc[t]);
                    acc[t] = fmaf(qv.z, x.z, acc[t]);
                    acc[t] = fmaf(qv.w, x.w, acc[t]);
                }
            }
        } else {
            for (uint32_t j = lane; j < rr.dim; j += 32) {
                const float qv = sq[j];
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = fmaf(qv, __ldg(rr.base + (uint64_t)b[t] * rr.dim + j), acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
            for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        float a = acc[0];
#pragma unroll
        for (int t = 1; t < 8; ++t) a = lane == t ? acc[t] : a;
        bool keep = false;
        unsigned long long key = 0;
        if (lane < 8 && i0 + lane < n) {
            const float d = rr.metric_l2 ? (-2.0f * a + rr.bn[myb]) + qq : -a;
            key = ((unsigned long long)f2ord(d) << 32) | myb;
            keep = key < kth;
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, keep);
        uint32_t base_slot = 0;
        if (lane == 0 && vote) base_slot = atomicAdd(&s_kept, (uint32_t)__popc(vote));
        base_slot = __shfl_sync(0xffffffffu, base_slot, 0);
        if (keep) sk[base_slot + __popc(vote & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    const uint32_t kept = s_kept;
    uint32_t m2 = 1;
    while (m2 < kept) m2 <<= 1;
    if (kept && kMergeByRank && m2 + K <= sn) {
        // The current list is already sorted: sort the `kept` new entries alone (typically ~K/2: a 64-entry network inside one
        // warp instead of a 256-entry one with a block barrier per step) and merge by RANK — entry i of one list lands at
        // i + (entries of the other list below it); the keys (distance, id) are unique, so this is the full sort's order.
        unsigned long long *st = sk + m2;
        for (uint32_t i = kept + threadIdx.x; i < m2; i += blockDim.x) sk[i] = ~0ull;
        for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) st[i] = top[q * K + i];
        __syncthreads();
        if (m2 <= 64) {
            if (warp == 0) bitonic_sort_warp(sk, m2);
            __syncthreads();
        } else {
            bitonic_sort_smem(sk, m2);
        }
        for (uint32_t i = threadIdx.x; i < kept + K; i += blockDim.x) {
            const bool fresh = i < kept;
            const unsigned long long key = fresh ? sk[i] : st[i - kept];
            const uint32_t pos = fresh ? i + count_below(st, K, key) : (i - kept) + count_below(sk, kept, key);
            if (pos < K) {
                top[q * K + pos] = key;
                // tau only tightens (it may have been seeded)
                if (pos == K - 1 && key != ~0ull) tau[q] = fminf(tau[q], ord2f((uint32_t)(key >> 32)));
            }
        }
    } else if (kept) {
        uint32_t m = 1;
        while (m < kept + K) m <<= 1;  // <= sn because kept <= n
        for (uint32_t i = kept + threadIdx.x; i < m; i += blockDim.x) sk[i] = i < kept + K ? top[q * K + (i - kept)] : ~0ull;
        __syncthreads();
        bitonic_sort_smem(sk, m);
        for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) top[q * K + i] = sk[i];
        if (threadIdx.x == 0) {
            const unsigned long long nk = sk[K - 1];
            if (nk != ~0ull) tau[q] = fminf(tau[q], ord2f((uint32_t)(nk >> 32)));
        }
    }
    if (threadIdx.x == 0) {
        cand_cnt[2 * q] = 0;
        cand_cnt[2 * q + 1] = 0;
    }
    }  // query loop
}

__global__ void knn_init_kernel(unsigned long long *top, uint64_t n_top, float *tau, uint32_t *cnt, uint64_t nq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_top) top[i] = ~0ull;
    if (i < nq) {
        tau[i] = INFINITY;
        cnt[2 * i] = 0;
        cnt[2 * i + 1] = 0;
    }
}

__global__ void knn_fill_kernel(float *x, uint64_t n, float v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void knn_min_kernel(float *x, const float *y, uint64_t n) {  // x = min(x, y)
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = fminf(x[i], y[i]);
}

// bound[q] = exact distance of the rank-th best row found so far for query q (+inf while the list is shorter)
__global__ void knn_rank_kernel(const unsigned long long *top, uint32_t K, uint32_t rank, uint64_t nq, float *bound) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const unsigned long long v = top[q * K + rank - 1];
    bound[q] = (v == ~0ull) ? INFINITY : ord2f((uint32_t)(v >> 32));
}

__global__ void knn_unpack_kernel(const unsigned long long *top, uint64_t n, int32_t id_offset, int32_t *ids, float *dists) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = top[i];
    ids[i] = (v == ~0ull) ? -1 : (int32_t)((uint32_t)v) + id_offset;
    dists[i] = (v == ~0ull) ? INFINITY : ord2f((uint32_t)(v >> 32));
}

// merge G sorted lists of K (dist, id) per query: one warp per query, G*K <= 32*32 handled by repeated min-extraction
__global__ void topk_merge_kernel(const float *in_d, const uint32_t *in_i, uint32_t G, uint64_t nq, uint32_t K, float *out_d,
                                  uint32_t *out_i) {
    const uint64_t q = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    // lane g owns the head pointer of list g (G <= 32)
    uint32_t head = 0;
    const bool own = (uint32_t)lane < G;
    const float *ld = in_d + ((uint64_t)lane * nq + q) * K;
    const uint32_t *li = in_i + ((uint64_t)lane * nq + q) * K;
    for (uint32_t o = 0; o < K; ++o) {
        unsigned long long key = ~0ull;
        if (own && head < K) key = ((unsigned long long)f2ord(ld[head]) << 32) | li[head];
        unsigned long long best = key;
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, s);
            best = other < best ? other : best;
        }
        const uint32_t winners = __ballot_sync(0xffffffffu, own && head < K && key == best);
        if (winners && lane == __ffs(winners) - 1) {  // keys embed the id, so ties only arise from identical entries
            out_d[q * K + o] = ld[head];
            out_i[q * K + o] = li[head];
            head++;
        } else if (!winners && lane == 0) {
            out_d[q * K + o] = INFINITY;
            out_i[q * K + o] = 0xffffffffu;
        }
        __syncwarp();
    }
}

// (distance, id) pairs packed into one 8-byte word per entry so that a shard's whole top-K block travels in ONE all-gather:
// high word = the distance's float bits, low word = the id, made global by the shard's first row (start_id,
// compute_groundtruth.cpp:413); the "no result" sentinel 0xffffffff stays what it is.
__global__ void topk_pack_kernel(const float *d, const uint32_t *id, uint64_t n, uint32_t id_offset, unsigned long long *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = id[i];
    out[i] = ((unsigned long long)__float_as_uint(d[i]) << 32) | (v == 0xffffffffu ? v : v + id_offset);
}

// merge of G packed lists of K_in entries per query (each ascending in (distance, id)) into the K_out best, for the queries
// [q_lo, q_hi) of a batch of nq: one warp per query, lane g owns the head of list g; entries with the sentinel id sort last
__global__ void topk_merge_packed_kernel(const unsigned long long *in, uint32_t G, uint64_t nq, uint32_t K_in, uint32_t K_out,
                                         uint64_t q_lo, uint64_t q_hi, float *out_d, uint32_t *out_i) {
    const uint64_t q = q_lo + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= q_hi) return;
    const int lane = threadIdx.x & 31;
    uint32_t head = 0;
    const bool own = (uint32_t)lane < G;
    const unsigned long long *l = in + ((uint64_t)lane * nq + q) * K_in;
    for (uint32_t o = 0; o < K_out; ++o) {
        unsigned long long key = ~0ull, raw = 0;
        if (own && head < K_in) {
            raw = l[head];
            if ((uint32_t)raw != 0xffffffffu) key = ((unsigned long long)f2ord(__uint_as_float((uint32_t)(raw >> 32))) << 32) | (uint32_t)raw;
        }
        unsigned long long best = key;
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, s);
            best = other < best ? other : best;
        }
        const uint32_t winners = best == ~0ull ? 0u : __ballot_sync(0xffffffffu, key == best);
        if (winners && lane == __ffs(winners) - 1) {  // keys embed the global id: equal keys are the same entry seen twice
            out_d[q * K_out + o] = __uint_as_float((uint32_t)(raw >> 32));
            out_i[q * K_out + o] = (uint32_t)raw;
            head++;
        } else if (!winners && lane == 0) {
            out_d[q * K_out + o] = INFINITY;
            out_i[q * K_out + o] = 0xffffffffu;
        }
        __syncwarp();
    }
}

// Scratch buffers of the kNN calls live in a grow-only per-device workspace: cudaMalloc / cudaFree of the multi-GB
// candidate lists and the bf16 base copy cost ~10 % of a one-batch call, and a 10 M-query run re-enters with the same
// sizes.  rg_knn_release_workspace() (knn_release_workspace) gives the memory of the current device back.
struct Workspace {
    static constexpr int kSlots = 20;
    void *p[kSlots] = {};
    size_t cap[kSlots] = {};
    void release() {
        for (int i = 0; i < kSlots; ++i) {
            if (p[i]) cudaFree(p[i]);
            p[i] = nullptr;
            cap[i] = 0;
        }
    }
    int get(int slot, size_t bytes, void **out) {
        if (cap[slot] < bytes) {
            if (p[slot]) cudaFree(p[slot]);
            p[slot] = nullptr;
            cap[slot] = 0;
            RG_CUDA(cudaMalloc(&p[slot], bytes));
            cap[slot] = bytes;
        }
        *out = p[slot];
        return RG_OK;
    }
};
constexpr int kMaxDevices = 64;
Workspace g_ws[kMaxDevices];  // one per device: a workspace is only touched by calls running on its device (one at a time)
enum { WS_BN, WS_QN, WS_BCOPY, WS_QCOPY, WS_TOP, WS_CAND, WS_CNT, WS_TAU, WS_OVF, WS_B16, WS_Q16, WS_MAX, WS_WORK, WS_SEED };

Workspace *current_ws() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_ws[dev];
}

#define RG_WS(ws, ptr, slot, bytes)                                          \
    do {                                                                     \
        void *p_ws_ = nullptr;                                               \
        int rc_ws_ = (ws)->get(slot, (size_t)(bytes), &p_ws_);               \
        if (rc_ws_ != RG_OK) return rc_ws_;                                  \
        (ptr) = reinterpret_cast<decltype(ptr)>(p_ws_);                      \
    } while (0)

// RG_KNN_TRACE=1: device time of every phase (events + a stream synchronisation per phase: a diagnostic, never the timed path)
int g_trace_override = -1;  // rg_knn_trace(): 1 / 0 force the per-phase timers on / off, -1 = follow RG_KNN_TRACE
struct PhaseTimer {
    bool on;
    cudaStream_t s;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    PhaseTimer(cudaStream_t st) : s(st) {
        static const bool env = getenv("RG_KNN_TRACE") != nullptr;
        on = g_trace_override < 0 ? env : g_trace_override != 0;
        if (on) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
        }
    }
    ~PhaseTimer() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    void begin() {
        if (on) cudaEventRecord(e0, s);
    }
    void end(double &acc) {
        if (!on) return;
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        acc += ms;
    }
};

thread_local KnnProfile t_profile;

}  // namespace

const KnnProfile &knn_last_profile() { return t_profile; }

static bool use_tensor_cores(uint32_t dim) {
    const char *m = getenv("RG_KNN_MODE");
    if (m && m[0] == 's') return false;  // RG_KNN_MODE=simt: the fp32 SIMT tile (also taken for dim > 512)
    return knn_tc_supported(dim);
}

}  // namespace rg

// The base side of exact_knn, prepared once for any number of query batches (compute_groundtruth walks the queries
// 1 024 at a time against a base part it keeps in memory, compute_groundtruth.cpp:187-200): squared norms, the
// normalised copy for COSINE (:146-174), the bf16 copy and max norm for the tensor-core filter.
struct rg_knn_base {
    const float *in = nullptr;  // caller's rows (device)
    float *own_in = nullptr;    // device copy of host rows (rg_knn_base_create with a host pointer)
    uint64_t nb = 0;
    uint32_t dim = 0;
    int metric = 0, metric_l2 = 0, device = 0;
    bool tc = false, from_ws = false;
    const float *B = nullptr;  // rows the distances are computed on (the normalised copy for COSINE)
    float *bcopy = nullptr, *bn = nullptr;
    void *b16 = nullptr;
    float bmax = 0.f;
};

namespace rg {

void knn_base_free(rg_knn_base *kb) {
    if (!kb) return;
    if (!kb->from_ws) {
        if (kb->bcopy) cudaFree(kb->bcopy);
        if (kb->bn) cudaFree(kb->bn);
        if (kb->b16) cudaFree(kb->b16);
    }
    if (kb->own_in) cudaFree(kb->own_in);
    delete kb;
}

int knn_base_new(rg_knn_base **out, const float *d_base, float *own_in, uint64_t nb, uint32_t dim, int metric) {
    rg_knn_base *kb = new rg_knn_base();
    kb->own_in = own_in;
    int rc = knn_base_prepare(kb, d_base, nb, dim, metric, /*from_ws=*/false, nullptr);
    if (rc != RG_OK) {
        knn_base_free(kb);
        return rc;
    }
    *out = kb;
    return RG_OK;
}

// from_ws: the derived arrays live in the per-device workspace (one-shot rg_knn_dev calls), else in their own allocations
int knn_base_prepare(rg_knn_base *kb, const float *d_base, uint64_t nb, uint32_t dim, int metric, bool from_ws, cudaStream_t s) {
    if (!d_base || nb == 0 || dim == 0) { set_error("rg_knn: null/empty base"); return RG_EINVAL; }
    if (nb >= 0x7fffffffull) { set_error("rg_knn: nb must be < 2^31 per call (the reference uses int ids)"); return RG_EINVAL; }
    const bool cosine = metric == RG_METRIC_COSINE;
    const int metric_l2 = (metric == RG_METRIC_L2 || cosine) ? 1 : 0;
    if (!metric_l2 && metric != RG_METRIC_INNER_PRODUCT) { set_error("rg_knn: unknown metric"); return RG_EINVAL; }
    RG_CUDA(cudaGetDevice(&kb->device));
    kb->in = d_base;
    kb->nb = nb;
    kb->dim = dim;
    kb->metric = metric;
    kb->metric_l2 = metric_l2;
    kb->tc = use_tensor_cores(dim);
    kb->from_ws = from_ws;
    kb->B = d_base;
    Workspace *ws = current_ws();
    if (!ws) { set_error("rg_knn: no current CUDA device"); return RG_ECUDA; }
    PhaseTimer pt(s);
    pt.begin();
    if (metric_l2 || kb->tc) {  // the tensor-core filter needs the norms for its error bound even for inner product
        if (from_ws) RG_WS(ws, kb->bn, WS_BN, nb * sizeof(float));
        else RG_CUDA(cudaMalloc(&kb->bn, nb * sizeof(float)));
        row_sqnorm_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, s>>>(d_base, nb, dim, kb->bn);
        if (cosine) {
            if (from_ws) RG_WS(ws, kb->bcopy, WS_BCOPY, nb * dim * sizeof(float));
            else RG_CUDA(cudaMalloc(&kb->bcopy, nb * dim * sizeof(float)));
            row_normalize_kernel<<<(unsigned)((nb * dim + 255) / 256), 256, 0, s>>>(d_base, nb, dim, kb->bn, kb->bcopy);
            kb->B = kb->bcopy;
            row_sqnorm_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, s>>>(kb->B, nb, dim, kb->bn);  // :170-171 recompute
        }
    }
    if (kb->tc) {
        const uint32_t kp = knn_tc_pitch(dim);
        if (from_ws) RG_WS(ws, kb->b16, WS_B16, nb * kp * 2);
        else RG_CUDA(cudaMalloc(&kb->b16, nb * kp * 2));
        float *d_max = nullptr;
        RG_WS(ws, d_max, WS_MAX, sizeof(float));
        int rc = knn_tc_convert(kb->B, nb, dim, kb->b16, nullptr, s);
        if (rc != RG_OK) return rc;
        if ((rc = knn_tc_max_norm(kb->bn, nb, d_max, s)) != RG_OK) return rc;
        float m2 = 0.f;
        RG_CUDA(cudaMemcpyAsync(&m2, d_max, sizeof(float), cudaMemcpyDeviceToHost, s));
        RG_CUDA(cudaStreamSynchronize(s));
        kb->bmax = sqrtf(m2);
    }
    RG_CUDA(cudaGetLastError());
    pt.end(t_profile.convert_base_ms);
    return RG_OK;
}

// One or more query batches against a prepared base.  tau_seed (optional, device, nq floats): an externally known upper
// bound of every query's K-th best distance (e.g. agreed between the shards of a sharded base); the ramp then starts from it.
// Seed bound for a SHARDED base: tau_out[q] = an upper bound of the distance of the `rank`-th best row of query q among this
// shard's seed rows (the `rank`-th largest of the bf16 group maxima, widened by the bf16 bound; +inf when the shard is too
// small for a seed pass).  With rank = ceil(K / G) on each of G shards, the MAXIMUM of the shards' values bounds the GLOBAL K-th
// best distance: every shard holds >= rank rows at or below its own value, hence at or below the maximum — G * rank >= K rows.
// One float per query crosses the ranks instead of per-shard candidate lists, and every shard then filters against a threshold
// as tight as K-th-of-(G x 65 536 rows).
int knn_base_seed(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, uint32_t rank, float *d_tau_out, cudaStream_t s) {
    if (!kb || !d_queries || !d_tau_out) { set_error("rg_knn_base_seed: null buffer"); return RG_EINVAL; }
    if (rank == 0 || rank > K) { set_error("rg_knn_base_seed: rank must be in [1, K]"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    {
        int dev = 0;
        RG_CUDA(cudaGetDevice(&dev));
        if (dev != kb->device) RG_CUDA(cudaSetDevice(kb->device));
    }
    Workspace *ws = current_ws();
    if (!ws) { set_error("rg_knn: no current CUDA device"); return RG_ECUDA; }
    const uint64_t n_seed = (kb->tc && kb->metric != RG_METRIC_COSINE) ? knn_tc_seed_rows(kb->nb, K) : 0;
    if (!n_seed) {
        knn_fill_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(d_tau_out, nq, INFINITY);
        RG_CUDA(cudaGetLastError());
        return RG_OK;
    }
    const uint32_t dim = kb->dim;
    const uint64_t qb_max = (uint64_t)sm_count() * 256u * 4u;
    const uint64_t qb_alloc = std::min<uint64_t>(nq, qb_max);
    float *qn = nullptr, *seed_buf = nullptr;
    void *q16 = nullptr;
    RG_WS(ws, qn, WS_QN, nq * sizeof(float));
    RG_WS(ws, q16, WS_Q16, qb_alloc * knn_tc_pitch(dim) * 2);
    RG_WS(ws, seed_buf, WS_SEED, (n_seed / 128u) * qb_alloc * sizeof(float));
    row_sqnorm_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(d_queries, nq, dim, qn);
    for (uint64_t qs = 0; qs < nq; qs += qb_max) {
        const uint64_t nqb = std::min<uint64_t>(qb_max, nq - qs);
        int rc = knn_tc_convert(d_queries + qs * dim, nqb, dim, q16, nullptr, s);
        if (rc != RG_OK) return rc;
        rc = knn_tc_seed(q16, nqb, kb->b16, kb->nb, dim, n_seed, rank, kb->metric_l2, qn + qs, kb->bn, kb->bmax, seed_buf, d_tau_out + qs,
                         sm_count(), s);
        if (rc != RG_OK) return rc;
    }
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int knn_base_search(rg_knn_base *kb, const float *d_queries, uint64_t nq, uint32_t K, int32_t id_offset, int32_t *d_ids,
                    float *d_dists, cudaStream_t s, const float *d_tau_ext, uint64_t bound_rows, const KnnResync *resync) {
    if (!kb || !d_queries || !d_ids || !d_dists) { set_error("rg_knn: null buffer"); return RG_EINVAL; }
    const uint64_t nb = kb->nb;
    const uint32_t dim = kb->dim;
    if (K == 0 || K > nb) { set_error("rg_knn: K must be in [1, nb]"); return RG_EINVAL; }
    if (K > 1024) { set_error("rg_knn: K > 1024 not supported"); return RG_EINVAL; }
    if (resync && (!resync->hook || !resync->d_bound || resync->rank == 0 || resync->rank > K)) {
        set_error("rg_knn: re-sync needs a hook, a bound buffer and a rank in [1, K]");
        return RG_EINVAL;
    }
    if (nq == 0) return RG_OK;
    {
        int dev = 0;
        RG_CUDA(cudaGetDevice(&dev));
        if (dev != kb->device) RG_CUDA(cudaSetDevice(kb->device));
    }
    Workspace *ws = current_ws();
    if (!ws) { set_error("rg_knn: no current CUDA device"); return RG_ECUDA; }
    const bool cosine = kb->metric == RG_METRIC_COSINE;
    const int metric_l2 = kb->metric_l2;
    const bool tc = kb->tc;
    const int sms = sm_count();
    const uint32_t cap = std::max<uint32_t>(2048u, 4u * K);
    // seed pass (tensor-core path, large bases): tau starts from a bound derived from the bf16 scores of the first rows
    // alone; otherwise the first chunk is small because everything in it survives (tau = inf) and chunks then grow so that
    // the expected survivors per chunk stay near K
    const uint64_t n_seed = tc ? knn_tc_seed_rows(nb, K) : 0;
    const uint64_t chunk0 = tc ? 256u : cap;
    // Largest chunk.  Measured on a 10 M-row base (profiles/r2_knn_chunk_ab.txt): every filter + merge launch pair costs a
    // fixed ~1.4 ms for a 151 552-query batch (one merge CTA per query), which outweighs what smaller chunks buy (tau
    // tightening more often; the chunk's fp32 rows staying in L2 for the re-rank): 4 M rows 87 %, 262 144 rows 85 %,
    // 104 704 rows 78 % of the sustained bf16 peak.  RG_KNN_CHUNK_MAX overrides (rows, rounded down to 256).
    uint64_t chunk_max = 1ull << 22;
    if (tc) {
        if (const char *cm = getenv("RG_KNN_CHUNK_MAX")) chunk_max = std::max<uint64_t>(256u, strtoull(cm, nullptr, 10) / 256u * 256u);
    }
    uint32_t sort_n = 1;
    while (sort_n < cap + K) sort_n <<= 1;
    // query batches so that the candidate lists stay within ~2 GiB
    uint64_t qb_max = std::max<uint64_t>(TQ, ((2ull << 30) / (cap * 8ull)) / TQ * TQ);
    if (tc) {  // whole waves of the tensor-core filter: one CTA = 256 queries, one CTA per SM
        const uint64_t wave = (uint64_t)sms * 256u;
        const uint64_t fit = ((2600ull << 20) / (cap * 8ull)) / wave * wave;
        if (fit) qb_max = fit;
    }
    const uint64_t qb_alloc = std::min<uint64_t>(nq, qb_max);
    PhaseTimer pt(s);

    float *qn = nullptr, *qcopy = nullptr;
    const float *Q = d_queries;
    pt.begin();
    if (metric_l2 || tc) {
        RG_WS(ws, qn, WS_QN, nq * sizeof(float));
        row_sqnorm_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(d_queries, nq, dim, qn);
        if (cosine) {
            RG_WS(ws, qcopy, WS_QCOPY, nq * dim * sizeof(float));
            row_normalize_kernel<<<(unsigned)((nq * dim + 255) / 256), 256, 0, s>>>(d_queries, nq, dim, qn, qcopy);
            Q = qcopy;
            row_sqnorm_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(Q, nq, dim, qn);
        }
    }
    void *q16 = nullptr;
    float *seed_buf = nullptr;
    if (tc) {
        RG_WS(ws, q16, WS_Q16, qb_alloc * knn_tc_pitch(dim) * 2);
        if (n_seed) RG_WS(ws, seed_buf, WS_SEED, (n_seed / 128u) * qb_alloc * sizeof(float));
    }
    pt.end(t_profile.convert_queries_ms);
    Rerank rr{};
    if (tc) {
        rr.base = kb->B;
        rr.qn = qn;
        rr.bn = kb->bn;
        rr.dim = dim;
        rr.metric_l2 = metric_l2;
    }
    const size_t merge_smem = (size_t)sort_n * 8 + (tc ? (size_t)dim * 4 : 0);
    unsigned long long *top = nullptr, *cand = nullptr;
    uint32_t *cnt = nullptr, *ovf = nullptr, *worklist = nullptr;
    float *tau = nullptr;
    RG_WS(ws, top, WS_TOP, qb_alloc * K * sizeof(unsigned long long));
    RG_WS(ws, cand, WS_CAND, qb_alloc * cap * sizeof(unsigned long long));
    RG_WS(ws, cnt, WS_CNT, 2 * qb_alloc * sizeof(uint32_t));
    RG_WS(ws, worklist, WS_WORK, (qb_alloc + 2) * sizeof(uint32_t));
    RG_WS(ws, tau, WS_TAU, qb_alloc * sizeof(float));
    RG_WS(ws, ovf, WS_OVF, sizeof(uint32_t));
    RG_CUDA(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem));
    uint32_t fast_sn = 1024;
    while (fast_sn < 4u * K) fast_sn <<= 1;
    if (fast_sn > sort_n) fast_sn = sort_n;
    const size_t fast_smem = (size_t)fast_sn * 8 + (size_t)dim * 4;
    const bool vec = dim % 4 == 0 && (reinterpret_cast<uintptr_t>(kb->B) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;
    rr.vec = vec ? 1 : 0;
    // RG_KNN_MERGE_RANK=0: the earlier full sort of new + current entries (A/B and a second implementation for the tests)
    const char *mr = getenv("RG_KNN_MERGE_RANK");
    const bool by_rank = !(mr && mr[0] == '0');
    auto fast_kernel = vec ? (by_rank ? knn_merge_fast_kernel<true, true> : knn_merge_fast_kernel<true, false>)
                           : (by_rank ? knn_merge_fast_kernel<false, true> : knn_merge_fast_kernel<false, false>);
    RG_CUDA(cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
    uint32_t *work_cnt = worklist + qb_alloc;

    for (uint64_t qs = 0; qs < nq; qs += qb_max) {
        const uint64_t nqb = std::min<uint64_t>(qb_max, nq - qs);
        if (tc) {
            pt.begin();
            int rc = knn_tc_convert(Q + qs * dim, nqb, dim, q16, nullptr, s);
            if (rc != RG_OK) return rc;
            pt.end(t_profile.convert_queries_ms);
            rr.queries = Q + qs * dim;
            rr.qn = qn + qs;
        }
        // Two query lanes (tensor-core path, batches of at least two full waves, no tracing): the batch is split in halves A
        // and B on two streams whose FILTER launches alternate (F(B,c) waits for F(A,c), F(A,c+1) for F(B,c)), so that the
        // re-rank / merge of one half — an HBM gather that leaves the tensor cores idle — runs under the other half's filter.
        // The filter leaves `reserve` bytes of shared memory per SM free so that merge CTAs fit beside its one CTA per SM.
        const uint64_t pair_q = 256;
        const uint64_t half = ((nqb / 2 + pair_q - 1) / pair_q) * pair_q;
        // Opt-in (RG_KNN_LANES=2): measured SLOWER than one lane (151 552 queries, K = 100: 99.6 vs 82.1 ms against a 1.25 M-row
        // shard, 516 vs 494 ms against 10 M rows, profiles/r2_knn_lanes_ab.txt).  With one merge CTA per query the small CTAs
        // flood every SM's shared memory and keep the 195 KB filter CTAs of the other lane out until the merge is over (no
        // overlap); with the persistent two-CTA-per-SM merge grid that does fit beside a filter CTA the gather is
        // latency-bound (5.7 ms per half chunk at 0.7 TB/s against 3.8 ms HBM-bound for the whole chunk) and no longer hides.
        static const bool two_lane_env = getenv("RG_KNN_LANES") && getenv("RG_KNN_LANES")[0] == '2';
        const int n_lanes = (tc && !pt.on && two_lane_env && !resync && nqb >= 2ull * (uint64_t)sms * 256u && half < nqb) ? 2 : 1;
        const uint32_t reserve = n_lanes == 2 ? 2u * (uint32_t)(fast_smem + 1536) : 0u;
        cudaStream_t ls[2] = {s, s};
        cudaEvent_t ev_f[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_init = nullptr;
        if (n_lanes == 2) {
            for (int l = 0; l < 2; ++l) {
                RG_CUDA(cudaStreamCreateWithFlags(&ls[l], cudaStreamNonBlocking));
                RG_CUDA(cudaEventCreateWithFlags(&ev_f[l], cudaEventDisableTiming));
                RG_CUDA(cudaEventCreateWithFlags(&ev_done[l], cudaEventDisableTiming));
            }
            RG_CUDA(cudaEventCreateWithFlags(&ev_init, cudaEventDisableTiming));
        }
        struct Lane { uint64_t off, n; };
        const Lane lanes[2] = {{0, n_lanes == 2 ? half : nqb}, {half, nqb - half}};
        const uint32_t kp16 = tc ? knn_tc_pitch(dim) : 0;
        int rc_lane = RG_OK;
        // Schedules, tried in turn until no candidate list overflows:
        //   kWhole   (only with an externally agreed bound, sharded base) the whole shard in ONE filter + merge launch pair: the
        //            bound is tight from row 0 and tightens locally only once this shard's own list fills, so chunking buys
        //            nothing and every launch pair costs a fixed ~1.4 ms per 151 552-query batch (expected survivors per query
        //            ~ rows / (G * 65 536) * K, a few hundred against a list capacity of 2 048);
        //   kGrowing seeded tau (or tau = inf and a 256-row first chunk), chunks growing with the rows seen (expected survivors
        //            per chunk ~K);
        //   kSafe    every chunk has at most `cap` rows, which cannot overflow.
        enum { kWhole, kGrowing, kSafe };
        // RG_KNN_WHOLE_SHARD=1 opts in to kWhole; RG_KNN_EXT_GROWTH = growth factor of the chunks under an external bound
        const char *we = getenv("RG_KNN_WHOLE_SHARD");
        const bool whole_env = we && we[0] == '1';
        const char *ge = getenv("RG_KNN_EXT_GROWTH");
        const uint64_t ext_growth = ge ? std::max<uint64_t>(2, strtoull(ge, nullptr, 10)) : 2;
        // expected survivors per query with the bound in force over the whole shard: rows * K / (rows the bound is the K-th
        // best of); x 2.5 for the bf16 band and the maximum over the shards, against one sub-list of cap / 2 entries
        const bool whole_fits = bound_rows != 0 && (double)nb * K / (double)bound_rows * 2.5 <= (double)(cap / 2);
        const int first_sched = (d_tau_ext && tc && whole_env && whole_fits) ? kWhole : kGrowing;
        // Re-sync of the shared bound (sharded base): ONCE per query batch, at the first chunk boundary at or past
        // resync->rows (or after the last chunk, so that every shard makes the same number of hook calls whatever its size and
        // schedule), the shard publishes its rank-th best exact distance so far, the hook reduces it to the maximum over the
        // shards, and tau = min(tau, that).  The reduced bound stays valid for a retry in safe mode, which re-applies it.
        bool resynced = false;
        for (int sched = first_sched; sched <= kSafe; ++sched) {
            const int attempt = sched == kSafe ? 1 : 0;
            const uint64_t n_init = std::max<uint64_t>(nqb * K, nqb);
            knn_init_kernel<<<(unsigned)((n_init + 255) / 256), 256, 0, s>>>(top, nqb * K, tau, cnt, nqb);
            RG_CUDA(cudaMemsetAsync(ovf, 0, sizeof(uint32_t), s));
            if (n_lanes == 2) {
                RG_CUDA(cudaEventRecord(ev_init, s));
                for (int l = 0; l < 2; ++l) RG_CUDA(cudaStreamWaitEvent(ls[l], ev_init, 0));
            }
            // an externally agreed bound (sharded base: the maximum over the shards of their rank-ceil(K/G) seed values) replaces
            // the local seed pass; it stays in force in the safe-mode retry as well (it is a bound, not a schedule)
            if (d_tau_ext) knn_min_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, s>>>(tau, d_tau_ext + qs, nqb);
            if (resynced) knn_min_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, s>>>(tau, resync->d_bound + qs, nqb);
            const bool seeded = attempt == 0 && n_seed != 0 && !d_tau_ext;
            if (seeded) {
                pt.begin();
                for (int l = 0; l < n_lanes; ++l) {
                    const Lane &ln = lanes[l];
                    int rc = knn_tc_seed((const char *)q16 + ln.off * kp16 * 2, ln.n, kb->b16, nb, dim, n_seed, K, metric_l2,
                                         qn + qs + ln.off, kb->bn, kb->bmax, seed_buf + (n_seed / 128u) * ln.off, tau + ln.off, sms, ls[l],
                                         reserve);
                    if (rc != RG_OK) return rc;
                }
                pt.end(t_profile.seed_ms);
            }
            uint64_t lo = 0;
            uint64_t csize = attempt == 1 ? (tc ? (uint64_t)(cap / 256u) * 256u : (uint64_t)cap)
                             : sched == kWhole ? nb
                                               : ((seeded || (d_tau_ext && tc)) ? std::max<uint64_t>(n_seed, 65536u) : chunk0);
            bool first_filter = true;
            while (lo < nb) {
                const uint64_t hi = std::min<uint64_t>(nb, lo + csize);
                for (int l = 0; l < n_lanes; ++l) {
                    const Lane &ln = lanes[l];
                    cudaStream_t st = ls[l];
                    unsigned long long *l_top = top + ln.off * K, *l_cand = cand + ln.off * cap;
                    uint32_t *l_cnt = cnt + 2 * ln.off, *l_work = worklist + ln.off, *l_work_cnt = work_cnt + l;
                    float *l_tau = tau + ln.off;
                    Rerank lr = rr;
                    if (tc) {
                        lr.queries = rr.queries + ln.off * dim;
                        lr.qn = rr.qn + ln.off;
                    }
                    pt.begin();
                    if (tc) {
                        // the other lane's latest filter has finished: the two filters never share the tensor cores
                        if (n_lanes == 2 && !first_filter) RG_CUDA(cudaStreamWaitEvent(st, ev_f[1 - l], 0));
                        int rc = knn_tc_filter((const char *)q16 + ln.off * kp16 * 2, ln.n, kb->b16, nb, dim, lo, hi, metric_l2, l_tau,
                                               qn + qs + ln.off, kb->bn, kb->bmax, l_cand, l_cnt, cap, ovf, sms, st, reserve);
                        if (rc != RG_OK) return rc;
                        if (n_lanes == 2) RG_CUDA(cudaEventRecord(ev_f[l], st));
                        first_filter = false;
                    } else {
                        dim3 grid((unsigned)((hi - lo + TB - 1) / TB), (unsigned)((nqb + TQ - 1) / TQ));
                        knn_gemm_filter_kernel<<<grid, KNN_THREADS, 0, s>>>(Q + qs * dim, kb->B, dim, nqb, lo, hi, qn ? qn + qs : nullptr,
                                                                            kb->bn, metric_l2, tau, cand, cnt, cap, ovf);
                    }
                    pt.end(t_profile.filter_ms);
                    pt.begin();
                    if (tc) {
                        // fast path per query, then the general kernel over the (normally empty) work list of long lists
                        RG_CUDA(cudaMemsetAsync(l_work_cnt, 0, sizeof(uint32_t), st));
                        // two lanes: a persistent grid of two CTAs per SM, which fits beside the other lane's filter CTA (one CTA per
                        // query would flood every SM's shared memory and keep the 195 KB filter CTAs out until the merge is over)
                        const unsigned mgrid = n_lanes == 2 ? (unsigned)std::min<uint64_t>(ln.n, 2ull * sms) : (unsigned)ln.n;
                        fast_kernel<<<mgrid, 128, fast_smem, st>>>(l_top, K, l_cand, l_cnt, cap, l_tau, fast_sn, lr, l_work, l_work_cnt, ln.n);
                        // (two lanes: a handful of CTAs — its 34 KB sort buffer does not fit beside a filter CTA, and the list is
                        // normally empty)
                        knn_merge_kernel<<<(unsigned)std::min<uint64_t>(ln.n, n_lanes == 2 ? 16ull : 2ull * sms), 256, merge_smem, st>>>(
                            l_top, K, l_cand, l_cnt, cap, l_tau, sort_n, lr, l_work, l_work_cnt);
                    } else {
                        knn_merge_kernel<<<(unsigned)nqb, 256, merge_smem, s>>>(top, K, cand, cnt, cap, tau, sort_n, rr, nullptr, nullptr);
                    }
                    pt.end(t_profile.merge_ms);
                }
                t_profile.chunks += 1;
                lo = hi;
                if (resync && !resynced && (lo >= resync->rows || lo >= nb)) {
                    pt.begin();
                    knn_rank_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, s>>>(top, K, resync->rank, nqb, resync->d_bound + qs);
                    const int hrc = resync->hook(resync->user, qs, nqb, (void *)s);
                    if (hrc != RG_OK) { set_error("rg_knn: the bound re-sync hook failed"); return hrc; }
                    knn_min_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, s>>>(tau, resync->d_bound + qs, nqb);
                    pt.end(t_profile.merge_ms);
                    resynced = true;
                }
                // next chunk: as many rows as seen so far (expected survivors ~K), three times as many while the rows seen
                // still fit in L2 (tensor-core path without a seed: the re-rank gather of those ~3K survivors never reaches
                // DRAM, and the fixed per-launch cost of filter + merge dominates there)
                // with an externally agreed bound (sharded base) tau is tight from row 0 and tightens locally only once this
                // shard's own list fills, so finer chunks buy nothing: the chunks grow 4x (three launch pairs over a 1.25 M-row
                // shard instead of six; every pair costs a fixed ~1.4 ms per 151 552-query batch)
                if (attempt == 0)
                    csize = std::min<uint64_t>(d_tau_ext ? (ext_growth - 1) * lo : (tc && !seeded && lo < 65536) ? 3 * lo : lo, chunk_max);
            }
            rc_lane = cudaGetLastError() == cudaSuccess ? RG_OK : RG_ECUDA;
            if (n_lanes == 2)
                for (int l = 0; l < 2; ++l) {
                    RG_CUDA(cudaEventRecord(ev_done[l], ls[l]));
                    RG_CUDA(cudaStreamWaitEvent(s, ev_done[l], 0));
                }
            uint32_t h_ovf = 0;
            RG_CUDA(cudaMemcpyAsync(&h_ovf, ovf, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            RG_CUDA(cudaStreamSynchronize(s));
            if (rc_lane != RG_OK) { set_error("rg_knn: kernel launch failed"); break; }
            if (!h_ovf) break;
            t_profile.overflow_retries += 1;
            if (sched == kSafe) { set_error("rg_knn: candidate overflow in safe mode (internal error)"); rc_lane = RG_ENOMEM; break; }
        }
        if (n_lanes == 2) {
            for (int l = 0; l < 2; ++l) {
                cudaStreamDestroy(ls[l]);
                cudaEventDestroy(ev_f[l]);
                cudaEventDestroy(ev_done[l]);
            }
            cudaEventDestroy(ev_init);
        }
        if (rc_lane != RG_OK) return rc_lane;
        knn_unpack_kernel<<<(unsigned)((nqb * K + 255) / 256), 256, 0, s>>>(top, nqb * K, id_offset, d_ids + qs * K, d_dists + qs * K);
    }
    RG_CUDA(cudaGetLastError());
    RG_CUDA(cudaStreamSynchronize(s));
    return RG_OK;
}

int knn_device(const float *d_base, uint64_t nb, const float *d_queries, uint64_t nq, uint32_t dim, uint32_t K,
               int metric, int32_t id_offset, int32_t *d_ids, float *d_dists, cudaStream_t s) {
    if (!d_base || !d_queries || !d_ids || !d_dists) { set_error("rg_knn: null buffer"); return RG_EINVAL; }
    if (K == 0 || K > nb) { set_error("rg_knn: K must be in [1, nb]"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    t_profile = KnnProfile{};
    rg_knn_base *kb = new rg_knn_base();
    int rc = knn_base_prepare(kb, d_base, nb, dim, metric, /*from_ws=*/true, s);
    if (rc == RG_OK) rc = knn_base_search(kb, d_queries, nq, K, id_offset, d_ids, d_dists, s, nullptr, 0, nullptr);
    knn_base_free(kb);
    return rc;
}

void knn_reset_profile() { t_profile = KnnProfile{}; }
void knn_set_trace(int on) { g_trace_override = on; }

int topk_pack_device(const float *d_dists, const uint32_t *d_ids, uint64_t n, uint32_t id_offset, unsigned long long *d_packed,
                     cudaStream_t s) {
    if (!d_dists || !d_ids || !d_packed) { set_error("rg_topk_pack: null buffer"); return RG_EINVAL; }
    if (n == 0) return RG_OK;
    topk_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_dists, d_ids, n, id_offset, d_packed);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int topk_merge_packed_device(const unsigned long long *d_packed, uint32_t G, uint64_t nq, uint32_t K_in, uint32_t K_out,
                             uint64_t q_lo, uint64_t q_hi, float *d_out_dists, uint32_t *d_out_ids, cudaStream_t s) {
    if (!d_packed || !d_out_dists || !d_out_ids) { set_error("rg_topk_merge_packed: null buffer"); return RG_EINVAL; }
    if (G == 0 || G > 32) { set_error("rg_topk_merge_packed: G must be in [1, 32]"); return RG_EINVAL; }
    if (K_in == 0 || K_out == 0 || q_hi > nq || q_lo > q_hi) { set_error("rg_topk_merge_packed: bad K / query range"); return RG_EINVAL; }
    if (q_hi == q_lo) return RG_OK;
    topk_merge_packed_kernel<<<(unsigned)((q_hi - q_lo + 3) / 4), 128, 0, s>>>(d_packed, G, nq, K_in, K_out, q_lo, q_hi, d_out_dists,
                                                                             d_out_ids);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

void knn_release_workspace() {
    if (Workspace *ws = current_ws()) ws->release();
}

int topk_merge_device(const float *d_in_dists, const uint32_t *d_in_ids, uint32_t G, uint64_t nq, uint32_t K,
                      float *d_out_dists, uint32_t *d_out_ids, cudaStream_t s) {
    if (!d_in_dists || !d_in_ids || !d_out_dists || !d_out_ids) { set_error("rg_topk_merge: null buffer"); return RG_EINVAL; }
    if (G == 0 || G > 32) { set_error("rg_topk_merge: G must be in [1, 32]"); return RG_EINVAL; }
    if (nq == 0) return RG_OK;
    topk_merge_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, s>>>(d_in_dists, d_in_ids, G, nq, K, d_out_dists, d_out_ids);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

}  // namespace rg
