// rg_knn_tc.cu — tensor-core candidate filter of the brute-force kNN (exact_knn, compute_groundtruth.cpp:126-248 of
// the reference's vendored DiskANN): Q x B^T on tcgen05 (bf16 operands, fp32 accumulation in TMEM), TMA-staged operands,
// and an epilogue that never writes the score tile anywhere — it only compares each score with the query's running
// K-th best distance (widened by a rigorous bf16 error bound) and appends the rare survivors to the per-query
// candidate list.  The exact fp32 distance of every survivor is recomputed by the merge kernel (rg_knn.cu), so the
// returned neighbours and distances are those of an fp32 computation.
//
// One CTA = 128 queries (UMMA M = 128) x a range of 256-row base tiles (UMMA N = 256), 192 threads:
//   warp 0      TMA producer: the query tile once (resident for the CTA's lifetime), then a ring of 64-column base
//               k-blocks (32 KB each, 128B-swizzled, K-major)
//   warp 1      MMA issuer: one elected lane, tcgen05.mma.cta_group::1.kind::f16, 13 k-steps of 16 for d = 200;
//               tcgen05.commit releases ring stages and publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns at a time from one of TWO 256-column accumulators (all 512
//               TMEM columns), so the scan of tile t overlaps the MMAs of tile t+1.  Thread = query row: its threshold
//               lives in a register.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "rg_internal.h"

namespace rg {
namespace {

constexpr int kBlockM = 128;   // queries per query tile (UMMA M)
constexpr int kAccCols = 256;  // TMEM columns per accumulator set; two sets = all 512 columns
constexpr int kBlockK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr uint32_t kABytesPerKb = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxKb = 8;      // dim <= 512

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, with the descriptors assembled from 32-bit halves inside the asm block (the high word is a launch constant and
// a k-step only moves the low word), so the issuing loop stays on 32-bit integer math
__device__ __forceinline__ void tc_mma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one lane of a converged warp (the same lane every time for a full mask)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// 32 lanes x 32 consecutive fp32 columns of the warp's TMEM lane quadrant (asynchronous: pair with tc_ld_wait)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- cluster / 2-CTA (cta_group::2) primitives -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 rem;\n"
        "mapa.shared::cluster.u32 rem, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n"
        "}\n" ::"r"(bar),
        "r"(rank)
        : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the barrier of the pair's leader (even
// rank): bit 24 of a shared::cluster address is the CTA's rank inside the pair.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(x), "r"(y)
        : "memory");
}
// tcgen05.commit of a 2-CTA MMA group: arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// three-input maximum (FMNMX3 on sm_100): a 32-value maximum is 16 instructions in a depth-4 tree
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// maximum of 32 values; g[0..3] = maxima of v[0..8], v[9..17], v[18..26], v[27..31] (the survivor scan descends only into
// groups that hold a survivor)
__device__ __forceinline__ float max32(const float (&v)[32], float (&g)[4]) {
    float a[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = fmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    g[0] = fmax3(a[0], a[1], a[2]);
    g[1] = fmax3(a[3], a[4], a[5]);
    g[2] = fmax3(a[6], a[7], a[8]);
    g[3] = fmax3(a[9], v[30], v[31]);
    return fmaxf(fmax3(g[0], g[1], g[2]), g[3]);
}
__device__ __forceinline__ float max32(const float (&v)[32]) {
    float g[4];
    return max32(v, g);
}
// bit e set iff v[e] > theta, looking only inside the groups whose maximum passes
__device__ __forceinline__ uint32_t survivor_mask(const float (&v)[32], const float (&g)[4], float theta) {
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (g[k] > theta) {
#pragma unroll
            for (int e = 9 * k; e < (k == 3 ? 32 : 9 * k + 9); ++e) mask |= (v[e] > theta ? 1u : 0u) << e;
        }
    }
    return mask;
}

// K-major, 128-byte swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); version 1 (sm_100);
// layout type 2 = SWIZZLE_128B.  The tile base must be 1024-byte aligned; a k-step of 16 bf16 advances the start by 32 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n, uint32_t m = kBlockM) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct TcParams {
    uint64_t nq;            // queries in this batch (rows of the A tensor map)
    uint64_t nb;            // base rows (rows of the B tensor map)
    uint32_t num_kb;        // 64-column k-blocks
    uint32_t ksteps_last;   // k-steps of 16 in the last k-block
    uint32_t stages;        // ring depth (k-blocks of B)
    uint32_t tile_lo, tile_hi;  // base tiles [tile_lo, tile_hi) of this chunk, in units of the kernel's BLOCK_N
    uint32_t tiles_per_cta;
    int metric_l2;
    int atomic_append;      // several CTAs share a query (base range split over blockIdx.y)
    const float *tau;       // running K-th best internal distance per query
    const float *qn;        // ||q||^2 (exact fp32)
    const float *bn;        // ||b||^2
    float bmax;             // max_b ||b||
    float err_rel;          // |dot_bf16 - dot| <= err_rel * ||q|| * ||b||
    unsigned long long *cand;
    uint32_t *cand_cnt;
    uint32_t cap;
    uint32_t *overflow;
    // seed pass (knn_tc_seed): instead of scanning for survivors every epilogue thread stores the maximum of its 128 columns
    // of each tile, seed_out[((tile - tile_lo) * 2 + half) * seed_stride + query]
    float *seed_out;
    uint64_t seed_stride;
    uint32_t smem_reserve;  // shared memory per SM left free for the merge CTAs of the other query lane (rg_knn.cu)
};

// NQT = query tiles per CTA.  NQT == 2 (dim <= 256): 256 queries share every base tile, which halves the L2 -> SM operand
// traffic per flop — the limiter of this small-K GEMM (profiles/r1_knn_tc_2M_ncu_full.txt); base tiles are then 128 rows
// so that {2 query tiles} x {2 tiles in flight} x 128 columns fill the 512 TMEM columns.  NQT == 1: one query tile, 256-row
// base tiles.
template <int NQT>
__global__ void __launch_bounds__(kThreads, 1) knn_tc_filter_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                     const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    constexpr int BLOCK_N = kAccCols / NQT;
    constexpr uint32_t kBBytesPerKb = BLOCK_N * kBlockK * 2;
    constexpr uint32_t kIdesc = make_idesc(BLOCK_N);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A: NQT x num_kb x 16 KB][B ring: stages x kBBytesPerKb][bn: 2 x BLOCK_N floats][barriers][tmem ptr]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)NQT * p.num_kb * kABytesPerKb;
    float *s_bn = reinterpret_cast<float *>(sB + (size_t)p.stages * kBBytesPerKb);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_bn + 2 * BLOCK_N);
    // bars: [0] a_full, [1..1+S) full, [1+S..1+2S) empty, then tmem_full[2], tmem_empty[2]
    const uint32_t S = p.stages;
    const uint32_t bar_a = smem_u32(bars);
    const uint32_t bar_full = bar_a + 8;
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_empty + 8 * S;
    const uint32_t bar_tempty = bar_tfull + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 1 + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t q0 = (uint64_t)blockIdx.x * (kBlockM * NQT);
    const uint32_t t_begin = p.tile_lo + blockIdx.y * p.tiles_per_cta;
    const uint32_t t_end = min(p.tile_hi, t_begin + p.tiles_per_cta);
    const uint32_t n_tiles = t_end > t_begin ? t_end - t_begin : 0;

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 8);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two accumulator sets); this warp also frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // TMA producer: warp-uniform control flow, one elected lane issues
        if (n_tiles) {
            if (elect_one()) {
                // query tiles, resident for the CTA's lifetime
                mbar_expect_tx(bar_a, NQT * p.num_kb * kABytesPerKb);
                for (int j = 0; j < NQT; ++j)
                    for (uint32_t kb = 0; kb < p.num_kb; ++kb)
                        tma_load_2d(smem_u32(sA + ((size_t)j * p.num_kb + kb) * kABytesPerKb), &map_q, bar_a, (int)(kb * kBlockK),
                                    (int)(q0 + (uint64_t)j * kBlockM));
            }
            __syncwarp();
            uint32_t stage = 0, phase = 0;
            const uint32_t sB_u32 = smem_u32(sB);
            for (uint32_t t = t_begin; t < t_end; ++t) {
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + 8 * stage, kBBytesPerKb);
                        tma_load_2d(sB_u32 + stage * kBBytesPerKb, &map_b, bar_full + 8 * stage, (int)(kb * kBlockK), (int)(t * BLOCK_N));
                    }
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the whole warp walks the pipeline (uniform control flow keeps the descriptor math off the
        // per-thread path), one elected lane issues the tcgen05.mma / tcgen05.commit instructions
        if (n_tiles) {
            mbar_wait(bar_a, 0);
            tc_fence_after();
            const uint64_t d0 = umma_desc_sw128(0);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32);
            const uint32_t a_lo0 = (uint32_t)d0 | ((smem_u32(sA) & 0x3ffffu) >> 4);
            const uint32_t b_lo0 = (uint32_t)d0 | ((smem_u32(sB) & 0x3ffffu) >> 4);
            uint32_t stage = 0, phase = 0;
            for (uint32_t i = 0; i < n_tiles; ++i) {
                const uint32_t buf = i & 1;
                mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1);  // epilogue has drained this accumulator set
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kAccCols;
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t b_lo = b_lo0 + stage * (kBBytesPerKb >> 4);
                    const uint32_t nk = (kb == p.num_kb - 1) ? p.ksteps_last : (uint32_t)(kBlockK / kUmmaK);
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < NQT; ++j) {
                            const uint32_t a_lo = a_lo0 + ((uint32_t)j * p.num_kb + kb) * (kABytesPerKb >> 4);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                if ((uint32_t)k < nk)
                                    tc_mma_bf16_lo(d_tmem + j * BLOCK_N, a_lo + 2 * k, b_lo + 2 * k, desc_hi, kIdesc,
                                                   k == 0 ? (kb != 0 ? 1u : 0u) : 1u);
                        }
                        tc_commit(bar_empty + 8 * stage);  // stage reusable once these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) tc_commit(bar_tfull + 8 * buf);  // accumulator set complete
                __syncwarp();
            }
        }
    } else {
        // epilogue, 8 warps: TMEM lane quadrant = warp id % 4 (hardware rule), column half = (warp - 2) / 4.
        // A thread owns one accumulator row and 128 of the 256 columns of a set: with NQT == 2 that is exactly one query
        // (tile `half`, row `row`), whose threshold and candidate count live in registers.
        const uint32_t quad = (uint32_t)warp & 3u;
        const uint32_t half = (uint32_t)(warp - 2) >> 2;
        const uint32_t row = quad * 32 + lane;
        const int epi_tid = threadIdx.x - 64;  // 0..255
        const uint32_t jq = (half * 128u) / BLOCK_N;       // query tile of this thread's columns
        const uint32_t cbase = (half * 128u) % BLOCK_N;    // first base-tile column of this thread's columns
        const uint64_t qrow = q0 + (uint64_t)jq * kBlockM + row;
        uint32_t cnt = 0;
        // keep a score iff its exact distance can still be below tau:  |dot_bf16 - dot| <= E
        float theta = INFINITY;  // rows past the batch never pass
        if (qrow < p.nq) {
            const float t = p.tau[qrow];
            const float qq = p.qn[qrow];
            const float E = p.err_rel * sqrtf(qq) * p.bmax;
            if (p.metric_l2) theta = (qq - t) - 2.f * E;               // keep iff 2 dot - ||b||^2 > theta
            else theta = (-t) - E;                                    // keep iff dot > theta
            if (t == INFINITY) theta = -INFINITY;
        }
        const bool atomic_append = p.atomic_append || NQT == 1;  // NQT == 1: the two column halves share the query
        for (uint32_t i = 0; i < n_tiles; ++i) {
            const uint32_t buf = i & 1;
            const uint32_t tile = t_begin + i;
            const uint64_t col0 = (uint64_t)tile * BLOCK_N + cbase;
            if (p.metric_l2) {
                // ||b||^2 of this tile's rows -> shared (double-buffered with the accumulator sets)
                for (int j = epi_tid; j < BLOCK_N; j += 256) {
                    const uint64_t b = (uint64_t)tile * BLOCK_N + j;
                    s_bn[buf * BLOCK_N + j] = b < p.nb ? __ldg(p.bn + b) : 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(bar_tfull + 8 * buf, (i >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + buf * kAccCols + half * 128u;
            uint32_t r[2][32];
            __syncwarp();
            tc_ld32(taddr, r[0]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                tc_ld_wait();
                if (cc + 1 < 4) tc_ld32(taddr + (cc + 1) * 32, r[(cc + 1) & 1]);  // overlaps the scan below
                const uint32_t(&rr)[32] = r[cc & 1];
                float v[32];
                if (p.metric_l2) {
                    const float *bn = s_bn + buf * BLOCK_N + cbase + cc * 32;
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = fmaf(2.f, __uint_as_float(rr[e]), -bn[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(rr[e]);
                }
                const float m = max32(v);
                if (m > theta) {  // rare once tau has tightened
                    uint32_t mask = 0;
#pragma unroll
                    for (int e = 0; e < 32; ++e) mask |= (v[e] > theta ? 1u : 0u) << e;
                    while (mask) {
                        const int e = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint64_t b = col0 + cc * 32 + e;
                        if (b < p.nb) {
                            const uint32_t slot = atomic_append ? atomicAdd(p.cand_cnt + 2 * qrow, 1u) : cnt++;
                            if (slot < p.cap) p.cand[qrow * p.cap + slot] = (unsigned long long)(uint32_t)b;
                            else *p.overflow = 1u;
                        }
                    }
                }
                __syncwarp();  // tcgen05.ld / wait are .sync.aligned: reconverge after a divergent survivor scan
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if (!atomic_append && qrow < p.nq && cnt) p.cand_cnt[2 * qrow] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---- 2-CTA variant (the default) -----------------------------------------------------------------------------------------
// A cluster of two CTAs on an SM pair runs ONE tcgen05.mma.cta_group::2 of shape M = 256 (128 queries per CTA) x N = 256
// (base rows, 128 staged by each CTA) per k-step.  Per CTA and k-step the tensor core reads 4 KB of A and 4 KB of B from
// shared memory in 128 cycles — half the shared-memory bandwidth of the 1-CTA 128 x 128 shape, which was the limiter of
// the single-CTA kernel (tensor pipe 64 % with the MMA warp never waiting, profiles/r1_knn_tc_10M_ncu_full.txt) — while
// the L2 -> SM operand traffic stays at 128 base rows per CTA and tile.  Each CTA keeps the accumulator rows of its own
// 128 queries: 2 x 256 TMEM columns, double-buffered against the epilogue exactly as above.
//   barriers:  full[s]   leader only, 2 arrivals (the leader's expect_tx + the peer's remote arrive) + both CTAs' TMA bytes
//              empty[s]  both CTAs, arrived by the multicast tcgen05.commit of the MMAs that read the stage
//              tfull[b]  both CTAs, multicast commit after a tile's last MMA
//              tempty[b] leader only, 16 arrivals (8 epilogue warps x 2 CTAs)
constexpr int kPairN = 256;                                    // UMMA N of the pair = base rows per tile
constexpr uint32_t kPairBBytesPerKb = (kPairN / 2) * kBlockK * 2;  // each CTA stages half of the tile's rows: 16 KB

template <bool kSeed>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    knn_tc_filter_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    constexpr uint32_t kIdesc = make_idesc(kPairN, 2 * kBlockM);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve (identical offsets in both CTAs): [A: num_kb x 16 KB][B ring: stages x 16 KB][bn: 2 x 256 floats][barriers][tmem ptr]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)p.num_kb * kABytesPerKb;
    float *s_bn = reinterpret_cast<float *>(sB + (size_t)p.stages * kPairBBytesPerKb);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_bn + 2 * kPairN);
    const uint32_t S = p.stages;
    const uint32_t bar_a = smem_u32(bars);
    const uint32_t bar_full = bar_a + 8;
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_empty + 8 * S;
    const uint32_t bar_tempty = bar_tfull + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 1 + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint64_t q0 = (uint64_t)blockIdx.x * kBlockM;  // blockIdx.x = 2 * pair + rank
    const uint32_t t_begin = p.tile_lo + blockIdx.y * p.tiles_per_cta;
    const uint32_t t_end = min(p.tile_hi, t_begin + p.tiles_per_cta);
    const uint32_t n_tiles = t_end > t_begin ? t_end - t_begin : 0;

    cluster_sync_all();  // both CTAs are resident before the pair-wide TMEM allocation
    if (threadIdx.x == 0) {
        mbar_init(bar_a, 2);
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 2);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 16);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs performs the pair-wide allocation of all 512 columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // TMA producer of this CTA: its 128 query rows once, then its 128-row half of every base tile
        if (n_tiles) {
            if (elect_one()) {
                if (leader) mbar_expect_tx(bar_a, 2 * p.num_kb * kABytesPerKb);
                else mbar_arrive_remote(bar_a, 0);
                for (uint32_t kb = 0; kb < p.num_kb; ++kb)
                    tma_load_2d_pair(smem_u32(sA + (size_t)kb * kABytesPerKb), &map_q, bar_a, (int)(kb * kBlockK), (int)q0);
            }
            __syncwarp();
            uint32_t stage = 0, phase = 0;
            const uint32_t sB_u32 = smem_u32(sB);
            for (uint32_t t = t_begin; t < t_end; ++t) {
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * kPairBBytesPerKb);
                        else mbar_arrive_remote(bar_full + 8 * stage, 0);
                        tma_load_2d_pair(sB_u32 + stage * kPairBBytesPerKb, &map_b, bar_full + 8 * stage, (int)(kb * kBlockK),
                                         (int)(t * kPairN + rank * (kPairN / 2)));
                    }
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: only the leader CTA issues; the instruction reads both CTAs' shared memory at the same offsets and
        // writes both CTAs' TMEM
        if (n_tiles && leader) {
            mbar_wait(bar_a, 0);
            tc_fence_after();
            const uint64_t d0 = umma_desc_sw128(0);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32);
            const uint32_t a_lo0 = (uint32_t)d0 | ((smem_u32(sA) & 0x3ffffu) >> 4);
            const uint32_t b_lo0 = (uint32_t)d0 | ((smem_u32(sB) & 0x3ffffu) >> 4);
            uint32_t stage = 0, phase = 0;
            for (uint32_t i = 0; i < n_tiles; ++i) {
                const uint32_t buf = i & 1;
                mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kAccCols;
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + kb * (kABytesPerKb >> 4);
                    const uint32_t b_lo = b_lo0 + stage * (kPairBBytesPerKb >> 4);
                    const uint32_t nk = (kb == p.num_kb - 1) ? p.ksteps_last : (uint32_t)(kBlockK / kUmmaK);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)
                            if ((uint32_t)k < nk)
                                tc_mma_bf16_pair(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, kIdesc, k == 0 ? (kb != 0 ? 1u : 0u) : 1u);
                        tc_commit_pair(bar_empty + 8 * stage);  // frees the stage in both CTAs
                    }
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) tc_commit_pair(bar_tfull + 8 * buf);  // accumulator complete in both CTAs
                __syncwarp();
            }
            // the peer's last remote arrivals must have landed before this CTA's barriers go away
            for (uint32_t i = n_tiles > 2 ? n_tiles - 2 : 0; i < n_tiles; ++i) mbar_wait(bar_tempty + 8 * (i & 1), (i >> 1) & 1);
        }
    } else {
        // epilogue, 8 warps per CTA: TMEM lane quadrant = warp id % 4, column half = (warp - 2) / 4.  A thread owns one
        // query row and 128 of the tile's 256 columns; the two halves of a row append through the global counter.
        const uint32_t quad = (uint32_t)warp & 3u;
        const uint32_t half = (uint32_t)(warp - 2) >> 2;
        const uint32_t row = quad * 32 + lane;
        const int epi_tid = threadIdx.x - 64;  // 0..255
        const uint64_t qrow = q0 + row;
        const uint32_t sub_cap = p.cap >> 1;
        unsigned long long *my_cand = p.cand + qrow * p.cap + half * sub_cap;
        uint32_t cnt = 0;
        float theta = INFINITY;  // rows past the batch never pass
        if (qrow < p.nq) {
            const float t = p.tau[qrow];
            const float qq = p.qn[qrow];
            const float E = p.err_rel * sqrtf(qq) * p.bmax;
            if (p.metric_l2) theta = (qq - t) - 2.f * E;  // keep iff 2 dot - ||b||^2 > theta
            else theta = (-t) - E;                        // keep iff dot > theta
            if (t == INFINITY) theta = -INFINITY;
        }
        for (uint32_t i = 0; i < n_tiles; ++i) {
            const uint32_t buf = i & 1;
            const uint32_t tile = t_begin + i;
            const uint64_t col0 = (uint64_t)tile * kPairN + half * 128u;
            if (p.metric_l2) {
                for (int j = epi_tid; j < kPairN; j += 256) {
                    const uint64_t b = (uint64_t)tile * kPairN + j;
                    s_bn[buf * kPairN + j] = b < p.nb ? __ldg(p.bn + b) : 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(bar_tfull + 8 * buf, (i >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + buf * kAccCols + half * 128u;
            uint32_t r[2][32];
            float tile_max = -INFINITY;  // seed pass only
            __syncwarp();
            tc_ld32(taddr, r[0]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                tc_ld_wait();
                if (cc + 1 < 4) tc_ld32(taddr + (cc + 1) * 32, r[(cc + 1) & 1]);  // overlaps the scan below
                const uint32_t(&rr)[32] = r[cc & 1];
                float v[32];
                if (p.metric_l2) {
                    const float4 *bn4 = reinterpret_cast<const float4 *>(s_bn + buf * kPairN + half * 128u + cc * 32);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 b4 = bn4[e];
                        v[4 * e + 0] = fmaf(2.f, __uint_as_float(rr[4 * e + 0]), -b4.x);
                        v[4 * e + 1] = fmaf(2.f, __uint_as_float(rr[4 * e + 1]), -b4.y);
                        v[4 * e + 2] = fmaf(2.f, __uint_as_float(rr[4 * e + 2]), -b4.z);
                        v[4 * e + 3] = fmaf(2.f, __uint_as_float(rr[4 * e + 3]), -b4.w);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(rr[e]);
                }
                float gmax[4];
                if (kSeed) {
                    tile_max = fmaxf(tile_max, max32(v, gmax));
                } else if (max32(v, gmax) > theta) {  // rare once tau has tightened
                    uint32_t mask = survivor_mask(v, gmax, theta);
                    while (mask) {
                        const int e = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint64_t b = col0 + cc * 32 + e;
                        if (b < p.nb) {
                            if (p.atomic_append) {  // few queries: the base range is split over blockIdx.y, lists are shared
                                const uint32_t slot = atomicAdd(p.cand_cnt + 2 * qrow, 1u);
                                if (slot < p.cap) p.cand[qrow * p.cap + slot] = (unsigned long long)(uint32_t)b;
                                else *p.overflow = 1u;
                            } else {  // private sub-list of this (query, column half): no atomic, no round trip
                                if (cnt < sub_cap) my_cand[cnt] = (unsigned long long)(uint32_t)b;
                                else *p.overflow = 1u;
                                ++cnt;
                            }
                        }
                    }
                }
                __syncwarp();  // tcgen05.ld / wait are .sync.aligned: reconverge after a divergent survivor scan
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
            // seed pass: one value per (tile, column half, query); consecutive lanes = consecutive queries -> 128-byte stores
            if (kSeed && qrow < p.nq) p.seed_out[(uint64_t)((tile - p.tile_lo) * 2u + half) * p.seed_stride + qrow] = tile_max;
        }
        if (!kSeed && !p.atomic_append && qrow < p.nq && cnt) p.cand_cnt[2 * qrow + half] = cnt;
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// fp32 rows -> bf16 rows of pitch kp (zero padded) + exact fp32 squared norms; one warp per row
__global__ void to_bf16_kernel(const float *x, uint64_t n, uint32_t dim, uint32_t kp, __nv_bfloat16 *out, float *sq) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float *r = x + row * dim;
    __nv_bfloat16 *o = out + row * kp;
    float s = 0.f;
    for (uint32_t j = lane; j < kp; j += 32) {
        const float v = j < dim ? r[j] : 0.f;
        s = fmaf(v, v, s);
        o[j] = __float2bfloat16_rn(v);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0 && sq) sq[row] = s;
}

__global__ void max_norm_kernel(const float *sq, uint64_t n, float *out) {
    float m = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = fmaxf(m, sq[i]);
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));  // non-negative floats order as ints
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap *m, const void *ptr, uint64_t rows, uint32_t kp, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RG_ECUDA; }
    const cuuint64_t dims[2] = {kp, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        set_error(buf);
        return RG_ECUDA;
    }
    return RG_OK;
}

}  // namespace

bool knn_tc_supported(uint32_t dim) { return dim >= 8 && ((dim + 7) / 8 * 8) <= (uint32_t)(kMaxKb * kBlockK) && encode_fn() != nullptr; }

uint32_t knn_tc_pitch(uint32_t dim) { return (dim + 7) / 8 * 8; }

int knn_tc_convert(const float *x, uint64_t n, uint32_t dim, void *out_bf16, float *sq, cudaStream_t s) {
    to_bf16_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(x, n, dim, knn_tc_pitch(dim), reinterpret_cast<__nv_bfloat16 *>(out_bf16), sq);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int knn_tc_max_norm(const float *sq, uint64_t n, float *d_out, cudaStream_t s) {
    RG_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), s));
    max_norm_kernel<<<296, 256, 0, s>>>(sq, n, d_out);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

// One chunk [row_lo, row_hi) of the base (both multiples of 256, or row_hi == nb) against one query batch.
template <int NQT>
static int launch_filter(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t kp, uint64_t row_lo,
                         uint64_t row_hi, TcParams p, int sm_count, cudaStream_t s) {
    constexpr int BLOCK_N = kAccCols / NQT;
    constexpr uint32_t kBBytesPerKb = BLOCK_N * kBlockK * 2;
    CUtensorMap map_q, map_b;
    int rc = make_map(&map_q, q_bf16, nq, kp, kBlockM);
    if (rc != RG_OK) return rc;
    rc = make_map(&map_b, b_bf16, nb, kp, BLOCK_N);
    if (rc != RG_OK) return rc;
    const size_t fixed = (size_t)NQT * p.num_kb * kABytesPerKb + 2 * BLOCK_N * 4 + 512 + 1024;
    const size_t budget = 227 * 1024;
    if (fixed + 2 * kBBytesPerKb > budget) { set_error("rg_knn: dimension too large for the tensor-core path"); return RG_EINVAL; }
    uint32_t stages = (uint32_t)((budget - fixed) / kBBytesPerKb);
    if (stages > 8) stages = 8;
    p.stages = stages;
    p.tile_lo = (uint32_t)(row_lo / BLOCK_N);
    p.tile_hi = (uint32_t)((row_hi + BLOCK_N - 1) / BLOCK_N);
    const uint32_t tiles = p.tile_hi - p.tile_lo;
    const uint32_t q_tiles = (uint32_t)((nq + kBlockM * NQT - 1) / (kBlockM * NQT));
    uint32_t splits = 1;
    if (q_tiles < 2u * (uint32_t)sm_count) splits = std::min<uint32_t>(tiles, (2u * (uint32_t)sm_count + q_tiles - 1) / q_tiles);
    p.tiles_per_cta = (tiles + splits - 1) / splits;
    splits = (tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    p.atomic_append = splits > 1 ? 1 : 0;
    const size_t smem = fixed + (size_t)stages * kBBytesPerKb;
    RG_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<NQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(q_tiles, splits);
    knn_tc_filter_kernel<NQT><<<grid, kThreads, smem, s>>>(map_q, map_b, p);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}


// 2-CTA launch: grid.x = 2 CTAs per 256 queries (one cluster), grid.y = base-range splits when there are few queries.
static int launch_filter_pair(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t kp, uint64_t row_lo,
                              uint64_t row_hi, TcParams p, int sm_count, cudaStream_t s) {
    CUtensorMap map_q, map_b;
    int rc = make_map(&map_q, q_bf16, nq, kp, kBlockM);
    if (rc != RG_OK) return rc;
    rc = make_map(&map_b, b_bf16, nb, kp, kPairN / 2);
    if (rc != RG_OK) return rc;
    const size_t fixed = (size_t)p.num_kb * kABytesPerKb + 2 * kPairN * 4 + 512 + 1024;
    size_t budget = 227 * 1024;
    if (p.smem_reserve && fixed + 4 * kPairBBytesPerKb + p.smem_reserve <= budget) budget -= p.smem_reserve;  // never below 4 stages
    uint32_t stages = (uint32_t)((budget - fixed) / kPairBBytesPerKb);
    if (stages > 12) stages = 12;
    p.stages = stages;
    p.tile_lo = (uint32_t)(row_lo / kPairN);
    p.tile_hi = (uint32_t)((row_hi + kPairN - 1) / kPairN);
    const uint32_t tiles = p.tile_hi - p.tile_lo;
    const uint32_t q_ctas = 2u * (uint32_t)((nq + 2 * kBlockM - 1) / (2 * kBlockM));
    uint32_t splits = 1;
    if (q_ctas < 2u * (uint32_t)sm_count) splits = std::min<uint32_t>(tiles, (2u * (uint32_t)sm_count + q_ctas - 1) / q_ctas);
    p.tiles_per_cta = (tiles + splits - 1) / splits;
    splits = (tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    p.atomic_append = splits > 1 ? 1 : 0;
    const size_t smem = fixed + (size_t)stages * kPairBBytesPerKb;
    dim3 grid(q_ctas, splits);
    if (p.seed_out) {
        RG_CUDA(cudaFuncSetAttribute(knn_tc_filter_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_tc_filter_pair_kernel<true><<<grid, kThreads, smem, s>>>(map_q, map_b, p);
    } else {
        RG_CUDA(cudaFuncSetAttribute(knn_tc_filter_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_tc_filter_pair_kernel<false><<<grid, kThreads, smem, s>>>(map_q, map_b, p);
    }
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

// |dot_bf16 - dot| <= kErrRel * ||q|| * ||b||: (2u + u^2) with u = 2^-8 (bf16 round-to-nearest), plus slack for the fp32 accumulation
constexpr float kErrRel = 1.01f * 0.0078125f;

namespace {
__device__ __forceinline__ uint32_t seed_f2ord(float f) {  // order-preserving float -> uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float seed_ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// tau seed: the K-th largest of a query's G group maxima (each the bf16 score of a DISTINCT base row) bounds the K-th best
// distance over the seed rows from above once the bf16 error bound is added.  One CTA = 32 queries, staged through shared
// memory ([g][query] in global memory: the filter kernel's stores are coalesced over queries); one warp = one query at a
// time, 32-step radix select over the order-preserving keys (G <= 1024 values, up to 32 per lane).
constexpr int kSeedMaxPerLane = 32;
__global__ void __launch_bounds__(256) knn_seed_select_kernel(const float *__restrict__ seed, uint64_t seed_stride, uint32_t G,
                                                              uint64_t nq, uint32_t K, int metric_l2, const float *__restrict__ qn,
                                                              float bmax, float *__restrict__ tau) {
    extern __shared__ float s_tile[];  // G x 33
    const uint64_t q0 = (uint64_t)blockIdx.x * 32u;
    for (uint32_t i = threadIdx.x; i < G * 32u; i += blockDim.x) {
        const uint32_t g = i >> 5, c = i & 31u;
        s_tile[g * 33u + c] = (q0 + c < nq) ? seed[(uint64_t)g * seed_stride + q0 + c] : -INFINITY;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t c = warp; c < 32u; c += blockDim.x >> 5) {
        const uint64_t q = q0 + c;
        if (q >= nq) break;
        uint32_t key[kSeedMaxPerLane];
#pragma unroll
        for (int j = 0; j < kSeedMaxPerLane; ++j) {
            const uint32_t g = (uint32_t)lane + 32u * j;
            key[j] = g < G ? seed_f2ord(s_tile[g * 33u + c]) : 0u;  // 0 = below every real key
        }
        uint32_t prefix = 0, remaining = K;
        for (int b = 31; b >= 0; --b) {
            const uint32_t want = prefix | (1u << b), mask = ~((1u << b) - 1u);
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < kSeedMaxPerLane; ++j) cnt += ((key[j] & mask) == want) ? 1u : 0u;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt >= remaining) prefix = want;
            else remaining -= cnt;
        }
        if (lane == 0) {
            const float vk = seed_ord2f(prefix);  // K-th largest bf16 score (IP: dot; L2: 2 dot - ||b||^2)
            const float qq = qn[q];
            const float E = kErrRel * sqrtf(qq) * bmax;
            // K distinct rows have a TRUE internal distance <= this value; 1e-30: a zero query (E = 0, all scores equal) keeps
            // every row instead of none
            float t = metric_l2 ? (qq - vk) + 2.f * E : (-vk) + E;
            t += fmaxf(fabsf(t) * 1e-6f, 1e-30f);
            tau[q] = t;
        }
    }
}
}  // namespace

// Seed pass over base rows [0, n_seed) (a multiple of 256): tau[q] <- an upper bound of the K-th best distance of query q
// over those rows, from the tensor-core scores alone (no candidate lists, no re-rank).  seed_buf: 2 * n_seed / 256 * nq floats.
int knn_tc_seed(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t dim, uint64_t n_seed, uint32_t K,
                int metric_l2, const float *qn, const float *bn, float bmax, float *seed_buf, float *tau, int sm_count,
                cudaStream_t s, uint32_t smem_reserve) {
    const uint32_t kp = knn_tc_pitch(dim);
    const uint32_t G = (uint32_t)(n_seed / 128u);
    if (n_seed == 0 || n_seed % 256u != 0 || n_seed > nb || G < K || G > 32u * kSeedMaxPerLane) {
        set_error("rg_knn: bad seed range (internal error)");
        return RG_EINVAL;
    }
    TcParams p{};
    p.nq = nq;
    p.nb = nb;
    p.num_kb = (kp + kBlockK - 1) / kBlockK;
    p.ksteps_last = ((kp - (p.num_kb - 1) * kBlockK) + kUmmaK - 1) / kUmmaK;
    p.metric_l2 = metric_l2;
    p.tau = tau;  // read but unused by the seed kernel
    p.qn = qn;
    p.bn = bn;
    p.bmax = bmax;
    p.err_rel = kErrRel;
    p.seed_out = seed_buf;
    p.seed_stride = nq;
    p.smem_reserve = smem_reserve;
    int rc = launch_filter_pair(q_bf16, nq, b_bf16, nb, kp, 0, n_seed, p, sm_count, s);
    if (rc != RG_OK) return rc;
    const size_t smem = (size_t)G * 33u * sizeof(float);
    RG_CUDA(cudaFuncSetAttribute(knn_seed_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_seed_select_kernel<<<(unsigned)((nq + 31) / 32), 256, smem, s>>>(seed_buf, nq, G, nq, K, metric_l2, qn, bmax, tau);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

// rows of the seed pass for a base of nb rows and top-K (0: too small to seed — the chunk ramp starts from tau = inf)
uint64_t knn_tc_seed_rows(uint64_t nb, uint32_t K) {
    const char *force = getenv("RG_KNN_QTILES");
    if (force && (force[0] == '1' || force[0] == '2')) return 0;  // only the 2-CTA kernel has the seed epilogue
    if (const char *v = getenv("RG_KNN_SEED")) {
        if (v[0] == '0') return 0;
    }
    uint64_t n = std::min<uint64_t>(65536u, (nb / 4u) / 256u * 256u);
    if (n / 128u < 2ull * K) return 0;  // the K-th of fewer than 2K group maxima is a poor bound
    return n;
}

int knn_tc_filter(const void *q_bf16, uint64_t nq, const void *b_bf16, uint64_t nb, uint32_t dim, uint64_t row_lo, uint64_t row_hi,
                  int metric_l2, const float *tau, const float *qn, const float *bn, float bmax, unsigned long long *cand,
                  uint32_t *cand_cnt, uint32_t cap, uint32_t *overflow, int sm_count, cudaStream_t s, uint32_t smem_reserve) {
    const uint32_t kp = knn_tc_pitch(dim);
    TcParams p{};
    p.smem_reserve = smem_reserve;
    p.nq = nq;
    p.nb = nb;
    p.num_kb = (kp + kBlockK - 1) / kBlockK;
    p.ksteps_last = ((kp - (p.num_kb - 1) * kBlockK) + kUmmaK - 1) / kUmmaK;
    p.metric_l2 = metric_l2;
    p.tau = tau;
    p.qn = qn;
    p.bn = bn;
    p.bmax = bmax;
    p.cand = cand;
    p.cand_cnt = cand_cnt;
    p.cap = cap;
    p.overflow = overflow;
    p.err_rel = kErrRel;
    // RG_KNN_QTILES=1|2 selects the single-CTA kernels (one or two resident query tiles); default: the 2-CTA kernel
    const char *force = getenv("RG_KNN_QTILES");
    if (!force || (force[0] != '1' && force[0] != '2'))
        return launch_filter_pair(q_bf16, nq, b_bf16, nb, kp, row_lo, row_hi, p, sm_count, s);
    const bool two = force[0] == '2';
    if (two && p.num_kb <= 4) return launch_filter<2>(q_bf16, nq, b_bf16, nb, kp, row_lo, row_hi, p, sm_count, s);
    return launch_filter<1>(q_bf16, nq, b_bf16, nb, kp, row_lo, row_hi, p, sm_count, s);
}

}  // namespace rg
