// ubench_gather.cu — how many DRAM bytes does a random 800-byte (or ROWB-byte) row gather really cost, per copy mechanism?
//   mode 0: cp.async.bulk (TMA 1-D bulk copy, what the search kernel uses), 8 rows per warp pass, mbarrier completion
//   mode 1: LDG.128 into registers (4 lanes per row)
//   mode 2: cp.async.cg 16 B (LDGSTS) into shared memory, 4 lanes per row
// Run under ncu for dram__bytes_read.sum; prints GB/s of useful bytes from CUDA events otherwise.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_gather scripts/ubench_gather.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nWLOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WDONE;\nbra WLOOP;\nWDONE:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 6) gather(const float *base, uint32_t row_floats, const uint32_t *ids, uint32_t n_ids, float *out,
                                                 uint8_t *tags, uint64_t tag_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rstride = row_floats + 16;  // floats
    float *rows = reinterpret_cast<float *>(smem) + (size_t)warp * (8 * rstride + 16);
    uint64_t *bar = reinterpret_cast<uint64_t *>(rows + 8 * rstride);
    const uint32_t bar_u = smem_u32(bar);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint32_t row_bytes = row_floats * 4;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    float acc = 0.f;
    uint32_t phase = 0;
    for (uint32_t i0 = gw * 8; i0 + 8 <= n_ids; i0 += nwarps * 8) {
        if (MODE == 3) {
            // the search kernel's visited-tag traffic: ~10 random 1-byte reads and 8 random 1-byte writes per 8 rows
            const uint64_t h = ((uint64_t)ids[i0 + (lane & 7)] * 0x9E3779B97F4A7C15ull + lane * 0xBF58476D1CE4E5B9ull) % tag_bytes;
            uint8_t t = 0;
            if (lane < 10) t = __ldcg(tags + h);
            if (lane < 8 && t != 77) __stcg(tags + h, (uint8_t)77);
        }
        if (MODE == 0 || MODE == 3) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u), "r"(8 * row_bytes) : "memory");
            __syncwarp();
            if (lane < 8) {
                const uint32_t id = ids[i0 + lane];
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(rows + lane * rstride)),
                             "l"(base + (size_t)id * row_floats), "r"(row_bytes), "r"(bar_u)
                             : "memory");
            }
            mbar_wait(bar_u, phase);
            phase ^= 1;
            acc += rows[(lane >> 2) * rstride + (lane & 3)];
            __syncwarp();
        } else if (MODE == 1) {
            const uint32_t id = ids[i0 + (lane >> 2)];
            const float4 *r = reinterpret_cast<const float4 *>(base + (size_t)id * row_floats) + (lane & 3);
            const uint32_t n4 = row_floats / 4;
#pragma unroll 13
            for (uint32_t c = 0; c * 4 + (lane & 3) < n4; ++c) {
                const float4 v = __ldg(r + 4 * c);
                acc += v.x + v.y + v.z + v.w;
            }
        } else {
            const uint32_t id = ids[i0 + (lane >> 2)];
            const float4 *r = reinterpret_cast<const float4 *>(base + (size_t)id * row_floats) + (lane & 3);
            float4 *d = reinterpret_cast<float4 *>(rows + (lane >> 2) * rstride) + (lane & 3);
            const uint32_t n4 = row_floats / 4;
            for (uint32_t c = 0; c * 4 + (lane & 3) < n4; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + 4 * c)), "l"(r + 4 * c) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            acc += rows[(lane >> 2) * rstride + (lane & 3)];
            __syncwarp();
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char **argv) {
    const uint32_t row_floats = argc > 1 ? atoi(argv[1]) : 200;
    const uint64_t nrows = 10000000;
    const uint32_t n_ids = 1u << 24;
    float *base, *out;
    uint32_t *ids;
    cudaMalloc(&base, nrows * row_floats * 4);
    cudaMemset(base, 0, nrows * row_floats * 4);
    cudaMalloc(&out, 4);
    cudaMalloc(&ids, n_ids * 4);
    std::vector<uint32_t> h(n_ids);
    uint64_t x = 88172645463325252ull;
    for (auto &v : h) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        v = (uint32_t)(x % nrows);
    }
    cudaMemcpy(ids, h.data(), n_ids * 4, cudaMemcpyHostToDevice);
    const int smem = 4 * (8 * (row_floats + 16) + 16) * 4;
    cudaFuncSetAttribute(gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gather<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    uint8_t *tags;
    const uint64_t tag_bytes = 20ull << 30;
    cudaMalloc(&tags, tag_bytes);
    cudaMemset(tags, 0, tag_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[4] = {"cp.async.bulk", "ldg.128", "cp.async.cg 16B", "bulk + tag bytes"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) gather<0><<<148 * 6, 128, smem>>>(base, row_floats, ids, n_ids, out, tags, tag_bytes);
            if (mode == 1) gather<1><<<148 * 6, 128, smem>>>(base, row_floats, ids, n_ids, out, tags, tag_bytes);
            if (mode == 3) gather<3><<<148 * 6, 128, smem>>>(base, row_floats, ids, n_ids, out, tags, tag_bytes);
            if (mode == 2) gather<2><<<148 * 6, 128, smem>>>(base, row_floats, ids, n_ids, out, tags, tag_bytes);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", names[mode], cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1) printf("%-18s row %u B: %.3f ms, %.1f GB/s useful\n", names[mode], row_floats * 4, ms, (double)n_ids * row_floats * 4 / ms / 1e6);
        }
    }
    return 0;
}
