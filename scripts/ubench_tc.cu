// ubench_tc.cu — microbenchmarks behind the kNN filter kernel's design (run on the GPU box):
//   1. tcgen05.mma SS-mode cadence (cycles per 128 x N x 16 bf16 MMA, operands in shared memory) for N = 128 / 256
//   2. tcgen05.ld 32x32b.x32 throughput with 4 and 8 warps
//   3. both at once (does the TMEM read port slow the MMA down?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tc scripts/ubench_tc.cu ; run: /tmp/ubench_tc
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
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
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n, uint32_t m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// mode bit 0: run MMAs (warp 0), bit 1: run LDTM in `ld_warps` warps (warps 4..4+ld_warps)
__global__ void __launch_bounds__(384, 1) ubench(int mode, int n_mma, uint32_t N, int ld_warps, int n_ld, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    long long t0 = 0, t1 = 0;
    if (warp == 0 && (mode & 1)) {
        if (lane == 0) {
            const uint64_t a = desc_sw128(smem_u32(smem)), b = desc_sw128(smem_u32(smem) + 32 * 1024);
            const uint32_t idesc = make_idesc(N, 128);
            t0 = clock64();
            for (int i = 0; i < n_mma; ++i) tc_mma(tmem + (i & 1) * 256, a + (uint64_t)((i & 3) * 2), b + (uint64_t)((i & 3) * 2), idesc, 1);
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            t1 = clock64();
            out[0] = t1 - t0;
        }
    } else if (warp >= 4 && warp < 4 + ld_warps && (mode & 2)) {
        uint32_t r[32];
        uint32_t sink = 0;
        const uint32_t taddr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + ((warp - 4) >> 2) * 128;
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < n_ld; ++i) {
            tc_ld32(taddr + (i & 3) * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sink ^= r[i & 31];
        }
        t1 = clock64();
        if (lane == 0) out[1 + (warp - 4)] = t1 - t0;
        if (sink == 0x12345678u) out[20] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d_out, h[32];
    cudaMalloc(&d_out, sizeof(h));
    const int smem = 100 * 1024;
    cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_mma = 4096, n_ld = 4096;
    struct Cfg { int mode; uint32_t N; int ldw; const char *name; } cfgs[] = {
        {1, 128, 0, "mma N=128 alone"}, {1, 256, 0, "mma N=256 alone"}, {1, 64, 0, "mma N=64 alone"},
        {2, 128, 4, "ld x32, 4 warps alone"}, {2, 128, 8, "ld x32, 8 warps alone"},
        {3, 128, 8, "mma N=128 + 8 ld warps"}, {3, 256, 8, "mma N=256 + 8 ld warps"}, {3, 256, 4, "mma N=256 + 4 ld warps"},
    };
    for (auto &c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_out, 0, sizeof(h));
            ubench<<<1, 384, smem>>>(c.mode, n_mma, c.N, c.ldw, n_ld, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-28s", c.name);
        if (c.mode & 1) printf("  mma: %.1f cyc/MMA", (double)h[0] / n_mma);
        if (c.mode & 2) {
            double mx = 0;
            for (int w = 0; w < c.ldw; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
            printf("  ld: %.1f cyc per x32 load per warp (%.1f B/clk/SM)", mx / n_ld, c.ldw * 4096.0 * n_ld / mx);
        }
        printf("\n");
    }
    return 0;
}
