// probe_mma_rate.cu -- hardware probe (GPU box): how many cycles does one tcgen05.mma (kind::tf32, cta_group::1, K = 32 bytes)
// take as a function of its shape and of where its operands sit in shared memory?  The block-0 / row-pair kernels issue many
// SMALL MMAs (M = 128, N = 48..192) on no-swizzle or narrow operands; the math floor (M*N/256 cycles) says 24 cycles for
// N = 48, the kernels measure ~100.  This separates: fixed per-instruction cost, A-operand fetch, B-operand fetch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mma_rate tools/probe_mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../dnn-inference-engine_b200/csrc/ptx.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Cfg {
    int m, n;
    uint32_t a_lbo, a_sbo, a_layout, a_step;   // a_step: bytes the A start advances from one MMA of a group to the next
    uint32_t b_lbo, b_sbo, b_layout, b_step;
    int group, same_acc;                       // MMAs per group; all into one accumulator or alternating two
    int issuers;                               // warps that issue concurrently, each its own stream of MMAs (own accumulator columns)
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(256, 1) rate_kernel(Cfg c, int iters, long long *cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    __shared__ long long t_issuer[8];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    const uint32_t b = ptx::smem_u32(&bars[warp]);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if ((threadIdx.x & 31) == 0 && warp < c.issuers) {
        const uint32_t sa = ptx::smem_u32(smem) + warp * 8192, sb = ptx::smem_u32(smem) + 96 * 1024 + warp * 2048;
        const uint32_t tmem_w = tmem + warp * (512 / (c.issuers > 4 ? 8 : c.issuers > 2 ? 4 : c.issuers));
        const uint32_t idesc = ptx::umma_idesc(2u, c.n, c.m);
        uint32_t phase = 0;
        // warm up
        ptx::umma<false>(tmem_w, mk_desc(sa, c.a_lbo, c.a_sbo, c.a_layout), mk_desc(sb, c.b_lbo, c.b_sbo, c.b_layout), idesc, 0u);
        ptx::umma_commit(b); ptx::mbar_wait(b, phase); phase ^= 1;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int j = 0; j < c.group; ++j)
                ptx::umma<false>(tmem_w + (c.same_acc ? 0 : (j & 1) * 64), mk_desc(sa + j * c.a_step, c.a_lbo, c.a_sbo, c.a_layout),
                                 mk_desc(sb + j * c.b_step, c.b_lbo, c.b_sbo, c.b_layout), idesc, 1u);
            if ((it & 7) == 7 || it == iters - 1) { ptx::umma_commit(b); ptx::mbar_wait(b, phase); phase ^= 1; }
        }
        t_issuer[warp] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long mx = 0;
        for (int i = 0; i < c.issuers; ++i) mx = t_issuer[i] > mx ? t_issuer[i] : mx;
        cycles[blockIdx.x] = mx;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int smem = 160 * 1024 + 1024;
    CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long *d;
    CHECK(cudaMalloc(&d, sms * sizeof(long long)));
    std::vector<long long> h(sms);
    struct Named { const char *name; Cfg c; };
    const uint32_t SW128 = 2, NONE = 0;
    std::vector<Named> cfgs;
    // standard SW128 operands (128-byte rows; one MMA = 32 bytes of K): N sweep at M = 128
    for (int n : {16, 32, 48, 64, 96, 128, 192, 256})
        cfgs.push_back({"M128 SW128 A, SW128 B", {128, n, 16, 1024, SW128, 32, 16, 1024, SW128, 32, 4, 1, 1}});
    for (int n : {64, 128, 256})
        cfgs.push_back({"M64  SW128 A, SW128 B", {64, n, 16, 1024, SW128, 32, 16, 1024, SW128, 32, 4, 1, 1}});
    // block 0, first generation: A = pixel patch, no swizzle, overlapping K columns (LBO 16); B = 48 stacked weight rows
    for (int n : {16, 32, 48})
        cfgs.push_back({"M128 noswz A lbo16, noswz B lbo768", {128, n, 16, 128, NONE, 2176, 768, 128, NONE, 0, 4, 1, 1}});
    // block 0, second generation: K columns from two planes (LBO 2240)
    cfgs.push_back({"M128 noswz A lbo2240, noswz B lbo768", {128, 48, 2240, 128, NONE, 4480, 768, 128, NONE, 0, 4, 1, 1}});
    cfgs.push_back({"M128 noswz A lbo2240, alternating acc", {128, 48, 2240, 128, NONE, 4480, 768, 128, NONE, 0, 4, 0, 1}});
    // swapped roles: weights as A (M = 64: 48 rows used), pixels as B (N = 128 / 256), no swizzle
    for (int n : {128, 256})
        cfgs.push_back({"M64  noswz A(weights) lbo1024, noswz B(pixels) lbo16", {64, n, 1024, 128, NONE, 0, 16, 128, NONE, 2176, 4, 1, 1}});
    for (int n : {128, 256})
        cfgs.push_back({"M128 noswz A(weights) lbo2048, noswz B(pixels) lbo16", {128, n, 2048, 128, NONE, 0, 16, 128, NONE, 2176, 4, 1, 1}});
    // row-pair kernels: SW64 / SW128 patches, N = 96 / 192 stacked weights
    cfgs.push_back({"M128 SW64 A (block 1), SW64 B N96", {128, 96, 16, 512, 4, 64, 16, 512, 4, 0, 6, 1, 1}});
    cfgs.push_back({"M128 SW128 A (block 2), SW128 B N192", {128, 192, 16, 1024, SW128, 128, 16, 1024, SW128, 0, 6, 1, 1}});
    for (int iss : {1, 2, 4, 8})
        for (int n : {48, 128})
            cfgs.push_back({iss == 1 ? "1 issuer  M128 noswz A lbo2240" : iss == 2 ? "2 issuers M128 noswz A lbo2240" : iss == 4 ? "4 issuers M128 noswz A lbo2240" : "8 issuers M128 noswz A lbo2240",
                            {128, n > 64 ? 64 : n, 2240, 128, NONE, 4480, 768, 128, NONE, 0, 4, 1, iss}});
    for (int iss : {1, 2, 4})
        cfgs.push_back({iss == 1 ? "1 issuer  M128 SW128 N256" : iss == 2 ? "2 issuers M128 SW128 N128" : "4 issuers M128 SW128 N128",
                        {128, iss == 1 ? 256 : 128, 16, 1024, SW128, 32, 16, 1024, SW128, 32, 4, 1, iss}});
    printf("%-58s %4s %4s %10s %10s %12s\n", "operands", "M", "N", "cyc/MMA", "floor", "MAC/cyc/SM");
    for (const Named &nc : cfgs) {
        const int iters = 512;
        for (int grid : {1, sms}) {
            rate_kernel<<<grid, 256, smem>>>(nc.c, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", nc.name, cudaGetErrorString(e)); return 3; }
            CHECK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            const double per = (double)mx / ((double)iters * nc.c.group * nc.c.issuers);
            if (grid == sms)
                printf("%-58s %4d %4d %10.1f %10.1f %12.0f\n", nc.name, nc.c.m, nc.c.n, per, (nc.c.m < 128 ? 128 : nc.c.m) * nc.c.n / 256.0,
                       (double)nc.c.m * nc.c.n * 8 / per);
        }
    }
    return 0;
}
