// probe_desc.cu -- hardware probe (GPU box) for three UMMA shared-memory descriptor questions the
// next kernels depend on.  For a TMA-written, swizzled A slab of 160 rows:
//   Q1  SW128: can the descriptor start at an arbitrary row s (start += s*128 B), and does the
//       descriptor's base_offset field have to be 0 or (start >> 7) & 7 ?
//   Q2  SW64 : same with 64-byte rows.
//   Q3  no-swizzle K-major with OVERLAPPING core matrices: rows 16 B apart, LBO = 16 B, i.e. the K
//       slice of row m is bytes [16 m, 16 m + 32) of one linear buffer (a pixel-major patch).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_desc tools/probe_desc.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../dnn-inference-engine_b200/csrc/ptx.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ROWS = 160;   // slab rows
constexpr int N = 16;

// mode 0: SW128 (32 floats / row)   mode 1: SW64 (16 floats / row)   mode 2: no swizzle, 4 floats / row, overlapping
struct Args {
    int mode, shift, base_offset_mode;   // base_offset_mode: 0 -> 0, 1 -> (start >> 7) & 7
    float *out;                          // [128][16]
    const float *a_lin;                  // mode 2: linear slab [ROWS][4]
    const float *b;                      // [N][Kmma] row-major, Kmma = 8
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a, Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem;                 // up to 160*128 = 20480 B
    uint8_t *sB = smem + 24576;         // B: no-swizzle canonical [k-core 2][n-group 2][8][4 floats]
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    const uint32_t bt = ptx::smem_u32(&bar_tma), bm = ptx::smem_u32(&bar_mma);
    if (tid == 0) { ptx::mbar_init(bt, 1); ptx::mbar_init(bm, 1); ptx::fence_barrier_init(); }
    if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(&slot), 32);
    // B: element (n, k) at (k/4)*256 + (n/8)*128 + (n%8)*16 + (k%4)*4   => LBO = 256, SBO = 128
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        *reinterpret_cast<float *>(sB + (k / 4) * 256 + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) = a.b[i];
    }
    if (a.mode == 2)
        for (int i = tid; i < ROWS * 4; i += 128) reinterpret_cast<float *>(sA)[i] = a.a_lin[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const int rowbytes = a.mode == 0 ? 128 : 64;
        if (a.mode != 2) {
            ptx::mbar_expect_tx(bt, ROWS * rowbytes);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(ptx::smem_u32(sA)), "l"(&map_a), "r"(bt), "r"(0), "r"(0) : "memory");
            ptx::mbar_wait(bt, 0);
        }
        ptx::tc_fence_after();
        uint64_t da, db;
        // B descriptor: no swizzle, LBO 256, SBO 128
        db = (uint64_t)((ptx::smem_u32(sB) & 0x3FFFF) >> 4) | ((uint64_t)(256 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        if (a.mode == 2) {
            const uint32_t start = ptx::smem_u32(sA) + a.shift * 16;
            da = (uint64_t)((start & 0x3FFFF) >> 4) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        } else {
            const uint32_t start = ptx::smem_u32(sA) + a.shift * rowbytes;
            da = ptx::umma_desc_kmajor(start, rowbytes);
            if (a.base_offset_mode == 1) da |= (uint64_t)((start >> 7) & 7) << 49;
        }
        const uint32_t idesc = ptx::umma_idesc(2u, N);
        ptx::umma<false>(tmem, da, db, idesc, 0u);
        ptx::umma_commit(bm);
    }
    ptx::mbar_wait(bm, 0);
    ptx::tc_fence_after();
    uint32_t r[16];
    ptx::tmem_ld_32x16(tmem + ((uint32_t)((tid / 32) * 32) << 16), r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) a.out[tid * 16 + j] = __uint_as_float(r[j]);
    ptx::tc_fence_before();
    __syncthreads();
    if (tid < 32) ptx::tmem_dealloc(tmem, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q));
    std::vector<float> hA(ROWS * 32), hB(N * 8);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = (float)((int)(i * 7 % 23) - 11);     // small integers: exact in tf32
    for (size_t i = 0; i < hB.size(); ++i) hB[i] = (float)((int)(i * 5 % 13) - 6);
    float *dA, *dB, *dOut;
    CHECK(cudaMalloc(&dA, hA.size() * 4));
    CHECK(cudaMalloc(&dB, hB.size() * 4));
    CHECK(cudaMalloc(&dOut, 128 * 16 * 4));
    CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 24576 + 1024 + 1024;
    CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> out(128 * 16);
    for (int mode = 0; mode < 3; ++mode) {
        const int width = mode == 0 ? 32 : mode == 1 ? 16 : 4;    // floats per slab row
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        if (mode != 2) {
            cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)ROWS};
            cuuint64_t strides[1] = {(cuuint64_t)width * 4};
            cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)ROWS};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
        }
        for (int bom = 0; bom < (mode == 2 ? 1 : 2); ++bom) {
            printf("mode %d (%s) base_offset_mode %d:", mode, mode == 0 ? "SW128" : mode == 1 ? "SW64" : "no-swizzle overlap", bom);
            for (int shift = 0; shift < 20; ++shift) {
                Args a{mode, shift, bom, dOut, dA, dB};
                CHECK(cudaMemset(dOut, 0xff, 128 * 16 * 4));
                probe_kernel<<<1, 128, smem>>>(map, a);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" [shift %d: %s]\n", shift, cudaGetErrorString(e)); return 3; }
                CHECK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < 8; ++k) {
                            // slab element feeding (row m, k): modes 0/1: row m+shift, col k; mode 2: linear float 4*(m+shift) + k
                            const float av = mode == 2 ? hA[4 * (m + shift) + k] : hA[(size_t)(m + shift) * width + k];
                            ref += (double)av * hB[n * 8 + k];
                        }
                        const double d = fabs(ref - out[m * 16 + n]);
                        if (!(d <= maxerr)) maxerr = d;
                    }
                printf(" s%d:%s", shift, maxerr == 0 ? "ok" : "BAD");
            }
            printf("\n");
        }
    }
    return 0;
}
