// host_util.cuh -- host-side helpers of libdnn_b200.so: the thread-local error string behind
// dnnb_last_error(), the CK() status macro, the driver's tensor-map encoders (resolved at run time, no
// link-time dependency on libcuda), the rounding the weight prepack uses, SAME/VALID geometry.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <cuda.h>
#include <cuda_runtime.h>

namespace {


thread_local std::string g_err;

int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(...)                                                                                       \
    do {                                                                                              \
        cudaError_t e_ = (__VA_ARGS__);                                                                    \
        if (e_ != cudaSuccess)                                                                        \
            return fail("%s:%d: %s -> %s", __FILE__, __LINE__, #__VA_ARGS__, cudaGetErrorString(e_));        \
    } while (0)

// ------------------------------------------------------------------------------------------
// driver entry points for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_fns() {
    if (g_encode_tiled && g_encode_im2col) return 0;
    cudaDriverEntryPointQueryResult q;
    void *f = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeIm2col not available");
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f);
    return 0;
}

CUtensorMapSwizzle swizzle_enum(int swz) {
    return swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// ------------------------------------------------------------------------------------------
// host-side rounding used by the weight prepack (round to nearest even)
float host_round_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0xfffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
uint16_t host_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

void out_pads(int in_size, int k, int s, int same, int *o, int *pf, int *pb) {
    // TF SAME / VALID, proj3/dnn_openblas.py:127-142
    if (same) {
        *o = (in_size + s - 1) / s;
        int pad = (*o - 1) * s + k - in_size;
        if (pad < 0) pad = 0;
        *pf = pad / 2;
        *pb = pad - pad / 2;
    } else {
        *o = (in_size - k + 1 + s - 1) / s;
        *pf = *pb = 0;
    }
}

// opt a kernel in to > 48 KB of dynamic shared memory, once per (kernel, device)
template <typename K>
cudaError_t set_smem_once(K kern, int bytes, bool (&done)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return cudaSuccess;
}
}  // namespace
