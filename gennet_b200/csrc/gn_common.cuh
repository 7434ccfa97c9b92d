// Shared helpers for the gennet_b200 sm_100a kernels and their C-ABI wrappers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gennet_b200.h"

namespace gn {

extern thread_local char g_err[512];

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

inline int cuda_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return GN_ERR_CUDA;
    }
    return GN_OK;
}

#define GN_REQUIRE(cond, msg)                                                          \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            snprintf(gn::g_err, sizeof(gn::g_err), "%s: invalid argument: %s", __func__, msg); \
            return GN_ERR_ARG;                                                         \
        }                                                                              \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem32) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    T r = 0;
    if (w == 0) {
        r = (lane < (int)((blockDim.x + 31) >> 5)) ? smem32[lane] : (T)0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// x / d and x % d for a non-negative 64-bit x: the 32-bit unsigned division (a handful of instructions) whenever the
// operands fit, which they do for every tensor that fits a GPU; the generic 64-bit division is a ~100-instruction
// subroutine and used to dominate the index math of the streaming kernels.
__device__ __forceinline__ long long fast_div(long long x, int d, int& rem) {
    if (((unsigned long long)x >> 32) == 0ull) {
        const unsigned q = (unsigned)x / (unsigned)d;
        rem = (int)((unsigned)x - q * (unsigned)d);
        return (long long)q;
    }
    const long long q = x / d;
    rem = (int)(x - q * d);
    return q;
}

// activation codes shared by fused epilogues and the elementwise kernels.  The *_t forms take the code at
// compile time; act_dispatch() turns a run-time code into one uniform branch OUTSIDE the element loops (a
// per-element `switch` compiles to an indirect branch per element, which serialises a fused epilogue).
template <int KIND>
__device__ __forceinline__ float act_fwd_t(float x, float a) {
    if (KIND == GN_ACT_RELU) return x > 0.f ? x : 0.f;
    if (KIND == GN_ACT_TANH) return tanhf(x);
    if (KIND == GN_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
    if (KIND == GN_ACT_LEAKY) return x >= 0.f ? x : a * x;
    if (KIND == GN_ACT_RELU_MAX) return fminf(fmaxf(x, 0.f), a);
    if (KIND == GN_ACT_ELU) return x > 0.f ? x : expm1f(x);
    return x;
}
// derivative expressed through the OUTPUT y (all supported activations allow it)
template <int KIND>
__device__ __forceinline__ float act_bwd_t(float y, float a) {
    if (KIND == GN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (KIND == GN_ACT_TANH) return 1.f - y * y;
    if (KIND == GN_ACT_SIGMOID) return y * (1.f - y);
    if (KIND == GN_ACT_LEAKY) return y >= 0.f ? 1.f : a;
    if (KIND == GN_ACT_RELU_MAX) return (y > 0.f && y < a) ? 1.f : 0.f;
    if (KIND == GN_ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
    return 1.f;
}
template <int K>
struct ActTag {
    static constexpr int kind = K;
};
template <typename F>
__device__ __forceinline__ void act_dispatch(int kind, F&& f) {
    if (kind == GN_ACT_RELU) f(ActTag<GN_ACT_RELU>());
    else if (kind == GN_ACT_TANH) f(ActTag<GN_ACT_TANH>());
    else if (kind == GN_ACT_SIGMOID) f(ActTag<GN_ACT_SIGMOID>());
    else if (kind == GN_ACT_LEAKY) f(ActTag<GN_ACT_LEAKY>());
    else if (kind == GN_ACT_RELU_MAX) f(ActTag<GN_ACT_RELU_MAX>());
    else if (kind == GN_ACT_ELU) f(ActTag<GN_ACT_ELU>());
    else f(ActTag<GN_ACT_NONE>());
}
// run-time code, one element (compare chain, no jump table); prefer act_dispatch around a loop
__device__ __forceinline__ float act_fwd(float x, int kind, float a) {
    if (kind == GN_ACT_RELU) return act_fwd_t<GN_ACT_RELU>(x, a);
    if (kind == GN_ACT_TANH) return act_fwd_t<GN_ACT_TANH>(x, a);
    if (kind == GN_ACT_SIGMOID) return act_fwd_t<GN_ACT_SIGMOID>(x, a);
    if (kind == GN_ACT_LEAKY) return act_fwd_t<GN_ACT_LEAKY>(x, a);
    if (kind == GN_ACT_RELU_MAX) return act_fwd_t<GN_ACT_RELU_MAX>(x, a);
    if (kind == GN_ACT_ELU) return act_fwd_t<GN_ACT_ELU>(x, a);
    return x;
}
__device__ __forceinline__ float act_bwd_from_y(float y, int kind, float a) {
    if (kind == GN_ACT_RELU) return act_bwd_t<GN_ACT_RELU>(y, a);
    if (kind == GN_ACT_TANH) return act_bwd_t<GN_ACT_TANH>(y, a);
    if (kind == GN_ACT_SIGMOID) return act_bwd_t<GN_ACT_SIGMOID>(y, a);
    if (kind == GN_ACT_LEAKY) return act_bwd_t<GN_ACT_LEAKY>(y, a);
    if (kind == GN_ACT_RELU_MAX) return act_bwd_t<GN_ACT_RELU_MAX>(y, a);
    if (kind == GN_ACT_ELU) return act_bwd_t<GN_ACT_ELU>(y, a);
    return 1.f;
}

// ---- scaled fp16 pair format of the split-operand tensor-core kernels (conv1d_tc3.cu): a float32 tensor t with
// max |t| = amax travels as T0 = fp16(t s), T1 = fp16((t s - T0) 2^11) with s = 2^f16s_exp(amax)
constexpr float F16S_LO = 2048.f;             // 2^11: weight of the low plane
__host__ __device__ __forceinline__ int f16s_exp(float amax) {
#ifdef __CUDA_ARCH__
    const uint32_t bits = __float_as_uint(amax);
#else
    uint32_t bits;
    memcpy(&bits, &amax, 4);
#endif
    int e = (int)((bits >> 23) & 0xffu) - 127;      // floor(log2(amax)) of a normal number
    if (e < -100) e = -100;                          // zero / tiny tensors: any scale will do, keep 2^se finite
    return 14 - e;                                   // in [-113, 114]
}
__device__ __forceinline__ float pow2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }
__device__ __forceinline__ void split2h(float xs, uint16_t& h0, uint16_t& h1) {
    const __half a = __float2half_rn(xs);
    const __half b = __float2half_rn((xs - __half2float(a)) * F16S_LO);
    h0 = __half_as_ushort(a);
    h1 = __half_as_ushort(b);
}

}  // namespace gn
