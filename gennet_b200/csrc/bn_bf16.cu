// BatchNormalization -> activation -> dropout chains (generator, bbhMahoGANy.py:235-289: every hidden layer is
// Conv1D/Dense -> BatchNormalization(momentum) -> tanh -> Dropout) as three streaming passes, for bf16 activations
// (throughput mode) and for float32 activations (float32 / split-operand modes: libm-accurate activations, statistics
// accumulated in double from the first element on):
//   forward :  per-channel sum / sum of squares of x (one pass)  ->  y = drop(act(gamma * xhat + beta))  (one pass)
//   backward:  g = dy * drop' * act'(a) recomputed from x;  per-channel sum g, sum g*xhat (one pass)  ->
//              dx = gamma * invstd * (g - sum_g/n - xhat * sum_gxhat/n), dgamma, dbeta (one pass)
// Nothing but x, y and dx touches HBM: the activation output a and the dropout mask are recomputed (the mask from
// the Philox counter of gn_noise_draw_f32, so fused and unfused runs draw identical masks), all arithmetic is fp32,
// statistics accumulate in double.  Thread = 8 consecutive channels (one 128-bit access).
#include "gn_common.cuh"
#include "philox.cuh"

#include <cuda_bf16.h>

#include <type_traits>

namespace gn {

struct ChainArgs {
    const float* mean;        // (C) batch or moving mean; null = no normalisation (pure activation / noise)
    const float* scale;       // (C) 1/sqrt(var+eps), or the variance when use_var
    const float* gamma;       // (C)
    const float* beta;        // (C)
    int use_var;
    float eps;
    int act;
    float act_param;
    int noise;                // -1 none, else GN_NOISE_*
    float rate;
    const float* r;           // fed noise tensor (rows*C) or null = Philox(seed, offset)
    unsigned long long seed, offset;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 pk = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 f = __bfloat1622float2(h[e]);
        v[2 * e] = f.x;
        v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// Activation in the bf16 chain: the result is rounded to 8 mantissa bits, so tanh / sigmoid use the SFU approximations
// (tanh.approx.f32: ~2^-11 relative error) instead of the multi-instruction libm forms.
template <int KIND, bool EXACT>
__device__ __forceinline__ float chain_act(float h, float p) {
    if (EXACT) {
        // float32 activations.  The chain kernels are instruction-issue bound once tanh is a ~30-instruction libm call
        // (it is evaluated in the forward pass and again in both backward passes), so tanh takes the exponential form
        // 1 - 2 / (exp(2h) + 1) on the SFU (ex2.approx, rcp.approx): absolute error <= 1.5e-7 for every h (the relative
        // error of exp(2h), ~2.4e-7 + 1.2e-7 |h|, is scaled by 2e/(e+1)^2 <= 1/2), saturating correctly to +-1.
        if (KIND == GN_ACT_TANH) return 1.f - __fdividef(2.f, __expf(2.f * h) + 1.f);
        return act_fwd_t<KIND>(h, p);                 // the same forms as the per-layer kernels
    }
    if (KIND == GN_ACT_TANH) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(h));
        return y;
    }
    if (KIND == GN_ACT_SIGMOID) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(0.5f * h));
        return fmaf(0.5f, y, 0.5f);
    }
    return act_fwd_t<KIND>(h, p);
}

// multiplicative noise factors of the 8 elements starting at flat index i0 (i0 % 8 == 0); GaussianNoise (additive) is
// not a factor and is not handled by the chain
__device__ __forceinline__ void noise_factors(const ChainArgs& a, long long i0, float (&f)[8]) {
    if (a.noise < 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 1.f;
        return;
    }
    float rv[8];
    if (a.r != nullptr) {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(a.r + i0));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(a.r + i0) + 1);
        rv[0] = r0.x; rv[1] = r0.y; rv[2] = r0.z; rv[3] = r0.w; rv[4] = r1.x; rv[5] = r1.y; rv[6] = r1.z; rv[7] = r1.w;
    } else {
        // same stream as gn_noise_draw_f32 (rng_fill_kernel, stream id 3): block (offset + i) / 4, 4 values per block
        const unsigned long long blk = (a.offset + (unsigned long long)i0) >> 2;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint4 u = philox_flat(a.seed, blk + q, 3u);
            if (a.noise == GN_NOISE_DROPOUT) {
                rv[4 * q + 0] = u01(u.x) >= a.rate ? 1.f : 0.f;
                rv[4 * q + 1] = u01(u.y) >= a.rate ? 1.f : 0.f;
                rv[4 * q + 2] = u01(u.z) >= a.rate ? 1.f : 0.f;
                rv[4 * q + 3] = u01(u.w) >= a.rate ? 1.f : 0.f;
            } else {
                const float2 g0 = box_muller(u.x, u.y), g1 = box_muller(u.z, u.w);
                rv[4 * q + 0] = g0.x; rv[4 * q + 1] = g0.y; rv[4 * q + 2] = g1.x; rv[4 * q + 3] = g1.y;
            }
        }
    }
    if (a.noise == GN_NOISE_DROPOUT) {
        const float k = 1.f / (1.f - a.rate);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = rv[e] * k;
    } else {
        const float sd = sqrtf(a.rate / (1.f - a.rate));
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(rv[e], sd, 1.f);
    }
}

// per-channel affine of the normalisation for the 8 channels starting at c0:  h = x * sc + sh, xhat = (x - mu) * is
__device__ __forceinline__ void channel_affine(const ChainArgs& a, int c0, float (&mu)[8], float (&is)[8], float (&sc)[8],
                                               float (&sh)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (a.mean != nullptr) {
            mu[e] = __ldg(&a.mean[c0 + e]);
            const float s = __ldg(&a.scale[c0 + e]);
            is[e] = a.use_var ? rsqrtf(s + a.eps) : s;
            const float g = a.gamma ? __ldg(&a.gamma[c0 + e]) : 1.f;
            const float b = a.beta ? __ldg(&a.beta[c0 + e]) : 0.f;
            sc[e] = g * is[e];
            sh[e] = b - mu[e] * sc[e];
        } else {
            mu[e] = 0.f; is[e] = 1.f; sc[e] = 1.f; sh[e] = 0.f;
        }
    }
}

// ---- statistics: sums (2C) double += (sum x, sum x^2)  [caller zeroes] -------------------------------------------------
// block = (channel groups) x (row lanes); the lanes of a block are folded in shared memory before the double atomics
template <bool BWD, int KIND, typename T>
__global__ void __launch_bounds__(256) chain_sums_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         ChainArgs a, long long rows, int C, int gpb, long long rows_per_block,
                                                         double* __restrict__ sums) {
    constexpr bool EXACT = sizeof(T) == 4;
    typedef typename std::conditional<EXACT, double, float>::type Acc;      // per-thread partial sums
    __shared__ Acc sm[2][256][9];
    const int g = threadIdx.x % gpb, rl = threadIdx.x / gpb, nrl = blockDim.x / gpb;
    const int cg = blockIdx.x * gpb + g;                 // channel group (8 channels)
    const bool live = cg < C / 8 && rl < nrl;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    Acc s0[8], s1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s0[e] = s1[e] = (Acc)0;
    if (live) {
        float mu[8], is[8], sc[8], sh[8];
        if (BWD) channel_affine(a, cg * 8, mu, is, sc, sh);
        for (long long r = r0 + rl; r < r1; r += nrl) {
            const long long i0 = r * C + (long long)cg * 8;
            float xv[8];
            load8(x + i0, xv);
            if (!BWD) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s0[e] += (Acc)xv[e];
                    s1[e] += (Acc)xv[e] * (Acc)xv[e];
                }
            } else {
                float gv[8], nf[8];
                load8(dy + i0, gv);
                noise_factors(a, i0, nf);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float av = chain_act<KIND, EXACT>(fmaf(xv[e], sc[e], sh[e]), a.act_param);
                    const float gg = gv[e] * nf[e] * act_bwd_t<KIND>(av, a.act_param);
                    s0[e] += (Acc)gg;
                    s1[e] += (Acc)gg * (Acc)((xv[e] - mu[e]) * is[e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sm[0][threadIdx.x][e] = s0[e];
        sm[1][threadIdx.x][e] = s1[e];
    }
    __syncthreads();
    // thread (which, g, e) folds the row lanes: 2 * gpb * 8 results per block
    for (int t = threadIdx.x; t < 2 * gpb * 8; t += blockDim.x) {
        const int which = t / (gpb * 8), rem = t - which * gpb * 8, gg = rem / 8, e = rem - gg * 8;
        const int cgo = blockIdx.x * gpb + gg;
        if (cgo >= C / 8) continue;
        double acc = 0.0;
        for (int l = 0; l < nrl; ++l) acc += (double)sm[which][l * gpb + gg][e];
        atomicAdd(&sums[(size_t)which * C + cgo * 8 + e], acc);
    }
}

// optional side output of the apply kernels: max |result| into a device scalar (zeroed by the launcher), for the
// consumer's operand split (gn_split_f32_f16x2 with have_amax); non-negative floats order like their bit patterns
__device__ __forceinline__ void warp_amax(float* amax, float m) {
    if (amax == nullptr) return;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));
}

// ---- forward apply: y = noise(act(bn(x))) --------------------------------------------------------------------------
// A thread keeps ONE channel group for its whole life (its per-channel affine is computed once) and walks the rows
// with a stride of `lanes` = (threads of the grid) / (C/8); consecutive threads hold consecutive channel groups, so
// every access of a warp is a contiguous 512-byte run of a row.
template <int KIND, typename T>
__global__ void __launch_bounds__(256) chain_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                        ChainArgs a, long long rows, int C, long long lanes,
                                                        float* __restrict__ amax, uint16_t* __restrict__ planes,
                                                        float pscale, float pbound) {
    // planes != nullptr (float32 activations): the result is also written as the scaled fp16 pair the next convolution
    // consumes, planes (2, rows, C), with the scale of the a-priori bound pbound >= max |y| (bounded activations:
    // tanh, sigmoid, ReLU(max_value), times the dropout factor), which is stored to amax[0] for the consumer
    constexpr bool EXACT = sizeof(T) == 4;
    const int C8 = C / 8;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = (int)(tid % C8);
    const long long lane = tid / C8;
    float m = 0.f;
    if (planes != nullptr && tid == 0) amax[0] = pbound;
    if (lane < lanes) {
        float mu[8], is[8], sc[8], sh[8];
        channel_affine(a, cg * 8, mu, is, sc, sh);
#pragma unroll 2
        for (long long r = lane; r < rows; r += lanes) {
            const long long i0 = r * C + (long long)cg * 8;
            float xv[8], nf[8], o[8];
            load8(x + i0, xv);
            noise_factors(a, i0, nf);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                o[e] = chain_act<KIND, EXACT>(fmaf(xv[e], sc[e], sh[e]), a.act_param) * nf[e];
                m = fmaxf(m, fabsf(o[e]));
            }
            store8(y + i0, o);
            if (planes != nullptr) {
                uint32_t p0[4], p1[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint16_t a0, a1, b0, b1;
                    split2h(o[2 * e] * pscale, a0, a1);
                    split2h(o[2 * e + 1] * pscale, b0, b1);
                    p0[e] = (uint32_t)a0 | ((uint32_t)b0 << 16);
                    p1[e] = (uint32_t)a1 | ((uint32_t)b1 << 16);
                }
                *reinterpret_cast<uint4*>(planes + i0) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
                *reinterpret_cast<uint4*>(planes + (size_t)rows * C + i0) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
            }
        }
    }
    if (planes == nullptr) warp_amax(amax, m);
}

// ---- backward apply: dx = sc * (g - sum_g/n - xhat * sum_gxhat/n);  without normalisation dx = g ---------------------
template <int KIND, typename T>
__global__ void __launch_bounds__(256) chain_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                        T* __restrict__ dx, ChainArgs a, const double* __restrict__ sums,
                                                        double n_total, long long rows, int C, long long lanes,
                                                        float* __restrict__ amax) {
    constexpr bool EXACT = sizeof(T) == 4;
    const int C8 = C / 8;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = (int)(tid % C8);
    const long long lane = tid / C8;
    float m = 0.f;
    if (lane < lanes) {
        float mu[8], is[8], sc[8], sh[8], m0[8], m1[8];
        channel_affine(a, cg * 8, mu, is, sc, sh);
        const bool bn = a.mean != nullptr;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            m0[e] = bn ? (float)(sums[cg * 8 + e] / n_total) : 0.f;
            m1[e] = bn ? (float)(sums[C + cg * 8 + e] / n_total) : 0.f;
        }
#pragma unroll 2
        for (long long r = lane; r < rows; r += lanes) {
            const long long i0 = r * C + (long long)cg * 8;
            float xv[8], gv[8], nf[8], o[8];
            load8(x + i0, xv);
            load8(dy + i0, gv);
            noise_factors(a, i0, nf);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float av = chain_act<KIND, EXACT>(fmaf(xv[e], sc[e], sh[e]), a.act_param);
                const float gg = gv[e] * nf[e] * act_bwd_t<KIND>(av, a.act_param);
                o[e] = bn ? sc[e] * (gg - m0[e] - (xv[e] - mu[e]) * is[e] * m1[e]) : gg;
                m = fmaxf(m, fabsf(o[e]));
            }
            store8(dx + i0, o);
        }
    }
    warp_amax(amax, m);
}

__global__ void __launch_bounds__(256) chain_param_grads_kernel(const double* __restrict__ sums, float* dgamma, float* dbeta,
                                                                int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[C + c];
}

static void sums_geometry(long long rows, int C, int* gpb, dim3* grid, long long* rpb) {
    const int groups = C / 8;
    int g = groups < 256 ? groups : 256;
    *gpb = g;
    const int bx = (groups + g - 1) / g;
    const int nrl = 256 / g;
    long long splits = (4LL * num_sms() + bx - 1) / bx;
    long long maxs = (rows + nrl - 1) / nrl;
    if (splits > maxs) splits = maxs;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
    long long per = (rows + splits - 1) / splits;
    splits = (rows + per - 1) / per;
    *rpb = per;
    *grid = dim3((unsigned)bx, (unsigned)splits);
}

#define GN_CHAIN_DISPATCH(KERNEL_CALL)                                   \
    switch (a.act) {                                                     \
        case GN_ACT_RELU: { constexpr int K_ = GN_ACT_RELU; KERNEL_CALL; break; }         \
        case GN_ACT_TANH: { constexpr int K_ = GN_ACT_TANH; KERNEL_CALL; break; }         \
        case GN_ACT_SIGMOID: { constexpr int K_ = GN_ACT_SIGMOID; KERNEL_CALL; break; }   \
        case GN_ACT_LEAKY: { constexpr int K_ = GN_ACT_LEAKY; KERNEL_CALL; break; }       \
        case GN_ACT_RELU_MAX: { constexpr int K_ = GN_ACT_RELU_MAX; KERNEL_CALL; break; } \
        case GN_ACT_ELU: { constexpr int K_ = GN_ACT_ELU; KERNEL_CALL; break; }           \
        default: { constexpr int K_ = GN_ACT_NONE; KERNEL_CALL; break; }                  \
    }

// grid of the apply kernels: `lanes` row lanes x C/8 channel groups, about 16 CTAs of 256 threads per SM
static void apply_geometry(long long rows, int C, unsigned* grid, long long* lanes) {
    const long long C8 = C / 8;
    long long want = 16LL * num_sms() * 256;
    long long l = want / C8;
    if (l < 1) l = 1;
    if (l > rows) l = rows;
    *lanes = l;
    *grid = (unsigned)((l * C8 + 255) / 256);
}

static int make_chain(ChainArgs* a, const float* mean, const float* scale, const float* gamma, const float* beta, int use_var,
                      float eps, int act, float act_param, int noise, float rate, const float* r, uint64_t seed,
                      uint64_t offset, int C) {
    GN_REQUIRE(C > 0 && C % 8 == 0, "C must be a positive multiple of 8");
    GN_REQUIRE(mean == nullptr || scale != nullptr, "mean given without scale");
    GN_REQUIRE(noise == -1 || noise == GN_NOISE_DROPOUT || noise == GN_NOISE_GDROPOUT, "noise must be -1, dropout or gaussian dropout");
    GN_REQUIRE(noise < 0 || (rate >= 0.f && rate < 1.f), "rate must be in [0, 1)");
    GN_REQUIRE(offset % 4 == 0, "offset must be a multiple of 4");
    a->mean = mean; a->scale = scale; a->gamma = gamma; a->beta = beta; a->use_var = use_var; a->eps = eps;
    a->act = act; a->act_param = act_param; a->noise = noise; a->rate = rate; a->r = r; a->seed = seed; a->offset = offset;
    return GN_OK;
}

}  // namespace gn

using namespace gn;

template <typename T>
static int bn_stats_t(const T* x, long long rows, int C, double* sums, void* stream) {
    GN_REQUIRE(x && sums && rows >= 0 && C > 0 && C % 8 == 0, "null pointer or bad size (C % 8 == 0)");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st);
    if (rows == 0) return GN_OK;
    int gpb; dim3 grid; long long rpb;
    sums_geometry(rows, C, &gpb, &grid, &rpb);
    ChainArgs a{};
    chain_sums_kernel<false, GN_ACT_NONE, T><<<grid, 256, 0, st>>>(x, nullptr, a, rows, C, gpb, rpb, sums);
    return cuda_status("chain_sums_kernel");
}

template <typename T>
static int chain_fwd_t(const T* x, T* y, const float* mean, const float* scale, const float* gamma, const float* beta,
                       int use_var, float eps, int act, float act_param, int noise, float rate, const float* r, uint64_t seed,
                       uint64_t offset, long long rows, int C, float* amax, void* stream, void* planes = nullptr,
                       float bound = 0.f) {
    GN_REQUIRE(x && y && rows >= 0, "null pointer or rows < 0");
    GN_REQUIRE(planes == nullptr || (sizeof(T) == 4 && amax != nullptr && bound > 0.f),
               "planes need float32 activations, the scale scalar and a positive bound");
    ChainArgs a{};
    int rc = make_chain(&a, mean, scale, gamma, beta, use_var, eps, act, act_param, noise, rate, r, seed, offset, C);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    if (amax != nullptr && planes == nullptr) cudaMemsetAsync(amax, 0, sizeof(float), st);
    if (rows == 0) return GN_OK;
    unsigned grid; long long lanes;
    apply_geometry(rows, C, &grid, &lanes);
    const float pscale = planes != nullptr ? ldexpf(1.f, f16s_exp(bound)) : 1.f;
    GN_CHAIN_DISPATCH((chain_fwd_kernel<K_, T><<<grid, 256, 0, st>>>(x, y, a, rows, C, lanes, amax, (uint16_t*)planes, pscale,
                                                                   bound)));
    return cuda_status("chain_fwd_kernel");
}

template <typename T>
static int chain_bwd_sums_t(const T* x, const T* dy, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, int act, float act_param, int noise, float rate, const float* r, uint64_t seed,
                            uint64_t offset, long long rows, int C, double* sums, void* stream) {
    GN_REQUIRE(x && dy && mean && invstd && sums && rows >= 0, "null pointer or rows < 0");
    ChainArgs a{};
    int rc = make_chain(&a, mean, invstd, gamma, beta, 0, 0.f, act, act_param, noise, rate, r, seed, offset, C);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st);
    if (rows == 0) return GN_OK;
    int gpb; dim3 grid; long long rpb;
    sums_geometry(rows, C, &gpb, &grid, &rpb);
    GN_CHAIN_DISPATCH((chain_sums_kernel<true, K_, T><<<grid, 256, 0, st>>>(x, dy, a, rows, C, gpb, rpb, sums)));
    return cuda_status("chain_sums_kernel(bwd)");
}

template <typename T>
static int chain_bwd_t(const T* x, const T* dy, T* dx, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, const double* sums, double n_total, int act, float act_param, int noise, float rate,
                       const float* r, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta, long long rows, int C,
                       float* amax, void* stream) {
    GN_REQUIRE(x && dy && dx && rows >= 0, "null pointer or rows < 0");
    GN_REQUIRE(mean == nullptr || (invstd && sums && n_total > 0), "normalisation needs invstd, sums and n_total");
    ChainArgs a{};
    int rc = make_chain(&a, mean, invstd, gamma, beta, 0, 0.f, act, act_param, noise, rate, r, seed, offset, C);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    if (amax != nullptr) cudaMemsetAsync(amax, 0, sizeof(float), st);
    if (rows > 0) {
        unsigned grid; long long lanes;
        apply_geometry(rows, C, &grid, &lanes);
        GN_CHAIN_DISPATCH((chain_bwd_kernel<K_, T><<<grid, 256, 0, st>>>(x, dy, dx, a, sums, n_total, rows, C, lanes, amax)));
    }
    if (mean != nullptr && (dgamma || dbeta))
        chain_param_grads_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, dgamma, dbeta, C);
    return cuda_status("chain_bwd_kernel");
}

typedef __nv_bfloat16 bf16_t;

extern "C" int gn_bn_stats_bf16(const void* x, long long rows, int C, double* sums, void* stream) {
    return bn_stats_t<bf16_t>((const bf16_t*)x, rows, C, sums, stream);
}
extern "C" int gn_chain_fwd_bf16(const void* x, void* y, const float* mean, const float* scale, const float* gamma,
                                 const float* beta, int use_var, float eps, int act, float act_param, int noise, float rate,
                                 const float* r, uint64_t seed, uint64_t offset, long long rows, int C, void* stream) {
    return chain_fwd_t<bf16_t>((const bf16_t*)x, (bf16_t*)y, mean, scale, gamma, beta, use_var, eps, act, act_param, noise, rate,
                               r, seed, offset, rows, C, nullptr, stream);
}
extern "C" int gn_chain_bwd_sums_bf16(const void* x, const void* dy, const float* mean, const float* invstd, const float* gamma,
                                      const float* beta, int act, float act_param, int noise, float rate, const float* r,
                                      uint64_t seed, uint64_t offset, long long rows, int C, double* sums, void* stream) {
    return chain_bwd_sums_t<bf16_t>((const bf16_t*)x, (const bf16_t*)dy, mean, invstd, gamma, beta, act, act_param, noise, rate, r,
                                    seed, offset, rows, C, sums, stream);
}
extern "C" int gn_chain_bwd_bf16(const void* x, const void* dy, void* dx, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, const double* sums, double n_total, int act,
                                 float act_param, int noise, float rate, const float* r, uint64_t seed, uint64_t offset,
                                 float* dgamma, float* dbeta, long long rows, int C, void* stream) {
    return chain_bwd_t<bf16_t>((const bf16_t*)x, (const bf16_t*)dy, (bf16_t*)dx, mean, invstd, gamma, beta, sums, n_total, act,
                               act_param, noise, rate, r, seed, offset, dgamma, dbeta, rows, C, nullptr, stream);
}

// float32 activations: same arguments, every activation tensor is float*
extern "C" int gn_bn_sums_f32(const float* x, long long rows, int C, double* sums, void* stream) {
    return bn_stats_t<float>(x, rows, C, sums, stream);
}
extern "C" int gn_chain_fwd_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma,
                                const float* beta, int use_var, float eps, int act, float act_param, int noise, float rate,
                                const float* r, uint64_t seed, uint64_t offset, long long rows, int C, void* stream) {
    return chain_fwd_t<float>(x, y, mean, scale, gamma, beta, use_var, eps, act, act_param, noise, rate, r, seed, offset, rows, C,
                              nullptr, stream);
}
extern "C" int gn_chain_fwd_amax_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma,
                                     const float* beta, int use_var, float eps, int act, float act_param, int noise, float rate,
                                     const float* r, uint64_t seed, uint64_t offset, long long rows, int C, float* y_amax,
                                     void* stream) {
    return chain_fwd_t<float>(x, y, mean, scale, gamma, beta, use_var, eps, act, act_param, noise, rate, r, seed, offset, rows, C,
                              y_amax, stream);
}
extern "C" int gn_chain_bwd_sums_f32(const float* x, const float* dy, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int act, float act_param, int noise, float rate,
                                     const float* r, uint64_t seed, uint64_t offset, long long rows, int C, double* sums,
                                     void* stream) {
    return chain_bwd_sums_t<float>(x, dy, mean, invstd, gamma, beta, act, act_param, noise, rate, r, seed, offset, rows, C, sums,
                                   stream);
}
extern "C" int gn_chain_bwd_f32(const float* x, const float* dy, float* dx, const float* mean, const float* invstd,
                                const float* gamma, const float* beta, const double* sums, double n_total, int act,
                                float act_param, int noise, float rate, const float* r, uint64_t seed, uint64_t offset,
                                float* dgamma, float* dbeta, long long rows, int C, void* stream) {
    return chain_bwd_t<float>(x, dy, dx, mean, invstd, gamma, beta, sums, n_total, act, act_param, noise, rate, r, seed, offset,
                              dgamma, dbeta, rows, C, nullptr, stream);
}
extern "C" int gn_chain_fwd_planes_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma,
                                       const float* beta, int use_var, float eps, int act, float act_param, int noise,
                                       float rate, const float* r, uint64_t seed, uint64_t offset, long long rows, int C,
                                       void* y_planes, float* y_amax, float bound, void* stream) {
    GN_REQUIRE(y_planes && y_amax, "null pointer");
    return chain_fwd_t<float>(x, y, mean, scale, gamma, beta, use_var, eps, act, act_param, noise, rate, r, seed, offset, rows, C,
                              y_amax, stream, y_planes, bound);
}
extern "C" int gn_chain_bwd_amax_f32(const float* x, const float* dy, float* dx, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, const double* sums, double n_total, int act,
                                     float act_param, int noise, float rate, const float* r, uint64_t seed, uint64_t offset,
                                     float* dgamma, float* dbeta, long long rows, int C, float* dx_amax, void* stream) {
    return chain_bwd_t<float>(x, dy, dx, mean, invstd, gamma, beta, sums, n_total, act, act_param, noise, rate, r, seed, offset,
                              dgamma, dbeta, rows, C, dx_amax, stream);
}
