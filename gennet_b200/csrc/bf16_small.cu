// Bandwidth-bound companions of the tensor-core path: the first convolution of every
// network (Cin = 1 or 2: K = 5..10, no GEMM worth the name) and the Dense heads with <= 4 outputs over the
// flattened feature map (a GEMV over up to 519 168 features).  HBM-bound streaming kernels: 128-bit accesses,
// warp-shuffle reductions, one pass over the big operand.
// Reference layers: bbhMahoGANy.py:362,382 (first Conv1D of both PE towers), :439 (first D conv, Cin'=2),
// :377,399,494 (Dense(1) heads); burstMahoGANy.py:272,310.
#include "gn_common.cuh"

#include <cuda_bf16.h>

namespace gn {

// The big operand of every kernel here (the activation / gradient tensor that is streamed once) is bf16 in the bf16
// throughput mode and float32 in the split-operand ("bf16x3") mode, whose activations stay float32: the kernels are
// templates over that element type and touch it eight elements at a time (one 128-bit access for bf16, two for float).
template <typename T> struct Act8;
template <> struct Act8<__nv_bfloat16> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[2 * e] = f.x;
            v[2 * e + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
    }
    static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }   // as stored
    static __device__ __forceinline__ float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};
template <> struct Act8<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw ldraw(const float* p) {
        Raw r;
        r.a = __ldg(reinterpret_cast<const float4*>(p));
        r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        return r;
    }
    static __device__ __forceinline__ Raw zero() { Raw r; r.a = r.b = make_float4(0.f, 0.f, 0.f, 0.f); return r; }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ float round(float x) { return x; }
    static __device__ __forceinline__ float get(const float* p) { return *p; }
};

// ---- first-layer convolution: x f32 (B,L,CIN), w f32 (k,CIN,Cout) -> y bf16 (B,Lout,Cout), fused bias + act ----
// thread = (row, 8 consecutive output channels); weights (<= 16*2*256 floats) live in shared memory
template <int CIN, typename T>
__global__ void __launch_bounds__(256) conv_smallcin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                T* __restrict__ y, int B, int L, int Lout,
                                                                int Cout, int k, int s, int p, int act, float ap) {
    extern __shared__ float sw[];     // k*CIN*Cout weights then Cout bias
    const int nw = k * CIN * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[nw + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int groups = Cout / 8;
    const long long total = (long long)B * Lout * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int g, l;
        const long long row = fast_div(i, groups, g);
        const int b = (int)fast_div(row, Lout, l);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = sw[nw + g * 8 + j];
        for (int t = 0; t < k; ++t) {
            const int pos = l * s + t - p;
            if (pos < 0 || pos >= L) continue;
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float xv = __ldg(&x[((size_t)b * L + pos) * CIN + c]);
                const float4* wr = reinterpret_cast<const float4*>(&sw[(t * CIN + c) * Cout + g * 8]);
                const float4 w0 = wr[0], w1 = wr[1];
                acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]);
                acc[3] = fmaf(xv, w0.w, acc[3]); acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
                acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = act_fwd(acc[j], act, ap);
        Act8<T>::store(y + (size_t)row * Cout + g * 8, acc);
    }
}

// Register-resident variant for k <= 5: a thread keeps the k*CIN*8 weights and 8 biases of its channel group in
// registers and walks a run of consecutive output rows of one sample, so a row costs its k*CIN input loads, the FMAs,
// the (compile-time) activation and one 128-bit store -- no shared-memory weight reads and no per-row index division.
template <int CIN, int KMAX, int KIND, typename T>
__global__ void __launch_bounds__(256) conv_smallcin_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ bias,
                                                                    T* __restrict__ y, int B, int L, int Lout,
                                                                    int Cout, int k, int s, int p, float ap, int gpb, int run,
                                                                    int runs_per_sample, long long n_runs) {
    const int g = blockIdx.y * gpb + threadIdx.x % gpb;          // channel group (8 channels)
    const int lane_r = threadIdx.x / gpb, lanes = blockDim.x / gpb;
    if (g >= Cout / 8 || lane_r >= lanes) return;
    float wr[KMAX * CIN][8], bv[8];
#pragma unroll
    for (int i = 0; i < KMAX * CIN; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[i][j] = (i < k * CIN) ? __ldg(&w[(size_t)i * Cout + g * 8 + j]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = bias ? __ldg(&bias[g * 8 + j]) : 0.f;
    for (long long r = (long long)blockIdx.x * lanes + lane_r; r < n_runs; r += (long long)gridDim.x * lanes) {
        int ri;
        const int b = (int)fast_div(r, runs_per_sample, ri);
        const int l0 = ri * run, l1 = min(Lout, l0 + run);
        const float* __restrict__ xb = x + (size_t)b * L * CIN;
        T* __restrict__ yb = y + ((size_t)b * Lout) * Cout + g * 8;
        for (int l = l0; l < l1; ++l) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = bv[j];
#pragma unroll
            for (int t = 0; t < KMAX; ++t) {
                const int pos = l * s + t - p;
                if (t < k && pos >= 0 && pos < L) {
#pragma unroll
                    for (int c = 0; c < CIN; ++c) {
                        const float xv = __ldg(&xb[(size_t)pos * CIN + c]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wr[t * CIN + c][j], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = act_fwd_t<KIND>(acc[j], ap);
            Act8<T>::store(yb + (size_t)l * Cout, acc);
        }
    }
}

// ---- first-layer weight gradient: dw f32 (k,CIN,Cout), db f32 (Cout) from x f32 and dy bf16 (pre-activation grad) ----
// thread = 8 consecutive output channels (one 128-bit dy load per row) x one row lane; k*CIN*8 (+8 bias) partial sums
template <int CIN, int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_smallcin_wgrad_kernel(const float* __restrict__ x,
                                                                  const T* __restrict__ dy,
                                                                  float* __restrict__ dw, float* __restrict__ db, int B,
                                                                  int L, int Lout, int Cout, int k, int s, int p,
                                                                  long long rows_per_block, int ld, int co0) {
    // Cout = width of the channel slice handled by this launch (<= 128), ld = channels per dy row, co0 = first channel
    const int groups = Cout / 8;
    const int g = threadIdx.x % groups;
    const int ry = threadIdx.x / groups, nry = blockDim.x / groups;
    float acc[KMAX * CIN][8];
    float accb[8];
#pragma unroll
    for (int i = 0; i < KMAX * CIN; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accb[j] = 0.f;
    const long long rows = (long long)B * Lout;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (ry < nry) {
        // every row lane walks its own contiguous run of the block's rows: (sample, position) advance incrementally
        const long long per_lane = (r1 - r0 + nry - 1) / nry;
        const long long q0 = r0 + (long long)ry * per_lane, q1 = min(r1, q0 + per_lane);
        int l = 0;
        int b = (q0 < q1) ? (int)fast_div(q0, Lout, l) : 0;
        // rows in batches of U: all U 128-bit dy loads of a thread are issued before the first is used (the kernel is
        // latency bound otherwise: two 256-thread blocks per SM at ~100 registers keep too few bytes in flight)
        // CIN = 2 already sits at the 128-register limit of two blocks per SM; float rows take twice the registers
        constexpr int U = (CIN == 1 ? 8 : 2) / (sizeof(T) == 4 ? 2 : 1);
        const T* __restrict__ dyg = dy + co0 + g * 8;
        for (long long row = q0; row < q1; row += U) {
            typename Act8<T>::Raw pkv[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                pkv[u] = (row + u < q1) ? Act8<T>::ldraw(dyg + (size_t)(row + u) * ld) : Act8<T>::zero();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (row + u < q1) {
                    if (l == Lout) { l = 0; ++b; }
                    float gv[8];
                    Act8<T>::unpack(pkv[u], gv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) accb[j] += gv[j];
#pragma unroll
                    for (int t = 0; t < KMAX; ++t) {
                        if (t < k) {
                            const int pos = l * s + t - p;
                            if (pos >= 0 && pos < L) {
#pragma unroll
                                for (int c = 0; c < CIN; ++c) {
                                    const float xv = __ldg(&x[((size_t)b * L + pos) * CIN + c]);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) acc[t * CIN + c][j] = fmaf(xv, gv[j], acc[t * CIN + c][j]);
                                }
                            }
                        }
                    }
                    ++l;
                }
            }
        }
    }
    // combine: row lanes that share a warp by shuffles, the 8 warps through shared memory, then one atomic per
    // (tap, cin, channel) and block
    constexpr int NV = KMAX * CIN + 1;
    __shared__ float sm[8][NV][128];            // Cout <= 128
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = (i < KMAX * CIN) ? acc[i < KMAX * CIN ? i : 0][j] : accb[j];
            for (int o = groups; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < groups && groups <= 32) sm[warp][i][(lane % groups) * 8 + j] = v;
        }
    }
    __syncthreads();
    const int gpw = groups < 32 ? groups : 32;       // column groups held by one warp's lanes
    (void)gpw;
    for (int e = threadIdx.x; e < NV * Cout; e += blockDim.x) {
        const int i = e / Cout, co = e - i * Cout;
        if (i < KMAX * CIN && i >= k * CIN) continue;
        float t = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += sm[w8][i][co];
        if (i < KMAX * CIN) atomicAdd(&dw[(size_t)i * ld + co0 + co], t);
        else if (db != nullptr) atomicAdd(&db[co0 + co], t);
    }
}

// ---- first-layer data gradient: dx f32 (B,L,CIN) from dy bf16 (B,Lout,Cout) and w f32 (k,CIN,Cout) ----------------
// Needed when the gradient flows on through a Cin <= 2 convolution (generator step through the frozen discriminator,
// bbhMahoGANy.py:1296).  One warp per dy row: the row (Cout bf16) is read once with 128-bit loads, dotted with the
// k*CIN weight rows held in shared memory, reduced by shuffles and scattered with k*CIN atomics into dx (zeroed by
// the caller): HBM traffic = dy once.
template <int CIN, int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_smallcin_dgrad_kernel(const T* __restrict__ dy,
                                                                  const float* __restrict__ w, float* __restrict__ dx,
                                                                  int B, int L, int Lout, int Cout, int k, int s, int p) {
    extern __shared__ float sw[];     // k*CIN*Cout
    const int nw = k * CIN * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * Lout;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = warp0; row < rows; row += nwarps) {
        int l;
        const int b = (int)fast_div(row, Lout, l);
        float acc[KMAX * CIN];
#pragma unroll
        for (int i = 0; i < KMAX * CIN; ++i) acc[i] = 0.f;
        for (int c8 = lane; c8 < Cout / 8; c8 += 32) {
            float gv[8];
            Act8<T>::load(dy + (size_t)row * Cout + c8 * 8, gv);
#pragma unroll
            for (int i = 0; i < KMAX * CIN; ++i) {
                if (i < k * CIN) {
                    const float4* wr = reinterpret_cast<const float4*>(&sw[(size_t)i * Cout + c8 * 8]);
                    const float4 w0 = wr[0], w1 = wr[1];
                    acc[i] = fmaf(gv[0], w0.x, acc[i]); acc[i] = fmaf(gv[1], w0.y, acc[i]);
                    acc[i] = fmaf(gv[2], w0.z, acc[i]); acc[i] = fmaf(gv[3], w0.w, acc[i]);
                    acc[i] = fmaf(gv[4], w1.x, acc[i]); acc[i] = fmaf(gv[5], w1.y, acc[i]);
                    acc[i] = fmaf(gv[6], w1.z, acc[i]); acc[i] = fmaf(gv[7], w1.w, acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < KMAX * CIN; ++i) {
            float v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[i] = v;
        }
        if (lane < k * CIN) {
            const int t = lane / CIN, c = lane - t * CIN;
            const int pos = l * s + t - p;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < KMAX * CIN; ++i) v = (i == lane) ? acc[i] : v;
            if (pos >= 0 && pos < L) atomicAdd(&dx[((size_t)b * L + pos) * CIN + c], v);
        }
    }
}

// ---- first-layer convolution, ANY filter count and taps up to 16 (float32 only): Conv1D(50, 16) on (out_dim, 1) of the
// 2_model_version discriminators (no_mode_collapse_network.py:117, subtract_model.py:134), Conv1D(25, 5) on (8192, 1)
// of train_on_wvf_version/nn.py:95.  With Cin <= 2 the layer is k*Cin FMAs per output and bound by writing y / reading dy;
// the implicit-GEMM kernels pad K = k*Cin to 16 and Cout to 128 and run it 10-30x below that bound.
//   fwd  : thread = one output element (row, co), consecutive threads = consecutive co (coalesced store, broadcast x)
//   wgrad: thread = (co, row lane): k*CIN accumulators in registers, dy read once (coalesced), x by broadcast loads
//   dgrad: one warp per dy row, k*CIN dot products reduced by shuffles and scattered with atomics (dx zeroed by the caller)
constexpr int SCG_KMAX = 16;
constexpr int SCG_R = 4;            // consecutive output rows per thread: one input window and one weight read feed all of them
// fwd: thread = (block of SCG_R consecutive output rows of one sample, co); consecutive threads = consecutive co
template <int CIN, int S>
__global__ void __launch_bounds__(256) conv_smallcin_gen_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, float* __restrict__ y,
                                                                    int B, int L, int Lout, int Cout, int k, int p,
                                                                    int act, float ap) {
    extern __shared__ float sw[];     // k*CIN*Cout weights then Cout bias
    const int nw = k * CIN * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[nw + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    constexpr int WIN = (SCG_R - 1) * S + SCG_KMAX;
    const int nblk = (Lout + SCG_R - 1) / SCG_R;
    const long long total = (long long)B * nblk * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int co, rb;
        const long long blk = fast_div(i, Cout, co);
        const int b = (int)fast_div(blk, nblk, rb);
        const int l0 = rb * SCG_R;
        const float* __restrict__ xb = x + (size_t)b * L * CIN;
        const int base = l0 * S - p, wlen = (SCG_R - 1) * S + k;
        float xw[WIN][CIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int pos = base + j;
            const bool ok = j < wlen && pos >= 0 && pos < L;
#pragma unroll
            for (int c = 0; c < CIN; ++c) xw[j][c] = ok ? __ldg(&xb[(size_t)pos * CIN + c]) : 0.f;
        }
        float acc[SCG_R];
#pragma unroll
        for (int r = 0; r < SCG_R; ++r) acc[r] = sw[nw + co];
#pragma unroll
        for (int t = 0; t < SCG_KMAX; ++t) {
            if (t < k) {
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float wv = sw[(t * CIN + c) * Cout + co];
#pragma unroll
                    for (int r = 0; r < SCG_R; ++r) acc[r] = fmaf(xw[r * S + t][c], wv, acc[r]);
                }
            }
        }
        float* __restrict__ yb = y + ((size_t)b * Lout + l0) * Cout + co;
#pragma unroll
        for (int r = 0; r < SCG_R; ++r)
            if (l0 + r < Lout) yb[(size_t)r * Cout] = act_fwd(acc[r], act, ap);
    }
}

// wgrad: block = 64 output channels x 4 row lanes (blockIdx.y = slice of 64 channels); a lane takes blocks of SCG_R
// consecutive rows of one sample: one input window and SCG_R gradient values feed k*CIN*SCG_R FMAs
template <int CIN, int S>
__global__ void __launch_bounds__(256) conv_smallcin_gen_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                      float* __restrict__ dw, float* __restrict__ db, int B,
                                                                      int L, int Lout, int Cout, int k, int p,
                                                                      long long blks_per_block) {
    __shared__ float red[4][SCG_KMAX * CIN + 1][64];
    const int cl = threadIdx.x & 63, ry = threadIdx.x >> 6;
    const int co = blockIdx.y * 64 + cl;
    const bool live = co < Cout;
    constexpr int WIN = (SCG_R - 1) * S + SCG_KMAX;
    float acc[SCG_KMAX * CIN], accb = 0.f;
#pragma unroll
    for (int i = 0; i < SCG_KMAX * CIN; ++i) acc[i] = 0.f;
    const int nblk = (Lout + SCG_R - 1) / SCG_R;
    const long long nb_total = (long long)B * nblk;
    const long long q0 = (long long)blockIdx.x * blks_per_block, q1 = min(nb_total, q0 + blks_per_block);
    if (live) {
        for (long long q = q0 + ry; q < q1; q += 4) {
            int rb;
            const int b = (int)fast_div(q, nblk, rb);
            const int l0 = rb * SCG_R;
            const float* __restrict__ xb = x + (size_t)b * L * CIN;
            const float* __restrict__ gp = dy + ((size_t)b * Lout + l0) * Cout + co;
            float g[SCG_R];
#pragma unroll
            for (int r = 0; r < SCG_R; ++r) {
                g[r] = (l0 + r < Lout) ? __ldg(&gp[(size_t)r * Cout]) : 0.f;
                accb += g[r];
            }
            const int base = l0 * S - p, wlen = (SCG_R - 1) * S + k;
            float xw[WIN][CIN];
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const int pos = base + j;
                const bool ok = j < wlen && pos >= 0 && pos < L;
#pragma unroll
                for (int c = 0; c < CIN; ++c) xw[j][c] = ok ? __ldg(&xb[(size_t)pos * CIN + c]) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < SCG_KMAX; ++t) {
                if (t < k) {
#pragma unroll
                    for (int c = 0; c < CIN; ++c)
#pragma unroll
                        for (int r = 0; r < SCG_R; ++r) acc[t * CIN + c] = fmaf(xw[r * S + t][c], g[r], acc[t * CIN + c]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < SCG_KMAX * CIN; ++i) red[ry][i][cl] = acc[i];
    red[ry][SCG_KMAX * CIN][cl] = accb;
    __syncthreads();
    for (int e = threadIdx.x; e < (SCG_KMAX * CIN + 1) * 64; e += blockDim.x) {
        const int i = e >> 6, c2 = e & 63, cg = blockIdx.y * 64 + c2;
        if (cg >= Cout || (i < SCG_KMAX * CIN && i >= k * CIN)) continue;
        const float v = red[0][i][c2] + red[1][i][c2] + red[2][i][c2] + red[3][i][c2];
        if (i < SCG_KMAX * CIN) atomicAdd(&dw[(size_t)i * Cout + cg], v);
        else if (db != nullptr) atomicAdd(&db[cg], v);
    }
}

template <int CIN>
__global__ void __launch_bounds__(256) conv_smallcin_gen_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                      float* __restrict__ dx, int B, int L, int Lout, int Cout,
                                                                      int k, int s, int p) {
    extern __shared__ float sw[];     // k*CIN*Cout
    const int nw = k * CIN * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * Lout;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = warp0; row < rows; row += nwarps) {
        int l;
        const int b = (int)fast_div(row, Lout, l);
        float acc[SCG_KMAX * CIN];
#pragma unroll
        for (int i = 0; i < SCG_KMAX * CIN; ++i) acc[i] = 0.f;
        for (int co = lane; co < Cout; co += 32) {
            const float g = __ldg(&dy[(size_t)row * Cout + co]);
#pragma unroll
            for (int i = 0; i < SCG_KMAX * CIN; ++i)
                if (i < k * CIN) acc[i] = fmaf(g, sw[(size_t)i * Cout + co], acc[i]);
        }
        // lane i ends up with the full sum of dot product i (k*CIN <= 32)
        float mine = 0.f;
#pragma unroll
        for (int i = 0; i < SCG_KMAX * CIN; ++i) {
            float v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (i == lane) mine = v;
        }
        if (lane < k * CIN) {
            const int t = lane / CIN, c = lane - t * CIN;
            const int pos = l * s + t - p;
            if (pos >= 0 && pos < L) atomicAdd(&dx[((size_t)b * L + pos) * CIN + c], mine);
        }
    }
}

// ---- Dense with N <= 4 outputs over bf16 features -------------------------------------------------------------
// fwd: y[m, j] = act(sum_k x[m,k] w[k,j] + b[j]) : one CTA per row m, 128-bit loads, block reduction
template <int NS, typename T>
__global__ void __launch_bounds__(512) dense_small_fwd_bf16_kernel(const T* __restrict__ x,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int K, int act, float ap) {
    __shared__ float sm[32];
    const int m = blockIdx.x;
    const T* xr = x + (size_t)m * K;
    float acc[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = 0.f;
    const int K8 = K / 8;
    for (int i = threadIdx.x; i < K8; i += blockDim.x) {
        float xv[8];
        Act8<T>::load(xr + (size_t)i * 8, xv);
        // the 8*NS weights of this chunk are contiguous: NS*2 128-bit loads
        float wv[8 * NS];
        const float4* wp = reinterpret_cast<const float4*>(w + (size_t)i * 8 * NS);
#pragma unroll
        for (int q4 = 0; q4 < 2 * NS; ++q4) {
            float4 t4 = __ldg(&wp[q4]);
            wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[j] = fmaf(xv[e], wv[e * NS + j], acc[j]);
        }
    }
    for (int kk = K8 * 8 + threadIdx.x; kk < K; kk += blockDim.x) {
        float v = Act8<T>::get(xr + kk);
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[j] = fmaf(v, __ldg(&w[(size_t)kk * NS + j]), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        float t = block_sum(acc[j], sm);
        if (threadIdx.x == 0) y[(size_t)m * NS + j] = act_fwd(t + (bias ? bias[j] : 0.f), act, ap);
    }
}

// dgrad: dx[m,k] (bf16) = act'(x[m,k]) * sum_j dy[m,j] w[k,j]   (x = the layer's own input, post-activation)
// thread = 8 consecutive features (its 8*NS weights stay in registers) x a slice of the rows; optionally the
// per-channel sum of dx (feature k belongs to channel k % C: the flattened (L, C) output of a convolution), which
// is the bias gradient of that convolution, accumulated with one atomic per thread and channel.
template <int NS, int KIND, typename T>
__global__ void __launch_bounds__(256) dense_small_dgrad_bf16_kernel(const float* __restrict__ dy,
                                                                     const float* __restrict__ w,
                                                                     const T* __restrict__ xin,
                                                                     T* __restrict__ dx, int M, int K,
                                                                     int m_per_split, float ap,
                                                                     float* __restrict__ colsum, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // chunk of 8 features
    const bool live = c < K / 8;
    const int mb = blockIdx.y * m_per_split, me = live ? min(M, mb + m_per_split) : mb;
    float wv[8 * NS];
    const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(live ? c : 0) * 8 * NS);
#pragma unroll
    for (int q4 = 0; q4 < 2 * NS; ++q4) {
        float4 t4 = __ldg(&wp[q4]);
        wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
    }
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
#pragma unroll 4
    for (int m = mb; m < me; ++m) {
        float g[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) g[j] = __ldg(&dy[(size_t)m * NS + j]);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) sacc = fmaf(g[j], wv[e * NS + j], sacc);
            o[e] = sacc;
        }
        if (KIND != GN_ACT_NONE) {
            float xv[8];
            Act8<T>::load(xin + (size_t)m * K + (size_t)c * 8, xv);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] *= act_bwd_t<KIND>(xv[e], ap);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[e] += Act8<T>::round(o[e]);      // the sum is over the values as stored
        Act8<T>::store(dx + (size_t)m * K + (size_t)c * 8, o);
    }
    if (colsum != nullptr) {
        // same-address atomics serialise in L2 (~tens of ns each), so fold inside the block first: the block's
        // 2048 consecutive features wrap around the C channels 2048/C times
        __shared__ float red[2048];
        const bool fold = (2048 % C) == 0;       // block-uniform
        if (fold) {
#pragma unroll
            for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = (c < K / 8) ? cs[e] : 0.f;
            __syncthreads();
            const int f0 = blockIdx.x * 2048;    // first feature of the block; f0 % C == 0
            for (int ch = threadIdx.x; ch < C; ch += 256) {
                float t = 0.f;
                for (int f = ch; f < 2048; f += C) t += red[f];
                if (f0 + ch < K) atomicAdd(&colsum[ch], t);
            }
        } else if (c < K / 8) {
            const int ch = (c * 8) % C;          // C % 8 == 0: the 8 features of a chunk are 8 consecutive channels
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&colsum[ch + e], cs[e]);
        }
    }
}

// wgrad: dw[k,j] = sum_m x[m,k] dy[m,j] : thread = 8 consecutive k, loop over a slice of m, atomics across slices
template <int NS, typename T>
__global__ void __launch_bounds__(256) dense_small_wgrad_bf16_kernel(const T* __restrict__ x,
                                                                     const float* __restrict__ dy,
                                                                     float* __restrict__ dw, int M, int K,
                                                                     int m_per_split) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // chunk of 8 features
    if (c >= K / 8) return;
    const int mb = blockIdx.y * m_per_split, me = min(M, mb + m_per_split);
    float acc[8][NS];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[e][j] = 0.f;
    for (int m = mb; m < me; ++m) {
        float xv[8];
        Act8<T>::load(x + (size_t)m * K + (size_t)c * 8, xv);
        float g[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) g[j] = __ldg(&dy[(size_t)m * NS + j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[e][j] = fmaf(xv[e], g[j], acc[e][j]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int j = 0; j < NS; ++j) atomicAdd(&dw[(size_t)(c * 8 + e) * NS + j], acc[e][j]);
}

__global__ void dense_small_bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, int M, int N) {
    const int j = threadIdx.x;
    if (j >= N) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += dy[(size_t)m * N + j];
    db[j] = s;
}

// elementwise act backward on bf16 tensors (used when the consumer could not fuse the mask)
__global__ void __launch_bounds__(256) act_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ dx, long long n, int act, float ap) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * act_bwd_from_y(__bfloat162float(y[i]), act, ap));
}

// ---- UpSampling1D on bf16 activations (generator, bbhMahoGANy.py:248,258): y[b, j, :] = x[b, j / size, :] and its
// adjoint dx[b, l, :] = sum_r dy[b, l*size + r, :]; 8 channels (16 bytes) per thread
__global__ void __launch_bounds__(256) upsample_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            long long rows_out, int L, int C8, int size) {
    const long long total = rows_out * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, j;
        const long long r = fast_div(i, C8, c);     // output row = b * (L*size) + j
        const long long b = fast_div(r, L * size, j);
        reinterpret_cast<uint4*>(y)[i] = __ldg(reinterpret_cast<const uint4*>(x) + (b * L + j / size) * C8 + c);
    }
}
__global__ void __launch_bounds__(256) upsample_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dx, long long rows_in, int C8,
                                                                int size) {
    const long long total = rows_in * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c;
        const long long r = fast_div(i, C8, c);     // input row = b * L + l ; its copies are rows r*size .. r*size+size-1
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int q = 0; q < size; ++q) {
            uint4 pk = __ldg(reinterpret_cast<const uint4*>(dy) + (r * size + q) * C8 + c);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 v = __bfloat1622float2(h[e]);
                acc[2 * e] += v.x;
                acc[2 * e + 1] += v.y;
            }
        }
        __nv_bfloat162 o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
        reinterpret_cast<uint4*>(dx)[i] = *reinterpret_cast<uint4*>(o);
    }
}

// ---- last convolution of the generator: Cout = 1, stride 1 (bbhMahoGANy.py:291).  With one output channel the layer
// is five dot products per input row: tap t of row pos contributes d_t = <x[b,pos,:], w[t,:]> to y[b, pos - t + p].
// fwd: one warp per input row of a 32-row output tile, 128-bit loads, shuffle reduction, tap products combined through
// shared memory in a fixed order (bit-reproducible).  dgrad: dx[b,pos,:] = sum_t dy[b,pos-t+p] w[t,:] streamed out row by row.  wgrad:
// dw[t,:] = sum_rows x[row,:] dy[row shifted by t]; thread = 8 channels, a slice of the rows, atomics across slices.
constexpr int COUT1_ROWS = 32;      // output rows per block of the Cout = 1 forward
template <int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_cout1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y, int B,
                                                             int L, int Lout, int Cin, int k, int p, int tiles_per_sample) {
    // One block = COUT1_ROWS consecutive output rows of one sample.  Every input row the tile touches (the tile plus a
    // k-1 row halo) is read once by one warp, which leaves its k tap products in shared memory; the outputs are then
    // summed from shared memory in a fixed order: no atomics, bit-reproducible, x is read (1 + (k-1)/32) times.
    // Weights in shared memory as float4 [tap][iteration][half][lane]: lane l of iteration i owns channels
    // 8 (l + 32 i) .. + 7, so a warp's 128-bit reads are contiguous (conflict-free); a warp works on two rows at a time
    // so that every weight read feeds two rows.
    extern __shared__ float4 sw4[];   // k * n_it * 2 * 32 float4 weights, then (COUT1_ROWS + KMAX - 1) * KMAX partial products
    const int C8 = Cin / 8, n_it = (C8 + 31) / 32;
    float* part = reinterpret_cast<float*>(sw4 + (size_t)k * n_it * 64);
    {
        float* swf = reinterpret_cast<float*>(sw4);
        for (int i = threadIdx.x; i < k * n_it * 256; i += blockDim.x) {
            const int j = i & 3, lane_ = (i >> 2) & 31, h = (i >> 7) & 1, ti = i >> 8;      // ti = t * n_it + it
            const int t = ti / n_it, it = ti - t * n_it;
            const int c = (lane_ + 32 * it) * 8 + h * 4 + j;
            swf[i] = (c < Cin) ? w[(size_t)t * Cin + c] : 0.f;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bv = bias ? bias[0] : 0.f;
    for (long long tile = blockIdx.x; tile < (long long)B * tiles_per_sample; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_sample);
        const int l0 = (int)(tile - (long long)b * tiles_per_sample) * COUT1_ROWS;
        const int nin = COUT1_ROWS + k - 1;                       // input rows l0 - p ... l0 - p + nin - 1
        for (int ri = 2 * warp; ri < nin; ri += 16) {
            float acc[2][KMAX];
#pragma unroll
            for (int t = 0; t < KMAX; ++t) acc[0][t] = acc[1][t] = 0.f;
            const int pos0 = l0 - p + ri;
            const bool ok0 = pos0 >= 0 && pos0 < L, ok1 = ri + 1 < nin && pos0 + 1 >= 0 && pos0 + 1 < L;
            const T* xr = x + ((size_t)b * L + pos0) * Cin;
            for (int it = 0; it < n_it; ++it) {
                const int c8 = lane + 32 * it;
                float xv[2][8];
#pragma unroll
                for (int e = 0; e < 8; ++e) xv[0][e] = xv[1][e] = 0.f;
                if (c8 < C8) {
                    if (ok0) Act8<T>::load(xr + c8 * 8, xv[0]);
                    if (ok1) Act8<T>::load(xr + Cin + c8 * 8, xv[1]);
                }
#pragma unroll
                for (int t = 0; t < KMAX; ++t) {
                    if (t < k) {
                        const float4 w0 = sw4[((t * n_it + it) * 2 + 0) * 32 + lane], w1 = sw4[((t * n_it + it) * 2 + 1) * 32 + lane];
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            acc[r][t] = fmaf(xv[r][0], w0.x, acc[r][t]); acc[r][t] = fmaf(xv[r][1], w0.y, acc[r][t]);
                            acc[r][t] = fmaf(xv[r][2], w0.z, acc[r][t]); acc[r][t] = fmaf(xv[r][3], w0.w, acc[r][t]);
                            acc[r][t] = fmaf(xv[r][4], w1.x, acc[r][t]); acc[r][t] = fmaf(xv[r][5], w1.y, acc[r][t]);
                            acc[r][t] = fmaf(xv[r][6], w1.z, acc[r][t]); acc[r][t] = fmaf(xv[r][7], w1.w, acc[r][t]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int t = 0; t < KMAX; ++t) {
                    float v = acc[r][t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0 && ri + r < nin) part[(ri + r) * KMAX + t] = v;
                }
        }
        __syncthreads();
        if (threadIdx.x < COUT1_ROWS) {
            const int l = l0 + threadIdx.x;
            if (l < Lout) {
                float v = bv;             // y[l] = bias + sum_t <x[l + t - p], w[t]>; input row l + t - p is tile row i + t
                for (int t = 0; t < k; ++t) v += part[(threadIdx.x + t) * KMAX + t];
                y[(size_t)b * Lout + l] = v;
            }
        }
        __syncthreads();
    }
}

// A thread keeps ONE group of 8 channels (its k x 8 weights in registers) and walks a contiguous run of rows with a
// sliding window of the k gradient values; consecutive threads hold consecutive channel groups, so a warp writes
// contiguous runs of a row.
template <int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_cout1_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                               T* __restrict__ dx, int B, int L, int Lout,
                                                               int Cin, int k, int p, long long lanes, long long run) {
    const int C8 = Cin / 8;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = (int)(tid % C8);
    const long long lane = tid / C8;
    if (lane >= lanes) return;
    const long long rows = (long long)B * L;
    const long long r0 = lane * run, r1 = min(rows, r0 + run);
    if (r0 >= r1) return;
    float wv[KMAX][8];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
        if (t < k) {
            const float4* wp = reinterpret_cast<const float4*>(w + (size_t)t * Cin + cg * 8);
            const float4 w0 = __ldg(&wp[0]), w1 = __ldg(&wp[1]);
            wv[t][0] = w0.x; wv[t][1] = w0.y; wv[t][2] = w0.z; wv[t][3] = w0.w;
            wv[t][4] = w1.x; wv[t][5] = w1.y; wv[t][6] = w1.z; wv[t][7] = w1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) wv[t][e] = 0.f;
        }
    }
    int pos;
    int b = (int)fast_div(r0, L, pos);
    float g[KMAX];      // g[t] = dy[b, pos - t + p]
    auto reload = [&]() {
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            const int l = pos - t + p;
            g[t] = (t < k && l >= 0 && l < Lout) ? __ldg(&dy[(size_t)b * Lout + l]) : 0.f;
        }
    };
    reload();
    for (long long row = r0; row < r1; ++row) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
        for (int t = 0; t < KMAX; ++t)
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(g[t], wv[t][e], o[e]);
        Act8<T>::store(dx + ((size_t)row * C8 + cg) * 8, o);
        if (++pos == L) {
            pos = 0;
            ++b;
            if (row + 1 < r1) reload();
        } else {
#pragma unroll
            for (int t = KMAX - 1; t > 0; --t) g[t] = g[t - 1];
            const int l = pos + p;
            g[0] = (l < Lout) ? __ldg(&dy[(size_t)b * Lout + l]) : 0.f;
        }
    }
}

// dw[t, c] = sum over rows of x[row, c] dy[row shifted by t].  Block = 32 channel groups x 8 row lanes (a warp reads a
// contiguous 1 KB run of one row); every warp walks a contiguous run of rows with the sliding gradient window, the eight
// warps of a block meet in shared memory, one atomic per (tap, channel) and block.  Needs (Cin / 8) % 32 == 0.
template <int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_cout1_wgrad_kernel(const T* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, int B, int L, int Lout, int Cin, int k,
                                                               int p, long long rows_per_block) {
    __shared__ float red[8][KMAX * 8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = blockIdx.x * 32 + lane;
    const long long rows = (long long)B * L;
    const long long b0 = (long long)blockIdx.y * rows_per_block, b1 = min(rows, b0 + rows_per_block);
    const long long per = (b1 - b0 + 7) / 8;
    const long long r0 = b0 + warp * per, r1 = min(b1, r0 + per);
    float acc[KMAX][8];
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
    if (r0 < r1) {
        int pos;
        int b = (int)fast_div(r0, L, pos);
        float g[KMAX];
        auto reload = [&]() {
#pragma unroll
            for (int t = 0; t < KMAX; ++t) {
                const int l = pos - t + p;
                g[t] = (t < k && l >= 0 && l < Lout) ? __ldg(&dy[(size_t)b * Lout + l]) : 0.f;
            }
        };
        reload();
#pragma unroll 4
        for (long long row = r0; row < r1; ++row) {
            float xv[8];
            Act8<T>::load(x + ((size_t)row * (Cin / 8) + cg) * 8, xv);
#pragma unroll
            for (int t = 0; t < KMAX; ++t)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(xv[e], g[t], acc[t][e]);
            if (++pos == L) {
                pos = 0;
                ++b;
                if (row + 1 < r1) reload();
            } else {
#pragma unroll
                for (int t = KMAX - 1; t > 0; --t) g[t] = g[t - 1];
                const int l = pos + p;
                g[0] = (l < Lout) ? __ldg(&dy[(size_t)b * Lout + l]) : 0.f;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) red[warp][t * 8 + e][lane] = acc[t][e];
    __syncthreads();
    for (int i = threadIdx.x; i < k * 8 * 32; i += blockDim.x) {
        const int l_ = i & 31, te = i >> 5;      // te = t * 8 + e
        float v = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) v += red[wq][te][l_];
        const int t = te >> 3, e = te & 7;
        atomicAdd(&dw[(size_t)t * Cin + (blockIdx.x * 32 + l_) * 8 + e], v);
    }
}

// general channel counts: thread = 8 channels, a slice of the rows, atomics across slices
template <int KMAX, typename T>
__global__ void __launch_bounds__(256) conv_cout1_wgrad_slices_kernel(const T* __restrict__ x, const float* __restrict__ dy,
                                                                      float* __restrict__ dw, int B, int L, int Lout, int Cin,
                                                                      int k, int p, long long rows_per_split) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // chunk of 8 channels
    if (c >= Cin / 8) return;
    const long long rows = (long long)B * L;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
    float acc[KMAX][8];
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
#pragma unroll 4
    for (long long row = r0; row < r1; ++row) {
        int pos;
        const int b = (int)fast_div(row, L, pos);
        float xv[8];
        Act8<T>::load(x + (size_t)row * Cin + (size_t)c * 8, xv);
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            if (t < k) {
                const int l = pos - t + p;
                const float g = (l >= 0 && l < Lout) ? __ldg(&dy[(size_t)b * Lout + l]) : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(xv[e], g, acc[t][e]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
        if (t < k)
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&dw[(size_t)t * Cin + c * 8 + e], acc[t][e]);
}

__global__ void __launch_bounds__(1024) sum_all_kernel(const float* __restrict__ x, long long n, float* out) {
    __shared__ float sm[32];
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    float t = block_sum(s, sm);
    if (threadIdx.x == 0) out[0] = t;
}

}  // namespace gn

using namespace gn;

extern "C" int gn_upsample1d_fwd_bf16(const void* x, void* y, int B, int L, int C, int size, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && L > 0 && C > 0 && C % 8 == 0 && size >= 1, "null pointer or bad size (C % 8 == 0)");
    if (B == 0) return GN_OK;
    const long long rows = (long long)B * L * size, total = rows * (C / 8);
    unsigned grid = (unsigned)((total + 255) / 256 < 16LL * num_sms() ? (total + 255) / 256 : 16LL * num_sms());
    upsample_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, rows, L, C / 8, size);
    return cuda_status("upsample_bf16_kernel");
}

extern "C" int gn_upsample1d_bwd_bf16(const void* dy, void* dx, int B, int L, int C, int size, void* stream) {
    GN_REQUIRE(dy && dx && B >= 0 && L > 0 && C > 0 && C % 8 == 0 && size >= 1, "null pointer or bad size (C % 8 == 0)");
    if (B == 0) return GN_OK;
    const long long rows = (long long)B * L, total = rows * (C / 8);
    unsigned grid = (unsigned)((total + 255) / 256 < 16LL * num_sms() ? (total + 255) / 256 : 16LL * num_sms());
    upsample_bwd_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, rows, C / 8, size);
    return cuda_status("upsample_bwd_bf16_kernel");
}

static int check_cout1(int B, int L, int Cin, int Lout, int k, int p) {
    GN_REQUIRE(B >= 0 && L > 0 && Lout > 0 && Cin > 0 && Cin % 8 == 0 && k > 0 && k <= 5 && p >= 0 && p < k,
               "needs Cin % 8 == 0, k <= 5, stride 1");
    GN_REQUIRE(Lout <= L + k - 1, "Lout too large for L");
    return GN_OK;
}

template <typename T>
static int cout1_fwd(const T* x, const float* w, const float* bias, float* y, int B, int L, int Cin, int Lout, int k,
                     int pad_left, void* stream) {
    GN_REQUIRE(x && w && y, "null pointer");
    int rc = check_cout1(B, L, Cin, Lout, k, pad_left);
    if (rc != GN_OK) return rc;
    const int n_it = (Cin / 8 + 31) / 32;
    const size_t smem = sizeof(float) * ((size_t)k * n_it * 256 + (size_t)(COUT1_ROWS + 4) * 5);
    GN_REQUIRE(smem <= 48 * 1024, "weights do not fit shared memory");
    cudaStream_t st = as_stream(stream);
    if (B == 0) return GN_OK;
    const int tps = (Lout + COUT1_ROWS - 1) / COUT1_ROWS;
    long long blocks = (long long)B * tps;
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
    conv_cout1_fwd_kernel<5, T><<<(unsigned)blocks, 256, smem, st>>>(x, w, bias, y, B, L, Lout, Cin, k, pad_left, tps);
    return cuda_status("conv_cout1_fwd_kernel");
}

template <typename T>
static int cout1_dgrad(const float* dy, const float* w, T* dx, int B, int L, int Cin, int Lout, int k, int pad_left,
                       void* stream) {
    GN_REQUIRE(dy && w && dx, "null pointer");
    int rc = check_cout1(B, L, Cin, Lout, k, pad_left);
    if (rc != GN_OK) return rc;
    if (B == 0) return GN_OK;
    // about 16 CTAs of 256 threads per SM; every thread walks `run` consecutive rows of its channel group
    const long long C8 = Cin / 8, rows = (long long)B * L;
    long long lanes = 16LL * num_sms() * 256 / C8;
    if (lanes < 1) lanes = 1;
    if (lanes > rows) lanes = rows;
    const long long run = (rows + lanes - 1) / lanes;
    lanes = (rows + run - 1) / run;
    const unsigned grid = (unsigned)((lanes * C8 + 255) / 256);
    conv_cout1_dgrad_kernel<5, T><<<grid, 256, 0, as_stream(stream)>>>(dy, w, dx, B, L, Lout, Cin, k, pad_left, lanes, run);
    return cuda_status("conv_cout1_dgrad_kernel");
}

template <typename T>
static int cout1_wgrad(const T* x, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout, int k,
                       int pad_left, void* stream) {
    GN_REQUIRE(x && dy && dw, "null pointer");
    int rc = check_cout1(B, L, Cin, Lout, k, pad_left);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * Cin, st);
    if (B > 0 && (Cin / 8) % 32 == 0) {
        const long long rows = (long long)B * L;
        const int bx = Cin / 256;
        long long splits = (4LL * num_sms() + bx - 1) / bx;
        if (splits > (rows + 63) / 64) splits = (rows + 63) / 64;      // at least 8 rows per warp
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
        long long per = (rows + splits - 1) / splits;
        splits = (rows + per - 1) / per;
        dim3 grid((unsigned)bx, (unsigned)splits);
        conv_cout1_wgrad_kernel<5, T><<<grid, 256, 0, st>>>(x, dy, dw, B, L, Lout, Cin, k, pad_left, per);
    } else if (B > 0) {
        const long long rows = (long long)B * L;
        const int bx = (Cin / 8 + 255) / 256;
        // few, long slices: every slice ends in k*8 atomics per thread onto the same k*Cin addresses, and same-address
        // atomics serialise in L2
        long long splits = (2LL * num_sms() + bx - 1) / bx;
        if (Cin / 8 < 256) splits = splits * 256 / (Cin / 8);
        if (splits > rows) splits = rows;
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
        long long per = (rows + splits - 1) / splits;
        splits = (rows + per - 1) / per;
        const int threads = Cin / 8 < 256 ? ((Cin / 8 + 31) / 32) * 32 : 256;
        dim3 grid((Cin / 8 + threads - 1) / threads, (unsigned)splits);
        conv_cout1_wgrad_slices_kernel<5, T><<<grid, threads, 0, st>>>(x, dy, dw, B, L, Lout, Cin, k, pad_left, per);
    }
    if (db != nullptr) sum_all_kernel<<<1, 1024, 0, st>>>(dy, (long long)B * Lout, db);
    return cuda_status("conv_cout1_wgrad_kernel");
}

template <typename T>
static int smallcin_fwd(const float* x, const float* w, const float* bias, T* yy, int B, int L, int Cin, int Lout, int Cout,
                        int k, int stride, int pad_left, int act, float act_param, void* stream) {
    GN_REQUIRE(x && w && yy, "null pointer");
    GN_REQUIRE(B >= 0 && L > 0 && Lout > 0 && k > 0 && k <= 16 && stride > 0 && pad_left >= 0, "bad geometry");
    GN_REQUIRE((Cin == 1 || Cin == 2) && Cout > 0 && Cout <= 1024, "needs Cin in {1,2} and Cout <= 1024");
    GN_REQUIRE(Cout % 8 == 0 || sizeof(T) == 4, "bf16 outputs need Cout % 8 == 0");
    if (B == 0) return GN_OK;
    const size_t smem = sizeof(float) * ((size_t)k * Cin * Cout + Cout);
    GN_REQUIRE(smem <= 48 * 1024, "weights do not fit shared memory");
    if constexpr (sizeof(T) == 4) {
        if (Cout % 8 != 0) {      // any filter count: one thread per output element
            GN_REQUIRE(stride == 1 || stride == 2, "filter counts that are not multiples of 8 need stride 1 or 2");
            const long long total = (long long)B * ((Lout + SCG_R - 1) / SCG_R) * Cout;
            const unsigned grid = (unsigned)((total + 255) / 256 < 16LL * num_sms() ? (total + 255) / 256 : 16LL * num_sms());
            cudaStream_t st = as_stream(stream);
#define GN_SCG_FWD(CI, ST) conv_smallcin_gen_fwd_kernel<CI, ST><<<grid, 256, smem, st>>>(x, w, bias, (float*)yy, B, L, Lout, Cout, k, pad_left, act, act_param)
            if (Cin == 1) { if (stride == 1) GN_SCG_FWD(1, 1); else GN_SCG_FWD(1, 2); }
            else { if (stride == 1) GN_SCG_FWD(2, 1); else GN_SCG_FWD(2, 2); }
#undef GN_SCG_FWD
            return cuda_status("conv_smallcin_gen_fwd_kernel");
        }
    }
    if (k <= 5) {
        // register-resident path: 256 threads = gpb channel groups x row lanes; a run of 32 rows per thread visit
        const int groups = Cout / 8;
        const int gpb = groups < 32 ? groups : 32;
        const int run = 32;
        const int rps = (Lout + run - 1) / run;
        const long long n_runs = (long long)B * rps;
        const int lanes = 256 / gpb;
        long long bx = (n_runs + lanes - 1) / lanes;
        const long long cap = 8LL * num_sms();
        if (bx > cap) bx = cap;
        dim3 grid((unsigned)bx, (unsigned)((groups + gpb - 1) / gpb));
        cudaStream_t st = as_stream(stream);
#define GN_SCF(CI, KIND) conv_smallcin_fwd_reg_kernel<CI, 5, KIND, T><<<grid, 256, 0, st>>>(x, w, bias, yy, B, L, Lout, Cout, k, stride, pad_left, act_param, gpb, run, rps, n_runs)
#define GN_SCF_ACT(CI)                                              \
        switch (act) {                                              \
            case GN_ACT_RELU: GN_SCF(CI, GN_ACT_RELU); break;       \
            case GN_ACT_TANH: GN_SCF(CI, GN_ACT_TANH); break;       \
            case GN_ACT_SIGMOID: GN_SCF(CI, GN_ACT_SIGMOID); break; \
            case GN_ACT_LEAKY: GN_SCF(CI, GN_ACT_LEAKY); break;     \
            case GN_ACT_RELU_MAX: GN_SCF(CI, GN_ACT_RELU_MAX); break; \
            case GN_ACT_ELU: GN_SCF(CI, GN_ACT_ELU); break;         \
            default: GN_SCF(CI, GN_ACT_NONE); break;                \
        }
        if (Cin == 1) { GN_SCF_ACT(1) } else { GN_SCF_ACT(2) }
#undef GN_SCF_ACT
#undef GN_SCF
        return cuda_status("conv_smallcin_fwd_reg_kernel");
    }
    long long total = (long long)B * Lout * (Cout / 8);
    unsigned grid = (unsigned)((total + 255) / 256 < 16LL * num_sms() ? (total + 255) / 256 : 16LL * num_sms());
    if (Cin == 1)
        conv_smallcin_fwd_kernel<1, T><<<grid, 256, smem, as_stream(stream)>>>(x, w, bias, yy, B, L, Lout, Cout, k, stride,
                                                                              pad_left, act, act_param);
    else
        conv_smallcin_fwd_kernel<2, T><<<grid, 256, smem, as_stream(stream)>>>(x, w, bias, yy, B, L, Lout, Cout, k, stride,
                                                                              pad_left, act, act_param);
    return cuda_status("conv_smallcin_fwd_kernel");
}

template <typename T>
static int smallcin_wgrad(const float* x, const T* dy, float* dw, float* db, int B, int L, int Cin, int Lout, int Cout,
                          int k, int stride, int pad_left, void* stream) {
    GN_REQUIRE(x && dy && dw, "null pointer");
    GN_REQUIRE(B >= 0 && L > 0 && Lout > 0 && k > 0 && k <= 16 && stride > 0 && pad_left >= 0, "bad geometry (k <= 16)");
    const bool fast = k <= 5 && (Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64 || (Cout % 128 == 0 && Cout <= 2048));
    GN_REQUIRE((Cin == 1 || Cin == 2) && Cout > 0 && (fast || (sizeof(T) == 4 && Cout <= 1024)),
               "needs Cin in {1,2}; bf16 gradients need k <= 5 and Cout in {8,16,32,64} or a multiple of 128");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * Cin * Cout, st);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, st);
    if (B == 0) return GN_OK;
    const long long rows = (long long)B * Lout;
    if constexpr (sizeof(T) == 4) {
        if (!fast) {      // any filter count / up to 16 taps
            GN_REQUIRE(stride == 1 || stride == 2, "this filter count / tap count needs stride 1 or 2");
            const long long nblk_total = (long long)B * ((Lout + SCG_R - 1) / SCG_R);
            long long nb = 4LL * num_sms();
            long long per = (nblk_total + nb - 1) / nb;
            if (per < 4) per = 4;
            nb = (nblk_total + per - 1) / per;
            dim3 grid((unsigned)nb, (unsigned)((Cout + 63) / 64));
#define GN_SCG_WG(CI, ST) conv_smallcin_gen_wgrad_kernel<CI, ST><<<grid, 256, 0, st>>>(x, (const float*)dy, dw, db, B, L, Lout, Cout, k, pad_left, per)
            if (Cin == 1) { if (stride == 1) GN_SCG_WG(1, 1); else GN_SCG_WG(1, 2); }
            else { if (stride == 1) GN_SCG_WG(2, 1); else GN_SCG_WG(2, 2); }
#undef GN_SCG_WG
            return cuda_status("conv_smallcin_gen_wgrad_kernel");
        }
    }
    long long blocks = 2LL * num_sms();      // few blocks: each ends in k*Cin*Cout same-address atomics, which serialise in L2
    long long per = (rows + blocks - 1) / blocks;
    if (per < 64) per = 64;
    blocks = (rows + per - 1) / per;
    // wide layers (the first discriminator convolution has 2*256 packed outputs) go slice by slice of 128 channels
    const int slice = Cout < 128 ? Cout : 128;
    for (int co0 = 0; co0 < Cout; co0 += slice) {
        if (Cin == 1)
            conv_smallcin_wgrad_kernel<1, 5, T><<<(unsigned)blocks, 256, 0, st>>>(x, dy, dw, db, B, L, Lout, slice, k, stride,
                                                                                pad_left, per, Cout, co0);
        else
            conv_smallcin_wgrad_kernel<2, 5, T><<<(unsigned)blocks, 256, 0, st>>>(x, dy, dw, db, B, L, Lout, slice, k, stride,
                                                                                pad_left, per, Cout, co0);
    }
    return cuda_status("conv_smallcin_wgrad_kernel");
}

template <typename T>
static int smallcin_dgrad(const T* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int Cout, int k,
                          int stride, int pad_left, void* stream) {
    GN_REQUIRE(dy && w && dx, "null pointer");
    GN_REQUIRE(B >= 0 && L > 0 && Lout > 0 && k > 0 && k <= 16 && stride > 0 && pad_left >= 0, "bad geometry (k <= 16)");
    const bool fast = k <= 5 && Cout % 8 == 0;
    GN_REQUIRE((Cin == 1 || Cin == 2) && Cout > 0 && (fast || (sizeof(T) == 4 && k * Cin <= 32)),
               "needs Cin in {1,2}; bf16 gradients need k <= 5 and Cout % 8 == 0");
    const size_t smem = sizeof(float) * (size_t)k * Cin * Cout;
    GN_REQUIRE(smem <= 48 * 1024, "weights do not fit shared memory");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * L * Cin, st);
    if (B == 0) return GN_OK;
    const long long rows = (long long)B * Lout;
    long long blocks = (rows + 7) / 8;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    if constexpr (sizeof(T) == 4) {
        if (!fast) {
            if (Cin == 1)
                conv_smallcin_gen_dgrad_kernel<1><<<(unsigned)blocks, 256, smem, st>>>((const float*)dy, w, dx, B, L, Lout, Cout, k,
                                                                                       stride, pad_left);
            else
                conv_smallcin_gen_dgrad_kernel<2><<<(unsigned)blocks, 256, smem, st>>>((const float*)dy, w, dx, B, L, Lout, Cout, k,
                                                                                       stride, pad_left);
            return cuda_status("conv_smallcin_gen_dgrad_kernel");
        }
    }
    if (Cin == 1)
        conv_smallcin_dgrad_kernel<1, 5, T><<<(unsigned)blocks, 256, smem, st>>>(dy, w, dx, B, L, Lout, Cout, k, stride, pad_left);
    else
        conv_smallcin_dgrad_kernel<2, 5, T><<<(unsigned)blocks, 256, smem, st>>>(dy, w, dx, B, L, Lout, Cout, k, stride, pad_left);
    return cuda_status("conv_smallcin_dgrad_kernel");
}

template <typename T>
static int dense_small_fwd(const T* xb, const float* w, const float* bias, float* y, int M, int K, int N, int act,
                           float act_param, void* stream) {
    GN_REQUIRE(xb && w && y, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N >= 1 && N <= 4 && K % 8 == 0, "needs 1 <= N <= 4 and K % 8 == 0");
    if (M == 0) return GN_OK;
    cudaStream_t st = as_stream(stream);
    switch (N) {
        case 1: dense_small_fwd_bf16_kernel<1, T><<<M, 512, 0, st>>>(xb, w, bias, y, K, act, act_param); break;
        case 2: dense_small_fwd_bf16_kernel<2, T><<<M, 512, 0, st>>>(xb, w, bias, y, K, act, act_param); break;
        case 3: dense_small_fwd_bf16_kernel<3, T><<<M, 512, 0, st>>>(xb, w, bias, y, K, act, act_param); break;
        default: dense_small_fwd_bf16_kernel<4, T><<<M, 512, 0, st>>>(xb, w, bias, y, K, act, act_param); break;
    }
    return cuda_status("dense_small_fwd_kernel");
}

template <int NS, typename T>
static void launch_dense_small_dgrad(dim3 grid, cudaStream_t st, const float* dy, const float* w, const T* xi, T* d, int M,
                                     int K, int per, int in_act, float ap, float* colsum, int C) {
    if (xi == nullptr) in_act = GN_ACT_NONE;
#define GN_DSD(KIND) dense_small_dgrad_bf16_kernel<NS, KIND, T><<<grid, 256, 0, st>>>(dy, w, xi, d, M, K, per, ap, colsum, C)
    switch (in_act) {
        case GN_ACT_RELU: GN_DSD(GN_ACT_RELU); break;
        case GN_ACT_TANH: GN_DSD(GN_ACT_TANH); break;
        case GN_ACT_SIGMOID: GN_DSD(GN_ACT_SIGMOID); break;
        case GN_ACT_LEAKY: GN_DSD(GN_ACT_LEAKY); break;
        case GN_ACT_RELU_MAX: GN_DSD(GN_ACT_RELU_MAX); break;
        case GN_ACT_ELU: GN_DSD(GN_ACT_ELU); break;
        default: GN_DSD(GN_ACT_NONE); break;
    }
#undef GN_DSD
}

template <typename T>
static int dense_small_dgrad(const float* dy, const float* w, const T* xi, T* d, float* dx_colsum, int colsum_channels,
                             int M, int K, int N, int in_act, float in_act_param, void* stream) {
    GN_REQUIRE(dy && w && d, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N >= 1 && N <= 4 && K % 8 == 0, "needs 1 <= N <= 4 and K % 8 == 0");
    GN_REQUIRE(dx_colsum == nullptr || (colsum_channels > 0 && colsum_channels % 8 == 0 && K % colsum_channels == 0),
               "column sums need channels % 8 == 0 and K % channels == 0");
    cudaStream_t st = as_stream(stream);
    if (dx_colsum != nullptr) cudaMemsetAsync(dx_colsum, 0, sizeof(float) * (size_t)colsum_channels, st);
    if (M == 0) return GN_OK;
    const int chunks = K / 8;
    const int bx = (chunks + 255) / 256;
    int splits = (int)((3LL * num_sms() + bx - 1) / bx);      // ~3 blocks per SM keeps the atomic count low
    if (splits > M) splits = M;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int per = (M + splits - 1) / splits;
    splits = (M + per - 1) / per;
    dim3 grid(bx, splits);
    switch (N) {
        case 1: launch_dense_small_dgrad<1, T>(grid, st, dy, w, xi, d, M, K, per, in_act, in_act_param, dx_colsum, colsum_channels); break;
        case 2: launch_dense_small_dgrad<2, T>(grid, st, dy, w, xi, d, M, K, per, in_act, in_act_param, dx_colsum, colsum_channels); break;
        case 3: launch_dense_small_dgrad<3, T>(grid, st, dy, w, xi, d, M, K, per, in_act, in_act_param, dx_colsum, colsum_channels); break;
        default: launch_dense_small_dgrad<4, T>(grid, st, dy, w, xi, d, M, K, per, in_act, in_act_param, dx_colsum, colsum_channels); break;
    }
    return cuda_status("dense_small_dgrad_kernel");
}

template <typename T>
static int dense_small_wgrad(const T* xb, const float* dy, float* dw, float* db, int M, int K, int N, void* stream) {
    GN_REQUIRE(xb && dy && dw, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N >= 1 && N <= 4 && K % 8 == 0, "needs 1 <= N <= 4 and K % 8 == 0");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * N, st);
    if (M > 0) {
        const int chunks = K / 8;
        const int bx = (chunks + 255) / 256;
        int splits = (int)((4LL * num_sms() + bx - 1) / bx);
        if (splits > M) splits = M;
        if (splits < 1) splits = 1;
        if (splits > 65535) splits = 65535;
        int per = (M + splits - 1) / splits;
        splits = (M + per - 1) / per;
        dim3 grid(bx, splits);
        switch (N) {
            case 1: dense_small_wgrad_bf16_kernel<1, T><<<grid, 256, 0, st>>>(xb, dy, dw, M, K, per); break;
            case 2: dense_small_wgrad_bf16_kernel<2, T><<<grid, 256, 0, st>>>(xb, dy, dw, M, K, per); break;
            case 3: dense_small_wgrad_bf16_kernel<3, T><<<grid, 256, 0, st>>>(xb, dy, dw, M, K, per); break;
            default: dense_small_wgrad_bf16_kernel<4, T><<<grid, 256, 0, st>>>(xb, dy, dw, M, K, per); break;
        }
    }
    if (db != nullptr) dense_small_bias_grad_kernel<<<1, 32, 0, st>>>(dy, db, M, N);
    return cuda_status("dense_small_wgrad_kernel");
}

typedef __nv_bfloat16 bf16_t;

// bf16 activations (throughput mode)
extern "C" int gn_conv1d_cout1_fwd_bf16(const void* x, const float* w, const float* bias, float* y, int B, int L, int Cin,
                                        int Lout, int k, int pad_left, void* stream) {
    return cout1_fwd<bf16_t>((const bf16_t*)x, w, bias, y, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_cout1_dgrad_bf16(const float* dy, const float* w, void* dx, int B, int L, int Cin, int Lout, int k,
                                          int pad_left, void* stream) {
    return cout1_dgrad<bf16_t>(dy, w, (bf16_t*)dx, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_cout1_wgrad_bf16(const void* x, const float* dy, float* dw, float* db, int B, int L, int Cin,
                                          int Lout, int k, int pad_left, void* stream) {
    return cout1_wgrad<bf16_t>((const bf16_t*)x, dy, dw, db, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_smallcin_fwd_bf16(const float* x, const float* w, const float* bias, void* y, int B, int L,
                                           int Cin, int Lout, int Cout, int k, int stride, int pad_left, int act,
                                           float act_param, void* stream) {
    return smallcin_fwd<bf16_t>(x, w, bias, (bf16_t*)y, B, L, Cin, Lout, Cout, k, stride, pad_left, act, act_param, stream);
}
extern "C" int gn_conv1d_smallcin_wgrad_bf16(const float* x, const void* dy, float* dw, float* db, int B, int L, int Cin,
                                             int Lout, int Cout, int k, int stride, int pad_left, void* stream) {
    return smallcin_wgrad<bf16_t>(x, (const bf16_t*)dy, dw, db, B, L, Cin, Lout, Cout, k, stride, pad_left, stream);
}
extern "C" int gn_conv1d_smallcin_dgrad_bf16(const void* dy, const float* w, float* dx, int B, int L, int Cin, int Lout,
                                             int Cout, int k, int stride, int pad_left, void* stream) {
    return smallcin_dgrad<bf16_t>((const bf16_t*)dy, w, dx, B, L, Cin, Lout, Cout, k, stride, pad_left, stream);
}
extern "C" int gn_dense_small_fwd_bf16(const void* x, const float* w, const float* bias, float* y, int M, int K, int N,
                                       int act, float act_param, void* stream) {
    return dense_small_fwd<bf16_t>((const bf16_t*)x, w, bias, y, M, K, N, act, act_param, stream);
}
extern "C" int gn_dense_small_dgrad_bf16(const float* dy, const float* w, const void* x_in, void* dx, float* dx_colsum,
                                         int colsum_channels, int M, int K, int N, int in_act, float in_act_param,
                                         void* stream) {
    return dense_small_dgrad<bf16_t>(dy, w, (const bf16_t*)x_in, (bf16_t*)dx, dx_colsum, colsum_channels, M, K, N, in_act,
                                     in_act_param, stream);
}
extern "C" int gn_dense_small_wgrad_bf16(const void* x, const float* dy, float* dw, float* db, int M, int K, int N,
                                         void* stream) {
    return dense_small_wgrad<bf16_t>((const bf16_t*)x, dy, dw, db, M, K, N, stream);
}

// float32 activations (float32 and split-operand modes): same kernels, same arguments
extern "C" int gn_conv1d_cout1_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L, int Cin,
                                       int Lout, int k, int pad_left, void* stream) {
    return cout1_fwd<float>(x, w, bias, y, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_cout1_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int k,
                                         int pad_left, void* stream) {
    return cout1_dgrad<float>(dy, w, dx, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_cout1_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin,
                                         int Lout, int k, int pad_left, void* stream) {
    return cout1_wgrad<float>(x, dy, dw, db, B, L, Cin, Lout, k, pad_left, stream);
}
extern "C" int gn_conv1d_smallcin_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L,
                                          int Cin, int Lout, int Cout, int k, int stride, int pad_left, int act,
                                          float act_param, void* stream) {
    return smallcin_fwd<float>(x, w, bias, y, B, L, Cin, Lout, Cout, k, stride, pad_left, act, act_param, stream);
}
extern "C" int gn_conv1d_smallcin_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin,
                                            int Lout, int Cout, int k, int stride, int pad_left, void* stream) {
    return smallcin_wgrad<float>(x, dy, dw, db, B, L, Cin, Lout, Cout, k, stride, pad_left, stream);
}
extern "C" int gn_conv1d_smallcin_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout,
                                            int Cout, int k, int stride, int pad_left, void* stream) {
    return smallcin_dgrad<float>(dy, w, dx, B, L, Cin, Lout, Cout, k, stride, pad_left, stream);
}
extern "C" int gn_dense_small_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int K, int N,
                                      int act, float act_param, void* stream) {
    return dense_small_fwd<float>(x, w, bias, y, M, K, N, act, act_param, stream);
}
extern "C" int gn_dense_small_dgrad_f32(const float* dy, const float* w, const float* x_in, float* dx, float* dx_colsum,
                                        int colsum_channels, int M, int K, int N, int in_act, float in_act_param,
                                        void* stream) {
    return dense_small_dgrad<float>(dy, w, x_in, dx, dx_colsum, colsum_channels, M, K, N, in_act, in_act_param, stream);
}
extern "C" int gn_dense_small_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int M, int K, int N,
                                        void* stream) {
    return dense_small_wgrad<float>(x, dy, dw, db, M, K, N, stream);
}

extern "C" int gn_act_bwd_bf16(const void* dy, const void* y, void* dx, long long n, int act, float param, void* stream) {
    GN_REQUIRE(dy && y && dx && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    unsigned grid = (unsigned)((n + 255) / 256 < 16LL * num_sms() ? (n + 255) / 256 : 16LL * num_sms());
    act_bwd_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                            (__nv_bfloat16*)dx, n, act, param);
    return cuda_status("act_bwd_bf16_kernel");
}
