// Exact-parity float32 Conv1D / Dense kernels (forward, data gradient, weight gradient) on top of the
// SIMT implicit-GEMM core.  Keras semantics: NLC activations, Conv1D kernel (k,Cin,Cout), Dense kernel
// (in,out), cross-correlation, caller-supplied left zero padding (TF 'SAME' rule).
// Reference call sites: bbhMahoGANy.py:234,250-292,362-399,439-447,494; burstMahoGANy.py:158-202,272-290,
// 310-357; nn.py:73-89.
#include "gemm_simt.cuh"

namespace gn {

struct ConvGeom {
    int B, L, Lp, Cin, Lout, Cout, k, s, p, up;
};

// A(m=(b,l), kk=(tap,ci)) = x_logical[b, l*s + tap - p, ci], zero outside [0,L)
struct ConvFwdA {
    const float* __restrict__ x;
    ConvGeom g;
    __device__ __forceinline__ float operator()(int m, int kk) const {
        int b = m / g.Lout, l = m - b * g.Lout;
        int tap = kk / g.Cin, ci = kk - tap * g.Cin;
        int pos = l * g.s + tap - g.p;
        if (pos < 0 || pos >= g.L) return 0.f;
        if (g.up > 1) pos /= g.up;
        return __ldg(&x[((size_t)b * g.Lp + pos) * g.Cin + ci]);
    }
};
struct RowMajorB {  // B(k, n) = w[k*ld + n]
    const float* __restrict__ w;
    int ld;
    __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(&w[(size_t)k * ld + n]); }
};
struct RowMajorA {  // A(m, k) = x[m*ld + k]
    const float* __restrict__ x;
    int ld;
    __device__ __forceinline__ float operator()(int m, int k) const { return __ldg(&x[(size_t)m * ld + k]); }
};
struct ColMajorA {  // A(m, k) = x[k*ld + m]
    const float* __restrict__ x;
    int ld;
    __device__ __forceinline__ float operator()(int m, int k) const { return __ldg(&x[(size_t)k * ld + m]); }
};
struct TransB {  // B(k, n) = w[n*ld + k]
    const float* __restrict__ w;
    int ld;
    __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(&w[(size_t)n * ld + k]); }
};

struct BiasActStore {
    float* __restrict__ y;
    const float* __restrict__ bias;
    int ld, act;
    float ap;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        if (bias) v += __ldg(&bias[n]);
        y[(size_t)m * ld + n] = act_fwd(v, act, ap);
    }
};
__global__ void __launch_bounds__(256) bias_act_inplace_kernel(float* __restrict__ y, const float* __restrict__ bias,
                                                               long long n, int N, int act, float ap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = y[i];
    if (bias) v += __ldg(&bias[i % N]);
    y[i] = act_fwd(v, act, ap);
}
struct PlainStore {
    float* __restrict__ y;
    int ld;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { y[(size_t)m * ld + n] = v; }
};
struct AtomicStore {
    float* __restrict__ y;
    int ld;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { atomicAdd(&y[(size_t)m * ld + n], v); }
};

// dgrad: A(m=(b,jp), kk=(u,tap,co)) = dy[b, (jp*up+u + p - tap)/s, co] when divisible and in range
struct ConvDgradA {
    const float* __restrict__ dy;
    ConvGeom g;
    __device__ __forceinline__ float operator()(int m, int kk) const {
        int b = m / g.Lp, jp = m - b * g.Lp;
        int kc = g.k * g.Cout;
        int u = kk / kc, rem = kk - u * kc;
        int tap = rem / g.Cout, co = rem - tap * g.Cout;
        int t = jp * g.up + u + g.p - tap;
        if (t < 0) return 0.f;
        int l = t / g.s;
        if (l * g.s != t || l >= g.Lout) return 0.f;
        return __ldg(&dy[((size_t)b * g.Lout + l) * g.Cout + co]);
    }
};
// B(kk=(u,tap,co), n=ci) = w[tap, ci, co]
struct ConvDgradB {
    const float* __restrict__ w;
    ConvGeom g;
    __device__ __forceinline__ float operator()(int kk, int n) const {
        int kc = g.k * g.Cout;
        int rem = kk % kc;
        int tap = rem / g.Cout, co = rem - tap * g.Cout;
        return __ldg(&w[((size_t)tap * g.Cin + n) * g.Cout + co]);
    }
};
// wgrad: A'(m'=(tap,ci), kk=(b,l)) = x_logical[b, l*s+tap-p, ci]
struct ConvWgradA {
    const float* __restrict__ x;
    ConvGeom g;
    __device__ __forceinline__ float operator()(int mp, int kk) const {
        int tap = mp / g.Cin, ci = mp - tap * g.Cin;
        int b = kk / g.Lout, l = kk - b * g.Lout;
        int pos = l * g.s + tap - g.p;
        if (pos < 0 || pos >= g.L) return 0.f;
        if (g.up > 1) pos /= g.up;
        return __ldg(&x[((size_t)b * g.Lp + pos) * g.Cin + ci]);
    }
};
struct ConvWgradA_T {  // same, argument order (kk, mp) for the column-reduce kernel
    ConvWgradA a;
    __device__ __forceinline__ float operator()(int kk, int mp) const { return a(mp, kk); }
};

// per-column sum of a (rows, C) matrix in double, result cast to float (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int C,
                                                     long long rows_per_split, float* __restrict__ out) {
    // block = 32 columns x 8 row lanes
    __shared__ double sm[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.y * rows_per_split;
    const long long r1 = min(rows, r0 + rows_per_split);
    double s = 0.0;
    if (c < C)
        for (long long r = r0 + ry; r < r1; r += 8) s += (double)x[r * C + c];
    sm[ry][threadIdx.x & 31] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x & 31];
        atomicAdd(&out[c], (float)t);
    }
}

int launch_colsum(const float* x, long long rows, int C, float* out, cudaStream_t st) {
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, st);
    if (rows <= 0 || C <= 0) return GN_OK;
    int cb = (C + 31) / 32;
    long long splits = (4LL * num_sms() + cb - 1) / cb;
    if (splits > (rows + 63) / 64) splits = (rows + 63) / 64;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long per = (rows + splits - 1) / splits;
    splits = (rows + per - 1) / per;
    colsum_kernel<<<dim3(cb, (unsigned)splits), 256, 0, st>>>(x, rows, C, per, out);
    return cuda_status("colsum_kernel");
}

// dx[m, k] = sum_{j<NS} dy[m, j] * w[k, j]   (data gradient of a Dense layer with <= 4 outputs)
template <int NS>
__global__ void __launch_bounds__(256) outer_small_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int M, int K) {
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int m = (int)(i / K), k = (int)(i - (long long)m * K);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NS; ++j) s = fmaf(__ldg(&dy[(size_t)m * NS + j]), __ldg(&w[(size_t)k * NS + j]), s);
        dx[i] = s;
    }
}

__global__ void conv2d_w2_pack_kernel(const float* __restrict__ w2, const float* __restrict__ b,
                                      float* __restrict__ w1, float* __restrict__ b1, int kh, int kw, int Cin,
                                      int Cout, int pw) {
    // w1 [kh][(wi,ci)][(wo,co)]
    const long long total = (long long)kh * 2 * Cin * 2 * Cout;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        int co = (int)(i % Cout);
        long long r = i / Cout;
        int wo = (int)(r % 2); r /= 2;
        int ci = (int)(r % Cin); r /= Cin;
        int wi = (int)(r % 2); r /= 2;
        int h = (int)r;
        int q = wi - wo + pw;
        w1[i] = (q >= 0 && q < kw) ? w2[(((size_t)h * kw + q) * Cin + ci) * Cout + co] : 0.f;
    }
    if (b1 != nullptr && i < 2LL * Cout) b1[i] = b[i % Cout];
}

__global__ void conv2d_w2_unpack_kernel(const float* __restrict__ dw1, const float* __restrict__ db1,
                                        float* __restrict__ dw2, float* __restrict__ db, int kh, int kw, int Cin,
                                        int Cout, int pw) {
    const long long total = (long long)kh * kw * Cin * Cout;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        int co = (int)(i % Cout);
        long long r = i / Cout;
        int ci = (int)(r % Cin); r /= Cin;
        int q = (int)(r % kw); r /= kw;
        int h = (int)r;
        float s = 0.f;
#pragma unroll
        for (int wo = 0; wo < 2; ++wo) {
            int wi = q + wo - pw;
            if (wi >= 0 && wi < 2) s += dw1[((((size_t)h * 2 + wi) * Cin + ci) * 2 + wo) * Cout + co];
        }
        dw2[i] = s;
    }
    if (db != nullptr && i < Cout) db[i] = db1[i] + db1[Cout + i];
}

static int check_geom(ConvGeom& g, int B, int L, int Cin, int Lout, int Cout, int k, int s, int p, int up) {
    GN_REQUIRE(B >= 0 && L > 0 && Cin > 0 && Lout > 0 && Cout > 0, "non-positive dimension");
    GN_REQUIRE(k > 0 && s > 0 && p >= 0 && p < k, "bad kernel/stride/padding");
    GN_REQUIRE(up == 1 || up == 2, "up must be 1 or 2");
    GN_REQUIRE(L % up == 0, "L must be a multiple of up");
    GN_REQUIRE((long long)(Lout - 1) * s + k - p <= (long long)L + (k - 1), "Lout too large for L");
    GN_REQUIRE((long long)B * (L > Lout ? L : Lout) < (1LL << 31), "B*L exceeds 2^31");
    g = ConvGeom{B, L, L / up, Cin, Lout, Cout, k, s, p, up};
    return GN_OK;
}

}  // namespace gn

using namespace gn;

extern "C" int gn_conv1d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L, int Cin,
                                 int Lout, int Cout, int k, int stride, int pad_left, int up, int act,
                                 float act_param, void* stream) {
    GN_REQUIRE(x && w && y, "null pointer");
    ConvGeom g;
    int rc = check_geom(g, B, L, Cin, Lout, Cout, k, stride, pad_left, up);
    if (rc != GN_OK) return rc;
    if (B == 0) return GN_OK;
    ConvFwdA fa{x, g};
    RowMajorB fb{w, Cout};
    BiasActStore epi{y, bias, Cout, act, act_param};
    if (Cout <= 4) return launch_gemv_rows(fa, fb, epi, B * Lout, k * Cin, Cout, as_stream(stream));
    return launch_gemm_simt<true, false>(fa, fb, epi, B * Lout, Cout, k * Cin, 1, as_stream(stream));
}

extern "C" int gn_conv1d_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout,
                                   int Cout, int k, int stride, int pad_left, int up, void* stream) {
    GN_REQUIRE(dy && w && dx, "null pointer");
    ConvGeom g;
    int rc = check_geom(g, B, L, Cin, Lout, Cout, k, stride, pad_left, up);
    if (rc != GN_OK) return rc;
    if (B == 0) return GN_OK;
    ConvDgradA fa{dy, g};
    ConvDgradB fb{w, g};
    PlainStore epi{dx, Cin};
    // a convolution that reads <= 4 channels (the first layer of a network when the gradient flows on through it, e.g. the
    // generator step through the discriminator of 2_model_version): one warp per input position instead of a 128-wide tile
    if (Cin <= 4) return launch_gemv_rows(fa, fb, epi, B * g.Lp, up * k * Cout, Cin, as_stream(stream));
    return launch_gemm_simt<true, true>(fa, fb, epi, B * g.Lp, Cin, up * k * Cout, 1, as_stream(stream));
}

extern "C" int gn_conv1d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin,
                                   int Lout, int Cout, int k, int stride, int pad_left, int up, void* stream) {
    GN_REQUIRE(x && dy && dw, "null pointer");
    ConvGeom g;
    int rc = check_geom(g, B, L, Cin, Lout, Cout, k, stride, pad_left, up);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const int Mp = k * Cin, Kp = B * Lout;
    if (Cout <= 4) {
        ConvWgradA_T fa{ConvWgradA{x, g}};
        RowMajorA fd{dy, Cout};
        rc = launch_colreduce_small(fa, fd, dw, Kp, Mp, Cout, st);
    } else {
        ConvWgradA fa{x, g};
        RowMajorB fb{dy, Cout};
        int splits = pick_splits(Mp, Cout, Kp);
        if (splits > 1) {
            cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Mp * Cout, st);
            rc = launch_gemm_simt<false, false>(fa, fb, AtomicStore{dw, Cout}, Mp, Cout, Kp, splits, st);
        } else {
            rc = launch_gemm_simt<false, false>(fa, fb, PlainStore{dw, Cout}, Mp, Cout, Kp, 1, st);
        }
    }
    if (rc != GN_OK) return rc;
    if (db != nullptr) return launch_colsum(dy, (long long)Kp, Cout, db, st);
    return GN_OK;
}

extern "C" int gn_conv2d_w2_pack_f32(const float* w2, const float* b, float* w1, float* b1, int kh, int kw, int Cin,
                                     int Cout, int pw, void* stream) {
    GN_REQUIRE(w2 && w1, "null pointer");
    GN_REQUIRE(b1 == nullptr || b != nullptr, "b1 given without b");
    GN_REQUIRE(kh > 0 && kw > 0 && Cin > 0 && Cout > 0 && pw >= 0, "bad dimension");
    long long total = (long long)kh * 4 * Cin * Cout;
    if (total < 2LL * Cout) total = 2LL * Cout;
    conv2d_w2_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(w2, b, w1, b1, kh, kw, Cin,
                                                                                         Cout, pw);
    return cuda_status("conv2d_w2_pack_kernel");
}

extern "C" int gn_conv2d_w2_unpack_f32(const float* dw1, const float* db1, float* dw2, float* db, int kh, int kw,
                                       int Cin, int Cout, int pw, void* stream) {
    GN_REQUIRE(dw1 && dw2, "null pointer");
    GN_REQUIRE(db == nullptr || db1 != nullptr, "db given without db1");
    GN_REQUIRE(kh > 0 && kw > 0 && Cin > 0 && Cout > 0 && pw >= 0, "bad dimension");
    long long total = (long long)kh * kw * Cin * Cout;
    if (total < Cout) total = Cout;
    conv2d_w2_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(dw1, db1, dw2, db, kh, kw,
                                                                                           Cin, Cout, pw);
    return cuda_status("conv2d_w2_unpack_kernel");
}

extern "C" int gn_dense_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int K, int N,
                                int act, float act_param, void* stream) {
    GN_REQUIRE(x && w && y, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N > 0, "bad dimension");
    if (M == 0) return GN_OK;
    RowMajorA fa{x, K};
    RowMajorB fb{w, N};
    BiasActStore epi{y, bias, N, act, act_param};
    cudaStream_t st = as_stream(stream);
    if (N <= 4) return launch_gemv_rows(fa, fb, epi, M, K, N, st);
    // few output tiles over a long contraction (e.g. Flatten(204700) -> Dense(25), train_on_wvf_version/nn.py:90):
    // split K across CTAs with atomic accumulation, then bias + activation in place
    const int splits = pick_splits(M, N, K);
    if (splits > 1) {
        cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, st);
        int rc = launch_gemm_simt<true, false>(fa, fb, AtomicStore{y, N}, M, N, K, splits, st);
        if (rc != GN_OK) return rc;
        const long long tot = (long long)M * N;
        bias_act_inplace_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(y, bias, tot, N, act, act_param);
        return cuda_status("bias_act_inplace_kernel");
    }
    return launch_gemm_simt<true, false>(fa, fb, epi, M, N, K, 1, st);
}

extern "C" int gn_dense_dgrad_f32(const float* dy, const float* w, float* dx, int M, int K, int N, void* stream) {
    GN_REQUIRE(dy && w && dx, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N > 0, "bad dimension");
    if (M == 0) return GN_OK;
    cudaStream_t st = as_stream(stream);
    if (N <= 4) {
        long long total = (long long)M * K;
        unsigned grid = (unsigned)((total + 255) / 256 < 16LL * num_sms() ? (total + 255) / 256 : 16LL * num_sms());
        switch (N) {
            case 1: outer_small_kernel<1><<<grid, 256, 0, st>>>(dy, w, dx, M, K); break;
            case 2: outer_small_kernel<2><<<grid, 256, 0, st>>>(dy, w, dx, M, K); break;
            case 3: outer_small_kernel<3><<<grid, 256, 0, st>>>(dy, w, dx, M, K); break;
            default: outer_small_kernel<4><<<grid, 256, 0, st>>>(dy, w, dx, M, K); break;
        }
        return cuda_status("outer_small_kernel");
    }
    RowMajorA fa{dy, N};
    TransB fb{w, N};  // B(n, k) = w[k*N + n]
    return launch_gemm_simt<true, true>(fa, fb, PlainStore{dx, K}, M, K, N, 1, st);
}

extern "C" int gn_dense_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int M, int K, int N,
                                  void* stream) {
    GN_REQUIRE(x && dy && dw, "null pointer");
    GN_REQUIRE(M >= 0 && K > 0 && N > 0, "bad dimension");
    cudaStream_t st = as_stream(stream);
    int rc;
    if (N <= 4) {
        rc = launch_colreduce_small(RowMajorA{x, K}, RowMajorA{dy, N}, dw, M, K, N, st);
    } else {
        ColMajorA fa{x, K};  // A'(k, m) = x[m*K + k]
        RowMajorB fb{dy, N};
        int splits = pick_splits(K, N, M);
        if (splits > 1) {
            cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * N, st);
            rc = launch_gemm_simt<false, false>(fa, fb, AtomicStore{dw, N}, K, N, M, splits, st);
        } else {
            rc = launch_gemm_simt<false, false>(fa, fb, PlainStore{dw, N}, K, N, M, 1, st);
        }
    }
    if (rc != GN_OK) return rc;
    if (db != nullptr) return launch_colsum(dy, (long long)M, N, db, st);
    return GN_OK;
}


// ---- Conv2DTranspose with a (1, kw) kernel, unit strides, 'valid' (2_model_version generators) -----------------
// y[w, co] = sum_{j, ci} x[w - j, ci] K[0, j, co, ci]  is the Conv1D  y[w] = sum_t x[w + t - (kw-1)] W1[t]  with
// W1[t, ci, co] = K[0, kw-1-t, co, ci], left padding kw-1 and Lout = L + kw - 1, so the layer runs on the Conv1D
// kernels; this helper converts weights (and, with the roles of the two inner axes swapped, weight gradients)
// between the Keras layout (kw, Cout, Cin) and the Conv1D layout (kw, Cin, Cout):  out[t, a, b] = in[k-1-t, b, a].
namespace gn {
__global__ void __launch_bounds__(256) flip_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int k,
                                                             int A, int Bn) {
    const long long n = (long long)k * A * Bn;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % Bn);
        const long long r = i / Bn;
        const int a = (int)(r % A);
        const int t = (int)(r / A);
        out[i] = in[((size_t)(k - 1 - t) * Bn + b) * A + a];
    }
}
}  // namespace gn

extern "C" int gn_flip_transpose_f32(const float* in, float* out, int k, int A, int B, void* stream) {
    GN_REQUIRE(in && out && k > 0 && A > 0 && B > 0, "null pointer or bad size");
    const long long n = (long long)k * A * B;
    unsigned grid = (unsigned)((n + 255) / 256 < 8LL * num_sms() ? (n + 255) / 256 : 8LL * num_sms());
    gn::flip_transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, out, k, A, B);
    return cuda_status("flip_transpose_kernel");
}
