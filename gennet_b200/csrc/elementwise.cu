// Memory-bound layer kernels of the hot path: BatchNormalization (two-pass statistics, data-parallel
// friendly), activations, dropout family, up/down-sampling, MyLayer, losses, Adam/SGD, counter-based RNG.
// All are HBM-bound streaming kernels: grid-stride loops sized to a multiple of the SM count.
// Keras semantics restated from bbhMahoGANy.py:164-188,212-295,1101-1119 and burstMahoGANy.py:100-125.
#include "gn_common.cuh"
#include "philox.cuh"

namespace gn {

static inline unsigned ew_grid(long long n, int threads = 256) {
    long long b = (n + threads - 1) / threads;
    long long cap = 16LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}
#define GN_EW_LOOP(i, n)                                                                       \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n);              \
         i += (long long)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- BatchNorm
// sums[c] += sum_r x[r,c] ; sums[C+c] += sum_r (x[r,c]-shift[c])^2      (block = 32 channels x 8 row lanes)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long rows, int C,
                                                       long long rows_per_split, const float* __restrict__ shift,
                                                       double* __restrict__ sums) {
    __shared__ double s0[8][33], s1[8][33];
    const int cl = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
    double a = 0.0, q = 0.0;
    if (c < C) {
        const float sh = shift ? shift[c] : 0.f;
        for (long long r = r0 + ry; r < r1; r += 8) {
            float v = x[r * C + c];
            float d = v - sh;
            a += (double)v;
            q += (double)d * (double)d;
        }
    }
    s0[ry][cl] = a;
    s1[ry][cl] = q;
    __syncthreads();
    if (ry == 0 && c < C) {
        double ta = 0.0, tq = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ta += s0[i][cl]; tq += s1[i][cl]; }
        atomicAdd(&sums[c], ta);
        atomicAdd(&sums[C + c], tq);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum_x, const double* __restrict__ sum_sq, double n,
                                   int C, float eps, float momentum, float* __restrict__ stats,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var, int phase,
                                   float* __restrict__ biased, double debias) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (phase == 0) {
        stats[c] = (float)(sum_x[c] / n);
        return;
    }
    double var = sum_sq[c] / n;
    if (var < 0.0) var = 0.0;
    stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (moving_mean != nullptr) {
        double m = (double)momentum;
        // Keras 2.2.4 feeds the moving variance the "sample variance" var*n/(n-(1+eps))
        double sv = var * (n / (n - (1.0 + (double)eps)));
        if (biased != nullptr) {
            // K.moving_average_update of Keras 2.2.4 on TF 1.12 = assign_moving_average(zero_debias=True): the
            // exponential average runs on a zero-initialised shadow accumulator and the moving statistic is
            // OVERWRITTEN with accumulator / (1 - momentum^step)
            const float bm = (float)((double)biased[c] * m + (double)stats[c] * (1.0 - m));
            const float bv = (float)((double)biased[C + c] * m + sv * (1.0 - m));
            biased[c] = bm;
            biased[C + c] = bv;
            moving_mean[c] = (float)((double)bm * debias);
            moving_var[c] = (float)((double)bv * debias);
        } else {
            moving_mean[c] = (float)((double)moving_mean[c] * m + (double)stats[c] * (1.0 - m));
            moving_var[c] = (float)((double)moving_var[c] * m + sv * (1.0 - m));
        }
    }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                       const float* __restrict__ inv, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ y,
                                                       long long n, int C, float eps, int use_var) {
    GN_EW_LOOP(i, n) {
        int c;
        (void)fast_div(i, C, c);
        float is = use_var ? 1.0f / sqrtf(inv[c] + eps) : inv[c];
        y[i] = (x[i] - mean[c]) * is * gamma[c] + beta[c];
    }
}

__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ stats, long long rows, int C,
                                                          long long rows_per_split, double* __restrict__ sums) {
    __shared__ double s0[8][33], s1[8][33];
    const int cl = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
    double a = 0.0, q = 0.0;
    if (c < C) {
        const float mu = stats[c], is = stats[C + c];
        for (long long r = r0 + ry; r < r1; r += 8) {
            float g = dy[r * C + c];
            float xh = (x[r * C + c] - mu) * is;
            a += (double)g;
            q += (double)g * (double)xh;
        }
    }
    s0[ry][cl] = a;
    s1[ry][cl] = q;
    __syncthreads();
    if (ry == 0 && c < C) {
        double ta = 0.0, tq = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ta += s0[i][cl]; tq += s1[i][cl]; }
        atomicAdd(&sums[c], ta);
        atomicAdd(&sums[C + c], tq);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ stats,
                                                           const float* __restrict__ gamma,
                                                           const double* __restrict__ sums, double n_total,
                                                           float* __restrict__ dx, long long n, int C) {
    GN_EW_LOOP(i, n) {
        int c;
        (void)fast_div(i, C, c);
        float mu = stats[c], is = stats[C + c];
        float xh = (x[i] - mu) * is;
        float sdy = (float)(sums[c] / n_total), sdx = (float)(sums[C + c] / n_total);
        dx[i] = gamma[c] * is * (dy[i] - sdy - xh * sdx);
    }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)sums[c];
    dgamma[c] = (float)sums[C + c];
}

static void col_split(long long rows, int C, int& cb, long long& per, long long& splits) {
    cb = (C + 31) / 32;
    splits = (4LL * num_sms() + cb - 1) / cb;
    if (splits > (rows + 63) / 64) splits = (rows + 63) / 64;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    per = (rows + splits - 1) / splits;
    splits = (rows + per - 1) / per;
    if (splits < 1) splits = 1;
}

// ---------------------------------------------------------------- activations / noise layers
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                      int act, float p) {
    GN_EW_LOOP(i, n) y[i] = act_fwd(x[i], act, p);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                      float* __restrict__ dx, long long n, int act, float p) {
    GN_EW_LOOP(i, n) dx[i] = dy[i] * act_bwd_from_y(y[i], act, p);
}
__device__ __forceinline__ float noise_factor(float r, int kind, float rate) {
    if (kind == GN_NOISE_DROPOUT) return r / (1.f - rate);
    return 1.f + r * sqrtf(rate / (1.f - rate));
}
__global__ void __launch_bounds__(256) noise_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                        float* __restrict__ y, long long n, int kind, float rate) {
    if (kind == GN_NOISE_GNOISE) {
        GN_EW_LOOP(i, n) y[i] = fmaf(r[i], rate, x[i]);
    } else {
        GN_EW_LOOP(i, n) y[i] = x[i] * noise_factor(r[i], kind, rate);
    }
}
__global__ void __launch_bounds__(256) noise_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ r,
                                                        float* __restrict__ dx, long long n, int kind, float rate) {
    if (kind == GN_NOISE_GNOISE) {
        GN_EW_LOOP(i, n) dx[i] = dy[i];
    } else {
        GN_EW_LOOP(i, n) dx[i] = dy[i] * noise_factor(r[i], kind, rate);
    }
}

// mode 0: uniform(lo,hi)  1: normal(mean,std)  2: keep mask with P(keep)=1-p0
__global__ void __launch_bounds__(256) rng_fill_kernel(float* __restrict__ out, long long n, int mode, float p0,
                                                       float p1, unsigned long long seed, unsigned long long blk0,
                                                       unsigned stream) {
    const long long nblk = (n + 3) / 4;
    GN_EW_LOOP(b, nblk) {
        uint4 r = philox_flat(seed, blk0 + (unsigned long long)b, stream);
        float v[4];
        if (mode == 1) {
            float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
            v[0] = fmaf(g0.x, p1, p0); v[1] = fmaf(g0.y, p1, p0);
            v[2] = fmaf(g1.x, p1, p0); v[3] = fmaf(g1.y, p1, p0);
        } else {
            float u[4] = {u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (mode == 0) ? fmaf(u[j], p1 - p0, p0) : (u[j] >= p0 ? 1.f : 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * b + j < n) out[4 * b + j] = v[j];
    }
}

// ---------------------------------------------------------------- resampling / plumbing
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int L,
                                                           int C, int size, long long n_out) {
    GN_EW_LOOP(i, n_out) {
        int c;
        long long r = fast_div(i, C, c);
        int lo = (int)(r % ((long long)L * size));
        long long b = r / ((long long)L * size);
        y[i] = x[(b * L + lo / size) * C + c];
    }
}
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                           int L, int C, int size, long long n_in) {
    GN_EW_LOOP(i, n_in) {
        int c;
        long long r = fast_div(i, C, c);
        int l = (int)(r % L);
        long long b = r / L;
        float s = 0.f;
        for (int u = 0; u < size; ++u) s += dy[((b * L + l) * size + u) * C + c];
        dx[i] = s;
    }
}
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int L,
                                                          int C, int pool, int Lo, long long n_out) {
    GN_EW_LOOP(i, n_out) {
        int c;
        long long r = fast_div(i, C, c);
        int lo = (int)(r % Lo);
        long long b = r / Lo;
        float m = x[(b * L + (long long)lo * pool) * C + c];
        for (int u = 1; u < pool; ++u) m = fmaxf(m, x[(b * L + (long long)lo * pool + u) * C + c]);
        y[i] = m;
    }
}
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ dy, float* __restrict__ dx,
                                                          int L, int C, int pool, int Lo, long long n_in) {
    GN_EW_LOOP(i, n_in) {
        int c;
        long long r = fast_div(i, C, c);
        int l = (int)(r % L);
        long long b = r / L;
        int lo = l / pool;
        float g = 0.f;
        if (lo < Lo) {
            float m = y[(b * Lo + lo) * C + c];
            // gradient goes to the first element of the window that attains the max
            bool first = (x[i] == m);
            for (int u = 0; u < l - lo * pool && first; ++u)
                if (x[(b * L + (long long)lo * pool + u) * C + c] == m) first = false;
            if (first) g = dy[(b * Lo + lo) * C + c];
        }
        dx[i] = g;
    }
}
__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ a, const float* __restrict__ b, float alpha,
                                                   long long n) {
    GN_EW_LOOP(i, n) a[i] = fmaf(alpha, b[i], a[i]);
}
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                          float* __restrict__ out, long long row_len, long long n) {
    GN_EW_LOOP(i, n) {
        long long r = i / row_len;
        out[i] = src[(long long)idx[r] * row_len + (i - r * row_len)];
    }
}
__global__ void __launch_bounds__(256) stack_residual_fwd_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ cst, float* __restrict__ y,
                                                                 int L, long long n) {
    GN_EW_LOOP(i, n) {
        float v = x[i];
        reinterpret_cast<float2*>(y)[i] = make_float2(v, cst[i % L] - v);
    }
}
__global__ void __launch_bounds__(256) stack_pair_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ y, long long n) {
    GN_EW_LOOP(i, n) { reinterpret_cast<float2*>(y)[i] = make_float2(a[i], b[i]); }
}
__global__ void __launch_bounds__(256) stack_residual_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                                 long long n) {
    GN_EW_LOOP(i, n) {
        float2 g = reinterpret_cast<const float2*>(dy)[i];
        dx[i] = g.x - g.y;
    }
}
__global__ void __launch_bounds__(256) residual_moments_fwd_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ cst, int L, long long n,
                                                                   double* __restrict__ sums) {
    __shared__ double sm[32];
    double a = 0.0, q = 0.0;
    GN_EW_LOOP(i, n) {
        double d = (double)cst[i % L] - (double)x[i];
        a += d;
        q += d * d;
    }
    double ta = block_sum(a, sm);
    double tq = block_sum(q, sm);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], ta);
        atomicAdd(&sums[1], tq);
    }
}
__global__ void __launch_bounds__(256) residual_moments_bwd_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ cst,
                                                                   const float* __restrict__ dout,
                                                                   float* __restrict__ dx, int L, long long n,
                                                                   double n_total) {
    const float g0 = dout[0], g1 = dout[1];
    const float inv = (float)(1.0 / n_total);
    GN_EW_LOOP(i, n) {
        float d = cst[i % L] - x[i];
        dx[i] = -(g0 + 2.f * g1 * d) * inv;
    }
}

// ---------------------------------------------------------------- losses
// one thread per batch row; out[0] += sum_b loss_b, out[1] += sum_b metric_b
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                   float* __restrict__ out, float* __restrict__ dpred, int B, int D,
                                                   int kind, float param, float inv_batch, int pred_is_vec,
                                                   int metric_kind) {
    __shared__ float sm[32];
    const float EPSK = 1e-7f;
    float loss = 0.f, hit = 0.f;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        int am_p = 0, am_t = 0;
        float best_p = -INFINITY, best_t = -INFINITY;
        float hits = 0.f;
        for (int d = 0; d < D; ++d) {
            float p = pred_is_vec ? pred[d] : pred[(size_t)b * D + d];
            float t = target[(size_t)b * D + d];
            float l, g;
            if (kind == GN_LOSS_BCE) {
                float pc = fminf(fmaxf(p, EPSK), 1.f - EPSK);
                float x = logf(pc / (1.f - pc));
                l = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
                g = (p >= EPSK && p <= 1.f - EPSK) ? (pc - t) / (pc * (1.f - pc)) : 0.f;
                l /= D; g /= D;
            } else if (kind == GN_LOSS_MSE) {
                float e = p - t;
                l = e * e / D;
                g = 2.f * e / D;
            } else {  // chi-square: sum over D of (t-p)^2/n_sig^2
                float e = p - t;
                l = e * e / (param * param);
                g = 2.f * e / (param * param);
            }
            loss += l;
            g *= inv_batch;
            if (dpred != nullptr) {
                if (pred_is_vec) atomicAdd(&dpred[d], g);
                else dpred[(size_t)b * D + d] = g;
            }
            hits += (t == rintf(p)) ? 1.f : 0.f;
            if (p > best_p) { best_p = p; am_p = d; }
            if (t > best_t) { best_t = t; am_t = d; }
        }
        hit = (metric_kind == 0) ? hits / D : (am_p == am_t ? 1.f : 0.f);
    }
    float tl = block_sum(loss, sm);
    float th = block_sum(hit, sm);
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], tl);
        atomicAdd(&out[1], th);
    }
}

// ---------------------------------------------------------------- optimizers
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float lr_t, float b1, float b2, float eps, float gs) {
    GN_EW_LOOP(i, n) {
        float gi = g[i] * gs;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
}
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n,
                                                  float lr, float gs) {
    GN_EW_LOOP(i, n) p[i] = p[i] - lr * (g[i] * gs);
}

}  // namespace gn

using namespace gn;

extern "C" int gn_bn_stats_f32(const float* x, long long rows, int C, double* sums, const float* shift, void* stream) {
    GN_REQUIRE(x && sums && rows >= 0 && C > 0, "null pointer or bad size");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st);
    if (rows == 0) return GN_OK;
    int cb; long long per, splits;
    col_split(rows, C, cb, per, splits);
    bn_stats_kernel<<<dim3(cb, (unsigned)splits), 256, 0, st>>>(x, rows, C, per, shift, sums);
    return cuda_status("bn_stats_kernel");
}

extern "C" int gn_bn_finalize_f32(const double* sum_x, const double* sum_sq, double n_total, int C, float eps,
                                  float momentum, float* stats, float* moving_mean, float* moving_var, int phase,
                                  float* biased, double debias, void* stream) {
    GN_REQUIRE(stats && C > 0 && n_total > 0, "null pointer or bad size");
    GN_REQUIRE(phase == 0 ? sum_x != nullptr : sum_sq != nullptr, "missing sums for this phase");
    GN_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "moving_mean/var must both be given or both NULL");
    bn_finalize_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(sum_x, sum_sq, n_total, C, eps, momentum, stats,
                                                                       moving_mean, moving_var, phase, biased, debias);
    return cuda_status("bn_finalize_kernel");
}

extern "C" int gn_bn_apply_f32(const float* x, const float* mean, const float* inv, const float* gamma,
                               const float* beta, float* y, long long rows, int C, float eps, int use_var,
                               void* stream) {
    GN_REQUIRE(x && mean && inv && gamma && beta && y && rows >= 0 && C > 0, "null pointer or bad size");
    long long n = rows * C;
    if (n == 0) return GN_OK;
    bn_apply_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, mean, inv, gamma, beta, y, n, C, eps, use_var);
    return cuda_status("bn_apply_kernel");
}

extern "C" int gn_bn_bwd_sums_f32(const float* x, const float* dy, const float* stats, long long rows, int C,
                                  double* sums, void* stream) {
    GN_REQUIRE(x && dy && stats && sums && rows >= 0 && C > 0, "null pointer or bad size");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st);
    if (rows == 0) return GN_OK;
    int cb; long long per, splits;
    col_split(rows, C, cb, per, splits);
    bn_bwd_sums_kernel<<<dim3(cb, (unsigned)splits), 256, 0, st>>>(x, dy, stats, rows, C, per, sums);
    return cuda_status("bn_bwd_sums_kernel");
}

extern "C" int gn_bn_bwd_apply_f32(const float* x, const float* dy, const float* stats, const float* gamma,
                                   const double* sums, double n_total, float* dx, float* dgamma, float* dbeta,
                                   long long rows, int C, void* stream) {
    GN_REQUIRE(x && dy && stats && gamma && sums && dx && rows >= 0 && C > 0 && n_total > 0, "null pointer or bad size");
    cudaStream_t st = as_stream(stream);
    long long n = rows * C;
    if (n > 0) bn_bwd_apply_kernel<<<ew_grid(n), 256, 0, st>>>(x, dy, stats, gamma, sums, n_total, dx, n, C);
    if (dgamma != nullptr && dbeta != nullptr) bn_param_grads_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, dgamma, dbeta, C);
    return cuda_status("bn_bwd_apply_kernel");
}

extern "C" int gn_act_fwd_f32(const float* x, float* y, long long n, int act, float param, void* stream) {
    GN_REQUIRE(x && y && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    act_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, n, act, param);
    return cuda_status("act_fwd_kernel");
}
extern "C" int gn_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int act, float param,
                              void* stream) {
    GN_REQUIRE(dy && y && dx && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    act_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, y, dx, n, act, param);
    return cuda_status("act_bwd_kernel");
}
extern "C" int gn_noise_fwd_f32(const float* x, const float* r, float* y, long long n, int kind, float rate,
                                void* stream) {
    GN_REQUIRE(x && r && y && n >= 0, "null pointer or n < 0");
    GN_REQUIRE(kind >= 0 && kind <= 2, "unknown noise kind");
    GN_REQUIRE(kind == GN_NOISE_GNOISE || (rate >= 0.f && rate < 1.f), "rate must be in [0,1)");
    if (n == 0) return GN_OK;
    noise_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, r, y, n, kind, rate);
    return cuda_status("noise_fwd_kernel");
}
extern "C" int gn_noise_bwd_f32(const float* dy, const float* r, float* dx, long long n, int kind, float rate,
                                void* stream) {
    GN_REQUIRE(dy && r && dx && n >= 0, "null pointer or n < 0");
    GN_REQUIRE(kind >= 0 && kind <= 2, "unknown noise kind");
    if (n == 0) return GN_OK;
    noise_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, r, dx, n, kind, rate);
    return cuda_status("noise_bwd_kernel");
}

static int rng_fill(float* r, long long n, int mode, float p0, float p1, uint64_t seed, uint64_t offset,
                    unsigned stream_id, void* stream) {
    GN_REQUIRE(r && n >= 0, "null pointer or n < 0");
    GN_REQUIRE(offset % 4 == 0, "offset must be a multiple of 4");
    if (n == 0) return GN_OK;
    rng_fill_kernel<<<ew_grid((n + 3) / 4), 256, 0, as_stream(stream)>>>(r, n, mode, p0, p1, seed, offset / 4,
                                                                         stream_id);
    return cuda_status("rng_fill_kernel");
}
extern "C" int gn_noise_draw_f32(float* r, long long n, int kind, float rate, uint64_t seed, uint64_t offset,
                                 void* stream) {
    GN_REQUIRE(kind >= 0 && kind <= 2, "unknown noise kind");
    if (kind == GN_NOISE_DROPOUT) return rng_fill(r, n, 2, rate, 0.f, seed, offset, 3u, stream);
    return rng_fill(r, n, 1, 0.f, 1.f, seed, offset, 3u, stream);
}
extern "C" int gn_uniform_f32(float* r, long long n, float lo, float hi, uint64_t seed, uint64_t offset,
                              void* stream) {
    return rng_fill(r, n, 0, lo, hi, seed, offset, 1u, stream);
}
extern "C" int gn_normal_f32(float* r, long long n, float mean, float std, uint64_t seed, uint64_t offset,
                             void* stream) {
    return rng_fill(r, n, 1, mean, std, seed, offset, 2u, stream);
}

extern "C" int gn_upsample1d_fwd_f32(const float* x, float* y, int B, int L, int C, int size, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && L > 0 && C > 0 && size > 0, "null pointer or bad size");
    long long n = (long long)B * L * size * C;
    if (n == 0) return GN_OK;
    upsample_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, L, C, size, n);
    return cuda_status("upsample_fwd_kernel");
}
extern "C" int gn_upsample1d_bwd_f32(const float* dy, float* dx, int B, int L, int C, int size, void* stream) {
    GN_REQUIRE(dy && dx && B >= 0 && L > 0 && C > 0 && size > 0, "null pointer or bad size");
    long long n = (long long)B * L * C;
    if (n == 0) return GN_OK;
    upsample_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, dx, L, C, size, n);
    return cuda_status("upsample_bwd_kernel");
}
extern "C" int gn_maxpool1d_fwd_f32(const float* x, float* y, int B, int L, int C, int pool, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && L > 0 && C > 0 && pool > 0 && L >= pool, "null pointer or bad size");
    int Lo = L / pool;
    long long n = (long long)B * Lo * C;
    if (n == 0) return GN_OK;
    maxpool_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, L, C, pool, Lo, n);
    return cuda_status("maxpool_fwd_kernel");
}
extern "C" int gn_maxpool1d_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int B, int L, int C,
                                    int pool, void* stream) {
    GN_REQUIRE(x && y && dy && dx && B >= 0 && L > 0 && C > 0 && pool > 0 && L >= pool, "null pointer or bad size");
    long long n = (long long)B * L * C;
    if (n == 0) return GN_OK;
    maxpool_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, dy, dx, L, C, pool, L / pool, n);
    return cuda_status("maxpool_bwd_kernel");
}
namespace gn {
// ---- layers of the 2_model_version networks (no_weight_code/subtract_model.py:330-390, weight_version/subtract_model.py:215-222)
// GlobalAveragePooling1D/2D: y[b,c] = mean_l x[b,l,c]; block = 32 channels x 8 row lanes of one sample
__global__ void __launch_bounds__(256) gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int L, int C) {
    __shared__ float sm[8][33];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    float s = 0.f;
    if (c < C)
        for (int l = ry; l < L; l += 8) s += x[((size_t)b * L + l) * C + c];
    sm[ry][threadIdx.x & 31] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x & 31];
        y[(size_t)b * C + c] = t / (float)L;
    }
}
__global__ void __launch_bounds__(256) gap_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int L, int C,
                                                      long long n) {
    const float inv = 1.f / (float)L;
    GN_EW_LOOP(i, n) {
        int c;
        const long long r = fast_div(i, C, c);
        dx[i] = dy[(r / L) * C + c] * inv;
    }
}
// batched transpose y[b, c, r] = x[b, r, c] through 32 x 32 shared-memory tiles (BatchNormalization(axis=1): the
// normalised axis is moved innermost, normalised by the channels-last kernels and moved back)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
    __shared__ float tile[32][33];
    const size_t base = (size_t)blockIdx.z * R * C;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < R && c0 + tx < C) tile[i][tx] = x[base + (size_t)(r0 + i) * C + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < C && r0 + tx < R) y[base + (size_t)(c0 + i) * R + r0 + tx] = tile[tx][i];
}
// keras.regularizers.l1 / l2 on a tensor x: loss += l1 * sum |x| + l2 * sum x^2 (double, accumulated) and, when g is
// given, g += l1 * sign(x) + 2 * l2 * x  (kernel_regularizer on the weight gradient, activity_regularizer on dL/dy)
__global__ void __launch_bounds__(256) reg_terms_kernel(const float* __restrict__ x, float* __restrict__ g, long long n,
                                                        float l1, float l2, double* __restrict__ loss) {
    __shared__ double sm[32];
    double acc = 0.0;
    GN_EW_LOOP(i, n) {
        const float v = x[i];
        acc += (double)(l1 * fabsf(v)) + (double)l2 * (double)v * (double)v;
        if (g != nullptr) g[i] += l1 * (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) + 2.f * l2 * v;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0 && loss != nullptr) atomicAdd(loss, acc);
}
}  // namespace gn

extern "C" int gn_gap_fwd_f32(const float* x, float* y, int B, int L, int C, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && L > 0 && C > 0 && B <= 65535, "null pointer or bad size");
    if (B == 0) return GN_OK;
    gn::gap_fwd_kernel<<<dim3((C + 31) / 32, B), 256, 0, as_stream(stream)>>>(x, y, L, C);
    return cuda_status("gap_fwd_kernel");
}
extern "C" int gn_gap_bwd_f32(const float* dy, float* dx, int B, int L, int C, void* stream) {
    GN_REQUIRE(dy && dx && B >= 0 && L > 0 && C > 0, "null pointer or bad size");
    const long long n = (long long)B * L * C;
    if (n == 0) return GN_OK;
    gn::gap_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, dx, L, C, n);
    return cuda_status("gap_bwd_kernel");
}
extern "C" int gn_transpose_f32(const float* x, float* y, int B, int R, int C, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && R > 0 && C > 0 && B <= 65535 && (R + 31) / 32 <= 65535, "null pointer or bad size");
    if (B == 0) return GN_OK;
    gn::transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32, B), 256, 0, as_stream(stream)>>>(x, y, R, C);
    return cuda_status("transpose_kernel");
}
extern "C" int gn_reg_terms_f32(const float* x, float* g, long long n, float l1, float l2, double* loss, void* stream) {
    GN_REQUIRE(x && n >= 0 && l1 >= 0.f && l2 >= 0.f, "null pointer, n < 0 or negative coefficient");
    if (n == 0) return GN_OK;
    gn::reg_terms_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, g, n, l1, l2, loss);
    return cuda_status("reg_terms_kernel");
}
extern "C" int gn_axpy_f32(float* a, const float* b, float alpha, long long n, void* stream) {
    GN_REQUIRE(a && b && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    axpy_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a, b, alpha, n);
    return cuda_status("axpy_kernel");
}
// waveform ingest (train_on_wvf_version/load_txtwfs.py:47-50,66-69): y = roll(x / max(x), offset) per row.
// One CTA per waveform: block max (plain maximum, as np.max -- not the absolute value), then the rotated store.
namespace gn {
__global__ void __launch_bounds__(256) maxnorm_roll_kernel(const float* __restrict__ x, const int* __restrict__ off,
                                                           float* __restrict__ y, int N) {
    __shared__ float sm[8];
    const float* xr = x + (size_t)blockIdx.x * N;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < N; j += blockDim.x) m = fmaxf(m, xr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    int sh = off ? off[blockIdx.x] % N : 0;
    if (sh < 0) sh += N;
    float* yr = y + (size_t)blockIdx.x * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        int d = j + sh;
        if (d >= N) d -= N;
        yr[d] = xr[j] / m;
    }
}
}  // namespace gn

extern "C" int gn_maxnorm_roll_f32(const float* x, const int* offsets, float* y, int B, int N, void* stream) {
    GN_REQUIRE(x && y && B >= 0 && N > 0, "null pointer or bad size");
    GN_REQUIRE(x != y, "in-place rotation is not supported");
    if (B == 0) return GN_OK;
    gn::maxnorm_roll_kernel<<<B, 256, 0, as_stream(stream)>>>(x, offsets, y, N);
    return cuda_status("maxnorm_roll_kernel");
}

extern "C" int gn_gather_rows_f32(const float* src, const int* idx, float* out, int n, long long row_len,
                                  void* stream) {
    GN_REQUIRE(src && idx && out && n >= 0 && row_len > 0, "null pointer or bad size");
    long long tot = (long long)n * row_len;
    if (tot == 0) return GN_OK;
    gather_rows_kernel<<<ew_grid(tot), 256, 0, as_stream(stream)>>>(src, idx, out, row_len, tot);
    return cuda_status("gather_rows_kernel");
}
extern "C" int gn_stack_residual_fwd_f32(const float* x, const float* cst, float* y, int B, int L, void* stream) {
    GN_REQUIRE(x && cst && y && B >= 0 && L > 0, "null pointer or bad size");
    long long n = (long long)B * L;
    if (n == 0) return GN_OK;
    stack_residual_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, cst, y, L, n);
    return cuda_status("stack_residual_fwd_kernel");
}
extern "C" int gn_stack_pair_f32(const float* a, const float* b, float* y, long long n, void* stream) {
    GN_REQUIRE(a && b && y && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    stack_pair_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a, b, y, n);
    return cuda_status("stack_pair_kernel");
}
extern "C" int gn_stack_residual_bwd_f32(const float* dy, float* dx, int B, int L, void* stream) {
    GN_REQUIRE(dy && dx && B >= 0 && L > 0, "null pointer or bad size");
    long long n = (long long)B * L;
    if (n == 0) return GN_OK;
    stack_residual_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, dx, n);
    return cuda_status("stack_residual_bwd_kernel");
}
extern "C" int gn_residual_moments_fwd_f32(const float* x, const float* cst, double* sums, int B, int L,
                                           void* stream) {
    GN_REQUIRE(x && cst && sums && B >= 0 && L > 0, "null pointer or bad size");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
    long long n = (long long)B * L;
    if (n == 0) return GN_OK;
    residual_moments_fwd_kernel<<<ew_grid(n), 256, 0, st>>>(x, cst, L, n, sums);
    return cuda_status("residual_moments_fwd_kernel");
}
extern "C" int gn_residual_moments_bwd_f32(const float* x, const float* cst, const float* dout, float* dx, int B,
                                           int L, double n_total, void* stream) {
    GN_REQUIRE(x && cst && dout && dx && B >= 0 && L > 0 && n_total > 0, "null pointer or bad size");
    long long n = (long long)B * L;
    if (n == 0) return GN_OK;
    residual_moments_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, cst, dout, dx, L, n, n_total);
    return cuda_status("residual_moments_bwd_kernel");
}

extern "C" int gn_loss_fwd_bwd_f32(const float* pred, const float* target, float* out, float* dpred, int B, int D,
                                   int kind, float param, float inv_batch, int pred_is_vec, int metric_kind,
                                   void* stream) {
    GN_REQUIRE(pred && target && out && B >= 0 && D > 0, "null pointer or bad size");
    GN_REQUIRE(kind >= 0 && kind <= 2, "unknown loss kind");
    GN_REQUIRE(kind != GN_LOSS_CHISQ || param != 0.f, "n_sig must be non-zero");
    if (B == 0) return GN_OK;
    loss_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(pred, target, out, dpred, B, D, kind, param, inv_batch,
                                                              pred_is_vec, metric_kind);
    return cuda_status("loss_kernel");
}

extern "C" int gn_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1,
                                float beta2, float eps, float grad_scale, void* stream) {
    GN_REQUIRE(p && g && m && v && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    adam_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale);
    return cuda_status("adam_kernel");
}
extern "C" int gn_sgd_step_f32(float* p, const float* g, long long n, float lr, float grad_scale, void* stream) {
    GN_REQUIRE(p && g && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    sgd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, n, lr, grad_scale);
    return cuda_status("sgd_kernel");
}
