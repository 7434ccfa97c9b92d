// Evaluation-stage statistics of the GAN loop on the device: the two-dimensional Gaussian kernel density estimate the
// reference builds with scipy.stats.gaussian_kde (bbhMahoGANy.py:787-791) evaluated on the 100 x 100 comparison grid,
// and the sums behind the overlap score beta of overlap_tests (bbhMahoGANy.py:853-870).
#include "gn_common.cuh"

namespace gn {

// pdf[j] = inv_norm * sum_i exp(-1/2 (x_i - p_j)^T C^-1 (x_i - p_j)); c11, c12, c22 carry the 1/2 and log2(e):
// energy*log2e = c11 dx^2 + c12 dx dy + c22 dy^2.  One thread per position, samples staged through shared memory;
// partial sums leave float after every tile of 256 samples.
constexpr int KDE_TILE = 256;
__global__ void __launch_bounds__(128) kde2d_pdf_kernel(const float2* __restrict__ data, int n,
                                                        const float2* __restrict__ pos, int m, float c11, float c12,
                                                        float c22, double inv_norm, float* __restrict__ pdf) {
    __shared__ float2 tile[KDE_TILE];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float2 p = j < m ? pos[j] : make_float2(0.f, 0.f);
    double acc = 0.0;
    for (int base = 0; base < n; base += KDE_TILE) {
        const int cnt = min(KDE_TILE, n - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) tile[i] = data[base + i];
        __syncthreads();
        float part = 0.f;
#pragma unroll 8
        for (int i = 0; i < cnt; ++i) {
            const float dx = tile[i].x - p.x, dy = tile[i].y - p.y;
            part += exp2f(-(c11 * dx * dx + c12 * dx * dy + c22 * dy * dy));
        }
        acc += (double)part;
        __syncthreads();
    }
    if (j < m) pdf[j] = (float)(acc * inv_norm);
}

// out = {sum a*b, sum a*a, sum b*b} in double
__global__ void __launch_bounds__(1024) overlap_sums_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            long long n, double* __restrict__ out) {
    __shared__ double sm[32];
    double sab = 0.0, saa = 0.0, sbb = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = (double)a[i], y = (double)b[i];
        sab += x * y;
        saa += x * x;
        sbb += y * y;
    }
    sab = block_sum(sab, sm);
    __syncthreads();
    saa = block_sum(saa, sm);
    __syncthreads();
    sbb = block_sum(sbb, sm);
    if (threadIdx.x == 0) {
        out[0] = sab;
        out[1] = saa;
        out[2] = sbb;
    }
}

// ---- percentile curves of plot_waveform_est (bbhMahoGANy.py:913-921): for every time sample l the p-th percentiles of
// the n generated waveforms, np.percentile's default linear interpolation: q = p/100 * (n-1), (1-g) a[j] + g a[j+1].
// One block per tile of TL adjacent time samples (a 32-byte sector of every row is used whole): the tile is staged
// column by column in shared memory (padded to a power of two with +inf), each column is sorted by a bitonic network
// run by the whole block, and the requested order statistics are interpolated in double.
__global__ void __launch_bounds__(256) percentile_kernel(const float* __restrict__ x, int n, int L, int n_pad, int TL,
                                                         const float* __restrict__ pcts, int npct, float* __restrict__ out) {
    extern __shared__ float col[];      // TL columns of n_pad floats
    const int l0 = blockIdx.x * TL;
    for (int i = threadIdx.x; i < n_pad * TL; i += blockDim.x) {
        const int r = i / TL, c = i - r * TL;
        col[c * n_pad + r] = (r < n && l0 + c < L) ? x[(size_t)r * L + l0 + c] : __int_as_float(0x7f800000);
    }
    __syncthreads();
    for (int c = 0; c < TL; ++c) {
        float* a = col + c * n_pad;
        for (int size = 2; size <= n_pad; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = threadIdx.x; t < n_pad / 2; t += blockDim.x) {
                    const int i = 2 * t - (t & (stride - 1));      // lower index of the pair
                    const int j = i + stride;
                    const bool up = (i & size) == 0;
                    const float u = a[i], v = a[j];
                    if ((u > v) == up) { a[i] = v; a[j] = u; }
                }
                __syncthreads();
            }
        }
    }
    for (int t = threadIdx.x; t < TL * npct; t += blockDim.x) {
        const int c = t / npct, q = t - c * npct;
        if (l0 + c >= L) continue;
        const double pos = (double)pcts[q] / 100.0 * (double)(n - 1);
        int j = (int)floor(pos);
        if (j < 0) j = 0;
        if (j > n - 1) j = n - 1;
        const int j1 = j + 1 < n ? j + 1 : n - 1;
        const double g = pos - (double)j;
        const double lo = (double)col[c * n_pad + j], hi = (double)col[c * n_pad + j1];
        out[(size_t)q * L + l0 + c] = (float)(lo + (hi - lo) * g);
    }
}

}  // namespace gn

using namespace gn;

extern "C" int gn_percentiles_f32(const float* x, int n, int L, const float* pcts, int npct, float* out, void* stream) {
    GN_REQUIRE(x && pcts && out, "null pointer");
    GN_REQUIRE(n > 0 && L > 0 && npct > 0 && npct <= 64, "needs n > 0, L > 0 and 1..64 percentiles");
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    GN_REQUIRE(n_pad <= 32768, "at most 32768 waveforms per call");
    int TL = 40960 / n_pad;             // <= 160 KB of shared memory
    if (TL > 8) TL = 8;
    if (TL < 1) TL = 1;
    const size_t smem = sizeof(float) * (size_t)n_pad * TL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(percentile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
        attr_set = true;
    }
    percentile_kernel<<<(L + TL - 1) / TL, 256, smem, as_stream(stream)>>>(x, n, L, n_pad, TL, pcts, npct, out);
    return cuda_status("percentile_kernel");
}

extern "C" int gn_kde2d_pdf_f32(const float* data_xy, int n, const float* pos_xy, int m, double a11, double a12,
                                double a22, double inv_norm, float* pdf, void* stream) {
    GN_REQUIRE(data_xy && pos_xy && pdf, "null pointer");
    GN_REQUIRE(n > 0 && m >= 0, "n must be > 0 and m >= 0");
    GN_REQUIRE(a11 > 0.0 && a22 > 0.0 && a11 * a22 - a12 * a12 > 0.0, "inverse covariance is not positive definite");
    if (m == 0) return GN_OK;
    const double l2e = 1.4426950408889634074;
    kde2d_pdf_kernel<<<(m + 127) / 128, 128, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(data_xy), n, reinterpret_cast<const float2*>(pos_xy), m, (float)(0.5 * a11 * l2e),
        (float)(a12 * l2e), (float)(0.5 * a22 * l2e), inv_norm, pdf);
    return cuda_status("kde2d_pdf_kernel");
}

extern "C" int gn_overlap_sums_f32(const float* a, const float* b, long long n, double* out3, void* stream) {
    GN_REQUIRE(a && b && out3, "null pointer");
    GN_REQUIRE(n > 0, "n must be > 0");
    overlap_sums_kernel<<<1, 1024, 0, as_stream(stream)>>>(a, b, n, out3);
    return cuda_status("overlap_sums_kernel");
}
