// fp32 SIMT implicit-GEMM core used by the exact-parity (float32) path of Conv1D and Dense.
//
//   C[m, n] (+)= sum_k A(m, k) * B(k, n)
//
// A and B are *functors* (implicit im2col for the convolutions), C goes through an epilogue functor.
// Tile 128x128x16, 256 threads, 8x8 register micro-tile per thread (two 4-wide strips per dimension so
// shared-memory reads are conflict-free float4), register-staged prefetch of the next k-slab.
// Split-K over blockIdx.z for the reductions over B*L of the weight gradients.
//
// This is the correctness anchor; the throughput path is the tcgen05 kernel in conv1d_tc.cu.
#pragma once
#include "gn_common.cuh"

namespace gn {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

// A_KFAST: consecutive threads fetch consecutive k of A (true when A(m, .) is contiguous in memory)
// B_KFAST: consecutive threads fetch consecutive k of B (true when B(., n) is contiguous in memory)
template <bool A_KFAST, bool B_KFAST, class FA, class FB, class EPI>
__global__ void __launch_bounds__(SG_THREADS, 2)
gemm_simt_kernel(FA fa, FB fb, EPI epi, int M, int N, int K, int k_per_split) {
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = tid + i * SG_THREADS;  // 0..2047
            int mm, kk;
            if (A_KFAST) { kk = e & 15; mm = e >> 4; } else { mm = e & 127; kk = e >> 7; }
            int gm = m0 + mm, gk = k0 + kk;
            ra[i] = (gm < M && gk < kend) ? fa(gm, gk) : 0.f;
            int nn, kb;
            if (B_KFAST) { kb = e & 15; nn = e >> 4; } else { nn = e & 127; kb = e >> 7; }
            int gn_ = n0 + nn, gkb = k0 + kb;
            rb[i] = (gn_ < N && gkb < kend) ? fb(gkb, gn_) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = tid + i * SG_THREADS;
            int mm, kk;
            if (A_KFAST) { kk = e & 15; mm = e >> 4; } else { mm = e & 127; kk = e >> 7; }
            As[kk][mm] = ra[i];
            int nn, kb;
            if (B_KFAST) { kb = e & 15; nn = e >> 4; } else { nn = e & 127; kb = e >> 7; }
            Bs[kb][nn] = rb[i];
        }
    };

    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
        stash();
        __syncthreads();
        if (k0 + SG_BK < kend) fetch(k0 + SG_BK);
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gn_ = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn_ < N) epi(gm, gn_, acc[i][j]);
        }
    }
}

template <bool A_KFAST, bool B_KFAST, class FA, class FB, class EPI>
inline int launch_gemm_simt(FA fa, FB fb, EPI epi, int M, int N, int K, int splits, cudaStream_t st) {
    if (M <= 0 || N <= 0) return GN_OK;
    if (splits < 1) splits = 1;
    int k_per = ((K + splits - 1) / splits + SG_BK - 1) / SG_BK * SG_BK;
    if (k_per < SG_BK) k_per = SG_BK;
    splits = (K + k_per - 1) / k_per;
    if (splits < 1) splits = 1;
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM, splits);
    if (grid.y > 65535u || grid.z > 65535u) return fail(GN_ERR_UNSUPPORTED, "gemm_simt: grid too large%s", "");
    gemm_simt_kernel<A_KFAST, B_KFAST><<<grid, SG_THREADS, 0, st>>>(fa, fb, epi, M, N, K, k_per);
    return cuda_status("gemm_simt_kernel");
}

// how many K-splits give every SM about two CTAs
inline int pick_splits(int M, int N, int K) {
    long long tiles = (long long)((M + SG_BM - 1) / SG_BM) * ((N + SG_BN - 1) / SG_BN);
    long long want = 2LL * num_sms();
    if (tiles >= want) return 1;
    long long s = (want + tiles - 1) / tiles;
    long long maxs = (K + 4 * SG_BK - 1) / (4 * SG_BK);
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    if (s > 4096) s = 4096;
    return (int)s;
}

// ---- small-N product: one warp per output row, N <= 4 (Dense heads, Cout=1 convolutions) -------------------
template <int NS, class FA, class FB, class EPI>
__global__ void __launch_bounds__(256) gemv_rows_kernel(FA fa, FB fb, EPI epi, int M, int K) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= M) return;
    float acc[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
        float a = fa(warp, k);
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[j] = fmaf(a, fb(k, j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        float s = warp_sum(acc[j]);
        if (lane == 0) epi(warp, j, s);
    }
}

template <class FA, class FB, class EPI>
inline int launch_gemv_rows(FA fa, FB fb, EPI epi, int M, int K, int N, cudaStream_t st) {
    if (M <= 0) return GN_OK;
    long long threads = (long long)M * 32;
    unsigned grid = (unsigned)((threads + 255) / 256);
    switch (N) {
        case 1: gemv_rows_kernel<1><<<grid, 256, 0, st>>>(fa, fb, epi, M, K); break;
        case 2: gemv_rows_kernel<2><<<grid, 256, 0, st>>>(fa, fb, epi, M, K); break;
        case 3: gemv_rows_kernel<3><<<grid, 256, 0, st>>>(fa, fb, epi, M, K); break;
        case 4: gemv_rows_kernel<4><<<grid, 256, 0, st>>>(fa, fb, epi, M, K); break;
        default: return fail(GN_ERR_UNSUPPORTED, "gemv_rows: N > 4%s", "");
    }
    return cuda_status("gemv_rows_kernel");
}

// ---- small-M reduction: out[k, j] = sum_m A(m, k) * D(m, j), j < NS; threads over k, loop over m chunks ----
// (weight gradient of the small-N layers: K is huge and contiguous, M = rows to reduce)
template <int NS, class FA, class FD>
__global__ void __launch_bounds__(256) colreduce_small_kernel(FA fa, FD fd, float* __restrict__ out, int M, int K,
                                                              int m_per_split) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int mb = blockIdx.y * m_per_split, me = min(M, mb + m_per_split);
    float acc[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = 0.f;
    for (int m = mb; m < me; ++m) {
        float a = fa(m, k);
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[j] = fmaf(a, fd(m, j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) atomicAdd(&out[(size_t)k * NS + j], acc[j]);
}

template <class FA, class FD>
inline int launch_colreduce_small(FA fa, FD fd, float* out, int M, int K, int N, cudaStream_t st) {
    if (K <= 0) return GN_OK;
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)K * N, st);
    if (M <= 0) return GN_OK;
    int kblocks = (K + 255) / 256;
    int splits = (int)((4LL * num_sms() + kblocks - 1) / kblocks);
    if (splits < 1) splits = 1;
    if (splits > M) splits = M;
    if (splits > 65535) splits = 65535;
    int m_per = (M + splits - 1) / splits;
    splits = (M + m_per - 1) / m_per;
    dim3 grid(kblocks, splits);
    switch (N) {
        case 1: colreduce_small_kernel<1><<<grid, 256, 0, st>>>(fa, fd, out, M, K, m_per); break;
        case 2: colreduce_small_kernel<2><<<grid, 256, 0, st>>>(fa, fd, out, M, K, m_per); break;
        case 3: colreduce_small_kernel<3><<<grid, 256, 0, st>>>(fa, fd, out, M, K, m_per); break;
        case 4: colreduce_small_kernel<4><<<grid, 256, 0, st>>>(fa, fd, out, M, K, m_per); break;
        default: return fail(GN_ERR_UNSUPPORTED, "colreduce_small: N > 4%s", "");
    }
    return cuda_status("colreduce_small_kernel");
}

}  // namespace gn
