// Sample synthesis kernels: PSD-coloured noise, Tukey-windowed rFFT whitening, injection, crop.
//
// One CTA owns one time series end to end: the N-point real transforms are done as M=N/2-point
// complex Stockham FFTs held in shared memory (radix-16 register butterflies, M/16 threads,
// 16 points per thread), so a series is read from HBM once and written once.
//   gn_whiten_td_f32 : window -> rfft -> x weights -> irfft -> crop -> scale
//   gn_irfft_f32     : x weights -> irfft -> roll -> scale
//   gn_synth_f32     : normals x amp -> irfft (coloured noise) -> + template -> whiten_td -> crop -> scale
// Reference arithmetic: BBH_version/gw_template_maker.py:161-193 (gen_noise), :243-286 (whiten_data),
// :695 (crop), :813-814 (norm constant).  Algorithmic HBM bytes per series: 8N (whiten), 8N+4L (synth,
// normals fed in), 4N+4L (synth, Philox normals).
#include "gn_common.cuh"
#include "philox.cuh"

#include <cstdlib>
#include <math.h>
#include <mutex>
#include <vector>

struct gn_fft_plan {
    int N;
    int log2M;
    float2* tw;   // device, exp(-2*pi*i*j/N), j in [0,N)   (real-FFT packing twiddles)
    float2* ptw;  // device, per-pass Stockham twiddles laid out [pass][r-1][k] = exp(-2*pi*i*r*k/(p*R)), k < p,
                  // so that the lanes of a warp (consecutive k) read consecutive 8-byte entries
    float2* tw64; // device, N == 8192 only: [k1][t] = exp(-2*pi*i*k1*t/4096), k1, t < 64 (64 x 64 decomposition)
    // Whitening coefficient tables (alpha_k, beta_k), M entries each, rebuilt by a prologue kernel on every call from
    // the caller's weights, window and output scale.  A slot is keyed by (weights pointer, window pointer, scale): calls
    // with the same key write the same values (the caller may not change weights or window while a call that reads them
    // is in flight), so they can share the slot on any number of streams; a slot handed to another key first waits for
    // the event that covers its earlier users.
    static constexpr int NSLOT = 8;
    float2* coef[NSLOT];
    const void* coef_key[NSLOT];
    const void* coef_win[NSLOT];
    float coef_scale[NSLOT];
    cudaStream_t coef_stream[NSLOT];     // a slot is only ever re-used AS IS by the stream that filled it
    cudaEvent_t coef_done[NSLOT];
    bool coef_used[NSLOT];
    int coef_next;
    std::mutex mu;
};

namespace gn {

// packed two-lane fp32 arithmetic (Blackwell FADD2 / FFMA2): a complex add or subtract is one instruction
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
// multiply by exp(DIR*i*pi/2): DIR<0 -> -i, DIR>0 -> +i
template <int DIR>
__device__ __forceinline__ float2 mul_wi(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

template <int DIR>
__device__ __forceinline__ void fft2(float2& a0, float2& a1) {
    float2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}
template <int DIR>
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_wi<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

// In-register DFT of R points. After run(v), output bin r lives in v[out_reg(r)].
template <int R, int DIR>
struct Dft;

template <int DIR>
struct Dft<2, DIR> {
    static __device__ __forceinline__ void run(float2* v) { fft2<DIR>(v[0], v[1]); }
    static __device__ __forceinline__ constexpr int out_reg(int r) { return r; }
};
template <int DIR>
struct Dft<4, DIR> {
    static __device__ __forceinline__ void run(float2* v) { fft4<DIR>(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ constexpr int out_reg(int r) { return r; }
};
template <int DIR>
struct Dft<8, DIR> {
    static __device__ __forceinline__ void run(float2* v) {
        const float h = 0.70710678118654752440f;
        fft4<DIR>(v[0], v[2], v[4], v[6]);
        fft4<DIR>(v[1], v[3], v[5], v[7]);
        // v[2k1+1] *= W8^{k1}
        v[3] = cmul(v[3], make_float2(h, DIR * h));
        v[5] = mul_wi<DIR>(v[5]);
        v[7] = cmul(v[7], make_float2(-h, DIR * h));
        fft2<DIR>(v[0], v[1]);
        fft2<DIR>(v[2], v[3]);
        fft2<DIR>(v[4], v[5]);
        fft2<DIR>(v[6], v[7]);
    }
    static __device__ __forceinline__ constexpr int out_reg(int r) { return r < 4 ? 2 * r : 2 * (r - 4) + 1; }
};
template <int DIR>
struct Dft<16, DIR> {
    static __device__ __forceinline__ void run(float2* v) {
        const float h = 0.70710678118654752440f;
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) fft4<DIR>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
        // v[4*k1+n2] *= W16^{n2*k1},  W16 = exp(DIR*2*pi*i/16)
        v[5] = cmul(v[5], make_float2(c1, DIR * s1));    // 1
        v[6] = cmul(v[6], make_float2(h, DIR * h));      // 2
        v[7] = cmul(v[7], make_float2(s1, DIR * c1));    // 3
        v[9] = cmul(v[9], make_float2(h, DIR * h));      // 2
        v[10] = mul_wi<DIR>(v[10]);                      // 4
        v[11] = cmul(v[11], make_float2(-h, DIR * h));   // 6
        v[13] = cmul(v[13], make_float2(s1, DIR * c1));  // 3
        v[14] = cmul(v[14], make_float2(-h, DIR * h));   // 6
        v[15] = cmul(v[15], make_float2(-c1, -DIR * s1));  // 9
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) fft4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    }
    // X[k1 + 4*k2] sits in v[4*k1 + k2]
    static __device__ __forceinline__ constexpr int out_reg(int r) { return 4 * (r & 3) + (r >> 2); }
};

// Table load (twiddles, whitening coefficients) that stays where it is written.  The tables are read-only, so a
// __ldg / ld.global.nc of them is free to move: loop-invariant code motion lifts the per-thread loads (identical for
// every series) out of the persistent loop, and ptxas hoists them over the preceding __syncthreads into the previous
// pass -- in both cases the values no longer fit in the 64-register budget and are spilled to local memory, which
// costs more L1 wavefronts than the L1-resident loads themselves.  A coherent ld.global is ordered by the barrier.
// Streaming load of a series element that is read exactly once: no L1 allocation (the tables keep the L1).
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_table(const float2* p) {
#ifdef GN_TABLE_LDG
    // experiment: movable read-only loads; with a register budget that holds the tables (GN_SYNTH_REGS >= 128) and
    // persistent CTAs (GN_WAVES = 1) the compiler keeps the per-thread twiddles / coefficients in registers across series
    return __ldg(p);
#else
    float2 v;
    asm volatile("ld.global.ca.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
#endif
}

// shared-memory index with one pad slot per 16 complex values (kills the stride-R store conflicts)
__device__ __forceinline__ int PADI(int i) { return i + (i >> 4); }

// Shared-memory views with compile-time strides: element (i + r*T) of the padded array is PADI(i) + r*(T + T/16)
// whenever T is a multiple of 16, so every access of a thread is base + immediate.
template <int STRIDE>
__device__ __forceinline__ int padi_off(int base_idx, int r) {
    if (STRIDE % 16 == 0) return PADI(base_idx) + r * (STRIDE + STRIDE / 16);
    return PADI(base_idx + r * STRIDE);
}

// Pass adapters.  get<T>(i, r, j) fetches logical element i + r*T of the pass input, put<P>(base, r, v, j) writes logical
// element base + r*P of the pass output.  j is the compile-time slot of that element in the thread's own 16 points
// whenever the element is tid + j*(M/16) (first-pass inputs, last-pass outputs): register adapters index with it.
struct SmemIn {
    const float2* buf;
    template <int T>
    __device__ __forceinline__ float2 get(int i, int r, int) const { return buf[padi_off<T>(i, r)]; }
    template <int T>
    __device__ __forceinline__ float2 fix(float2 v, int, int, int) const { return v; }
};
struct SmemOut {
    float2* buf;
    template <int P>
    __device__ __forceinline__ void put(int base, int r, float2 v, int) const { buf[padi_off<P>(base, r)] = v; }
};
// the thread's 16 points z[j] = element tid + j*(M/16), kept in registers between two transforms
struct RegIn {
    const float2* z;
    template <int T>
    __device__ __forceinline__ float2 get(int, int, int j) const { return z[j]; }
    template <int T>
    __device__ __forceinline__ float2 fix(float2 v, int, int, int) const { return v; }
};
struct RegOut {
    float2* z;
    template <int P>
    __device__ __forceinline__ void put(int, int, float2 v, int j) const { z[j] = v; }
};

// adapters for pass inputs/outputs expressed on the logical index
template <class F>
struct IdxIn {
    F f;
    template <int T>
    __device__ __forceinline__ float2 get(int i, int r, int j) const { return f(i + r * T, j); }
    template <int T>
    __device__ __forceinline__ float2 fix(float2 v, int, int, int) const { return v; }
};
// two-phase input: f issues the global load of every point of the pass first, g finishes the values (window, noise)
// in a second sweep, so that all loads of a thread are in flight together
template <class F, class G>
struct IdxIn2 {
    F f;
    G g;
    template <int T>
    __device__ __forceinline__ float2 get(int i, int r, int j) const { return f(i + r * T, j); }
    template <int T>
    __device__ __forceinline__ float2 fix(float2 v, int i, int r, int j) const { return g(v, i + r * T, j); }
};
template <class F, class G>
__device__ __forceinline__ IdxIn2<F, G> make_in2(F f, G g) { return IdxIn2<F, G>{f, g}; }
template <class F>
struct IdxOut {
    F f;
    template <int P>
    __device__ __forceinline__ void put(int base, int r, float2 v, int j) const { f(base + r * P, v, j); }
};
template <class F>
__device__ __forceinline__ IdxIn<F> make_in(F f) { return IdxIn<F>{f}; }
template <class F>
__device__ __forceinline__ IdxOut<F> make_out(F f) { return IdxOut<F>{f}; }

// One Stockham pass of radix R with sub-transform length P over the M-point array; each thread owns 16/R
// butterflies (16 points).  A __syncthreads separates loads from stores so that the pass may run in place.
// ptw = this pass's twiddle table [r-1][k].  TWP = 1 (radix 16 only): six table rows (r = 1,2,3,4,8,12) are loaded and
// the other nine twiddles are formed as w^(4a) * w^b -- 2.5x fewer L1 wavefronts for 36 more FP32 operations.
template <int R, int P, int DIR, int LOG2M, int TWP, class IN, class OUT>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ ptw, const IN& in, const OUT& out,
                                         bool sync_before_store) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = M / R;
    constexpr int NT = M / 16;
    constexpr int IT = 16 / R;
    float2 v[IT][R];
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int i = tid + it * NT;
#pragma unroll
        for (int r = 0; r < R; ++r) v[it][r] = in.template get<T>(i, r, it + r * IT);
#pragma unroll
        for (int r = 0; r < R; ++r) v[it][r] = in.template fix<T>(v[it][r], i, r, it + r * IT);
        if (P > 1) {
            const int k = i & (P - 1);
            if constexpr (TWP == 1 && R == 16) {
                float2 wl[4], wh[4];
#pragma unroll
                for (int b = 1; b < 4; ++b) {
                    wl[b] = ld_table(&ptw[(b - 1) * P + k]);
                    wh[b] = ld_table(&ptw[(4 * b - 1) * P + k]);
                }
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const int a = r >> 2, b = r & 3;
                    const float2 w = a == 0 ? wl[b] : (b == 0 ? wh[a] : cmul(wh[a], wl[b]));
                    v[it][r] = DIR > 0 ? cmulc(v[it][r], w) : cmul(v[it][r], w);
                }
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const float2 w = ld_table(&ptw[(r - 1) * P + k]);
                    v[it][r] = DIR > 0 ? cmulc(v[it][r], w) : cmul(v[it][r], w);
                }
            }
        }
        Dft<R, DIR>::run(v[it]);
    }
    if (sync_before_store) __syncthreads();
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int i = tid + it * NT;
        const int k = i & (P - 1);
        const int base = (i - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out.template put<P>(base, r, v[it][Dft<R, DIR>::out_reg(r)], r + it * R);
    }
}

// 64-point DFT in registers as 4 x 16 (Cooley-Tukey): n = 16*n1 + n2, k = k1 + 4*k2.  v[16*k1 + n2] holds the radix-4
// outputs after step 1, is rotated by W_64^(n2*k1) and transformed by Dft<16> per k1; bin k ends in v[out_reg(k)].
__device__ constexpr float kCos64[64] = {1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f, 9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f, 7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f, 3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f, 6.123233996e-17f, -9.801714033e-02f, -1.950903220e-01f, -2.902846773e-01f, -3.826834324e-01f, -4.713967368e-01f, -5.555702330e-01f, -6.343932842e-01f, -7.071067812e-01f, -7.730104534e-01f, -8.314696123e-01f, -8.819212643e-01f, -9.238795325e-01f, -9.569403357e-01f, -9.807852804e-01f, -9.951847267e-01f, -1.000000000e+00f, -9.951847267e-01f, -9.807852804e-01f, -9.569403357e-01f, -9.238795325e-01f, -8.819212643e-01f, -8.314696123e-01f, -7.730104534e-01f, -7.071067812e-01f, -6.343932842e-01f, -5.555702330e-01f, -4.713967368e-01f, -3.826834324e-01f, -2.902846773e-01f, -1.950903220e-01f, -9.801714033e-02f, -1.836970199e-16f, 9.801714033e-02f, 1.950903220e-01f, 2.902846773e-01f, 3.826834324e-01f, 4.713967368e-01f, 5.555702330e-01f, 6.343932842e-01f, 7.071067812e-01f, 7.730104534e-01f, 8.314696123e-01f, 8.819212643e-01f, 9.238795325e-01f, 9.569403357e-01f, 9.807852804e-01f, 9.951847267e-01f};
__device__ constexpr float kSin64[64] = {0.000000000e+00f, 9.801714033e-02f, 1.950903220e-01f, 2.902846773e-01f, 3.826834324e-01f, 4.713967368e-01f, 5.555702330e-01f, 6.343932842e-01f, 7.071067812e-01f, 7.730104534e-01f, 8.314696123e-01f, 8.819212643e-01f, 9.238795325e-01f, 9.569403357e-01f, 9.807852804e-01f, 9.951847267e-01f, 1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f, 9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f, 7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f, 3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f, 1.224646799e-16f, -9.801714033e-02f, -1.950903220e-01f, -2.902846773e-01f, -3.826834324e-01f, -4.713967368e-01f, -5.555702330e-01f, -6.343932842e-01f, -7.071067812e-01f, -7.730104534e-01f, -8.314696123e-01f, -8.819212643e-01f, -9.238795325e-01f, -9.569403357e-01f, -9.807852804e-01f, -9.951847267e-01f, -1.000000000e+00f, -9.951847267e-01f, -9.807852804e-01f, -9.569403357e-01f, -9.238795325e-01f, -8.819212643e-01f, -8.314696123e-01f, -7.730104534e-01f, -7.071067812e-01f, -6.343932842e-01f, -5.555702330e-01f, -4.713967368e-01f, -3.826834324e-01f, -2.902846773e-01f, -1.950903220e-01f, -9.801714033e-02f};
template <int DIR>
struct Dft<64, DIR> {
    static __device__ __forceinline__ void run(float2* v) {
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) fft4<DIR>(v[n2], v[16 + n2], v[32 + n2], v[48 + n2]);
#pragma unroll
        for (int k1 = 1; k1 < 4; ++k1) {
#pragma unroll
            for (int n2 = 1; n2 < 16; ++n2) {
                const int m = (n2 * k1) & 63;            // W_64^m = cos - i sin (forward), conjugate for the inverse
                v[16 * k1 + n2] = cmul(v[16 * k1 + n2], make_float2(kCos64[m], DIR < 0 ? -kSin64[m] : kSin64[m]));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) Dft<16, DIR>::run(v + 16 * k1);
    }
    static __device__ __forceinline__ constexpr int out_reg(int k) {
        return 16 * (k & 3) + Dft<16, DIR>::out_reg(k >> 2);
    }
};

// remainder radix so that M = REM * 16^a
template <int LOG2M>
struct Plan {
    static constexpr int REM = 1 << (LOG2M & 3);
    static constexpr int N16 = LOG2M >> 2;
    // offset (in entries) of the twiddle table of the j-th radix-16 pass (sub-transform length p = REM*16^j)
    static constexpr int table_offset(int j) {
        int off = 0, p = REM;
        for (int q = 0; q < j; ++q) {
            if (p > 1) off += 15 * p;
            p *= 16;
        }
        return off;
    }
    static constexpr int table_size() { return table_offset(N16); }
};

template <int DIR, int LOG2M, int TWP, int J, class IN, class OUTL>
__device__ __forceinline__ void fft16_passes(float2* buf, const float2* __restrict__ ptw, const IN& first_in,
                                             bool first_is_custom, bool first_from_smem, const OUTL& last_out,
                                             bool last_custom) {
    constexpr int REM = Plan<LOG2M>::REM;
    constexpr int N16 = Plan<LOG2M>::N16;
    if constexpr (J < N16) {
        constexpr int P = REM << (4 * J);
        const float2* tab = ptw + Plan<LOG2M>::table_offset(J);
        constexpr bool last = (J == N16 - 1);
        SmemIn sin{buf};
        SmemOut sout{buf};
        if (J == 0 && first_is_custom) {
            if (last && last_custom) fft_pass<16, P, DIR, LOG2M, TWP>(tab, first_in, last_out, first_from_smem);
            else fft_pass<16, P, DIR, LOG2M, TWP>(tab, first_in, sout, first_from_smem);
        } else {
            if (last && last_custom) fft_pass<16, P, DIR, LOG2M, TWP>(tab, sin, last_out, false);
            else fft_pass<16, P, DIR, LOG2M, TWP>(tab, sin, sout, true);
        }
        if (!(last && last_custom)) __syncthreads();
        fft16_passes<DIR, LOG2M, TWP, J + 1>(buf, ptw, first_in, first_is_custom, first_from_smem, last_out,
                                             last_custom);
    }
}

// Full M-point complex FFT.  first_in reads the logical input of the first pass (first_from_smem: it reads `buf`, so
// the pass synchronises before it stores); the result ends in smem `buf` (padded natural order) unless last_out is
// used for the final pass (last_custom).
template <int DIR, int LOG2M, int TWP = 0, class IN, class OUTL>
__device__ __forceinline__ void fft_full(float2* buf, const float2* __restrict__ ptw, const IN& first_in,
                                         bool first_from_smem, const OUTL& last_out, bool last_custom) {
    constexpr int REM = Plan<LOG2M>::REM;
    if constexpr (REM > 1) {
        // a REM pass is never the last one (N16 >= 1 for every supported size); it has P = 1: no twiddles
        SmemOut sout{buf};
        fft_pass<REM, 1, DIR, LOG2M, TWP>(ptw, first_in, sout, first_from_smem);
        __syncthreads();
        fft16_passes<DIR, LOG2M, TWP, 0>(buf, ptw, first_in, false, false, last_out, last_custom);
    } else {
        fft16_passes<DIR, LOG2M, TWP, 0>(buf, ptw, first_in, true, first_from_smem, last_out, last_custom);
    }
}

// rfft-of-packed post-processing, spectral weighting and irfft pre-processing for the pair (k, M-k):
// in: Z = FFT_M(x[2n] + i x[2n+1]); out: Z' with IFFT_M(Z') = y[2n] + i y[2n+1], y = irfft(rfft(x)*w)*M
template <int LOG2M>
__device__ __forceinline__ void whiten_pointwise(float2* buf, const float2* __restrict__ tw,
                                                 const float* __restrict__ wts) {
    constexpr int M = 1 << LOG2M;
    constexpr int NT = M / 16;
    for (int k = threadIdx.x; k <= M / 2; k += NT) {
        if (k == 0) {
            float2 z = buf[0];
            float y0 = __ldg(&wts[0]) * (z.x + z.y);
            float yM = __ldg(&wts[M]) * (z.x - z.y);
            buf[0] = make_float2(0.5f * (y0 + yM), 0.5f * (y0 - yM));
            continue;
        }
        const int mk = M - k;
        float2 zk = buf[PADI(k)], zm = buf[PADI(mk)];
        // with A = (Zk + conj(Zm))/2, O = (Zk - conj(Zm))/(2i), t = exp(-2*pi*i*k/N):  Y[k] = wk (A + tO),
        // conj(Y[M-k]) = wm (A - tO);  E = (Y[k] + conj(Y[M-k]))/2 = s A + d tO,  Op = conj(t) (Y[k] - conj(Y[M-k]))/2
        // = s O + d conj(t) A  with s = (wk+wm)/2, d = (wk-wm)/2  (|t| = 1).  The halves fold into s, d.
        const float2 A2 = make_float2(zk.x + zm.x, zk.y - zm.y);
        const float2 O2 = make_float2(zk.y + zm.y, zm.x - zk.x);
        const float2 t = __ldg(&tw[k]);  // exp(-2*pi*i*k/N)
        const float wk = __ldg(&wts[k]), wm = __ldg(&wts[mk]);
        const float s = 0.25f * (wk + wm), d = 0.25f * (wk - wm);
        const float2 tO = cmul(t, O2), ctA = cmulc(A2, t);
        const float2 E = make_float2(fmaf(s, A2.x, d * tO.x), fmaf(s, A2.y, d * tO.y));
        const float2 Op = make_float2(fmaf(s, O2.x, d * ctA.x), fmaf(s, O2.y, d * ctA.y));
        // Z'[k] = E + i*Op ; Z'[M-k] = conj(E) + i*conj(Op)
        buf[PADI(k)] = make_float2(E.x - Op.y, E.y + Op.x);
        buf[PADI(mk)] = make_float2(E.x + Op.y, -E.y + Op.x);
    }
}

// irfft pre-processing from an explicit half spectrum Y[0..M] (weights applied on load).
template <int LOG2M, class SPEC>
__device__ __forceinline__ void irfft_pre(float2* buf, const float2* __restrict__ tw, SPEC Y, bool drop_dc) {
    constexpr int M = 1 << LOG2M;
    constexpr int NT = M / 16;
    for (int k = threadIdx.x; k <= M / 2; k += NT) {
        if (k == 0) {
            float y0 = drop_dc ? 0.f : Y(0).x;
            float yM = Y(M).x;
            buf[0] = make_float2(0.5f * (y0 + yM), 0.5f * (y0 - yM));
            continue;
        }
        const int mk = M - k;
        float2 P = Y(k), Q = cconj(Y(mk));
        float2 t = __ldg(&tw[k]);
        float2 E = cscale(cadd(P, Q), 0.5f);
        float2 Op = cmul(cconj(t), cscale(csub(P, Q), 0.5f));
        buf[PADI(k)] = make_float2(E.x - Op.y, E.y + Op.x);
        buf[PADI(mk)] = make_float2(E.x + Op.y, -E.y + Op.x);
    }
}

// resident CTAs per SM the register allocator must allow (<= ~85 registers per thread)
#ifndef GN_SYNTH_REGS
#define GN_SYNTH_REGS 64
#endif
#define GN_SYNTH_MINB(L2) (((1 << (L2)) / 16) >= 512 ? 1 : (65536 / ((((1 << (L2)) / 16) < 32 ? 32 : ((1 << (L2)) / 16)) * GN_SYNTH_REGS) > 8 ? 8 : 65536 / ((((1 << (L2)) / 16) < 32 ? 32 : ((1 << (L2)) / 16)) * GN_SYNTH_REGS)))
constexpr int MODE_WHITEN = 0, MODE_IRFFT = 1, MODE_SYNTH = 2;
#ifndef GN_COEF_PREFETCH
#define GN_COEF_PREFETCH 6
#endif


struct SynthArgs {
    const float* x;          // WHITEN: (batch,N) input; IRFFT: (batch,Nf) complex; SYNTH: normals (batch,2,Nf) or null
    const float* amp;        // SYNTH: (Nf)
    const float* templates;  // SYNTH: (n_templates,N) or null
    const int* tidx;         // SYNTH: (batch) or null
    const float* window;     // (N)
    const float* weights;    // (Nf) (IRFFT: may be null)
    float* y;
    const float2* tw;
    const float2* ptw;
    const float2* tw64;
    const float2* coef;      // (M + 1) whitening coefficients (alpha_k, beta_k) and flat window range of whiten_coef_kernel
    int batch, n_templates, crop_lo, crop_len, roll, drop_dc;
    float noise_scale, out_scale;
    unsigned long long seed, sample_offset;
};

// ---- gn_whiten_td_f32 for N = 8192 (M = 4096 = 64 x 64): two register-resident radix-64 passes per transform --------
// 64 threads own one series.  Element n = 64*n1 + n2 of the packed complex series, bin k = k1 + 64*k2:
//   pass A (thread = n2): DFT-64 over n1, times W_4096^(n2*k1), to shared memory S[k1][n2]
//   pass B (thread = k1): DFT-64 over n2 -> bins k1 + 64*k2 in the thread's registers
// which is also what pass A of the inverse transform asks of thread k1, so the whitening step happens in registers:
// only the partner bins (M-k) cross through shared memory (unpadded [k2][k1] layout, conflict-free both ways) and the
// step itself is Z'[k] = alpha_k Z[k] + beta_k i conj(Z[M-k]) from the prologue's coefficient table.
// Rows of S are 65 complex apart, so the column writes of pass A and the row reads of pass B are conflict free.
// Six shared-memory sweeps per series against ten for the radix-16 organisation, one twiddled pass per transform; the
// 63 pass twiddles of a thread are formed from 14 table entries (W^(t*b), W^(8*t*a)) because at 6 CTAs per SM the L1
// keeps under 30 KB.  Selected with GN_WHITEN_RADIX=64.
constexpr int W64_LD = 65;
template <int DIR>
__device__ __forceinline__ void twiddle_store64(float2* S, const float2* __restrict__ tw64, float2* v) {
    const int t = threadIdx.x;
    float2 lo[8];
#pragma unroll
    for (int b = 1; b < 8; ++b) lo[b] = ld_table(&tw64[b * 64 + t]);
#pragma unroll
    for (int a8 = 0; a8 < 8; ++a8) {
        float2 hi = make_float2(1.f, 0.f);
        if (a8 > 0) hi = ld_table(&tw64[a8 * 8 * 64 + t]);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k1 = 8 * a8 + b;
            float2 y = v[Dft<64, DIR>::out_reg(k1)];
            if (k1 > 0) {
                const float2 w = a8 == 0 ? lo[b] : (b == 0 ? hi : cmul(hi, lo[b]));
                y = DIR > 0 ? cmulc(y, w) : cmul(y, w);
            }
            S[k1 * W64_LD + t] = y;
        }
    }
}

__global__ void __launch_bounds__(64, 6) whiten64_kernel(SynthArgs a) {
    constexpr int M = 4096, N = 8192;
    __shared__ float2 S[64 * W64_LD];
    const float2* __restrict__ tw64 = a.tw64;
    const float2* __restrict__ win2 = reinterpret_cast<const float2*>(a.window);
    const float2* __restrict__ ab = a.coef;
    const int t = threadIdx.x;
    // the window is exactly 1 on element 64*n1 + t when (unsigned)(wf0 + 64*n1) < wflen (flat range from the prologue)
    const int2 fr = *reinterpret_cast<const int2*>(a.coef + M);
    const int wf0 = t - fr.x;
    const unsigned wflen = (unsigned)(fr.y - fr.x);
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        const float2* __restrict__ src2 = reinterpret_cast<const float2*>(a.x + (size_t)b * N);
        {   // pull the series this CTA handles next into L2
            const int bn = b + gridDim.x;
            if (bn < a.batch) {
                const char* nx = reinterpret_cast<const char*>(a.x + (size_t)bn * N);
                for (int ln = threadIdx.x; ln < N * 4 / 128; ln += blockDim.x)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)ln * 128));
            }
        }
        float2 v[64];
        // forward pass A: window, packed real -> complex
#pragma unroll
        for (int n1 = 0; n1 < 64; ++n1) v[n1] = ld_stream(&src2[64 * n1 + t]);
#pragma unroll
        for (int n1 = 0; n1 < 64; ++n1) {
            if (!((unsigned)(wf0 + 64 * n1) < wflen)) {
                const float2 w = ld_table(&win2[64 * n1 + t]);
                v[n1] = make_float2(v[n1].x * w.x, v[n1].y * w.y);
            }
        }
        Dft<64, -1>::run(v);
        twiddle_store64<-1>(S, tw64, v);
        __syncthreads();
        // forward pass B
#pragma unroll
        for (int n2 = 0; n2 < 64; ++n2) v[n2] = S[t * W64_LD + n2];
        __syncthreads();          // every row has been read: the exchange below reuses the buffer
        Dft<64, -1>::run(v);
        // whitening step: bin k = t + 64*k2 sits in v[out_reg(k2)]; partners through the unpadded exchange layout
#pragma unroll
        for (int k2 = 0; k2 < 64; ++k2) S[64 * k2 + t] = v[Dft<64, -1>::out_reg(k2)];
        __syncthreads();
        float2 u[64];
#pragma unroll
        for (int k2 = 0; k2 < 64; ++k2) {
            const int k = t + 64 * k2;
            const float2 zm = S[(M - k) & (M - 1)];
            const float2 c = ld_table(&ab[k]);
            const float2 z = v[Dft<64, -1>::out_reg(k2)];
            u[k2] = make_float2(fmaf(c.x, z.x, c.y * zm.y), fmaf(c.x, z.y, c.y * zm.x));
        }
        __syncthreads();          // partner reads complete before the inverse pass A overwrites S
        // inverse pass A: thread t owns elements 64*n1 + t = the bins it just whitened
        Dft<64, +1>::run(u);
        twiddle_store64<+1>(S, tw64, u);
        __syncthreads();
#pragma unroll
        for (int n2 = 0; n2 < 64; ++n2) v[n2] = S[t * W64_LD + n2];
        __syncthreads();          // rows read: the next series' pass A may store
        Dft<64, +1>::run(v);
        float* __restrict__ yb = a.y + (size_t)b * a.crop_len;
        // the output scale is already in the coefficients; the host sends odd crop windows to the radix-16 kernel
        const int nb = 2 * t - a.crop_lo;
        float* __restrict__ yt = yb + nb;
        const unsigned clen = (unsigned)a.crop_len;
#pragma unroll
        for (int k2 = 0; k2 < 64; ++k2)
            if ((unsigned)(nb + 128 * k2) < clen) *reinterpret_cast<float2*>(yt + 128 * k2) = v[Dft<64, +1>::out_reg(k2)];
    }
}

// The rfft split, the whitening weights and the irfft packing collapse to two real coefficients per bin: with
// Z = FFT_M(x[2n] + i x[2n+1]) and t_k = exp(-2 pi i k / N) = c_k + i s_k,
//     Z'[k] = alpha_k Z[k] + beta_k * i conj(Z[(M-k) mod M]),   alpha_k = (w_k + w_{M-k})/2 + (w_k - w_{M-k})/2 * s_k,
//                                                                beta_k  = (w_k - w_{M-k})/2 * c_k,
// and IFFT_M(Z') = y[2n] + i y[2n+1] with y = irfft(rfft(x) * w) (k = 0 included: its partner is itself and w_M the
// Nyquist weight).  Same arithmetic as gw_template_maker.py:277-283 (rfft, multiply, irfft).
// `scale` (the caller's output scale and the 1/M of the unnormalised inverse) is folded into both coefficients.
// The same launch finds the flat part of the window: ab[M] holds (as two ints) the range [lo, hi) of packed elements
// n (samples 2n, 2n+1) on which the window is exactly 1 -- the main kernels skip the window load and multiply there.
// A window whose unit elements are not one contiguous run gets the empty range (every element is then multiplied).
__global__ void __launch_bounds__(1024) whiten_coef_kernel(const float* __restrict__ wts, const float* __restrict__ window,
                                                           float2* __restrict__ ab, int M, float scale) {
    const double sc = (double)scale;
    int lo = M, hi = 0, cnt = 0;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const double wk = (double)wts[k], wm = (double)wts[M - k];
        double sn, cs;
        sincospi((double)k / (double)M, &sn, &cs);      // angle 2 pi k / N
        ab[k] = make_float2((float)(sc * (0.5 * (wk + wm) - 0.5 * (wk - wm) * sn)), (float)(sc * 0.5 * (wk - wm) * cs));
        if (window[2 * k] == 1.f && window[2 * k + 1] == 1.f) {
            lo = min(lo, k);
            hi = max(hi, k + 1);
            ++cnt;
        }
    }
    __shared__ int s_lo, s_hi, s_cnt;
    if (threadIdx.x == 0) { s_lo = M; s_hi = 0; s_cnt = 0; }
    __syncthreads();
    atomicMin(&s_lo, lo);
    atomicMax(&s_hi, hi);
    atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        int2 r = make_int2(0, 0);
        if (s_cnt > 0 && s_hi - s_lo == s_cnt) r = make_int2(s_lo, s_hi);
        *reinterpret_cast<int2*>(ab + M) = r;
    }
}

// VAR 0: transforms meet in shared memory (forward result -> whiten_pointwise -> inverse).
// VAR 1: a thread keeps its 16 points (bins tid + j*M/16: what the last forward pass leaves it and what the first inverse
//        pass asks of it) in registers across the whitening step; only the partner bins M-k cross through shared memory
//        (one conflict-free write + read instead of four padded sweeps), the pointwise step is two FMAs per component
//        from the (alpha, beta) table, window loads are skipped where the window is exactly 1, and in SYNTH mode the
//        coloured noise goes from its inverse transform to the forward transform in registers as well; six-row product
//        twiddles (fft_pass TWP = 1).
template <int LOG2M, int MODE, int VAR>
__global__ void __launch_bounds__((1 << LOG2M) / 16, GN_SYNTH_MINB(LOG2M)) synth_kernel(SynthArgs a) {
    constexpr int M = 1 << LOG2M;
    constexpr int N = 2 * M;
    constexpr int Nf = M + 1;
    constexpr int NT = M / 16;
    constexpr int TWP = VAR >= 1 ? 1 : 0;
    extern __shared__ float2 buf[];  // PADI(M) complex
    const float2* __restrict__ tw = a.tw;
    const float2* __restrict__ ptw = a.ptw;
    constexpr bool fast_out = VAR >= 1 && MODE != MODE_IRFFT;      // the host sends odd crop windows to VAR 0
    // VAR >= 1: the thread's j-th point is packed element tid + j*NT.  The window is exactly 1 on it when
    // (unsigned)(wf0 + j*NT) < wflen (flat range from the prologue), and its output pair lies inside the (even) crop
    // window when (unsigned)(nb + 2*j*NT) < crop_len: one add and one compare each, no per-thread mask to keep.
    int wf0 = 0;
    unsigned wflen = 0;
    if (VAR >= 1 && MODE != MODE_IRFFT) {
        const int2 fr = *reinterpret_cast<const int2*>(a.coef + M);
        wf0 = (int)threadIdx.x - fr.x;
        wflen = (unsigned)(fr.y - fr.x);
    }
    const int nb = 2 * (int)threadIdx.x - a.crop_lo;
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        // output of the final inverse pass: packed (y[2j], y[2j+1]) at logical index j
        float* __restrict__ yb = a.y + (size_t)b * (MODE == MODE_IRFFT ? N : a.crop_len);
        const float oscale = (VAR >= 1 && MODE != MODE_IRFFT) ? 1.f : a.out_scale;      // VAR >= 1: in the coefficients
        float* __restrict__ ycrop = yb + nb;          // dereferenced inside the crop window only
        auto out_store = [&](int j, float2 val, int slot) {
            if (fast_out) {
                if ((unsigned)(nb + 2 * slot * NT) < (unsigned)a.crop_len)
                    *reinterpret_cast<float2*>(ycrop + 2 * slot * NT) = val;
                return;
            }
            if (MODE == MODE_IRFFT) {
                int n0 = (2 * j + a.roll) & (N - 1);
                if ((a.roll & 1) == 0) {
                    *reinterpret_cast<float2*>(yb + n0) = make_float2(val.x * oscale, val.y * oscale);
                } else {
                    yb[n0] = val.x * oscale;
                    yb[(n0 + 1) & (N - 1)] = val.y * oscale;
                }
            } else {
                int n0 = 2 * j - a.crop_lo;
                if (((a.crop_lo | a.crop_len) & 1) == 0) {
                    if (n0 >= 0 && n0 < a.crop_len) {
                        if (n0 + 1 < a.crop_len)
                            *reinterpret_cast<float2*>(yb + n0) = make_float2(val.x * oscale, val.y * oscale);
                        else
                            yb[n0] = val.x * oscale;
                    }
                } else {
                    if (n0 >= 0 && n0 < a.crop_len) yb[n0] = val.x * oscale;
                    if (n0 + 1 >= 0 && n0 + 1 < a.crop_len) yb[n0 + 1] = val.y * oscale;
                }
            }
        };

        if (MODE == MODE_IRFFT) {
            const float2* __restrict__ xf = reinterpret_cast<const float2*>(a.x) + (size_t)b * Nf;
            const float* __restrict__ w = a.weights;
            auto spec = [&](int k) {
                float2 v = __ldg(&xf[k]);
                float s = w ? __ldg(&w[k]) : 1.f;
                return make_float2(v.x * s, v.y * s);
            };
            irfft_pre<LOG2M>(buf, tw, spec, a.drop_dc != 0);
            __syncthreads();
            fft_full<+1, LOG2M, TWP>(buf, ptw, SmemIn{buf}, true, make_out(out_store), true);
            __syncthreads();
            continue;
        }

        float2 z[16];      // VAR >= 1: the thread's points tid + j*NT between transforms

        if (MODE == MODE_SYNTH) {
            // coloured noise: Y[k] = amp[k]*(re[k] + i*im[k]), DC dropped (gen_noise :187-190)
            const float* __restrict__ amp = a.amp;
            if (a.x != nullptr) {
                const float* __restrict__ nre = a.x + (size_t)b * 2 * Nf;
                const float* __restrict__ nim = nre + Nf;
                auto spec = [&](int k) {
                    float s = __ldg(&amp[k]);
                    return make_float2(__ldg(&nre[k]) * s, __ldg(&nim[k]) * s);
                };
                irfft_pre<LOG2M>(buf, tw, spec, true);
            } else {
                const unsigned long long sample = a.sample_offset + (unsigned long long)b;
                auto spec = [&](int k) {
                    float2 g = philox_normal2(a.seed, sample, (unsigned)k);
                    float s = __ldg(&amp[k]);
                    return make_float2(g.x * s, g.y * s);
                };
                irfft_pre<LOG2M>(buf, tw, spec, true);
            }
            __syncthreads();
            if (VAR >= 1) fft_full<+1, LOG2M, TWP>(buf, ptw, SmemIn{buf}, true, RegOut{z}, true);
            else fft_full<+1, LOG2M, TWP>(buf, ptw, SmemIn{buf}, true, SmemOut{buf}, false);
            // fft_full ends with a __syncthreads when the last pass goes to smem
        }

        // forward transform of window * (noise + template | input)
        {
            const float* __restrict__ src = nullptr;
            if (MODE == MODE_WHITEN) {
                src = a.x + (size_t)b * N;
                // pull the series this CTA handles next into L2 while this one is being transformed
                const int bn = b + gridDim.x;
                if (bn < a.batch) {
                    const char* nx = reinterpret_cast<const char*>(a.x + (size_t)bn * N);
                    for (int ln = threadIdx.x; ln < N * 4 / 128; ln += blockDim.x)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)ln * 128));
                }
            } else if (a.templates != nullptr) {
                int t = a.tidx ? __ldg(&a.tidx[b]) : b;
                t = min(max(t, 0), a.n_templates - 1);
                src = a.templates + (size_t)t * N;
            }
            const float2* __restrict__ win2 = reinterpret_cast<const float2*>(a.window);
            const float2* __restrict__ src2 = reinterpret_cast<const float2*>(src);
            const float nscale = a.noise_scale;
            auto first_load = [&](int idx, int j) {
                float2 v = make_float2(0.f, 0.f);
                if (MODE == MODE_SYNTH) {
                    float2 n = VAR >= 1 ? z[j] : buf[PADI(idx)];
                    v = make_float2(n.x * nscale, n.y * nscale);
                }
                if (src2 != nullptr) {
                    float2 s = __ldg(&src2[idx]);
                    v.x += s.x;
                    v.y += s.y;
                }
                float2 w = __ldg(&win2[idx]);
                return make_float2(v.x * w.x, v.y * w.y);
            };
            // VAR >= 1: the same in two phases -- every global load of the pass first, then noise, window
            // (SYNTH: the noise already fills the registers, so the template loads stay with the second phase)
            auto load_src = [&](int idx, int j) {
                if (MODE == MODE_WHITEN) return ld_stream(&src2[idx]);
                return make_float2(z[j].x * nscale, z[j].y * nscale);
            };
            auto finish = [&](float2 v, int idx, int j) {
                if (MODE == MODE_SYNTH && src2 != nullptr) {
                    const float2 s = __ldg(&src2[idx]);
                    v.x += s.x;
                    v.y += s.y;
                }
                if ((unsigned)(wf0 + j * NT) < wflen) return v;
                const float2 w = ld_table(&win2[idx]);
                return make_float2(v.x * w.x, v.y * w.y);
            };
            // MODE_SYNTH: the first pass must not store into buf before every thread has read the noise transform's
            // last-pass input (VAR >= 1) / its own noise points (VAR 0)
            if (VAR >= 1) fft_full<-1, LOG2M, TWP>(buf, ptw, make_in2(load_src, finish), MODE == MODE_SYNTH, RegOut{z}, true);
            else fft_full<-1, LOG2M, TWP>(buf, ptw, make_in(first_load), MODE == MODE_SYNTH, SmemOut{buf}, false);
        }
        if (VAR >= 1) {
            __syncthreads();      // the last forward pass has been read by every thread
#pragma unroll
            for (int j = 0; j < 16; ++j) buf[j * NT + threadIdx.x] = z[j];
            const float2* __restrict__ ab = a.coef;
            // the first coefficients do not depend on the exchange: their loads fly while the CTA meets at the barrier
            constexpr int NPRE = GN_COEF_PREFETCH;
            float2 cpre[NPRE > 0 ? NPRE : 1];
#pragma unroll
            for (int j = 0; j < NPRE; ++j) cpre[j] = ld_table(&ab[threadIdx.x + j * NT]);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = threadIdx.x + j * NT;
                const float2 zm = buf[(M - k) & (M - 1)];      // a warp reads 32 consecutive slots, descending
                const float2 c = j < NPRE ? cpre[j < NPRE ? j : 0] : ld_table(&ab[k]);
                z[j] = make_float2(fmaf(c.x, z[j].x, c.y * zm.y), fmaf(c.x, z[j].y, c.y * zm.x));
            }
            // the first inverse pass synchronises before it stores: the partner reads above are complete by then
            fft_full<+1, LOG2M, TWP>(buf, ptw, RegIn{z}, true, make_out(out_store), true);
        } else {
            whiten_pointwise<LOG2M>(buf, tw, a.weights);
            __syncthreads();
            fft_full<+1, LOG2M, TWP>(buf, ptw, SmemIn{buf}, true, make_out(out_store), true);
        }
        __syncthreads();
    }
}

static int whiten_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GN_WHITEN_RADIX");
        v = e ? atoi(e) : 16;      // 64 selects the experimental two-pass kernel below (measured slower)
    }
    return v;
}

// Coefficient slot for this call's key (see gn_fft_plan): the prologue kernel may be launched on `st` afterwards.
// The caller holds pl->mu from here until coef_release has returned, so that no other host thread can recycle the slot
// between this call's prologue and the event that covers its main kernel.
static int coef_acquire(gn_fft_plan* pl, const float* weights, const float* window, float scale, cudaStream_t st) {
    for (int i = 0; i < gn_fft_plan::NSLOT; ++i)
        if (pl->coef_key[i] == (const void*)weights && pl->coef_win[i] == (const void*)window &&
            pl->coef_scale[i] == scale && pl->coef_stream[i] == st && pl->coef_used[i])
            return i;      // same stream: the prologue that rewrites the slot is ordered after every earlier use of it
    const int slot = pl->coef_next;
    pl->coef_next = (pl->coef_next + 1) % gn_fft_plan::NSLOT;
    if (pl->coef_used[slot]) cudaStreamWaitEvent(st, pl->coef_done[slot], 0);
    pl->coef_key[slot] = (const void*)weights;
    pl->coef_win[slot] = (const void*)window;
    pl->coef_scale[slot] = scale;
    pl->coef_stream[slot] = st;
    pl->coef_used[slot] = false;
    return slot;
}
// After the main kernel: the slot's event must cover this use and every earlier one.
static void coef_release(gn_fft_plan* pl, int slot, cudaStream_t st) {
    if (pl->coef_used[slot]) cudaStreamWaitEvent(st, pl->coef_done[slot], 0);      // orders later work on st only
    cudaEventRecord(pl->coef_done[slot], st);
    pl->coef_used[slot] = true;
}

// GN_SYNTH_VAR selects the kernel organisation (see synth_kernel); every variant is parity-tested.
static int synth_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GN_SYNTH_VAR");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 1) v = 1;
    }
    return v;
}

template <int MODE, int VAR>
static int launch_synth_var(const gn_fft_plan* plan, SynthArgs a, cudaStream_t st) {
    const int M = plan->N / 2;
    const size_t smem = (size_t)(M + (M >> 4) + 1) * sizeof(float2);
    const int threads = M / 16;
    // one CTA per series; beyond GN_WAVES waves of resident CTAs the CTAs stride over the batch
    int grid = a.batch;
    {
        const int per_sm = threads >= 512 ? 1 : (threads >= 256 ? 65536 / (256 * GN_SYNTH_REGS) : 6);
#ifndef GN_WAVES
#define GN_WAVES 64     // CTAs per resident slot: the per-CTA prologue is a few instructions, so one series per CTA (up to
                         // 64 waves) lets the hardware scheduler balance the tail: 599 -> 575 us on 32768 series against 4 waves
#endif
        const int cap = num_sms() * per_sm * GN_WAVES;
        if (grid > cap) grid = cap;
    }
    int slot = -1;
    gn_fft_plan* pl = const_cast<gn_fft_plan*>(plan);      // the coefficient slots are the plan's own scratch
    std::unique_lock<std::mutex> lock(pl->mu, std::defer_lock);
    if (VAR >= 1 && MODE != MODE_IRFFT) {
        lock.lock();
        slot = coef_acquire(pl, a.weights, a.window, a.out_scale, st);
        whiten_coef_kernel<<<1, 1024, 0, st>>>(a.weights, a.window, pl->coef[slot], M, a.out_scale);
        a.coef = pl->coef[slot];
    }
    auto release = [&]() {
        if (slot >= 0) coef_release(pl, slot, st);
    };
#define GN_SYNTH_CASE(L2)                                                                                   \
    case L2: {                                                                                              \
        auto kfn = synth_kernel<L2, MODE, VAR>;                                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kfn<<<grid, threads, smem, st>>>(a);                                                                \
        break;                                                                                              \
    }
    switch (plan->log2M) {
#ifndef GN_QUICK      // -DGN_QUICK: N = 8192 only, for fast compile-and-inspect iterations (never shipped)
        GN_SYNTH_CASE(8)
        GN_SYNTH_CASE(9)
        GN_SYNTH_CASE(10)
        GN_SYNTH_CASE(11)
        GN_SYNTH_CASE(13)
        GN_SYNTH_CASE(14)
#endif
        GN_SYNTH_CASE(12)
        default:
            release();
            return fail(GN_ERR_UNSUPPORTED, "synth: unsupported FFT length N=%s%lld", "", plan->N);
    }
#undef GN_SYNTH_CASE
    release();
    return cuda_status("synth_kernel");
}

template <int MODE>
static int launch_synth(const gn_fft_plan* plan, SynthArgs a, cudaStream_t st) {
    a.tw = plan->tw;
    a.ptw = plan->ptw;
    a.tw64 = plan->tw64;
    if (MODE == MODE_WHITEN && plan->log2M == 12 && plan->tw64 != nullptr && whiten_variant() == 64 &&
        ((a.crop_lo | a.crop_len) & 1) == 0) {
        int grid = a.batch;
        const int cap = num_sms() * 6 * 4;
        if (grid > cap) grid = cap;
        gn_fft_plan* pl = const_cast<gn_fft_plan*>(plan);
        std::lock_guard<std::mutex> lock(pl->mu);
        const int slot = coef_acquire(pl, a.weights, a.window, a.out_scale, st);
        whiten_coef_kernel<<<1, 1024, 0, st>>>(a.weights, a.window, pl->coef[slot], plan->N / 2, a.out_scale);
        a.coef = pl->coef[slot];
        whiten64_kernel<<<grid, 64, 0, st>>>(a);
        coef_release(pl, slot, st);
        return cuda_status("whiten64_kernel");
    }
    if constexpr (MODE == MODE_IRFFT) {
        return launch_synth_var<MODE, 0>(plan, a, st);      // no whitening step: one organisation
    } else {
        // VAR >= 1 stores whole sample pairs: crop windows with an odd start or length take the VAR 0 organisation
        const int var = ((a.crop_lo | a.crop_len) & 1) ? 0 : synth_variant();
        if (var == 0) return launch_synth_var<MODE, 0>(plan, a, st);
        return launch_synth_var<MODE, 1>(plan, a, st);
    }
}

__global__ void mean_std_kernel(const float* __restrict__ x, long long n, float* out) {
    __shared__ double sm[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
    double tot = block_sum(s, sm);
    __shared__ double mean_s;
    if (threadIdx.x == 0) mean_s = tot / (double)n;
    __syncthreads();
    const double mean = mean_s;
    double q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        double d = (double)x[i] - mean;
        q += d * d;
    }
    double qt = block_sum(q, sm);
    if (threadIdx.x == 0) {
        out[0] = (float)mean;
        out[1] = (float)sqrt(qt / (double)n);
    }
}

__global__ void add_scaled_kernel(float* __restrict__ x, const float* __restrict__ r, float sigma, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] = fmaf(sigma, r[i], x[i]);
}

__global__ void burst_kernel(const double* __restrict__ pars, float* __restrict__ out, int n, int N, float amp,
                             float freq, float dt, float phi) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * N) return;
    int s = (int)(i / N), j = (int)(i - (long long)s * N);
    // double precision: the reference evaluates this in float64 (burstMahoGANy.py:90-93) and the phase
    // 2*pi*f*(t-t0) reaches ~300 rad, where float32 sin would lose 1e-5
    double t0 = pars[2 * s], tau = pars[2 * s + 1];
    double t = (double)dt * j - t0;
    out[i] = (float)((double)amp * sin(2.0 * 3.14159265358979323846 * (double)freq * t + (double)phi) *
                     exp(-(t * t) / (tau * tau)));
}


// gen_bbh tail (gw_template_maker.py:528-571): ref_idx = argmax(hp^2+hc^2) (first maximum), ht = Fp*hp + Fc*hc,
// ts[:len] = ht[ref_idx-idx-lead:] with Python's negative-start semantics, ts *= win, then crop + scale.
__global__ void __launch_bounds__(256) bbh_assemble_kernel(const float* __restrict__ hp, const float* __restrict__ hc,
                                                           const float* __restrict__ Fp, const float* __restrict__ Fc,
                                                           const int* __restrict__ idx, int lead,
                                                           const float* __restrict__ win, float* __restrict__ out,
                                                           int* __restrict__ ref_out, int N, int crop_lo, int crop_len,
                                                           float scale) {
    __shared__ float sv[256];
    __shared__ int si[256];
    const int b = blockIdx.x;
    const float* p = hp + (size_t)b * N;
    const float* c = hc + (size_t)b * N;
    float best = -1.f;
    int bi = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float v = p[j] * p[j] + c[j] * c[j];
        if (v > best) { best = v; bi = j; }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            float v2 = sv[threadIdx.x + o];
            int i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    const int ref = si[0];
    if (threadIdx.x == 0 && ref_out != nullptr) ref_out[b] = ref;
    int start = ref - idx[b] - lead;
    int s0 = start >= 0 ? start : (start < -N ? 0 : N + start);
    if (s0 > N) s0 = N;
    const float fp = Fp[b], fc = Fc[b];
    float* o = out + (size_t)b * crop_len;
    for (int j = threadIdx.x; j < crop_len; j += blockDim.x) {
        int t = crop_lo + j;
        int src = s0 + t;
        float v = 0.f;
        if (src < N) v = (fp * p[src] + fc * c[src]) * win[t];
        o[j] = v * scale;
    }
}

}  // namespace gn

using namespace gn;

extern "C" int gn_fft_plan_create(int N, gn_fft_plan** out) {
    GN_REQUIRE(out != nullptr, "plan pointer is null");
    GN_REQUIRE(N >= 512 && N <= 32768 && (N & (N - 1)) == 0, "N must be a power of two in [512, 32768]");
    gn_fft_plan* p = new gn_fft_plan;
    p->N = N;
    int l = 0;
    while ((1 << l) < N / 2) ++l;
    p->log2M = l;
    std::vector<float2> h(N);
    for (int j = 0; j < N; ++j) {
        double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)j / (double)N;
        h[j] = make_float2((float)cos(ang), (float)sin(ang));
    }
    if (cudaMalloc(&p->tw, sizeof(float2) * N) != cudaSuccess) {
        delete p;
        return cuda_status("gn_fft_plan_create(cudaMalloc)") == GN_OK ? GN_ERR_CUDA : GN_ERR_CUDA;
    }
    if (cudaMemcpy(p->tw, h.data(), sizeof(float2) * N, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p->tw);
        delete p;
        cuda_status("gn_fft_plan_create(cudaMemcpy)");
        return GN_ERR_CUDA;
    }
    // per-pass Stockham twiddles: M = REM * 16^a; radix-16 pass j has sub-transform length pj = REM*16^j
    {
        const int M = N / 2;
        const int rem = 1 << (p->log2M & 3), n16 = p->log2M >> 2;
        std::vector<float2> t;
        int pj = rem;
        for (int j = 0; j < n16; ++j) {
            if (pj > 1) {
                for (int r = 1; r < 16; ++r)
                    for (int k = 0; k < pj; ++k) {
                        double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)r * (double)k / ((double)pj * 16.0);
                        t.push_back(make_float2((float)cos(ang), (float)sin(ang)));
                    }
            }
            pj *= 16;
        }
        (void)M;
        if (t.empty()) t.push_back(make_float2(1.f, 0.f));
        if (cudaMalloc(&p->ptw, sizeof(float2) * t.size()) != cudaSuccess ||
            cudaMemcpy(p->ptw, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(p->tw);
            delete p;
            cuda_status("gn_fft_plan_create(twiddle tables)");
            return GN_ERR_CUDA;
        }
    }
    for (int i = 0; i < gn_fft_plan::NSLOT; ++i) {
        p->coef[i] = nullptr;
        p->coef_key[i] = nullptr;
        p->coef_win[i] = nullptr;
        p->coef_scale[i] = 0.f;
        p->coef_stream[i] = nullptr;
        p->coef_used[i] = false;
        p->coef_done[i] = nullptr;
    }
    p->coef_next = 0;
    for (int i = 0; i < gn_fft_plan::NSLOT; ++i) {
        if (cudaMalloc(&p->coef[i], sizeof(float2) * (size_t)(N / 2 + 1)) != cudaSuccess ||      // + the flat range
            cudaEventCreateWithFlags(&p->coef_done[i], cudaEventDisableTiming) != cudaSuccess) {
            for (int q = 0; q <= i; ++q) {
                if (p->coef[q]) cudaFree(p->coef[q]);
                if (p->coef_done[q]) cudaEventDestroy(p->coef_done[q]);
            }
            cudaFree(p->tw);
            cudaFree(p->ptw);
            delete p;
            cuda_status("gn_fft_plan_create(coefficient slots)");
            return GN_ERR_CUDA;
        }
    }
    p->tw64 = nullptr;
    if (N == 8192) {
        std::vector<float2> t(64 * 64);
        for (int k1 = 0; k1 < 64; ++k1)
            for (int tt = 0; tt < 64; ++tt) {
                double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)(k1 * tt) / 4096.0;
                t[k1 * 64 + tt] = make_float2((float)cos(ang), (float)sin(ang));
            }
        if (cudaMalloc(&p->tw64, sizeof(float2) * t.size()) != cudaSuccess ||
            cudaMemcpy(p->tw64, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            for (int i = 0; i < gn_fft_plan::NSLOT; ++i) {
                cudaFree(p->coef[i]);
                cudaEventDestroy(p->coef_done[i]);
            }
            cudaFree(p->tw);
            cudaFree(p->ptw);
            delete p;
            cuda_status("gn_fft_plan_create(tw64)");
            return GN_ERR_CUDA;
        }
    }
    *out = p;
    return GN_OK;
}

extern "C" int gn_fft_plan_destroy(gn_fft_plan* p) {
    if (p == nullptr) return GN_OK;
    cudaFree(p->tw);
    cudaFree(p->ptw);
    if (p->tw64) cudaFree(p->tw64);
    for (int i = 0; i < gn_fft_plan::NSLOT; ++i) {
        cudaFree(p->coef[i]);
        cudaEventDestroy(p->coef_done[i]);
    }
    delete p;
    return GN_OK;
}

extern "C" int gn_whiten_td_f32(const gn_fft_plan* plan, const float* x, const float* window, const float* weights,
                                float* y, int batch, int crop_lo, int crop_len, float scale, void* stream) {
    GN_REQUIRE(plan && x && window && weights && y, "null pointer");
    GN_REQUIRE(batch >= 0, "batch < 0");
    GN_REQUIRE(crop_lo >= 0 && crop_len > 0 && crop_lo + crop_len <= plan->N, "crop window outside the series");
    if (batch == 0) return GN_OK;
    SynthArgs a{};
    a.x = x; a.window = window; a.weights = weights; a.y = y; a.batch = batch;
    a.crop_lo = crop_lo; a.crop_len = crop_len; a.out_scale = scale / (float)(plan->N / 2);
    return launch_synth<MODE_WHITEN>(plan, a, as_stream(stream));
}

extern "C" int gn_irfft_f32(const gn_fft_plan* plan, const float* xf, const float* weights, float* y, int batch,
                            float scale, int roll, int drop_dc, void* stream) {
    GN_REQUIRE(plan && xf && y, "null pointer");
    GN_REQUIRE(batch >= 0, "batch < 0");
    if (batch == 0) return GN_OK;
    SynthArgs a{};
    a.x = xf; a.weights = weights; a.y = y; a.batch = batch; a.drop_dc = drop_dc;
    a.roll = ((roll % plan->N) + plan->N) % plan->N;
    a.out_scale = scale / (float)(plan->N / 2);
    return launch_synth<MODE_IRFFT>(plan, a, as_stream(stream));
}

extern "C" int gn_synth_f32(const gn_fft_plan* plan, const float* normals, const float* amp, const float* templates,
                            const int* tidx, const float* window, const float* weights, float* out, int batch,
                            int n_templates, int crop_lo, int crop_len, float noise_scale, float out_scale,
                            uint64_t seed, uint64_t sample_offset, void* stream) {
    GN_REQUIRE(plan && amp && window && weights && out, "null pointer");
    GN_REQUIRE(batch >= 0, "batch < 0");
    GN_REQUIRE(templates == nullptr || n_templates > 0, "n_templates must be > 0 when templates are given");
    GN_REQUIRE(crop_lo >= 0 && crop_len > 0 && crop_lo + crop_len <= plan->N, "crop window outside the series");
    if (batch == 0) return GN_OK;
    SynthArgs a{};
    a.x = normals; a.amp = amp; a.templates = templates; a.tidx = tidx; a.window = window; a.weights = weights;
    a.y = out; a.batch = batch; a.n_templates = n_templates; a.crop_lo = crop_lo; a.crop_len = crop_len;
    // irfft of the noise spectrum: unnormalised inverse / M ; gen_noise multiplies by N*df (= noise_scale)
    a.noise_scale = noise_scale / (float)(plan->N / 2);
    a.out_scale = out_scale / (float)(plan->N / 2);
    a.seed = seed; a.sample_offset = sample_offset;
    return launch_synth<MODE_SYNTH>(plan, a, as_stream(stream));
}

extern "C" int gn_mean_std_f32(const float* x, long long n, float* out, void* stream) {
    GN_REQUIRE(x && out && n > 0, "null pointer or n <= 0");
    mean_std_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, n, out);
    return cuda_status("mean_std_kernel");
}

extern "C" int gn_add_scaled_f32(float* x, const float* r, float sigma, long long n, void* stream) {
    GN_REQUIRE(x && r && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    int grid = (int)((n + 255) / 256 < (long long)num_sms() * 8 ? (n + 255) / 256 : (long long)num_sms() * 8);
    add_scaled_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, r, sigma, n);
    return cuda_status("add_scaled_kernel");
}

extern "C" int gn_burst_waveforms_f32(const double* pars, float* out, int n, int N, float amp, float freq, float dt,
                                      float phi, void* stream) {
    GN_REQUIRE(pars && out && n >= 0 && N > 0, "null pointer or bad size");
    if (n == 0) return GN_OK;
    long long tot = (long long)n * N;
    burst_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(pars, out, n, N, amp, freq, dt, phi);
    return cuda_status("burst_kernel");
}

extern "C" int gn_bbh_assemble_f32(const float* hp, const float* hc, const float* Fp, const float* Fc, const int* idx,
                                   int lead, const float* win, float* out, int* ref_idx, int batch, int N, int crop_lo,
                                   int crop_len, float scale, void* stream) {
    GN_REQUIRE(hp && hc && Fp && Fc && idx && win && out, "null pointer");
    GN_REQUIRE(batch >= 0 && N > 0, "bad size");
    GN_REQUIRE(crop_lo >= 0 && crop_len > 0 && crop_lo + crop_len <= N, "crop window outside the series");
    if (batch == 0) return GN_OK;
    bbh_assemble_kernel<<<batch, 256, 0, as_stream(stream)>>>(hp, hc, Fp, Fc, idx, lead, win, out, ref_idx, N, crop_lo,
                                                             crop_len, scale);
    return cuda_status("bbh_assemble_kernel");
}
