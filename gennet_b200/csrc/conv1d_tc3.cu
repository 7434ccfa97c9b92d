// Conv1D as implicit GEMM on the 5th-generation tensor cores (tcgen05) AT FLOAT32 ACCURACY: split 16-bit operands.
//
//   The reference trains in float32 (cuDNN / Eigen fp32 convolutions behind the Conv1D layers of
//   bbhMahoGANy.py:250-292,362-395).  A float32 tensor is carried as NC 16-bit "planes" in one of two formats:
//     bf16 planes        x = x0 + x1 + x2,  x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)   (NC = 3: 24 mantissa bits)
//       a product a*b = sum of the plane products a_i*b_j with i + j < NC (NC = 3: six tcgen05.mma per K step, dropped
//       terms <= 2^-23 |a||b|; NC = 2: three, <= 2^-15)
//     scaled fp16 pair   t = (T0 + 2^-11 T1) / s, s = power of two from max |t| (see f16s_exp below; NC = 2)
//       three tcgen05.mma per K step, dropped term <= 2^-22 |a||b|, 22-23 bits relative to the TENSOR's scale
//   Plane products are exact in fp32, so the result differs from an fp32 FMA chain only by how the sums are rounded.
//   The tensor-memory accumulator TRUNCATES on every tcgen05.mma (measured: results biased towards zero by about
//   0.3 ulp per accumulation), so with NC > 1 a tile keeps TWO accumulators: MAIN takes only the leading products
//   (K/16 accumulations), CORR the 2^-8 (bf16) / 2^-11 (fp16) times smaller cross products (whose truncation is then
//   negligible); the epilogue adds them in float32.  Double-buffered that is 4 x BN tensor-memory columns, hence BN <= 128.
//
//   Layout: every operand tensor is stored as (NC, B, L, C) 16-bit (plane-major).  One 4-D TMA load per operand and
//   pipeline stage brings the (128 | BN) rows x BK channels x NC planes box of one tap (BK = 32: SWIZZLE_64B, three to
//   six stages; BK = 64: SWIZZLE_128B, three stages, for the two-plane formats).  Zero padding, sample boundaries, ragged
//   tails and the stride-2 sampling come from TMA out-of-bounds fill / traversal stride exactly as in conv1d_tc.cu.
//   Warp roles (320 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2-9 two epilogue sets
//   (tcgen05.ld -> bias / activation | activation-derivative mask -> fp32 rows stored straight from registers: a
//   thread owns 32 consecutive channels of one output row = one full 128-byte line -> optional re-split into the
//   bf16 planes the next convolution consumes -> optional per-channel column sums -> optional max |result|).
//
//   forward : Y[b,l,co]  = act( sum_{t,ci} X[b, l*s+t-p, ci] * W[t,ci,co] + bias[co] )      A=X   (K-major)  B=Wt[t][co][ci]
//   dgrad   : dX[b,j,ci] = act'(Xin[b,j,ci]) * sum_{t,co} dY[b,(j+p-t)/s,co] * W[t,ci,co]    A=dY  (K-major)  B=W [t][ci][co]
//   wgrad   : dW[t,ci,co] = sum_{b,l} X[b,l*s+t-p,ci] * dY[b,l,co]                           A=X^T, B=dY (both MN-major)
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace gn {

int launch_colsum(const float* x, long long rows, int C, float* out, cudaStream_t st);      // conv_dense_f32.cu

constexpr int T3_THREADS = 320;      // warps: 0 TMA producer, 1 MMA issuer, 2-5 and 6-9 epilogue (alternate 32-column slabs)
constexpr int T3_EPI_SETS = 2;
constexpr int T3_WG_THREADS = 192;   // wgrad: warps 0 producer, 1 MMA, 2-5 epilogue
constexpr int T3_SMEM_BUDGET = 232448 - 1024 /*align slack*/ - 256 /*barriers*/;

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// float32 -> NC bf16 planes (round to nearest each step; the residuals are exact in fp32)
template <int NC>
__device__ __forceinline__ void split3(float x, __nv_bfloat16 (&p)[3]) {
    p[0] = __float2bfloat16_rn(x);
    if (NC > 1) {
        const float r1 = x - __bfloat162float(p[0]);
        p[1] = __float2bfloat16_rn(r1);
        if (NC > 2) p[2] = __float2bfloat16_rn(r1 - __bfloat162float(p[1]));
    }
}

// ---- the scaled fp16 pair format ("f16x2"): a float32 tensor t with max |t| = amax is carried as two fp16 planes
//        t = (T0 + 2^-11 T1) / s,   s = 2^f16s_exp(amax),   T0 = fp16(t s),   T1 = fp16((t s - T0) 2^11)
//   s places amax in [2^14, 2^15), so T0 keeps 11 significant bits for every element down to 2^-28 amax and T1 the next
//   11 (fp16 is a normal number from 2^-14 up): 22-23 bits of t relative to the TENSOR's scale.  A product needs three
//   plane products, T0 W0 into MAIN and T0 W1 + T1 W0 into CORR (the dropped T1 W1 is <= 2^-22 |t||w|); the epilogue
//   forms (MAIN + 2^-11 CORR) / (s_t s_w).
// one element -> the 16-bit patterns of its planes in either format (F16S: `scale` = s, NC == 2)
template <int NC, bool F16S>
__device__ __forceinline__ void split_bits(float x, float scale, uint16_t (&p)[3]) {
    if (F16S) {
        split2h(x * scale, p[0], p[1]);
    } else {
        __nv_bfloat16 b[3];
        split3<NC>(x, b);
#pragma unroll
        for (int i = 0; i < NC; ++i) p[i] = __bfloat16_as_ushort(b[i]);
    }
}

struct Tc3Args {
    int B, L, Lout, Cin, Cout, k, s, p;   // convolution geometry (L = input length, Lout = output length)
    int mode;                              // 0 fwd, 1 dgrad
    int act;                               // fwd: output activation; dgrad: derivative mask taken from `aux`
    float act_param;
    const float* bias;                     // fwd, may be null
    const float* aux;                      // dgrad: the conv's own input X in fp32 (output of the previous activation) or null
    float* out;                            // fp32 result: fwd Y (B,Lout,Cout); dgrad dX (B,L,Cin); may be null
    __nv_bfloat16* planes;                 // optional: the same result re-split as (NC,B,rows,C) bf16 planes
    long long plane_stride;                // elements between planes of `planes`
    float* colsum;                         // optional: per-channel sum of the result over (b, row), atomically accumulated
    int m_tiles;                           // tiles of 128 rows per (sample, parity)
    // split-K (Dense layers with a long contraction and few output tiles; B == 1, k == 1, forward only): the "sample"
    // digit of the tile index counts K chunks of `kchunk` channels; every chunk adds its partial result into the
    // zeroed fp32 output with vector reductions, bias / activation follow in bias_act_kernel
    int kchunks, kchunk;
    // scaled fp16 pair operands (NC == 2): max |.| of the tensors behind A and B (device scalars), else null
    const float* amax_a;
    const float* amax_b;
    float* amax_out;                       // optional: max |result| over the valid elements, atomically (bits as uint)
    // optional (forward): per-channel (sum y, sum y^2) of the stored result, added into 2 * Cout doubles (zeroed by the
    // launcher) -- the BatchNormalization statistics of the layer that follows, straight from the accumulators.  A CTA
    // gathers them in shared memory (16 bytes per channel behind the barriers) and flushes once at the end.
    double* bn_sums;
};

// operand formats: the instruction descriptor's A / B format fields (bits 7, 10) are 1 for bf16, 0 for fp16
__device__ __forceinline__ uint32_t idesc_fmt(uint32_t idesc_bf16, bool f16) {
    return f16 ? (idesc_bf16 & ~((1u << 7) | (1u << 10))) : idesc_bf16;
}
// factors of the epilogue: result = (MAIN + cs * CORR) * ia * ib
struct OutScale {
    float cs, ia, ib;
};
__device__ __forceinline__ OutScale out_scale(const float* amax_a, const float* amax_b) {
    OutScale o{1.f, 1.f, 1.f};
    if (amax_a != nullptr) {
        o.cs = 1.f / F16S_LO;
        o.ia = pow2i(-f16s_exp(__ldg(amax_a)));
        o.ib = pow2i(-f16s_exp(__ldg(amax_b)));
    }
    return o;
}

// BK = contraction channels per pipeline stage: 32 (64-byte rows, SWIZZLE_64B) or 64 (128-byte rows, SWIZZLE_128B).
// The TMA unit moves a box row by row, and 64-byte rows reach only ~40 B/clk per SM against >= 60 B/clk for 128-byte
// rows (profiles/r02_conv_f16x2_step_traffic.json: the 128-wide tiles of the two-plane formats are operand-fetch bound),
// so the two-plane formats use BK = 64 wherever three stages fit; three planes keep BK = 32 for pipeline depth.
template <int BN, int NC, int BK = 32>
struct T3Smem {
    static constexpr int A_PLANE = TC_BM * BK * 2;       // 128 rows x (64 | 128) B
    static constexpr int B_PLANE = BN * BK * 2;
    static constexpr int STAGE_BYTES = NC * (A_PLANE + B_PLANE);
    static constexpr int STAGES_MAX = T3_SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

// plane pairs (i, j), i + j < NC, smallest products first
template <int NC> struct PlanePairs;
template <> struct PlanePairs<1> { static constexpr int N = 1; __host__ __device__ static constexpr int a(int) { return 0; } __host__ __device__ static constexpr int b(int) { return 0; } };
template <> struct PlanePairs<2> {
    static constexpr int N = 3;
    __host__ __device__ static constexpr int a(int q) { return q == 0 ? 1 : 0; }
    __host__ __device__ static constexpr int b(int q) { return q == 1 ? 1 : 0; }
};
template <> struct PlanePairs<3> {
    static constexpr int N = 6;      // (2,0) (1,1) (0,2) (1,0) (0,1) (0,0)
    __host__ __device__ static constexpr int a(int q) { return q == 0 ? 2 : ((q == 1 || q == 3) ? 1 : 0); }
    __host__ __device__ static constexpr int b(int q) { return q == 2 ? 2 : ((q == 1 || q == 4) ? 1 : 0); }
};

// ------------------------------------------------------------------------------------------------ fwd / dgrad
// Persistent: grid = min(#tiles, #SMs), tile = blockIdx.x + i*gridDim.x with the n-tile fastest.  Pipelines:
//   shared-memory ring (full/empty mbarriers)      TMA producer -> MMA issuer
//   TMEM accumulator ring (2 x BN fp32 columns)    MMA issuer   -> epilogue warps
template <int BN, int NC, bool AUX, int BK>
__global__ void __launch_bounds__(T3_THREADS)
conv_tc3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Tc3Args a) {
    extern __shared__ uint8_t smem_raw[];
    using S = T3Smem<BN, NC, BK>;
    constexpr int STAGES = S::STAGES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    double* sstat = reinterpret_cast<double*>(tiles + STAGES * S::STAGE_BYTES + 256);      // [2][cols] when a.bn_sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npar = (a.mode == 1) ? a.s : 1;
    const bool splitk = a.kchunks > 1;
    const int kdim = splitk ? a.kchunk : ((a.mode == 0) ? a.Cin : a.Cout);   // contraction channels (of one chunk)
    const int nkb = kdim / BK;
    const int cols = (a.mode == 0) ? a.Cout : a.Cin;   // channels of the result
    const int n_nt = cols / BN;
    const int sshift = (a.s == 2) ? 1 : 0;             // the stride is 1 or 2 (checked on the host)
    const int walk_b = splitk ? a.kchunks : a.B;       // extent of the slowest tile digit

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * T3_EPI_SETS);      // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (a.bn_sums != nullptr)
        for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) sstat[i] = 0.0;
    constexpr int ACC_COLS = (NC > 1 ? 2 : 1) * BN;      // MAIN [+ CORR] accumulator of one tile
    // two accumulator buffers (the epilogue of tile i overlaps the MMAs of tile i+1) when they fit the 512 columns;
    // MAIN + CORR at BN = 256 is single-buffered: chosen on the host only for tiles with >= 80 K steps, where the
    // exposed epilogue (~7 %) costs less than the second read of the activation tile that BN = 128 needs
    constexpr int BUFS = (2 * ACC_COLS <= 512) ? 2 : 1;
    static_assert(BUFS * ACC_COLS <= 512, "tensor memory holds 512 columns");
    if (warp == 1) tmem_alloc<BUFS * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // TMA producer: the whole warp walks the schedule (all values warp-uniform), one elected lane issues
        uint32_t st = 0, ph = 0;
        TileWalker<BN> tw;
        for (tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, walk_b); tw.valid(); tw.next()) {
            const TileCoord c = tw.coord();
            int tap_first = 0, tap_step = 1, ntaps = a.k;
            if (a.mode == 1) {
                tap_first = (c.par + a.p) & (a.s - 1);
                tap_step = a.s;
                ntaps = (a.k - tap_first + a.s - 1) >> sshift;
            }
            for (int ti = 0; ti < ntaps; ++ti) {
                const int tap = tap_first + ti * tap_step;
                // first row coordinate (global, pre-stride) along the length axis of A; in the data gradient
                // par + p - tap is a multiple of the stride by construction, so the arithmetic shift is exact
                const int rowc = (a.mode == 0) ? (c.m0 * a.s + tap - a.p) : (c.m0 + ((c.par + a.p - tap) >> sshift));
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[st], ph ^ 1);
                    if (elect_one()) {
                        uint8_t* sA = tiles + st * S::STAGE_BYTES;
                        uint8_t* sB = sA + NC * S::A_PLANE;
                        mbar_expect_tx(&full[st], S::STAGE_BYTES);
                        const int kc = kb * BK + (splitk ? c.b * a.kchunk : 0);
                        tma_load_4d(sA, &mapA, &full[st], kc, rowc, splitk ? 0 : c.b, 0);
                        tma_load_4d(sB, &mapB, &full[st], kc, c.n0, tap, 0);
                    }
                    __syncwarp();
                    if (++st == STAGES) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: per stage two K = 16 steps x the plane pairs, all into the same fp32 accumulator
        const uint32_t idesc = idesc_fmt(make_idesc(TC_BM, BN, 0, 0), a.amax_a != nullptr);
        // K-major descriptor: hi = SBO (8 rows x 64 | 128 B) | version 1 | layout type (4 = SWIZZLE_64B, 2 = SWIZZLE_128B);
        // lo = addr >> 4 | LBO; a K = 16 step advances the start address by 32 B inside the swizzle row
        constexpr uint32_t desc_hi = (BK == 32) ? ((512u >> 4) | (1u << 14) | (4u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
        using PP = PlanePairs<NC>;
        uint32_t st = 0, ph = 0;
        int ti_local = 0;
        TileWalker<BN> tw;
        for (tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, walk_b); tw.valid(); tw.next(), ++ti_local) {
            int ntaps = a.k;
            if (a.mode == 1) {
                const int tap_first = (tw.par + a.p) & (a.s - 1);
                ntaps = (a.k - tap_first + a.s - 1) >> sshift;
            }
            const int niter = ntaps * nkb;
            const int acc = (BUFS == 2) ? (ti_local & 1) : 0, acc_ph = ((BUFS == 2) ? (ti_local >> 1) : ti_local) & 1;
            mbar_wait(&tmem_empty[acc], acc_ph ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * ACC_COLS);
            for (int it = 0; it < niter; ++it) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = ((base + st * (uint32_t)S::STAGE_BYTES) >> 4);
                    const uint32_t sb = sa + (uint32_t)((NC * S::A_PLANE) >> 4);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
#pragma unroll
                        for (int q = 0; q < PP::N; ++q) {
                            const uint32_t la = ((sa + (uint32_t)((PP::a(q) * S::A_PLANE) >> 4) + 2u * ks) & 0x3FFFu) | (1u << 16);
                            const uint32_t lb = ((sb + (uint32_t)((PP::b(q) * S::B_PLANE) >> 4) + 2u * ks) & 0x3FFFu) | (1u << 16);
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)la;
                            const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)lb;
                            if (q == PP::N - 1) tc_mma_bf16(tacc, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);      // MAIN
                            else tc_mma_bf16(tacc + BN, da, db, idesc, (it > 0 || ks > 0 || q > 0) ? 1u : 0u);       // CORR
                        }
                    }
                    tc_commit(&empty[st]);      // frees the smem stage once the MMAs have read it
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
            if (elect_one()) tc_commit(&tmem_full[acc]);     // accumulator complete
            __syncwarp();
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (rows of the tile); set (w-2)/4 takes slabs set, set+2, ...
        const int q = warp & 3;
        const int eset = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        constexpr int NS = BN / 32;
        const int rows_out = (a.mode == 0) ? a.Lout : a.L;
        const OutScale os = out_scale(a.amax_a, a.amax_b);
        float amax_run = 0.f;
        int ti_local = 0;
        TileWalker<BN> tw;
        for (tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, walk_b); tw.valid(); tw.next(), ++ti_local) {
            const TileCoord c = tw.coord();
            // output row of this thread: forward l = m0 + row; data gradient j = (m0 + row) * stride + parity
            const int j = (c.m0 + row) * npar + c.par;
            const bool valid = j < rows_out;
            const size_t roff = ((size_t)(splitk ? 0 : c.b) * rows_out + (valid ? j : 0)) * cols + c.n0;
            const int acc = (BUFS == 2) ? (ti_local & 1) : 0, acc_ph = ((BUFS == 2) ? (ti_local >> 1) : ti_local) & 1;
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * ACC_COLS) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int sl = eset; sl < NS; sl += T3_EPI_SETS) {
                float mk[32];
                if (AUX) {
                    const float4* mp = reinterpret_cast<const float4*>(a.aux + roff + sl * 32);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 m = valid ? __ldg(&mp[g4]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        mk[4 * g4 + 0] = m.x; mk[4 * g4 + 1] = m.y; mk[4 * g4 + 2] = m.z; mk[4 * g4 + 3] = m.w;
                    }
                }
                float f[32];
                {
                    uint32_t v[32];
                    tmem_ld32(tacc + (uint32_t)(sl * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    if (NC > 1) {
                        tmem_ld32(tacc + (uint32_t)(BN + sl * 32), v);
                        // bf16 planes: cs = ia = ib = 1, i.e. MAIN + CORR rounded once as before
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = fmaf(__uint_as_float(v[i]), os.cs, f[i]) * os.ia * os.ib;
                    }
                }
                if (sl + T3_EPI_SETS >= NS) {
                    // all of this warp's TMEM reads are complete: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (splitk) {
                    if (valid) {
                        float* op = a.out + roff + sl * 32;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(op + 4 * g4), "f"(f[4 * g4]),
                                         "f"(f[4 * g4 + 1]), "f"(f[4 * g4 + 2]), "f"(f[4 * g4 + 3]) : "memory");
                    }
                    continue;
                }
                if (a.mode == 0) {
                    if (a.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(a.bias + c.n0 + sl * 32);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            float4 bv = __ldg(&bp[g4]);
                            f[4 * g4 + 0] += bv.x; f[4 * g4 + 1] += bv.y; f[4 * g4 + 2] += bv.z; f[4 * g4 + 3] += bv.w;
                        }
                    }
                    act_dispatch(a.act, [&](auto tag) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = act_fwd_t<decltype(tag)::kind>(f[i], a.act_param);
                    });
                } else if (AUX) {
                    act_dispatch(a.act, [&](auto tag) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] *= act_bwd_t<decltype(tag)::kind>(mk[i], a.act_param);
                    });
                }
                if (a.amax_out != nullptr && valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) amax_run = fmaxf(amax_run, fabsf(f[i]));
                }
                if (valid) {
                    if (a.out != nullptr) {
                        float4* op = reinterpret_cast<float4*>(a.out + roff + sl * 32);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) op[g4] = make_float4(f[4 * g4], f[4 * g4 + 1], f[4 * g4 + 2], f[4 * g4 + 3]);
                    }
                    if (a.planes != nullptr) {
                        uint32_t pk[3][16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            __nv_bfloat16 p0[3], p1[3];
                            split3<NC>(f[2 * i], p0);
                            split3<NC>(f[2 * i + 1], p1);
#pragma unroll
                            for (int pl = 0; pl < NC; ++pl) {
                                __nv_bfloat162 h;
                                h.x = p0[pl];
                                h.y = p1[pl];
                                pk[pl][i] = *reinterpret_cast<uint32_t*>(&h);
                            }
                        }
#pragma unroll
                        for (int pl = 0; pl < NC; ++pl) {
                            uint4* pp = reinterpret_cast<uint4*>(a.planes + (size_t)pl * a.plane_stride + roff + sl * 32);
#pragma unroll
                            for (int g = 0; g < 4; ++g) pp[g] = make_uint4(pk[pl][4 * g], pk[pl][4 * g + 1], pk[pl][4 * g + 2], pk[pl][4 * g + 3]);
                        }
                    }
                }
                if (a.colsum != nullptr) {
                    // per-channel sums over the 32 rows of this warp by a transposing butterfly (31 shuffles): after
                    // the step with offset o a lane holds partial sums of the o channels whose bit o matches its own,
                    // at the end lane l holds the full sum of channel l
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = 0.f;
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float send = up ? f[i] : f[i + o];
                            const float keep = up ? f[i + o] : f[i];
                            f[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    atomicAdd(a.colsum + c.n0 + sl * 32 + lane, f[0]);
                }
                if (a.bn_sums != nullptr) {
                    // the same butterfly twice: lane l ends with sum y and sum y^2 of channel l over the warp's 32 rows
                    // (float32 over 32 values), which go into the CTA's double accumulators
                    float g[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        f[i] = valid ? f[i] : 0.f;
                        g[i] = f[i] * f[i];
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float s1 = up ? f[i] : f[i + o], k1 = up ? f[i + o] : f[i];
                            const float s2 = up ? g[i] : g[i + o], k2 = up ? g[i + o] : g[i];
                            f[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, o);
                            g[i] = k2 + __shfl_xor_sync(0xffffffffu, s2, o);
                        }
                    }
                    atomicAdd(&sstat[c.n0 + sl * 32 + lane], (double)f[0]);
                    atomicAdd(&sstat[cols + c.n0 + sl * 32 + lane], (double)g[0]);
                }
            }
        }
        if (a.amax_out != nullptr) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) amax_run = fmaxf(amax_run, __shfl_xor_sync(0xffffffffu, amax_run, o));
            if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(a.amax_out), __float_as_uint(amax_run));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.bn_sums != nullptr)
        for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x)
            if (sstat[i] != 0.0) atomicAdd(&a.bn_sums[i], sstat[i]);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BUFS * ACC_COLS>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[t, ci, co] (fp32, atomically accumulated; caller zeroes) = sum over (b, l) of X[b, l*s+t-p, ci] * dY[b, l, co]
// Persistent split-K as conv_tc_wgrad_kernel: work unit = (K chunk, tap, 128 x BN block of (ci, co)), output tile
// fastest.  K = 32 positions per stage; both operands MN-major SWIZZLE_128B: a 64-channel block of one plane is
// 32 rows x 128 B = 4 KB, the NC planes of a block arrive with one 4-D TMA load (plane stride 4 KB), blocks are
// NC * 4 KB apart (= LBO).  The K extent of one accumulator is capped on the host (T3_WG_MAX_ITERS) so that the
// accumulation error of the fp32 tensor-memory adder stays far below the 1e-4 gradient tolerance.
struct Tc3WgradArgs {
    int B, L, Lout, Cin, Cout, k, s, p;
    int lblocks;        // ceil(Lout / 32)
    int iters_total;    // B * lblocks
    int iters_per_chunk;
    int n_chunks;
    int n_tiles_n;      // number of N tiles
    int out_tiles;      // k * m_tiles * n_tiles_n
    int rows_store;     // input channels (rows of dW per tap) that exist in dw: < Cin when the planes are zero-padded (Dense)
    float* dw;          // (k, rows_store, Cout) fp32
    const float* amax_x;   // scaled fp16 pair operands (NC == 2): max |.| of x and dy (device scalars), else null
    const float* amax_dy;
};

struct Wg3Unit {
    int tap, m0, n0, it0, niter;
};
template <int BN>
__device__ __forceinline__ Wg3Unit decode_wg3_unit(int u, const Tc3WgradArgs& a) {
    Wg3Unit w;
    const int kc = u / a.out_tiles, t = u - kc * a.out_tiles;
    w.tap = t % a.k;
    const int r = t / a.k;
    const int mt = r / a.n_tiles_n, nt = r - mt * a.n_tiles_n;
    w.m0 = mt * TC_BM;
    w.n0 = nt * BN;
    w.it0 = kc * a.iters_per_chunk;
    w.niter = min(a.iters_total - w.it0, a.iters_per_chunk);
    return w;
}

template <int BN, int NC>
struct T3WgSmem {
    static constexpr int BLK = 32 * 128;                       // one 64-channel block of one plane: 32 positions x 128 B
    static constexpr int A_BYTES = 2 * NC * BLK;               // 128 channels
    static constexpr int B_BYTES = (BN / 64) * NC * BLK;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_MAX = T3_SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int NC, bool SWAP>
__global__ void __launch_bounds__(T3_WG_THREADS)
conv_tc3_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                      Tc3WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    using S = T3WgSmem<BN, NC>;
    constexpr int STAGES = S::STAGES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = a.out_tiles * a.n_chunks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapDY);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    constexpr int ACC_COLS = (NC > 1 ? 2 : 1) * BN;      // MAIN [+ CORR] accumulator of one unit
    // two accumulator buffers when they fit the 512 tensor-memory columns; MAIN + CORR at BN = 256 runs single-buffered
    // (the reductions of a unit are exposed, about a tenth of its 64 iterations, but the 256-wide tile needs a quarter
    // less operand fetch per MMA than the 128-wide one, which is fetch bound with three MMAs per K step)
    constexpr int BUFS = (2 * ACC_COLS <= 512) ? 2 : 1;
    static_assert(BUFS * ACC_COLS <= 512, "tensor memory holds 512 columns");
    if (warp == 1) tmem_alloc<BUFS * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Wg3Unit w = decode_wg3_unit<BN>(u, a);
            int bb = w.it0 / a.lblocks, lb = w.it0 - bb * a.lblocks;
            for (int i = 0; i < w.niter; ++i) {
                mbar_wait(&empty[st], ph ^ 1);
                if (elect_one()) {
                    const int l0 = lb * 32;
                    uint8_t* sA = tiles + st * S::STAGE_BYTES;
                    uint8_t* sB = sA + S::A_BYTES;
                    mbar_expect_tx(&full[st], S::STAGE_BYTES);
                    const int xrow = l0 * a.s + w.tap - a.p;     // X position of output position l0 for this tap
                    if (!SWAP) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) tma_load_4d(sA + h * NC * S::BLK, &mapX, &full[st], w.m0 + h * 64, xrow, bb, 0);
#pragma unroll
                        for (int h = 0; h < BN / 64; ++h) tma_load_4d(sB + h * NC * S::BLK, &mapDY, &full[st], w.n0 + h * 64, l0, bb, 0);
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) tma_load_4d(sA + h * NC * S::BLK, &mapDY, &full[st], w.m0 + h * 64, l0, bb, 0);
#pragma unroll
                        for (int h = 0; h < BN / 64; ++h) tma_load_4d(sB + h * NC * S::BLK, &mapX, &full[st], w.n0 + h * 64, xrow, bb, 0);
                    }
                }
                __syncwarp();
                if (++lb == a.lblocks) { lb = 0; ++bb; }
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = idesc_fmt(make_idesc(TC_BM, BN, 1, 1), a.amax_x != nullptr);     // both operands MN-major
        // MN-major SWIZZLE_128B: 64-element MN blocks LBO = NC * 4 KB apart, 8-row K groups SBO = 1024 B apart,
        // 16 K rows per instruction = 2048 B
        constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t lbo = (uint32_t)((NC * S::BLK) >> 4) << 16;
        using PP = PlanePairs<NC>;
        uint32_t st = 0, ph = 0;
        int ul = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ul) {
            const Wg3Unit w = decode_wg3_unit<BN>(u, a);
            const int acc = (BUFS == 2) ? (ul & 1) : 0, acc_ph = ((BUFS == 2) ? (ul >> 1) : ul) & 1;
            mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * ACC_COLS);
            for (int i = 0; i < w.niter; ++i) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = ((base + st * (uint32_t)S::STAGE_BYTES) >> 4);
                    const uint32_t sb = sa + (uint32_t)(S::A_BYTES >> 4);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                        for (int q = 0; q < PP::N; ++q) {
                            const uint32_t la = ((sa + (uint32_t)((PP::a(q) * S::BLK) >> 4) + (2048u >> 4) * ks) & 0x3FFFu) | lbo;
                            const uint32_t lb = ((sb + (uint32_t)((PP::b(q) * S::BLK) >> 4) + (2048u >> 4) * ks) & 0x3FFFu) | lbo;
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)la;
                            const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)lb;
                            if (q == PP::N - 1) tc_mma_bf16(tacc, da, db, idesc, (i > 0 || ks > 0) ? 1u : 0u);        // MAIN
                            else tc_mma_bf16(tacc + BN, da, db, idesc, (i > 0 || ks > 0 || q > 0) ? 1u : 0u);         // CORR
                        }
                    }
                    tc_commit(&empty[st]);
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
            if (elect_one()) tc_commit(&tmem_full[acc]);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const OutScale os = out_scale(a.amax_x, a.amax_dy);
        int ul = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ul) {
            const Wg3Unit w = decode_wg3_unit<BN>(u, a);
            const int acc = (BUFS == 2) ? (ul & 1) : 0, acc_ph = ((BUFS == 2) ? (ul >> 1) : ul) & 1;
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * ACC_COLS) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)c0, v);
                if (NC > 1) {
                    uint32_t vc[32];
                    tmem_ld32(tacc + (uint32_t)(BN + c0), vc);
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        v[i] = __float_as_uint(fmaf(__uint_as_float(vc[i]), os.cs, __uint_as_float(v[i])) * os.ia * os.ib);
                }
                if (c0 + 32 >= BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (!SWAP) {
                    // thread = input channel, 32 consecutive output channels: eight 16-byte vector reductions
                    if (w.m0 + row < a.rows_store) {
                        float* dst = a.dw + ((size_t)w.tap * a.rows_store + (w.m0 + row)) * a.Cout + w.n0 + c0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            red_add_v4_f32(dst + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                           __uint_as_float(v[i + 3]));
                    }
                } else {
                    // thread = output channel (consecutive across the warp), columns = input channels
                    float* dst = a.dw + ((size_t)w.tap * a.rows_store + (w.n0 + c0)) * a.Cout + w.m0 + row;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (w.n0 + c0 + i < a.rows_store) atomicAdd(dst + (size_t)i * a.Cout, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BUFS * ACC_COLS>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ helpers
// max |x| of a float32 tensor into a device scalar (zeroed by the caller): non-negative floats order like their bits
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
    const long long n4 = n >> 2;
    const long long step = (long long)gridDim.x * blockDim.x;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}
static int launch_amax(const float* x, long long n, float* amax, cudaStream_t st) {
    cudaMemsetAsync(amax, 0, sizeof(float), st);
    if (n == 0) return GN_OK;
    const long long want = (n / 4 + 255) / 256;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want < 8LL * num_sms() ? want : 8LL * num_sms()));
    amax_kernel<<<grid, 256, 0, st>>>(x, n, amax);
    return cuda_status("amax_kernel");
}

// float32 tensor -> 16-bit planes (plane p at planes + p * n): NC bf16 planes, or the scaled fp16 pair (F16S, NC == 2,
// `amax` = device scalar holding max |x|).  Eight elements per thread and step.
template <int NC, bool F16S>
__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes,
                                                        long long n, const float* __restrict__ amax) {
    const float scale = F16S ? pow2i(f16s_exp(__ldg(amax))) : 1.f;
    const long long n8 = n >> 3;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += step) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
        const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        uint32_t pk[3][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint16_t p0[3], p1[3];
            split_bits<NC, F16S>(f[2 * e], scale, p0);
            split_bits<NC, F16S>(f[2 * e + 1], scale, p1);
#pragma unroll
            for (int pl = 0; pl < NC; ++pl) pk[pl][e] = (uint32_t)p0[pl] | ((uint32_t)p1[pl] << 16);
        }
#pragma unroll
        for (int pl = 0; pl < NC; ++pl)
            reinterpret_cast<uint4*>(planes + (size_t)pl * n)[i] = make_uint4(pk[pl][0], pk[pl][1], pk[pl][2], pk[pl][3]);
    }
    // tail (n % 8 elements)
    const long long t0 = n8 << 3;
    for (long long i = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        uint16_t p[3];
        split_bits<NC, F16S>(x[i], scale, p);
#pragma unroll
        for (int pl = 0; pl < NC; ++pl) planes[(size_t)pl * n + i] = p[pl];
    }
}

// The same split for a gradient tensor dy (rows, C) that ALSO leaves its per-channel column sums = the bias gradient of
// the layer whose output gradient it is (the weight-gradient call then skips its own pass over the float32 dy).  A thread
// keeps one group of 8 channels and walks the rows with a stride of `lanes`; the row lanes of a block meet in shared
// memory, one atomic per (block, channel) onto colsum (zeroed by the launcher).  C % 8 == 0, 256 % (C / 8) == 0 or
// (C / 8) % 256 == 0.
__global__ void __launch_bounds__(256) split_colsum_f16s_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes,
                                                                long long rows, int C, long long lanes,
                                                                const float* __restrict__ amax, float* __restrict__ colsum) {
    __shared__ float red[256][9];
    const float scale = pow2i(f16s_exp(__ldg(amax)));
    const int C8 = C / 8;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = (int)(tid % C8);
    const long long lane = tid / C8;
    const size_t n = (size_t)rows * C;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    if (lane < lanes) {
#pragma unroll 2
        for (long long r = lane; r < rows; r += lanes) {
            const size_t i0 = (size_t)r * C + (size_t)cg * 8;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(x + i0));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(x + i0) + 1);
            const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            uint32_t p0[4], p1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint16_t u0, u1, v0, v1;
                split2h(f[2 * e] * scale, u0, u1);
                split2h(f[2 * e + 1] * scale, v0, v1);
                p0[e] = (uint32_t)u0 | ((uint32_t)v0 << 16);
                p1[e] = (uint32_t)u1 | ((uint32_t)v1 << 16);
                s[2 * e] += f[2 * e];
                s[2 * e + 1] += f[2 * e + 1];
            }
            *reinterpret_cast<uint4*>(planes + i0) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
            *reinterpret_cast<uint4*>(planes + n + i0) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
        }
    }
    // fold the row lanes of this block that share a channel group (threads t, t + C8, t + 2 C8, ... when C8 < 256)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = s[e];
    __syncthreads();
    const int per = C8 < 256 ? 256 / C8 : 1;      // row lanes per block
    const int ngroups = C8 < 256 ? C8 : 256;       // distinct channel groups in this block
    for (int t = threadIdx.x; t < ngroups * 8; t += blockDim.x) {
        const int g = t >> 3, e = t & 7;
        float v = 0.f;
        for (int q = 0; q < per; ++q) v += red[g + q * C8][e];
        const int cgo = (int)(((long long)blockIdx.x * blockDim.x + g) % C8);
        if (v != 0.f) atomicAdd(&colsum[cgo * 8 + e], v);
    }
}

// y[r, c] = act(y[r, c] + bias[c]) in place (second pass of a split-K Dense forward); C % 4 == 0
__global__ void __launch_bounds__(256) bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n4, int C4,
                                                       int act, float ap) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(y)[i];
        if (bias != nullptr) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % C4));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        v.x = act_fwd(v.x, act, ap); v.y = act_fwd(v.y, act, ap); v.z = act_fwd(v.z, act, ap); v.w = act_fwd(v.w, act, ap);
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// float32 matrix (rows, K) -> planes of a (rows, Kp) matrix, columns K..Kp-1 zero (Dense operands whose feature
// count is not a multiple of the 64-channel tile, e.g. the generator's Dense(100 -> 128 n_pix), bbhMahoGANy.py:234)
template <int NC, bool F16S>
__global__ void __launch_bounds__(256) split_pad_f32_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes,
                                                            long long rows, int K, int Kp, const float* __restrict__ amax) {
    const float scale = F16S ? pow2i(f16s_exp(__ldg(amax))) : 1.f;
    const long long n = rows * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / Kp;
        const int c = (int)(i - r * Kp);
        uint16_t p[3];
        split_bits<NC, F16S>(c < K ? x[r * K + c] : 0.f, scale, p);
#pragma unroll
        for (int pl = 0; pl < NC; ++pl) planes[(size_t)pl * n + i] = p[pl];
    }
}

// weights: f32 (k,Cin,Cout) -> planes of the same layout (dgrad B operand) and planes of the transposed layout
// (k,Cout,Cin) (forward B operand); one 32 x 32 (ci, co) tile of one tap per block, transposed through shared memory
// Cin_src <= Cin: rows Cin_src..Cin-1 of every tap are written as zeros (w holds (k, Cin_src, Cout))
template <int NC, bool F16S>
__global__ void __launch_bounds__(256) conv_w_split_kernel(const float* __restrict__ w, uint16_t* __restrict__ wk,
                                                           uint16_t* __restrict__ wt, int k, int Cin, int Cout,
                                                           int Cin_src, const float* __restrict__ amax) {
    __shared__ uint16_t tile[NC][32][33];
    const float scale = F16S ? pow2i(f16s_exp(__ldg(amax))) : 1.f;
    const int t = blockIdx.z;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const size_t plane = (size_t)k * Cin * Cout;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) {
            const size_t i = ((size_t)t * Cin + ci) * Cout + co;
            uint16_t p[3];
            split_bits<NC, F16S>(ci < Cin_src ? w[((size_t)t * Cin_src + ci) * Cout + co] : 0.f, scale, p);
#pragma unroll
            for (int pl = 0; pl < NC; ++pl) {
                wk[pl * plane + i] = p[pl];
                tile[pl][r][tx] = p[pl];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        if (ci < Cin && co < Cout) {
#pragma unroll
            for (int pl = 0; pl < NC; ++pl) wt[pl * plane + ((size_t)t * Cout + co) * Cin + ci] = tile[pl][tx][r];
        }
    }
}

// 4-D bf16 tensor map over a plane-major tensor (NC, d2, d1, d0): dims (d0 contiguous, d1, d2, NC); box
// (b0, b1 rows, 1, NC); traversal stride es1 along d1 (box extent b1*es1 in global coordinates -> b1 rows in smem)
static int make_map4(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, int nc, uint64_t st1,
                     uint64_t st2, uint64_t st3, uint32_t b0, uint32_t b1, uint32_t es1, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = tc_encode_fn();
    if (fn == nullptr) return fail(GN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available%s", "");
    cuuint64_t dims[4] = {d0, d1, d2, (cuuint64_t)nc};
    cuuint64_t strides[3] = {st1 * 2, st2 * 2, st3 * 2};
    cuuint32_t box[4] = {b0, b1 * es1, 1, (cuuint32_t)nc};
    cuuint32_t estr[4] = {1, es1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GN_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%s code %lld)", "", (long long)r);
    return GN_OK;
}

template <int BN, int NC, bool AUX, int BK>
static int launch_conv_tc3(const CUtensorMap& mA, const CUtensorMap& mB, const Tc3Args& a, long long total_tiles,
                           cudaStream_t st) {
    auto kfn = conv_tc3_kernel<BN, NC, AUX, BK>;
    constexpr int smem = T3Smem<BN, NC, BK>::TOTAL;
    static_assert(smem <= 232448, "shared memory budget exceeded");
    static_assert(T3Smem<BN, NC, BK>::STAGES >= 3, "pipeline too shallow");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        attr_set = true;
    }
    const int grid = (int)(total_tiles < (long long)num_sms() ? total_tiles : (long long)num_sms());
    // the BatchNormalization statistics take 16 bytes of shared memory per output channel behind the barriers
    const int extra = (a.bn_sums != nullptr) ? 16 * ((a.mode == 0) ? a.Cout : a.Cin) : 0;
    if (smem + extra > 232448) return fail(GN_ERR_UNSUPPORTED, "statistics accumulators do not fit shared memory%s", "");
    kfn<<<grid, T3_THREADS, smem + extra, st>>>(mA, mB, a);
    return cuda_status("conv_tc3_kernel");
}

// channels per stage: 64 for the two-plane formats on tiles up to 128 wide (three 64 KB stages), else 32
static int pick_bk3(int BN, int nc) { return (nc == 2 && BN <= 128) ? 64 : 32; }

template <int NC>
static int dispatch_conv_tc3(int BN, bool aux, const CUtensorMap& mA, const CUtensorMap& mB, const Tc3Args& a,
                             long long tiles, cudaStream_t st) {
    if constexpr (NC == 2) {
        if (BN == 128) return aux ? launch_conv_tc3<128, NC, true, 64>(mA, mB, a, tiles, st)
                                  : launch_conv_tc3<128, NC, false, 64>(mA, mB, a, tiles, st);
        if (BN == 64) return aux ? launch_conv_tc3<64, NC, true, 64>(mA, mB, a, tiles, st)
                                 : launch_conv_tc3<64, NC, false, 64>(mA, mB, a, tiles, st);
    }
    if (!aux) {
        if (BN == 256) return launch_conv_tc3<256, NC, false, 32>(mA, mB, a, tiles, st);
        if (BN == 128) return launch_conv_tc3<128, NC, false, 32>(mA, mB, a, tiles, st);
        return launch_conv_tc3<64, NC, false, 32>(mA, mB, a, tiles, st);
    }
    if (BN == 256) return launch_conv_tc3<256, NC, true, 32>(mA, mB, a, tiles, st);
    if (BN == 128) return launch_conv_tc3<128, NC, true, 32>(mA, mB, a, tiles, st);
    return launch_conv_tc3<64, NC, true, 32>(mA, mB, a, tiles, st);
}
// widest N tile.  nc > 1: MAIN + CORR accumulators double-buffered need 4 x BN tensor-memory columns, so BN = 256 runs
// single-buffered and is taken only when a tile has enough K steps (ksteps = taps * K / 16) to hide its epilogue
static int pick_bn3(int C, int nc, int ksteps) {
    if (C % 256 == 0 && (nc == 1 || ksteps >= 80)) return 256;
    return (C % 128 == 0) ? 128 : 64;
}

template <int BN, int NC, bool SWAP>
static int launch_wgrad_tc3(const CUtensorMap& mX, const CUtensorMap& mDY, const Tc3WgradArgs& a, int grid,
                            cudaStream_t st) {
    auto kfn = conv_tc3_wgrad_kernel<BN, NC, SWAP>;
    constexpr int smem = T3WgSmem<BN, NC>::TOTAL;
    static_assert(smem <= 232448, "shared memory budget exceeded");
    static_assert(T3WgSmem<BN, NC>::STAGES >= 3, "pipeline too shallow");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    kfn<<<grid, T3_WG_THREADS, smem, st>>>(mX, mDY, a);
    return cuda_status("conv_tc3_wgrad_kernel");
}

template <int NC>
static int dispatch_wgrad_tc3(int BN, bool swap, const CUtensorMap& mX, const CUtensorMap& mDY, const Tc3WgradArgs& a,
                              int grid, cudaStream_t st) {
    if (swap) return launch_wgrad_tc3<64, NC, true>(mX, mDY, a, grid, st);
    if constexpr (NC <= 2) {
        if (BN == 256) return launch_wgrad_tc3<256, NC, false>(mX, mDY, a, grid, st);
    }
    if (BN == 128) return launch_wgrad_tc3<128, NC, false>(mX, mDY, a, grid, st);
    return launch_wgrad_tc3<64, NC, false>(mX, mDY, a, grid, st);
}

// K chunks of the persistent split-K wgrad: minimise waves * (iterations per chunk + fixed cost per unit), with the
// iterations of one accumulator capped at T3_WG_MAX_ITERS (K = 32 positions each)
constexpr int T3_WG_MAX_ITERS = 64;      // 2048 positions = 128 truncating accumulations of the MAIN products
static int pick_wgrad3_chunks(int out_tiles, int iters_total, int grid) {
    const int ovh = 8;
    int lo = (iters_total + T3_WG_MAX_ITERS - 1) / T3_WG_MAX_ITERS, hi = (iters_total + 31) / 32;
    if (lo < 1) lo = 1;
    if (hi < lo) hi = lo;
    if (hi > 8192) hi = 8192;
    long long best_cost = -1;
    int best = lo;
    for (int nk = lo; nk <= hi; ++nk) {
        const int ipc = (iters_total + nk - 1) / nk;
        const int nk_eff = (iters_total + ipc - 1) / ipc;
        const long long units = (long long)out_tiles * nk_eff;
        const long long waves = (units + grid - 1) / grid;
        const long long cost = waves * (ipc + ovh);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nk_eff; }
    }
    return best;
}

static int check_tc3_geom(int B, int L, int Cin, int Lout, int Cout, int k, int s, int p, int nc) {
    GN_REQUIRE(B > 0 && L > 0 && Lout > 0 && k > 0 && k <= 16 && (s == 1 || s == 2) && p >= 0 && p < k, "bad geometry");
    GN_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tensor-core path needs Cin and Cout to be multiples of 64");
    GN_REQUIRE(B <= 65535, "batch too large for one launch");
    GN_REQUIRE(nc >= 1 && nc <= 3, "number of bf16 planes must be 1, 2 or 3");
    return GN_OK;
}

}  // namespace gn

using namespace gn;

// ---- operand conversion.  amax == nullptr: nc bf16 planes; else the scaled fp16 pair (two planes) with the tensor's
// max |x| in the device scalar `amax` (computed here unless have_amax: a producer already accumulated it)
static int split_impl(const float* x, void* planes, long long n, int nc, float* amax, int have_amax, void* stream) {
    GN_REQUIRE(x && planes && n >= 0 && nc >= 1 && nc <= 3, "null pointer, n < 0 or planes not in 1..3");
    GN_REQUIRE(n % 8 == 0 || nc == 1, "element count must be a multiple of 8 (16-byte aligned planes)");
    if (n == 0) return GN_OK;
    const long long want = (n / 8 + 255) / 256;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want < 16LL * num_sms() ? want : 16LL * num_sms()));
    cudaStream_t st = as_stream(stream);
    uint16_t* pl = (uint16_t*)planes;
    if (amax != nullptr) {
        if (!have_amax) {
            int rc = launch_amax(x, n, amax, st);
            if (rc != GN_OK) return rc;
        }
        split_f32_kernel<2, true><<<grid, 256, 0, st>>>(x, pl, n, amax);
    } else if (nc == 3) split_f32_kernel<3, false><<<grid, 256, 0, st>>>(x, pl, n, nullptr);
    else if (nc == 2) split_f32_kernel<2, false><<<grid, 256, 0, st>>>(x, pl, n, nullptr);
    else split_f32_kernel<1, false><<<grid, 256, 0, st>>>(x, pl, n, nullptr);
    return cuda_status("split_f32_kernel");
}
extern "C" int gn_split_f32_bf16(const float* x, void* planes, long long n, int nc, void* stream) {
    return split_impl(x, planes, n, nc, nullptr, 0, stream);
}
extern "C" int gn_split_f32_f16x2(const float* x, void* planes, float* amax, int have_amax, long long n, void* stream) {
    GN_REQUIRE(amax, "null pointer");
    return split_impl(x, planes, n, 2, amax, have_amax, stream);
}
extern "C" int gn_amax_f32(const float* x, long long n, float* amax, void* stream) {
    GN_REQUIRE(x && amax && n >= 0, "null pointer or n < 0");
    return launch_amax(x, n, amax, as_stream(stream));
}

extern "C" int gn_split_colsum_f32_f16x2(const float* x, void* planes, float* amax, int have_amax, long long rows, int C,
                                         float* colsum, void* stream) {
    GN_REQUIRE(x && planes && amax && colsum && rows >= 0 && C > 0, "null pointer or bad size");
    GN_REQUIRE(C % 8 == 0 && (256 % (C / 8) == 0 || (C / 8) % 256 == 0),
               "C / 8 must divide 256 or be a multiple of it");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)C, st);
    if (rows == 0) return GN_OK;
    if (!have_amax) {
        int rc = launch_amax(x, rows * C, amax, st);
        if (rc != GN_OK) return rc;
    }
    // about four CTAs of 256 threads per SM: long-lived threads keep the atomics per channel in the hundreds
    const long long C8 = C / 8;
    long long lanes = 4LL * num_sms() * 256 / C8;
    if (lanes < 1) lanes = 1;
    if (lanes > rows) lanes = rows;
    if (C8 < 256) lanes = (lanes + 256 / C8 - 1) / (256 / C8) * (256 / C8);      // whole blocks of row lanes
    const unsigned grid = (unsigned)((lanes * C8 + 255) / 256);
    split_colsum_f16s_kernel<<<grid, 256, 0, st>>>(x, (uint16_t*)planes, rows, C, lanes, amax, colsum);
    return cuda_status("split_colsum_f16s_kernel");
}

static int w_split_impl(const float* w, void* wk, void* wt, int k, int Cin, int Cout, int Cin_src, int nc, float* amax,
                        void* stream) {
    GN_REQUIRE(w && wk && wt && k > 0 && Cin > 0 && Cout > 0 && Cin_src > 0 && Cin_src <= Cin && nc >= 1 && nc <= 3,
               "null pointer or bad size");
    GN_REQUIRE(k <= 65535 && (Cin + 31) / 32 <= 65535, "weight tensor too large for the split grid");
    dim3 grid((unsigned)((Cout + 31) / 32), (unsigned)((Cin + 31) / 32), (unsigned)k);
    cudaStream_t st = as_stream(stream);
    uint16_t *a = (uint16_t*)wk, *b = (uint16_t*)wt;
    if (amax != nullptr) {
        int rc = launch_amax(w, (long long)k * Cin_src * Cout, amax, st);
        if (rc != GN_OK) return rc;
        conv_w_split_kernel<2, true><<<grid, 256, 0, st>>>(w, a, b, k, Cin, Cout, Cin_src, amax);
    } else if (nc == 3) conv_w_split_kernel<3, false><<<grid, 256, 0, st>>>(w, a, b, k, Cin, Cout, Cin_src, nullptr);
    else if (nc == 2) conv_w_split_kernel<2, false><<<grid, 256, 0, st>>>(w, a, b, k, Cin, Cout, Cin_src, nullptr);
    else conv_w_split_kernel<1, false><<<grid, 256, 0, st>>>(w, a, b, k, Cin, Cout, Cin_src, nullptr);
    return cuda_status("conv_w_split_kernel");
}
extern "C" int gn_conv_w_split_bf16(const float* w, void* wk, void* wt, int k, int Cin, int Cout, int nc, void* stream) {
    return w_split_impl(w, wk, wt, k, Cin, Cout, Cin, nc, nullptr, stream);
}
extern "C" int gn_conv_w_split_f16x2(const float* w, void* wk, void* wt, float* amax, int k, int Cin, int Cout,
                                     void* stream) {
    GN_REQUIRE(amax, "null pointer");
    return w_split_impl(w, wk, wt, k, Cin, Cout, Cin, 2, amax, stream);
}

// operand format of a compute call: nc bf16 planes (amax pointers null) or scaled fp16 pairs
struct Fmt3 {
    int nc;
    const float* amax_a;      // tensor behind the A operand (x or dy)
    const float* amax_b;      // tensor behind the B operand (w; dy in the weight gradient)
};
static int check_fmt(const Fmt3& f) {
    GN_REQUIRE((f.amax_a == nullptr) == (f.amax_b == nullptr), "both operands must be in the same plane format");
    GN_REQUIRE(f.amax_a == nullptr || f.nc == 2, "scaled fp16 operands come as two planes");
    return GN_OK;
}

static int fwd_tc3(const void* xs, const void* wts, const float* bias, float* y, void* ys, float* y_amax, double* bn_sums,
                   int B, int L, int Cin, int Lout, int Cout, int k, int stride, int pad_left, int act, float act_param, Fmt3 f,
                   int kchunks, void* stream) {
    GN_REQUIRE(xs && wts && (y || ys), "null pointer");
    const int nc = f.nc;
    int rc = check_tc3_geom(B, L, Cin, Lout, Cout, k, stride, pad_left, nc);
    if (rc != GN_OK) return rc;
    if ((rc = check_fmt(f)) != GN_OK) return rc;
    GN_REQUIRE(f.amax_a == nullptr || ys == nullptr, "the epilogue re-split exists for bf16 planes only");
    CUtensorMap mA, mB;
    const int BN = pick_bn3(Cout, nc, k * Cin / 16);
    // A: X planes viewed as (Cin, L, B, NC); 128 output rows per tile, traversal stride = conv stride
    const int BK = pick_bk3(BN, nc);
    const CUtensorMapSwizzle swz = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    rc = make_map4(&mA, xs, Cin, L, B, nc, Cin, (uint64_t)L * Cin, (uint64_t)B * L * Cin, BK, TC_BM, stride, swz);
    if (rc != GN_OK) return rc;
    // B: Wt planes (NC, k, Cout, Cin) viewed as (Cin, Cout, k, NC)
    rc = make_map4(&mB, wts, Cin, Cout, k, nc, Cin, (uint64_t)Cout * Cin, (uint64_t)k * Cout * Cin, BK, BN, 1, swz);
    if (rc != GN_OK) return rc;
    Tc3Args a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.mode = 0; a.act = act; a.act_param = act_param; a.bias = bias; a.out = y;
    a.planes = (__nv_bfloat16*)ys; a.plane_stride = (long long)B * Lout * Cout;
    a.amax_a = f.amax_a; a.amax_b = f.amax_b; a.amax_out = y_amax; a.bn_sums = bn_sums;
    a.m_tiles = (Lout + TC_BM - 1) / TC_BM;
    a.kchunks = kchunks;
    a.kchunk = Cin / kchunks;
    GN_REQUIRE(kchunks == 1 || (B == 1 && k == 1 && ys == nullptr && y != nullptr && y_amax == nullptr && bn_sums == nullptr &&
                                a.kchunk % BK == 0),
               "split-K needs B == 1, k == 1, a float32 output and chunks of whole channel blocks");
    const long long tiles = (long long)(kchunks > 1 ? kchunks : B) * a.m_tiles * (Cout / BN);
    cudaStream_t st = as_stream(stream);
    if (y_amax != nullptr) cudaMemsetAsync(y_amax, 0, sizeof(float), st);
    if (bn_sums != nullptr) cudaMemsetAsync(bn_sums, 0, sizeof(double) * 2 * (size_t)Cout, st);
    if (nc == 3) return dispatch_conv_tc3<3>(BN, false, mA, mB, a, tiles, st);
    if (nc == 2) return dispatch_conv_tc3<2>(BN, false, mA, mB, a, tiles, st);
    return dispatch_conv_tc3<1>(BN, false, mA, mB, a, tiles, st);
}

extern "C" int gn_conv1d_fwd_bf16x3(const void* xs, const void* wts, const float* bias, float* y, void* ys, int B, int L,
                                    int Cin, int Lout, int Cout, int k, int stride, int pad_left, int act,
                                    float act_param, int nc, void* stream) {
    return fwd_tc3(xs, wts, bias, y, ys, nullptr, nullptr, B, L, Cin, Lout, Cout, k, stride, pad_left, act, act_param,
                   Fmt3{nc, nullptr, nullptr}, 1, stream);
}
extern "C" int gn_conv1d_fwd_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax,
                                   const float* bias, float* y, float* y_amax, int B, int L, int Cin, int Lout, int Cout,
                                   int k, int stride, int pad_left, int act, float act_param, void* stream) {
    GN_REQUIRE(x_amax && w_amax, "null pointer");
    return fwd_tc3(xs, wts, bias, y, nullptr, y_amax, nullptr, B, L, Cin, Lout, Cout, k, stride, pad_left, act, act_param,
                   Fmt3{2, x_amax, w_amax}, 1, stream);
}
extern "C" int gn_conv1d_fwd_stats_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax,
                                         const float* bias, float* y, double* y_sums, int B, int L, int Cin, int Lout,
                                         int Cout, int k, int stride, int pad_left, int act, float act_param, void* stream) {
    GN_REQUIRE(x_amax && w_amax && y && y_sums, "null pointer");
    GN_REQUIRE(Cout <= 1024, "the statistics accumulators hold up to 1024 channels");
    return fwd_tc3(xs, wts, bias, y, nullptr, nullptr, y_sums, B, L, Cin, Lout, Cout, k, stride, pad_left, act, act_param,
                   Fmt3{2, x_amax, w_amax}, 1, stream);
}

static int dgrad_tc3(const void* dys, const void* wks, const float* x_in, float* dx, void* dxs, float* dx_colsum,
                     float* dx_amax, int B, int L, int Cin, int Lout, int Cout, int k, int stride, int pad_left, int in_act,
                     float in_act_param, Fmt3 f, void* stream) {
    GN_REQUIRE(dys && wks && (dx || dxs), "null pointer");
    const int nc = f.nc;
    int rc = check_tc3_geom(B, L, Cin, Lout, Cout, k, stride, pad_left, nc);
    if (rc != GN_OK) return rc;
    if ((rc = check_fmt(f)) != GN_OK) return rc;
    GN_REQUIRE(f.amax_a == nullptr || dxs == nullptr, "the epilogue re-split exists for bf16 planes only");
    CUtensorMap mA, mB;
    // mask, column sums and re-split make this epilogue long: the single-buffered 256-wide tile pays only for the
    // two-plane formats (three MMAs per K step leave the 128-wide tile operand-fetch bound) on tiles with >= 160 K steps
    // (measured: 512<-1024 and 256<-512 stride 1 gain 8-10 %, 256<-512 stride 2 with 80 steps loses 18 %)
    const int BN = (nc == 2 && Cin % 256 == 0 && k * Cout / 16 / stride >= 160) ? 256 : pick_bn3(Cin, nc, 0);
    // A: dY planes viewed as (Cout, Lout, B, NC), 128 rows, unit traversal stride (parity classes handle the conv stride)
    const int BK = pick_bk3(BN, nc);
    const CUtensorMapSwizzle swz = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    rc = make_map4(&mA, dys, Cout, Lout, B, nc, Cout, (uint64_t)Lout * Cout, (uint64_t)B * Lout * Cout, BK, TC_BM, 1, swz);
    if (rc != GN_OK) return rc;
    // B: W planes (NC, k, Cin, Cout) viewed as (Cout, Cin, k, NC)
    rc = make_map4(&mB, wks, Cout, Cin, k, nc, Cout, (uint64_t)Cin * Cout, (uint64_t)k * Cin * Cout, BK, BN, 1, swz);
    if (rc != GN_OK) return rc;
    cudaStream_t st = as_stream(stream);
    Tc3Args a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.mode = 1; a.act = in_act; a.act_param = in_act_param; a.aux = x_in; a.out = dx;
    a.planes = (__nv_bfloat16*)dxs; a.plane_stride = (long long)B * L * Cin;
    a.colsum = dx_colsum;
    a.amax_a = f.amax_a; a.amax_b = f.amax_b; a.amax_out = dx_amax;
    if (dx_colsum != nullptr) cudaMemsetAsync(dx_colsum, 0, sizeof(float) * (size_t)Cin, st);
    if (dx_amax != nullptr) cudaMemsetAsync(dx_amax, 0, sizeof(float), st);
    const int rows = (L + stride - 1) / stride;      // rows of the largest parity class
    a.m_tiles = (rows + TC_BM - 1) / TC_BM;
    const bool aux = (x_in != nullptr && in_act != GN_ACT_NONE);
    const long long tiles = (long long)B * stride * a.m_tiles * (Cin / BN);
    if (nc == 3) return dispatch_conv_tc3<3>(BN, aux, mA, mB, a, tiles, st);
    if (nc == 2) return dispatch_conv_tc3<2>(BN, aux, mA, mB, a, tiles, st);
    return dispatch_conv_tc3<1>(BN, aux, mA, mB, a, tiles, st);
}

extern "C" int gn_conv1d_dgrad_bf16x3(const void* dys, const void* wks, const float* x_in, float* dx, void* dxs,
                                      float* dx_colsum, int B, int L, int Cin, int Lout, int Cout, int k, int stride,
                                      int pad_left, int in_act, float in_act_param, int nc, void* stream) {
    return dgrad_tc3(dys, wks, x_in, dx, dxs, dx_colsum, nullptr, B, L, Cin, Lout, Cout, k, stride, pad_left, in_act,
                     in_act_param, Fmt3{nc, nullptr, nullptr}, stream);
}
extern "C" int gn_conv1d_dgrad_f16x2(const void* dys, const float* dy_amax, const void* wks, const float* w_amax,
                                     const float* x_in, float* dx, float* dx_colsum, float* dx_amax, int B, int L, int Cin,
                                     int Lout, int Cout, int k, int stride, int pad_left, int in_act, float in_act_param,
                                     void* stream) {
    GN_REQUIRE(dy_amax && w_amax, "null pointer");
    return dgrad_tc3(dys, wks, x_in, dx, nullptr, dx_colsum, dx_amax, B, L, Cin, Lout, Cout, k, stride, pad_left, in_act,
                     in_act_param, Fmt3{2, dy_amax, w_amax}, stream);
}

static int wgrad_tc3(const void* xs, const void* dys, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                     int Cout, int k, int stride, int pad_left, Fmt3 f, int rows_store, void* stream) {
    GN_REQUIRE(xs && dys && dw, "null pointer");
    GN_REQUIRE(db == nullptr || dy != nullptr, "the bias gradient needs the float32 dy");
    const int nc = f.nc;
    int rc = check_tc3_geom(B, L, Cin, Lout, Cout, k, stride, pad_left, nc);
    if (rc != GN_OK) return rc;
    if ((rc = check_fmt(f)) != GN_OK) return rc;
    GN_REQUIRE(Cin % 128 == 0 || (Cin == 64 && Cout % 128 == 0), "wgrad needs Cin % 128 == 0, or Cin == 64 with Cout % 128 == 0");
    cudaStream_t st = as_stream(stream);
    CUtensorMap mX, mDY;
    // X planes (Cin, L, B, NC): 64 channels x 32 positions (traversal stride = conv stride); dY planes (Cout, Lout, B, NC)
    rc = make_map4(&mX, xs, Cin, L, B, nc, Cin, (uint64_t)L * Cin, (uint64_t)B * L * Cin, 64, 32, stride,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != GN_OK) return rc;
    rc = make_map4(&mDY, dys, Cout, Lout, B, nc, Cout, (uint64_t)Lout * Cout, (uint64_t)B * Lout * Cout, 64, 32, 1,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != GN_OK) return rc;
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * rows_store * Cout, st);
    Tc3WgradArgs a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.rows_store = rows_store;
    a.lblocks = (Lout + 31) / 32;
    a.iters_total = B * a.lblocks;
    a.dw = dw;
    a.amax_x = f.amax_a; a.amax_dy = f.amax_b;
    const bool swap = (Cin == 64);
    int m_tiles, BN;
    if (!swap) {
        m_tiles = Cin / 128;
        BN = pick_bn3(Cout, nc, 0);
        // two-plane formats: the single-buffered 256-wide tile (a quarter less operand fetch per MMA) beats the
        // double-buffered 128-wide one on every layer measured (GAN -0.7 ms, PE -0.9 ms per step)
        if (nc == 2 && Cout % 256 == 0) BN = 256;
        a.n_tiles_n = Cout / BN;
    }
    else { m_tiles = Cout / 128; BN = 64; a.n_tiles_n = 1; }
    a.out_tiles = k * m_tiles * a.n_tiles_n;
    const int nsm = num_sms();
    int nk = pick_wgrad3_chunks(a.out_tiles, a.iters_total, nsm);
    a.iters_per_chunk = (a.iters_total + nk - 1) / nk;
    a.n_chunks = (a.iters_total + a.iters_per_chunk - 1) / a.iters_per_chunk;
    const long long units = (long long)a.out_tiles * a.n_chunks;
    const int grid = (int)(units < nsm ? units : nsm);
    if (nc == 3) rc = dispatch_wgrad_tc3<3>(BN, swap, mX, mDY, a, grid, st);
    else if (nc == 2) rc = dispatch_wgrad_tc3<2>(BN, swap, mX, mDY, a, grid, st);
    else rc = dispatch_wgrad_tc3<1>(BN, swap, mX, mDY, a, grid, st);
    if (rc != GN_OK) return rc;
    if (db != nullptr) return launch_colsum(dy, (long long)B * Lout, Cout, db, st);
    return GN_OK;
}

extern "C" int gn_conv1d_wgrad_bf16x3(const void* xs, const void* dys, const float* dy, float* dw, float* db, int B, int L,
                                      int Cin, int Lout, int Cout, int k, int stride, int pad_left, int nc, void* stream) {
    return wgrad_tc3(xs, dys, dy, dw, db, B, L, Cin, Lout, Cout, k, stride, pad_left, Fmt3{nc, nullptr, nullptr}, Cin, stream);
}
extern "C" int gn_conv1d_wgrad_f16x2(const void* xs, const float* x_amax, const void* dys, const float* dy_amax,
                                     const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout, int Cout, int k,
                                     int stride, int pad_left, void* stream) {
    GN_REQUIRE(x_amax && dy_amax, "null pointer");
    return wgrad_tc3(xs, dys, dy, dw, db, B, L, Cin, Lout, Cout, k, stride, pad_left, Fmt3{2, x_amax, dy_amax}, Cin, stream);
}

// ---- Dense layers on the same kernels: a Dense GEMM is a convolution with one tap over a single "sample" whose
// positions are the batch rows (M).  Kp = feature count padded to a multiple of 64 in the PLANES only.
static int split_pad_impl(const float* x, void* planes, long long rows, int K, int Kp, int nc, float* amax, void* stream) {
    GN_REQUIRE(x && planes && rows >= 0 && K > 0 && Kp >= K && nc >= 1 && nc <= 3, "null pointer or bad size");
    if (rows == 0) return GN_OK;
    const long long n = rows * Kp, want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < 16LL * num_sms() ? want : 16LL * num_sms());
    cudaStream_t st = as_stream(stream);
    uint16_t* pl = (uint16_t*)planes;
    if (amax != nullptr) {
        int rc = launch_amax(x, rows * K, amax, st);
        if (rc != GN_OK) return rc;
        split_pad_f32_kernel<2, true><<<grid, 256, 0, st>>>(x, pl, rows, K, Kp, amax);
    } else if (nc == 3) split_pad_f32_kernel<3, false><<<grid, 256, 0, st>>>(x, pl, rows, K, Kp, nullptr);
    else if (nc == 2) split_pad_f32_kernel<2, false><<<grid, 256, 0, st>>>(x, pl, rows, K, Kp, nullptr);
    else split_pad_f32_kernel<1, false><<<grid, 256, 0, st>>>(x, pl, rows, K, Kp, nullptr);
    return cuda_status("split_pad_f32_kernel");
}
extern "C" int gn_split_pad_f32_bf16(const float* x, void* planes, long long rows, int K, int Kp, int nc, void* stream) {
    return split_pad_impl(x, planes, rows, K, Kp, nc, nullptr, stream);
}
extern "C" int gn_split_pad_f32_f16x2(const float* x, void* planes, float* amax, long long rows, int K, int Kp, void* stream) {
    GN_REQUIRE(amax, "null pointer");
    return split_pad_impl(x, planes, rows, K, Kp, 2, amax, stream);
}

extern "C" int gn_dense_w_split_bf16(const float* w, void* wk, void* wt, int K, int Kp, int N, int nc, void* stream) {
    GN_REQUIRE(K > 0 && Kp >= K, "bad feature counts");
    return w_split_impl(w, wk, wt, 1, Kp, N, K, nc, nullptr, stream);
}
extern "C" int gn_dense_w_split_f16x2(const float* w, void* wk, void* wt, float* amax, int K, int Kp, int N, void* stream) {
    GN_REQUIRE(amax && K > 0 && Kp >= K, "null pointer or bad feature counts");
    return w_split_impl(w, wk, wt, 1, Kp, N, K, 2, amax, stream);
}

static int dense_fwd_tc3(const void* xs, const void* wts, const float* bias, float* y, void* ys, int M, int Kp, int N, int act,
                         float act_param, Fmt3 f, void* stream) {
    GN_REQUIRE(M > 0 && M <= 65535 * 128, "bad batch size");
    // split-K: a long contraction with few output tiles (burst discriminator Dense(16128 -> 1024) at batch 16-128 has
    // 8 tiles for 148 SMs) is cut into chunks of >= 256 channels that run as separate tiles and meet in the output by
    // vector reductions; it also keeps the truncating tensor-memory accumulator short (<= 4096 channels per chunk)
    int chunks = 1;
    if (y != nullptr && ys == nullptr && Kp >= 1024 && Kp % 64 == 0 && N % 4 == 0) {
        const long long tiles = (long long)((M + 127) / 128) * (N / (N % 128 == 0 ? 128 : 64));
        const int units = Kp / 64;      // chunks of whole 64-channel blocks (the widest stage)
        for (int c = 2; c <= units; ++c) {
            if (units % c != 0) continue;
            const int kc = Kp / c;
            if (kc < 256) break;
            if (tiles * c <= 2LL * num_sms() || kc > 4096) chunks = c;
        }
    }
    if (chunks == 1) return fwd_tc3(xs, wts, bias, y, ys, nullptr, nullptr, 1, M, Kp, M, N, 1, 1, 0, act, act_param, f, 1, stream);
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, st);
    int rc = fwd_tc3(xs, wts, nullptr, y, nullptr, nullptr, nullptr, 1, M, Kp, M, N, 1, 1, 0, GN_ACT_NONE, 0.f, f, chunks, stream);
    if (rc != GN_OK) return rc;
    if (bias != nullptr || act != GN_ACT_NONE) {
        const long long n4 = (long long)M * N / 4;
        const unsigned grid = (unsigned)((n4 + 255) / 256 < 16LL * num_sms() ? (n4 + 255) / 256 : 16LL * num_sms());
        bias_act_kernel<<<grid, 256, 0, st>>>(y, bias, n4, N / 4, act, act_param);
        return cuda_status("bias_act_kernel");
    }
    return GN_OK;
}
extern "C" int gn_dense_fwd_bf16x3(const void* xs, const void* wts, const float* bias, float* y, void* ys, int M, int Kp,
                                   int N, int act, float act_param, int nc, void* stream) {
    return dense_fwd_tc3(xs, wts, bias, y, ys, M, Kp, N, act, act_param, Fmt3{nc, nullptr, nullptr}, stream);
}
extern "C" int gn_dense_fwd_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax,
                                  const float* bias, float* y, int M, int Kp, int N, int act, float act_param, void* stream) {
    GN_REQUIRE(x_amax && w_amax, "null pointer");
    return dense_fwd_tc3(xs, wts, bias, y, nullptr, M, Kp, N, act, act_param, Fmt3{2, x_amax, w_amax}, stream);
}

extern "C" int gn_dense_dgrad_bf16x3(const void* dys, const void* wks, const float* x_in, float* dx, float* dx_colsum, int M,
                                     int K, int N, int in_act, float in_act_param, int nc, void* stream) {
    GN_REQUIRE(K % 64 == 0, "the data gradient needs an unpadded feature count (K % 64 == 0)");
    return dgrad_tc3(dys, wks, x_in, dx, nullptr, dx_colsum, nullptr, 1, M, K, M, N, 1, 1, 0, in_act, in_act_param,
                     Fmt3{nc, nullptr, nullptr}, stream);
}
extern "C" int gn_dense_dgrad_f16x2(const void* dys, const float* dy_amax, const void* wks, const float* w_amax,
                                    const float* x_in, float* dx, float* dx_colsum, int M, int K, int N, int in_act,
                                    float in_act_param, void* stream) {
    GN_REQUIRE(dy_amax && w_amax, "null pointer");
    GN_REQUIRE(K % 64 == 0, "the data gradient needs an unpadded feature count (K % 64 == 0)");
    return dgrad_tc3(dys, wks, x_in, dx, nullptr, dx_colsum, nullptr, 1, M, K, M, N, 1, 1, 0, in_act, in_act_param,
                     Fmt3{2, dy_amax, w_amax}, stream);
}

extern "C" int gn_dense_wgrad_bf16x3(const void* xs, const void* dys, const float* dy, float* dw, float* db, int M, int K,
                                     int N, int Kp, int nc, void* stream) {
    GN_REQUIRE(K > 0 && Kp >= K, "bad feature counts");
    return wgrad_tc3(xs, dys, dy, dw, db, 1, M, Kp, M, N, 1, 1, 0, Fmt3{nc, nullptr, nullptr}, K, stream);
}
extern "C" int gn_dense_wgrad_f16x2(const void* xs, const float* x_amax, const void* dys, const float* dy_amax,
                                    const float* dy, float* dw, float* db, int M, int K, int N, int Kp, void* stream) {
    GN_REQUIRE(x_amax && dy_amax && K > 0 && Kp >= K, "null pointer or bad feature counts");
    return wgrad_tc3(xs, dys, dy, dw, db, 1, M, Kp, M, N, 1, 1, 0, Fmt3{2, x_amax, dy_amax}, K, stream);
}
