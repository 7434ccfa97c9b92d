// Conv1D as implicit GEMM on the 5th-generation tensor cores (tcgen05) -- the throughput path.
//
//   bf16 activations (channels-last) and weights, fp32 accumulation in tensor memory (TMEM).
//   Operand tiles are fetched by TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight out of the NLC
//   activation tensor: one (tap, 64-channel) slab per pipeline stage, zero padding / sample boundaries /
//   ragged tails come from TMA out-of-bounds zero fill, stride-2 convolutions from the TMA traversal stride.
//   Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
//   tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld -> bias / activation / mask -> bf16 global store).
//
//   forward : Y[b,l,co]  = act( sum_{t,ci} X[b, l*s+t-p, ci] * W[t,ci,co] + bias[co] )      A=X   (K-major)  B=Wt[t][co][ci]
//   dgrad   : dX[b,j,ci] = act'(Xin[b,j,ci]) * sum_{t,co} dY[b,(j+p-t)/s,co] * W[t,ci,co]    A=dY  (K-major)  B=W [t][ci][co]
//   wgrad   : dW[t,ci,co] = sum_{b,l} X[b,l*s+t-p,ci] * dY[b,l,co]                           A=X^T, B=dY (both MN-major)
//
// Reference call sites: Conv1D layers of bbhMahoGANy.py:250-292,362-395 (cuDNN fp32 in the reference).

#include "tc_common.cuh"

namespace gn {


constexpr int TC_BK = 64;           // bf16 elements per 128-byte swizzle row
constexpr int TC_THREADS = 416;     // warps: 0 TMA producer, 1 MMA, 2-5 and 6-9 epilogue (two sets that take alternate
                                    // slabs: a lone warp per TMEM lane quarter needs ~900 cycles per slab, which bounds
                                    // every layer whose tile has little MMA work), 10 I/O (bulk stores), 11-12 column sums (alternate slabs)
constexpr int TC_EPI_SETS = 2;
constexpr int TC_WG_THREADS = 192;  // wgrad kernel: warps 0 TMA producer, 1 MMA, 2-5 epilogue

struct TcArgs {
    int B, L, Lout, Cin, Cout, k, s, p;   // convolution geometry (L = input length, Lout = output length)
    int mode;                              // 0 fwd, 1 dgrad
    int act;                               // fwd: output activation; dgrad: derivative mask of `aux` (input activation)
    float act_param;
    const float* bias;                     // fwd
    const __nv_bfloat16* aux;              // dgrad: the conv's own input X (post-activation of the previous layer) or null
    __nv_bfloat16* out;                    // fwd: Y (B,Lout,Cout); dgrad: dX (B,L,Cin)
    float* colsum;                         // dgrad: per-channel sum of dX over (b, l) (= bias gradient of the layer that
                                           // produced x_in), accumulated with atomics; null = not wanted
    int m_tiles;                           // tiles of 128 rows per (sample, parity)
};

constexpr int TC_SLAB_COLS = 32;                          // channels per epilogue slab (64-byte rows, SWIZZLE_64B)
constexpr int TC_SLAB_BYTES = TC_BM * TC_SLAB_COLS * 2;   // 8 KB
constexpr int TC_RING = 4;                                // slabs in the epilogue ring

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RING_BYTES = TC_RING * TC_SLAB_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + RING_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------ fwd / dgrad
// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles tile = blockIdx.x + i*gridDim.x with
// n-tile fastest (CTAs running side by side share the activation tile through L2).  Three pipelines:
//   shared-memory ring   (full/empty mbarriers, STAGES deep)   TMA producer -> MMA issuer
//   TMEM accumulator ring (2 x BN fp32 columns)                 MMA issuer   -> epilogue warps
//   epilogue slab ring   (4 x [128 rows x 32 channels] bf16)    I/O thread  <-> epilogue warps
// The epilogue of tile i (TMEM -> registers -> bias/act/mask -> bf16 -> slab -> TMA store) overlaps the MMAs of
// tile i+1.  For dgrad the epilogue threads read the mask source (the conv's own input, same shape as the
// output: 64 contiguous bytes per thread and slab) straight from global memory, prefetched into L2 one tile
// ahead and into registers one slab ahead, so its latency never sits on the epilogue's critical path.


// walks the (tile, slab) sequence of one CTA
template <int BN>
struct SlabIter {
    TileWalker<BN> w;
    int sl;
    TileCoord c;
    __device__ __forceinline__ void init(int first, int step, int n_nt, int m_tiles, int npar, int B) {
        w.init(first, step, n_nt, m_tiles, npar, B);
        sl = 0;
        c = w.coord();
    }
    __device__ __forceinline__ bool valid() const { return w.valid(); }
    __device__ __forceinline__ void next() {
        if (++sl == BN / TC_SLAB_COLS) {
            sl = 0;
            w.next();
            c = w.coord();
        }
    }
};

template <int BN, int STAGES, bool AUX>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapO0, const __grid_constant__ CUtensorMap mapO1, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    using S = TcSmem<BN, STAGES>;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* ring = tiles + STAGES * S::STAGE_BYTES;            // 1024-byte aligned (stage sizes are multiples of 1 KB)
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + S::RING_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* slab_ready = tmem_empty + 2;         // [TC_RING]  I/O thread -> epilogue: slab free (and mask landed)
    uint64_t* slab_done = slab_ready + TC_RING;    // [TC_RING]  epilogue -> I/O thread: slab holds the bf16 result
    uint64_t* slab_summed = slab_done + TC_RING;   // [TC_RING]  column-sum warp -> I/O thread: slab has been read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slab_summed + TC_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npar = (a.mode == 1) ? a.s : 1;
    const int kdim = (a.mode == 0) ? a.Cin : a.Cout;   // contraction channels
    const int nkb = kdim / TC_BK;
    const int n_nt = ((a.mode == 0) ? a.Cout : a.Cin) / BN;
    const int total_tiles = a.B * npar * a.m_tiles * n_nt;
    const int sshift = (a.s == 2) ? 1 : 0;             // the stride is 1 or 2 (check_tc_geom)
    (void)total_tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * TC_EPI_SETS);      // one arrival per epilogue warp
        }
        for (int i = 0; i < TC_RING; ++i) {
            mbar_init(&slab_ready[i], 1);
            mbar_init(&slab_done[i], 4);
            mbar_init(&slab_summed[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // TMA producer: the whole warp walks the schedule (all values warp-uniform), one elected lane issues
        uint32_t st = 0, ph = 0;
        TileWalker<BN> tw;
        for (tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, a.B); tw.valid(); tw.next()) {
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
                        uint8_t* sB = sA + S::A_BYTES;
                        mbar_expect_tx(&full[st], S::STAGE_BYTES);
                        tma_load_3d(sA, &mapA, &full[st], kb * TC_BK, rowc, c.b);
                        tma_load_3d(sB, &mapB, &full[st], kb * TC_BK, c.n0, tap);
                    }
                    __syncwarp();
                    if (++st == STAGES) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: same structure; the elected lane issues the four K=16 instructions of a stage and the commits
        constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
        // descriptor words: lo = (addr >> 4) | LBO(16 B) << 16, hi = SBO(1024 B) | version 1 | SWIZZLE_128B; only the
        // address field changes, by (stage bytes >> 4) per stage and 2 per K=16 step (all below 2^14, no carry)
        constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lo_a0 = ((base >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t lo_b0 = (((base + S::A_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
        uint32_t st = 0, ph = 0;
        int ti_local = 0;
        TileWalker<BN> tw;
        for (tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, a.B); tw.valid(); tw.next(), ++ti_local) {
            int ntaps = a.k;
            if (a.mode == 1) {
                const int tap_first = (tw.par + a.p) & (a.s - 1);
                ntaps = (a.k - tap_first + a.s - 1) >> sshift;
            }
            const int niter = ntaps * nkb;
            const int acc = ti_local & 1, acc_ph = (ti_local >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_ph ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * BN);
            for (int it = 0; it < niter; ++it) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t la = lo_a0 + st * (uint32_t)(S::STAGE_BYTES >> 4);
                    const uint32_t lb = lo_b0 + st * (uint32_t)(S::STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(la + 2u * k);
                        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(lb + 2u * k);
                        tc_mma_bf16(tacc, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[st]);      // frees the smem stage once the MMAs have read it
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
            if (elect_one()) tc_commit(&tmem_full[acc]);     // accumulator complete
            __syncwarp();
        }
    } else if (warp == 10) {
        // I/O thread: owns every bulk store (bulk groups are per thread) and, for dgrad, the mask loads.
        // Slab step s uses ring buffer s % 4.  After committing store s, store s-1 has been read out of shared
        // memory once at most one group is pending, which frees buffer (s-1) % 4 == (s+3) % 4 for slab s+3.
        if (lane == 0) {
            SlabIter<BN> ld, st;
            ld.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, a.B);
            st = ld;
            auto prepare = [&](int s_idx) {
                mbar_arrive(&slab_ready[s_idx & (TC_RING - 1)]);
                ld.next();
            };
            for (int i = 0; i < TC_RING - 1 && ld.valid(); ++i) prepare(i);
            for (int s = 0; st.valid(); ++s, st.next()) {
                const int b = s & (TC_RING - 1);
                mbar_wait(&slab_done[b], (s / TC_RING) & 1);
                tma_store_3d((st.c.par == 0) ? &mapO0 : &mapO1, ring + b * TC_SLAB_BYTES, st.c.n0 + st.sl * TC_SLAB_COLS,
                             st.c.m0, st.c.b);
                tma_store_commit();
                if (ld.valid()) {
                    tma_store_wait_read<1>();
                    // the buffer about to be handed out again held slab s-1: the column-sum warp must be done with it
                    if (a.colsum != nullptr && s >= 1) mbar_wait(&slab_summed[(s - 1) & (TC_RING - 1)], ((s - 1) / TC_RING) & 1);
                    prepare(s + TC_RING - 1);
                }
            }
            tma_store_wait_read<0>();
        }
    } else if (warp >= 11) {
        // column sums of the finished bf16 slabs (dgrad only): the per-channel sum of dX over (b, l) is the bias
        // gradient of the layer that produced this conv's input, so that layer needs no separate pass over dX.
        // lane = (row parity, channel pair): 64 conflict-free 4-byte loads per slab, rows past the end excluded.
        if (a.colsum != nullptr) {
            SlabIter<BN> it;
            it.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, a.B);
            const int pr = lane & 15, half = lane >> 4;
            const int cset = warp - 11;                        // this warp takes slabs with s % 2 == cset
            for (int s = 0; it.valid(); ++s, it.next()) {
                if ((s & 1) != cset) continue;
                const int b = s & (TC_RING - 1);
                const int rows_class = (a.L - it.c.par + npar - 1) >> ((npar == 2) ? 1 : 0);   // rows of this parity class
                const int nrows = min(TC_BM, rows_class - it.c.m0);
                mbar_wait(&slab_done[b], (s / TC_RING) & 1);
                const uint8_t* slab = ring + b * TC_SLAB_BYTES;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                for (int r = half; r < nrows; r += 2) {
                    const uint32_t off = (uint32_t)r * 64u + ((((uint32_t)pr >> 2) ^ (((uint32_t)r >> 1) & 3u)) << 4) +
                                         (((uint32_t)pr & 3u) << 2);
                    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(slab + off));
                    s0 += v.x;
                    s1 += v.y;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&slab_summed[b]);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                if (half == 0) {
                    float* dst = a.colsum + it.c.n0 + it.sl * TC_SLAB_COLS + 2 * pr;
                    atomicAdd(dst, s0);
                    atomicAdd(dst + 1, s1);
                }
            }
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (rows of the tile).  Each 32-column slab goes
        // TMEM -> registers -> bias / activation / mask -> bf16 -> swizzled slab -> (I/O thread) TMA store;
        // rows past the end of the sample are clipped by the tensor map.
        const int q = warp & 3;
        const int eset = (warp - 2) >> 2;                       // epilogue set: slabs eset, eset + 2, ... of every tile
        const int row = q * 32 + lane;
        const uint32_t sw = (uint32_t)(row >> 1) & 3u;          // SWIZZLE_64B: 16-byte chunk index XOR address bits 7-8
        constexpr int NS = BN / TC_SLAB_COLS;
        // dgrad mask source: row (m0 + row) of parity class par of sample b is position j = (m0 + row) * s + par
        // of the conv's own input; this thread needs BN consecutive channels of it (n0 ..), 64 bytes per slab
        auto mask_row = [&](const TileCoord& c) -> const uint8_t* {
            const int j = (c.m0 + row) * npar + c.par;
            if (!AUX || j >= a.L) return nullptr;
            return reinterpret_cast<const uint8_t*>(a.aux + ((size_t)c.b * a.L + j) * a.Cin + c.n0);
        };
        auto mask_prefetch_l2 = [&](const uint8_t* p) {
            if (p != nullptr) {
#pragma unroll
                for (int i = 0; i < BN * 2 / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
            }
        };
        auto mask_load = [&](uint4 (&m)[4], const uint8_t* p, int sl) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                m[i] = (p != nullptr) ? __ldg(reinterpret_cast<const uint4*>(p + sl * 64) + i) : make_uint4(0, 0, 0, 0);
        };
        int ti_local = 0;
        TileWalker<BN> tw, tw_next;
        tw.init(blockIdx.x, gridDim.x, n_nt, a.m_tiles, npar, a.B);
        tw_next = tw;
        tw_next.next();
        if (AUX && tw.valid()) mask_prefetch_l2(mask_row(tw.coord()));
        for (; tw.valid(); tw.next(), tw_next.next(), ++ti_local) {
            const TileCoord c = tw.coord();
            const uint8_t* mrow = mask_row(c);
            uint4 mk[4];
            if (AUX) {
                mask_load(mk, mrow, eset);
                // pull the next tile's mask rows into L2 while this tile is being processed
                if (tw_next.valid()) mask_prefetch_l2(mask_row(tw_next.coord()));
            }
            const int acc = ti_local & 1, acc_ph = (ti_local >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int sl = eset; sl < NS; sl += TC_EPI_SETS) {
                const int s = ti_local * NS + sl;          // position of this slab in the CTA's slab sequence
                uint4 mk_next[4];
                if (AUX && sl + TC_EPI_SETS < NS) mask_load(mk_next, mrow, sl + TC_EPI_SETS);
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)(sl * TC_SLAB_COLS), v);
                if (sl + TC_EPI_SETS >= NS) {
                    // all of this warp's TMEM reads are complete: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                const int b = s & (TC_RING - 1);
                uint8_t* srow = ring + b * TC_SLAB_BYTES + row * 64;
                if (a.mode == 0) {
                    if (a.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(a.bias + c.n0 + sl * TC_SLAB_COLS);
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
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&mk[g8]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 y = __bfloat1622float2(h[e]);
                                f[g8 * 8 + 2 * e] *= act_bwd_t<decltype(tag)::kind>(y.x, a.act_param);
                                f[g8 * 8 + 2 * e + 1] *= act_bwd_t<decltype(tag)::kind>(y.y, a.act_param);
                            }
                        }
                    });
#pragma unroll
                    for (int i = 0; i < 4; ++i) mk[i] = mk_next[i];
                }
                mbar_wait(&slab_ready[b], (s / TC_RING) & 1);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[g8 * 8 + 0], f[g8 * 8 + 1]);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(f[g8 * 8 + 2], f[g8 * 8 + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[g8 * 8 + 4], f[g8 * 8 + 5]);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(f[g8 * 8 + 6], f[g8 * 8 + 7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2);
                    pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(srow + (((uint32_t)g8 ^ sw) << 4)) = pk;
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) mbar_arrive(&slab_done[b]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[t, ci, co] (fp32, atomically accumulated; caller zeroes) = sum over (b, l) of X[b, l*s+t-p, ci] * dY[b, l, co]
// Persistent split-K: a work unit is (K chunk, output tile) with the output tile = (tap, 128 x BN block of
// (ci, co)) varying fastest, so the CTAs running at any moment cover every output tile of the same one or two
// K chunks and the activations / gradients of that chunk are read from HBM once and shared through L2.
// The number of K chunks is chosen on the host so that the unit count fills whole waves of the grid.
// TMEM is double buffered: the fp32 red.add epilogue of unit i overlaps the MMAs of unit i+1.
// SWAP (Cin == 64): the 128-row operand is dY (co) and the BN(=64)-column operand is X (ci).
struct TcWgradArgs {
    int B, L, Lout, Cin, Cout, k, s, p;
    int lblocks;        // ceil(Lout / 64)
    int iters_total;    // B * lblocks
    int iters_per_chunk;
    int n_chunks;
    int n_tiles_n;      // number of N tiles
    int out_tiles;      // k * m_tiles * n_tiles_n
    float* dw;          // (k, Cin, Cout) fp32
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgUnit {
    int tap, m0, n0, it0, niter;
};
template <int BN>
__device__ __forceinline__ WgUnit decode_wg_unit(int u, const TcWgradArgs& a) {
    WgUnit w;
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

template <int BN, int STAGES, bool SWAP>
__global__ void __launch_bounds__(TC_WG_THREADS)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                     TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = TC_BM * 128;      // 128 (MN) x 64 (K) bf16 = two 64x64 MN-major blocks of 8 KB
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
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
    if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // producer: the whole warp walks the schedule, one elected lane issues (operands stay in uniform registers)
        uint32_t st = 0, ph = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WgUnit w = decode_wg_unit<BN>(u, a);
            int bb = w.it0 / a.lblocks, lb = w.it0 - bb * a.lblocks;
            for (int i = 0; i < w.niter; ++i) {
                mbar_wait(&empty[st], ph ^ 1);
                if (elect_one()) {
                    const int l0 = lb * 64;
                    uint8_t* sA = tiles + st * STAGE_BYTES;
                    uint8_t* sB = sA + A_BYTES;
                    mbar_expect_tx(&full[st], STAGE_BYTES);
                    const int xrow = l0 * a.s + w.tap - a.p;     // X position of output position l0 for this tap
                    // 128-row operand: two 64-channel blocks; BN-column operand: BN/64 blocks
                    if (!SWAP) {
                        tma_load_3d(sA, &mapX, &full[st], w.m0, xrow, bb);
                        tma_load_3d(sA + 8192, &mapX, &full[st], w.m0 + 64, xrow, bb);
#pragma unroll
                        for (int h = 0; h < BN / 64; ++h)
                            tma_load_3d(sB + h * 8192, &mapDY, &full[st], w.n0 + h * 64, l0, bb);
                    } else {
                        tma_load_3d(sA, &mapDY, &full[st], w.m0, l0, bb);
                        tma_load_3d(sA + 8192, &mapDY, &full[st], w.m0 + 64, l0, bb);
#pragma unroll
                        for (int h = 0; h < BN / 64; ++h)
                            tma_load_3d(sB + h * 8192, &mapX, &full[st], w.n0 + h * 64, xrow, bb);
                    }
                }
                __syncwarp();
                if (++lb == a.lblocks) { lb = 0; ++bb; }
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(TC_BM, BN, 1, 1);     // both operands MN-major
        // MN-major SW128 descriptors: 64-element MN blocks LBO = 8192 B apart, 8-row K groups SBO = 1024 B apart, 16 K
        // rows per instruction = 2048 B; only the 14-bit address field changes between instructions
        constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lo_a0 = ((base >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);
        const uint32_t lo_b0 = (((base + A_BYTES) >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);
        uint32_t st = 0, ph = 0;
        int ul = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ul) {
            const WgUnit w = decode_wg_unit<BN>(u, a);
            const int acc = ul & 1, acc_ph = (ul >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * BN);
            for (int i = 0; i < w.niter; ++i) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t la = lo_a0 + st * (uint32_t)(STAGE_BYTES >> 4);
                    const uint32_t lb = lo_b0 + st * (uint32_t)(STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 64 / 16; ++k) {
                        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(la + (2048u >> 4) * k);
                        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(lb + (2048u >> 4) * k);
                        tc_mma_bf16(tacc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
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
        int ul = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ul) {
            const WgUnit w = decode_wg_unit<BN>(u, a);
            const int acc = ul & 1, acc_ph = (ul >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)c0, v);
                if (c0 + 32 >= BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (!SWAP) {
                    // thread = input channel, 32 consecutive output channels: eight 16-byte vector reductions
                    float* dst = a.dw + ((size_t)w.tap * a.Cin + (w.m0 + row)) * a.Cout + w.n0 + c0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        red_add_v4(dst + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                   __uint_as_float(v[i + 3]));
                } else {
                    // thread = output channel (consecutive across the warp), columns = input channels
                    float* dst = a.dw + ((size_t)w.tap * a.Cin + (w.n0 + c0)) * a.Cout + w.m0 + row;
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(dst + (size_t)i * a.Cout, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ helpers
// weights: f32 (k,Cin,Cout) -> bf16 same layout (dgrad B operand: K = co contiguous) and bf16 transposed
// (k,Cout,Cin) (forward B operand: K = ci contiguous)
// One 32 x 32 (ci, co) tile of one tap per block, transposed through shared memory: both outputs are written with
// consecutive lanes on consecutive addresses (the direct form scattered 2-byte stores Cin elements apart).
__global__ void __launch_bounds__(256) conv_w_cast_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wk,
                                                          __nv_bfloat16* __restrict__ wt, int k, int Cin, int Cout) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int t = blockIdx.z;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) {
            const size_t i = ((size_t)t * Cin + ci) * Cout + co;
            const __nv_bfloat16 h = __float2bfloat16_rn(w[i]);
            wk[i] = h;
            tile[r][tx] = h;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        if (ci < Cin && co < Cout) wt[((size_t)t * Cout + co) * Cin + ci] = tile[tx][r];
    }
}
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                            long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __bfloat162float(x[i]);
}
// per-channel sum of a (rows, C) bf16 matrix -> f32 (bias gradient); out must be zeroed by the caller.
// thread = 8 consecutive channels (one 128-bit load per row), RL row lanes per block, shared-memory combine.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C,
                                                          long long rows_per_block, float* __restrict__ out) {
    __shared__ float sm[256][9];
    const int groups = C / 8;                       // C % 8 == 0
    const int g = threadIdx.x % groups;
    const int rl = threadIdx.x / groups, nrl = blockDim.x / groups;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (rl < nrl) {
        for (long long r = r0 + rl; r < r1; r += nrl) {
            uint4 pk = __ldg(reinterpret_cast<const uint4*>(x + r * C) + g);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 v = __bfloat1622float2(h[e]);
                acc[2 * e] += v.x;
                acc[2 * e + 1] += v.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < groups) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        for (int l = 0; l < nrl; ++l)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] += sm[l * groups + threadIdx.x][j];
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&out[threadIdx.x * 8 + j], t[j]);
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------------

EncodeTiledFn tc_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D bf16 tensor map: dims (d0 contiguous, d1, d2), strides in elements for d1, d2; box (b0, b1, 1); traversal
// stride es1 along d1 (box extent b1*es1 in global coordinates -> b1 rows in shared memory)
static int make_map3(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t st1, uint64_t st2,
                     uint32_t b0, uint32_t b1, uint32_t es1, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = tc_encode_fn();
    if (fn == nullptr) return fail(GN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available%s", "");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {st1 * 2, st2 * 2};
    cuuint32_t box[3] = {b0, b1 * es1, 1};
    cuuint32_t estr[3] = {1, es1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%s code %lld)", "", (long long)r);
    return GN_OK;
}

template <int BN, int STAGES, bool AUX>
static int launch_conv_tc(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mO0, const CUtensorMap& mO1,
                          const TcArgs& a, long long total_tiles, cudaStream_t st) {
    auto kfn = conv_tc_kernel<BN, STAGES, AUX>;
    constexpr int smem = TcSmem<BN, STAGES>::TOTAL;
    static_assert(smem <= 232448, "shared memory budget exceeded");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    const int grid = (int)(total_tiles < (long long)num_sms() ? total_tiles : (long long)num_sms());
    kfn<<<grid, TC_THREADS, smem, st>>>(mA, mB, mO0, mO1, a);
    return cuda_status("conv_tc_kernel");
}

static int dispatch_conv_tc(int BN, bool aux, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mO0,
                            const CUtensorMap& mO1, const TcArgs& a, long long tiles, cudaStream_t st) {
    if (!aux) {
        if (BN == 256) return launch_conv_tc<256, 4, false>(mA, mB, mO0, mO1, a, tiles, st);
        if (BN == 128) return launch_conv_tc<128, 6, false>(mA, mB, mO0, mO1, a, tiles, st);
        return launch_conv_tc<64, 8, false>(mA, mB, mO0, mO1, a, tiles, st);
    }
    if (BN == 256) return launch_conv_tc<256, 4, true>(mA, mB, mO0, mO1, a, tiles, st);
    if (BN == 128) return launch_conv_tc<128, 6, true>(mA, mB, mO0, mO1, a, tiles, st);
    return launch_conv_tc<64, 8, true>(mA, mB, mO0, mO1, a, tiles, st);
}
static int pick_bn(int C) { return (C % 256 == 0) ? 256 : ((C % 128 == 0) ? 128 : 64); }

template <int BN, int STAGES, bool SWAP>
static int launch_wgrad_tc(const CUtensorMap& mX, const CUtensorMap& mDY, const TcWgradArgs& a, int grid,
                           cudaStream_t st) {
    auto kfn = conv_tc_wgrad_kernel<BN, STAGES, SWAP>;
    constexpr int smem = STAGES * (TC_BM * 128 + BN * 128) + 1024 + 256;
    static_assert(smem <= 232448, "shared memory budget exceeded");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    kfn<<<grid, TC_WG_THREADS, smem, st>>>(mX, mDY, a);
    return cuda_status("conv_tc_wgrad_kernel");
}

// number of K chunks of the persistent split-K wgrad: minimise  waves * (iterations per chunk + fixed cost per unit)
static int pick_wgrad_chunks(int out_tiles, int iters_total, int grid) {
    const int ovh = 6;                                   // pipeline refill + epilogue hand-over, in K iterations
    int lo = (iters_total + 511) / 512, hi = (iters_total + 23) / 24;
    if (lo < 1) lo = 1;
    if (hi < lo) hi = lo;
    if (hi > 4096) hi = 4096;
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

static int check_tc_geom(int B, int L, int Cin, int Lout, int Cout, int k, int s, int p) {
    GN_REQUIRE(B > 0 && L > 0 && Lout > 0 && k > 0 && k <= 16 && (s == 1 || s == 2) && p >= 0 && p < k, "bad geometry");
    GN_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tensor-core path needs Cin and Cout to be multiples of 64");
    GN_REQUIRE(B <= 65535, "batch too large for one launch");
    return GN_OK;
}

}  // namespace gn

using namespace gn;

extern "C" int gn_conv_w_to_bf16(const float* w, void* wk, void* wt, int k, int Cin, int Cout, void* stream) {
    GN_REQUIRE(w && wk && wt && k > 0 && Cin > 0 && Cout > 0, "null pointer or bad size");
    GN_REQUIRE(k <= 65535 && (Cin + 31) / 32 <= 65535, "weight tensor too large for the cast grid");
    dim3 grid((unsigned)((Cout + 31) / 32), (unsigned)((Cin + 31) / 32), (unsigned)k);
    conv_w_cast_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wk, (__nv_bfloat16*)wt, k, Cin, Cout);
    return cuda_status("conv_w_cast_kernel");
}

extern "C" int gn_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
    GN_REQUIRE(x && y && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    unsigned grid = (unsigned)((n + 255) / 256 < 16LL * num_sms() ? (n + 255) / 256 : 16LL * num_sms());
    cast_f32_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)y, n);
    return cuda_status("cast_f32_bf16_kernel");
}

extern "C" int gn_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream) {
    GN_REQUIRE(x && y && n >= 0, "null pointer or n < 0");
    if (n == 0) return GN_OK;
    unsigned grid = (unsigned)((n + 255) / 256 < 16LL * num_sms() ? (n + 255) / 256 : 16LL * num_sms());
    cast_bf16_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, y, n);
    return cuda_status("cast_bf16_f32_kernel");
}

extern "C" int gn_conv1d_fwd_bf16(const void* x, const void* wt, const float* bias, void* y, int B, int L, int Cin,
                                  int Lout, int Cout, int k, int stride, int pad_left, int act, float act_param,
                                  void* stream) {
    GN_REQUIRE(x && wt && y, "null pointer");
    int rc = check_tc_geom(B, L, Cin, Lout, Cout, k, stride, pad_left);
    if (rc != GN_OK) return rc;
    CUtensorMap mA, mB;
    const int BN = pick_bn(Cout);
    // A: X viewed as (Cin, L, B); 128 output rows per tile, traversal stride = conv stride
    rc = make_map3(&mA, x, Cin, L, B, Cin, (uint64_t)L * Cin, TC_BK, TC_BM, stride);
    if (rc != GN_OK) return rc;
    // B: Wt (k, Cout, Cin) viewed as (Cin, Cout, k)
    rc = make_map3(&mB, wt, Cin, Cout, k, Cin, (uint64_t)Cout * Cin, TC_BK, BN, 1);
    if (rc != GN_OK) return rc;
    TcArgs a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.mode = 0; a.act = act; a.act_param = act_param; a.bias = bias; a.out = (__nv_bfloat16*)y;
    a.m_tiles = (Lout + TC_BM - 1) / TC_BM;
    // output: Y viewed as (Cout, Lout, B), 32-channel x 128-row store boxes (SWIZZLE_64B slabs)
    CUtensorMap mO;
    rc = make_map3(&mO, y, Cout, Lout, B, Cout, (uint64_t)Lout * Cout, TC_SLAB_COLS, TC_BM, 1, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != GN_OK) return rc;
    return dispatch_conv_tc(BN, false, mA, mB, mO, mO, a, (long long)B * a.m_tiles * (Cout / BN), as_stream(stream));
}

extern "C" int gn_conv1d_dgrad_bf16(const void* dy, const void* wk, const void* x_in, void* dx, float* dx_colsum, int B,
                                    int L, int Cin, int Lout, int Cout, int k, int stride, int pad_left, int in_act,
                                    float in_act_param, void* stream) {
    GN_REQUIRE(dy && wk && dx, "null pointer");
    int rc = check_tc_geom(B, L, Cin, Lout, Cout, k, stride, pad_left);
    if (rc != GN_OK) return rc;
    CUtensorMap mA, mB;
    const int BN = pick_bn(Cin);
    // A: dY viewed as (Cout, Lout, B), 128 rows, unit traversal stride (parity classes handle the conv stride)
    rc = make_map3(&mA, dy, Cout, Lout, B, Cout, (uint64_t)Lout * Cout, TC_BK, TC_BM, 1);
    if (rc != GN_OK) return rc;
    // B: W (k, Cin, Cout) viewed as (Cout, Cin, k)
    rc = make_map3(&mB, wk, Cout, Cin, k, Cout, (uint64_t)Cin * Cout, TC_BK, BN, 1);
    if (rc != GN_OK) return rc;
    TcArgs a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.mode = 1; a.act = in_act; a.act_param = in_act_param; a.aux = (const __nv_bfloat16*)x_in;
    a.out = (__nv_bfloat16*)dx;
    a.colsum = dx_colsum;
    if (dx_colsum != nullptr) cudaMemsetAsync(dx_colsum, 0, sizeof(float) * (size_t)Cin, as_stream(stream));
    const int rows = (L + stride - 1) / stride;      // rows of the largest parity class
    a.m_tiles = (rows + TC_BM - 1) / TC_BM;
    // output: one map per parity class r: rows j = i*stride + r of dX, viewed as (Cin, rows_r, B)
    CUtensorMap mO[2];
    const bool aux = (x_in != nullptr && in_act != GN_ACT_NONE);
    for (int r = 0; r < 2; ++r) {
        const int rr = (r < stride) ? r : 0;
        const int rows_r = (L - rr + stride - 1) / stride;
        rc = make_map3(&mO[r], (const __nv_bfloat16*)dx + (size_t)rr * Cin, Cin, rows_r, B, (uint64_t)stride * Cin,
                       (uint64_t)L * Cin, TC_SLAB_COLS, TC_BM, 1, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc != GN_OK) return rc;
    }
    return dispatch_conv_tc(BN, aux, mA, mB, mO[0], mO[1], a,
                            (long long)B * stride * a.m_tiles * (Cin / BN), as_stream(stream));
}

extern "C" int gn_conv1d_wgrad_bf16(const void* x, const void* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                                    int Cout, int k, int stride, int pad_left, void* stream) {
    GN_REQUIRE(x && dy && dw, "null pointer");
    int rc = check_tc_geom(B, L, Cin, Lout, Cout, k, stride, pad_left);
    if (rc != GN_OK) return rc;
    GN_REQUIRE(Cin % 128 == 0 || (Cin == 64 && Cout % 128 == 0), "wgrad needs Cin % 128 == 0, or Cin == 64 with Cout % 128 == 0");
    cudaStream_t st = as_stream(stream);
    CUtensorMap mX, mDY;
    // X (Cin, L, B): 64 channels x 64 positions (traversal stride = conv stride); dY (Cout, Lout, B): 64 x 64
    rc = make_map3(&mX, x, Cin, L, B, Cin, (uint64_t)L * Cin, 64, 64, stride);
    if (rc != GN_OK) return rc;
    rc = make_map3(&mDY, dy, Cout, Lout, B, Cout, (uint64_t)Lout * Cout, 64, 64, 1);
    if (rc != GN_OK) return rc;
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * Cin * Cout, st);
    TcWgradArgs a{};
    a.B = B; a.L = L; a.Lout = Lout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.s = stride; a.p = pad_left;
    a.lblocks = (Lout + 63) / 64;
    a.iters_total = B * a.lblocks;
    a.dw = dw;
    const bool swap = (Cin == 64);
    int m_tiles, BN;
    if (!swap) { m_tiles = Cin / 128; BN = pick_bn(Cout); a.n_tiles_n = Cout / BN; }
    else { m_tiles = Cout / 128; BN = 64; a.n_tiles_n = 1; }
    a.out_tiles = k * m_tiles * a.n_tiles_n;
    const int nsm = num_sms();
    int nk = pick_wgrad_chunks(a.out_tiles, a.iters_total, nsm);
    a.iters_per_chunk = (a.iters_total + nk - 1) / nk;
    a.n_chunks = (a.iters_total + a.iters_per_chunk - 1) / a.iters_per_chunk;
    const long long units = (long long)a.out_tiles * a.n_chunks;
    const int grid = (int)(units < nsm ? units : nsm);
    if (swap) rc = launch_wgrad_tc<64, 8, true>(mX, mDY, a, grid, st);
    else if (BN == 256) rc = launch_wgrad_tc<256, 4, false>(mX, mDY, a, grid, st);
    else if (BN == 128) rc = launch_wgrad_tc<128, 6, false>(mX, mDY, a, grid, st);
    else rc = launch_wgrad_tc<64, 8, false>(mX, mDY, a, grid, st);
    if (rc != GN_OK) return rc;
    if (db != nullptr) {
        cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, st);
        const long long rows = (long long)B * Lout;
        // Cout is a multiple of 64 and <= 2048 here: Cout/8 column groups per block, 256/(Cout/8) row lanes
        GN_REQUIRE(Cout / 8 <= 256, "Cout too large for the bias-gradient kernel");
        long long blocks = 4LL * num_sms();
        long long per = (rows + blocks - 1) / blocks;
        if (per < 32) per = 32;
        blocks = (rows + per - 1) / per;
        colsum_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, rows, Cout, per, db);
        return cuda_status("colsum_bf16_kernel");
    }
    return GN_OK;
}
