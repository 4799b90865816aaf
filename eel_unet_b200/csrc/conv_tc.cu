// bf16 implicit-GEMM 3x3 convolution (fprop + data gradient) for sm_100a, second generation.
//
// Measured on B200 (tools/umma_probe.cu): a tcgen05 K-major SWIZZLE_128B operand is addressed through the ABSOLUTE
// shared-memory address bits, so the 8-row groups of an operand view need not sit on 1024-byte boundaries.  That
// lets ONE dense halo block per 64-channel K chunk serve all nine taps:
//
//     smem A stage = [TH+2 image rows][10 columns][64 ch] = (TH+2) * 1280 B   (one 4-D TMA box, OOB = zero padding)
//     tap (dy, dx), accumulator j:  start = stage + ((16 j + 1 + dy) * 10 + 1 + dx) * 128,  SBO = 1280 B
//
// i.e. activations cross L2 -> SM 1.3-1.4x instead of 3.4x (three column-shifted copies in generation one).
// An output tile is TH x 8 pixels, TH = 16 * MT: MT accumulators of 128 rows share every weight tile, which
// divides the weight traffic per pixel by MT.  When the whole [9][Cin][BN] weight block fits (<= 144 KB:
// 64->64, 128->64, 64->128 -- the full-resolution layers that hold most of the FLOPs) it is loaded ONCE per CTA and
// stays resident (RES), so the steady state only streams activations.
//
//   warp 0  TMA producer   warp 1  MMA issuer (one thread)   warps 2-5 / 2-9  epilogue (TMEM -> +bias/ReLU -> bf16 -> smem transposition -> global)
// Two TMEM accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1.  Persistent, one CTA per SM.
#include "tc_common.cuh"

namespace eel {
namespace tc {

struct ConvParams {
    int kchunks;            // 64-wide input-channel chunks
    int m_tiles, n_tiles;
    int N, H, W;
    int tiles_h, tiles_w;
    int Ntot;               // output channels
    int flip;               // mirror the tap offsets (data gradient)
    int relu;
    int na;                 // A stages in shared memory
    const float* bias;
    bf16* out;
    float* bn_sums;         // optional [2][Ntot]: per-channel sum / sum of squares of the stored output (grid % n_tiles == 0)
    // STATS == 2 (data gradient in front of a BatchNorm + ReLU): bn_sums receives the BatchNorm-BACKWARD sums
    // {sum g, sum g * xhat}, g = stored dy * ReLU mask, from the BatchNorm's input bz (layout of `out`) and bcst[Ntot] = {scale, -shift};
    // bn_sums[1] then holds sum g * z (bn_sums_fix_kernel turns it into sum g * xhat)
    const bf16* bz;
    const float2* bcst;
    int brelu;
    // split > 0 (data gradient of the conv that reads a skip bridge): output columns [0, split) go to `out`, columns
    // [split, Ntot) to `out2`, both [N,H,W,split] tensors -- the weight operand's rows were de-interleaved, so the two halves
    // are the gradients of (upconv + edge feature) and of the encoder skip and no interleaved gradient tensor is ever written.
    // Fused BatchNorm-backward sums (STATS == 2) then cover the first half only (bz / bcst / bn_sums are [.., split]).
    bf16* out2;
    int split;
    // kc1 < kchunks (forward of the conv that reads a skip bridge): the first kc1 K chunks come from the first input tensor
    // (tensor map tmA), the rest from a second one (tmA2) -- the Cin axis of the weight operand was de-interleaved, the
    // interleaved 2C-channel activation tensor is never built
    int kc1;
};

constexpr int kMaxStages = 8;

template <int BN, int MT, bool RES> struct ConvCfg {
    static constexpr int TH = 16 * MT;
    static constexpr int A_TX = (TH + 2) * 10 * 128;
    static constexpr int A_BYTES = (A_TX + 1023) & ~1023;
    static constexpr int B_TILE = BN * 128;                                   // one (chunk, tap) weight tile
    static constexpr int NB = RES ? 0 : (BN == 256 ? 3 : (BN == 128 ? 5 : 8));
    static constexpr int TMEM_COLS = 2 * MT * BN;
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "accumulators must fill a power of two of TMEM");
};

constexpr int kSmemLimit = 232448 - 1024;   // 227 KB per CTA minus alignment slack

template <int BN, int MT, bool RES>
__device__ __forceinline__ void conv_mma_loop(const ConvParams& p, uint8_t* sA, uint8_t* sB, uint64_t* fullA, uint64_t* emptyA,
                                              uint64_t* fullB, uint64_t* emptyB, uint64_t* tmemFull, uint64_t* tmemEmpty,
                                              const uint32_t tmem_base, const int total_tiles) {
    typedef ConvCfg<BN, MT, RES> Cfg;
    {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        // The issuing thread must sustain one MMA per 32-128 cycles: descriptors are built ONCE (stage 0) and every
        // operand view is that descriptor plus a 16-byte-unit offset in its start-address field (no carry: < 256 KB).
        const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), 16, 1280, false);
        const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 16, 1024, false);
        uint32_t tap_off[9];                       // (view start - stage start) / 16 per tap
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            int dy = t / 3 - 1, dx = t % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
            tap_off[t] = (uint32_t)(((1 + dy) * 10 + 1 + dx) * 128) >> 4;
        }
        if (RES) {
            mbar_wait(&fullB[0], 0);
            tc_fence_after();
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            mbar_wait(&tmemEmpty[acc], acc_par ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (MT * BN);
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&fullA[sa], pa);
                tc_fence_after();
                const uint64_t a_stage = a_desc0 + (uint32_t)((sa * Cfg::A_BYTES) >> 4);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    uint64_t b_tile;
                    if (RES) b_tile = b_desc0 + (uint32_t)(((c * 9 + t) * Cfg::B_TILE) >> 4);
                    else {
                        mbar_wait(&fullB[sb], pb);
                        tc_fence_after();
                        b_tile = b_desc0 + (uint32_t)((sb * Cfg::B_TILE) >> 4);
                    }
                    const uint64_t a_tap = a_stage + tap_off[t];
#pragma unroll
                    for (int j = 0; j < MT; ++j) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_d + j * BN, a_tap + (uint32_t)((j * 16 * 1280 + k * 32) >> 4), b_tile + (uint32_t)(k * 2), idesc,
                                      (t | k) != 0 ? 1u : (uint32_t)(c != 0));
                    }
                    if (!RES) {
                        umma_commit(&emptyB[sb]);
                        if (++sb == Cfg::NB) { sb = 0; pb ^= 1; }
                    }
                }
                umma_commit(&emptyA[sa]);
                if (++sa == p.na) { sa = 0; pa ^= 1; }
            }
            umma_commit(&tmemFull[acc]);
        }
    }
}


template <int BN, int MT, bool RES, int EW, int STATS>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
               const ConvParams p) {
    typedef ConvCfg<BN, MT, RES> Cfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + p.na * Cfg::A_BYTES;
    const int b_tiles = RES ? 9 * p.kchunks : Cfg::NB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + b_tiles * Cfg::B_TILE);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + kMaxStages;
    uint64_t* fullB = emptyA + kMaxStages;      // RES: fullB[0] = "weights resident"
    uint64_t* emptyB = fullB + kMaxStages;
    uint64_t* tmemFull = emptyB + kMaxStages;
    uint64_t* tmemEmpty = tmemFull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmemEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1);
            mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmemFull[i], 1); mbar_init(&tmemEmpty[i], EW); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        if (RES) {
            mbar_expect_tx(&fullB[0], (uint32_t)(b_tiles * Cfg::B_TILE));
            for (int c = 0; c < p.kchunks; ++c)
                for (int t = 0; t < 9; ++t)
                    tma_load_3d(sB + (c * 9 + t) * Cfg::B_TILE, &tmB, &fullB[0], c * 64, 0, t);
        }
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int n0 = nt * BN;
            const int tw = mt % p.tiles_w, r = mt / p.tiles_w;
            const int th = r % p.tiles_h, c_n = r / p.tiles_h;
            const int c_h = th * Cfg::TH - 1, c_w = tw * 8 - 1;
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&emptyA[sa], pa ^ 1);
                mbar_expect_tx(&fullA[sa], Cfg::A_TX);
                if (c < p.kc1) tma_load_4d(sA + sa * Cfg::A_BYTES, &tmA, &fullA[sa], c * 64, c_w, c_h, c_n);
                else tma_load_4d(sA + sa * Cfg::A_BYTES, &tmA2, &fullA[sa], (c - p.kc1) * 64, c_w, c_h, c_n);
                if (++sa == p.na) { sa = 0; pa ^= 1; }
                if (!RES) {
                    for (int t = 0; t < 9; ++t) {
                        mbar_wait(&emptyB[sb], pb ^ 1);
                        mbar_expect_tx(&fullB[sb], Cfg::B_TILE);
                        tma_load_3d(sB + sb * Cfg::B_TILE, &tmB, &fullB[sb], c * 64, n0, t);
                        if (++sb == Cfg::NB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        // The first allocation of a CTA that owns its SM starts at TMEM column 0; with the base a literal the
        // accumulator addresses are immediates and an MMA costs ~4 issue slots instead of ~9 (N = 64 MMAs last 32 cycles).
        if (tmem_base == 0) conv_mma_loop<BN, MT, RES>(p, sA, sB, fullA, emptyA, fullB, emptyB, tmemFull, tmemEmpty, 0u, total_tiles);
        else conv_mma_loop<BN, MT, RES>(p, sA, sB, fullA, emptyA, fullB, emptyB, tmemFull, tmemEmpty, tmem_base, total_tiles);
    } else if (warp >= 2) {
        // ===================================================================== epilogue (EW = 4 or 8 warps)
        // With 8 warps two warps share each TMEM lane quarter and split the tile's MT * BN / 32 column chunks.  The TMEM
        // read of chunk i+1 is in flight while chunk i is converted, transposed (epi_store_chunk) and stored.
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;       // 0 when EW == 4
        const EpiLane L = epi_lane(reinterpret_cast<uint8_t*>(tmem_ptr) + 64 + (warp - 2) * 2048, lane);
        constexpr int NCH_ALL = MT * BN / 32;
        constexpr int NCH = NCH_ALL / (EW / 4);   // chunks per warp and tile
        float st[STATS ? NCH : 1][2];             // fused BatchNorm statistics: two finished values per chunk (tc_common.cuh);
#pragma unroll                                    // a template flag, so that the data-gradient launches carry none of it
        for (int ci = 0; ci < (STATS ? NCH : 1); ++ci) st[ci][0] = st[ci][1] = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            // unsigned 32-bit tile arithmetic; destinations in units of one 16-byte vector (8 bf16) -- the epilogue, not the
            // MMA, bounds the N = 64 layers and half of its instructions used to be 64-bit address arithmetic
            const uint32_t mt = (uint32_t)tile / (uint32_t)p.n_tiles, nt = (uint32_t)tile - mt * (uint32_t)p.n_tiles;
            const int n0 = (int)nt * BN;
            const uint32_t rr = mt / (uint32_t)p.tiles_w, tw = mt - rr * (uint32_t)p.tiles_w;
            const uint32_t n = rr / (uint32_t)p.tiles_h, th = rr - n * (uint32_t)p.tiles_h;
            const int w = (int)tw * 8 + L.row_lo;
            // rows (row_lo + 8 i) of this warp's quarter are image rows h0 + i of accumulator j (+ 16 j)
            const int h0 = (int)th * Cfg::TH + q * 4;
            const uint32_t ntot_v = (uint32_t)(p.split ? p.split : p.Ntot) >> 3, rowstep_v = (uint32_t)p.W * ntot_v;
            const uint32_t pix_v = ((n * (uint32_t)p.H + (uint32_t)h0) * (uint32_t)p.W + (uint32_t)w) * ntot_v + (uint32_t)L.slot;
            bf16* const pix = p.out + (size_t)pix_v * 8;
            bf16* const pix2 = p.split ? p.out2 + (size_t)pix_v * 8 : nullptr;
            const bool wok = w < p.W;
            // chunk ci lies in the second output tensor (split launches only)
            auto chunk_second = [&](int ci) -> bool {
                const int col = (half * NCH + ci) * 32;
                return p.split && n0 + col % BN >= p.split;
            };
            // element offset (from `pix` / `pix2`) of row i of chunk ci, or -1 when the pixel lies outside the image
            auto chunk_off = [&](int ci, int i) -> long long {
                const int col = (half * NCH + ci) * 32;
                const int j = col / BN, cc = col - j * BN;
                const int h = h0 + j * 16 + i;
                int g = n0 + cc;
                if (p.split && g >= p.split) g -= p.split;
                return (wok && h < p.H) ? (long long)(((uint32_t)(j * 16 + i) * rowstep_v + ((uint32_t)g >> 3)) << 3) : -1;
            };
            // STATS == 2: the BatchNorm input z is needed at every stored position.  Its lines are pulled into L2 for the
            // whole tile and the first chunk's vectors are loaded BEFORE waiting for the accumulator; the next chunk's are
            // loaded while the current one is processed -- no exposed DRAM round trip.
            const bf16* const zpix = STATS == 2 ? p.bz + (pix - p.out) : nullptr;
            uint4 zbuf[STATS == 2 ? 2 : 1][4];
            auto load_z = [&](int ci, uint4 (&dstv)[4]) {
                const bool sec = chunk_second(ci);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long o = chunk_off(ci, i);
                    dstv[i] = (o >= 0 && !sec) ? ldg_early(zpix + o) : make_uint4(0, 0, 0, 0);
                }
            };
            if (STATS == 2) {
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long o = chunk_off(ci, i);
                        if (o >= 0 && !chunk_second(ci)) prefetch_l2(zpix + o);
                    }
                load_z(0, zbuf[0]);
            }
            mbar_wait(&tmemFull[acc], acc_par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * BN) + half * NCH * 32;
            // STATS == 2 spends its registers on the prefetched z vectors instead of a second TMEM read buffer
            constexpr int TB = STATS == 2 ? 1 : 2;
            uint32_t buf[TB][32];
            tmem_ld32_async(taddr, buf[0]);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                tmem_ld_wait();
                if (TB == 2 && ci + 1 < NCH) tmem_ld32_async(taddr + (ci + 1) * 32, buf[(ci + 1) % TB]);
                if (STATS == 2 && ci + 1 < NCH) load_z(ci + 1, zbuf[STATS == 2 ? (ci + 1) & 1 : 0]);
                const int col = (half * NCH + ci) * 32;       // column within the MT * BN accumulator block
                const int cc = col % BN;
                bf16* dst[4];
                const bool sec = chunk_second(ci);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long o = chunk_off(ci, i);
                    dst[i] = o >= 0 ? (sec ? pix2 : pix) + o : nullptr;
                }
                if (STATS == 2) {
                    if (!sec) {
                        const EpiBnBwd bb{zbuf[STATS == 2 ? ci & 1 : 0], p.bcst + n0 + cc + L.slot * 8, p.brelu};
                        epi_store_chunk(L, buf[0], nullptr, 0, dst, &st[STATS ? ci : 0], lane, &bb);
                    } else epi_store_chunk(L, buf[0], nullptr, 0, dst, nullptr, lane);      // (the skip's half: no BatchNorm behind it here)
                    if (ci + 1 < NCH) tmem_ld32_async(taddr + (ci + 1) * 32, buf[0]);
                } else
                    epi_store_chunk(L, buf[ci % TB], p.bias ? p.bias + n0 + cc : nullptr, p.relu, dst, STATS ? &st[STATS ? ci : 0] : nullptr, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmemEmpty[acc]);
        }
        if (STATS) {
            // the grid is a multiple of n_tiles (host check): all tiles of this CTA lie in N tile blockIdx.x % n_tiles
            const int n_own = (blockIdx.x % p.n_tiles) * BN;
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int gcol = n_own + ((half * NCH + ci) * 32) % BN;
                if (!p.split || gcol < p.split) epi_stats_flush(p.bn_sums, p.split ? p.split : p.Ntot, gcol, lane, st[STATS ? ci : 0]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int MT, bool RES>
static int launch_conv(const void* x, const void* wk, ConvParams p, int Cin, int Cout, cudaStream_t st, const char* what,
                       const void* x2 = nullptr, int Cin1 = 0) {
    typedef ConvCfg<BN, MT, RES> Cfg;
    const int b_bytes = (RES ? 9 * p.kchunks : Cfg::NB) * Cfg::B_TILE;
    // 8 epilogue warps (16 KB of transposition buffers) unless that would leave fewer than three activation stages
    int ew = 8;
    int na = (kSmemLimit - 1024 - ew * 2048 - b_bytes) / Cfg::A_BYTES;
    if (na < 3) {
        ew = 4;
        na = (kSmemLimit - 1024 - ew * 2048 - b_bytes) / Cfg::A_BYTES;
    }
    if (na > 6) na = 6;
    if (na < 2) {
        set_error("%s: weights (%d B) leave no room for two activation stages", what, b_bytes);
        return EEL_ERR_INVALID;
    }
    p.na = na;
    const int smem = na * Cfg::A_BYTES + b_bytes + 1024 /* barriers */ + ew * 2048 + 1024 /* alignment slack */;
    const int stats = p.bn_sums == nullptr ? 0 : (p.bz != nullptr ? 2 : 1);
    static SmemOptIn configured[6];
    const int variant = (ew == 8 ? 1 : 0) + 2 * stats;
    const void* fns[6] = {(const void*)tc_conv_kernel<BN, MT, RES, 4, 0>, (const void*)tc_conv_kernel<BN, MT, RES, 8, 0>,
                          (const void*)tc_conv_kernel<BN, MT, RES, 4, 1>, (const void*)tc_conv_kernel<BN, MT, RES, 8, 1>,
                          (const void*)tc_conv_kernel<BN, MT, RES, 4, 2>, (const void*)tc_conv_kernel<BN, MT, RES, 8, 2>};
    if (!configured[variant].ensure(fns[variant], smem)) {
        set_error("%s: cannot raise dynamic shared memory to %d", what, smem);
        return EEL_ERR_CUDA;
    }
    CUtensorMap tmA, tmA2, tmB;
    const int Ca = x2 != nullptr ? Cin1 : Cin;      // channels of the first (or only) input tensor
    {
        uint64_t dims[4] = {(uint64_t)Ca, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
        uint64_t str[4] = {1, (uint64_t)Ca, (uint64_t)p.W * Ca, (uint64_t)p.H * p.W * Ca};
        uint32_t box[4] = {64, 10, (uint32_t)(Cfg::TH + 2), 1};
        if (int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, what)) return rc;
    }
    tmA2 = tmA;
    p.kc1 = p.kchunks;
    if (x2 != nullptr) {
        const int Cb = Cin - Cin1;
        uint64_t dims[4] = {(uint64_t)Cb, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
        uint64_t str[4] = {1, (uint64_t)Cb, (uint64_t)p.W * Cb, (uint64_t)p.H * p.W * Cb};
        uint32_t box[4] = {64, 10, (uint32_t)(Cfg::TH + 2), 1};
        if (int rc = make_tmap_bf16(&tmA2, x2, 4, dims, str, box, what)) return rc;
        p.kc1 = Cin1 / 64;
    }
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
        uint64_t str[3] = {1, (uint64_t)Cin, (uint64_t)Cin * Cout};
        uint32_t box[3] = {64, (uint32_t)BN, 1};
        if (int rc = make_tmap_bf16(&tmB, wk, 3, dims, str, box, what)) return rc;
    }
    p.tiles_h = cdiv(p.H, Cfg::TH);
    p.tiles_w = cdiv(p.W, 8);
    p.m_tiles = p.N * p.tiles_h * p.tiles_w;
    p.n_tiles = Cout / BN;
    if (p.bn_sums != nullptr) {
        const int all = p.m_tiles * p.n_tiles;
        if ((all < kNumSMs ? all : kNumSMs) % p.n_tiles != 0) {
            set_error("%s: fused BatchNorm statistics need a grid that is a multiple of the %d N tiles (Cout %d)", what, p.n_tiles, Cout);
            return EEL_ERR_INVALID;
        }
        if (cudaMemsetAsync(p.bn_sums, 0, sizeof(float) * 2 * (p.split ? p.split : Cout), st) != cudaSuccess) {
            set_error("%s: memset failed", what);
            return EEL_ERR_CUDA;
        }
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    void* args[4] = {(void*)&tmA, (void*)&tmA2, (void*)&tmB, (void*)&p};
    if (cudaLaunchKernel(fns[variant], dim3(grid), dim3(64 + 32 * ew), args, smem, st) != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        return EEL_ERR_CUDA;
    }
    return check_launch(what);
}

}  // namespace tc
}  // namespace eel

using namespace eel;
using namespace eel::tc;

// {scale, -shift} per channel (BatchNorm output > 0  <=>  scale * z > -shift), as the fused BatchNorm-backward epilogue reads them
__global__ void bn_consts_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float2* __restrict__ out, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] * rstd[c];
    out[c] = make_float2(sc, mean[c] * sc - beta[c]);
}
// the epilogue accumulated {sum g, sum g * z}: sum g * xhat = rstd * (sum g z - mean * sum g)
__global__ void bn_sums_fix_kernel(float* __restrict__ sums, const float* __restrict__ mean, const float* __restrict__ rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    sums[C + c] = rstd[c] * (sums[C + c] - mean[c] * sums[c]);
}

static int conv3x3_dispatch(const void* x, const void* wk, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                            int relu, int flip, float* bn_sums, const bf16* bz, const float2* bcst, int brelu, cudaStream_t st,
                            void* y2 = nullptr, int split = 0, const void* x2 = nullptr, int Cin1 = 0) {
    EEL_REQUIRE(x && wk && y && N > 0 && H > 0 && W > 0, "tc_conv3x3: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_conv3x3: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    EEL_REQUIRE((long long)N * H * W * Cout / 8 < (1LL << 32), "tc_conv3x3: output too large for 32-bit vector offsets");
    ConvParams p{};
    p.kchunks = Cin / 64;
    p.Ntot = Cout;
    p.N = N; p.H = H; p.W = W;
    p.flip = flip; p.relu = relu;
    p.bias = bias; p.out = (bf16*)y;
    p.bn_sums = bn_sums;
    p.bz = bz; p.bcst = bcst; p.brelu = brelu;
    p.out2 = (bf16*)y2; p.split = split;
    const bool tall = H > 16;                  // a 32-row tile would be half empty on 16-row maps
    if (Cin == 64 && Cout == 64)
        return tall ? launch_conv<64, 2, true>(x, wk, p, Cin, Cout, st, "tc_conv3x3(res 64x64)", x2, Cin1)
                    : launch_conv<64, 1, true>(x, wk, p, Cin, Cout, st, "tc_conv3x3(res 64x64)", x2, Cin1);
    if (Cin == 128 && Cout == 64) return launch_conv<64, 1, true>(x, wk, p, Cin, Cout, st, "tc_conv3x3(res 128x64)", x2, Cin1);
    if (Cin == 64 && Cout == 128) return launch_conv<128, 1, true>(x, wk, p, Cin, Cout, st, "tc_conv3x3(res 64x128)", x2, Cin1);
    // N = 256 MMAs read the least shared memory per FLOP (A 4 KB + B 8 KB per 128 cycles): preferred when Cout allows
    if (Cout % 256 == 0) return launch_conv<256, 1, false>(x, wk, p, Cin, Cout, st, "tc_conv3x3(256x1)", x2, Cin1);
    if (Cout % 128 == 0)
        return tall ? launch_conv<128, 2, false>(x, wk, p, Cin, Cout, st, "tc_conv3x3(128x2)", x2, Cin1)
                    : launch_conv<128, 1, false>(x, wk, p, Cin, Cout, st, "tc_conv3x3(128x1)", x2, Cin1);
    return tall ? launch_conv<64, 2, false>(x, wk, p, Cin, Cout, st, "tc_conv3x3(64x2)", x2, Cin1)
                : launch_conv<64, 1, false>(x, wk, p, Cin, Cout, st, "tc_conv3x3(64x1)", x2, Cin1);
}

extern "C" {

int eel_tc_conv3x3(const void* x, const void* wk, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                   int relu, int flip, float* bn_sums, eel_stream s) {
    return conv3x3_dispatch(x, wk, bias, y, N, H, W, Cin, Cout, relu, flip, bn_sums, nullptr, nullptr, 0, (cudaStream_t)s);
}

// Data gradient of a conv3x3 whose INPUT came from BatchNorm -> ReLU (models/EELUnet.py:338-344): the same launch as
// eel_tc_conv3x3(dy, wk, flip = 1) that also accumulates that BatchNorm's backward sums in its epilogue.
int eel_tc_conv3x3_dgrad_bnsums(const void* dy, const void* wk, void* dx, int N, int H, int W, int Cin, int Cout, const void* z,
                                const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                                float* sums, void* consts_ws, eel_stream s) {
    EEL_REQUIRE(dy && wk && dx && z && mean && rstd && gamma && beta && sums && consts_ws && N > 0 && H > 0 && W > 0,
                "tc_conv3x3_dgrad_bnsums: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    bn_consts_kernel<<<cdiv(Cout, 128), 128, 0, st>>>(mean, rstd, gamma, beta, (float2*)consts_ws, Cout);
    if (int rc = check_launch("tc_conv3x3_dgrad_bnsums.consts")) return rc;
    if (int rc = conv3x3_dispatch(dy, wk, nullptr, dx, N, H, W, Cin, Cout, 0, 1, sums, (const bf16*)z, (const float2*)consts_ws, relu, st)) return rc;
    bn_sums_fix_kernel<<<cdiv(Cout, 128), 128, 0, st>>>(sums, mean, rstd, Cout);
    return check_launch("tc_conv3x3_dgrad_bnsums.fix");
}

// conv3x3 whose input channels come from TWO tensors (x1: C1 channels, x2: C2): the forward of the conv that reads a skip bridge.
// wk is the forward operand [ky][kx][co][ci] with its ci columns de-interleaved (eel_cols_deinterleave): x1 = the even input
// channels (BatchNorm(upconv) + edge feature), x2 = the odd ones (the encoder skip).
int eel_tc_conv3x3_2src(const void* x1, const void* x2, const void* wk, const float* bias, void* y, int N, int H, int W, int C1, int C2,
                        int Cout, int relu, float* bn_sums, eel_stream s) {
    EEL_REQUIRE(x1 && x2 && C1 > 0 && C2 > 0 && C1 % 64 == 0 && C2 % 64 == 0, "tc_conv3x3_2src: both inputs need multiples of 64 channels");
    return conv3x3_dispatch(x1, wk, bias, y, N, H, W, C1 + C2, Cout, relu, 0, bn_sums, nullptr, nullptr, 0, (cudaStream_t)s, nullptr, 0, x2, C1);
}

// Data gradient of the conv that reads a skip bridge (channels interleaved as (upconv + edge feature, encoder skip), models/EELUnet.py:
// 132-141 then :338): wk is the data-gradient operand with its Cout rows DE-INTERLEAVED (even channels first), so the two halves of
// the result are stored as two [N,H,W,Cout/2] tensors -- the gradient of the sum and the gradient of the skip -- and
// eel_add_interleave_bwd never runs; with z given the epilogue also accumulates the upconv BatchNorm's backward sums over the first half.
int eel_tc_conv3x3_dgrad_split(const void* dy, const void* wk, void* dx0, void* dx1, int N, int H, int W, int Cin, int Cout,
                               const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                               float* sums, void* consts_ws, eel_stream s) {
    EEL_REQUIRE(dy && wk && dx0 && dx1 && N > 0 && H > 0 && W > 0, "tc_conv3x3_dgrad_split: bad argument");
    EEL_REQUIRE(Cout % 128 == 0, "tc_conv3x3_dgrad_split: the two halves must be multiples of 64 channels (Cout %d)", Cout);
    EEL_REQUIRE(z == nullptr || (mean && rstd && gamma && beta && sums && consts_ws), "tc_conv3x3_dgrad_split: the BatchNorm is incomplete");
    cudaStream_t st = (cudaStream_t)s;
    const int C = Cout / 2;
    if (z != nullptr) {
        bn_consts_kernel<<<cdiv(C, 128), 128, 0, st>>>(mean, rstd, gamma, beta, (float2*)consts_ws, C);
        if (int rc = check_launch("tc_conv3x3_dgrad_split.consts")) return rc;
    }
    if (int rc = conv3x3_dispatch(dy, wk, nullptr, dx0, N, H, W, Cin, Cout, 0, 1, z != nullptr ? sums : nullptr, (const bf16*)z,
                                  (const float2*)consts_ws, relu, st, dx1, C)) return rc;
    if (z != nullptr) {
        bn_sums_fix_kernel<<<cdiv(C, 128), 128, 0, st>>>(sums, mean, rstd, C);
        return check_launch("tc_conv3x3_dgrad_split.fix");
    }
    return EEL_OK;
}

}  // extern "C"
