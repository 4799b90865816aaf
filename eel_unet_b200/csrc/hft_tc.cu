// HighFourierTransform on the tensor cores (bf16 mode): the two large projections of hft.cu -- step 1
// (along W: [2F x W] . x[n,h]) and step 4 (back along W: [2W x 2F] . T3[n,h], fused with |x - low| or with the
// gradient subtraction) -- as tcgen05 GEMMs whose M dimension is the CHANNEL axis of the NHWC tensor:
//
//     D[(row j, c)][col] = sum_k A[n, h0 + j, k, c] * Bmat[col][k]
//
// A (MN-major: channels contiguous, K = pixel / frequency rows) streams through a TMA ring; the small DFT matrix
// Bmat (K-major, bf16) is loaded ONCE per CTA and stays resident in shared memory; the accumulators live in TMEM.
// With C = 64 two image rows share one M = 128 tile, with C = 128 one row does.  Both steps are HBM-bound: x is
// read once per step and y / phase are written once.
#include "tc_common.cuh"

namespace eel {
namespace tc {

enum { HEPI_T = 0, HEPI_ABS = 1, HEPI_SUB = 2 };

struct HftTcParams {
    int items;          // N * H / rows_per_item
    int H, C, W;
    int rows_per_item;  // 128 / C
    int kchunks;        // ceil(K / 64)
    int k16_last;       // MMAs (K = 16) in the last chunk
    int nblocks;        // column blocks of NB
    int n_stages;
    int ncols;          // valid output columns (T step: 2F)
    int mode;           // 0: step 1 (rows of x / g);  1: step 2 (T1b[n] with K = (re/im, h), M = 128-blocks of (f, c))
    int mtiles;         // mode 1: 128-blocks of the F*C axis
    int hc;             // mode 1: 64-row chunks per re/im half (H / 64)
    long long Cst;      // output channel extent (C, or F*C in mode 1)
    int kc_begin;       // first GLOBAL 64-wide K chunk of this launch (K split over launches when the matrix does not fit)
    int accumulate;     // HEPI_T: add the fp32 T already in memory (second launch of a K split) before storing
    float* T;           // fp32 output [rows][ncols][Cst] (or null)
    bf16* Tb;           // bf16 output, same layout (or null)
    const bf16* x;      // HEPI_ABS: input x [N,H,W,C];  HEPI_SUB: g [N,H,W,2,C]
    bf16* y;            // HEPI_ABS: |z| ; HEPI_SUB: dx
    bf16* phase;        // HEPI_ABS: z/|z| [N,H,W,2,C]
    bf16* gre;          // hft_tc1g_kernel: Re(dy * phase) [N,H,W,C], kept for the last step of the backward
};

constexpr int kHThreads = 192;
constexpr int kAStage = 16384;   // 2 atoms x 64 rows x 128 B

template <int NB, int EPI>
__global__ void __launch_bounds__(kHThreads, 1)
hft_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HftTcParams p) {
    constexpr int BT = NB * 128;   // bytes of one resident B tile (NB rows x 128 B)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.kchunks * p.nblocks * BT;
    uint8_t* sB = smem;
    uint8_t* sA = smem + ((b_bytes + 1023) & ~1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(sA + p.n_stages * kAStage);
    uint64_t* empty = full + p.n_stages;
    uint64_t* bFull = empty + p.n_stages;
    uint64_t* accFull = bFull + 1;     // [2]
    uint64_t* accEmpty = accFull + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int acc_cols = p.nblocks * (NB == 80 ? 128 : NB);   // column pitch 128 for the 80-wide tiles
    const int nacc = 2 * acc_cols <= 512 ? 2 : 1;
    const uint32_t tmem_cols = nacc * acc_cols <= 128 ? 128 : (nacc * acc_cols <= 256 ? 256 : 512);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(bFull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        mbar_expect_tx(bFull, b_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
            for (int nb = 0; nb < p.nblocks; ++nb) tma_load_2d(sB + (kc * p.nblocks + nb) * BT, &tmB, bFull, (p.kc_begin + kc) * 64, nb * NB);
        int st = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int hb = p.mode == 0 ? p.H / p.rows_per_item : p.mtiles;
            const int n = item / hb, h0 = (item - n * hb) * p.rows_per_item;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], kAStage);
                uint8_t* a = sA + st * kAStage;
                if (p.mode == 1) {
                    const int mt = item - n * hb;
                    const int g = p.kc_begin + kc;     // global K chunk: (ri, h) = (g / hc, 64 * (g % hc))
                    tma_load_4d(a, &tmA, &full[st], mt * 128, (g % p.hc) * 64, g / p.hc, n);
                    tma_load_4d(a + 8192, &tmA, &full[st], mt * 128 + 64, (g % p.hc) * 64, g / p.hc, n);
                } else if (p.rows_per_item == 2) {
                    tma_load_4d(a, &tmA, &full[st], 0, kc * 64, h0, n);
                    tma_load_4d(a + 8192, &tmA, &full[st], 0, kc * 64, h0 + 1, n);
                } else {
                    tma_load_4d(a, &tmA, &full[st], 0, kc * 64, h0, n);
                    tma_load_4d(a + 8192, &tmA, &full[st], 64, kc * 64, h0, n);
                }
                if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, NB, 1, 0);   // A: MN-major, B: K-major
        mbar_wait(bFull, 0);
        int st = 0, it = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int acc = nacc == 2 ? (it & 1) : 0;
            const uint32_t par = nacc == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&accEmpty[acc], par ^ 1);
            tc_fence_after();
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + st * kAStage);
                const int nk = kc == p.kchunks - 1 ? p.k16_last : 4;
                for (int nb = 0; nb < p.nblocks; ++nb) {
                    const uint32_t b = smem_u32(sB + (kc * p.nblocks + nb) * BT);
                    const uint32_t d = tmem_base + acc * acc_cols + nb * (NB == 80 ? 128 : NB);
                    for (int ks = 0; ks < nk; ++ks)
                        umma_bf16(d, make_smem_desc(a + ks * 2048, 8192, 1024, false), make_smem_desc(b + ks * 32, 16, 1024, false),
                                  idesc, (kc | ks) != 0);
                }
                umma_commit(&empty[st]);
                if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
            umma_commit(&accFull[acc]);
        }
    } else if (warp >= 2) {
        // ===================================================================== epilogue
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int j = p.rows_per_item == 2 ? (r >> 6) : 0;
        const int c = p.rows_per_item == 2 ? (r & 63) : r;
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int acc = nacc == 2 ? (it & 1) : 0;
            const uint32_t par = nacc == 2 ? ((it >> 1) & 1) : (it & 1);
            const int hb = p.mode == 0 ? p.H / p.rows_per_item : p.mtiles;
            const int n = item / hb;
            long long row, ch;
            if (p.mode == 0) { row = (long long)n * p.H + (item - n * hb) * p.rows_per_item + j; ch = c; }
            else { row = n; ch = (long long)(item - n * hb) * 128 + r; }
            mbar_wait(&accFull[acc], par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
            if (EPI == HEPI_T) {
                const long long off = row * p.ncols * p.Cst + ch;
#pragma unroll 1
                for (int cc = 0; cc < p.ncols; cc += 16) {
                    float v[32];
                    tmem_ld32(taddr + cc, v);    // reads 32 columns; only the first 16 are consumed per step (80 = 5 x 16)
                    if (p.accumulate) {
#pragma unroll
                        for (int t = 0; t < 16; ++t) v[t] += p.T[off + (long long)(cc + t) * p.Cst];
                    }
                    if (p.Tb != nullptr) {
#pragma unroll
                        for (int t = 0; t < 16; ++t) p.Tb[off + (long long)(cc + t) * p.Cst] = __float2bfloat16_rn(v[t]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 16; ++t) p.T[off + (long long)(cc + t) * p.Cst] = v[t];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accEmpty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- step 4 with the PIXEL axis as M -------------------------------------------------------------------------
// D[(w, re/im)][c] = sum_k Bmat[(w, re/im)][k] * T3b[n, h, k, c].  The DFT matrix (K-major, rows = output pixels) is
// the resident A operand, T3b[n,h] (MN-major, N = C) streams.  A thread of the epilogue owns one (pixel, re/im) row and
// 32 consecutive channels, so x / y / phase move as 64-byte vectors; re and im of a pixel sit in adjacent lanes and
// meet through one shuffle.
struct Hft4Params {
    int items;        // N * H
    int H, W, C;
    int mtiles;       // 128-row tiles handled by this launch
    int m_begin;      // first tile (the resident matrix holds tiles m_begin .. m_begin + mtiles)
    int stage_bytes;  // 2 k-chunks x (C/64) atoms x 8192
    int n_stages;
    int nacc;         // accumulator slots of C columns
    const bf16* x;    // fwd: x [N,H,W,C];  bwd: g [N,H,W,2,C]
    bf16* y;          // fwd: |z|;  bwd: dx;  step 3: T3b [N][H][2F][Cc]
    bf16* phase;      // fwd: z/|z| [N,H,W,2,C]
    int ntiles;       // step 3: 128-blocks of the F*Cc axis (items = N * ntiles)
    int F, Cc;        // step 3: frequencies, channels
    int xmul;         // bwd: row pitch of x in units of C (2: the real rows of [.., 2, C] gradient pairs;  1: a [.., C] tensor)
};

enum { H4_FWD = 0, H4_BWD = 1, H4_T3 = 2 };

// one MUFU.RSQ (2 ulp), no denormal fix-up: the epilogue is instruction-issue bound and its results are rounded to bf16
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int kH4Threads = 320;   // TMA warp, MMA warp, two epilogue warpgroups of 4 warps that alternate tiles

template <int C, int EPI>
__global__ void __launch_bounds__(kH4Threads, 1)
hft_tc4_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmT, const Hft4Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int m_bytes = p.mtiles * 2 * 16384;
    uint8_t* sM = smem;                 // [mtile][kchunk][128 rows x 128 B]
    uint8_t* sT = smem + m_bytes;       // ring of T3b[n,h]: [kchunk][atom][64 rows x 128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sT + p.n_stages * p.stage_bytes);
    uint64_t* empty = full + p.n_stages;
    uint64_t* mFull = empty + p.n_stages;
    uint64_t* accFull = mFull + 1;      // [8]
    uint64_t* accEmpty = accFull + 8;   // [8]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 8);
    constexpr int ATOMS = C / 64;
    constexpr bool FWD = EPI == H4_FWD;
    const int idiv = EPI == H4_T3 ? p.ntiles : p.H;   // items = N * idiv

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmM);
        tma_prefetch_desc(&tmT);
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(mFull, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        mbar_expect_tx(mFull, m_bytes);
        for (int mt = 0; mt < p.mtiles; ++mt)
            for (int kc = 0; kc < 2; ++kc) tma_load_2d(sM + (mt * 2 + kc) * 16384, &tmM, mFull, kc * 64, (p.m_begin + mt) * 128);
        int st = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / idiv, h = item - n * idiv;
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], p.stage_bytes);
            uint8_t* t = sT + st * p.stage_bytes;
            for (int kc = 0; kc < 2; ++kc)
                for (int a = 0; a < ATOMS; ++a) {
                    if (EPI == H4_T3) tma_load_4d(t + (kc * ATOMS + a) * 8192, &tmT, &full[st], h * 128 + a * 64, kc * 64, 0, n);
                    else tma_load_4d(t + (kc * ATOMS + a) * 8192, &tmT, &full[st], a * 64, kc * 64, h, n);
                }
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1 && elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, C, 0, 1);   // A: K-major (matrix), B: MN-major (T3b)
        mbar_wait(mFull, 0);
        int st = 0;
        uint32_t ph = 0;
        int tile_no = 0;                     // tiles per CTA stay far below 2^31: 32-bit divisions in the hot loops
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t t = smem_u32(sT + st * p.stage_bytes);
            for (int mt = 0; mt < p.mtiles; ++mt, ++tile_no) {
                const int slot = (int)(tile_no % p.nacc);
                const uint32_t par = (uint32_t)((tile_no / p.nacc) & 1);
                mbar_wait(&accEmpty[slot], par ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + slot * C;
#pragma unroll
                for (int k = 0; k < 5; ++k) {      // K = 80 = 5 x 16: chunk 0 holds k-steps 0..3, chunk 1 k-step 4
                    const int kc = k >> 2, ks = k & 3;
                    const uint32_t a = smem_u32(sM + (mt * 2 + kc) * 16384) + ks * 32;
                    const uint32_t b = t + kc * ATOMS * 8192 + ks * 2048;
                    umma_bf16(d, make_smem_desc(a, 16, 1024, false), make_smem_desc(b, 8192, 1024, false), idesc, k != 0);
                }
                umma_commit(&accFull[slot]);
            }
            umma_commit(&empty[st]);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;     // epilogue warpgroup 0 / 1
        const int r = q * 32 + lane;
        // output rows leave through the per-warp transposition buffer (tc_common.cuh): full 32-byte sectors per store
        const EpiLane L = epi_lane(reinterpret_cast<uint8_t*>(full) + 1024 + (warp - 2) * 2048, lane);
        int tile_no = 0;                     // tiles per CTA stay far below 2^31: 32-bit divisions in the hot loops
        // C == 64: the epilogue's global operand is double-buffered in registers one tile ahead (32 registers);
        // C == 128 would need 64 more registers than there are, so there the next tile's rows are pulled into L2
        constexpr bool DB = EPI != H4_T3 && C <= 64;
        uint4 nxt[DB ? C / 8 : 1];
        const int item_first = blockIdx.x + (grp / p.mtiles) * gridDim.x;     // this warpgroup's first tile: tile_no == grp
        if (DB && item_first < p.items) {
            const int n0 = item_first / idiv, h0 = item_first - n0 * idiv;
            const long long rb0 = ((long long)n0 * p.H + h0) * p.W;
            const int m0 = (p.m_begin + grp % p.mtiles) * 128 + r;
            const bf16* s0 = FWD ? p.x + (rb0 + (m0 >> 1)) * C : p.x + ((rb0 + m0) * p.xmul) * C;
            if (!FWD || (m0 & 1) == 0) {
#pragma unroll
                for (int i = 0; i < (DB ? C / 8 : 1); ++i) nxt[i] = reinterpret_cast<const uint4*>(s0)[i];
            }
        }
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / idiv, h = item - n * idiv;
            const long long rowbase = ((long long)n * p.H + h) * p.W;
            for (int mt = 0; mt < p.mtiles; ++mt, ++tile_no) {
                if ((int)(tile_no & 1) != grp) continue;
                const int slot = (int)(tile_no % p.nacc);
                const uint32_t par = (uint32_t)((tile_no / p.nacc) & 1);
                const int m = (p.m_begin + mt) * 128 + r;
                if (EPI == H4_T3) {
                    // rows m = (re/im, h'), columns = 128 of the (f, c) axis: T3b[n][h'][ro*F + f][c]
                    mbar_wait(&accFull[slot], par);
                    tc_fence_after();
                    const uint32_t taddr3 = tmem_base + ((uint32_t)(q * 32) << 16) + slot * C;
                    const int ro = m / p.H, hh = m - ro * p.H;
                    (void)ro; (void)hh;
                    bf16* dst3[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int mm = (p.m_begin + mt) * 128 + q * 32 + L.row_lo + 8 * i;
                        const int ro2 = mm / p.H, hh2 = mm - ro2 * p.H;
                        dst3[i] = p.y + (((long long)n * p.H + hh2) * 2 * p.F + (long long)ro2 * p.F) * p.Cc + (long long)h * 128 + L.slot * 8;
                    }
#pragma unroll
                    for (int cc = 0; cc < C; cc += 32) {
                        uint32_t rr[32];
                        tmem_ld32_async(taddr3 + cc, rr);
                        tmem_ld_wait();
                        bf16* d4[4] = {dst3[0] + cc, dst3[1] + cc, dst3[2] + cc, dst3[3] + cc};
                        epi_store_chunk(L, rr, nullptr, 0, d4);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accEmpty[slot]);
                    continue;
                }
                // the global operand of the epilogue (x, or the real half of g) is fetched BEFORE waiting for the MMAs; the
                // rows the NEXT tile of this warpgroup will need are pulled into L2 now (the MMAs are far ahead of the
                // epilogue, so without this every tile would expose a full DRAM round trip on its first use of `pre`)
                uint4 pre[C / 8];
                if (DB) {
#pragma unroll
                    for (int i = 0; i < C / 8; ++i) pre[i] = nxt[i < (DB ? C / 8 : 1) ? i : 0];
                }
                {
                    // this warpgroup's next tile is tile_no + 2 in the CTA's linear (item, mt) order
                    const int t2 = tile_no + 2;
                    const int mt2 = t2 % p.mtiles;
                    const int item2 = (int)blockIdx.x + (t2 / p.mtiles) * (int)gridDim.x;
                    if (item2 < p.items) {
                        const int n2 = item2 / idiv, h2 = item2 - n2 * idiv;
                        const long long rb2 = ((long long)n2 * p.H + h2) * p.W;
                        const int m2 = (p.m_begin + mt2) * 128 + r;
                        const bf16* nx = FWD ? p.x + (rb2 + (m2 >> 1)) * C : p.x + ((rb2 + m2) * p.xmul) * C;
                        if (!FWD || (m2 & 1) == 0) {
                            if (DB) {
#pragma unroll
                                for (int i = 0; i < (DB ? C / 8 : 1); ++i) nxt[i] = reinterpret_cast<const uint4*>(nx)[i];
                            } else {
#pragma unroll
                                for (int b = 0; b < C * 2; b += 128)
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(nx) + b));
                            }
                        }
                    }
                }
                if (!DB) {
                    const bf16* src = FWD ? p.x + (rowbase + (m >> 1)) * C : p.x + ((rowbase + m) * p.xmul) * C;
                    if (!FWD || (m & 1) == 0) {
#pragma unroll
                        for (int i = 0; i < C / 8; ++i) pre[i] = reinterpret_cast<const uint4*>(src)[i];
                    }
                }
                mbar_wait(&accFull[slot], par);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * C;
#pragma unroll
                for (int cc = 0; cc < C; cc += 32) {
                    float v[32];
                    tmem_ld32(taddr + cc, v);
                    if (FWD) {
                        const int w = m >> 1, ro = m & 1;
                        const long long e = (rowbase + w) * C + cc;
                        float mine[32];
                        if (ro == 0) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                Vec16<bf16> xv; xv.raw = pre[cc / 8 + i];
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj) mine[i * 8 + jj] = xv.get(jj) - v[i * 8 + jj];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) mine[i] = -v[i];
                        }
                        // re and im of a pixel sit in adjacent lanes: one shuffle per value, then |z| and z/|z| packed two
                        // at a time (one F2FP per pair instead of a convert + bit insert per element)
                        uint32_t pkp[16], pkm[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float o0 = __shfl_xor_sync(0xffffffffu, mine[i], 1);
                            const float o1 = __shfl_xor_sync(0xffffffffu, mine[i + 1], 1);
                            const float sq0 = fmaf(mine[i], mine[i], o0 * o0), sq1 = fmaf(mine[i + 1], mine[i + 1], o1 * o1);
                            const float inv0 = sq0 > 0.f ? rsqrt_approx(sq0) : 0.f, inv1 = sq1 > 0.f ? rsqrt_approx(sq1) : 0.f;
                            __nv_bfloat162 hm = __floats2bfloat162_rn(sq0 * inv0, sq1 * inv1);
                            __nv_bfloat162 hp = __floats2bfloat162_rn(mine[i] * inv0, mine[i + 1] * inv1);
                            pkm[i >> 1] = *reinterpret_cast<uint32_t*>(&hm);
                            pkp[i >> 1] = *reinterpret_cast<uint32_t*>(&hp);
                        }
                        bf16 *dp[4], *dm[4];
                        {
                            // rows (row_lo + 8 i) of this warp: phase rows are 8 * C apart, |z| rows 4 * C (even rows only)
                            const int mm0 = (p.m_begin + mt) * 128 + q * 32 + L.row_lo;
                            // (p.phase null: inference, the unit phase is not kept -- two thirds of this step's writes)
                            bf16* pb = p.phase != nullptr ? p.phase + (rowbase * 2 + mm0) * C + cc + L.slot * 8 : nullptr;
                            bf16* yb = p.y + (rowbase + (mm0 >> 1)) * C + cc + L.slot * 8;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                dp[i] = pb != nullptr ? pb + i * 8 * C : nullptr;
                                dm[i] = (mm0 & 1) ? nullptr : yb + i * 4 * C;
                            }
                        }
                        epi_store_packed(L, pkp, dp);
                        epi_store_packed(L, pkm, dm);
                    } else {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            Vec16<bf16> gv; gv.raw = pre[cc / 8 + i];   // real half of the (re, im) pair
#pragma unroll
                            for (int jj = 0; jj < 8; jj += 2) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(gv.get(jj) - v[i * 8 + jj], gv.get(jj + 1) - v[i * 8 + jj + 1]);
                                pk[4 * i + (jj >> 1)] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                        }
                        bf16* dd[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int mm = (p.m_begin + mt) * 128 + q * 32 + L.row_lo + 8 * i;
                            dd[i] = p.y + (rowbase + mm) * C + cc + L.slot * 8;
                        }
                        epi_store_packed(L, pk, dd);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&accEmpty[slot]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int C, int EPI>
static int launch_hft4(const CUtensorMap& m, const CUtensorMap& t, Hft4Params& p, cudaStream_t st, const char* what) {
    static SmemOptIn configured;
    const int kMax = 227 * 1024;
    if (!configured.ensure(hft_tc4_kernel<C, EPI>, kMax)) {
        set_error("%s: cannot raise dynamic shared memory", what);
        return EEL_ERR_CUDA;
    }
    p.stage_bytes = 2 * (C / 64) * 8192;
    p.nacc = 512 / C;
    const int total_tiles = p.mtiles;
    for (int mb = 0; mb < total_tiles; mb += 4) {      // at most 4 resident tiles (128 KB) per pass
        p.m_begin = mb;
        p.mtiles = total_tiles - mb < 4 ? total_tiles - mb : 4;
        const int m_bytes = p.mtiles * 2 * 16384;
        int ns = (kMax - 3072 - 16384 - m_bytes) / p.stage_bytes;   // 16 KB: epilogue transposition buffers
        if (ns > 6) ns = 6;
        if (ns < 2) { set_error("%s: resident matrix leaves no room for the operand ring", what); return EEL_ERR_INVALID; }
        p.n_stages = ns;
        const int smem = m_bytes + ns * p.stage_bytes + 3072 + 16384;
        const int grid = p.items < kNumSMs ? p.items : kNumSMs;
        hft_tc4_kernel<C, EPI><<<grid, kH4Threads, smem, st>>>(m, t, p);
        if (int rc = check_launch(what)) return rc;
    }
    return EEL_OK;
}

// =====================================================================================================================
// Second generation (training shapes, W <= 256): the unit phase z/|z| that the backward needs is kept as ONE 16-bit code
// per element instead of a (re, im) pair of bf16 -- half the bytes and ~40x finer than bf16 near |component| = 1:
//     bit 15: the SMALLER component is re;  bit 14: the larger component is negative;
//     bits 13..0: q + 8192, q = the smaller component in fixed point (step sqrt(1/2) / 8191, |q| <= 8191);  field 0 = zero vector
// (|smaller| <= sqrt(1/2), the larger one is +-sqrt(1 - smaller^2): its error is at most the smaller one's).
// With it (a) the forward's last step writes |z| and the code from an epilogue in which a thread owns one PIXEL (re and im
// of the low-pass part come from two accumulators of the same TMEM row: no shuffles, no duplicated rsqrt, no idle lanes),
// (b) the backward's first step forms the complex upstream gradient g = dy * phase INSIDE the kernel (transform warps between
// the TMA ring and the MMA: dy and code tiles in, the swizzled MN-major operand tile out) and its last step reads dy and the
// code again -- the [N,H,W,2,C] tensor of gradient pairs (written once, read twice) is gone.
constexpr float kPhaseStep = 0.70710678118f / 8191.0f;
constexpr float kPhaseInvStep = 8191.0f / 0.70710678118f;

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t phase_encode(float zr, float zi, float inv) {
    const float re = zr * inv, im = zi * inv;
    const bool sel = fabsf(re) < fabsf(im);
    const float small = sel ? re : im, large = sel ? im : re;
    int q = __float2int_rn(small * kPhaseInvStep);
    q = q > 8191 ? 8191 : (q < -8191 ? -8191 : q);
    const uint32_t c = (sel ? 0x8000u : 0u) | (large < 0.f ? 0x4000u : 0u) | (uint32_t)(q + 8192);
    return inv > 0.f ? c : 0u;            // biased field 0 (q = -8192, never produced otherwise) = the zero vector
}
// (re, im) * d for one code: the biased 14-bit field becomes a float through the 2^23 mantissa trick (no I2F), the larger
// component through one MUFU.SQRT; the zero vector zeroes d instead of both components
__device__ __forceinline__ void phase_apply(uint32_t code, float d, float& gre, float& gim) {
    const uint32_t u = code & 0x3FFFu;
    const float small = fmaf(__uint_as_float(0x4B000000u | u), kPhaseStep, -(8388608.f + 8192.f) * kPhaseStep);
    float large = sqrt_approx(fmaf(-small, small, 1.f));
    large = __uint_as_float(__float_as_uint(large) ^ ((code << 17) & 0x80000000u));
    const bool sel = (code & 0x8000u) != 0;
    d = u == 0 ? 0.f : d;
    gre = d * (sel ? small : large);
    gim = d * (sel ? large : small);
}
__device__ __forceinline__ float phase_apply_re(uint32_t code, float d) {
    const uint32_t u = code & 0x3FFFu;
    const float small = fmaf(__uint_as_float(0x4B000000u | u), kPhaseStep, -(8388608.f + 8192.f) * kPhaseStep);
    float large = sqrt_approx(fmaf(-small, small, 1.f));
    large = __uint_as_float(__float_as_uint(large) ^ ((code << 17) & 0x80000000u));
    d = u == 0 ? 0.f : d;
    return d * ((code & 0x8000u) ? small : large);
}

// ---- last step, pixel rows: D[w][(part, c)] = sum_k Mat_part[w][k] * T3b[n, h, k, c], part = re (fwd: and im) -----------
struct Hft4pParams {
    int items;            // N * H
    int H, W;
    int ptiles;           // 128-pixel tiles resident in this pass
    int pt_begin;         // first of them
    int stage_bytes, n_stages, nacc;
    const bf16* x;        // fwd: x;  bwd: dy
    const uint16_t* code_in;   // bwd
    bf16* y;              // fwd: |z|;  bwd: dx
    uint16_t* code_out;   // fwd
};

template <int C, bool FWD, int NG>
__global__ void __launch_bounds__(64 + 128 * NG, 1)
hft_tc4p_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmT, const Hft4pParams p) {
    constexpr int MT = FWD ? 2 : 1;          // matrix tiles per pixel tile (re, im | re)
    constexpr int SW = MT * C;               // accumulator slot width in TMEM columns
    constexpr int ATOMS = C / 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int m_bytes = p.ptiles * MT * 2 * 16384;
    uint8_t* sM = smem;                 // [ptile][part][kchunk][128 rows x 128 B]
    uint8_t* sT = smem + m_bytes;       // ring of T3b[n,h]: [kchunk][atom][64 rows x 128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sT + p.n_stages * p.stage_bytes);
    uint64_t* empty = full + p.n_stages;
    uint64_t* mFull = empty + p.n_stages;
    uint64_t* accFull = mFull + 1;      // [8]
    uint64_t* accEmpty = accFull + 8;   // [8]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmM);
        tma_prefetch_desc(&tmT);
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(mFull, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        mbar_expect_tx(mFull, m_bytes);
        for (int lt = 0; lt < p.ptiles * MT; ++lt)
            for (int kc = 0; kc < 2; ++kc) tma_load_2d(sM + (lt * 2 + kc) * 16384, &tmM, mFull, kc * 64, (p.pt_begin * MT + lt) * 128);
        int st = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / p.H, h = item - n * p.H;
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], p.stage_bytes);
            uint8_t* t = sT + st * p.stage_bytes;
            for (int kc = 0; kc < 2; ++kc)
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(t + (kc * ATOMS + a) * 8192, &tmT, &full[st], a * 64, kc * 64, h, n);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1 && elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, C, 0, 1);   // A: K-major (matrix), B: MN-major (T3b)
        mbar_wait(mFull, 0);
        int st = 0;
        uint32_t ph = 0;
        int tile_no = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t t = smem_u32(sT + st * p.stage_bytes);
            for (int lt = 0; lt < p.ptiles; ++lt, ++tile_no) {
                const int slot = tile_no % p.nacc;
                const uint32_t par = (uint32_t)((tile_no / p.nacc) & 1);
                mbar_wait(&accEmpty[slot], par ^ 1);
                tc_fence_after();
#pragma unroll
                for (int part = 0; part < MT; ++part) {
                    const uint32_t d = tmem_base + slot * SW + part * C;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {      // K = 80 = 5 x 16: chunk 0 holds k-steps 0..3, chunk 1 k-step 4
                        const int kc = k >> 2, ks = k & 3;
                        const uint32_t a = smem_u32(sM + ((lt * MT + part) * 2 + kc) * 16384) + ks * 32;
                        const uint32_t b = t + kc * ATOMS * 8192 + ks * 2048;
                        umma_bf16(d, make_smem_desc(a, 16, 1024, false), make_smem_desc(b, 8192, 1024, false), idesc, k != 0);
                    }
                }
                umma_commit(&accFull[slot]);
            }
            umma_commit(&empty[st]);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;     // epilogue warpgroups take tiles round robin
        const int r = q * 32 + lane;         // this thread's pixel inside the tile
        const EpiLane L = epi_lane(reinterpret_cast<uint8_t*>(full) + 1024 + (warp - 2) * 2048, lane);
        int tile_no = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / p.H, h = item - n * p.H;
            const long long rowbase = ((long long)n * p.H + h) * p.W;
            for (int lt = 0; lt < p.ptiles; ++lt, ++tile_no) {
                if (tile_no % NG != grp) continue;
                const int slot = tile_no % p.nacc;
                const uint32_t par = (uint32_t)((tile_no / p.nacc) & 1);
                const long long pix = rowbase + (p.pt_begin + lt) * 128 + r;
                // this thread's row of the epilogue's global operands, fetched BEFORE waiting for the MMAs; the rows of this
                // warpgroup's NEXT tile (tile_no + 2) are pulled into L2 now so that their load finds them there
                // (one 32-channel chunk at a time, the next chunk's loads in flight while this one is processed)
                uint4 pre[4], cd[4], pre_n[4], cd_n[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pre[i] = ldg_early(p.x + pix * C + i * 8);
                    if (!FWD) cd[i] = ldg_early(p.code_in + pix * C + i * 8);
                }
                {
                    const int t2 = tile_no + NG;
                    const int lt2 = t2 % p.ptiles;
                    const int item2 = (int)blockIdx.x + (t2 / p.ptiles) * (int)gridDim.x;
                    if (item2 < p.items) {
                        const long long pix2 = (long long)item2 * p.W + (p.pt_begin + lt2) * 128 + r;
#pragma unroll
                        for (int b = 0; b < C * 2; b += 128) {
                            prefetch_l2(reinterpret_cast<const uint8_t*>(p.x + pix2 * C) + b);
                            if (!FWD) prefetch_l2(reinterpret_cast<const uint8_t*>(p.code_in + pix2 * C) + b);
                        }
                    }
                }
                mbar_wait(&accFull[slot], par);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SW;
                // rows (row_lo + 8 i) of this warp after the transposition
                const long long prow = rowbase + (p.pt_begin + lt) * 128 + q * 32 + L.row_lo;
#pragma unroll
                for (int cc = 0; cc < C; cc += 32) {
                    bf16* dy_[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dy_[i] = p.y + (prow + 8 * i) * C + cc + L.slot * 8;
                    if (cc + 32 < C) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            pre_n[i] = ldg_early(p.x + pix * C + cc + 32 + i * 8);
                            if (!FWD) cd_n[i] = ldg_early(p.code_in + pix * C + cc + 32 + i * 8);
                        }
                    }
                    if (FWD) {
                        uint32_t rre[32], rim[32];
                        tmem_ld32_async(taddr + cc, rre);
                        tmem_ld32_async(taddr + C + cc, rim);
                        tmem_ld_wait();
                        uint32_t pkm[16], pkc[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            Vec16<bf16> xv; xv.raw = pre[i];
#pragma unroll
                            for (int jj = 0; jj < 8; jj += 2) {
                                float mg[2];
                                uint32_t co[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int e = i * 8 + jj + u;
                                    const float zr = xv.get(jj + u) - __uint_as_float(rre[e]);
                                    const float zi = -__uint_as_float(rim[e]);
                                    const float sq = fmaf(zr, zr, zi * zi);
                                    const float inv = sq > 0.f ? rsqrt_approx(sq) : 0.f;
                                    mg[u] = sq * inv;
                                    co[u] = phase_encode(zr, zi, inv);
                                }
                                __nv_bfloat162 hm = __floats2bfloat162_rn(mg[0], mg[1]);
                                pkm[(i * 8 + jj) >> 1] = *reinterpret_cast<uint32_t*>(&hm);
                                pkc[(i * 8 + jj) >> 1] = co[0] | (co[1] << 16);
                            }
                        }
                        bf16* dc_[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            dc_[i] = p.code_out != nullptr ? reinterpret_cast<bf16*>(p.code_out) + (prow + 8 * i) * C + cc + L.slot * 8 : nullptr;
                        epi_store_packed(L, pkm, dy_);
                        epi_store_packed(L, pkc, dc_);
                    } else {
                        float v[32];
                        tmem_ld32(taddr + cc, v);
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            Vec16<bf16> gv; gv.raw = pre[i];
                            const uint32_t cw[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w};
#pragma unroll
                            for (int jj = 0; jj < 8; jj += 2) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(phase_apply_re(cw[jj >> 1], gv.get(jj)) - v[i * 8 + jj],
                                                                          phase_apply_re(cw[jj >> 1] >> 16, gv.get(jj + 1)) - v[i * 8 + jj + 1]);
                                pk[4 * i + (jj >> 1)] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                        }
                        epi_store_packed(L, pk, dy_);
                    }
                    if (cc + 32 < C) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { pre[i] = pre_n[i]; if (!FWD) cd[i] = cd_n[i]; }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&accEmpty[slot]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int C, bool FWD>
static int launch_hft4p(const CUtensorMap& m, const CUtensorMap& t, Hft4pParams& p, int total_ptiles, cudaStream_t st, const char* what) {
    static SmemOptIn configured;
    const int kMax = 227 * 1024;
    constexpr int NG = (512 / ((FWD ? 2 : 1) * C)) >= 3 ? 3 : 2;     // epilogue warpgroups: as many as TMEM holds tiles, at most 3
    if (!configured.ensure(hft_tc4p_kernel<C, FWD, NG>, kMax)) {
        set_error("%s: cannot raise dynamic shared memory", what);
        return EEL_ERR_CUDA;
    }
    constexpr int MT = FWD ? 2 : 1;
    p.stage_bytes = 2 * (C / 64) * 8192;
    p.nacc = 512 / (MT * C);
    const int per_pass = 4 / MT;                        // at most 4 resident matrix tiles (128 KB) per pass
    for (int pb = 0; pb < total_ptiles; pb += per_pass) {
        p.pt_begin = pb;
        p.ptiles = total_ptiles - pb < per_pass ? total_ptiles - pb : per_pass;
        const int m_bytes = p.ptiles * MT * 2 * 16384;
        int ns = (kMax - 3072 - 8192 * NG - m_bytes) / p.stage_bytes;   // 8 KB per warpgroup: epilogue transposition buffers
        if (ns > 6) ns = 6;
        if (ns < 2) { set_error("%s: resident matrix leaves no room for the operand ring", what); return EEL_ERR_INVALID; }
        p.n_stages = ns;
        const int smem = m_bytes + ns * p.stage_bytes + 3072 + 8192 * NG;
        const int grid = p.items < kNumSMs ? p.items : kNumSMs;
        hft_tc4p_kernel<C, FWD, NG><<<grid, 64 + 128 * NG, smem, st>>>(m, t, p);
        if (int rc = check_launch(what)) return rc;
    }
    return EEL_OK;
}

// ---- backward step 1 with the complex gradient formed in the kernel ------------------------------------------------------
//     T1b[n][h][(ro, f)][c] = sum_(w, ri) Bmat[(ro, f)][2w + ri] * (dy[n,h,w,c] * phase[n,h,w,ri,c])
// Per K chunk (32 pixels = 64 operand rows) the TMA ring delivers the dy and code tiles of the item's two 64-channel atoms;
// eight transform warps decode, multiply and write the swizzled MN-major operand tile the MMA reads.
constexpr int kG1Threads = 32 * 14;     // TMA warp, MMA warp, 4 epilogue warps, 8 transform warps
constexpr int kG1Stage = 32768;         // 16 KB operand + 8 KB dy + 8 KB code

template <int C>
__global__ void __launch_bounds__(kG1Threads, 1)
hft_tc1g_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmCode,
                const __grid_constant__ CUtensorMap tmB, const HftTcParams p) {
    constexpr int NB = 80, BT = NB * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.kchunks * BT;
    uint8_t* sB = smem;
    uint8_t* sA = smem + ((b_bytes + 1023) & ~1023);
    uint64_t* rawFull = reinterpret_cast<uint64_t*>(sA + p.n_stages * kG1Stage);
    uint64_t* aFull = rawFull + p.n_stages;
    uint64_t* empty = aFull + p.n_stages;
    uint64_t* bFull = empty + p.n_stages;
    uint64_t* accFull = bFull + 1;     // [2]
    uint64_t* accEmpty = accFull + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int acc_cols = 128;      // 80 used
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmCode);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.n_stages; ++i) { mbar_init(&rawFull[i], 1); mbar_init(&aFull[i], 8); mbar_init(&empty[i], 1); }
        mbar_init(bFull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int hb = p.H / p.rows_per_item;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        mbar_expect_tx(bFull, b_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc) tma_load_2d(sB + kc * BT, &tmB, bFull, kc * 64, 0);
        int st = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / hb, h0 = (item - n * hb) * p.rows_per_item;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&rawFull[st], 16384);
                uint8_t* raw = sA + st * kG1Stage + 16384;
                for (int a = 0; a < 2; ++a) {
                    const int c0 = C == 64 ? 0 : a * 64, hh = C == 64 ? h0 + a : h0;
                    tma_load_4d(raw + a * 4096, &tmDy, &rawFull[st], c0, kc * 32, hh, n);
                    tma_load_4d(raw + 8192 + a * 4096, &tmCode, &rawFull[st], c0, kc * 32, hh, n);
                }
                if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, NB, 1, 0);   // A: MN-major, B: K-major
        mbar_wait(bFull, 0);
        int st = 0, it = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t par = (it >> 1) & 1;
            mbar_wait(&accEmpty[acc], par ^ 1);
            tc_fence_after();
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&aFull[st], ph);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + st * kG1Stage);
                const uint32_t b = smem_u32(sB + kc * BT);
                const uint32_t d = tmem_base + acc * acc_cols;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(d, make_smem_desc(a + ks * 2048, 8192, 1024, false), make_smem_desc(b + ks * 32, 16, 1024, false), idesc,
                              (kc | ks) != 0);
                umma_commit(&empty[st]);
                if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
            umma_commit(&accFull[acc]);
        }
    } else if (warp >= 2 && warp < 6) {
        // ===================================================================== epilogue: T1b[row][col][ch] (bf16)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int j = p.rows_per_item == 2 ? (r >> 6) : 0;
        const int c = p.rows_per_item == 2 ? (r & 63) : r;
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int n = item / hb;
            const long long row = (long long)n * p.H + (item - n * hb) * p.rows_per_item + j;
            mbar_wait(&accFull[acc], par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
            const long long off = row * p.ncols * p.Cst + c;
#pragma unroll 1
            for (int cc = 0; cc < p.ncols; cc += 16) {
                float v[32];
                tmem_ld32(taddr + cc, v);
#pragma unroll
                for (int t = 0; t < 16; ++t) p.Tb[off + (long long)(cc + t) * p.Cst] = __float2bfloat16_rn(v[t]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accEmpty[acc]);
        }
    } else if (warp >= 6) {
        // ===================================================================== transform: (dy, code) -> g = dy * phase
        const int tt = threadIdx.x - 192;     // 0 .. 255
        int st = 0;
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int n = item / hb, h0 = (item - n * hb) * p.rows_per_item;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&rawFull[st], ph);
                uint8_t* A = sA + st * kG1Stage;
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int itw = tt + rep * 256;
                    const int a = itw >> 8, pr = (itw >> 3) & 31, jc = itw & 7;
                    const uint32_t src = a * 4096 + pr * 128 + ((jc ^ (pr & 7)) << 4);
                    const uint4 dv = *reinterpret_cast<const uint4*>(A + 16384 + src);
                    const uint4 cv = *reinterpret_cast<const uint4*>(A + 24576 + src);
                    const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, cw[4] = {cv.x, cv.y, cv.z, cv.w};
                    uint32_t ore[4], oim[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float d0 = __uint_as_float(dw[k] << 16), d1 = __uint_as_float(dw[k] & 0xffff0000u);
                        float gr0, gi0, gr1, gi1;
                        phase_apply(cw[k], d0, gr0, gi0);
                        phase_apply(cw[k] >> 16, d1, gr1, gi1);
                        __nv_bfloat162 hr = __floats2bfloat162_rn(gr0, gr1);
                        __nv_bfloat162 hi = __floats2bfloat162_rn(gi0, gi1);
                        ore[k] = *reinterpret_cast<uint32_t*>(&hr);
                        oim[k] = *reinterpret_cast<uint32_t*>(&hi);
                    }
                    const int r0 = 2 * pr, r1 = 2 * pr + 1;
                    {   // Re(g) also goes to memory (8 lanes = one pixel's 128 contiguous bytes): the last step subtracts Re(low) from it
                        const int hh = C == 64 ? h0 + a : h0, c0 = (C == 64 ? 0 : a * 64) + jc * 8;
                        *reinterpret_cast<uint4*>(p.gre + (((long long)n * p.H + hh) * p.W + kc * 32 + pr) * C + c0) =
                            make_uint4(ore[0], ore[1], ore[2], ore[3]);
                    }
                    *reinterpret_cast<uint4*>(A + a * 8192 + r0 * 128 + ((jc ^ (r0 & 7)) << 4)) = make_uint4(ore[0], ore[1], ore[2], ore[3]);
                    *reinterpret_cast<uint4*>(A + a * 8192 + r1 * 128 + ((jc ^ (r1 & 7)) << 4)) = make_uint4(oim[0], oim[1], oim[2], oim[3]);
                }
                // the operand tile was written through the generic proxy and is read by tcgen05.mma through the async proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&aFull[st]);
                if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// real-expanded DFT matrices in bf16 (rows = output index, K contiguous, K padded to Kp)
//   kind 0 (step 1 fwd):  rows (ro, f), cols w            [[C], [-S]]
//   kind 1 (step 1 bwd):  rows (ro, f), cols 2w + ri      [[C, S], [-S, C]]
//   kind 2 (step 4 fwd):  rows 2w + ro, cols ri*F + f     [[C, -S], [S, C]]
//   kind 3 (step 4 bwd):  rows w,       cols ri*F + f     [C, -S]
//   kind 6 (step 4 fwd, pixel rows): per 128-pixel tile 128 rows [C, -S] (re) then 128 rows [S, C] (im)
__global__ void hft_tc_matrix_kernel(bf16* __restrict__ out, int kind, int rows, int Kp, int F, int r, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * Kp) return;
    const int row = i / Kp, k = i - row * Kp;
    int f = -1, w = 0, ro = 0, ri = 0;
    if (kind == 0) { ro = row / F; f = row % F; if (k < W) w = k; else f = -1; }
    else if (kind == 1) { ro = row / F; f = row % F; if (k < 2 * W) { w = k >> 1; ri = k & 1; } else f = -1; }
    else if (kind == 2) { w = row >> 1; ro = row & 1; if (k < 2 * F) { ri = k / F; f = k % F; } }
    else if (kind == 3) { w = row; ro = 0; if (k < 2 * F) { ri = k / F; f = k % F; } }
    else if (kind == 4) { ro = row / W; w = row % W; if (k < 2 * F) { ri = k / F; f = k % F; } }
    else if (kind == 6) { ro = (row >> 7) & 1; w = (row >> 8) * 128 + (row & 127); if (k < 2 * F) { ri = k / F; f = k % F; } }
    else { ro = row / F; f = row % F; if (k < 2 * W) { ri = k / W; w = k % W; } else f = -1; }
    float v = 0.f;
    if (f >= 0) {
        long long kk = f - r;
        long long ph = ((kk * w) % W + W) % W;
        double s, c;
        sincospi(2.0 * (double)ph / (double)W, &s, &c);
        const double sc = 1.0 / sqrt((double)W);
        const bool useS = ro != ri;
        const bool table1 = (kind >= 2 && kind <= 4) || kind == 6;
        double val = useS ? s : c;
        if (useS && (table1 ? (ro == 0) : (ro == 1))) val = -val;
        v = (float)(val * sc);
    }
    out[i] = __float2bfloat16_rn(v);
}

template <int NB, int EPI>
static int launch_hft(const CUtensorMap& a, const CUtensorMap& b, HftTcParams& p, cudaStream_t st, const char* what) {
    static SmemOptIn configured;
    const int kMax = 227 * 1024;
    if (!configured.ensure(hft_tc_kernel<NB, EPI>, kMax)) {
        set_error("%s: cannot raise dynamic shared memory", what);
        return EEL_ERR_CUDA;
    }
    const int b_bytes = ((p.kchunks * p.nblocks * NB * 128) + 1023) & ~1023;
    int ns = (kMax - 2048 - 1024 - b_bytes) / kAStage;
    if (ns > 8) ns = 8;
    if (ns < 2) { set_error("%s: resident matrix leaves no room for the operand ring", what); return EEL_ERR_INVALID; }
    p.n_stages = ns;
    const int smem = b_bytes + ns * kAStage + 2048 + 1024;
    const int grid = p.items < kNumSMs ? p.items : kNumSMs;
    hft_tc_kernel<NB, EPI><<<grid, kHThreads, smem, st>>>(a, b, p);
    return check_launch(what);
}

static int make_a_map(CUtensorMap* m, const void* base, int C, int R, int H, int N, const char* what) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)R, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {1, (uint64_t)C, (uint64_t)R * C, (uint64_t)H * R * C};
    uint32_t box[4] = {64, 64, 1, 1};
    return make_tmap_bf16(m, base, 4, dims, str, box, what);
}
static int make_pix_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, int box_rows, const char* what) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
    uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
    return make_tmap_bf16(m, base, 4, dims, str, box, what);
}
static int make_b_map(CUtensorMap* m, const void* base, int Kp, int rows, int nb, const char* what) {
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)rows};
    uint64_t str[2] = {1, (uint64_t)Kp};
    uint32_t box[2] = {64, (uint32_t)nb};
    return make_tmap_bf16(m, base, 2, dims, str, box, what);
}

bool hft_tc_supported(int H, int W, int C, int r) {
    return (C == 64 || C == 128) && r == 20 && (W == 128 || W == 256 || W == 512) && (H == 128 || H == 256 || H == 512);
}
// forward only: 1024-wide planes too (step 1's [80][W] matrix still fits; step 2 splits K; steps 3 / 4 already run in passes
// of four resident tiles).  The backward's step 1 reduces over 2W and would need the same split.
bool hft_tc_supported_fwd(int H, int W, int C, int r) {
    auto ok = [](int v) { return v == 128 || v == 256 || v == 512 || v == 1024; };
    return (C == 64 || C == 128) && r == 20 && ok(W) && ok(H);
}

// the 16-bit phase code + pixel-row kernels: the training shapes
bool hft_tc_code_path(int H, int W, int C, int r) { return hft_tc_supported(H, W, C, r) && W <= 256; }

size_t hft_tc_matrix_elems(int W) { return (size_t)80 * 2 * W + (size_t)2 * W * 128; }

// step 1: T[n][h][2F][c] (fp32) = Bmat . rows, rows = x (R = W, kind 0) or g pairs (R = 2W, kind 1)
int hft_tc_step1(const bf16* rows_in, int R, bf16* mat_ws, int kind, float* T, bf16* Tb, int N, int H, int W, int C, int r, cudaStream_t st) {
    const int F = 2 * r;
    const int Kp = (R + 63) / 64 * 64;
    hft_tc_matrix_kernel<<<cdiv(2 * F * Kp, 256), 256, 0, st>>>(mat_ws, kind, 2 * F, Kp, F, r, W);
    if (int rc = check_launch("hft_tc.matrix1")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_a_map(&tmA, rows_in, C, R, H, N, "hft_tc.step1(A)")) return rc;
    if (int rc = make_b_map(&tmB, mat_ws, Kp, 2 * F, 2 * F, "hft_tc.step1(B)")) return rc;
    HftTcParams p{};
    p.rows_per_item = 128 / C;
    p.items = N * H / p.rows_per_item;
    p.H = H; p.C = C; p.W = W;
    p.kchunks = Kp / 64; p.k16_last = 4;
    p.nblocks = 1; p.ncols = 2 * F; p.T = T; p.Tb = Tb; p.Cst = C; p.mode = 0;
    return launch_hft<80, HEPI_T>(tmA, tmB, p, st, "hft_tc.step1");
}

// step 2: T2b[n][(ro,g)][(f,c)] = sum_(ri,h) A2[(ro,g)][(ri,h)] * T1b[n][h][ri*F + f][c]      (bf16 in, bf16 out)
int hft_tc_step2(const bf16* T1b, bf16* mat_ws, bf16* T2b, float* T2f, int N, int H, int C, int r, cudaStream_t st) {
    const int F = 2 * r;
    const long long FC = (long long)F * C;
    hft_tc_matrix_kernel<<<cdiv(2 * F * 2 * H, 256), 256, 0, st>>>(mat_ws, 5, 2 * F, 2 * H, F, r, H);
    if (int rc = check_launch("hft_tc.matrix2")) return rc;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)FC, (uint64_t)H, 2, (uint64_t)N};
        uint64_t str[4] = {1, (uint64_t)(2 * FC), (uint64_t)FC, (uint64_t)H * 2 * FC};
        uint32_t box[4] = {64, 64, 1, 1};
        if (int rc = make_tmap_bf16(&tmA, T1b, 4, dims, str, box, "hft_tc.step2(A)")) return rc;
    }
    if (int rc = make_b_map(&tmB, mat_ws, 2 * H, 2 * F, 2 * F, "hft_tc.step2(B)")) return rc;
    HftTcParams p{};
    p.rows_per_item = 1;
    p.mode = 1; p.mtiles = (int)(FC / 128); p.hc = H / 64;
    p.items = N * p.mtiles;
    p.H = H; p.C = C;
    p.k16_last = 4;
    p.nblocks = 1; p.ncols = 2 * F; p.Cst = FC;
    if (H <= 512) {
        p.kchunks = 2 * H / 64; p.Tb = T2b;
        return launch_hft<80, HEPI_T>(tmA, tmB, p, st, "hft_tc.step2");
    }
    // K = 2H = (ri, h) no longer fits next to the operand ring (80 x 2H x 2 B resident): split over ri -- the first launch
    // leaves fp32 partial sums, the second adds its half and stores bf16
    p.kchunks = H / 64;
    p.kc_begin = 0; p.accumulate = 0; p.T = T2f; p.Tb = nullptr;
    if (int rc = launch_hft<80, HEPI_T>(tmA, tmB, p, st, "hft_tc.step2(ri=0)")) return rc;
    p.kc_begin = H / 64; p.accumulate = 1; p.T = T2f; p.Tb = T2b;
    return launch_hft<80, HEPI_T>(tmA, tmB, p, st, "hft_tc.step2(ri=1)");
}

// step 3: T3b[n][h][ro*F + f][c] = sum_(ri,g) A3[(ro,h)][(ri,g)] * T2b[n][(ri,g)][(f,c)]
int hft_tc_step3(const bf16* T2b, bf16* mat_ws, bf16* T3b, int N, int H, int C, int r, cudaStream_t st) {
    const int F = 2 * r;
    const long long FC = (long long)F * C;
    hft_tc_matrix_kernel<<<cdiv(2 * H * 128, 256), 256, 0, st>>>(mat_ws, 4, 2 * H, 128, F, r, H);
    if (int rc = check_launch("hft_tc.matrix3")) return rc;
    CUtensorMap tmM, tmT;
    if (int rc = make_b_map(&tmM, mat_ws, 128, 2 * H, 128, "hft_tc.step3(M)")) return rc;
    {
        uint64_t dims[4] = {(uint64_t)FC, (uint64_t)(2 * F), 1, (uint64_t)N};
        uint64_t str[4] = {1, (uint64_t)FC, (uint64_t)(2 * F) * FC, (uint64_t)(2 * F) * FC};
        uint32_t box[4] = {64, 64, 1, 1};
        if (int rc = make_tmap_bf16(&tmT, T2b, 4, dims, str, box, "hft_tc.step3(T)")) return rc;
    }
    Hft4Params p{};
    p.ntiles = (int)(FC / 128);
    p.items = N * p.ntiles;
    p.H = H; p.W = 0; p.C = 128;
    p.mtiles = 2 * H / 128;
    p.y = T3b; p.F = F; p.Cc = C;
    return launch_hft4<128, H4_T3>(tmM, tmT, p, st, "hft_tc.step3");
}

// step 4: low = Bmat . T3b[n,h] (K = 2F);  fwd: y = |x - low|, phase;  bwd: dx = g_re - Re(low)
int hft_tc_step4(const bf16* T3b, bf16* mat_ws, bool fwd, const bf16* x_or_g, bf16* y_or_dx, bf16* phase, int N, int H, int W, int C,
                 int r, cudaStream_t st, bool g_re_only) {
    const int F = 2 * r;
    const int rows = fwd ? 2 * W : W;
    hft_tc_matrix_kernel<<<cdiv(rows * 128, 256), 256, 0, st>>>(mat_ws, fwd ? 2 : 3, rows, 128, F, r, W);
    if (int rc = check_launch("hft_tc.matrix4")) return rc;
    CUtensorMap tmM, tmT;
    if (int rc = make_b_map(&tmM, mat_ws, 128, rows, 128, "hft_tc.step4(M)")) return rc;
    if (int rc = make_a_map(&tmT, T3b, C, 2 * F, H, N, "hft_tc.step4(T)")) return rc;
    Hft4Params p{};
    p.items = N * H;
    p.H = H; p.W = W; p.C = C;
    p.mtiles = rows / 128;
    p.x = x_or_g; p.y = y_or_dx; p.phase = phase;
    p.xmul = g_re_only ? 1 : 2;
    if (C == 64) return fwd ? launch_hft4<64, H4_FWD>(tmM, tmT, p, st, "hft_tc.step4_fwd") : launch_hft4<64, H4_BWD>(tmM, tmT, p, st, "hft_tc.step4_bwd");
    return fwd ? launch_hft4<128, H4_FWD>(tmM, tmT, p, st, "hft_tc.step4_fwd") : launch_hft4<128, H4_BWD>(tmM, tmT, p, st, "hft_tc.step4_bwd");
}

// backward step 1 from (dy, phase code): T1b[n][h][2F][c] (bf16)
int hft_tc_step1g(const bf16* dy, const uint16_t* code, bf16* mat_ws, bf16* T1b, bf16* gre, int N, int H, int W, int C, int r, cudaStream_t st) {
    const int F = 2 * r;
    const int Kp = 2 * W;
    hft_tc_matrix_kernel<<<cdiv(2 * F * Kp, 256), 256, 0, st>>>(mat_ws, 1, 2 * F, Kp, F, r, W);
    if (int rc = check_launch("hft_tc.matrix1g")) return rc;
    CUtensorMap tmDy, tmCode, tmB;
    if (int rc = make_pix_map(&tmDy, dy, C, W, H, N, 32, "hft_tc.step1g(dy)")) return rc;
    if (int rc = make_pix_map(&tmCode, code, C, W, H, N, 32, "hft_tc.step1g(code)")) return rc;
    if (int rc = make_b_map(&tmB, mat_ws, Kp, 2 * F, 2 * F, "hft_tc.step1g(B)")) return rc;
    HftTcParams p{};
    p.rows_per_item = 128 / C;
    p.items = N * H / p.rows_per_item;
    p.H = H; p.C = C; p.W = W;
    p.kchunks = Kp / 64; p.k16_last = 4;
    p.nblocks = 1; p.ncols = 2 * F; p.Tb = T1b; p.Cst = C; p.mode = 0; p.gre = gre;
    const int kMax = 227 * 1024;
    const int b_bytes = ((p.kchunks * 80 * 128) + 1023) & ~1023;
    int ns = (kMax - 2048 - 1024 - b_bytes) / kG1Stage;
    if (ns > 6) ns = 6;
    if (ns < 2) { set_error("hft_tc.step1g: resident matrix leaves no room for the operand ring"); return EEL_ERR_INVALID; }
    p.n_stages = ns;
    const int smem = b_bytes + ns * kG1Stage + 2048 + 1024;
    const int grid = p.items < kNumSMs ? p.items : kNumSMs;
    if (C == 64) {
        static SmemOptIn configured;
        if (!configured.ensure(hft_tc1g_kernel<64>, kMax)) { set_error("hft_tc.step1g: cannot raise dynamic shared memory"); return EEL_ERR_CUDA; }
        hft_tc1g_kernel<64><<<grid, kG1Threads, smem, st>>>(tmDy, tmCode, tmB, p);
    } else {
        static SmemOptIn configured;
        if (!configured.ensure(hft_tc1g_kernel<128>, kMax)) { set_error("hft_tc.step1g: cannot raise dynamic shared memory"); return EEL_ERR_CUDA; }
        hft_tc1g_kernel<128><<<grid, kG1Threads, smem, st>>>(tmDy, tmCode, tmB, p);
    }
    return check_launch("hft_tc.step1g");
}

// last step with pixel rows: fwd: y = |x - low|, code = phase(x - low);  bwd: dx = dy * phase_re - Re(low)
int hft_tc_step4p(const bf16* T3b, bf16* mat_ws, bool fwd, const bf16* x_or_dy, const uint16_t* code_in, bf16* y_or_dx,
                  uint16_t* code_out, int N, int H, int W, int C, int r, cudaStream_t st) {
    const int F = 2 * r;
    const int rows = fwd ? 2 * W : W;
    hft_tc_matrix_kernel<<<cdiv(rows * 128, 256), 256, 0, st>>>(mat_ws, fwd ? 6 : 3, rows, 128, F, r, W);
    if (int rc = check_launch("hft_tc.matrix4p")) return rc;
    CUtensorMap tmM, tmT;
    if (int rc = make_b_map(&tmM, mat_ws, 128, rows, 128, "hft_tc.step4p(M)")) return rc;
    if (int rc = make_a_map(&tmT, T3b, C, 2 * F, H, N, "hft_tc.step4p(T)")) return rc;
    Hft4pParams p{};
    p.items = N * H;
    p.H = H; p.W = W;
    p.x = x_or_dy; p.code_in = code_in; p.y = y_or_dx; p.code_out = code_out;
    const int pt = W / 128;
    if (C == 64) return fwd ? launch_hft4p<64, true>(tmM, tmT, p, pt, st, "hft_tc.step4p_fwd") : launch_hft4p<64, false>(tmM, tmT, p, pt, st, "hft_tc.step4p_bwd");
    return fwd ? launch_hft4p<128, true>(tmM, tmT, p, pt, st, "hft_tc.step4p_fwd") : launch_hft4p<128, false>(tmM, tmT, p, pt, st, "hft_tc.step4p_bwd");
}

}  // namespace tc
}  // namespace eel
