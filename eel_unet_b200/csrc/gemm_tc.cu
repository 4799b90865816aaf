// bf16 tensor-core GEMM kernels for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma with fp32
// accumulators in TMEM -> tcgen05.ld epilogue.  One persistent, warp-specialised kernel template:
//
//   warp 0   TMA producer (one elected thread)          warp 1   TMEM allocator + MMA issuer (one elected thread)
//   warps 2-9  epilogue: TMEM -> registers -> (+bias, ReLU) -> bf16 -> transposed through smem -> global
//
//   C[M, N] = A[M, K] . B[N, K]^T     (1x1 conv / Linear; ConvTranspose2d(k2,s2) forward with a scatter epilogue and
//                                      its data gradient with a gathered A operand)
// A streams through a ring of 128 x 64 tiles.  B (the weights) either streams through its own ring or -- when the whole
// [N][K] block of the single N tile fits beside the A ring (<= 128 KB: every CAPMLP layer of the 256-channel stages) --
// is loaded ONCE per CTA and stays resident, so the steady state moves activations only.
// Two TMEM accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1.
// (The 3x3 convolution lives in conv_tc.cu.)
#include "tc_common.cuh"

#include <dlfcn.h>
#include <mutex>

namespace eel {
namespace tc {

// ------------------------------------------------------------------------------------------------ host helpers
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, const char* what) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) {
        set_error("%s: cuTensorMapEncodeTiled unavailable (driver too old?)", what);
        return EEL_ERR_CUDA;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * 2;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
        return EEL_ERR_CUDA;
    }
    return EEL_OK;
}

// ------------------------------------------------------------------------------------------------ kernel
enum { EPI_DENSE = 0, EPI_SCATTER = 1 /* dense + adjoint ShiftedChannel row scatter */, EPI_CONVT = 2 };

struct TcParams {
    int kchunks;            // 64-wide K chunks
    int m_tiles, n_tiles;
    long long M;            // GEMM rows
    int Ntot;               // total output columns
    int W;                  // convT: input width
    int relu;
    int Co;                 // convT: output channels
    int na, nb;             // ring depths (nb unused when res)
    int res;                // weights resident (n_tiles == 1)
    const float* bias;
    bf16* out;
    int shH, shW;           // dense: > 0 scatters output rows through the adjoint of ShiftedChannel (rows = pixels of [*, shH, shW])
    float* bn_sums;         // optional [2][Ntot]: per-channel sum / sum of squares of the stored output (grid % n_tiles == 0)
};

constexpr int kMaxStages = 8;
constexpr int kABytes = 16384;                    // 128 rows x 64 bf16
// EW epilogue warps (8 or 16): these kernels' wide-output launches (64 -> 256 linears, ConvTranspose) are bound by the LATENCY of
// the epilogue chain (TMEM read -> convert -> transposition -> store), not by its instruction count or by HBM: with 8 warps the
// four schedulers sit idle three cycles out of four.  Sixteen warps (four per TMEM lane quarter, a quarter of the tile's columns
// each) double the chains in flight; BN = 64 tiles keep 8.
constexpr int kSmemBudget8 = 232448 - 1024 /* alignment */ - 1024 /* barriers */ - 16384 /* epilogue staging */;
constexpr int kSmemBudget16 = kSmemBudget8 - 16384;

template <int BN, int EPI, int AGATHER, bool STATS, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    constexpr int B_BYTES = BN * 128;
    constexpr int TMEM_COLS = 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + p.na * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (p.res ? p.kchunks : p.nb) * B_BYTES);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + kMaxStages;
    uint64_t* fullB = emptyA + kMaxStages;      // res: fullB[0] = "weights resident"
    uint64_t* emptyB = fullB + kMaxStages;
    uint64_t* tmemFull = emptyB + kMaxStages;
    uint64_t* tmemEmpty = tmemFull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmemEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1);
            mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmemFull[i], 1); mbar_init(&tmemEmpty[i], EW); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        if (p.res) {
            mbar_expect_tx(&fullB[0], (uint32_t)(p.kchunks * B_BYTES));
            for (int c = 0; c < p.kchunks; ++c) tma_load_2d(sB + c * B_BYTES, &tmB, &fullB[0], c * 64, 0);
        }
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int n0 = nt * BN;
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&emptyA[sa], pa ^ 1);
                mbar_expect_tx(&fullA[sa], kABytes);
                if (AGATHER) {
                    // ConvTranspose data gradient: A(m, k) = g[n, 2y+dy, 2x+dx, co], k = (dy, dx, co); the tensor map
                    // views g as {2*Co, w, 2 (dy), N*h}; a tile is gw columns x 128/gw (n,y) rows
                    const int per_dy = p.kchunks >> 1;
                    const int dy = c / per_dy, kc = (c - dy * per_dy) * 64;
                    const long long m0 = (long long)mt * 128;
                    tma_load_4d(sA + sa * kABytes, &tmA, &fullA[sa], kc, (int)(m0 % p.W), dy, (int)(m0 / p.W));
                } else tma_load_2d(sA + sa * kABytes, &tmA, &fullA[sa], c * 64, mt * 128);
                if (++sa == p.na) { sa = 0; pa ^= 1; }
                if (!p.res) {
                    mbar_wait(&emptyB[sb], pb ^ 1);
                    mbar_expect_tx(&fullB[sb], B_BYTES);
                    tma_load_2d(sB + sb * B_BYTES, &tmB, &fullB[sb], c * 64, n0);
                    if (++sb == p.nb) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        // descriptors built once; a view = descriptor + (byte offset >> 4) in the start-address field
        const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), 16, 1024, false);
        const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 16, 1024, false);
        if (p.res) {
            mbar_wait(&fullB[0], 0);
            tc_fence_after();
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            mbar_wait(&tmemEmpty[acc], acc_par ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&fullA[sa], pa);
                uint64_t b_view;
                if (p.res) b_view = b_desc0 + (uint32_t)((c * B_BYTES) >> 4);
                else {
                    mbar_wait(&fullB[sb], pb);
                    b_view = b_desc0 + (uint32_t)((sb * B_BYTES) >> 4);
                }
                tc_fence_after();
                const uint64_t a_view = a_desc0 + (uint32_t)((sa * kABytes) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_d, a_view + (uint32_t)(k * 2), b_view + (uint32_t)(k * 2), idesc, (uint32_t)((c | k) != 0));
                if (!p.res) {
                    umma_commit(&emptyB[sb]);
                    if (++sb == p.nb) { sb = 0; pb ^= 1; }
                }
                umma_commit(&emptyA[sa]);
                if (++sa == p.na) { sa = 0; pa ^= 1; }
            }
            umma_commit(&tmemFull[acc]);
        }
    } else if (warp >= 2) {
        // ===================================================================== epilogue (EW warps)
        // EW / 4 warps share each TMEM lane quarter and split the tile's columns; the TMEM read of chunk i+1 is in flight
        // while chunk i is converted, transposed (epi_store_chunk) and stored.
        constexpr int PARTS = EW / 4;           // column parts of a tile
        constexpr int PW = BN / PARTS;          // columns per warp
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;       // which part of the columns
        const EpiLane L = epi_lane(reinterpret_cast<uint8_t*>(tmem_ptr) + 64 + (warp - 2) * 2048, lane);
        constexpr int NCH = PW / 32;            // 32-column chunks per warp and tile
        float st[STATS ? NCH : 1][2];           // fused BatchNorm statistics (tc_common.cuh); STATS is a template flag so
#pragma unroll                                  // that the common no-statistics launches carry none of its registers / code
        for (int ci = 0; ci < (STATS ? NCH : 1); ++ci) st[ci][0] = st[ci][1] = 0.f;
        // Addressing in 32-bit units of one 16-byte vector (8 bf16): ncu showed the epilogue -- not HBM -- bounds the output-heavy
        // launches, with half of its instructions 64-bit address arithmetic and per-row integer divisions.  Everything that
        // depends only on the column chunk is computed when the N tile changes (once per CTA when the grid is a multiple of
        // n_tiles); per tile ONE 32-bit division locates the tile's first row, the rows follow by add / wrap.
        const uint32_t ntot_v = (uint32_t)p.Ntot >> 3, co_v = (uint32_t)p.Co >> 3;
        uint32_t coff_v[NCH];                   // column part of the destination offset of chunk ci
        const float* bp[NCH];                   // its 32 bias values (or null)
        int sdh[NCH], sdw[NCH];                 // EPI_SCATTER: the ShiftedChannel adjoint of the chunk's channel quarter
        int nt_cur = -1;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            if (nt != nt_cur) {
                nt_cur = nt;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int gcol = nt * BN + half * PW + ci * 32;
                    sdh[ci] = sdw[ci] = 0;
                    if (EPI == EPI_CONVT) {
                        const int dy = gcol / (2 * p.Co), rem = gcol - dy * 2 * p.Co;
                        coff_v[ci] = (uint32_t)dy * (2u * (uint32_t)p.W) * co_v + ((uint32_t)rem >> 3);
                        bp[ci] = p.bias ? p.bias + (rem % p.Co) : nullptr;
                    } else {
                        coff_v[ci] = (uint32_t)gcol >> 3;
                        bp[ci] = p.bias ? p.bias + gcol : nullptr;
                        if (EPI == EPI_SCATTER) {
                            const int qd = gcol / (p.Ntot >> 2);
                            sdh[ci] = qd == 0 ? -1 : (qd == 1 ? 1 : 0);
                            sdw[ci] = qd == 2 ? -1 : 0;
                        }
                    }
                }
            }
            const uint32_t m0 = (uint32_t)mt * 128u;
            const uint32_t rows_left = p.M - (long long)m0 >= 128 ? 128u : (uint32_t)(p.M - (long long)m0);
            uint32_t roff_v[4];          // vector offset of output row (row_lo + 8 i)
            bool rok[4];
            int hrow[4], wcol[4];        // shH > 0: the row's (h, w) inside its image
            {
                // the tile's first row: (image row index, column) for the layouts that need them
                const uint32_t Wd = EPI == EPI_CONVT ? (uint32_t)p.W : (EPI == EPI_SCATTER ? (uint32_t)p.shW : 1u);
                const uint32_t r0 = EPI == EPI_DENSE ? 0u : m0 / Wd, x0 = EPI == EPI_DENSE ? 0u : m0 - r0 * Wd;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t r = (uint32_t)(q * 32 + L.row_lo + 8 * i);
                    rok[i] = r < rows_left;
                    hrow[i] = 0; wcol[i] = 0;
                    if (EPI == EPI_DENSE) roff_v[i] = (m0 + r) * ntot_v;
                    else {
                        uint32_t x = x0 + r, ry = r0;
                        if (x >= Wd) { const uint32_t k = x / Wd; ry += k; x -= k * Wd; }
                        if (EPI == EPI_CONVT) roff_v[i] = ((ry * 2u) * (2u * Wd) + 2u * x) * co_v;      // (2*ny, 2*x, 0)
                        else {
                            roff_v[i] = (m0 + r) * ntot_v;
                            wcol[i] = (int)x;
                            hrow[i] = (int)(ry % (uint32_t)p.shH);
                        }
                    }
                }
            }
            mbar_wait(&tmemFull[acc], acc_par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * PW;
            uint32_t buf[2][32];
            tmem_ld32_async(taddr, buf[0]);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                tmem_ld_wait();
                if (ci + 1 < NCH) tmem_ld32_async(taddr + (ci + 1) * 32, buf[(ci + 1) & 1]);
                bf16* dst[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t o = roff_v[i] + coff_v[ci] + (uint32_t)L.slot;
                    if (EPI == EPI_SCATTER) {
                        int hh = hrow[i] + sdh[ci], ww = wcol[i] + sdw[ci];
                        hh = hh < 0 ? hh + p.shH : (hh >= p.shH ? hh - p.shH : hh);
                        ww = ww < 0 ? ww + p.shW : ww;
                        o += (uint32_t)(((hh - hrow[i]) * p.shW + (ww - wcol[i])) * (int)ntot_v);
                    }
                    dst[i] = rok[i] ? p.out + (size_t)o * 8 : nullptr;
                }
                epi_store_chunk(L, buf[ci & 1], bp[ci], p.relu, dst, STATS ? &st[STATS ? ci : 0] : nullptr, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmemEmpty[acc]);
        }
        // the grid is a multiple of n_tiles (host check), so every tile of this CTA lies in the SAME N tile and the
        // statistics registers belong to fixed columns
        const int nt_own = blockIdx.x % p.n_tiles;
        if (STATS && EPI == EPI_DENSE) {
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) epi_stats_flush(p.bn_sums, p.Ntot, nt_own * BN + half * PW + ci * 32, lane, st[STATS ? ci : 0]);
        }
        if (STATS && EPI == EPI_CONVT) {
            // columns are (dy, dx, co): the four taps of a channel add into the same [2][Co] sums
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) epi_stats_flush(p.bn_sums, p.Co, (nt_own * BN + half * PW + ci * 32) % p.Co, lane, st[STATS ? ci : 0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int EPI, int AGATHER, bool STATS>
static int launch_tc_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, cudaStream_t st, const char* what) {
    constexpr int B_BYTES = BN * 128;
    // Measured on B200 (tools/op_bench.py, batch 64): 16 warps take the bias-only ConvTranspose forwards from 261 / 142 / 85 us to
    // 204 / 111 / 70 us, but inside the model most wide launches also carry the fused BatchNorm statistics, whose registers no longer
    // fit the 113-register budget of 576 threads (spills): tc_linear 1.46 -> 1.64 ms, ConvTranspose forward 0.86 -> 1.0 ms per step.
    // Until the statistics epilogue is slimmed down, 8 warps stay the default.
    constexpr int EW = 8;
    constexpr int kSmemBudget = EW == 16 ? kSmemBudget16 : kSmemBudget8;
    // resident weights: single N tile whose whole K extent fits beside >= 4 A stages
    p.res = (p.n_tiles == 1 && p.kchunks * B_BYTES + 4 * kABytes <= kSmemBudget) ? 1 : 0;
    int b_bytes;
    if (p.res) {
        b_bytes = p.kchunks * B_BYTES;
        p.na = (kSmemBudget - b_bytes) / kABytes;
        if (p.na > kMaxStages) p.na = kMaxStages;
        p.nb = 0;
    } else {
        p.na = p.nb = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
        b_bytes = p.nb * B_BYTES;
    }
    const int smem = p.na * kABytes + b_bytes + 1024 + 2048 * EW + 1024;
    static SmemOptIn configured;
    if (!configured.ensure(tc_gemm_kernel<BN, EPI, AGATHER, STATS, EW>, smem)) {
        set_error("%s: cannot raise dynamic shared memory to %d", what, smem);
        return EEL_ERR_CUDA;
    }
    int tiles = p.m_tiles * p.n_tiles;
    int grid = tiles < kNumSMs ? tiles : kNumSMs;
    tc_gemm_kernel<BN, EPI, AGATHER, STATS, EW><<<grid, 64 + 32 * EW, smem, st>>>(tmA, tmB, p);
    return check_launch(what);
}

template <int BN, int EPI, int AGATHER>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st, const char* what) {
    if (EPI != EPI_SCATTER && AGATHER == 0 && p.bn_sums != nullptr) return launch_tc_impl<BN, EPI, AGATHER, (EPI != EPI_SCATTER && AGATHER == 0)>(tmA, tmB, p, st, what);
    return launch_tc_impl<BN, EPI, AGATHER, false>(tmA, tmB, p, st, what);
}

template <int EPI, int AGATHER = 0>
static int dispatch_bn(int bn, const CUtensorMap& a, const CUtensorMap& b, const TcParams& p, cudaStream_t st, const char* what) {
    if (bn == 256) return launch_tc<256, EPI, AGATHER>(a, b, p, st, what);
    if (bn == 128) return launch_tc<128, EPI, AGATHER>(a, b, p, st, what);
    return launch_tc<64, EPI, AGATHER>(a, b, p, st, what);
}

// fused BatchNorm statistics keep per-column sums in registers across a CTA's tiles: every tile of a CTA must lie in the
// same N tile, i.e. the grid (min(tiles, SMs)) must be a multiple of n_tiles
static bool stats_grid_ok(const TcParams& p) {
    const int tiles = p.m_tiles * p.n_tiles;
    return (tiles < kNumSMs ? tiles : kNumSMs) % p.n_tiles == 0;
}

static int pick_bn(int ncols) { return ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64); }

}  // namespace tc
}  // namespace eel

using namespace eel;
using namespace eel::tc;

extern "C" {

int eel_tc_linear(const void* x, const void* w, const float* bias, void* y, long long P, int K, int Nout, int relu,
                  float* bn_sums, int scatterH, int scatterW, eel_stream s) {
    EEL_REQUIRE(x && w && y && P > 0, "tc_linear: bad argument");
    EEL_REQUIRE(K % 64 == 0 && Nout % 64 == 0, "tc_linear: K and Nout must be multiples of 64 (got %d, %d)", K, Nout);
    EEL_REQUIRE(P * (long long)Nout / 8 < (1LL << 32), "tc_linear: output too large for 32-bit vector offsets");
    const int bn = pick_bn(Nout);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)P};
        uint64_t str[2] = {1, (uint64_t)K};
        uint32_t box[2] = {64, 128};
        if (int rc = make_tmap_bf16(&tmA, x, 2, dims, str, box, "tc_linear(A)")) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)Nout};
        uint64_t str[2] = {1, (uint64_t)K};
        uint32_t box[2] = {64, (uint32_t)bn};
        if (int rc = make_tmap_bf16(&tmB, w, 2, dims, str, box, "tc_linear(B)")) return rc;
    }
    TcParams p{};
    p.kchunks = K / 64;
    p.m_tiles = cdiv(P, 128);
    p.n_tiles = Nout / bn;
    p.M = P; p.Ntot = Nout;
    p.relu = relu; p.bias = bias; p.out = (bf16*)y;
    if (scatterH > 0 || scatterW > 0) {
        EEL_REQUIRE(scatterH > 0 && scatterW > 0 && P % ((long long)scatterH * scatterW) == 0 && Nout % 128 == 0,
                    "tc_linear: scatter needs P = images * H * W and Nout a multiple of 128 (32-column chunks inside one quarter)");
        p.shH = scatterH; p.shW = scatterW;
    }
    if (bn_sums != nullptr) {
        EEL_REQUIRE(stats_grid_ok(p), "tc_linear: fused BatchNorm statistics need 1, 2 or 4 N tiles (Nout %d)", Nout);
        if (cudaMemsetAsync(bn_sums, 0, sizeof(float) * 2 * Nout, (cudaStream_t)s) != cudaSuccess) {
            set_error("tc_linear: memset failed");
            return EEL_ERR_CUDA;
        }
        p.bn_sums = bn_sums;
    }
    if (p.shH > 0) return dispatch_bn<EPI_SCATTER>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_linear(scatter)");
    return dispatch_bn<EPI_DENSE>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_linear");
}

int eel_tc_convt2x2_fwd(const void* x, const void* wk, const float* bias, void* y, int N, int h, int w, int Cin, int Cout,
                        float* bn_sums, eel_stream s) {
    EEL_REQUIRE(x && wk && y && N > 0 && h > 0 && w > 0, "tc_convt2x2_fwd: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convt2x2_fwd: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    const long long P = (long long)N * h * w;
    const int ncols = 4 * Cout;
    EEL_REQUIRE(P * (long long)ncols / 8 < (1LL << 32), "tc_convt2x2_fwd: output too large for 32-bit vector offsets");
    const int bn = pick_bn(ncols);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)P};
        uint64_t str[2] = {1, (uint64_t)Cin};
        uint32_t box[2] = {64, 128};
        if (int rc = make_tmap_bf16(&tmA, x, 2, dims, str, box, "tc_convt2x2_fwd(A)")) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)ncols};
        uint64_t str[2] = {1, (uint64_t)Cin};
        uint32_t box[2] = {64, (uint32_t)bn};
        if (int rc = make_tmap_bf16(&tmB, wk, 2, dims, str, box, "tc_convt2x2_fwd(B)")) return rc;
    }
    TcParams p{};
    p.kchunks = Cin / 64;
    p.m_tiles = cdiv(P, 128);
    p.n_tiles = ncols / bn;
    p.M = P; p.Ntot = ncols; p.W = w; p.Co = Cout;
    p.bias = bias; p.out = (bf16*)y;
    if (bn_sums != nullptr) {
        EEL_REQUIRE(stats_grid_ok(p), "tc_convt2x2_fwd: fused BatchNorm statistics need 1, 2 or 4 N tiles (got Cout %d)", Cout);
        if (cudaMemsetAsync(bn_sums, 0, sizeof(float) * 2 * Cout, (cudaStream_t)s) != cudaSuccess) {
            set_error("tc_convt2x2_fwd: memset failed");
            return EEL_ERR_CUDA;
        }
        p.bn_sums = bn_sums;
    }
    return dispatch_bn<EPI_CONVT>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_convt2x2_fwd");
}


int eel_tc_convt2x2_dgrad(const void* dy, const void* wp, void* dx, int N, int h, int w, int Cin, int Cout, eel_stream s) {
    EEL_REQUIRE(dy && wp && dx && N > 0 && h > 0 && w > 0, "tc_convt2x2_dgrad: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convt2x2_dgrad: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    const int gw = w >= 128 ? 128 : w;
    EEL_REQUIRE(128 % gw == 0 && w % gw == 0, "tc_convt2x2_dgrad: input width %d must divide or be a multiple of 128", w);
    const long long P = (long long)N * h * w;
    EEL_REQUIRE(P * (long long)Cin / 8 < (1LL << 32), "tc_convt2x2_dgrad: output too large for 32-bit vector offsets");
    const int bn = pick_bn(Cin);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)2 * Cout, (uint64_t)w, 2, (uint64_t)N * h};
        uint64_t str[4] = {1, (uint64_t)2 * Cout, (uint64_t)2 * w * Cout, (uint64_t)4 * w * Cout};
        uint32_t box[4] = {64, (uint32_t)gw, 1, (uint32_t)(128 / gw)};
        if (int rc = make_tmap_bf16(&tmA, dy, 4, dims, str, box, "tc_convt2x2_dgrad(A)")) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)4 * Cout, (uint64_t)Cin};
        uint64_t str[2] = {1, (uint64_t)4 * Cout};
        uint32_t box[2] = {64, (uint32_t)bn};
        if (int rc = make_tmap_bf16(&tmB, wp, 2, dims, str, box, "tc_convt2x2_dgrad(B)")) return rc;
    }
    TcParams p{};
    p.kchunks = 4 * Cout / 64;
    p.m_tiles = cdiv(P, 128);
    p.n_tiles = Cin / bn;
    p.M = P; p.Ntot = Cin; p.W = w;
    p.out = (bf16*)dx;
    return dispatch_bn<EPI_DENSE, 1>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_convt2x2_dgrad");
}

}  // extern "C"
