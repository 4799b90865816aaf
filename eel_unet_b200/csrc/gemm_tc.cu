// bf16 tensor-core GEMM-class kernels for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma with
// fp32 accumulators in TMEM -> tcgen05.ld epilogue.  One persistent, warp-specialised kernel template:
//
//   warp 0   TMA producer (one elected thread)          warp 1   TMEM allocator + MMA issuer (one thread)
//   warps 2-5  epilogue: TMEM -> registers -> (+bias, ReLU) -> bf16 -> global
//
// TAPS = 1 : plain GEMM   C[M, N] = A[M, K] . B[N, K]^T          (1x1 conv / Linear / ConvTranspose2d as GEMM)
// TAPS = 9 : implicit-GEMM 3x3 convolution over NHWC.  An output tile is 16 x 8 pixels (= 128 GEMM rows).  Per
//            64-channel K chunk THREE 18 x 8 row-halo tiles of the input (column offsets -1, 0, +1) are brought
//            into shared memory by 4-D TMA box loads (out-of-bounds = zero = the conv padding) and the nine taps
//            are nine *views*: tap (dy, dx) is copy dx+1 with the A descriptor's start address advanced by
//            (1+dy) image rows = (1+dy) * 1024 B, so every view starts on a swizzle-pattern boundary.
//            Activations cross L2 -> SM 3.4x instead of 9x.  Weights stream per (chunk, tap) through their own
//            ring: B tile = [BN output channels][64 input channels] of tap t, K-major.
// Two rings (A, B) with full/empty mbarriers; two TMEM accumulator stages so the epilogue of tile i overlaps
// the MMAs of tile i+1.
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
enum { EPI_DENSE = 0, EPI_CONV = 1, EPI_CONVT = 2 };

struct TcParams {
    int kchunks;            // 64-wide K chunks (per tap)
    int m_tiles, n_tiles;
    long long M;            // GEMM rows (dense / convT)
    int Ntot;               // total output columns
    int N, H, W;            // conv: image batch / height / width; convT: input w in W
    int tiles_h, tiles_w;
    int flip;               // conv: mirror the tap offsets (data gradient)
    int relu;
    int Co;                 // convT: output channels
    const float* bias;
    bf16* out;
};

constexpr int kThreads = 192;
constexpr int kHaloH = 18;                       // 16 + 2 rows, 8 columns per copy
constexpr int kCopyBytes = 8 * kHaloH * 128;      // 18432 per column-shifted copy
constexpr int kHaloBytes = 3 * kCopyBytes;        // 55296

template <int BN, int TAPS> struct TcCfg {
    static constexpr int A_BYTES = TAPS == 9 ? kHaloBytes : 16384;   // 1024-aligned stage sizes
    static constexpr int A_TX = TAPS == 9 ? kHaloBytes : 16384;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int NA = TAPS == 9 ? 2 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
    static constexpr int NB = TAPS == 9 ? (BN == 256 ? 3 : (BN == 128 ? 6 : 8)) : NA;
    static constexpr int SMEM = NA * A_BYTES + NB * B_BYTES + 1024 /* barriers */ + 1024 /* alignment slack */;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // 128 / 256 / 512: powers of two
};

template <int BN, int TAPS, int EPI, int AGATHER>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    typedef TcCfg<BN, TAPS> Cfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::NA * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::NB * Cfg::B_BYTES);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + Cfg::NA;
    uint64_t* fullB = emptyA + Cfg::NA;
    uint64_t* emptyB = fullB + Cfg::NB;
    uint64_t* tmemFull = emptyB + Cfg::NB;
    uint64_t* tmemEmpty = tmemFull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmemEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < Cfg::NB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmemFull[i], 1); mbar_init(&tmemEmpty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int n0 = nt * BN;
            int c_w = 0, c_h = 0, c_n = 0;
            if (TAPS == 9) {
                const int tw = mt % p.tiles_w, r = mt / p.tiles_w;
                const int th = r % p.tiles_h;
                c_n = r / p.tiles_h;
                c_h = th * 16 - 1;
                c_w = tw * 8;
            }
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&emptyA[sa], pa ^ 1);
                mbar_expect_tx(&fullA[sa], Cfg::A_TX);
                if (TAPS == 9) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        tma_load_4d(sA + sa * Cfg::A_BYTES + j * kCopyBytes, &tmA, &fullA[sa], c * 64, c_w + j - 1, c_h, c_n);
                } else if (AGATHER) {
                    // ConvTranspose data gradient: A(m, k) = g[n, 2y+dy, 2x+dx, co], k = (dy, dx, co); the tensor map
                    // views g as {2*Co, w, 2 (dy), N*h}; a tile is p.gw columns x 128/p.gw (n,y) rows
                    const int per_dy = p.kchunks >> 1;
                    const int dy = c / per_dy, kc = (c - dy * per_dy) * 64;
                    const long long m0 = (long long)mt * 128;
                    tma_load_4d(sA + sa * Cfg::A_BYTES, &tmA, &fullA[sa], kc, (int)(m0 % p.W), dy, (int)(m0 / p.W));
                } else tma_load_2d(sA + sa * Cfg::A_BYTES, &tmA, &fullA[sa], c * 64, mt * 128);
                if (++sa == Cfg::NA) { sa = 0; pa ^= 1; }
                for (int t = 0; t < TAPS; ++t) {
                    mbar_wait(&emptyB[sb], pb ^ 1);
                    mbar_expect_tx(&fullB[sb], Cfg::B_BYTES);
                    if (TAPS == 9) tma_load_3d(sB + sb * Cfg::B_BYTES, &tmB, &fullB[sb], c * 64, n0, t);
                    else tma_load_2d(sB + sb * Cfg::B_BYTES, &tmB, &fullB[sb], c * 64, n0);
                    if (++sb == Cfg::NB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            mbar_wait(&tmemEmpty[acc], acc_par ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            uint32_t accum = 0;
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(&fullA[sa], pa);
                const uint32_t a_base = smem_u32(sA + sa * Cfg::A_BYTES);
                for (int t = 0; t < TAPS; ++t) {
                    mbar_wait(&fullB[sb], pb);
                    tc_fence_after();
                    uint32_t a_view = a_base;
                    if (TAPS == 9) {
                        int dy = t / 3 - 1, dx = t % 3 - 1;
                        if (p.flip) { dy = -dy; dx = -dx; }
                        a_view += (dx + 1) * kCopyBytes + (1 + dy) * 1024;
                    }
                    const uint32_t b_base = smem_u32(sB + sb * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_smem_desc(a_view + k * 32, 16, 1024, false);
                        const uint64_t db = make_smem_desc(b_base + k * 32, 16, 1024, false);
                        umma_bf16(tmem_d, da, db, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&emptyB[sb]);
                    if (++sb == Cfg::NB) { sb = 0; pb ^= 1; }
                }
                umma_commit(&emptyA[sa]);
                if (++sa == Cfg::NA) { sa = 0; pa ^= 1; }
            }
            umma_commit(&tmemFull[acc]);
        }
    } else if (warp >= 2) {
        // ===================================================================== epilogue
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;            // accumulator row = GEMM row within the tile
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_par = (it >> 1) & 1;
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int n0 = nt * BN;
            bool valid;
            long long row_off = 0;       // element offset of this row's output (dense / conv)
            long long ct_base = 0;       // convT: offset of (2*ny, 2*x, 0)
            if (EPI == EPI_CONV) {
                const int tw = mt % p.tiles_w, rr = mt / p.tiles_w;
                const int th = rr % p.tiles_h, n = rr / p.tiles_h;
                const int h = th * 16 + (r >> 3), w = tw * 8 + (r & 7);
                valid = h < p.H && w < p.W;
                row_off = (((long long)n * p.H + h) * p.W + w) * p.Ntot;
            } else {
                const long long m = (long long)mt * 128 + r;
                valid = m < p.M;
                if (EPI == EPI_DENSE) row_off = m * p.Ntot;
                else {
                    const long long ny = m / p.W;
                    const int x = (int)(m - ny * p.W);
                    ct_base = ((ny * 2) * (2LL * p.W) + 2 * x) * p.Co;
                }
            }
            mbar_wait(&tmemFull[acc], acc_par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 32) {
                float v[32];
                tmem_ld32(taddr + cc, v);
                if (valid) {
                    const int gcol = n0 + cc;
                    bf16* dst;
                    const float* bp = nullptr;
                    if (EPI == EPI_CONVT) {
                        const int dy = gcol / (2 * p.Co), rem = gcol - dy * 2 * p.Co;
                        dst = p.out + ct_base + (long long)dy * (2LL * p.W) * p.Co + rem;
                        if (p.bias) bp = p.bias + (rem % p.Co);
                    } else {
                        dst = p.out + row_off + gcol;
                        if (p.bias) bp = p.bias + gcol;
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float a = v[j] + (bp ? __ldg(bp + j) : 0.f);
                        float b = v[j + 1] + (bp ? __ldg(bp + j + 1) : 0.f);
                        if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmemEmpty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int TAPS, int EPI, int AGATHER>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st, const char* what) {
    typedef TcCfg<BN, TAPS> Cfg;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_kernel<BN, TAPS, EPI, AGATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
            set_error("%s: cannot raise dynamic shared memory to %d", what, Cfg::SMEM);
            return EEL_ERR_CUDA;
        }
        configured = true;
    }
    int tiles = p.m_tiles * p.n_tiles;
    int grid = tiles < kNumSMs ? tiles : kNumSMs;
    tc_gemm_kernel<BN, TAPS, EPI, AGATHER><<<grid, kThreads, Cfg::SMEM, st>>>(tmA, tmB, p);
    return check_launch(what);
}

template <int TAPS, int EPI, int AGATHER = 0>
static int dispatch_bn(int bn, const CUtensorMap& a, const CUtensorMap& b, const TcParams& p, cudaStream_t st, const char* what) {
    if (bn == 256) return launch_tc<256, TAPS, EPI, AGATHER>(a, b, p, st, what);
    if (bn == 128) return launch_tc<128, TAPS, EPI, AGATHER>(a, b, p, st, what);
    return launch_tc<64, TAPS, EPI, AGATHER>(a, b, p, st, what);
}

static int pick_bn(int ncols) { return ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64); }

}  // namespace tc
}  // namespace eel

using namespace eel;
using namespace eel::tc;

extern "C" {

int eel_tc_conv3x3_v1(const void* x, const void* wk, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                      int relu, int flip, eel_stream s) {
    EEL_REQUIRE(x && wk && y && N > 0 && H > 0 && W > 0, "tc_conv3x3: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_conv3x3: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    const int bn = pick_bn(Cout);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
        uint32_t box[4] = {64, 8, (uint32_t)kHaloH, 1};
        if (int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, "tc_conv3x3(A)")) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
        uint64_t str[3] = {1, (uint64_t)Cin, (uint64_t)Cin * Cout};
        uint32_t box[3] = {64, (uint32_t)bn, 1};
        if (int rc = make_tmap_bf16(&tmB, wk, 3, dims, str, box, "tc_conv3x3(B)")) return rc;
    }
    TcParams p{};
    p.kchunks = Cin / 64;
    p.tiles_h = cdiv(H, 16);
    p.tiles_w = cdiv(W, 8);
    p.m_tiles = N * p.tiles_h * p.tiles_w;
    p.n_tiles = Cout / bn;
    p.Ntot = Cout;
    p.N = N; p.H = H; p.W = W;
    p.flip = flip; p.relu = relu;
    p.bias = bias; p.out = (bf16*)y;
    return dispatch_bn<9, EPI_CONV>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_conv3x3");
}

int eel_tc_linear(const void* x, const void* w, const float* bias, void* y, long long P, int K, int Nout, int relu,
                  eel_stream s) {
    EEL_REQUIRE(x && w && y && P > 0, "tc_linear: bad argument");
    EEL_REQUIRE(K % 64 == 0 && Nout % 64 == 0, "tc_linear: K and Nout must be multiples of 64 (got %d, %d)", K, Nout);
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
    return dispatch_bn<1, EPI_DENSE>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_linear");
}

int eel_tc_convt2x2_fwd(const void* x, const void* wk, const float* bias, void* y, int N, int h, int w, int Cin, int Cout,
                        eel_stream s) {
    EEL_REQUIRE(x && wk && y && N > 0 && h > 0 && w > 0, "tc_convt2x2_fwd: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convt2x2_fwd: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    const long long P = (long long)N * h * w;
    const int ncols = 4 * Cout;
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
    return dispatch_bn<1, EPI_CONVT>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_convt2x2_fwd");
}


int eel_tc_convt2x2_dgrad(const void* dy, const void* wp, void* dx, int N, int h, int w, int Cin, int Cout, eel_stream s) {
    EEL_REQUIRE(dy && wp && dx && N > 0 && h > 0 && w > 0, "tc_convt2x2_dgrad: bad argument");
    EEL_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convt2x2_dgrad: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    const int gw = w >= 128 ? 128 : w;
    EEL_REQUIRE(128 % gw == 0 && w % gw == 0, "tc_convt2x2_dgrad: input width %d must divide or be a multiple of 128", w);
    const long long P = (long long)N * h * w;
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
    return dispatch_bn<1, EPI_DENSE, 1>(bn, tmA, tmB, p, (cudaStream_t)s, "tc_convt2x2_dgrad");
}

}  // extern "C"
