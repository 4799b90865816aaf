// bf16 tensor-core weight-gradient kernels (sm_100a): D[m][n] = sum_p A[p][m] * B[p][n], the reduction running
// over PIXELS.  Both operands are "MN-major" for tcgen05 (the pixel index is K; channels are contiguous), staged by
// TMA as [64-channel atom][pixel rows][128 B] slabs with the 128-byte swizzle; fp32 accumulators live in TMEM for
// the whole pixel range a CTA owns and are added to the fp32 gradient with red.global.add at the end (split-K over
// CTAs).
//
//  PLAIN   dW = A^T B for row-major A[P][Ma], B[P][Nb]   (1x1 conv / Linear / ConvTranspose2d weight gradients)
//  CONV_B  3x3 conv, Cin >= 128: a CTA owns (128 input channels) x (BN output channels) x (one column offset dx);
//          per 16x8-pixel patch it loads ONE 18x8 row-halo copy of x (column offset dx) and the dy patch; the three
//          taps dy = -1,0,+1 are views of that copy shifted by whole image rows (1024 B): 3 accumulators.
//  CONV_A  3x3 conv, Cin == 64: M = 128 is filled by stacking TWO taps (2 x 64 input channels: the second atom of the
//          A descriptor is the same slab seen LBO bytes further = one image row lower, or the next column copy);
//          9 taps = 5 accumulators of 64 output channels.
#include "tc_common.cuh"

namespace eel {
namespace tc {

enum { WG_PLAIN = 0, WG_CONV_B = 1, WG_CONV_A = 2 };

struct WgParams {
    int stages_total;        // patches (conv) or 64-row chunks (plain)
    int splits;              // split-K factor
    int m_blocks, n_blocks;  // 128-row blocks of M, BN-col blocks of N
    int N, H, W, tiles_h, tiles_w;   // conv geometry
    int gw;                  // plain + gathered B (ConvTranspose): tile width in input pixels; 0 = dense B
    int gW;                  // ... input width
    int per_dy;              // ... 64-column atoms per dy (= 2*Co/64)
    int Ma, Nb;              // valid extents of D
    long long ldm, ldn;      // plain: element (m, n) lives at out[m*ldm + n*ldn]
    int Cin, Cout;           // conv: out is dwp[3][3][Cin][Cout]
    float* out;
};

constexpr int kWgThreads = 192;
constexpr int kCopy = 18432;   // 18 rows x 8 cols x 128 B
constexpr int kTile = 16384;   // 16 rows x 8 cols x 128 B

template <int MODE, int BN> struct WgCfg {
    static constexpr int A_BYTES = MODE == WG_PLAIN ? 2 * 8192 : (MODE == WG_CONV_B ? 2 * kCopy : 3 * kCopy);
    static constexpr int B_BYTES = MODE == WG_PLAIN ? (BN / 64) * 8192 : (BN / 64) * kTile;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int NS = (220 * 1024) / STAGE > 6 ? 6 : (220 * 1024) / STAGE;
    static constexpr int SMEM = NS * STAGE + 2048;
    static constexpr int GROUPS = MODE == WG_PLAIN ? 1 : (MODE == WG_CONV_B ? 3 : 5);
    static constexpr int COLS = GROUPS * BN;
    static constexpr int TMEM_COLS = COLS <= 32 ? 32 : (COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512)));
    static_assert(COLS <= 512, "accumulators exceed TMEM");
    static_assert(NS >= 2, "need at least two stages");
    static_assert(STAGE % 1024 == 0, "stage must keep 1024-byte alignment");
};

template <int MODE, int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
    typedef WgCfg<MODE, BN> Cfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::NS * Cfg::STAGE);
    uint64_t* empty = full + Cfg::NS;
    uint64_t* accFull = empty + Cfg::NS;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accFull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- which unit of work is this CTA?
    const int split = blockIdx.x % p.splits;
    int u = blockIdx.x / p.splits;
    const int nb = u % p.n_blocks; u /= p.n_blocks;
    const int mb = u % p.m_blocks; u /= p.m_blocks;
    const int grp = u;   // CONV_B: column offset index (dx = grp - 1)
    const int per = (p.stages_total + p.splits - 1) / p.splits;
    const int s_begin = split * per;
    const int s_end = min(p.stages_total, s_begin + per);
    const int nstages = s_end - s_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accFull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (nstages > 0) {
        if (warp == 0 && lane == 0) {
            // ================================================================= TMA producer
            int st = 0;
            uint32_t ph = 0;
            for (int s = s_begin; s < s_end; ++s) {
                mbar_wait(&empty[st], ph ^ 1);
                uint8_t* a = smem + st * Cfg::STAGE;
                uint8_t* b = a + Cfg::A_BYTES;
                mbar_expect_tx(&full[st], Cfg::STAGE);
                if (MODE == WG_PLAIN) {
                    const int r0 = s * 64;
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d(a + j * 8192, &tmA, &full[st], mb * 128 + j * 64, r0);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) {
                        if (p.gw == 0) tma_load_2d(b + j * 8192, &tmB, &full[st], nb * BN + j * 64, r0);
                        else {
                            const int atom = nb * (BN / 64) + j;
                            const int dy = atom / p.per_dy, kc = (atom - dy * p.per_dy) * 64;
                            tma_load_4d(b + j * 8192, &tmB, &full[st], kc, r0 % p.gW, dy, r0 / p.gW);
                        }
                    }
                } else {
                    const int tw = s % p.tiles_w, r = s / p.tiles_w;
                    const int th = r % p.tiles_h, n = r / p.tiles_h;
                    if (MODE == WG_CONV_B) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            tma_load_4d(a + j * kCopy, &tmA, &full[st], mb * 128 + j * 64, tw * 8 + grp - 1, th * 16 - 1, n);
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; ++j) tma_load_4d(a + j * kCopy, &tmA, &full[st], 0, tw * 8 + j - 1, th * 16 - 1, n);
                    }
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) tma_load_4d(b + j * kTile, &tmB, &full[st], nb * BN + j * 64, tw * 8, th * 16, n);
                }
                if (++st == Cfg::NS) { st = 0; ph ^= 1; }
            }
        } else if (warp == 1 && elect_one()) {
            // ================================================================= MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
            int st = 0;
            uint32_t ph = 0;
            // descriptors are built once for stage 0; every operand view is one of them plus a 16-byte-unit offset in
            // the start-address field, so an MMA costs a few issue slots (N = 64 MMAs last only 32 cycles)
            const uint32_t s0 = smem_u32(smem);
            const uint64_t dB = make_smem_desc(s0 + Cfg::A_BYTES, MODE == WG_PLAIN ? 8192 : kTile, 1024, false);
            const uint64_t dA0 = make_smem_desc(s0, MODE == WG_PLAIN ? 8192 : (MODE == WG_CONV_B ? kCopy : 1024), 1024, false);
            const uint64_t dA3 = make_smem_desc(s0, kCopy, 1024, false);   // CONV_A group 3: second atom = next column copy
            const uint64_t dA4 = make_smem_desc(s0, 0, 1024, false);       // CONV_A group 4: second atom = first (unused half)
            for (int s = 0; s < nstages; ++s) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t so = (uint32_t)(st * Cfg::STAGE) >> 4;
                if (MODE == WG_PLAIN) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem_base, dA0 + (so + ks * 128), dB + (so + ks * 128), idesc, (s | ks) != 0);
                } else if (MODE == WG_CONV_B) {
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_bf16(tmem_base + t * BN, dA0 + (so + t * 64 + ks * 128), dB + (so + ks * 128), idesc, (s | ks) != 0);
                } else {
                    // accumulator g: rows 0-63 = tap t1, rows 64-127 = tap t2 (second atom = first + LBO)
                    //   g0..g2: copy g, dy = -1 and 0 (LBO = one image row);  g3: (dy=+1, dx=-1) and (dy=+1, dx=0);
                    //   g4: (dy=+1, dx=+1) twice (upper half unused)
#pragma unroll
                    for (int g = 0; g < 5; ++g) {
                        const uint32_t off = (g < 3 ? g * kCopy : (g == 3 ? 2 * 1024 : 2 * kCopy + 2 * 1024)) >> 4;
                        const uint64_t dA = g < 3 ? dA0 : (g == 3 ? dA3 : dA4);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_bf16(tmem_base + g * BN, dA + (so + off + ks * 128), dB + (so + ks * 128), idesc, (s | ks) != 0);
                    }
                }
                umma_commit(&empty[st]);
                if (++st == Cfg::NS) { st = 0; ph ^= 1; }
            }
            umma_commit(accFull);
        } else if (warp >= 2) {
            // ================================================================= epilogue: TMEM -> red.global.add
            const int q = warp & 3;
            const int r = q * 32 + lane;
            mbar_wait(accFull, 0);
            tc_fence_after();
#pragma unroll 1
            for (int g = 0; g < Cfg::GROUPS; ++g) {
                float* row = nullptr;
                long long cstride = 1;
                int ncols_valid = BN;
                if (MODE == WG_PLAIN) {
                    const int m = mb * 128 + r;
                    if (m < p.Ma && p.out) row = p.out + m * p.ldm + (long long)(nb * BN) * p.ldn;
                    cstride = p.ldn;
                    ncols_valid = min(BN, p.Nb - nb * BN);
                } else if (MODE == WG_CONV_B) {
                    const int tap = g * 3 + grp;   // (dy = g - 1, dx = grp - 1)
                    if (p.out) row = p.out + ((long long)tap * p.Cin + mb * 128 + r) * p.Cout + nb * BN;
                } else {
                    const int half = r >> 6, ci = r & 63;
                    int tap;
                    if (g < 3) tap = half * 3 + g;              // dy = -1 / 0, dx = g - 1
                    else if (g == 3) tap = 6 + half;            // dy = +1, dx = -1 / 0
                    else tap = half == 0 ? 8 : -1;              // dy = +1, dx = +1
                    if (tap >= 0 && p.out) row = p.out + ((long long)tap * p.Cin + ci) * p.Cout + nb * BN;
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + g * BN;
#pragma unroll 1
                for (int cc = 0; cc < BN; cc += 32) {
                    float v[32];
                    tmem_ld32(taddr + cc, v);
                    if (row != nullptr) {
                        if (cstride == 1 && cc + 32 <= ncols_valid) {
                            // contiguous columns: 16-byte vector reductions (4x fewer L2 atomic operations)
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + cc + j), "f"(v[j]), "f"(v[j + 1]),
                                             "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (cc + j < ncols_valid) atomicAdd(row + (long long)(cc + j) * cstride, v[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int MODE, int BN>
static int launch_wg(const CUtensorMap& a, const CUtensorMap& b, WgParams& p, int units, cudaStream_t st, const char* what) {
    typedef WgCfg<MODE, BN> Cfg;
    static SmemOptIn configured;
    if (!configured.ensure(tc_wgrad_kernel<MODE, BN>, Cfg::SMEM)) {
        set_error("%s: cannot raise dynamic shared memory to %d", what, Cfg::SMEM);
        return EEL_ERR_CUDA;
    }
    // split-K factor: minimise waves x stages-per-CTA
    long long best = -1;
    int bestS = 1;
    const int maxS = p.stages_total < kNumSMs ? p.stages_total : kNumSMs;
    for (int S = 1; S <= maxS; ++S) {
        long long waves = ((long long)units * S + kNumSMs - 1) / kNumSMs;
        long long cost = waves * ((p.stages_total + S - 1) / S + 6);   // +6: fixed prologue / epilogue per CTA
        if (best < 0 || cost < best) { best = cost; bestS = S; }
    }
    p.splits = bestS;
    tc_wgrad_kernel<MODE, BN><<<units * bestS, kWgThreads, Cfg::SMEM, st>>>(a, b, p);
    return check_launch(what);
}

}  // namespace tc
}  // namespace eel

using namespace eel;
using namespace eel::tc;

extern "C" {

/* dwp:[3][3][Cin][Cout] fp32 (overwritten).  x:[N,H,W,Cin], dy:[N,H,W,Cout] bf16; Cin == 64 or Cin % 128 == 0; Cout % 64 == 0 */
int eel_tc_conv3x3_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout, eel_stream s) {
    EEL_REQUIRE(x && dy && dwp && N > 0 && H > 0 && W > 0, "tc_conv3x3_wgrad: bad argument");
    EEL_REQUIRE((Cin == 64 || Cin % 128 == 0) && Cout % 64 == 0, "tc_conv3x3_wgrad: unsupported channels (%d, %d)", Cin, Cout);
    cudaStream_t st = (cudaStream_t)s;
    if (cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * (size_t)Cin * Cout, st) != cudaSuccess) {
        set_error("tc_conv3x3_wgrad: memset failed");
        return EEL_ERR_CUDA;
    }
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
        uint32_t box[4] = {64, 8, 18, 1};
        if (int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, "tc_conv3x3_wgrad(x)")) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {1, (uint64_t)Cout, (uint64_t)W * Cout, (uint64_t)H * W * Cout};
        uint32_t box[4] = {64, 8, 16, 1};
        if (int rc = make_tmap_bf16(&tmB, dy, 4, dims, str, box, "tc_conv3x3_wgrad(dy)")) return rc;
    }
    WgParams p{};
    p.N = N; p.H = H; p.W = W;
    p.tiles_h = cdiv(H, 16);
    p.tiles_w = cdiv(W, 8);
    p.stages_total = N * p.tiles_h * p.tiles_w;
    p.Cin = Cin; p.Cout = Cout; p.out = dwp;
    if (Cin == 64) {
        p.m_blocks = 1;
        p.n_blocks = Cout / 64;
        return launch_wg<WG_CONV_A, 64>(tmA, tmB, p, p.n_blocks, st, "tc_conv3x3_wgrad(A)");
    }
    p.m_blocks = Cin / 128;
    if (Cout % 128 == 0) {
        p.n_blocks = Cout / 128;
        return launch_wg<WG_CONV_B, 128>(tmA, tmB, p, 3 * p.m_blocks * p.n_blocks, st, "tc_conv3x3_wgrad(B128)");
    }
    p.n_blocks = Cout / 64;
    return launch_wg<WG_CONV_B, 64>(tmA, tmB, p, 3 * p.m_blocks * p.n_blocks, st, "tc_conv3x3_wgrad(B64)");
}

/* out[m*ldm + n*ldn] = sum_p a[p][m] * b[p][n]  (fp32, overwritten; `out_elems` floats are zeroed first).
 * a:[P][Ma], b:[P][Nb] bf16 row-major, Ma % 128 == 0, Nb % 64 == 0.
 * gather_w > 0: b is the [N,2h,2w,Co] output-side tensor of a ConvTranspose2d(k2,s2) read through its input pixel
 * (row p = input pixel, column = (dy, dx, co), Nb = 4*Co, gather_w = input width w, w | 64 or 64 | w). */
int eel_tc_wgrad(const void* a, const void* b, float* out, long long P, int Ma, int Nb, long long ldm, long long ldn,
                 long long out_elems, int gather_w, eel_stream s) {
    EEL_REQUIRE(a && b && out && P > 0 && out_elems > 0, "tc_wgrad: bad argument");
    EEL_REQUIRE(Ma % 128 == 0 && Nb % 64 == 0, "tc_wgrad: Ma must be a multiple of 128 and Nb of 64 (got %d, %d)", Ma, Nb);
    cudaStream_t st = (cudaStream_t)s;
    if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)out_elems, st) != cudaSuccess) {
        set_error("tc_wgrad: memset failed");
        return EEL_ERR_CUDA;
    }
    const int bn = Nb % 256 == 0 ? 256 : (Nb % 128 == 0 ? 128 : 64);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)Ma, (uint64_t)P};
        uint64_t str[2] = {1, (uint64_t)Ma};
        uint32_t box[2] = {64, 64};
        if (int rc = make_tmap_bf16(&tmA, a, 2, dims, str, box, "tc_wgrad(a)")) return rc;
    }
    WgParams p{};
    if (gather_w > 0) {
        const int Co = Nb / 4, w = gather_w;
        const int gw = w >= 64 ? 64 : w;
        EEL_REQUIRE(64 % gw == 0 && w % gw == 0 && P % w == 0, "tc_wgrad: gathered width %d must divide or be a multiple of 64", w);
        uint64_t dims[4] = {(uint64_t)2 * Co, (uint64_t)w, 2, (uint64_t)(P / w)};
        uint64_t str[4] = {1, (uint64_t)2 * Co, (uint64_t)2 * w * Co, (uint64_t)4 * w * Co};
        uint32_t box[4] = {64, (uint32_t)gw, 1, (uint32_t)(64 / gw)};
        if (int rc = make_tmap_bf16(&tmB, b, 4, dims, str, box, "tc_wgrad(b gather)")) return rc;
        p.gw = gw; p.gW = w; p.per_dy = 2 * Co / 64;
    } else {
        uint64_t dims[2] = {(uint64_t)Nb, (uint64_t)P};
        uint64_t str[2] = {1, (uint64_t)Nb};
        uint32_t box[2] = {64, 64};
        if (int rc = make_tmap_bf16(&tmB, b, 2, dims, str, box, "tc_wgrad(b)")) return rc;
    }
    p.stages_total = cdiv(P, 64);
    p.m_blocks = Ma / 128;
    p.n_blocks = Nb / bn;
    p.Ma = Ma; p.Nb = Nb; p.ldm = ldm; p.ldn = ldn; p.out = out;
    const int units = p.m_blocks * p.n_blocks;
    if (bn == 256) return launch_wg<WG_PLAIN, 256>(tmA, tmB, p, units, st, "tc_wgrad");
    if (bn == 128) return launch_wg<WG_PLAIN, 128>(tmA, tmB, p, units, st, "tc_wgrad");
    return launch_wg<WG_PLAIN, 64>(tmA, tmB, p, units, st, "tc_wgrad");
}

}  // extern "C"
