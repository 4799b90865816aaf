// ChannelAwarePatchedMLP token MLP as ONE kernel (bf16 mode): the three ops that follow the squeeze-excite scaling,
//
//     h = mlp[0](u)            Linear 64 -> 256        (models/EELUnet.py:107)
//     a = GELU(h)              erf form                (:108)
//     z = to_space(mlp[2](a))  the composed 256 -> C   (:109-111,121-122; compose.cu)
//
// run back to back on a 128-pixel tile without the 256-channel intermediates making a round trip through HBM between them:
//
//   warp 0   TMA producer: u tiles (ring of 2), the composed weight Wc in 256 x 64 chunks (ring of 2); W0 is resident
//   warp 1   MMA issuer: acc1[128 x 256] = u . W0^T (TMEM columns 0..255), then per 256-column N tile of the output
//            acc2[128 x 256] = a . Wc^T (columns 256..511) with the A operand read from the shared-memory tile the epilogue wrote
//   warps 2-9  epilogue 1: acc1 + b0 -> h (bf16, stored: the backward's GELU needs it) -> a = GELU(h) (stored: the weight gradient
//            of the composed layer needs it) and written, 128-byte swizzled K-major, into the operand tile of the second GEMM;
//            epilogue 2: acc2 (+ bias) -> z, with the BatchNorm statistics of the following BatchNorm when asked for.
//
// Algorithmic traffic per pixel: 64 read + 256 + 256 + C written (x 2 B), against 64 + 256 | 256 + 256 | 256 + C for the three
// separate launches.  h, a and z are bit-identical to what eel_tc_linear -> eel_gelu_fwd -> eel_tc_linear produce (same MMA order,
// same rounding points), so the existing backward kernels consume them unchanged.
#include "tc_common.cuh"

#include <stdlib.h>

namespace eel {
namespace tc {

constexpr int kMlpHidden = 256, kMlpIn = 64;

struct MlpParams {
    int m_tiles;
    const float* b0;     // [256]
    const float* bc;     // [C] or null (a training-mode BatchNorm follows: the bias cancels in it)
    bf16* h;             // [P][256]
    bf16* a;             // [P][256]
    bf16* z;             // [P][C]
    int C;
    int relu;            // inference: ReLU on z (the folded BatchNorm + ReLU that follow)
    float* bn_sums;      // [2][C] or null
};

// nn.GELU() in the bf16 path: the same rational erfc approximation as elementwise.cu (gelu_fwd_value<bf16>), so that the fused
// and the separate launches agree bit for bit
__device__ __forceinline__ float gelu_bf16_path(float a) {
    const float z = fabsf(a) * 0.70710678118654752f;
    float d = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.078108f, 0.000972f), 0.230389f), 0.278393f), 1.0f);
    d *= d;
    d *= d;
    const float half_erfc = __fdividef(0.5f, d);
    return a * (a >= 0.f ? 1.0f - half_erfc : half_erfc);
}

template <int NT2, bool STATS, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
tc_capmlp_fwd_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW0,
                     const __grid_constant__ CUtensorMap tmWc, const MlpParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW0 = smem;                    // 32 KB   [256 n][64 k]
    uint8_t* sU = sW0 + 32768;              // 2 x 16 KB
    uint8_t* sA2 = sU + 2 * 16384;          // 64 KB   4 atoms [128 rows][64 k]
    uint8_t* sB2 = sA2 + 65536;             // 2 x 32 KB  [256 n][64 k]
    uint8_t* stage = sB2 + 2 * 32768;       // EW x 2 KB epilogue transposition buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2048 * EW);
    uint64_t* w0Full = bars;                // [1]
    uint64_t* uFull = bars + 1;             // [2]
    uint64_t* uEmpty = bars + 3;            // [2]
    uint64_t* b2Full = bars + 5;            // [2]
    uint64_t* b2Empty = bars + 7;           // [2]
    uint64_t* acc1Full = bars + 9;
    uint64_t* acc1Empty = bars + 10;
    uint64_t* a2Full = bars + 11;
    uint64_t* a2Empty = bars + 12;
    uint64_t* acc2Full = bars + 13;
    uint64_t* acc2Empty = bars + 14;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmU);
        tma_prefetch_desc(&tmW0);
        tma_prefetch_desc(&tmWc);
        mbar_init(w0Full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&uFull[i], 1); mbar_init(&uEmpty[i], 1); mbar_init(&b2Full[i], 1); mbar_init(&b2Empty[i], 1); }
        mbar_init(acc1Full, 1); mbar_init(acc1Empty, EW);
        mbar_init(a2Full, EW); mbar_init(a2Empty, 1);
        mbar_init(acc2Full, 1); mbar_init(acc2Empty, EW);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int my_tiles = ((int)p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        mbar_expect_tx(w0Full, 32768);
        tma_load_2d(sW0, &tmW0, w0Full, 0, 0);
        int sb = 0;
        uint32_t pb = 0;
        auto load_u = [&](int it) {
            const int s = it & 1;
            mbar_wait(&uEmpty[s], ((it >> 1) & 1) ^ 1);
            mbar_expect_tx(&uFull[s], 16384);
            tma_load_2d(sU + s * 16384, &tmU, &uFull[s], 0, ((int)blockIdx.x + it * (int)gridDim.x) * 128);
        };
        if (my_tiles > 0) load_u(0);
        for (int it = 0; it < my_tiles; ++it) {
            if (it + 1 < my_tiles) load_u(it + 1);      // the next tile's input is on its way before this tile's weights
            for (int nt = 0; nt < NT2; ++nt)
                for (int kc = 0; kc < 4; ++kc) {
                    mbar_wait(&b2Empty[sb], pb ^ 1);
                    mbar_expect_tx(&b2Full[sb], 32768);
                    tma_load_2d(sB2 + sb * 32768, &tmWc, &b2Full[sb], kc * 64, nt * 256);
                    if (++sb == 2) { sb = 0; pb ^= 1; }
                }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
        const uint64_t w0_desc = make_smem_desc(smem_u32(sW0), 16, 1024, false);
        const uint64_t u_desc0 = make_smem_desc(smem_u32(sU), 16, 1024, false);
        const uint64_t a2_desc0 = make_smem_desc(smem_u32(sA2), 16, 1024, false);
        const uint64_t b2_desc0 = make_smem_desc(smem_u32(sB2), 16, 1024, false);
        mbar_wait(w0Full, 0);
        tc_fence_after();
        int sb = 0, n2 = 0;
        uint32_t pb = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            // ---- GEMM 1: acc1 = u . W0^T
            mbar_wait(&uFull[s], (it >> 1) & 1);
            mbar_wait(acc1Empty, (it & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base, u_desc0 + (uint32_t)((s * 16384) >> 4) + (uint32_t)(k * 2), w0_desc + (uint32_t)(k * 2), idesc, (uint32_t)(k != 0));
            umma_commit(&uEmpty[s]);
            umma_commit(acc1Full);
            // ---- GEMM 2: acc2 = a . Wc^T, one 256-column N tile at a time
            mbar_wait(a2Full, it & 1);
            tc_fence_after();
            for (int nt = 0; nt < NT2; ++nt, ++n2) {
                mbar_wait(acc2Empty, (n2 & 1) ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < 4; ++kc) {
                    mbar_wait(&b2Full[sb], pb);
                    tc_fence_after();
                    const uint64_t a_view = a2_desc0 + (uint32_t)((kc * 16384) >> 4);
                    const uint64_t b_view = b2_desc0 + (uint32_t)((sb * 32768) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + 256, a_view + (uint32_t)(k * 2), b_view + (uint32_t)(k * 2), idesc, (uint32_t)((kc | k) != 0));
                    umma_commit(&b2Empty[sb]);
                    if (++sb == 2) { sb = 0; pb ^= 1; }
                }
                umma_commit(acc2Full);
            }
            umma_commit(a2Empty);
        }
    } else if (warp >= 2) {
        // ===================================================================== epilogues (EW warps)
        // The chain TMEM read -> convert (-> GELU) -> transposition -> store is bound by its latency, not by its instruction count:
        // EW / 4 warps per TMEM lane quarter, each with PW = 1024 / EW of the 256 columns
        constexpr int PW = 1024 / EW, NCH = PW / 32;
        const int q = warp & 3;                 // TMEM lane quarter
        const int half = (warp - 2) >> 2;       // which PW of the 256 columns
        const EpiLane L = epi_lane(stage + (warp - 2) * 2048, lane);
        const int row = q * 32 + lane;          // this thread's accumulator row
        float st[STATS ? NT2 * NCH : 1][2];
#pragma unroll
        for (int i = 0; i < (STATS ? NT2 * NCH : 1); ++i) st[i][0] = st[i][1] = 0.f;
        int n2 = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const uint32_t m0 = (uint32_t)((int)blockIdx.x + it * (int)gridDim.x) * 128u;
            const uint32_t rbase = m0 + (uint32_t)(q * 32 + L.row_lo);       // row of dst[0] after the transposition
            // ---- epilogue 1
            mbar_wait(acc1Full, it & 1);
            mbar_wait(a2Empty, (it & 1) ^ 1);
            tc_fence_after();
            const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + half * PW;
#pragma unroll 1
            for (int ci = 0; ci < NCH; ++ci) {
                const int col = half * PW + ci * 32;
                float v[32];
                tmem_ld32(t1 + ci * 32, v);
                const float4* b4p = reinterpret_cast<const float4*>(p.b0 + col);
                uint32_t pkh[16], pka[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = __ldg(b4p + i);
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[4 * i] + b4.x, v[4 * i + 1] + b4.y);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
                    pkh[2 * i] = *reinterpret_cast<uint32_t*>(&h0);
                    pkh[2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
                    // GELU of the ROUNDED h (what a separate GELU launch would read back)
                    __nv_bfloat162 a0 = __floats2bfloat162_rn(gelu_bf16_path(__low2float(h0)), gelu_bf16_path(__high2float(h0)));
                    __nv_bfloat162 a1 = __floats2bfloat162_rn(gelu_bf16_path(__low2float(h1)), gelu_bf16_path(__high2float(h1)));
                    pka[2 * i] = *reinterpret_cast<uint32_t*>(&a0);
                    pka[2 * i + 1] = *reinterpret_cast<uint32_t*>(&a1);
                }
                // operand tile of GEMM 2: row `row`, K columns col .. col+31 -> atom col/64, four 16-byte slots, 128-byte swizzle
                {
                    uint8_t* atom = sA2 + (col >> 6) * 16384 + row * 128;
                    const int j0 = (col & 63) >> 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(atom + (((j0 + j) ^ (row & 7)) << 4)) =
                            make_uint4(pka[4 * j], pka[4 * j + 1], pka[4 * j + 2], pka[4 * j + 3]);
                }
                if (p.h != nullptr) {       // (inference keeps neither intermediate)
                    bf16 *dh[4], *da[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const size_t o = (size_t)(rbase + 8 * i) * kMlpHidden + col + L.slot * 8;
                        dh[i] = p.h + o;
                        da[i] = p.a + o;
                    }
                    epi_store_packed(L, pkh, dh);
                    epi_store_packed(L, pka, da);
                }
            }
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of sA2 -> tcgen05.mma reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(acc1Empty); mbar_arrive(a2Full); }
            // ---- epilogue 2
            for (int nt = 0; nt < NT2; ++nt, ++n2) {
                mbar_wait(acc2Full, n2 & 1);
                tc_fence_after();
                const uint32_t t2 = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + half * PW;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    uint32_t buf[32];
                    tmem_ld32_async(t2 + ci * 32, buf);
                    tmem_ld_wait();
                    const int col = nt * 256 + half * PW + ci * 32;
                    bf16* dz[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dz[i] = p.z + (size_t)(rbase + 8 * i) * p.C + col + L.slot * 8;
                    epi_store_chunk(L, buf, p.bc ? p.bc + col : nullptr, p.relu, dz, STATS ? &st[STATS ? nt * NCH + ci : 0] : nullptr, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc2Empty);
            }
        }
        if (STATS) {
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci)
                    epi_stats_flush(p.bn_sums, p.C, nt * 256 + half * PW + ci * 32, lane, st[STATS ? nt * NCH + ci : 0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int NT2, bool STATS, int EW>
static int launch_mlp_ew(const CUtensorMap& u, const CUtensorMap& w0, const CUtensorMap& wc, const MlpParams& p, cudaStream_t st) {
    constexpr int smem = 32768 + 2 * 16384 + 65536 + 2 * 32768 + 2048 * EW + 256 + 1024;
    static SmemOptIn configured;
    if (!configured.ensure(tc_capmlp_fwd_kernel<NT2, STATS, EW>, smem)) {
        set_error("tc_capmlp_fwd: cannot raise dynamic shared memory to %d", smem);
        return EEL_ERR_CUDA;
    }
    const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
    tc_capmlp_fwd_kernel<NT2, STATS, EW><<<grid, 64 + 32 * EW, smem, st>>>(u, w0, wc, p);
    return check_launch("tc_capmlp_fwd");
}
template <int NT2, bool STATS>
static int launch_mlp(const CUtensorMap& u, const CUtensorMap& w0, const CUtensorMap& wc, const MlpParams& p, cudaStream_t st) {
    static const int ew = [] { const char* e = getenv("EEL_MLP_EW"); return e ? atoi(e) : 16; }();
    if (ew == 8) return launch_mlp_ew<NT2, STATS, 8>(u, w0, wc, p, st);
    return launch_mlp_ew<NT2, STATS, 16>(u, w0, wc, p, st);
}

}  // namespace tc
}  // namespace eel

using namespace eel;
using namespace eel::tc;

extern "C" {

int eel_tc_capmlp_fwd(const void* u, const void* w0, const float* b0, const void* wc, const float* bc, void* h, void* a, void* z,
                      long long P, int C, int relu, float* bn_sums, eel_stream s) {
    EEL_REQUIRE(u && w0 && b0 && wc && z && P > 0 && ((h == nullptr) == (a == nullptr)), "tc_capmlp_fwd: bad argument");
    EEL_REQUIRE(P % 128 == 0 && (C == 256 || C == 512 || C == 1024), "tc_capmlp_fwd: needs P %% 128 == 0 and C in {256, 512, 1024} (got %lld, %d)", P, C);
    EEL_REQUIRE(P * (long long)C < (1LL << 40), "tc_capmlp_fwd: output too large");
    cudaStream_t st = (cudaStream_t)s;
    CUtensorMap tmU, tmW0, tmWc;
    {
        uint64_t dims[2] = {(uint64_t)kMlpIn, (uint64_t)P};
        uint64_t str[2] = {1, (uint64_t)kMlpIn};
        uint32_t box[2] = {64, 128};
        if (int rc = make_tmap_bf16(&tmU, u, 2, dims, str, box, "tc_capmlp_fwd(u)")) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)kMlpIn, (uint64_t)kMlpHidden};
        uint64_t str[2] = {1, (uint64_t)kMlpIn};
        uint32_t box[2] = {64, 256};
        if (int rc = make_tmap_bf16(&tmW0, w0, 2, dims, str, box, "tc_capmlp_fwd(W0)")) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)kMlpHidden, (uint64_t)C};
        uint64_t str[2] = {1, (uint64_t)kMlpHidden};
        uint32_t box[2] = {64, 256};
        if (int rc = make_tmap_bf16(&tmWc, wc, 2, dims, str, box, "tc_capmlp_fwd(Wc)")) return rc;
    }
    MlpParams p{};
    p.m_tiles = (int)(P / 128);
    p.b0 = b0; p.bc = bc; p.h = (bf16*)h; p.a = (bf16*)a; p.z = (bf16*)z; p.C = C; p.relu = relu; p.bn_sums = bn_sums;
    if (bn_sums != nullptr) {
        if (cudaMemsetAsync(bn_sums, 0, sizeof(float) * 2 * C, st) != cudaSuccess) { set_error("tc_capmlp_fwd: memset failed"); return EEL_ERR_CUDA; }
        if (C == 256) return launch_mlp<1, true>(tmU, tmW0, tmWc, p, st);
        if (C == 512) return launch_mlp<2, true>(tmU, tmW0, tmWc, p, st);
        return launch_mlp<4, true>(tmU, tmW0, tmWc, p, st);
    }
    if (C == 256) return launch_mlp<1, false>(tmU, tmW0, tmWc, p, st);
    if (C == 512) return launch_mlp<2, false>(tmU, tmW0, tmWc, p, st);
    return launch_mlp<4, false>(tmU, tmW0, tmWc, p, st);
}

}  // extern "C"
