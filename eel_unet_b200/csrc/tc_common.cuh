// sm_100a building blocks for the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// TMEM allocation, tcgen05.mma / commit / ld, shared-memory and instruction descriptors.
// Inline PTX only (no CUTLASS dependency).  Bit layouts follow the sm_100 UMMA descriptor formats.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eel {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xFFFFFFFF;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (reported error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("eel tc kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- TMEM + tcgen05 -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}


// same load without the wait: lets the next chunk's TMEM read overlap the processing of the current one
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- epilogue: 32 rows x 32 fp32 columns (thread = row) -> bf16 -> coalesced global stores -----------------------
// A row's 32 channels are 64 contiguous output bytes, so storing straight from the TMEM layout would touch 32 lines
// with 16 bytes each per instruction.  The chunk is transposed through a 2 KB per-warp buffer (16-byte slots
// XOR-swizzled, conflict free both ways): afterwards four lanes write one row's 64 bytes and a warp-wide store covers
// full 32-byte sectors only.  After the transposition lane l holds slot (l & 3) of rows (l >> 2) + 8 i, i = 0..3.
struct EpiLane {
    uint32_t wr_base, wr_sw, rd_base;
    int row_lo, slot;
};
__device__ __forceinline__ EpiLane epi_lane(uint8_t* staging_2k, int lane) {
    EpiLane L;
    L.wr_base = smem_u32(staging_2k) + lane * 64;
    L.wr_sw = (lane >> 1) & 3;
    L.row_lo = lane >> 2;
    L.slot = lane & 3;
    L.rd_base = smem_u32(staging_2k) + L.row_lo * 64 + ((L.slot ^ ((lane >> 3) & 3)) << 4);
    return L;
}
// Per-channel sum / sum of squares of a stored chunk (BatchNorm statistics fused into the producer).  After the
// transposition a lane holds 8 channels (its slot) of 4 rows; the 16 partial values {sum[8], sumsq[8]} are
// reduce-scattered over the 8 lanes that share the slot (lane bits 2-4), leaving each lane with TWO finished values:
//     quantity q = bit 4, channel-in-slot c = 4 * bit3 + 2 * bit2 + {0, 1}
// which the caller accumulates across tiles and adds to global memory once per CTA (epi_stats_flush).
// reduce-scatter of 16 per-lane partial values {q0[8], q1[8]} over the 8 lanes that share a slot (lane bits 2-4)
__device__ __forceinline__ void epi_reduce16(const float (&a)[16], int lane, float (&acc)[2]) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
    float b[8], c[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float recv = __shfl_xor_sync(0xffffffffu, h4 ? a[k] : a[k + 8], 16);
        b[k] = (h4 ? a[k + 8] : a[k]) + recv;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float recv = __shfl_xor_sync(0xffffffffu, h3 ? b[k] : b[k + 4], 8);
        c[k] = (h3 ? b[k + 4] : b[k]) + recv;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float recv = __shfl_xor_sync(0xffffffffu, h2 ? c[k] : c[k + 2], 4);
        acc[k] += (h2 ? c[k + 2] : c[k]) + recv;
    }
}
__device__ __forceinline__ void epi_stats_chunk(const uint4 (&o)[4], const bool (&ok)[4], int lane, float (&acc)[2]) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!ok[i]) continue;
        const uint32_t w[4] = {o[i].x, o[i].y, o[i].z, o[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
            a[2 * j] += lo; a[2 * j + 1] += hi;
            a[8 + 2 * j] = fmaf(lo, lo, a[8 + 2 * j]); a[8 + 2 * j + 1] = fmaf(hi, hi, a[8 + 2 * j + 1]);
        }
    }
    epi_reduce16(a, lane, acc);
}
// BatchNorm BACKWARD sums fused into the producer of dy (the data-gradient GEMM in front of a BatchNorm + ReLU): with z the
// BatchNorm's input at the positions of the stored chunk and cst = this lane's 8 channels x {scale, -shift},
//     g = dy * [scale * z > -shift],   quantity 0 = sum g,   quantity 1 = sum g * z
// in the same reduce-scatter layout as epi_stats_chunk (flushed by epi_stats_flush).  The caller turns quantity 1 into
// sum g * xhat = rstd * (sum g z - mean * sum g) afterwards: two constants per channel and seven instructions per element
// keep the epilogue under the tile's MMA time.
__device__ __forceinline__ void epi_bnbwd_chunk(const uint4 (&o)[4], const uint4 (&zv)[4], const bool (&ok)[4],
                                                const float2* __restrict__ cst, int relu, int lane, float (&acc)[2]) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 0.f;
    float2 cc[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(cst) + k);
        cc[2 * k] = make_float2(t.x, t.y);
        cc[2 * k + 1] = make_float2(t.z, t.w);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!ok[i]) continue;
        const uint32_t w[4] = {o[i].x, o[i].y, o[i].z, o[i].w};
        const uint32_t zw[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 2 * j + h;
                float g = h ? __uint_as_float(w[j] & 0xffff0000u) : __uint_as_float(w[j] << 16);
                const float zz = h ? __uint_as_float(zw[j] & 0xffff0000u) : __uint_as_float(zw[j] << 16);
                if (relu && !(zz * cc[k].x > cc[k].y)) g = 0.f;
                a[k] += g;
                a[8 + k] = fmaf(g, zz, a[8 + k]);
            }
        }
    }
    epi_reduce16(a, lane, acc);
}
// sums: [2][C] fp32 (zeroed by the host); col0 = first channel of the chunk these accumulators belong to
__device__ __forceinline__ void epi_stats_flush(float* sums, int C, int col0, int lane, const float (&acc)[2]) {
    const int q = (lane >> 4) & 1;
    const int ch = col0 + (lane & 3) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2;
    atomicAdd(sums + (long long)q * C + ch, acc[0]);
    atomicAdd(sums + (long long)q * C + ch + 1, acc[1]);
}

// pk: this thread's row chunk already packed (32 bf16 = 16 words); dst[i]: where row (row_lo + 8 i) keeps these 32
// columns (already offset by slot * 8 elements), null = skip.  stats != null: fold the stored values into (*stats)[2].
struct EpiBnBwd {              // optional: BatchNorm-backward sums of the stored values (epi_bnbwd_chunk)
    const uint4* zv;           // the BatchNorm's input at the four positions this lane stores (loaded by the caller well ahead:
                               // a DRAM round trip per chunk would otherwise serialise the epilogue)
    const float2* cst;         // this lane's 8 channels x {scale, -shift} for this chunk
    int relu;
};
__device__ __forceinline__ void epi_store_packed(const EpiLane& L, const uint32_t (&pk)[16], __nv_bfloat16* const (&dst)[4],
                                                 float (*stats)[2] = nullptr, int lane = 0, const EpiBnBwd* bb = nullptr) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(L.wr_base + ((c ^ L.wr_sw) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
    __syncwarp();
    uint4 o[4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o[i].x), "=r"(o[i].y), "=r"(o[i].z), "=r"(o[i].w) : "r"(L.rd_base + i * 512) : "memory");
        ok[i] = dst[i] != nullptr;
        if (ok[i]) *reinterpret_cast<uint4*>(dst[i]) = o[i];
    }
    if (stats != nullptr) {
        if (bb != nullptr) {
            const uint4 zv[4] = {bb->zv[0], bb->zv[1], bb->zv[2], bb->zv[3]};
            epi_bnbwd_chunk(o, zv, ok, bb->cst, bb->relu, lane, *stats);
        } else epi_stats_chunk(o, ok, lane, *stats);
    }
    __syncwarp();
}
// r: the 32 accumulator columns of this thread's row; bias32: 32 floats (16-byte aligned) or null
__device__ __forceinline__ void epi_store_chunk(const EpiLane& L, const uint32_t (&r)[32], const float* bias32, int relu,
                                                __nv_bfloat16* const (&dst)[4], float (*stats)[2] = nullptr, int lane = 0,
                                                const EpiBnBwd* bb = nullptr) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (bias32) {
        const float4* b4p = reinterpret_cast<const float4*>(bias32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(b4p + i);
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
        }
    }
    if (relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    epi_store_packed(L, pk, dst, stats, lane, bb);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// a 16-byte read-only load the compiler may not sink towards its use (it is issued HERE, ahead of a long wait)
__device__ __forceinline__ uint4 ldg_early(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory operand descriptor, 128-byte swizzle.  lbo/sbo in bytes.
//   K-major  : rows of 128 B (64 bf16 of K); 8-row groups `sbo` apart; lbo unused.
//   MN-major : rows of 128 B (64 bf16 of M/N) per K index; 8-K-row groups `sbo` apart; 64-wide M/N atoms `lbo` apart.
// base_offset = (start address >> 7) & 7 when the start is not on a 1024-byte swizzle-pattern boundary.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, bool with_base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
    if (with_base_offset) d |= (uint64_t)((saddr >> 7) & 7) << 49;
    d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// bf16 tensor, `rank` dims innermost first, strides in ELEMENTS for dims 1..rank-1, 128B swizzle, zero fill
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, const char* what);

}  // namespace tc
}  // namespace eel
