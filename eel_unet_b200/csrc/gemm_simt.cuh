// Generic fp32-accumulate SIMT GEMM over "problem" functors.
//
// This is the exact-arithmetic (FFMA) engine behind fp32 mode: every GEMM-class op of the hot path
// (conv3x3 fprop/dgrad/wgrad, ConvTranspose2d k2s2, 1x1 conv / Linear, the HFT projections) is a
// Problem struct that says how to fetch A(m,k), B(k,n) and what to do with a finished micro-tile.
// bf16 mode uses the tcgen05 kernels in gemm_tc.cu for the heavy shapes; this engine is also the
// on-device cross-check for those.
//
//   C[z][m][n] = sum_{k in krange(z)} A[z](m,k) * B[z](k,n)
//
// Problem interface (all __device__):
//   int M, N;  int gridZ();                        problem extents
//   void krange(int z, int& kb, int& ke)           reduction range of slice z (split-K or batch)
//   ARow prepA(int z, int m); AK decA(int k);  float loadA(const ARow&, const AK&)
//   BCol prepB(int z, int n); BK decB(int k);  float loadB(const BCol&, const BK&)
//   static constexpr bool A_KCONTIG, B_NCONTIG     which index is contiguous in memory (coalescing)
//   void epilogue(int z, int m, int n, const float (&acc)[TM][TN])   rows m..m+TM-1, cols n..n+TN-1
#pragma once
#include <type_traits>
#include "common.cuh"

namespace eel {

template <class P, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_simt_kernel(const P p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_PER = BM * BK / NT;
    constexpr int B_PER = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
    static_assert(NT % BK == 0 && NT % BM == 0 && NT % BN == 0, "loader mapping");
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const int t = threadIdx.x;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    int kb, ke;
    p.krange(z, kb, ke);

    // ---- loader mappings: each thread owns a fixed set of rows (cols) for the whole K loop
    // A, K-contiguous: k = t % BK, rows m = t / BK + i * (NT / BK)
    // A, M-contiguous: m = t % BM, ks   = t / BM + i * (NT / BM)
    constexpr int A_ROWS = P::A_KCONTIG ? A_PER : 1;
    typename P::ARow arow[A_ROWS];
    if (P::A_KCONTIG) {
#pragma unroll
        for (int i = 0; i < A_ROWS; ++i) arow[i] = p.prepA(z, m0 + t / BK + i * (NT / BK));
    } else {
        arow[0] = p.prepA(z, m0 + t % BM);
    }
    constexpr int B_COLS = P::B_NCONTIG ? 1 : B_PER;
    typename P::BCol bcol[B_COLS];
    if (P::B_NCONTIG) {
        bcol[0] = p.prepB(z, n0 + t % BN);
    } else {
#pragma unroll
        for (int i = 0; i < B_COLS; ++i) bcol[i] = p.prepB(z, n0 + t / BK + i * (NT / BK));
    }

    const int ty = t / (BN / TN), tx = t % (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = kb; k0 < ke; k0 += BK) {
        float ra[A_PER], rb[B_PER];
        if (P::A_KCONTIG) {
            const int k = k0 + t % BK;
            if (k < ke) {
                const typename P::AK ak = p.decA(k);
#pragma unroll
                for (int i = 0; i < A_PER; ++i) ra[i] = p.loadA(arow[i], ak);
            } else {
#pragma unroll
                for (int i = 0; i < A_PER; ++i) ra[i] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < A_PER; ++i) {
                const int k = k0 + t / BM + i * (NT / BM);
                ra[i] = (k < ke) ? p.loadA(arow[0], p.decA(k)) : 0.f;
            }
        }
        if (P::B_NCONTIG) {
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                const int k = k0 + t / BN + i * (NT / BN);
                rb[i] = (k < ke) ? p.loadB(bcol[0], p.decB(k)) : 0.f;
            }
        } else {
            const int k = k0 + t % BK;
            if (k < ke) {
                const typename P::BK bk = p.decB(k);
#pragma unroll
                for (int i = 0; i < B_PER; ++i) rb[i] = p.loadB(bcol[i], bk);
            } else {
#pragma unroll
                for (int i = 0; i < B_PER; ++i) rb[i] = 0.f;
            }
        }
        __syncthreads();  // previous tile fully consumed
        if (P::A_KCONTIG) {
#pragma unroll
            for (int i = 0; i < A_PER; ++i) As[t % BK][t / BK + i * (NT / BK)] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < A_PER; ++i) As[t / BM + i * (NT / BM)][t % BM] = ra[i];
        }
        if (P::B_NCONTIG) {
#pragma unroll
            for (int i = 0; i < B_PER; ++i) Bs[t / BN + i * (NT / BN)][t % BN] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < B_PER; ++i) Bs[t % BK][t / BK + i * (NT / BK)] = rb[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    p.epilogue(z, m0 + ty * TM, n0 + tx * TN, acc);
}

// Launch helper.  Tile 128x64x16 with 8x4 micro-tiles (256 threads).
template <class P>
int launch_gemm_simt(const P& p, cudaStream_t stream, const char* what) {
    constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
    if (p.M <= 0 || p.N <= 0 || p.gridZ() <= 0) return EEL_OK;
    dim3 grid(cdiv(p.M, BM), cdiv(p.N, BN), p.gridZ());
    if (grid.y > 65535 || grid.z > 65535) {
        set_error("%s: grid too large (%u,%u,%u)", what, grid.x, grid.y, grid.z);
        return EEL_ERR_INVALID;
    }
    gemm_simt_kernel<P, BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, stream>>>(p);
    return check_launch(what);
}

constexpr int kSimtTM = 8, kSimtTN = 4;

}  // namespace eel
