// The first convolution (models/EELUnet.py:338 with in_channels = 3: K = 27) on the tensor cores, bf16 mode.
//
// The layer is HBM-bound on its 64-channel output (K = 27 is far too thin for an implicit-GEMM tile of its own), so it is
// fed to the generic tcgen05 GEMM / weight-gradient kernels through a compact im2col:
//
//   col[p][32]        = the 27 taps (ky, kx, c) of pixel p, zero padded to 32        (half the bytes of the output)
//   col2[p/2][64]     = two neighbouring pixels side by side  ->  K = 64
//   wblk[128][64]     = block-diagonal: rows 0..63 hold W in columns 0..26, rows 64..127 hold W in columns 32..58
//   y2[p/2][128]      = col2 . wblk^T  =  y[p][64]   (eel_tc_linear, bias duplicated, BatchNorm sums [2][128] folded after)
//   dwblk[128][64]    = y-gradient2^T . col2          (eel_tc_wgrad);  dW = diagonal block 0 + diagonal block 1
//
// The zero blocks double the (negligible) FLOPs; the bytes moved are the im2col (1/2 of y) + y.
#include "common.cuh"

namespace eel {

// one thread per pixel: 27 bf16 taps (+5 zeros) = 64 bytes
__global__ void __launch_bounds__(256) stem_im2col_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int N, int H, int W) {
    const long long P = (long long)N * H * W;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(p % W);
        const int h = (int)((p / W) % H);
        uint32_t pk[16];
        unsigned short v[32];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
            const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
            const bf16* src = x + (p + (long long)(t / 3 - 1) * W + (t % 3 - 1)) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[t * 3 + c] = in ? __bfloat16_as_ushort(src[c]) : (unsigned short)0;
        }
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = (uint32_t)v[2 * k] | ((uint32_t)v[2 * k + 1] << 16);
        uint4* dst = reinterpret_cast<uint4*>(col + p * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
}

// w:[64][3][3][3] (Cout, Cin, ky, kx) fp32 -> wblk:[128][64] bf16 block diagonal, bias2:[128] = {bias, bias}
__global__ void stem_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ wblk,
                                 float* __restrict__ bias2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 128 * 64) {
        const int n = i / 64, k = i % 64;
        const int co = n & 63, kk = k - (n >> 6) * 32;         // tap index inside this row's diagonal block
        float v = 0.f;
        if (kk >= 0 && kk < 27) {
            const int t = kk / 3, c = kk % 3;
            v = w[((co * 3 + c) * 3 + t / 3) * 3 + t % 3];
        }
        wblk[i] = __float2bfloat16_rn(v);
    }
    if (i < 128) bias2[i] = bias != nullptr ? bias[i & 63] : 0.f;
}

// dwblk:[128][64] fp32 -> dw:[64][3][3][3]: the two diagonal blocks added
__global__ void stem_unpack_dw_kernel(const float* __restrict__ dwblk, float* __restrict__ dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 27) return;
    const int co = i / 27, r = i % 27;
    const int c = r / 9, t = r % 9;                              // dw index (co, c, ky, kx), t = ky * 3 + kx
    const int k = t * 3 + c;
    dw[i] = dwblk[co * 64 + k] + dwblk[(64 + co) * 64 + 32 + k];
}

// sums:[2][128] -> out:[2][64]
__global__ void stem_fold_sums_kernel(const float* __restrict__ sums, float* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < 128) out[i] = sums[(i >> 6) * 128 + (i & 63)] + sums[(i >> 6) * 128 + 64 + (i & 63)];
}

}  // namespace eel

using namespace eel;

extern "C" {

int eel_stem_im2col(const void* x, void* col, int N, int H, int W, eel_stream s) {
    EEL_REQUIRE(x && col && N > 0 && H > 0 && W > 0, "stem_im2col: bad argument");
    const long long P = (long long)N * H * W;
    long long blocks = (P + 255) / 256;
    if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
    stem_im2col_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>((const bf16*)x, (bf16*)col, N, H, W);
    return check_launch("stem_im2col");
}

int eel_stem_pack(const float* w, const float* bias, void* wblk, float* bias2, eel_stream s) {
    EEL_REQUIRE(w && wblk && bias2, "stem_pack: bad argument");
    stem_pack_kernel<<<32, 256, 0, (cudaStream_t)s>>>(w, bias, (bf16*)wblk, bias2);
    return check_launch("stem_pack");
}

int eel_stem_unpack_dw(const float* dwblk, float* dw, eel_stream s) {
    EEL_REQUIRE(dwblk && dw, "stem_unpack_dw: bad argument");
    stem_unpack_dw_kernel<<<cdiv(64 * 27, 256), 256, 0, (cudaStream_t)s>>>(dwblk, dw);
    return check_launch("stem_unpack_dw");
}

int eel_stem_fold_sums(const float* sums128, float* sums64, eel_stream s) {
    EEL_REQUIRE(sums128 && sums64, "stem_fold_sums: bad argument");
    stem_fold_sums_kernel<<<1, 128, 0, (cudaStream_t)s>>>(sums128, sums64);
    return check_launch("stem_fold_sums");
}

}  // extern "C"
