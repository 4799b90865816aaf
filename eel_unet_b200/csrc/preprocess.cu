// GPU input pipeline (reference data/ToothDataset.py:58-61 with train.py:249-252): what the reference does per sample on
// the host with torchvision + Pillow --
//     transforms.Resize((H, W)) -> PIL.Image.resize(BILINEAR)   (antialiased: filter support scales with the down-scale)
//     transforms.ToTensor()     -> uint8 / 255, CHW float32
//     transforms.Normalize(mean, std)                            (image only)
// -- on a whole uint8 NHWC batch in one pass.  Pillow's 8-bit resampling is reproduced bit-exactly (Resample.c:
// double-precision coefficients, 22-bit fixed point, horizontal pass rounded to uint8, then vertical pass).
#include "common.cuh"

namespace eel {

constexpr int kPrecBits = 32 - 8 - 2;

// coefficient tables of one axis: bounds[out][2] = (first tap, tap count), kk[out][ksize] fixed-point weights
__global__ void resize_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double ss = 1.0 / filterscale;
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = (x + xmin - center + 0.5) * ss;
        if (a < 0.0) a = -a;
        ww += a < 1.0 ? 1.0 - a : 0.0;
    }
    for (int x = 0; x < ksize; ++x) {
        double w = 0.0;
        if (x < xmax) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w = a < 1.0 ? 1.0 - a : 0.0;
            if (ww != 0.0) w /= ww;
        }
        kk[xx * ksize + x] = w < 0.0 ? (int)(-0.5 + w * (double)(1 << kPrecBits)) : (int)(0.5 + w * (double)(1 << kPrecBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// one thread = one output pixel, all channels
template <int C>
__global__ void __launch_bounds__(256) preprocess_kernel(const unsigned char* __restrict__ in, int N, int Hs, int Ws, int H, int W,
                                                         const int* __restrict__ xb, const int* __restrict__ xk, int kx,
                                                         const int* __restrict__ yb, const int* __restrict__ yk, int ky,
                                                         const float* __restrict__ mean, const float* __restrict__ stdv,
                                                         float* __restrict__ out, unsigned char* __restrict__ out_u8) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)N * H * W) return;
    const int ox = (int)(idx % W);
    const int oy = (int)((idx / W) % H);
    const int n = (int)(idx / ((long long)W * H));
    const int xmin = xb[2 * ox], nx = xb[2 * ox + 1], ymin = yb[2 * oy], ny = yb[2 * oy + 1];
    int vacc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) vacc[c] = 1 << (kPrecBits - 1);
    for (int j = 0; j < ny; ++j) {
        const unsigned char* row = in + (((long long)n * Hs + ymin + j) * Ws + xmin) * C;
        int hacc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) hacc[c] = 1 << (kPrecBits - 1);
        for (int i = 0; i < nx; ++i) {
            const int k = xk[ox * kx + i];
#pragma unroll
            for (int c = 0; c < C; ++c) hacc[c] += (int)row[i * C + c] * k;
        }
        const int kv = yk[oy * ky + j];
#pragma unroll
        for (int c = 0; c < C; ++c) vacc[c] += clip8(hacc[c]) * kv;      // the horizontal pass rounds to uint8 first
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int v = clip8(vacc[c]);
        if (out_u8 != nullptr) out_u8[idx * C + c] = (unsigned char)v;
        if (out != nullptr) {
            float f = (float)v / 255.0f;                                 // ToTensor
            if (mean != nullptr) f = (f - mean[c]) / stdv[c];            // Normalize
            out[(((long long)n * C + c) * H + oy) * W + ox] = f;
        }
    }
}

static int ksize_of(int in_size, int out_size) {
    double scale = (double)in_size / (double)out_size;
    if (scale < 1.0) scale = 1.0;
    double support = scale;
    int c = (int)support;
    if ((double)c < support) ++c;      // ceil
    return c * 2 + 1;
}

}  // namespace eel

using namespace eel;

extern "C" {

size_t eel_preprocess_workspace_bytes(int Hs, int Ws, int H, int W) {
    if (Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(int) * ((size_t)W * (2 + ksize_of(Ws, W)) + (size_t)H * (2 + ksize_of(Hs, H)));
}

int eel_preprocess_u8(const unsigned char* in, int N, int Hs, int Ws, int C, int H, int W, const float* mean, const float* stdv,
                      float* out_nchw, unsigned char* out_u8_nhwc, void* ws, size_t ws_bytes, eel_stream s) {
    EEL_REQUIRE(in && (out_nchw || out_u8_nhwc) && N > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "preprocess_u8: bad argument");
    EEL_REQUIRE(C == 1 || C == 3 || C == 4, "preprocess_u8: 1, 3 or 4 channels (got %d)", C);
    EEL_REQUIRE((mean == nullptr) == (stdv == nullptr), "preprocess_u8: mean and std go together");
    const size_t need = eel_preprocess_workspace_bytes(Hs, Ws, H, W);
    if (!ws || ws_bytes < need) {
        set_error("preprocess_u8: workspace too small (%zu > %zu)", need, ws_bytes);
        return EEL_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)s;
    const int kx = ksize_of(Ws, W), ky = ksize_of(Hs, H);
    int* xb = (int*)ws;
    int* xk = xb + 2 * W;
    int* yb = xk + (size_t)W * kx;
    int* yk = yb + 2 * H;
    resize_coeffs_kernel<<<cdiv(W, 128), 128, 0, st>>>(Ws, W, kx, xb, xk);
    if (int rc = check_launch("preprocess_u8.coeffs_x")) return rc;
    resize_coeffs_kernel<<<cdiv(H, 128), 128, 0, st>>>(Hs, H, ky, yb, yk);
    if (int rc = check_launch("preprocess_u8.coeffs_y")) return rc;
    const long long total = (long long)N * H * W;
    const int grid = cdiv(total, 256);
    if (C == 1) preprocess_kernel<1><<<grid, 256, 0, st>>>(in, N, Hs, Ws, H, W, xb, xk, kx, yb, yk, ky, mean, stdv, out_nchw, out_u8_nhwc);
    else if (C == 3) preprocess_kernel<3><<<grid, 256, 0, st>>>(in, N, Hs, Ws, H, W, xb, xk, kx, yb, yk, ky, mean, stdv, out_nchw, out_u8_nhwc);
    else preprocess_kernel<4><<<grid, 256, 0, st>>>(in, N, Hs, Ws, H, W, xb, xk, kx, yb, yk, ky, mean, stdv, out_nchw, out_u8_nhwc);
    return check_launch("preprocess_u8");
}

}  // extern "C"
