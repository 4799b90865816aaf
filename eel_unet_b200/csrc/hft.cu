// HighFourierTransform (models/EELUnet.py:153-191) without an FFT.
//
// The reference zeroes the centred 2r x 2r block of the shifted 2-D spectrum (r = min(20, H/2, W/2))
// and takes |ifft2|.  Zeroing frequencies kh, kw in [-r, r) of BOTH axes is the rank-(2r)^2
// projection  P x = U_H (U_H^H x conj(U_W)) U_W^T  with U_N[n, f] = exp(2 pi i n k_f / N) / sqrt(N),
// so   y = | x - P x |   exactly.  P x is four small complex GEMMs per
// image; complex arithmetic is carried as stacked real/imaginary rows.  All steps run on the SIMT
// engine of gemm_simt.cuh in fp32; the channel axis C of the NHWC tensor is the GEMM N dimension.
//
//   step 1  T1 = X conj(U_W)        per (n, h):  [2F x W]  . [W x C]
//   step 2  T2 = U_H^H T1           per n:       [2F x 2H] . [2H x F*C]
//   step 3  T3 = U_H T2             per n:       [2H x 2F] . [2F x F*C]
//   step 4  low = T3 U_W^T          per (n, h):  [2W x 2F] . [2F x C]   + epilogue |x - low|
// Backward: dx = Re[(I - P)(dy * z/|z|)] -- the same four steps on a complex input.
#include "gemm_simt.cuh"

namespace eel {

__global__ void hft_init_mats_kernel(float* __restrict__ Cm, float* __restrict__ Sm, int F, int r, int X) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * X) return;
    int f = i / X, x = i - f * X;
    long long k = f - r;
    long long ph = ((k * x) % X + X) % X;   // exact phase index
    double s, c;
    sincospi(2.0 * (double)ph / (double)X, &s, &c);
    double sc = 1.0 / sqrt((double)X);
    Cm[i] = (float)(c * sc);
    Sm[i] = (float)(s * sc);
}

// how an index splits into (real/imag selector r, position)
enum IdxMode { IDX_R_F = 0 /* r = i / F */, IDX_X = 1 /* r = 0 */, IDX_X_R = 2 /* x = i>>1, r = i&1 */, IDX_R_X = 3 /* r = i / X */ };

struct Idx { int pos; int r; };

__device__ __forceinline__ Idx decode_idx(int i, int mode, int F, int X, int limit) {
    if (i >= limit) return Idx{-1, 0};
    switch (mode) {
        case IDX_R_F: return Idx{i % F, i / F};
        case IDX_X: return Idx{i, 0};
        case IDX_X_R: return Idx{i >> 1, i & 1};
        default: return Idx{i % X, i / X};
    }
}

// A(m, k) = sign(ro, ri) * {C or S}[f][x];  table 0: [[+C, +S], [-S, +C]],  table 1: [[+C, -S], [+S, +C]]
struct SmallA {
    const float* Cm; const float* Sm; int F, X; int m_mode, k_mode; int m_is_f; int table; int M, K;
    __device__ Idx decM(int m) const { return decode_idx(m, m_mode, F, X, M); }
    __device__ Idx decK(int k) const { return decode_idx(k, k_mode, F, X, K); }
    __device__ float at(const Idx& a, const Idx& b) const {
        if (a.pos < 0 || b.pos < 0) return 0.f;
        int f = m_is_f ? a.pos : b.pos, x = m_is_f ? b.pos : a.pos;
        int ro = a.r, ri = b.r;
        bool useS = ro != ri;
        float v = useS ? Sm[f * X + x] : Cm[f * X + x];
        bool neg = useS && ((table == 0) ? (ro == 1) : (ro == 0));
        return neg ? -v : v;
    }
};

// ---- B operand loaders ----------------------------------------------------------------------
// fp32 buffer: value = base[z * zstride + koff(k) + n];  koff(k) = (k % kdiv) * s0 + (k / kdiv) * s1
struct BufB {
    const float* base; long long zstride; int K, N; int kdiv; long long s0, s1;
    struct Col { const float* p; };
    struct Kh { long long off; };
    __device__ Col prep(int z, int n) const { return Col{n < N ? base + (long long)z * zstride + n : nullptr}; }
    __device__ Kh dec(int k) const { return Kh{k < K ? (long long)(k % kdiv) * s0 + (long long)(k / kdiv) * s1 : -1}; }
    __device__ float at(const Col& c, const Kh& k) const { return (c.p && k.off >= 0) ? c.p[k.off] : 0.f; }
};
// forward step 1: the real input rows x[z][w][c]
template <class T> struct InB {
    const T* x; int W, C;
    struct Col { const T* p; };
    struct Kh { long long off; };
    __device__ Col prep(int z, int n) const { return Col{n < C ? x + (long long)z * W * C + n : nullptr}; }
    __device__ Kh dec(int k) const { return Kh{k < W ? (long long)k * C : -1}; }
    __device__ float at(const Col& c, const Kh& k) const { return (c.p && k.off >= 0) ? to_f32(c.p[k.off]) : 0.f; }
};
// backward step 1: complex input g = dy * phase, k = 2w + ri
template <class T> struct GradB {
    const T* dy; const T* ph; int W, C;
    struct Col { const T* d; const T* p; };
    struct Kh { long long off; };
    __device__ Col prep(int z, int n) const {
        if (n >= C) return Col{nullptr, nullptr};
        return Col{dy + (long long)z * W * C + n, ph + (long long)z * 2 * W * C + n};
    }
    __device__ Kh dec(int k) const { return Kh{k < 2 * W ? (long long)k : -1}; }
    __device__ float at(const Col& c, const Kh& k) const {
        if (!c.d || k.off < 0) return 0.f;
        return to_f32(c.d[(k.off >> 1) * C]) * to_f32(c.p[k.off * C]);
    }
};

// ---- epilogues ------------------------------------------------------------------------------
// out[z * zstride + roff(m) + n], roff(m) = (m % rdiv) * s0 + (m / rdiv) * s1
struct BufEpi {
    float* out; long long zstride; int M, N; int rdiv; long long s0, s1;
    __device__ void operator()(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            if (m + i >= M) break;
            float* row = out + (long long)z * zstride + (long long)((m + i) % rdiv) * s0 + (long long)((m + i) / rdiv) * s1;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j)
                if (n + j < N) row[n + j] = acc[i][j];
        }
    }
};
// same addressing, bf16 output (feeds the tensor-core step 4 of hft_tc.cu)
struct BufEpiB {
    bf16* out; long long zstride; int M, N; int rdiv; long long s0, s1;
    __device__ void operator()(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            if (m + i >= M) break;
            bf16* row = out + (long long)z * zstride + (long long)((m + i) % rdiv) * s0 + (long long)((m + i) / rdiv) * s1;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j)
                if (n + j < N) row[n + j] = __float2bfloat16_rn(acc[i][j]);
        }
    }
};
// out[i] = dy[pixel] * phase[pixel][re/im]  (complex upstream gradient as (re, im) row pairs)
__global__ void hft_grad_pairs_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ ph, bf16* __restrict__ g, long long nvec, int C) {
    constexpr int V = 8;
    const int cv = C / V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        long long pix2 = i / cv;          // pixel * 2 + ri
        int c = (int)(i - pix2 * cv) * V;
        Vec16<bf16> a = ld16(dy + (pix2 >> 1) * C + c), b = ld16(ph + i * V), o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.set(j, a.get(j) * b.get(j));
        st16(g + i * V, o);
    }
}
// forward step 4: rows m = 2w + {re, im} of low;  y = |x - low|, phase = (x - low)/|x - low|
template <class T> struct AbsEpi {
    const T* x; T* y; T* ph; int W, C;
    __device__ void operator()(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; i += 2) {
            int w = (m + i) >> 1;
            if (w >= W) break;
            long long base = ((long long)z * W + w) * C;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                if (n + j >= C) break;
                float zr = to_f32(x[base + n + j]) - acc[i][j];
                float zi = -acc[i + 1][j];
                float mag = sqrtf(zr * zr + zi * zi);
                float inv = mag > 0.f ? 1.0f / mag : 0.f;
                y[base + n + j] = from_f32<T>(mag);
                if (ph != nullptr) {
                    ph[base * 2 + n + j] = from_f32<T>(zr * inv);
                    ph[base * 2 + C + n + j] = from_f32<T>(zi * inv);
                }
            }
        }
    }
};
// backward step 4: dx = dy * phase_re - Re(low)
template <class T> struct GradEpi {
    const T* dy; const T* ph; T* dx; int W, C;
    __device__ void operator()(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            int w = m + i;
            if (w >= W) break;
            long long base = ((long long)z * W + w) * C;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                if (n + j >= C) break;
                float g = to_f32(dy[base + n + j]) * to_f32(ph[base * 2 + n + j]);
                dx[base + n + j] = from_f32<T>(g - acc[i][j]);
            }
        }
    }
};

template <class BL, class Epi, bool AKC> struct HftProblem {
    SmallA a; BL b; Epi epi; int M, N, K, Z;
    static constexpr bool A_KCONTIG = AKC;
    static constexpr bool B_NCONTIG = true;
    typedef Idx ARow; typedef Idx AK;
    typedef typename BL::Col BCol; typedef typename BL::Kh BK;
    __host__ __device__ int gridZ() const { return Z; }
    __device__ void krange(int z, int& kb, int& ke) const { kb = 0; ke = K; }
    __device__ ARow prepA(int z, int m) const { return a.decM(m); }
    __device__ AK decA(int k) const { return a.decK(k); }
    __device__ float loadA(const ARow& r, const AK& k) const { return a.at(r, k); }
    __device__ BCol prepB(int z, int n) const { return b.prep(z, n); }
    __device__ BK decB(int k) const { return b.dec(k); }
    __device__ float loadB(const BCol& c, const BK& k) const { return b.at(c, k); }
    __device__ void epilogue(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const { epi(z, m, n, acc); }
};

template <class BL, class Epi, bool AKC>
static int hft_run(const SmallA& a, const BL& b, const Epi& e, int M, int N, int K, long long Z, cudaStream_t st, const char* what) {
    // z runs over up to N*H slices: fold it over grid.z chunks of 65535
    if (Z > 65535) {
        set_error("%s: more than 65535 slices in one launch", what);
        return EEL_ERR_INVALID;
    }
    HftProblem<BL, Epi, AKC> p{a, b, e, M, N, K, (int)Z};
    return launch_gemm_simt(p, st, what);
}

namespace tc {
bool hft_tc_supported(int H, int W, int C, int r);
bool hft_tc_supported_fwd(int H, int W, int C, int r);
size_t hft_tc_matrix_elems(int W);
int hft_tc_step1(const bf16* rows_in, int R, bf16* mat_ws, int kind, float* T, bf16* Tb, int N, int H, int W, int C, int r, cudaStream_t st);
int hft_tc_step2(const bf16* T1b, bf16* mat_ws, bf16* T2b, float* T2f, int N, int H, int C, int r, cudaStream_t st);
int hft_tc_step3(const bf16* T2b, bf16* mat_ws, bf16* T3b, int N, int H, int C, int r, cudaStream_t st);
int hft_tc_step4(const bf16* T3b, bf16* mat_ws, bool fwd, const bf16* x_or_g, bf16* y_or_dx, bf16* phase, int N, int H, int W, int C,
                 int r, cudaStream_t st, bool g_re_only = false);
bool hft_tc_code_path(int H, int W, int C, int r);
int hft_tc_step1g(const bf16* dy, const uint16_t* code, bf16* mat_ws, bf16* T1b, bf16* gre, int N, int H, int W, int C, int r,
                  cudaStream_t st);
int hft_tc_step4p(const bf16* T3b, bf16* mat_ws, bool fwd, const bf16* x_or_dy, const uint16_t* code_in, bf16* y_or_dx,
                  uint16_t* code_out, int N, int H, int W, int C, int r, cudaStream_t st);
}  // namespace tc

struct HftWs { float *Cw, *Sw, *Ch, *Sh, *T1, *T2; bf16 *T3b, *G, *M1, *M4, *M2, *M3; };

static size_t hft_carve(int N, int H, int W, int C, int r, HftWs* ws, void* base, bool need_pairs = true) {
    int F = 2 * r;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n * sizeof(float) + 255) / 256 * 256; return o; };
    size_t oCw = take((size_t)F * W), oSw = take((size_t)F * W), oCh = take((size_t)F * H), oSh = take((size_t)F * H);
    size_t oT1 = take((size_t)N * H * 2 * F * C), oT2 = take((size_t)N * 2 * F * F * C);
    // tensor-core path (bf16): T3 in bf16, the (re, im) gradient pairs, the two resident DFT matrices
    size_t oT3b = take(((size_t)N * H * 2 * F * C + 1) / 2), oG = take(need_pairs ? (size_t)N * H * W * C : ((size_t)N * H * W * C + 1) / 2);   // (re, im) pairs, or Re(g) only
    size_t oM1 = take(((size_t)80 * 2 * W + 1) / 2), oM4 = take(((size_t)2 * W * 128 + 1) / 2);
    size_t oM2 = take(((size_t)80 * 2 * H + 1) / 2), oM3 = take(((size_t)2 * H * 128 + 1) / 2);
    if (ws && base) {
        char* b = (char*)base;
        ws->Cw = (float*)(b + oCw); ws->Sw = (float*)(b + oSw); ws->Ch = (float*)(b + oCh); ws->Sh = (float*)(b + oSh);
        ws->T1 = (float*)(b + oT1); ws->T2 = (float*)(b + oT2);
        ws->T3b = (bf16*)(b + oT3b); ws->G = (bf16*)(b + oG); ws->M1 = (bf16*)(b + oM1); ws->M4 = (bf16*)(b + oM4); ws->M2 = (bf16*)(b + oM2); ws->M3 = (bf16*)(b + oM3);
    }
    return off;
}

static int hft_radius(int H, int W, int mask_range) {
    int r = mask_range;
    if (H / 2 < r) r = H / 2;
    if (W / 2 < r) r = W / 2;
    return r;
}

// steps 2 and 3 are shared by forward and backward
static int hft_middle(const HftWs& w, int N, int H, int C, int F, cudaStream_t st, bool t3_bf16 = false) {
    long long FC = (long long)F * C;
    {   // T2[n][(ro,g)][(f,c)] = sum_(ri,h) A2 * T1[n][h][ri*F+f][c]
        SmallA a{w.Ch, w.Sh, F, H, IDX_R_F, IDX_R_X, 1, 0, 2 * F, 2 * H};
        BufB b{w.T1, (long long)H * 2 * FC, 2 * H, (int)FC, H, 2 * FC, FC};
        BufEpi e{w.T2, 2 * F * FC, 2 * F, (int)FC, 2 * F, FC, 0};
        if (int rc = hft_run<BufB, BufEpi, true>(a, b, e, 2 * F, (int)FC, 2 * H, N, st, "hft.step2")) return rc;
    }
    {   // T3[n][h][ro*F+f][c] = sum_(ri,g) A3 * T2[n][(ri,g)][(f,c)]   (T3 aliases T1)
        SmallA a{w.Ch, w.Sh, F, H, IDX_R_X, IDX_R_F, 0, 1, 2 * H, 2 * F};
        BufB b{w.T2, 2 * F * FC, 2 * F, (int)FC, 2 * F, FC, 0};
        if (t3_bf16) {
            BufEpiB e{w.T3b, (long long)H * 2 * FC, 2 * H, (int)FC, H, 2 * FC, FC};
            if (int rc = hft_run<BufB, BufEpiB, false>(a, b, e, 2 * H, (int)FC, 2 * F, N, st, "hft.step3(bf16)")) return rc;
            return EEL_OK;
        }
        BufEpi e{w.T1, (long long)H * 2 * FC, 2 * H, (int)FC, H, 2 * FC, FC};
        if (int rc = hft_run<BufB, BufEpi, false>(a, b, e, 2 * H, (int)FC, 2 * F, N, st, "hft.step3")) return rc;
    }
    return EEL_OK;
}

}  // namespace eel

using namespace eel;

extern "C" {

size_t eel_hft_workspace_bytes(int N, int H, int W, int C, int mask_range) {
    int r = hft_radius(H, W, mask_range);
    return hft_carve(N, H, W, C, r, nullptr, nullptr, !tc::hft_tc_code_path(H, W, C, r)) + 256;
}

size_t eel_hft_phase_elems(int N, int H, int W, int C, int mask_range, int dtype) {
    int r = hft_radius(H, W, mask_range);
    const size_t px = (size_t)N * H * W * C;
    return (dtype == EEL_BF16 && tc::hft_tc_code_path(H, W, C, r)) ? px : 2 * px;
}

static int hft_prepare(int N, int H, int W, int C, int mask_range, void* ws, size_t ws_bytes, HftWs* w, int* F, cudaStream_t st,
                       bool simt_mats) {
    EEL_REQUIRE(N > 0 && H > 1 && W > 1 && C > 0 && mask_range > 0, "hft: bad argument");
    EEL_REQUIRE((long long)N * H <= 65535, "hft: N*H must be <= 65535 per call");
    int r = hft_radius(H, W, mask_range);
    *F = 2 * r;
    const bool pairs = !tc::hft_tc_code_path(H, W, C, r);
    size_t need = hft_carve(N, H, W, C, r, nullptr, nullptr, pairs);
    uintptr_t al = ((uintptr_t)ws + 255) / 256 * 256;
    if (!ws || ws_bytes < need + (al - (uintptr_t)ws)) { set_error("hft: workspace too small (%zu > %zu)", need, ws_bytes); return EEL_ERR_WORKSPACE; }
    hft_carve(N, H, W, C, r, w, (void*)al, pairs);
    if (!simt_mats) return EEL_OK;      // the tensor-core steps build their own bf16 matrices
    hft_init_mats_kernel<<<cdiv(*F * W, 256), 256, 0, st>>>(w->Cw, w->Sw, *F, r, W);
    if (int rc = check_launch("hft.init_w")) return rc;
    hft_init_mats_kernel<<<cdiv(*F * H, 256), 256, 0, st>>>(w->Ch, w->Sh, *F, r, H);
    return check_launch("hft.init_h");
}

int eel_hft_fwd(const void* x, void* y, void* phase, int N, int H, int W, int C, int mask_range, void* ws, size_t ws_bytes,
                int dtype, eel_stream s) {
    EEL_REQUIRE(x && y, "hft_fwd: null pointer");       // phase may be null (inference: nothing is kept for a backward)
    cudaStream_t st = (cudaStream_t)s;
    HftWs w; int F;
    const bool tcf = dtype == EEL_BF16 && tc::hft_tc_supported_fwd(H, W, C, hft_radius(H, W, mask_range));
    if (int rc = hft_prepare(N, H, W, C, mask_range, ws, ws_bytes, &w, &F, st, !tcf)) return rc;
    if (tcf) {
        // every step on the tensor cores; T1 / T2 / T3 live in bf16 (the fp32 regions T1, T2 are reused as storage).
        // H > 512: step 2 splits its reduction and needs the whole fp32 T2 region for partial sums, so the bf16 T2 goes to
        // the unused second half of the T1 region.
        bf16* T1b = (bf16*)w.T1;
        bf16* T2b = H > 512 ? T1b + (size_t)N * H * 2 * F * C : (bf16*)w.T2;
        if (int rc = tc::hft_tc_step1((const bf16*)x, W, w.M1, 0, nullptr, T1b, N, H, W, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step2(T1b, w.M2, T2b, w.T2, N, H, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step3(T2b, w.M3, w.T3b, N, H, C, F / 2, st)) return rc;
        if (tc::hft_tc_code_path(H, W, C, F / 2))      // |z| and the 16-bit phase code from the pixel-row epilogue
            return tc::hft_tc_step4p(w.T3b, w.M4, true, (const bf16*)x, nullptr, (bf16*)y, (uint16_t*)phase, N, H, W, C, F / 2, st);
        return tc::hft_tc_step4(w.T3b, w.M4, true, (const bf16*)x, (bf16*)y, (bf16*)phase, N, H, W, C, F / 2, st);
    }
    EEL_DISPATCH_DTYPE(dtype, {
        {   // T1[(n,h)][(ro,f)][c] = sum_w A1 * x[n,h,w,c],  A1 = [C; -S]
            SmallA a{w.Cw, w.Sw, F, W, IDX_R_F, IDX_X, 1, 0, 2 * F, W};
            InB<T> b{(const T*)x, W, C};
            BufEpi e{w.T1, 2LL * F * C, 2 * F, C, 2 * F, C, 0};
            if (int rc = hft_run<InB<T>, BufEpi, true>(a, b, e, 2 * F, C, W, (long long)N * H, st, "hft_fwd.step1")) return rc;
        }
        if (int rc = hft_middle(w, N, H, C, F, st)) return rc;
        {   // low[(n,h)][2w+ro][c] = sum_(ri,f) A4 * T3[(n,h)][(ri,f)][c];  epilogue |x - low|
            SmallA a{w.Cw, w.Sw, F, W, IDX_X_R, IDX_R_F, 0, 1, 2 * W, 2 * F};
            BufB b{w.T1, 2LL * F * C, 2 * F, C, 2 * F, C, 0};
            AbsEpi<T> e{(const T*)x, (T*)y, (T*)phase, W, C};
            return (hft_run<BufB, AbsEpi<T>, false>(a, b, e, 2 * W, C, 2 * F, (long long)N * H, st, "hft_fwd.step4"));
        }
    });
}

int eel_hft_bwd(const void* dy, const void* phase, void* dx, int N, int H, int W, int C, int mask_range, void* ws,
                size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(dy && phase && dx, "hft_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)s;
    HftWs w; int F;
    const bool tcb = dtype == EEL_BF16 && tc::hft_tc_supported(H, W, C, hft_radius(H, W, mask_range));
    if (int rc = hft_prepare(N, H, W, C, mask_range, ws, ws_bytes, &w, &F, st, !tcb)) return rc;
    if (dtype == EEL_BF16 && tc::hft_tc_code_path(H, W, C, F / 2)) {
        // phase arrives as the 16-bit code; the complex gradient dy * phase is formed inside the first kernel
        // (the first kernel also leaves Re(g) in the workspace: the last step then has the light epilogue dx = Re(g) - Re(low))
        if (int rc = tc::hft_tc_step1g((const bf16*)dy, (const uint16_t*)phase, w.M1, (bf16*)w.T1, w.G, N, H, W, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step2((const bf16*)w.T1, w.M2, (bf16*)w.T2, w.T2, N, H, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step3((const bf16*)w.T2, w.M3, w.T3b, N, H, C, F / 2, st)) return rc;
        return tc::hft_tc_step4(w.T3b, w.M4, false, w.G, (bf16*)dx, nullptr, N, H, W, C, F / 2, st, true);
    }
    if (dtype == EEL_BF16 && tc::hft_tc_supported(H, W, C, F / 2)) {
        const long long nvec = (long long)N * H * W * 2 * C / 8;
        long long blocks = (nvec + 255) / 256;
        if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
        hft_grad_pairs_kernel<<<(int)blocks, 256, 0, st>>>((const bf16*)dy, (const bf16*)phase, w.G, nvec, C);
        if (int rc = check_launch("hft_bwd.pairs")) return rc;
        if (int rc = tc::hft_tc_step1(w.G, 2 * W, w.M1, 1, nullptr, (bf16*)w.T1, N, H, W, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step2((const bf16*)w.T1, w.M2, (bf16*)w.T2, w.T2, N, H, C, F / 2, st)) return rc;
        if (int rc = tc::hft_tc_step3((const bf16*)w.T2, w.M3, w.T3b, N, H, C, F / 2, st)) return rc;
        return tc::hft_tc_step4(w.T3b, w.M4, false, w.G, (bf16*)dx, nullptr, N, H, W, C, F / 2, st);
    }
    EEL_DISPATCH_DTYPE(dtype, {
        {   // complex input g = dy * phase, k = 2w + ri, A = [[C, S], [-S, C]]
            SmallA a{w.Cw, w.Sw, F, W, IDX_R_F, IDX_X_R, 1, 0, 2 * F, 2 * W};
            GradB<T> b{(const T*)dy, (const T*)phase, W, C};
            BufEpi e{w.T1, 2LL * F * C, 2 * F, C, 2 * F, C, 0};
            if (int rc = hft_run<GradB<T>, BufEpi, true>(a, b, e, 2 * F, C, 2 * W, (long long)N * H, st, "hft_bwd.step1")) return rc;
        }
        if (int rc = hft_middle(w, N, H, C, F, st)) return rc;
        {   // Re(low)[(n,h)][w][c], A = [C, -S];  dx = dy * phase_re - Re(low)
            SmallA a{w.Cw, w.Sw, F, W, IDX_X, IDX_R_F, 0, 1, W, 2 * F};
            BufB b{w.T1, 2LL * F * C, 2 * F, C, 2 * F, C, 0};
            GradEpi<T> e{(const T*)dy, (const T*)phase, (T*)dx, W, C};
            return (hft_run<BufB, GradEpi<T>, false>(a, b, e, W, C, 2 * F, (long long)N * H, st, "hft_bwd.step4"));
        }
    });
}

}  // extern "C"
