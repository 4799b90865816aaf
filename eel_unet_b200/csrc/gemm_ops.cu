// GEMM-class ops of the hot path on the exact SIMT engine (gemm_simt.cuh):
// conv3x3 fprop/dgrad/wgrad, ConvTranspose2d(k2,s2) fprop/dgrad/wgrad, 1x1 conv / Linear (+ folded
// ShiftedChannel).  All operands are 2-D views X(i, j) -- i a pixel-like index, j a channel-like
// (memory-contiguous) index -- described by small accessor structs; a Problem wires two accessors
// and an epilogue into the engine.
#include "gemm_simt.cuh"

#include <type_traits>

namespace eel {

// ------------------------------------------------------------------------------------ accessors
// Dense row-major matrix X(i, j) = base[i * ld + j]
template <class T> struct DenseAcc {
    const T* base; long long I; int J; long long ld;
    struct IH { const T* p; };
    struct JH { int j; };
    __device__ IH prepI(long long i) const { return IH{i < I ? base + i * ld : nullptr}; }
    __device__ JH prepJ(int j) const { return JH{j < J ? j : -1}; }
    __device__ float at(const IH& a, const JH& b) const {
        return (a.p != nullptr && b.j >= 0) ? to_f32(a.p[b.j]) : 0.f;
    }
};

// im2col view of an NHWC tensor for a 3x3 / pad 1 convolution: i = (n, y, x), j = (tap, c)
template <class T> struct Conv3Acc {
    const T* base; int N, H, W, C; int flip;
    struct IH { const T* p; int y, x; };
    struct JH { int off, dy, dx; };
    __device__ IH prepI(long long i) const {
        if (i >= (long long)N * H * W) return IH{nullptr, 0, 0};
        int x = (int)(i % W);
        int y = (int)((i / W) % H);
        return IH{base + i * C, y, x};
    }
    __device__ JH prepJ(int j) const {
        if (j >= 9 * C) return JH{0, 100, 100};
        int tap = j / C, c = j - tap * C;
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (flip) { dy = -dy; dx = -dx; }
        return JH{(dy * W + dx) * C + c, dy, dx};
    }
    __device__ float at(const IH& a, const JH& b) const {
        int yy = a.y + b.dy, xx = a.x + b.dx;
        if (a.p == nullptr || (unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) return 0.f;
        return to_f32(a.p[b.off]);
    }
};

// ShiftedChannel view (models/EELUnet.py:88-97) of an [N,H,W,C] tensor: i = pixel, j = channel.
// quarter 0 reads row y-1, quarter 1 row y+1, quarter 2 column x-1 (all circular), rest unshifted.
template <class T> struct ShiftAcc {
    const T* base; long long P; int H, W, C;
    struct IH { long long img; int y, x; };   // img = n*H*W, or -1
    struct JH { int c, dy, dx; };
    __device__ IH prepI(long long i) const {
        if (i >= P) return IH{-1, 0, 0};
        int x = (int)(i % W);
        int y = (int)((i / W) % H);
        return IH{i - (long long)y * W - x, y, x};
    }
    __device__ JH prepJ(int j) const {
        if (j >= C) return JH{-1, 0, 0};
        int q = j / (C / 4);
        return JH{j, q == 0 ? -1 : (q == 1 ? 1 : 0), q == 2 ? -1 : 0};
    }
    __device__ long long index(const IH& a, const JH& b) const {
        int yy = a.y + b.dy, xx = a.x + b.dx;
        yy = yy < 0 ? yy + H : (yy >= H ? yy - H : yy);
        xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
        return (a.img + (long long)yy * W + xx) * C + b.c;
    }
    __device__ float at(const IH& a, const JH& b) const {
        return (a.img >= 0 && b.c >= 0) ? to_f32(base[index(a, b)]) : 0.f;
    }
};

// View of the [N,2h,2w,Co] output-side tensor of a ConvTranspose2d(k2,s2) indexed by the INPUT pixel:
// i = (n, y, x) over h x w, j = (dy, dx, co)
template <class T> struct ConvTAcc {
    const T* base; int N, h, w, Co;
    struct IH { const T* p; };
    struct JH { long long off; };
    __device__ IH prepI(long long i) const {
        if (i >= (long long)N * h * w) return IH{nullptr};
        int x = (int)(i % w);
        long long r = i / w;   // n*h + y
        return IH{base + ((r * 2) * (2LL * w) + 2 * x) * Co};
    }
    __device__ JH prepJ(int j) const {
        if (j >= 4 * Co) return JH{-1};
        int dy = j / (2 * Co), rem = j - dy * 2 * Co;
        return JH{(long long)dy * 2 * w * Co + rem};
    }
    __device__ float at(const IH& a, const JH& b) const {
        return (a.p != nullptr && b.off >= 0) ? to_f32(a.p[b.off]) : 0.f;
    }
};

// ------------------------------------------------------------------------------------ epilogues
template <class T> struct StoreEpi {
    T* out; long long M; int N; long long ld; const float* bias; int relu;
    __device__ void operator()(int z, long long m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            if (m + i >= M) break;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                if (n + j >= N) break;
                float v = acc[i][j] + (bias ? bias[n + j] : 0.f);
                if (relu) v = fmaxf(v, 0.f);
                out[(m + i) * ld + n + j] = from_f32<T>(v);
            }
        }
    }
};

struct AtomicEpi {
    float* out; int M, N; long long ld;
    __device__ void operator()(int z, long long m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            if (m + i >= M) break;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                if (n + j >= N) break;
                atomicAdd(out + (m + i) * ld + n + j, acc[i][j]);
            }
        }
    }
};

// ConvTranspose scatter: row m = input pixel, col n = (dy, dx, co)
template <class T> struct ConvTScatterEpi {
    T* out; const float* bias; int N_, h, w, Co;
    __device__ void operator()(int z, long long m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
        ConvTAcc<T> v{out, N_, h, w, Co};
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            typename ConvTAcc<T>::IH ih = v.prepI(m + i);
            if (ih.p == nullptr) break;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                typename ConvTAcc<T>::JH jh = v.prepJ(n + j);
                if (jh.off < 0) break;
                int co = (n + j) % Co;
                const_cast<T*>(ih.p)[jh.off] = from_f32<T>(acc[i][j] + (bias ? bias[co] : 0.f));
            }
        }
    }
};

// adjoint of ShiftAcc: the gradient of shifted element (i, j) lands where the forward read it
template <class T> struct ShiftScatterEpi {
    T* out; long long P; int H, W, C;
    __device__ void operator()(int z, long long m, int n, const float (&acc)[kSimtTM][kSimtTN]) const {
        ShiftAcc<T> v{out, P, H, W, C};
#pragma unroll
        for (int i = 0; i < kSimtTM; ++i) {
            typename ShiftAcc<T>::IH ih = v.prepI(m + i);
            if (ih.img < 0) break;
#pragma unroll
            for (int j = 0; j < kSimtTN; ++j) {
                typename ShiftAcc<T>::JH jh = v.prepJ(n + j);
                if (jh.c < 0) break;
                out[v.index(ih, jh)] = from_f32<T>(acc[i][j]);
            }
        }
    }
};

// ------------------------------------------------------------------------------------ problem
// A(m,k) = XA(m,k) or XA(k,m) [TA]; B(k,n) = XB(k,n) or XB(n,k) [TB]; K split over gridZ slices.
template <class XA, bool TA, class XB, bool TB, class Epi> struct Problem {
    XA xa; XB xb; Epi epi;
    int M, N; long long K; int splits; long long kchunk;
    static constexpr bool A_KCONTIG = !TA;
    static constexpr bool B_NCONTIG = !TB;
    typedef typename std::conditional<TA, typename XA::JH, typename XA::IH>::type ARow;
    typedef typename std::conditional<TA, typename XA::IH, typename XA::JH>::type AK;
    typedef typename std::conditional<TB, typename XB::IH, typename XB::JH>::type BCol;
    typedef typename std::conditional<TB, typename XB::JH, typename XB::IH>::type BK;
    __host__ __device__ int gridZ() const { return splits; }
    __device__ void krange(int z, int& kb, int& ke) const {
        long long b = (long long)z * kchunk, e = b + kchunk;
        kb = (int)b; ke = (int)(e < K ? e : K);
    }
    __device__ ARow prepA(int z, int m) const { if constexpr (TA) return xa.prepJ(m); else return xa.prepI(m); }
    __device__ AK decA(int k) const { if constexpr (TA) return xa.prepI(k); else return xa.prepJ(k); }
    __device__ float loadA(const ARow& r, const AK& k) const { if constexpr (TA) return xa.at(k, r); else return xa.at(r, k); }
    __device__ BCol prepB(int z, int n) const { if constexpr (TB) return xb.prepI(n); else return xb.prepJ(n); }
    __device__ BK decB(int k) const { if constexpr (TB) return xb.prepJ(k); else return xb.prepI(k); }
    __device__ float loadB(const BCol& c, const BK& k) const { if constexpr (TB) return xb.at(c, k); else return xb.at(k, c); }
    __device__ void epilogue(int z, int m, int n, const float (&acc)[kSimtTM][kSimtTN]) const { epi(z, m, n, acc); }
};

template <class XA, bool TA, class XB, bool TB, class Epi>
static int run(const XA& xa, const XB& xb, const Epi& epi, long long M, long long N, long long K, int splits,
               cudaStream_t st, const char* what) {
    if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) {
        set_error("%s: extent over 2^31", what);
        return EEL_ERR_INVALID;
    }
    Problem<XA, TA, XB, TB, Epi> p{xa, xb, epi, (int)M, (int)N, K, 1, K};
    if (splits > 1) {
        long long chunk = (K + splits - 1) / splits;
        chunk = (chunk + 15) / 16 * 16;
        p.splits = (int)((K + chunk - 1) / chunk);
        p.kchunk = chunk;
    }
    return launch_gemm_simt(p, st, what);
}

// split-K factor for weight gradients: enough blocks for ~3 waves, at least 256 reduction steps each
static int pick_splits(long long M, long long N, long long K) {
    long long tiles = (long long)cdiv(M, 128) * cdiv(N, 64);
    long long want = (3LL * kNumSMs * 2 + tiles - 1) / tiles;
    long long maxs = K / 256 > 0 ? K / 256 : 1;
    long long s = want < maxs ? want : maxs;
    return (int)(s < 1 ? 1 : (s > 4096 ? 4096 : s));
}

// ------------------------------------------------------------------------------------ stem forward
// conv3x3 forward for the 3(4)-channel input layer (K = 27): HBM-bound on its 64-channel output.  One block walks image
// rows; the three input rows live in shared memory (fp32); a thread owns 4 output channels with all 9*CIN weights in
// registers and strides over the row's pixels; 16 threads write one pixel's 64 channels as one contiguous line.
// smem row pitch in floats: a 4-pixel window (6 columns x CIN floats starting at pixel w, w % 4 == 0) is 16-byte aligned
// 128-thread blocks: ~150-220 registers per thread still leave 2-3 blocks per SM, so one block's row load and store
// phases overlap another's FMA phase
constexpr int kStemThreads = 128;
__host__ __device__ constexpr int stem_pitch(int W, int CIN) { return ((W + 2) * CIN + 3) / 4 * 4 + 4; }

template <class T, int CIN>
__device__ __forceinline__ void stem_load_rows(float* sh, const T* __restrict__ x, int n, int h, int H, int W) {
    const int RP = stem_pitch(W, CIN);
    for (int i = threadIdx.x; i < 3 * (W + 2) * CIN; i += kStemThreads) {
        const int c = i % CIN, r = i / CIN;
        const int col = r % (W + 2), dy = r / (W + 2);
        const int ww = col - 1, hh = h + dy - 1;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f32(x[(((long long)n * H + hh) * W + ww) * CIN + c]);
        sh[dy * RP + col * CIN + c] = v;
    }
}

// window of pixels w .. w+3 (columns w-1 .. w+4) of the three rows -> xw[3][6 * CIN] by 16-byte loads
template <int CIN>
__device__ __forceinline__ void stem_window(const float* sh, int RP, int w, float (&xw)[3][6 * CIN + 2]) {
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const float4* p4 = reinterpret_cast<const float4*>(sh + dy * RP + w * CIN);
#pragma unroll
        for (int k = 0; k < (6 * CIN + 3) / 4; ++k) {
            const float4 v = p4[k];
            xw[dy][4 * k] = v.x; xw[dy][4 * k + 1] = v.y;
            if (4 * k + 2 < 6 * CIN + 2) { xw[dy][4 * k + 2] = v.z; xw[dy][4 * k + 3] = v.w; }
        }
    }
}

template <class T, int CIN>
__global__ void __launch_bounds__(kStemThreads) stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
                                                      T* __restrict__ y, int N, int H, int W, int Cout, int relu) {
    extern __shared__ __align__(16) float sh[];   // [3][pitch]
    const int RP = stem_pitch(W, CIN);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int cb = 0; cb < Cout; cb += 64) {
        float wr[9 * CIN][4], b4[4];
#pragma unroll
        for (int k = 0; k < 9 * CIN; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) wr[k][j] = to_f32(wp[(long long)k * Cout + cb + tx * 4 + j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) b4[j] = bias ? bias[cb + tx * 4 + j] : 0.f;
        for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
            const int n = row / H, h = row - n * H;
            __syncthreads();
            stem_load_rows<T, CIN>(sh, x, n, h, H, W);
            __syncthreads();
            T* yrow = y + ((long long)row * W) * Cout + cb + tx * 4;
            for (int w = ty * 4; w < W; w += kStemThreads / 4) {   // 4 adjacent pixels per thread: 16 independent FMA chains
                float xw[3][6 * CIN + 2];
                stem_window<CIN>(sh, RP, w, xw);
                float acc[4][4];
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[p][j] = b4[j];
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int c = 0; c < CIN; ++c)
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float xv = xw[t / 3][(p + t % 3) * CIN + c];
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[p][j] = fmaf(xv, wr[t * CIN + c][j], acc[p][j]);
                        }
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[p][j] = fmaxf(acc[p][j], 0.f);
                    }
                    T* dst = yrow + (long long)(w + p) * Cout;
                    if (sizeof(T) == 4) {
                        *reinterpret_cast<float4*>(dst) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
                    } else {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(acc[p][0], acc[p][1]), hi = __floats2bfloat162_rn(acc[p][2], acc[p][3]);
                        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    }
                }
            }
        }
    }
}

template <class T, int CIN>
static int launch_stem_fwd(const void* x, const void* wp, const float* bias, void* y, int N, int H, int W, int Cout, int relu,
                           cudaStream_t st) {
    size_t smem = sizeof(float) * 3 * (size_t)stem_pitch(W, CIN);
    if (smem > 200 * 1024 || W % 4 != 0) return 1;   // caller falls back to the generic engine
    static SmemOptIn configured;
    configured.ensure(stem_fwd_kernel<T, CIN>, 200 * 1024);
    int rows = N * H;
    int grid = rows < kNumSMs * 12 ? rows : kNumSMs * 12;
    stem_fwd_kernel<T, CIN><<<grid, kStemThreads, smem, st>>>((const T*)x, (const T*)wp, bias, (T*)y, N, H, W, Cout, relu);
    return check_launch("conv3x3_fwd(stem)");
}

// ------------------------------------------------------------------------------------ stem weight gradient
// conv3x3 wgrad for the 3(4)-channel input layer (K = 27: far too thin for a GEMM tile): one block walks image rows,
// the three input rows live in shared memory, a thread owns 4 output channels x all 9*CIN taps in registers.
template <class T, int CIN>
__global__ void __launch_bounds__(kStemThreads) stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dwp,
                                                        int N, int H, int W, int Cout) {
    extern __shared__ __align__(16) float sh[];
    const int RP = stem_pitch(W, CIN);
    float* sdw = sh + 3 * RP;                     // [9 * CIN][64]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int cb = 0; cb < Cout; cb += 64) {
        float acc[9 * CIN][4];
#pragma unroll
        for (int k = 0; k < 9 * CIN; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
        for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
            const int n = row / H, h = row - n * H;
            __syncthreads();
            stem_load_rows<T, CIN>(sh, x, n, h, H, W);
            __syncthreads();
            const T* drow = dy + ((long long)row * W) * Cout + cb + tx * 4;
            // the four dy vectors of the NEXT window are fetched while this one is multiplied (two 4-warp blocks per SM
            // cannot hide a DRAM round trip per window otherwise)
            typedef typename std::conditional<sizeof(T) == 4, float4, uint2>::type DV;
            DV dnext[4];
            if (ty * 4 < W) {
#pragma unroll
                for (int p = 0; p < 4; ++p) dnext[p] = *reinterpret_cast<const DV*>(drow + (long long)(ty * 4 + p) * Cout);
            }
            for (int w = ty * 4; w < W; w += kStemThreads / 4) {   // 4 adjacent pixels per thread
                float d[4][4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (sizeof(T) == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(&dnext[p]);
                        d[p][0] = v.x; d[p][1] = v.y; d[p][2] = v.z; d[p][3] = v.w;
                    } else {
                        const uint2 v = *reinterpret_cast<const uint2*>(&dnext[p]);
                        d[p][0] = __uint_as_float(v.x << 16); d[p][1] = __uint_as_float(v.x & 0xffff0000u);
                        d[p][2] = __uint_as_float(v.y << 16); d[p][3] = __uint_as_float(v.y & 0xffff0000u);
                    }
                }
                if (w + kStemThreads / 4 < W) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) dnext[p] = *reinterpret_cast<const DV*>(drow + (long long)(w + kStemThreads / 4 + p) * Cout);
                }
                float xw[3][6 * CIN + 2];
                stem_window<CIN>(sh, RP, w, xw);
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int c = 0; c < CIN; ++c)
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float xv = xw[t / 3][(p + t % 3) * CIN + c];
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[t * CIN + c][j] = fmaf(xv, d[p][j], acc[t * CIN + c][j]);
                        }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 9 * CIN * 64; i += kStemThreads) sdw[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 9 * CIN; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&sdw[k * 64 + tx * 4 + j], acc[k][j]);
        __syncthreads();
        for (int i = threadIdx.x; i < 9 * CIN * 64; i += kStemThreads) atomicAdd(dwp + (long long)(i / 64) * Cout + cb + (i % 64), sdw[i]);
    }
}

template <class T, int CIN>
static int launch_stem_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cout, cudaStream_t st) {
    size_t smem = sizeof(float) * (3 * (size_t)stem_pitch(W, CIN) + 9 * CIN * 64);
    if (smem > 200 * 1024 || W % 4 != 0) return 1;   // caller falls back to the generic engine
    static SmemOptIn configured;
    configured.ensure(stem_wgrad_kernel<T, CIN>, 200 * 1024);
    int rows = N * H;
    int grid = rows < kNumSMs * 3 ? rows : kNumSMs * 3;   // three 128-thread blocks (168 registers) are resident per SM: one wave
    stem_wgrad_kernel<T, CIN><<<grid, kStemThreads, smem, st>>>((const T*)x, (const T*)dy, dwp, N, H, W, Cout);
    return check_launch("conv3x3_wgrad(stem)");
}

}  // namespace eel

using namespace eel;

extern "C" {

int eel_conv3x3_fwd(const void* x, const void* wp, const float* bias, void* y, int N, int H, int W, int Cin,
                    int Cout, int relu, int flip, int dtype, eel_stream s) {
    EEL_REQUIRE(x && wp && y && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_fwd: bad argument");
    long long P = (long long)N * H * W;
    if ((Cin == 3 || Cin == 4) && Cout % 64 == 0 && !flip) {
        int rc = 1;
        cudaStream_t st = (cudaStream_t)s;
        if (dtype == EEL_F32) rc = Cin == 3 ? launch_stem_fwd<float, 3>(x, wp, bias, y, N, H, W, Cout, relu, st)
                                            : launch_stem_fwd<float, 4>(x, wp, bias, y, N, H, W, Cout, relu, st);
        else if (dtype == EEL_BF16) rc = Cin == 3 ? launch_stem_fwd<bf16, 3>(x, wp, bias, y, N, H, W, Cout, relu, st)
                                                  : launch_stem_fwd<bf16, 4>(x, wp, bias, y, N, H, W, Cout, relu, st);
        if (rc <= 0) return rc;
    }
    EEL_DISPATCH_DTYPE(dtype, {
        Conv3Acc<T> a{(const T*)x, N, H, W, Cin, flip};
        DenseAcc<T> b{(const T*)wp, 9LL * Cin, Cout, Cout};
        StoreEpi<T> e{(T*)y, P, Cout, Cout, bias, relu};
        return (run<Conv3Acc<T>, false, DenseAcc<T>, false, StoreEpi<T>>(a, b, e, P, Cout, 9LL * Cin, 1,
                                                                        (cudaStream_t)s, "conv3x3_fwd"));
    });
}

int eel_conv3x3_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout,
                      int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dwp && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_wgrad: bad argument");
    long long P = (long long)N * H * W;
    if (cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * Cin * Cout, (cudaStream_t)s) != cudaSuccess) {
        set_error("conv3x3_wgrad: memset failed");
        return EEL_ERR_CUDA;
    }
    if ((Cin == 3 || Cin == 4) && Cout % 64 == 0) {
        int rc = 1;
        if (dtype == EEL_F32) rc = Cin == 3 ? launch_stem_wgrad<float, 3>(x, dy, dwp, N, H, W, Cout, (cudaStream_t)s)
                                            : launch_stem_wgrad<float, 4>(x, dy, dwp, N, H, W, Cout, (cudaStream_t)s);
        else if (dtype == EEL_BF16) rc = Cin == 3 ? launch_stem_wgrad<bf16, 3>(x, dy, dwp, N, H, W, Cout, (cudaStream_t)s)
                                                  : launch_stem_wgrad<bf16, 4>(x, dy, dwp, N, H, W, Cout, (cudaStream_t)s);
        if (rc <= 0) return rc;
    }
    EEL_DISPATCH_DTYPE(dtype, {
        Conv3Acc<T> a{(const T*)x, N, H, W, Cin, 0};
        DenseAcc<T> b{(const T*)dy, P, Cout, Cout};
        AtomicEpi e{dwp, 9 * Cin, Cout, Cout};
        return (run<Conv3Acc<T>, true, DenseAcc<T>, false, AtomicEpi>(a, b, e, 9LL * Cin, Cout, P,
                                                                     pick_splits(9LL * Cin, Cout, P),
                                                                     (cudaStream_t)s, "conv3x3_wgrad"));
    });
}

int eel_convt2x2_fwd(const void* x, const void* wp, const float* bias, void* y, int N, int h, int w, int Cin,
                     int Cout, int dtype, eel_stream s) {
    EEL_REQUIRE(x && wp && y && N > 0 && h > 0 && w > 0 && Cin > 0 && Cout > 0, "convt2x2_fwd: bad argument");
    long long P = (long long)N * h * w;
    EEL_DISPATCH_DTYPE(dtype, {
        DenseAcc<T> a{(const T*)x, P, Cin, Cin};
        DenseAcc<T> b{(const T*)wp, Cin, 4 * Cout, 4LL * Cout};
        ConvTScatterEpi<T> e{(T*)y, bias, N, h, w, Cout};
        return (run<DenseAcc<T>, false, DenseAcc<T>, false, ConvTScatterEpi<T>>(a, b, e, P, 4LL * Cout, Cin, 1,
                                                                               (cudaStream_t)s, "convt2x2_fwd"));
    });
}

int eel_convt2x2_dgrad(const void* dy, const void* wp, void* dx, int N, int h, int w, int Cin, int Cout,
                       int dtype, eel_stream s) {
    EEL_REQUIRE(dy && wp && dx && N > 0 && h > 0 && w > 0 && Cin > 0 && Cout > 0, "convt2x2_dgrad: bad argument");
    long long P = (long long)N * h * w;
    EEL_DISPATCH_DTYPE(dtype, {
        ConvTAcc<T> a{(const T*)dy, N, h, w, Cout};
        DenseAcc<T> b{(const T*)wp, Cin, 4 * Cout, 4LL * Cout};   // B(k, n) = wp[n][k]
        StoreEpi<T> e{(T*)dx, P, Cin, Cin, nullptr, 0};
        return (run<ConvTAcc<T>, false, DenseAcc<T>, true, StoreEpi<T>>(a, b, e, P, Cin, 4LL * Cout, 1,
                                                                       (cudaStream_t)s, "convt2x2_dgrad"));
    });
}

int eel_convt2x2_wgrad(const void* x, const void* dy, float* dwp, int N, int h, int w, int Cin, int Cout,
                       int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dwp && N > 0 && h > 0 && w > 0 && Cin > 0 && Cout > 0, "convt2x2_wgrad: bad argument");
    long long P = (long long)N * h * w;
    if (cudaMemsetAsync(dwp, 0, sizeof(float) * 4 * Cin * Cout, (cudaStream_t)s) != cudaSuccess) {
        set_error("convt2x2_wgrad: memset failed");
        return EEL_ERR_CUDA;
    }
    EEL_DISPATCH_DTYPE(dtype, {
        DenseAcc<T> a{(const T*)x, P, Cin, Cin};
        ConvTAcc<T> b{(const T*)dy, N, h, w, Cout};
        AtomicEpi e{dwp, Cin, 4 * Cout, 4LL * Cout};
        return (run<DenseAcc<T>, true, ConvTAcc<T>, false, AtomicEpi>(a, b, e, Cin, 4LL * Cout, P,
                                                                     pick_splits(Cin, 4LL * Cout, P),
                                                                     (cudaStream_t)s, "convt2x2_wgrad"));
    });
}

static int check_shift(long long P, int K, int shiftH, int shiftW, const char* what) {
    if (shiftH < 0 || shiftW < 0 || (shiftH > 0) != (shiftW > 0)) {
        set_error("%s: shiftH/shiftW must both be 0 or both positive", what);
        return EEL_ERR_INVALID;
    }
    if (shiftH > 0 && (P % ((long long)shiftH * shiftW) != 0 || K % 4 != 0)) {
        set_error("%s: rows must be whole %dx%d images and K a multiple of 4 for the folded shift", what, shiftH, shiftW);
        return EEL_ERR_INVALID;
    }
    return EEL_OK;
}

int eel_linear_fwd(const void* x, const void* w, const float* bias, void* y, long long P, int K, int Nout,
                   int shiftH, int shiftW, int dtype, eel_stream s) {
    EEL_REQUIRE(x && w && y && P > 0 && K > 0 && Nout > 0, "linear_fwd: bad argument");
    if (int rc = check_shift(P, K, shiftH, shiftW, "linear_fwd")) return rc;
    EEL_DISPATCH_DTYPE(dtype, {
        DenseAcc<T> b{(const T*)w, Nout, K, K};   // B(k, n) = w[n][k]
        StoreEpi<T> e{(T*)y, P, Nout, Nout, bias, 0};
        if (shiftH > 0) {
            ShiftAcc<T> a{(const T*)x, P, shiftH, shiftW, K};
            return (run<ShiftAcc<T>, false, DenseAcc<T>, true, StoreEpi<T>>(a, b, e, P, Nout, K, 1, (cudaStream_t)s,
                                                                           "linear_fwd(shift)"));
        }
        DenseAcc<T> a{(const T*)x, P, K, K};
        return (run<DenseAcc<T>, false, DenseAcc<T>, true, StoreEpi<T>>(a, b, e, P, Nout, K, 1, (cudaStream_t)s,
                                                                       "linear_fwd"));
    });
}

int eel_linear_dgrad(const void* dy, const void* w, void* dx, long long P, int K, int Nout, int shiftH,
                     int shiftW, int dtype, eel_stream s) {
    EEL_REQUIRE(dy && w && dx && P > 0 && K > 0 && Nout > 0, "linear_dgrad: bad argument");
    if (int rc = check_shift(P, K, shiftH, shiftW, "linear_dgrad")) return rc;
    EEL_DISPATCH_DTYPE(dtype, {
        DenseAcc<T> a{(const T*)dy, P, Nout, Nout};
        DenseAcc<T> b{(const T*)w, Nout, K, K};   // B(k = nout, n = kin) = w[k][n]
        if (shiftH > 0) {
            ShiftScatterEpi<T> e{(T*)dx, P, shiftH, shiftW, K};
            return (run<DenseAcc<T>, false, DenseAcc<T>, false, ShiftScatterEpi<T>>(a, b, e, P, K, Nout, 1,
                                                                                   (cudaStream_t)s, "linear_dgrad(shift)"));
        }
        StoreEpi<T> e{(T*)dx, P, K, K, nullptr, 0};
        return (run<DenseAcc<T>, false, DenseAcc<T>, false, StoreEpi<T>>(a, b, e, P, K, Nout, 1, (cudaStream_t)s,
                                                                        "linear_dgrad"));
    });
}

int eel_linear_wgrad(const void* x, const void* dy, float* dw, long long P, int K, int Nout, int shiftH,
                     int shiftW, int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dw && P > 0 && K > 0 && Nout > 0, "linear_wgrad: bad argument");
    if (int rc = check_shift(P, K, shiftH, shiftW, "linear_wgrad")) return rc;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Nout * K, (cudaStream_t)s) != cudaSuccess) {
        set_error("linear_wgrad: memset failed");
        return EEL_ERR_CUDA;
    }
    EEL_DISPATCH_DTYPE(dtype, {
        DenseAcc<T> a{(const T*)dy, P, Nout, Nout};   // A(m = nout, k = p) = dy[p][m]
        AtomicEpi e{dw, Nout, K, K};
        int splits = pick_splits(Nout, K, P);
        if (shiftH > 0) {
            ShiftAcc<T> b{(const T*)x, P, shiftH, shiftW, K};
            return (run<DenseAcc<T>, true, ShiftAcc<T>, false, AtomicEpi>(a, b, e, Nout, K, P, splits, (cudaStream_t)s,
                                                                         "linear_wgrad(shift)"));
        }
        DenseAcc<T> b{(const T*)x, P, K, K};
        return (run<DenseAcc<T>, true, DenseAcc<T>, false, AtomicEpi>(a, b, e, Nout, K, P, splits, (cudaStream_t)s,
                                                                     "linear_wgrad"));
    });
}

}  // extern "C"
