// Weight-space composition of ChannelAwarePatchedMLP's last two layers.
//
// models/EELUnet.py:106-112,121-123: `mlp[2]` (nn.Linear 256 -> Cout) is followed by `to_space` (1x1 conv Cout -> Cout)
// with nothing in between, so per pixel
//
//     y = W2 (W1 g + b1) + b2 = (W2 W1) g + (W2 b1 + b2)            W1 = mlp[2].weight, W2 = to_space.weight
//
// The hot path therefore runs ONE pixel-space GEMM with Wc = W2 W1 (Cout x 256) instead of two (the Cout x Cout one
// being the larger), and its backward ONE data-gradient and ONE weight-gradient GEMM; the parameter gradients of the
// two reference layers follow exactly (linearity) from dWc = sum_p dy_p g_p^T and s = sum_p dy_p:
//
//     dW2 = dWc W1^T + s b1^T      dW1 = W2^T dWc      db1 = W2^T s      db2 = s
//
// These are weight-sized products (<= 1024 x 1024 x 256), done here in fp32 FFMA.  The forward composition of all
// eleven blocks is one launch per step (job table, like eel_pack_batch) and can fold an eval-mode BatchNorm in.
#include "common.cuh"

namespace eel {

struct ComposeJob {            // mirrors eel_compose_job in eel.h (128 bytes)
    const float* w2;           // [Cout][Cmid]
    const float* b2;           // [Cout]
    const float* w1;           // [Cmid][K]
    const float* b1;           // [Cmid]
    void* out_fwd;             // [Cout][K]   storage dtype, or null
    void* out_dgrad;           // [K][Cout]   storage dtype, or null
    float* bias_out;           // [Cout] fp32
    const float* rmean;        // optional eval-mode BatchNorm over Cout folded in (all four or none)
    const float* rvar;
    const float* gamma;
    const float* beta;
    int Cout, Cmid, K, dtype;
    float eps;
    int pad[5];
};
static_assert(sizeof(ComposeJob) == 128, "ComposeJob must stay 128 bytes (eel.h)");

constexpr int kCT = 64;        // output tile edge
constexpr int kCK = 16;        // reduction step
constexpr int kCThreads = 256; // 16 x 16 threads, 4 x 4 outputs each

// acc[i][j] += sum_k A(m0 + ty*4 + i, k) * B(k, n0 + tx*4 + j);  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
// One of each operand's strides is 1; the loader walks that index fastest.
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, long long sam, long long sak,
                                          const float* __restrict__ B, long long sbk, long long sbn, int M, int N, int Kd,
                                          int m0, int n0, float (*As)[kCT + 4], float (*Bs)[kCT + 4], float (&acc)[4][4]) {
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // the global loads of step k0 + 16 are in flight while step k0 is multiplied (these products are latency-bound:
    // a few dozen blocks, 16-64 steps each)
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m, k;
            if (sak == 1) { k = t & 15; m = (t >> 4) + 16 * i; } else { m = t & 63; k = (t >> 6) + 4 * i; }
            ra[i] = (m0 + m < M && k0 + k < Kd) ? A[(long long)(m0 + m) * sam + (long long)(k0 + k) * sak] : 0.f;
            int n, kb;
            if (sbn == 1) { n = t & 63; kb = (t >> 6) + 4 * i; } else { kb = t & 15; n = (t >> 4) + 16 * i; }
            rb[i] = (n0 + n < N && k0 + kb < Kd) ? B[(long long)(k0 + kb) * sbk + (long long)(n0 + n) * sbn] : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < Kd; k0 += kCK) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (sak == 1) As[t & 15][(t >> 4) + 16 * i] = ra[i]; else As[(t >> 6) + 4 * i][t & 63] = ra[i];
            if (sbn == 1) Bs[(t >> 6) + 4 * i][t & 63] = rb[i]; else Bs[t & 15][(t >> 4) + 16 * i] = rb[i];
        }
        __syncthreads();
        if (k0 + kCK < Kd) fetch(k0 + kCK);
#pragma unroll
        for (int kk = 0; kk < kCK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
}

template <class T> __device__ __forceinline__ void store_composed(const ComposeJob& j, int m, int n, float v) {
    if (j.out_fwd != nullptr) reinterpret_cast<T*>(j.out_fwd)[(long long)m * j.K + n] = from_f32<T>(v);
    if (j.out_dgrad != nullptr) reinterpret_cast<T*>(j.out_dgrad)[(long long)n * j.Cout + m] = from_f32<T>(v);
}

// grid = (work items, jobs): items 0 .. tiles-1 are 64 x 64 tiles of Wc, the items after them are 64-row slices of the bias
__global__ void __launch_bounds__(kCThreads) compose_batch_kernel(const ComposeJob* __restrict__ jobs) {
    __shared__ __align__(16) float As[kCK][kCT + 4];
    __shared__ __align__(16) float Bs[kCK][kCT + 4];
    const ComposeJob j = jobs[blockIdx.y];
    const int tm = (j.Cout + kCT - 1) / kCT, tn = (j.K + kCT - 1) / kCT;
    const bool fold = j.rmean != nullptr;
    for (int item = blockIdx.x; item < tm * tn + tm; item += gridDim.x) {
        if (item < tm * tn) {
            const int m0 = (item / tn) * kCT, n0 = (item % tn) * kCT;
            float acc[4][4];
            __syncthreads();
            tile_gemm(j.w2, j.Cmid, 1, j.w1, j.K, 1, j.Cout, j.K, j.Cmid, m0, n0, As, Bs, acc);
            const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty * 4 + i;
                if (m >= j.Cout) break;
                const float sc = fold ? j.gamma[m] / sqrtf(j.rvar[m] + j.eps) : 1.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + tx * 4 + q;
                    if (n >= j.K) break;
                    if (j.dtype == EEL_BF16) store_composed<bf16>(j, m, n, acc[i][q] * sc);
                    else store_composed<float>(j, m, n, acc[i][q] * sc);
                }
            }
        } else {
            // bias rows: one warp per row, lanes over Cmid
            const int m0 = (item - tm * tn) * kCT;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int r = warp; r < kCT; r += kCThreads / 32) {
                const int m = m0 + r;
                if (m >= j.Cout) break;
                float s = 0.f;
                for (int k = lane; k < j.Cmid; k += 32) s = fmaf(j.w2[(long long)m * j.Cmid + k], j.b1[k], s);
                s = warp_sum(s);
                if (lane == 0) {
                    float b = s + j.b2[m];
                    if (fold) b = (b - j.rmean[m]) * (j.gamma[m] / sqrtf(j.rvar[m] + j.eps)) + j.beta[m];
                    j.bias_out[m] = b;
                }
            }
        }
    }
}

// One launch for the three parameter-gradient products of a composed pair.  Work items:
//   [0, t2)        tiles of dW2[Cout][Cmid] = dWc W1^T + s b1^T      (reduction over K)
//   [t2, t2 + t1)  tiles of dW1[Cmid][K]    = W2^T dWc               (reduction over Cout)
//   the rest       64-row slices of db1[Cmid] = W2^T s
__global__ void __launch_bounds__(kCThreads) compose_bwd_kernel(const float* __restrict__ dwc, const float* __restrict__ s,
                                                              const float* __restrict__ w2, const float* __restrict__ w1,
                                                              const float* __restrict__ b1, float* __restrict__ dw2,
                                                              float* __restrict__ dw1, float* __restrict__ db1, int Cout,
                                                              int Cmid, int K) {
    __shared__ __align__(16) float As[kCK][kCT + 4];
    __shared__ __align__(16) float Bs[kCK][kCT + 4];
    const int to = (Cout + kCT - 1) / kCT, tmid = (Cmid + kCT - 1) / kCT, tk = (K + kCT - 1) / kCT;
    const int t2 = to * tmid, t1 = tmid * tk;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    for (int item = blockIdx.x; item < t2 + t1 + tmid; item += gridDim.x) {
        float acc[4][4];
        if (item < t2) {
            const int m0 = (item / tmid) * kCT, n0 = (item % tmid) * kCT;
            __syncthreads();
            tile_gemm(dwc, K, 1, w1, 1, K, Cout, Cmid, K, m0, n0, As, Bs, acc);      // B(k, n) = W1[n][k]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty * 4 + i;
                if (m >= Cout) break;
                const float sm = s[m];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + tx * 4 + q;
                    if (n < Cmid) dw2[(long long)m * Cmid + n] = fmaf(sm, b1[n], acc[i][q]);
                }
            }
        } else if (item < t2 + t1) {
            const int it = item - t2;
            const int m0 = (it / tk) * kCT, n0 = (it % tk) * kCT;
            __syncthreads();
            tile_gemm(w2, 1, Cmid, dwc, K, 1, Cmid, K, Cout, m0, n0, As, Bs, acc);    // A(m, k) = W2[k][m]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty * 4 + i;
                if (m >= Cmid) break;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + tx * 4 + q;
                    if (n < K) dw1[(long long)m * K + n] = acc[i][q];
                }
            }
        } else {
            // db1[m] = sum_i W2[i][m] s[i]: 64 columns per item, 4 row groups of 64 threads meet through shared memory
            const int m = (item - t2 - t1) * kCT + (threadIdx.x & 63);
            const int g = threadIdx.x >> 6;
            float v = 0.f;
            if (m < Cmid)
                for (int i = g; i < Cout; i += 4) v = fmaf(w2[(long long)i * Cmid + m], s[i], v);
            __syncthreads();
            As[g][threadIdx.x & 63] = v;
            __syncthreads();
            if (g == 0 && m < Cmid) db1[m] = As[0][threadIdx.x] + As[1][threadIdx.x] + As[2][threadIdx.x] + As[3][threadIdx.x];
        }
    }
}

}  // namespace eel

using namespace eel;

extern "C" {

int eel_compose_batch(const void* jobs_device, int njobs, int blocks_per_job, eel_stream s) {
    EEL_REQUIRE(jobs_device && njobs > 0 && blocks_per_job > 0, "compose_batch: bad argument");
    dim3 grid(blocks_per_job, njobs);
    compose_batch_kernel<<<grid, kCThreads, 0, (cudaStream_t)s>>>((const ComposeJob*)jobs_device);
    return check_launch("compose_batch");
}

int eel_compose_linear_bwd(const float* dwc, const float* colsum, const float* w2, const float* w1, const float* b1,
                           float* dw2, float* dw1, float* db1, int Cout, int Cmid, int K, eel_stream s) {
    EEL_REQUIRE(dwc && colsum && w2 && w1 && b1 && dw2 && dw1 && db1 && Cout > 0 && Cmid > 0 && K > 0, "compose_linear_bwd: bad argument");
    const int to = cdiv(Cout, kCT), tmid = cdiv(Cmid, kCT), tk = cdiv(K, kCT);
    const int items = to * tmid + tmid * tk + tmid;
    compose_bwd_kernel<<<items, kCThreads, 0, (cudaStream_t)s>>>(dwc, colsum, w2, w1, b1, dw2, dw1, db1, Cout, Cmid, K);
    return check_launch("compose_linear_bwd");
}

}  // extern "C"
