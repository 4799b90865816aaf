// Per-pixel fused ops: PredictionGuidedRefinement and the LayerNorm+1x1+sigmoid head, forward and
// backward.  A group of G = min(32, C/vec) lanes owns one pixel; its channel reductions are warp
// shuffles; parameter gradients are accumulated in registers, folded through shared memory and
// written as one partial row per block (summed in fp64 by finalize_partials).
#include "common.cuh"

namespace eel {

__device__ __forceinline__ float group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block = 32 columns x 32 row-slices (parallel over the per-block partial rows).  Column j of the finished row goes to the
// output segment that contains it (the parameter gradients are separate tensors: no device-to-device copies afterwards).
struct RowSegs { float* p[4]; int end[4]; };   // segment k holds columns [end[k-1], end[k])
__global__ void __launch_bounds__(1024) finalize_rows_kernel(const float* __restrict__ partial, int nrows, int width, RowSegs segs) {
    __shared__ double sm[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (j < width)
        for (int r = ty; r < nrows; r += 32 * 8) {    // 8 independent loads in flight
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = r + 32 * u < nrows ? partial[(long long)(r + 32 * u) * width + j] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < width) {
        double t = 0.0;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sm[y][tx];
        int k = 0, start = 0;
        while (k < 3 && j >= segs.end[k]) { start = segs.end[k]; ++k; }
        segs.p[k][j - start] = (float)t;
    }
}

constexpr int kPixThreads = 256;
constexpr int kMaxIter = 8;   // C <= 1024

// ------------------------------------------------------------------------------------ PGR
template <class T>
__global__ void __launch_bounds__(kPixThreads) pgr_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, T* __restrict__ y,
                                                            float* __restrict__ sg, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    const int nv = C / V, iters = nv / G;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const float bias = b[0];
    for (long long p0 = (long long)blockIdx.x * groups_per_block; p0 < P; p0 += (long long)gridDim.x * groups_per_block) {
        long long p = p0 + g_in_block;
        bool ok = p < P;
        float dot = 0.f;
        if (ok)
            for (int k = 0; k < iters; ++k) {
                int c0 = (gl + k * G) * V;
                Vec16<T> v = ld16(x + p * C + c0);
#pragma unroll
                for (int j = 0; j < V; ++j) dot += v.get(j) * w[c0 + j];
            }
        dot = group_sum(dot, G);
        float sv = sigmoidf_(dot + bias);
        if (ok) {
            if (gl == 0) sg[p] = sv;
            for (int k = 0; k < iters; ++k) {
                int c0 = (gl + k * G) * V;
                Vec16<T> v = ld16(x + p * C + c0), o;
#pragma unroll
                for (int j = 0; j < V; ++j) o.set(j, v.get(j) * (1.f + sv));
                st16(y + p * C + c0, o);
            }
        }
    }
}

// ITERS vectors per lane and pixel, U pixels per lane group in flight: x is loaded once into registers (U*ITERS independent
// 16-byte loads per thread) and reused for the gate and the output
template <class T, int ITERS, int U>
__global__ void __launch_bounds__(kPixThreads) pgr_fwd_kernel2(const T* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, T* __restrict__ y,
                                                             float* __restrict__ sg, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const float bias = b[0];
    float wreg[ITERS][V];
#pragma unroll
    for (int k = 0; k < ITERS; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) wreg[k][j] = w[(gl + k * G) * V + j];
    const long long stride = (long long)gridDim.x * groups_per_block * U;
    for (long long p0 = (long long)blockIdx.x * groups_per_block * U; p0 < P; p0 += stride) {
        Vec16<T> vx[U][ITERS];
        float dot[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            ok[u] = p < P;
            if (ok[u])
#pragma unroll
                for (int k = 0; k < ITERS; ++k) vx[u][k] = ld16(x + p * C + (gl + k * G) * V);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            dot[u] = 0.f;
            if (ok[u])
#pragma unroll
                for (int k = 0; k < ITERS; ++k)
#pragma unroll
                    for (int j = 0; j < V; ++j) dot[u] += vx[u][k].get(j) * wreg[k][j];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dot[u] = group_sum(dot[u], G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            const float sv = sigmoidf_(dot[u] + bias);
            if (gl == 0) sg[p] = sv;
#pragma unroll
            for (int k = 0; k < ITERS; ++k) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) o.set(j, vx[u][k].get(j) * (1.f + sv));
                st16(y + p * C + (gl + k * G) * V, o);
            }
        }
    }
}

template <class T>
__global__ void __launch_bounds__(kPixThreads) pgr_bwd_kernel(const T* __restrict__ x, const float* __restrict__ sg,
                                                            const float* __restrict__ w, const T* __restrict__ dy,
                                                            const float* __restrict__ dsg, T* __restrict__ dx,
                                                            float* __restrict__ partial, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ float sh[];   // [C + 1]
    const int nv = C / V, iters = nv / G;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    for (int i = threadIdx.x; i <= C; i += kPixThreads) sh[i] = 0.f;
    __syncthreads();
    float dwacc[kMaxIter][V];
#pragma unroll
    for (int k = 0; k < kMaxIter; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) dwacc[k][j] = 0.f;
    float dbacc = 0.f;
    for (long long p0 = (long long)blockIdx.x * groups_per_block; p0 < P; p0 += (long long)gridDim.x * groups_per_block) {
        long long p = p0 + g_in_block;
        bool ok = p < P;
        float dot = 0.f;
        if (ok)
#pragma unroll
            for (int k = 0; k < kMaxIter; ++k) {
                if (k < iters) {
                    int c0 = (gl + k * G) * V;
                    Vec16<T> vx = ld16(x + p * C + c0), vd = ld16(dy + p * C + c0);
#pragma unroll
                    for (int j = 0; j < V; ++j) dot += vx.get(j) * vd.get(j);
                }
            }
        dot = group_sum(dot, G);
        if (ok) {
            float sv = sg[p];
            float dg = (dot + (dsg ? dsg[p] : 0.f)) * sv * (1.f - sv);
            if (gl == 0) dbacc += dg;
#pragma unroll
            for (int k = 0; k < kMaxIter; ++k) {
                if (k < iters) {
                    int c0 = (gl + k * G) * V;
                    Vec16<T> vx = ld16(x + p * C + c0), vd = ld16(dy + p * C + c0), o;
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        o.set(j, vd.get(j) * (1.f + sv) + w[c0 + j] * dg);
                        dwacc[k][j] += vx.get(j) * dg;
                    }
                    st16(dx + p * C + c0, o);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxIter; ++k) {
        if (k < iters) {
            int c0 = (gl + k * G) * V;
#pragma unroll
            for (int j = 0; j < V; ++j) atomicAdd(&sh[c0 + j], dwacc[k][j]);
        }
    }
    if (gl == 0) atomicAdd(&sh[C], dbacc);
    __syncthreads();
    for (int i = threadIdx.x; i <= C; i += kPixThreads) partial[(long long)blockIdx.x * (C + 1) + i] = sh[i];
}

// Same op with ITERS (vectors per lane and pixel) a compile-time constant and U pixels per group in flight:
// every x / dy vector is loaded ONCE into registers (2*U*ITERS independent 16-byte loads per thread) before any
// arithmetic, which is what it takes to keep HBM busy at two blocks per SM.
template <class T, int ITERS, int U>
__global__ void __launch_bounds__(kPixThreads) pgr_bwd_kernel2(const T* __restrict__ x, const float* __restrict__ sg,
                                                             const float* __restrict__ w, const T* __restrict__ dy,
                                                             const float* __restrict__ dsg, T* __restrict__ dx,
                                                             float* __restrict__ partial, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ float sh[];   // [C + 1]
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    for (int i = threadIdx.x; i <= C; i += kPixThreads) sh[i] = 0.f;
    __syncthreads();
    float dwacc[ITERS][V], wreg[ITERS][V];
#pragma unroll
    for (int k = 0; k < ITERS; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) { dwacc[k][j] = 0.f; wreg[k][j] = w[(gl + k * G) * V + j]; }
    float dbacc = 0.f;
    const long long stride = (long long)gridDim.x * groups_per_block * U;
    for (long long p0 = (long long)blockIdx.x * groups_per_block * U; p0 < P; p0 += stride) {
        Vec16<T> vx[U][ITERS], vd[U][ITERS];
        float sv[U], ds[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            ok[u] = p < P;
            sv[u] = 0.f; ds[u] = 0.f;
            if (ok[u]) {
#pragma unroll
                for (int k = 0; k < ITERS; ++k) {
                    const int c0 = (gl + k * G) * V;
                    vx[u][k] = ld16(x + p * C + c0);
                    vd[u][k] = ld16(dy + p * C + c0);
                }
                sv[u] = sg[p];
                if (dsg) ds[u] = dsg[p];
            }
        }
        float dot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            dot[u] = 0.f;
            if (ok[u])
#pragma unroll
                for (int k = 0; k < ITERS; ++k)
#pragma unroll
                    for (int j = 0; j < V; ++j) dot[u] += vx[u][k].get(j) * vd[u][k].get(j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dot[u] = group_sum(dot[u], G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            const float dg = (dot[u] + ds[u]) * sv[u] * (1.f - sv[u]);
            if (gl == 0) dbacc += dg;
#pragma unroll
            for (int k = 0; k < ITERS; ++k) {
                const int c0 = (gl + k * G) * V;
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    o.set(j, vd[u][k].get(j) * (1.f + sv[u]) + wreg[k][j] * dg);
                    dwacc[k][j] += vx[u][k].get(j) * dg;
                }
                st16(dx + p * C + c0, o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int c0 = (gl + k * G) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) atomicAdd(&sh[c0 + j], dwacc[k][j]);
    }
    if (gl == 0) atomicAdd(&sh[C], dbacc);
    __syncthreads();
    for (int i = threadIdx.x; i <= C; i += kPixThreads) partial[(long long)blockIdx.x * (C + 1) + i] = sh[i];
}

// ---- BatchNorm + ReLU + PredictionGuidedRefinement fused (the end of every decoder block, models/EELUnet.py:343-344 then
// :200-203).  Forward: the input is the PRE-BatchNorm tensor z; x = relu(gamma * (z - mean) * rstd + beta) exists only in
// registers.  Backward: x is recomputed from z, and the kernel also accumulates the two per-channel sums the BatchNorm
// backward needs (sum g, sum g * xhat with g = dx * relu_mask), so that BatchNorm's own reduction pass over (dy, z)
// disappears -- eel_bn_act_bwd_apply finishes with one pass.
struct BnConst { const float* mean; const float* rstd; const float* gamma; const float* beta; };

// sums = {sum g, sum g * z} -> {sum g, sum g * xhat}
__global__ void bn_raw_sums_fix_kernel(float* __restrict__ sums, const float* __restrict__ mean, const float* __restrict__ rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    sums[C + c] = rstd[c] * (sums[C + c] - mean[c] * sums[c]);
}

template <class T, int ITERS, int U>
__global__ void __launch_bounds__(kPixThreads) bn_pgr_fwd_kernel(const T* __restrict__ z, BnConst bn, const float* __restrict__ w,
                                                               const float* __restrict__ b, T* __restrict__ y,
                                                               float* __restrict__ sg, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const float bias = b[0];
    // x = relu(sc * z + sh): two constants per channel (one FMA per element; fewer registers leave room for U pixels in flight)
    float wreg[ITERS][V], sc[ITERS][V], sh[ITERS][V];
#pragma unroll
    for (int k = 0; k < ITERS; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int c = (gl + k * G) * V + j;
            wreg[k][j] = w[c];
            sc[k][j] = bn.gamma[c] * bn.rstd[c];
            sh[k][j] = bn.beta[c] - bn.mean[c] * sc[k][j];
        }
    const long long stride = (long long)gridDim.x * groups_per_block * U;
    for (long long p0 = (long long)blockIdx.x * groups_per_block * U; p0 < P; p0 += stride) {
        Vec16<T> vz[U][ITERS];
        float dot[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            ok[u] = p < P;
            if (ok[u])
#pragma unroll
                for (int k = 0; k < ITERS; ++k) vz[u][k] = ld16(z + p * C + (gl + k * G) * V);
        }
        float xf[U][ITERS][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            dot[u] = 0.f;
#pragma unroll
            for (int k = 0; k < ITERS; ++k)
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float x = ok[u] ? fmaxf(fmaf(vz[u][k].get(j), sc[k][j], sh[k][j]), 0.f) : 0.f;
                    xf[u][k][j] = x;
                    dot[u] = fmaf(x, wreg[k][j], dot[u]);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dot[u] = group_sum(dot[u], G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            const float sv = sigmoidf_(dot[u] + bias);
            if (gl == 0) sg[p] = sv;
#pragma unroll
            for (int k = 0; k < ITERS; ++k) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) o.set(j, xf[u][k][j] * (1.f + sv));
                st16(y + p * C + (gl + k * G) * V, o);
            }
        }
    }
}

// partial row layout: dw[C] db[1] sum_g[C] sum_gx[C]
template <class T, int ITERS, int U>
__global__ void __launch_bounds__(kPixThreads) bn_pgr_bwd_kernel(const T* __restrict__ z, BnConst bn, const float* __restrict__ sg,
                                                               const float* __restrict__ w, const T* __restrict__ dy,
                                                               const float* __restrict__ dsg, T* __restrict__ dx,
                                                               float* __restrict__ partial, long long P, int C, int G) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ float sh[];   // [3 C + 1]
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const int width = 3 * C + 1;
    for (int i = threadIdx.x; i < width; i += kPixThreads) sh[i] = 0.f;
    __syncthreads();
    // sums over raw z (sgx = sum g * z; the finalize turns it into sum g * xhat = rstd * (sum g z - mean * sum g)) and
    // x = relu(sc * z + sh): the kernel was instruction-bound with the four-constant form
    float dwacc[ITERS][V], sga[ITERS][V], sgx[ITERS][V], wreg[ITERS][V], sc[ITERS][V], shf[ITERS][V];
#pragma unroll
    for (int k = 0; k < ITERS; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int c = (gl + k * G) * V + j;
            dwacc[k][j] = 0.f; sga[k][j] = 0.f; sgx[k][j] = 0.f;
            wreg[k][j] = w[c];
            sc[k][j] = bn.gamma[c] * bn.rstd[c];
            shf[k][j] = bn.beta[c] - bn.mean[c] * sc[k][j];
        }
    float dbacc = 0.f;
    const long long stride = (long long)gridDim.x * groups_per_block * U;
    for (long long p0 = (long long)blockIdx.x * groups_per_block * U; p0 < P; p0 += stride) {
        Vec16<T> vz[U][ITERS], vd[U][ITERS];
        float sv[U], ds[U], dot[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            ok[u] = p < P;
            sv[u] = 0.f; ds[u] = 0.f;
            if (ok[u]) {
#pragma unroll
                for (int k = 0; k < ITERS; ++k) {
                    const int c0 = (gl + k * G) * V;
                    vz[u][k] = ld16(z + p * C + c0);
                    vd[u][k] = ld16(dy + p * C + c0);
                }
                sv[u] = sg[p];
                if (dsg) ds[u] = dsg[p];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            dot[u] = 0.f;
            if (ok[u])
#pragma unroll
                for (int k = 0; k < ITERS; ++k)
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const float x = fmaxf(fmaf(vz[u][k].get(j), sc[k][j], shf[k][j]), 0.f);
                        dot[u] = fmaf(x, vd[u][k].get(j), dot[u]);
                    }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dot[u] = group_sum(dot[u], G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long p = p0 + (long long)u * groups_per_block + g_in_block;
            const float dg = (dot[u] + ds[u]) * sv[u] * (1.f - sv[u]);
            if (gl == 0) dbacc += dg;
#pragma unroll
            for (int k = 0; k < ITERS; ++k) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float zz = vz[u][k].get(j);
                    const float pre = fmaf(zz, sc[k][j], shf[k][j]);
                    const float x = fmaxf(pre, 0.f);
                    o.set(j, fmaf(vd[u][k].get(j), 1.f + sv[u], wreg[k][j] * dg));
                    dwacc[k][j] = fmaf(x, dg, dwacc[k][j]);
                    const float g = pre > 0.f ? o.get(j) : 0.f;     // the BatchNorm backward sees the STORED (rounded) dx
                    sga[k][j] += g;
                    sgx[k][j] = fmaf(g, zz, sgx[k][j]);
                }
                st16(dx + p * C + (gl + k * G) * V, o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int c0 = (gl + k * G) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            atomicAdd(&sh[c0 + j], dwacc[k][j]);
            atomicAdd(&sh[C + 1 + c0 + j], sga[k][j]);
            atomicAdd(&sh[2 * C + 1 + c0 + j], sgx[k][j]);
        }
    }
    if (gl == 0) atomicAdd(&sh[C], dbacc);
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += kPixThreads) partial[(long long)blockIdx.x * width + i] = sh[i];
}

// ------------------------------------------------------------------------------------ head (C = 64)
constexpr int kHeadC = 64;
constexpr int kHeadMaxO = 4;

template <class T>
__global__ void __launch_bounds__(kPixThreads) head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ lnw,
                                                             const float* __restrict__ lnb, const float* __restrict__ w,
                                                             const float* __restrict__ b, float* __restrict__ prob,
                                                             long long P, long long HW, int O) {
    constexpr int V = Vec16<T>::N, G = kHeadC / V;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const int c0 = gl * V;
#pragma unroll 4
    for (long long p0 = (long long)blockIdx.x * groups_per_block; p0 < P; p0 += (long long)gridDim.x * groups_per_block) {
        long long p = p0 + g_in_block;
        bool ok = p < P;
        float xv[V];
        float s = 0.f;
        if (ok) {
            Vec16<T> v = ld16(x + p * kHeadC + c0);
#pragma unroll
            for (int j = 0; j < V; ++j) { xv[j] = v.get(j); s += xv[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) xv[j] = 0.f;
        }
        float mu = group_sum(s, G) * (1.f / kHeadC);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) { float d = xv[j] - mu; q += d * d; }
        float var = group_sum(q, G) * (1.f / kHeadC);
        float r = 1.f / sqrtf(var + 1e-6f);
        float nv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) nv[j] = lnw[c0 + j] * ((xv[j] - mu) * r) + lnb[c0 + j];
        for (int o = 0; o < O; ++o) {
            float d = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) d += w[o * kHeadC + c0 + j] * nv[j];
            d = group_sum(d, G);
            if (ok && gl == 0) {
                // P < 2^31 (checked by the host): a 32-bit division instead of the ~100-instruction 64-bit one per pixel
                const unsigned nu = (unsigned)p / (unsigned)HW;
                const long long n = nu, hw = (long long)((unsigned)p - nu * (unsigned)HW);
                prob[(n * O + o) * HW + hw] = sigmoidf_(d + b[o]);
            }
        }
    }
}

// partial row layout: dlnw[64] dlnb[64] dw[O][64] db[O]
// MO = compile-time bound on the output channels O (1 for the reference's binary segmentation: the per-output register
// arrays then cost 9 registers instead of 36, which is the difference between 2 and 3-4 resident blocks per SM)
template <class T, int MO>
__global__ void __launch_bounds__(kPixThreads, 3) head_bwd_kernel(const T* __restrict__ x, const float* __restrict__ lnw,
                                                             const float* __restrict__ lnb, const float* __restrict__ w,
                                                             const float* __restrict__ prob, const float* __restrict__ dprob,
                                                             T* __restrict__ dx, float* __restrict__ partial, long long P,
                                                             long long HW, int O) {
    constexpr int V = Vec16<T>::N, G = kHeadC / V;
    extern __shared__ float sh[];
    const int width = (2 + O) * kHeadC + O;
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int groups_per_block = kPixThreads / G;
    const int g_in_block = (threadIdx.x >> 5) * (32 / G) + gi;
    const int c0 = gl * V;
    for (int i = threadIdx.x; i < width; i += kPixThreads) sh[i] = 0.f;
    __syncthreads();
    float aw[V], ab[V], awo[MO][V], abo[MO];
#pragma unroll
    for (int j = 0; j < V; ++j) { aw[j] = 0.f; ab[j] = 0.f; }
#pragma unroll
    for (int o = 0; o < MO; ++o) {
        abo[o] = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) awo[o][j] = 0.f;
    }
    for (long long p0 = (long long)blockIdx.x * groups_per_block; p0 < P; p0 += (long long)gridDim.x * groups_per_block) {
        long long p = p0 + g_in_block;
        bool ok = p < P;
        float xv[V];
        float s = 0.f;
        if (ok) {
            Vec16<T> v = ld16(x + p * kHeadC + c0);
#pragma unroll
            for (int j = 0; j < V; ++j) { xv[j] = v.get(j); s += xv[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) xv[j] = 0.f;
        }
        float mu = group_sum(s, G) * (1.f / kHeadC);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) { float d = xv[j] - mu; q += d * d; }
        float var = group_sum(q, G) * (1.f / kHeadC);
        float r = 1.f / sqrtf(var + 1e-6f);
        float xh[V], dn[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { xh[j] = (xv[j] - mu) * r; dn[j] = 0.f; }
        const unsigned nu = ok ? (unsigned)p / (unsigned)HW : 0u;      // P < 2^31 (checked by the host)
        const long long n = nu, hw = ok ? (long long)((unsigned)p - nu * (unsigned)HW) : 0;
#pragma unroll
        for (int o = 0; o < MO; ++o) {
            if (o < O) {
                float dl = 0.f;
                if (ok) {
                    float pr = prob[(n * O + o) * HW + hw];
                    dl = dprob[(n * O + o) * HW + hw] * pr * (1.f - pr);
                }
                if (gl == 0) abo[o] += dl;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    dn[j] += w[o * kHeadC + c0 + j] * dl;
                    awo[o][j] += dl * (lnw[c0 + j] * xh[j] + lnb[c0 + j]);
                }
            }
        }
        float m1 = 0.f, m2 = 0.f;
        float dxh[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            aw[j] += dn[j] * xh[j];
            ab[j] += dn[j];
            dxh[j] = dn[j] * lnw[c0 + j];
            m1 += dxh[j];
            m2 += dxh[j] * xh[j];
        }
        m1 = group_sum(m1, G) * (1.f / kHeadC);
        m2 = group_sum(m2, G) * (1.f / kHeadC);
        if (ok) {
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < V; ++j) o.set(j, r * (dxh[j] - m1 - xh[j] * m2));
            st16(dx + p * kHeadC + c0, o);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        atomicAdd(&sh[c0 + j], aw[j]);
        atomicAdd(&sh[kHeadC + c0 + j], ab[j]);
    }
#pragma unroll
    for (int o = 0; o < MO; ++o) {
        if (o < O) {
#pragma unroll
            for (int j = 0; j < V; ++j) atomicAdd(&sh[(2 + o) * kHeadC + c0 + j], awo[o][j]);
            if (gl == 0) atomicAdd(&sh[(2 + O) * kHeadC + o], abo[o]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += kPixThreads) partial[(long long)blockIdx.x * width + i] = sh[i];
}

// ---- head forward, one THREAD per pixel (O == 1) -------------------------------------------------------------------------
// (measured inside the step, where the input was just written and partly sits in L2, this simpler kernel beats the warp-tile
// forward below: 0.16 against 0.19 ms; the backward is the other way round, 0.47 against 0.39 ms)
constexpr int kTpThreads = 128;

template <class T> __device__ __forceinline__ void tp_load64(const T* __restrict__ src, float (&v)[kHeadC]) {
    constexpr int V = Vec16<T>::N;
#pragma unroll
    for (int k = 0; k < kHeadC / V; ++k) {
        const Vec16<T> q = ld16(src + k * V);
#pragma unroll
        for (int j = 0; j < V; ++j) v[k * V + j] = q.get(j);
    }
}

template <class T>
__global__ void __launch_bounds__(kTpThreads) head_fwd_tp_kernel(const T* __restrict__ x, const float* __restrict__ lnw,
                                                               const float* __restrict__ lnb, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ prob, long long P) {
    __shared__ float swl[kHeadC];   // w[c] * lnw[c]
    __shared__ float sc[2];         // sum_c w[c] * lnb[c] + b,  unused
    if (threadIdx.x < kHeadC) swl[threadIdx.x] = w[threadIdx.x] * lnw[threadIdx.x];
    if (threadIdx.x == 0) {
        float t = b[0];
        for (int c = 0; c < kHeadC; ++c) t += w[c] * lnb[c];
        sc[0] = t;
    }
    __syncthreads();
    const float cb = sc[0];
    for (long long p = blockIdx.x * (long long)kTpThreads + threadIdx.x; p < P; p += (long long)gridDim.x * kTpThreads) {
        float v[kHeadC];
        tp_load64<T>(x + p * kHeadC, v);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kHeadC; ++c) s += v[c];
        const float mu = s * (1.f / kHeadC);
        float q = 0.f, d = 0.f;
#pragma unroll
        for (int c = 0; c < kHeadC; ++c) {
            const float t = v[c] - mu;
            q = fmaf(t, t, q);
            d = fmaf(swl[c], t, d);
        }
        const float r = 1.f / sqrtf(q * (1.f / kHeadC) + 1e-6f);
        prob[p] = sigmoidf_(fmaf(d, r, cb));       // sum_c w (lnw xhat + lnb) + b
    }
}

// ---- head, one WARP per 32-pixel tile (O == 1, the reference's binary segmentation) -----------------------------------
// The lane-group kernels above spend most of their issue slots on shuffles and on per-lane copies of scalar work (ncu:
// 67 % issue-active at 23 % of DRAM bandwidth).  With a pixel's 64 channels in ONE thread's registers LayerNorm needs no
// shuffle at all, and for O == 1 every parameter gradient follows from S0 = sum_p dl_p and S1[c] = sum_p dl_p * xhat_pc:
//     dlnw[c] = w[c] S1[c],  dlnb[c] = w[c] S0,  dw[c] = lnw[c] S1[c] + lnb[c] S0,  db = S0      (dl = dprob * p (1 - p))
// A first version gave each thread its pixel straight from global memory: 128 (256) contiguous bytes per lane, so every
// 16-byte access of a warp touched 32 different lines, and the backward's 64 S1 accumulators per thread left 12 warps per
// SM (36 % of the HBM roofline, 0.47 ms at 64 x 256^2).  Here a warp moves its 32 pixels as 16-byte
// vectors in lane order (whole 512-byte runs per access) through a shared-memory tile [32][68] fp32 (rows 16-byte aligned,
// 128-bit accesses conflict-free both ways): phase A stores the vectors, phase B is the thread-per-pixel LayerNorm arithmetic
// on the pixel's 64 values in registers, phase C walks the tile in vector order again for the dx stores and for S1, of which
// a lane now only accumulates the 8 (4) channels of its vector slot.  The next tile's vectors are loaded one round ahead.
constexpr int kHwThreads = 128;               // 4 warps, one tile each per round
constexpr int kHwRow = kHeadC + 4;

template <class T> __device__ __forceinline__ void hw_store_vec(float* row_slot, const Vec16<T>& v, bool ok) {
    constexpr int V = Vec16<T>::N;
#pragma unroll
    for (int k = 0; k < V / 4; ++k)
        *reinterpret_cast<float4*>(row_slot + 4 * k) = ok ? make_float4(v.get(4 * k), v.get(4 * k + 1), v.get(4 * k + 2), v.get(4 * k + 3))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
}

// partial row layout: dlnw[64] dlnb[64] dw[64] db[1]
template <class T>
__global__ void __launch_bounds__(kHwThreads, 3) head_bwd_warp_kernel(const T* __restrict__ x, const float* __restrict__ lnw,
                                                                 const float* __restrict__ lnb, const float* __restrict__ w,
                                                                 const float* __restrict__ prob, const float* __restrict__ dprob,
                                                                 T* __restrict__ dx, float* __restrict__ partial, long long P) {
    constexpr int V = Vec16<T>::N, VPP = kHeadC / V;
    __shared__ __align__(16) float tile[kHwThreads / 32][32 * kHwRow];
    __shared__ float pk[kHwThreads / 32][3][32];        // per pixel of a warp's tile: dl, r * dl, m2
    __shared__ __align__(16) float swl[kHeadC];
    __shared__ float sacc[kHeadC + 1];                  // S1[c], S0
    __shared__ float swm;
    if (threadIdx.x < kHeadC) swl[threadIdx.x] = w[threadIdx.x] * lnw[threadIdx.x];
    if (threadIdx.x <= kHeadC) sacc[threadIdx.x] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int c = 0; c < kHeadC; ++c) t += swl[c];
        swm = t * (1.f / kHeadC);
    }
    __syncthreads();
    const float wm = swm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tl = tile[warp];
    float (*pw)[32] = pk[warp];
    const int slot = lane % VPP;                        // (32 is a multiple of VPP: a lane meets the same slot in every round)
    float swl_l[V];
#pragma unroll
    for (int j = 0; j < V; ++j) swl_l[j] = swl[slot * V + j] - wm;
    float S1[V], S0 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) S1[j] = 0.f;
    const long long ntiles = (P + 31) / 32;
    const long long tstep = (long long)gridDim.x * (kHwThreads / 32);
    long long t = (long long)blockIdx.x * (kHwThreads / 32) + warp;
    Vec16<T> v[VPP];
    float pr_n = 0.f, dp_n = 0.f;
    auto load_tile = [&](long long tt) {
        const long long q0 = tt * 32;
#pragma unroll
        for (int it = 0; it < VPP; ++it) {
            const int vi = it * 32 + lane;
            if (q0 + vi / VPP < P) v[it] = ld16(x + q0 * kHeadC + (long long)vi * V);
        }
        const bool ok = q0 + lane < P;
        pr_n = ok ? prob[q0 + lane] : 0.f;
        dp_n = ok ? dprob[q0 + lane] : 0.f;
    };
    if (t < ntiles) load_tile(t);
    for (; t < ntiles; t += tstep) {
        const long long p0 = t * 32;
        // ---- A: the tile's vectors (loaded in lane order one round ahead) go to shared memory
        const float dl = dp_n * pr_n * (1.f - pr_n);
#pragma unroll
        for (int it = 0; it < VPP; ++it) {
            const int pix = (it * 32 + lane) / VPP;
            hw_store_vec<T>(tl + pix * kHwRow + slot * V, v[it], p0 + pix < P);
        }
        if (t + tstep < ntiles) load_tile(t + tstep);      // the next tile's vectors travel while this one is processed
        __syncwarp();
        // ---- B: LayerNorm statistics of pixel `lane` in registers; the tile row becomes xhat
        {
            float xr[kHeadC];
#pragma unroll
            for (int k = 0; k < kHeadC / 4; ++k) {
                const float4 q4 = *reinterpret_cast<const float4*>(tl + lane * kHwRow + 4 * k);
                xr[4 * k] = q4.x; xr[4 * k + 1] = q4.y; xr[4 * k + 2] = q4.z; xr[4 * k + 3] = q4.w;
            }
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < kHeadC; ++c) s += xr[c];
            const float mu = s * (1.f / kHeadC);
            float q = 0.f;
#pragma unroll
            for (int c = 0; c < kHeadC; ++c) { xr[c] -= mu; q = fmaf(xr[c], xr[c], q); }
            const float r = 1.f / sqrtf(q * (1.f / kHeadC) + 1e-6f);
            float m2 = 0.f;
#pragma unroll
            for (int k = 0; k < kHeadC / 4; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(swl + 4 * k);
                const float4 h4 = make_float4(xr[4 * k] * r, xr[4 * k + 1] * r, xr[4 * k + 2] * r, xr[4 * k + 3] * r);
                m2 = fmaf(w4.x, h4.x, m2); m2 = fmaf(w4.y, h4.y, m2); m2 = fmaf(w4.z, h4.z, m2); m2 = fmaf(w4.w, h4.w, m2);
                *reinterpret_cast<float4*>(tl + lane * kHwRow + 4 * k) = h4;
            }
            pw[0][lane] = dl;
            pw[1][lane] = r * dl;
            pw[2][lane] = m2 * (1.f / kHeadC);
            S0 += dl;
        }
        __syncwarp();
        // ---- C: dx in vector order; S1 over the lane's slot
        T* dst = dx + p0 * kHeadC;
#pragma unroll
        for (int it = 0; it < VPP; ++it) {
            const int vi = it * 32 + lane, pix = vi / VPP;
            const float dlp = pw[0][pix], kk = pw[1][pix], mm = pw[2][pix];
            float xh[V];
#pragma unroll
            for (int k = 0; k < V / 4; ++k) {
                const float4 q4 = *reinterpret_cast<const float4*>(tl + pix * kHwRow + slot * V + 4 * k);
                xh[4 * k] = q4.x; xh[4 * k + 1] = q4.y; xh[4 * k + 2] = q4.z; xh[4 * k + 3] = q4.w;
            }
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                S1[j] = fmaf(dlp, xh[j], S1[j]);
                o.set(j, kk * fmaf(-xh[j], mm, swl_l[j]));
            }
            if (p0 + pix < P) st16(dst + (long long)vi * V, o);
        }
        __syncwarp();
    }
    // lanes with the same slot hold partial sums of the same channels (lane bits above log2(VPP))
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float t2 = S1[j];
#pragma unroll
        for (int o = VPP; o < 32; o <<= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        if (lane < VPP) atomicAdd(&sacc[slot * V + j], t2);
    }
    {
        const float t2 = warp_sum(S0);
        if (lane == 0) atomicAdd(&sacc[kHeadC], t2);
    }
    __syncthreads();
    float* row = partial + (long long)blockIdx.x * (3 * kHeadC + 1);
    if (threadIdx.x < kHeadC) {
        const int c = threadIdx.x;
        const float s1 = sacc[c], s0 = sacc[kHeadC];
        row[c] = w[c] * s1;
        row[kHeadC + c] = w[c] * s0;
        row[2 * kHeadC + c] = lnw[c] * s1 + lnb[c] * s0;
        if (c == 0) row[3 * kHeadC] = s0;
    }
}

static int pix_group(int C, int V) {
    int nv = C / V;
    return nv >= 32 ? 32 : nv;
}
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace eel

using namespace eel;

extern "C" {

int eel_pgr_fwd(const void* x, const float* w, const float* b, void* y, float* sgm, long long P, int C, int dtype,
                eel_stream s) {
    EEL_REQUIRE(x && w && b && y && sgm && P > 0 && C > 0, "pgr_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(C % V == 0 && pow2(C / V) && C / V <= 32 * kMaxIter, "pgr_fwd: C/vec must be a power of two <= 256");
        int G = pix_group(C, V);
        int gpb = kPixThreads / G;
        long long blocks = (P + gpb - 1) / gpb;
        int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
        const int iters = C / V / G;
        cudaStream_t st2 = (cudaStream_t)s;
        if (iters == 1) pgr_fwd_kernel2<T, 1, 4><<<grid, kPixThreads, 0, st2>>>((const T*)x, w, b, (T*)y, sgm, P, C, G);
        else if (iters == 2) pgr_fwd_kernel2<T, 2, 2><<<grid, kPixThreads, 0, st2>>>((const T*)x, w, b, (T*)y, sgm, P, C, G);
        else if (iters == 4) pgr_fwd_kernel2<T, 4, 1><<<grid, kPixThreads, 0, st2>>>((const T*)x, w, b, (T*)y, sgm, P, C, G);
        else pgr_fwd_kernel<T><<<grid, kPixThreads, 0, st2>>>((const T*)x, w, b, (T*)y, sgm, P, C, G);
        return check_launch("pgr_fwd");
    });
}

int eel_pgr_bwd(const void* x, const float* sgm, const float* w, const void* dy, const float* dsgm, void* dx, float* dw,
                float* db, long long P, int C, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(x && sgm && w && dy && dx && dw && db && P > 0 && C > 0, "pgr_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(C % V == 0 && pow2(C / V) && C / V <= 32 * kMaxIter, "pgr_bwd: C/vec must be a power of two <= 256");
        int G = pix_group(C, V);
        int gpb = kPixThreads / G;
        long long blocks = (P + gpb - 1) / gpb;
        int grid = (int)(blocks < (long long)kNumSMs * 2 ? blocks : (long long)kNumSMs * 2);
        size_t need = sizeof(float) * ((size_t)grid + 1) * (C + 1);
        if (need > ws_bytes || !ws) { set_error("pgr_bwd: workspace too small (%zu > %zu)", need, ws_bytes); return EEL_ERR_WORKSPACE; }
        float* partial = (float*)ws;
        const int iters = C / V / G;
        const size_t shb = sizeof(float) * (C + 1);
        cudaStream_t st2 = (cudaStream_t)s;
        if (iters == 1) pgr_bwd_kernel2<T, 1, 4><<<grid, kPixThreads, shb, st2>>>((const T*)x, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        else if (iters == 2) pgr_bwd_kernel2<T, 2, 2><<<grid, kPixThreads, shb, st2>>>((const T*)x, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        else if (iters == 4) pgr_bwd_kernel2<T, 4, 1><<<grid, kPixThreads, shb, st2>>>((const T*)x, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        else pgr_bwd_kernel<T><<<grid, kPixThreads, shb, st2>>>((const T*)x, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        if (int rc = check_launch("pgr_bwd")) return rc;
        RowSegs segs{{dw, db, db, db}, {C, C + 1, C + 1, C + 1}};
        finalize_rows_kernel<<<cdiv(C + 1, 32), 1024, 0, (cudaStream_t)s>>>(partial, grid, C + 1, segs);
        return check_launch("pgr_bwd.finalize");
    });
}

int eel_bn_pgr_fwd(const void* z, const float* bn_mean, const float* bn_rstd, const float* bn_gamma, const float* bn_beta,
                   const float* w, const float* b, void* y, float* sgm, long long P, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(z && bn_mean && bn_rstd && bn_gamma && bn_beta && w && b && y && sgm && P > 0 && C > 0, "bn_pgr_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(C % V == 0 && pow2(C / V), "bn_pgr_fwd: C/vec must be a power of two");
        const int G = pix_group(C, V), iters = C / V / G;
        EEL_REQUIRE(iters == 1 || iters == 2, "bn_pgr_fwd: at most 64 channel vectors (C <= %d)", 64 * V);
        const int gpb = kPixThreads / G;
        const long long blocks = (P + gpb - 1) / gpb;
        const int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
        const BnConst bn{bn_mean, bn_rstd, bn_gamma, bn_beta};
        cudaStream_t st2 = (cudaStream_t)s;
        if (iters == 1) bn_pgr_fwd_kernel<T, 1, 4><<<grid, kPixThreads, 0, st2>>>((const T*)z, bn, w, b, (T*)y, sgm, P, C, G);
        else bn_pgr_fwd_kernel<T, 2, 2><<<grid, kPixThreads, 0, st2>>>((const T*)z, bn, w, b, (T*)y, sgm, P, C, G);
        return check_launch("bn_pgr_fwd");
    });
}

int eel_bn_pgr_bwd(const void* z, const float* bn_mean, const float* bn_rstd, const float* bn_gamma, const float* bn_beta,
                   const float* sgm, const float* w, const void* dy, const float* dsgm, void* dx, float* dw, float* db,
                   float* bn_sums, long long P, int C, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(z && bn_mean && bn_rstd && bn_gamma && bn_beta && sgm && w && dy && dx && dw && db && bn_sums && P > 0 && C > 0,
                "bn_pgr_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(C % V == 0 && pow2(C / V), "bn_pgr_bwd: C/vec must be a power of two");
        const int G = pix_group(C, V), iters = C / V / G;
        EEL_REQUIRE(iters == 1 || iters == 2, "bn_pgr_bwd: at most 64 channel vectors (C <= %d)", 64 * V);
        const int gpb = kPixThreads / G;
        const long long blocks = (P + gpb - 1) / gpb;
        const int grid = (int)(blocks < (long long)kNumSMs * 2 ? blocks : (long long)kNumSMs * 2);
        const int width = 3 * C + 1;
        const size_t need = sizeof(float) * (size_t)grid * width;
        if (need > ws_bytes || !ws) { set_error("bn_pgr_bwd: workspace too small (%zu > %zu)", need, ws_bytes); return EEL_ERR_WORKSPACE; }
        float* partial = (float*)ws;
        const BnConst bn{bn_mean, bn_rstd, bn_gamma, bn_beta};
        const size_t shb = sizeof(float) * width;
        cudaStream_t st2 = (cudaStream_t)s;
        if (iters == 1) bn_pgr_bwd_kernel<T, 1, 4><<<grid, kPixThreads, shb, st2>>>((const T*)z, bn, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        else bn_pgr_bwd_kernel<T, 2, 2><<<grid, kPixThreads, shb, st2>>>((const T*)z, bn, sgm, w, (const T*)dy, dsgm, (T*)dx, partial, P, C, G);
        if (int rc = check_launch("bn_pgr_bwd")) return rc;
        // bn_sums = [2][C]: {sum g, sum g * xhat} (= dbeta, dgamma of the BatchNorm)
        RowSegs segs{{dw, db, bn_sums, bn_sums + C}, {C, C + 1, 2 * C + 1, 3 * C + 1}};
        finalize_rows_kernel<<<cdiv(width, 32), 1024, 0, st2>>>(partial, grid, width, segs);
        if (int rc = check_launch("bn_pgr_bwd.finalize")) return rc;
        bn_raw_sums_fix_kernel<<<cdiv(C, 128), 128, 0, st2>>>(bn_sums, bn_mean, bn_rstd, C);
        return check_launch("bn_pgr_bwd.fix");
    });
}

int eel_head_fwd(const void* x, const float* lnw, const float* lnb, const float* w, const float* b, float* prob, int N,
                 long long HW, int O, int dtype, eel_stream s) {
    EEL_REQUIRE(x && lnw && lnb && w && b && prob && N > 0 && HW > 0 && O > 0 && O <= kHeadMaxO, "head_fwd: bad argument (1 <= O <= 4)");
    long long P = (long long)N * HW;
    EEL_REQUIRE(P < (1LL << 31), "head_fwd: more than 2^31 pixels");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int G = kHeadC / Vec16<T>::N;
        int gpb = kPixThreads / G;
        long long blocks = (P + gpb - 1) / gpb;
        int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
        if (O == 1) {
            long long tb = (P + kTpThreads - 1) / kTpThreads;
            int g1 = (int)(tb < (long long)kNumSMs * 16 ? tb : (long long)kNumSMs * 16);
            head_fwd_tp_kernel<T><<<g1, kTpThreads, 0, (cudaStream_t)s>>>((const T*)x, lnw, lnb, w, b, prob, P);
            return check_launch("head_fwd(tp)");
        }
        head_fwd_kernel<T><<<grid, kPixThreads, 0, (cudaStream_t)s>>>((const T*)x, lnw, lnb, w, b, prob, P, HW, O);
        return check_launch("head_fwd");
    });
}

int eel_head_bwd(const void* x, const float* lnw, const float* lnb, const float* w, const float* b, const float* prob,
                 const float* dprob, void* dx, float* dlnw, float* dlnb, float* dw, float* db, int N, long long HW,
                 int O, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    (void)b;
    EEL_REQUIRE(x && lnw && lnb && w && prob && dprob && dx && dlnw && dlnb && dw && db && N > 0 && HW > 0 && O > 0 && O <= kHeadMaxO,
                "head_bwd: bad argument (1 <= O <= 4)");
    long long P = (long long)N * HW;
    EEL_REQUIRE(P < (1LL << 31), "head_bwd: more than 2^31 pixels");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int G = kHeadC / Vec16<T>::N;
        int gpb = kPixThreads / G;
        long long blocks = (P + gpb - 1) / gpb;
        int grid = (int)(blocks < (long long)kNumSMs * 4 ? blocks : (long long)kNumSMs * 4);
        int width = (2 + O) * kHeadC + O;
        size_t need = sizeof(float) * ((size_t)grid + 1) * width;
        if (need > ws_bytes || !ws) { set_error("head_bwd: workspace too small (%zu > %zu)", need, ws_bytes); return EEL_ERR_WORKSPACE; }
        float* partial = (float*)ws;
        if (O == 1) {
            // thread-per-pixel kernel: the grid is bounded by the partial rows the workspace holds (same bound as above)
            long long tb = (P + kHwThreads - 1) / kHwThreads;
            grid = (int)(tb < (long long)grid ? tb : (long long)grid);
            head_bwd_warp_kernel<T><<<grid, kHwThreads, 0, (cudaStream_t)s>>>((const T*)x, lnw, lnb, w, prob, dprob, (T*)dx, partial, P);
        } else head_bwd_kernel<T, kHeadMaxO><<<grid, kPixThreads, sizeof(float) * width, (cudaStream_t)s>>>((const T*)x, lnw, lnb, w, prob, dprob, (T*)dx, partial, P, HW, O);
        if (int rc = check_launch("head_bwd")) return rc;
        RowSegs segs{{dlnw, dlnb, dw, db}, {kHeadC, 2 * kHeadC, (2 + O) * kHeadC, (2 + O) * kHeadC + O}};
        finalize_rows_kernel<<<cdiv(width, 32), 1024, 0, (cudaStream_t)s>>>(partial, grid, width, segs);
        return check_launch("head_bwd.finalize");
    });
}

}  // extern "C"
