// GPU evaluate() metrics (reference evaluate.py:62-124): thresholded confusion counts and the boundary-F1 ingredients
// (seg2bnd = mask minus its erosion by a 3x3 square iterated d times = a (2d+1)^2 square minimum with "outside = max"
// border, evaluate.py:25-41), all integer counts so the host-side formulas reproduce the reference exactly.
#include "common.cuh"

namespace eel {

// counts[0..3] += {TP, TN, FP, FN}; pred = seg > 0.5, label compared with == 1 / == 0 (evaluate.py:91-100)
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ seg, const float* __restrict__ lab, long long n,
                                                      unsigned long long* __restrict__ counts) {
    unsigned tp = 0, tn = 0, fp = 0, fn = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool p = seg[i] > 0.5f;
        const float l = lab[i];
        tp += p && l == 1.f; tn += !p && l == 0.f; fp += p && l == 0.f; fn += !p && l == 1.f;
    }
    __shared__ unsigned red[4][8];
    unsigned v[4] = {tp, tn, fp, fn};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        if (s) atomicAdd(counts + threadIdx.x, s);
    }
}

// horizontal pass: hmin[which][n][y][x] = min over |dx| <= d (inside the image) of the uint8 mask
//   which = 0: prediction mask (seg > 0.5) * 255;  which = 1: (label * 255) truncated to uint8
__global__ void erode_h_kernel(const float* __restrict__ seg, const float* __restrict__ lab, uint8_t* __restrict__ hmin, int N, int H,
                               int W, int d) {
    const long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * total; i += (long long)gridDim.x * blockDim.x) {
        const int which = i >= total;
        const long long j = which ? i - total : i;
        const int x = (int)(j % W);
        const long long row = j - x;
        int m = 255;
        for (int xx = max(0, x - d); xx <= min(W - 1, x + d); ++xx) {
            int v = which ? (int)(uint8_t)(int)(lab[row + xx] * 255.f) : (seg[row + xx] > 0.5f ? 255 : 0);
            m = min(m, v);
        }
        hmin[i] = (uint8_t)m;
    }
}

// vertical pass + boundary + per-sample counts: out[n][0..2] += {sum(pred_b & gt_b), sum(pred_b), sum(gt_b)}
__global__ void __launch_bounds__(256) boundary_kernel(const float* __restrict__ seg, const float* __restrict__ lab,
                                                     const uint8_t* __restrict__ hmin, int N, int H, int W, int d,
                                                     unsigned long long* __restrict__ out) {
    const int n = blockIdx.y;
    const long long HW = (long long)H * W, total = (long long)N * HW;
    unsigned c0 = 0, c1 = 0, c2 = 0;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < HW; j += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(j / W), x = (int)(j - (long long)y * W);
        const long long base = (long long)n * HW;
        int ep = 255, eg = 255;
        for (int yy = max(0, y - d); yy <= min(H - 1, y + d); ++yy) {
            ep = min(ep, (int)hmin[base + (long long)yy * W + x]);
            eg = min(eg, (int)hmin[total + base + (long long)yy * W + x]);
        }
        const int mp = seg[base + j] > 0.5f ? 255 : 0;
        const int mg = (int)(uint8_t)(int)(lab[base + j] * 255.f);
        const bool bp = mp - ep > 0, bg = mg - eg > 0;
        c0 += bp && bg; c1 += bp; c2 += bg;
    }
    __shared__ unsigned red[3][8];
    unsigned v[3] = {c0, c1, c2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        if (s) atomicAdd(out + n * 3 + threadIdx.x, s);
    }
}

}  // namespace eel

using namespace eel;

extern "C" {

int eel_confusion_counts(const float* seg, const float* labels, long long n, unsigned long long* counts, eel_stream s) {
    EEL_REQUIRE(seg && labels && counts && n > 0, "confusion_counts: bad argument");
    long long b = (n + 255) / 256;
    int grid = (int)(b < (long long)kNumSMs * 8 ? b : (long long)kNumSMs * 8);
    confusion_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(seg, labels, n, counts);
    return check_launch("confusion_counts");
}

size_t eel_boundary_workspace_bytes(int N, int H, int W) { return 2 * (size_t)N * H * W; }

int eel_boundary_counts(const float* seg, const float* labels, int N, int H, int W, int iterations, unsigned long long* per_sample,
                        void* ws, size_t ws_bytes, eel_stream s) {
    EEL_REQUIRE(seg && labels && per_sample && N > 0 && H > 0 && W > 0 && iterations > 0, "boundary_counts: bad argument");
    EEL_REQUIRE(ws && ws_bytes >= 2 * (size_t)N * H * W, "boundary_counts: workspace too small");
    cudaStream_t st = (cudaStream_t)s;
    const long long total = (long long)N * H * W;
    long long b = (2 * total + 255) / 256;
    int grid = (int)(b < (long long)kNumSMs * 16 ? b : (long long)kNumSMs * 16);
    erode_h_kernel<<<grid, 256, 0, st>>>(seg, labels, (uint8_t*)ws, N, H, W, iterations);
    if (int rc = check_launch("boundary_counts.h")) return rc;
    long long hb = ((long long)H * W + 255) / 256;
    dim3 g2((unsigned)(hb < 64 ? hb : 64), (unsigned)N);
    boundary_kernel<<<g2, 256, 0, st>>>(seg, labels, (const uint8_t*)ws, N, H, W, iterations, per_sample);
    return check_launch("boundary_counts.v");
}

}  // extern "C"
