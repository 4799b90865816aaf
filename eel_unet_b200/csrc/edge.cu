// Integer edge maps with OpenCV semantics, bit-exact:
//   gray  = (9798 R + 19235 G + 3735 B + 16384) >> 15
//   Canny = 3x3 Sobel (replicated border) -> L1 magnitude (zero outside) -> tangent-table
//           non-maximum suppression -> hysteresis (8-connected components that contain a strong pixel)
// Stage 1 is one fused shared-memory tile kernel (gray halo 2, magnitude halo 1) that reads the RGB
// image once and writes a {0 none, 1 weak, 2 strong} byte per pixel into the OUTPUT image, appending
// the sparse candidate pixels to a list.  Stage 2 is a lock-free union-find over that list only, so
// the dense traffic stays at the algorithmic 3 B read + 1 B written per pixel.
#include "common.cuh"

namespace eel {

constexpr int kTW = 64, kTH = 16;          // output tile
constexpr int kGW = kTW + 4, kGH = kTH + 4;  // gray tile (halo 2)
constexpr int kMW = kTW + 2, kMH = kTH + 2;  // magnitude tile (halo 1)

__device__ __forceinline__ int gray_of(int r, int g, int b) { return (9798 * r + 19235 * g + 3735 * b + 16384) >> 15; }

template <bool RGB>
__global__ void __launch_bounds__(256) canny_stage1_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ out,
                                                         int* __restrict__ labels, int* __restrict__ list,
                                                         int* __restrict__ count, int H, int W, int low, int high) {
    __shared__ uint8_t g[kGH][kGW];
    __shared__ short mag[kMH][kMW + 2];
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const long long img = (long long)n * H * W;
    const uint8_t* s = src + img * (RGB ? 3 : 1);

    for (int i = threadIdx.x; i < kGH * kGW; i += 256) {
        int ly = i / kGW, lx = i - ly * kGW;
        int y = min(max(y0 + ly - 2, 0), H - 1), x = min(max(x0 + lx - 2, 0), W - 1);   // BORDER_REPLICATE
        long long o = (long long)y * W + x;
        g[ly][lx] = RGB ? (uint8_t)gray_of(s[o * 3], s[o * 3 + 1], s[o * 3 + 2]) : s[o];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMH * kMW; i += 256) {
        int ly = i / kMW, lx = i - ly * kMW;
        int y = y0 + ly - 1, x = x0 + lx - 1;
        int m = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const int gy = ly + 1, gx = lx + 1;   // centre in gray-tile coordinates
            int a = g[gy - 1][gx - 1], b = g[gy - 1][gx], c = g[gy - 1][gx + 1];
            int d = g[gy][gx - 1], f = g[gy][gx + 1];
            int p = g[gy + 1][gx - 1], q = g[gy + 1][gx], r = g[gy + 1][gx + 1];
            int dx = (c + 2 * f + r) - (a + 2 * d + p);
            int dy = (p + 2 * q + r) - (a + 2 * b + c);
            m = abs(dx) + abs(dy);
        }
        mag[ly][lx] = (short)m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTH * kTW; i += 256) {
        int ly = i / kTW, lx = i - ly * kTW;
        int y = y0 + ly, x = x0 + lx;
        bool inside = y < H && x < W;
        int v = 0;
        if (inside) {
            const int gy = ly + 2, gx = lx + 2, my = ly + 1, mx = lx + 1;
            int m = mag[my][mx];
            if (m > low) {
                int a = g[gy - 1][gx - 1], b = g[gy - 1][gx], c = g[gy - 1][gx + 1];
                int d = g[gy][gx - 1], f = g[gy][gx + 1];
                int p = g[gy + 1][gx - 1], q = g[gy + 1][gx], r = g[gy + 1][gx + 1];
                int dx = (c + 2 * f + r) - (a + 2 * d + p);
                int dy = (p + 2 * q + r) - (a + 2 * b + c);
                int ax = abs(dx), ay = abs(dy) << 15;
                int t22 = ax * 13573;
                bool keep;
                if (ay < t22) {
                    keep = m > mag[my][mx - 1] && m >= mag[my][mx + 1];
                } else {
                    int t67 = t22 + (ax << 16);
                    if (ay > t67) {
                        keep = m > mag[my - 1][mx] && m >= mag[my + 1][mx];
                    } else {
                        int sgn = (dx ^ dy) < 0 ? -1 : 1;
                        keep = m > mag[my - 1][mx - sgn] && m > mag[my + 1][mx + sgn];
                    }
                }
                if (keep) v = m > high ? 2 : 1;
            }
            out[img + (long long)y * W + x] = (uint8_t)v;
        }
        // warp-aggregated append of the candidates
        unsigned ballot = __ballot_sync(0xffffffffu, v != 0);
        if (ballot) {
            int lane = threadIdx.x & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(count, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (v != 0) {
                int id = (int)(img + (long long)y * W + x);
                list[base + __popc(ballot & ((1u << lane) - 1))] = id;
                labels[id] = id;
            }
        }
    }
}

__device__ __forceinline__ int uf_find(const int* L, int i) {
    int r = __ldcg(L + i);
    while (r != i) { i = r; r = __ldcg(L + i); }
    return r;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // link the larger root under the smaller
        int old = atomicMin(L + a, b);
        if (old == a) return;
        a = old;
    }
}

// union every candidate with its E, SW, S, SE candidate neighbours (each 8-neighbour pair once)
__global__ void canny_merge_kernel(const uint8_t* __restrict__ out, int* __restrict__ labels, const int* __restrict__ list,
                                   const int* __restrict__ count, int H, int W) {
    const int cnt = *count;
    const long long HW = (long long)H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        int id = list[i];
        int r = (int)(id % HW);
        int y = r / W, x = r - y * W;
        if (x + 1 < W && out[id + 1]) uf_union(labels, id, id + 1);
        if (y + 1 < H) {
            if (x > 0 && out[id + W - 1]) uf_union(labels, id, id + W - 1);
            if (out[id + W]) uf_union(labels, id, id + W);
            if (x + 1 < W && out[id + W + 1]) uf_union(labels, id, id + W + 1);
        }
    }
}

// a strong pixel marks the root of its component as a final edge (255)
__global__ void canny_mark_kernel(uint8_t* __restrict__ out, const int* __restrict__ labels, const int* __restrict__ list,
                                  const int* __restrict__ count) {
    const int cnt = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        int id = list[i];
        if (out[id] == 2) out[uf_find(labels, id)] = 255;
    }
}

// every candidate takes its root's verdict: 255 (component holds a strong pixel) or 0
__global__ void canny_resolve_kernel(uint8_t* __restrict__ out, const int* __restrict__ labels, const int* __restrict__ list,
                                     const int* __restrict__ count) {
    const int cnt = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        int id = list[i];
        int root = uf_find(labels, id);
        uint8_t v = __ldcg(out + root);
        out[id] = v == 255 ? 255 : 0;
    }
}

__global__ void gray_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, long long npix) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x)
        gray[i] = (uint8_t)gray_of(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2]);
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// LAPLACE == false: convertScaleAbs(magnitude(Sobel_x, Sobel_y)) ; true: convertScaleAbs(Laplacian)
template <bool LAPLACE>
__global__ void deriv_map_kernel(const uint8_t* __restrict__ gray, uint8_t* __restrict__ out, int H, int W, long long npix) {
    const long long HW = (long long)H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / HW;
        int r = (int)(i - n * HW);
        int y = r / W, x = r - y * W;
        const uint8_t* s = gray + n * HW;
        int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
        int res;
        if (LAPLACE) {
            int v = s[(long long)ym * W + x] + s[(long long)yp * W + x] + s[(long long)y * W + xm] + s[(long long)y * W + xp] -
                    4 * s[(long long)y * W + x];
            res = min(abs(v), 255);
        } else {
            int a = s[(long long)ym * W + xm], b = s[(long long)ym * W + x], c = s[(long long)ym * W + xp];
            int d = s[(long long)y * W + xm], f = s[(long long)y * W + xp];
            int p = s[(long long)yp * W + xm], q = s[(long long)yp * W + x], rr = s[(long long)yp * W + xp];
            int dx = (c + 2 * f + rr) - (a + 2 * d + p);
            int dy = (p + 2 * q + rr) - (a + 2 * b + c);
            int v = dx * dx + dy * dy;                 // <= 2 * 1020^2, exact in int
            int rt = (int)sqrtf((float)v);
            while (rt * rt > v) --rt;
            while ((rt + 1) * (rt + 1) <= v) ++rt;     // rt = floor(sqrt(v))
            rt += (v > rt * rt + rt) ? 1 : 0;          // round to nearest (ties cannot occur for integer v)
            res = min(rt, 255);
        }
        out[i] = (uint8_t)res;
    }
}

__global__ void canny_enhance_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ edges,
                                     uint8_t* __restrict__ out, long long npix, float cr, float cg, float cb, float alpha) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        bool e = edges[i] != 0;
        float col[3] = {cr, cg, cb};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // addWeighted(img, 1, overlay, alpha, 0) evaluated in fp32, rounded half-to-even, saturated
            float t = __fadd_rn(__fadd_rn(__fmul_rn((float)rgb[i * 3 + k], 1.0f), __fmul_rn(e ? col[k] : 0.f, alpha)), 0.0f);
            int v = __float2int_rn(t);
            out[i * 3 + k] = (uint8_t)min(max(v, 0), 255);
        }
    }
}

static int ew_grid1(long long n) {
    long long b = (n + 255) / 256, cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <bool RGB>
static int canny_impl(const uint8_t* src, uint8_t* edges, int N, int H, int W, int low, int high, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
    long long P = (long long)N * H * W;
    if (P >= (1LL << 31)) { set_error("canny: more than 2^31 pixels in one call"); return EEL_ERR_INVALID; }
    size_t need = eel_canny_workspace_bytes(N, H, W);
    if (!ws || ws_bytes < need) { set_error("canny: workspace too small (%zu > %zu)", need, ws_bytes); return EEL_ERR_WORKSPACE; }
    int* count = (int*)ws;
    int* labels = count + 4;
    int* list = labels + P;
    if (cudaMemsetAsync(count, 0, 16, st) != cudaSuccess) { set_error("canny: memset failed"); return EEL_ERR_CUDA; }
    dim3 grid(cdiv(W, kTW), cdiv(H, kTH), N);
    if (grid.y > 65535 || grid.z > 65535) { set_error("canny: image too large"); return EEL_ERR_INVALID; }
    canny_stage1_kernel<RGB><<<grid, 256, 0, st>>>(src, edges, labels, list, count, H, W, low, high);
    if (int rc = check_launch("canny.stage1")) return rc;
    const int g = kNumSMs * 8;
    canny_merge_kernel<<<g, 256, 0, st>>>(edges, labels, list, count, H, W);
    if (int rc = check_launch("canny.merge")) return rc;
    canny_mark_kernel<<<g, 256, 0, st>>>(edges, labels, list, count);
    if (int rc = check_launch("canny.mark")) return rc;
    canny_resolve_kernel<<<g, 256, 0, st>>>(edges, labels, list, count);
    return check_launch("canny.resolve");
}

}  // namespace eel

using namespace eel;

extern "C" {

size_t eel_canny_workspace_bytes(int N, int H, int W) {
    return 16 + 2 * sizeof(int) * (size_t)N * (size_t)H * (size_t)W;
}

int eel_gray_u8(const uint8_t* rgb, uint8_t* gray, int N, int H, int W, eel_stream s) {
    EEL_REQUIRE(rgb && gray && N > 0 && H > 0 && W > 0, "gray_u8: bad argument");
    long long P = (long long)N * H * W;
    gray_kernel<<<ew_grid1(P), 256, 0, (cudaStream_t)s>>>(rgb, gray, P);
    return check_launch("gray_u8");
}

int eel_canny_rgb(const uint8_t* rgb, uint8_t* edges, int N, int H, int W, int low, int high, void* ws, size_t ws_bytes,
                  eel_stream s) {
    EEL_REQUIRE(rgb && edges && N > 0 && H > 0 && W > 0, "canny_rgb: bad argument");
    return canny_impl<true>(rgb, edges, N, H, W, low, high, ws, ws_bytes, (cudaStream_t)s);
}

int eel_canny_gray(const uint8_t* gray, uint8_t* edges, int N, int H, int W, int low, int high, void* ws, size_t ws_bytes,
                   eel_stream s) {
    EEL_REQUIRE(gray && edges && N > 0 && H > 0 && W > 0, "canny_gray: bad argument");
    return canny_impl<false>(gray, edges, N, H, W, low, high, ws, ws_bytes, (cudaStream_t)s);
}

int eel_sobel_map(const uint8_t* gray, uint8_t* out, int N, int H, int W, eel_stream s) {
    EEL_REQUIRE(gray && out && N > 0 && H > 0 && W > 0, "sobel_map: bad argument");
    long long P = (long long)N * H * W;
    deriv_map_kernel<false><<<ew_grid1(P), 256, 0, (cudaStream_t)s>>>(gray, out, H, W, P);
    return check_launch("sobel_map");
}

int eel_laplacian_map(const uint8_t* gray, uint8_t* out, int N, int H, int W, eel_stream s) {
    EEL_REQUIRE(gray && out && N > 0 && H > 0 && W > 0, "laplacian_map: bad argument");
    long long P = (long long)N * H * W;
    deriv_map_kernel<true><<<ew_grid1(P), 256, 0, (cudaStream_t)s>>>(gray, out, H, W, P);
    return check_launch("laplacian_map");
}

int eel_canny_enhance(const uint8_t* rgb, const uint8_t* edges, uint8_t* out, int N, int H, int W, int cr, int cg, int cb,
                      float alpha, eel_stream s) {
    EEL_REQUIRE(rgb && edges && out && N > 0 && H > 0 && W > 0, "canny_enhance: bad argument");
    long long P = (long long)N * H * W;
    canny_enhance_kernel<<<ew_grid1(P), 256, 0, (cudaStream_t)s>>>(rgb, edges, out, P, (float)cr, (float)cg, (float)cb, alpha);
    return check_launch("canny_enhance");
}

}  // extern "C"
