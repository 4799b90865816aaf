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

// ---- second generation of stage 1 + hysteresis (W % 4 == 0) ------------------------------------------------------------------
// Tile 128 x 32 (halo overhead 16 %), 256 threads.  The RGB (or gray) bytes of the tile's rows come in as aligned 32-bit words
// (coalesced 400-byte row segments) and are converted from shared memory; the {0, 1, 2} map leaves as 4-pixel words.  The
// candidates of a tile are merged by a union-find IN SHARED MEMORY (8-connectivity inside the tile), so a candidate's global
// label starts out as its tile-local root: the global lock-free union-find that follows only has to join components ACROSS
// tile borders (about one candidate in seven, and every chain it walks is at most one hop deep inside a tile).
constexpr int kT2W = 128, kT2H = 32;
constexpr int kG2W = kT2W + 4, kG2H = kT2H + 4;
constexpr int kM2W = kT2W + 2, kM2H = kT2H + 2;
constexpr int kRawWords = 100;                 // RGB: bytes [x0*3 - 8, x0*3 + 392) of a row hold pixels x0-2 .. x0+129 (+ slack)
constexpr int kRawWordsGray = 34;              // gray: bytes [x0 - 4, x0 + 132)

__device__ __forceinline__ int suf_find(const int* L, int i) {
    int r = L[i];
    while (r != i) { i = r; r = L[i]; }
    return r;
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
    while (true) {
        a = suf_find(L, a);
        b = suf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(L + a, b);
        if (old == a) return;
        a = old;
    }
}

constexpr int kStage1Smem = kG2H * kG2W + kM2H * (kM2W + 2) * 2 + kT2H * kT2W + kT2H * kT2W * 4 + 64;

template <bool RGB>
__global__ void __launch_bounds__(256) canny_stage1_v2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ out,
                                                            int* __restrict__ labels, int* __restrict__ list,
                                                            int* __restrict__ count, int H, int W, int low, int high) {
    constexpr int RW = RGB ? kRawWords : kRawWordsGray;
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    int* scratch = reinterpret_cast<int*>(smem_dyn);                                   // raw row words (kG2H x RW <= 3600), then tile-local labels
    short (*mag)[kM2W + 2] = reinterpret_cast<short (*)[kM2W + 2]>(scratch + kT2H * kT2W);
    // the raw words are dead once the gray tile exists: the first half of that buffer then holds the list of pixels above the low
    // threshold, which the suppression pass compacts IN PLACE into the list of surviving candidates (a survivor's slot is never
    // ahead of the entries already consumed: one barrier per round of 256); the labels are initialised only afterwards
    unsigned short* alist = reinterpret_cast<unsigned short*>(scratch) + kT2H * kT2W;   // (second half of the 16 KB: the labels' init below writes the first entries late)
    unsigned short* clist = alist;
    uint8_t (*cand)[kT2W] = reinterpret_cast<uint8_t (*)[kT2W]>(&mag[0][0] + kM2H * (kM2W + 2));
    uint8_t (*g)[kG2W] = reinterpret_cast<uint8_t (*)[kG2W]>(&cand[0][0] + kT2H * kT2W);
    int* ctr = reinterpret_cast<int*>(&g[0][0] + kG2H * kG2W);                          // nact, ncand, gbase
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * kT2W, y0 = blockIdx.y * kT2H;
    const long long img = (long long)n * H * W;
    const int bpp = RGB ? 3 : 1;
    const int rowbytes = W * bpp;
    const uint8_t* s = src + img * bpp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 3) ctr[threadIdx.x] = 0;

    // ---- rows y0-2 .. y0+33 (clamped), aligned words around the tile's byte range; words outside the row are skipped
    const int b0 = x0 * bpp - (RGB ? 8 : 4);
    {
        const int wlo = b0 < 0 ? (-b0) >> 2 : 0;                                // first / one-past-last word inside the row
        const int whi = min(RW, (rowbytes - b0) >> 2);
        for (int ly = warp; ly < kG2H; ly += 8) {
            const int y = min(max(y0 + ly - 2, 0), H - 1);
            const uint32_t* row = reinterpret_cast<const uint32_t*>(s + (long long)y * rowbytes + b0);
            for (int wi = lane; wi < RW; wi += 32) scratch[ly * RW + wi] = (wi >= wlo && wi < whi) ? (int)__ldg(row + wi) : 0;
        }
    }
    __syncthreads();
    {
        const uint8_t* raw = reinterpret_cast<const uint8_t*>(scratch);
        for (int ly = warp; ly < kG2H; ly += 8) {
            const uint8_t* rrow = raw + ly * (RW * 4) - b0;
            for (int lx = lane; lx < kG2W; lx += 32) {
                const int x = min(max(x0 + lx - 2, 0), W - 1);             // BORDER_REPLICATE (rows were clamped while loading)
                const uint8_t* p = rrow + x * bpp;
                g[ly][lx] = RGB ? (uint8_t)gray_of(p[0], p[1], p[2]) : p[0];
            }
        }
    }
    __syncthreads();
    // ---- L1 gradient magnitude, runs of 4 pixels: per column the vertical smooth s = a + 2b + c and the vertical difference
    // d = c - a are formed once and shared by the three outputs that use them
    for (int run = threadIdx.x; run < kM2H * 33; run += 256) {
        const int ly = run / 33, lx4 = (run - ly * 33) * 4;
        const int y = y0 + ly - 1;
        int sm_[6], df[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int gx = min(lx4 + j, kG2W - 1);
            const int a = g[ly][gx], bq = g[ly + 1][gx], c = g[ly + 2][gx];
            sm_[j] = a + 2 * bq + c;
            df[j] = c - a;
        }
        const bool yin = y >= 0 && y < H;
        short m4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = x0 + lx4 + u - 1;
            const int dx = sm_[u + 2] - sm_[u];
            const int dy = df[u] + 2 * df[u + 1] + df[u + 2];
            m4[u] = (yin && x >= 0 && x < W) ? (short)(abs(dx) + abs(dy)) : (short)0;
        }
        *reinterpret_cast<uint2*>(&mag[ly][lx4]) = make_uint2((uint32_t)(uint16_t)m4[0] | ((uint32_t)(uint16_t)m4[1] << 16),
                                                              (uint32_t)(uint16_t)m4[2] | ((uint32_t)(uint16_t)m4[3] << 16));
    }
    __syncthreads();
    // ---- pixels above the low threshold (a few per cent) are collected first, so that the direction test below runs on a dense
    // list instead of dragging whole warps through it for one or two lanes
    for (int i = threadIdx.x; i < kT2H * kT2W / 4; i += 256) {
        const int ly = i >> 5, lx4 = (i & 31) * 4;
        reinterpret_cast<uint32_t*>(&cand[0][0])[i] = 0;
        const uint2 mm = *reinterpret_cast<const uint2*>(&mag[ly + 1][lx4]);       // columns lx4 .. lx4+3 of the row (tile col = lx + 1)
        const short m0 = (short)(mm.x >> 16), m1 = (short)(mm.y & 0xffff), m2 = (short)(mm.y >> 16), m3 = mag[ly + 1][lx4 + 4];
        const bool in = (y0 + ly) < H && x0 + lx4 < W;
        const int f0 = in && m0 > low, f1 = in && m1 > low, f2 = in && m2 > low, f3 = in && m3 > low;
        const int mine = f0 + f1 + f2 + f3;
        // one shared-memory atomic per warp: exclusive prefix sum of the lanes' counts
        int pre = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        if (total == 0) continue;
        int base = 0;
        if (lane == 31) base = atomicAdd(&ctr[0], total);
        base = __shfl_sync(0xffffffffu, base, 31) + pre - mine;
        const int idx0 = ly * kT2W + lx4;
        if (f0) alist[base++] = (unsigned short)idx0;
        if (f1) alist[base++] = (unsigned short)(idx0 + 1);
        if (f2) alist[base++] = (unsigned short)(idx0 + 2);
        if (f3) alist[base++] = (unsigned short)(idx0 + 3);
    }
    __syncthreads();
    int* lab = scratch;
    const int nact = ctr[0];
    for (int k0 = 0; k0 < nact; k0 += 256) {
        const int k = k0 + threadIdx.x;
        const int idx = k < nact ? alist[k] : 0;
        __syncthreads();                       // every entry of this round has been read: survivors may now be written below it
        bool keep = false;
        if (k < nact) {
        const int ly = idx >> 7, lx = idx & (kT2W - 1);
        const int my = ly + 1, mx = lx + 1, gy = ly + 2, gx = lx + 2;
        const int m = mag[my][mx];
        const int a = g[gy - 1][gx - 1], b = g[gy - 1][gx], c = g[gy - 1][gx + 1];
        const int d = g[gy][gx - 1], f = g[gy][gx + 1];
        const int p = g[gy + 1][gx - 1], q = g[gy + 1][gx], r = g[gy + 1][gx + 1];
        const int dx = (c + 2 * f + r) - (a + 2 * d + p);
        const int dy = (p + 2 * q + r) - (a + 2 * b + c);
        const int ax = abs(dx), ay = abs(dy) << 15;
        const int t22 = ax * 13573;
        if (ay < t22) {
            keep = m > mag[my][mx - 1] && m >= mag[my][mx + 1];
        } else {
            const int t67 = t22 + (ax << 16);
            if (ay > t67) {
                keep = m > mag[my - 1][mx] && m >= mag[my + 1][mx];
            } else {
                const int sgn = (dx ^ dy) < 0 ? -1 : 1;
                keep = m > mag[my - 1][mx - sgn] && m > mag[my + 1][mx + sgn];
            }
        }
        if (keep) cand[ly][lx] = m > high ? 2 : 1;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctr[1], __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) clist[base + __popc(ballot & ((1u << lane) - 1))] = (unsigned short)idx;
        }
    }
    __syncthreads();
    // ---- the {0, 1, 2} map leaves as 4-pixel words
    for (int i = threadIdx.x; i < kT2H * kT2W / 4; i += 256) {
        const int ly = i >> 5, lx4 = (i & 31) * 4;
        const int y = y0 + ly;
        if (y < H && x0 + lx4 < W)          // W % 4 == 0: a 4-pixel group is inside or outside as a whole
            *reinterpret_cast<uint32_t*>(out + img + (long long)y * W + x0 + lx4) = reinterpret_cast<const uint32_t*>(&cand[0][0])[i];
    }
    // ---- tile-local union-find over the candidates (8-connectivity inside the tile).  The candidate list moves into registers
    // / the magnitude buffer first: the labels are about to overwrite the buffer it lives in.
    const int ncand = ctr[1];
    unsigned short* clist2 = reinterpret_cast<unsigned short*>(&mag[0][0]);
    for (int k = threadIdx.x; k < ncand; k += 256) clist2[k] = clist[k];
    __syncthreads();
    clist = clist2;
    for (int k = threadIdx.x; k < ncand; k += 256) lab[clist[k]] = clist[k];
    __syncthreads();
    for (int k = threadIdx.x; k < ncand; k += 256) {
        const int idx = clist[k];
        const int ly = idx >> 7, lx = idx & (kT2W - 1);
        if (lx + 1 < kT2W && cand[ly][lx + 1]) suf_union(lab, idx, idx + 1);
        if (ly + 1 < kT2H) {
            if (lx > 0 && cand[ly + 1][lx - 1]) suf_union(lab, idx, idx + kT2W - 1);
            if (cand[ly + 1][lx]) suf_union(lab, idx, idx + kT2W);
            if (lx + 1 < kT2W && cand[ly + 1][lx + 1]) suf_union(lab, idx, idx + kT2W + 1);
        }
    }
    // ---- one slot reservation per tile; global labels start at the tile-local roots
    if (threadIdx.x == 0 && ncand > 0) ctr[2] = atomicAdd(count, ncand);
    __syncthreads();
    const int gbase = ctr[2];
    for (int k = threadIdx.x; k < ncand; k += 256) {
        const int idx = clist[k];
        const int ly = idx >> 7, lx = idx & (kT2W - 1);
        const int root = suf_find(lab, idx);
        const int id = (int)(img + (long long)(y0 + ly) * W + x0 + lx);
        list[gbase + k] = id;
        labels[id] = (int)(img + (long long)(y0 + (root >> 7)) * W + x0 + (root & (kT2W - 1)));
    }
}

// joins across tile borders only: every in-tile pair was merged in shared memory by stage 1
__global__ void canny_merge_v2_kernel(const uint8_t* __restrict__ out, int* __restrict__ labels, const int* __restrict__ list,
                                      const int* __restrict__ count, int H, int W) {
    const int cnt = *count;
    const long long HW = (long long)H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int id = list[i];
        const int r = (int)(id % HW);
        const int y = r / W, x = r - y * W;
        const int lx = x % kT2W, ly = y % kT2H;
        const bool right = lx == kT2W - 1, left = lx == 0, bottom = ly == kT2H - 1;
        if (!(right || left || bottom)) continue;
        if (right && x + 1 < W && out[id + 1]) uf_union(labels, id, id + 1);
        if (y + 1 < H) {
            if ((bottom || left) && x > 0 && out[id + W - 1]) uf_union(labels, id, id + W - 1);
            if (bottom && out[id + W]) uf_union(labels, id, id + W);
            if ((bottom || right) && x + 1 < W && out[id + W + 1]) uf_union(labels, id, id + W + 1);
        }
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
    const int g = kNumSMs * 8;
    const bool v2 = W % 4 == 0 && ((uintptr_t)src % 4) == 0 && ((uintptr_t)edges % 4) == 0;
    if (v2) {
        dim3 grid(cdiv(W, kT2W), cdiv(H, kT2H), N);
        if (grid.y > 65535 || grid.z > 65535) { set_error("canny: image too large"); return EEL_ERR_INVALID; }
        static SmemOptIn configured;
        if (!configured.ensure(canny_stage1_v2_kernel<RGB>, kStage1Smem)) { set_error("canny: cannot raise dynamic shared memory"); return EEL_ERR_CUDA; }
        canny_stage1_v2_kernel<RGB><<<grid, 256, kStage1Smem, st>>>(src, edges, labels, list, count, H, W, low, high);
        if (int rc = check_launch("canny.stage1")) return rc;
        canny_merge_v2_kernel<<<g, 256, 0, st>>>(edges, labels, list, count, H, W);
        if (int rc = check_launch("canny.merge")) return rc;
    } else {
        dim3 grid(cdiv(W, kTW), cdiv(H, kTH), N);
        if (grid.y > 65535 || grid.z > 65535) { set_error("canny: image too large"); return EEL_ERR_INVALID; }
        canny_stage1_kernel<RGB><<<grid, 256, 0, st>>>(src, edges, labels, list, count, H, W, low, high);
        if (int rc = check_launch("canny.stage1")) return rc;
        canny_merge_kernel<<<g, 256, 0, st>>>(edges, labels, list, count, H, W);
        if (int rc = check_launch("canny.merge")) return rc;
    }
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
