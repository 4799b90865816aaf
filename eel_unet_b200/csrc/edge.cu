// Integer edge maps with OpenCV semantics, bit-exact:
//   gray  = (9798 R + 19235 G + 3735 B + 16384) >> 15
//   Canny = 3x3 Sobel (replicated border) -> L1 magnitude (zero outside) -> tangent-table
//           non-maximum suppression -> hysteresis (8-connected components that contain a strong pixel)
// Three generations live here, chosen by canny_impl from the image width:
//   W % 32 == 0, W <= 512   register-resident row bands (one warp per band, no shared memory) writing two bitmaps, hysteresis as
//                           a bitmap flood fill (column strips in registers for H <= 512, a warp per 32-row band above that)
//   W % 4 == 0              fused shared-memory tile kernel (gray halo 2, magnitude halo 1) that reads the image once and writes a
//                           {0 none, 1 weak, 2 strong} byte per pixel plus a sparse candidate list; hysteresis = lock-free union-find
//                           over that list (tile-local in shared memory first, then across tile borders)
//   otherwise               the first, byte-wise version of the tile kernel
// The dense traffic stays at the algorithmic 3 B read + 1 B written per pixel in all of them.
#include <algorithm>

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

// ---- third generation (W % 32 == 0, W <= 512): no shared-memory tiles, no labels ------------------------------------------------
// Stage 1: one WARP owns a band of BR image rows over the full row width: lane l holds pixels [16 l, 16 l + 16) of the current
// row and slides down the band.  Two gray rows, two magnitude rows and the row being formed live in registers (the row loop is
// unrolled by three so that the rotation is a renaming); the only horizontal exchange is one value per side of the vertical
// smooth / difference sums and of the magnitude (6 shuffles per row of 512 pixels).  The replicated border needs no halo: a row
// starts and ends inside the warp.  Pixels above the low threshold (a few per cent) branch into the direction test when their
// magnitude is formed and into the suppression test one row later, everything else is ~25 instructions per pixel (the tile kernel
// above: ~125).  Output: two BITMAPS per image (suppressed candidates; candidates above the high threshold), 16 bits per lane.
// Stage 2: hysteresis = flood fill of the strong bitmap through the candidate bitmap, one CTA per image with both bitmaps in
// shared memory (512 x 512: 2 x 32 KB).  A warp holds 32 rows (lane = row, 16 words = 512 pixels per lane): a whole row floods
// in one carry-propagating pass per direction ((w & ~(w + s)) | s extends every seed through its run of candidates), rows
// exchange through shuffles, 32-row bands through shared memory until nothing changes; the CTA then expands its bitmap into the
// 0 / 255 byte image.  No global atomics, no memset, 0.25 bytes of workspace per pixel.
struct RawRow { uint4 a, b, c; };

template <bool RGB>
__device__ __forceinline__ RawRow canny_load_row(const uint8_t* p) {
    RawRow r;
    r.a = __ldg(reinterpret_cast<const uint4*>(p));
    if (RGB) {
        r.b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        r.c = __ldg(reinterpret_cast<const uint4*>(p) + 2);
    } else {
        r.b = r.a; r.c = r.a;
    }
    return r;
}

template <bool RGB>
__device__ __forceinline__ void canny_gray16(const RawRow& r, int (&g)[16]) {
    const uint32_t w[12] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w, r.c.x, r.c.y, r.c.z, r.c.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (RGB) {
            const int i0 = 3 * j, i1 = 3 * j + 1, i2 = 3 * j + 2;
            const int R = (int)((w[i0 >> 2] >> (8 * (i0 & 3))) & 0xffu);
            const int G = (int)((w[i1 >> 2] >> (8 * (i1 & 3))) & 0xffu);
            const int B = (int)((w[i2 >> 2] >> (8 * (i2 & 3))) & 0xffu);
            g[j] = gray_of(R, G, B);
        } else {
            g[j] = (int)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
        }
    }
}

// per-row masks, bit j = pixel j of the lane: above the low / high threshold, gradient direction classes (horizontal: compare
// left / right; vertical: up / down; otherwise a diagonal, `neg`: dx and dy of opposite sign = up-right / down-left), and the
// comparisons of the row's magnitudes with the row above (gt_u: M[j] > up[j]; gt_ul / gt_ur: > up[j-1] / up[j+1]) and with the
// left neighbour (gt_l, 17 bits: bit 16 belongs to the first pixel of the next lane)
struct RowMasks { unsigned cand, strong, hor, ver, neg, gt_l, gt_u, gt_ul, gt_ur; };

// one row step: gC (gray row t) is formed from `raw`; for t >= 2 the magnitude row MC (image row y0 + t - 3) from gA, gB, gC with
// its masks mC; for t >= 4 the suppression of image row y0 + t - 4 (magnitudes MB, masks mB, row below MC) is stored.  Straight-line
// code: every comparison is made for every pixel once and combined as 16-bit masks (a divergent per-pixel branch for the few
// per cent above the low threshold runs for nearly every j of nearly every row, with two or three lanes in it).
template <bool RGB>
__device__ __forceinline__ void canny_row_step(const int t, const RawRow& raw, const int (&gA)[16], const int (&gB)[16], int (&gC)[16],
                                               const int (&MB)[18], int (&MC)[18], const RowMasks& mB, RowMasks& mC,
                                               const int y0, const int H, const int low, const int high, const bool first,
                                               const bool last, const bool act, uint16_t* __restrict__ wrow,
                                               uint16_t* __restrict__ srow, const int row_stride16) {
    canny_gray16<RGB>(raw, gC);
    if (t < 2) return;
    int s[18], d[18];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        s[j + 1] = gA[j] + 2 * gB[j] + gC[j];
        d[j + 1] = gC[j] - gA[j];
    }
    {
        const int sl = __shfl_up_sync(0xffffffffu, s[16], 1), dl = __shfl_up_sync(0xffffffffu, d[16], 1);
        const int sr = __shfl_down_sync(0xffffffffu, s[1], 1), dr = __shfl_down_sync(0xffffffffu, d[1], 1);
        s[0] = first ? s[1] : sl;  d[0] = first ? d[1] : dl;          // BORDER_REPLICATE
        s[17] = last ? s[16] : sr; d[17] = last ? d[16] : dr;
    }
    const int ry = y0 + t - 3;
    const int rowmask = (ry >= 0 && ry < H) ? -1 : 0;                 // the magnitude is zero outside the image
    // every mask is built by shifting the SIGN of a difference in at bit 0 (a > b  <=>  b - a < 0; all operands < 2^27), last
    // pixel first, so that pixel j ends at bit j: two instructions per comparison
#define EEL_PUSH(mask, diff) mask = __funnelshift_l((unsigned)(diff), mask, 1)
    unsigned cand = 0, strong = 0, hor = 0, ver = 0, neg = 0;
#pragma unroll
    for (int j = 15; j >= 0; --j) {
        const int dx = s[j + 2] - s[j];
        const int dy = d[j] + 2 * d[j + 1] + d[j + 2];
        const int ax = abs(dx), ay = abs(dy);
        const int m = (ax + ay) & rowmask;
        MC[j + 1] = m;
        const int dh = (ay << 15) - ax * 13573;          // < 0: closer than 22.5 degrees to the x axis
        const int dv = (ax << 16) - dh;                  // < 0: beyond 67.5 degrees  (t67 = t22 + (ax << 16))
        EEL_PUSH(cand, low - m);
        EEL_PUSH(strong, high - m);
        EEL_PUSH(hor, dh);
        EEL_PUSH(ver, dv);
        EEL_PUSH(neg, dx ^ dy);
    }
    {
        const int ml = __shfl_up_sync(0xffffffffu, MC[16], 1), mr = __shfl_down_sync(0xffffffffu, MC[1], 1);
        MC[0] = first ? 0 : ml;
        MC[17] = last ? 0 : mr;
    }
    unsigned gt_l = 0, gt_u = 0, gt_ul = 0, gt_ur = 0, gt_d = 0, gt_dl = 0, gt_dr = 0;
#pragma unroll
    for (int j = 16; j >= 0; --j) EEL_PUSH(gt_l, MC[j] - MC[j + 1]);
#pragma unroll
    for (int j = 15; j >= 0; --j) {
        EEL_PUSH(gt_u, MB[j + 1] - MC[j + 1]);            // the new row against the row above it ...
        EEL_PUSH(gt_ul, MB[j] - MC[j + 1]);
        EEL_PUSH(gt_ur, MB[j + 2] - MC[j + 1]);
        EEL_PUSH(gt_dl, MC[j] - MB[j + 1]);               // ... and the row above against the new row (strict on the diagonals)
        EEL_PUSH(gt_dr, MC[j + 2] - MB[j + 1]);
    }
#undef EEL_PUSH
    gt_d = gt_u;                                                       // bit j: below[j] > M[j], i.e. NOT (M[j] >= below[j])
    mC.cand = cand; mC.strong = strong; mC.hor = hor; mC.ver = ver & ~hor; mC.neg = neg;
    mC.gt_l = gt_l; mC.gt_u = gt_u; mC.gt_ul = gt_ul; mC.gt_ur = gt_ur;
    if (t < 4) return;
    const int rn = y0 + t - 4;
    const unsigned keep_h = mB.gt_l & ~(mB.gt_l >> 1);                 // m > left  && m >= right
    const unsigned keep_v = mB.gt_u & ~gt_d;                           // m > up    && m >= down
    const unsigned keep_d2 = mB.gt_ul & gt_dr;                         // m > up-left  && m > down-right
    const unsigned keep_d3 = mB.gt_ur & gt_dl;                         // m > up-right && m > down-left
    const unsigned diag = ~(mB.hor | mB.ver);
    const unsigned wk = mB.cand & 0xffffu &
                        ((mB.hor & keep_h) | (mB.ver & keep_v) | (diag & ((~mB.neg & keep_d2) | (mB.neg & keep_d3))));
    const unsigned st = wk & mB.strong;
    if (act && rn < H) {
        wrow[(size_t)rn * row_stride16] = (uint16_t)wk;
        srow[(size_t)rn * row_stride16] = (uint16_t)st;
    }
}

template <bool RGB>
__global__ void __launch_bounds__(128, 4) canny_rows_kernel(const uint8_t* __restrict__ src, uint16_t* __restrict__ weak,
                                                       uint16_t* __restrict__ strong, int N, int H, int W, int BR, int bands,
                                                       int low, int high) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= N * bands) return;                          // (the whole warp leaves together)
    const int n = gw / bands, band = gw - n * bands;
    const int y0 = band * BR;
    const int nl = W >> 4;                                // lanes that hold pixels
    const bool act = lane < nl;
    const bool first = lane == 0, last = lane == nl - 1;
    constexpr int bpp = RGB ? 3 : 1;
    const size_t rowbytes = (size_t)W * bpp;
    // lanes beyond the row re-read lane 0's pixels (they take part in the shuffles, their results are never used or stored)
    const uint8_t* base = src + (size_t)n * H * rowbytes + (size_t)(act ? lane : 0) * 16 * bpp;
    const int row_stride16 = W >> 4;
    uint16_t* wrow = weak + (size_t)n * H * row_stride16 + lane;
    uint16_t* srow = strong + (size_t)n * H * row_stride16 + lane;
    const int T = BR + 4;                                 // gray rows y0 - 2 .. y0 + BR + 1 (clamped)

    int g0[16], g1[16], g2[16], M0[18], M1[18];
    RowMasks m0 = {}, m1 = {};
#pragma unroll
    for (int j = 0; j < 18; ++j) { M0[j] = 0; M1[j] = 0; }
#pragma unroll
    for (int j = 0; j < 16; ++j) { g0[j] = 0; g1[j] = 0; g2[j] = 0; }

    auto rowptr = [&](int t) { return base + (size_t)min(max(y0 - 2 + t, 0), H - 1) * rowbytes; };
    RawRow nxt = canny_load_row<RGB>(rowptr(0));
    // two steps per trip: magnitude rows and masks alternate between two register sets, the gray rows are renamed by 32 moves
    // per trip.  (Six steps per trip would make every rotation a pure renaming, but that loop body is 76 KB of code: with the
    // warps of an SM at different places in it, 46 % of the stall samples were instruction-cache misses; two steps are 26 KB.)
#define EEL_CANNY_STEP(tt, GA, GB, GC, MB_, MC_, mB_, mC_)                                                                       \
    {                                                                                                                            \
        if ((tt) >= T) break;                                                                                                    \
        const RawRow cur = nxt;                                                                                                  \
        nxt = canny_load_row<RGB>(rowptr((tt) + 1)); /* (clamped: the last prefetch re-reads a valid row) */                     \
        canny_row_step<RGB>((tt), cur, GA, GB, GC, MB_, MC_, mB_, mC_, y0, H, low, high, first, last, act, wrow, srow, row_stride16); \
    }
#pragma unroll 1
    for (int t = 0; t < T; t += 2) {
        EEL_CANNY_STEP(t, g0, g1, g2, M1, M0, m1, m0)          // rows t-2, t-1 -> row t in g2
        EEL_CANNY_STEP(t + 1, g1, g2, g0, M0, M1, m0, m1)      // rows t-1, t -> row t+1 in g0
#pragma unroll
        for (int j = 0; j < 16; ++j) { g1[j] = g0[j]; g0[j] = g2[j]; }
    }
#undef EEL_CANNY_STEP
}

constexpr int kHystWords = 16;        // words per lane = 512 pixels per row
constexpr int kFloodWarps = 16;       // warps per CTA of the per-image fix-up kernel
constexpr int kMaxBands = 1024;       // 32-row bands per image the fix-up kernel can track (H <= 32768)

// one bitmap row of WW words into 16 registers (words beyond the row are zero); CG: bypass L1 (the row may have been rewritten by
// another warp of this kernel)
template <bool CG>
__device__ __forceinline__ void bitmap_row_load(const uint32_t* __restrict__ p, const int WW, const bool valid, uint32_t (&v)[kHystWords]) {
    if ((WW & 3) == 0) {
#pragma unroll
        for (int q = 0; q < kHystWords / 4; ++q) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (valid && 4 * q < WW) t = CG ? __ldcg(reinterpret_cast<const uint4*>(p) + q) : __ldg(reinterpret_cast<const uint4*>(p) + q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kHystWords; ++k) v[k] = (valid && k < WW) ? (CG ? __ldcg(p + k) : __ldg(p + k)) : 0u;
    }
}

__device__ __forceinline__ void bitmap_row_store(uint32_t* __restrict__ p, const int WW, const uint32_t (&v)[kHystWords]) {
    if ((WW & 3) == 0) {
#pragma unroll
        for (int q = 0; q < kHystWords / 4; ++q)
            if (4 * q < WW) __stcg(reinterpret_cast<uint4*>(p) + q, make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else {
#pragma unroll
        for (int k = 0; k < kHystWords; ++k)
            if (k < WW) __stcg(p + k, v[k]);
    }
}

// Flood step of a 32-row band held by one warp (lane = row, 16 words = 512 pixels per lane; w: candidates, rw: bit-reversed
// candidates, s: set pixels, hu / hd: the rows just above lane 0 / below lane 31), repeated until nothing changes inside the band.
// One step: seeds = candidates next to a set pixel (8-neighbourhood, across rows through shuffles, across words through the
// neighbouring words' end bits), then every seed runs through its whole run of candidates INSIDE its word in both directions:
// (w & ~(w + sd)) | sd towards higher x, the same on the bit-reversed word towards lower x.  The 16 words are independent (a run
// that crosses a word border continues in the next step).  Returns the lane's changed bits (OR over the words).
__device__ __forceinline__ uint32_t canny_flood_steps(const uint32_t (&w)[kHystWords], const uint32_t (&rw)[kHystWords],
                                                      uint32_t (&s)[kHystWords], const uint32_t (&hu)[kHystWords],
                                                      const uint32_t (&hd)[kHystWords], const int lane) {
    uint32_t changed = 0u;
    for (;;) {
        uint32_t nb[kHystWords];
#pragma unroll
        for (int k = 0; k < kHystWords; ++k) {
            uint32_t up = __shfl_up_sync(0xffffffffu, s[k], 1), dn = __shfl_down_sync(0xffffffffu, s[k], 1);
            if (lane == 0) up = hu[k];
            if (lane == 31) dn = hd[k];
            nb[k] = s[k] | up | dn;
        }
        uint32_t diff = 0u;
#pragma unroll
        for (int k = 0; k < kHystWords; ++k) {
            uint32_t n3 = nb[k] | (nb[k] << 1) | (nb[k] >> 1);
            if (k > 0) n3 |= nb[k - 1] >> 31;
            if (k + 1 < kHystWords) n3 |= nb[k + 1] << 31;
            const uint32_t sd = s[k] | (w[k] & n3);
            const uint32_t f = (w[k] & ~(w[k] + sd)) | sd;               // towards higher x
            const uint32_t rs = __brev(f);
            const uint32_t rf = (rw[k] & ~(rw[k] + rs)) | rs;            // towards lower x
            const uint32_t nf = __brev(rf);
            diff |= nf ^ s[k];
            s[k] = nf;
        }
        if (!__any_sync(0xffffffffu, diff != 0u)) break;
        changed |= diff;
    }
    return changed;
}

// A warp floods the strong bitmap of one 32-row band in global memory through the candidate bitmap until nothing changes inside
// the band; the rows just outside the band only seed it.  Returns (warp-uniform) a 2-bit code: bit 0: the band's first row
// changed, bit 1: its last row changed.  Rows of other bands may be rewritten by other warps while they are read here: every value
// read is a subset of the final answer (bits are only ever set), so the result can only be incomplete, never wrong, and the caller
// iterates until nothing changes.
__device__ __forceinline__ unsigned canny_flood_band(const uint32_t* __restrict__ weak, uint32_t* __restrict__ strong, const int H,
                                                     const int WW, const int band, const int lane) {
    const int row = band * 32 + lane;
    const bool valid = row < H;
    uint32_t w[kHystWords], s[kHystWords];
    bitmap_row_load<false>(weak + (size_t)row * WW, WW, valid, w);
    bitmap_row_load<true>(strong + (size_t)row * WW, WW, valid, s);
    // (the rows just outside the band are requested together with the band: one memory round trip per visit)
    uint32_t hu[kHystWords], hd[kHystWords];
    bitmap_row_load<true>(strong + (size_t)(row - 1) * WW, WW, lane == 0 && row > 0, hu);
    bitmap_row_load<true>(strong + (size_t)(row + 1) * WW, WW, lane == 31 && row + 1 < H, hd);
    uint32_t open = 0u;
#pragma unroll
    for (int k = 0; k < kHystWords; ++k) open |= w[k] & ~s[k];
    if (!__any_sync(0xffffffffu, open != 0u)) return 0u;                 // nothing left to grow into
    uint32_t rw[kHystWords];
#pragma unroll
    for (int k = 0; k < kHystWords; ++k) rw[k] = __brev(w[k]);
    const uint32_t changed = canny_flood_steps(w, rw, s, hu, hd, lane);
    const unsigned rows_changed = __ballot_sync(0xffffffffu, changed != 0u);
    if (rows_changed == 0u) return 0u;
    if (changed != 0u) bitmap_row_store(strong + (size_t)row * WW, WW, s);
    const int last_lane = min(31, H - 1 - band * 32);
    return (rows_changed & 1u) | (((rows_changed >> last_lane) & 1u) << 1);
}

// H <= 512: one CTA per image, thread = (segment g of 16 rows, word column k): the thread keeps its 16 rows x 32 pixels of both
// bitmaps in registers and sweeps them down and up (Gauss-Seidel: a chain runs through the whole column strip in one sweep, and
// through each row's word in both directions by the carry trick); what crosses into another thread's strip -- the first / last row
// of a strip and the first / last bit of every row -- is published in shared memory after each iteration (Jacobi between threads),
// so the number of iterations is the number of strip borders the longest chain crosses (2-8 on the synthetic batches; the
// row-per-lane flood above needs one step per ROW a chain climbs).  The converged strips leave directly as 0 / 255 bytes.
// A thread may read a neighbour's published words while that neighbour is already publishing its next state: bits are only ever
// set, so whatever it reads is a subset of the final answer, and any change raises the flag for another iteration (the loop ends
// only after an iteration in which no thread changed anything, i.e. every thread has seen the final state of its neighbours).
// tests/test_canny_design_cpu.py runs a thread-level transcription of this kernel against the oracle.
constexpr int kSweepRows = 16;
constexpr int kSweepSegs = 32;

__global__ void __launch_bounds__(kSweepSegs * 16) canny_flood_sweep_kernel(const uint32_t* __restrict__ weak, const uint32_t* __restrict__ strong,
                                                                           uint8_t* __restrict__ out, int H, int WW) {
    constexpr int R = kSweepRows;
    __shared__ uint32_t sT[kSweepSegs][16], sB[kSweepSegs][16];           // first / last row of every strip
    __shared__ uint32_t sM[kSweepSegs][16], sL[kSweepSegs][16];           // bit r: last / first pixel of the strip's row r
    __shared__ int flags[3];
    volatile uint32_t (*T)[16] = sT;
    volatile uint32_t (*B)[16] = sB;
    volatile uint32_t (*Mm)[16] = sM;
    volatile uint32_t (*Lm)[16] = sL;
    volatile int* vflags = flags;
    const int k = threadIdx.x & 15, g = threadIdx.x >> 4;
    const int nseg = (H + R - 1) / R;
    const size_t img = (size_t)blockIdx.x * H * WW;
    uint32_t w[R], s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int y = g * R + r;
        const bool in = y < H && k < WW;
        w[r] = in ? __ldg(weak + img + (size_t)y * WW + k) : 0u;
        s[r] = in ? __ldg(strong + img + (size_t)y * WW + k) : 0u;
    }
    auto publish = [&]() {
        uint32_t mm = 0u, lm = 0u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mm |= (s[r] >> 31) << r;
            lm |= (s[r] & 1u) << r;
        }
        T[g][k] = s[0];
        B[g][k] = s[R - 1];
        Mm[g][k] = mm;
        Lm[g][k] = lm;
    };
    if (threadIdx.x < 3) flags[threadIdx.x] = 0;
    publish();
    __syncthreads();
    for (int it = 0;; ++it) {
        const bool hasU = g > 0, hasD = g + 1 < nseg, hasL = k > 0, hasR = k < 15;
        const uint32_t up = hasU ? B[g - 1][k] : 0u;
        const uint32_t dn = hasD ? T[g + 1][k] : 0u;
        const uint32_t upL = (hasU && hasL) ? B[g - 1][k - 1] >> 31 : 0u, upR = (hasU && hasR) ? B[g - 1][k + 1] & 1u : 0u;
        const uint32_t dnL = (hasD && hasL) ? T[g + 1][k - 1] >> 31 : 0u, dnR = (hasD && hasR) ? T[g + 1][k + 1] & 1u : 0u;
        // bit r + 1: row r of the neighbouring column (bit 0: the row above the strip, bit R + 1: the row below);
        // cL / cR bit r: any of rows r - 1, r, r + 1 there = the three neighbours of this strip's first / last pixel of row r
        const uint32_t mLx = ((hasL ? Mm[g][k - 1] : 0u) << 1) | upL | (dnL << (R + 1));
        const uint32_t mRx = ((hasR ? Lm[g][k + 1] : 0u) << 1) | upR | (dnR << (R + 1));
        const uint32_t cL = mLx | (mLx >> 1) | (mLx >> 2);
        const uint32_t cR = mRx | (mRx >> 1) | (mRx >> 2);
        if (threadIdx.x == 0) vflags[(it + 1) % 3] = 0;                  // (last read two barriers ago)
        uint32_t ch = 0u;
#define EEL_SWEEP_ROW(r)                                                                                       \
        {                                                                                                      \
            const uint32_t above = (r) == 0 ? up : s[(r) > 0 ? (r) - 1 : 0];                                  \
            const uint32_t below = (r) == R - 1 ? dn : s[(r) < R - 1 ? (r) + 1 : R - 1];                      \
            const uint32_t nb = above | s[r] | below;                                                         \
            const uint32_t n3 = nb | (nb << 1) | (nb >> 1) | ((cL >> (r)) & 1u) | (((cR >> (r)) & 1u) << 31); \
            const uint32_t sd = s[r] | (w[r] & n3);                                                           \
            const uint32_t f = (w[r] & ~(w[r] + sd)) | sd;                                                    \
            const uint32_t rw = __brev(w[r]), rs = __brev(f);                                                 \
            const uint32_t nf = __brev((rw & ~(rw + rs)) | rs);                                               \
            ch |= nf ^ s[r];                                                                                  \
            s[r] = nf;                                                                                        \
        }
#pragma unroll
        for (int r = 0; r < R; ++r) EEL_SWEEP_ROW(r)
#pragma unroll
        for (int r = R - 1; r >= 0; --r) EEL_SWEEP_ROW(r)
#undef EEL_SWEEP_ROW
        if (ch != 0u) {
            publish();
            vflags[it % 3] = 1;
        }
        __syncthreads();
        if (!vflags[it % 3]) break;
    }
    // ---- the strip as 0 / 255 bytes: 32 pixels = two 16-byte stores per row (the 16 columns of a row are consecutive threads)
    if (k < WW) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int y = g * R + r;
            if (y < H) {
                uint4* o = reinterpret_cast<uint4*>(out + ((size_t)blockIdx.x * H + y) * ((size_t)WW * 32) + k * 32);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t x = (s[r] >> (16 * h + 4 * q)) & 0xfu;
                        v[q] = ((x | (x << 7) | (x << 14) | (x << 21)) & 0x01010101u) * 255u;
                    }
                    o[h] = make_uint4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
}

// first pass over all SMs: every (image, band) once
__global__ void __launch_bounds__(128) canny_flood_bands_kernel(const uint32_t* __restrict__ weak, uint32_t* __restrict__ strong, int N, int H,
                                                              int WW, int nbands) {
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= N * nbands) return;
    const int n = gw / nbands, band = gw - n * nbands;
    const size_t img = (size_t)n * H * WW;
    canny_flood_band(weak + img, strong + img, H, WW, band, threadIdx.x & 31);
}

// fix-up across band borders, one CTA per image: bands whose neighbouring rows changed are flooded again until nothing changes
__global__ void __launch_bounds__(kFloodWarps * 32) canny_flood_image_kernel(const uint32_t* __restrict__ weak, uint32_t* __restrict__ strong,
                                                                            int H, int WW, int nbands) {
    __shared__ int dirty[2][kMaxBands];
    __shared__ int flag;
    const size_t img = (size_t)blockIdx.x * H * WW;
    weak += img;
    strong += img;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = threadIdx.x; b < nbands; b += kFloodWarps * 32) { dirty[0][b] = 1; dirty[1][b] = 0; }
    for (int round = 0;; ++round) {
        const int cur = round & 1, nxt = cur ^ 1;
        __syncthreads();                       // bitmap writes (L2) and dirty marks of the previous round are visible; the flag has been read
        if (threadIdx.x == 0) flag = 0;
        __syncthreads();
        for (int band = warp; band < nbands; band += kFloodWarps) {
            if (!dirty[cur][band]) continue;                             // (warp-uniform)
            __syncwarp();
            if (lane == 0) dirty[cur][band] = 0;
            const unsigned edge = canny_flood_band(weak, strong, H, WW, band, lane);
            if (lane == 0) {
                if ((edge & 1u) && band > 0) { dirty[nxt][band - 1] = 1; flag = 1; }
                if ((edge & 2u) && band + 1 < nbands) { dirty[nxt][band + 1] = 1; flag = 1; }
            }
        }
        __threadfence_block();
        __syncthreads();
        if (!flag) break;
    }
}

// the final bitmap as 0 / 255 bytes, 16 pixels per thread and store
__global__ void __launch_bounds__(256) canny_expand_kernel(const uint32_t* __restrict__ strong, uint8_t* __restrict__ out, long long groups) {
    for (long long gidx = blockIdx.x * 256LL + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * 256) {
        const uint32_t bits = (__ldcg(strong + (gidx >> 1)) >> ((int)(gidx & 1) * 16)) & 0xffffu;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t x = (bits >> (4 * q)) & 0xfu;
            v[q] = ((x | (x << 7) | (x << 14) | (x << 21)) & 0x01010101u) * 255u;
        }
        reinterpret_cast<uint4*>(out)[gidx] = make_uint4(v[0], v[1], v[2], v[3]);
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
    // third generation: register-resident row bands + bitmap flood-fill hysteresis
    const size_t bitmap_bytes = (size_t)H * (W / 32) * 4;
    const int nbands = cdiv(H, 32);
    if (W % 32 == 0 && W <= 32 * kHystWords && nbands <= kMaxBands && ((uintptr_t)src % 16) == 0 && ((uintptr_t)edges % 16) == 0 &&
        ((uintptr_t)ws % 16) == 0) {
        uint16_t* weak = (uint16_t*)ws;
        uint16_t* strong = (uint16_t*)((uint8_t*)ws + (size_t)N * bitmap_bytes);
        // rows per warp: every band costs 4 extra gray rows, so the bands are as long as still fills the GPU once (one wave)
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, canny_rows_kernel<RGB>, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
        const long long capacity = (long long)kNumSMs * per_sm * 4;
        int BR = 32;
        while (BR > 8 && (long long)N * cdiv(H, BR / 2) <= capacity) BR >>= 1;
        const int bands = cdiv(H, BR);
        canny_rows_kernel<RGB><<<cdiv((long long)N * bands, 4), 128, 0, st>>>(src, weak, strong, N, H, W, BR, bands, low, high);
        if (int rc = check_launch("canny.rows")) return rc;
        if (H <= kSweepRows * kSweepSegs) {
            canny_flood_sweep_kernel<<<N, kSweepSegs * 16, 0, st>>>((const uint32_t*)weak, (const uint32_t*)strong, edges, H, W / 32);
            return check_launch("canny.flood_sweep");
        }
        canny_flood_bands_kernel<<<cdiv((long long)N * nbands, 4), 128, 0, st>>>((const uint32_t*)weak, (uint32_t*)strong, N, H, W / 32, nbands);
        if (int rc = check_launch("canny.flood_bands")) return rc;
        canny_flood_image_kernel<<<N, kFloodWarps * 32, 0, st>>>((const uint32_t*)weak, (uint32_t*)strong, H, W / 32, nbands);
        if (int rc = check_launch("canny.flood_image")) return rc;
        const long long groups = P / 16;
        canny_expand_kernel<<<(int)std::min<long long>(cdiv(groups, 256), (long long)kNumSMs * 16), 256, 0, st>>>((const uint32_t*)strong, edges, groups);
        return check_launch("canny.expand");
    }
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
