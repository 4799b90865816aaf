/* Plain-C restatement of the integer edge-map path (gray + Canny) the reference gets from OpenCV.
 *
 * TEST INFRASTRUCTURE ONLY -- checker / timed CPU baseline, never a fallback (see oracle/__init__.py).
 * Call sites it stands for: augmentation/AddCannyEdge.py:25-27, augmentation/CannyEnhance.py:32-35,
 * utils/tools.py:145 (cv2.cvtColor(RGB2GRAY) + cv2.Canny(gray, low, high), aperture 3, L1 norm).
 * Algorithm (OpenCV imgproc, restated): 15-bit fixed-point luma; 3x3 Sobel with replicated border;
 * |dx|+|dy| magnitude with a zero ring; non-maximum suppression with the tan(22.5) fixed-point test;
 * hysteresis by depth-first flood fill from the strong pixels over 8-connected weak candidates.
 * Pinned bit-exactly against cv2 and oracle/edge_np.py by tests/test_oracle_c.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void eel_oracle_gray(const uint8_t* rgb, uint8_t* gray, long npix) {
    for (long i = 0; i < npix; ++i)
        gray[i] = (uint8_t)((9798 * rgb[3 * i] + 19235 * rgb[3 * i + 1] + 3735 * rgb[3 * i + 2] + 16384) >> 15);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* gray: H x W, out: H x W (0 / 255). returns 0, or -1 on allocation failure */
int eel_oracle_canny(const uint8_t* gray, uint8_t* out, int H, int W, int low, int high) {
    const int pw = W + 2;
    int* mag = (int*)calloc((size_t)(H + 2) * pw, sizeof(int));
    short* gx = (short*)malloc((size_t)H * W * sizeof(short));
    short* gy = (short*)malloc((size_t)H * W * sizeof(short));
    uint8_t* map = (uint8_t*)malloc((size_t)(H + 2) * pw);
    int* stack = (int*)malloc((size_t)H * W * sizeof(int));
    if (!mag || !gx || !gy || !map || !stack) { free(mag); free(gx); free(gy); free(map); free(stack); return -1; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int ym = clampi(y - 1, 0, H - 1), yp = clampi(y + 1, 0, H - 1), xm = clampi(x - 1, 0, W - 1), xp = clampi(x + 1, 0, W - 1);
            int a = gray[ym * W + xm], b = gray[ym * W + x], c = gray[ym * W + xp];
            int d = gray[y * W + xm], f = gray[y * W + xp];
            int p = gray[yp * W + xm], q = gray[yp * W + x], r = gray[yp * W + xp];
            int dx = (c + 2 * f + r) - (a + 2 * d + p), dy = (p + 2 * q + r) - (a + 2 * b + c);
            gx[y * W + x] = (short)dx; gy[y * W + x] = (short)dy;
            mag[(y + 1) * pw + x + 1] = abs(dx) + abs(dy);
        }
    memset(map, 1, (size_t)(H + 2) * pw);
    int sp = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int* m = mag + (y + 1) * pw + x + 1;
            int v = m[0];
            if (v <= low) continue;
            int dx = gx[y * W + x], dy = gy[y * W + x];
            int ax = abs(dx), ay = abs(dy) << 15, t22 = ax * 13573;
            int keep;
            if (ay < t22) keep = v > m[-1] && v >= m[1];
            else {
                int t67 = t22 + (ax << 16);
                if (ay > t67) keep = v > m[-pw] && v >= m[pw];
                else { int s = (dx ^ dy) < 0 ? -1 : 1; keep = v > m[-pw - s] && v > m[pw + s]; }
            }
            if (!keep) continue;
            if (v > high) { map[(y + 1) * pw + x + 1] = 2; stack[sp++] = (y + 1) * pw + x + 1; }
            else map[(y + 1) * pw + x + 1] = 0;
        }
    while (sp > 0) {
        int i = stack[--sp];
        static const int dxs[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, dys[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
        for (int k = 0; k < 8; ++k) {
            int j = i + dys[k] * pw + dxs[k];
            if (map[j] == 0) { map[j] = 2; stack[sp++] = j; }
        }
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) out[y * W + x] = map[(y + 1) * pw + x + 1] == 2 ? 255 : 0;
    free(mag); free(gx); free(gy); free(map); free(stack);
    return 0;
}

/* batch: rgb [N][H][W][3] -> edges [N][H][W]; scratch gray [H*W] */
int eel_oracle_canny_rgb_batch(const uint8_t* rgb, uint8_t* edges, int N, int H, int W, int low, int high) {
    uint8_t* g = (uint8_t*)malloc((size_t)H * W);
    if (!g) return -1;
    for (int n = 0; n < N; ++n) {
        eel_oracle_gray(rgb + (size_t)n * H * W * 3, g, (long)H * W);
        if (eel_oracle_canny(g, edges + (size_t)n * H * W, H, W, low, high)) { free(g); return -1; }
    }
    free(g);
    return 0;
}
