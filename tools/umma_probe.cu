// Probe (run on a B200): does a tcgen05 K-major SWIZZLE_128B operand view work when its 8-row groups are NOT on
// 1024-byte boundaries?  Layout under test = ONE dense halo block [18 rows][10 cols][64 ch] written by a single TMA
// box load; tap (dy, dx) of a 3x3 convolution is then the view start = ((1+dy)*10 + 1+dx)*128 B, SBO = 1280 B.
// D = A_view * I  (B = 64x64 identity)  must reproduce the shifted pixels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe tools/umma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../eel_unet_b200/csrc/tc_common.cuh"

namespace eel {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace eel
using namespace eel;
using namespace eel::tc;

constexpr int HR = 18, HC = 10;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int dy, int dx, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 23040 B
    uint8_t* sB = smem + 23552;               // 1024-aligned, 8192 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
    uint64_t* done = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, HR * HC * 128 + 8192);
        tma_load_2d(sA, &tmA, bar, 0, 0);
        tma_load_2d(sB, &tmB, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t a0 = smem_u32(sA) + ((1 + dy) * HC + 1 + dx) * 128;
        for (int k = 0; k < 4; ++k) {
            uint64_t da = make_smem_desc(a0 + k * 32, 16, HC * 128, mode == 1);
            uint64_t db = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024, false);
            umma_bf16(tmem, da, db, idesc, k != 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc, v);
        for (int j = 0; j < 32; ++j) out[r * 64 + cc + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static EncodeTiledFn enc() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (EncodeTiledFn)p;
}

static void tmap2d(CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint32_t bc, uint32_t br) {
    cuuint64_t gd[2] = {cols, rows}, gs[1] = {cols * 2};
    cuuint32_t bx[2] = {bc, br}, es[2] = {1, 1};
    CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    std::vector<__nv_bfloat16> hA(HR * HC * 64), hB(64 * 64);
    std::vector<float> fA(HR * HC * 64);
    unsigned s = 12345;
    for (size_t i = 0; i < hA.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        fA[i] = (float)((int)((s >> 16) & 255) - 128);
        hA[i] = __float2bfloat16(fA[i]);
    }
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB;
    float* dO;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    tmap2d(&tA, dA, 64, HR * HC, 64, HR * HC);
    tmap2d(&tB, dB, 64, 64, 64, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    std::vector<float> hO(128 * 64);
    for (int mode = 0; mode < 2; ++mode)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                cudaMemset(dO, 0xff, 128 * 64 * 4);
                probe<<<1, 128, 40960>>>(tA, tB, dO, dy, dx, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d dy %d dx %d: CUDA error %s\n", mode, dy, dx, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0, bad_rows = 0;
                for (int m = 0; m < 128; ++m) {
                    int rb = 0;
                    const int pr = (m / 8 + 1 + dy) * HC + (m % 8) + 1 + dx;
                    for (int n = 0; n < 64; ++n) rb += hO[m * 64 + n] != fA[pr * 64 + n];
                    bad += rb;
                    bad_rows += rb != 0;
                }
                printf("mode %d (base_offset %s) dy %+d dx %+d : %d mismatching elements in %d rows\n", mode, mode ? "set" : "0", dy, dx,
                       bad, bad_rows);
            }
    return 0;
}
