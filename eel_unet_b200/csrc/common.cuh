// Shared helpers for the eel_unet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eel.h"

namespace eel {

// ---- error plumbing: the C-ABI never throws; it returns a code and keeps a message -------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);

#define EEL_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::eel::set_error(__VA_ARGS__);                       \
            return EEL_ERR_INVALID;                              \
        }                                                        \
    } while (0)

// ---- storage-type helpers ---------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <class T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
    __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <> struct Vec16<bf16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ float get(int i) const {
        uint32_t w = (&raw.x)[i >> 1];
        return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
    }
    __device__ __forceinline__ void set(int i, float v) {
        uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
        uint32_t& w = (&raw.x)[i >> 1];
        w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
    }
};
template <class T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
    Vec16<T> v;
    v.raw = *reinterpret_cast<const decltype(v.raw)*>(p);
    return v;
}
template <class T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& v) {
    *reinterpret_cast<decltype(Vec16<T>::raw)*>(p) = v.raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE (per-context) function attribute: a process that drives several
// GPUs must opt in once on each of them.  One of these per (kernel instantiation); remembers the largest size set per device.
struct SmemOptIn {
    int have[32] = {};
    template <class F> bool ensure(F fn, int bytes) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return false;
        if (dev < 32 && have[dev] >= bytes) return true;
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
        if (dev < 32) have[dev] = bytes;       // (a benign race between host threads: the attribute call is idempotent)
        return true;
    }
};

// dispatch on the storage dtype
#define EEL_DISPATCH_DTYPE(dtype, ...)                                  \
    do {                                                                \
        if ((dtype) == EEL_F32) {                                       \
            typedef float T;                                            \
            __VA_ARGS__;                                                \
        } else if ((dtype) == EEL_BF16) {                               \
            typedef ::eel::bf16 T;                                      \
            __VA_ARGS__;                                                \
        } else {                                                        \
            ::eel::set_error("unsupported dtype %d", (int)(dtype));     \
            return EEL_ERR_INVALID;                                     \
        }                                                               \
    } while (0)

}  // namespace eel
