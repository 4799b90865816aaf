// Bandwidth-bound ops of the hot path: layout conversion, per-channel reductions, BatchNorm
// (statistics / apply / backward), max-pool, add+interleave, GELU, squeeze-excite, Adam.
// Everything moves 16-byte vectors along the contiguous channel axis of NHWC tensors.
#include "common.cuh"

#include <cooperative_groups.h>

#include <stdlib.h>

namespace eel {

static inline int ew_grid(long long work_items, int block) {
    long long b = (work_items + block - 1) / block;
    long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------ layout
template <class T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, long long NHW, int C, long long HW) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NHW; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / HW, r = i - n * HW;
        const float* src = x + n * C * HW + r;
        T* dst = y + i * C;
        for (int c = 0; c < C; ++c) dst[c] = from_f32<T>(src[(long long)c * HW]);
    }
}

template <class TI, class TO>
__global__ void permute4_kernel(const TI* __restrict__ in, TO* __restrict__ out, int d0, int d1, int d2, int d3,
                                int p0, int p1, int p2, int p3) {
    const int d[4] = {d0, d1, d2, d3};
    const int od[4] = {d[p0], d[p1], d[p2], d[p3]};
    const long long istr[4] = {(long long)d1 * d2 * d3, (long long)d2 * d3, d3, 1};
    const long long total = (long long)d0 * d1 * d2 * d3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int o3 = (int)(r % od[3]); r /= od[3];
        int o2 = (int)(r % od[2]); r /= od[2];
        int o1 = (int)(r % od[1]); r /= od[1];
        int o0 = (int)r;
        long long src = o0 * istr[p0] + o1 * istr[p1] + o2 * istr[p2] + o3 * istr[p3];
        out[i] = from_f32<TO>(to_f32(in[src]));
    }
}


// ---- tiled 3-D transposes ------------------------------------------------------------------------------------------------
// Every weight repacking of a step is a permutation of a [A][B][T] view (A, B channel counts, T = kernel taps: 9, 4 or 1):
// reference layout <-> operand layouts, and weight-gradient layout -> reference layout.  The generic index-arithmetic kernels
// above read with strides of T (or A*B) elements -- 4-byte accesses scattered over 36-byte / kilobyte strides, 8 % of the HBM
// roofline.  Here a block moves a 32 x 32 x T tile through shared memory: both the global reads and the global writes run
// along whichever index is contiguous on their side.
struct Tile3 {             // element offsets: a * sa + b * sb + t * st on each side
    long long sa_in, sb_in, st_in, sa_out, sb_out, st_out;
    int A, B, T;
};
constexpr int kTileT = 9, kTilePitch = 33;

// decompose a 4-D permutation into the [A][B][T] form; false when it is not one of the weight patterns (or dims are ragged)
__host__ __device__ inline bool tile3_from_perm(const int (&d)[4], const int (&p)[4], Tile3* t) {
    const long long istr[4] = {(long long)d[1] * d[2] * d[3], (long long)d[2] * d[3], d[3], 1};
    long long ostr[4];          // stride of SOURCE dim k in the output
    {
        long long run = 1;
        for (int o = 3; o >= 0; --o) { ostr[p[o]] = run; run *= d[p[o]]; }
    }
    int ia, ib, t0, t1;         // which source dims play A, B and the (adjacent, in-order on both sides) tap pair
    if (p[0] == 2 && p[1] == 3 && p[2] == 0 && p[3] == 1) { ia = 0; ib = 1; t0 = 2; t1 = 3; }        // [A][B][T] -> [T][A][B]
    else if (p[0] == 2 && p[1] == 3 && p[2] == 1 && p[3] == 0) { ia = 0; ib = 1; t0 = 2; t1 = 3; }   // [A][B][T] -> [T][B][A]
    else if (p[0] == 0 && p[1] == 2 && p[2] == 3 && p[3] == 1) { ia = 0; ib = 1; t0 = 2; t1 = 3; }   // [A][B][T] -> [A][T][B]
    else if (p[0] == 3 && p[1] == 2 && p[2] == 0 && p[3] == 1) { ia = 3; ib = 2; t0 = 0; t1 = 1; }   // [T][B][A] -> [A][B][T]
    else if (p[0] == 0 && p[1] == 3 && p[2] == 1 && p[3] == 2) { ia = 0; ib = 3; t0 = 1; t1 = 2; }   // [A][T][B] -> [A][B][T]
    else return false;
    t->A = d[ia]; t->B = d[ib]; t->T = d[t0] * d[t1];
    if (t->A % 32 != 0 || t->B % 32 != 0 || t->T > kTileT) return false;
    t->sa_in = istr[ia]; t->sb_in = istr[ib]; t->st_in = istr[t1];
    t->sa_out = ostr[ia]; t->sb_out = ostr[ib]; t->st_out = ostr[t1];
    return true;
}

// sm[(t * 32 + a) * 33 + b]
template <class TI, class TO>
__device__ __forceinline__ void tile3_move(const TI* __restrict__ in, TO* __restrict__ out, const Tile3& g, int a0, int b0,
                                           const float* __restrict__ scale, int scale_on_a, float* sm) {
    const int T = g.T, n = 32 * 32 * T;
    // ---- read along the side's contiguous index
    if (g.st_in == 1 && g.sb_in == T) {                       // a-rows of 32 * T contiguous elements
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int a = i / (32 * T), rem = i - a * 32 * T, b = rem / T, t = rem - b * T;
            sm[(t * 32 + a) * kTilePitch + b] = to_f32(in[(a0 + a) * g.sa_in + (long long)b0 * T + rem]);
        }
    } else if (g.sa_in == 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int a = i & 31, b = (i >> 5) & 31, t = i >> 10;
            sm[(t * 32 + a) * kTilePitch + b] = to_f32(in[(a0 + a) + (b0 + b) * g.sb_in + t * g.st_in]);
        }
    } else {                                                  // b contiguous
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int b = i & 31, a = (i >> 5) & 31, t = i >> 10;
            sm[(t * 32 + a) * kTilePitch + b] = to_f32(in[(a0 + a) * g.sa_in + (b0 + b) * g.sb_in + t * g.st_in]);
        }
    }
    __syncthreads();
    // ---- write along the other side's contiguous index
    if (g.st_out == 1 && g.sb_out == T) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int a = i / (32 * T), rem = i - a * 32 * T, b = rem / T, t = rem - b * T;
            float v = sm[(t * 32 + a) * kTilePitch + b];
            if (scale) v *= scale[scale_on_a ? a0 + a : b0 + b];
            out[(a0 + a) * g.sa_out + (long long)b0 * T + rem] = from_f32<TO>(v);
        }
    } else if (g.sa_out == 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int a = i & 31, b = (i >> 5) & 31, t = i >> 10;
            float v = sm[(t * 32 + a) * kTilePitch + b];
            if (scale) v *= scale[scale_on_a ? a0 + a : b0 + b];
            out[(a0 + a) + (b0 + b) * g.sb_out + t * g.st_out] = from_f32<TO>(v);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int b = i & 31, a = (i >> 5) & 31, t = i >> 10;
            float v = sm[(t * 32 + a) * kTilePitch + b];
            if (scale) v *= scale[scale_on_a ? a0 + a : b0 + b];
            out[(a0 + a) * g.sa_out + (b0 + b) * g.sb_out + t * g.st_out] = from_f32<TO>(v);
        }
    }
}

template <class TI, class TO>
__global__ void __launch_bounds__(256) permute_tiled_kernel(const TI* __restrict__ in, TO* __restrict__ out, const Tile3 g) {
    __shared__ float sm[kTileT * 32 * kTilePitch];
    const int tb = g.B / 32, tiles = (g.A / 32) * tb;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();
        tile3_move(in, out, g, (tile / tb) * 32, (tile % tb) * 32, nullptr, 0, sm);
    }
}

// One launch packs a whole table of weights (fp32, reference layout) into their bf16 operand layouts: blockIdx.y = job.
struct PackJob {           // mirrors eel_pack_job in eel.h (64 bytes)
    const float* src;
    bf16* dst;
    int d[4];
    int p[4];
    const float* scale;    // optional: multiply by scale[index of output dimension scale_pos] (BatchNorm folded into the weight)
    int scale_pos;
    int pad;
};
__global__ void __launch_bounds__(256) pack_batch_kernel(const PackJob* __restrict__ jobs) {
    __shared__ float sm[kTileT * 32 * kTilePitch];
    const PackJob j = jobs[blockIdx.y];
    Tile3 g;
    const int dd[4] = {j.d[0], j.d[1], j.d[2], j.d[3]}, pp[4] = {j.p[0], j.p[1], j.p[2], j.p[3]};
    // (reference layout [A][B][taps] -> operand layout; a scaled output dimension must be A or B of the tile view)
    if ((pp[0] == 2 || (pp[0] == 0 && pp[1] == 2)) && (j.scale == nullptr || j.p[j.scale_pos] <= 1) && tile3_from_perm(dd, pp, &g)) {
        const int scale_on_a = j.scale != nullptr && j.p[j.scale_pos] == 0;
        const int tb = g.B / 32, tiles = (g.A / 32) * tb;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            __syncthreads();
            tile3_move(j.src, j.dst, g, (tile / tb) * 32, (tile % tb) * 32, j.scale, scale_on_a, sm);
        }
        return;
    }
    const int od1 = j.d[j.p[1]], od2 = j.d[j.p[2]], od3 = j.d[j.p[3]];
    const long long istr[4] = {(long long)j.d[1] * j.d[2] * j.d[3], (long long)j.d[2] * j.d[3], j.d[3], 1};
    const long long s0 = istr[j.p[0]], s1 = istr[j.p[1]], s2 = istr[j.p[2]], s3 = istr[j.p[3]];
    const long long total = (long long)j.d[0] * j.d[1] * j.d[2] * j.d[3];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int o3 = (int)(r % od3); r /= od3;
        const int o2 = (int)(r % od2); r /= od2;
        const int o1 = (int)(r % od1); r /= od1;
        float v = j.src[r * s0 + o1 * s1 + o2 * s2 + o3 * s3];
        if (j.scale != nullptr) v *= j.scale[j.scale_pos == 0 ? (int)r : (j.scale_pos == 1 ? o1 : (j.scale_pos == 2 ? o2 : o3))];
        j.dst[i] = __float2bfloat16_rn(v);
    }
}

// Inference-time BatchNorm folding: scale[c] = gamma / sqrt(running_var + eps), bias_out[c] = (bias - running_mean) * scale + beta
struct FoldJob {           // mirrors eel_fold_job in eel.h (64 bytes)
    const float* rmean;
    const float* rvar;
    const float* gamma;
    const float* beta;
    const float* bias;     // producer's bias or null
    float* scale;
    float* bias_out;
    int C;
    float eps;
};
__global__ void bn_fold_batch_kernel(const FoldJob* __restrict__ jobs) {
    const FoldJob j = jobs[blockIdx.x];
    for (int c = threadIdx.x; c < j.C; c += blockDim.x) {
        const float sc = j.gamma[c] / sqrtf(j.rvar[c] + j.eps);
        j.scale[c] = sc;
        j.bias_out[c] = ((j.bias ? j.bias[c] : 0.f) - j.rmean[c]) * sc + j.beta[c];
    }
}

// ------------------------------------------------------------------------------------ column reductions
// partial[b][rb][q][c] = sum over the rows of row-block rb (of image-batch b) of f.q-th quantity.
// Threads: TX lanes across channel vectors, TY = 256/TX across rows.
constexpr int kRedThreads = 256;
constexpr int kRedMaxRowBlocks = 4096;
// Row-block targets (blocks per SM), measured on B200 with tools/op_bench.py at the model's shapes: the two-input
// reductions and the BatchNorm backward apply pass run fastest with few, long-lived blocks (less tail and a shorter
// partial-sum finalize); the one-input BatchNorm apply prefers the full 8 x 256 threads per SM.
constexpr int kRedBlocksPerSM = 4;
constexpr int kStreamBpsFwd = 8;
constexpr int kStreamBpsBwd = 2;

struct RedPlan { int TX, TY, ncb, nrb; long long rows_per_rb; };

template <class T> static RedPlan plan_reduce(long long rows, int C, int batch) {
    RedPlan p;
    int cv = C / Vec16<T>::N;
    p.TX = cv >= 32 ? 32 : (cv >= 16 ? 16 : (cv >= 8 ? 8 : (cv >= 4 ? 4 : (cv >= 2 ? 2 : 1))));
    p.TY = kRedThreads / p.TX;
    p.ncb = cdiv(cv, p.TX);
    long long want = (kRedBlocksPerSM * (long long)kNumSMs) / ((long long)p.ncb * batch);
    if (want < 1) want = 1;
    long long maxrb = cdiv(rows, p.TY * 4);
    if (maxrb < 1) maxrb = 1;
    long long nrb = want < maxrb ? want : maxrb;
    if (nrb * batch > kRedMaxRowBlocks) nrb = kRedMaxRowBlocks / batch > 0 ? kRedMaxRowBlocks / batch : 1;
    p.nrb = (int)nrb;
    p.rows_per_rb = (rows + nrb - 1) / nrb;
    return p;
}

// same thread layout for pure streaming kernels: more row blocks (no partial buffer to bound them)
template <class T> static RedPlan plan_stream(long long rows, int C, int blocks_per_sm = kStreamBpsFwd) {
    RedPlan p = plan_reduce<T>(rows, C, 1);
    long long want = (blocks_per_sm * (long long)kNumSMs) / p.ncb;
    if (want < 1) want = 1;
    long long maxrb = cdiv(rows, p.TY * 4);
    if (maxrb < 1) maxrb = 1;
    long long nrb = want < maxrb ? want : maxrb;
    if (nrb > 65535) nrb = 65535;
    p.nrb = (int)nrb;
    p.rows_per_rb = (rows + nrb - 1) / nrb;
    return p;
}

template <class T, class F, int Q>
__global__ void __launch_bounds__(kRedThreads) colreduce_kernel(const F f, long long rows, int C, int TX,
                                                              long long rows_per_rb, float* __restrict__ partial) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 4;   // independent 16-byte loads in flight per thread and input tensor
    __shared__ float sm[kRedThreads][Q * V + 1];
    const int TY = kRedThreads / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cvec = blockIdx.x * TX + tx;
    const int c0 = cvec * V;
    const int rb = blockIdx.y, b = blockIdx.z;
    float acc[Q][V];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[q][v] = 0.f;
    if (c0 < C) {
        typename F::State stt = f.init(c0);
        long long r0 = rb * rows_per_rb, r1 = r0 + rows_per_rb;
        if (r1 > rows) r1 = rows;
        for (long long r = r0 + ty; r < r1; r += (long long)U * TY) {
            Vec16<T> in[U][F::NIN];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + (long long)u * TY < r1) f.load(b * rows + r + (long long)u * TY, c0, in[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + (long long)u * TY < r1) f.template accum<V>(stt, in[u], acc);
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) sm[threadIdx.x][q * V + v] = acc[q][v];
    __syncthreads();
    // thread (tx, ty) sums column tx over ty for a subset of the Q*V values
    for (int j = ty; j < Q * V; j += TY) {
        float s = 0.f;
        for (int y = 0; y < TY; ++y) s += sm[y * TX + tx][j];
        if (c0 < C) {
            int q = j / V, v = j % V;
            partial[(((long long)b * gridDim.y + rb) * Q + q) * C + c0 + v] = s;
        }
    }
}

// out[b][j] = scale * sum_rb partial[b][rb][j]   (fp64 accumulation).  Block = 32 columns x 32 row-slices: the sum over
// row blocks is itself parallel (a serial loop over ~1000 partials per column was a ~100 us latency chain per call).
__global__ void __launch_bounds__(1024) finalize_partials_kernel(const float* __restrict__ partial, int nrb, int width,
                                                               float* __restrict__ out, float scale) {
    __shared__ double sm[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const int b = blockIdx.y;
    double s = 0.0;
    if (j < width)
        for (int r = ty; r < nrb; r += 32 * 8) {      // 8 independent loads in flight (the partials sit in L2: ~1 us each)
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = r + 32 * u < nrb ? partial[((long long)b * nrb + r + 32 * u) * width + j] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < width) {
        double t = 0.0;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sm[y][tx];
        out[(long long)b * width + j] = (float)(t * scale);
    }
}


// BatchNorm backward: sums[j] = sum_rb partial[rb][j] for j in [0, 2C) = {sum g | sum g*xhat}; also written straight to
// dbeta / dgamma (no device-to-device copies) and the optional dz column-sum accumulator is cleared for the apply pass.
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nrb, int C, float* __restrict__ sums,
                                                             float* __restrict__ dbeta, float* __restrict__ dgamma,
                                                             float* __restrict__ dz_colsum) {
    __shared__ double sm[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const int width = 2 * C;
    double s = 0.0;
    if (j < width)
        for (int r = ty; r < nrb; r += 32 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = r + 32 * u < nrb ? partial[(long long)(r + 32 * u) * width + j] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < width) {
        double t = 0.0;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sm[y][tx];
        const float f = (float)t;
        sums[j] = f;
        if (j < C) { dbeta[j] = f; if (dz_colsum != nullptr) dz_colsum[j] = 0.f; }
        else dgamma[j - C] = f;
    }
}

template <class T, class F, int Q>
static int run_colreduce(const F& f, long long rows, int C, int batch, float* partial, size_t ws_bytes, RedPlan& pl,
                         cudaStream_t st, const char* what) {
    if (C % Vec16<T>::N != 0) {
        set_error("%s: channel count %d not a multiple of the vector width", what, C);
        return EEL_ERR_INVALID;
    }
    pl = plan_reduce<T>(rows, C, batch);
    size_t need = sizeof(float) * (size_t)batch * pl.nrb * Q * C;
    if (need > ws_bytes || partial == nullptr) {
        set_error("%s: workspace too small (%zu > %zu)", what, need, ws_bytes);
        return EEL_ERR_WORKSPACE;
    }
    dim3 grid(pl.ncb, pl.nrb, batch);
    colreduce_kernel<T, F, Q><<<grid, kRedThreads, 0, st>>>(f, rows, C, pl.TX, pl.rows_per_rb, partial);
    return check_launch(what);
}

static int run_finalize(const float* partial, int nrb, int width, int batch, float* out, float scale, cudaStream_t st,
                        const char* what) {
    dim3 grid(cdiv(width, 32), batch);
    finalize_partials_kernel<<<grid, 1024, 0, st>>>(partial, nrb, width, out, scale);
    return check_launch(what);
}

struct NoState {};

template <class T> struct SumF {
    const T* x; int C;
    static constexpr int NIN = 1;
    typedef NoState State;
    __device__ State init(int) const { return State{}; }
    __device__ void load(long long r, int c0, Vec16<T> (&in)[1]) const { in[0] = ld16(x + r * C + c0); }
    template <int V> __device__ void accum(const State&, const Vec16<T> (&in)[1], float (&acc)[1][V]) const {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[0][i] += in[0].get(i);
    }
};

// sum of a*b (squeeze-excite backward: d att = sum_hw dout * t)
template <class T> struct DotF {
    const T* a; const T* b; int C;
    static constexpr int NIN = 2;
    typedef NoState State;
    __device__ State init(int) const { return State{}; }
    __device__ void load(long long r, int c0, Vec16<T> (&in)[2]) const { in[0] = ld16(a + r * C + c0); in[1] = ld16(b + r * C + c0); }
    template <int V> __device__ void accum(const State&, const Vec16<T> (&in)[2], float (&acc)[1][V]) const {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[0][i] += in[0].get(i) * in[1].get(i);
    }
};

// the same plus the plain sum of a: {sum a*b, sum a} (squeeze-excite backward: d att and, through sum_hw dout, the column
// sums of the op's input gradient = the bias gradient of the to_patch conv in front of it)
template <class T> struct Dot2F {
    const T* a; const T* b; int C;
    static constexpr int NIN = 2;
    typedef NoState State;
    __device__ State init(int) const { return State{}; }
    __device__ void load(long long r, int c0, Vec16<T> (&in)[2]) const { in[0] = ld16(a + r * C + c0); in[1] = ld16(b + r * C + c0); }
    template <int V> __device__ void accum(const State&, const Vec16<T> (&in)[2], float (&acc)[2][V]) const {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float av = in[0].get(i);
            acc[0][i] = fmaf(av, in[1].get(i), acc[0][i]);
            acc[1][i] += av;
        }
    }
};

// shifted moments for BatchNorm: d = z - z[0][c]  ->  sum d, sum d^2 (no catastrophic cancellation)
template <class T> struct MomentF {
    const T* z; int C;
    static constexpr int NIN = 1;
    struct State { float pivot[Vec16<T>::N]; };
    __device__ State init(int c0) const {
        State s;
        Vec16<T> p = ld16(z + c0);
#pragma unroll
        for (int i = 0; i < Vec16<T>::N; ++i) s.pivot[i] = p.get(i);
        return s;
    }
    __device__ void load(long long r, int c0, Vec16<T> (&in)[1]) const { in[0] = ld16(z + r * C + c0); }
    template <int V> __device__ void accum(const State& s, const Vec16<T> (&in)[1], float (&acc)[2][V]) const {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float d = in[0].get(i) - s.pivot[i];
            acc[0][i] += d;
            acc[1][i] += d * d;
        }
    }
};

// BatchNorm backward sums: g = dy * relu_mask;  sum g, sum g * xhat.  Per-channel constants live in registers.
template <class T> struct BnBwdF {
    const T* dy; const T* z; const float* mean; const float* rstd; const float* gamma; const float* beta; int C; int relu;
    static constexpr int NIN = 2;
    struct State { float m[Vec16<T>::N], r[Vec16<T>::N], g[Vec16<T>::N], b[Vec16<T>::N]; };
    __device__ State init(int c0) const {
        State s;
#pragma unroll
        for (int i = 0; i < Vec16<T>::N; ++i) { s.m[i] = mean[c0 + i]; s.r[i] = rstd[c0 + i]; s.g[i] = gamma[c0 + i]; s.b[i] = beta[c0 + i]; }
        return s;
    }
    __device__ void load(long long r, int c0, Vec16<T> (&in)[2]) const { in[0] = ld16(dy + r * C + c0); in[1] = ld16(z + r * C + c0); }
    template <int V> __device__ void accum(const State& s, const Vec16<T> (&in)[2], float (&acc)[2][V]) const {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float xh = (in[1].get(i) - s.m[i]) * s.r[i];
            float g = in[0].get(i);
            if (relu && !(s.g[i] * xh + s.b[i] > 0.f)) g = 0.f;
            acc[0][i] += g;
            acc[1][i] += g * xh;
        }
    }
};

template <class T>
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int nrb, int C, const T* __restrict__ z,
                                                         double count, float* mean, float* rstd, float* rmean, float* rvar,
                                                         float momentum, float eps) {
    __shared__ double sm1[32][33], sm2[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double s1 = 0.0, s2 = 0.0;
    if (c < C)
        for (int r = ty; r < nrb; r += 32 * 4) {      // 8 independent loads in flight
            float v1[4], v2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = r + 32 * u < nrb;
                v1[u] = ok ? partial[((long long)(r + 32 * u) * 2 + 0) * C + c] : 0.f;
                v2[u] = ok ? partial[((long long)(r + 32 * u) * 2 + 1) * C + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s1 += (double)v1[u]; s2 += (double)v2[u]; }
        }
    sm1[ty][tx] = s1;
    sm2[ty][tx] = s2;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    s1 = 0.0; s2 = 0.0;
#pragma unroll
    for (int y = 0; y < 32; ++y) { s1 += sm1[y][tx]; s2 += sm2[y][tx]; }
    double pivot = (double)to_f32(z[c]);
    double md = s1 / count;
    double var = s2 / count - md * md;
    if (var < 0.0) var = 0.0;
    double m = pivot + md;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean != nullptr) rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * m);
    if (rvar != nullptr) {
        double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unb);
    }
}

// BatchNorm statistics from sums the producing GEMM kernel accumulated in its epilogue (sums = [2][C]: sum z, sum z^2)
__global__ void bn_from_sums_kernel(const float* __restrict__ sums, int C, double count, float* mean, float* rstd, float* rmean,
                                    float* rvar, float momentum, float eps, const float* __restrict__ skipped_bias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = (double)sums[c] / count;
    double var = (double)sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    // the producer left its bias out of z (a per-channel constant cancels in a training-mode BatchNorm): only the running
    // mean, which describes z + bias, has to see it
    const double mb = skipped_bias != nullptr ? m + (double)skipped_bias[c] : m;
    if (rmean != nullptr) rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * mb);
    if (rvar != nullptr) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unb);
    }
}

__global__ void bn_eval_stats_kernel(const float* rmean, const float* rvar, float eps, float* mean, float* rstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rmean[c];
    rstd[c] = 1.0f / sqrtf(rvar[c] + eps);
}

// Row-streaming layout shared by the BatchNorm apply kernels: a thread owns one 16-byte channel vector (its
// per-channel constants stay in registers) and walks down the rows with 4 independent loads in flight.
// grid = (channel-vector blocks, row blocks); block = TX x TY threads.
template <class T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const T* __restrict__ z, T* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, long long rows, int C, int TX,
                                                       long long rows_per_rb, int relu) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 4;
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = (blockIdx.x * TX + tx) * V;
    if (c0 >= C) return;
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = gamma[c0 + j] * rstd[c0 + j];
        sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
    }
    long long r0 = blockIdx.y * rows_per_rb, r1 = r0 + rows_per_rb;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + ty; r < r1; r += (long long)U * TY) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + (long long)u * TY < r1) v[u] = ld16(z + (r + (long long)u * TY) * C + c0);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + (long long)u * TY < r1) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float t = fmaf(v[u].get(j), sc[j], sh[j]);
                    o.set(j, relu ? fmaxf(t, 0.f) : t);
                }
                st16(y + (r + (long long)u * TY) * C + c0, o);
            }
    }
}

// BatchNorm (+ReLU) whose result is stored through ShiftedChannel (models/EELUnet.py:88-97 in front of to_patch, :118):
// y[n,h,w,c] = act(bn(z[n,(h+dh)%H,(w+dw)%W,c])) with the quarter shifts of shift_channels_kernel -- the activation the
// token MLP reads is written in its shifted layout, the separate gather copy (one read + one write of the tensor) disappears.
template <class T>
__global__ void __launch_bounds__(256) bn_act_shift_fwd_kernel(const T* __restrict__ z, T* __restrict__ y, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, unsigned int npix, int H, int W, int C, int TX,
                                                             unsigned int pix_per_block, int relu) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 4;
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c = (blockIdx.x * TX + tx) * V;
    if (c >= C) return;
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = gamma[c + j] * rstd[c + j];
        sh[j] = beta[c + j] - mean[c + j] * sc[j];
    }
    const int q = c / (C / 4);
    const int dh = q == 0 ? -1 : (q == 1 ? 1 : 0), dw = q == 2 ? -1 : 0;
    const unsigned int p0 = blockIdx.y * pix_per_block;
    const unsigned int p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
    for (unsigned int p = p0 + ty; p < p1; p += U * TY) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int pp = p + u * TY;
            if (pp < p1) {
                const unsigned int r = pp / (unsigned int)W, w = pp - r * (unsigned int)W;
                const unsigned int n = r / (unsigned int)H, h = r - n * (unsigned int)H;
                int hh = (int)h + dh, ww = (int)w + dw;
                hh = hh < 0 ? hh + H : (hh >= H ? hh - H : hh);
                ww = ww < 0 ? ww + W : (ww >= W ? ww - W : ww);
                v[u] = ld16(z + ((size_t)(n * (unsigned int)H + (unsigned int)hh) * (unsigned int)W + (unsigned int)ww) * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int pp = p + u * TY;
            if (pp < p1) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float t = fmaf(v[u].get(j), sc[j], sh[j]);
                    o.set(j, relu ? fmaxf(t, 0.f) : t);
                }
                st16(y + (size_t)pp * C + c, o);
            }
        }
    }
}

// dz = gamma*rstd*(g - sum_g/n - xhat*sum_gx/n)  (train)   or   gamma*rstd*g  (frozen statistics)
template <class T>
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, T* __restrict__ dz,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ sums /* [2][C]: dbeta, dgamma */, float inv_count,
                                                       long long rows, int C, int TX, long long rows_per_rb, int relu, int train,
                                                       float* __restrict__ dzsum /* [C] or null: += column sums of dz */) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 4;
    __shared__ float red[256][V + 1];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = (blockIdx.x * TX + tx) * V;
    const bool live = c0 < C;
    float m[V], rs[V], gm[V], bt[V], k1[V], k2[V], cs[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = live ? c0 + j : 0;
        m[j] = mean[c]; rs[j] = rstd[c]; gm[j] = gamma[c]; bt[j] = beta[c];
        k1[j] = train ? sums[c] * inv_count : 0.f;
        k2[j] = train ? sums[C + c] * inv_count : 0.f;
        cs[j] = 0.f;
    }
    // row blocks are walked from the END of the tensor: the reduction pass that ran just before finished there, so the
    // last ~100 MB of dy / z it read are still in the 126 MB L2 when this pass starts
    long long r0 = (long long)(gridDim.y - 1 - blockIdx.y) * rows_per_rb, r1 = r0 + rows_per_rb;
    if (r1 > rows) r1 = rows;
    if (!live) r1 = r0;
    for (long long r = r0 + ty; r < r1; r += (long long)U * TY) {
        Vec16<T> vd[U], vz[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + (long long)u * TY < r1) {
                vd[u] = ld16(dy + (r + (long long)u * TY) * C + c0);
                vz[u] = ld16(z + (r + (long long)u * TY) * C + c0);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + (long long)u * TY < r1) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float xh = (vz[u].get(j) - m[j]) * rs[j];
                    float g = vd[u].get(j);
                    if (relu && !(gm[j] * xh + bt[j] > 0.f)) g = 0.f;
                    const float dzv = gm[j] * rs[j] * (g - k1[j] - xh * k2[j]);
                    cs[j] += dzv;
                    o.set(j, dzv);
                }
                st16(dz + (r + (long long)u * TY) * C + c0, o);
            }
    }
    if (dzsum != nullptr) {   // bias gradient of the producing conv / linear, for free
#pragma unroll
        for (int j = 0; j < V; ++j) red[threadIdx.x][j] = cs[j];
        __syncthreads();
        if (ty == 0 && live) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float t = 0.f;
                for (int y = 0; y < TY; ++y) t += red[y * TX + tx][j];
                atomicAdd(dzsum + c0 + j, t);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ max-pool 2x2
template <class T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec_out, int Ho, int Wo, int C) {
    constexpr int V = Vec16<T>::N;
    const int cv = C / V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_out; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % cv);
        long long p = i / cv;
        int xo = (int)(p % Wo);
        long long r = p / Wo;   // n*Ho + yo
        const T* s = x + ((r * 2) * (2LL * Wo) + 2 * xo) * C + c * V;
        Vec16<T> a = ld16(s), b = ld16(s + C), d = ld16(s + 2LL * Wo * C), e = ld16(s + 2LL * Wo * C + C), o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.set(j, fmaxf(fmaxf(a.get(j), b.get(j)), fmaxf(d.get(j), e.get(j))));
        st16(y + i * V, o);
    }
}

// gradient goes to the first maximum in scan order (ATen max_pool2d semantics)
template <class T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                    long long nvec_out, int Ho, int Wo, int C) {
    constexpr int V = Vec16<T>::N;
    const int cv = C / V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_out; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % cv);
        long long p = i / cv;
        int xo = (int)(p % Wo);
        long long r = p / Wo;
        long long off = ((r * 2) * (2LL * Wo) + 2 * xo) * C + c * V;
        long long o2 = 2LL * Wo * C;
        Vec16<T> a = ld16(x + off), b = ld16(x + off + C), d = ld16(x + off + o2), e = ld16(x + off + o2 + C);
        Vec16<T> g = ld16(dy + i * V), ga, gb, gd, ge;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float va = a.get(j), vb = b.get(j), vd = d.get(j), ve = e.get(j);
            int k = 0; float m = va;
            if (vb > m) { m = vb; k = 1; }
            if (vd > m) { m = vd; k = 2; }
            if (ve > m) { m = ve; k = 3; }
            float gv = g.get(j);
            ga.set(j, k == 0 ? gv : 0.f); gb.set(j, k == 1 ? gv : 0.f);
            gd.set(j, k == 2 ? gv : 0.f); ge.set(j, k == 3 ? gv : 0.f);
        }
        st16(dx + off, ga); st16(dx + off + C, gb); st16(dx + off + o2, gd); st16(dx + off + o2 + C, ge);
    }
}

// ------------------------------------------------------------------------------------ BatchNorm + ReLU + max-pool 2x2
// The encoder stages end in BatchNorm -> ReLU whose result feeds BOTH the 2x2 max-pool (next stage) and the decoder's skip
// bridge (models/EELUnet.py:387-406, 423-459).  Forward: one pass over z writes the activation `a` and the pooled tensor.
// Backward: the two incoming gradients (da from the bridge, dp from the pool) meet INSIDE the BatchNorm backward passes --
// no max-pool backward tensor, no gradient-accumulation pass: the pooled gradient is routed to the recomputed arg-max
// (first maximum in scan order, ATen semantics, on the activation values as stored) while z is read anyway.
// A "row" is a 2x2 window; a thread owns one 16-byte channel vector of it (per-channel constants in registers).
template <class T> struct PoolWin {
    long long base, down;    // element offset of the window's top-left pixel (without the channel offset); +down = next image row
    __device__ PoolWin(long long wi, int Wo, int C) {
        const int xo = (int)(wi % Wo);
        const long long r = wi / Wo;                       // n * Ho + yo
        base = ((r * 2) * (2LL * Wo) + 2 * xo) * C;
        down = 2LL * Wo * C;
    }
    __device__ long long pix(int k, int C) const { return base + (k >> 1) * down + (k & 1) * C; }
};

template <class T>
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const T* __restrict__ z, T* __restrict__ a, T* __restrict__ pooled,
                                                             unsigned short* __restrict__ argmax,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             long long wins, int Wo, int C, int TX, long long wins_per_rb) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 2;
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cvi = blockIdx.x * TX + tx;
    const int c0 = cvi * V;
    if (c0 >= C) return;
    const int cvecs = C / V;
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = gamma[c0 + j] * rstd[c0 + j];
        sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
    }
    long long w0 = blockIdx.y * wins_per_rb, w1 = w0 + wins_per_rb;
    if (w1 > wins) w1 = wins;
    for (long long w = w0 + ty; w < w1; w += (long long)U * TY) {
        Vec16<T> v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (w + (long long)u * TY < w1) {
                const PoolWin<T> pw(w + (long long)u * TY, Wo, C);
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = ld16(z + pw.pix(k, C) + c0);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (w + (long long)u * TY < w1) {
                const PoolWin<T> pw(w + (long long)u * TY, Wo, C);
                Vec16<T> o[4], mx;
                unsigned int idx = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < V; ++j) o[k].set(j, fmaxf(fmaf(v[u][k].get(j), sc[j], sh[j]), 0.f));
#pragma unroll
                for (int j = 0; j < V; ++j) {    // the maximum of the STORED (rounded) activations, first one in scan order
                    float best = o[0].get(j);
                    unsigned int kmax = 0;
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                        const float t = o[k].get(j);
                        if (t > best) { best = t; kmax = k; }
                    }
                    mx.set(j, best);
                    idx |= kmax << (2 * j);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) st16(a + pw.pix(k, C) + c0, o[k]);
                st16(pooled + (w + (long long)u * TY) * C + c0, mx);
                argmax[(w + (long long)u * TY) * cvecs + cvi] = (unsigned short)idx;
            }
    }
}

// With g = (da + [k == arg-max] dp) * [scale * z + shift > 0]:
// APPLY = false: partial[rb][2][C] = {sum g, sum g * z} over the block's windows (the finalize turns the second into
//                sum g * xhat = rstd * (sum g z - mean * sum g)).
// APPLY = true : dz = scale * (g - sum_g / n - xhat * sum_gx / n) = scale * g + A * z + B with per-channel A, B (train);
//                scale * g with frozen statistics.  Two FMAs per element: the kernel is instruction-bound otherwise.
template <class T, bool APPLY>
__global__ void __launch_bounds__(256, 2) bn_relu_pool_bwd_kernel(const T* __restrict__ da, const T* __restrict__ dp, const T* __restrict__ z,
                                                             const unsigned short* __restrict__ argmax,
                                                             T* __restrict__ dz, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ sums,
                                                             float inv_count, long long wins, int Wo, int C, int TX,
                                                             long long wins_per_rb, int train, float* __restrict__ partial,
                                                             float* __restrict__ dzsum) {
    constexpr int V = Vec16<T>::N;
    __shared__ float red[256][2 * V + 1];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cvi = blockIdx.x * TX + tx;
    const int c0 = cvi * V;
    const bool live = c0 < C;
    const int cvecs = C / V;
    float sc[V], sh[V], ka[V], kb[V], acc0[V], acc1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = live ? c0 + j : 0;
        const float m = mean[c], rs = rstd[c];
        sc[j] = gamma[c] * rs;
        sh[j] = beta[c] - m * sc[j];
        const float k1 = (APPLY && train) ? sums[c] * inv_count : 0.f;
        const float k2 = (APPLY && train) ? sums[C + c] * inv_count : 0.f;
        ka[j] = -sc[j] * k2 * rs;
        kb[j] = -sc[j] * k1 - ka[j] * m;
        acc0[j] = acc1[j] = 0.f;
    }
    // the apply pass walks the row blocks from the END (the reduction that ran just before left its tail in L2)
    const long long rb = APPLY ? (long long)(gridDim.y - 1 - blockIdx.y) : (long long)blockIdx.y;
    long long w0 = rb * wins_per_rb, w1 = w0 + wins_per_rb;
    if (w1 > wins) w1 = wins;
    if (!live) w1 = w0;
    for (long long w = w0 + ty; w < w1; w += TY) {
        const PoolWin<T> pw(w, Wo, C);
        Vec16<T> vz[4], vd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { vz[k] = ld16(z + pw.pix(k, C) + c0); vd[k] = ld16(da + pw.pix(k, C) + c0); }
        const Vec16<T> vp = ld16(dp + w * C + c0);
        const unsigned int idx = argmax[w * cvecs + cvi];
        Vec16<T> o[4];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const unsigned int kmax = (idx >> (2 * j)) & 3u;
            const float gp = vp.get(j);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float zz = vz[k].get(j);
                float g = vd[k].get(j);
                if (k == (int)kmax) g += gp;
                if (!(fmaf(zz, sc[j], sh[j]) > 0.f)) g = 0.f;
                if (APPLY) {
                    const float dzv = fmaf(sc[j], g, fmaf(ka[j], zz, kb[j]));
                    acc0[j] += dzv;
                    o[k].set(j, dzv);
                } else {
                    acc0[j] += g;
                    acc1[j] = fmaf(g, zz, acc1[j]);
                }
            }
        }
        if (APPLY) {
#pragma unroll
            for (int k = 0; k < 4; ++k) st16(dz + pw.pix(k, C) + c0, o[k]);
        }
    }
    if (!APPLY || dzsum != nullptr) {
#pragma unroll
        for (int j = 0; j < V; ++j) { red[threadIdx.x][j] = acc0[j]; red[threadIdx.x][V + j] = acc1[j]; }
        __syncthreads();
        if (ty == 0 && live) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float t0 = 0.f, t1 = 0.f;
                for (int y = 0; y < TY; ++y) { t0 += red[y * TX + tx][j]; t1 += red[y * TX + tx][V + j]; }
                if (APPLY) atomicAdd(dzsum + c0 + j, t0);
                else {
                    partial[((long long)blockIdx.y * 2 + 0) * C + c0 + j] = t0;
                    partial[((long long)blockIdx.y * 2 + 1) * C + c0 + j] = t1;
                }
            }
        }
    }
}

// sums = {sum g, sum g * z} -> {sum g, sum g * xhat}, in place and in dgamma (bn_bwd_finalize_kernel wrote the raw values)
__global__ void bn_pool_fix_kernel(float* __restrict__ sums, float* __restrict__ dgamma, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float v = rstd[c] * (sums[C + c] - mean[c] * sums[c]);
    sums[C + c] = v;
    dgamma[c] = v;
}

// ------------------------------------------------------------------------------------ add + interleave
// out[:, 2c] = a[:, c] + b[:, c], out[:, 2c+1] = e[:, c].  mean != null: `a` is a pre-BatchNorm tensor and is normalised on
// the fly (a * gamma * rstd + beta - mean * gamma * rstd, no ReLU: the BatchNorm that ends an upconv block).  The launch uses
// a multiple of C / V threads, so a thread always meets the same channel vector and keeps its constants in registers.
template <class T>
__global__ void add_interleave_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ e,
                                          T* __restrict__ out, long long nvec, int C, const float* __restrict__ mean,
                                          const float* __restrict__ rstd, const float* __restrict__ gamma,
                                          const float* __restrict__ beta) {
    constexpr int V = Vec16<T>::N;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float sc[V], sh[V];
    if (mean != nullptr) {
        const int c0 = (int)(i0 % (C / V)) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            sc[j] = gamma[c0 + j] * rstd[c0 + j];
            sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
    }
    for (long long i = i0; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        Vec16<T> va = ld16(a + i * V), vb = ld16(b + i * V), ve = ld16(e + i * V), o0, o1;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float s = fmaf(va.get(j), sc[j], sh[j]) + vb.get(j);
            if (j < V / 2) { o0.set(2 * j, s); o0.set(2 * j + 1, ve.get(j)); }
            else { o1.set(2 * j - V, s); o1.set(2 * j + 1 - V, ve.get(j)); }
        }
        st16(out + i * 2 * V, o0);
        st16(out + i * 2 * V + V, o1);
    }
}

template <class T>
__global__ void add_interleave_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dab, T* __restrict__ de, long long nvec) {
    constexpr int V = Vec16<T>::N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        Vec16<T> i0 = ld16(dout + i * 2 * V), i1 = ld16(dout + i * 2 * V + V), oa, oe;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (j < V / 2) { oa.set(j, i0.get(2 * j)); oe.set(j, i0.get(2 * j + 1)); }
            else { oa.set(j, i1.get(2 * j - V)); oe.set(j, i1.get(2 * j + 1 - V)); }
        }
        st16(dab + i * V, oa);
        st16(de + i * V, oe);
    }
}


// add + interleave backward that ALSO accumulates the backward sums of the BatchNorms whose whole upstream gradient is dab
// (the BatchNorm that ends the upconv block, fused into the bridge's forward; at the top level also the BatchNorm + ReLU that
// ends the edge branch): their reduction passes over (dab, z) disappear, one extra read of z rides on this pass instead.
// Thread layout of colreduce_kernel (a thread owns one 16-byte channel vector of dab / de, i.e. 32 contiguous bytes of dout,
// and walks down its row block); partial[rb][2 * NBN][C] = per BatchNorm {sum g, sum g * z}, g = dab * [scale z + shift > 0].
template <class T> struct BnSumsRef {
    const T* z; const float* mean; const float* rstd; const float* gamma; const float* beta; float* sums; int relu;
};

template <class T, int NBN>
__global__ void __launch_bounds__(kRedThreads) add_interleave_bwd_bn_kernel(const T* __restrict__ dout, T* __restrict__ dab,
                                                                          T* __restrict__ de, const BnSumsRef<T> b0,
                                                                          const BnSumsRef<T> b1, long long rows, int C, int TX,
                                                                          long long rows_per_rb, float* __restrict__ partial) {
    constexpr int V = Vec16<T>::N;
    constexpr int U = 2;
    constexpr int Q = 2 * NBN;
    __shared__ float sm[kRedThreads][Q * V + 1];
    const int TY = kRedThreads / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = (blockIdx.x * TX + tx) * V;
    const bool live = c0 < C;
    const T* zs[2] = {b0.z, b1.z};
    float sc[NBN][V], sh[NBN][V], acc[Q][V];
#pragma unroll
    for (int k = 0; k < NBN; ++k) {
        const BnSumsRef<T>& b = k == 0 ? b0 : b1;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int c = live ? c0 + j : 0;
            // no ReLU: the mask test (scale * z + shift > 0) must always pass
            sc[k][j] = b.relu ? b.gamma[c] * b.rstd[c] : 0.f;
            sh[k][j] = b.relu ? b.beta[c] - b.mean[c] * sc[k][j] : 1.f;
            acc[2 * k][j] = acc[2 * k + 1][j] = 0.f;
        }
    }
    long long r0 = blockIdx.y * rows_per_rb, r1 = r0 + rows_per_rb;
    if (r1 > rows) r1 = rows;
    if (!live) r1 = r0;
    for (long long r = r0 + ty; r < r1; r += (long long)U * TY) {
        Vec16<T> i0[U], i1[U], vz[U][NBN];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + (long long)u * TY;
            if (rr < r1) {
                i0[u] = ld16(dout + (rr * C + c0) * 2);
                i1[u] = ld16(dout + (rr * C + c0) * 2 + V);
#pragma unroll
                for (int k = 0; k < NBN; ++k) vz[u][k] = ld16(zs[k] + rr * C + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + (long long)u * TY;
            if (rr < r1) {
                Vec16<T> oa, oe;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (j < V / 2) { oa.set(j, i0[u].get(2 * j)); oe.set(j, i0[u].get(2 * j + 1)); }
                    else { oa.set(j, i1[u].get(2 * j - V)); oe.set(j, i1[u].get(2 * j + 1 - V)); }
                }
                st16(dab + rr * C + c0, oa);
                st16(de + rr * C + c0, oe);
#pragma unroll
                for (int k = 0; k < NBN; ++k)
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const float zz = vz[u][k].get(j);
                        float g = oa.get(j);          // the STORED (rounded) gradient: what the apply pass will read
                        if (!(fmaf(zz, sc[k][j], sh[k][j]) > 0.f)) g = 0.f;
                        acc[2 * k][j] += g;
                        acc[2 * k + 1][j] = fmaf(g, zz, acc[2 * k + 1][j]);
                    }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) sm[threadIdx.x][q * V + v] = acc[q][v];
    __syncthreads();
    for (int j = ty; j < Q * V; j += TY) {
        float t = 0.f;
        for (int y = 0; y < TY; ++y) t += sm[y * TX + tx][j];
        if (live) partial[((long long)blockIdx.y * Q + j / V) * C + c0 + j % V] = t;
    }
}

// BatchNorm blockIdx.y: sums = {sum g, sum g * xhat} from partial[rb][Q][C] rows (2 * blockIdx.y, 2 * blockIdx.y + 1) =
// raw {sum g, sum g * z}: sum g * xhat = rstd * (sum g z - mean * sum g)   (fp64 accumulation over the row blocks)
template <class T>
__global__ void __launch_bounds__(1024) bn_bwd_finalize_raw_kernel(const float* __restrict__ partial, int nrb, int C, int Q,
                                                                 const BnSumsRef<T> b0, const BnSumsRef<T> b1) {
    __shared__ double sm0[32][33], sm1[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int k = blockIdx.y;
    const BnSumsRef<T>& b = k == 0 ? b0 : b1;
    double s0 = 0.0, s1 = 0.0;
    if (c < C)
        for (int r = ty; r < nrb; r += 32 * 4) {
            float v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = r + 32 * u < nrb;
                v0[u] = ok ? partial[((long long)(r + 32 * u) * Q + 2 * k) * C + c] : 0.f;
                v1[u] = ok ? partial[((long long)(r + 32 * u) * Q + 2 * k + 1) * C + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s0 += (double)v0[u]; s1 += (double)v1[u]; }
        }
    sm0[ty][tx] = s0;
    sm1[ty][tx] = s1;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    s0 = 0.0; s1 = 0.0;
#pragma unroll
    for (int y = 0; y < 32; ++y) { s0 += sm0[y][tx]; s1 += sm1[y][tx]; }
    b.sums[c] = (float)s0;
    b.sums[C + c] = (float)((double)b.rstd[c] * (s1 - (double)b.mean[c] * s0));
}

// dst[g][r'][:] = src[g][perm(r')][:], perm(r') = 2 r' (r' < rows / 2) or 2 (r' - rows / 2) + 1: the rows of every group
// de-interleaved (even rows first).  Used on the data-gradient operand of the convs that read a skip bridge: their output
// channels then come out as (gradient of the sum | gradient of the skip) halves (eel_tc_conv3x3_dgrad_split).
__global__ void rows_deinterleave_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long nvec, int rows, int vec_per_row) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / vec_per_row;
        const int v = (int)(i - row * vec_per_row);
        const long long g = row / rows;
        const int r = (int)(row - g * rows);
        const int sr = r < rows / 2 ? 2 * r : 2 * (r - rows / 2) + 1;
        dst[i] = src[(g * rows + sr) * vec_per_row + v];
    }
}

// dst[r][c'] = src[r][perm(c')], perm(c') = 2 c' (c' < cols / 2) or 2 (c' - cols / 2) + 1 (2-byte elements): the forward operand
// [ky][kx][co][ci] of a conv that reads a skip bridge, its input channels de-interleaved (eel_tc_conv3x3_2src)
__global__ void cols_deinterleave_kernel(const unsigned short* __restrict__ src, unsigned short* __restrict__ dst, long long n, int cols) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        const int sc = c < cols / 2 ? 2 * c : 2 * (c - cols / 2) + 1;
        dst[i] = src[r * cols + sc];
    }
}

// dw[co][2 c + h][t] = dwp_h[t][c][co]: the two half weight gradients of a two-source conv (taps x C x Cout each, fp32) written
// into the reference layout [Cout][2C][3][3] with the input channels interleaved again
__global__ void dw_interleave_kernel(const float* __restrict__ dwp0, const float* __restrict__ dwp1, float* __restrict__ dw, int taps, int C,
                                     int Cout) {
    const long long n = (long long)Cout * 2 * C * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % taps);
        long long r = i / taps;
        const int ci = (int)(r % (2 * C));
        const int co = (int)(r / (2 * C));
        const float* srcp = (ci & 1) ? dwp1 : dwp0;
        dw[i] = srcp[((long long)t * C + (ci >> 1)) * Cout + co];
    }
}

// out = BatchNorm(z) + b  (the first half of a skip bridge that is never interleaved: eel_tc_conv3x3_2src reads it next to the skip)
template <class T>
__global__ void bn_add_fwd_kernel(const T* __restrict__ z, const T* __restrict__ b, T* __restrict__ out, long long nvec, int C,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                  const float* __restrict__ beta) {
    constexpr int V = Vec16<T>::N;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float sc[V], sh[V];
    const int c0 = (int)(i0 % (C / V)) * V;          // (the launch uses a multiple of C / V threads: a thread keeps its channel vector)
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = gamma[c0 + j] * rstd[c0 + j];
        sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
    }
    for (long long i = i0; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const Vec16<T> vz = ld16(z + i * V), vb = ld16(b + i * V);
        Vec16<T> o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.set(j, fmaf(vz.get(j), sc[j], sh[j]) + vb.get(j));
        st16(out + i * V, o);
    }
}

// ------------------------------------------------------------------------------------ column-block copy (torch.concat on C)
template <class T>
__global__ void copy_cols_kernel(const T* __restrict__ src, long long src_ld, int src_c0, T* __restrict__ dst, long long dst_ld,
                                 int dst_c0, long long nvec, int ncols) {
    constexpr int V = Vec16<T>::N;
    const int cv = ncols / V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        long long p = i / cv;
        int c = (int)(i - p * cv) * V;
        st16(dst + p * dst_ld + dst_c0 + c, ld16(src + p * src_ld + src_c0 + c));
    }
}

// ------------------------------------------------------------------------------------ ShiftedChannel
// y[n,h,w,c] = x[n,(h+dh)%H,(w+dw)%W,c]; quarter 0: dh=-1, quarter 1: dh=+1, quarter 2: dw=-1 (signs flip for the adjoint)
template <class T>
__global__ void __launch_bounds__(256) shift_channels_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned int npix, int H, int W, int C,
                                                           int inverse, int TX, unsigned int pix_per_block) {
    // Row-streaming layout of the BatchNorm apply kernels: a thread owns one 16-byte channel vector -- so its channel quarter
    // and with it its (dh, dw) are fixed -- and walks down the pixels with 4 independent loads in flight; 32-bit index
    // arithmetic (the first version spent ~300 instructions per vector on three 64-bit divisions: 58 % of the HBM roofline).
    constexpr int V = Vec16<T>::N;
    constexpr int U = 4;
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c = (blockIdx.x * TX + tx) * V;
    if (c >= C) return;
    const int q = c / (C / 4);
    int dh = q == 0 ? -1 : (q == 1 ? 1 : 0), dw = q == 2 ? -1 : 0;
    if (inverse) { dh = -dh; dw = -dw; }
    const unsigned int p0 = blockIdx.y * pix_per_block;
    const unsigned int p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
    for (unsigned int p = p0 + ty; p < p1; p += U * TY) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int pp = p + u * TY;
            if (pp < p1) {
                const unsigned int r = pp / (unsigned int)W, w = pp - r * (unsigned int)W;
                const unsigned int n = r / (unsigned int)H, h = r - n * (unsigned int)H;
                int hh = (int)h + dh, ww = (int)w + dw;
                hh = hh < 0 ? hh + H : (hh >= H ? hh - H : hh);
                ww = ww < 0 ? ww + W : (ww >= W ? ww - W : ww);
                v[u] = ld16(x + ((size_t)(n * (unsigned int)H + (unsigned int)hh) * (unsigned int)W + (unsigned int)ww) * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int pp = p + u * TY;
            if (pp < p1) st16(y + (size_t)pp * C + c, v[u]);
        }
    }
}

// ------------------------------------------------------------------------------------ ReLU
template <class T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec) {
    constexpr int V = Vec16<T>::N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        Vec16<T> v = ld16(x + i * V), o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.set(j, fmaxf(v.get(j), 0.f));
        st16(y + i * V, o);
    }
}

template <class T>
__global__ void relu_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long long nvec) {
    constexpr int V = Vec16<T>::N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        Vec16<T> v = ld16(y + i * V), g = ld16(dy + i * V), o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.set(j, v.get(j) > 0.f ? g.get(j) : 0.f);
        st16(dx + i * V, o);
    }
}

// ------------------------------------------------------------------------------------ GELU (erf)
// cdf(a) = 0.5 (1 + erf(a / sqrt 2)) and e = exp(-a^2 / 2).  fp32 storage: erff.  bf16 storage: Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7, far below bf16 rounding) on the SAME exponential the derivative's density term needs -- the exact
// erff made these kernels ALU-bound at half of the HBM roofline.
template <class T> __device__ __forceinline__ void gelu_cdf(float a, float& cdf, float& e) {
    e = __expf(-0.5f * a * a);
    if (sizeof(T) == 4) {
        cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752f));
    } else {
        const float z = fabsf(a) * 0.70710678118654752f;
        const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
        const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
        const float half_erfc = 0.5f * poly * e;          // 0.5 * erfc(|a| / sqrt 2)
        cdf = a >= 0.f ? 1.0f - half_erfc : half_erfc;
    }
}
// forward only needs the cdf: bf16 storage uses Abramowitz-Stegun 7.1.27 (no exponential, |error| < 5e-4 in erf, i.e. below
// half a bf16 ulp of the result for every |a|)
template <class T> __device__ __forceinline__ float gelu_fwd_value(float a) {
    if (sizeof(T) == 4) return 0.5f * a * (1.0f + erff(a * 0.70710678118654752f));
    const float z = fabsf(a) * 0.70710678118654752f;
    float d = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.078108f, 0.000972f), 0.230389f), 0.278393f), 1.0f);
    d *= d;
    d *= d;
    const float half_erfc = __fdividef(0.5f, d);          // 0.5 * erfc(z)
    return a * (a >= 0.f ? 1.0f - half_erfc : half_erfc);
}

template <class T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec) {
    constexpr int V = Vec16<T>::N, U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += U * stride) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < nvec) v[u] = ld16(x + (i + u * stride) * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) break;
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                o.set(j, gelu_fwd_value<T>(v[u].get(j)));
            }
            st16(y + (i + u * stride) * V, o);
        }
    }
}

template <class T>
__global__ void gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long nvec) {
    constexpr int V = Vec16<T>::N, U = 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += U * stride) {
        Vec16<T> v[U], g[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < nvec) { v[u] = ld16(x + (i + u * stride) * V); g[u] = ld16(dy + (i + u * stride) * V); }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) break;
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float a = v[u].get(j);
                float cdf, e;
                gelu_cdf<T>(a, cdf, e);
                o.set(j, g[u].get(j) * (cdf + a * 0.39894228040143268f * e));
            }
            st16(dx + (i + u * stride) * V, o);
        }
    }
}

// GELU backward that also leaves the per-channel column sums of dx (the bias gradient of the Linear that produced x:
// mlp[0] of ChannelAwarePatchedMLP) -- saves the separate column-sum pass over dx.  The grid stride is a multiple of the
// C / V channel vectors of a row, so a thread stays on ONE channel vector and accumulates in registers.
template <class T>
__global__ void __launch_bounds__(256) gelu_bwd_colsum_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                            float* __restrict__ colsum, long long nvec, int cvec) {
    constexpr int V = Vec16<T>::N, U = 2;
    __shared__ float sm[256][V + 1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += U * stride) {
        Vec16<T> v[U], g[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < nvec) { v[u] = ld16(x + (i + u * stride) * V); g[u] = ld16(dy + (i + u * stride) * V); }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) break;
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float a = v[u].get(j);
                float cdf, e;
                gelu_cdf<T>(a, cdf, e);
                o.set(j, g[u].get(j) * (cdf + a * 0.39894228040143268f * e));
                acc[j] += o.get(j);           // the STORED (rounded) value, as a separate pass over dx would see it
            }
            st16(dx + (i + u * stride) * V, o);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) sm[threadIdx.x][j] = acc[j];
    __syncthreads();
    // threads t, t + cvec, t + 2 cvec, ... of the block share a channel vector (256 % cvec == 0)
    for (int k = threadIdx.x; k < cvec * V; k += blockDim.x) {
        const int cv = k / V, j = k - cv * V;
        float sum = 0.f;
        for (int t = cv; t < 256; t += cvec) sum += sm[t][j];
        atomicAdd(colsum + k, sum);
    }
}

// ------------------------------------------------------------------------------------ squeeze-excite
// one block per image: hid = relu(W1 mean + b1), att = sigmoid(W2 hid + b2)
__global__ void se_mlp_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
                              const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ att,
                              float* __restrict__ hid, int C, int R) {
    extern __shared__ float sh[];   // [R]
    const int n = blockIdx.x;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float s = b1[r];
        for (int c = 0; c < C; ++c) s += w1[r * C + c] * mean[n * C + c];
        s = fmaxf(s, 0.f);
        sh[r] = s;
        hid[n * R + r] = s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = b2[c];
        for (int r = 0; r < R; ++r) s += w2[c * R + r] * sh[r];
        att[n * C + c] = sigmoidf_(s);
    }
}

// out[n][p][c] = t[n][p][c] * att[n][c]  (+ add[n][c] if add != nullptr)
template <class T>
__global__ void scale_rows_kernel(const T* __restrict__ t, const float* __restrict__ att, const float* __restrict__ add,
                                  T* __restrict__ out, long long nvec, long long vec_per_img, int C) {
    constexpr int V = Vec16<T>::N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / vec_per_img;
        int c0 = (int)((i * V) % C);
        Vec16<T> v = ld16(t + i * V), o;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float r = v.get(j) * att[n * C + c0 + j];
            if (add != nullptr) r += add[n * C + c0 + j];
            o.set(j, r);
        }
        st16(out + i * V, o);
    }
}

// Squeeze-excite backward, phase A: one block per image does that image's tiny backward algebra.
// in:  datt[n][c] = sum_hw dout*t ; out: dmean_scaled[n][c] = (W1^T dhid_pre)[c] / HW, dpre[n][c], dhid[n][r]
__global__ void se_mlp_bwd_kernel(const float* __restrict__ datt, const float* __restrict__ att, const float* __restrict__ hid,
                                  const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ dmean_scaled,
                                  float* __restrict__ dpre_out, float* __restrict__ dhid_out, int C, int R, float inv_hw,
                                  int datt_stride) {
    extern __shared__ float sh[];   // dpre[C] then dhid[R]
    float* dpre = sh;
    float* dhid = sh + C;
    const int t = threadIdx.x, n = blockIdx.x;
    for (int c = t; c < C; c += blockDim.x) {
        float a = att[n * C + c];
        float d = datt[n * datt_stride + c] * a * (1.f - a);
        dpre[c] = d;
        dpre_out[n * C + c] = d;
    }
    __syncthreads();
    for (int r = t; r < R; r += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += w2[c * R + r] * dpre[c];
        s = hid[n * R + r] > 0.f ? s : 0.f;
        dhid[r] = s;
        dhid_out[n * R + r] = s;
    }
    __syncthreads();
    for (int c = t; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += w1[r * C + c] * dhid[r];
        dmean_scaled[n * C + c] = s * inv_hw;
    }
}

// phase B: one thread per parameter-gradient element sums its contributions over the images (deterministic order).  The arrays
// that other blocks of a cooperative launch have just written are read with ld.global.cg (L2 is the point of coherence).
__device__ void se_param_grad_element(int i, const float* __restrict__ dpre, const float* __restrict__ dhid, const float* __restrict__ hid,
                                     const float* __restrict__ mean, float* __restrict__ dw1, float* __restrict__ db1,
                                     float* __restrict__ dw2, float* __restrict__ db2, int N, int C, int R,
                                     const float* __restrict__ att, const float* __restrict__ sums, const float* __restrict__ dmean_scaled,
                                     float hw, float* __restrict__ dt_colsum) {
    const int CR = C * R;
    float s = 0.f;
    if (i < CR) {                       // dw2[c][r] = sum_n dpre[n][c] * hid[n][r]
        const int c = i / R, r = i % R;
        for (int n = 0; n < N; ++n) s += __ldcg(dpre + n * C + c) * hid[n * R + r];
        dw2[i] = s;
    } else if (i < 2 * CR) {            // dw1[r][c] = sum_n dhid[n][r] * mean[n][c]
        const int j = i - CR, r = j / C, c = j % C;
        for (int n = 0; n < N; ++n) s += __ldcg(dhid + n * R + r) * mean[n * C + c];
        dw1[j] = s;
    } else if (i < 2 * CR + C) {        // db2[c]
        const int c = i - 2 * CR;
        for (int n = 0; n < N; ++n) s += __ldcg(dpre + n * C + c);
        db2[c] = s;
    } else if (i < 2 * CR + C + R) {    // db1[r]
        const int r = i - 2 * CR - C;
        for (int n = 0; n < N; ++n) s += __ldcg(dhid + n * R + r);
        db1[r] = s;
    } else if (dt_colsum != nullptr && i < 2 * CR + 2 * C + R) {
        // column sums of dt = dout * att + dmean over all pixels: sum_n att[n][c] * (sum_hw dout)[n][c] + HW * dmean[n][c]
        const int c = i - 2 * CR - C - R;
        for (int n = 0; n < N; ++n) s += att[n * C + c] * __ldcg(sums + (long long)n * 2 * C + C + c) + hw * __ldcg(dmean_scaled + n * C + c);
        dt_colsum[c] = s;
    }
}


__global__ void se_param_grad_kernel(const float* __restrict__ dpre, const float* __restrict__ dhid, const float* __restrict__ hid,
                                     const float* __restrict__ mean, float* __restrict__ dw1, float* __restrict__ db1,
                                     float* __restrict__ dw2, float* __restrict__ db2, int N, int C, int R,
                                     const float* __restrict__ att, const float* __restrict__ sums, const float* __restrict__ dmean_scaled,
                                     float hw, float* __restrict__ dt_colsum) {
    se_param_grad_element(blockIdx.x * blockDim.x + threadIdx.x, dpre, dhid, hid, mean, dw1, db1, dw2, db2, N, C, R, att, sums,
                          dmean_scaled, hw, dt_colsum);
}

// ---- squeeze-excite as ONE cooperative launch per direction ---------------------------------------------------------------
// The multi-launch versions below are latency chains: on the model's 33 MB token tensors the four (forward) / five (backward)
// dependent launches take 36 / 64 us where the traffic needs 10 / 20.  Here k blocks per image (k * N <= 2 blocks per SM, so
// that the launch fits beside a weight-gradient kernel of the other stream) reduce their slice, meet at a grid barrier, each
// repeat the image's tiny 64 -> R -> 64 algebra in shared memory, and scale their slice -- which they have just read, so the
// second read comes from L2.  The backward needs a second barrier for the parameter gradients (sums over the images).

template <class T, int Q, class LOADACC>
__device__ __forceinline__ void se_slice_sums(const LOADACC& f, long long row0, long long rows, int C, int TX, float* red /* [256][Q*V+1] */,
                                              float* __restrict__ dst /* [Q][C] */) {
    constexpr int V = Vec16<T>::N;
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    float acc[Q][V];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[q][v] = 0.f;
    for (long long r = ty; r < rows; r += 4LL * TY) {
        Vec16<T> in[4][Q];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + (long long)u * TY < rows) f.load((row0 + r + (long long)u * TY) * C + tx * V, in[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + (long long)u * TY < rows) f.accum(in[u], acc);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) red[threadIdx.x * (Q * V + 1) + q * V + v] = acc[q][v];
    __syncthreads();
    for (int j = ty; j < Q * V; j += TY) {
        float t = 0.f;
        for (int y = 0; y < TY; ++y) t += red[(y * TX + tx) * (Q * V + 1) + j];
        dst[(j / V) * C + tx * V + j % V] = t;
    }
}

template <class T> struct SeSumLoad {
    const T* t;
    __device__ void load(long long off, Vec16<T> (&in)[1]) const { in[0] = ld16(t + off); }
    __device__ void accum(const Vec16<T> (&in)[1], float (&acc)[1][Vec16<T>::N]) const {
#pragma unroll
        for (int i = 0; i < Vec16<T>::N; ++i) acc[0][i] += in[0].get(i);
    }
};
template <class T> struct SeDotLoad {      // {sum dout * t, sum dout}
    const T* dout; const T* t;
    __device__ void load(long long off, Vec16<T> (&in)[2]) const { in[0] = ld16(dout + off); in[1] = ld16(t + off); }
    __device__ void accum(const Vec16<T> (&in)[2], float (&acc)[2][Vec16<T>::N]) const {
#pragma unroll
        for (int i = 0; i < Vec16<T>::N; ++i) {
            const float d = in[0].get(i);
            acc[0][i] = fmaf(d, in[1].get(i), acc[0][i]);
            acc[1][i] += d;
        }
    }
};

// dynamic shared memory: red[256][Q*V+1] floats, then 3*C + R floats of per-image vectors
template <class T>
__global__ void __launch_bounds__(256) se_fwd_coop_kernel(const T* __restrict__ t, const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2, T* __restrict__ out,
                                                        float* __restrict__ mean, float* __restrict__ att, float* __restrict__ hid,
                                                        float* __restrict__ partial /* [N][k][C] */, long long HW, int C, int R, int k) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ float sh[];
    float* red = sh;
    float* s_mean = sh + 256 * (V + 1);
    float* s_att = s_mean + C;
    float* s_hid = s_att + C;
    const int TX = C / V, TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.x / k, part = blockIdx.x % k;
    const long long rows = HW / k, row0 = (long long)n * HW + (long long)part * rows;
    se_slice_sums<T, 1>(SeSumLoad<T>{t}, row0, rows, C, TX, red, partial + (long long)blockIdx.x * C);
    __threadfence();
    cooperative_groups::this_grid().sync();
    const float inv_hw = 1.0f / (float)HW;
    for (int c = threadIdx.x; c < C; c += 256) {
        float m = 0.f;
        for (int p = 0; p < k; ++p) m += __ldcg(partial + ((long long)n * k + p) * C + c);
        m *= inv_hw;
        s_mean[c] = m;
        if (part == 0) mean[(long long)n * C + c] = m;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += 256) {
        float a = b1[r];
        for (int c = 0; c < C; ++c) a += w1[r * C + c] * s_mean[c];
        a = fmaxf(a, 0.f);
        s_hid[r] = a;
        if (part == 0) hid[(long long)n * R + r] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = b2[c];
        for (int r = 0; r < R; ++r) a += w2[c * R + r] * s_hid[r];
        a = sigmoidf_(a);
        s_att[c] = a;
        if (part == 0) att[(long long)n * C + c] = a;
    }
    __syncthreads();
    float av[V];
#pragma unroll
    for (int j = 0; j < V; ++j) av[j] = s_att[tx * V + j];
    // the slice is walked from its END: the reduction finished there, those lines are the most likely to still sit in L2
    for (long long r = rows - 1 - ty; r >= 0; r -= 4LL * TY) {
        Vec16<T> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r - (long long)u * TY >= 0) v[u] = ld16(t + (row0 + r - (long long)u * TY) * C + tx * V);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r - (long long)u * TY >= 0) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) o.set(j, v[u].get(j) * av[j]);
                st16(out + (row0 + r - (long long)u * TY) * C + tx * V, o);
            }
    }
}

// workspace layout as eel_se_bwd: sums[N][2C] | dmean[N][C] | dpre[N][C] | dhid[N][C] | partial[N][k][2][C]
template <class T>
__global__ void __launch_bounds__(256) se_bwd_coop_kernel(const T* __restrict__ t, const T* __restrict__ dout, const float* __restrict__ att,
                                                        const float* __restrict__ hid, const float* __restrict__ mean,
                                                        const float* __restrict__ w1, const float* __restrict__ w2, T* __restrict__ dt,
                                                        float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                                        float* __restrict__ db2, float* __restrict__ dt_colsum, float* __restrict__ sums,
                                                        float* __restrict__ dmean, float* __restrict__ dpre_g, float* __restrict__ dhid_g,
                                                        float* __restrict__ partial, int N, long long HW, int C, int R, int k) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ float sh[];
    float* red = sh;
    float* s_dpre = sh + 256 * (2 * V + 1);
    float* s_dmean = s_dpre + C;
    float* s_dhid = s_dmean + C;
    const int TX = C / V, TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.x / k, part = blockIdx.x % k;
    const long long rows = HW / k, row0 = (long long)n * HW + (long long)part * rows;
    se_slice_sums<T, 2>(SeDotLoad<T>{dout, t}, row0, rows, C, TX, red, partial + (long long)blockIdx.x * 2 * C);
    __threadfence();
    cooperative_groups::this_grid().sync();
    const float inv_hw = 1.0f / (float)HW;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s_dt = 0.f, s_d = 0.f;
        for (int p = 0; p < k; ++p) {
            s_dt += __ldcg(partial + (((long long)n * k + p) * 2 + 0) * C + c);
            s_d += __ldcg(partial + (((long long)n * k + p) * 2 + 1) * C + c);
        }
        const float a = att[(long long)n * C + c];
        const float d = s_dt * a * (1.f - a);
        s_dpre[c] = d;
        if (part == 0) {
            sums[(long long)n * 2 * C + c] = s_dt;
            sums[(long long)n * 2 * C + C + c] = s_d;
            dpre_g[(long long)n * C + c] = d;
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += 256) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a += w2[c * R + r] * s_dpre[c];
        a = hid[(long long)n * R + r] > 0.f ? a : 0.f;
        s_dhid[r] = a;
        if (part == 0) dhid_g[(long long)n * R + r] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int r = 0; r < R; ++r) a += w1[r * C + c] * s_dhid[r];
        a *= inv_hw;
        s_dmean[c] = a;
        if (part == 0) dmean[(long long)n * C + c] = a;
    }
    __syncthreads();
    float av[V], dm[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { av[j] = att[(long long)n * C + tx * V + j]; dm[j] = s_dmean[tx * V + j]; }
    for (long long r = rows - 1 - ty; r >= 0; r -= 4LL * TY) {
        Vec16<T> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r - (long long)u * TY >= 0) v[u] = ld16(dout + (row0 + r - (long long)u * TY) * C + tx * V);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r - (long long)u * TY >= 0) {
                Vec16<T> o;
#pragma unroll
                for (int j = 0; j < V; ++j) o.set(j, fmaf(v[u].get(j), av[j], dm[j]));
                st16(dt + (row0 + r - (long long)u * TY) * C + tx * V, o);
            }
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    const int nout = 2 * C * R + 2 * C + R;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < nout) se_param_grad_element(i, dpre_g, dhid_g, hid, mean, dw1, db1, dw2, db2, N, C, R, att, sums, dmean, (float)HW, dt_colsum);
}

// ------------------------------------------------------------------------------------ Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] + wd * p[i];
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// blocks per image of the cooperative squeeze-excite launches: a power of two that divides HW, leaves every thread row of a
// block at least one row, and keeps the grid at <= 2 blocks per SM (0: use the multi-launch path)
static int se_coop_parts(int N, long long HW, int TY) {
    if (N <= 0 || N > 2 * kNumSMs) return 0;
    int k = 1;
    while (2 * k * N <= 2 * kNumSMs && HW % (2 * k) == 0 && HW / (2 * k) >= TY) k *= 2;
    return k;
}

template <class K> static bool se_coop_fits(K kernel, int blocks, int smem) {
    int dev = 0, per_sm = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess) return false;
    return (long long)per_sm * sms >= blocks;
}


}  // namespace eel

using namespace eel;

#define EEL_VEC_CHECK(T, n, what) \
    EEL_REQUIRE((n) % Vec16<T>::N == 0, what ": element count / channels must be a multiple of the 16-byte vector")

extern "C" {

size_t eel_reduce_workspace_bytes(int channels, int quantities) {
    return sizeof(float) * (size_t)kRedMaxRowBlocks * (size_t)quantities * (size_t)channels;
}

int eel_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int dtype, eel_stream s) {
    EEL_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad argument");
    long long HW = (long long)H * W, NHW = HW * N;
    EEL_DISPATCH_DTYPE(dtype, {
        nchw_to_nhwc_kernel<T><<<ew_grid(NHW, 256), 256, 0, (cudaStream_t)s>>>(x, (T*)y, NHW, C, HW);
        return check_launch("nchw_to_nhwc");
    });
}

int eel_pack_batch(const void* jobs_device, int njobs, int blocks_per_job, eel_stream s) {
    EEL_REQUIRE(jobs_device && njobs > 0 && blocks_per_job > 0 && njobs <= 65535, "pack_batch: bad argument");
    static_assert(sizeof(PackJob) == 64, "PackJob must match eel_pack_job");
    dim3 grid(blocks_per_job, njobs);
    pack_batch_kernel<<<grid, 256, 0, (cudaStream_t)s>>>((const PackJob*)jobs_device);
    return check_launch("pack_batch");
}

int eel_bn_fold_batch(const void* jobs_device, int njobs, eel_stream s) {
    EEL_REQUIRE(jobs_device && njobs > 0 && njobs <= 65535, "bn_fold_batch: bad argument");
    static_assert(sizeof(FoldJob) == 64, "FoldJob must match eel_fold_job");
    bn_fold_batch_kernel<<<njobs, 256, 0, (cudaStream_t)s>>>((const FoldJob*)jobs_device);
    return check_launch("bn_fold_batch");
}

int eel_permute4(const void* in, int in_dtype, void* out, int out_dtype, int d0, int d1, int d2, int d3, int p0,
                 int p1, int p2, int p3, eel_stream s) {
    EEL_REQUIRE(in && out && d0 > 0 && d1 > 0 && d2 > 0 && d3 > 0, "permute4: bad argument");
    int seen = (1 << p0) | (1 << p1) | (1 << p2) | (1 << p3);
    EEL_REQUIRE(p0 >= 0 && p0 < 4 && p1 >= 0 && p1 < 4 && p2 >= 0 && p2 < 4 && p3 >= 0 && p3 < 4 && seen == 15,
                "permute4: not a permutation");
    long long total = (long long)d0 * d1 * d2 * d3;
    int g = ew_grid(total, 256);
    cudaStream_t st = (cudaStream_t)s;
    {   // the weight / weight-gradient patterns go through shared-memory tiles (coalesced on both sides)
        Tile3 t3;
        const int dd[4] = {d0, d1, d2, d3}, pp[4] = {p0, p1, p2, p3};
        if (in_dtype == EEL_F32 && tile3_from_perm(dd, pp, &t3)) {
            const int tiles = (t3.A / 32) * (t3.B / 32);
            const int grid = tiles < kNumSMs * 4 ? tiles : kNumSMs * 4;
            if (out_dtype == EEL_F32) permute_tiled_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, t3);
            else if (out_dtype == EEL_BF16) permute_tiled_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)in, (bf16*)out, t3);
            else { set_error("permute4: unsupported dtype"); return EEL_ERR_INVALID; }
            return check_launch("permute4(tiled)");
        }
    }
    if (in_dtype == EEL_F32 && out_dtype == EEL_F32)
        permute4_kernel<float, float><<<g, 256, 0, st>>>((const float*)in, (float*)out, d0, d1, d2, d3, p0, p1, p2, p3);
    else if (in_dtype == EEL_F32 && out_dtype == EEL_BF16)
        permute4_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)in, (bf16*)out, d0, d1, d2, d3, p0, p1, p2, p3);
    else if (in_dtype == EEL_BF16 && out_dtype == EEL_F32)
        permute4_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)in, (float*)out, d0, d1, d2, d3, p0, p1, p2, p3);
    else if (in_dtype == EEL_BF16 && out_dtype == EEL_BF16)
        permute4_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)in, (bf16*)out, d0, d1, d2, d3, p0, p1, p2, p3);
    else {
        set_error("permute4: unsupported dtype");
        return EEL_ERR_INVALID;
    }
    return check_launch("permute4");
}

int eel_colsum(const void* x, float* out, long long P, int C, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(x && out && P > 0 && C > 0, "colsum: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        RedPlan pl;
        SumF<T> f{(const T*)x, C};
        if (int rc = run_colreduce<T, SumF<T>, 1>(f, P, C, 1, (float*)ws, ws_bytes, pl, (cudaStream_t)s, "colsum")) return rc;
        return run_finalize((const float*)ws, pl.nrb, C, 1, out, 1.0f, (cudaStream_t)s, "colsum.finalize");
    });
}

int eel_bn_stats(const void* z, long long P, int C, float* mean, float* rstd, float* running_mean, float* running_var,
                 float momentum, float eps, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(z && mean && rstd && P > 0 && C > 0, "bn_stats: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        RedPlan pl;
        MomentF<T> f{(const T*)z, C};
        if (int rc = run_colreduce<T, MomentF<T>, 2>(f, P, C, 1, (float*)ws, ws_bytes, pl, (cudaStream_t)s, "bn_stats")) return rc;
        bn_finalize_kernel<T><<<cdiv(C, 32), 1024, 0, (cudaStream_t)s>>>((const float*)ws, pl.nrb, C, (const T*)z, (double)P, mean,
                                                                      rstd, running_mean, running_var, momentum, eps);
        return check_launch("bn_stats.finalize");
    });
}

int eel_bn_stats_from_sums(const float* sums, long long P, int C, float* mean, float* rstd, float* running_mean,
                           float* running_var, float momentum, float eps, const float* skipped_bias, eel_stream s) {
    EEL_REQUIRE(sums && mean && rstd && P > 0 && C > 0, "bn_stats_from_sums: bad argument");
    bn_from_sums_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)s>>>(sums, C, (double)P, mean, rstd, running_mean, running_var, momentum, eps,
                                                                   skipped_bias);
    return check_launch("bn_stats_from_sums");
}

int eel_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean, float* rstd, int C,
                      eel_stream s) {
    EEL_REQUIRE(running_mean && running_var && mean && rstd && C > 0, "bn_eval_stats: bad argument");
    bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)s>>>(running_mean, running_var, eps, mean, rstd, C);
    return check_launch("bn_eval_stats");
}

int eel_bn_act_fwd(const void* z, void* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                   long long P, int C, int relu, int dtype, eel_stream s) {
    EEL_REQUIRE(z && y && mean && rstd && gamma && beta && P > 0 && C > 0, "bn_act_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "bn_act_fwd");
        RedPlan pl = plan_stream<T>(P, C);
        dim3 grid(pl.ncb, pl.nrb);
        bn_act_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)z, (T*)y, mean, rstd, gamma, beta, P, C, pl.TX, pl.rows_per_rb, relu);
        return check_launch("bn_act_fwd");
    });
}

int eel_bn_act_shift_fwd(const void* z, void* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         int N, int H, int W, int C, int relu, int dtype, eel_stream s) {
    EEL_REQUIRE(z && y && mean && rstd && gamma && beta && N > 0 && H > 0 && W > 0 && C > 0, "bn_act_shift_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_REQUIRE(C % (4 * Vec16<T>::N) == 0, "bn_act_shift_fwd: C/4 must be a multiple of the 16-byte vector");
        const long long npix = (long long)N * H * W;
        EEL_REQUIRE(npix < (1LL << 32), "bn_act_shift_fwd: more than 2^32 pixels");
        RedPlan pl = plan_stream<T>(npix, C);
        dim3 grid(pl.ncb, pl.nrb);
        bn_act_shift_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)z, (T*)y, mean, rstd, gamma, beta, (unsigned int)npix, H, W, C,
                                                                      pl.TX, (unsigned int)pl.rows_per_rb, relu);
        return check_launch("bn_act_shift_fwd");
    });
}

int eel_bn_act_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, void* dz, float* dgamma, float* dbeta, float* dz_colsum, long long P, int C, int relu,
                   int train, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(dy && z && mean && rstd && gamma && beta && dz && dgamma && dbeta && P > 0 && C > 0, "bn_act_bwd: bad argument");
    EEL_REQUIRE(ws_bytes >= sizeof(float) * 2 * (size_t)C, "bn_act_bwd: workspace too small");
    EEL_DISPATCH_DTYPE(dtype, {
        RedPlan pl;
        // first 2*C floats of ws hold the finished sums {sum g, sum g*xhat}; partials follow
        float* sums = (float*)ws;
        float* partial = sums + 2 * C;
        BnBwdF<T> f{(const T*)dy, (const T*)z, mean, rstd, gamma, beta, C, relu};
        if (int rc = run_colreduce<T, BnBwdF<T>, 2>(f, P, C, 1, partial, ws_bytes - sizeof(float) * 2 * C, pl, (cudaStream_t)s,
                                                   "bn_act_bwd.reduce")) return rc;
        bn_bwd_finalize_kernel<<<cdiv(2 * C, 32), 1024, 0, (cudaStream_t)s>>>(partial, pl.nrb, C, sums, dbeta, dgamma, dz_colsum);
        if (int rc = check_launch("bn_act_bwd.finalize")) return rc;
        RedPlan ps = plan_stream<T>(P, C, kStreamBpsBwd);
        dim3 grid(ps.ncb, ps.nrb);
        bn_act_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)z, (T*)dz, mean, rstd, gamma, beta, sums,
                                                            1.0f / (float)P, P, C, ps.TX, ps.rows_per_rb, relu, train, dz_colsum);
        return check_launch("bn_act_bwd");
    });
}

int eel_bn_act_bwd_apply(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         const float* sums, void* dz, float* dz_colsum, long long P, int C, int relu, int train, int dtype,
                         eel_stream s) {
    EEL_REQUIRE(dy && z && mean && rstd && gamma && beta && sums && dz && P > 0 && C > 0, "bn_act_bwd_apply: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "bn_act_bwd_apply");
        RedPlan ps = plan_stream<T>(P, C, kStreamBpsBwd);
        dim3 grid(ps.ncb, ps.nrb);
        if (dz_colsum != nullptr && cudaMemsetAsync(dz_colsum, 0, sizeof(float) * C, (cudaStream_t)s) != cudaSuccess) {
            set_error("bn_act_bwd_apply: memset failed");
            return EEL_ERR_CUDA;
        }
        bn_act_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)z, (T*)dz, mean, rstd, gamma, beta, sums,
                                                            1.0f / (float)P, P, C, ps.TX, ps.rows_per_rb, relu, train, dz_colsum);
        return check_launch("bn_act_bwd_apply");
    });
}

int eel_bn_relu_pool_fwd(const void* z, void* a, void* pooled, void* argmax, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, int N, int H, int W, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(z && a && pooled && argmax && mean && rstd && gamma && beta && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 &&
                    W % 2 == 0, "bn_relu_pool_fwd: bad argument (H, W must be even)");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "bn_relu_pool_fwd");
        const long long wins = (long long)N * (H / 2) * (W / 2);
        RedPlan pl = plan_stream<T>(wins, C, kStreamBpsFwd);
        dim3 grid(pl.ncb, pl.nrb);
        bn_relu_pool_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)z, (T*)a, (T*)pooled, (unsigned short*)argmax, mean, rstd,
                                                                  gamma, beta, wins, W / 2, C, pl.TX, pl.rows_per_rb);
        return check_launch("bn_relu_pool_fwd");
    });
}

int eel_bn_relu_pool_bwd(const void* da, const void* dp, const void* z, const void* argmax, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, void* dz, float* dgamma, float* dbeta, float* dz_colsum, int N,
                         int H, int W, int C, int train, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(da && dp && z && argmax && mean && rstd && gamma && beta && dz && dgamma && dbeta && N > 0 && H > 0 && W > 0 && C > 0 &&
                    H % 2 == 0 && W % 2 == 0, "bn_relu_pool_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "bn_relu_pool_bwd");
        const long long wins = (long long)N * (H / 2) * (W / 2);
        const long long P = wins * 4;
        float* sums = (float*)ws;                   // finished {sum g, sum g*xhat}; partials follow
        float* partial = sums + 2 * C;
        RedPlan pr = plan_reduce<T>(wins, C, 1);
        EEL_REQUIRE(ws_bytes >= sizeof(float) * (2 * (size_t)C + (size_t)pr.nrb * 2 * C), "bn_relu_pool_bwd: workspace too small");
        dim3 gr(pr.ncb, pr.nrb);
        bn_relu_pool_bwd_kernel<T, false><<<gr, 256, 0, st>>>((const T*)da, (const T*)dp, (const T*)z, (const unsigned short*)argmax, nullptr,
                                                            mean, rstd, gamma, beta, nullptr, 0.f, wins, W / 2, C, pr.TX, pr.rows_per_rb,
                                                            train, partial, nullptr);
        if (int rc = check_launch("bn_relu_pool_bwd.reduce")) return rc;
        bn_bwd_finalize_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(partial, pr.nrb, C, sums, dbeta, dgamma, dz_colsum);
        if (int rc = check_launch("bn_relu_pool_bwd.finalize")) return rc;
        bn_pool_fix_kernel<<<cdiv(C, 128), 128, 0, st>>>(sums, dgamma, mean, rstd, C);
        if (int rc = check_launch("bn_relu_pool_bwd.fix")) return rc;
        RedPlan ps = plan_stream<T>(wins, C, kStreamBpsBwd);
        dim3 ga(ps.ncb, ps.nrb);
        bn_relu_pool_bwd_kernel<T, true><<<ga, 256, 0, st>>>((const T*)da, (const T*)dp, (const T*)z, (const unsigned short*)argmax, (T*)dz,
                                                           mean, rstd, gamma, beta, sums, 1.0f / (float)P, wins, W / 2, C, ps.TX,
                                                           ps.rows_per_rb, train, nullptr, dz_colsum);
        return check_launch("bn_relu_pool_bwd");
    });
}

int eel_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_fwd: bad argument (H, W must be even)");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "maxpool2_fwd");
        long long nvec = (long long)N * (H / 2) * (W / 2) * C / Vec16<T>::N;
        maxpool2_fwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, nvec, H / 2, W / 2, C);
        return check_launch("maxpool2_fwd");
    });
}

int eel_maxpool2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "maxpool2_bwd");
        long long nvec = (long long)N * (H / 2) * (W / 2) * C / Vec16<T>::N;
        maxpool2_bwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, (T*)dx, nvec, H / 2, W / 2, C);
        return check_launch("maxpool2_bwd");
    });
}

int eel_add_interleave_fwd(const void* a, const void* b, const void* e, void* out, long long P, int C, const float* a_mean,
                           const float* a_rstd, const float* a_gamma, const float* a_beta, int dtype, eel_stream s) {
    EEL_REQUIRE(a && b && e && out && P > 0 && C > 0, "add_interleave_fwd: bad argument");
    EEL_REQUIRE(a_mean == nullptr || (a_rstd && a_gamma && a_beta), "add_interleave_fwd: mean / rstd / gamma / beta go together");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "add_interleave_fwd");
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(a_mean == nullptr || 256 % (C / V) == 0, "add_interleave_fwd: fused BatchNorm needs C / %d to divide 256", V);
        long long nvec = P * C / V;
        add_interleave_fwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)a, (const T*)b, (const T*)e, (T*)out, nvec, C,
                                                                                   a_mean, a_rstd, a_gamma, a_beta);
        return check_launch("add_interleave_fwd");
    });
}

int eel_add_interleave_bwd(const void* dout, void* dab, void* de, long long P, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(dout && dab && de && P > 0 && C > 0, "add_interleave_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "add_interleave_bwd");
        long long nvec = P * C / Vec16<T>::N;
        add_interleave_bwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)dout, (T*)dab, (T*)de, nvec);
        return check_launch("add_interleave_bwd");
    });
}

int eel_add_interleave_bwd_bnsums(const void* dout, void* dab, void* de, long long P, int C,
                                  const void* z0, const float* mean0, const float* rstd0, const float* gamma0, const float* beta0,
                                  int relu0, float* sums0,
                                  const void* z1, const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                                  int relu1, float* sums1, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(dout && dab && de && P > 0 && C > 0, "add_interleave_bwd_bnsums: bad argument");
    EEL_REQUIRE(z0 && mean0 && rstd0 && gamma0 && beta0 && sums0, "add_interleave_bwd_bnsums: the first BatchNorm is incomplete");
    EEL_REQUIRE(z1 == nullptr || (mean1 && rstd1 && gamma1 && beta1 && sums1), "add_interleave_bwd_bnsums: the second BatchNorm is incomplete");
    cudaStream_t st = (cudaStream_t)s;
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "add_interleave_bwd_bnsums");
        const int nbn = z1 != nullptr ? 2 : 1;
        const int Q = 2 * nbn;
        RedPlan pl = plan_reduce<T>(P, C, 1);
        EEL_REQUIRE(ws != nullptr && ws_bytes >= sizeof(float) * (size_t)pl.nrb * Q * C, "add_interleave_bwd_bnsums: workspace too small");
        float* partial = (float*)ws;
        BnSumsRef<T> b0{(const T*)z0, mean0, rstd0, gamma0, beta0, sums0, relu0};
        BnSumsRef<T> b1{(const T*)z1, mean1, rstd1, gamma1, beta1, sums1, relu1};
        dim3 grid(pl.ncb, pl.nrb);
        if (nbn == 2)
            add_interleave_bwd_bn_kernel<T, 2><<<grid, kRedThreads, 0, st>>>((const T*)dout, (T*)dab, (T*)de, b0, b1, P, C, pl.TX,
                                                                            pl.rows_per_rb, partial);
        else
            add_interleave_bwd_bn_kernel<T, 1><<<grid, kRedThreads, 0, st>>>((const T*)dout, (T*)dab, (T*)de, b0, b0, P, C, pl.TX,
                                                                            pl.rows_per_rb, partial);
        if (int rc = check_launch("add_interleave_bwd_bnsums")) return rc;
        bn_bwd_finalize_raw_kernel<T><<<dim3(cdiv(C, 32), nbn), 1024, 0, st>>>(partial, pl.nrb, C, Q, b0, b1);
        return check_launch("add_interleave_bwd_bnsums.finalize");
    });
}

int eel_rows_deinterleave(const void* src, void* dst, long long groups, int rows, long long row_bytes, eel_stream s) {
    EEL_REQUIRE(src && dst && groups > 0 && rows > 0 && rows % 2 == 0 && row_bytes > 0 && row_bytes % 16 == 0,
                "rows_deinterleave: bad argument (even row count, rows of whole 16-byte vectors)");
    const long long nvec = groups * rows * (row_bytes / 16);
    rows_deinterleave_kernel<<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const uint4*)src, (uint4*)dst, nvec, rows, (int)(row_bytes / 16));
    return check_launch("rows_deinterleave");
}

int eel_cols_deinterleave(const void* src, void* dst, long long rows, int cols, eel_stream s) {
    EEL_REQUIRE(src && dst && rows > 0 && cols > 0 && cols % 2 == 0, "cols_deinterleave: bad argument (even column count)");
    const long long n = rows * cols;
    cols_deinterleave_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)s>>>((const unsigned short*)src, (unsigned short*)dst, n, cols);
    return check_launch("cols_deinterleave");
}

int eel_dw_interleave(const float* dwp0, const float* dwp1, float* dw, int taps, int C, int Cout, eel_stream s) {
    EEL_REQUIRE(dwp0 && dwp1 && dw && taps > 0 && C > 0 && Cout > 0, "dw_interleave: bad argument");
    const long long n = (long long)Cout * 2 * C * taps;
    dw_interleave_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)s>>>(dwp0, dwp1, dw, taps, C, Cout);
    return check_launch("dw_interleave");
}

int eel_bn_add_fwd(const void* z, const void* b, void* out, long long P, int C, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, int dtype, eel_stream s) {
    EEL_REQUIRE(z && b && out && mean && rstd && gamma && beta && P > 0 && C > 0, "bn_add_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "bn_add_fwd");
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(256 % (C / V) == 0, "bn_add_fwd: C / %d must divide 256", V);
        long long nvec = P * C / V;
        bn_add_fwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)z, (const T*)b, (T*)out, nvec, C, mean, rstd, gamma, beta);
        return check_launch("bn_add_fwd");
    });
}

int eel_copy_cols(const void* src, long long src_ld, int src_c0, void* dst, long long dst_ld, int dst_c0, long long P, int ncols,
                  int dtype, eel_stream s) {
    EEL_REQUIRE(src && dst && P > 0 && ncols > 0 && src_c0 >= 0 && dst_c0 >= 0, "copy_cols: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        EEL_REQUIRE(ncols % V == 0 && src_c0 % V == 0 && dst_c0 % V == 0 && src_ld % V == 0 && dst_ld % V == 0,
                    "copy_cols: columns and strides must be multiples of the 16-byte vector");
        long long nvec = P * (ncols / V);
        copy_cols_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)src, src_ld, src_c0, (T*)dst, dst_ld, dst_c0, nvec, ncols);
        return check_launch("copy_cols");
    });
}

int eel_shift_channels(const void* x, void* y, int N, int H, int W, int C, int inverse, int dtype, eel_stream s) {
    EEL_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "shift_channels: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_REQUIRE(C % (4 * Vec16<T>::N) == 0, "shift_channels: C/4 must be a multiple of the 16-byte vector");
        const long long npix = (long long)N * H * W;
        EEL_REQUIRE(npix < (1LL << 32), "shift_channels: more than 2^32 pixels");
        RedPlan pl = plan_stream<T>(npix, C);
        dim3 grid(pl.ncb, pl.nrb);
        shift_channels_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, (unsigned int)npix, H, W, C, inverse, pl.TX,
                                                                    (unsigned int)pl.rows_per_rb);
        return check_launch("shift_channels");
    });
}

int eel_relu_fwd(const void* x, void* y, long long n, int dtype, eel_stream s) {
    EEL_REQUIRE(x && y && n > 0, "relu_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, n, "relu_fwd");
        long long nvec = n / Vec16<T>::N;
        relu_fwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, nvec);
        return check_launch("relu_fwd");
    });
}

int eel_relu_bwd(const void* y, const void* dy, void* dx, long long n, int dtype, eel_stream s) {
    EEL_REQUIRE(y && dy && dx && n > 0, "relu_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, n, "relu_bwd");
        long long nvec = n / Vec16<T>::N;
        relu_bwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)y, (const T*)dy, (T*)dx, nvec);
        return check_launch("relu_bwd");
    });
}

int eel_gelu_fwd(const void* x, void* y, long long n, int dtype, eel_stream s) {
    EEL_REQUIRE(x && y && n > 0, "gelu_fwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, n, "gelu_fwd");
        long long nvec = n / Vec16<T>::N;
        gelu_fwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, nvec);
        return check_launch("gelu_fwd");
    });
}

int eel_gelu_bwd(const void* x, const void* dy, void* dx, long long n, int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dx && n > 0, "gelu_bwd: bad argument");
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, n, "gelu_bwd");
        long long nvec = n / Vec16<T>::N;
        gelu_bwd_kernel<T><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, (T*)dx, nvec);
        return check_launch("gelu_bwd");
    });
}

int eel_gelu_bwd_colsum(const void* x, const void* dy, void* dx, float* colsum, long long n, int C, int dtype, eel_stream s) {
    EEL_REQUIRE(x && dy && dx && colsum && n > 0 && C > 0 && n % C == 0, "gelu_bwd_colsum: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    EEL_DISPATCH_DTYPE(dtype, {
        EEL_VEC_CHECK(T, C, "gelu_bwd_colsum");
        const int cvec = C / Vec16<T>::N;
        EEL_REQUIRE(256 % cvec == 0, "gelu_bwd_colsum: %d channels do not tile a 256-thread block", C);
        if (cudaMemsetAsync(colsum, 0, sizeof(float) * C, st) != cudaSuccess) {
            set_error("gelu_bwd_colsum: memset failed");
            return EEL_ERR_CUDA;
        }
        long long nvec = n / Vec16<T>::N;
        long long blocks = (nvec + 511) / 512;
        if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
        gelu_bwd_colsum_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)x, (const T*)dy, (T*)dx, colsum, nvec, cvec);
        return check_launch("gelu_bwd_colsum");
    });
}


int eel_se_fwd(const void* t, const float* w1, const float* b1, const float* w2, const float* b2, void* out,
               float* mean, float* att, float* hid, int N, long long HW, int C, int R, void* ws, size_t ws_bytes,
               int dtype, eel_stream s) {
    EEL_REQUIRE(t && w1 && b1 && w2 && b2 && out && mean && att && hid && N > 0 && HW > 0 && C > 0 && R > 0, "se_fwd: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    EEL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (C % V == 0 && 256 % (C / V) == 0 && C / V <= 256) {
            const int k = se_coop_parts(N, HW, 256 / (C / V));
            const int smem = (int)sizeof(float) * (256 * (V + 1) + 2 * C + R);
            if (k > 0 && smem <= 48 * 1024 && ws != nullptr && ws_bytes >= sizeof(float) * (size_t)N * k * C &&
                se_coop_fits(se_fwd_coop_kernel<T>, N * k, smem)) {
                const T* tp = (const T*)t; T* op = (T*)out; float* part = (float*)ws; long long hw = HW; int kk = k;
                void* args[] = {&tp, &w1, &b1, &w2, &b2, &op, &mean, &att, &hid, &part, &hw, &C, &R, &kk};
                if (cudaLaunchCooperativeKernel((void*)se_fwd_coop_kernel<T>, dim3(N * k), dim3(256), args, smem, st) == cudaSuccess)
                    return check_launch("se_fwd(cooperative)");
                (void)cudaGetLastError();       // (fall through to the multi-launch path)
            }
        }
        RedPlan pl;
        SumF<T> f{(const T*)t, C};
        if (int rc = run_colreduce<T, SumF<T>, 1>(f, HW, C, N, (float*)ws, ws_bytes, pl, st, "se_fwd.mean")) return rc;
        if (int rc = run_finalize((const float*)ws, pl.nrb, C, N, mean, 1.0f / (float)HW, st, "se_fwd.finalize")) return rc;
        se_mlp_kernel<<<N, 64, sizeof(float) * R, st>>>(mean, w1, b1, w2, b2, att, hid, C, R);
        if (int rc = check_launch("se_fwd.mlp")) return rc;
        long long nvec = (long long)N * HW * C / Vec16<T>::N;
        scale_rows_kernel<T><<<ew_grid(nvec, 256), 256, 0, st>>>((const T*)t, att, nullptr, (T*)out, nvec, HW * C / Vec16<T>::N, C);
        return check_launch("se_fwd.scale");
    });
}

int eel_se_bwd(const void* t, const void* dout, const float* att, const float* hid, const float* mean, const float* w1,
               const float* w2, void* dt, float* dw1, float* db1, float* dw2, float* db2, float* dt_colsum, int N, long long HW,
               int C, int R, void* ws, size_t ws_bytes, int dtype, eel_stream s) {
    EEL_REQUIRE(t && dout && att && hid && mean && w1 && w2 && dt && dw1 && db1 && dw2 && db2 && N > 0 && HW > 0 && C > 0 && R > 0,
                "se_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    size_t head = sizeof(float) * 5 * (size_t)N * C;   // sums[N][2C] = {sum dout*t | sum dout}, dmean_scaled[N][C], dpre[N][C], dhid[N][R] (R <= C)
    EEL_REQUIRE(ws_bytes > head && R <= C, "se_bwd: workspace too small");
    EEL_DISPATCH_DTYPE(dtype, {
        float* sums = (float*)ws;
        float* dmean = sums + (size_t)2 * N * C;
        float* dpre = dmean + (size_t)N * C;
        float* dhid = dpre + (size_t)N * C;
        float* partial = dhid + (size_t)N * C;
        constexpr int V = Vec16<T>::N;
        if (C % V == 0 && 256 % (C / V) == 0 && C / V <= 256) {
            const int k = se_coop_parts(N, HW, 256 / (C / V));
            const int smem = (int)sizeof(float) * (256 * (2 * V + 1) + 2 * C + R);
            const int nout = 2 * C * R + 2 * C + R;
            if (k > 0 && smem <= 48 * 1024 && (long long)N * k * 256 >= nout && ws_bytes - head >= sizeof(float) * (size_t)N * k * 2 * C &&
                se_coop_fits(se_bwd_coop_kernel<T>, N * k, smem)) {
                const T* tp = (const T*)t; const T* dp = (const T*)dout; T* dtp = (T*)dt; long long hw = HW; int kk = k;
                void* args[] = {&tp, &dp, &att, &hid, &mean, &w1, &w2, &dtp, &dw1, &db1, &dw2, &db2, &dt_colsum, &sums, &dmean, &dpre, &dhid,
                                &partial, &N, &hw, &C, &R, &kk};
                if (cudaLaunchCooperativeKernel((void*)se_bwd_coop_kernel<T>, dim3(N * k), dim3(256), args, smem, st) == cudaSuccess)
                    return check_launch("se_bwd(cooperative)");
                (void)cudaGetLastError();
            }
        }
        RedPlan pl;
        Dot2F<T> f{(const T*)dout, (const T*)t, C};
        if (int rc = run_colreduce<T, Dot2F<T>, 2>(f, HW, C, N, partial, ws_bytes - head, pl, st, "se_bwd.dot")) return rc;
        if (int rc = run_finalize(partial, pl.nrb, 2 * C, N, sums, 1.0f, st, "se_bwd.finalize")) return rc;
        se_mlp_bwd_kernel<<<N, 64, sizeof(float) * (C + R), st>>>(sums, att, hid, w1, w2, dmean, dpre, dhid, C, R, 1.0f / (float)HW, 2 * C);
        if (int rc = check_launch("se_bwd.mlp")) return rc;
        const int nout = 2 * C * R + 2 * C + R;
        se_param_grad_kernel<<<cdiv(nout, 128), 128, 0, st>>>(dpre, dhid, hid, mean, dw1, db1, dw2, db2, N, C, R, att, sums, dmean,
                                                               (float)HW, dt_colsum);
        if (int rc = check_launch("se_bwd.param_grad")) return rc;
        long long nvec = (long long)N * HW * C / Vec16<T>::N;
        scale_rows_kernel<T><<<ew_grid(nvec, 256), 256, 0, st>>>((const T*)dout, att, dmean, (T*)dt, nvec, HW * C / Vec16<T>::N, C);
        return check_launch("se_bwd.scale");
    });
}

int eel_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, eel_stream s) {
    EEL_REQUIRE(p && g && m && v && n > 0 && step > 0, "adam_step: bad argument");
    float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
    adam_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2);
    return check_launch("adam_step");
}

}  // extern "C"
