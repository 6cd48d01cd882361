// Spectrogram representations around the STFT pair, one pass each (HBM-bound elementwise):
//   split: complex64 X -> two real planes      join: two real planes -> complex64 X
//   mode 0  (Re X, Im X)                   STFT.forward(return_type='real_imag')   stft.py:91-94
//   mode 1  (|X|, angle X)                 return_type='mag_phase' / input_type     stft.py:95-110
//   mode 2  (log1p(|X| + eps), angle X)    MetricGAN-OKD's stft / istft wrappers    metricganokd.py:185-195
// and their adjoints (torch convention: the gradient of a complex tensor is dL/dRe + i dL/dIm).
// The planes are elementwise images of X's memory: the caller passes the dense buffers, so the
// (F, T) views with strides (1, F) the reference's `.abs()` / `.angle()` return come for free.
#include <math.h>

#include "brv_common.cuh"

namespace {

constexpr int SF_THREADS = 256;
constexpr int SF_PER_THREAD = 4;

template <int MODE>
__global__ void __launch_bounds__(SF_THREADS)
spec_split_kernel(const float2* __restrict__ X, int64_t n, float eps, float* __restrict__ a,
                  float* __restrict__ b) {
    const int64_t base = ((int64_t)blockIdx.x * SF_THREADS + threadIdx.x) * SF_PER_THREAD;
    if (base + SF_PER_THREAD <= n) {
        // two 16-byte loads, two 16-byte stores per plane pair
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(X + base));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(X + base + 2));
        const float re[4] = {v0.x, v0.z, v1.x, v1.z}, im[4] = {v0.y, v0.w, v1.y, v1.w};
        float oa[4], ob[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (MODE == 0) {
                oa[i] = re[i];
                ob[i] = im[i];
            } else {
                const float m = hypotf(re[i], im[i]);
                oa[i] = MODE == 1 ? m : log1pf(m + eps);
                ob[i] = atan2f(im[i], re[i]);
            }
        }
        *reinterpret_cast<float4*>(a + base) = make_float4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<float4*>(b + base) = make_float4(ob[0], ob[1], ob[2], ob[3]);
    } else {
        for (int64_t i = base; i < n; ++i) {
            const float2 v = __ldg(X + i);
            if (MODE == 0) {
                a[i] = v.x;
                b[i] = v.y;
            } else {
                const float m = hypotf(v.x, v.y);
                a[i] = MODE == 1 ? m : log1pf(m + eps);
                b[i] = atan2f(v.y, v.x);
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ float2 join_one(float a, float b) {
    if (MODE == 0) return make_float2(a, b);
    const float m = MODE == 1 ? a : expm1f(a);
    float s, c;
    sincosf(b, &s, &c);
    return make_float2(m * c, m * s);
}

template <int MODE>
__global__ void __launch_bounds__(SF_THREADS)
spec_join_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                 float2* __restrict__ X) {
    const int64_t base = ((int64_t)blockIdx.x * SF_THREADS + threadIdx.x) * SF_PER_THREAD;
    if (base + SF_PER_THREAD <= n && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
        const float4 va = __ldg(reinterpret_cast<const float4*>(a + base));
        const float4 vb = __ldg(reinterpret_cast<const float4*>(b + base));
        const float2 x0 = join_one<MODE>(va.x, vb.x), x1 = join_one<MODE>(va.y, vb.y);
        const float2 x2 = join_one<MODE>(va.z, vb.z), x3 = join_one<MODE>(va.w, vb.w);
        *reinterpret_cast<float4*>(X + base) = make_float4(x0.x, x0.y, x1.x, x1.y);
        *reinterpret_cast<float4*>(X + base + 2) = make_float4(x2.x, x2.y, x3.x, x3.y);
    } else {
        for (int64_t i = base; i < n && i < base + SF_PER_THREAD; ++i) X[i] = join_one<MODE>(__ldg(a + i), __ldg(b + i));
    }
}

// gX from (ga, gb): mode 1 / 2 need X itself (|X| = 0 gets a zero gradient, as torch's sgn does)
template <int MODE>
__global__ void __launch_bounds__(SF_THREADS)
spec_split_grad_kernel(const float* __restrict__ ga, const float* __restrict__ gb,
                       const float2* __restrict__ X, int64_t n, float eps, float2* __restrict__ gX) {
    for (int64_t i = ((int64_t)blockIdx.x * SF_THREADS + threadIdx.x); i < n;
         i += (int64_t)gridDim.x * SF_THREADS) {
        const float da = ga ? __ldg(ga + i) : 0.f, db = gb ? __ldg(gb + i) : 0.f;
        if (MODE == 0) {
            gX[i] = make_float2(da, db);
            continue;
        }
        const float2 v = __ldg(X + i);
        const float m2 = v.x * v.x + v.y * v.y;
        if (!(m2 > 0.f)) {
            gX[i] = make_float2(0.f, 0.f);
            continue;
        }
        const float m = sqrtf(m2);
        const float dm = MODE == 1 ? da : da / (1.f + m + eps);
        gX[i] = make_float2(dm * v.x / m - db * v.y / m2, dm * v.y / m + db * v.x / m2);
    }
}

// (ga, gb) from gX: X = m(a) e^{i b}
template <int MODE>
__global__ void __launch_bounds__(SF_THREADS)
spec_join_grad_kernel(const float2* __restrict__ gX, const float* __restrict__ a,
                      const float* __restrict__ b, int64_t n, float* __restrict__ ga,
                      float* __restrict__ gb) {
    for (int64_t i = ((int64_t)blockIdx.x * SF_THREADS + threadIdx.x); i < n;
         i += (int64_t)gridDim.x * SF_THREADS) {
        const float2 g = __ldg(gX + i);
        if (MODE == 0) {
            ga[i] = g.x;
            gb[i] = g.y;
            continue;
        }
        const float av = __ldg(a + i);
        const float m = MODE == 1 ? av : expm1f(av);
        float s, c;
        sincosf(__ldg(b + i), &s, &c);
        const float dm = g.x * c + g.y * s;
        ga[i] = MODE == 1 ? dm : dm * expf(av);
        gb[i] = m * (g.y * c - g.x * s);
    }
}

unsigned blocks_vec(int64_t n) { return (unsigned)brv_ceil_div(n, (int64_t)SF_THREADS * SF_PER_THREAD); }
unsigned blocks_loop(int64_t n) {
    const int64_t b = brv_ceil_div(n, SF_THREADS);
    return (unsigned)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace

#define BRV_MODE_SWITCH(KERNEL, GRID, ...)                                             \
    switch (mode) {                                                                    \
        case 0: KERNEL<0><<<GRID, SF_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break; \
        case 1: KERNEL<1><<<GRID, SF_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break; \
        default: KERNEL<2><<<GRID, SF_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break; \
    }

extern "C" int brv_spec_split(const void* X, int64_t n, int mode, float eps, float* a, float* b,
                              void* stream) {
    BRV_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "bad spectrogram split arguments");
    if (n == 0) return BRV_OK;
    BRV_REQUIRE(X && a && b, "null pointer argument");
    BRV_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(a) |
                  reinterpret_cast<uintptr_t>(b)) & 15) == 0, "spectrogram planes must be 16-byte aligned");
    BRV_MODE_SWITCH(spec_split_kernel, blocks_vec(n), (const float2*)X, n, eps, a, b)
    BRV_LAUNCH_CHECK("spec_split_kernel");
    return BRV_OK;
}

extern "C" int brv_spec_join(const float* a, const float* b, int64_t n, int mode, void* X,
                             void* stream) {
    BRV_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "bad spectrogram join arguments");
    if (n == 0) return BRV_OK;
    BRV_REQUIRE(X && a && b, "null pointer argument");
    BRV_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, "spectrogram must be 16-byte aligned");
    BRV_MODE_SWITCH(spec_join_kernel, blocks_vec(n), a, b, n, (float2*)X)
    BRV_LAUNCH_CHECK("spec_join_kernel");
    return BRV_OK;
}

extern "C" int brv_spec_split_grad(const float* ga, const float* gb, const void* X, int64_t n,
                                   int mode, float eps, void* gX, void* stream) {
    BRV_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "bad spectrogram split arguments");
    if (n == 0) return BRV_OK;
    BRV_REQUIRE(gX && (mode == 0 || X), "null pointer argument");
    BRV_MODE_SWITCH(spec_split_grad_kernel, blocks_loop(n), ga, gb, (const float2*)X, n, eps, (float2*)gX)
    BRV_LAUNCH_CHECK("spec_split_grad_kernel");
    return BRV_OK;
}

extern "C" int brv_spec_join_grad(const void* gX, const float* a, const float* b, int64_t n, int mode,
                                  float* ga, float* gb, void* stream) {
    BRV_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "bad spectrogram join arguments");
    if (n == 0) return BRV_OK;
    BRV_REQUIRE(gX && ga && gb && (mode == 0 || (a && b)), "null pointer argument");
    BRV_MODE_SWITCH(spec_join_grad_kernel, blocks_loop(n), (const float2*)gX, a, b, n, ga, gb)
    BRV_LAUNCH_CHECK("spec_join_grad_kernel");
    return BRV_OK;
}

// ---- mean over channels times a real per-bin mask, in one pass ------------------------------
// FFNN._enhance (models/ffnn/ffnn.py:107-110): `x = x.mean(1); x = stft.backward(x * mask)` -- the
// reference (and an eager mirror) runs a complex mean and a broadcast multiply over the whole
// spectrogram.  out[b, t, f] = mask[b, f, t] / C * sum_c X[b, c, f, t], written frame-major (the
// layout the iSTFT kernels stream), X and mask with arbitrary element strides.
namespace {
constexpr int CM_ITEMS = 4;        // (frame, bin) elements per thread, all loads issued first
__global__ void __launch_bounds__(256)
channel_mean_mask_kernel(const float2* __restrict__ X, int64_t xb, int64_t xc, int64_t xf, int64_t xt,
                         const float* __restrict__ mask, int64_t mb, int64_t mf, int64_t mt,
                         int n_channels, int n_bins, int64_t n_frames, float2* __restrict__ out) {
    const int64_t b = blockIdx.y;
    const int64_t total = n_frames * n_bins;
    const float inv_c = 1.f / (float)n_channels;
    const float2* xs = X + b * xb;
    float re[CM_ITEMS], im[CM_ITEMS], m[CM_ITEMS];
#pragma unroll
    for (int k = 0; k < CM_ITEMS; ++k) {
        const int64_t e = ((int64_t)blockIdx.x * CM_ITEMS + k) * 256 + threadIdx.x;
        re[k] = im[k] = 0.f;
        m[k] = inv_c;
        if (e < total) {
            const int64_t t = e / n_bins, f = e - t * n_bins;
            const float2* src = xs + f * xf + t * xt;
            for (int c = 0; c < n_channels; ++c) {
                const float2 v = __ldg(src + (int64_t)c * xc);
                re[k] += v.x;
                im[k] += v.y;
            }
            if (mask) m[k] = __ldg(mask + b * mb + f * mf + t * mt) * inv_c;
        }
    }
#pragma unroll
    for (int k = 0; k < CM_ITEMS; ++k) {
        const int64_t e = ((int64_t)blockIdx.x * CM_ITEMS + k) * 256 + threadIdx.x;
        if (e < total) out[b * total + e] = make_float2(re[k] * m[k], im[k] * m[k]);
    }
}

// The common case -- frame-major spectrogram (bins contiguous) and a bin-major mask (frames
// contiguous: what MelFilterbank.backward returns) -- reads the mask through a 32 x 32 shared
// memory tile so that both streams and the output are coalesced (the element-wise kernel above
// touches one 32-byte sector per mask value there).
__global__ void __launch_bounds__(256)
channel_mean_mask_tiled_kernel(const float2* __restrict__ X, int64_t xb, int64_t xc, int64_t xt,
                               const float* __restrict__ mask, int64_t mb, int64_t mf, int n_channels,
                               int n_bins, int64_t n_frames, float2* __restrict__ out) {
    __shared__ float tile[32][33];                         // [bin][frame]
    const int64_t b = blockIdx.z;
    const int64_t t0 = (int64_t)blockIdx.y * 32;
    const int f0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const float inv_c = 1.f / (float)n_channels;
    const float* mp = mask + b * mb;
#pragma unroll
    for (int r = 0; r < 4; ++r) {                          // rows = bins, lanes along frames
        const int f = f0 + ty + 8 * r;
        const int64_t t = t0 + tx;
        tile[ty + 8 * r][tx] = (f < n_bins && t < n_frames) ? __ldg(mp + (int64_t)f * mf + t) : 0.f;
    }
    __syncthreads();
    const float2* xs = X + b * xb;
    float2 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                          // rows = frames, lanes along bins
        const int64_t t = t0 + ty + 8 * r;
        const int f = f0 + tx;
        acc[r] = make_float2(0.f, 0.f);
        if (f < n_bins && t < n_frames)
            for (int c = 0; c < n_channels; ++c) {
                const float2 v = __ldg(xs + (int64_t)c * xc + t * xt + f);
                acc[r].x += v.x;
                acc[r].y += v.y;
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t t = t0 + ty + 8 * r;
        const int f = f0 + tx;
        if (f < n_bins && t < n_frames) {
            const float m = tile[tx][ty + 8 * r] * inv_c;
            out[(b * n_frames + t) * n_bins + f] = make_float2(acc[r].x * m, acc[r].y * m);
        }
    }
}

// total += mean(v): the running metric of the training loop (training.py:369-373) without two
// ATen launches per step
__global__ void accumulate_mean_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ total) {
    double acc = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)v[i];
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (int)blockDim.x / 32; ++k) t += red[k];
        *total += (float)(t / (double)n);
    }
}
}  // namespace

extern "C" int brv_channel_mean_mask(const void* X, int64_t xb, int64_t xc, int64_t xf, int64_t xt,
                                     const float* mask, int64_t mb, int64_t mf, int64_t mt,
                                     int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                                     void* out, void* stream) {
    BRV_REQUIRE(n_batch >= 0 && n_channels >= 1 && n_bins >= 1 && n_frames >= 0, "bad shape");
    if (n_batch == 0 || n_frames == 0) return BRV_OK;
    BRV_REQUIRE(X && out, "null pointer argument");
    BRV_REQUIRE(n_batch < 65536, "more than 65535 batch items per call");
    BRV_REQUIRE(n_frames * n_bins < (1LL << 40), "spectrogram too large");
    if (mask && xf == 1 && mt == 1 && brv_ceil_div(n_frames, 32) < 65536) {
        dim3 tgrid((unsigned)brv_ceil_div(n_bins, 32), (unsigned)brv_ceil_div(n_frames, 32), (unsigned)n_batch);
        channel_mean_mask_tiled_kernel<<<tgrid, 256, 0, (cudaStream_t)stream>>>(
            (const float2*)X, xb, xc, xt, mask, mb, mf, n_channels, n_bins, n_frames, (float2*)out);
        BRV_LAUNCH_CHECK("channel_mean_mask_tiled_kernel");
        return BRV_OK;
    }
    dim3 grid((unsigned)brv_ceil_div(n_frames * n_bins, 256 * CM_ITEMS), (unsigned)n_batch);
    channel_mean_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float2*)X, xb, xc, xf, xt, mask, mb, mf, mt, n_channels, n_bins, n_frames, (float2*)out);
    BRV_LAUNCH_CHECK("channel_mean_mask_kernel");
    return BRV_OK;
}

extern "C" int brv_accumulate_mean(const float* v, int64_t n, float* total, void* stream) {
    BRV_REQUIRE(n >= 0, "bad shape");
    if (n == 0) return BRV_OK;
    BRV_REQUIRE(v && total, "null pointer argument");
    accumulate_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(v, n, total);
    BRV_LAUNCH_CHECK("accumulate_mean_kernel");
    return BRV_OK;
}
