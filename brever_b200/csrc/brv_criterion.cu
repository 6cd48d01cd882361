// Fused SNR / SI-SNR reduction and its masked-affine gradient.
//
// Reference: brever/criterion.py:21-72 (sisnr), :75-101 (snr), :229-234
// (apply_mask).  The reference builds a float32 mask in a Python loop, multiplies
// both tensors by it (twice for sisnr), materialises (B,S,S,L) temporaries and
// launches ~25 kernels.  Here one pass reads each (estimate, target) row pair
// once, masks by comparing the sample index with lengths[b], and accumulates
// six moments in float64:
//     sum x, sum y, sum xy, sum x^2, sum y^2, sum (y-x)^2   over n < lengths[b]
// from which both criteria follow in closed form.  float64
// keeps the SI-SNR closed form ||e||^2 = ||a||^2 - <a,b>^2/||b||^2 exact to
// ~1e-15 relative, so it does not cancel at high SNR.
//
// HBM-bound: 8 algorithmic bytes per (estimate, target) sample pair.
#include <stdlib.h>

#include "brv_common.cuh"

namespace {

constexpr int CR_THREADS = 256;
constexpr int CR_UNROLL = 4;                            // 16-byte loads in flight per array per thread
constexpr int CR_BLOCK = CR_THREADS * 4 * CR_UNROLL;    // samples per CTA per iteration (4096)
constexpr int CR_MOMENTS = 6;

struct Moments {
    double sx, sy, sxy, sxx, syy, sdd;
    __device__ void zero() { sx = sy = sxy = sxx = syy = sdd = 0.0; }
    template <bool PAIRWISE>
    __device__ __forceinline__ void add(float xf, float yf) {
        const double x = xf, y = yf;
        syy = fma(y, y, syy);
        if (PAIRWISE) {                 // SI-SNR: the five centred-moment sums
            sx += x;
            sy += y;
            sxy = fma(x, y, sxy);
            sxx = fma(x, x, sxx);
        } else {                        // SNR: sum y^2 and sum (y - x)^2 (difference in float32,
            const double d = (double)(yf - xf);   // as the reference forms it)
            sdd = fma(d, d, sdd);
        }
    }
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}

// Loads the 4 samples i .. i+3 of a row masked to `length` samples (exact zeros beyond it).
__device__ __forceinline__ float4 load4(const float* row, int64_t i, int64_t length, bool vec) {
    if (vec && i + 4 <= length) return ld_stream4(row + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < length) v.x = __ldcs(row + i);
    if (i + 1 < length) v.y = __ldcs(row + i + 1);
    if (i + 2 < length) v.z = __ldcs(row + i + 2);
    if (i + 3 < length) v.w = __ldcs(row + i + 3);
    return v;
}

// workspace layout: partial[n_pairs][chunks][6] doubles | ticket[n_pairs] uint32
//
// One CTA = one 4096-sample chunk of one (estimate, target) row pair.  Every thread
// first issues all of its loads (4 x 16 B per array, 32 KB in flight per CTA, several
// CTAs per SM: enough bytes in flight to cover HBM latency), then masks by index and
// accumulates in float64.  Samples at or beyond lengths[b] contribute exact zeros.
template <bool PAIRWISE, int ITERS>
__global__ void __launch_bounds__(CR_THREADS)
snr_moments_kernel(const float* __restrict__ x, const float* __restrict__ y,
                   const int64_t* __restrict__ lengths, int64_t n_rows, int64_t length,
                   int64_t xsb, int64_t xsr, int64_t ysb, int64_t ysr,
                   float eps, float out_sign, int chunks, float* __restrict__ out_db,
                   double* __restrict__ moments, double* __restrict__ partial,
                   unsigned int* __restrict__ ticket) {
    const int64_t pair = blockIdx.x;
    const int chunk = blockIdx.y;
    int64_t b, xr, yr;
    if (PAIRWISE) {                      // pair = (b, target i, estimate j)
        b = pair / (n_rows * n_rows);
        int64_t ij = pair % (n_rows * n_rows);
        yr = ij / n_rows;
        xr = ij % n_rows;
    } else {
        b = pair / n_rows;
        xr = yr = pair % n_rows;
    }
    int64_t valid = lengths[b];
    if (valid > length) valid = length;
    if (valid < 0) valid = 0;
    const float* xp = x + b * xsb + xr * xsr;
    const float* yp = y + b * ysb + yr * ysr;

    Moments m;
    m.zero();
    const bool vx = (((uintptr_t)xp) & 15) == 0, vy = (((uintptr_t)yp) & 15) == 0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int64_t begin = ((int64_t)chunk * ITERS + it) * CR_BLOCK;
        if (begin >= valid) break;
        float4 a[CR_UNROLL], c[CR_UNROLL];
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            const int64_t i = begin + 4 * (int64_t)(u * CR_THREADS + threadIdx.x);
            a[u] = load4(xp, i, valid, vx);
            c[u] = load4(yp, i, valid, vy);
        }
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            m.add<PAIRWISE>(a[u].x, c[u].x);
            m.add<PAIRWISE>(a[u].y, c[u].y);
            m.add<PAIRWISE>(a[u].z, c[u].z);
            m.add<PAIRWISE>(a[u].w, c[u].w);
        }
    }

    __shared__ double red[CR_THREADS / 32][CR_MOMENTS];
    __shared__ bool is_last;
    double vals[CR_MOMENTS] = {m.sx, m.sy, m.sxy, m.sxx, m.syy, m.sdd};
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
    for (int q = 0; q < CR_MOMENTS; ++q) {
        double v = warp_sum_d(vals[q]);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < CR_MOMENTS) {
        double v = 0;
        for (int w = 0; w < CR_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partial[(pair * chunks + chunk) * CR_MOMENTS + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(ticket + pair, 1u);
        is_last = (prev == (unsigned int)chunks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    // last CTA of this pair: fixed-order sum of the partials (deterministic)
    __threadfence();
    if (threadIdx.x < CR_MOMENTS) {
        double v = 0;
        for (int c = 0; c < chunks; ++c)
            v += __ldcg(partial + (pair * chunks + c) * CR_MOMENTS + threadIdx.x);
        red[0][threadIdx.x] = v;
        moments[pair * CR_MOMENTS + threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ticket[pair] = 0;                // leave the workspace zeroed for the next call
        const double e = (double)eps;
        double ratio;
        if (PAIRWISE) {
            // zero-mean over the valid length (criterion.py:48-49), closed form
            const double L = (double)lengths[b];   // the reference divides by lengths[b]
            const double sx = red[0][0], sy = red[0][1], sxy = red[0][2];
            const double sxx = red[0][3], syy = red[0][4];
            const double mx = sx / L, my = sy / L;
            const double n = (double)valid;
            // sums over the valid samples of (x-mx)(y-my) etc.
            const double dot = sxy - mx * sy - my * sx + n * mx * my;
            const double ea = sxx - 2 * mx * sx + n * mx * mx;
            const double eb = syy - 2 * my * sy + n * my * my;
            const double tgt = dot * dot / eb;             // ||s_target||^2
            double noise = ea - tgt;                       // ||e_noise||^2
            if (noise < 0) noise = 0;
            ratio = tgt / (noise + e);
        } else {
            ratio = red[0][4] / (red[0][5] + e);           // criterion.py:99
        }
        out_db[pair] = out_sign * (float)(10.0 * log10(ratio + e));   // criterion.py:61,100
    }
}

// ---- L1 reductions for MultiResYuLoss (criterion.py:135-226) and the plain masked MSE ----
// One scalar per row: partial sums per chunk, fixed-order final sum by the last CTA of the row
// (same ticket scheme as above, so results are deterministic).
__device__ __forceinline__ void finish_row_sum(double v, int64_t row, int chunk, int chunks,
                                               double* __restrict__ partial,
                                               unsigned int* __restrict__ ticket, double post,
                                               float* __restrict__ out) {
    __shared__ double red1[CR_THREADS / 32];
    __shared__ bool last1;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    v = warp_sum_d(v);
    if (lane == 0) red1[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < CR_THREADS / 32; ++w) t += red1[w];
        partial[row * chunks + chunk] = t;
        __threadfence();
        const unsigned int prev = atomicAdd(ticket + row, 1u);
        last1 = (prev == (unsigned int)chunks - 1);
    }
    __syncthreads();
    if (!last1 || threadIdx.x != 0) return;
    __threadfence();
    double t = 0;
    for (int c = 0; c < chunks; ++c) t += __ldcg(partial + row * chunks + c);
    ticket[row] = 0;
    out[row] = (float)(t * post);
}

// out[row] = sum_{n < lengths[b]} | s[row] * x[n] - y[n] |      (time-domain term, :207-209)
__global__ void __launch_bounds__(CR_THREADS)
l1_rows_kernel(const float* __restrict__ x, const float* __restrict__ y,
               const int64_t* __restrict__ lengths, const float* __restrict__ scale,
               int64_t n_rows, int64_t length, int64_t xsb, int64_t xsr, int64_t ysb, int64_t ysr,
               int chunks, float* __restrict__ out, double* __restrict__ partial,
               unsigned int* __restrict__ ticket) {
    const int64_t row = blockIdx.x;
    const int chunk = blockIdx.y;
    const int64_t b = row / n_rows, r = row % n_rows;
    int64_t valid = lengths[b];
    if (valid > length) valid = length;
    if (valid < 0) valid = 0;
    const float* xp = x + b * xsb + r * xsr;
    const float* yp = y + b * ysb + r * ysr;
    const float s = scale ? __ldg(scale + row) : 1.f;
    const bool vx = (((uintptr_t)xp) & 15) == 0, vy = (((uintptr_t)yp) & 15) == 0;
    double acc = 0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int64_t begin = ((int64_t)chunk * 2 + it) * CR_BLOCK;
        if (begin >= valid) break;
        float4 a[CR_UNROLL], c[CR_UNROLL];
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            const int64_t i = begin + 4 * (int64_t)(u * CR_THREADS + threadIdx.x);
            a[u] = load4(xp, i, valid, vx);
            c[u] = load4(yp, i, valid, vy);
        }
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            const float t = (fabsf(s * a[u].x - c[u].x) + fabsf(s * a[u].y - c[u].y)) +
                            (fabsf(s * a[u].z - c[u].z) + fabsf(s * a[u].w - c[u].w));
            acc += (double)t;
        }
    }
    finish_row_sum(acc, row, chunk, chunks, partial, ticket, 1.0, out);
}

constexpr int MAG_BLOCK = 8 * CR_THREADS;     // complex bins per CTA
// out[sig] = sum over the (contiguous) spectrogram of | |X| - |Y| |        (spectral term, :213-216)
__global__ void __launch_bounds__(CR_THREADS)
mag_l1_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int64_t n_elems,
              int chunks, float* __restrict__ out, double* __restrict__ partial,
              unsigned int* __restrict__ ticket) {
    const int64_t sig = blockIdx.x;
    const int chunk = blockIdx.y;
    const float2* xp = X + sig * n_elems;
    const float2* yp = Y + sig * n_elems;
    const int64_t begin = (int64_t)chunk * MAG_BLOCK;
    double acc = 0;
    float2 a[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = begin + u * CR_THREADS + threadIdx.x;
        a[u] = i < n_elems ? __ldcs(xp + i) : make_float2(0.f, 0.f);
        c[u] = i < n_elems ? __ldcs(yp + i) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
        const float t0 = fabsf(hypotf(a[u].x, a[u].y) - hypotf(c[u].x, c[u].y));
        const float t1 = fabsf(hypotf(a[u + 1].x, a[u + 1].y) - hypotf(c[u + 1].x, c[u + 1].y));
        acc += (double)(t0 + t1);
    }
    finish_row_sum(acc, sig, chunk, chunks, partial, ticket, 1.0, out);
}

// gX = coef[sig] * sign(|X| - |Y|) * X / |X|   (torch's complex gradient of the spectral term)
__global__ void mag_l1_grad_kernel(const float2* __restrict__ X, const float2* __restrict__ Y,
                                   const float* __restrict__ coef, int64_t n_elems,
                                   float2* __restrict__ gX) {
    const int64_t sig = blockIdx.x;
    const float cf = coef[sig];
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n_elems;
         i += (int64_t)gridDim.y * blockDim.x) {
        const float2 a = __ldcs(X + sig * n_elems + i), c = __ldcs(Y + sig * n_elems + i);
        const float ma = hypotf(a.x, a.y), mc = hypotf(c.x, c.y);
        float2 g = make_float2(0.f, 0.f);
        if (ma > 0.f && ma != mc) {
            const float k = (ma > mc ? cf : -cf) / ma;
            g = make_float2(k * a.x, k * a.y);
        }
        gX[sig * n_elems + i] = g;
    }
}

// MANNER's multi-resolution STFT loss (models/manner/stft_loss.py:22-77) on two spectrograms of the
// same resolution: with m(.) = sqrt(max(re^2 + im^2, 1e-7)) the per-signal sums
//   s0 = sum (m(Y) - m(X))^2   s1 = sum m(Y)^2   s2 = sum |log m(Y) - log m(X)|
// (spectral convergence = sqrt(s0 / s1), log-magnitude = s2 / n); double accumulation.
constexpr float MR_CLAMP = 1e-7f;
__global__ void __launch_bounds__(CR_THREADS)
mrstft_sums_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int64_t n_elems,
                   double* __restrict__ sums) {
    const int64_t sig = blockIdx.x;
    const float2* xp = X + sig * n_elems;
    const float2* yp = Y + sig * n_elems;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int64_t i0 = (int64_t)blockIdx.y * (4 * CR_THREADS); i0 < n_elems;
         i0 += (int64_t)gridDim.y * (4 * CR_THREADS)) {
        float2 a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * CR_THREADS + threadIdx.x;
            a[u] = i < n_elems ? __ldcs(xp + i) : make_float2(0.f, 0.f);
            c[u] = i < n_elems ? __ldcs(yp + i) : make_float2(0.f, 0.f);
        }
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * CR_THREADS + threadIdx.x < n_elems) {
                const float px = fmaxf(a[u].x * a[u].x + a[u].y * a[u].y, MR_CLAMP);
                const float py = fmaxf(c[u].x * c[u].x + c[u].y * c[u].y, MR_CLAMP);
                const float mx = sqrtf(px), my = sqrtf(py);
                t0 += (my - mx) * (my - mx);
                t1 += py;
                t2 += fabsf(0.5f * (logf(py) - logf(px)));
            }
        }
        a0 += t0;
        a1 += t1;
        a2 += t2;
    }
    __shared__ double red[3][CR_THREADS / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const int w = threadIdx.x / 32;
    if ((threadIdx.x & 31) == 0) {
        red[0][w] = a0;
        red[1][w] = a1;
        red[2][w] = a2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int k = 0; k < CR_THREADS / 32; ++k) t += red[threadIdx.x][k];
        atomicAdd(sums + sig * 3 + threadIdx.x, t);
    }
}

// gX = (k_sc (m(X) - m(Y)) + k_mag sign(m(X) - m(Y)) / m(X)) X / m(X); zero where the clamp is active
__global__ void mrstft_grad_kernel(const float2* __restrict__ X, const float2* __restrict__ Y,
                                   const float* __restrict__ k_sc, const float* __restrict__ k_mag,
                                   int64_t n_elems, float2* __restrict__ gX) {
    const int64_t sig = blockIdx.x;
    const float ks = k_sc[sig], km = k_mag[sig];
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n_elems;
         i += (int64_t)gridDim.y * blockDim.x) {
        const float2 a = __ldcs(X + sig * n_elems + i), c = __ldcs(Y + sig * n_elems + i);
        const float px = a.x * a.x + a.y * a.y;
        float2 g = make_float2(0.f, 0.f);
        if (px > MR_CLAMP) {
            const float mx = sqrtf(px), my = sqrtf(fmaxf(c.x * c.x + c.y * c.y, MR_CLAMP));
            const float d = mx - my;
            const float k = (ks * d + (d > 0.f ? km : (d < 0.f ? -km : 0.f)) / mx) / mx;
            g = make_float2(k * a.x, k * a.y);
        }
        gX[sig * n_elems + i] = g;
    }
}

// gx[row, n] = coef[row] * sign(s x - y) for n < lengths[b], else 0
__global__ void l1_rows_grad_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const int64_t* __restrict__ lengths,
                                    const float* __restrict__ scale, const float* __restrict__ coef,
                                    int64_t n_rows, int64_t length, int64_t xsb, int64_t xsr,
                                    int64_t ysb, int64_t ysr, float* __restrict__ gx) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / n_rows, r = row % n_rows;
    int64_t valid = lengths[b];
    if (valid > length) valid = length;
    const float s = scale ? scale[row] : 1.f, cf = coef[row];
    const float* xp = x + b * xsb + r * xsr;
    const float* yp = y + b * ysb + r * ysr;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < length;
         i += (int64_t)gridDim.y * blockDim.x) {
        float v = 0.f;
        if (i < valid) {
            const float d = s * __ldg(xp + i) - __ldg(yp + i);
            v = d > 0.f ? cf : (d < 0.f ? -cf : 0.f);
        }
        gx[row * length + i] = v;
    }
}

__global__ void masked_affine_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const int64_t* __restrict__ lengths, int64_t n_rows,
                                     int64_t length, int64_t xsb, int64_t xsr, int64_t ysb,
                                     int64_t ysr, const float* __restrict__ ca,
                                     const float* __restrict__ cb, const float* __restrict__ c0,
                                     const int32_t* __restrict__ ymap, float* __restrict__ gx) {
    const int64_t row = blockIdx.x;               // b * n_rows + r
    const int64_t b = row / n_rows, r = row % n_rows;
    const int64_t yr = ymap ? ymap[row] : r;
    int64_t valid = lengths[b];
    if (valid > length) valid = length;
    const float a = ca[row], bb = cb[row], c = c0[row];
    const float* xp = x + b * xsb + r * xsr;
    const float* yp = y + b * ysb + yr * ysr;
    float* gp = gx + row * length;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < length;
         i += (int64_t)gridDim.y * blockDim.x) {
        float v = 0.f;
        if (i < valid) v = fmaf(a, __ldg(xp + i), fmaf(bb, __ldg(yp + i), c));
        gp[i] = v;
    }
}

// Criterion gradient with the per-row coefficients evaluated in the kernel from the saved
// float64 moments (one launch instead of ~25 small tensor ops, SURVEY 8a' closed forms):
//   mode 0 (snr):   gx = g coef (x - y),  coef = K 2P / ((r + eps)(D + eps)^2), r = P / (D + eps)
//   mode 1 (sisnr): gx = g (beta (x - mx) + alpha (y - my)) for the matched (target, estimate) pair
// g = -gout[b] * gscale for sisnr (loss = -dB), +gout[b] * gscale * ... see the host wrapper.
__global__ void __launch_bounds__(256)
criterion_grad_kernel(const float* __restrict__ x, const float* __restrict__ y,
                      const int64_t* __restrict__ lengths, const double* __restrict__ moments,
                      const float* __restrict__ gout, int64_t gout_stride, float gscale,
                      const int32_t* __restrict__ ymap, int mode, int64_t n_rows, int64_t length,
                      int64_t xsb, int64_t xsr, int64_t ysb, int64_t ysr, float eps,
                      float* __restrict__ gx) {
    const int64_t row = blockIdx.x;               // b * n_rows + r (estimate row)
    const int64_t b = row / n_rows, r = row % n_rows;
    const int64_t yr = ymap ? ymap[row] : r;
    int64_t valid = lengths[b];
    if (valid > length) valid = length;
    if (valid < 0) valid = 0;
    __shared__ float coef[3];
    if (threadIdx.x == 0) {
        const double K = 4.342944819032518;       // 10 / ln 10
        const double e = (double)eps;
        const double g = (double)gout[b * gout_stride] * (double)gscale;
        if (mode == 0) {
            const double* m = moments + row * CR_MOMENTS;
            const double p = m[4], d = m[5];
            const double ratio = p / (d + e);
            const double c = g * K * 2.0 * p / ((ratio + e) * (d + e) * (d + e));
            coef[0] = (float)c;
            coef[1] = (float)(-c);
            coef[2] = 0.f;
        } else {
            const double* m = moments + ((b * n_rows + yr) * n_rows + r) * CR_MOMENTS;
            const double sx = m[0], sy = m[1], sxy = m[2], sxx = m[3], syy = m[4];
            const double L = (double)lengths[b], n = (double)valid;
            const double mx = sx / L, my = sy / L;
            const double dot = sxy - mx * sy - my * sx + n * mx * my;
            const double ea = sxx - 2 * mx * sx + n * mx * mx;
            const double eb = syy - 2 * my * sy + n * my * my;
            const double t = dot * dot / eb;
            double en = ea - t;
            if (en < 0) en = 0;
            const double ratio = t / (en + e);
            const double common = K / ((ratio + e) * (en + e) * (en + e));
            const double alpha = common * (en + e + t) * (2 * dot / eb);   // multiplies (y - my)
            const double beta = -2 * t * common;                            // multiplies (x - mx)
            coef[0] = (float)(-g * beta);
            coef[1] = (float)(-g * alpha);
            coef[2] = (float)(-g * (-alpha * my - beta * mx));
        }
    }
    __syncthreads();
    const float ca = coef[0], cb = coef[1], c0 = coef[2];
    const float* xp = x + b * xsb + r * xsr;
    const float* yp = y + b * ysb + yr * ysr;
    float* gp = gx + row * length;
    const bool vec = (((((uintptr_t)xp) | ((uintptr_t)yp) | ((uintptr_t)gp)) & 15) == 0);
    const int64_t base = (int64_t)blockIdx.y * 2048 * 4;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = base + 4 * (int64_t)(u * 256 + threadIdx.x);
        if (i >= length) break;
        if (vec && i + 4 <= length) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < valid) {
                const float4 a = load4(xp, i, valid, true), c = load4(yp, i, valid, true);
                o.x = fmaf(ca, a.x, fmaf(cb, c.x, c0));
                o.y = i + 1 < valid ? fmaf(ca, a.y, fmaf(cb, c.y, c0)) : 0.f;
                o.z = i + 2 < valid ? fmaf(ca, a.z, fmaf(cb, c.z, c0)) : 0.f;
                o.w = i + 3 < valid ? fmaf(ca, a.w, fmaf(cb, c.w, c0)) : 0.f;
            }
            *reinterpret_cast<float4*>(gp + i) = o;
        } else {
            for (int64_t j = i; j < i + 4 && j < length; ++j)
                gp[j] = j < valid ? fmaf(ca, __ldg(xp + j), fmaf(cb, __ldg(yp + j), c0)) : 0.f;
        }
    }
}

__global__ void apply_mask_kernel(const float* __restrict__ x, const int64_t* __restrict__ lengths,
                                  int64_t inner, int64_t length, float* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / inner;
    const int64_t valid = lengths[b];
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < length;
         i += (int64_t)gridDim.y * blockDim.x)
        out[row * length + i] = i < valid ? x[row * length + i] : 0.f;
}

}  // namespace

// samples per CTA = ITERS * 4096.  Measured on B200: 2 for problems of a few MB per array (64 x 4 s:
// 11.4 us against 13.4 / 12.4 with 1 / 4), 4 once the arrays reach tens of MB (256 x 4 s: 27.2 us
// against 28.8).  BRV_CR_ITERS (1, 2 or 4) overrides.
static int cr_iters(int64_t n_pairs, int64_t length) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("BRV_CR_ITERS");
        const int v = e ? atoi(e) : 0;
        forced = (v == 1 || v == 2 || v == 4) ? v : 0;
    }
    if (forced) return forced;
    return n_pairs * length >= 12LL * 1000 * 1000 ? 4 : 2;
}
static int chunks_for(int64_t n_pairs, int64_t length) {
    int64_t c = brv_ceil_div(length, (int64_t)CR_BLOCK * cr_iters(n_pairs, length));
    return (int)(c < 1 ? 1 : c);
}

extern "C" size_t brv_snr_workspace_bytes(int64_t n_pairs, int64_t length) {
    if (n_pairs <= 0) return 256;
    size_t tickets = ((size_t)n_pairs * sizeof(unsigned int) + 255) & ~(size_t)255;
    return tickets + (size_t)n_pairs * chunks_for(n_pairs, length) * CR_MOMENTS * sizeof(double);
}

extern "C" int brv_snr_forward(const float* x, const float* y, const int64_t* lengths,
                               int64_t n_batch, int64_t n_rows, int64_t length,
                               int64_t xsb, int64_t xsr, int64_t ysb, int64_t ysr,
                               int pairwise, float eps, float out_sign, float* out_db,
                               double* moments, void* workspace, size_t workspace_bytes,
                               void* stream) {
    BRV_REQUIRE(x && y && lengths && out_db && moments && workspace, "null pointer argument");
    BRV_REQUIRE(n_batch >= 0 && n_rows >= 1 && length >= 0, "bad criterion shape");
    const int64_t n_pairs = pairwise ? n_batch * n_rows * n_rows : n_batch * n_rows;
    if (n_pairs == 0) return BRV_OK;
    const int chunks = chunks_for(n_pairs, length);
    BRV_REQUIRE(n_pairs <= 2147483647LL && chunks < 65536, "criterion problem too large");
    BRV_REQUIRE(workspace_bytes >= brv_snr_workspace_bytes(n_pairs, length),
                "criterion workspace too small: %zu < %zu", workspace_bytes,
                brv_snr_workspace_bytes(n_pairs, length));
    // ticket counters first (they must start zeroed), partial sums after
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    size_t off = ((size_t)n_pairs * sizeof(unsigned int) + 255) & ~(size_t)255;
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + off);
    dim3 grid((unsigned)n_pairs, (unsigned)chunks);
#define BRV_LAUNCH_SNR(PW_, IT_)                                                              \
    snr_moments_kernel<PW_, IT_><<<grid, CR_THREADS, 0, (cudaStream_t)stream>>>(                \
        x, y, lengths, n_rows, length, xsb, xsr, ysb, ysr, eps, out_sign, chunks, out_db,       \
        moments, partial, ticket)
    switch (cr_iters(n_pairs, length) * 2 + (pairwise ? 1 : 0)) {
        case 2: BRV_LAUNCH_SNR(false, 1); break;
        case 3: BRV_LAUNCH_SNR(true, 1); break;
        case 4: BRV_LAUNCH_SNR(false, 2); break;
        case 5: BRV_LAUNCH_SNR(true, 2); break;
        case 8: BRV_LAUNCH_SNR(false, 4); break;
        default: BRV_LAUNCH_SNR(true, 4); break;
    }
#undef BRV_LAUNCH_SNR
    BRV_LAUNCH_CHECK("snr_moments_kernel");
    return BRV_OK;
}

extern "C" int brv_masked_affine(const float* x, const float* y, const int64_t* lengths,
                                 int64_t n_batch, int64_t n_rows, int64_t length, int64_t xsb,
                                 int64_t xsr, int64_t ysb, int64_t ysr, const float* ca,
                                 const float* cb, const float* c0, const int32_t* ymap,
                                 float* gx, void* stream) {
    BRV_REQUIRE(x && y && lengths && ca && cb && c0 && gx, "null pointer argument");
    const int64_t rows = n_batch * n_rows;
    if (rows == 0 || length == 0) return BRV_OK;
    BRV_REQUIRE(rows <= 2147483647LL, "too many rows per call");
    unsigned gx_blocks = (unsigned)brv_ceil_div(length, 256 * 8);
    if (gx_blocks < 1) gx_blocks = 1;
    if (gx_blocks > 65535) gx_blocks = 65535;                  // grid-stride loop inside
    masked_affine_kernel<<<dim3((unsigned)rows, gx_blocks), 256, 0, (cudaStream_t)stream>>>(
        x, y, lengths, n_rows, length, xsb, xsr, ysb, ysr, ca, cb, c0, ymap, gx);
    BRV_LAUNCH_CHECK("masked_affine_kernel");
    return BRV_OK;
}

extern "C" int brv_apply_mask(const float* x, const int64_t* lengths, int64_t n_batch,
                              int64_t inner, int64_t length, float* out, void* stream) {
    BRV_REQUIRE(x && lengths && out, "null pointer argument");
    const int64_t rows = n_batch * inner;
    if (rows == 0 || length == 0) return BRV_OK;
    BRV_REQUIRE(rows <= 2147483647LL, "too many rows per call");
    unsigned blocks = (unsigned)brv_ceil_div(length, 256 * 8);
    if (blocks > 65535) blocks = 65535;                        // grid-stride loop inside
    apply_mask_kernel<<<dim3((unsigned)rows, blocks), 256, 0, (cudaStream_t)stream>>>(
        x, lengths, inner, length, out);
    BRV_LAUNCH_CHECK("apply_mask_kernel");
    return BRV_OK;
}

// ---- L1 terms of MultiResYuLoss ---------------------------------------------------------
static int l1_chunks(int64_t n, int64_t per_cta) {
    int64_t c = brv_ceil_div(n, per_cta);
    return (int)(c < 1 ? 1 : c);
}

extern "C" size_t brv_l1_workspace_bytes(int64_t n_rows_total, int64_t length) {
    if (n_rows_total <= 0) return 256;
    size_t tickets = ((size_t)n_rows_total * sizeof(unsigned int) + 255) & ~(size_t)255;
    return tickets + (size_t)n_rows_total * l1_chunks(length, MAG_BLOCK) * sizeof(double);
}

static void l1_split_ws(void* workspace, int64_t rows, unsigned int** ticket, double** partial) {
    *ticket = reinterpret_cast<unsigned int*>(workspace);
    size_t off = ((size_t)rows * sizeof(unsigned int) + 255) & ~(size_t)255;
    *partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + off);
}

extern "C" int brv_l1_forward(const float* x, const float* y, const int64_t* lengths,
                              const float* scale, int64_t n_batch, int64_t n_rows, int64_t length,
                              int64_t xsb, int64_t xsr, int64_t ysb, int64_t ysr, float* out,
                              void* workspace, size_t workspace_bytes, void* stream) {
    BRV_REQUIRE(x && y && lengths && out && workspace, "null pointer argument");
    const int64_t rows = n_batch * n_rows;
    if (rows == 0) return BRV_OK;
    BRV_REQUIRE(workspace_bytes >= brv_l1_workspace_bytes(rows, length), "L1 workspace too small");
    const int chunks = l1_chunks(length, 2 * CR_BLOCK);
    BRV_REQUIRE(rows <= 2147483647LL && chunks < 65536, "L1 problem too large");
    unsigned int* ticket;
    double* partial;
    l1_split_ws(workspace, rows, &ticket, &partial);
    l1_rows_kernel<<<dim3((unsigned)rows, (unsigned)chunks), CR_THREADS, 0, (cudaStream_t)stream>>>(
        x, y, lengths, scale, n_rows, length, xsb, xsr, ysb, ysr, chunks, out, partial, ticket);
    BRV_LAUNCH_CHECK("l1_rows_kernel");
    return BRV_OK;
}

extern "C" int brv_l1_backward(const float* x, const float* y, const int64_t* lengths,
                               const float* scale, const float* coef, int64_t n_batch,
                               int64_t n_rows, int64_t length, int64_t xsb, int64_t xsr,
                               int64_t ysb, int64_t ysr, float* gx, void* stream) {
    BRV_REQUIRE(x && y && lengths && coef && gx, "null pointer argument");
    const int64_t rows = n_batch * n_rows;
    if (rows == 0 || length == 0) return BRV_OK;
    BRV_REQUIRE(rows <= 2147483647LL, "too many rows per call");
    unsigned blocks = (unsigned)brv_ceil_div(length, 256 * 8);
    if (blocks > 65535) blocks = 65535;                        // grid-stride loop inside
    l1_rows_grad_kernel<<<dim3((unsigned)rows, blocks), 256, 0, (cudaStream_t)stream>>>(
        x, y, lengths, scale, coef, n_rows, length, xsb, xsr, ysb, ysr, gx);
    BRV_LAUNCH_CHECK("l1_rows_grad_kernel");
    return BRV_OK;
}

extern "C" int brv_mag_l1_forward(const void* X, const void* Y, int64_t n_signals, int64_t n_elems,
                                  float* out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    BRV_REQUIRE(X && Y && out && workspace, "null pointer argument");
    if (n_signals == 0) return BRV_OK;
    BRV_REQUIRE(workspace_bytes >= brv_l1_workspace_bytes(n_signals, n_elems),
                "L1 workspace too small");
    const int chunks = l1_chunks(n_elems, MAG_BLOCK);
    BRV_REQUIRE(n_signals <= 2147483647LL && chunks < 65536, "L1 problem too large");
    unsigned int* ticket;
    double* partial;
    l1_split_ws(workspace, n_signals, &ticket, &partial);
    mag_l1_kernel<<<dim3((unsigned)n_signals, (unsigned)chunks), CR_THREADS, 0,
                    (cudaStream_t)stream>>>((const float2*)X, (const float2*)Y, n_elems, chunks,
                                            out, partial, ticket);
    BRV_LAUNCH_CHECK("mag_l1_kernel");
    return BRV_OK;
}

extern "C" int brv_mag_l1_backward(const void* X, const void* Y, const float* coef,
                                   int64_t n_signals, int64_t n_elems, void* gX, void* stream) {
    BRV_REQUIRE(X && Y && coef && gX, "null pointer argument");
    if (n_signals == 0 || n_elems == 0) return BRV_OK;
    BRV_REQUIRE(n_signals <= 2147483647LL, "too many signals per call");
    unsigned blocks = (unsigned)brv_ceil_div(n_elems, 256 * 8);
    if (blocks > 4096) blocks = 4096;
    mag_l1_grad_kernel<<<dim3((unsigned)n_signals, blocks), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)X, (const float2*)Y, coef, n_elems, (float2*)gX);
    BRV_LAUNCH_CHECK("mag_l1_grad_kernel");
    return BRV_OK;
}

extern "C" int brv_mrstft_forward(const void* X, const void* Y, int64_t n_signals, int64_t n_elems,
                                  double* sums, void* stream) {
    BRV_REQUIRE(n_signals >= 0 && n_elems >= 0, "bad shape");
    if (n_signals == 0) return BRV_OK;
    BRV_REQUIRE(sums && (n_elems == 0 || (X && Y)), "null pointer argument");
    BRV_REQUIRE(n_signals <= 2147483647LL, "too many signals per call");
    BRV_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_signals * 3 * sizeof(double), (cudaStream_t)stream));
    if (n_elems == 0) return BRV_OK;
    int64_t blocks = brv_ceil_div(n_elems, 4 * CR_THREADS);
    const int64_t cap = brv_ceil_div(148 * 8, n_signals);          // ~8 CTAs per SM in total
    if (blocks > cap) blocks = cap < 1 ? 1 : cap;
    mrstft_sums_kernel<<<dim3((unsigned)n_signals, (unsigned)blocks), CR_THREADS, 0, (cudaStream_t)stream>>>(
        (const float2*)X, (const float2*)Y, n_elems, sums);
    BRV_LAUNCH_CHECK("mrstft_sums_kernel");
    return BRV_OK;
}

extern "C" int brv_mrstft_backward(const void* X, const void* Y, const float* k_sc, const float* k_mag,
                                   int64_t n_signals, int64_t n_elems, void* gX, void* stream) {
    BRV_REQUIRE(n_signals >= 0 && n_elems >= 0, "bad shape");
    if (n_signals == 0 || n_elems == 0) return BRV_OK;
    BRV_REQUIRE(X && Y && k_sc && k_mag && gX, "null pointer argument");
    BRV_REQUIRE(n_signals <= 2147483647LL, "too many signals per call");
    unsigned blocks = (unsigned)brv_ceil_div(n_elems, 256 * 8);
    if (blocks > 4096) blocks = 4096;
    mrstft_grad_kernel<<<dim3((unsigned)n_signals, blocks), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)X, (const float2*)Y, k_sc, k_mag, n_elems, (float2*)gX);
    BRV_LAUNCH_CHECK("mrstft_grad_kernel");
    return BRV_OK;
}

extern "C" int brv_criterion_backward(const float* x, const float* y, const int64_t* lengths,
                                      const double* moments, const float* gout,
                                      int64_t gout_stride, float gscale, const int32_t* ymap,
                                      int pairwise, int64_t n_batch, int64_t n_rows,
                                      int64_t length, int64_t xsb, int64_t xsr, int64_t ysb,
                                      int64_t ysr, float eps, float* gx, void* stream) {
    BRV_REQUIRE(x && y && lengths && moments && gout && gx, "null pointer argument");
    const int64_t rows = n_batch * n_rows;
    if (rows == 0 || length == 0) return BRV_OK;
    BRV_REQUIRE(rows <= 2147483647LL, "too many rows per call");
    const unsigned blocks = (unsigned)brv_ceil_div(length, 2048 * 4);
    BRV_REQUIRE(blocks <= 65535, "rows longer than 2^29 samples are not supported");
    criterion_grad_kernel<<<dim3((unsigned)rows, blocks), 256, 0, (cudaStream_t)stream>>>(
        x, y, lengths, moments, gout, gout_stride, gscale, ymap, pairwise ? 1 : 0, n_rows, length,
        xsb, xsr, ysb, ysr, eps, gx);
    BRV_LAUNCH_CHECK("criterion_grad_kernel");
    return BRV_OK;
}
