// Float64 compute path of the STFT pair (brever/modules/stft.py:59-138 with float64 tensors:
// torch.stft / torch.istft then run cuFFT / pocketfft in double precision).
//
// The tensor-core kernels are fp32-grade (two fp16 planes, 2^-22); a float64 caller -- gradcheck, a
// numerical study -- gets the dtype AND the precision from this file instead: the same four
// operators as direct O(N) sums per output value in double arithmetic, with an exact twiddle
// table indexed by (k n) mod N.  Two kernels cover the four entry points:
//
//   analysis   out[s, t, k] = post * C( gain * w_k * sum_n win[n] m[j] in[s, j] e^{-2 pi i k n / N} ),
//              j = t hop + n - left, m = 1 or 1 / envelope, C = |.|^c e^{i arg(.)} or identity
//                  STFT.forward            : gain 1 / norm, w 1, m 1, c, post = scale_factor
//                  d STFT.backward / dX    : gain norm / (N scale), w = (1, 2, ..., 2, 1), m = 1 / envelope
//   synthesis  y[s, i] = gain * m[i] * sum_t win[n] sum_k w_k Re( D(pre X[s, k, t]) e^{+2 pi i k n / N} ),
//              n = i + left - t hop, D = |.|^(1/c) e^{i arg(.)} or identity
//                  STFT.backward           : pre 1 / scale, D, w = (1, 2, ..., 2, 1), gain norm / N, m = 1 / envelope
//                  d STFT.forward / dx     : pre 1, w 1, gain scale / norm, m 1
//
// (Im X[0] and Im X[N/2] drop out of the synthesis by themselves: sin(0) = sin(pi n) = 0.)
// Throughput is that of the fp64 pipe on a dense O(N) sum -- milliseconds where the tensor-core
// path takes tens of microseconds; this is a fidelity path, not a fast one.
#include "brv_common.cuh"

#include <math.h>

namespace {

constexpr int F64_THREADS = 256;

// sum over the frames t in [0, n_frames) that cover position pos (= sample index + left) of win[pos - t hop]^2
__device__ double ola_envelope_f64(const double* __restrict__ win, int N, int H, int64_t n_frames, int64_t pos) {
    int64_t t_hi = pos / H;
    if (t_hi > n_frames - 1) t_hi = n_frames - 1;
    double e = 0.0;
    for (int64_t t = t_hi; t >= 0; --t) {
        const int64_t n = pos - t * H;
        if (n >= N) break;
        e += win[n] * win[n];
    }
    return e;
}

template <typename T> struct Cplx;
template <> struct Cplx<double> { typedef double2 type; };
template <> struct Cplx<float> { typedef float2 type; };

// T = double: the float64 entry points; T = float: the generic ConvSTFT sizes (float tensors,
// double arithmetic)
template <typename T>
__global__ void __launch_bounds__(F64_THREADS)
analysis_f64_kernel(const T* __restrict__ in, int64_t in_stride, int64_t in_len, int left, int H, int N,
                    int F, int64_t n_frames, const double* __restrict__ win, const double2* __restrict__ tw,
                    double gain, double w_dc, double w_ny, double w_mid, double compression, double post,
                    int use_env, typename Cplx<T>::type* __restrict__ out) {
    extern __shared__ double fr[];                         // windowed frame, N doubles
    const int64_t t = blockIdx.x, s = blockIdx.y;
    const T* x = in + s * in_stride;
    for (int n = threadIdx.x; n < N; n += F64_THREADS) {
        const int64_t j = t * H + n - left;
        double v = (j >= 0 && j < in_len) ? (double)x[j] : 0.0;
        if (use_env && v != 0.0) v /= ola_envelope_f64(win, N, H, n_frames, j + left);
        fr[n] = v * win[n];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < F; k += F64_THREADS) {
        double re = 0.0, im = 0.0;
        int idx = 0;                                       // (k n) mod N
        for (int n = 0; n < N; ++n) {
            const double2 c = tw[idx];
            re = fma(fr[n], c.x, re);
            im = fma(-fr[n], c.y, im);
            idx += k;
            if (idx >= N) idx -= N;
        }
        const double w = gain * (k == 0 ? w_dc : (2 * k == N ? w_ny : w_mid));
        re *= w;
        im *= w;
        if (compression != 1.0) {
            const double mag = hypot(re, im);
            const double g = mag > 0.0 ? pow(mag, compression - 1.0) : 0.0;
            re *= g;
            im *= g;
        }
        typename Cplx<T>::type o;
        o.x = (T)(re * post);
        o.y = (T)(im * post);
        out[(s * n_frames + t) * F + k] = o;
    }
}

template <typename T>
__global__ void __launch_bounds__(F64_THREADS)
synthesis_f64_kernel(const typename Cplx<T>::type* __restrict__ X, int64_t ss, int64_t sb, int64_t sf, int F,
                     int64_t n_frames, int left, int H, int N, const double* __restrict__ win,
                     const double2* __restrict__ tw, double pre, double decompression, double w_dc, double w_ny,
                     double w_mid, double gain, int use_env, int64_t out_len, T* __restrict__ y) {
    extern __shared__ double2 row[];                       // one frame's weighted bins, F values
    const int64_t s = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * F64_THREADS;
    const int64_t i = i0 + threadIdx.x;
    const int64_t pos = i + left;
    // frames that cover any position of this block
    int64_t t_lo = (i0 + left - N + 1 + H - 1) / H;        // ceil((first pos - N + 1) / H), may be negative
    if (i0 + left - N + 1 < 0) t_lo = 0;
    int64_t t_hi = (i0 + left + F64_THREADS - 1) / H;
    if (t_hi > n_frames - 1) t_hi = n_frames - 1;
    double acc = 0.0;
    for (int64_t t = t_lo; t <= t_hi; ++t) {
        __syncthreads();
        for (int k = threadIdx.x; k < F; k += F64_THREADS) {
            const typename Cplx<T>::type xv = X[s * ss + (int64_t)k * sb + t * sf];
            double2 v = make_double2((double)xv.x * pre, (double)xv.y * pre);
            if (decompression != 1.0) {
                const double mag = hypot(v.x, v.y);
                const double g = mag > 0.0 ? pow(mag, decompression - 1.0) : 0.0;
                v.x *= g;
                v.y *= g;
            }
            const double w = k == 0 ? w_dc : (2 * k == N ? w_ny : w_mid);
            row[k] = make_double2(v.x * w, v.y * w);
        }
        __syncthreads();
        const int64_t n = pos - t * H;
        if (i < out_len && n >= 0 && n < N) {
            double f = 0.0;
            int idx = 0;                                   // (k n) mod N
            for (int k = 0; k < F; ++k) {
                const double2 c = tw[idx];
                f = fma(row[k].x, c.x, f);
                f = fma(-row[k].y, c.y, f);
                idx += (int)n;
                if (idx >= N) idx -= N;
            }
            acc = fma(f, win[n], acc);
        }
    }
    if (i < out_len) {
        double v = acc * gain;
        if (use_env) v /= ola_envelope_f64(win, N, H, n_frames, pos);
        y[s * out_len + i] = (T)v;
    }
}

// double window and twiddle table of a plan, built on first use
int f64_tables(const brv_stft_plan* cp, const double** win, const double2** tw) {
    brv_stft_plan* p = const_cast<brv_stft_plan*>(cp);
    std::lock_guard<std::mutex> lock(p->mu);
    if (!p->win64) {
        const int N = p->n_fft;
        std::vector<double> t(2 * (size_t)N);
        const double pi = 3.14159265358979323846264338327950288;
        for (int j = 0; j < N; ++j) {
            // exact octant symmetries are not needed at 1e-16: cos / sin of a reduced argument
            t[2 * j] = cos(2.0 * pi * j / N);
            t[2 * j + 1] = sin(2.0 * pi * j / N);
        }
        double *dw = nullptr, *dt = nullptr;
        BRV_CUDA(cudaMalloc((void**)&dw, N * sizeof(double)));
        if (cudaMalloc((void**)&dt, 2 * (size_t)N * sizeof(double)) != cudaSuccess) {
            cudaFree(dw);
            return brv_fail_cuda(cudaGetLastError(), "cudaMalloc(float64 twiddle table)");
        }
        if (cudaMemcpy(dw, p->window.data(), N * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(dt, t.data(), 2 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(dw);
            cudaFree(dt);
            return brv_fail_cuda(cudaGetLastError(), "cudaMemcpy(float64 tables; first float64 call of a plan "
                                                     "must not be inside a stream capture)");
        }
        p->win64 = dw;
        p->tw64 = dt;
    }
    *win = p->win64;
    *tw = reinterpret_cast<const double2*>(p->tw64);
    return BRV_OK;
}

template <typename T>
int launch_analysis(const brv_stft_plan* p, const T* in, int64_t n_sig, int64_t in_stride, int64_t in_len, int left,
                    int F, int64_t n_frames, double gain, double w_dc, double w_ny, double w_mid,
                    double compression, double post, int use_env, void* out, cudaStream_t st) {
    const double* win;
    const double2* tw;
    int rc = f64_tables(p, &win, &tw);
    if (rc != BRV_OK) return rc;
    BRV_REQUIRE(n_frames < (1LL << 31) && n_sig < 65536, "direct-sum path: too many frames / signals per call");
    dim3 grid((unsigned)n_frames, (unsigned)n_sig);
    if (p->n_fft * sizeof(double) > 48 * 1024)
        BRV_CUDA(cudaFuncSetAttribute(analysis_f64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    analysis_f64_kernel<T><<<grid, F64_THREADS, p->n_fft * sizeof(double), st>>>(
        in, in_stride, in_len, left, p->hop, p->n_fft, F, n_frames, win, tw, gain, w_dc, w_ny, w_mid,
        compression, post, use_env, (typename Cplx<T>::type*)out);
    BRV_LAUNCH_CHECK("analysis_f64_kernel");
    return BRV_OK;
}

template <typename T>
int launch_synthesis(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb, int64_t sf, int64_t n_sig,
                     int left, int F, int64_t n_frames, double pre, double decompression, double w_dc, double w_ny,
                     double w_mid, double gain, int use_env, int64_t out_len, T* y, cudaStream_t st) {
    const double* win;
    const double2* tw;
    int rc = f64_tables(p, &win, &tw);
    if (rc != BRV_OK) return rc;
    const int64_t blocks = brv_ceil_div(out_len, F64_THREADS);
    BRV_REQUIRE(blocks < (1LL << 31) && n_sig < 65536, "direct-sum path: too many samples / signals per call");
    dim3 grid((unsigned)blocks, (unsigned)n_sig);
    if (F * sizeof(double2) > 48 * 1024)
        BRV_CUDA(cudaFuncSetAttribute(synthesis_f64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    synthesis_f64_kernel<T><<<grid, F64_THREADS, F * sizeof(double2), st>>>(
        (const typename Cplx<T>::type*)X, ss, sb, sf, F, n_frames, left, p->hop, p->n_fft, win, tw, pre,
        decompression, w_dc, w_ny, w_mid, gain, use_env, out_len, y);
    BRV_LAUNCH_CHECK("synthesis_f64_kernel");
    return BRV_OK;
}

}  // namespace

// ConvSTFT (brever/modules/stft.py:201-319) at sizes the folded tensor-core kernels do not cover:
// float tensors, the same two direct-sum kernels.  Analysis: frames start L - H before t hop, DC
// row 1 / sqrt(2), `gain` = 1 / normalisation; synthesis = its adjoint (no envelope), cut by L - H.
int brv_direct_conv_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                            int64_t x_stride, double gain, void* out, int64_t n_frames, cudaStream_t st) {
    return launch_analysis<float>(p, x, n_sig, x_stride, samples, p->frame_length - p->hop, p->n_bins, n_frames,
                                  gain, sqrt(0.5), 1.0, 1.0, p->compression, p->scale, 0, out, st);
}
int brv_direct_conv_backward(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb, int64_t sf,
                             int64_t n_sig, int64_t n_frames, int64_t out_len, double gain, float* y,
                             cudaStream_t st) {
    return launch_synthesis<float>(p, X, ss, sb, sf, n_sig, p->frame_length - p->hop, p->n_bins, n_frames,
                                   1.0 / p->scale, 1.0 / p->compression, sqrt(0.5), 1.0, 1.0, gain, 0, out_len, y,
                                   st);
}

void brv_f64_plan_free(brv_stft_plan* p) {
    cudaFree(p->win64);
    cudaFree(p->tw64);
    p->win64 = p->tw64 = nullptr;
}

extern "C" int brv_stft_forward_f64(const brv_stft_plan* p, const double* x, int64_t n_signals, int64_t samples,
                                    int64_t x_stride, void* out, void* stream) {
    BRV_REQUIRE(p && out, "null pointer argument");
    BRV_REQUIRE(n_signals >= 0 && samples >= 0, "negative shape");
    int64_t n_frames = 0;
    int rc = brv_stft_geometry(p, samples, &n_frames, nullptr, nullptr);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0) return BRV_OK;
    BRV_REQUIRE(x, "input pointer is null");
    return launch_analysis<double>(p, x, n_signals, x_stride, samples, brv_left(p), p->n_bins, n_frames,
                                   1.0 / p->norm, 1.0, 1.0, 1.0, p->compression, p->scale, 0, out,
                                   (cudaStream_t)stream);
}

extern "C" int brv_istft_forward_f64(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb, int64_t sf,
                                     int64_t n_signals, int64_t n_frames, double* y, void* stream) {
    BRV_REQUIRE(p && X, "null pointer argument");
    BRV_REQUIRE(n_signals >= 0 && n_frames >= 1, "bad shape");
    if (!p->center)
        return brv_fail(BRV_ERR_UNSUPPORTED, "the inverse transform is implemented for center=True plans only");
    int64_t out_len = 0;
    int rc = brv_istft_geometry(p, n_frames, &out_len);
    if (rc != BRV_OK) return rc;
    rc = brv_check_nola(p, n_frames);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0 || out_len == 0) return BRV_OK;
    BRV_REQUIRE(y, "output pointer is null");
    return launch_synthesis<double>(p, X, ss, sb, sf, n_signals, brv_left(p), p->n_bins_inv, n_frames,
                                    1.0 / p->scale, 1.0 / p->compression, 1.0, 1.0, 2.0, p->norm / p->n_fft, 1,
                                    out_len, y, (cudaStream_t)stream);
}

extern "C" int brv_stft_forward_grad_f64(const brv_stft_plan* p, const void* gX, int64_t ss, int64_t sb,
                                         int64_t sf, int64_t n_signals, int64_t samples, double* gx,
                                         void* stream) {
    BRV_REQUIRE(p && gX && gx, "null pointer argument");
    if (p->compression != 1.0)
        return brv_fail(BRV_ERR_UNSUPPORTED, "gradient of the compressed STFT (compression_factor != 1) is not "
                                             "implemented: no reference model back-propagates through it");
    int64_t n_frames = 0;
    int rc = brv_stft_geometry(p, samples, &n_frames, nullptr, nullptr);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0 || samples == 0) return BRV_OK;
    return launch_synthesis<double>(p, gX, ss, sb, sf, n_signals, brv_left(p), p->n_bins, n_frames, 1.0, 1.0, 1.0,
                                    1.0, 1.0, p->scale / p->norm, 0, samples, gx, (cudaStream_t)stream);
}

extern "C" int brv_istft_forward_grad_f64(const brv_stft_plan* p, const double* gy, int64_t n_signals,
                                          int64_t n_frames, void* gX, void* stream) {
    BRV_REQUIRE(p && gy && gX, "null pointer argument");
    if (!p->center)
        return brv_fail(BRV_ERR_UNSUPPORTED, "the inverse transform is implemented for center=True plans only");
    if (p->compression != 1.0)
        return brv_fail(BRV_ERR_UNSUPPORTED, "gradient of the decompressing iSTFT (compression_factor != 1) is "
                                             "not implemented: the reference only runs it under no_grad");
    BRV_REQUIRE(p->n_bins == p->n_bins_inv, "float64 iSTFT gradient: one-sided plans only");
    int64_t out_len = 0;
    int rc = brv_istft_geometry(p, n_frames, &out_len);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0) return BRV_OK;
    return launch_analysis<double>(p, gy, n_signals, out_len, out_len, brv_left(p), p->n_bins_inv, n_frames,
                                   p->norm / ((double)p->n_fft * p->scale), 1.0, 1.0, 2.0, 1.0, 1.0, 1, gX,
                                   (cudaStream_t)stream);
}
