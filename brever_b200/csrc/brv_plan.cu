// Library-level entry points, error reporting and the STFT plan (host side).
//
// The plan restates the constructor state of brever.modules.STFT
// (brever/modules/stft.py:32-54) and the integer frame arithmetic of
// STFT.pad / STFT.frame_count (stft.py:140-149).  The DFT bases are generated
// here in float64 with exact angle reduction (k*n mod N) and uploaded once.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>

#include "brv_common.cuh"

static thread_local char g_last_error[512] = "";

int brv_fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return status;
}

int brv_fail_cuda(cudaError_t err, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "CUDA error %d (%s) at %s",
             (int)err, cudaGetErrorString(err), where);
    return BRV_ERR_CUDA;
}

unsigned long long g_brv_launches = 0;

extern "C" int brv_abi_version(void) { return BRV_ABI_VERSION; }

extern "C" uint64_t brv_launch_count(void) {
    return __atomic_load_n(&g_brv_launches, __ATOMIC_RELAXED);
}

extern "C" const char* brv_last_error(void) { return g_last_error; }

extern "C" const char* brv_status_string(int status) {
    switch (status) {
        case BRV_OK: return "ok";
        case BRV_ERR_INVALID: return "invalid argument";
        case BRV_ERR_UNSUPPORTED: return "unsupported configuration";
        case BRV_ERR_CUDA: return "CUDA runtime error";
        case BRV_ERR_NOLA: return "window overlap add min: 1";
        case BRV_ERR_ALLOC: return "allocation failure";
        default: return "unknown status";
    }
}

extern "C" int brv_device_query(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    BRV_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    BRV_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return BRV_OK;
}

// cos/sin of 2*pi*m/N with exact values on the axes
static void unit_root(long long m, int N, double* c, double* s) {
    m %= N;
    if ((4 * m) % N == 0) {
        static const double cs[4] = {1, 0, -1, 0}, sn[4] = {0, 1, 0, -1};
        int q = (int)((4 * m) / N);
        *c = cs[q];
        *s = sn[q];
        return;
    }
    double ang = 2.0 * M_PI * (double)m / (double)N;
    *c = cos(ang);
    *s = sin(ang);
}

static int upload(float** dst, const std::vector<float>& src) {
    BRV_CUDA(cudaMalloc((void**)dst, src.size() * sizeof(float)));
    BRV_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(float),
                        cudaMemcpyHostToDevice));
    return BRV_OK;
}

int brv_tc_plan_init(brv_stft_plan* p, const std::vector<double>& fwd,
                     const std::vector<double>& inv);  // brv_stft_tc.cu
void brv_tc_plan_free(brv_stft_plan* p);
int brv_fold_plan_init(brv_stft_plan* p);                 // brv_stft_fold.cu
void brv_fold_plan_free(brv_stft_plan* p);

extern "C" int brv_stft_plan_create(brv_stft_plan** out, int frame_length,
                                    int hop_length, int n_fft,
                                    const double* window, int normalized,
                                    int onesided, double compression_factor,
                                    double scale_factor) {
    BRV_REQUIRE(out != nullptr, "plan output pointer is null");
    *out = nullptr;
    BRV_REQUIRE(frame_length >= 1 && hop_length >= 1, "frame_length and hop_length must be positive");
    if (n_fft <= 0) n_fft = frame_length;
    BRV_REQUIRE(n_fft >= frame_length, "n_fft (%d) must be >= frame_length (%d)", n_fft, frame_length);
    BRV_REQUIRE(n_fft <= 8192, "n_fft (%d) larger than 8192 is not supported", n_fft);
    BRV_REQUIRE(window != nullptr, "window is null");
    BRV_REQUIRE(compression_factor > 0 && scale_factor != 0, "compression_factor must be > 0 and scale_factor != 0");

    brv_stft_plan* p = new brv_stft_plan();
    p->frame_length = frame_length;
    p->hop = hop_length;
    p->n_fft = n_fft;
    p->onesided = onesided ? 1 : 0;
    p->normalized = normalized ? 1 : 0;
    p->center = 1;
    p->pad_frames = 1;
    p->n_bins = onesided ? n_fft / 2 + 1 : n_fft;
    p->n_bins_inv = n_fft / 2 + 1;
    p->compression = compression_factor;
    p->scale = scale_factor;
    p->basis_fwd = p->basis_fwd_t = p->basis_inv = p->basis_inv_t = p->window_sq = nullptr;
    p->tc_fwd = p->tc_inv = nullptr;
    p->fold = nullptr;
    p->win64 = p->tw64 = nullptr;
    p->tc_fwd_cols = p->tc_inv_k = 0;
    if (cudaGetDevice(&p->device) != cudaSuccess) {
        delete p;
        return brv_fail_cuda(cudaGetLastError(), "cudaGetDevice (no CUDA device: this library has no CPU path)");
    }

    const int N = n_fft, F = p->n_bins, Fi = p->n_bins_inv;
    // torch.stft centres a short window inside n_fft
    p->window.assign(N, 0.0);
    const int left = (N - frame_length) / 2;
    double sumsq = 0;
    for (int i = 0; i < frame_length; ++i) {
        p->window[left + i] = window[i];
        sumsq += window[i] * window[i];
    }
    p->norm = normalized ? sqrt(sumsq) : 1.0;
    BRV_REQUIRE(p->norm > 0, "window has zero energy");

    std::vector<double> fwd((size_t)N * 2 * F), inv((size_t)2 * Fi * N);
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < F; ++k) {
            double c, s;
            unit_root((long long)k * n, N, &c, &s);
            fwd[(size_t)n * 2 * F + 2 * k] = p->window[n] * c / p->norm;
            fwd[(size_t)n * 2 * F + 2 * k + 1] = -p->window[n] * s / p->norm;
        }
    }
    for (int k = 0; k < Fi; ++k) {
        const bool self_conj = (k == 0) || (N % 2 == 0 && k == N / 2);
        const double ck = (self_conj ? 1.0 : 2.0) / N;
        for (int n = 0; n < N; ++n) {
            double c, s;
            unit_root((long long)k * n, N, &c, &s);
            inv[(size_t)(2 * k) * N + n] = ck * c * p->window[n] * p->norm;
            // imaginary parts of DC / Nyquist are ignored by the c2r inverse
            inv[(size_t)(2 * k + 1) * N + n] = self_conj ? 0.0 : -ck * s * p->window[n] * p->norm;
        }
    }
    std::vector<float> h_fwd(fwd.size()), h_fwd_t(fwd.size()), h_inv(inv.size()), h_inv_t(inv.size()), h_wsq(N);
    for (int n = 0; n < N; ++n)
        for (int j = 0; j < 2 * F; ++j) {
            float v = (float)fwd[(size_t)n * 2 * F + j];
            h_fwd[(size_t)n * 2 * F + j] = v;
            h_fwd_t[(size_t)j * N + n] = v;
        }
    for (int j = 0; j < 2 * Fi; ++j)
        for (int n = 0; n < N; ++n) {
            float v = (float)inv[(size_t)j * N + n];
            h_inv[(size_t)j * N + n] = v;
            h_inv_t[(size_t)n * 2 * Fi + j] = v;
        }
    for (int n = 0; n < N; ++n) h_wsq[n] = (float)(p->window[n] * p->window[n]);

    int rc = upload(&p->basis_fwd, h_fwd);
    if (rc == BRV_OK) rc = upload(&p->basis_fwd_t, h_fwd_t);
    if (rc == BRV_OK) rc = upload(&p->basis_inv, h_inv);
    if (rc == BRV_OK) rc = upload(&p->basis_inv_t, h_inv_t);
    if (rc == BRV_OK) rc = upload(&p->window_sq, h_wsq);
    if (rc == BRV_OK) rc = brv_tc_plan_init(p, fwd, inv);
    if (rc == BRV_OK) rc = brv_fold_plan_init(p);
    if (rc != BRV_OK) {
        brv_stft_plan_destroy(p);
        return rc;
    }
    *out = p;
    return BRV_OK;
}

extern "C" int brv_stft_plan_destroy(brv_stft_plan* p) {
    if (!p) return BRV_OK;
    cudaFree(p->basis_fwd);
    cudaFree(p->basis_fwd_t);
    cudaFree(p->basis_inv);
    cudaFree(p->basis_inv_t);
    cudaFree(p->window_sq);
    brv_tc_plan_free(p);
    brv_fold_plan_free(p);
    brv_f64_plan_free(p);
    delete p;
    return BRV_OK;
}

extern "C" int brv_stft_geometry(const brv_stft_plan* p, int64_t samples,
                                 int64_t* n_frames, int64_t* n_bins,
                                 int64_t* pad_right) {
    BRV_REQUIRE(p != nullptr, "plan is null");
    BRV_REQUIRE(samples >= 0, "negative sample count");
    // stft.py:146-149: ceil(max(S - L, 0) / H) + 1 frames before centre padding
    int64_t over = samples > p->frame_length ? samples - p->frame_length : 0;
    int64_t frames0 = brv_ceil_div(over, p->hop) + 1;
    int64_t pad = p->pad_frames ? (frames0 - 1) * p->hop + p->frame_length - samples : 0;  // stft.py:143
    int64_t padded = samples + pad + 2 * (int64_t)brv_left(p);         // centre pad
    BRV_REQUIRE(padded >= p->n_fft, "input (%lld samples) shorter than one frame of %d",
                (long long)samples, p->n_fft);
    if (n_frames) *n_frames = 1 + (padded - p->n_fft) / p->hop;
    if (n_bins) *n_bins = p->n_bins;
    if (pad_right) *pad_right = pad;
    return BRV_OK;
}

extern "C" int brv_stft_plan_set_framing(brv_stft_plan* p, int center, int pad_to_frames) {
    BRV_REQUIRE(p != nullptr, "plan is null");
    p->center = center ? 1 : 0;
    p->pad_frames = pad_to_frames ? 1 : 0;
    return BRV_OK;
}

// out[s, j] = z[s, reflect(j - left)] where z is the signal right-padded to `padded` samples --
// with zeros (right_reflect = 0) or by mirroring its tail (right_reflect = 1: STFT.pad uses
// F.pad(mode=pad_mode), stft.py:140-144) -- and reflect() mirrors indices below 0 and beyond
// padded - 1 without repeating the edge sample (torch.stft pad_mode='reflect', stft.py:66-77).
__global__ void reflect_pad_kernel(const float* __restrict__ x, int64_t samples, int64_t x_stride,
                                   int64_t padded, int left, int right_reflect,
                                   float* __restrict__ out) {
    const int64_t total = padded + 2 * (int64_t)left;
    const int64_t sig = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = j - left;
        if (r < 0) r = -r;
        if (r >= padded) r = 2 * (padded - 1) - r;
        if (r >= samples && right_reflect) r = 2 * (samples - 1) - r;
        out[sig * total + j] = (r >= 0 && r < samples) ? __ldg(x + sig * x_stride + r) : 0.f;
    }
}
// adjoint, as a gather: x[r] feeds z[r] and (right_reflect) z[2 (S - 1) - r]; z[q] feeds the padded
// positions q + left, left - q (1 <= q <= left) and left + 2 (padded - 1) - q (right mirror)
__global__ void reflect_pad_grad_kernel(const float* __restrict__ g, int64_t samples, int64_t padded,
                                        int left, int right_reflect, float* __restrict__ gx) {
    const int64_t total = padded + 2 * (int64_t)left;
    const int64_t sig = blockIdx.y;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < samples;
         r += (int64_t)gridDim.x * blockDim.x) {
        const float* gs = g + sig * total;
        float acc = 0.f;
        for (int k = 0; k < 2; ++k) {
            const int64_t q = k == 0 ? r : 2 * (samples - 1) - r;
            if (k == 1 && !(right_reflect && q >= samples && q < padded)) continue;
            acc += gs[q + left];
            if (q >= 1 && q <= left) acc += gs[left - q];
            const int64_t m = 2 * (padded - 1) - q;
            if (m >= padded && m < padded + left) acc += gs[m + left];
        }
        gx[sig * samples + r] = acc;
    }
}

extern "C" int brv_reflect_pad(const float* x, int64_t n_signals, int64_t samples, int64_t x_stride,
                               int64_t padded, int left, int right_reflect, float* out, void* stream) {
    BRV_REQUIRE(n_signals >= 0 && samples >= 0 && padded >= samples && left >= 0, "bad reflect-pad shape");
    // torch raises for a mirror that is not shorter than the signal it mirrors
    BRV_REQUIRE(left < padded || left == 0, "reflect padding (%d) must be smaller than the signal (%lld)",
                left, (long long)padded);
    BRV_REQUIRE(!right_reflect || padded - samples < samples || padded == samples,
                "reflect padding (%lld) must be smaller than the signal (%lld)",
                (long long)(padded - samples), (long long)samples);
    if (n_signals == 0 || padded + 2 * (int64_t)left == 0) return BRV_OK;
    BRV_REQUIRE(out && (x || samples == 0), "null pointer argument");
    BRV_REQUIRE(n_signals < 65536, "more than 65535 signals per call");
    const int64_t total = padded + 2 * (int64_t)left;
    const unsigned blocks = (unsigned)(brv_ceil_div(total, 256) < 1024 ? brv_ceil_div(total, 256) : 1024);
    reflect_pad_kernel<<<dim3(blocks, (unsigned)n_signals), 256, 0, (cudaStream_t)stream>>>(
        x, samples, x_stride, padded, left, right_reflect, out);
    BRV_LAUNCH_CHECK("reflect_pad_kernel");
    return BRV_OK;
}

extern "C" int brv_reflect_pad_grad(const float* g, int64_t n_signals, int64_t samples, int64_t padded,
                                    int left, int right_reflect, float* gx, void* stream) {
    BRV_REQUIRE(n_signals >= 0 && samples >= 0 && padded >= samples && left >= 0, "bad reflect-pad shape");
    if (n_signals == 0 || samples == 0) return BRV_OK;
    BRV_REQUIRE(g && gx, "null pointer argument");
    BRV_REQUIRE(n_signals < 65536, "more than 65535 signals per call");
    const unsigned blocks = (unsigned)(brv_ceil_div(samples, 256) < 1024 ? brv_ceil_div(samples, 256) : 1024);
    reflect_pad_grad_kernel<<<dim3(blocks, (unsigned)n_signals), 256, 0, (cudaStream_t)stream>>>(
        g, samples, padded, left, right_reflect, gx);
    BRV_LAUNCH_CHECK("reflect_pad_grad_kernel");
    return BRV_OK;
}

extern "C" int brv_istft_geometry(const brv_stft_plan* p, int64_t n_frames,
                                  int64_t* samples) {
    BRV_REQUIRE(p != nullptr, "plan is null");
    BRV_REQUIRE(n_frames >= 1, "need at least one frame");
    if (samples) *samples = (int64_t)p->hop * (n_frames - 1) + p->n_fft - 2 * (int64_t)(p->n_fft / 2);
    return BRV_OK;
}

extern "C" size_t brv_stft_workspace_bytes(const brv_stft_plan* p, int64_t n_signals,
                                           int64_t n_frames) {
    if (!p || n_signals <= 0 || n_frames <= 0) return 0;
    // generic path: one fp32 time-domain frame per (signal, frame) + 1/envelope
    size_t frames = (size_t)n_signals * (size_t)n_frames * (size_t)p->n_fft * sizeof(float);
    size_t env = ((size_t)p->hop * (size_t)n_frames + (size_t)p->n_fft) * sizeof(float);
    return frames + env + 256;
}

// torch.istft refuses to divide by an envelope below 1e-11 (NOLA).
int brv_check_nola(const brv_stft_plan* p, int64_t n_frames) {
    brv_stft_plan* mp = const_cast<brv_stft_plan*>(p);
    std::lock_guard<std::mutex> lock(mp->mu);
    auto& cache = mp->nola_cache;
    auto it = cache.find(n_frames);
    if (it == cache.end()) {
        const int N = p->n_fft, H = p->hop;
        const int64_t full = N + (int64_t)H * (n_frames - 1);
        bool ok = true;
        // interior is periodic in H: check the first/last N positions and one period
        auto env_at = [&](int64_t pos) {
            double e = 0;
            int64_t t_hi = pos / H;
            if (t_hi > n_frames - 1) t_hi = n_frames - 1;
            int64_t t_lo = pos - N + 1 <= 0 ? 0 : (pos - N + 1 + H - 1) / H;
            for (int64_t t = t_lo; t <= t_hi; ++t) {
                double w = p->window[pos - t * H];
                e += w * w;
            }
            return e;
        };
        const int64_t lo = N / 2, hi = full - N / 2;  // [lo, hi) survives the trim
        for (int64_t pos = lo; pos < hi && ok; ++pos) {
            if (pos >= lo + 2 * (int64_t)N && pos < hi - 2 * (int64_t)N) {
                pos = hi - 2 * (int64_t)N - 1;  // skip the periodic interior
                continue;
            }
            if (fabs(env_at(pos)) < 1e-11) ok = false;
        }
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(n_frames, ok).first;
    }
    return it->second ? BRV_OK : brv_fail(BRV_ERR_NOLA, "window overlap add min: 1");
}
