// Shared declarations for the brever_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../include/brever_b200.h"

int brv_fail(int status, const char* fmt, ...);
int brv_fail_cuda(cudaError_t err, const char* where);

#define BRV_CUDA(call)                                         \
    do {                                                       \
        cudaError_t brv_e_ = (call);                           \
        if (brv_e_ != cudaSuccess) return brv_fail_cuda(brv_e_, #call); \
    } while (0)

extern unsigned long long g_brv_launches;  // kernels launched by this library

#define BRV_LAUNCH_CHECK(name)                                 \
    do {                                                       \
        cudaError_t brv_e_ = cudaGetLastError();               \
        if (brv_e_ != cudaSuccess) return brv_fail_cuda(brv_e_, name); \
        __atomic_add_fetch(&g_brv_launches, 1ULL, __ATOMIC_RELAXED); \
    } while (0)

#define BRV_REQUIRE(cond, ...)                                 \
    do {                                                       \
        if (!(cond)) return brv_fail(BRV_ERR_INVALID, __VA_ARGS__); \
    } while (0)

__host__ __device__ static inline int64_t brv_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Plan: immutable STFT parameters + device-resident DFT bases.
struct brv_stft_plan {
    int frame_length, hop, n_fft;
    int n_bins;       // bins produced by the forward transform (N/2+1 or N)
    int n_bins_inv;   // bins consumed by the inverse (always N/2+1)
    int normalized, onesided;
    int center;       // frames start n_fft/2 before t*hop (torch.stft center=True) or at t*hop
    int pad_frames;   // right-pad to whole frames first (STFT.pad, stft.py:140-144) or floor framing
    double compression, scale, norm;  // norm = sqrt(sum w^2) or 1
    int device;
    std::vector<double> window;       // zero-padded, centred, n_fft points
    // fp32 bases, row-major
    float* basis_fwd;    // (n_fft, 2*n_bins)      x-frame  -> spectrum
    float* basis_fwd_t;  // (2*n_bins, n_fft)      adjoint of the above
    float* basis_inv;    // (2*n_bins_inv, n_fft)  spectrum -> windowed frame
    float* basis_inv_t;  // (n_fft, 2*n_bins_inv)  adjoint of the above
    float* window_sq;    // (n_fft) squared window, for the OLA envelope
    // tensor-core operands (fp16 hi/lo splits, K-major), see brv_stft_tc.cu
    void* tc_fwd;        // forward basis^T  [2 parts][cols_pad][n_fft] half
    void* tc_inv;        // inverse basis^T  [2 parts][n_fft][k_pad] half
    int tc_fwd_cols, tc_inv_k;
    void* fold;          // symmetry-folded tensor-core plan, see brv_stft_fold.cu
    double* win64;       // float64 path (brv_stft_f64.cu): window and (cos, sin)(2 pi j / N), built on first use
    double* tw64;
    std::map<int64_t, bool> nola_cache;  // n_frames -> envelope is invertible
    std::mutex mu;
};

int brv_check_nola(const brv_stft_plan* p, int64_t n_frames);
void brv_f64_plan_free(brv_stft_plan* p);
static inline int brv_left(const brv_stft_plan* p) { return p->center ? p->n_fft / 2 : 0; }

// simt (generic) path, brv_stft_simt.cu
int brv_simt_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig,
                          int64_t samples, int64_t x_stride, float2* out,
                          int64_t n_frames, cudaStream_t st);
int brv_simt_spec_to_signal(const brv_stft_plan* p, const float2* X, int64_t ss,
                            int64_t sb, int64_t sf, int64_t n_sig, int64_t n_frames,
                            int64_t out_len, bool inverse, float* y, float* ws,
                            cudaStream_t st);
int brv_overlap_add(const brv_stft_plan* p, const float* frames, int64_t n_sig, int64_t n_frames,
                    int64_t out_len, bool inverse, float* y, float* inv_env, cudaStream_t st);
int brv_simt_istft_grad(const brv_stft_plan* p, const float* gy, int64_t n_sig,
                        int64_t n_frames, float2* gX, float* ws, cudaStream_t st);
