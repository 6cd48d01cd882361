// Generic (any n_fft / hop / window / one- or two-sided) STFT and iSTFT on the
// fp32 CUDA-core pipes: a register-tiled contraction of implicit frame tiles
// against the plan's DFT basis, with the reference's pre/post processing fused
// into the loaders and epilogues.  This path serves shapes the tensor-core
// kernels (brv_stft_tc.cu) do not cover and is the on-device cross-check for
// them; both sit behind the same C entry points.
//
// Reference semantics: brever/modules/stft.py:59-89 (forward) and :101-138
// (backward), torch.stft / torch.istft with center=True, pad_mode='constant'.
#include "brv_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

// ---- A-operand loaders: element (signal, frame t, k) of the implicit matrix --
struct FrameLoader {  // rows are overlapping frames of a zero-padded signal
    const float* x;
    int64_t x_stride, samples;
    int hop, left;               // left = n_fft / 2 (centre padding)
    const float* sample_scale;   // nullable per-sample multiplier (1/envelope)
    static constexpr bool kRowFast = false;
    __device__ float operator()(int64_t sig, int64_t t, int k) const {
        int64_t i = t * hop + k - left;
        if (i < 0 || i >= samples) return 0.f;   // both paddings are zeros
        float v = __ldg(x + sig * x_stride + i);
        if (sample_scale) v *= __ldg(sample_scale + i);
        return v;
    }
};

template <bool ROW_FAST>
struct SpecLoader {  // rows are frames of a complex spectrogram, k = 2*bin + (re|im)
    const float2* X;
    int64_t ss, sb, sf;          // element strides: signal, bin, frame
    float pre_scale;             // 1 / scale_factor
    float expo;                  // 1/c - 1 (decompression), 0 = none
    static constexpr bool kRowFast = ROW_FAST;
    __device__ float operator()(int64_t sig, int64_t t, int k) const {
        float2 v = __ldg(X + sig * ss + (int64_t)(k >> 1) * sb + t * sf);
        v.x *= pre_scale;
        v.y *= pre_scale;
        if (expo != 0.f) {
            float m2 = v.x * v.x + v.y * v.y;
            float g = m2 > 0.f ? powf(m2, 0.5f * expo) : 0.f;
            v.x *= g;
            v.y *= g;
        }
        return (k & 1) ? v.y : v.x;
    }
};

// ---- epilogues: receive 4 consecutive columns of one output row ---------------
struct SpecEpilogue {  // (signal, frame, 2F) interleaved complex, compress + scale
    float* out;
    int64_t n_frames;
    int n_cols;                  // 2 * (bins computed)
    int row_stride;              // 2 * (bins stored per frame)
    float expo;                  // c - 1 (0 = none)
    float post_scale;
    __device__ void operator()(int64_t sig, int64_t t, int col, const float* v) const {
        float* row = out + (sig * n_frames + t) * (int64_t)row_stride;
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
            if (col + q >= n_cols) break;
            float re = v[q], im = v[q + 1];
            if (expo != 0.f) {
                float m2 = re * re + im * im;
                float g = m2 > 0.f ? powf(m2, 0.5f * expo) : 0.f;
                re *= g;
                im *= g;
            }
            *reinterpret_cast<float2*>(row + col + q) = make_float2(re * post_scale, im * post_scale);
        }
    }
};

struct FrameEpilogue {  // (signal, frame, n_fft) time-domain frames into the workspace
    float* ws;
    int64_t n_frames;
    int n_cols;                  // n_fft
    __device__ void operator()(int64_t sig, int64_t t, int col, const float* v) const {
        float* row = ws + (sig * n_frames + t) * (int64_t)n_cols;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (col + q < n_cols) row[col + q] = v[q];
    }
};

// C tile = A tile (implicit) x basis, accumulated in float64: products of two
// float32 values are exact in float64, so the result carries only the final
// rounding to float32 (the reference's FFT has ~log2(N) roundings; a float32
// dot product over n_fft terms would have ~n_fft/sqrt(2)).
template <class ALoad, class Epi>
__global__ void __launch_bounds__(THREADS)
frame_gemm_kernel(ALoad aload, const float* __restrict__ B, int K, int n_cols,
                  int64_t n_frames, int tiles_per_signal, Epi epi) {
    __shared__ double As[BK][BM + 2];
    __shared__ __align__(16) double Bs[BK][BN];

    const int tid = threadIdx.x;
    const int64_t sig = blockIdx.x / tiles_per_signal;
    const int64_t t0 = (int64_t)(blockIdx.x % tiles_per_signal) * BM;
    const int c0 = blockIdx.y * BN;
    const int tx = tid % 16, ty = tid / 16;

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int r = 0; r < (BM * BK) / THREADS; ++r) {
            int e = tid + r * THREADS;
            int k, m;
            if (ALoad::kRowFast) { m = e % BM; k = e / BM; }
            else                 { k = e % BK; m = e / BK; }
            float v = 0.f;
            if (k0 + k < K && t0 + m < n_frames) v = aload(sig, t0 + m, k0 + k);
            As[k][m] = (double)v;
        }
#pragma unroll
        for (int r = 0; r < (BN * BK) / THREADS; ++r) {
            int e = tid + r * THREADS;
            int n = e % BN, k = e / BN;
            float v = 0.f;
            if (k0 + k < K && c0 + n < n_cols) v = __ldg(B + (int64_t)(k0 + k) * n_cols + c0 + n);
            Bs[k][n] = (double)v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t t = t0 + ty * 4 + i;
        int col = c0 + tx * 4;
        float v[4] = {(float)acc[i][0], (float)acc[i][1], (float)acc[i][2], (float)acc[i][3]};
        if (t < n_frames && col < n_cols) epi(sig, t, col, v);
    }
}

// 1 / (overlap-added squared window) on the trimmed output grid (torch.istft).
__global__ void inv_envelope_kernel(const float* __restrict__ wsq, int n_fft, int hop,
                                    int64_t n_frames, int64_t out_len, float* inv_env) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    int64_t pos = i + n_fft / 2;
    int64_t t_hi = pos / hop;
    if (t_hi > n_frames - 1) t_hi = n_frames - 1;
    int64_t t_lo = pos - n_fft + 1 <= 0 ? 0 : (pos - n_fft + hop) / hop;
    float e = 0.f;
    for (int64_t t = t_lo; t <= t_hi; ++t) e += wsq[pos - t * hop];
    inv_env[i] = 1.f / e;
}

// Overlap-add of the workspace frames at stride hop, centre trim, and either the
// window-sum-square division (inverse) or a plain scale (adjoint of the forward).
__global__ void overlap_add_kernel(const float* __restrict__ ws, int n_fft, int hop, int left,
                                   int64_t n_frames, int64_t out_len,
                                   const float* __restrict__ inv_env, float scale,
                                   float* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t sig = blockIdx.y;
    if (i >= out_len) return;
    int64_t pos = i + left;
    int64_t t_hi = pos / hop;
    if (t_hi > n_frames - 1) t_hi = n_frames - 1;
    int64_t t_lo = pos - n_fft + 1 <= 0 ? 0 : (pos - n_fft + hop) / hop;
    const float* base = ws + sig * n_frames * (int64_t)n_fft;
    float acc = 0.f;
    for (int64_t t = t_lo; t <= t_hi; ++t) acc += base[t * n_fft + (pos - t * hop)];
    y[sig * out_len + i] = acc * (inv_env ? inv_env[i] : scale);
}

template <class ALoad, class Epi>
int launch_gemm(ALoad a, const float* B, int K, int n_cols, int64_t n_sig,
                int64_t n_frames, Epi epi, cudaStream_t st) {
    if (n_sig == 0 || n_frames == 0) return BRV_OK;
    int64_t tiles = brv_ceil_div(n_frames, BM);
    int64_t gx = n_sig * tiles;
    BRV_REQUIRE(gx < (1LL << 31), "too many frame tiles (%lld)", (long long)gx);
    dim3 grid((unsigned)gx, (unsigned)brv_ceil_div(n_cols, BN));
    frame_gemm_kernel<<<grid, THREADS, 0, st>>>(a, B, K, n_cols, n_frames, (int)tiles, epi);
    BRV_LAUNCH_CHECK("frame_gemm_kernel");
    return BRV_OK;
}

}  // namespace

int brv_simt_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig,
                          int64_t samples, int64_t x_stride, float2* out,
                          int64_t n_frames, cudaStream_t st) {
    FrameLoader a{x, x_stride, samples, p->hop, brv_left(p), nullptr};
    SpecEpilogue e{reinterpret_cast<float*>(out), n_frames, 2 * p->n_bins, 2 * p->n_bins,
                   (float)(p->compression - 1.0), (float)p->scale};
    return launch_gemm(a, p->basis_fwd, p->n_fft, 2 * p->n_bins, n_sig, n_frames, e, st);
}

// inverse == true : STFT.backward (iSTFT); false : adjoint of STFT.forward.
int brv_simt_spec_to_signal(const brv_stft_plan* p, const float2* X, int64_t ss,
                            int64_t sb, int64_t sf, int64_t n_sig, int64_t n_frames,
                            int64_t out_len, bool inverse, float* y, float* ws,
                            cudaStream_t st) {
    if (n_sig == 0 || out_len == 0) return BRV_OK;
    const int N = p->n_fft;
    float* frames = ws;
    float* inv_env = ws + (size_t)n_sig * n_frames * N;
    const float* basis = inverse ? p->basis_inv : p->basis_fwd_t;
    const int K = inverse ? 2 * p->n_bins_inv : 2 * p->n_bins;
    const float pre = inverse ? (float)(1.0 / p->scale) : 1.f;
    const float expo = inverse ? (float)(1.0 / p->compression - 1.0) : 0.f;
    FrameEpilogue e{frames, n_frames, N};
    int rc;
    if (sf == 1 && sb != 1) {
        SpecLoader<true> a{X, ss, sb, sf, pre, expo};
        rc = launch_gemm(a, basis, K, N, n_sig, n_frames, e, st);
    } else {
        SpecLoader<false> a{X, ss, sb, sf, pre, expo};
        rc = launch_gemm(a, basis, K, N, n_sig, n_frames, e, st);
    }
    if (rc != BRV_OK) return rc;
    return brv_overlap_add(p, frames, n_sig, n_frames, out_len, inverse, y, inv_env, st);
}

// Overlap-add of (n_sig, T, n_fft) frames at stride hop with the centre trim;
// inverse == true divides by the overlap-added squared window (torch.istft),
// otherwise multiplies by scale_factor (adjoint of the forward transform).
int brv_overlap_add(const brv_stft_plan* p, const float* frames, int64_t n_sig, int64_t n_frames,
                    int64_t out_len, bool inverse, float* y, float* inv_env, cudaStream_t st) {
    if (n_sig == 0 || out_len == 0) return BRV_OK;
    const int N = p->n_fft;
    const int threads = 256;
    unsigned blocks = (unsigned)brv_ceil_div(out_len, threads);
    if (inverse) {
        inv_envelope_kernel<<<blocks, threads, 0, st>>>(p->window_sq, N, p->hop, n_frames, out_len, inv_env);
        BRV_LAUNCH_CHECK("inv_envelope_kernel");
    }
    BRV_REQUIRE(n_sig < 65536, "more than 65535 signals per call");
    overlap_add_kernel<<<dim3(blocks, (unsigned)n_sig), threads, 0, st>>>(
        frames, N, p->hop, inverse ? N / 2 : brv_left(p), n_frames, out_len, inverse ? inv_env : nullptr,
        (float)p->scale, y);
    BRV_LAUNCH_CHECK("overlap_add_kernel");
    return BRV_OK;
}

// Adjoint of the inverse: frames of (gy / envelope) against basis_inv^T, / scale.
int brv_simt_istft_grad(const brv_stft_plan* p, const float* gy, int64_t n_sig,
                        int64_t n_frames, float2* gX, float* ws, cudaStream_t st) {
    const int N = p->n_fft;
    const int64_t out_len = (int64_t)p->hop * (n_frames - 1) + N - 2 * (int64_t)(N / 2);
    if (n_sig == 0) return BRV_OK;
    float* inv_env = ws;
    if (out_len > 0) {
        const int threads = 256;
        inv_envelope_kernel<<<(unsigned)brv_ceil_div(out_len, threads), threads, 0, st>>>(
            p->window_sq, N, p->hop, n_frames, out_len, inv_env);
        BRV_LAUNCH_CHECK("inv_envelope_kernel");
    }
    FrameLoader a{gy, out_len, out_len, p->hop, N / 2, inv_env};
    if (p->n_bins != p->n_bins_inv)  // two-sided: the upper half never reaches the output
        BRV_CUDA(cudaMemsetAsync(gX, 0, (size_t)n_sig * n_frames * p->n_bins * sizeof(float2), st));
    SpecEpilogue e{reinterpret_cast<float*>(gX), n_frames, 2 * p->n_bins_inv, 2 * p->n_bins,
                   0.f, (float)(1.0 / p->scale)};
    return launch_gemm(a, p->basis_inv_t, N, 2 * p->n_bins_inv, n_sig, n_frames, e, st);
}
