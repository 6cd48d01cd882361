// C entry points for the STFT / iSTFT pair: argument checks, geometry, and the
// choice between the tensor-core kernels (brv_stft_tc.cu) and the generic
// CUDA-core kernels (brv_stft_simt.cu).  There is no CPU path.
#include <math.h>
#include <stdlib.h>

#include "brv_common.cuh"

// brv_stft_tc.cu
bool brv_tc_supported(const brv_stft_plan* p);
int brv_tc_spec_to_frames(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb,
                          int64_t sf, int64_t n_sig, int64_t n_frames, float* frames,
                          cudaStream_t st);
int brv_tc_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                        int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st);

// brv_stft_fold.cu
bool brv_fold_supported(const brv_stft_plan* p);
int brv_fold_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                          int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st);
bool brv_fold_inverse_supported(const brv_stft_plan* p);
bool brv_fold_grad_supported(const brv_stft_plan* p);
int brv_fold_istft(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb, int64_t sf,
                   int64_t n_sig, int64_t n_frames, int64_t out_len, float* y, cudaStream_t st);

int brv_fold_istft_grad(const brv_stft_plan* p, const float* gy, int64_t n_sig, int64_t n_frames,
                        int64_t out_len, float2* gX, cudaStream_t st);

int brv_fold_stft_grad(const brv_stft_plan* p, const float2* gX, int64_t ss, int64_t sb, int64_t sf,
                       int64_t n_sig, int64_t n_frames, int64_t samples, float* gx,
                       cudaStream_t st);

bool brv_fold_conv_supported(const brv_stft_plan* p);
int brv_fold_conv_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                          int64_t x_stride, double gain, float2* out, int64_t n_frames,
                          cudaStream_t st);
int brv_fold_conv_backward(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb,
                           int64_t sf, int64_t n_sig, int64_t n_frames, int64_t out_len,
                           double gain, float* y, cudaStream_t st);

static int g_force_generic = -1;
extern int g_brv_fold_variant;
// 0: folded kernels where supported (transposed strip kernels first), 1: dense contraction only,
// 4: the one-tile-per-TMEM folded kernels (forward picked by tile count), 2 / 3: those with the
// forward forced to one tile per CTA / persistent two-pass
static int g_tc_variant = -1;

static int tc_variant() {
    if (g_tc_variant < 0) {
        const char* e = getenv("BRV_TC_VARIANT");
        const int v = e ? atoi(e) : 0;
        g_tc_variant = (v >= 1 && v <= 8) ? v : 0;
        g_brv_fold_variant = g_tc_variant >= 2 ? g_tc_variant : 0;
    }
    return g_tc_variant;
}

extern "C" int brv_set_tc_variant(int variant) {
    int prev = tc_variant();
    g_tc_variant = (variant >= 1 && variant <= 8) ? variant : 0;
    g_brv_fold_variant = g_tc_variant >= 2 ? g_tc_variant : 0;
    return prev;
}


static int force_generic() {
    if (g_force_generic < 0) {
        const char* e = getenv("BRV_FORCE_GENERIC");
        g_force_generic = (e && e[0] == '1') ? 1 : 0;
    }
    return g_force_generic;
}

extern "C" int brv_set_force_generic(int on) {
    int prev = force_generic();
    g_force_generic = on ? 1 : 0;
    return prev;
}

extern "C" size_t brv_stft_workspace_bytes_op(const brv_stft_plan* p, int64_t n_signals,
                                              int64_t n_frames, int op) {
    if (!p || n_signals <= 0 || n_frames <= 0) return 0;
    // mirrors the dispatch conditions of the three entry points below
    const bool fold_ok = !force_generic() && tc_variant() != 1;
    bool fused = false;
    if (op == 0)
        fused = fold_ok && brv_fold_inverse_supported(p);
    else if (op == 1)
        fused = fold_ok && brv_fold_inverse_supported(p) && brv_fold_grad_supported(p) &&
                p->n_bins == p->n_bins_inv;
    else if (op == 2)
        fused = fold_ok && brv_fold_grad_supported(p) && p->n_bins == p->n_bins_inv &&
                (int64_t)p->hop * (n_frames - 1) + p->n_fft - 2 * (int64_t)(p->n_fft / 2) > 0;
    return fused ? 256 : brv_stft_workspace_bytes(p, n_signals, n_frames);
}

extern "C" int brv_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_signals,
                                int64_t samples, int64_t x_stride, void* out, void* stream) {
    BRV_REQUIRE(p && out, "null pointer argument");
    BRV_REQUIRE(n_signals >= 0 && samples >= 0, "negative shape");
    BRV_REQUIRE(x || n_signals * samples == 0, "input pointer is null");
    BRV_REQUIRE(x_stride >= samples || n_signals <= 1, "row stride smaller than the row");
    int64_t n_frames = 0;
    int rc = brv_stft_geometry(p, samples, &n_frames, nullptr, nullptr);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0) return BRV_OK;
    if (!force_generic() && tc_variant() != 1 && brv_fold_supported(p))
        return brv_fold_stft_forward(p, x, n_signals, samples, x_stride, (float2*)out, n_frames,
                                     (cudaStream_t)stream);
    if (!force_generic() && brv_tc_supported(p))
        return brv_tc_stft_forward(p, x, n_signals, samples, x_stride, (float2*)out, n_frames,
                                   (cudaStream_t)stream);
    return brv_simt_stft_forward(p, x, n_signals, samples, x_stride, (float2*)out, n_frames,
                                 (cudaStream_t)stream);
}

extern "C" int brv_stft_forward_grad(const brv_stft_plan* p, const void* gX, int64_t ss,
                                     int64_t sb, int64_t sf, int64_t n_signals, int64_t samples,
                                     float* gx, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    BRV_REQUIRE(p && gX && gx, "null pointer argument");
    if (p->compression != 1.0)
        return brv_fail(BRV_ERR_UNSUPPORTED,
                        "gradient of the compressed STFT (compression_factor != 1) is not "
                        "implemented: no reference model back-propagates through it");
    int64_t n_frames = 0;
    int rc = brv_stft_geometry(p, samples, &n_frames, nullptr, nullptr);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0 || samples == 0) return BRV_OK;
    if (!force_generic() && tc_variant() != 1 && brv_fold_inverse_supported(p) &&
        brv_fold_grad_supported(p) && p->n_bins == p->n_bins_inv)
        return brv_fold_stft_grad(p, (const float2*)gX, ss, sb, sf, n_signals, n_frames, samples,
                                  gx, (cudaStream_t)stream);
    BRV_REQUIRE(workspace && workspace_bytes >= brv_stft_workspace_bytes(p, n_signals, n_frames),
                "workspace too small");
    return brv_simt_spec_to_signal(p, (const float2*)gX, ss, sb, sf, n_signals, n_frames, samples,
                                   false, gx, (float*)workspace, (cudaStream_t)stream);
}

extern "C" int brv_istft_forward(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb,
                                 int64_t sf, int64_t n_signals, int64_t n_frames, float* y,
                                 void* workspace, size_t workspace_bytes, void* stream) {
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
    if (!force_generic() && tc_variant() != 1 && brv_fold_inverse_supported(p))
        return brv_fold_istft(p, (const float2*)X, ss, sb, sf, n_signals, n_frames, out_len, y,
                              (cudaStream_t)stream);   // fused overlap-add: no workspace
    BRV_REQUIRE(workspace && workspace_bytes >= brv_stft_workspace_bytes(p, n_signals, n_frames),
                "workspace too small");
    if (!force_generic() && brv_tc_supported(p)) {
        float* frames = (float*)workspace;
        float* inv_env = frames + (size_t)n_signals * n_frames * p->n_fft;
        rc = brv_tc_spec_to_frames(p, (const float2*)X, ss, sb, sf, n_signals, n_frames, frames,
                                   (cudaStream_t)stream);
        if (rc != BRV_OK) return rc;
        return brv_overlap_add(p, frames, n_signals, n_frames, out_len, true, y, inv_env,
                               (cudaStream_t)stream);
    }
    return brv_simt_spec_to_signal(p, (const float2*)X, ss, sb, sf, n_signals, n_frames, out_len,
                                   true, y, (float*)workspace, (cudaStream_t)stream);
}

extern "C" int brv_istft_forward_grad(const brv_stft_plan* p, const float* gy, int64_t n_signals,
                                      int64_t n_frames, void* gX, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    BRV_REQUIRE(p && gy && gX, "null pointer argument");
    if (!p->center)
        return brv_fail(BRV_ERR_UNSUPPORTED, "the inverse transform is implemented for center=True plans only");
    if (p->compression != 1.0)
        return brv_fail(BRV_ERR_UNSUPPORTED,
                        "gradient of the decompressing iSTFT (compression_factor != 1) is not "
                        "implemented: the reference only runs it under no_grad");
    if (n_signals == 0) return BRV_OK;
    if (!force_generic() && tc_variant() != 1 && brv_fold_grad_supported(p) && p->n_bins == p->n_bins_inv) {
        int64_t out_len = 0;
        int rc = brv_istft_geometry(p, n_frames, &out_len);
        if (rc != BRV_OK) return rc;
        if (out_len > 0)
            return brv_fold_istft_grad(p, gy, n_signals, n_frames, out_len, (float2*)gX,
                                       (cudaStream_t)stream);
    }
    BRV_REQUIRE(workspace && workspace_bytes >= brv_stft_workspace_bytes(p, n_signals, n_frames),
                "workspace too small");
    return brv_simt_istft_grad(p, gy, n_signals, n_frames, (float2*)gX, (float*)workspace,
                               (cudaStream_t)stream);
}

// ---- ConvSTFT (brever/modules/stft.py:201-319) ------------------------------------------
int brv_direct_conv_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                            int64_t x_stride, double gain, void* out, int64_t n_frames, cudaStream_t st);
int brv_direct_conv_backward(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb, int64_t sf,
                             int64_t n_sig, int64_t n_frames, int64_t out_len, double gain, float* y,
                             cudaStream_t st);

// frame_length in {128, 256, 384, 512} with hop = L/4, L/2 or L runs on the folded tensor-core
// kernels; any other size on the direct-sum kernels of brv_stft_f64.cu (correct, not fast)
static int conv_check(const brv_stft_plan* p) {
    BRV_REQUIRE(p, "plan is null");
    if (p->normalized || p->n_fft != p->frame_length || p->frame_length < p->hop || !p->onesided)
        return brv_fail(BRV_ERR_UNSUPPORTED,
                        "ConvSTFT needs a one-sided plan created with normalized = 0, n_fft = frame_length "
                        "and hop_length <= frame_length");
    return BRV_OK;
}

// 0.5 L / sqrt(H), stft.py:232
static double conv_normalization(const brv_stft_plan* p) {
    return 0.5 * p->frame_length / sqrt((double)p->hop);
}

extern "C" int brv_convstft_geometry(const brv_stft_plan* p, int64_t samples, int64_t* n_frames) {
    int rc = conv_check(p);
    if (rc != BRV_OK) return rc;
    BRV_REQUIRE(samples >= 0, "negative sample count");
    // ConvSTFT.pad (stft.py:305-315): right pad to whole frames, then L - H on both sides;
    // F.conv1d with stride H yields (len - L) / H + 1 frames
    const int64_t over = samples > p->frame_length ? samples - p->frame_length : 0;
    const int64_t frames0 = brv_ceil_div(over, p->hop) + 1;
    const int64_t len = (frames0 - 1) * p->hop + p->frame_length +
                        2 * (int64_t)(p->frame_length - p->hop);
    if (n_frames) *n_frames = (len - p->frame_length) / p->hop + 1;
    return BRV_OK;
}

extern "C" int brv_convstft_forward(const brv_stft_plan* p, const float* x, int64_t n_signals,
                                    int64_t samples, int64_t x_stride, int normalized, void* out,
                                    void* stream) {
    int rc = conv_check(p);
    if (rc != BRV_OK) return rc;
    BRV_REQUIRE(out && (x || n_signals * samples == 0), "null pointer argument");
    BRV_REQUIRE(n_signals >= 0 && samples >= 0, "negative shape");
    int64_t n_frames = 0;
    rc = brv_convstft_geometry(p, samples, &n_frames);
    if (rc != BRV_OK) return rc;
    if (n_signals == 0) return BRV_OK;
    const double gain = normalized ? 1.0 / conv_normalization(p) : 1.0;
    if (force_generic() || !brv_fold_conv_supported(p))
        return brv_direct_conv_forward(p, x, n_signals, samples, x_stride, gain, out, n_frames,
                                       (cudaStream_t)stream);
    return brv_fold_conv_forward(p, x, n_signals, samples, x_stride, gain, (float2*)out, n_frames,
                                 (cudaStream_t)stream);
}

extern "C" int brv_convstft_backward(const brv_stft_plan* p, const void* X, int64_t ss, int64_t sb,
                                     int64_t sf, int64_t n_signals, int64_t n_frames,
                                     int normalized, float* y, void* stream) {
    int rc = conv_check(p);
    if (rc != BRV_OK) return rc;
    BRV_REQUIRE(X && n_signals >= 0 && n_frames >= 1, "bad arguments");
    // conv_transpose1d gives (T - 1) H + L samples; L - H are cut from both ends (stft.py:294-298)
    const int64_t out_len = (n_frames + 1) * p->hop - p->frame_length;
    if (n_signals == 0 || out_len <= 0) return BRV_OK;
    BRV_REQUIRE(y, "output pointer is null");
    const double nf = conv_normalization(p);
    const double gain = normalized ? 1.0 / nf : 1.0 / (nf * nf);
    if (force_generic() || !brv_fold_conv_supported(p))
        return brv_direct_conv_backward(p, X, ss, sb, sf, n_signals, n_frames, out_len, gain, y,
                                        (cudaStream_t)stream);
    return brv_fold_conv_backward(p, (const float2*)X, ss, sb, sf, n_signals, n_frames, out_len,
                                  gain, y, (cudaStream_t)stream);
}
