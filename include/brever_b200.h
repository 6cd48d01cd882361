/*
 * brever_b200.h — C ABI of the B200-native time-frequency front-end.
 *
 * This is the drop-in boundary for the hot path of philgzl/brever named in
 * BASELINE.json (STFT / iSTFT, mel projection, FFNN feature extractor, SNR /
 * SI-SNR criterion).  The reference has no FFI today (it is pure Python over
 * torch); every entry point below states the reference interface (file:line,
 * relative to the brever repository root) whose arithmetic it replaces.  The
 * thin Python mirror of the reference classes lives in brever_b200/ and calls
 * these symbols through ctypes.
 *
 * Conventions
 *   - plain C types only; every tensor pointer is a DEVICE pointer owned by the
 *     caller (PyTorch), unless the parameter is documented as host memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *     and nothing synchronises the device;
 *   - return value: BRV_OK (0) or a negative brv_status; brv_last_error() gives
 *     a human-readable message for the calling thread;
 *   - complex tensors are interleaved (re, im) float32 pairs ("float2");
 *   - "frame-major" spectrogram layout = (signal, frame, bin) contiguous, which
 *     is exactly the memory torch.stft produces behind its (..., bin, frame) view.
 */
#ifndef BREVER_B200_H
#define BREVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRV_ABI_VERSION 1

typedef enum brv_status {
    BRV_OK = 0,
    BRV_ERR_INVALID = -1,     /* bad argument (maps to ValueError)            */
    BRV_ERR_UNSUPPORTED = -2, /* valid but not implemented (NotImplementedError) */
    BRV_ERR_CUDA = -3,        /* CUDA runtime error (RuntimeError)            */
    BRV_ERR_NOLA = -4,        /* window overlap-add envelope ~ 0 (RuntimeError,
                                 same condition torch.istft raises on)        */
    BRV_ERR_ALLOC = -5
} brv_status;

typedef struct brv_stft_plan brv_stft_plan; /* opaque */

/* ---- library ----------------------------------------------------------- */
int brv_abi_version(void);
const char* brv_status_string(int status);
const char* brv_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's
 * `gpu_launches` evidence). */
uint64_t brv_launch_count(void);
/* Testing hook: route STFT calls through the generic CUDA-core kernels instead
 * of the tensor-core kernels (also: env BRV_FORCE_GENERIC=1).  Returns the
 * previous setting.  Both are CUDA paths; there is no CPU path to select. */
int brv_set_force_generic(int on);
/* Testing hook (also: env BRV_TC_VARIANT).  0 (default) = symmetry-folded tensor-core
 * kernels where the plan supports them, picked per geometry and launch size; 1 = the
 * dense DFT contraction only; 2 / 3 = forward tile kernel: one tile per CTA / persistent
 * two-pass; 4 = the one-tile-per-TMEM kernels only; 5 = forward strip kernel at any size;
 * 6 / 7 = inverse strip kernel with 64- / 32-frame tiles at any size; 8 = forward strip
 * kernel with 32-frame tiles.  Every variant computes the same result to fp32 rounding;
 * the parity tests run them against each other.  Returns the previous value. */
int brv_set_tc_variant(int variant);
/* SM count / compute capability of the current device; fails without a GPU. */
int brv_device_query(int* sm_count, int* cc_major, int* cc_minor);

/* ---- STFT plan ---------------------------------------------------------
 * Replaces the constructor state of brever.modules.STFT
 * (brever/modules/stft.py:32-54).  `window` is HOST memory: `frame_length`
 * float64 samples exactly as the reference stores `self.window`
 * (scipy.signal.get_window(name, frame_length), periodic).  The plan owns the
 * device-resident DFT bases (window, normalisation 1/sqrt(sum w^2) and the
 * Hermitian weights folded in) on the device that is current at creation.
 * Created for center=True, pad_mode='constant' (stft.py:33); see
 * brv_stft_plan_set_framing / brv_reflect_pad for the other modes.           */
int brv_stft_plan_create(brv_stft_plan** plan, int frame_length, int hop_length,
                         int n_fft, const double* window, int normalized,
                         int onesided, double compression_factor,
                         double scale_factor);
int brv_stft_plan_destroy(brv_stft_plan* plan);
/* Framing of the forward transform, to be set right after creation (default 1, 1 = the
 * reference's STFT with center=True): center = 0 starts frame t at sample t * hop
 * (torch.stft center=False, stft.py:66-77); pad_to_frames = 0 skips STFT.pad's right
 * zero padding (stft.py:140-144), i.e. raw torch.stft framing, T = 1 + (S' - n_fft) / hop
 * -- what MANNER's loss uses (models/manner/stft_loss.py:33).  The inverse entry points
 * need center = 1.                                                                    */
int brv_stft_plan_set_framing(brv_stft_plan* plan, int center, int pad_to_frames);
/* torch.stft's pad_mode='reflect': out (n_signals, padded + 2 * left) = the signal, right
 * padded to `padded` samples (zeros, or its mirrored tail when right_reflect: STFT.pad uses
 * F.pad(mode=pad_mode), stft.py:140-144), mirrored by `left` samples on both sides (edge
 * sample not repeated); brv_reflect_pad_grad is its adjoint (gx: (n_signals, samples) dense). */
int brv_reflect_pad(const float* x, int64_t n_signals, int64_t samples, int64_t x_stride,
                    int64_t padded, int left, int right_reflect, float* out, void* stream);
int brv_reflect_pad_grad(const float* g, int64_t n_signals, int64_t samples, int64_t padded,
                         int left, int right_reflect, float* gx, void* stream);

/* Integer frame arithmetic of STFT.pad / STFT.frame_count + torch.stft
 * (stft.py:140-149): for `samples` input samples returns the number of frames
 * T, bins F and the right zero padding.  Bit-exact contract.                 */
int brv_stft_geometry(const brv_stft_plan* plan, int64_t samples,
                      int64_t* n_frames, int64_t* n_bins, int64_t* pad_right);
/* Output samples of the inverse for T frames: hop*(T-1) (torch.istft, centre). */
int brv_istft_geometry(const brv_stft_plan* plan, int64_t n_frames,
                       int64_t* samples);

/* ---- STFT.forward (stft.py:59-89) ---------------------------------------
 * x   : (n_signals, samples) float32, row stride `x_stride` elements.
 * out : (n_signals, T, F) complex64 frame-major.  Framing, zero padding,
 *       windowing, DFT, 1/sqrt(sum w^2), |X|^c e^{j angle X}, *scale fused.  */
int brv_stft_forward(const brv_stft_plan* plan, const float* x,
                     int64_t n_signals, int64_t samples, int64_t x_stride,
                     void* out, void* stream);

/* Adjoint of brv_stft_forward w.r.t. x for compression_factor == 1 (what
 * autograd gives the reference; needed by MultiResYuLoss, criterion.py:219).
 * gX : complex64 (n_signals, F, T) with element strides (complex units).
 * gx : (n_signals, samples) float32 contiguous.
 * workspace: brv_stft_workspace_bytes(plan, n_signals, T) bytes.            */
int brv_stft_forward_grad(const brv_stft_plan* plan, const void* gX,
                          int64_t stride_signal, int64_t stride_bin,
                          int64_t stride_frame, int64_t n_signals,
                          int64_t samples, float* gx, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- STFT.backward = iSTFT (stft.py:101-138) ------------------------------
 * X : complex64 (n_signals, F, T) with arbitrary element strides (complex
 *     units) — callers hand both frame-major views and bin-major tensors.
 * y : (n_signals, hop*(T-1)) float32 contiguous.
 * /scale, |X|^(1/c), *sqrt(sum w^2), inverse real DFT (imag of DC/Nyquist
 * ignored), window, overlap-add, / overlap-added w^2, centre trim — fused.
 * Returns BRV_ERR_NOLA where torch.istft raises.  X is NOT modified: the
 * reference divides a complex input by scale_factor in place (stft.py:114);
 * neither this entry point nor the Python mirror reproduce that side effect
 * ("Deliberate differences").                                     */
int brv_istft_forward(const brv_stft_plan* plan, const void* X,
                      int64_t stride_signal, int64_t stride_bin,
                      int64_t stride_frame, int64_t n_signals, int64_t n_frames,
                      float* y, void* workspace, size_t workspace_bytes,
                      void* stream);
/* Adjoint of brv_istft_forward w.r.t. X (compression_factor == 1).
 * gy : (n_signals, hop*(T-1)); gX : (n_signals, T, F) complex64 frame-major. */
int brv_istft_forward_grad(const brv_stft_plan* plan, const float* gy,
                           int64_t n_signals, int64_t n_frames, void* gX,
                           void* workspace, size_t workspace_bytes,
                           void* stream);
size_t brv_stft_workspace_bytes(const brv_stft_plan* plan, int64_t n_signals,
                                int64_t n_frames);
/* The same bound for one operation under the current dispatch settings: a few
 * bytes when the call will run on a fused kernel that needs no frames workspace
 * (op: 0 = brv_istft_forward, 1 = brv_stft_forward_grad, 2 = brv_istft_forward_grad),
 * else brv_stft_workspace_bytes.  Lets the caller skip the (n_signals, T, n_fft)
 * float allocation — 525 MB for 1024 x 2 x 4 s at 256 / 128 — on the common path. */
size_t brv_stft_workspace_bytes_op(const brv_stft_plan* plan, int64_t n_signals,
                                   int64_t n_frames, int op);

/* ---- float64 tensors (stft.py:59-138 called with double precision input) ----
 * torch.stft / torch.istft compute in the input's precision; the entries above are
 * fp32-grade.  These four run the same operators as direct sums in double arithmetic
 * (x, y, gx, gy: double; X, gX: complex128, same layouts and strides as above; no
 * workspace).  A fidelity path: milliseconds, not microseconds.                  */
int brv_stft_forward_f64(const brv_stft_plan* plan, const double* x,
                         int64_t n_signals, int64_t samples, int64_t x_stride,
                         void* out, void* stream);
int brv_istft_forward_f64(const brv_stft_plan* plan, const void* X,
                          int64_t stride_signal, int64_t stride_bin,
                          int64_t stride_frame, int64_t n_signals, int64_t n_frames,
                          double* y, void* stream);
int brv_stft_forward_grad_f64(const brv_stft_plan* plan, const void* gX,
                              int64_t stride_signal, int64_t stride_bin,
                              int64_t stride_frame, int64_t n_signals,
                              int64_t samples, double* gx, void* stream);
int brv_istft_forward_grad_f64(const brv_stft_plan* plan, const double* gy,
                               int64_t n_signals, int64_t n_frames, void* gX,
                               void* stream);

/* ---- ConvSTFT (stft.py:201-319) ---------------------------------------------
 * The convolutional STFT pair: analysis = F.conv1d with the windowed one-sided
 * DFT rows (DC row / sqrt(2), all / (0.5 L / sqrt(H)) when `normalized`) after
 * ConvSTFT.pad (right pad to whole frames, then L - H zeros on both sides);
 * synthesis = F.conv_transpose1d with the same filters (/ normalisation^2 when
 * not `normalized`), trimmed by L - H on both sides: (T + 1) H - L samples.
 * The plan is created by brv_stft_plan_create with the (square-root) window,
 * normalized = 0, n_fft = frame_length; compression_factor / scale_factor as in
 * the STFT entries.  Runs on the folded tensor-core kernels only (frame_length
 * in {128,256,384,512}, hop = L/4, L/2 or L): otherwise BRV_ERR_UNSUPPORTED.   */
int brv_convstft_geometry(const brv_stft_plan* plan, int64_t samples,
                          int64_t* n_frames);
int brv_convstft_forward(const brv_stft_plan* plan, const float* x,
                         int64_t n_signals, int64_t samples, int64_t x_stride,
                         int normalized, void* out /* (n_signals, T, L/2+1) complex64 */,
                         void* stream);
int brv_convstft_backward(const brv_stft_plan* plan, const void* X,
                          int64_t stride_signal, int64_t stride_bin,
                          int64_t stride_frame, int64_t n_signals,
                          int64_t n_frames, int normalized,
                          float* y /* (n_signals, (T+1)H - L) */, void* stream);

/* ---- mel filterbank (stft.py:152-198) -------------------------------------
 * Sparse (CSR) form of MelFilterbank.filters / inverse_filters applied along
 * the second-to-last axis:  out[b, r, t] = sum_j vals[j] * x[b, cols[j], t],
 * j in [rowptr[r], rowptr[r+1]).  x strides in elements; out contiguous
 * (n_batch, n_rows_out, n_frames).                                           */
int brv_mel_apply(const float* x, int64_t stride_batch, int64_t stride_row,
                  int64_t stride_frame, int64_t n_batch, int n_rows_in,
                  int64_t n_frames, const float* vals, const int32_t* cols,
                  const int32_t* rowptr, int n_rows_out, float* out,
                  void* stream);

/* ---- FeatureExtractor.fbe family + FFNN.stack/decimate + StaticNormalizer ---
 * (features.py:186-198, ffnn.py:122-135,175-187) in one pass over X:
 *   P[f,t]   = mean_c |X[b,c,f,t]|^2
 *   E[m,t]   = sum_f mel[m,f] P[f,t]             (CSR mel)
 *   E       /= sum_m E + eps                      if normalize
 *   E        = log(E + eps) | cbrt(E) | E         compression 1 | 2 | 0
 *   out[b, k*n_mel + m, t'] = (E[m, max(t'*dec - k, 0)] - mean) / std,
 *                              k = 0..stacks (mean/std nullable).
 * X : complex64 (B, C, F, T) with element strides (complex units).
 * mel_vals / mel_cols : mel_nnz CSR entries; mel_rowptr : n_mel + 1 offsets.
 * out : (B, n_mel*(stacks+1), ceil(T/dec)) float32 contiguous.               */
int brv_fbe_features(const void* X, int64_t stride_b, int64_t stride_c,
                     int64_t stride_f, int64_t stride_t, int64_t n_batch,
                     int n_channels, int n_bins, int64_t n_frames,
                     const float* mel_vals, const int32_t* mel_cols,
                     const int32_t* mel_rowptr, int n_mel, int mel_nnz,
                     int normalize, int compression, float eps, int stacks,
                     int decimation, const float* mean, const float* std,
                     float* out, void* stream);

/* ---- the other FeatureExtractor features (features.py:199-296) --------------
 * The same one-pass kernel with a selectable per-bin quantity and an optional
 * DCT stage; replaces FeatureExtractor.ild / .ipd / .ic and the dct=True branch
 * of .fbe (mfcc, cubicmfcc, pdfcc):
 *   mode 0  P = mean_c |X|^2                               (fbe family, as above)
 *   mode 1  P = 20 log10((|X[:,1]| + eps) / (|X[:,0]| + eps))   ild, features.py:238-240
 *   mode 2  P = angle(X[:,1]) - angle(X[:,0])                   ipd, features.py:258-260
 *   mode 3  P = X read as a real (B, T, F) map, strides in floats (n_channels = 1):
 *           the interaural coherence from brv_ic_coherence, then compression 3
 *           (square root) gives ic, features.py:295-296
 *   E = mel(P), normalize / compression as in brv_fbe_features (compression 3 = sqrt).
 *   n_dct > 0: cc[k] = sum_m dct_basis[k, m] E[m] (k < n_dct; the caller passes rows
 *   1..13 of the orthonormal DCT-II matrix, scipy.fft.dct(type=2, norm='ortho'),
 *   features.py:201-206) and out is (B, 3*n_dct, T) = [cc | first difference |
 *   second difference along frames], differences zero-padded on the left
 *   (features.py:207-215).  n_dct == 0: out is (B, n_mel, T).                  */
int brv_mel_features(const void* X, int64_t stride_b, int64_t stride_c,
                     int64_t stride_f, int64_t stride_t, int64_t n_batch,
                     int n_channels, int n_bins, int64_t n_frames, int mode,
                     const float* mel_vals, const int32_t* mel_cols,
                     const int32_t* mel_rowptr, int n_mel, int mel_nnz,
                     int normalize, int compression, float eps,
                     const float* dct_basis, int n_dct, float* out, void* stream);

/* Interaural coherence before the mel projection (FeatureExtractor.ic,
 * features.py:286-295): exponentially weighted auto / cross power spectra
 * phi[t] = b0 x[t] - a1 phi[t-1] along frames (torchaudio lfilter with
 * a = [1, a1 = -alpha], b = [b0 = 1 - alpha, 0], output clamped to [-1, 1]),
 * out[b, t, f] = |phi_lr|^2 / (phi_ll phi_rr), float32 (B, T, F) contiguous.  */
int brv_ic_coherence(const void* X, int64_t stride_b, int64_t stride_c,
                     int64_t stride_f, int64_t stride_t, int64_t n_batch,
                     int n_channels, int n_bins, int64_t n_frames, float b0,
                     float a1, float* out, void* stream);

/* FFNN.stack + decimate + StaticNormalizer on an existing feature tensor
 * (ffnn.py:122-135,186-187): x (B, nf, T) contiguous -> (B, nf*(stacks+1), T'). */
int brv_stack_normalize(const float* x, int64_t n_batch, int n_features,
                        int64_t n_frames, int stacks, int decimation,
                        const float* mean, const float* std, float* out,
                        void* stream);
/* CumulativeNormalizer.forward (ffnn.py:195-203): rows of length n_frames.   */
int brv_cumulative_normalize(const float* x, int64_t n_rows, int64_t n_frames,
                             float eps, float* out, void* stream);

/* ---- criteria (criterion.py:21-101,229-234) --------------------------------
 * One pass over estimate/target rows, masked to lengths[b], float64 moment
 * accumulation (no (B,S,S,L) temporaries, no Python mask loop).
 *   pairwise == 0 ("snr", criterion.py:96-101): x, y are (n_batch, n_rows, L);
 *       out_db[b, r] = 10 log10( sum y^2 / (sum (y-x)^2 + eps) + eps ).
 *   pairwise == 1 ("sisnr", criterion.py:45-61): x, y are (n_batch, S, L);
 *       out_db[b, i, j] = SI-SNR of estimate j against target i (zero-mean over
 *       the valid length).  The S! permutation max stays in the Python mirror.
 * Row r of batch b starts at b*stride_batch + r*stride_row (elements).
 * out_sign : multiplies the dB value (-1 gives the loss directly).
 * moments : (n_pairs, 6) float64 = [sum x, sum y, sum xy, sum x^2, sum y^2,
 *           sum (y-x)^2] over the valid length, kept for the backward pass
 *           (pairwise: the first five, [5] = 0; otherwise [4] and [5], rest 0).
 * workspace : brv_snr_workspace_bytes(n_pairs, length) bytes whose leading
 *           4*n_pairs bytes (rounded up to 256) must be ZERO on entry; the
 *           kernel leaves them zeroed again, so one cached buffer serves
 *           every call on a stream.                                          */
int brv_snr_forward(const float* x, const float* y, const int64_t* lengths,
                    int64_t n_batch, int64_t n_rows, int64_t length,
                    int64_t x_stride_batch, int64_t x_stride_row,
                    int64_t y_stride_batch, int64_t y_stride_row, int pairwise,
                    float eps, float out_sign, float* out_db, double* moments,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t brv_snr_workspace_bytes(int64_t n_pairs, int64_t length);

/* Gradient w.r.t. the estimate as an affine masked map per estimate row:
 *   gx[b, r, n] = ca[b,r]*x[b,r,n] + cb[b,r]*y[b, ymap[b,r], n] + c0[b,r], n < lengths[b]
 *   gx[b, r, n] = 0 otherwise.   (closed forms: SURVEY.md §8a')
 * ca/cb/c0 : (n_batch*n_rows) float32; ymap : nullable int32 (identity).      */
int brv_masked_affine(const float* x, const float* y, const int64_t* lengths,
                      int64_t n_batch, int64_t n_rows, int64_t length,
                      int64_t x_stride_batch, int64_t x_stride_row,
                      int64_t y_stride_batch, int64_t y_stride_row,
                      const float* ca, const float* cb, const float* c0,
                      const int32_t* ymap, float* gx, void* stream);

/* ---- L1 terms of MultiResYuLoss (criterion.py:135-226) ------------------------------
 * brv_l1_forward : out[b, r] = sum_{n < lengths[b]} | scale[b,r] * x[b,r,n] - y[b,r,n] |
 *                  (time-domain term, :207-209; scale nullable = 1).
 * brv_l1_backward: gx[b,r,n] = coef[b,r] * sign(scale x - y) under the length mask.
 * brv_mag_l1_forward : out[s] = sum over a contiguous complex64 spectrogram of n_elems bins
 *                  of | |X| - |Y| |  (spectral term, :213-216), X = STFT(estimate), Y = STFT(target).
 * brv_mag_l1_backward: gX = coef[s] * sign(|X| - |Y|) * X / |X| (torch's complex gradient).
 * workspace : brv_l1_workspace_bytes(rows, length) bytes, leading tickets zero on entry and
 *             left zeroed (same contract as brv_snr_forward).                              */
size_t brv_l1_workspace_bytes(int64_t n_rows_total, int64_t length);
int brv_l1_forward(const float* x, const float* y, const int64_t* lengths,
                   const float* scale, int64_t n_batch, int64_t n_rows, int64_t length,
                   int64_t x_stride_batch, int64_t x_stride_row, int64_t y_stride_batch,
                   int64_t y_stride_row, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);
int brv_l1_backward(const float* x, const float* y, const int64_t* lengths,
                    const float* scale, const float* coef, int64_t n_batch,
                    int64_t n_rows, int64_t length, int64_t x_stride_batch,
                    int64_t x_stride_row, int64_t y_stride_batch, int64_t y_stride_row,
                    float* gx, void* stream);
int brv_mag_l1_forward(const void* X, const void* Y, int64_t n_signals, int64_t n_elems,
                       float* out, void* workspace, size_t workspace_bytes, void* stream);
int brv_mag_l1_backward(const void* X, const void* Y, const float* coef,
                        int64_t n_signals, int64_t n_elems, void* gX, void* stream);

/* Gradient of the snr / sisnr loss w.r.t. the estimate in one launch: the per-row
 * coefficients of the closed forms (SURVEY.md 8a') are evaluated in the kernel from the
 * float64 `moments` brv_snr_forward saved.
 *   gout : upstream gradient of the (negated) loss, read at gout[b * gout_stride]
 *          (stride 0 broadcasts one scalar); gscale multiplies it (1 / rows for the row mean).
 *   pairwise == 0 (snr): estimate row r pairs with target row r.
 *   pairwise == 1 (sisnr): estimate row r pairs with target row ymap[b, r] (nullable =
 *          identity), moments indexed [(b, target, estimate)].                       */
int brv_criterion_backward(const float* x, const float* y, const int64_t* lengths,
                           const double* moments, const float* gout,
                           int64_t gout_stride, float gscale, const int32_t* ymap,
                           int pairwise, int64_t n_batch, int64_t n_rows,
                           int64_t length, int64_t x_stride_batch, int64_t x_stride_row,
                           int64_t y_stride_batch, int64_t y_stride_row, float eps,
                           float* gx, void* stream);

/* apply_mask (criterion.py:229-234) for callers that need the masked tensors
 * themselves: out[b, ..., n] = n < lengths[b] ? x[b, ..., n] : 0.            */
int brv_apply_mask(const float* x, const int64_t* lengths, int64_t n_batch,
                   int64_t inner, int64_t length, float* out, void* stream);

/* ---- MANNER's multi-resolution STFT loss (models/manner/stft_loss.py:22-151) ----
 * One resolution, two spectrograms (n_signals, n_elems) complex64 dense: with
 * m(.) = sqrt(max(re^2 + im^2, 1e-7)) fills sums[s] = {sum (m(Y) - m(X))^2, sum m(Y)^2,
 * sum |log m(Y) - log m(X)|} (float64; the entry zeroes `sums` itself), from which
 * spectral convergence = sqrt(s0 / s1) and log-magnitude loss = s2 / n_elems.
 * backward: gX = (k_sc (m(X) - m(Y)) + k_mag sign(m(X) - m(Y)) / m(X)) X / m(X), zero where
 * the clamp is active, with per-signal coefficients k_sc = g_sc / sqrt(s0 s1),
 * k_mag = g_mag / n_elems.                                                        */
int brv_mrstft_forward(const void* X, const void* Y, int64_t n_signals, int64_t n_elems,
                       double* sums, void* stream);
int brv_mrstft_backward(const void* X, const void* Y, const float* k_sc, const float* k_mag,
                        int64_t n_signals, int64_t n_elems, void* gX, void* stream);

/* ---- spectrogram representations (stft.py:91-110, metricganokd.py:185-195) -----
 * One elementwise pass over n complex64 values and two float32 planes of the same
 * memory order.  mode 0: (Re, Im) -- return_type / input_type 'real_imag';
 * mode 1: (|X|, angle X) -- 'mag_phase'; mode 2: (log1p(|X| + eps), angle X) and
 * X = expm1(a) e^{i b} -- MetricGAN-OKD's stft / istft wrappers.  X, a, b 16-byte
 * aligned for split; the *_grad entries are the adjoints (gradient of a complex
 * tensor = dL/dRe + i dL/dIm; |X| = 0 gets a zero gradient; ga / gb nullable in
 * split_grad).                                                                    */
int brv_spec_split(const void* X, int64_t n, int mode, float eps, float* a, float* b,
                   void* stream);
int brv_spec_join(const float* a, const float* b, int64_t n, int mode, void* X,
                  void* stream);
int brv_spec_split_grad(const float* ga, const float* gb, const void* X, int64_t n,
                        int mode, float eps, void* gX, void* stream);
int brv_spec_join_grad(const void* gX, const float* a, const float* b, int64_t n, int mode,
                       float* ga, float* gb, void* stream);

/* FFNN._enhance's `x.mean(1) * mask` (models/ffnn/ffnn.py:107-110) in one pass:
 * out[b, t, f] = mask[b, f, t] / C * sum_c X[b, c, f, t]; X complex64 and mask float32
 * (nullable) with element strides, out (n_batch, n_frames, n_bins) complex64 dense --
 * frame-major, the layout the iSTFT kernels stream.                                  */
int brv_channel_mean_mask(const void* X, int64_t x_stride_batch, int64_t x_stride_channel,
                          int64_t x_stride_bin, int64_t x_stride_frame, const float* mask,
                          int64_t m_stride_batch, int64_t m_stride_bin, int64_t m_stride_frame,
                          int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                          void* out, void* stream);
/* *total += mean(v[0..n)): the running metric of the training loop (training.py:369-373). */
int brv_accumulate_mean(const float* v, int64_t n, float* total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BREVER_B200_H */
