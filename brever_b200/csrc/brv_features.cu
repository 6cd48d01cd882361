// Mel projection and the FFNN feature extractor: |X|^2 -> channel mean -> banded
// (CSR) mel -> optional pdf normalisation -> log / cubic-root -> context stacking
// with edge replication -> decimation -> static normalisation, in one kernel.
//
// Reference: brever/modules/stft.py:152-198 (MelFilterbank),
// brever/modules/features.py:186-198 (FeatureExtractor.fbe),
// brever/models/ffnn/ffnn.py:122-135,175-203 (stack, decimate, normalisers).
//
// The same kernel also produces the binaural cues (features.py:222-296: ild, ipd read both
// channels in phase 1; ic arrives as a real per-bin coherence from ic_coherence_kernel) and
// the DCT features (features.py:199-219: mfcc, cubicmfcc, pdfcc = DCT-II of the compressed
// energies, coefficients 1..13, plus first / second frame differences).
//
// The kernel is HBM-bound: per frame it reads C*F complex bins once (8 B each)
// and writes n_mel*(stacks+1) floats.  The mel matrix is 97 % zeros (<= 2
// filters per bin), so it is applied as a CSR gather from a shared-memory
// power spectrum instead of a dense GEMM.
#include "brv_common.cuh"

namespace {

constexpr int FT_THREADS = 256;
constexpr int FT_MAX_STACK = 32;
enum { FT_MODE_POWER = 0, FT_MODE_ILD = 1, FT_MODE_IPD = 2, FT_MODE_REAL = 3 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One CTA = `tile` output frames (pre-decimation) of one batch item, plus the `stacks`
// frames of left context the stacker reaches back to.
//   phase 1  all warps: |X|^2, channel mean -> power[staged][n_bins]   (coalesced 8-byte loads,
//            four independent loads in flight per thread)
//   phase 2  one thread per (frame, mel band): CSR row (values / columns staged in shared
//            memory once per CTA) dotted with the frame's power row
//   phase 3  optional pdf normalisation, log / cubic root
//   phase 4  stack + decimate + (x - mean) / std, frames contiguous in the output
// dynamic smem: vals[nnz] | cols[nnz] | rowptr[n_mel + 1] | power[staged][ldP] | energy[staged][n_mel + 1]
__global__ void __launch_bounds__(FT_THREADS)
fbe_features_kernel(const float2* __restrict__ X, int64_t sb, int64_t sc, int64_t sf,
                    int64_t st, int C, int n_bins, int64_t n_frames,
                    const float* __restrict__ mel_vals, const int32_t* __restrict__ mel_cols,
                    const int32_t* __restrict__ mel_rowptr, int n_mel, int nnz, int normalize,
                    int compression, float eps, int stacks, int decimation,
                    const float* __restrict__ mean, const float* __restrict__ stdv,
                    float* __restrict__ out, int64_t out_frames, int tile, int tiles_per_item,
                    int mode, const float* __restrict__ dct_basis, int n_dct) {
    extern __shared__ float smem[];
    const int warps = FT_THREADS / 32;
    const int ctx = n_dct ? 2 : stacks;                    // frames of left context staged
    const int staged = tile + ctx;
    const int ldP = n_bins | 1;                            // odd pitch: frames hit distinct banks
    const int ldE = n_mel + 1;
    float* vals = smem;
    int32_t* cols = reinterpret_cast<int32_t*>(smem + nnz);
    int32_t* rowptr = cols + nnz;
    float* power = reinterpret_cast<float*>(rowptr + n_mel + 1);
    float* energy = power + (size_t)staged * ldP;
    float* norm_tab = energy + (size_t)staged * ldE;       // mean[rows] | std[rows] (stack path)

    const int64_t b = blockIdx.x / tiles_per_item;
    const int64_t t0 = (int64_t)(blockIdx.x % tiles_per_item) * tile;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t first = t0 - ctx;                        // first frame staged (may be < 0)
    const float inv_c = 1.f / (float)C;

    for (int j = threadIdx.x; j < nnz; j += FT_THREADS) {
        vals[j] = __ldg(mel_vals + j);
        cols[j] = __ldg(mel_cols + j);
    }
    for (int j = threadIdx.x; j <= n_mel; j += FT_THREADS) rowptr[j] = __ldg(mel_rowptr + j);
    if (!n_dct) {
        const int rows_n = n_mel * (stacks + 1);
        for (int j = threadIdx.x; j < rows_n; j += FT_THREADS) {
            norm_tab[j] = mean ? __ldg(mean + j) : 0.f;
            norm_tab[rows_n + j] = stdv ? __ldg(stdv + j) : 1.f;
        }
    }

    // ---- phase 1: per-bin quantity -> power[staged][n_bins] ----
    const float2* xb = X + b * sb;
    if (mode == FT_MODE_REAL) {
        // a real (frames, bins) map computed upstream (interaural coherence): strides in floats
        const float* rb = reinterpret_cast<const float*>(X) + b * sb;
        for (int s = warp; s < staged; s += warps) {
            const int64_t t = first + s;
            if (t < 0 || t >= n_frames) continue;
            for (int f = lane; f < n_bins; f += 32)
                power[(size_t)s * ldP + f] = __ldg(rb + t * st + (int64_t)f * sf);
        }
    } else if (mode != FT_MODE_POWER) {
        // ild: 20 log10((|R| + eps) / (|L| + eps)), features.py:238-239
        // ipd: angle(R) - angle(L), features.py:258-259          (L = channel 0, R = channel 1)
        const bool bins_fast = sf <= st;
        const int n_outer = bins_fast ? staged : n_bins, n_inner = bins_fast ? n_bins : staged;
        for (int o = warp; o < n_outer; o += warps) {
            for (int i = lane; i < n_inner; i += 32) {
                const int s = bins_fast ? o : i, f = bins_fast ? i : o;
                const int64_t t = first + s;
                if (t < 0 || t >= n_frames) continue;
                const float2 l = __ldg(xb + (int64_t)f * sf + t * st);
                const float2 r = __ldg(xb + sc + (int64_t)f * sf + t * st);
                float v;
                if (mode == FT_MODE_ILD)
                    v = 20.f * log10f((hypotf(r.x, r.y) + eps) / (hypotf(l.x, l.y) + eps));
                else
                    v = atan2f(r.y, r.x) - atan2f(l.y, l.x);
                power[(size_t)s * ldP + f] = v;
            }
        }
    } else if (sf <= st) {
        // power spectrum, channel mean (features.py:186-188)
        // bins contiguous (frame-major spectrogram, what STFT.forward returns): lanes along bins
        for (int s = warp; s < staged; s += warps) {
            const int64_t t = first + s;
            if (t < 0 || t >= n_frames) continue;          // t < 0 is never read: the stacker
                                                           // clamps to frame 0 (ffnn.py:126)
            const float2* base = xb + t * st;
            float* pw = power + (size_t)s * ldP;
            for (int f0 = 0; f0 < n_bins; f0 += 128) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int c = 0; c < C; ++c) {
                    float2 v[4];
                    const float2* bc = base + (int64_t)c * sc;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int f = f0 + u * 32 + lane;
                        v[u] = f < n_bins ? __ldg(sf == 1 ? bc + f : bc + (int64_t)f * sf)
                                          : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = fmaf(v[u].y, v[u].y, fmaf(v[u].x, v[u].x, acc[u]));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int f = f0 + u * 32 + lane;
                    if (f < n_bins) pw[f] = acc[u] * inv_c;
                }
            }
        }
    } else {
        // frames contiguous (bin-major spectrogram): lanes along frames
        for (int f = warp; f < n_bins; f += warps) {
            for (int s = lane; s < staged; s += 32) {
                const int64_t t = first + s;
                if (t < 0 || t >= n_frames) continue;
                float acc = 0.f;
                for (int c = 0; c < C; ++c) {
                    const float2 v = __ldg(xb + (int64_t)c * sc + (int64_t)f * sf + t * st);
                    acc = fmaf(v.y, v.y, fmaf(v.x, v.x, acc));
                }
                power[(size_t)s * ldP + f] = acc * inv_c;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: banded mel projection (stft.py:189-190) ----
    // lanes along the staged frames of one band (padded to a power of two), so that a warp's lanes
    // walk CSR rows of (nearly) the same length: the filters are 2 .. 35 bins wide
    const int sp_shift = 32 - __clz(staged - 1 > 0 ? staged - 1 : 1);
    for (int idx = threadIdx.x; idx < (n_mel << sp_shift); idx += FT_THREADS) {
        const int s = idx & ((1 << sp_shift) - 1), m = idx >> sp_shift;
        const int64_t t = first + s;
        if (s >= staged || t < 0 || t >= n_frames) continue;
        const float* pw = power + (size_t)s * ldP;
        float e = 0.f;
        for (int j = rowptr[m]; j < rowptr[m + 1]; ++j) e = fmaf(vals[j], pw[cols[j]], e);
        energy[s * ldE + m] = e;
    }
    __syncthreads();

    // ---- phase 3: pdf normalisation (features.py:192-193), compression (:195-198) ----
    if (normalize) {
        for (int s = warp; s < staged; s += warps) {
            float part = 0.f;
            for (int m = lane; m < n_mel; m += 32) part += energy[s * ldE + m];
            const float total = warp_sum(part) + eps;
            for (int m = lane; m < n_mel; m += 32) energy[s * ldE + m] /= total;
        }
        __syncthreads();
    }
    if (compression) {
        // (the pad column of every row is transformed too: never read)
        for (int idx = threadIdx.x; idx < staged * ldE; idx += FT_THREADS) {
            float e = energy[idx];
            e = compression == 1 ? logf(e + eps) : (compression == 2 ? cbrtf(e) : sqrtf(e));
            energy[idx] = e;
        }
        __syncthreads();
    }

    if (n_dct) {
        // ---- DCT-II (ortho) coefficients 1..n_dct of every staged frame, then
        //      out[b, k] = cc[k, t], out[b, n_dct + k] = cc[k, t] - cc[k, t-1] (0 for t < 1),
        //      out[b, 2 n_dct + k] = cc[k, t] - 2 cc[k, t-1] + cc[k, t-2] (0 for t < 2)
        //      (scipy.fft.dct + np.diff + left zero pad, features.py:201-215) ----
        float* cc = power;                                 // the power rows are dead by now
        const int ldC = n_dct + 1;
        for (int idx = threadIdx.x; idx < staged * n_dct; idx += FT_THREADS) {
            const int s = idx / n_dct, k = idx - s * n_dct;
            const int64_t t = first + s;
            float acc = 0.f;
            if (t >= 0 && t < n_frames)
                for (int m = 0; m < n_mel; ++m)
                    acc = fmaf(__ldg(dct_basis + k * n_mel + m), energy[s * ldE + m], acc);
            cc[s * ldC + k] = acc;
        }
        __syncthreads();
        int64_t t_end = t0 + tile < n_frames ? t0 + tile : n_frames;
        const int width = (int)(t_end - t0);
        for (int idx = threadIdx.x; idx < 3 * n_dct * width; idx += FT_THREADS) {
            const int r = idx / width, w = idx - r * width;
            const int order = r / n_dct, k = r - order * n_dct;
            const int64_t t = t0 + w;
            const int s = w + ctx;
            float v = cc[s * ldC + k];
            if (order == 1) v = t >= 1 ? v - cc[(s - 1) * ldC + k] : 0.f;
            if (order == 2)
                v = t >= 2 ? v - 2.f * cc[(s - 1) * ldC + k] + cc[(s - 2) * ldC + k] : 0.f;
            out[(b * 3 * n_dct + r) * n_frames + t] = v;
        }
        return;
    }

    // ---- phase 4: out[b, k*n_mel + m, t/dec] = (E[m, max(t - k, 0)] - mean) / std ----
    const int rows = n_mel * (stacks + 1);
    const int64_t o0 = brv_ceil_div(t0, decimation);       // first output frame of this tile
    int64_t t_end = t0 + tile < n_frames ? t0 + tile : n_frames;
    const int64_t o1 = brv_ceil_div(t_end, decimation);
    const int width = (int)(o1 - o0);
    if (width <= 0) return;
    // thread = (band m, output frame w) with the frame count padded to a power of two: no integer
    // divisions per element; mean / std come from shared memory
    const int wp_shift = 32 - __clz(width - 1 > 0 ? width - 1 : 1);
    float* ob = out + b * rows * out_frames + o0;
    for (int k = 0; k <= stacks; ++k) {
        for (int idx = threadIdx.x; idx < (n_mel << wp_shift); idx += FT_THREADS) {
            const int w = idx & ((1 << wp_shift) - 1), m = idx >> wp_shift;
            if (w >= width) continue;
            const int r = k * n_mel + m;
            int64_t src = (o0 + w) * decimation - k;
            if (src < 0) src = 0;
            float v = energy[(int)(src - first) * ldE + m] - norm_tab[r];
            if (stdv) v /= norm_tab[rows + r];
            ob[r * out_frames + w] = v;
        }
    }
}

__global__ void mel_apply_kernel(const float* __restrict__ x, int64_t sb, int64_t sr, int64_t st,
                                 int64_t n_frames, const float* __restrict__ vals,
                                 const int32_t* __restrict__ cols,
                                 const int32_t* __restrict__ rowptr, int n_rows_out,
                                 float* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    int64_t b = blockIdx.z;
    if (t >= n_frames) return;
    const float* base = x + b * sb + t * st;
    float acc = 0.f;
    for (int j = rowptr[r]; j < rowptr[r + 1]; ++j)
        acc = fmaf(__ldg(vals + j), __ldg(base + (int64_t)cols[j] * sr), acc);
    out[(b * n_rows_out + r) * n_frames + t] = acc;
}

__global__ void stack_normalize_kernel(const float* __restrict__ x, int nf, int64_t n_frames,
                                       int stacks, int decimation,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ stdv, float* __restrict__ out,
                                       int64_t out_frames) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;                      // output row = k*nf + f
    int64_t b = blockIdx.z;
    if (o >= out_frames) return;
    int k = r / nf, f = r % nf;
    int64_t src = o * decimation - k;
    if (src < 0) src = 0;
    float v = __ldg(x + (b * nf + f) * n_frames + src);
    if (mean) v -= __ldg(mean + r);
    if (stdv) v /= __ldg(stdv + r);
    out[(b * (int64_t)nf * (stacks + 1) + r) * out_frames + o] = v;
}

// One warp per row: running mean / variance along frames (ffnn.py:195-203),
// accumulated in float64 (the float32 reference cancels in E[x^2]-E[x]^2).
__global__ void cumulative_normalize_kernel(const float* __restrict__ x, int64_t n_rows,
                                            int64_t n_frames, float eps,
                                            float* __restrict__ out) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= n_rows) return;
    const float* src = x + row * n_frames;
    float* dst = out + row * n_frames;
    double run_s = 0, run_q = 0;
    for (int64_t t0 = 0; t0 < n_frames; t0 += 32) {
        int64_t t = t0 + lane;
        float v = t < n_frames ? src[t] : 0.f;
        double s = v, q = (double)v * v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {            // inclusive warp scan
            double s2 = __shfl_up_sync(0xffffffffu, s, o);
            double q2 = __shfl_up_sync(0xffffffffu, q, o);
            if (lane >= o) { s += s2; q += q2; }
        }
        s += run_s;
        q += run_q;
        if (t < n_frames) {
            double cnt = (double)(t + 1);
            double mu = s / cnt;
            double var = q / cnt - mu * mu;
            dst[t] = (float)(((double)v - mu) / sqrt(var + (double)eps));
        }
        run_s = __shfl_sync(0xffffffffu, s, 31);
        run_q = __shfl_sync(0xffffffffu, q, 31);
    }
}

// Interaural coherence before the mel projection (features.py:262-296): one thread per
// (batch item, bin) walks the frames with the four recursions
//   phi[t] = (1 - alpha) x[t] + alpha phi[t-1],  x in {|L|^2, |R|^2, Re L conj(R), Im L conj(R)}
// (torchaudio.functional.lfilter, a = [1, -alpha], b = [1 - alpha, 0]); the filter OUTPUT is
// clamped to [-1, 1] (lfilter's default clamp=True) and coh = |phi_lr|^2 / (phi_ll phi_rr).
// Loads of IC_UNROLL frames are issued together; out is (B, T, F) float, bins contiguous.
constexpr int IC_UNROLL = 8;
__global__ void __launch_bounds__(128)
ic_coherence_kernel(const float2* __restrict__ X, int64_t sb, int64_t sc, int64_t sf, int64_t st,
                    int n_bins, int64_t n_frames, float b0, float a1, float* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (f >= n_bins) return;
    const float2* xl = X + b * sb + (int64_t)f * sf;
    const float2* xr = xl + sc;
    float* o = out + b * n_frames * n_bins + f;
    float ll = 0.f, rr = 0.f, cre = 0.f, cim = 0.f;
    for (int64_t t0 = 0; t0 < n_frames; t0 += IC_UNROLL) {
        float2 l[IC_UNROLL], r[IC_UNROLL];
#pragma unroll
        for (int u = 0; u < IC_UNROLL; ++u) {
            const bool ok = t0 + u < n_frames;
            l[u] = ok ? __ldg(xl + (t0 + u) * st) : make_float2(0.f, 0.f);
            r[u] = ok ? __ldg(xr + (t0 + u) * st) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < IC_UNROLL; ++u) {
            if (t0 + u >= n_frames) break;
            ll = fmaf(b0, l[u].x * l[u].x + l[u].y * l[u].y, -a1 * ll);
            rr = fmaf(b0, r[u].x * r[u].x + r[u].y * r[u].y, -a1 * rr);
            cre = fmaf(b0, l[u].x * r[u].x + l[u].y * r[u].y, -a1 * cre);
            cim = fmaf(b0, l[u].y * r[u].x - l[u].x * r[u].y, -a1 * cim);
            const float cl = fminf(fmaxf(ll, -1.f), 1.f), cr = fminf(fmaxf(rr, -1.f), 1.f);
            const float cx = fminf(fmaxf(cre, -1.f), 1.f), cy = fminf(fmaxf(cim, -1.f), 1.f);
            o[(t0 + u) * n_bins] = (cx * cx + cy * cy) / (cl * cr);
        }
    }
}

}  // namespace

static int launch_features(const void* X, int64_t sb, int64_t sc, int64_t sf, int64_t st,
                           int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                           const float* mel_vals, const int32_t* mel_cols,
                           const int32_t* mel_rowptr, int n_mel, int nnz, int normalize,
                           int compression, float eps, int stacks, int decimation,
                           const float* mean, const float* stdv, float* out, void* stream,
                           int mode, const float* dct_basis, int n_dct) {
    BRV_REQUIRE(X && mel_vals && mel_cols && mel_rowptr && out, "null pointer argument");
    BRV_REQUIRE(n_channels >= 1 && n_bins >= 1 && n_mel >= 1, "bad feature dimensions");
    BRV_REQUIRE(compression >= 0 && compression <= 3,
                "compression must be 0 (none), 1 (log), 2 (cubic) or 3 (square root)");
    BRV_REQUIRE(mode >= FT_MODE_POWER && mode <= FT_MODE_REAL, "unknown feature mode %d", mode);
    BRV_REQUIRE((mode != FT_MODE_ILD && mode != FT_MODE_IPD) || n_channels >= 2,
                "interaural features need two channels, got %d", n_channels);
    BRV_REQUIRE(n_dct >= 0 && n_dct <= n_mel && (n_dct == 0 || dct_basis), "bad DCT arguments");
    BRV_REQUIRE(n_dct == 0 || (stacks == 0 && decimation == 1 && !mean && !stdv),
                "the DCT features are not fused with stacking / decimation / normalisation");
    BRV_REQUIRE(stacks >= 0 && stacks <= FT_MAX_STACK, "stacks must be in [0, %d]", FT_MAX_STACK);
    BRV_REQUIRE(decimation >= 1, "decimation must be >= 1");
    if (n_batch == 0 || n_frames == 0) return BRV_OK;
    const int64_t out_frames = brv_ceil_div(n_frames, decimation);
    // frames per CTA: the largest tile that still fills the machine (>= 4 CTAs per SM)
    int sm_count = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    int tile = 64;
    while (tile > 8 && n_batch * brv_ceil_div(n_frames, tile) < 4LL * sm_count) tile /= 2;
    const int tiles = (int)brv_ceil_div(n_frames, tile);
    const int64_t grid = n_batch * tiles;
    BRV_REQUIRE(grid < (1LL << 31), "too many feature tiles");
    BRV_REQUIRE(nnz >= 0 && nnz <= n_mel * n_bins, "bad CSR row pointer");
    const int staged = tile + (n_dct ? 2 : stacks);
    size_t smem = ((size_t)2 * nnz + n_mel + 1 + (size_t)staged * (n_bins | 1) +
                   (size_t)staged * (n_mel + 1) + (n_dct ? 0 : (size_t)2 * n_mel * (stacks + 1))) * sizeof(float);
    BRV_REQUIRE(smem <= 200 * 1024, "feature tile does not fit shared memory (%zu bytes)", smem);
    if (smem > 48 * 1024)
        BRV_CUDA(cudaFuncSetAttribute(fbe_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    fbe_features_kernel<<<(unsigned)grid, FT_THREADS, smem, (cudaStream_t)stream>>>(
        (const float2*)X, sb, sc, sf, st, n_channels, n_bins, n_frames, mel_vals, mel_cols,
        mel_rowptr, n_mel, nnz, normalize, compression, eps, stacks, decimation, mean, stdv, out,
        out_frames, tile, tiles, mode, dct_basis, n_dct);
    BRV_LAUNCH_CHECK("fbe_features_kernel");
    return BRV_OK;
}

extern "C" int brv_fbe_features(const void* X, int64_t sb, int64_t sc, int64_t sf, int64_t st,
                                int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                                const float* mel_vals, const int32_t* mel_cols,
                                const int32_t* mel_rowptr, int n_mel, int nnz, int normalize,
                                int compression, float eps, int stacks, int decimation,
                                const float* mean, const float* stdv, float* out, void* stream) {
    return launch_features(X, sb, sc, sf, st, n_batch, n_channels, n_bins, n_frames, mel_vals,
                           mel_cols, mel_rowptr, n_mel, nnz, normalize, compression, eps, stacks,
                           decimation, mean, stdv, out, stream, FT_MODE_POWER, nullptr, 0);
}

extern "C" int brv_mel_features(const void* X, int64_t sb, int64_t sc, int64_t sf, int64_t st,
                                int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                                int mode, const float* mel_vals, const int32_t* mel_cols,
                                const int32_t* mel_rowptr, int n_mel, int nnz, int normalize,
                                int compression, float eps, const float* dct_basis, int n_dct,
                                float* out, void* stream) {
    return launch_features(X, sb, sc, sf, st, n_batch, n_channels, n_bins, n_frames, mel_vals,
                           mel_cols, mel_rowptr, n_mel, nnz, normalize, compression, eps, 0, 1,
                           nullptr, nullptr, out, stream, mode, dct_basis, n_dct);
}

extern "C" int brv_ic_coherence(const void* X, int64_t sb, int64_t sc, int64_t sf, int64_t st,
                                int64_t n_batch, int n_channels, int n_bins, int64_t n_frames,
                                float b0, float a1, float* out, void* stream) {
    BRV_REQUIRE(X && out, "null pointer argument");
    BRV_REQUIRE(n_channels >= 2, "interaural coherence needs two channels, got %d", n_channels);
    BRV_REQUIRE(n_bins >= 1 && n_batch < 65536, "bad shape");
    if (n_batch == 0 || n_frames == 0) return BRV_OK;
    // b0 = float32(1 - alpha), a1 = float32(-alpha): the coefficients the reference hands to lfilter
    dim3 grid((unsigned)brv_ceil_div(n_bins, 128), (unsigned)n_batch);
    ic_coherence_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float2*)X, sb, sc, sf, st,
                                                               n_bins, n_frames, b0, a1, out);
    BRV_LAUNCH_CHECK("ic_coherence_kernel");
    return BRV_OK;
}

extern "C" int brv_mel_apply(const float* x, int64_t sb, int64_t sr, int64_t st, int64_t n_batch,
                             int n_rows_in, int64_t n_frames, const float* vals,
                             const int32_t* cols, const int32_t* rowptr, int n_rows_out,
                             float* out, void* stream) {
    BRV_REQUIRE(x && vals && cols && rowptr && out, "null pointer argument");
    BRV_REQUIRE(n_rows_in >= 1 && n_rows_out >= 1 && n_rows_out < 65536, "bad row counts");
    BRV_REQUIRE(n_batch < 65536, "more than 65535 batch items per call");
    if (n_batch == 0 || n_frames == 0) return BRV_OK;
    dim3 grid((unsigned)brv_ceil_div(n_frames, 128), (unsigned)n_rows_out, (unsigned)n_batch);
    mel_apply_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, sb, sr, st, n_frames, vals, cols,
                                                            rowptr, n_rows_out, out);
    BRV_LAUNCH_CHECK("mel_apply_kernel");
    return BRV_OK;
}

extern "C" int brv_stack_normalize(const float* x, int64_t n_batch, int n_features,
                                   int64_t n_frames, int stacks, int decimation,
                                   const float* mean, const float* stdv, float* out,
                                   void* stream) {
    BRV_REQUIRE(x && out, "null pointer argument");
    BRV_REQUIRE(n_features >= 1 && stacks >= 0 && decimation >= 1, "bad stacking parameters");
    const int64_t rows = (int64_t)n_features * (stacks + 1);
    BRV_REQUIRE(rows < 65536 && n_batch < 65536, "too many rows / batch items per call");
    if (n_batch == 0 || n_frames == 0) return BRV_OK;
    const int64_t out_frames = brv_ceil_div(n_frames, decimation);
    dim3 grid((unsigned)brv_ceil_div(out_frames, 128), (unsigned)rows, (unsigned)n_batch);
    stack_normalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        x, n_features, n_frames, stacks, decimation, mean, stdv, out, out_frames);
    BRV_LAUNCH_CHECK("stack_normalize_kernel");
    return BRV_OK;
}

extern "C" int brv_cumulative_normalize(const float* x, int64_t n_rows, int64_t n_frames,
                                        float eps, float* out, void* stream) {
    BRV_REQUIRE(x && out, "null pointer argument");
    if (n_rows == 0 || n_frames == 0) return BRV_OK;
    const int warps = 4;
    cumulative_normalize_kernel<<<(unsigned)brv_ceil_div(n_rows, warps), warps * 32, 0,
                                  (cudaStream_t)stream>>>(x, n_rows, n_frames, eps, out);
    BRV_LAUNCH_CHECK("cumulative_normalize_kernel");
    return BRV_OK;
}
