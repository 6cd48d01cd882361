// Symmetry-folded tensor-core STFT for sm_100a (tcgen05 + TMEM + TMA), fp32-grade.
//
// STFT.forward (brever/modules/stft.py:59-89) for one-sided transforms with
// n_fft in {128, 256, 384, 512} and {126, 254, 382, 510} (N = 4Q - 2: the same fold
// with N/2 odd — SGMSE's 510-point transform; see FoldFwdParams::odd).  ConvSTFT
// (stft.py:201-319) runs on the same kernels with another frame origin (end of file).
// A real DFT of length N = 4Q splits, by the even/odd symmetries of cos and sin about n = N/2 and n = N/4 (two radix-2
// decimation-in-frequency steps done on the *input* side), into four independent
// Q x Q contractions:
//
//   with xw[n] = w[n] x[n] / sqrt(sum w^2),  a = xw[n], b = xw[N/2-n],
//        c = xw[N/2+n], d = xw[N-n]   (0 < n < Q)
//   ee[n] = a+d+b+c   Re X[2m]   =  sum_n ee[n] cos(2 pi 2m n / N)     (+ (-1)^m ee[Q])
//   eo[n] = a+d-b-c   Re X[2m+1] =  sum_n eo[n] cos(2 pi (2m+1) n / N)
//   oe[n] = a-d-b+c   Im X[2m]   = -sum_n oe[n] sin(2 pi 2m n / N)
//   oo[n] = a-d+b-c   Im X[2m+1] = -sum_n oo[n] sin(2 pi (2m+1) n / N) (- (-1)^m oo[Q])
//
// n, m = 0..Q-1; the n = Q terms are rank-1 corrections applied in fp32 in the
// epilogue and the Nyquist bin Re X[N/2] = sum_n (-1)^n ee[n] is accumulated in
// fp32 by the operand builders.  This is 4x fewer tensor-core flops than the
// dense N x N DFT contraction (brv_stft_tc.cu), which moves the kernel from the
// tensor roofline to the HBM roofline (the spectrogram write).
//
// One CTA = up to 128 frames of one signal x all bins:
//   * the frames' sample span is staged once in shared memory (coalesced float4
//     loads; centre / right zero padding by predication — nothing padded or
//     framed is ever written to HBM);
//   * 8 builder warps window, fold, scale each row by its own power of two and
//     split it into fp16 hi/lo planes (22 significant bits) in the K-major
//     SWIZZLE_64B layout tcgen05 reads;
//   * the pure-trigonometric basis (split the same way once per plan) streams
//     in by TMA through a 2-stage mbarrier ring;
//   * three products per k-step (hi*hi + lo*hi + hi*lo) accumulate in fp32 in
//     four TMEM accumulators (4 x Q = N columns);
//   * epilogue: TMEM -> registers, undo the power-of-two scalings, rank-1
//     corrections, |X|^(c-1) compression, scale_factor, interleave
//     (Re, Im) of even / odd bins, per-warp shared-memory transpose, 256-byte
//     contiguous row stores into the frame-major complex64 output.
#include <type_traits>
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "brv_common.cuh"
#include "brv_tc_ptx.cuh"

// Dev-only phase timing (compile with -DBRV_PHASE_TIMING): builder thread 0 of CTA 0 stamps
// %globaltimer at the phase boundaries of its first tiles; read back with brv_debug_phase_times.
#ifdef BRV_PHASE_TIMING
__device__ unsigned long long g_brv_phase_ts[64];
#define BRV_STAMP(i)                                                                    \
    do {                                                                                \
        if (blockIdx.x == 0 && bt == 0 && (i) < 64) {                                   \
            unsigned long long t_;                                                      \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
            g_brv_phase_ts[(i)] = t_;                                                   \
        }                                                                               \
    } while (0)
extern "C" int brv_debug_phase_times(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_brv_phase_ts, sizeof(unsigned long long) * (n < 64 ? n : 64));
}
#else
#define BRV_STAMP(i) do {} while (0)
#endif
#ifdef BRV_PHASE_TIMING
extern "C" int brv_debug_t_times(unsigned long long* out, int n);
#endif

namespace {

using namespace brv_ptx;

constexpr int TILE_M = 128;                    // frames per CTA (UMMA M)
constexpr int BK = 32;                         // k per stage: 64-byte rows (SWIZZLE_64B)
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int SUB_TILE = TILE_M * BK * 2;      // 8 KB: one (sub-GEMM, plane) operand block
constexpr int STAGE_A = 4 * SUB_TILE;          // 2 sub-GEMMs x {hi, lo}
constexpr int STAGE_BYTES = 2 * STAGE_A;       // A + B: 64 KB
constexpr int SPAN_MAX = 16768;                // floats: (128 - 1) * 128 + 512
constexpr int SPAN_ALLOC = SPAN_MAX + 32;
constexpr int BUILDERS = 8;                    // builder / epilogue warps
constexpr int BUILDER_THREADS = BUILDERS * 32;
constexpr int NUM_THREADS = 64 + BUILDER_THREADS;
constexpr int MAX_Q = 128;
constexpr int EPI_PITCH = 68;                  // floats per staged epilogue row (64 + 4)

constexpr int SMEM_STAGES = STAGES * STAGE_BYTES;              // 131072
constexpr int SMEM_SPAN = SPAN_ALLOC * 4;                      // 67200
constexpr int SMEM_WTAB = MAX_Q * 16;                          // 2048
constexpr int SMEM_ROWINFO = TILE_M * 8;                       // 1024 (scale, nyquist)
constexpr int SMEM_BYTES = 1024 + SMEM_STAGES + SMEM_SPAN + SMEM_WTAB + SMEM_ROWINFO + 2112;   // + per-32-sample maxima

struct FoldFwdParams {
    const float* x;
    int64_t x_stride, samples;
    float* out;                  // (sig, T, n_bins) complex64 as floats
    int64_t n_frames;
    const float4* wtab;          // Q entries: (w[n], w[N/2-n], w[N/2+n], w[N-n]) / norm
    int n_fft, hop, n_bins, q;
    int rows;                    // frames per tile (<= 128, span fits SPAN_MAX)
    int tiles_per_signal;
    int64_t total_tiles;         // n_signals * tiles_per_signal (persistent kernel)
    int tmem_cols;
    float wq, w3q;               // w[Q] / norm, w[3Q] / norm
    float wmax;                  // max |w| / norm
    float basis_scale_inv;
    float post_scale, post_expo;
    // gradient of the inverse transform (brv_fold_istft_grad): the input is multiplied by
    // in_mul[sample] (1 / overlap-added w^2) while it is staged, and the DC and Nyquist bins,
    // which carry Hermitian weight 1 instead of 2, are multiplied by edge_scale = 1/2
    const float* in_mul;
    float edge_scale;            // Nyquist bin
    float dc_scale;              // DC bin (equal to edge_scale except for ConvSTFT's 1/sqrt(2) row)
    // frame t starts origin samples before sample t * hop: n_fft / 2 (torch.stft centre
    // padding) or frame_length - hop_length (ConvSTFT.pad, stft.py:311-313)
    int origin;
    // n_fft = 4Q - 2 (N/2 odd, e.g. 510): same four contractions over n = 0..Q-1, but there is
    // no self-paired n = Q column and no separate Nyquist bin (k = N/2 is the last odd bin)
    int odd;
    // the staged span starts `shift` samples early so that it is 16-byte aligned in HBM
    int shift;
    // transposed strip kernel (brv_fold_t.cuh): total_tiles counts frames; grad_env = multiply the
    // input by 1 / overlap-added w^2 (periodic table + edges summed from the squared window)
    // instead of by an in_mul table
    int grad_env;
    const float* env_per;
    const float* wsq;
};

// 1 / (overlap-added squared window) at offset `off` of hop block `u` (torch.istft divides by
// this envelope; NOLA was checked by the caller).  Interior blocks see all ceil(N / H) frames:
// periodic table; edge blocks are summed from the squared window.  No per-frame-count table.
__device__ __forceinline__ float ola_inv_envelope(const float* env_per, const float* wsq, int N,
                                                  int H, int64_t n_frames, int64_t u, int off) {
    const int rmax = (N + H - 1) / H;
    if (u >= rmax - 1 && u <= n_frames - 1) return __ldg(env_per + off);
    float acc = 0.f;
    for (int j = 0; j < rmax; ++j) {
        const int64_t t = u - j;
        const int pos = j * H + off;
        if (t >= 0 && t <= n_frames - 1 && pos < N) acc += __ldg(wsq + pos);
    }
    return 1.f / acc;
}

__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
// table[idx .. idx+3] with 1 outside [0, n) (positions there hold zero-padding anyway)
__device__ __forceinline__ float4 load4_clamped(const float* table, int64_t idx, int64_t n) {
    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    if (idx >= 0 && idx < n) v.x = __ldg(table + idx);
    if (idx + 1 >= 0 && idx + 1 < n) v.y = __ldg(table + idx + 1);
    if (idx + 2 >= 0 && idx + 2 < n) v.z = __ldg(table + idx + 2);
    if (idx + 3 >= 0 && idx + 3 < n) v.w = __ldg(table + idx + 3);
    return v;
}
__device__ __forceinline__ void split_store(uint8_t* hi_ptr, uint8_t* lo_ptr, float v0, float v1) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__half2*>(hi_ptr) = h;
    *reinterpret_cast<__half2*>(lo_ptr) = l;
}

// COMPRESS: magnitude compression in the epilogue (32 inlined power evaluations: a third of the
// kernel's code, kept out of the plain instantiation -- instruction fetch is a top stall here)
template <bool COMPRESS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
stft_fold_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // aligned by pointer arithmetic (not an integer round trip): the compiler keeps the
    // shared address space and emits LDS / STS instead of generic loads and stores
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* span = reinterpret_cast<float*>(stages + SMEM_STAGES);
    float4* wtab = reinterpret_cast<float4*>(stages + SMEM_STAGES + SMEM_SPAN);
    float2* rowinfo = reinterpret_cast<float2*>(stages + SMEM_STAGES + SMEM_SPAN + SMEM_WTAB);
    uint32_t* bmax = reinterpret_cast<uint32_t*>(stages + SMEM_STAGES + SMEM_SPAN + SMEM_WTAB +
                                                 SMEM_ROWINFO);

    const int64_t sig = blockIdx.x / p.tiles_per_signal;
    const int64_t t0 = (int64_t)(blockIdx.x % p.tiles_per_signal) * p.rows;
    const int rows_eff = (int)min((int64_t)p.rows, p.n_frames - t0);
    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_it = 2 * (Q / BK);             // pipeline iterations: (k-chunk, sub-GEMM pair)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + BUILDERS);      // TMA arrive + one per builder warp
            mbar_init(&empty_bar[s], 1);                // one tcgen05.commit
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer: basis k-chunks =====================
        if (elect_one()) {
            for (int it = 0; it < n_it; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int kc = it >> 1, pair = it & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)Q * BK * 2);
                uint8_t* sb = stages + (size_t)s * STAGE_BYTES + STAGE_A;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
                        tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE), &basis_map,
                                    &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(TILE_M, Q);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int kc = it >> 1, pair = it & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(stages + (size_t)s * STAGE_BYTES);
                const uint32_t b0 = a0 + STAGE_A;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t d = tmem_base + (uint32_t)((pair * 2 + j) * Q);
#pragma unroll
                    for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                        const uint32_t off = ks * UMMA_K * 2;
                        const uint64_t dah = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                        const uint64_t dal = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                        const uint64_t dbh = umma_desc_sw64(b0 + (j * 2) * SUB_TILE + off);
                        const uint64_t dbl = umma_desc_sw64(b0 + (j * 2 + 1) * SUB_TILE + off);
                        umma_f16(d, dah, dbh, idesc, (kc | ks) != 0);
                        umma_f16(d, dal, dbh, idesc, 1);
                        umma_f16(d, dah, dbl, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&accum_bar);
        }
    } else {
        // ===================== builders, then epilogue ==========================
        const int bw = warp - 2;                   // 0..7
        const int bt = bw * 32 + lane;             // 0..255
        const float* xs = p.x + sig * p.x_stride;
        const int shift = p.shift;
        const int64_t span0 = t0 * H - p.origin - shift; // first sample of the span (may be < 0)
        const int span_len = (rows_eff - 1) * H + N + shift;
        const int span_pad = (span_len + 31) & ~31;

        BRV_STAMP(32);
        for (int j = bt; j < Q; j += BUILDER_THREADS) wtab[j] = __ldg(p.wtab + j);

        // ---- stage the sample span + per-32-sample maxima ------------------------
        // SPAN_UNROLL independent 16-byte loads in flight per thread (the span is a
        // first touch: HBM latency, not bandwidth, would bound a one-at-a-time loop)
        const bool vec = ((((uintptr_t)xs) & 15) == 0) && ((span0 & 3) == 0);
        constexpr int SPAN_UNROLL = 6;
        for (int wb0 = bw * 128; wb0 < span_pad; wb0 += SPAN_UNROLL * BUILDERS * 128) {
            float4 f[SPAN_UNROLL];
#pragma unroll
            for (int u = 0; u < SPAN_UNROLL; ++u) {
                const int i = wb0 + u * BUILDERS * 128 + lane * 4;
                const int64_t idx = span0 + i;
                f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < span_pad) {
                    if (vec && idx >= 0 && idx + 4 <= p.samples) {
                        f[u] = __ldg(reinterpret_cast<const float4*>(xs + idx));
                        if (p.in_mul)
                            f[u] = mul4(f[u], __ldg(reinterpret_cast<const float4*>(p.in_mul + idx)));
                    } else {
                        if (idx >= 0 && idx < p.samples) f[u].x = __ldg(xs + idx);
                        if (idx + 1 >= 0 && idx + 1 < p.samples) f[u].y = __ldg(xs + idx + 1);
                        if (idx + 2 >= 0 && idx + 2 < p.samples) f[u].z = __ldg(xs + idx + 2);
                        if (idx + 3 >= 0 && idx + 3 < p.samples) f[u].w = __ldg(xs + idx + 3);
                        if (p.in_mul) f[u] = mul4(f[u], load4_clamped(p.in_mul, idx, p.samples));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < SPAN_UNROLL; ++u) {
                const int i = wb0 + u * BUILDERS * 128 + lane * 4;
                if (i < span_pad) *reinterpret_cast<float4*>(span + i) = f[u];
                float m = fmaxf(fmaxf(finite_abs(f[u].x), finite_abs(f[u].y)),
                                fmaxf(finite_abs(f[u].z), finite_abs(f[u].w)));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                if ((lane & 7) == 0 && i < span_pad) bmax[i >> 5] = __float_as_uint(m);
            }
        }
        named_bar_sync(1, BUILDER_THREADS);
        BRV_STAMP(33);

        // ---- per-row power-of-two scale from a bound on the folded magnitudes ----
        if (bt < rows_eff) {
            const int b0 = (bt * H + shift) >> 5, b1 = (bt * H + shift + N - 1) >> 5;
            uint32_t mx = 0u;
            for (int b = b0; b <= b1; ++b) mx = max(mx, bmax[b]);
            rowinfo[bt].x = row_scale(4.f * p.wmax * __uint_as_float(mx));
        }
        named_bar_sync(1, BUILDER_THREADS);
        BRV_STAMP(34);

        // ---- main loop: window, fold, scale, split, store ------------------------
        const int half = lane >> 4;                // which of the warp's two rows per pass
        const int pr = lane & 15;                  // n pair inside the 32-wide k-chunk
        const uint32_t chunk = (uint32_t)(pr >> 2);
        float nyq[8];
        float rscale[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            nyq[i] = 0.f;
            const int row = bw * 16 + 2 * i + half;
            rscale[i] = row < rows_eff ? rowinfo[row].x : 0.f;
        }
        for (int kc = 0; kc < Q / BK; ++kc) {
            const int n0 = kc * BK + 2 * pr;
            const float4 w0 = wtab[n0], w1 = wtab[n0 + 1];
            // window and first fold once per k-chunk, before waiting for a stage: both sub-GEMM
            // pairs read the same four samples per (row, n), and the loads run under the wait
            float sp0[8], sp1[8], rp0[8], rp1[8], sm0[8], sm1[8], rm0[8], rm1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = bw * 16 + 2 * i + half;
                // rows >= rows_eff are built too (scale 0, accumulator rows never read): a branch
                // here would split the unrolled rows into basic blocks and serialise their chains
                const float* fr = span + shift + (row < rows_eff ? row : 0) * H;   // stay inside the span
                const float a0 = fr[n0] * w0.x, a1 = fr[n0 + 1] * w1.x;
                const float b0 = fr[Hf - n0] * w0.y, b1 = fr[Hf - n0 - 1] * w1.y;
                const float c0 = fr[Hf + n0] * w0.z, c1 = fr[Hf + n0 + 1] * w1.z;
                const float d0 = n0 ? fr[N - n0] * w0.w : 0.f, d1 = fr[N - n0 - 1] * w1.w;
                sp0[i] = a0 + d0; sp1[i] = a1 + d1; rp0[i] = b0 + c0; rp1[i] = b1 + c1;
                sm0[i] = a0 - d0; sm1[i] = a1 - d1; rm0[i] = b0 - c0; rm1[i] = b1 - c1;
            }
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
                const int it = kc * 2 + pair;
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = stages + (size_t)s * STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = bw * 16 + 2 * i + half;
                    float u0, u1, v0, v1;          // the pair's two folded sequences at n0, n0+1
                    if (pair == 0) {
                        u0 = sp0[i] + rp0[i]; u1 = sp1[i] + rp1[i];    // ee
                        v0 = sp0[i] - rp0[i]; v1 = sp1[i] - rp1[i];    // eo
                        nyq[i] += u0 - u1;                             // (-1)^n ee[n], n0 even
                    } else {
                        u0 = sm0[i] - rm0[i]; u1 = sm1[i] - rm1[i];    // oe
                        v0 = sm0[i] + rm0[i]; v1 = sm1[i] + rm1[i];    // oo
                    }
                    const float sc = rscale[i];
                    uint8_t* dst = sa + row * (BK * 2) +
                                   ((chunk ^ (uint32_t)((row >> 1) & 3)) << 4) + (pr & 3) * 4;
                    split_store(dst, dst + SUB_TILE, u0 * sc, u1 * sc);
                    split_store(dst + 2 * SUB_TILE, dst + 3 * SUB_TILE, v0 * sc, v1 * sc);
                }
                fence_proxy_async();                           // generic -> async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = nyq[i];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            const int row = bw * 16 + 2 * i + half;
            if (pr == 0 && row < rows_eff) rowinfo[row].y = v;
        }
        named_bar_sync(1, BUILDER_THREADS);
        BRV_STAMP(35);

        // ---- epilogue ---------------------------------------------------------------
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int hsel = bw >> 2;                  // which half of the m range
        const int row = q * 32 + lane;
        const bool live = row < rows_eff;
        float g0 = 0.f, eeq = 0.f, ooq = 0.f, ny = 0.f;
        if (live) {
            const float2 ri = rowinfo[row];
            g0 = p.basis_scale_inv / ri.x;
            const float xq = span[shift + row * H + Q] * p.wq, x3q = span[shift + row * H + 3 * Q] * p.w3q;
            eeq = xq + x3q;
            ooq = xq - x3q;
            ny = ri.y + eeq;                       // Q is even: (-1)^Q = +1
        }
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        BRV_STAMP(36);
        // every MMA has retired: the pipeline stages are free, reuse them as staging
        float* stg = reinterpret_cast<float*>(stages) + (size_t)bw * 32 * EPI_PITCH;
        const int pitch = 2 * p.n_bins;
        float* obase = p.out + (sig * p.n_frames + t0 + q * 32) * (int64_t)pitch;
        const int rows_w = min(32, rows_eff - q * 32);      // rows of this warp that exist
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < Q / 32; ++c) {
            const int m0 = hsel * (Q / 2) + 16 * c;
            uint32_t r0[16], r1[16], r2[16], r3[16];
            tmem_ld16_nowait(tq + (uint32_t)(m0), r0);
            tmem_ld16_nowait(tq + (uint32_t)(Q + m0), r1);
            tmem_ld16_nowait(tq + (uint32_t)(2 * Q + m0), r2);
            tmem_ld16_nowait(tq + (uint32_t)(3 * Q + m0), r3);
            tmem_ld_wait();
            // three straight-line passes over the registers (a uniform `if (compress)` inside one
            // loop would cut the 16 independent columns into basic blocks and serialise them)
            float* f0 = reinterpret_cast<float*>(r0);      // Re X[2m]
            float* f1 = reinterpret_cast<float*>(r1);      // Re X[2m+1]
            float* f2 = reinterpret_cast<float*>(r2);      // Im X[2m]
            float* f3 = reinterpret_cast<float*>(r3);      // Im X[2m+1]
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float sg = (j & 1) ? -1.f : 1.f;        // m0 is even
                f0[j] = f0[j] * g0 + sg * eeq;
                if (j == 0 && m0 == 0) f0[j] *= p.dc_scale;        // DC bin (its Im is exactly 0)
                f1[j] = f1[j] * g0;
                // n_fft = 4Q - 2: the Nyquist bin is the last odd bin (Hermitian weight 1 in the gradient)
                if (j == 15 && p.odd && m0 + 16 == p.q) f1[j] *= p.edge_scale;
                f2[j] = f2[j] * g0;
                f3[j] = f3[j] * g0 - sg * ooq;
            }
            if (COMPRESS) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    compress(f0[j], f2[j], p.post_expo);
                    compress(f1[j], f3[j], p.post_expo);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
                *reinterpret_cast<float4*>(stg + lane * EPI_PITCH + 4 * j) =
                    make_float4(f0[j] * p.post_scale, f2[j] * p.post_scale, f1[j] * p.post_scale,
                                f3[j] * p.post_scale);
            __syncwarp();
            float* ocol = obase + 4 * m0 + 2 * lane;
            const float* srow = stg + 2 * lane;
#pragma unroll 8
            for (int rr = 0; rr < rows_w; ++rr) {
                *reinterpret_cast<float2*>(ocol) = *reinterpret_cast<const float2*>(srow);
                ocol += pitch;
                srow += EPI_PITCH;
            }
            __syncwarp();
        }
        if (hsel == 1 && live && !p.odd) {         // Nyquist bin: purely real
            float v = ny * p.edge_scale;
            if (COMPRESS) v = compress_real(v, p.post_expo);
            *reinterpret_cast<float2*>(obase + (int64_t)lane * pitch + 2 * Hf) =
                make_float2(v * p.post_scale, 0.f);
        }
        BRV_STAMP(37);
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}


__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory");
}

// =====================================================================================
// Pipelined persistent forward (default).  Same arithmetic as stft_fold_kernel, but the
// four sub-GEMMs are grouped by bin parity into two passes with their own halves of
// TMEM, so that every phase overlaps another one:
//
//   pass A  even bins  ee -> Re X[2m]    oe -> Im X[2m]      TMEM columns [0, 2Q)
//   pass B  odd bins   eo -> Re X[2m+1]  oo -> Im X[2m+1]    TMEM columns [2Q, 4Q)
//
//   warp 0       TMA producer: basis k-chunks of the current pass (2-stage ring)
//   warp 1       TMEM owner + MMA issuer; waits for the epilogue to hand a half back
//   warps 4-7    epilogue: drain one half (TMEM -> registers -> per-warp transpose ->
//                complex64 stores of that parity's bins) while the tensor core and the
//                builders work on the other half / the next tile
//   warps 8-15   builders: stage the tile's sample span (the next tile's span is
//                prefetched into L2 a tile ahead), window, fold, scale, split
//
// One CTA per SM walks the tile list, so barrier setup, TMEM allocation and the tensor
// map fetch are paid once per SM and a tile's stores drain under the next tile's loads.
constexpr int F2_BUILDERS = 16;                                   // builder warps
constexpr int F2_BUILDER_THREADS = F2_BUILDERS * 32;
constexpr int F2_ROWS_PER_WARP = TILE_M / F2_BUILDERS;            // 8
constexpr int F2_THREADS = 256 + F2_BUILDER_THREADS;
constexpr int F2_EPI_WARP0 = 4;
constexpr int F2_EPI_WARPS = 4;
constexpr int F2_BUILD_WARP0 = 8;
constexpr int F2_EPI_PITCH = 34;                                   // floats: 16 complex + pad
constexpr int F2_OFF_SPAN = SMEM_STAGES;
constexpr int F2_OFF_WTAB = F2_OFF_SPAN + SMEM_SPAN;
constexpr int F2_OFF_ROWINFO = F2_OFF_WTAB + SMEM_WTAB;             // 2 slots x 128 float4
constexpr int F2_OFF_BMAX = F2_OFF_ROWINFO + 2 * TILE_M * 16;
constexpr int F2_OFF_EPI = F2_OFF_BMAX + 2112;
constexpr int F2_SMEM_EPI = F2_EPI_WARPS * 32 * F2_EPI_PITCH * 4;
constexpr int F2_SMEM_BYTES = 1024 + F2_OFF_EPI + F2_SMEM_EPI;
static_assert(F2_SMEM_BYTES <= 227 * 1024, "forward kernel shared memory");

template <bool COMPRESS>
__global__ void __launch_bounds__(F2_THREADS, 1)
stft_fold2_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];     // MMA -> epilogue, per pass
    __shared__ __align__(8) uint64_t tmem_empty[2];    // epilogue -> MMA, per pass
    __shared__ __align__(8) uint64_t ri_full[2];       // builders -> epilogue (row info slot)
    __shared__ __align__(8) uint64_t ri_empty[2];      // epilogue -> builders
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* span = reinterpret_cast<float*>(stages + F2_OFF_SPAN);
    float4* wtab = reinterpret_cast<float4*>(stages + F2_OFF_WTAB);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + F2_OFF_ROWINFO);
    uint32_t* bmax = reinterpret_cast<uint32_t*>(stages + F2_OFF_BMAX);
    float* epi = reinterpret_cast<float*>(stages + F2_OFF_EPI);

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_kc = Q / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + F2_BUILDERS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], F2_EPI_WARPS);
            mbar_init(&ri_full[b], F2_BUILDERS);
            mbar_init(&ri_empty[b], F2_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    for (int j = threadIdx.x; j < Q; j += F2_THREADS) wtab[j] = __ldg(p.wtab + j);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer ====================================
        if (elect_one()) {
            int g = 0;
            for (int64_t tile_id = blockIdx.x; tile_id < p.total_tiles; tile_id += gridDim.x)
                for (int pass = 0; pass < 2; ++pass)
                    for (int kc = 0; kc < n_kc; ++kc, ++g) {
                        const int s = g % STAGES;
                        const uint32_t ph = (g / STAGES) & 1;
                        mbar_wait_relaxed(&empty_bar[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)Q * BK * 2);
                        uint8_t* sb = stages + (size_t)s * STAGE_BYTES + STAGE_A;
#pragma unroll
                        for (int j = 0; j < 2; ++j)      // j = 0: cos-type sub, 1: sin-type
#pragma unroll
                            for (int pl = 0; pl < 2; ++pl)
                                tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE), &basis_map,
                                            &full_bar[s], kc * BK, (pl * 4 + 2 * j + pass) * Q);
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(TILE_M, Q);
            int g = 0, n = 0;
            for (int64_t tile_id = blockIdx.x; tile_id < p.total_tiles; tile_id += gridDim.x, ++n)
                for (int pass = 0; pass < 2; ++pass) {
                    mbar_wait_relaxed(&tmem_empty[pass], (uint32_t)((n & 1) ^ 1));
                    tcgen05_fence_after();
                    for (int kc = 0; kc < n_kc; ++kc, ++g) {
                        const int s = g % STAGES;
                        const uint32_t ph = (g / STAGES) & 1;
                        mbar_wait_relaxed(&full_bar[s], ph, 32);
                        tcgen05_fence_after();
                        const uint32_t a0 = smem_u32(stages + (size_t)s * STAGE_BYTES);
                        const uint32_t b0 = a0 + STAGE_A;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t d = tmem_base + (uint32_t)((pass * 2 + j) * Q);
#pragma unroll
                            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                                const uint32_t off = ks * UMMA_K * 2;
                                const uint64_t dah = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                                const uint64_t dal = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                                const uint64_t dbh = umma_desc_sw64(b0 + (j * 2) * SUB_TILE + off);
                                const uint64_t dbl = umma_desc_sw64(b0 + (j * 2 + 1) * SUB_TILE + off);
                                umma_f16(d, dah, dbh, idesc, (kc | ks) != 0);
                                umma_f16(d, dal, dbh, idesc, 1);
                                umma_f16(d, dah, dbl, idesc, 1);
                            }
                        }
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&tmem_full[pass]);
                }
        }
    } else if (warp >= F2_EPI_WARP0 && warp < F2_EPI_WARP0 + F2_EPI_WARPS) {
        // ===================== epilogue ========================================
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        float* stg = epi + (size_t)q * 32 * F2_EPI_PITCH;
        const int pitch = 2 * p.n_bins;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        const int sub_row = lane >> 4, cl = lane & 15;
        int n = 0;
        for (int64_t tile_id = blockIdx.x; tile_id < p.total_tiles; tile_id += gridDim.x, ++n) {
            const int slot = n & 1;
            const int64_t sig = tile_id / p.tiles_per_signal;
            const int64_t t0 = (int64_t)(tile_id % p.tiles_per_signal) * p.rows;
            const int rows_eff = (int)min((int64_t)p.rows, p.n_frames - t0);
            const int rows_w = min(32, rows_eff - q * 32);
            const bool live = row < rows_eff;
            mbar_wait_relaxed(&ri_full[slot], (uint32_t)((n >> 1) & 1));
            const float4 ri = rowinfo2[slot * TILE_M + row];   // scale, nyquist sum, ee[Q], oo[Q]
            __syncwarp();
            if (lane == 0) mbar_arrive(&ri_empty[slot]);
            // without compression scale_factor folds into the per-row factors
            const float ps = COMPRESS ? 1.f : p.post_scale;
            const float g0 = live ? ps * p.basis_scale_inv / ri.x : 0.f;
            const float eeq = live ? ps * ri.z : 0.f, ooq = live ? ps * ri.w : 0.f;
            float* obase = p.out + (sig * p.n_frames + t0 + q * 32) * (int64_t)pitch;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                mbar_wait_relaxed(&tmem_full[pass], (uint32_t)(n & 1));
                tcgen05_fence_after();
                const float rq = pass == 0 ? eeq : 0.f;    // rank-1 term of Re (even bins)
                const float iq = pass == 1 ? -ooq : 0.f;   // rank-1 term of Im (odd bins)
                const uint32_t tp = tq + (uint32_t)(pass * 2 * Q);
#pragma unroll 1
                for (int c = 0; c < Q / 16; ++c) {
                    const int m0 = 16 * c;
                    uint32_t r0[16], r1[16];
                    tmem_ld16_nowait(tp + (uint32_t)m0, r0);
                    tmem_ld16_nowait(tp + (uint32_t)(Q + m0), r1);
                    tmem_ld_wait();
                    if (c == Q / 16 - 1) {                 // last read of this half: hand it back
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[pass]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float re = fmaf(__uint_as_float(r0[j]), g0, (j & 1) ? -rq : rq);   // m0 is even
                        float im = fmaf(__uint_as_float(r1[j]), g0, (j & 1) ? -iq : iq);
                        if (j == 0 && c == 0 && pass == 0) re *= p.dc_scale;     // DC bin
                        if (j == 15 && pass == 1 && p.odd && c == Q / 16 - 1) re *= p.edge_scale;   // Nyquist (4Q - 2)
                        if (COMPRESS) {
                            compress(re, im, p.post_expo);
                            re *= p.post_scale;
                            im *= p.post_scale;
                        }
                        *reinterpret_cast<float2*>(stg + lane * F2_EPI_PITCH + 2 * j) =
                            make_float2(re, im);
                    }
                    __syncwarp();
                    // bins 2 (m0 + cl) + pass of rows sub_row, sub_row + 2, ...
                    float* ocol = obase + 2 * (2 * (m0 + cl) + pass) + (int64_t)sub_row * pitch;
                    const float* srow = stg + sub_row * F2_EPI_PITCH + 2 * cl;
#pragma unroll 4
                    for (int rr = sub_row; rr < rows_w; rr += 2) {
                        *reinterpret_cast<float2*>(ocol) = *reinterpret_cast<const float2*>(srow);
                        ocol += 2 * (int64_t)pitch;
                        srow += 2 * F2_EPI_PITCH;
                    }
                    __syncwarp();
                }
                if (pass == 0 && live && !p.odd) {         // Nyquist bin: purely real
                    float v = (ri.y + ri.z) * p.edge_scale;   // Q is even: (-1)^Q = +1
                    if (COMPRESS) v = compress_real(v, p.post_expo);
                    *reinterpret_cast<float2*>(obase + (int64_t)lane * pitch + 2 * Hf) =
                        make_float2(v * p.post_scale, 0.f);
                }
            }
        }
    } else if (warp >= F2_BUILD_WARP0) {
        // ===================== builders ========================================
        const int bw = warp - F2_BUILD_WARP0;      // 0..15
        const int bt = bw * 32 + lane;             // 0..511
        const int half = lane >> 4;                // which of the warp's two rows per pass
        const int pr = lane & 15;                  // n pair inside the 32-wide k-chunk
        const uint32_t chunk = (uint32_t)(pr >> 2);
        int g = 0, n = 0;
        for (int64_t tile_id = blockIdx.x; tile_id < p.total_tiles; tile_id += gridDim.x, ++n) {
            const int slot = n & 1;
            float4* rowinfo = rowinfo2 + slot * TILE_M;
            const int64_t sig = tile_id / p.tiles_per_signal;
            const int64_t t0 = (int64_t)(tile_id % p.tiles_per_signal) * p.rows;
            const int rows_eff = (int)min((int64_t)p.rows, p.n_frames - t0);
            const float* xs = p.x + sig * p.x_stride;
            const int shift = p.shift;
            const int64_t span0 = t0 * H - p.origin - shift; // first sample of the span (may be < 0)
            const int span_len = (rows_eff - 1) * H + N + shift;
            const int span_pad = (span_len + 31) & ~31;
            if (n > 0) named_bar_sync(1, F2_BUILDER_THREADS);   // previous tile's span fully read

            // ---- stage the sample span + per-32-sample maxima ------------------------
            // aligned rows: every thread fires all of its 16-byte cp.async copies at once
            // (zero-filled outside [0, samples)), so one memory latency covers the whole span
            const bool vec = ((((uintptr_t)xs) & 15) == 0) && ((span0 & 3) == 0);
            if (vec) {
                for (int i = bt * 4; i < span_pad; i += F2_BUILDER_THREADS * 4) {
                    const int64_t idx = span0 + i;
                    const int64_t nb = (p.samples - idx) * 4;           // valid bytes from idx on
                    const uint32_t src_bytes = idx < 0 ? 0u : (uint32_t)(nb < 0 ? 0 : (nb > 16 ? 16 : nb));
                    const float* src = src_bytes ? xs + idx : xs;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(span + i)),
                                 "l"(src), "r"(src_bytes)
                                 : "memory");
                }
                cp_async_commit();
            } else {
                for (int i = bt * 4; i < span_pad; i += F2_BUILDER_THREADS * 4) {
                    const int64_t idx = span0 + i;
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx >= 0 && idx < p.samples) f.x = __ldg(xs + idx);
                    if (idx + 1 >= 0 && idx + 1 < p.samples) f.y = __ldg(xs + idx + 1);
                    if (idx + 2 >= 0 && idx + 2 < p.samples) f.z = __ldg(xs + idx + 2);
                    if (idx + 3 >= 0 && idx + 3 < p.samples) f.w = __ldg(xs + idx + 3);
                    *reinterpret_cast<float4*>(span + i) = f;
                }
            }
            // ---- pull the NEXT tile's span into L2 while this tile is built -----------
            {
                const int64_t nt = tile_id + gridDim.x;
                if (nt < p.total_tiles) {
                    const int64_t nsig = nt / p.tiles_per_signal;
                    const int64_t nt0 = (int64_t)(nt % p.tiles_per_signal) * p.rows;
                    const int nrows = (int)min((int64_t)p.rows, p.n_frames - nt0);
                    int64_t lo = nt0 * H - p.origin, hi = lo + (int64_t)(nrows - 1) * H + N;
                    if (lo < 0) lo = 0;
                    if (hi > p.samples) hi = p.samples;
                    const float* nx = p.x + nsig * p.x_stride;
                    for (int64_t i = lo + (int64_t)bt * 32; i < hi; i += F2_BUILDER_THREADS * 32)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
                }
            }
            cp_async_wait<0>();
            for (int i0 = bw * 128; i0 < span_pad; i0 += F2_BUILDER_THREADS * 4) {   // own copies only
                const int i = i0 + lane * 4;
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < span_pad) {
                    f = *reinterpret_cast<const float4*>(span + i);
                    if (p.in_mul) {
                        const int64_t idx = span0 + i;
                        if (vec && idx >= 0 && idx + 4 <= p.samples)
                            f = mul4(f, __ldg(reinterpret_cast<const float4*>(p.in_mul + idx)));
                        else
                            f = mul4(f, load4_clamped(p.in_mul, idx, p.samples));
                        *reinterpret_cast<float4*>(span + i) = f;
                    }
                }
                // fmaxf drops NaNs; an inf is removed by the slow path
                float m = fmaxf(fmaxf(fabsf(f.x), fabsf(f.y)), fmaxf(fabsf(f.z), fabsf(f.w)));
                if (!(m <= 3.0e38f))
                    m = fmaxf(fmaxf(finite_abs(f.x), finite_abs(f.y)),
                              fmaxf(finite_abs(f.z), finite_abs(f.w)));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                if ((lane & 7) == 0 && i < span_pad) bmax[i >> 5] = __float_as_uint(m);
            }
            named_bar_sync(1, F2_BUILDER_THREADS);

            // ---- per-row scale, rank-1 terms ------------------------------------------
            mbar_wait(&ri_empty[slot], (uint32_t)(((n >> 1) & 1) ^ 1));
            if (bt < TILE_M) {
                float4 ri = make_float4(1.f, 0.f, 0.f, 0.f);
                if (bt < rows_eff) {
                    const int b0 = (bt * H + shift) >> 5, b1 = (bt * H + shift + N - 1) >> 5;
                    uint32_t mx = 0u;
                    for (int b = b0; b <= b1; ++b) mx = max(mx, bmax[b]);
                    ri.x = row_scale(4.f * p.wmax * __uint_as_float(mx));
                    const float xq = span[shift + bt * H + Q] * p.wq,
                                x3q = span[shift + bt * H + 3 * Q] * p.w3q;
                    ri.z = xq + x3q;                   // ee[Q]
                    ri.w = xq - x3q;                   // oo[Q]
                }
                rowinfo[bt] = ri;
            }
            named_bar_sync(1, F2_BUILDER_THREADS);

            constexpr int RI = F2_ROWS_PER_WARP / 2;      // row pairs per warp
            float rscale[RI];
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                const int row = bw * F2_ROWS_PER_WARP + 2 * i + half;
                rscale[i] = row < rows_eff ? rowinfo[row].x : 0.f;
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const float psign = pass == 0 ? 1.f : -1.f, pnyq = pass == 0 ? 1.f : 0.f;
                float nyq[RI];
#pragma unroll
                for (int i = 0; i < RI; ++i) nyq[i] = 0.f;
                for (int kc = 0; kc < n_kc; ++kc, ++g) {
                    const int n0 = kc * BK + 2 * pr;
                    const float4 w0 = wtab[n0], w1 = wtab[n0 + 1];
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = stages + (size_t)s * STAGE_BYTES;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const int row = bw * F2_ROWS_PER_WARP + 2 * i + half;
                        // rows >= rows_eff are built too (their scale is 0 / their inputs are zero or stale,
                        // their accumulator rows are never read): a `continue` here would split the unrolled rows
                        // into separate basic blocks and serialise their dependent chains
                        const float* fr = span + shift + (row < rows_eff ? row : 0) * H;   // stay inside the span
                        const float a0 = fr[n0] * w0.x, a1 = fr[n0 + 1] * w1.x;
                        const float b0 = fr[Hf - n0] * w0.y, b1 = fr[Hf - n0 - 1] * w1.y;
                        const float c0 = fr[Hf + n0] * w0.z, c1 = fr[Hf + n0 + 1] * w1.z;
                        const float d0 = n0 ? fr[N - n0] * w0.w : 0.f, d1 = fr[N - n0 - 1] * w1.w;
                        const float s0 = a0 + d0, r0 = b0 + c0, s1 = a1 + d1, r1 = b1 + c1;
                        const float sd0 = a0 - d0, rd0 = b0 - c0, sd1 = a1 - d1, rd1 = b1 - c1;
                        // pass 0: ee = s + r, oe = sd - rd; pass 1: eo = s - r, oo = sd + rd -- by
                        // sign, not by branch (a branch would serialise the unrolled rows)
                        const float u0 = fmaf(psign, r0, s0), u1 = fmaf(psign, r1, s1);
                        const float v0 = fmaf(-psign, rd0, sd0), v1 = fmaf(-psign, rd1, sd1);
                        nyq[i] = fmaf(pnyq, u0 - u1, nyq[i]);      // (-1)^n ee[n], n0 even (pass 0 only)
                        const float sc = rscale[i];
                        uint8_t* dst = sa + row * (BK * 2) +
                                       ((chunk ^ (uint32_t)((row >> 1) & 3)) << 4) + (pr & 3) * 4;
                        split_store(dst, dst + SUB_TILE, u0 * sc, u1 * sc);
                        split_store(dst + 2 * SUB_TILE, dst + 3 * SUB_TILE, v0 * sc, v1 * sc);
                    }
                    if (pass == 0 && kc == n_kc - 1) {
                        // Nyquist sums must be visible before the epilogue can be released
#pragma unroll
                        for (int i = 0; i < RI; ++i) {
                            float v = nyq[i];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            const int row = bw * F2_ROWS_PER_WARP + 2 * i + half;
                            if (pr == 0 && row < rows_eff) rowinfo[row].y = v;
                        }
                    }
                    fence_proxy_async();                           // generic -> async proxy
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&full_bar[s]);
                        if (pass == 0 && kc == n_kc - 1) mbar_arrive(&ri_full[slot]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// =====================================================================================
// Folded inverse: STFT.backward (brever/modules/stft.py:101-138) = per-frame real
// inverse DFT, window, overlap-add, / overlap-added w^2, centre trim — one persistent
// kernel, no frames workspace in HBM.
//
// The transpose of the forward fold: with k = 2m / 2m+1 and n in [0, Q)
//   Ce[n] = sum_m c_2m Re X[2m]   cos(2 pi 2m n / N)       (c_0 = 1, else 2)
//   Co[n] = sum_m 2    Re X[2m+1] cos(2 pi (2m+1) n / N)
//   Se[n] = sum_m 2    Im X[2m]   sin(2 pi 2m n / N)        (Im X[0] never contributes)
//   So[n] = sum_m 2    Im X[2m+1] sin(2 pi (2m+1) n / N)
//   f[n]      = Ce + Co - Se - So + (-1)^n Re X[N/2]
//   f[N/2-n]  = Ce - Co + Se - So + (-1)^n Re X[N/2]
//   f[N/2+n]  = Ce - Co - Se + So + (-1)^n Re X[N/2]
//   f[N-n]    = Ce + Co + Se + So + (-1)^n Re X[N/2]
// f[Q] and f[3Q] (the n = Q column) are accumulated in fp32 by the operand builders.
// The frame is then multiplied by w[n] * sqrt(sum w^2) / N (window table).
//
// Overlap-add for hop = Q, 2Q or 4Q: in the epilogue a thread owns one frame (its
// TMEM lane) and, per offset o in [0, Q), the four Q-long segments f[sQ + o].
// Hop block u receives segment s of frame u - s / HQ: a lane rotation by s / HQ
// (warp shuffle).  Lanes that rotate across the warp boundary deposit their partial
// sum in a spill slot that is added at copy-out, so every output sample is written
// once, in a fixed order (deterministic).  Tiles overlap by R - 1 frames so tile
// boundaries need no exchange.  The finished hop blocks sit in shared memory (skewed
// rows: conflict-free for lanes = frames) and leave as coalesced rows multiplied by
// 1 / envelope (periodic in the interior: held in registers; edges from a cached table).
//
// Persistent CTAs (one per SM) walk the tile list.  Warp roles (512 threads, registers
// re-balanced with setmaxnreg: 64 for warps 0-7, 192 for the builders):
//   warp 0      TMA producer (basis k-chunks, 2-stage mbarrier ring)
//   warp 1      TMEM owner + MMA issuer          (warps 2-3 idle)
//   warps 8-15  operand builders (spectrogram -> scaled fp16 hi/lo planes), then the
//               epilogue and copy-out of the same tile
//   warps 4-7   scouts: stream the NEXT tile's spectrogram HBM -> shared-memory ring
//               with cp.async (no registers held, ~48 KB in flight), take the per-row
//               maxima the power-of-two scales need, and leave the tile L2-resident
//               for the builders.  HBM streaming of tile i+1 overlaps the tensor-core
//               work, epilogue and stores of tile i.
constexpr int INV_REGION = 132096;             // stages (128 KB) aliased with the output rows
constexpr int INV_SPILL = 12 * 256 * 4;        // 4 warp quarters x 3 halo blocks x H floats
constexpr int INV_SCOUT_WARPS = 4;
constexpr int INV_THREADS = 512;              // 4 warpgroups: control, scouts, 2 x builders
constexpr int INV_FIRST_BUILDER = 8;
constexpr int RING_DEPTH = 4;
constexpr int RING_ROW = 258;                  // float2 per staged spectrogram row (257 + 1)
constexpr int RING_CHUNK = 8 * RING_ROW * 8;   // 16512 B: 8 frames (or 16 bins x 128 frames)
constexpr int INV_OFF_WTAB = INV_REGION;
constexpr int INV_OFF_ROWINFO = INV_OFF_WTAB + SMEM_WTAB;
constexpr int INV_OFF_SPILL = INV_OFF_ROWINFO + 2 * TILE_M * 16;
constexpr int INV_OFF_SCRATCH = INV_OFF_SPILL + INV_SPILL;
constexpr int INV_OFF_RING = INV_OFF_SCRATCH + 2048;
constexpr int INV_SMEM_BYTES = 1024 + INV_OFF_RING + RING_DEPTH * RING_CHUNK;

struct FoldInvParams {
    const float2* spec;          // (sig, bin, frame) with element strides (complex units)
    int64_t ss, sb, sf;
    float pre_scale, pre_expo;   // 1 / scale_factor, 1 / c - 1
    float* y;                    // (sig, out_len)
    int64_t out_len;
    int64_t total_cols;          // n_signals * n_blocks (strip scheduling)
    int strips;                  // one strip of hop blocks per CTA instead of fixed tiles
    const float* env_per;        // hop: 1 / overlap-added w^2 of interior hop blocks (periodic)
    const float* wsq;            // n_fft: squared window (edge hop blocks are summed on the fly)
    int no_env;                  // gradient / ConvSTFT use: no division by the envelope
    const float4* wtab;          // Q entries: (w[n], w[N/2-n], w[N/2+n], w[N-n]) * norm / N
    int64_t n_frames;
    int n_fft, hop, q;
    int halo, adv;               // R - 1 frames of overlap between tiles; 128 - halo
    int tiles_per_signal;
    int64_t total_tiles;
    int n_blocks;                // hop blocks that reach the output
    int tmem_cols;
    float wq, w3q;               // w[Q] * norm / N, w[3Q] * norm / N
    float basis_scale_inv;
    // gradient of the forward transform (brv_fold_stft_grad): no envelope (no_env),
    // all bins weigh 1: the window table carries 1/2 and the DC / Nyquist inputs edge_gain = 2
    float edge_gain;
    float dc_gain;               // DC input (equal to edge_gain except for ConvSTFT: sqrt(2))
    int origin;                  // output sample i sits at overlap-add position i + origin
    // contiguous rows (sb == 1, sf == n_fft / 2 + 1): the scouts stage 8 rows with one bulk copy;
    // a chunk whose 16-byte rounded range would leave [spec_lo, spec_hi) (the tensor's own
    // bytes) is copied element-wise instead
    int bulk;
    uintptr_t spec_lo, spec_hi;
};

// The hop blocks of all signals are cut into one contiguous strip per CTA (CTA c of P owns the
// global columns [c G / P, (c+1) G / P)), walked in tiles of <= 128 frames that never cross a
// signal: every SM gets the same number of frames (256 fixed tiles on 148 SMs ran as 2 + 1).
// A tile that starts inside a signal recomputes the `halo` frames before it (`skip` rows whose
// hop blocks belong to the previous tile).
struct TileStrip {
    uint32_t g, g1, per_signal;
    int halo;
    // fixed tiling (small launches: measured faster there): tile ids cta, cta + ctas, ... of
    // `tiles_per_signal` tiles of TILE_M - halo new columns each
    uint32_t tile_id, n_tiles, stride;
    int tiles_per_signal;
    __device__ TileStrip(int64_t total, int64_t per_signal_, int halo_, int tiles_per_signal_, int64_t n_tiles_,
                         int cta, int ctas)
        : g((uint32_t)(total * cta / ctas)), g1((uint32_t)(total * (cta + 1) / ctas)),
          per_signal((uint32_t)per_signal_), halo(halo_), tile_id((uint32_t)cta), n_tiles((uint32_t)n_tiles_),
          stride((uint32_t)ctas), tiles_per_signal(tiles_per_signal_) {}
    __device__ __forceinline__ bool next(int64_t& sig, int64_t& t0, int& ncols, int& skip) {
        if (tiles_per_signal > 0) {
            if (tile_id >= n_tiles) return false;
            const uint32_t s = tile_id / (uint32_t)tiles_per_signal;
            const uint32_t tile = tile_id - s * (uint32_t)tiles_per_signal;
            sig = s;
            t0 = (int64_t)tile * (TILE_M - halo);
            skip = tile ? halo : 0;
            ncols = (int)min((uint32_t)TILE_M, per_signal - (uint32_t)t0);
            tile_id += stride;
            return true;
        }
        if (g >= g1) return false;
        const uint32_t s = g / per_signal, v = g - s * per_signal;
        sig = s;
        skip = (int)min((uint32_t)halo, v);
        t0 = v - skip;
        const uint32_t m = min((uint32_t)TILE_M, min(per_signal - (v - skip), g1 - g + skip));
        ncols = (int)m;
        g += m - skip;
        return true;
    }
};

template <bool DECOMP>
__device__ __forceinline__ float2 prep_bin(float2 c, float pre_scale, float pre_expo) {
    c.x *= pre_scale;
    c.y *= pre_scale;
    if (DECOMP) compress(c.x, c.y, pre_expo);
    return c;
}
__device__ __forceinline__ float abs2_finite(float2 c) {
    return fmaxf(finite_abs(c.x), finite_abs(c.y));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr));
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier;
// addresses and size are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ float rot(float v, int lane, int s) {
    return __shfl_sync(0xffffffffu, v, (lane - s) & 31);
}
// max over |re|, |im| as unsigned bit patterns (|x| ordering == uint ordering for
// non-negative floats); a result >= 0x7f800000 flags an inf / nan in the set
__device__ __forceinline__ uint32_t absbits_max(uint32_t m, float2 v) {
    return max(m, max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu));
}
// HQ = hop / Q (1, 2 or 4); FRAMES_FAST: lanes run along frames when loading the
// spectrogram (bin-major or arbitrary strides), else along bins (frame-major input);
// DECOMP: |X|^(1/c - 1) decompression in the loaders (keeps powf out of the common path).
// ODD: n_fft = 4Q - 2 with hop = Q (SGMSE's 510 / 128).  Same contractions over n = 0..Q-1, no
// Nyquist term and no n = Q column; the four formulas land on positions n, N/2 - n, N/2 + n,
// N - n with N/2 = 2Q - 1, so hop block s (positions sQ .. sQ + Q - 1) at offset o takes
//   s = 0: f[n = o]            s = 1: f[N/2 - (Q-1-o)]
//   s = 2: f[N/2 + (o+1)]  (o = Q-1: f[N - (Q-1)])      s = 3: f[N - (Q-2-o)]  (o >= Q-2: nothing)
template <int HQ, bool FRAMES_FAST, bool DECOMP, bool ODD = false>
__global__ void __launch_bounds__(INV_THREADS, 1)
istft_fold_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldInvParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ __align__(8) uint64_t region_free;      // output rows copied out: stages reusable
    __shared__ __align__(8) uint64_t scale_full[2];    // scouts -> builders (row scales ready)
    __shared__ __align__(8) uint64_t scale_empty[2];   // builders -> scouts (slot reusable)
    __shared__ __align__(8) uint64_t chunk_bar[RING_DEPTH];   // bulk chunk copy landed in the ring slot
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* orow = reinterpret_cast<float*>(stages);                 // aliases the stages
    float4* wtab = reinterpret_cast<float4*>(stages + INV_OFF_WTAB);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + INV_OFF_ROWINFO);
    float* spill = reinterpret_cast<float*>(stages + INV_OFF_SPILL);
    float* scratch = reinterpret_cast<float*>(stages + INV_OFF_SCRATCH);
    uint8_t* ring = stages + INV_OFF_RING;

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_it = 2 * (Q / BK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + BUILDERS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        mbar_init(&region_free, BUILDERS);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&scale_full[b], INV_SCOUT_WARPS);
            mbar_init(&scale_empty[b], BUILDERS);
        }
        for (int b = 0; b < RING_DEPTH; ++b) mbar_init(&chunk_bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    for (int j = threadIdx.x; j < Q; j += INV_THREADS) wtab[j] = __ldg(p.wtab + j);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    TileStrip strip(p.total_cols, p.n_blocks, p.halo, p.strips ? 0 : p.tiles_per_signal, p.total_tiles,
                    (int)blockIdx.x, (int)gridDim.x);
    int64_t sig, t0;
    int ncols, skip;

    if (warp < INV_FIRST_BUILDER) {
    // ---- control + scout warpgroups: give registers back to the builders ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ===================== TMA producer: basis k-chunks =====================
        if (elect_one()) {
            int g = 0;
            int n = 0;
            for (; strip.next(sig, t0, ncols, skip); ++n) {
                if (n > 0) mbar_wait_relaxed(&region_free, (uint32_t)((n - 1) & 1));
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    mbar_wait_relaxed(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)Q * BK * 2);
                    uint8_t* sb = stages + (size_t)s * STAGE_BYTES + STAGE_A;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl)
                            tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE), &basis_map,
                                        &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(TILE_M, Q);
            int g = 0;
            while (strip.next(sig, t0, ncols, skip)) {
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    mbar_wait_relaxed(&full_bar[s], ph, 32);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(stages + (size_t)s * STAGE_BYTES);
                    const uint32_t b0 = a0 + STAGE_A;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t d = tmem_base + (uint32_t)((pair * 2 + j) * Q);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            const uint32_t off = ks * UMMA_K * 2;
                            const uint64_t dah = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                            const uint64_t dal = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                            const uint64_t dbh = umma_desc_sw64(b0 + (j * 2) * SUB_TILE + off);
                            const uint64_t dbl = umma_desc_sw64(b0 + (j * 2 + 1) * SUB_TILE + off);
                            umma_f16(d, dah, dbh, idesc, (kc | ks) != 0);
                            umma_f16(d, dal, dbh, idesc, 1);
                            umma_f16(d, dah, dbl, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&accum_bar);
            }
        }
    } else if (warp >= 4) {
        // ===================== scouts: row maxima of the next tile ===================
        const int sw = warp - 4;                   // 0..3
        const int st = sw * 32 + lane;             // 0..127
        int n = 0;
        for (; strip.next(sig, t0, ncols, skip); ++n) {
            const int slot = n & 1;
            mbar_wait_relaxed(&scale_empty[slot], (uint32_t)(((n >> 1) & 1) ^ 1));
            if (n == 0) {
                // nobody runs ahead of the first tile: the builders (twice the threads, plain
                // 8-byte loads instead of a cp.async ring) take its row maxima themselves;
                // the scouts start on tile 1
                __syncwarp();
                if (lane == 0) mbar_arrive(&scale_full[slot]);
                continue;
            }
            float4* ri = rowinfo2 + slot * TILE_M;
            const int rows_eff = (int)max((int64_t)0, min((int64_t)ncols, p.n_frames - t0));
            const float2* xs = p.spec + sig * p.ss;
            if (FRAMES_FAST) {
                // thread = frame; chunk = 16 bins x 128 frames; a thread scans only what it
                // copied itself: cp.async group waits are the only synchronisation
                const int nch = Q / 8;                 // 2Q bins besides the Nyquist bin
                const bool live = st < rows_eff;
                const float2* col = xs + (t0 + (live ? st : 0)) * p.sf;
                float m = 0.f;
                for (int c = 0; c < nch + RING_DEPTH - 1; ++c) {
                    if (c < nch && live) {
                        float2* dst = reinterpret_cast<float2*>(ring + (c % RING_DEPTH) * RING_CHUNK) + st;
                        const float2* src = col + (int64_t)(16 * c) * p.sb;
#pragma unroll
                        for (int e = 0; e < 16; ++e) cp_async8(dst + e * TILE_M, src + (int64_t)e * p.sb);
                    }
                    cp_async_commit();
                    if (c >= RING_DEPTH - 1) {
                        cp_async_wait<RING_DEPTH - 1>();
                        const int cc = c - (RING_DEPTH - 1);
                        const float2* src =
                            reinterpret_cast<const float2*>(ring + (cc % RING_DEPTH) * RING_CHUNK) + st;
                        if (live) {
                            if (DECOMP) {
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    float2 v = prep_bin<true>(src[e * TILE_M], p.pre_scale, p.pre_expo);
                                    if (cc == 0 && e == 0) v.y = 0.f;
                                    m = fmaxf(m, abs2_finite(v));
                                }
                            } else {
                                float2 v[16];
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[e] = src[e * TILE_M];
                                if (cc == 0) v[0].y = 0.f;
                                uint32_t mb = 0u;
#pragma unroll
                                for (int e = 0; e < 16; ++e) mb = absbits_max(mb, v[e]);
                                if (mb >= 0x7f800000u) {      // inf / nan: exclude them
                                    float mf = 0.f;
#pragma unroll
                                    for (int e = 0; e < 16; ++e) mf = fmaxf(mf, abs2_finite(v[e]));
                                    mb = __float_as_uint(mf);
                                }
                                m = fmaxf(m, __uint_as_float(mb) * fabsf(p.pre_scale));
                            }
                        }
                    }
                }
                float ny = 0.f;
                if (live && !ODD)
                    ny = prep_bin<DECOMP>(__ldg(col + (int64_t)Hf * p.sb), p.pre_scale, p.pre_expo).x *
                         p.edge_gain;
                ri[st] = make_float4(row_scale(m), 0.f, 0.f, ny);
            } else {
                // chunk = 8 frames; warp sw owns rows sw and sw + 4 of every chunk (bins contiguous),
                // lanes along bins, so only warp-level synchronisation is needed
                if (p.bulk) {
                    const int nch = TILE_M / 8;
                    constexpr int ROWS_PER_WARP = 8 / INV_SCOUT_WARPS;
                    // Chunk copies: when the rows are contiguous (sf == Hf + 1) the 8 rows of a chunk are
                    // one run of bytes and ONE bulk copy (TMA engine, no per-element instructions) brings
                    // them in, over the run's 16-byte rounded range (rows sit at 8-byte alignment: the
                    // data starts 0 or 8 bytes into the ring slot, rows at pitch Hf + 1).  Per-row bulk
                    // copies were tried first: 128 small requests per tile made the scouts the critical
                    // path (cfg5 510 -> 667 us).  A chunk whose rounded range would leave the tensor,
                    // and inputs that are not plain contiguous rows, use 8-byte cp.async as before.
                    // 16 chunks per tile = 4 uses of each ring slot: the mbarrier parity of chunk cc is
                    // (cc >> 2) & 1 in every tile.
                    const int f_in = Hf + 1;
                    auto chunk_range = [&](int c, uintptr_t& a0, uint32_t& sz, int& off) {
                        const int nrows = min(8, rows_eff - 8 * c);
                        const uintptr_t a = (uintptr_t)(xs + (t0 + 8 * c) * p.sf);
                        a0 = a & ~(uintptr_t)15;
                        off = (int)((a - a0) >> 3);
                        sz = ((uint32_t)(a - a0) + (uint32_t)(nrows * f_in) * 8u + 15u) & ~15u;
                        return p.bulk && nrows > 0 && a0 >= p.spec_lo && a0 + sz <= p.spec_hi;
                    };
                    for (int c = 0; c < nch + RING_DEPTH - 1; ++c) {
                        if (c < nch) {
                            float2* dst = reinterpret_cast<float2*>(ring + (c % RING_DEPTH) * RING_CHUNK);
                            uintptr_t a0;
                            uint32_t sz;
                            int off;
                            const bool bulk_chunk = chunk_range(c, a0, sz, off);
                            if (bulk_chunk) {
                                if (st == 0) {
                                    mbar_arrive_expect_tx(&chunk_bar[c % RING_DEPTH], sz);
                                    bulk_g2s(dst, reinterpret_cast<const void*>(a0), sz, &chunk_bar[c % RING_DEPTH]);
                                }
                            } else {
                                // (the chunk barriers are used only when the scout warps run in lock
                                //  step, i.e. with bulk copies: free-running warps could lap the 1-bit
                                //  phase parity of a barrier that another warp arrives on)
                                if (p.bulk && st == 0) mbar_arrive_expect_tx(&chunk_bar[c % RING_DEPTH], 0u);
#pragma unroll
                                for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                                    const int r = sw + INV_SCOUT_WARPS * rr, row = 8 * c + r;
                                    if (row < rows_eff) {
                                        const float2* xr = xs + (t0 + row) * p.sf;
#pragma unroll
                                        for (int j = 0; j < 9; ++j) {
                                            const int b = j * 32 + lane;
                                            if (b <= Hf) cp_async8(dst + r * RING_ROW + b, xr + b);   // sb == 1
                                        }
                                    }
                                }
                            }
                        }
                        cp_async_commit();
                        if (c >= RING_DEPTH - 1) {
                            cp_async_wait<RING_DEPTH - 1>();
                            const int cc = c - (RING_DEPTH - 1);
                            if (p.bulk) mbar_wait_relaxed(&chunk_bar[cc % RING_DEPTH], (uint32_t)((cc >> 2) & 1), 32);
                            __syncwarp();
                            const float2* ring_c =
                                reinterpret_cast<const float2*>(ring + (cc % RING_DEPTH) * RING_CHUNK);
                            uintptr_t a0c;
                            uint32_t szc;
                            int offc;
                            const bool bulk_chunk = chunk_range(cc, a0c, szc, offc);
                            const int pitch_c = bulk_chunk ? f_in : RING_ROW;
#pragma unroll
                            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                                const int r = sw + INV_SCOUT_WARPS * rr, row = 8 * cc + r;
                                float m = 0.f, ny = 0.f;
                                if (row < rows_eff) {
                                    // data starts 0 or 1 elements into a bulk-copied slot
                                    const float2* src = ring_c + (bulk_chunk ? offc : 0) + r * pitch_c;
                                    float2 v[8];                 // bins lane + 32 j < Hf (Hf <= 256)
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        v[j] = j * 32 < Hf ? src[j * 32 + lane] : make_float2(0.f, 0.f);
                                    if (lane == 0) v[0].y = 0.f; // Im X[0] never reaches the output
                                    if (DECOMP) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            m = fmaxf(m, abs2_finite(prep_bin<true>(v[j], p.pre_scale, p.pre_expo)));
                                    } else {
                                        uint32_t mb = 0u;
#pragma unroll
                                        for (int j = 0; j < 8; ++j) mb = absbits_max(mb, v[j]);
                                        if (mb >= 0x7f800000u) {  // inf / nan: exclude them
                                            float mf = 0.f;
#pragma unroll
                                            for (int j = 0; j < 8; ++j) mf = fmaxf(mf, abs2_finite(v[j]));
                                            mb = __float_as_uint(mf);
                                        }
                                        m = __uint_as_float(mb) * fabsf(p.pre_scale);
                                    }
                                    if (lane == 0 && !ODD)
                                        ny = prep_bin<DECOMP>(src[Hf], p.pre_scale, p.pre_expo).x * p.edge_gain;
                                }
#pragma unroll
                                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                                if (lane == 0) ri[row] = make_float4(row_scale(m), 0.f, 0.f, ny);
                            }
                            // a bulk-copied chunk is read by every scout warp: all of them are done with
                            // the slot before scout thread 0 refills it (the cp.async path refills only
                            // the warp's own rows: lock-stepping the warps there cost cfg5 510 -> 606 us)
                            if (p.bulk) named_bar_sync(2, INV_SCOUT_WARPS * 32);
                            else __syncwarp();
                        }
                    }
                } else {
                    // (verbatim the pre-bulk loop: compile-time ring offsets and free-running warps;
                    //  sharing one loop with runtime pitches cost cfg5 510 -> 600 us)
                    const int nch = TILE_M / 8;
                    constexpr int ROWS_PER_WARP = 8 / INV_SCOUT_WARPS;
                    for (int c = 0; c < nch + RING_DEPTH - 1; ++c) {
                        if (c < nch) {
                            float2* dst = reinterpret_cast<float2*>(ring + (c % RING_DEPTH) * RING_CHUNK);
#pragma unroll
                            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                                const int r = sw + INV_SCOUT_WARPS * rr, row = 8 * c + r;
                                if (row < rows_eff) {
                                    const float2* xr = xs + (t0 + row) * p.sf;
#pragma unroll
                                    for (int j = 0; j < 9; ++j) {
                                        const int b = j * 32 + lane;
                                        if (b <= Hf) cp_async8(dst + r * RING_ROW + b, xr + b);   // sb == 1
                                    }
                                }
                            }
                        }
                        cp_async_commit();
                        if (c >= RING_DEPTH - 1) {
                            cp_async_wait<RING_DEPTH - 1>();
                            __syncwarp();
                            const int cc = c - (RING_DEPTH - 1);
                            const float2* src =
                                reinterpret_cast<const float2*>(ring + (cc % RING_DEPTH) * RING_CHUNK);
#pragma unroll
                            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                                const int r = sw + INV_SCOUT_WARPS * rr, row = 8 * cc + r;
                                float m = 0.f, ny = 0.f;
                                if (row < rows_eff) {
                                    float2 v[8];                 // bins lane + 32 j < Hf (Hf <= 256)
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        v[j] = j * 32 < Hf ? src[r * RING_ROW + j * 32 + lane]
                                                           : make_float2(0.f, 0.f);
                                    if (lane == 0) v[0].y = 0.f; // Im X[0] never reaches the output
                                    if (DECOMP) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            m = fmaxf(m, abs2_finite(prep_bin<true>(v[j], p.pre_scale, p.pre_expo)));
                                    } else {
                                        uint32_t mb = 0u;
#pragma unroll
                                        for (int j = 0; j < 8; ++j) mb = absbits_max(mb, v[j]);
                                        if (mb >= 0x7f800000u) {  // inf / nan: exclude them
                                            float mf = 0.f;
#pragma unroll
                                            for (int j = 0; j < 8; ++j) mf = fmaxf(mf, abs2_finite(v[j]));
                                            mb = __float_as_uint(mf);
                                        }
                                        m = __uint_as_float(mb) * fabsf(p.pre_scale);
                                    }
                                    if (lane == 0 && !ODD)
                                        ny = prep_bin<DECOMP>(src[r * RING_ROW + Hf], p.pre_scale, p.pre_expo).x *
                                             p.edge_gain;
                                }
#pragma unroll
                                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                                if (lane == 0) ri[row] = make_float4(row_scale(m), 0.f, 0.f, ny);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&scale_full[slot]);
        }
    }
    } else {
        // ===================== builders, then epilogue ==========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        const int bw = warp - INV_FIRST_BUILDER;   // 0..7
        const int bt = bw * 32 + lane;             // 0..255
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int hsel = bw >> 2;                  // which half of the offsets (epilogue)
        const int pitch = H + 1;                   // skew: lanes (= frames) hit distinct banks
        constexpr int R = 4 / HQ;                  // frames overlapping one hop block
        float env_reg[8];                          // 1 / envelope of interior hop blocks
#pragma unroll
        for (int j = 0; j < 8; ++j)
            env_reg[j] = p.no_env ? 1.f : (lane + 32 * j < H ? __ldg(p.env_per + lane + 32 * j) : 0.f);

        int g = 0, n = 0;
        for (; strip.next(sig, t0, ncols, skip); ++n) {
            const int slot = n & 1;
            float4* rowinfo = rowinfo2 + slot * TILE_M;
            const int rows_eff = (int)max((int64_t)0, min((int64_t)ncols, p.n_frames - t0));
            const float2* xs = p.spec + sig * p.ss;
            BRV_STAMP(n * 8 + 0);
            mbar_wait(&scale_full[slot], (uint32_t)((n >> 1) & 1));
            if (n == 0) {
                // ---- first tile: row maxima (power-of-two scales) and Nyquist terms by the
                //      builders; the tile is L2-resident afterwards for the main loop ----------
                if (FRAMES_FAST) {
                    const int row = bt & 127, kh = bt >> 7;
                    const bool live = row < rows_eff;
                    const float2* xr = xs + (t0 + (live ? row : 0)) * p.sf;
                    float m = 0.f;
                    for (int b0 = kh * Q; b0 < (kh + 1) * Q; b0 += 32) {
                        float2 v[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            v[e] = live ? __ldg(xr + (int64_t)(b0 + e) * p.sb) : make_float2(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float2 c = prep_bin<DECOMP>(v[e], p.pre_scale, p.pre_expo);
                            if (b0 + e == 0) c.y = 0.f;
                            m = fmaxf(m, abs2_finite(c));
                        }
                    }
                    scratch[kh * 128 + row] = m;
                    named_bar_sync(1, BUILDER_THREADS);
                    if (kh == 0) {
                        float ny = 0.f;
                        if (live && !ODD)
                            ny = prep_bin<DECOMP>(__ldg(xr + (int64_t)Hf * p.sb), p.pre_scale, p.pre_expo).x *
                                 p.edge_gain;
                        rowinfo[row] = make_float4(row_scale(fmaxf(scratch[row], scratch[128 + row])),
                                                   0.f, 0.f, ny);
                    }
                } else {
                    const int nj = Q / 16;             // 32-bin groups below the Nyquist bin
#pragma unroll 1
                    for (int rb = 0; rb < 16; rb += 4) {
                        float2 v[4][9];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int row = bw * 16 + rb + r;
                            const float2* xr = xs + (t0 + row) * p.sf;     // sb == 1
#pragma unroll
                            for (int j = 0; j < 9; ++j) {
                                const int b = j * 32 + lane;
                                v[r][j] = (row < rows_eff && b <= Hf) ? __ldg(xr + b)
                                                                      : make_float2(0.f, 0.f);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int row = bw * 16 + rb + r;
                            float m = 0.f, ny = 0.f;
#pragma unroll
                            for (int j = 0; j < 9; ++j) {
                                float2 c = prep_bin<DECOMP>(v[r][j], p.pre_scale, p.pre_expo);
                                if (j == 0 && lane == 0) c.y = 0.f;   // Im X[0] never reaches the output
                                if (j < nj) m = fmaxf(m, abs2_finite(c));
                                if (!ODD && j == nj && lane == 0) ny = c.x * p.edge_gain;
                            }
#pragma unroll
                            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                            if (lane == 0) rowinfo[row] = make_float4(row_scale(m), 0.f, 0.f, ny);
                        }
                    }
                }
                named_bar_sync(1, BUILDER_THREADS);
            }
            BRV_STAMP(n * 8 + 1);

            // ---- main loop: load (L2-resident after the scouts), scale, split, store ----
            if (FRAMES_FAST) {
                const int row = bt & 127, kh = bt >> 7;
                const bool live = row < rows_eff;
                const float2* xr = xs + (t0 + (live ? row : 0)) * p.sf;
                const float sc = rowinfo[row].x;
                const uint32_t sw = (uint32_t)((row >> 1) & 3);
                float pacc = 0.f, racc = 0.f;
                for (int kc = 0; kc < Q / BK; ++kc) {
                    float2 c[32];                      // bins 64 kc + 32 kh + e
                    const int bin0 = 64 * kc + 32 * kh;
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) c[e] = __ldg(xr + (int64_t)(bin0 + e) * p.sb);
#pragma unroll
                        for (int e = 0; e < 32; ++e) c[e] = prep_bin<DECOMP>(c[e], p.pre_scale, p.pre_expo);
                        if (bin0 == 0) {
                            c[0].y = 0.f;              // Im X[0] is ignored by the c2r inverse
                            c[0].x *= p.dc_gain;
                        }
                        if (ODD && bin0 + 32 == 2 * Q) {
                            c[31].y = 0.f;             // so is Im X[N/2]
                            c[31].x *= p.edge_gain;    // (1 except in the STFT gradient: unit bin weights)
                        }
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            pacc += c[2 * j].x - c[2 * j + 2].x;          // (-1)^m Re X[2m]
                            racc += c[2 * j + 1].y - c[2 * j + 3].y;      // (-1)^m Im X[2m+1]
                        }
                        if (bin0 == 0) pacc -= 0.5f * c[0].x;             // c_0 = 1, the others 2
                    }
#pragma unroll
                    for (int pair = 0; pair < 2; ++pair, ++g) {
                        const int s = g % STAGES;
                        const uint32_t ph = (g / STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (live) {
                            uint8_t* sa = stages + (size_t)s * STAGE_BYTES + row * (BK * 2);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {          // sub-GEMM: even / odd bins
#pragma unroll
                                for (int ch = 0; ch < 2; ++ch) {   // 16-byte chunk = 8 k values
                                    uint32_t hi[4], lo[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const int m0 = ch * 8 + 2 * e;     // local m of the k pair
                                        const float2 ca = c[2 * m0 + j], cb = c[2 * m0 + 2 + j];
                                        const float v0 = (pair ? ca.y : ca.x) * sc;
                                        const float v1 = (pair ? cb.y : cb.x) * sc;
                                        const __half2 h = __floats2half2_rn(v0, v1);
                                        const float2 hf = __half22float2(h);
                                        const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                                        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
                                        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
                                    }
                                    const uint32_t dst = (((uint32_t)(2 * kh + ch)) ^ sw) << 4;
                                    *reinterpret_cast<uint4*>(sa + (j * 2) * SUB_TILE + dst) =
                                        make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                    *reinterpret_cast<uint4*>(sa + (j * 2 + 1) * SUB_TILE + dst) =
                                        make_uint4(lo[0], lo[1], lo[2], lo[3]);
                                }
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[s]);
                    }
                }
                scratch[kh * 128 + row] = pacc;
                scratch[256 + kh * 128 + row] = racc;
                named_bar_sync(1, BUILDER_THREADS);
                if (kh == 0) {
                    rowinfo[row].y = 2.f * (scratch[row] + scratch[128 + row]);
                    rowinfo[row].z = 2.f * (scratch[256 + row] + scratch[384 + row]);
                }
            } else {
                const int half = lane >> 4;            // which of the warp's two rows per pass
                const int pr = lane & 15;              // m pair inside the 32-wide k-chunk
                const uint32_t chunk = (uint32_t)(pr >> 2);
                float pacc[8], racc[8], rscale[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    pacc[i] = racc[i] = 0.f;
                    const int row = bw * 16 + 2 * i + half;
                    rscale[i] = row < rows_eff ? rowinfo[row].x : 0.f;
                }
                // bins 2 m0 .. 2 m0 + 3 of the warp's 16 rows for k-chunk kc (L2-resident after the
                // row-maxima pass).  Issuing the next chunk's loads before this one is consumed was
                // tried: 64 more live registers spill and cost more than the latency they hide.
                auto load_chunk = [&](float2 (&dst)[8][4], int kc) {
                    const int m0 = kc * BK + 2 * pr;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = bw * 16 + 2 * i + half;
                        const float2* xr = xs + (t0 + row) * p.sf + (int64_t)(2 * m0) * p.sb;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            dst[i][e] = row < rows_eff ? __ldg(xr + (int64_t)e * p.sb)
                                                       : make_float2(0.f, 0.f);
                    }
                };
                for (int kc = 0; kc < Q / BK; ++kc) {
                    const int m0 = kc * BK + 2 * pr;   // bins 2 m0 .. 2 m0 + 3
                    float2 c[8][4];
                    load_chunk(c, kc);
                    // (selects, not lane-dependent branches, inside the unrolled rows)
                    const bool dc = m0 == 0, ny_im = ODD && m0 == Q - 2;
                    const float dc_mul = dc ? p.dc_gain : 1.f, dc_half = dc ? 0.5f : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            c[i][e] = prep_bin<DECOMP>(c[i][e], p.pre_scale, p.pre_expo);
                        c[i][0].y = dc ? 0.f : c[i][0].y;      // Im X[0] is ignored by the c2r inverse
                        c[i][0].x *= dc_mul;
                        c[i][3].y = ny_im ? 0.f : c[i][3].y;   // so is Im X[N/2]
                        if (ODD) c[i][3].x *= ny_im ? p.edge_gain : 1.f;
                        pacc[i] += c[i][0].x - c[i][2].x;
                        racc[i] += c[i][1].y - c[i][3].y;
                        pacc[i] -= dc_half * c[i][0].x;
                    }
#pragma unroll
                    for (int pair = 0; pair < 2; ++pair, ++g) {
                        const int s = g % STAGES;
                        const uint32_t ph = (g / STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* sa = stages + (size_t)s * STAGE_BYTES;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = bw * 16 + 2 * i + half;
                            // rows >= rows_eff are built too (their scale is 0 / their inputs are zero or stale,
                            // their accumulator rows are never read): a `continue` here would split the unrolled rows
                            // into separate basic blocks and serialise their dependent chains
                            const float sc = rscale[i];
                            uint8_t* dst = sa + row * (BK * 2) +
                                           ((chunk ^ (uint32_t)((row >> 1) & 3)) << 4) + (pr & 3) * 4;
                            if (pair == 0) {
                                split_store(dst, dst + SUB_TILE, c[i][0].x * sc, c[i][2].x * sc);
                                split_store(dst + 2 * SUB_TILE, dst + 3 * SUB_TILE, c[i][1].x * sc,
                                            c[i][3].x * sc);
                            } else {
                                split_store(dst, dst + SUB_TILE, c[i][0].y * sc, c[i][2].y * sc);
                                split_store(dst + 2 * SUB_TILE, dst + 3 * SUB_TILE, c[i][1].y * sc,
                                            c[i][3].y * sc);
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[s]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = pacc[i], b = racc[i];
#pragma unroll
                    for (int o = 8; o; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        b += __shfl_xor_sync(0xffffffffu, b, o);
                    }
                    const int row = bw * 16 + 2 * i + half;
                    if (pr == 0 && row < rows_eff) {
                        rowinfo[row].y = 2.f * a;
                        rowinfo[row].z = 2.f * b;
                    }
                }
            }
            named_bar_sync(1, BUILDER_THREADS);
            BRV_STAMP(n * 8 + 2);

            // ---- epilogue: TMEM -> segments -> lane-rotated overlap-add -> skewed rows ----
            {
                const int row = q * 32 + lane;         // frame t0 + row == hop block t0 + row
                const bool live = row < rows_eff;
                const float4 ri = rowinfo[row];
                const float g0 = live ? p.basis_scale_inv / ri.x : 0.f;
                const float xny = live ? ri.w : 0.f;
                // f[Q], f[3Q]: Q is even, so the Nyquist term enters with +1
                const float fq = live ? (ri.y - ri.z + ri.w) * p.wq : 0.f;
                const float f3q = live ? (ri.y + ri.z + ri.w) * p.w3q : 0.f;
                mbar_wait(&accum_bar, (uint32_t)(n & 1));
                BRV_STAMP(n * 8 + 3);
                tcgen05_fence_after();
                float* my_row = orow + row * pitch;
                float* my_spill = spill + ((q + 1) * 3 + lane) * H;
                const bool do_spill = q < 3 && lane < 3;
                const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
                const int o_begin = hsel * (Q / 2), o_end = o_begin + Q / 2;
                uint32_t carry[4];                     // column Q - c0 of Ce, Co, Se, So
                if (o_begin > 0) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        tmem_ld1_nowait(tq + (uint32_t)(a * Q + Q - o_begin), &carry[a]);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) carry[a] = 0u;
                }
#pragma unroll 1
                for (int c0 = o_begin; c0 < o_end; c0 += 8) {
                    uint32_t A[4][8], B[4][8];
                    uint32_t XA[4] = {0u, 0u, 0u, 0u}, XB[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        tmem_ld8_nowait(tq + (uint32_t)(a * Q + c0), A[a]);
                        tmem_ld8_nowait(tq + (uint32_t)(a * Q + Q - c0 - 8), B[a]);
                        if (ODD && c0 + 8 < Q) {       // one column of look-ahead on either side
                            tmem_ld1_nowait(tq + (uint32_t)(a * Q + c0 + 8), &XA[a]);
                            tmem_ld1_nowait(tq + (uint32_t)(a * Q + Q - c0 - 9), &XB[a]);
                        }
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int o = c0 + i;
                        const float sg = (i & 1) ? -xny : xny;         // c0 is even
                        const float4 wa = wtab[o];
                        const float ce = __uint_as_float(A[0][i]), co = __uint_as_float(A[1][i]);
                        const float se = __uint_as_float(A[2][i]), so = __uint_as_float(A[3][i]);
                        float seg0 = (((ce + co) - (se + so)) * g0 + sg) * wa.x;
                        float seg2 = (((ce - co) - (se - so)) * g0 + sg) * wa.z;
                        float seg1, seg3;
                        // (selects, not branches: a branch inside the unrolled offsets would split
                        //  them into basic blocks and serialise their dependent chains; the table
                        //  reads one entry past either end land in valid shared memory and are dropped)
                        if (ODD) {
                            // column Q-1-o -> position N/2 - n
                            const float ce1 = __uint_as_float(B[0][7 - i]), co1 = __uint_as_float(B[1][7 - i]);
                            const float se1 = __uint_as_float(B[2][7 - i]), so1 = __uint_as_float(B[3][7 - i]);
                            seg1 = ((ce1 - co1) + (se1 - so1)) * g0 * wtab[Q - 1 - o].y;
                            const float ce2 = __uint_as_float(i < 7 ? A[0][(i + 1) & 7] : XA[0]);
                            const float co2 = __uint_as_float(i < 7 ? A[1][(i + 1) & 7] : XA[1]);
                            const float se2 = __uint_as_float(i < 7 ? A[2][(i + 1) & 7] : XA[2]);
                            const float so2 = __uint_as_float(i < 7 ? A[3][(i + 1) & 7] : XA[3]);
                            // column Q-2-o -> position N - n (column 0 carries weight 0)
                            const float ce3 = __uint_as_float(i < 7 ? B[0][(6 - i) & 7] : XB[0]);
                            const float co3 = __uint_as_float(i < 7 ? B[1][(6 - i) & 7] : XB[1]);
                            const float se3 = __uint_as_float(i < 7 ? B[2][(6 - i) & 7] : XB[2]);
                            const float so3 = __uint_as_float(i < 7 ? B[3][(6 - i) & 7] : XB[3]);
                            const bool last = i == 7 && o == Q - 1;    // position N - (Q-1): column Q-1 = A[.][7]
                            seg2 = last ? ((ce + co) + (se + so)) * g0 * wa.w
                                        : ((ce2 - co2) - (se2 - so2)) * g0 * wtab[o + 1].z;
                            seg3 = last ? 0.f : ((ce3 + co3) + (se3 + so3)) * g0 * wtab[Q - 2 - o].w;
                        } else {
                            const float4 wb = wtab[Q - o];
                            const float ce2 = __uint_as_float(i == 0 ? carry[0] : B[0][8 - i]);
                            const float co2 = __uint_as_float(i == 0 ? carry[1] : B[1][8 - i]);
                            const float se2 = __uint_as_float(i == 0 ? carry[2] : B[2][8 - i]);
                            const float so2 = __uint_as_float(i == 0 ? carry[3] : B[3][8 - i]);
                            seg1 = (((ce2 - co2) + (se2 - so2)) * g0 + sg) * wb.y;
                            seg3 = (((ce2 + co2) + (se2 + so2)) * g0 + sg) * wb.w;
                            if (i == 0) {                              // o == 0 only happens at i == 0
                                seg1 = o == 0 ? fq : seg1;
                                seg3 = o == 0 ? f3q : seg3;
                            }
                        }
                        if (!live) seg0 = seg1 = seg2 = seg3 = 0.f;    // dead rows hold garbage
                        if (HQ == 1) {
                            const float r1 = rot(seg1, lane, 1), r2 = rot(seg2, lane, 2),
                                        r3 = rot(seg3, lane, 3);
                            float acc = seg0;
                            acc += lane >= 1 ? r1 : 0.f;
                            acc += lane >= 2 ? r2 : 0.f;
                            acc += lane >= 3 ? r3 : 0.f;
                            my_row[o] = acc;
                            float sp = lane < 1 ? r1 : 0.f;
                            sp += lane < 2 ? r2 : 0.f;
                            sp += r3;
                            if (do_spill) my_spill[o] = sp;
                        } else if (HQ == 2) {
                            const float r2 = rot(seg2, lane, 1), r3 = rot(seg3, lane, 1);
                            my_row[o] = lane >= 1 ? seg0 + r2 : seg0;
                            my_row[Q + o] = lane >= 1 ? seg1 + r3 : seg1;
                            if (do_spill && lane == 0) {
                                my_spill[o] = r2;
                                my_spill[Q + o] = r3;
                            }
                        } else {
                            my_row[o] = seg0;
                            my_row[Q + o] = seg1;
                            my_row[2 * Q + o] = seg2;
                            my_row[3 * Q + o] = seg3;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a) carry[a] = B[a][0];
                }
                tcgen05_fence_before();
                BRV_STAMP(n * 8 + 4);
            }
            named_bar_sync(1, BUILDER_THREADS);
            BRV_STAMP(n * 8 + 5);

            // ---- copy-out: finished hop blocks, * 1 / envelope, centre trim ---------------
            {
                float* ys = p.y + sig * p.out_len;
                for (int r = skip + bw; r < ncols; r += BUILDERS) {
                    const int64_t u = t0 + r;
                    const float* src = orow + r * pitch;
                    const int qq = r >> 5, lr = r & 31;
                    const float* sp = (qq > 0 && lr < p.halo) ? spill + (qq * 3 + lr) * H : nullptr;
                    const int64_t i0 = u * H - p.origin;
                    const bool use_reg = p.no_env || (u >= R - 1 && u <= p.n_frames - 1);
                    // offsets [lo, hi) of this hop block that exist in the output, as 32-bit
                    // values computed once per row (the warps of this phase run dependent
                    // chains at ~1 instruction per 5 cycles: per-element 64-bit range checks
                    // made the copy-out 4x longer than its loads and stores)
                    const int lo = i0 < 0 ? (int)min((int64_t)H, -i0) : 0;
                    const int hi = (int)max((int64_t)0, min((int64_t)H, p.out_len - i0));
                    float* yrow = ys + i0;             // dereferenced inside [lo, hi) only
                    if (lo == 0 && hi == H && use_reg && !sp) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int off = lane + 32 * j;
                            if (off < H) yrow[off] = src[off] * env_reg[j];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int off = lane + 32 * j;
                            if (off >= lo && off < hi) {
                                float v = src[off];
                                if (sp) v += sp[off];
                                yrow[off] = v * (use_reg ? env_reg[j] : ola_inv_envelope(p.env_per, p.wsq, N, H, p.n_frames, u, off));
                            }
                        }
                    }
                }
            }
            BRV_STAMP(n * 8 + 7);
            // the output rows alias the operand stages: nobody may start building the
            // next tile before every warp has copied its rows out
            named_bar_sync(1, BUILDER_THREADS);
            BRV_STAMP(n * 8 + 6);
            if (lane == 0) {
                mbar_arrive(&region_free);
                mbar_arrive(&scale_empty[slot]);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

#include "brv_fold_t.cuh"

struct FoldBasis {
    __half* data = nullptr;      // [plane 2][sub 4][m Q][n Q]
    CUtensorMap map;
    float scale_inv = 1.f;
};
struct FoldPlan {
    FoldBasis fwd, inv;
    float4* wtab = nullptr;      // Q permuted, normalised window entries
    float4* wtab_inv = nullptr;  // the same permutation of w * sqrt(sum w^2) / N
    float4* wtab_grad = nullptr; // forward-convention table of 2 w sqrt(sum w^2) / N (iSTFT gradient)
    float wq_g = 0.f, w3q_g = 0.f, wmax_g = 0.f;
    float4* wtab_fgrad = nullptr;   // inverse-convention table of w / (2 sqrt(sum w^2)) (STFT gradient)
    float wq_fg = 0.f, w3q_fg = 0.f;
    float wq = 0.f, w3q = 0.f, wmax = 0.f;
    float wq_inv = 0.f, w3q_inv = 0.f;
    int q = 0, tmem_cols = 0;
    int odd = 0;                 // n_fft = 4Q - 2 (see FoldFwdParams::odd)
    int hq = 0;                  // hop / Q when the fused overlap-add applies (1, 2, 4), else 0
    float* env_per = nullptr;    // hop: 1 / overlap-added w^2 for interior hop blocks
    int sm_count = 0;
    std::map<int64_t, float*> inv_env;   // n_frames -> 1 / overlap-added w^2 (device)
};

// cos / sin of 2*pi*m/N with exact values on the axes
void unit_root(long long m, int N, double* c, double* s) {
    m %= N;
    if ((4 * m) % N == 0) {
        static const double cs[4] = {1, 0, -1, 0}, sn[4] = {0, 1, 0, -1};
        const int qd = (int)((4 * m) / N);
        *c = cs[qd];
        *s = sn[qd];
        return;
    }
    const double ang = 2.0 * M_PI * (double)m / (double)N;
    *c = cos(ang);
    *s = sin(ang);
}

// value(sub, m, n) split into scaled fp16 hi / lo planes behind a SWIZZLE_64B
// tensor map of (BK x Q) boxes.
template <class F>
int build_fold_basis(FoldBasis* b, int Q, F value) {
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return BRV_ERR_UNSUPPORTED;
    double mx = 0;
    for (int sub = 0; sub < 4; ++sub)
        for (int m = 0; m < Q; ++m)
            for (int n = 0; n < Q; ++n) mx = fmax(mx, fabs(value(sub, m, n)));
    if (!(mx > 0)) return BRV_ERR_UNSUPPORTED;
    int e;
    frexp(mx, &e);
    const double sB = ldexp(1.0, 13 - e);              // mx * sB in [2^12, 2^13)
    b->scale_inv = (float)(1.0 / sB);
    const size_t plane = (size_t)4 * Q * Q;
    std::vector<__half> host(2 * plane);
    for (int sub = 0; sub < 4; ++sub)
        for (int m = 0; m < Q; ++m)
            for (int n = 0; n < Q; ++n) {
                const double v = value(sub, m, n) * sB;
                const __half h = __float2half_rn((float)v);
                const __half l = __float2half_rn((float)(v - (double)__half2float(h)));
                const size_t at = ((size_t)sub * Q + m) * Q + n;
                host[at] = h;
                host[plane + at] = l;
            }
    if (cudaMalloc((void**)&b->data, host.size() * sizeof(__half)) != cudaSuccess)
        return brv_fail_cuda(cudaGetLastError(), "cudaMalloc(folded basis)");
    if (cudaMemcpy(b->data, host.data(), host.size() * sizeof(__half), cudaMemcpyHostToDevice) !=
        cudaSuccess)
        return brv_fail_cuda(cudaGetLastError(), "cudaMemcpy(folded basis)");
    cuuint64_t dims[2] = {(cuuint64_t)Q, (cuuint64_t)(8 * Q)};
    cuuint64_t strides[1] = {(cuuint64_t)Q * sizeof(__half)};
    cuuint32_t box[2] = {BK, (cuuint32_t)Q};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&b->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b->data, dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        return brv_fail(BRV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return BRV_OK;
}

void free_fold(FoldPlan* fp) {
    if (!fp) return;
    cudaFree(fp->fwd.data);
    cudaFree(fp->inv.data);
    cudaFree(fp->wtab);
    cudaFree(fp->wtab_inv);
    cudaFree(fp->wtab_grad);
    cudaFree(fp->wtab_fgrad);
    cudaFree(fp->env_per);
    for (auto& kv : fp->inv_env) cudaFree(kv.second);
    delete fp;
}

}  // namespace

// 0: transposed strip kernels (brv_fold_t.cuh) where they apply; 4: the one-tile-per-TMEM kernels,
// forward picked by tile count; 2 / 3: those with the forward forced to one tile per CTA /
// persistent two-pass; 5 / 6: transposed forward only / transposed inverse only (brv_set_tc_variant)
int g_brv_fold_variant = 0;

bool brv_fold_supported(const brv_stft_plan* p) { return p->fold != nullptr; }

// the gradient entry points reuse the kernels with edge-bin weights (DC and Nyquist count once in
// the Hermitian sums); for n_fft = 4Q - 2 the Nyquist bin is the last odd bin of the contraction
bool brv_fold_grad_supported(const brv_stft_plan* p) { return p->fold != nullptr; }

int brv_fold_plan_init(brv_stft_plan* p) {
    const int N = p->n_fft;
    // N = 4Q, or N = 4Q - 2 (SGMSE's 510): Q a multiple of the 32-wide k-chunk
    const bool odd = N % 4 == 2 && (N + 2) % 128 == 0;
    if (!p->onesided || !(N % 128 == 0 || odd) || N > 512) return BRV_OK;
    if (!encode_tiled()) return BRV_OK;
    const int Q = (N + 2) / 4, Hf = N / 2;
    FoldPlan* fp = new FoldPlan();
    fp->q = Q;
    fp->odd = odd ? 1 : 0;
    fp->tmem_cols = 4 * Q <= 128 ? 128 : (4 * Q <= 256 ? 256 : 512);
    int rc = build_fold_basis(&fp->fwd, Q, [&](int sub, int m, int n) {
        const long long k = (sub & 1) ? 2 * m + 1 : 2 * m;
        double c, s;
        unit_root(k * n, N, &c, &s);
        return sub < 2 ? c : -s;
    });
    if (rc == BRV_OK) {
        std::vector<float4> wt(Q);
        double wmax = 0;
        for (int n = 0; n < N; ++n) wmax = fmax(wmax, fabs(p->window[n] / p->norm));
        for (int n = 0; n < Q; ++n) {
            wt[n].x = (float)(p->window[n] / p->norm);
            wt[n].y = (float)(p->window[Hf - n] / p->norm);
            wt[n].z = n ? (float)(p->window[Hf + n] / p->norm) : 0.f;
            wt[n].w = n ? (float)(p->window[N - n] / p->norm) : 0.f;
        }
        fp->wq = odd ? 0.f : (float)(p->window[Q] / p->norm);
        fp->w3q = odd ? 0.f : (float)(p->window[3 * Q] / p->norm);
        fp->wmax = (float)wmax;
        if (cudaMalloc((void**)&fp->wtab, Q * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(fp->wtab, wt.data(), Q * sizeof(float4), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded window table");
    }
    if (rc == BRV_OK) {
        // gradient of the inverse transform = this forward kernel on gy / envelope with the
        // inverse's window scaling and Hermitian weights (2, folded in here; DC / Nyquist get 1/2)
        const double gs = 2.0 * p->norm / N;
        std::vector<float4> wt(Q);
        double wmax = 0;
        for (int n = 0; n < N; ++n) wmax = fmax(wmax, fabs(p->window[n] * gs));
        for (int n = 0; n < Q; ++n) {
            wt[n].x = (float)(p->window[n] * gs);
            wt[n].y = (float)(p->window[Hf - n] * gs);
            wt[n].z = n ? (float)(p->window[Hf + n] * gs) : 0.f;
            wt[n].w = n ? (float)(p->window[N - n] * gs) : 0.f;
        }
        fp->wq_g = odd ? 0.f : (float)(p->window[Q] * gs);
        fp->w3q_g = odd ? 0.f : (float)(p->window[3 * Q] * gs);
        fp->wmax_g = (float)wmax;
        if (cudaMalloc((void**)&fp->wtab_grad, Q * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(fp->wtab_grad, wt.data(), Q * sizeof(float4), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded gradient window table");
    }
    if (rc == BRV_OK &&
        (cudaFuncSetAttribute(stft_fold_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(stft_fold_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SMEM_BYTES) != cudaSuccess))
        rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(stft_fold_kernel)");
    if (rc == BRV_OK &&
        (cudaFuncSetAttribute(stft_fold2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              F2_SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(stft_fold2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              F2_SMEM_BYTES) != cudaSuccess))
        rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(stft_fold2_kernel)");
    if (rc == BRV_OK &&
        (cudaFuncSetAttribute(stft_t_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              FtLayout<64>::SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(stft_t_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              FtLayout<64>::SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(stft_t_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              FtLayout<32>::SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(stft_t_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              FtLayout<32>::SMEM_BYTES) != cudaSuccess))
        rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(stft_t_kernel)");
    if (rc == BRV_OK &&
        cudaDeviceGetAttribute(&fp->sm_count, cudaDevAttrMultiProcessorCount, p->device) !=
            cudaSuccess)
        rc = brv_fail_cuda(cudaGetLastError(), "cudaDeviceGetAttribute(SM count)");
    // inverse: trigonometric basis with the Hermitian weights (c_0 = 1, else 2); the
    // window, sqrt(sum w^2) and 1/N live in the permuted window table
    if (rc == BRV_OK)
        rc = build_fold_basis(&fp->inv, Q, [&](int sub, int n, int m) {
            const long long k = (sub & 1) ? 2 * m + 1 : 2 * m;
            double c, s;
            unit_root(k * n, N, &c, &s);
            const double wk = (k == 0 || 2 * k == N) ? 1.0 : 2.0;
            return wk * (sub < 2 ? c : s);
        });
    if (rc == BRV_OK) {
        const double gw = p->norm / N;
        std::vector<float4> wt(Q);
        for (int n = 0; n < Q; ++n) {
            wt[n].x = (float)(p->window[n] * gw);
            wt[n].y = (float)(p->window[Hf - n] * gw);
            wt[n].z = (float)(p->window[Hf + n] * gw);
            wt[n].w = n ? (float)(p->window[N - n] * gw) : 0.f;
        }
        fp->wq_inv = odd ? 0.f : (float)(p->window[Q] * gw);
        fp->w3q_inv = odd ? 0.f : (float)(p->window[3 * Q] * gw);
        // gradient of the forward transform: window / norm, and 1/2 against the basis' weight 2
        const double fg = 0.5 / p->norm;
        std::vector<float4> wf(Q);
        for (int n = 0; n < Q; ++n) {
            wf[n].x = (float)(p->window[n] * fg);
            wf[n].y = (float)(p->window[Hf - n] * fg);
            wf[n].z = (float)(p->window[Hf + n] * fg);
            wf[n].w = n ? (float)(p->window[N - n] * fg) : 0.f;
        }
        fp->wq_fg = odd ? 0.f : (float)(p->window[Q] * fg);
        fp->w3q_fg = odd ? 0.f : (float)(p->window[3 * Q] * fg);
        if (cudaMalloc((void**)&fp->wtab_fgrad, Q * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(fp->wtab_fgrad, wf.data(), Q * sizeof(float4), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded forward-gradient window table");
        if (cudaMalloc((void**)&fp->wtab_inv, Q * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(fp->wtab_inv, wt.data(), Q * sizeof(float4), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded inverse window table");
        const int H = p->hop;
        if (H == Q || (!odd && (H == 2 * Q || H == 4 * Q)))
            if ((size_t)TILE_M * (H + 1) * sizeof(float) <= (size_t)INV_REGION && H <= 256)
                fp->hq = H / Q;
    }
    if (rc == BRV_OK) {
        // interior hop blocks see all ceil(N / hop) frames: the envelope is periodic (any hop:
        // the iSTFT gradient runs on the forward kernel, which has no hop restriction)
        const int H = p->hop;
        std::vector<float> per(H);
        for (int off = 0; off < H; ++off) {
            double e = 0;
            for (int pos = off; pos < N; pos += H) e += p->window[pos] * p->window[pos];
            per[off] = e > 0 ? (float)(1.0 / e) : 0.f;
        }
        if (cudaMalloc((void**)&fp->env_per, H * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(fp->env_per, per.data(), H * sizeof(float), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded inverse periodic envelope");
        if (cudaDeviceGetAttribute(&fp->sm_count, cudaDevAttrMultiProcessorCount, p->device) !=
            cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "cudaDeviceGetAttribute(SM count)");
    }
    if (rc == BRV_OK) {
        const void* kernels[16] = {
            (const void*)istft_fold_kernel<1, false, false, true>, (const void*)istft_fold_kernel<1, true, false, true>,
            (const void*)istft_fold_kernel<1, false, true, true>, (const void*)istft_fold_kernel<1, true, true, true>,
            (const void*)istft_fold_kernel<1, false, false>, (const void*)istft_fold_kernel<1, true, false>,
            (const void*)istft_fold_kernel<2, false, false>, (const void*)istft_fold_kernel<2, true, false>,
            (const void*)istft_fold_kernel<4, false, false>, (const void*)istft_fold_kernel<4, true, false>,
            (const void*)istft_fold_kernel<1, false, true>, (const void*)istft_fold_kernel<1, true, true>,
            (const void*)istft_fold_kernel<2, false, true>, (const void*)istft_fold_kernel<2, true, true>,
            (const void*)istft_fold_kernel<4, false, true>, (const void*)istft_fold_kernel<4, true, true>};
        for (const void* k : kernels)
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     INV_SMEM_BYTES) != cudaSuccess)
                rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(istft_fold_kernel)");
        const void* tkernels[8] = {
            (const void*)istft_t_kernel<1, false, false, 64>, (const void*)istft_t_kernel<1, true, false, 64>,
            (const void*)istft_t_kernel<2, false, false, 64>, (const void*)istft_t_kernel<2, true, false, 64>,
            (const void*)istft_t_kernel<1, false, true, 64>, (const void*)istft_t_kernel<1, true, true, 64>,
            (const void*)istft_t_kernel<2, false, true, 64>, (const void*)istft_t_kernel<2, true, true, 64>};
        for (const void* k : tkernels)
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ItLayout<64>::SMEM_BYTES) != cudaSuccess)
                rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(istft_t_kernel)");
        const void* tkernels32[8] = {
            (const void*)istft_t_kernel<1, false, false, 32>, (const void*)istft_t_kernel<2, false, false, 32>,
            (const void*)istft_t_kernel<1, false, true, 32>, (const void*)istft_t_kernel<2, false, true, 32>,
            (const void*)istft_t_kernel<1, true, false, 32>, (const void*)istft_t_kernel<2, true, false, 32>,
            (const void*)istft_t_kernel<1, true, true, 32>, (const void*)istft_t_kernel<2, true, true, 32>};
        for (const void* k : tkernels32)
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ItLayout<32>::SMEM_BYTES) != cudaSuccess)
                rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(istft_t_kernel<32>)");
        const void* tkernels_dup[4] = {
            (const void*)istft_t_kernel<2, false, false, 32, false, true>, (const void*)istft_t_kernel<2, true, false, 32, false, true>,
            (const void*)istft_t_kernel<2, false, true, 32, false, true>, (const void*)istft_t_kernel<2, true, true, 32, false, true>};
        for (const void* k : tkernels_dup)
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ItLayout<32>::SMEM_BYTES) != cudaSuccess)
                rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(istft_t_kernel<dup>)");
        const void* tkernels_odd[4] = {
            (const void*)istft_t_kernel<1, false, false, 32, true>, (const void*)istft_t_kernel<1, true, false, 32, true>,
            (const void*)istft_t_kernel<1, false, true, 32, true>, (const void*)istft_t_kernel<1, true, true, 32, true>};
        for (const void* k : tkernels_odd)
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ItLayout<32, true>::SMEM_BYTES) != cudaSuccess)
                rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(istft_t_kernel<odd>)");
    }
    if (rc != BRV_OK) {
        free_fold(fp);
        return rc == BRV_ERR_UNSUPPORTED ? BRV_OK : rc;
    }
    p->fold = fp;
    return BRV_OK;
}

void brv_fold_plan_free(brv_stft_plan* p) {
    free_fold((FoldPlan*)p->fold);
    p->fold = nullptr;
}

// Transposed strip kernel (brv_fold_t.cuh): every SM walks the same number of frames in <= 64-frame
// tiles, the epilogue of a tile runs under the build + MMA of the next one.  Measured (B200):
// 128 x 8 s at 510 / 128 compressed 210 -> 181 us, 2048 x 4 s at 256 / 128 625 -> 530 us, but
// 64 x 4 s at 512 / 128 40 -> 44 us: with only a few tiles per SM its longer pipeline fill
// loses to one CTA per tile.  Variant 5 forces it, variant 4 disables it.
static bool fold_forward_uses_t(const brv_stft_plan* p, int64_t n_sig, int64_t n_frames, int origin) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    if (origin < 0) origin = brv_left(p);
    const int shift = p->hop % 4 == 0 ? (4 - (origin & 3)) & 3 : 0;
    if (ft_tile_frames(p->n_fft, p->hop, shift, 32) < 16) return false;
    return g_brv_fold_variant == 5 || g_brv_fold_variant == 8 ||
           (g_brv_fold_variant == 0 && n_sig * n_frames >= 512LL * fp->sm_count);
}

static int fold_forward_launch(const brv_stft_plan* p, FoldFwdParams prm, bool compress,
                               int64_t n_sig, int64_t n_frames, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    prm.n_frames = n_frames;
    prm.n_fft = p->n_fft;
    prm.hop = p->hop;
    prm.n_bins = p->n_bins;
    prm.q = fp->q;
    prm.odd = fp->odd;
    if (prm.origin < 0) prm.origin = brv_left(p);
    // 16-byte aligned span when the hop allows it (n_fft / 2 = 255 would otherwise force scalar loads)
    prm.shift = p->hop % 4 == 0 ? (4 - (prm.origin & 3)) & 3 : 0;
    prm.tmem_cols = fp->tmem_cols;
    prm.basis_scale_inv = fp->fwd.scale_inv;
    if (fold_forward_uses_t(p, n_sig, n_frames, prm.origin)) {
        // variant 8: 32-frame tiles, four of them resident in TMEM (shorter pipeline fill)
        const bool nf32 = g_brv_fold_variant == 8;
        const int nf = ft_tile_frames(p->n_fft, p->hop, prm.shift, nf32 ? 32 : 64);
        {
            prm.rows = nf;
            prm.tiles_per_signal = 0;
            prm.total_tiles = n_sig * n_frames;              // columns (frames), not tiles
            BRV_REQUIRE(prm.total_tiles < (1LL << 31), "too many frames (%lld)", (long long)prm.total_tiles);
            const int64_t want = brv_ceil_div(prm.total_tiles, 16);
            const unsigned ctas = (unsigned)(want < fp->sm_count ? want : fp->sm_count);
            if (nf32) {
                if (compress)
                    stft_t_kernel<true, 32><<<ctas, FT_THREADS, FtLayout<32>::SMEM_BYTES, st>>>(fp->fwd.map, prm);
                else
                    stft_t_kernel<false, 32><<<ctas, FT_THREADS, FtLayout<32>::SMEM_BYTES, st>>>(fp->fwd.map, prm);
            } else if (compress)
                stft_t_kernel<true, 64><<<ctas, FT_THREADS, FtLayout<64>::SMEM_BYTES, st>>>(fp->fwd.map, prm);
            else
                stft_t_kernel<false, 64><<<ctas, FT_THREADS, FtLayout<64>::SMEM_BYTES, st>>>(fp->fwd.map, prm);
            BRV_LAUNCH_CHECK("stft_t_kernel");
            return BRV_OK;
        }
    }
    int rows = (SPAN_MAX - p->n_fft - prm.shift) / p->hop + 1;
    if (rows > TILE_M) rows = TILE_M;
    if (rows > n_frames) rows = (int)n_frames;
    prm.rows = rows;
    prm.tiles_per_signal = (int)brv_ceil_div(n_frames, rows);
    prm.tmem_cols = fp->tmem_cols;
    prm.basis_scale_inv = fp->fwd.scale_inv;
    const int64_t grid = n_sig * prm.tiles_per_signal;
    BRV_REQUIRE(grid < (1LL << 31), "too many tiles (%lld)", (long long)grid);
    prm.total_tiles = grid;
    // Both kernels are bound by shared-memory bandwidth per tile (operand build + MMA operand
    // reads + epilogue transpose).  The persistent two-pass kernel overlaps the phases of
    // consecutive tiles and wins once every SM has a few tiles to pipeline (cfg5: 760 -> 620 us);
    // with one or two tiles per SM its fill / drain latency loses to one CTA per tile
    // (cfg2: 46 vs 60 us).  Variant 2 / 3 force one or the other for A/B runs.
    // For Q = 128 with full 128-row tiles (n_fft 512 / 510 at hop <= 128) the persistent kernel's
    // two operand-build passes per tile cost more than its overlap wins at any tile count
    // (128 x 8 s: 163 vs 187 us plain, 224 vs 276 us compressed): one CTA per tile there too.
    const bool one_per_cta =
        g_brv_fold_variant == 2 ||
        (g_brv_fold_variant != 3 &&
         (grid < 3LL * fp->sm_count || (fp->q == MAX_Q && prm.rows == TILE_M)));
    if (one_per_cta) {
        if (prm.post_expo != 0.f)
            stft_fold_kernel<true><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(fp->fwd.map, prm);
        else
            stft_fold_kernel<false><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(fp->fwd.map, prm);
        BRV_LAUNCH_CHECK("stft_fold_kernel");
        return BRV_OK;
    }
    const unsigned ctas = (unsigned)(grid < fp->sm_count ? grid : fp->sm_count);
    if (compress)
        stft_fold2_kernel<true><<<ctas, F2_THREADS, F2_SMEM_BYTES, st>>>(fp->fwd.map, prm);
    else
        stft_fold2_kernel<false><<<ctas, F2_THREADS, F2_SMEM_BYTES, st>>>(fp->fwd.map, prm);
    BRV_LAUNCH_CHECK("stft_fold2_kernel");
    return BRV_OK;
}

int brv_fold_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                          int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldFwdParams prm = {};
    prm.x = x;
    prm.x_stride = x_stride;
    prm.samples = samples;
    prm.out = reinterpret_cast<float*>(out);
    prm.wtab = fp->wtab;
    prm.wq = fp->wq;
    prm.w3q = fp->w3q;
    prm.wmax = fp->wmax;
    prm.post_scale = (float)p->scale;
    prm.post_expo = (float)(p->compression - 1.0);
    prm.in_mul = nullptr;
    prm.edge_scale = prm.dc_scale = 1.f;
    prm.origin = -1;
    return fold_forward_launch(p, prm, p->compression != 1.0, n_sig, n_frames, st);
}

bool brv_fold_inverse_supported(const brv_stft_plan* p) {
    return p->fold != nullptr && ((const FoldPlan*)p->fold)->hq != 0;
}

// 1 / (overlap-added squared window) on the trimmed output grid, cached per frame
// count (torch.istft divides by this envelope; NOLA was checked by the caller).
static int fold_inv_envelope(const brv_stft_plan* p, int64_t n_frames, int64_t out_len,
                             const float** out) {
    brv_stft_plan* mp = const_cast<brv_stft_plan*>(p);
    FoldPlan* fp = (FoldPlan*)p->fold;
    std::lock_guard<std::mutex> lock(mp->mu);
    auto it = fp->inv_env.find(n_frames);
    if (it == fp->inv_env.end()) {
        const int N = p->n_fft, H = p->hop;
        std::vector<float> host((size_t)out_len);
        for (int64_t i = 0; i < out_len; ++i) {
            const int64_t pos = i + N / 2;
            int64_t t_hi = pos / H;
            if (t_hi > n_frames - 1) t_hi = n_frames - 1;
            const int64_t t_lo = pos - N + 1 <= 0 ? 0 : (pos - N + H) / H;
            double e = 0;
            for (int64_t t = t_lo; t <= t_hi; ++t) {
                const double w = p->window[pos - t * H];
                e += w * w;
            }
            host[(size_t)i] = (float)(1.0 / e);
        }
        // (only the A/B variants 2-4 of the iSTFT gradient still use this table; entries are never
        //  freed while the plan lives: a captured CUDA graph may hold the pointer)
        BRV_REQUIRE(fp->inv_env.size() < 1024, "too many distinct frame counts for the tabulated envelope");
        float* dev = nullptr;
        BRV_CUDA(cudaMalloc((void**)&dev, (size_t)out_len * sizeof(float)));
        cudaError_t e = cudaMemcpy(dev, host.data(), (size_t)out_len * sizeof(float),
                                   cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(dev);
            return brv_fail_cuda(e, "cudaMemcpy(inverse envelope; first call for this frame "
                                    "count must not be inside a stream capture)");
        }
        it = fp->inv_env.emplace(n_frames, dev).first;
    }
    *out = it->second;
    return BRV_OK;
}

static int fold_inverse_launch(const brv_stft_plan* p, FoldInvParams prm, bool decomp,
                               int64_t n_sig, int64_t n_frames, int64_t out_len, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    prm.out_len = out_len;
    prm.n_frames = n_frames;
    prm.n_fft = p->n_fft;
    prm.hop = p->hop;
    prm.q = fp->q;
    prm.halo = 4 / fp->hq - 1;
    prm.adv = TILE_M - prm.halo;
    if (prm.origin < 0) prm.origin = brv_left(p);
    prm.n_blocks = (int)brv_ceil_div(prm.origin + out_len, p->hop);
    prm.tiles_per_signal =
        prm.n_blocks <= TILE_M ? 1 : 1 + (int)brv_ceil_div(prm.n_blocks - TILE_M, prm.adv);
    prm.tmem_cols = fp->tmem_cols;
    prm.basis_scale_inv = fp->inv.scale_inv;
    prm.env_per = fp->env_per;
    prm.wsq = p->window_sq;
    prm.total_tiles = n_sig * prm.tiles_per_signal;
    const bool frames_fast = prm.sb != 1;
    // Which kernel (measured on B200, tools/t_sweep.py):
    //   * the transposed strip kernel (istft_t_kernel, 32-frame tiles) for Q = 128 (n_fft 512-class)
    //     and for Q = 64 with hop = Q, at any launch size and for both spectrogram layouts;
    //     for Q = 64 with hop = 2Q its duplicated-lane flavour (2048 x 4 s: 1034 -> 856 us, bin-major 908 -> 860 us);
    //   * the one-tile-per-TMEM kernel for everything else: hop = 4Q, Q = 32, Q = 96.
    // Variants 6 / 7 force the strip kernel (64- / 32-frame tiles), variant 4 the tile kernel.
    const int64_t cols = n_sig * (int64_t)prm.n_blocks;
    bool use_t = false, nf32 = true;
    if ((fp->hq == 1 || (fp->hq == 2 && !fp->odd)) && cols < (1LL << 31)) {
        if (g_brv_fold_variant == 6 || g_brv_fold_variant == 7) {
            use_t = true;
            nf32 = g_brv_fold_variant == 7;
        } else if (g_brv_fold_variant == 0) {
            use_t = fp->q == 128 || fp->q == 64;
        }
    }
    if (use_t) {
        prm.total_tiles = cols;                              // columns (hop blocks), not tiles
        const int64_t want = brv_ceil_div(prm.total_tiles, 16);
        const unsigned ctas = (unsigned)(want < fp->sm_count ? want : fp->sm_count);
#define BRV_LAUNCH_INV_T(HQ_, FF_, NF_)                                                           \
    do {                                                                                          \
        if (decomp)                                                                               \
            istft_t_kernel<HQ_, FF_, true, NF_>                                                   \
                <<<ctas, IT_THREADS, ItLayout<NF_>::SMEM_BYTES, st>>>(fp->inv.map, prm);          \
        else                                                                                      \
            istft_t_kernel<HQ_, FF_, false, NF_>                                                  \
                <<<ctas, IT_THREADS, ItLayout<NF_>::SMEM_BYTES, st>>>(fp->inv.map, prm);          \
    } while (0)
        if (fp->odd) {                               // n_fft = 4Q - 2: 32-frame tiles only
            if (decomp) {
                if (frames_fast)
                    istft_t_kernel<1, true, true, 32, true>
                        <<<ctas, IT_THREADS, ItLayout<32, true>::SMEM_BYTES, st>>>(fp->inv.map, prm);
                else
                    istft_t_kernel<1, false, true, 32, true>
                        <<<ctas, IT_THREADS, ItLayout<32, true>::SMEM_BYTES, st>>>(fp->inv.map, prm);
            } else {
                if (frames_fast)
                    istft_t_kernel<1, true, false, 32, true>
                        <<<ctas, IT_THREADS, ItLayout<32, true>::SMEM_BYTES, st>>>(fp->inv.map, prm);
                else
                    istft_t_kernel<1, false, false, 32, true>
                        <<<ctas, IT_THREADS, ItLayout<32, true>::SMEM_BYTES, st>>>(fp->inv.map, prm);
            }
        } else if (fp->hq == 2 && fp->q == 64 && nf32) {
            // duplicated accumulator lanes: all eight epilogue warps work although Q = 64
#define BRV_LAUNCH_INV_DUP(FF_, DC_)                                                              \
    istft_t_kernel<2, FF_, DC_, 32, false, true>                                                  \
        <<<ctas, IT_THREADS, ItLayout<32>::SMEM_BYTES, st>>>(fp->inv.map, prm)
            if (decomp) { if (frames_fast) BRV_LAUNCH_INV_DUP(true, true); else BRV_LAUNCH_INV_DUP(false, true); }
            else { if (frames_fast) BRV_LAUNCH_INV_DUP(true, false); else BRV_LAUNCH_INV_DUP(false, false); }
#undef BRV_LAUNCH_INV_DUP
        } else if (fp->hq == 1) {
            if (frames_fast) { if (nf32) BRV_LAUNCH_INV_T(1, true, 32); else BRV_LAUNCH_INV_T(1, true, 64); }
            else if (nf32) BRV_LAUNCH_INV_T(1, false, 32);
            else BRV_LAUNCH_INV_T(1, false, 64);
        } else {
            if (frames_fast) { if (nf32) BRV_LAUNCH_INV_T(2, true, 32); else BRV_LAUNCH_INV_T(2, true, 64); }
            else if (nf32) BRV_LAUNCH_INV_T(2, false, 32);
            else BRV_LAUNCH_INV_T(2, false, 64);
        }
#undef BRV_LAUNCH_INV_T
        BRV_LAUNCH_CHECK("istft_t_kernel");
        return BRV_OK;
    }
    // large launches: one strip of hop blocks per CTA (TileStrip; cfg5-size: 5 % faster); small
    // ones keep the fixed tiles (measured: strips cost 60 -> 73 us on 256 tiles)
    prm.total_cols = n_sig * (int64_t)prm.n_blocks;
    BRV_REQUIRE(prm.total_cols < (1LL << 31) && prm.total_tiles < (1LL << 31), "too many hop blocks (%lld)",
                (long long)prm.total_cols);
    prm.strips = prm.total_tiles >= 8LL * fp->sm_count ? 1 : 0;
    const unsigned grid = (unsigned)(prm.total_tiles < fp->sm_count ? prm.total_tiles : fp->sm_count);
    {   // the bytes the (signal, frame, bin) view itself covers: bulk row copies stay inside them
        const int64_t f_in = p->n_fft / 2 + 1;
        // (16 KB chunks only: with the 8 KB chunks of n_fft = 256 three copies in flight do not
        //  cover the bulk-copy latency and the scouts fall behind the builders, cfg5 510 -> 542 us)
        prm.bulk = (!frames_fast && prm.sf == f_in && f_in >= 256 && (prm.ss >= 0 || n_sig == 1)) ? 1 : 0;
        prm.spec_lo = (uintptr_t)prm.spec;
        prm.spec_hi = prm.spec_lo +
                      (uintptr_t)(((n_sig - 1) * (prm.ss > 0 ? prm.ss : 0) + (n_frames - 1) * prm.sf + f_in) * 8);
    }
#define BRV_LAUNCH_INV(HQ_, FF_)                                                                  \
    do {                                                                                          \
        if (decomp)                                                                               \
            istft_fold_kernel<HQ_, FF_, true><<<grid, INV_THREADS, INV_SMEM_BYTES, st>>>(fp->inv.map, prm); \
        else                                                                                      \
            istft_fold_kernel<HQ_, FF_, false><<<grid, INV_THREADS, INV_SMEM_BYTES, st>>>(fp->inv.map, prm); \
    } while (0)
#define BRV_LAUNCH_INV_ODD(FF_)                                                                   \
    do {                                                                                          \
        if (decomp)                                                                               \
            istft_fold_kernel<1, FF_, true, true><<<grid, INV_THREADS, INV_SMEM_BYTES, st>>>(fp->inv.map, prm); \
        else                                                                                      \
            istft_fold_kernel<1, FF_, false, true><<<grid, INV_THREADS, INV_SMEM_BYTES, st>>>(fp->inv.map, prm); \
    } while (0)
    if (fp->odd) {
        if (frames_fast) BRV_LAUNCH_INV_ODD(true); else BRV_LAUNCH_INV_ODD(false);
    } else
    switch (fp->hq * 2 + (frames_fast ? 1 : 0)) {
        case 2: BRV_LAUNCH_INV(1, false); break;
        case 3: BRV_LAUNCH_INV(1, true); break;
        case 4: BRV_LAUNCH_INV(2, false); break;
        case 5: BRV_LAUNCH_INV(2, true); break;
        case 8: BRV_LAUNCH_INV(4, false); break;
        default: BRV_LAUNCH_INV(4, true); break;
    }
#undef BRV_LAUNCH_INV
#undef BRV_LAUNCH_INV_ODD
    BRV_LAUNCH_CHECK("istft_fold_kernel");
    return BRV_OK;
}

int brv_fold_istft(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb, int64_t sf,
                   int64_t n_sig, int64_t n_frames, int64_t out_len, float* y, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldInvParams prm = {};
    prm.no_env = 0;
    prm.spec = X;
    prm.ss = ss;
    prm.sb = sb;
    prm.sf = sf;
    prm.pre_scale = (float)(1.0 / p->scale);
    prm.pre_expo = (float)(1.0 / p->compression - 1.0);
    prm.y = y;
    prm.wtab = fp->wtab_inv;
    prm.wq = fp->wq_inv;
    prm.w3q = fp->w3q_inv;
    prm.edge_gain = prm.dc_gain = 1.f;
    prm.origin = -1;
    return fold_inverse_launch(p, prm, p->compression != 1.0, n_sig, n_frames, out_len, st);
}

// Gradient of STFT.forward w.r.t. the waveform (SURVEY 8a', compression_factor == 1): the
// inverse kernel with every bin weighing 1, window * scale / sqrt(sum w^2), no envelope, and
// the overlap-added frames cut to the `samples` the forward transform was given.
int brv_fold_stft_grad(const brv_stft_plan* p, const float2* gX, int64_t ss, int64_t sb, int64_t sf,
                       int64_t n_sig, int64_t n_frames, int64_t samples, float* gx,
                       cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldInvParams prm = {};
    prm.no_env = 1;
    prm.spec = gX;
    prm.ss = ss;
    prm.sb = sb;
    prm.sf = sf;
    prm.pre_scale = (float)p->scale;
    prm.pre_expo = 0.f;
    prm.y = gx;
    prm.wtab = fp->wtab_fgrad;
    prm.wq = fp->wq_fg;
    prm.w3q = fp->w3q_fg;
    prm.edge_gain = prm.dc_gain = 2.f;
    prm.origin = -1;
    return fold_inverse_launch(p, prm, false, n_sig, n_frames, samples, st);
}

// Gradient of STFT.backward w.r.t. its spectrogram input (SURVEY 8a'): the forward kernel
// on gy / envelope, window * sqrt(sum w^2) / N, Hermitian weights (1, 2, ..., 2, 1), 1 / scale.
int brv_fold_istft_grad(const brv_stft_plan* p, const float* gy, int64_t n_sig, int64_t n_frames,
                        int64_t out_len, float2* gX, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldFwdParams prm = {};
    if (fold_forward_uses_t(p, n_sig, n_frames, -1)) {
        prm.grad_env = 1;                          // the transposed kernel sums the envelope itself
        prm.env_per = fp->env_per;
        prm.wsq = p->window_sq;
    } else {
        int rc = fold_inv_envelope(p, n_frames, out_len, &prm.in_mul);
        if (rc != BRV_OK) return rc;
    }
    prm.x = gy;
    prm.x_stride = out_len;
    prm.samples = out_len;
    prm.out = reinterpret_cast<float*>(gX);
    prm.wtab = fp->wtab_grad;
    prm.wq = fp->wq_g;
    prm.w3q = fp->w3q_g;
    prm.wmax = fp->wmax_g;
    prm.post_scale = (float)(1.0 / p->scale);
    prm.post_expo = 0.f;
    prm.edge_scale = prm.dc_scale = 0.5f;
    prm.origin = -1;
    return fold_forward_launch(p, prm, false, n_sig, n_frames, st);
}

// ---- ConvSTFT (brever/modules/stft.py:201-319) on the same kernels ------------------------
// ConvSTFT.forward is a strided convolution with the rows k = 0..L/2 of the DFT matrix times
// the window, the DC row divided by sqrt(2), all divided by 0.5 L / sqrt(H) when normalised
// (stft.py:221-235): the folded forward kernel with frames starting L - H samples before
// t * hop (ConvSTFT.pad, stft.py:305-315, instead of torch.stft's n_fft / 2), a DC gain of
// 1 / sqrt(2) and `gain` folded into the post-compression scale ((g X)|g X|^(c-1) = g^c X|X|^(c-1)).
bool brv_fold_conv_supported(const brv_stft_plan* p) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    return fp && !fp->odd && fp->hq != 0 && !p->normalized && p->n_fft == p->frame_length &&
           p->frame_length >= p->hop;
}

int brv_fold_conv_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                          int64_t x_stride, double gain, float2* out, int64_t n_frames,
                          cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldFwdParams prm = {};
    prm.x = x;
    prm.x_stride = x_stride;
    prm.samples = samples;
    prm.out = reinterpret_cast<float*>(out);
    prm.wtab = fp->wtab;
    prm.wq = fp->wq;
    prm.w3q = fp->w3q;
    prm.wmax = fp->wmax;
    prm.post_scale = (float)(p->scale * pow(gain, p->compression));
    prm.post_expo = (float)(p->compression - 1.0);
    prm.in_mul = nullptr;
    prm.edge_scale = 1.f;
    prm.dc_scale = (float)sqrt(0.5);
    prm.origin = p->frame_length - p->hop;
    return fold_forward_launch(p, prm, p->compression != 1.0, n_sig, n_frames, st);
}

// ConvSTFT.backward = conv_transpose1d with the same filters (stft.py:281-300): the adjoint of
// the forward map, no envelope division, trimmed by L - H on both sides.  The folded inverse
// kernel in its gradient configuration (every bin weighs 1; DC 1 / sqrt(2)), after / scale and
// |X|^(1/c - 1) decompression in the loaders; `gain` (1 / normalisation, squared when the
// filters are not normalised, stft.py:291-292) enters before the decompression as gain^c.
int brv_fold_conv_backward(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb,
                           int64_t sf, int64_t n_sig, int64_t n_frames, int64_t out_len,
                           double gain, float* y, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldInvParams prm = {};
    prm.no_env = 1;
    prm.spec = X;
    prm.ss = ss;
    prm.sb = sb;
    prm.sf = sf;
    prm.pre_scale = (float)(pow(gain, p->compression) / p->scale);
    prm.pre_expo = (float)(1.0 / p->compression - 1.0);
    prm.y = y;
    prm.wtab = fp->wtab_fgrad;
    prm.wq = fp->wq_fg;
    prm.w3q = fp->w3q_fg;
    prm.edge_gain = 2.f;
    prm.dc_gain = (float)sqrt(2.0);
    prm.origin = p->frame_length - p->hop;
    return fold_inverse_launch(p, prm, p->compression != 1.0, n_sig, n_frames, out_len, st);
}

#ifdef BRV_PHASE_TIMING
extern "C" int brv_debug_t_times(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_brv_t_ts, sizeof(unsigned long long) * (n < 160 ? n : 160));
}
extern "C" int brv_debug_t_waits(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_brv_t_wait, sizeof(unsigned long long) * (n < 32 ? n : 32));
}
#endif
