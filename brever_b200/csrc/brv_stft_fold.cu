// Symmetry-folded tensor-core STFT for sm_100a (tcgen05 + TMEM + TMA), fp32-grade.
//
// STFT.forward (brever/modules/stft.py:59-89) for one-sided transforms with
// n_fft in {128, 256, 384, 512}.  A real DFT of length N = 4Q splits, by the
// even/odd symmetries of cos and sin about n = N/2 and n = N/4 (two radix-2
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
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "brv_common.cuh"
#include "brv_tc_ptx.cuh"

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
constexpr int SMEM_BMAX = (SPAN_ALLOC / 32) * 4;               // 2100
constexpr int SMEM_BYTES = 1024 + SMEM_STAGES + SMEM_SPAN + SMEM_WTAB + SMEM_ROWINFO + 2112;

struct FoldFwdParams {
    const float* x;
    int64_t x_stride, samples;
    float* out;                  // (sig, T, n_bins) complex64 as floats
    int64_t n_frames;
    const float4* wtab;          // Q entries: (w[n], w[N/2-n], w[N/2+n], w[N-n]) / norm
    int n_fft, hop, n_bins, q;
    int rows;                    // frames per tile (<= 128, span fits SPAN_MAX)
    int tiles_per_signal;
    int tmem_cols;
    float wq, w3q;               // w[Q] / norm, w[3Q] / norm
    float wmax;                  // max |w| / norm
    float basis_scale_inv;
    float post_scale, post_expo;
};

__device__ __forceinline__ void split_store(uint8_t* hi_ptr, uint8_t* lo_ptr, float v0, float v1) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__half2*>(hi_ptr) = h;
    *reinterpret_cast<__half2*>(lo_ptr) = l;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
stft_fold_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
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
        const int64_t span0 = t0 * H - Hf;         // first sample of the span (may be < 0)
        const int span_len = (rows_eff - 1) * H + N;
        const int span_pad = (span_len + 31) & ~31;

        for (int j = bt; j < Q; j += BUILDER_THREADS) wtab[j] = __ldg(p.wtab + j);

        // ---- stage the sample span + per-32-sample maxima ------------------------
        const bool vec = ((((uintptr_t)xs) & 15) == 0) && ((span0 & 3) == 0);
        for (int wb = bw * 128; wb < span_pad; wb += BUILDERS * 128) {
            const int i = wb + lane * 4;
            const int64_t idx = span0 + i;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < span_pad) {
                if (vec && idx >= 0 && idx + 4 <= p.samples) {
                    f = __ldg(reinterpret_cast<const float4*>(xs + idx));
                } else {
                    if (idx >= 0 && idx < p.samples) f.x = __ldg(xs + idx);
                    if (idx + 1 >= 0 && idx + 1 < p.samples) f.y = __ldg(xs + idx + 1);
                    if (idx + 2 >= 0 && idx + 2 < p.samples) f.z = __ldg(xs + idx + 2);
                    if (idx + 3 >= 0 && idx + 3 < p.samples) f.w = __ldg(xs + idx + 3);
                }
                *reinterpret_cast<float4*>(span + i) = f;
            }
            float m = fmaxf(fmaxf(finite_abs(f.x), finite_abs(f.y)),
                            fmaxf(finite_abs(f.z), finite_abs(f.w)));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            if ((lane & 7) == 0 && i < span_pad) bmax[i >> 5] = __float_as_uint(m);
        }
        named_bar_sync(1, BUILDER_THREADS);

        // ---- per-row power-of-two scale from a bound on the folded magnitudes ----
        if (bt < rows_eff) {
            const int b0 = (bt * H) >> 5, b1 = (bt * H + N - 1) >> 5;
            uint32_t mx = 0u;
            for (int b = b0; b <= b1; ++b) mx = max(mx, bmax[b]);
            rowinfo[bt].x = row_scale(4.f * p.wmax * __uint_as_float(mx));
        }
        named_bar_sync(1, BUILDER_THREADS);

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
                    if (row >= rows_eff) continue;
                    const float* fr = span + row * H;
                    const float a0 = fr[n0] * w0.x, a1 = fr[n0 + 1] * w1.x;
                    const float b0 = fr[Hf - n0] * w0.y, b1 = fr[Hf - n0 - 1] * w1.y;
                    const float c0 = fr[Hf + n0] * w0.z, c1 = fr[Hf + n0 + 1] * w1.z;
                    const float d0 = n0 ? fr[N - n0] * w0.w : 0.f, d1 = fr[N - n0 - 1] * w1.w;
                    float u0, u1, v0, v1;          // the pair's two folded sequences at n0, n0+1
                    if (pair == 0) {
                        const float s0 = a0 + d0, r0 = b0 + c0, s1 = a1 + d1, r1 = b1 + c1;
                        u0 = s0 + r0; u1 = s1 + r1;            // ee
                        v0 = s0 - r0; v1 = s1 - r1;            // eo
                        nyq[i] += u0 - u1;                     // (-1)^n ee[n], n0 even
                    } else {
                        const float s0 = a0 - d0, r0 = b0 - c0, s1 = a1 - d1, r1 = b1 - c1;
                        u0 = s0 - r0; u1 = s1 - r1;            // oe
                        v0 = s0 + r0; v1 = s1 + r1;            // oo
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

        // ---- epilogue ---------------------------------------------------------------
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int hsel = bw >> 2;                  // which half of the m range
        const int row = q * 32 + lane;
        const bool live = row < rows_eff;
        float g0 = 0.f, eeq = 0.f, ooq = 0.f, ny = 0.f;
        if (live) {
            const float2 ri = rowinfo[row];
            g0 = p.basis_scale_inv / ri.x;
            const float xq = span[row * H + Q] * p.wq, x3q = span[row * H + 3 * Q] * p.w3q;
            eeq = xq + x3q;
            ooq = xq - x3q;
            ny = ri.y + eeq;                       // Q is even: (-1)^Q = +1
        }
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
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
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float sg = (j & 1) ? -1.f : 1.f;        // m0 is even
                float re_e = __uint_as_float(r0[j]) * g0 + sg * eeq;
                float re_o = __uint_as_float(r1[j]) * g0;
                float im_e = __uint_as_float(r2[j]) * g0;
                float im_o = __uint_as_float(r3[j]) * g0 - sg * ooq;
                if (p.post_expo != 0.f) {
                    compress(re_e, im_e, p.post_expo);
                    compress(re_o, im_o, p.post_expo);
                }
                *reinterpret_cast<float4*>(stg + lane * EPI_PITCH + 4 * j) =
                    make_float4(re_e * p.post_scale, im_e * p.post_scale, re_o * p.post_scale,
                                im_o * p.post_scale);
            }
            __syncwarp();
            float* ocol = obase + 4 * m0 + 2 * lane;
#pragma unroll 4
            for (int rr = 0; rr < rows_w; ++rr) {
                const float2 v = *reinterpret_cast<const float2*>(stg + rr * EPI_PITCH + 2 * lane);
                *reinterpret_cast<float2*>(ocol + (int64_t)rr * pitch) = v;
            }
            __syncwarp();
        }
        if (hsel == 1 && live) {                   // Nyquist bin: purely real
            float v = ny;
            if (p.post_expo != 0.f) v = compress_real(v, p.post_expo);
            *reinterpret_cast<float2*>(obase + (int64_t)lane * pitch + 2 * Hf) =
                make_float2(v * p.post_scale, 0.f);
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

struct FoldBasis {
    __half* data = nullptr;      // [plane 2][sub 4][m Q][n Q]
    CUtensorMap map;
    float scale_inv = 1.f;
};
struct FoldPlan {
    FoldBasis fwd;
    float4* wtab = nullptr;      // Q permuted, normalised window entries
    float wq = 0.f, w3q = 0.f, wmax = 0.f;
    int q = 0, tmem_cols = 0;
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
    cudaFree(fp->wtab);
    delete fp;
}

}  // namespace

bool brv_fold_supported(const brv_stft_plan* p) { return p->fold != nullptr; }

int brv_fold_plan_init(brv_stft_plan* p) {
    const int N = p->n_fft;
    if (!p->onesided || N % 128 != 0 || N > 512) return BRV_OK;
    if (!encode_tiled()) return BRV_OK;
    const int Q = N / 4, Hf = N / 2;
    FoldPlan* fp = new FoldPlan();
    fp->q = Q;
    fp->tmem_cols = N <= 128 ? 128 : (N <= 256 ? 256 : 512);
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
        fp->wq = (float)(p->window[Q] / p->norm);
        fp->w3q = (float)(p->window[3 * Q] / p->norm);
        fp->wmax = (float)wmax;
        if (cudaMalloc((void**)&fp->wtab, Q * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(fp->wtab, wt.data(), Q * sizeof(float4), cudaMemcpyHostToDevice) !=
                cudaSuccess)
            rc = brv_fail_cuda(cudaGetLastError(), "folded window table");
    }
    if (rc == BRV_OK &&
        cudaFuncSetAttribute(stft_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess)
        rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(stft_fold_kernel)");
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

int brv_fold_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                          int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st) {
    const FoldPlan* fp = (const FoldPlan*)p->fold;
    FoldFwdParams prm = {};
    prm.x = x;
    prm.x_stride = x_stride;
    prm.samples = samples;
    prm.out = reinterpret_cast<float*>(out);
    prm.n_frames = n_frames;
    prm.wtab = fp->wtab;
    prm.n_fft = p->n_fft;
    prm.hop = p->hop;
    prm.n_bins = p->n_bins;
    prm.q = fp->q;
    int rows = (SPAN_MAX - p->n_fft) / p->hop + 1;
    if (rows > TILE_M) rows = TILE_M;
    if (rows > n_frames) rows = (int)n_frames;
    prm.rows = rows;
    prm.tiles_per_signal = (int)brv_ceil_div(n_frames, rows);
    prm.tmem_cols = fp->tmem_cols;
    prm.wq = fp->wq;
    prm.w3q = fp->w3q;
    prm.wmax = fp->wmax;
    prm.basis_scale_inv = fp->fwd.scale_inv;
    prm.post_scale = (float)p->scale;
    prm.post_expo = (float)(p->compression - 1.0);
    const int64_t grid = n_sig * prm.tiles_per_signal;
    BRV_REQUIRE(grid < (1LL << 31), "too many tiles (%lld)", (long long)grid);
    stft_fold_kernel<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(fp->fwd.map, prm);
    BRV_LAUNCH_CHECK("stft_fold_kernel");
    return BRV_OK;
}
