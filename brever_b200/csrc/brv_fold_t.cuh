// Transposed, strip-scheduled variants of the symmetry-folded STFT kernels (included by
// brv_stft_fold.cu inside its anonymous namespace, after the parameter structs).
//
// Same arithmetic as stft_fold_kernel / istft_fold_kernel (four Q x Q contractions, fp16
// hi / lo operands with per-frame power-of-two scales, three products per k-step), but the
// roles of the MMA operands are swapped:
//
//     D[n or m (TMEM lane), frame (TMEM column)] = basis[row, k] . data[frame, k]^T
//
// i.e. the trigonometric basis is the M-side operand (M = 128 lanes, rows >= Q are never
// read back) and the frames are the N side.  Three things follow:
//   * a tile is up to 64 frames = 4 x 64 accumulator columns, so TWO tiles fit in the 512
//     TMEM columns: the tensor core works on tile i+1 while the epilogue drains tile i, and the
//     operand builders never wait for an epilogue (the one-tile-per-TMEM kernels serialise
//     build -> MMA -> epilogue on every tile);
//   * a thread of the epilogue owns one bin pair (forward) or one sample offset (inverse) and
//     walks along the frames: the forward epilogue stores 16 contiguous bytes per thread and
//     512 per warp straight from registers (no shared-memory transpose), the inverse epilogue
//     overlap-adds consecutive frames in registers (no lane rotation, no spill slots);
//   * the frame axis is cut into strips, not tiles: CTA c of P owns the global (signal, frame)
//     range [c G / P, (c+1) G / P), walked in tiles of <= 64 frames that never cross a signal,
//     so every SM gets the same number of frames (256 tiles on 148 SMs ran as 2 + 1 before).

constexpr int T_NF = 64;                                   // frames per tile (UMMA N)
constexpr int T_DATA_TILE = T_NF * BK * 2;                 // 4 KB: one (sub-GEMM, plane) data block
constexpr int T_STAGE_BASIS = 4 * SUB_TILE;                // 32 KB: 2 sub-GEMMs x {hi, lo} basis boxes
constexpr int T_STAGE_BYTES = T_STAGE_BASIS + 4 * T_DATA_TILE;   // 48 KB
constexpr int T_STAGES = 3;
constexpr int T_SMEM_STAGES = T_STAGES * T_STAGE_BYTES;    // 144 KB
constexpr int T_TMEM_COLS = 512;                           // 2 buffers x 4 accumulators x 64 frames

// the strip of CTA `cta` out of `ctas` over `total` columns, walked in tiles
struct StripIter {
    int64_t g, g1, per_signal;
    int nf;
    __device__ StripIter(int64_t total, int64_t per_signal_, int nf_, int cta, int ctas)
        : g(total * cta / ctas), g1(total * (cta + 1) / ctas), per_signal(per_signal_), nf(nf_) {}
    __device__ bool next(int64_t& sig, int64_t& c0, int& ncols) {
        if (g >= g1) return false;
        sig = g / per_signal;
        c0 = g - sig * per_signal;
        const int64_t m = min((int64_t)nf, min(per_signal - c0, g1 - g));
        ncols = (int)m;
        g += m;
        return true;
    }
};

// 1 / s for a power of two s (exact)
__device__ __forceinline__ float pow2_inv(float s) {
    return __uint_as_float(0x7f000000u - __float_as_uint(s));
}
__device__ __forceinline__ void bulk_g2s_t(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// forward: STFT.forward (brever/modules/stft.py:59-89)
//
//   warp 0        TMA producer: basis k-chunks (3-stage mbarrier ring)
//   warp 1        TMEM owner + MMA issuer
//   warps 2-3     span loaders: bulk-copy the NEXT tile's sample span into the other span
//                 buffer, per-32-sample maxima, per-frame power-of-two scales and rank-1 terms
//   warps 4-11    epilogue: thread = bin pair (TMEM lane), 32 frames per warp
//   warps 12-19   operand builders: window, fold, scale, split (from the staged span)
constexpr int FT_THREADS = 640;
constexpr int FT_LOADER_WARP0 = 2, FT_LOADER_THREADS = 64;
constexpr int FT_EPI_WARP0 = 4, FT_EPI_WARPS = 8;
constexpr int FT_BUILD_WARP0 = 12, FT_BUILD_WARPS = 8;
constexpr int FT_SPAN = (T_NF - 1) * 128 + 512 + 64;        // floats per span buffer (8640)
constexpr int FT_BMAX = FT_SPAN / 32;                       // 270
constexpr int FT_OFF_SPAN = T_SMEM_STAGES;
constexpr int FT_OFF_WTAB = FT_OFF_SPAN + 2 * FT_SPAN * 4;
constexpr int FT_OFF_ROWINFO = FT_OFF_WTAB + SMEM_WTAB;      // 2 slots x 64 float4
constexpr int FT_OFF_BMAX = FT_OFF_ROWINFO + 2 * T_NF * 16;
constexpr int FT_SMEM_BYTES = 1024 + FT_OFF_BMAX + 2 * FT_BMAX * 4;
static_assert(FT_SMEM_BYTES <= 227 * 1024, "transposed forward kernel shared memory");

// frames per tile for a given geometry: the tile's span must fit one span buffer
static inline int ft_tile_frames(int n_fft, int hop, int shift) {
    int nf = (FT_SPAN - 32 - n_fft - shift) / hop + 1;
    return nf > T_NF ? T_NF : nf;
}

template <bool COMPRESS>
__global__ void __launch_bounds__(FT_THREADS, 1)
stft_t_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[T_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[T_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];     // MMA -> epilogue
    __shared__ __align__(8) uint64_t tmem_empty[2];    // epilogue -> MMA
    __shared__ __align__(8) uint64_t span_full[2];     // loaders -> builders (span + row info)
    __shared__ __align__(8) uint64_t span_empty[2];    // builders -> loaders
    __shared__ __align__(8) uint64_t ri_full[2];       // builders -> epilogue (Nyquist sums written)
    __shared__ __align__(8) uint64_t ri_empty[2];      // epilogue -> loaders (row info slot)
    __shared__ __align__(8) uint64_t span_landed[2];   // bulk copy of the span
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* span2 = reinterpret_cast<float*>(stages + FT_OFF_SPAN);
    float4* wtab = reinterpret_cast<float4*>(stages + FT_OFF_WTAB);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + FT_OFF_ROWINFO);
    uint32_t* bmax2 = reinterpret_cast<uint32_t*>(stages + FT_OFF_BMAX);

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_kc = Q / BK;
    const int n_it = 2 * n_kc;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + FT_BUILD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], FT_EPI_WARPS);
            mbar_init(&span_full[b], FT_LOADER_THREADS / 32);
            mbar_init(&span_empty[b], FT_BUILD_WARPS);
            mbar_init(&ri_full[b], FT_BUILD_WARPS);
            mbar_init(&ri_empty[b], FT_EPI_WARPS);
            mbar_init(&span_landed[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)T_TMEM_COLS);
    for (int j = threadIdx.x; j < Q; j += FT_THREADS) wtab[j] = __ldg(p.wtab + j);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    StripIter strip(p.total_tiles, p.n_frames, p.rows, (int)blockIdx.x, (int)gridDim.x);
    int64_t sig, t0;
    int ncols;

    if (warp == 0) {
        // ===================== TMA producer: basis k-chunks =====================
        if (elect_one()) {
            int g = 0;
            while (strip.next(sig, t0, ncols))
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % T_STAGES;
                    const uint32_t ph = (g / T_STAGES) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    mbar_wait_relaxed(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)Q * BK * 2);
                    uint8_t* sb = stages + (size_t)s * T_STAGE_BYTES;
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
            int g = 0, n = 0;
            for (; strip.next(sig, t0, ncols); ++n) {
                const int buf = n & 1;
                const uint32_t idesc = umma_idesc_f16(TILE_M, (ncols + 15) & ~15);
                mbar_wait_relaxed(&tmem_empty[buf], (uint32_t)(((n >> 1) & 1) ^ 1));
                tcgen05_fence_after();
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % T_STAGES;
                    const uint32_t ph = (g / T_STAGES) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    mbar_wait_relaxed(&full_bar[s], ph, 32);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(stages + (size_t)s * T_STAGE_BYTES);
                    const uint32_t b0 = a0 + T_STAGE_BASIS;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t d = tmem_base + (uint32_t)(buf * 4 * T_NF + (pair * 2 + j) * T_NF);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            const uint32_t off = ks * UMMA_K * 2;
                            const uint64_t bh = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                            const uint64_t bl = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                            const uint64_t dh = umma_desc_sw64(b0 + (j * 2) * T_DATA_TILE + off);
                            const uint64_t dl = umma_desc_sw64(b0 + (j * 2 + 1) * T_DATA_TILE + off);
                            umma_f16(d, bh, dh, idesc, (kc | ks) != 0);
                            umma_f16(d, bh, dl, idesc, 1);
                            umma_f16(d, bl, dh, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp < FT_EPI_WARP0) {
        // ===================== span loaders ====================================
        const int lt = threadIdx.x - FT_LOADER_WARP0 * 32;         // 0..63
        const int shift = p.shift;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int b = n & 1;
            float* span = span2 + b * FT_SPAN;
            uint32_t* bmax = bmax2 + b * FT_BMAX;
            float4* rowinfo = rowinfo2 + b * T_NF;
            const float* xs = p.x + sig * p.x_stride;
            const int64_t span0 = t0 * H - p.origin - shift;       // first sample of the span (may be < 0)
            const int span_len = (ncols - 1) * H + N + shift;
            const int span_pad = (span_len + 31) & ~31;
            const uint32_t kph = (uint32_t)((n >> 1) & 1);
            mbar_wait_relaxed(&span_empty[b], kph ^ 1);
            mbar_wait_relaxed(&ri_empty[b], kph ^ 1);
            // valid samples are span indices [lo, hi); [lo4, hi4) leaves by one bulk copy
            const int lo = (int)min((int64_t)span_pad, max((int64_t)0, -span0));
            const int hi = (int)max((int64_t)lo, min((int64_t)span_pad, p.samples - span0));
            const bool vec = ((((uintptr_t)xs) & 15) == 0) && ((span0 & 3) == 0);
            int lo4 = lo, hi4 = lo;
            if (vec) {
                lo4 = (lo + 3) & ~3;
                hi4 = hi & ~3;
                if (hi4 < lo4) hi4 = lo4;
            }
            if (lt == 0) {
                const uint32_t bytes = (uint32_t)(hi4 - lo4) * 4u;
                mbar_arrive_expect_tx(&span_landed[b], bytes);
                if (bytes) bulk_g2s_t(span + lo4, xs + span0 + lo4, bytes, &span_landed[b]);
            }
            // everything outside the bulk range: zero padding, ragged edges, unaligned rows
            for (int i = lt; i < lo4; i += FT_LOADER_THREADS)
                span[i] = i >= lo ? __ldg(xs + span0 + i) : 0.f;
            for (int i = hi4 + lt; i < span_pad; i += FT_LOADER_THREADS)
                span[i] = i < hi ? __ldg(xs + span0 + i) : 0.f;
            mbar_wait_relaxed(&span_landed[b], kph, 32);
            named_bar_sync(2, FT_LOADER_THREADS);
            // ---- per-32-sample maxima (and the gradient's input multiplier) ----------
            for (int k = lt; k < span_pad / 32; k += FT_LOADER_THREADS) {
                float4* blk = reinterpret_cast<float4*>(span + 32 * k);
                float m = 0.f;
                // iSTFT gradient: gy / (overlap-added w^2); sample i sits at position i + origin =
                // hop block eu, offset eoff (one division per 32 samples, then incremental)
                int64_t eu = 0;
                int eoff = 0;
                if (p.grad_env) {
                    const int64_t kk = (N + H - 1) / H + 1;
                    const int64_t pb = span0 + 32 * k + p.origin + kk * H;     // >= 0
                    eu = pb / H - kk;
                    eoff = (int)(pb % H);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float4 f = blk[e];
                    if (p.in_mul) {
                        f = mul4(f, load4_clamped(p.in_mul, span0 + 32 * k + 4 * e, p.samples));
                        blk[e] = f;
                    } else if (p.grad_env) {
                        float* fe = reinterpret_cast<float*>(&f);
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int64_t i = span0 + 32 * k + 4 * e + cc;
                            if (i >= 0 && i < p.samples)
                                fe[cc] *= ola_inv_envelope(p.env_per, p.wsq, N, H, p.n_frames, eu, eoff);
                            if (++eoff >= H) {
                                eoff = 0;
                                ++eu;
                            }
                        }
                        blk[e] = f;
                    }
                    m = fmaxf(m, fmaxf(fmaxf(finite_abs(f.x), finite_abs(f.y)),
                                       fmaxf(finite_abs(f.z), finite_abs(f.w))));
                }
                bmax[k] = __float_as_uint(m);
            }
            named_bar_sync(2, FT_LOADER_THREADS);
            // ---- per-frame scale and rank-1 terms -------------------------------------
            if (lt < T_NF) {
                float4 ri = make_float4(1.f, 0.f, 0.f, 0.f);
                if (lt < ncols) {
                    const int b0 = (lt * H + shift) >> 5, b1 = (lt * H + shift + N - 1) >> 5;
                    uint32_t mx = 0u;
                    for (int k = b0; k <= b1; ++k) mx = max(mx, bmax[k]);
                    ri.x = row_scale(4.f * p.wmax * __uint_as_float(mx));
                    const float xq = span[shift + lt * H + Q] * p.wq,
                                x3q = span[shift + lt * H + 3 * Q] * p.w3q;
                    ri.z = xq + x3q;                   // ee[Q]
                    ri.w = xq - x3q;                   // oo[Q]
                }
                rowinfo[lt] = ri;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&span_full[b]);
        }
    } else if (warp < FT_BUILD_WARP0) {
        // ===================== epilogue ========================================
        const int e = warp - FT_EPI_WARP0;         // 0..7
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int chalf = e >> 2;                  // which 32 frames of the tile
        const int m = q * 32 + lane;               // bin pair: bins 2m, 2m+1
        const bool valid = m < Q;
        const float sgn = (m & 1) ? -1.f : 1.f;
        const int pitch = 2 * p.n_bins;
        const float ps = COMPRESS ? 1.f : p.post_scale;   // without compression scale_factor folds in
        const float dcs = m == 0 ? p.dc_scale : 1.f;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int buf = n & 1;
            const float4* rowinfo = rowinfo2 + buf * T_NF;
            const uint32_t kph = (uint32_t)((n >> 1) & 1);
            mbar_wait_relaxed(&ri_full[buf], kph);
            mbar_wait_relaxed(&tmem_full[buf], kph);
            tcgen05_fence_after();
            float* obase = p.out + ((sig * p.n_frames + t0) * (int64_t)pitch + 4 * m);
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * T_NF);
#pragma unroll 1
            for (int cb = 0; cb < 32; cb += 16) {
                const int c0 = chalf * 32 + cb;
                if (c0 >= ncols) break;
                uint32_t r0[16], r1[16], r2[16], r3[16];
                tmem_ld16_nowait(tq + (uint32_t)(c0), r0);              // Re X[2m]
                tmem_ld16_nowait(tq + (uint32_t)(T_NF + c0), r1);       // Re X[2m+1]
                tmem_ld16_nowait(tq + (uint32_t)(2 * T_NF + c0), r2);   // Im X[2m]
                tmem_ld16_nowait(tq + (uint32_t)(3 * T_NF + c0), r3);   // Im X[2m+1]
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = c0 + j;
                    if (c < ncols && valid) {
                        const float4 ri = rowinfo[c];              // scale, nyquist sum, ee[Q], oo[Q]
                        const float g0 = ps * p.basis_scale_inv * pow2_inv(ri.x);
                        float re_e = fmaf(__uint_as_float(r0[j]), g0, sgn * ps * ri.z) * dcs;
                        float re_o = __uint_as_float(r1[j]) * g0;
                        float im_e = __uint_as_float(r2[j]) * g0;
                        float im_o = fmaf(__uint_as_float(r3[j]), g0, -sgn * ps * ri.w);
                        if (COMPRESS) {
                            compress(re_e, im_e, p.post_expo);
                            compress(re_o, im_o, p.post_expo);
                            re_e *= p.post_scale; im_e *= p.post_scale;
                            re_o *= p.post_scale; im_o *= p.post_scale;
                        }
                        float* o = obase + (int64_t)c * pitch;
                        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                            *reinterpret_cast<float4*>(o) = make_float4(re_e, im_e, re_o, im_o);
                        } else {
                            *reinterpret_cast<float2*>(o) = make_float2(re_e, im_e);
                            *reinterpret_cast<float2*>(o + 2) = make_float2(re_o, im_o);
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            // Nyquist bin (purely real): one frame per epilogue thread
            const int c = e * 32 + lane;
            if (!p.odd && c < ncols) {
                const float4 ri = rowinfo[c];
                float v = (ri.y + ri.z) * p.edge_scale;            // Q is even: (-1)^Q = +1
                if (COMPRESS) v = compress_real(v, p.post_expo);
                *reinterpret_cast<float2*>(p.out + (sig * p.n_frames + t0 + c) * (int64_t)pitch + 2 * Hf) =
                    make_float2(v * p.post_scale, 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ri_empty[buf]);
        }
    } else {
        // ===================== builders ========================================
        const int bw = warp - FT_BUILD_WARP0;      // 0..7
        const int half = lane >> 4;                // which of the warp's two rows per pass
        const int pr = lane & 15;                  // n pair inside the 32-wide k-chunk
        const uint32_t chunk = (uint32_t)(pr >> 2);
        const int shift = p.shift;
        constexpr int RI = T_NF / FT_BUILD_WARPS / 2;     // row pairs per warp (4)
        int g = 0;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int b = n & 1;
            const float* span = span2 + b * FT_SPAN;
            float4* rowinfo = rowinfo2 + b * T_NF;
            mbar_wait(&span_full[b], (uint32_t)((n >> 1) & 1));
            float rscale[RI], nyq[RI];
            const float* frow[RI];
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                const int row = bw * (2 * RI) + 2 * i + half;
                rscale[i] = row < ncols ? rowinfo[row].x : 0.f;
                nyq[i] = 0.f;
                // rows >= ncols are built too (scale 0, accumulator columns never read)
                frow[i] = span + shift + (row < ncols ? row : 0) * H;
            }
            for (int kc = 0; kc < n_kc; ++kc) {
                const int n0 = kc * BK + 2 * pr;
                const float4 w0 = wtab[n0], w1 = wtab[n0 + 1];
                float sp0[RI], sp1[RI], rp0[RI], rp1[RI], sm0[RI], sm1[RI], rm0[RI], rm1[RI];
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    const float* fr = frow[i];
                    const float a0 = fr[n0] * w0.x, a1 = fr[n0 + 1] * w1.x;
                    const float b0 = fr[Hf - n0] * w0.y, b1 = fr[Hf - n0 - 1] * w1.y;
                    const float c0 = fr[Hf + n0] * w0.z, c1 = fr[Hf + n0 + 1] * w1.z;
                    const float d0 = n0 ? fr[N - n0] * w0.w : 0.f, d1 = fr[N - n0 - 1] * w1.w;
                    sp0[i] = a0 + d0; sp1[i] = a1 + d1; rp0[i] = b0 + c0; rp1[i] = b1 + c1;
                    sm0[i] = a0 - d0; sm1[i] = a1 - d1; rm0[i] = b0 - c0; rm1[i] = b1 - c1;
                }
#pragma unroll
                for (int pair = 0; pair < 2; ++pair, ++g) {
                    const int s = g % T_STAGES;
                    const uint32_t ph = (g / T_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = stages + (size_t)s * T_STAGE_BYTES + T_STAGE_BASIS;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const int row = bw * (2 * RI) + 2 * i + half;
                        float u0, u1, v0, v1;
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
                        split_store(dst, dst + T_DATA_TILE, u0 * sc, u1 * sc);
                        split_store(dst + 2 * T_DATA_TILE, dst + 3 * T_DATA_TILE, v0 * sc, v1 * sc);
                    }
                    if (pair == 1 && kc == n_kc - 1) {
                        // the Nyquist sums must be visible before the tile's last stage is released
#pragma unroll
                        for (int i = 0; i < RI; ++i) {
                            float v = nyq[i];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            const int row = bw * (2 * RI) + 2 * i + half;
                            if (pr == 0 && row < ncols) rowinfo[row].y = v;
                        }
                    }
                    fence_proxy_async();                           // generic -> async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ri_full[b]);
                mbar_arrive(&span_empty[b]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)T_TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// inverse: STFT.backward (brever/modules/stft.py:101-138), hop = Q or 2Q, n_fft = 4Q
//
// Columns are hop blocks v = 0 .. n_blocks-1 of a signal; column v carries frame v (zero data
// when v >= n_frames).  An epilogue thread owns sample offset n (its TMEM lane, row n of Ce, Co,
// Se, So) and produces, per frame, the four values
//     a = f[n]  b = f[N/2 - n]  c = f[N/2 + n]  d = f[N - n]      (thread 0: b = f[Q], d = f[3Q])
// which land in hop blocks v .. v + R - 1 at offset n (a, c) or at the mirrored offset (b, d).
// Walking along the frames it keeps the values of the previous R - 1 frames in registers, so a
// hop block is finished with two shared-memory writes (direct + mirrored) and no exchange
// between threads; the carried values survive from one tile to the next of the same strip.
// A strip that starts inside a signal recomputes the R - 1 frames before it (`skip` columns
// whose hop blocks belong to the previous CTA).
//
//   warp 0        TMA producer: basis k-chunks (3-stage ring)      warps 2-3 idle
//   warp 1        TMEM owner + MMA issuer
//   warps 4-7     epilogue + copy-out (1 / envelope, centre trim)
//   warps 8-15    operand builders: spectrogram (L2-resident after the scouts) -> scaled fp16
//                 hi / lo planes; the next k-chunk's loads are in flight while one is converted
//   warps 16-23   scouts: per-frame maxima of the NEXT tile straight from HBM (18 loads in
//                 flight per lane), which also leaves that tile L2-resident for the builders
constexpr int IT_THREADS = 768;
constexpr int IT_EPI_WARP0 = 4, IT_EPI_WARPS = 4;
constexpr int IT_BUILD_WARP0 = 8, IT_BUILD_WARPS = 8, IT_BUILD_THREADS = IT_BUILD_WARPS * 32;
constexpr int IT_SCOUT_WARP0 = 16, IT_SCOUT_WARPS = 8, IT_SCOUT_THREADS = IT_SCOUT_WARPS * 32;
constexpr int IT_OUT_FLOATS = T_NF * 258;                   // HQ = 1: 2 x [64][129]; HQ = 2: [64][257]
constexpr int IT_OFF_OUT = T_SMEM_STAGES;
constexpr int IT_OFF_ROWINFO = IT_OFF_OUT + IT_OUT_FLOATS * 4;   // 2 slots x 64 float4
constexpr int IT_OFF_SCRATCH = IT_OFF_ROWINFO + 2 * T_NF * 16;   // [12][64] floats (builders 0..7, scouts 8..11)
constexpr int IT_SMEM_BYTES = 1024 + IT_OFF_SCRATCH + 12 * T_NF * 4;
static_assert(IT_SMEM_BYTES <= 227 * 1024, "transposed inverse kernel shared memory");

struct InvStrip {
    int64_t g, g1, per_signal;
    int nf, halo;
    bool first;
    __device__ InvStrip(int64_t total, int64_t per_signal_, int nf_, int halo_, int cta, int ctas)
        : g(total * cta / ctas), g1(total * (cta + 1) / ctas), per_signal(per_signal_), nf(nf_),
          halo(halo_), first(true) {}
    // tile = columns [c0, c0 + ncols) of signal `sig`; the first `skip` columns only warm up
    // the carried state (their hop blocks belong to the previous strip); `fresh`: no carried
    // state from the previous tile
    __device__ bool next(int64_t& sig, int64_t& c0, int& ncols, int& skip, bool& fresh) {
        if (g >= g1) return false;
        sig = g / per_signal;
        const int64_t v = g - sig * per_signal;
        fresh = first || v == 0;
        skip = first ? (int)min((int64_t)halo, v) : 0;
        c0 = v - skip;
        const int64_t m = min((int64_t)nf, min(per_signal - c0, g1 - g + skip));
        ncols = (int)m;
        g += m - skip;
        first = false;
        return true;
    }
};

template <int HQ, bool FRAMES_FAST, bool DECOMP>
__global__ void __launch_bounds__(IT_THREADS, 1)
istft_t_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldInvParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[T_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[T_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];     // MMA -> epilogue
    __shared__ __align__(8) uint64_t tmem_empty[2];    // epilogue -> MMA
    __shared__ __align__(8) uint64_t scale_full[2];    // scouts -> builders (frame scales ready)
    __shared__ __align__(8) uint64_t scale_empty[2];   // epilogue -> scouts (row info slot reusable)
    __shared__ __align__(8) uint64_t ri_full[2];       // builders -> epilogue (rank-1 sums written)
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* outbuf = reinterpret_cast<float*>(stages + IT_OFF_OUT);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + IT_OFF_ROWINFO);
    float* scratch = reinterpret_cast<float*>(stages + IT_OFF_SCRATCH);

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_kc = Q / BK;
    const int n_it = 2 * n_kc;
    constexpr int R = 4 / HQ;                      // frames overlapping one hop block

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + IT_BUILD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], IT_EPI_WARPS);
            mbar_init(&scale_full[b], IT_SCOUT_WARPS);
            mbar_init(&scale_empty[b], IT_EPI_WARPS);
            mbar_init(&ri_full[b], IT_BUILD_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)T_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    InvStrip strip(p.total_tiles, p.n_blocks, T_NF, R - 1, (int)blockIdx.x, (int)gridDim.x);
    int64_t sig, c0;
    int ncols, skip;
    bool fresh;

    if (warp < IT_EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 0) {
            // ===================== TMA producer: basis k-chunks =====================
            if (elect_one()) {
                int g = 0;
                while (strip.next(sig, c0, ncols, skip, fresh))
                    for (int it = 0; it < n_it; ++it, ++g) {
                        const int s = g % T_STAGES;
                        const uint32_t ph = (g / T_STAGES) & 1;
                        const int kc = it >> 1, pair = it & 1;
                        mbar_wait_relaxed(&empty_bar[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)Q * BK * 2);
                        uint8_t* sb = stages + (size_t)s * T_STAGE_BYTES;
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
                int g = 0, n = 0;
                for (; strip.next(sig, c0, ncols, skip, fresh); ++n) {
                    const int buf = n & 1;
                    const uint32_t idesc = umma_idesc_f16(TILE_M, (ncols + 15) & ~15);
                    mbar_wait_relaxed(&tmem_empty[buf], (uint32_t)(((n >> 1) & 1) ^ 1));
                    tcgen05_fence_after();
                    for (int it = 0; it < n_it; ++it, ++g) {
                        const int s = g % T_STAGES;
                        const uint32_t ph = (g / T_STAGES) & 1;
                        const int kc = it >> 1, pair = it & 1;
                        mbar_wait_relaxed(&full_bar[s], ph, 32);
                        tcgen05_fence_after();
                        const uint32_t a0 = smem_u32(stages + (size_t)s * T_STAGE_BYTES);
                        const uint32_t b0 = a0 + T_STAGE_BASIS;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t d = tmem_base + (uint32_t)(buf * 4 * T_NF + (pair * 2 + j) * T_NF);
#pragma unroll
                            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                                const uint32_t off = ks * UMMA_K * 2;
                                const uint64_t bh = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                                const uint64_t bl = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                                const uint64_t dh = umma_desc_sw64(b0 + (j * 2) * T_DATA_TILE + off);
                                const uint64_t dl = umma_desc_sw64(b0 + (j * 2 + 1) * T_DATA_TILE + off);
                                umma_f16(d, bh, dh, idesc, (kc | ks) != 0);
                                umma_f16(d, bh, dl, idesc, 1);
                                umma_f16(d, bl, dh, idesc, 1);
                            }
                        }
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&tmem_full[buf]);
                }
            }
        }
    } else if (warp < IT_BUILD_WARP0) {
        // ===================== epilogue + copy-out ==============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int nn = q * 32 + lane;              // sample offset n (row of Ce, Co, Se, So)
        const bool valid = nn < Q;
        const float4 wn = valid ? __ldg(p.wtab + nn) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float sgn = (nn & 1) ? -1.f : 1.f;
        const bool t0 = nn == 0;                   // carries f[Q], f[3Q] in its b, d slots
        const int pitch = H + 1;                   // row skew (copy-out reads lanes along offsets)
        float* outD = outbuf;
        float* outM = HQ == 1 ? outbuf + T_NF * (Q + 1) : outbuf;
        const int moff = HQ == 1 ? (t0 ? 0 : Q - nn) : (t0 ? Q : 2 * Q - nn);   // mirrored offset
        float env_reg[8];                          // 1 / envelope of interior hop blocks
#pragma unroll
        for (int j = 0; j < 8; ++j)
            env_reg[j] = p.no_env ? 1.f : (lane + 32 * j < H ? __ldg(p.env_per + lane + 32 * j) : 0.f);
        float c1 = 0.f, c2 = 0.f, b1 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;   // previous frames' values
        for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
            const int buf = n & 1;
            const float4* rowinfo = rowinfo2 + buf * T_NF;
            const uint32_t kph = (uint32_t)((n >> 1) & 1);
            if (fresh) c1 = c2 = b1 = d1 = d2 = d3 = 0.f;
            mbar_wait_relaxed(&ri_full[buf], kph);
            mbar_wait_relaxed(&tmem_full[buf], kph);
            tcgen05_fence_after();
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * T_NF);
#pragma unroll 1
            for (int cb = 0; cb < ncols; cb += 8) {
                uint32_t A[4][8];
#pragma unroll
                for (int a = 0; a < 4; ++a) tmem_ld8_nowait(tq + (uint32_t)(a * T_NF + cb), A[a]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = cb + j;
                    // columns >= ncols hold stale accumulators: select, do not multiply, them away
                    const bool live = col < ncols && valid;
                    const float4 ri = rowinfo[min(col, T_NF - 1)];   // scale, 2 pacc, 2 racc, nyquist term
                    const float g0 = p.basis_scale_inv * pow2_inv(ri.x);
                    const float ce = live ? __uint_as_float(A[0][j]) * g0 : 0.f;
                    const float co = live ? __uint_as_float(A[1][j]) * g0 : 0.f;
                    const float se = live ? __uint_as_float(A[2][j]) * g0 : 0.f;
                    const float so = live ? __uint_as_float(A[3][j]) * g0 : 0.f;
                    const float ny = live ? sgn * ri.w : 0.f;
                    const float cp = ce + co, cm = ce - co, sp = se + so, sm = se - so;
                    const float a = ((cp - sp) + ny) * wn.x;
                    float b = ((cm + sm) + ny) * wn.y;
                    const float c = ((cm - sm) + ny) * wn.z;
                    float d = ((cp + sp) + ny) * wn.w;
                    // thread 0: f[Q], f[3Q] (Q is even: the Nyquist term enters with +1); selects,
                    // not a branch, inside the unrolled columns
                    const float fq = live ? (ri.y - ri.z + ri.w) * p.wq : 0.f;
                    const float f3q = live ? (ri.y + ri.z + ri.w) * p.w3q : 0.f;
                    b = t0 ? fq : b;
                    d = t0 ? f3q : d;
                    float dv, mv;
                    if (HQ == 1) {
                        dv = a + c2; mv = b1 + d3;
                        c2 = c1; c1 = c; d3 = d2; d2 = d1; d1 = d; b1 = b;
                    } else {
                        dv = a + c1; mv = b + d1;
                        c1 = c; d1 = d;
                    }
                    if (col < ncols && valid) {
                        outD[col * pitch + nn] = dv;
                        outM[col * pitch + moff] = mv;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tmem_empty[buf]);
                mbar_arrive(&scale_empty[buf]);
            }
            named_bar_sync(3, IT_EPI_WARPS * 32);
            // ---- copy-out: finished hop blocks, * 1 / envelope, centre trim ---------------
            {
                float* ys = p.y + sig * p.out_len;
                for (int r = skip + q; r < ncols; r += IT_EPI_WARPS) {
                    const int64_t u = c0 + r;
                    const float* srcD = outD + r * pitch;
                    const float* srcM = outM + r * pitch;
                    const int64_t i0 = u * H - p.origin;
                    const bool use_reg = p.no_env || (u >= R - 1 && u <= p.n_frames - 1);
                    const int lo = i0 < 0 ? (int)min((int64_t)H, -i0) : 0;
                    const int hi = (int)max((int64_t)0, min((int64_t)H, p.out_len - i0));
                    float* yrow = ys + i0;             // dereferenced inside [lo, hi) only
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int off = lane + 32 * j;
                        if (off >= lo && off < hi) {
                            float v = HQ == 1 ? srcD[off] + srcM[off] : srcD[off];
                            float e = env_reg[j];
                            if (!use_reg) {
                                // edge hop blocks see fewer than R frames: overlap-added w^2 on the fly
                                float acc = 0.f;
                                const int64_t t_lo = max((int64_t)0, u - (R - 1));
                                const int64_t t_hi = min(u, p.n_frames - 1);
                                for (int64_t t = t_lo; t <= t_hi; ++t)
                                    acc += __ldg(p.wsq + (int)(u - t) * H + off);
                                e = 1.f / acc;
                            }
                            yrow[off] = v * e;
                        }
                    }
                }
            }
            named_bar_sync(3, IT_EPI_WARPS * 32);      // rows are rewritten by the next tile
        }
    } else if (warp < IT_SCOUT_WARP0) {
        // ===================== builders ========================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
        const int bw = warp - IT_BUILD_WARP0;      // 0..7
        const int bt = bw * 32 + lane;             // 0..255
        int g = 0;
        if (FRAMES_FAST) {
            // lanes along frames: thread = (frame row, 8 k = 16 bins of every 32-wide k-chunk)
            const int row = bt & (T_NF - 1), kq = bt >> 6;
            const uint32_t sw = (uint32_t)((row >> 1) & 3);
            for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
                const int slot = n & 1;
                float4* rowinfo = rowinfo2 + slot * T_NF;
                const bool live = row < ncols && c0 + row < p.n_frames;
                const float2* xr = p.spec + sig * p.ss + (c0 + (live ? row : 0)) * p.sf;
                mbar_wait(&scale_full[slot], (uint32_t)((n >> 1) & 1));
                const float sc = live ? rowinfo[row].x : 0.f;
                float pacc = 0.f, racc = 0.f;
                for (int kc = 0; kc < n_kc; ++kc) {
                    float2 c[16];                      // bins 64 kc + 16 kq + e
                    const int bin0 = 64 * kc + 16 * kq;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        c[e] = live ? __ldg(xr + (int64_t)(bin0 + e) * p.sb) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int e = 0; e < 16; ++e) c[e] = prep_bin<DECOMP>(c[e], p.pre_scale, p.pre_expo);
                    if (bin0 == 0) {
                        c[0].y = 0.f;                  // Im X[0] is ignored by the c2r inverse
                        c[0].x *= p.dc_gain;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        pacc += c[2 * j].x - c[2 * j + 2].x;          // (-1)^m Re X[2m]
                        racc += c[2 * j + 1].y - c[2 * j + 3].y;      // (-1)^m Im X[2m+1]
                    }
                    if (bin0 == 0) pacc -= 0.5f * c[0].x;             // c_0 = 1, the others 2
#pragma unroll
                    for (int pair = 0; pair < 2; ++pair, ++g) {
                        const int s = g % T_STAGES;
                        const uint32_t ph = (g / T_STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* sa = stages + (size_t)s * T_STAGE_BYTES + T_STAGE_BASIS + row * (BK * 2);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {          // sub-GEMM: even / odd bins
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 ca = c[4 * e + j], cb = c[4 * e + 2 + j];
                                const float v0 = (pair ? ca.y : ca.x) * sc;
                                const float v1 = (pair ? cb.y : cb.x) * sc;
                                const __half2 h = __floats2half2_rn(v0, v1);
                                const float2 hf = __half22float2(h);
                                const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                                hi[e] = *reinterpret_cast<const uint32_t*>(&h);
                                lo[e] = *reinterpret_cast<const uint32_t*>(&l);
                            }
                            const uint32_t dst = (((uint32_t)kq) ^ sw) << 4;
                            *reinterpret_cast<uint4*>(sa + (j * 2) * T_DATA_TILE + dst) =
                                make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(sa + (j * 2 + 1) * T_DATA_TILE + dst) =
                                make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                        if (pair == 1 && kc == n_kc - 1) {
                            scratch[kq * T_NF + row] = pacc;
                            scratch[4 * T_NF + kq * T_NF + row] = racc;
                            named_bar_sync(1, IT_BUILD_THREADS);
                            if (kq == 0) {
                                rowinfo[row].y = 2.f * ((scratch[row] + scratch[T_NF + row]) +
                                                        (scratch[2 * T_NF + row] + scratch[3 * T_NF + row]));
                                rowinfo[row].z = 2.f * ((scratch[4 * T_NF + row] + scratch[5 * T_NF + row]) +
                                                        (scratch[6 * T_NF + row] + scratch[7 * T_NF + row]));
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[s]);
                    }
                }
                named_bar_sync(1, IT_BUILD_THREADS);       // scratch is rewritten by the next tile
                if (lane == 0) mbar_arrive(&ri_full[slot]);
            }
        } else {
            // lanes along bins: a warp owns 8 frames; lane = (one of two frames, 4 bins)
            const int half = lane >> 4;
            const int pr = lane & 15;              // m pair inside the 32-wide k-chunk
            const uint32_t chunk = (uint32_t)(pr >> 2);
            constexpr int RI = T_NF / IT_BUILD_WARPS / 2;     // 4 row pairs per warp
            // bins 2 m0 .. 2 m0 + 3 (m0 = 32 kc + 2 pr) of the warp's frames of k-chunk kc
            auto load_chunk = [&](float2 (&dst)[RI][4], int64_t sig_, int64_t c0_, int ncols_, int kc) {
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    const int row = bw * (2 * RI) + 2 * i + half;
                    const bool live = row < ncols_ && c0_ + row < p.n_frames;
                    const float2* xr = p.spec + sig_ * p.ss + (c0_ + row) * p.sf + 2 * (kc * BK + 2 * pr);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        dst[i][e] = live ? __ldg(xr + e) : make_float2(0.f, 0.f);      // sb == 1
                }
            };
            float2 cur[RI][4], nxt[RI][4];
            InvStrip ahead = strip;                // runs one tile ahead (prefetch of its first k-chunk)
            int64_t sig_n, c0_n;
            int ncols_n, skip_n;
            bool fresh_n;
            bool more = ahead.next(sig_n, c0_n, ncols_n, skip_n, fresh_n);
            if (more) load_chunk(cur, sig_n, c0_n, ncols_n, 0);
            for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
                const int slot = n & 1;
                float4* rowinfo = rowinfo2 + slot * T_NF;
                more = ahead.next(sig_n, c0_n, ncols_n, skip_n, fresh_n);
                mbar_wait(&scale_full[slot], (uint32_t)((n >> 1) & 1));
                float pacc[RI], racc[RI], rscale[RI];
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    pacc[i] = racc[i] = 0.f;
                    const int row = bw * (2 * RI) + 2 * i + half;
                    rscale[i] = (row < ncols && c0 + row < p.n_frames) ? rowinfo[row].x : 0.f;
                }
                for (int kc = 0; kc < n_kc; ++kc) {
                    // the next chunk's loads (of this tile, or the first of the next tile) fly
                    // while this one is converted
                    if (kc + 1 < n_kc) load_chunk(nxt, sig, c0, ncols, kc + 1);
                    else if (more) load_chunk(nxt, sig_n, c0_n, ncols_n, 0);
                    const int m0 = kc * BK + 2 * pr;   // bins 2 m0 .. 2 m0 + 3
                    const bool dc = m0 == 0;
                    const float dc_mul = dc ? p.dc_gain : 1.f, dc_half = dc ? 0.5f : 0.f;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            cur[i][e] = prep_bin<DECOMP>(cur[i][e], p.pre_scale, p.pre_expo);
                        cur[i][0].y = dc ? 0.f : cur[i][0].y;      // Im X[0] is ignored by the c2r inverse
                        cur[i][0].x *= dc_mul;
                        pacc[i] += cur[i][0].x - cur[i][2].x;
                        racc[i] += cur[i][1].y - cur[i][3].y;
                        pacc[i] -= dc_half * cur[i][0].x;
                    }
#pragma unroll
                    for (int pair = 0; pair < 2; ++pair, ++g) {
                        const int s = g % T_STAGES;
                        const uint32_t ph = (g / T_STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* sa = stages + (size_t)s * T_STAGE_BYTES + T_STAGE_BASIS;
#pragma unroll
                        for (int i = 0; i < RI; ++i) {
                            const int row = bw * (2 * RI) + 2 * i + half;
                            const float sc = rscale[i];
                            uint8_t* dst = sa + row * (BK * 2) +
                                           ((chunk ^ (uint32_t)((row >> 1) & 3)) << 4) + (pr & 3) * 4;
                            if (pair == 0) {
                                split_store(dst, dst + T_DATA_TILE, cur[i][0].x * sc, cur[i][2].x * sc);
                                split_store(dst + 2 * T_DATA_TILE, dst + 3 * T_DATA_TILE, cur[i][1].x * sc,
                                            cur[i][3].x * sc);
                            } else {
                                split_store(dst, dst + T_DATA_TILE, cur[i][0].y * sc, cur[i][2].y * sc);
                                split_store(dst + 2 * T_DATA_TILE, dst + 3 * T_DATA_TILE, cur[i][1].y * sc,
                                            cur[i][3].y * sc);
                            }
                        }
                        if (pair == 1 && kc == n_kc - 1) {
                            // rank-1 sums must be visible before the tile's last stage is released
#pragma unroll
                            for (int i = 0; i < RI; ++i) {
                                float a = pacc[i], b = racc[i];
#pragma unroll
                                for (int o = 8; o; o >>= 1) {
                                    a += __shfl_xor_sync(0xffffffffu, a, o);
                                    b += __shfl_xor_sync(0xffffffffu, b, o);
                                }
                                const int row = bw * (2 * RI) + 2 * i + half;
                                if (pr == 0) {
                                    rowinfo[row].y = 2.f * a;
                                    rowinfo[row].z = 2.f * b;
                                }
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[s]);
                    }
#pragma unroll
                    for (int i = 0; i < RI; ++i)
#pragma unroll
                        for (int e = 0; e < 4; ++e) cur[i][e] = nxt[i][e];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ri_full[slot]);
            }
        }
    } else {
        // ===================== scouts: frame maxima of the next tile ==================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int sw = warp - IT_SCOUT_WARP0;      // 0..7
        for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
            const int slot = n & 1;
            float4* ri = rowinfo2 + slot * T_NF;
            const float2* xs = p.spec + sig * p.ss;
            mbar_wait_relaxed(&scale_empty[slot], (uint32_t)(((n >> 1) & 1) ^ 1));
            if (FRAMES_FAST) {
                // thread = (frame, quarter of the bins); lanes along frames
                const int st = sw * 32 + lane;
                const int row = st & (T_NF - 1), kq = st >> 6;
                const bool live = row < ncols && c0 + row < p.n_frames;
                const float2* col = xs + (c0 + (live ? row : 0)) * p.sf;
                float m = 0.f;
                const int nb = Q / 2;              // bins per quarter (2Q bins below the Nyquist bin)
                for (int b0 = kq * nb; b0 < (kq + 1) * nb; b0 += 16) {
                    float2 v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        v[e] = live ? __ldg(col + (int64_t)(b0 + e) * p.sb) : make_float2(0.f, 0.f);
                    if (b0 == 0) v[0].y = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        m = fmaxf(m, abs2_finite(prep_bin<DECOMP>(v[e], p.pre_scale, p.pre_expo)));
                }
                // scratch rows 8.. are the scouts' (the builders use rows 0..7)
                float* sc = scratch + 8 * T_NF;
                sc[kq * T_NF + row] = m;
                named_bar_sync(2, IT_SCOUT_THREADS);
                if (kq == 0) {
                    float ny = 0.f;
                    if (live)
                        ny = prep_bin<DECOMP>(__ldg(col + (int64_t)Hf * p.sb), p.pre_scale, p.pre_expo).x *
                             p.edge_gain;
                    const float mm = fmaxf(fmaxf(sc[row], sc[T_NF + row]),
                                           fmaxf(sc[2 * T_NF + row], sc[3 * T_NF + row]));
                    ri[row] = make_float4(live ? row_scale(mm) : 1.f, 0.f, 0.f, ny);
                }
                named_bar_sync(2, IT_SCOUT_THREADS);
            } else {
                // warp = frame (bins contiguous), two frames in flight per warp
                const int nj = Q / 16;             // 32-bin groups below the Nyquist bin
#pragma unroll 1
                for (int r0 = 2 * sw; r0 < T_NF; r0 += 2 * IT_SCOUT_WARPS) {
                    float2 v[2][8];
                    float2 vn[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = r0 + r;
                        const bool live = row < ncols && c0 + row < p.n_frames;
                        const float2* xr = xs + (c0 + row) * p.sf;      // sb == 1
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[r][j] = (live && j < nj) ? __ldg(xr + j * 32 + lane) : make_float2(0.f, 0.f);
                        vn[r] = (live && lane == 0) ? __ldg(xr + Hf) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = r0 + r;
                        if (lane == 0) v[r][0].y = 0.f;   // Im X[0] never reaches the output
                        float m = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            m = fmaxf(m, abs2_finite(prep_bin<DECOMP>(v[r][j], p.pre_scale, p.pre_expo)));
#pragma unroll
                        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                        if (lane == 0) {
                            const bool live = row < ncols && c0 + row < p.n_frames;
                            const float ny = prep_bin<DECOMP>(vn[r], p.pre_scale, p.pre_expo).x * p.edge_gain;
                            ri[row] = make_float4(live ? row_scale(m) : 1.f, 0.f, 0.f, live ? ny : 0.f);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&scale_full[slot]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)T_TMEM_COLS);
    }
}
