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
//     TMEM columns (the inverse kernel's default is 32 frames and FOUR resident tiles): the
//     tensor core works on tile i+1 while the epilogue drains tile i, and the operand builders
//     never wait for an epilogue (the one-tile-per-TMEM kernels serialise build -> MMA ->
//     epilogue on every tile);
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

// Dev-only role timing (-DBRV_PHASE_TIMING): lane 0 of each role's first warp of CTA 0 stamps
// %globaltimer when it starts waiting for a tile (0), starts working on it (1) and is done (2);
// tools/t_phase.py reads them back through brv_debug_t_times.
#ifdef BRV_PHASE_TIMING
__device__ unsigned long long g_brv_t_ts[5 * 8 * 4];
#define T_STAMP(role, tile, which)                                                      \
    do {                                                                                \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile) < 8) {                 \
            unsigned long long t_;                                                      \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
            g_brv_t_ts[((role) * 8 + (tile)) * 4 + (which)] = t_;                       \
        }                                                                               \
    } while (0)
// cycles a role's lane 0 spent inside one kind of wait, summed over the kernel (slot 0..31)
__device__ unsigned long long g_brv_t_wait[32];
#define T_WAIT_DECL long long t_wacc_[4] = {0, 0, 0, 0}
#define T_WAITED(k, stmt)                                                               \
    do {                                                                                \
        const long long t0_ = clock64();                                                \
        stmt;                                                                           \
        t_wacc_[k] += clock64() - t0_;                                                  \
    } while (0)
#define T_WAIT_FLUSH(role)                                                              \
    do {                                                                                \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)                                 \
            for (int k_ = 0; k_ < 4; ++k_) g_brv_t_wait[(role) * 4 + k_] = (unsigned long long)t_wacc_[k_]; \
    } while (0)
#else
#define T_STAMP(role, tile, which) do {} while (0)
#define T_WAIT_DECL do {} while (0)
#define T_WAITED(k, stmt) stmt
#define T_WAIT_FLUSH(role) do {} while (0)
#endif

// the strip of CTA `cta` out of `ctas` over `total` columns (< 2^31), walked in tiles
struct StripIter {
    uint32_t g, g1, per_signal;
    int nf;
    __device__ StripIter(int64_t total, int64_t per_signal_, int nf_, int cta, int ctas)
        : g((uint32_t)(total * cta / ctas)), g1((uint32_t)(total * (cta + 1) / ctas)),
          per_signal((uint32_t)per_signal_), nf(nf_) {}
    __device__ __forceinline__ bool next(int64_t& sig, int64_t& c0, int& ncols) {
        if (g >= g1) return false;
        const uint32_t s = g / per_signal, c = g - s * per_signal;
        sig = s;
        c0 = c;
        const uint32_t m = min((uint32_t)nf, min(per_signal - c, g1 - g));
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
//   warp 2        span loader: bulk-copies the NEXT tile's sample span into the other span buffer
//                 (zero padding and ragged edges by its lanes; the iSTFT gradient's 1 / envelope
//                 multiplier is applied here too); warp 3 idle
//   warps 4-11    epilogue: thread = bin pair (TMEM lane), 32 frames per warp, 16-byte stores
//   warps 12-19   operand builders: a warp owns 8 frames: per-32-sample maxima of their samples ->
//                 per-frame power-of-two scales and rank-1 terms (no CTA-wide barrier), then
//                 window, fold, scale, split from the staged span
constexpr int FT_THREADS = 896;
constexpr int FT_LOADER_WARP0 = 2;
constexpr int FT_EPI_WARP0 = 4, FT_EPI_WARPS = 8;
constexpr int FT_BUILD_WARP0 = 12, FT_BUILD_WARPS = 16;
constexpr int FT_BMAX_W = 72;                               // per-builder-warp maxima (<= (3 H + N) / 32 + 2)
// Shared memory of the forward kernel for NF frames per tile (64: two TMEM buffers, 3 stages; 32: four
// TMEM buffers, 4 stages -- a shorter pipeline fill for launches of a few tiles per SM)
template <int NF>
struct FtLayout {
    static constexpr int NBUF = T_TMEM_COLS / (4 * NF);
    static constexpr int DATA_TILE = NF * BK * 2;
    static constexpr int STAGE_BYTES = T_STAGE_BASIS + 4 * DATA_TILE;
    static constexpr int STAGES = NF == 64 ? 3 : 4;
    static constexpr int SPAN = (NF - 1) * 128 + 512 + 64;      // floats per span buffer
    static constexpr int OFF_SPAN = STAGES * STAGE_BYTES;
    static constexpr int OFF_WTAB = OFF_SPAN + 2 * SPAN * 4;
    static constexpr int OFF_ROWINFO = OFF_WTAB + SMEM_WTAB;     // NBUF slots x NF float4
    static constexpr int OFF_BMAX = OFF_ROWINFO + NBUF * NF * 16;
    static constexpr int SMEM_BYTES = 1024 + OFF_BMAX + FT_BUILD_WARPS * FT_BMAX_W * 4 + 32;
    static_assert(SMEM_BYTES <= 227 * 1024, "transposed forward kernel shared memory");
};

// frames per tile for a given geometry: the tile's span must fit one span buffer
static inline int ft_tile_frames(int n_fft, int hop, int shift, int nf_max) {
    const int span = (nf_max - 1) * 128 + 512 + 64;
    int nf = (span - 32 - n_fft - shift) / hop + 1;
    return nf > nf_max ? nf_max : nf;
}

// iSTFT gradient: span[i] *= 1 / (overlap-added w^2) of sample span0 + i (position + origin)
__device__ __noinline__ void ft_apply_envelope(float* span, int i_begin, int i_end, int i_step, int64_t span0,
                                               const float* in_mul, const float* env_per, const float* wsq,
                                               int N, int H, int64_t n_frames, int64_t samples, int origin) {
    for (int i = i_begin; i < i_end; i += i_step) {
        const int64_t j = span0 + i;
        if (j < 0 || j >= samples) continue;
        float m;
        if (in_mul) {
            m = __ldg(in_mul + j);
        } else {
            const int64_t pos = j + origin;
            const int64_t u = pos / H;
            m = ola_inv_envelope(env_per, wsq, N, H, n_frames, u, (int)(pos - u * H));
        }
        span[i] *= m;
    }
}

template <bool COMPRESS, int NF>
__global__ void __launch_bounds__(FT_THREADS, 1)
stft_t_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldFwdParams p) {
    using L = FtLayout<NF>;
    constexpr int NBUF = L::NBUF, NSTAGE = L::STAGES, STAGE_BYTES = L::STAGE_BYTES, DATA_TILE = L::DATA_TILE;
    constexpr int SPAN = L::SPAN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4];
    __shared__ __align__(8) uint64_t empty_bar[4];
    __shared__ __align__(8) uint64_t tmem_full[4];     // MMA -> epilogue
    __shared__ __align__(8) uint64_t tmem_empty[4];    // epilogue -> MMA
    __shared__ __align__(8) uint64_t span_empty[2];    // builders -> loader
    __shared__ __align__(8) uint64_t ri_full[4];       // builders -> epilogue (row info complete)
    __shared__ __align__(8) uint64_t ri_empty[4];      // epilogue -> builders (row info slot)
    __shared__ __align__(8) uint64_t span_landed[2];   // loader -> builders: span staged (bulk copy + edges)
    __shared__ __align__(8) uint64_t span_copied[2];   // bulk copy landed (loader's own, gradient use)
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* span2 = reinterpret_cast<float*>(stages + L::OFF_SPAN);
    float4* wtab = reinterpret_cast<float4*>(stages + L::OFF_WTAB);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + L::OFF_ROWINFO);
    uint32_t* bmax_all = reinterpret_cast<uint32_t*>(stages + L::OFF_BMAX);

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_kc = Q / BK;
    const int n_it = 2 * n_kc;
    // Q = 64: the basis rows are staged twice, TMEM lanes 64..127 hold a copy of the accumulators and
    // the epilogue warps of those lanes (idle otherwise) take half of the columns
    const bool dup = Q == 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1 + FT_BUILD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], FT_EPI_WARPS);
            mbar_init(&ri_full[b], FT_BUILD_WARPS);
            mbar_init(&ri_empty[b], FT_EPI_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&span_empty[b], FT_BUILD_WARPS);
            mbar_init(&span_landed[b], 2);              // expect_tx arrive + edge fills done
            mbar_init(&span_copied[b], 1);
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

    // register budget (72 x 896 at launch): control warpgroup 40, epilogue 80, builders 72
    if (warp < FT_EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ===================== TMA producer: basis k-chunks =====================
        if (elect_one()) {
            T_WAIT_DECL;
            int g = 0;
            while (strip.next(sig, t0, ncols))
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % NSTAGE;
                    const uint32_t ph = (g / NSTAGE) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    T_WAITED(0, mbar_wait_relaxed(&empty_bar[s], ph ^ 1));
#ifdef BRV_T_NO_TMA                                     // dev experiment: timing without the basis stream
                    mbar_arrive(&full_bar[s]);
                    (void)kc; (void)pair;
#else
                    mbar_arrive_expect_tx(&full_bar[s], (dup ? 8u : 4u) * (uint32_t)Q * BK * 2);
                    uint8_t* sb = stages + (size_t)s * STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {
                            tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE), &basis_map,
                                        &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
                            if (dup)                   // the same Q rows again below them (rows Q .. 2Q - 1)
                                tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE + Q * (BK * 2)), &basis_map,
                                            &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
                        }
#endif
                    T_WAIT_FLUSH(4);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            T_WAIT_DECL;
            int g = 0, n = 0;
            for (; strip.next(sig, t0, ncols); ++n) {
                const int buf = n % NBUF;
                const uint32_t idesc = umma_idesc_f16(TILE_M, (ncols + 15) & ~15);
                T_STAMP(2, n, 0);
                T_WAITED(0, mbar_wait_relaxed(&tmem_empty[buf], (uint32_t)(((n / NBUF) & 1) ^ 1)));
                tcgen05_fence_after();
                T_STAMP(2, n, 1);
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % NSTAGE;
                    const uint32_t ph = (g / NSTAGE) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    T_WAITED(1, mbar_wait_relaxed(&full_bar[s], ph, 32));
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(stages + (size_t)s * STAGE_BYTES);
                    const uint32_t b0 = a0 + T_STAGE_BASIS;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t d = tmem_base + (uint32_t)(buf * 4 * NF + (pair * 2 + j) * NF);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            const uint32_t off = ks * UMMA_K * 2;
                            const uint64_t bh = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                            const uint64_t bl = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                            const uint64_t dh = umma_desc_sw64(b0 + (j * 2) * DATA_TILE + off);
                            const uint64_t dl = umma_desc_sw64(b0 + (j * 2 + 1) * DATA_TILE + off);
#ifndef BRV_T_NO_MMA                                    // dev experiments: MMA count
                            umma_f16(d, bh, dh, idesc, (kc | ks) != 0);
#ifndef BRV_T_ONE_PRODUCT
                            umma_f16(d, bh, dl, idesc, 1);
                            umma_f16(d, bl, dh, idesc, 1);
#endif
#else
                            (void)d; (void)bh; (void)bl; (void)dh; (void)dl; (void)idesc;
#endif
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
                T_STAMP(2, n, 2);
                T_WAIT_FLUSH(2);
            }
        }
    } else if (warp == FT_LOADER_WARP0) {
        // ===================== span loader =====================================
        T_WAIT_DECL;
        const int shift = p.shift;
        const bool mul = p.in_mul != nullptr || p.grad_env;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int b = n & 1;
            float* span = span2 + b * SPAN;
            const float* xs = p.x + sig * p.x_stride;
            const int64_t span0 = t0 * H - p.origin - shift;       // first sample of the span (may be < 0)
            const int span_len = (ncols - 1) * H + N + shift;
            const int span_pad = (span_len + 31) & ~31;
            const uint32_t kph = (uint32_t)((n >> 1) & 1);
            T_STAMP(0, n, 0);
            T_WAITED(0, mbar_wait_relaxed(&span_empty[b], kph ^ 1));
            T_STAMP(0, n, 1);
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
            uint64_t* landing = mul ? &span_copied[b] : &span_landed[b];
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(hi4 - lo4) * 4u;
                mbar_arrive_expect_tx(landing, bytes);
                if (bytes) bulk_g2s_t(span + lo4, xs + span0 + lo4, bytes, landing);
            }
            // everything outside the bulk range: zero padding, ragged edges, unaligned rows
            for (int i = lane; i < lo4; i += 32) span[i] = i >= lo ? __ldg(xs + span0 + i) : 0.f;
            for (int i = hi4 + lane; i < span_pad; i += 32) span[i] = i < hi ? __ldg(xs + span0 + i) : 0.f;
            if (mul) {
                // gradient of the inverse transform: gy / envelope, once the copy has landed
                mbar_wait_relaxed(&span_copied[b], kph, 32);
                __syncwarp();
                ft_apply_envelope(span, lane, span_pad, 32, span0, p.in_mul, p.env_per, p.wsq, N, H, p.n_frames,
                                  p.samples, p.origin);
                __syncwarp();
                if (lane == 0) mbar_arrive(&span_landed[b]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&span_landed[b]);
            T_STAMP(0, n, 2);
            T_WAIT_FLUSH(0);
        }
    }                                                  // (warp 3 idles)
    } else if (warp < FT_BUILD_WARP0) {
        // ===================== epilogue ========================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
        T_WAIT_DECL;
        const int e = warp - FT_EPI_WARP0;         // 0..7
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int chalf = e >> 2;                  // which half of the tile's frames
        const int m = dup ? (q & 1) * 32 + lane : q * 32 + lane;       // bin pair: bins 2m, 2m+1
        const int c_lo = chalf * (NF / 2) + (dup ? (q >> 1) * (NF / 4) : 0);       // this warp's columns
        const int c_n = dup ? NF / 4 : NF / 2;
        const bool valid = m < Q;
        const float sgn = (m & 1) ? -1.f : 1.f;
        const int pitch = 2 * p.n_bins;
        const bool alt = (pitch & 3) != 0;         // row alignment alternates between 16 and 8 bytes
        const float ps = COMPRESS ? 1.f : p.post_scale;   // without compression scale_factor folds in
        const float dcs = m == 0 ? p.dc_scale : 1.f;
        const float nys = (p.odd && m == Q - 1) ? p.edge_scale : 1.f;   // n_fft = 4Q - 2: Nyquist = last odd bin
        const float gb = ps * p.basis_scale_inv;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int buf = n % NBUF;
            const float4* rowinfo = rowinfo2 + buf * NF;
            const uint32_t kph = (uint32_t)((n / NBUF) & 1);
            if (warp == FT_EPI_WARP0) T_STAMP(3, n, 0);
            T_WAITED(0, mbar_wait_relaxed(&ri_full[buf], kph));
            T_WAITED(1, mbar_wait_relaxed(&tmem_full[buf], kph));
            tcgen05_fence_after();
            if (warp == FT_EPI_WARP0) T_STAMP(3, n, 1);
            float* obase = p.out + ((sig * p.n_frames + t0) * (int64_t)pitch + 4 * m);
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * NF);
#pragma unroll 1
            for (int cb = 0; cb < c_n; cb += 8) {
                const int c0 = c_lo + cb;
#ifdef BRV_T_NO_EPI                                     // dev experiment: timing without the epilogue stores
                if (p.n_fft > 0) break;
#endif
                if (c0 >= ncols) break;
                uint32_t r0[8], r1[8], r2[8], r3[8];
                tmem_ld8_nowait(tq + (uint32_t)(c0), r0);              // Re X[2m]
                tmem_ld8_nowait(tq + (uint32_t)(NF + c0), r1);       // Re X[2m+1]
                tmem_ld8_nowait(tq + (uint32_t)(2 * NF + c0), r2);   // Im X[2m]
                tmem_ld8_nowait(tq + (uint32_t)(3 * NF + c0), r3);   // Im X[2m+1]
                float* o0 = obase + (int64_t)c0 * pitch;
                const bool a0 = (reinterpret_cast<uintptr_t>(o0) & 15) == 0;    // c0 is even
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c0 + j < ncols && valid) {
                        const float4 ri = rowinfo[c0 + j];         // scale, nyquist sum, ee[Q], oo[Q]
                        const float g0 = gb * pow2_inv(ri.x);
                        float re_e = fmaf(__uint_as_float(r0[j]), g0, sgn * ps * ri.z) * dcs;
                        float re_o = __uint_as_float(r1[j]) * g0 * nys;
                        float im_e = __uint_as_float(r2[j]) * g0;
                        float im_o = fmaf(__uint_as_float(r3[j]), g0, -sgn * ps * ri.w);
                        if (COMPRESS) {
                            compress(re_e, im_e, p.post_expo);
                            compress(re_o, im_o, p.post_expo);
                            re_e *= p.post_scale; im_e *= p.post_scale;
                            re_o *= p.post_scale; im_o *= p.post_scale;
                        }
                        float* o = o0 + j * pitch;
                        if ((j & 1) && alt ? !a0 : a0) {
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
            if (warp == FT_EPI_WARP0) T_STAMP(3, n, 2);
            if (warp == FT_EPI_WARP0) T_WAIT_FLUSH(3);
        }
    } else {
        // ===================== builders ========================================
        T_WAIT_DECL;
        const int bw = warp - FT_BUILD_WARP0;      // 0..15
        const int half = lane >> 4;                // which of the warp's two rows per pass
        const int pr = lane & 15;                  // n pair inside the 32-wide k-chunk
        const uint32_t chunk = (uint32_t)(pr >> 2);
        const int shift = p.shift;
        constexpr int RI = NF / FT_BUILD_WARPS / 2;     // row pairs per warp (2)
        constexpr int ROWS_W = 2 * RI;                    // rows per warp (4)
        uint32_t soff[RI];                                // swizzled byte offset of this lane's 4 bytes per row
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const uint32_t row = (uint32_t)(bw * ROWS_W + 2 * i + half);
            soff[i] = row * (BK * 2) + ((chunk ^ ((row >> 1) & 3u)) << 4) + (uint32_t)(pr & 3) * 4u;
        }
        int g = 0;
        for (int n = 0; strip.next(sig, t0, ncols); ++n) {
            const int b = n & 1, slot = n % NBUF;
            const float* span = span2 + b * SPAN;
            float4* rowinfo = rowinfo2 + slot * NF;
            if (bw == 0) T_STAMP(1, n, 0);
            const uint32_t kph = (uint32_t)((n >> 1) & 1);
            mbar_wait_relaxed(&span_landed[b], kph, 20);
            mbar_wait_relaxed(&ri_empty[slot], (uint32_t)(((n / NBUF) & 1) ^ 1), 20);
            if (bw == 0) T_STAMP(1, n, 1);
            if (bw * ROWS_W < ncols) {
                // ---- this warp's frames: per-32-sample maxima of the samples they cover (the
                //      neighbouring warps scan the overlap again into their own tables), then
                //      the per-frame scale and rank-1 terms ----------------------------------
                uint32_t* bmax = bmax_all + bw * FT_BMAX_W - (((bw * ROWS_W * H + shift) & ~31) >> 5);
                const int rows_w = min(ROWS_W, ncols - bw * ROWS_W);
                const int s_lo = (bw * ROWS_W * H + shift) & ~31;
                const int s_hi = ((bw * ROWS_W + rows_w - 1) * H + shift + N + 31) & ~31;   // whole blocks (<= span_pad)
#pragma unroll 4
                for (int i0 = s_lo; i0 < s_hi; i0 += 128) {
                    const int i = i0 + lane * 4;
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < s_hi) f = *reinterpret_cast<const float4*>(span + i);
                    // fmaxf drops NaNs; an inf is removed by the slow path
                    float m = fmaxf(fmaxf(fabsf(f.x), fabsf(f.y)), fmaxf(fabsf(f.z), fabsf(f.w)));
                    if (!(m <= 3.0e38f))
                        m = fmaxf(fmaxf(finite_abs(f.x), finite_abs(f.y)),
                                  fmaxf(finite_abs(f.z), finite_abs(f.w)));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    if ((lane & 7) == 0 && i < s_hi) bmax[i >> 5] = __float_as_uint(m);
                }
                __syncwarp();
                {
                    // eight lanes per frame walk its block maxima (a frame covers up to 17 blocks)
                    const int rr = lane >> 3, part = lane & 7;
                    const int row = bw * ROWS_W + rr;
                    const int b0 = (row * H + shift) >> 5, b1 = (row * H + shift + N - 1) >> 5;
                    uint32_t mx = 0u;
                    if (rr < rows_w)
                        for (int k = b0 + part; k <= b1; k += 8) mx = max(mx, bmax[k]);
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                    if (part == 0 && rr < rows_w) {
                        const float xq = span[shift + row * H + Q] * p.wq,
                                    x3q = span[shift + row * H + 3 * Q] * p.w3q;
                        // scale, (nyquist sum: below), ee[Q], oo[Q]
                        rowinfo[row] = make_float4(row_scale(4.f * p.wmax * __uint_as_float(mx)), 0.f,
                                                   xq + x3q, xq - x3q);
                    }
                }
                __syncwarp();
            }
            float rscale[RI], nyq[RI];
            const float* frow[RI];
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                const int row = bw * ROWS_W + 2 * i + half;
                rscale[i] = row < ncols ? rowinfo[row].x : 0.f;
                nyq[i] = 0.f;
                // rows >= ncols are built too (scale 0, accumulator columns never read)
                frow[i] = span + shift + (row < ncols ? row : 0) * H;
            }
            for (int kc = 0; kc < n_kc; ++kc) {
#ifdef BRV_PHASE_TIMING
                const long long tf0_ = clock64();
#endif
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
#ifdef BRV_PHASE_TIMING
                t_wacc_[0] += clock64() - tf0_;            // (timing build: kind 0 = window + fold part)
#endif
#pragma unroll
                for (int pair = 0; pair < 2; ++pair, ++g) {
                    const int s = g % NSTAGE;
                    const uint32_t ph = (g / NSTAGE) & 1;
                    T_WAITED(2, mbar_wait_relaxed(&empty_bar[s], ph ^ 1, 20));
#ifdef BRV_PHASE_TIMING
                    const long long ts0_ = clock64();
#endif
                    uint8_t* sa = stages + (size_t)s * STAGE_BYTES + T_STAGE_BASIS;
#ifdef BRV_T_NO_BUILD                                   // dev experiment: timing without the operand build
                    if (p.n_fft < 0)
#endif
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
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
                        uint8_t* dst = sa + soff[i];
                        split_store(dst, dst + DATA_TILE, u0 * sc, u1 * sc);
                        split_store(dst + 2 * DATA_TILE, dst + 3 * DATA_TILE, v0 * sc, v1 * sc);
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
                            const int row = bw * ROWS_W + 2 * i + half;
                            if (pr == 0 && row < ncols) rowinfo[row].y = v;
                        }
                    }
#ifdef BRV_PHASE_TIMING
                    const long long ts1_ = clock64();
                    t_wacc_[1] += ts1_ - ts0_;                 // (kind 1 = split + store loop)
#endif
                    fence_proxy_async();                           // generic -> async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
#ifdef BRV_PHASE_TIMING
                    t_wacc_[3] += clock64() - ts1_;            // (kind 3 = fence + arrive)
#endif
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ri_full[slot]);
                mbar_arrive(&span_empty[b]);
            }
            if (bw == 0) T_STAMP(1, n, 2);
            if (bw == 0) T_WAIT_FLUSH(1);
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
// which land in hop blocks v .. v + R - 1 at offset n (a, c: "direct") or at the mirrored offset
// (b, d).  Walking along the frames it keeps the values of the previous R - 1 frames in registers:
//     hop = Q :  direct(v) = a(v) + c(v-2)   mirrored(v) = b(v-1) + d(v-3)   at offset (Q - n) % Q
//     hop = 2Q:  direct(v) = a(v) + c(v-1)   mirrored(v) = b(v) + d(v-1)     at offset 2Q - n (Q for n = 0)
// For hop = 2Q the two land on different offsets and go straight to HBM; for hop = Q the mirrored
// sums of 8 frames cross the 128 threads through a small shared-memory exchange, are added to the
// direct sums in registers and leave as 128-byte rows per warp, multiplied by 1 / envelope (periodic
// in the interior, summed on the fly at the edges) -- no staged output tile, no copy-out phase.
// Two sets of four epilogue warps alternate tiles (set = TMEM buffer); the carried values of a
// tile's last R - 1 frames are computed FIRST and handed to the other set, so both sets run
// concurrently.  A strip that starts inside a signal recomputes the R - 1 frames before it (`skip`
// columns whose hop blocks belong to the previous CTA).
//
//   warp 0        TMA producer: basis k-chunks (3- or 4-stage ring)
//   warp 1        TMEM owner + MMA issuer
//   warps 4-7     epilogue set 0 (even tiles)      warps 8-11  epilogue set 1 (odd tiles)
//   warps 12-19   operand builders: spectrogram (L2-resident after the scouts) -> scaled fp16
//                 hi / lo planes; frame-major input: the k-chunks of all tiles form one stream
//                 over two register buffers, so the next chunk's loads (also across tiles) fly
//                 while one is converted; bin-major input: two groups of threads alternate chunks
//   warps 20-27   scouts: per-frame maxima of the tiles ahead straight from HBM (18 loads in
//                 flight per lane), which also leaves them L2-resident for the builders
// Template flavours: NF = 64 / 32 frames per tile (2 / 4 tiles resident in TMEM), ODD (n_fft =
// 4Q - 2: three exchange planes for the one-sample shifts of the segments), DUP (Q = 64, hop 2Q:
// accumulators duplicated on lanes 64..127 so that every epilogue warp has columns to work on).
constexpr int IT_THREADS = 896;
constexpr int IT_EPI_WARP0 = 4, IT_EPI_SET_WARPS = 4;
constexpr int IT_BUILD_WARP0 = 12, IT_BUILD_WARPS = 8, IT_BUILD_THREADS = IT_BUILD_WARPS * 32;
constexpr int IT_SCOUT_WARP0 = 20, IT_SCOUT_WARPS = 8, IT_SCOUT_THREADS = IT_SCOUT_WARPS * 32;
constexpr int IT_XP = MAX_Q + 1;                              // exchange row pitch (floats)
// Shared memory of the inverse kernel for NF frames per tile (64: two TMEM buffers, 3 stages;
// 32: four TMEM buffers, 4 stages -- more tiles in flight for launches of a few tiles per SM)
template <int NF, bool ODD = false>
struct ItLayout {
    static constexpr int NBUF = T_TMEM_COLS / (4 * NF);          // tiles resident in TMEM
    static constexpr int DATA_TILE = NF * BK * 2;
    static constexpr int STAGE_BYTES = T_STAGE_BASIS + 4 * DATA_TILE;
    static constexpr int STAGES = (NF == 64 || ODD) ? 3 : 4;
    static constexpr int XPLANES = ODD ? 3 : 1;                  // exchange planes (see the odd-fold epilogue)
    static constexpr int OFF_EXCH = STAGES * STAGE_BYTES;                   // [set 2][buffer 2][plane][8 frames][IT_XP]
    static constexpr int OFF_CARRY = OFF_EXCH + 2 * 2 * XPLANES * 8 * IT_XP * 4;   // [consumer set 2][slot 2][6][128]
    static constexpr int OFF_ROWINFO = OFF_CARRY + 2 * 2 * 6 * 128 * 4;     // NBUF slots x NF float4
    static constexpr int OFF_SCRATCH = OFF_ROWINFO + NBUF * NF * 16;        // [12][64] floats (builders 0..7, scouts 8..11)
    static constexpr int SMEM_BYTES = 1024 + OFF_SCRATCH + 12 * 64 * 4;
    static_assert(SMEM_BYTES <= 227 * 1024, "transposed inverse kernel shared memory");
};
constexpr int IT_MAX_BUF = 4;

// Scouts of a compressed spectrogram: the frame scale comes from a bound on the decompressed
// magnitudes, max_k (|pre_scale X_k|)^(1 + expo), from one |X|^2 per bin and one power per frame,
// instead of decompressing every bin a second time (the bound is at most sqrt(2) above the largest
// component: half a bit of the split's 22)
__device__ __forceinline__ float norm2_finite(float2 c) {
    const float fx = finite_abs(c.x), fy = finite_abs(c.y);
    return fminf(fmaf(fx, fx, fy * fy), 3.0e38f);
}
__device__ __forceinline__ float decomp_bound(float m2, float pre_scale, float expo) {
    const float mag2 = fminf(pre_scale * pre_scale * m2, 3.0e38f);
    return mag2 > 0.f ? fminf(sqrtf(mag2) * pow_half_expo(mag2, expo), 3.0e38f) : 0.f;
}

struct InvStrip {
    uint32_t g, g1, per_signal;
    uint32_t s, v;                                     // signal and hop block of column g (one division, at construction)
    int nf, halo;
    bool first;
    __device__ InvStrip(int64_t total, int64_t per_signal_, int nf_, int halo_, int cta, int ctas)
        : g((uint32_t)(total * cta / ctas)), g1((uint32_t)(total * (cta + 1) / ctas)),
          per_signal((uint32_t)per_signal_), nf(nf_), halo(halo_), first(true) {
        s = g / per_signal;
        v = g - s * per_signal;
    }
    // tile = columns [c0, c0 + ncols) of signal `sig`; the first `skip` columns only warm up
    // the carried state (their hop blocks belong to the previous strip); `fresh`: no carried
    // state from the previous tile.  A tile that is followed by a non-fresh one is always full
    // (ncols == nf): only the end of a signal or of the strip cuts a tile short.
    __device__ __forceinline__ bool next(int64_t& sig, int64_t& c0, int& ncols, int& skip, bool& fresh) {
        if (g >= g1) return false;
        sig = s;
        fresh = first || v == 0;
        skip = first ? (int)min((uint32_t)halo, v) : 0;
        c0 = v - skip;
        const uint32_t m = min((uint32_t)nf, min(per_signal - (v - skip), g1 - g + skip));
        ncols = (int)m;
        g += m - skip;
        v += m - skip;
        if (v >= per_signal) {                         // the tile ended at the signal's end
            v = 0;
            ++s;
        }
        first = false;
        return true;
    }
};

__device__ __noinline__ float it_edge_envelope(const float* env_per, const float* wsq, int N, int H,
                                               int64_t n_frames, int64_t u, int off) {
    return ola_inv_envelope(env_per, wsq, N, H, n_frames, u, off);
}

// the four values of one frame for sample offset n from its row of Ce, Co, Se, So.
// ri = { basis_scale_inv / frame scale, f[Q], f[3Q], Nyquist term } (the builders finish the first
// three).  Nothing is masked here: stale accumulators of columns past the tile's end, and lanes
// n >= Q, produce values that are never stored nor carried into a stored sum.
template <bool ODD>
__device__ __forceinline__ void it_frame_values(uint32_t ace, uint32_t aco, uint32_t ase, uint32_t aso,
                                                const float4 ri, bool t0, float sgn, const float4 wn,
                                                float& a, float& b, float& c, float& d) {
    const float fce = __uint_as_float(ace), fco = __uint_as_float(aco);
    const float fse = __uint_as_float(ase), fso = __uint_as_float(aso);
    if (ODD) {
        // n_fft = 4Q - 2: no Nyquist term outside the contraction, no self-paired samples
        // (1 / scales travel in ri.y here: the bin-major builders write it without a barrier)
        const float cp = (fce + fco) * ri.y, cm = (fce - fco) * ri.y;
        const float sp = (fse + fso) * ri.y, sm = (fse - fso) * ri.y;
        a = (cp - sp) * wn.x;           // f[n]
        b = (cm + sm) * wn.y;           // f[N/2 - n]
        c = (cm - sm) * wn.z;           // f[N/2 + n]   (n >= 1)
        d = (cp + sp) * wn.w;           // f[N - n]     (n >= 1; the table holds 0 for n = 0)
        return;
    }
    const float ny = sgn * ri.w;
    const float cpn = fmaf(fce + fco, ri.x, ny), cmn = fmaf(fce - fco, ri.x, ny);
    const float spg = (fse + fso) * ri.x, smg = (fse - fso) * ri.x;
    a = (cpn - spg) * wn.x;
    b = (cmn + smg) * wn.y;
    c = (cmn - smg) * wn.z;
    d = (cpn + spg) * wn.w;
    b = t0 ? ri.y : b;                  // thread 0 carries f[Q], f[3Q] in its b, d slots
    d = t0 ? ri.z : d;
}

// columns [lo, hi) of a tile whose sample index ibase + col * H lies in [0, out_len)
__device__ __forceinline__ void it_store_range(int64_t ibase, int64_t out_len, int H, int skip, int ncols,
                                               bool valid, int& lo, int& hi) {
    lo = skip;
    hi = ncols;
    if (ibase < 0) lo = max(lo, (int)(((uint32_t)(-ibase) + (uint32_t)H - 1u) / (uint32_t)H));
    const int64_t rem = out_len - ibase;            // col * H < rem
    if (rem < (int64_t)ncols * H) hi = rem <= 0 ? 0 : (int)(((uint32_t)rem + (uint32_t)H - 1u) / (uint32_t)H);
    if (!valid) hi = 0;
}

#ifdef BRV_T_NO_FENCE      // dev ablation: wrong results, timing only
#define IT_BUILD_FENCE() do {} while (0)
#else
#define IT_BUILD_FENCE() fence_proxy_async()
#endif
// DUP (Q = 64, hop = 2Q): the basis rows are staged twice, so TMEM lanes 64..127 hold a copy of the
// accumulators and the epilogue warps that would idle (sample offsets >= Q) take the second half of
// every tile's columns (one column of warm-up for the carried values).
template <int HQ, bool FRAMES_FAST, bool DECOMP, int NF, bool ODD = false, bool DUP = false>
__global__ void __launch_bounds__(IT_THREADS, 1)
istft_t_kernel(const __grid_constant__ CUtensorMap basis_map, const FoldInvParams p) {
    using L = ItLayout<NF, ODD>;
    static_assert(!ODD || HQ == 1, "n_fft = 4Q - 2 runs with hop = Q only");
    static_assert(!DUP || (HQ == 2 && !ODD && NF == 32), "duplicated lanes: hop = 2Q, 32-frame tiles");
    constexpr int NBUF = L::NBUF, NSTAGE = L::STAGES, STAGE_BYTES = L::STAGE_BYTES, DATA_TILE = L::DATA_TILE;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4];
    __shared__ __align__(8) uint64_t empty_bar[4];
    __shared__ __align__(8) uint64_t tmem_full[IT_MAX_BUF];     // MMA -> epilogue set
    __shared__ __align__(8) uint64_t tmem_empty[IT_MAX_BUF];    // epilogue set -> MMA
    __shared__ __align__(8) uint64_t scale_full[IT_MAX_BUF];    // scouts -> builders (frame scales ready)
    __shared__ __align__(8) uint64_t scale_empty[IT_MAX_BUF];   // epilogue set -> scouts (row info slot reusable)
    __shared__ __align__(8) uint64_t ri_full[IT_MAX_BUF];       // builders -> epilogue (rank-1 sums written)
    __shared__ __align__(8) uint64_t carry_full[2];    // other set -> set: carried values written
    __shared__ __align__(8) uint64_t carry_empty[2];   // set -> other set: carried values read
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* exch_base = reinterpret_cast<float*>(stages + L::OFF_EXCH);
    float* carry_base = reinterpret_cast<float*>(stages + L::OFF_CARRY);
    float4* rowinfo2 = reinterpret_cast<float4*>(stages + L::OFF_ROWINFO);
    float* scratch = reinterpret_cast<float*>(stages + L::OFF_SCRATCH);

    const int N = p.n_fft, H = p.hop, Q = p.q, Hf = N / 2;
    const int n_kc = Q / BK;
    const int n_it = 2 * n_kc;
    constexpr int R = 4 / HQ;                      // frames overlapping one hop block

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            // arrivals per stage: the TMA producer + the builder warps that write it
            mbar_init(&full_bar[s], 1 + (FRAMES_FAST ? IT_BUILD_THREADS / 32 / (IT_BUILD_THREADS / (4 * NF)) : IT_BUILD_WARPS));
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], IT_EPI_SET_WARPS);
            mbar_init(&scale_full[b], IT_SCOUT_WARPS);
            mbar_init(&scale_empty[b], IT_EPI_SET_WARPS);
            mbar_init(&ri_full[b], IT_BUILD_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&carry_full[b], IT_EPI_SET_WARPS);
            mbar_init(&carry_empty[b], IT_EPI_SET_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, (uint32_t)T_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    InvStrip strip(p.total_tiles, p.n_blocks, NF, R - 1, (int)blockIdx.x, (int)gridDim.x);
    int64_t sig, c0;
    int ncols, skip;
    bool fresh;

    // register budget (72 x 896 at launch): control 24, scouts 48, epilogue 80, builders 112
    if (warp < IT_EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == 0) {
        // ===================== TMA producer: basis k-chunks =====================
        if (elect_one()) {
            T_WAIT_DECL;
            int g = 0;
            while (strip.next(sig, c0, ncols, skip, fresh))
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % NSTAGE;
                    const uint32_t ph = (g / NSTAGE) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    T_WAITED(0, mbar_wait_relaxed(&empty_bar[s], ph ^ 1));
                    mbar_arrive_expect_tx(&full_bar[s], (DUP ? 8u : 4u) * (uint32_t)Q * BK * 2);
                    uint8_t* sb = stages + (size_t)s * STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {
                            tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE), &basis_map,
                                        &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
                            if (DUP)                   // the same Q rows again below them (rows Q .. 2Q - 1)
                                tma_load_2d(smem_u32(sb + (j * 2 + pl) * SUB_TILE + Q * (BK * 2)), &basis_map,
                                            &full_bar[s], kc * BK, (pl * 4 + pair * 2 + j) * Q);
                        }
                    T_WAIT_FLUSH(4);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            T_WAIT_DECL;
            int g = 0, n = 0;
            for (; strip.next(sig, c0, ncols, skip, fresh); ++n) {
                const int buf = n % NBUF;
                const uint32_t idesc = umma_idesc_f16(TILE_M, (ncols + 15) & ~15);
                T_STAMP(2, n, 0);
                T_WAITED(0, mbar_wait_relaxed(&tmem_empty[buf], (uint32_t)(((n / NBUF) & 1) ^ 1)));
                tcgen05_fence_after();
                T_STAMP(2, n, 1);
                for (int it = 0; it < n_it; ++it, ++g) {
                    const int s = g % NSTAGE;
                    const uint32_t ph = (g / NSTAGE) & 1;
                    const int kc = it >> 1, pair = it & 1;
                    T_WAITED(1, mbar_wait_relaxed(&full_bar[s], ph, 32));
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(stages + (size_t)s * STAGE_BYTES);
                    const uint32_t b0 = a0 + T_STAGE_BASIS;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t d = tmem_base + (uint32_t)(buf * 4 * NF + (pair * 2 + j) * NF);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            const uint32_t off = ks * UMMA_K * 2;
                            const uint64_t bh = umma_desc_sw64(a0 + (j * 2) * SUB_TILE + off);
                            const uint64_t bl = umma_desc_sw64(a0 + (j * 2 + 1) * SUB_TILE + off);
                            const uint64_t dh = umma_desc_sw64(b0 + (j * 2) * DATA_TILE + off);
                            const uint64_t dl = umma_desc_sw64(b0 + (j * 2 + 1) * DATA_TILE + off);
                            umma_f16(d, bh, dh, idesc, (kc | ks) != 0);
                            umma_f16(d, bh, dl, idesc, 1);
                            umma_f16(d, bl, dh, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
                T_STAMP(2, n, 2);
                T_WAIT_FLUSH(2);
            }
        }
    }
    } else if (warp < IT_BUILD_WARP0) {
        // ===================== epilogue sets ===================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
        T_WAIT_DECL;
        const int set = (warp - IT_EPI_WARP0) >> 2;      // even / odd tiles; also the TMEM buffer
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int lane_row = q * 32 + lane;        // TMEM lane
        const int half = DUP ? (lane_row >= Q ? 1 : 0) : 0;      // DUP: which half of a tile's columns
        const int nn = DUP ? lane_row - half * Q : lane_row;     // sample offset n (row of Ce, Co, Se, So)
        const bool valid = nn < Q;
        const float4 wn = valid ? __ldg(p.wtab + nn) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float sgn = (nn & 1) ? -1.f : 1.f;
        const bool t0 = nn == 0;                   // carries f[Q], f[3Q] in its b, d slots
        const int moff = HQ == 1 ? (t0 ? 0 : Q - nn) : (t0 ? Q : 2 * Q - nn);   // mirrored offset
        const float env_n = p.no_env ? 1.f : (valid ? __ldg(p.env_per + nn) : 0.f);
        const float env_m = p.no_env ? 1.f : ((valid && HQ == 2) ? __ldg(p.env_per + moff) : 0.f);
        constexpr int XPL = L::XPLANES * 8 * IT_XP;          // floats per exchange buffer
        float* exch = exch_base + set * (2 * XPL);
        float* carry_in = carry_base + set * (2 * 6 * 128);
        float* carry_out = carry_base + (set ^ 1) * (2 * 6 * 128);
        uint32_t n_in = 0, n_out = 0;              // carried-state hand-offs received / sent
        int pp = 0;                                // exchange buffer
        for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
            if ((n & 1) != set) continue;
            // the next tile continues this signal (then this tile is full)
            const bool cont = strip.g < strip.g1 && c0 + ncols < (int64_t)strip.per_signal;
            const int buf = n % NBUF;
            const float4* rowinfo = rowinfo2 + buf * NF;
            const uint32_t kph = (uint32_t)((n >> 1) & 1);         // carried-state slot
            const uint32_t bph = (uint32_t)((n / NBUF) & 1);       // use parity of the TMEM buffer
            if (q == 0) T_STAMP(3, n, 0);
            T_WAITED(0, mbar_wait_relaxed(&ri_full[buf], bph));
            T_WAITED(1, mbar_wait_relaxed(&tmem_full[buf], bph));
            tcgen05_fence_after();
            if (q == 0) T_STAMP(3, n, 1);
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * NF);
            if (cont && DUP && half == 0) {
                // (the other half of the warps owns the tile's last columns; the waits keep every
                //  warp's arrivals one per barrier phase)
                T_WAITED(2, mbar_wait_relaxed(&carry_empty[set ^ 1], (n_out & 1) ^ 1));
                ++n_out;
                __syncwarp();
                if (lane == 0) mbar_arrive(&carry_full[set ^ 1]);
            } else if (cont) {
                // ---- the values the next tile's first hop blocks need from this tile's last
                //      R - 1 frames go out first, so that the other set can start ----
                uint32_t A[4][8];
#pragma unroll
                for (int a = 0; a < 4; ++a) tmem_ld8_nowait(tq + (uint32_t)(a * NF + NF - 8), A[a]);
                tmem_ld_wait();
                float va[3], vb[3], vc[3], vd[3];
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    it_frame_values<ODD>(A[0][5 + j], A[1][5 + j], A[2][5 + j], A[3][5 + j],
                                    rowinfo[NF - 3 + j], t0, sgn, wn, va[j], vb[j], vc[j], vd[j]);
                T_WAITED(2, mbar_wait_relaxed(&carry_empty[set ^ 1], (n_out & 1) ^ 1));
                float* co = carry_out + (int)kph * (6 * 128) + nn;
                co[0 * 128] = vc[2];               // c(v-1)
                co[1 * 128] = vc[1];               // c(v-2)
                co[2 * 128] = vb[2];               // b(v-1)
                co[3 * 128] = vd[2];               // d(v-1)
                co[4 * 128] = vd[1];               // d(v-2)
                co[5 * 128] = vd[0];               // d(v-3)
                ++n_out;
                __syncwarp();
                if (lane == 0) mbar_arrive(&carry_full[set ^ 1]);
            }
            float c1 = 0.f, c2 = 0.f, b1 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;   // previous frames' values
            // DUP: columns [c_begin, c_end) of the tile are this warp's
            const int c_begin = half * (NF / 2);
            const int c_end = DUP ? (half == 0 ? min(ncols, NF / 2) : ncols) : ncols;
            if (DUP && half == 1) {
                if (!fresh) {                      // the hand-off is for the first half: only keep the count
                    T_WAITED(2, mbar_wait_relaxed(&carry_full[set], n_in & 1));
                    ++n_in;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&carry_empty[set]);
                }
                if (c_begin < c_end) {             // warm-up: the values of the column before this half
                    uint32_t A[4][8];
#pragma unroll
                    for (int a = 0; a < 4; ++a) tmem_ld8_nowait(tq + (uint32_t)(a * NF + c_begin - 8), A[a]);
                    tmem_ld_wait();
                    float a_, b_;
                    it_frame_values<ODD>(A[0][7], A[1][7], A[2][7], A[3][7], rowinfo[c_begin - 1], t0, sgn, wn,
                                         a_, b_, c1, d1);
                }
            } else if (!fresh) {
                T_WAITED(2, mbar_wait_relaxed(&carry_full[set], n_in & 1));
                ++n_in;
                const float* ci = carry_in + (int)(((n - 1) >> 1) & 1) * (6 * 128) + nn;
                c1 = ci[0 * 128]; c2 = ci[1 * 128]; b1 = ci[2 * 128];
                d1 = ci[3 * 128]; d2 = ci[4 * 128]; d3 = ci[5 * 128];
                __syncwarp();
                if (lane == 0) mbar_arrive(&carry_empty[set]);
            }
            // per tile, not per column: the store ranges (sample index inside the output, column
            // owned by this strip) and the columns whose envelope is the periodic interior one
            const int64_t ibase = c0 * H + nn - p.origin;          // output index of (column 0, offset n)
            float* yd = p.y + sig * p.out_len + ibase;             // never dereferenced outside [lo, hi)
            float* ym = yd + (moff - nn);
            int lo, hi, lo_m = 0, hi_m = 0;
            it_store_range(ibase, p.out_len, H, skip, ncols, valid, lo, hi);
            if (HQ == 2) it_store_range(ibase + (moff - nn), p.out_len, H, skip, ncols, valid, lo_m, hi_m);
            if (DUP) {
                lo = max(lo, c_begin); hi = min(hi, c_end);
                lo_m = max(lo_m, c_begin); hi_m = min(hi_m, c_end);
                if (c_begin >= c_end) {            // nothing to read in this tile: hand it back at once
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&tmem_empty[buf]);
                        mbar_arrive(&scale_empty[buf]);
                    }
                }
            }
            const int e_lo = p.no_env ? 0 : (int)max((int64_t)0, (int64_t)(R - 1) - c0);
            const int e_hi = p.no_env ? ncols : (int)max((int64_t)0, min((int64_t)ncols, p.n_frames - c0));
#pragma unroll 1
            for (int cb = c_begin; cb < c_end; cb += 8) {
                uint32_t A[4][8];
#pragma unroll
                for (int a = 0; a < 4; ++a) tmem_ld8_nowait(tq + (uint32_t)(a * NF + cb), A[a]);
                tmem_ld_wait();
                float dv[8], mv[8];
                float* xb = exch + pp * XPL;           // (hop = Q only)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a, b, c, d;
                    it_frame_values<ODD>(A[0][j], A[1][j], A[2][j], A[3][j], rowinfo[min(cb + j, NF - 1)], t0, sgn,
                                    wn, a, b, c, d);
                    if (ODD) {
                        // n_fft = 4Q - 2: the other three segments land one sample apart from each
                        // other's mirror (b at Q-1-n, c at n-1, d at Q-2-n; d of thread Q-1 one block
                        // early at Q-1): three exchange planes, every slot written by exactly one thread
                        dv[j] = a;
                        if (valid) {
                            float* x0 = xb + j * IT_XP;
                            x0[Q - 1 - nn] = b1;
                            if (nn >= 1) x0[8 * IT_XP + nn - 1] = c2;
                            if (nn == Q - 1) x0[8 * IT_XP + Q - 1] = d2;
                            x0[16 * IT_XP + (nn <= Q - 2 ? Q - 2 - nn : Q - 1)] = nn <= Q - 2 ? d3 : 0.f;
                        }
                        c2 = c1; c1 = c; d3 = d2; d2 = d1; d1 = d; b1 = b;
                    } else if (HQ == 1) {
                        dv[j] = a + c2; mv[j] = b1 + d3;
                        c2 = c1; c1 = c; d3 = d2; d2 = d1; d1 = d; b1 = b;
                    } else {
                        dv[j] = a + c1; mv[j] = b + d1;
                        c1 = c; d1 = d;
                    }
                }
                if (cb + 8 >= c_end) {
                    // last read of this tile's accumulators and row info: hand both back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&tmem_empty[buf]);
                        mbar_arrive(&scale_empty[buf]);
                    }
                }
                const bool interior = cb >= e_lo && cb + 8 <= e_hi;      // the whole batch (warp-uniform)
                if (HQ == 1) {
                    pp ^= 1;
                    if (valid && !ODD) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) xb[j * IT_XP + moff] = mv[j];
                    }
                    named_bar_sync(3 + set, IT_EPI_SET_WARPS * 32);
                    if (ODD) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dv[j] += xb[(8 + j) * IT_XP + nn] + xb[(16 + j) * IT_XP + nn];
                    }
                    if (interior) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = cb + j;
                            if (col >= lo && col < hi) yd[col * H] = (dv[j] + xb[j * IT_XP + nn]) * env_n;
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 8; ++j) {
                            const int col = cb + j;
                            if (col >= lo && col < hi) {
                                float v = dv[0];
#pragma unroll
                                for (int jj = 1; jj < 8; ++jj) v = j == jj ? dv[jj] : v;
                                const bool in = col >= e_lo && col < e_hi;
                                yd[col * H] = (v + xb[j * IT_XP + nn]) *
                                              (in ? env_n : it_edge_envelope(p.env_per, p.wsq, N, H, p.n_frames,
                                                                             c0 + col, nn));
                            }
                        }
                    }
                } else {
                    if (interior) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = cb + j;
                            if (col >= lo && col < hi) yd[col * H] = dv[j] * env_n;
                            if (col >= lo_m && col < hi_m) ym[col * H] = mv[j] * env_m;
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 8; ++j) {
                            const int col = cb + j;
                            float vd_ = dv[0], vm_ = mv[0];
#pragma unroll
                            for (int jj = 1; jj < 8; ++jj) {
                                vd_ = j == jj ? dv[jj] : vd_;
                                vm_ = j == jj ? mv[jj] : vm_;
                            }
                            const bool in = col >= e_lo && col < e_hi;
                            if (col >= lo && col < hi)
                                yd[col * H] = vd_ * (in ? env_n : it_edge_envelope(p.env_per, p.wsq, N, H,
                                                                                   p.n_frames, c0 + col, nn));
                            if (col >= lo_m && col < hi_m)
                                ym[col * H] = vm_ * (in ? env_m : it_edge_envelope(p.env_per, p.wsq, N, H,
                                                                                   p.n_frames, c0 + col, moff));
                        }
                    }
                }
            }
            if (q == 0) T_STAMP(3, n, 2);
            if (q == 0) T_WAIT_FLUSH(3);
        }
    } else if (warp < IT_SCOUT_WARP0) {
        // ===================== builders ========================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        T_WAIT_DECL;
        const int bw = warp - IT_BUILD_WARP0;      // 0..7
        const int bt = bw * 32 + lane;             // 0..255
        int g = 0;
        if (FRAMES_FAST) {
            // lanes along frames: thread = (frame row, 8 k = 16 bins of a 32-wide k-chunk, group);
            // with 32-frame tiles the 256 threads form two groups that alternate k-chunks (a group's
            // loads are in flight while the other one converts)
            constexpr int NGRP = IT_BUILD_THREADS / (4 * NF);
            const int row = bt & (NF - 1), kq = (bt / NF) & 3, grp = bt / (4 * NF);
            const uint32_t sw = (uint32_t)((row >> 1) & 3);
            const uint32_t dst = (((uint32_t)kq) ^ sw) << 4;
            for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n, g += n_it) {
                const int slot = n % NBUF;
                float4* rowinfo = rowinfo2 + slot * NF;
                const bool live = row < ncols && c0 + row < p.n_frames;
                const float2* xr = p.spec + sig * p.ss + (c0 + (live ? row : 0)) * p.sf;
                if (bw == 0) T_STAMP(1, n, 0);
                T_WAITED(0, mbar_wait_relaxed(&scale_full[slot], (uint32_t)((n / NBUF) & 1), 20));
                if (bw == 0) T_STAMP(1, n, 1);
                const float sc = live ? rowinfo[row].x : 0.f;
                float pacc = 0.f, racc = 0.f;
                for (int kc = grp; kc < n_kc; kc += NGRP) {
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
                    if (ODD && bin0 + 16 == 2 * Q) {   // n_fft = 4Q - 2: the Nyquist bin is the last odd bin
                        c[15].y = 0.f;
                        c[15].x *= p.edge_gain;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        pacc += c[2 * j].x - c[2 * j + 2].x;          // (-1)^m Re X[2m]
                        racc += c[2 * j + 1].y - c[2 * j + 3].y;      // (-1)^m Im X[2m+1]
                    }
                    if (bin0 == 0) pacc -= 0.5f * c[0].x;             // c_0 = 1, the others 2
                    // both sub-GEMM pairs of the chunk (two stages) are handed over together
                    const int g0 = g + 2 * kc;
                    const int s0 = g0 % NSTAGE, s1 = (g0 + 1) % NSTAGE;
                    T_WAITED(2, mbar_wait_relaxed(&empty_bar[s0], (uint32_t)(((g0 / NSTAGE) & 1) ^ 1), 20));
                    T_WAITED(2, mbar_wait_relaxed(&empty_bar[s1], (uint32_t)((((g0 + 1) / NSTAGE) & 1) ^ 1), 20));
#pragma unroll
                    for (int pair = 0; pair < 2; ++pair) {
                        uint8_t* sa = stages + (size_t)(pair ? s1 : s0) * STAGE_BYTES + T_STAGE_BASIS +
                                      row * (BK * 2) + dst;
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
                            *reinterpret_cast<uint4*>(sa + (j * 2) * DATA_TILE) =
                                make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(sa + (j * 2 + 1) * DATA_TILE) =
                                make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    IT_BUILD_FENCE();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&full_bar[s0]);
                        mbar_arrive(&full_bar[s1]);
                    }
                }
                if (ODD) {
                    // n_fft = 4Q - 2 has no rank-1 terms: only 1 / scales for the epilogue, in a field
                    // nobody reads during the tile -- no CTA-wide barrier needed
                    if (kq == 0 && grp == 0) rowinfo[row].y = sc > 0.f ? p.basis_scale_inv * pow2_inv(sc) : 0.f;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ri_full[slot]);
                    if (bw == 0) T_STAMP(1, n, 2);
                    if (bw == 0) T_WAIT_FLUSH(1);
                    continue;
                }
                // rank-1 sums: 4 NGRP partials per frame through the scratch rows
                scratch[(grp * 4 + kq) * NF + row] = pacc;
                scratch[(4 * NGRP + grp * 4 + kq) * NF + row] = racc;
                T_WAITED(3, named_bar_sync(1, IT_BUILD_THREADS));
                if (kq == 0 && grp == 0) {
                    float pa = 0.f, ra = 0.f;
#pragma unroll
                    for (int k = 0; k < 4 * NGRP; ++k) {
                        pa += scratch[k * NF + row];
                        ra += scratch[(4 * NGRP + k) * NF + row];
                    }
                    pa *= 2.f;
                    ra *= 2.f;
                    const float ny = rowinfo[row].w;
                    rowinfo[row] = make_float4(sc > 0.f ? p.basis_scale_inv * pow2_inv(sc) : 0.f,
                                               (pa - ra + ny) * p.wq, (pa + ra + ny) * p.w3q, ny);
                }
                T_WAITED(3, named_bar_sync(1, IT_BUILD_THREADS));       // scratch is rewritten by the next tile
                if (lane == 0) mbar_arrive(&ri_full[slot]);
                if (bw == 0) T_STAMP(1, n, 2);
                if (bw == 0) T_WAIT_FLUSH(1);
            }
        } else {
            // lanes along bins: a warp owns 8 frames; lane = (one of two frames, 4 bins)
            const int half = lane >> 4;
            const int pr = lane & 15;              // m pair inside the 32-wide k-chunk
            const uint32_t chunk = (uint32_t)(pr >> 2);
            constexpr int RI = NF / IT_BUILD_WARPS / 2;     // 4 row pairs per warp
            // per-thread constants: first row, swizzled byte offset of its 4 bytes inside a data block
            const int row0 = bw * (2 * RI) + half;             // rows row0 + 2 i
            uint32_t soff[RI];
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                const uint32_t row = (uint32_t)(row0 + 2 * i);
                soff[i] = row * (BK * 2) + ((chunk ^ ((row >> 1) & 3u)) << 4) + (uint32_t)(pr & 3) * 4u;
            }
            const int64_t sf2 = 2 * p.sf;
            // the plain path folds pre_scale into the frame scales (a power of two times it: same bits)
            const float fold_scale = DECOMP ? 1.f : p.pre_scale;
            float pacc[RI], racc[RI], rscale[RI];
            float4* rowinfo = rowinfo2;
            // one k-chunk: scale, split, store both sub-GEMM pairs
            auto convert = [&](float2 (&cur)[RI][4], int kc) {
                if (DECOMP) {
#pragma unroll
                    for (int i = 0; i < RI; ++i)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            cur[i][e] = prep_bin<true>(cur[i][e], p.pre_scale, p.pre_expo);
                }
                if (kc == 0) {                     // the DC bin lives in lane pr = 0 of the first chunk
                    const bool dc = pr == 0;
                    const float dc_mul = dc ? p.dc_gain : 1.f, dc_half = dc ? 0.5f : 0.f;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        cur[i][0].y = dc ? 0.f : cur[i][0].y;      // Im X[0] is ignored by the c2r inverse
                        cur[i][0].x *= dc_mul;
                        pacc[i] -= dc_half * cur[i][0].x;
                    }
                }
                if (ODD && kc == n_kc - 1 && pr == 15) {
                    // n_fft = 4Q - 2: the Nyquist bin is the last odd bin (Im ignored, unit weight
                    // except in the STFT gradient)
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        cur[i][3].y = 0.f;
                        cur[i][3].x *= p.edge_gain;
                    }
                }
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    pacc[i] += cur[i][0].x - cur[i][2].x;
                    racc[i] += cur[i][1].y - cur[i][3].y;
                }
                // both sub-GEMM pairs of the chunk (two stages) are handed over together: one
                // proxy fence and one warp barrier per k-chunk instead of per stage
                const int s0 = g % NSTAGE, s1 = (g + 1) % NSTAGE;
                T_WAITED(2, mbar_wait_relaxed(&empty_bar[s0], (uint32_t)(((g / NSTAGE) & 1) ^ 1), 20));
                T_WAITED(2, mbar_wait_relaxed(&empty_bar[s1], (uint32_t)((((g + 1) / NSTAGE) & 1) ^ 1), 20));
                g += 2;
#pragma unroll
                for (int pair = 0; pair < 2; ++pair) {
                    uint8_t* sa = stages + (size_t)(pair ? s1 : s0) * STAGE_BYTES + T_STAGE_BASIS;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const float sc = rscale[i];
                        uint8_t* dst = sa + soff[i];
                        if (pair == 0) {
                            split_store(dst, dst + DATA_TILE, cur[i][0].x * sc, cur[i][2].x * sc);
                            split_store(dst + 2 * DATA_TILE, dst + 3 * DATA_TILE, cur[i][1].x * sc,
                                        cur[i][3].x * sc);
                        } else {
                            split_store(dst, dst + DATA_TILE, cur[i][0].y * sc, cur[i][2].y * sc);
                            split_store(dst + 2 * DATA_TILE, dst + 3 * DATA_TILE, cur[i][1].y * sc,
                                        cur[i][3].y * sc);
                        }
                    }
                }
                if (kc == n_kc - 1) {
                    // rank-1 sums must be visible before the tile's last stage is released
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        float a = pacc[i], b = racc[i];
#pragma unroll
                        for (int o = 8; o; o >>= 1) {
                            a += __shfl_xor_sync(0xffffffffu, a, o);
                            b += __shfl_xor_sync(0xffffffffu, b, o);
                        }
                        if (pr == 0) {
                            // what the epilogue needs: 1 / scales, f[Q] and f[3Q] (Q is even: the
                            // Nyquist term enters them with +1), the Nyquist term
                            const float4 ri = rowinfo[row0 + 2 * i];
                            a *= 2.f * fold_scale;
                            b *= 2.f * fold_scale;
                            const float g0 = rscale[i] != 0.f ? p.basis_scale_inv * pow2_inv(ri.x) : 0.f;
                            rowinfo[row0 + 2 * i] = ODD ? make_float4(g0, g0, 0.f, 0.f)
                                                        : make_float4(g0, (a - b + ri.w) * p.wq,
                                                                      (a + b + ri.w) * p.w3q, ri.w);
                        }
                    }
                }
                IT_BUILD_FENCE();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&full_bar[s0]);
                    mbar_arrive(&full_bar[s1]);
                }
            };
            // The k-chunks of all the strip's tiles form one stream over NB register buffers (NB > 1:
            // the loads of the next NB - 1 chunks, also across tiles, fly while one is converted).
            // Measured: NB = 1 is the fastest (cfg4 247 -> 218 us, cfg5 914 -> 824 us against NB = 2):
            // the other warps hide the L2 latency, and every extra buffer is another inlined copy of
            // the conversion in kernels whose top stall is instruction fetch.
#ifndef BRV_T_NB
#define BRV_T_NB 1
#endif
            constexpr int NB = NF == 32 ? BRV_T_NB : 2;
            float2 ring[NB][RI][4];
            InvStrip ahead = strip;                // load cursor
            int64_t sig_l, c0_l;
            int ncols_l, skip_l, kc_l = 0;
            bool fresh_l;
            bool more_l = ahead.next(sig_l, c0_l, ncols_l, skip_l, fresh_l);
            // per load tile: pointer to this lane's 4 bins of k-chunk 0 of row0, liveness bit per row
            const float2* base_l = p.spec;
            uint32_t live_l = 0;
            auto set_load_tile = [&]() {
                base_l = p.spec + sig_l * p.ss + (c0_l + row0) * p.sf + 4 * pr;
                live_l = 0;
#pragma unroll
                for (int i = 0; i < RI; ++i)
                    if (row0 + 2 * i < ncols_l && c0_l + row0 + 2 * i < p.n_frames) live_l |= 1u << i;
            };
            if (more_l) set_load_tile();
            // bins 2 m0 .. 2 m0 + 3 (m0 = 32 kc + 2 pr) of the warp's frames, next k-chunk of the stream
            auto load_next = [&](float2 (&dst)[RI][4]) {
                if (!more_l) return;
                const float2* xr = base_l + kc_l * (2 * BK);
#pragma unroll
                for (int i = 0; i < RI; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e)       // sb == 1
                        dst[i][e] = ((live_l >> i) & 1u) ? __ldg(xr + i * sf2 + e) : make_float2(0.f, 0.f);
                if (++kc_l == n_kc) {
                    kc_l = 0;
                    more_l = ahead.next(sig_l, c0_l, ncols_l, skip_l, fresh_l);
                    if (more_l) set_load_tile();
                }
            };
#pragma unroll
            for (int u = 0; u < NB - 1; ++u) load_next(ring[u]);
            bool more = strip.next(sig, c0, ncols, skip, fresh);
            int n = 0, kc = 0, slot = 0;
            while (more) {
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    if (!more) break;
                    load_next(ring[(u + NB - 1) % NB]);
                    if (kc == 0) {
                        slot = n % NBUF;
                        rowinfo = rowinfo2 + slot * NF;
                        if (bw == 0) T_STAMP(1, n, 0);
                        T_WAITED(0, mbar_wait_relaxed(&scale_full[slot], (uint32_t)((n / NBUF) & 1), 20));
                        if (bw == 0) T_STAMP(1, n, 1);
#pragma unroll
                        for (int i = 0; i < RI; ++i) {
                            pacc[i] = racc[i] = 0.f;
                            const int row = row0 + 2 * i;
                            rscale[i] = (row < ncols && c0 + row < p.n_frames) ? rowinfo[row].x * fold_scale : 0.f;
                        }
                    }
                    convert(ring[u], kc);
                    if (++kc == n_kc) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ri_full[slot]);
                        if (bw == 0) T_STAMP(1, n, 2);
                        if (bw == 0) T_WAIT_FLUSH(1);
                        kc = 0;
                        ++n;
                        more = strip.next(sig, c0, ncols, skip, fresh);
                    }
                }
            }
        }
    } else {
        // ===================== scouts: frame maxima of the next tile ==================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        T_WAIT_DECL;
        const int sw = warp - IT_SCOUT_WARP0;      // 0..7
        for (int n = 0; strip.next(sig, c0, ncols, skip, fresh); ++n) {
            const int slot = n % NBUF;
            float4* ri = rowinfo2 + slot * NF;
            const float2* xs = p.spec + sig * p.ss;
            if (sw == 0) T_STAMP(0, n, 0);
            T_WAITED(0, mbar_wait_relaxed(&scale_empty[slot], (uint32_t)(((n / NBUF) & 1) ^ 1)));
            if (sw == 0) T_STAMP(0, n, 1);
            if (FRAMES_FAST) {
                // thread = (frame, one of NPART interleaved sets of 16-bin groups); lanes along frames
                constexpr int NPART = IT_SCOUT_THREADS / NF;
                const int st = sw * 32 + lane;
                const int row = st & (NF - 1), part = st / NF;
                const bool live = row < ncols && c0 + row < p.n_frames;
                const float2* col = xs + (c0 + (live ? row : 0)) * p.sf;
                float m = 0.f;
                for (int b0 = 16 * part; b0 < 2 * Q; b0 += 16 * NPART) {
                    float2 v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        v[e] = live ? __ldg(col + (int64_t)(b0 + e) * p.sb) : make_float2(0.f, 0.f);
                    if (b0 == 0) v[0].y = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        m = fmaxf(m, DECOMP ? norm2_finite(v[e])
                                            : abs2_finite(prep_bin<false>(v[e], p.pre_scale, p.pre_expo)));
                }
                // scratch floats 512.. are the scouts' (the builders use 0..511)
                float* sc = scratch + 512;
                sc[part * NF + row] = m;
                T_WAITED(1, named_bar_sync(2, IT_SCOUT_THREADS));
                if (part == 0) {
                    float ny = 0.f;
                    if (live && !ODD)
                        ny = prep_bin<DECOMP>(__ldg(col + (int64_t)Hf * p.sb), p.pre_scale, p.pre_expo).x *
                             p.edge_gain;
                    float mm = 0.f;
#pragma unroll
                    for (int k = 0; k < NPART; ++k) mm = fmaxf(mm, sc[k * NF + row]);
                    if (DECOMP) mm = decomp_bound(mm, p.pre_scale, p.pre_expo);
                    ri[row] = make_float4(live ? row_scale(mm) : 1.f, 0.f, 0.f, ny);
                }
                T_WAITED(1, named_bar_sync(2, IT_SCOUT_THREADS));
            } else {
                // warp = frame (bins contiguous), SR frames in flight per warp: one for Q = 128 (the
                // builders take long enough per tile, and the smaller loop is measurably faster: cfg4
                // 184 -> 174 us), two for Q <= 64 (short tiles: the scouts need the bytes in flight,
                // cfg5 462 -> 384 us)
                const int nj = Q / 16;             // 32-bin groups below the Nyquist bin
                auto scan_rows = [&](auto sr_c) {
                    constexpr int SR = decltype(sr_c)::value;
#pragma unroll 1
                    for (int r0 = SR * sw; r0 < NF; r0 += SR * IT_SCOUT_WARPS) {
                        float2 v[SR][8];
                        float2 vn[SR];
#pragma unroll
                        for (int r = 0; r < SR; ++r) {
                            const int row = r0 + r;
                            const bool live = row < ncols && c0 + row < p.n_frames;
                            const float2* xr = xs + (c0 + row) * p.sf;      // sb == 1
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                v[r][j] = (live && j < nj) ? __ldg(xr + j * 32 + lane) : make_float2(0.f, 0.f);
                            vn[r] = (live && lane == 0 && !ODD) ? __ldg(xr + Hf) : make_float2(0.f, 0.f);
                        }
#pragma unroll
                        for (int r = 0; r < SR; ++r) {
                            const int row = r0 + r;
                            if (lane == 0) v[r][0].y = 0.f;   // Im X[0] never reaches the output
                            float m = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                m = fmaxf(m, DECOMP ? norm2_finite(v[r][j])
                                                    : abs2_finite(prep_bin<false>(v[r][j], p.pre_scale, p.pre_expo)));
#pragma unroll
                            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                            if (DECOMP) m = decomp_bound(m, p.pre_scale, p.pre_expo);
                            if (lane == 0 && row < NF) {
                                const bool live = row < ncols && c0 + row < p.n_frames;
                                const float ny = prep_bin<DECOMP>(vn[r], p.pre_scale, p.pre_expo).x * p.edge_gain;
                                ri[row] = make_float4(live ? row_scale(m) : 1.f, 0.f, 0.f, live ? ny : 0.f);
                            }
                        }
                    }
                };
#ifdef BRV_T_SCOUT_FIXED2              // dev A/B switch
                scan_rows(std::integral_constant<int, 2>{});
#else
                if (Q <= 64) scan_rows(std::integral_constant<int, 2>{});
                else scan_rows(std::integral_constant<int, 1>{});
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&scale_full[slot]);
            if (sw == 0) T_STAMP(0, n, 2);
            if (sw == 0) T_WAIT_FLUSH(0);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)T_TMEM_COLS);
    }
}
