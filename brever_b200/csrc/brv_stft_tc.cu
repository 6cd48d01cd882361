// Tensor-core STFT for sm_100a: framing + windowing + DFT as one tcgen05
// contraction with fp32-grade accuracy.
//
//   D[frame, col] = sum_k A[frame, k] * Bt[col, k]
//
//   A  : 128 overlapping frames of one signal, built on the fly from the raw
//        samples (centre / right zero padding by predication, never materialised
//        in HBM), scaled per frame by a power of two and split into two fp16
//        planes  a = a_hi + a_lo  (22 significant bits);
//   Bt : the DFT basis with the window and 1/sqrt(sum w^2) folded in, transposed
//        (K-major), split the same way at plan creation, streamed by TMA
//        (SWIZZLE_64B) from L2 where it stays resident;
//   D  : fp32 accumulators in tensor memory, three products per k-step
//        (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo; the dropped a_lo*b_lo is 2^-22).
//
// Output columns are packed so that one-sided spectra of even n_fft need exactly
// n_fft columns: col 0 = Re X[0], col 1 = Re X[N/2] (their imaginary parts are
// identically zero), cols 2q, 2q+1 = Re, Im X[q].  The epilogue unpacks, applies
// the per-frame scale, |X|^(c-1) compression and scale_factor, and writes the
// frame-major complex64 layout torch.stft produces.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer (one elected lane), warps 2-5 = A-tile builders, then epilogue.
// Two CTAs are resident per SM (96 KB smem, 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
//
// Reference semantics: brever/modules/stft.py:59-89.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "brv_common.cuh"

namespace {

constexpr int TILE_M = 128;          // frames per CTA (UMMA M)
constexpr int TILE_N = 256;          // packed output columns per CTA (UMMA N)
constexpr int BK = 32;               // k per stage: 64-byte rows (SWIZZLE_64B)
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int A_PLANE = TILE_M * BK * 2;   // 8 KB  (one fp16 plane of the A tile)
constexpr int B_PLANE = TILE_N * BK * 2;   // 16 KB
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;   // 48 KB
constexpr int NUM_THREADS = 192;
constexpr int LOADER_THREADS = 128;
constexpr int MAX_BLOCKS = 160;      // hop blocks spanned by one tile (127 + n_fft/hop)

struct TcParams {
    const float* x;
    int64_t x_stride, samples;
    float2* out;
    int64_t n_frames;
    int n_fft, hop, n_bins;
    int k_blocks;            // ceil(n_fft / BK)
    int col_blocks;          // packed columns / TILE_N
    int cols_pad;            // col_blocks * TILE_N (lo plane starts at this row)
    int tiles_per_signal;
    float basis_scale_inv;   // 1 / sB
    float post_scale;        // scale_factor
    float expo;              // compression_factor - 1
};

// ---- PTX helpers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long start = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - start > 4000000000LL) __trap();   // never hang the device
    }
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_64B operand tile: rows of 64 bytes, 8-row groups 512 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)0 << 16;                                // LBO (unused: one swizzle atom along K)
    d |= (uint64_t)(512 >> 4) << 32;                       // SBO: 8 rows * 64 B
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                                // layout: SWIZZLE_64B
    return d;
}
// kind::f16, fp16 x fp16 -> fp32, both operands K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// power of two s with max*s in [2^13, 2^14): fp16 keeps 11 bits of a_hi and the
// residual a_lo stays far above the fp16 subnormal floor.
__device__ __forceinline__ float frame_scale(float mx) {
    if (!(mx > 0.f)) return 1.f;
    int e = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;   // floor(log2(mx)), normal range
    if (e < -100) e = -100;
    int se = 13 - e;                                            // in [-114, 113]
    return __uint_as_float((uint32_t)(se + 127) << 23);
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
stft_tc_kernel(const __grid_constant__ CUtensorMap basis_map, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float block_max[MAX_BLOCKS];

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    // work decomposition: blockIdx.x = (signal * tiles_per_signal + tile) * col_blocks + cb
    const int cb = blockIdx.x % p.col_blocks;
    const int64_t tile_id = blockIdx.x / p.col_blocks;
    const int64_t sig = tile_id / p.tiles_per_signal;
    const int64_t t0 = (tile_id % p.tiles_per_signal) * TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + LOADER_THREADS / 32);   // TMA arrive + 4 builder warps
            mbar_init(&empty_bar[s], 1);                        // one tcgen05.commit
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 256 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_slot)),
                     "r"(TILE_N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer: basis k-blocks =====================
        if (elect_one()) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], 2 * B_PLANE);
                uint8_t* st = tiles + (size_t)s * STAGE_BYTES;
                tma_load_2d(smem_u32(st + 2 * A_PLANE), &basis_map, &full_bar[s], kb * BK,
                            cb * TILE_N);
                tma_load_2d(smem_u32(st + 2 * A_PLANE + B_PLANE), &basis_map, &full_bar[s],
                            kb * BK, p.cols_pad + cb * TILE_N);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ======================================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(TILE_M, TILE_N);
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                const uint32_t a_lo = a_hi + A_PLANE;
                const uint32_t b_hi = a_hi + 2 * A_PLANE;
                const uint32_t b_lo = b_hi + B_PLANE;
#pragma unroll
                for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                    const uint32_t off = ks * UMMA_K * 2;     // bytes along K inside the 64 B row
                    const uint64_t dah = umma_desc_sw64(a_hi + off), dal = umma_desc_sw64(a_lo + off);
                    const uint64_t dbh = umma_desc_sw64(b_hi + off), dbl = umma_desc_sw64(b_lo + off);
                    umma_f16(tmem_base, dah, dbh, idesc, (kb | ks) != 0);
                    umma_f16(tmem_base, dal, dbh, idesc, 1);
                    umma_f16(tmem_base, dah, dbl, idesc, 1);
                }
                umma_commit(&empty_bar[s]);        // frees the stage once these MMAs retire
            }
            umma_commit(&accum_bar);               // accumulator complete
        }
    } else {
        // ===================== A-tile builders, then epilogue ===================
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;             // frame within the tile == TMEM lane
        const int lt = (warp - 2) * 32 + lane;     // 0..127 builder-thread index
        const int64_t t = t0 + row;
        const float* xs = p.x + sig * p.x_stride;
        const int half = p.n_fft / 2;

        // per-hop-block maxima over the samples this tile touches
        const int n_span = (p.n_fft + p.hop - 1) / p.hop;          // blocks per frame
        const int n_blocks = TILE_M - 1 + n_span;
        for (int j = lt; j < n_blocks; j += LOADER_THREADS) {
            const int64_t i0 = (t0 + j) * p.hop - half;
            float m = 0.f;
            for (int i = 0; i < p.hop; ++i) {
                const int64_t idx = i0 + i;
                if (idx >= 0 && idx < p.samples) {
                    float a = fabsf(__ldg(xs + idx));
                    if (a <= 3.0e38f) m = fmaxf(m, a);           // ignore inf / nan
                }
            }
            block_max[j] = m;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(LOADER_THREADS) : "memory");
        float mx = 0.f;
        for (int j = 0; j < n_span; ++j) mx = fmaxf(mx, block_max[row + j]);
        const float scale = frame_scale(mx);

        const int64_t frame_start = t * p.hop - half;              // sample index of k = 0
        const bool aligned = ((((uintptr_t)xs) & 15) == 0) && ((frame_start & 3) == 0);
        const uint32_t sw = (uint32_t)((row >> 1) & 3);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            float v[BK];
            const int64_t i0 = frame_start + (int64_t)kb * BK;
            const int k_left = p.n_fft - kb * BK;                  // valid k in this block
            if (aligned && i0 >= 0 && i0 + BK <= p.samples && k_left >= BK) {
#pragma unroll
                for (int c = 0; c < BK / 4; ++c) {
                    float4 f = __ldg(reinterpret_cast<const float4*>(xs + i0) + c);
                    v[4 * c] = f.x; v[4 * c + 1] = f.y; v[4 * c + 2] = f.z; v[4 * c + 3] = f.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < BK; ++j) {
                    const int64_t idx = i0 + j;
                    v[j] = (j < k_left && idx >= 0 && idx < p.samples) ? __ldg(xs + idx) : 0.f;
                }
            }
            uint32_t hi[BK / 2], lo[BK / 2];
#pragma unroll
            for (int j = 0; j < BK; j += 2) {
                const float a0 = v[j] * scale, a1 = v[j + 1] * scale;
                const __half2 h = __floats2half2_rn(a0, a1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                hi[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j / 2] = *reinterpret_cast<const uint32_t*>(&l);
            }
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* a_hi = tiles + (size_t)s * STAGE_BYTES + row * (BK * 2);
            uint8_t* a_lo = a_hi + A_PLANE;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                          // 4 x 16-byte chunks per row
                const uint32_t dst = (uint32_t)((c ^ sw) * 16);
                *reinterpret_cast<uint4*>(a_hi + dst) =
                    make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                *reinterpret_cast<uint4*>(a_lo + dst) =
                    make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            fence_proxy_async();                                   // generic -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }

        // ---- epilogue: TMEM -> registers -> unpack / scale / compress -> HBM ----
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        const float g0 = p.basis_scale_inv / scale;   // undo the operand scalings
        const float post = p.post_scale;             // scale_factor applies after compression
        const bool live = t < p.n_frames;
        float2* orow = p.out + (sig * p.n_frames + (live ? t : 0)) * (int64_t)p.n_bins;
#pragma unroll 1
        for (int c = 0; c < TILE_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            if (!live) continue;
            const int col0 = cb * TILE_N + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const int col = col0 + i;
                if (col >= p.n_fft) break;                         // packed width == n_fft
                float re = __uint_as_float(r[i]) * g0, im = __uint_as_float(r[i + 1]) * g0;
                if (col == 0) {
                    // packed pair: Re X[0] and Re X[N/2], both purely real
                    float d = re, ny = im;
                    if (p.expo != 0.f) {
                        d = d != 0.f ? d * powf(fabsf(d), p.expo) : 0.f;
                        ny = ny != 0.f ? ny * powf(fabsf(ny), p.expo) : 0.f;
                    }
                    orow[0] = make_float2(d * post, 0.f);
                    orow[half] = make_float2(ny * post, 0.f);
                } else {
                    if (p.expo != 0.f) {
                        const float m2 = re * re + im * im;
                        const float g = m2 > 0.f ? powf(m2, 0.5f * p.expo) : 0.f;
                        re *= g;
                        im *= g;
                    }
                    orow[col >> 1] = make_float2(re * post, im * post);
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TILE_N)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

struct TcPlan {
    __half* basis;        // [2 planes][cols_pad][k_pad]
    CUtensorMap map;
    int cols_pad, k_pad;
    float scale_inv;
};

}  // namespace

bool brv_tc_supports_forward(const brv_stft_plan* p) {
    return p->tc_fwd != nullptr;
}

int brv_tc_plan_init(brv_stft_plan* p, const std::vector<double>& fwd,
                     const std::vector<double>& inv) {
    (void)inv;
    const int N = p->n_fft;
    // one-sided, even n_fft: packed width == n_fft; tiles must fit the hop-block table
    if (!p->onesided || (N % 2) != 0 || N < 32 || N > 4096) return BRV_OK;
    if (TILE_M - 1 + (N + p->hop - 1) / p->hop > MAX_BLOCKS) return BRV_OK;
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return BRV_OK;   // driver too old for tensor maps: generic path only

    const int F = p->n_bins;
    const int cols_pad = (int)brv_ceil_div(N, TILE_N) * TILE_N;
    const int k_pad = (int)brv_ceil_div(N, BK) * BK;
    double mx = 0;
    for (double v : fwd) mx = fmax(mx, fabs(v));
    if (!(mx > 0)) return BRV_OK;
    int e;
    frexp(mx, &e);                                     // mx = m * 2^e, m in [0.5, 1)
    const double sB = ldexp(1.0, 13 - e);              // mx * sB in [2^12, 2^13)
    std::vector<__half> host((size_t)2 * cols_pad * k_pad, __float2half_rn(0.f));
    for (int c = 0; c < N; ++c) {
        // packed column c -> column of the interleaved (re, im) basis
        int src = (c == 0) ? 0 : (c == 1) ? 2 * (N / 2) : c;
        for (int k = 0; k < N; ++k) {
            const double v = fwd[(size_t)k * 2 * F + src] * sB;
            const __half h = __float2half_rn((float)v);
            const __half l = __float2half_rn((float)(v - (double)__half2float(h)));
            host[((size_t)c) * k_pad + k] = h;
            host[((size_t)cols_pad + c) * k_pad + k] = l;
        }
    }
    TcPlan* tp = new TcPlan();
    tp->cols_pad = cols_pad;
    tp->k_pad = k_pad;
    tp->scale_inv = (float)(1.0 / sB);
    if (cudaMalloc((void**)&tp->basis, host.size() * sizeof(__half)) != cudaSuccess) {
        delete tp;
        return brv_fail_cuda(cudaGetLastError(), "cudaMalloc(tc basis)");
    }
    cudaMemcpy(tp->basis, host.data(), host.size() * sizeof(__half), cudaMemcpyHostToDevice);
    cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)(2 * cols_pad)};
    cuuint64_t strides[1] = {(cuuint64_t)k_pad * sizeof(__half)};
    cuuint32_t box[2] = {BK, TILE_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tp->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, tp->basis, dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        cudaFree(tp->basis);
        delete tp;
        return brv_fail(BRV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    }
    if (cudaFuncSetAttribute(stft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             STAGES * STAGE_BYTES + 1024) != cudaSuccess) {
        cudaFree(tp->basis);
        delete tp;
        return brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(stft_tc_kernel)");
    }
    p->tc_fwd = tp;
    p->tc_fwd_cols = cols_pad;
    return BRV_OK;
}

void brv_tc_plan_free(brv_stft_plan* p) {
    TcPlan* tp = (TcPlan*)p->tc_fwd;
    if (tp) {
        cudaFree(tp->basis);
        delete tp;
        p->tc_fwd = nullptr;
    }
}

int brv_tc_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                        int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st) {
    const TcPlan* tp = (const TcPlan*)p->tc_fwd;
    TcParams prm;
    prm.x = x;
    prm.x_stride = x_stride;
    prm.samples = samples;
    prm.out = out;
    prm.n_frames = n_frames;
    prm.n_fft = p->n_fft;
    prm.hop = p->hop;
    prm.n_bins = p->n_bins;
    prm.k_blocks = tp->k_pad / BK;
    prm.col_blocks = tp->cols_pad / TILE_N;
    prm.cols_pad = tp->cols_pad;
    prm.tiles_per_signal = (int)brv_ceil_div(n_frames, TILE_M);
    prm.basis_scale_inv = tp->scale_inv;
    prm.post_scale = (float)p->scale;
    prm.expo = (float)(p->compression - 1.0);
    const int64_t grid = n_sig * prm.tiles_per_signal * prm.col_blocks;
    BRV_REQUIRE(grid < (1LL << 31), "too many tiles (%lld)", (long long)grid);
    stft_tc_kernel<<<(unsigned)grid, NUM_THREADS, STAGES * STAGE_BYTES + 1024, st>>>(tp->map, prm);
    BRV_LAUNCH_CHECK("stft_tc_kernel");
    return BRV_OK;
}
