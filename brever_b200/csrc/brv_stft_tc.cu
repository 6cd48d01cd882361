// Tensor-core DFT contractions for sm_100a (tcgen05 + TMEM + TMA), fp32-grade.
//
//   D[row, col] = sum_k A[row, k] * Bt[col, k]          (one 128 x 256 tile per CTA)
//
// MODE_FWD  STFT.forward (brever/modules/stft.py:59-89)
//   A   : 128 overlapping frames of one signal, built on the fly from the raw
//         samples (centre / right zero padding by predication — never
//         materialised in HBM);
//   Bt  : DFT basis with the window and 1/sqrt(sum w^2) folded in;
//   out : frame-major complex64 (signal, frame, bin) — the memory torch.stft
//         returns — after |X|^(c-1) compression and scale_factor.
// MODE_INV  the contraction inside STFT.backward (stft.py:111-136)
//   A   : 128 spectrogram frames (any strides), /scale_factor and |X|^(1/c-1)
//         applied in the loader;
//   Bt  : inverse real-DFT basis with the Hermitian weights, 1/N, the window and
//         sqrt(sum w^2) folded in;
//   out : windowed time-domain frames (signal, frame, n_fft) for the overlap-add.
//
// Precision: every A row is scaled by its own power of two and split into two
// fp16 planes a = a_hi + a_lo (22 significant bits); the basis is split the same
// way once at plan creation.  Three tensor-core products per k-step
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, the dropped a_lo*b_lo is 2^-22) accumulate
// in fp32 in tensor memory.  Measured max error vs float64: ~2e-6 of max|X|.
//
// Columns are packed so a one-sided spectrum of even n_fft is exactly n_fft wide:
// col 0 = Re X[0], col 1 = Re X[N/2] (their imaginary parts are identically zero /
// ignored), cols 2q, 2q+1 = Re, Im X[q].
//
// Warp roles (192 threads): warp 0 = TMA producer (basis k-blocks, SWIZZLE_64B),
// warp 1 = TMEM owner + MMA issuer (one elected lane), warps 2-5 = A-tile
// builders (global -> registers (prefetched) -> split -> swizzled smem), then the
// epilogue (TMEM -> registers -> per-warp smem transpose -> coalesced stores).
// Two CTAs are resident per SM (96 KB smem, 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "brv_common.cuh"
#include "brv_tc_ptx.cuh"

namespace {

constexpr int MODE_FWD = 0, MODE_INV = 1;
constexpr int TILE_M = 128;          // rows (frames) per CTA (UMMA M)
constexpr int TILE_N = 256;          // output columns per CTA (UMMA N)
constexpr int BK = 32;               // k per stage: 64-byte rows (SWIZZLE_64B)
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int A_PLANE = TILE_M * BK * 2;   // 8 KB  (one fp16 plane of the A tile)
constexpr int B_PLANE = TILE_N * BK * 2;   // 16 KB
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;   // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int NUM_THREADS = 192;
constexpr int LOADER_THREADS = 128;
constexpr int MAX_BLOCKS = 160;      // hop blocks spanned by one tile (127 + n_fft/hop)
constexpr int EPI_PITCH = 36;        // floats per row of the per-warp transpose tile

struct TcParams {
    // forward: raw signal
    const float* x;
    int64_t x_stride, samples;
    // inverse: spectrogram with element strides (complex units)
    const float2* spec;
    int64_t ss, sb, sf;
    float pre_scale, pre_expo;   // 1/scale_factor, 1/c - 1
    // output
    float* out;                  // fwd: (sig, T, n_bins) complex ; inv: (sig, T, n_fft) floats
    int64_t n_frames;
    int n_fft, hop, n_bins;
    int left;                    // frame t starts `left` samples before t * hop (n_fft / 2 when centred)
    int k_blocks;                // ceil(K / BK)
    int col_blocks;              // packed columns / TILE_N
    int cols_pad;                // col_blocks * TILE_N (lo plane starts at this row)
    int tiles_per_signal;
    float basis_scale_inv;       // 1 / sB
    float post_scale;            // fwd: scale_factor
    float post_expo;             // fwd: compression_factor - 1
};

using namespace brv_ptx;

// ---- A-operand sources ----------------------------------------------------------
// Forward: 32 consecutive samples of frame t starting at k0.
struct FrameSource {
    const float* xs;
    int64_t samples, frame_start;     // sample index of k = 0 (may be negative)
    int n_fft;
    bool aligned;
    __device__ __forceinline__ void load(int kb, float* v) const {
        const int64_t i0 = frame_start + (int64_t)kb * BK;
        const int k_left = n_fft - kb * BK;
        if (aligned && i0 >= 0 && i0 + BK <= samples && k_left >= BK) {
#pragma unroll
            for (int c = 0; c < BK / 4; ++c) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(xs + i0) + c);
                v[4 * c] = f.x; v[4 * c + 1] = f.y; v[4 * c + 2] = f.z; v[4 * c + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < BK; ++j) {
                const int64_t idx = i0 + j;
                v[j] = (j < k_left && idx >= 0 && idx < samples) ? __ldg(xs + idx) : 0.f;
            }
        }
    }
};
// Inverse: 16 packed complex bins of spectrogram frame t starting at bin kb*16.
struct SpectrumSource {
    const float2* row;                // X[sig, 0, t]
    int64_t sb;                       // bin stride (complex units)
    int half;                         // n_fft / 2
    float pre_scale, pre_expo;
    bool live;                        // t < n_frames
    __device__ __forceinline__ void load(int kb, float* v) const {
#pragma unroll
        for (int j = 0; j < BK / 2; ++j) {
            const int qb = kb * (BK / 2) + j;
            float re = 0.f, im = 0.f;
            if (live && qb < half) {
                const float2 c = __ldg(row + (int64_t)qb * sb);
                re = c.x * pre_scale;
                im = c.y * pre_scale;
                if (qb == 0) {
                    // packed pair: Re X[0], Re X[N/2].  The reference decompresses the
                    // complex values (their imaginary parts count towards |X|) and the
                    // c2r inverse then ignores the imaginary parts of both bins.
                    float2 ny = __ldg(row + (int64_t)half * sb);
                    ny.x *= pre_scale;
                    ny.y *= pre_scale;
                    if (pre_expo != 0.f) {
                        compress(re, im, pre_expo);
                        compress(ny.x, ny.y, pre_expo);
                    }
                    im = ny.x;
                } else if (pre_expo != 0.f) {
                    compress(re, im, pre_expo);
                }
            }
            v[2 * j] = re;
            v[2 * j + 1] = im;
        }
    }
};

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 2)
dft_tc_kernel(const __grid_constant__ CUtensorMap basis_map, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ uint32_t block_max[MAX_BLOCKS];

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    // blockIdx.x = (signal * tiles_per_signal + tile) * col_blocks + cb
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
    if (MODE == MODE_FWD)
        for (int j = threadIdx.x; j < MAX_BLOCKS; j += NUM_THREADS) block_max[j] = 0u;
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
        const int row = q * 32 + lane;             // row within the tile == TMEM lane
        const int lt = (warp - 2) * 32 + lane;     // 0..127 builder-thread index
        const int64_t t = t0 + row;
        const int half = p.n_fft / 2;
        float scale;

        FrameSource fsrc;
        SpectrumSource ssrc;
        if (MODE == MODE_FWD) {
            const float* xs = p.x + sig * p.x_stride;
            // --- per-hop-block maxima over the samples this tile touches: one
            //     coalesced pass, warp-reduced, then max over the blocks of a frame
            const int64_t span0 = t0 * p.hop - p.left;                 // first sample (may be < 0)
            const int span_len = (TILE_M - 1) * p.hop + p.n_fft;
            const bool vec = ((((uintptr_t)xs) & 15) == 0) && ((span0 & 3) == 0) && ((p.hop & 3) == 0);
            if (vec) {
                // warp-uniform trip count: the warp votes inside the loop
                for (int base = 0; base < span_len; base += 4 * LOADER_THREADS) {
                    const int i = base + 4 * lt;
                    const int64_t idx = span0 + i;
                    float m = 0.f;
                    int j = -1;
                    if (i < span_len) {
                        j = i / p.hop;                                  // 4 | hop: one block per float4
                        if (idx >= 0 && idx + 4 <= p.samples) {
                            const float4 f = __ldg(reinterpret_cast<const float4*>(xs + idx));
                            m = fmaxf(fmaxf(finite_abs(f.x), finite_abs(f.y)),
                                      fmaxf(finite_abs(f.z), finite_abs(f.w)));
                        } else {
                            for (int e = 0; e < 4; ++e)
                                if (idx + e >= 0 && idx + e < p.samples)
                                    m = fmaxf(m, finite_abs(__ldg(xs + idx + e)));
                        }
                    }
                    const int j0 = __shfl_sync(0xffffffffu, j, 0);
                    if (__all_sync(0xffffffffu, j == j0)) {
                        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
                        if (lane == 0 && j0 >= 0) atomicMax(&block_max[j0], wm);
                    } else if (j >= 0) {
                        atomicMax(&block_max[j], __float_as_uint(m));
                    }
                }
            } else {
                for (int i = lt; i < span_len; i += LOADER_THREADS) {
                    const int64_t idx = span0 + i;
                    if (idx >= 0 && idx < p.samples)
                        atomicMax(&block_max[i / p.hop], __float_as_uint(finite_abs(__ldg(xs + idx))));
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(LOADER_THREADS) : "memory");
            const int n_span = (p.n_fft + p.hop - 1) / p.hop;          // blocks per frame
            uint32_t mx = 0u;
            for (int j = 0; j < n_span; ++j) mx = max(mx, block_max[row + j]);
            scale = row_scale(__uint_as_float(mx));
            fsrc.xs = xs;
            fsrc.samples = p.samples;
            fsrc.frame_start = t * p.hop - p.left;
            fsrc.n_fft = p.n_fft;
            fsrc.aligned = ((((uintptr_t)xs) & 15) == 0) && ((fsrc.frame_start & 3) == 0);
        } else {
            ssrc.live = t < p.n_frames;
            ssrc.row = p.spec + sig * p.ss + (ssrc.live ? t : 0) * p.sf;
            ssrc.sb = p.sb;
            ssrc.half = half;
            ssrc.pre_scale = p.pre_scale;
            ssrc.pre_expo = p.pre_expo;
            // --- row maximum (after the loader's own pre-processing)
            float m = 0.f;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                float v[BK];
                ssrc.load(kb, v);
#pragma unroll
                for (int j = 0; j < BK; ++j) m = fmaxf(m, finite_abs(v[j]));
            }
            scale = row_scale(m);
        }

        const uint32_t sw = (uint32_t)((row >> 1) & 3);
        float cur[BK], nxt[BK];
        if (MODE == MODE_FWD) fsrc.load(0, cur); else ssrc.load(0, cur);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            if (kb + 1 < p.k_blocks) {             // prefetch the next k-block into registers
                if (MODE == MODE_FWD) fsrc.load(kb + 1, nxt); else ssrc.load(kb + 1, nxt);
            }
            uint32_t hi[BK / 2], lo[BK / 2];
#pragma unroll
            for (int j = 0; j < BK; j += 2) {
                const float a0 = cur[j] * scale, a1 = cur[j + 1] * scale;
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
#pragma unroll
            for (int j = 0; j < BK; ++j) cur[j] = nxt[j];
        }

        // ---- epilogue: TMEM -> registers -> per-warp transpose -> coalesced stores ----
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        // all MMAs have retired: the pipeline stages are free, reuse them as staging
        float* stage = reinterpret_cast<float*>(tiles) + (size_t)(warp - 2) * 32 * EPI_PITCH;
        const float g0 = p.basis_scale_inv / scale;                // undo the operand scalings
        const int64_t row0 = t0 + q * 32;                          // first row of this warp
        const int pitch = (MODE == MODE_FWD) ? 2 * p.n_bins : p.n_fft;   // floats per output row
        float* obase = p.out + (sig * p.n_frames) * (int64_t)pitch;
#pragma unroll 1
        for (int c = 0; c < TILE_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            const int col0 = cb * TILE_N + c * 32;
            if (col0 >= p.n_fft) break;                            // padded column block (uniform)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + 4 * j) =
                    make_float4(__uint_as_float(r[4 * j]) * g0, __uint_as_float(r[4 * j + 1]) * g0,
                                __uint_as_float(r[4 * j + 2]) * g0, __uint_as_float(r[4 * j + 3]) * g0);
            __syncwarp();
            // 16 lanes cover one 32-float row (128 contiguous bytes): two rows per pass
            const int cp = (lane & 15) * 2;                        // column pair inside the chunk
#pragma unroll 4
            for (int rr = 0; rr < 32; rr += 2) {
                const int rl = rr + (lane >> 4);
                const int64_t tr = row0 + rl;
                float2 v = *reinterpret_cast<const float2*>(stage + rl * EPI_PITCH + cp);
                const int col = col0 + cp;
                if (tr >= p.n_frames || col >= p.n_fft) continue;
                float* orow = obase + tr * pitch;
                if (MODE == MODE_FWD) {
                    if (col == 0) {        // packed pair: Re X[0] and Re X[N/2], purely real
                        float d = v.x, ny = v.y;
                        if (p.post_expo != 0.f) {
                            d = compress_real(d, p.post_expo);
                            ny = compress_real(ny, p.post_expo);
                        }
                        *reinterpret_cast<float2*>(orow) = make_float2(d * p.post_scale, 0.f);
                        *reinterpret_cast<float2*>(orow + 2 * half) = make_float2(ny * p.post_scale, 0.f);
                    } else {
                        if (p.post_expo != 0.f) compress(v.x, v.y, p.post_expo);
                        *reinterpret_cast<float2*>(orow + col) =
                            make_float2(v.x * p.post_scale, v.y * p.post_scale);
                    }
                } else {
                    *reinterpret_cast<float2*>(orow + col) = v;
                }
            }
            __syncwarp();
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

struct TcBasis {
    __half* data = nullptr;   // [2 planes][cols_pad][k_pad]
    CUtensorMap map;
    int cols_pad = 0, k_pad = 0;
    float scale_inv = 1.f;
};
struct TcPlan {
    TcBasis fwd, inv;
};

// Split `value(col, k)` (col < n_cols, k < n_k) into scaled fp16 hi/lo planes and
// describe it with a SWIZZLE_64B tensor map of (BK x TILE_N) boxes.
template <class F>
int build_basis(TcBasis* b, int n_cols, int n_k, F value) {
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return BRV_ERR_UNSUPPORTED;
    b->cols_pad = (int)brv_ceil_div(n_cols, TILE_N) * TILE_N;
    b->k_pad = (int)brv_ceil_div(n_k, BK) * BK;
    double mx = 0;
    for (int c = 0; c < n_cols; ++c)
        for (int k = 0; k < n_k; ++k) mx = fmax(mx, fabs(value(c, k)));
    if (!(mx > 0)) return BRV_ERR_UNSUPPORTED;
    int e;
    frexp(mx, &e);                                     // mx = m * 2^e, m in [0.5, 1)
    const double sB = ldexp(1.0, 13 - e);              // mx * sB in [2^12, 2^13)
    b->scale_inv = (float)(1.0 / sB);
    std::vector<__half> host((size_t)2 * b->cols_pad * b->k_pad, __float2half_rn(0.f));
    for (int c = 0; c < n_cols; ++c)
        for (int k = 0; k < n_k; ++k) {
            const double v = value(c, k) * sB;
            const __half h = __float2half_rn((float)v);
            const __half l = __float2half_rn((float)(v - (double)__half2float(h)));
            host[(size_t)c * b->k_pad + k] = h;
            host[((size_t)b->cols_pad + c) * b->k_pad + k] = l;
        }
    if (cudaMalloc((void**)&b->data, host.size() * sizeof(__half)) != cudaSuccess)
        return brv_fail_cuda(cudaGetLastError(), "cudaMalloc(tensor-core basis)");
    if (cudaMemcpy(b->data, host.data(), host.size() * sizeof(__half), cudaMemcpyHostToDevice) !=
        cudaSuccess)
        return brv_fail_cuda(cudaGetLastError(), "cudaMemcpy(tensor-core basis)");
    cuuint64_t dims[2] = {(cuuint64_t)b->k_pad, (cuuint64_t)(2 * b->cols_pad)};
    cuuint64_t strides[1] = {(cuuint64_t)b->k_pad * sizeof(__half)};
    cuuint32_t box[2] = {BK, TILE_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&b->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b->data, dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        return brv_fail(BRV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return BRV_OK;
}

void fill_common(TcParams& prm, const brv_stft_plan* p, const TcBasis& b, int64_t n_frames) {
    prm.n_frames = n_frames;
    prm.n_fft = p->n_fft;
    prm.hop = p->hop;
    prm.left = brv_left(p);
    prm.n_bins = p->n_bins;
    prm.k_blocks = b.k_pad / BK;
    prm.col_blocks = b.cols_pad / TILE_N;
    prm.cols_pad = b.cols_pad;
    prm.tiles_per_signal = (int)brv_ceil_div(n_frames, TILE_M);
    prm.basis_scale_inv = b.scale_inv;
}

}  // namespace

bool brv_tc_supported(const brv_stft_plan* p) { return p->tc_fwd != nullptr; }

int brv_tc_plan_init(brv_stft_plan* p, const std::vector<double>& fwd,
                     const std::vector<double>& inv) {
    const int N = p->n_fft;
    // one-sided, even n_fft: packed width == n_fft; tiles must fit the hop-block table
    if (!p->onesided || (N % 2) != 0 || N < 32 || N > 4096) return BRV_OK;
    if (TILE_M - 1 + (N + p->hop - 1) / p->hop > MAX_BLOCKS) return BRV_OK;
    if (!encode_tiled()) return BRV_OK;   // no tensor-map support: generic path only

    const int F = p->n_bins, half = N / 2;
    // packed index c -> row / column of the interleaved (re, im) bases
    auto unpack = [half](int c) { return c == 0 ? 0 : (c == 1 ? 2 * half : c); };
    TcPlan* tp = new TcPlan();
    // forward: Bt[col c][k] = fwd[k][unpack(c)]
    int rc = build_basis(&tp->fwd, N, N, [&](int c, int k) {
        return fwd[(size_t)k * 2 * F + unpack(c)];
    });
    // inverse: Bt[col n][k c] = inv[unpack(c)][n]
    if (rc == BRV_OK)
        rc = build_basis(&tp->inv, N, N, [&](int n, int c) {
            return inv[(size_t)unpack(c) * N + n];
        });
    if (rc == BRV_OK &&
        (cudaFuncSetAttribute(dft_tc_kernel<MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SMEM_BYTES) != cudaSuccess ||
         cudaFuncSetAttribute(dft_tc_kernel<MODE_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SMEM_BYTES) != cudaSuccess))
        rc = brv_fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(dft_tc_kernel)");
    if (rc != BRV_OK) {
        cudaFree(tp->fwd.data);
        cudaFree(tp->inv.data);
        delete tp;
        return rc == BRV_ERR_UNSUPPORTED ? BRV_OK : rc;
    }
    p->tc_fwd = tp;
    return BRV_OK;
}

void brv_tc_plan_free(brv_stft_plan* p) {
    TcPlan* tp = (TcPlan*)p->tc_fwd;
    if (tp) {
        cudaFree(tp->fwd.data);
        cudaFree(tp->inv.data);
        delete tp;
        p->tc_fwd = nullptr;
    }
}

int brv_tc_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                        int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st) {
    const TcPlan* tp = (const TcPlan*)p->tc_fwd;
    TcParams prm = {};
    fill_common(prm, p, tp->fwd, n_frames);
    prm.x = x;
    prm.x_stride = x_stride;
    prm.samples = samples;
    prm.out = reinterpret_cast<float*>(out);
    prm.post_scale = (float)p->scale;
    prm.post_expo = (float)(p->compression - 1.0);
    const int64_t grid = n_sig * prm.tiles_per_signal * prm.col_blocks;
    BRV_REQUIRE(grid < (1LL << 31), "too many tiles (%lld)", (long long)grid);
    dft_tc_kernel<MODE_FWD><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(tp->fwd.map, prm);
    BRV_LAUNCH_CHECK("dft_tc_kernel<fwd>");
    return BRV_OK;
}

// Spectrogram (any strides) -> windowed time-domain frames (n_sig, T, n_fft) fp32.
int brv_tc_spec_to_frames(const brv_stft_plan* p, const float2* X, int64_t ss, int64_t sb,
                          int64_t sf, int64_t n_sig, int64_t n_frames, float* frames,
                          cudaStream_t st) {
    const TcPlan* tp = (const TcPlan*)p->tc_fwd;
    TcParams prm = {};
    fill_common(prm, p, tp->inv, n_frames);
    prm.spec = X;
    prm.ss = ss;
    prm.sb = sb;
    prm.sf = sf;
    prm.pre_scale = (float)(1.0 / p->scale);
    prm.pre_expo = (float)(1.0 / p->compression - 1.0);
    prm.out = frames;
    const int64_t grid = n_sig * prm.tiles_per_signal * prm.col_blocks;
    BRV_REQUIRE(grid < (1LL << 31), "too many tiles (%lld)", (long long)grid);
    dft_tc_kernel<MODE_INV><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(tp->inv.map, prm);
    BRV_LAUNCH_CHECK("dft_tc_kernel<inv>");
    return BRV_OK;
}
