// PTX wrappers shared by the tensor-core kernels (sm_100a: mbarrier, TMA,
// tcgen05 MMA / TMEM) and the fp16 hi/lo operand-splitting helpers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace brv_ptx {

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
    long long start = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spins & 1023) == 1023) {                        // never hang the device
            if (start == 0) start = clock64();
            else if (clock64() - start > 4000000000LL) __trap();
        }
    }
}
#ifndef BRV_WAIT_NAP_MAX
#define BRV_WAIT_NAP_MAX 256u
#endif
// Wait used by the roles that are NOT on the critical path of the CUDA cores (TMA producer,
// MMA issuer, epilogue / scout warps waiting for work): backs off with nanosleep so the
// polling does not take issue slots from the warps that do the arithmetic.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok;
}
// exponential back-off: a short wait is answered at once, a long one (a role that is a tile ahead
// of its producer) polls a few times per microsecond instead of every ~100 ns -- ~20 waiting warps
// polling at the short period took a quarter of the SM's issue slots.  Kept small (one counter as
// the never-hang-the-device guard): the warp-specialised kernels wait at ~25 sites each and their
// code size shows up as instruction-fetch stalls.  (An out-of-line polling loop does not compile
// inside the setmaxnreg regions: ptxas C7600.)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns = 64) {
    const uint32_t addr = smem_u32(bar);
    uint32_t nap = ns;
    for (uint32_t spins = 0; !mbar_try_wait(addr, parity); ++spins) {
        __nanosleep(nap);
        nap = min(nap * 2u, BRV_WAIT_NAP_MAX);
        if (spins == (1u << 24)) __trap();                   // ~4 s at the capped nap: a lost arrival
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
    d |= (uint64_t)(512 >> 4) << 32;                       // SBO: 8 rows * 64 B (LBO unused)
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                                // layout: SWIZZLE_64B
    return d;
}
// kind::f16, fp16 x fp16 -> fp32, both operands K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

// power of two s with max*s in [2^13, 2^14): fp16 keeps 11 bits in a_hi and the
// residual a_lo stays far above the fp16 subnormal floor.
__device__ __forceinline__ float row_scale(float mx) {
    if (!(mx > 0.f)) return 1.f;
    int e = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;   // floor(log2(mx)), normal range
    if (e < -100) e = -100;
    return __uint_as_float((uint32_t)(13 - e + 127) << 23);
}
__device__ __forceinline__ float finite_abs(float v) {       // |v|, or 0 for inf / nan
    float a = fabsf(v);
    return a <= 3.0e38f ? a : 0.f;
}
// X * |X|^expo for a complex value (expo = c - 1 or 1/c - 1); 0 stays 0
// The two exponents SGMSE's compression_factor = 0.5 produces (c - 1 = -1/2 forward,
// 1/c - 1 = 1 inverse) need two MUFU.RSQ instead of powf: sqrt(m2) = m2 * rsqrt(m2),
// m2^-1/4 = rsqrt(sqrt(m2)).  The general exponent stays out of line so that the unrolled
// epilogue / loader loops do not carry 16 inlined copies of powf (instruction-cache misses
// made the compressed transforms 2 - 2.6x slower than the plain ones).
static __device__ __noinline__ float pow_general(float m2, float half_expo) { return powf(m2, half_expo); }
__device__ __forceinline__ float pow_half_expo(float m2, float expo) {
    if (expo == -0.5f) return rsqrtf(m2 * rsqrtf(m2));
    if (expo == 1.f) return m2 * rsqrtf(m2);
    return pow_general(m2, 0.5f * expo);
}
__device__ __forceinline__ void compress(float& re, float& im, float expo) {
    const float m2 = re * re + im * im;
    const float g = m2 > 0.f ? pow_half_expo(m2, expo) : 0.f;
    re *= g;
    im *= g;
}
__device__ __forceinline__ float compress_real(float v, float expo) {
    return v != 0.f ? v * pow_half_expo(v * v, expo) : 0.f;
}


__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, SWIZZLE_64B operand tile (64-byte rows, 8-row groups 512 bytes apart)
// with `n` the UMMA N extent: instruction descriptor for kind::f16, fp32 accumulate.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
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

}  // namespace brv_ptx
