// common.cuh -- small device helpers shared by the sm_100a kernels:
// mbarrier + 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP), Philox4x32-10,
// double-double arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Host emulation (tests/host_emul, nvcc -DSB_HOST_EMUL): the arithmetic helpers and the kernels' per-thread /
// per-warp functions are also compiled as plain host functions so the CPU test tier can run the same source
// against the oracle.  In a device build SB_DEV is exactly `__device__ __forceinline__`.
#ifdef SB_HOST_EMUL
#define SB_DEV inline
#else
#define SB_DEV __device__ __forceinline__
#endif

namespace sb {

#ifndef SB_HOST_EMUL   // device-only: mbarrier / TMA bulk copies, Philox (the oracle has the host Philox)
// ---------------------------------------------------------------- mbarrier / TMA bulk
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make the barrier init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// 1-D bulk tensor-memory-accelerator copy global -> shared; dst, src and
// bytes must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------- Philox4x32-10
// Same function as oracle/scoary_oracle.c so_philox4x32_10 (Salmon et al. 2011).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#endif  // !SB_HOST_EMUL

// ---------------------------------------------------------------- double-double
struct dd {
    double hi, lo;
};

SB_DEV dd dd_make(double2 v) { return dd{v.x, v.y}; }

SB_DEV dd dd_add(dd x, dd y)
{
    double s = x.hi + y.hi;
    double bb = s - x.hi;
    double e = (x.hi - (s - bb)) + (y.hi - bb);
    e += x.lo + y.lo;
    double hi = s + e;
    double lo = e - (hi - s);
    return dd{hi, lo};
}

SB_DEV dd dd_neg(dd x) { return dd{-x.hi, -x.lo}; }
SB_DEV dd dd_sub(dd x, dd y) { return dd_add(x, dd_neg(y)); }

SB_DEV uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

}  // namespace sb
