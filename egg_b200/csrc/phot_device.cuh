// phot_device.cuh — device-only helpers shared by the two flux kernels (phot_kernels.cu, phot_lane_kernels.cu):
// mbarrier / TMA bulk-copy wrappers and the log2-table lookups of the SED grids.
#ifndef EGGPHOT_PHOT_DEVICE_CUH
#define EGGPHOT_PHOT_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "phot_tables.hpp"

namespace eggphot {
namespace {

constexpr unsigned kFull = 0xffffffffu;

// ---- PTX wrappers (mbarrier + TMA bulk copy) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- lookups in a wavelength grid ---------------------------------------------------------------------
// last j in [0, n) with x[j] <= v (-1 if none), started from the log2 lookup of the grid: the table
// entry is at most a couple of nodes away, and the two scans make the answer exact in double.
__device__ __forceinline__ float lut_pos(double v, int e0) { return (__log2f((float)v) - (float)e0) * (float)kLutPerOctave; }
__device__ __forceinline__ int lut_last_le(const double *__restrict__ x, int n, const uint16_t *__restrict__ lut,
                                           int nlut, float kpos, double v) {
    int k = (int)kpos;
    k = max(0, min(k, nlut - 1));
    int j = __ldg(lut + k);
    // both neighbours of the guess are requested at once: the guess is normally right or off by one
    double a = __ldg(x + max(j, 0)), b = __ldg(x + min(j + 1, n - 1));
    while (j >= 0 && a > v) { --j; b = a; a = __ldg(x + max(j, 0)); }
    while (j + 1 < n && b <= v) { ++j; b = __ldg(x + min(j + 1, n - 1)); }
    return j;
}
__device__ __forceinline__ int lut_first_ge(const double *__restrict__ x, int n, const uint16_t *__restrict__ lut,
                                            int nlut, float kpos, double v) {
    int k = (int)kpos;
    k = max(0, min(k, nlut - 1));
    int j = __ldg(lut + k); // candidate for the last node < v
    double a = __ldg(x + max(j, 0)), b = __ldg(x + min(j + 1, n - 1));
    while (j >= 0 && a >= v) { --j; b = a; a = __ldg(x + max(j, 0)); }
    while (j + 1 < n && b < v) { ++j; b = __ldg(x + min(j + 1, n - 1)); }
    return j + 1;
}

} // namespace
} // namespace eggphot
#endif
