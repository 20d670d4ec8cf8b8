// TMA (cp.async.bulk.tensor) + mbarrier helpers for the sm_100a kernels.
//
// Host: tensor maps are encoded per call with cuTensorMapEncodeTiled (looked up through the runtime, so the library
// does not link libcuda) and travel to the kernel inside its __grid_constant__ parameter block.
// Device: one elected thread issues a tile copy into an mbarrier-guarded shared-memory stage; coordinates outside the
// tensor are zero-filled by the copy engine (= the 'zero' padding mode, and every image border, for free).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef __CUDACC__
#include "common.cuh"   // B200W_CHK / B200W_CHK_S (no-ops unless -DB200W_BOUNDS)
#endif

namespace b200w {

// ---- host ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tma_encode_fn();   // null when the driver does not export it (api.cu)

// fp32 tensor of rank 1..4, dims / box in elements (dim 0 = the contiguous one), strides in BYTES for dims 1.. (each a
// multiple of 16).  Returns false when the driver rejects the map (the caller then takes another kernel).
bool tma_make_map_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

// ---- device --------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy ones (a stage that
// was read / patched with ordinary loads and stores is handed back to the copy engine)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(unsigned bar, unsigned count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (no hardware suspend): for polling loops that serve several barriers
__device__ __forceinline__ bool mbar_test_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocks until the phase with the given parity has completed.  The retry count is bounded: a protocol error traps
// (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}

__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0) : "memory");
}
// plain bulk copy (no tensor map): `bytes` contiguous bytes, source / destination / size multiples of 16 (UBLKCP)
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    B200W_CHK_S(dst, 16);
    B200W_CHK_S(dst + bytes - 16, 16);
    B200W_CHK(src, 16);
    B200W_CHK(reinterpret_cast<const char*>(src) + bytes - 16, 16);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
#endif

}  // namespace b200w
