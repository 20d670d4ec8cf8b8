// Shared device/host helpers for the b200wave kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b200wave.h"

namespace b200w {

constexpr int kMaxTaps = B200W_MAX_TAPS;
constexpr int kThreads = 256;

// Filter taps travel by value in the kernel parameter block (constant bank): with the tap loops
// fully unrolled every FFMA takes its coefficient as a c[0x0][..] operand, no load instruction.
constexpr int kMaxTemplTaps = 16;   // tap counts the templated kernels are instantiated for
struct Taps {
    float w_lo[kMaxTaps];  // along W
    float w_hi[kMaxTaps];
    float h_lo[kMaxTaps];  // along H
    float h_hi[kMaxTaps];
    // the first kMaxTemplTaps H taps once more as (t, t) pairs: operands of the packed FFMA2 (fma.rn.f32x2), which
    // takes them straight from uniform registers
    float2 h_lo2[kMaxTemplTaps];
    float2 h_hi2[kMaxTemplTaps];
};

// Blackwell packed fp32 FMA: (a.x*b.x + c.x, a.y*b.y + c.y) in one issue slot (SASS FFMA2)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

// Extension index maps of mypad (pw/dwt/lowlevel.py:28-88) and of the periodization branch of
// afb1d (:134-150), folded into one function: returns the source index in [0,n) or -1 for "zero".
__host__ __device__ __forceinline__ int ext_index(int s, int n, int mode) {
    if ((unsigned)s < (unsigned)n) return s;
    switch (mode) {
        case B200W_MODE_SYMMETRIC: {  // half-sample symmetric, period 2n (utils.reflect(-0.5, n-0.5))
            int p = 2 * n;
            int m = s % p;
            if (m < 0) m += p;
            return m < n ? m : p - 1 - m;
        }
        case B200W_MODE_REFLECT: {  // whole-sample, period 2n-2 (F.pad reflect)
            if (n == 1) return 0;
            int p = 2 * n - 2;
            int m = s % p;
            if (m < 0) m += p;
            return m < n ? m : p - m;
        }
        case B200W_MODE_PERIODIC: {
            int m = s % n;
            if (m < 0) m += n;
            return m;
        }
        case B200W_MODE_PERIODIZATION: {  // odd n: last sample repeated once, then period n+1
            int p = n + (n & 1);
            int m = s % p;
            if (m < 0) m += p;
            return m < n ? m : n - 1;
        }
        default:
            return -1;  // zero
    }
}

// index of a coefficient for the synthesis bank: outside [0,m) is zero, except for periodization
// where the coefficient sequence is m-periodic (the wrap-add + roll of sfb1d, :252-261)
__device__ __forceinline__ int coef_index(int k, int m, bool periodic) {
    if ((unsigned)k < (unsigned)m) return k;
    if (!periodic) return -1;
    int r = k % m;
    if (r < 0) r += m;
    return r;
}

// Out-of-line variants for the stream kernels: the maps are only needed for the few rows / columns outside the
// image, and every inlined copy costs ~100 instructions (four integer divisions) of instruction-cache footprint.
static __device__ __noinline__ int ext_index_far(int s, int n, int mode) { return ext_index(s, n, mode); }
static __device__ __noinline__ int coef_index_far(int k, int m, int periodic) { return coef_index(k, m, periodic != 0); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// cp.async (LDGSTS): global -> shared without register staging.  `dst` is a 32-bit shared-window address.
template <int V>
__device__ __forceinline__ void cp_async(unsigned dst, const float* src) {
    if (V == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else if (V == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// src-size 0: nothing is read, the destination is zero-filled (`src` only has to be a valid address)
template <int V>
__device__ __forceinline__ void cp_async_zero(unsigned dst, const float* src) {
    const int z = 0;
    if (V == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(z) : "memory");
    else if (V == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(z) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(z) : "memory");
}
__device__ __forceinline__ void cp_async4_if(unsigned dst, const float* src, bool valid) {
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor in the stream
// is still running; it must call pdl_wait() before it touches global memory (reads of the predecessor's results, and
// writes -- the allocator may have handed it memory the predecessor still reads).  pdl_trigger() in the predecessor
// lets the dependent grid be scheduled as soon as every CTA has issued it (or exited).  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
bool pdl_enabled();   // B200W_PDL=0 switches the attribute off

template <typename Kernel, typename Params>
inline cudaError_t launch_pdl(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t st, const Params& prm) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, prm);
}

int set_last_cuda_error(cudaError_t e);
// diagnostics: counts the launch and remembers the kernel's name (b200w_kernel_launches / b200w_kernel_log)
void note_launch(const char* kernel);

}  // namespace b200w
