// 1-D analysis / synthesis filter banks for sm_100a.
//
// Replace the bodies of AFB1D.forward / SFB1D.forward (pw/dwt/lowlevel.py:389-405, 719-730: afb1d / sfb1d, :91-172 and
// :226-271, on an (N, C, 1, L) view -- gather padding + F.conv2d with stride (1,2) + two strided .contiguous() copies /
// two F.conv_transpose2d + add) and, with the other bank's taps, each other's hand-written backward (:407-424,
// :732-743).  The 2-D kernels of this library always decimate both axes, so the 1-D transform has its own pair.
// Both are one pass over the data, HBM-bound: a thread owns one output position of one signal, reads its L-tap window
// through the read-only path (adjacent threads read windows that overlap by L - 2 samples: the loads coalesce and hit
// L1), and writes lo / hi (analysis) or y (synthesis) with coalesced 32-bit stores.  Windows that leave the signal take
// the same index maps as the 2-D kernels (ext_index / coef_index, common.cuh): zero, symmetric, reflect, periodic,
// periodization.  Algorithmic bytes: 8 B per input sample (4 read + 4 written) for either direction.
#include "common.cuh"

namespace b200w {

struct Dwt1dParams {
    const float* a;        // analysis: x ; synthesis: lo
    const float* b;        // synthesis: hi (or null = zeros)
    float* o0;             // analysis: lo ; synthesis: y
    float* o1;             // analysis: hi
    long long a_rs;        // row stride of `a` in floats (views: the 'unpad' crop of DWT1DInverse is free)
    int rows, n, m, off, mode, L, periodic;   // n = signal length (analysis in / synthesis out), m = coefficients
    float t0[kMaxTaps], t1[kMaxTaps];
};

__global__ void __launch_bounds__(kThreads) afb1d_kernel(const __grid_constant__ Dwt1dParams p) {
    // rows on grid.y, positions on grid.x: no 64-bit division per output
    for (int row = blockIdx.y; row < p.rows; row += gridDim.y)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p.m; k += gridDim.x * blockDim.x) {
        const size_t idx = (size_t)row * p.m + k;
        const float* __restrict__ xr = p.a + (long long)row * p.a_rs;
        const int s0 = 2 * k - p.off;
        float lo = 0.f, hi = 0.f;
        if (s0 >= 0 && s0 + p.L <= p.n) {            // interior window: no index maps
#pragma unroll 4
            for (int j = 0; j < p.L; ++j) {
                const float v = __ldg(xr + s0 + j);
                lo = fmaf(p.t0[j], v, lo);
                hi = fmaf(p.t1[j], v, hi);
            }
        } else {
            for (int j = 0; j < p.L; ++j) {
                const int s = ext_index(s0 + j, p.n, p.mode);
                if (s < 0) continue;
                const float v = __ldg(xr + s);
                lo = fmaf(p.t0[j], v, lo);
                hi = fmaf(p.t1[j], v, hi);
            }
        }
        p.o0[idx] = lo;
        p.o1[idx] = hi;
    }
}

// "A-space" as in the 2-D synthesis kernels: a = i + off, y[i] = sum_{t = a&1, a&1+2, ..} c[(a - t) / 2] g[t].
// The even output a = 2q and the odd output a = 2q + 1 read the same coefficients c[q - u], u = 0 .. L/2 (even taps /
// odd taps), so a thread owns the pair q: one set of loads, two outputs.
__global__ void __launch_bounds__(kThreads) sfb1d_kernel(const __grid_constant__ Dwt1dParams p) {
    const int q0 = p.off >> 1;                               // pair of output i = 0 (a = off)
    const int npairs = ((p.n - 1 + p.off) >> 1) - q0 + 1;    // pairs covering outputs 0 .. n-1
    const int H2 = (p.L + 1) / 2;
    for (int row = blockIdx.y; row < p.rows; row += gridDim.y)
    for (int pq = blockIdx.x * blockDim.x + threadIdx.x; pq < npairs; pq += gridDim.x * blockDim.x) {
        const int q = q0 + pq;
        const float* __restrict__ lo = p.a + (long long)row * p.a_rs;
        const float* __restrict__ hi = p.b ? p.b + (size_t)row * p.m : nullptr;
        float y0 = 0.f, y1 = 0.f;
        if (q - (H2 - 1) >= 0 && q < p.m) {                  // interior: every coefficient exists
            for (int u = 0; u < H2; ++u) {
                const float cl = __ldg(lo + q - u);
                const float g0e = p.t0[2 * u], g0o = 2 * u + 1 < p.L ? p.t0[2 * u + 1] : 0.f;
                y0 = fmaf(cl, g0e, y0);
                y1 = fmaf(cl, g0o, y1);
                if (hi) {
                    const float ch = __ldg(hi + q - u);
                    const float g1e = p.t1[2 * u], g1o = 2 * u + 1 < p.L ? p.t1[2 * u + 1] : 0.f;
                    y0 = fmaf(ch, g1e, y0);
                    y1 = fmaf(ch, g1o, y1);
                }
            }
        } else {
            for (int u = 0; u < H2; ++u) {
                const int k = coef_index(q - u, p.m, p.periodic != 0);
                if (k < 0) continue;
                const float cl = __ldg(lo + k);
                const float ch = hi ? __ldg(hi + k) : 0.f;
                y0 = fmaf(cl, p.t0[2 * u], y0);
                y0 = fmaf(ch, p.t1[2 * u], y0);
                if (2 * u + 1 < p.L) {
                    y1 = fmaf(cl, p.t0[2 * u + 1], y1);
                    y1 = fmaf(ch, p.t1[2 * u + 1], y1);
                }
            }
        }
        const int i0 = 2 * q - p.off;
        float* yr = p.o0 + (size_t)row * p.n;
        if (i0 >= 0 && i0 < p.n) yr[i0] = y0;
        if (i0 + 1 >= 0 && i0 + 1 < p.n) yr[i0 + 1] = y1;
    }
}

// grid.x covers the positions of one signal, grid.y the signals (the kernels loop over what the limits cut off): many
// small CTAs balance better here than one resident wave of looping ones (measured: 136 vs 213 us at 64 x 2^20, db3)
static dim3 grid_1d(int positions, int rows) {
    unsigned gx = (unsigned)((positions + kThreads - 1) / kThreads);
    if (gx < 1) gx = 1;
    unsigned gy = (unsigned)rows;
    if (gy > 65535) gy = 65535;
    return dim3(gx, gy, 1);
}

static bool mode_ok(int mode) {
    return mode == B200W_MODE_ZERO || mode == B200W_MODE_SYMMETRIC || mode == B200W_MODE_PERIODIZATION ||
           mode == B200W_MODE_REFLECT || mode == B200W_MODE_PERIODIC;
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_afb1d_f32(const float* x, int64_t x_rs, int rows, int n, const float* h0, const float* h1, int L,
                               int mode, float* lo, float* hi, void* stream) {
    if (!x || !h0 || !h1 || !lo || !hi) return B200W_ERR_NULL_POINTER;
    if (!mode_ok(mode)) return B200W_ERR_BAD_MODE;
    if (rows < 1 || n < 1) return B200W_ERR_BAD_SHAPE;
    if (L < 1 || L > kMaxTaps) return B200W_ERR_BAD_TAPS;
    Dwt1dParams p = {};
    const bool per = mode == B200W_MODE_PERIODIZATION;
    if (per) {
        if (n + (n & 1) < L) return B200W_ERR_PER_TOO_SHORT;
        p.m = (n + 1) / 2;
        p.off = L - 1 - L / 2;
    } else {
        p.m = (n + L - 1) / 2;
        const int pad = 2 * (p.m - 1) - n + L;
        if (mode == B200W_MODE_REFLECT && pad > 0 && (pad + 1) / 2 >= n) return B200W_ERR_REFLECT_PAD;
        p.off = pad / 2;
    }
    if (p.m < 1) return B200W_ERR_BAD_SHAPE;
    p.a = x;
    p.a_rs = x_rs;
    p.o0 = lo;
    p.o1 = hi;
    p.rows = rows;
    p.n = n;
    p.mode = mode;
    p.L = L;
    p.periodic = per;
    for (int j = 0; j < L; ++j) { p.t0[j] = h0[j]; p.t1[j] = h1[j]; }
    afb1d_kernel<<<grid_1d(p.m, rows), kThreads, 0, (cudaStream_t)stream>>>(p);
    note_launch("afb1d_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

extern "C" int b200w_sfb1d_f32(const float* lo, int64_t lo_rs, const float* hi, int rows, int m, const float* g0,
                               const float* g1, int L, int mode, int out_len, float* y, void* stream) {
    if (!lo || !g0 || !g1 || !y) return B200W_ERR_NULL_POINTER;
    if (!mode_ok(mode)) return B200W_ERR_BAD_MODE;
    if (rows < 1 || m < 1 || out_len < 1) return B200W_ERR_BAD_SHAPE;
    if (L < 1 || L > kMaxTaps) return B200W_ERR_BAD_TAPS;
    const bool per = mode == B200W_MODE_PERIODIZATION;
    if (per && 2 * m < L) return B200W_ERR_PER_TOO_SHORT;
    if (out_len > (per ? 2 * m : 2 * m - L + 2)) return B200W_ERR_BAD_SHAPE;
    Dwt1dParams p = {};
    p.a = lo;
    p.a_rs = lo_rs;
    p.b = hi;
    p.o0 = y;
    p.rows = rows;
    p.n = out_len;
    p.m = m;
    p.off = per ? L / 2 - 1 : L - 2;
    p.mode = mode;
    p.L = L;
    p.periodic = per;
    for (int j = 0; j < L; ++j) { p.t0[j] = g0[j]; p.t1[j] = g1[j]; }
    sfb1d_kernel<<<grid_1d(out_len / 2 + 2, rows), kThreads, 0, (cudaStream_t)stream>>>(p);
    note_launch("sfb1d_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}
