// Double-precision single-level analysis / synthesis banks (the reference's fp64 mode: modules constructed under
// torch.set_default_dtype(torch.float64), pw tests/test_dwt.py:132-160 `test_equal_double`).
//
// Same arithmetic as AFB2D.forward / SFB2D.forward (pw/dwt/lowlevel.py:336-347, 671-680) and, with the other bank's
// taps, their backward passes -- one thread per output position, any tap count <= B200W_MAX_TAPS, every padding mode,
// the same index maps as the fp32 kernels (ext_index / coef_index, common.cuh).  Precision, not speed, is the point of
// this path: FP64 throughput on B200 is 1/64 of FP32, so the kernels are simple and the multi-level transforms run
// level by level through them.  The fp32 kernels are untouched.
#include "common.cuh"

namespace b200w {

struct Dwt64Params {
    const double* a;       // analysis: x ; synthesis: low
    const double* b;       // synthesis: highs (planes, 3, h, w) or null
    double* o0;            // analysis: low ; synthesis: y
    double* o1;            // analysis: highs (planes, 3, Ho, Wo)
    long long a_ps, a_rs;  // plane / row stride of `a` in elements
    int planes, H, W, Ho, Wo, offH, offW, mode, Lw, Lh, periodic;
    double w_lo[kMaxTaps], w_hi[kMaxTaps], h_lo[kMaxTaps], h_hi[kMaxTaps];
};

__global__ void __launch_bounds__(kThreads) afb2d_f64_kernel(const __grid_constant__ Dwt64Params p) {
    const unsigned band = (unsigned)p.Ho * (unsigned)p.Wo;
    for (int plane = blockIdx.y; plane < p.planes; plane += gridDim.y)
    for (unsigned px = blockIdx.x * blockDim.x + threadIdx.x; px < band; px += gridDim.x * blockDim.x) {
        const int i = (int)(px / (unsigned)p.Wo), k = (int)(px - (unsigned)i * (unsigned)p.Wo);
        const double* __restrict__ xp = p.a + (long long)plane * p.a_ps;
        double ll = 0.0, lh = 0.0, hl = 0.0, hh = 0.0;
        for (int jh = 0; jh < p.Lh; ++jh) {
            const int sr = ext_index(2 * i + jh - p.offH, p.H, p.mode);
            if (sr < 0) continue;
            double lo = 0.0, hi = 0.0;
            for (int jw = 0; jw < p.Lw; ++jw) {
                const int sc = ext_index(2 * k + jw - p.offW, p.W, p.mode);
                if (sc < 0) continue;
                const double v = xp[(long long)sr * p.a_rs + sc];
                lo = fma(p.w_lo[jw], v, lo);
                hi = fma(p.w_hi[jw], v, hi);
            }
            ll = fma(p.h_lo[jh], lo, ll);
            lh = fma(p.h_hi[jh], lo, lh);
            hl = fma(p.h_lo[jh], hi, hl);
            hh = fma(p.h_hi[jh], hi, hh);
        }
        p.o0[(size_t)plane * band + px] = ll;
        double* hp = p.o1 + (size_t)plane * 3 * band + px;
        hp[0] = lh;
        hp[band] = hl;
        hp[2 * (size_t)band] = hh;
    }
}

// "A-space" as in the fp32 synthesis kernels: a = n + off; output a uses taps of parity a & 1
__global__ void __launch_bounds__(kThreads) sfb2d_f64_kernel(const __grid_constant__ Dwt64Params p) {
    const unsigned outpx = (unsigned)p.Ho * (unsigned)p.Wo;   // here Ho x Wo = out_h x out_w, H x W = h x w
    const size_t band = (size_t)p.H * p.W;
    for (int plane = blockIdx.y; plane < p.planes; plane += gridDim.y)
    for (unsigned px = blockIdx.x * blockDim.x + threadIdx.x; px < outpx; px += gridDim.x * blockDim.x) {
        const int nH = (int)(px / (unsigned)p.Wo), nW = (int)(px - (unsigned)nH * (unsigned)p.Wo);
        const double* __restrict__ lowp = p.a + (long long)plane * p.a_ps;
        const double* __restrict__ hip = p.b ? p.b + (size_t)plane * 3 * band : nullptr;
        const int AH = nH + p.offH, AW = nW + p.offW;
        double y = 0.0;
        for (int tH = AH & 1; tH < p.Lh; tH += 2) {
            const int kr = coef_index((AH - tH) / 2, p.H, p.periodic != 0);
            if (kr < 0) continue;
            double lo = 0.0, hi = 0.0;
            for (int tW = AW & 1; tW < p.Lw; tW += 2) {
                const int kc = coef_index((AW - tW) / 2, p.W, p.periodic != 0);
                if (kc < 0) continue;
                lo = fma(lowp[(long long)kr * p.a_rs + kc], p.w_lo[tW], lo);
                if (hip) {
                    const double* q = hip + (size_t)kr * p.W + kc;
                    hi = fma(q[0], p.w_lo[tW], hi);            // LH
                    lo = fma(q[band], p.w_hi[tW], lo);         // HL
                    hi = fma(q[2 * band], p.w_hi[tW], hi);     // HH
                }
            }
            y = fma(lo, p.h_lo[tH], y);
            y = fma(hi, p.h_hi[tH], y);
        }
        p.o0[(size_t)plane * outpx + px] = y;
    }
}

static dim3 grid64(int planes, size_t px) {
    unsigned gx = (unsigned)((px + kThreads - 1) / kThreads);
    unsigned gy = (unsigned)planes;
    if (gy > 65535) gy = 65535;
    return dim3(gx < 1 ? 1 : gx, gy, 1);
}

static bool mode_ok64(int mode) {
    return mode == B200W_MODE_ZERO || mode == B200W_MODE_SYMMETRIC || mode == B200W_MODE_PERIODIZATION ||
           mode == B200W_MODE_REFLECT || mode == B200W_MODE_PERIODIC;
}

// output length and left extension of the analysis bank along one axis (pw/dwt/lowlevel.py:134-168)
static int analysis_axis(int n, int l, int mode, int* m, int* off) {
    if (mode == B200W_MODE_PERIODIZATION) {
        if (n + (n & 1) < l) return B200W_ERR_PER_TOO_SHORT;
        *m = (n + 1) / 2;
        *off = l - 1 - l / 2;
        return B200W_OK;
    }
    *m = (n + l - 1) / 2;
    const int pad = 2 * (*m - 1) - n + l;
    if (mode == B200W_MODE_REFLECT && pad > 0 && (pad + 1) / 2 >= n) return B200W_ERR_REFLECT_PAD;
    *off = pad / 2;
    return B200W_OK;
}

static int fill64(Dwt64Params& p, const double* w_lo, const double* w_hi, int Lw, const double* h_lo, const double* h_hi,
                  int Lh) {
    if (!w_lo || !w_hi || !h_lo || !h_hi) return B200W_ERR_NULL_POINTER;
    if (Lw < 1 || Lw > kMaxTaps || Lh < 1 || Lh > kMaxTaps) return B200W_ERR_BAD_TAPS;
    for (int j = 0; j < Lw; ++j) { p.w_lo[j] = w_lo[j]; p.w_hi[j] = w_hi[j]; }
    for (int j = 0; j < Lh; ++j) { p.h_lo[j] = h_lo[j]; p.h_hi[j] = h_hi[j]; }
    p.Lw = Lw;
    p.Lh = Lh;
    return B200W_OK;
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_afb2d_f64(const double* x, int64_t x_ps, int64_t x_rs, int planes, int H, int W, const double* w_lo,
                               const double* w_hi, int Lw, const double* h_lo, const double* h_hi, int Lh, int mode,
                               double* low, double* highs, void* stream) {
    if (!x || !low || !highs) return B200W_ERR_NULL_POINTER;
    if (!mode_ok64(mode)) return B200W_ERR_BAD_MODE;
    if (planes < 1 || H < 1 || W < 1) return B200W_ERR_BAD_SHAPE;
    Dwt64Params p = {};
    int rc = fill64(p, w_lo, w_hi, Lw, h_lo, h_hi, Lh);
    if (rc) return rc;
    if ((rc = analysis_axis(H, Lh, mode, &p.Ho, &p.offH))) return rc;
    if ((rc = analysis_axis(W, Lw, mode, &p.Wo, &p.offW))) return rc;
    if (p.Ho < 1 || p.Wo < 1 || (long long)p.Ho * p.Wo > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    p.a = x;
    p.a_ps = x_ps;
    p.a_rs = x_rs;
    p.o0 = low;
    p.o1 = highs;
    p.planes = planes;
    p.H = H;
    p.W = W;
    p.mode = mode;
    afb2d_f64_kernel<<<grid64(planes, (size_t)p.Ho * p.Wo), kThreads, 0, (cudaStream_t)stream>>>(p);
    note_launch("afb2d_f64_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

extern "C" int b200w_sfb2d_f64(const double* low, int64_t low_ps, int64_t low_rs, const double* highs, int planes, int h,
                               int w, const double* w_lo, const double* w_hi, int Lw, const double* h_lo,
                               const double* h_hi, int Lh, int mode, double* y, int out_h, int out_w, void* stream) {
    if (!low || !y) return B200W_ERR_NULL_POINTER;
    if (!mode_ok64(mode)) return B200W_ERR_BAD_MODE;
    if (planes < 1 || h < 1 || w < 1 || out_h < 1 || out_w < 1) return B200W_ERR_BAD_SHAPE;
    Dwt64Params p = {};
    const int rc = fill64(p, w_lo, w_hi, Lw, h_lo, h_hi, Lh);
    if (rc) return rc;
    const bool per = mode == B200W_MODE_PERIODIZATION;
    if (per && (2 * h < Lh || 2 * w < Lw)) return B200W_ERR_PER_TOO_SHORT;
    if (out_h > (per ? 2 * h : 2 * h - Lh + 2) || out_w > (per ? 2 * w : 2 * w - Lw + 2)) return B200W_ERR_BAD_SHAPE;
    if ((long long)out_h * out_w > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    p.a = low;
    p.a_ps = low_ps;
    p.a_rs = low_rs;
    p.b = highs;
    p.o0 = y;
    p.planes = planes;
    p.H = h;
    p.W = w;
    p.Ho = out_h;
    p.Wo = out_w;
    p.offH = per ? Lh / 2 - 1 : Lh - 2;
    p.offW = per ? Lw / 2 - 1 : Lw - 2;
    p.mode = mode;
    p.periodic = per;
    sfb2d_f64_kernel<<<grid64(planes, (size_t)out_h * out_w), kThreads, 0, (cudaStream_t)stream>>>(p);
    note_launch("sfb2d_f64_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}
