// 2-D DWT analysis / synthesis levels for sm_100a.
//
// One fused kernel per level (BASELINE.json north_star): the analysis kernel stages an input tile
// with its halo in shared memory (the padding mode is applied as an index map while staging), runs
// the row (W) pass with stride-2 decimation into shared memory, then the column (H) pass, and writes
// LL into `low` and LH/HL/HH straight into `highs[:, :, 0..2]` -- replacing the reference's
// pad-gather + 2x F.conv2d + reshape + 2x .contiguous() (pw/dwt/lowlevel.py:336-347).
// The synthesis kernel fuses upsample + filter + accumulate of all four sub-bands in polyphase form,
// replacing 6x F.conv_transpose2d + 3 adds (pw/dwt/lowlevel.py:671-680).
//
// Closed forms (validated against the reference to 1e-15 by oracle/dwt_oracle.py):
//   analysis   y_c[k] = sum_j w_c[j] * x_ext[2k + j - off],  off = p//2 with p = 2(M-1) - N + L,
//              periodization: off = L-1 - L//2 on the (even-extended) N'-periodic signal
//   synthesis  y[n]   = sum_k lo[k] g0[t] + hi[k] g1[t],  t = n + off - 2k in [0,L),
//              off = L-2 (coefficients outside [0,M) are zero), periodization: off = L//2 - 1 and the
//              coefficient sequence is M-periodic
#include "common.cuh"

namespace b200w {

struct AfbParams {
    const float* x;
    float* low;
    float* highs;
    long long x_ps, x_rs;
    int planes, H, W, Ho, Wo;
    int mode, offW, offH, Lw, Lh;
    int tiles_w, tiles_h;
    Taps t;
};

struct SfbParams {
    const float* low;
    const float* highs;  // may be null (= zeros)
    float* y;
    long long low_ps, low_rs;
    int planes, h, w, out_h, out_w;
    int periodic, offW, offH, Lw, Lh;
    int a0W, a0H;  // first A-space coordinate (even) covered by tile 0
    int tiles_w, tiles_h;
    Taps t;
};

// ------------------------------------------------------------------------------------------------
// analysis, tiled.  Output tile TH x TW per CTA (x4 sub-bands), 256 threads.
// ------------------------------------------------------------------------------------------------
template <int L, int TW, int TH>
struct AfbCfg {
    static constexpr int PC = 2 * TW + L - 2;  // staged patch columns (even)
    static constexpr int PR = 2 * TH + L - 2;  // staged patch rows
    static constexpr int NS = kThreads / TW;   // row strips in the column pass
    static constexpr int RS = TH / NS;         // output rows per thread in the column pass
    static constexpr size_t smem = sizeof(float) * (size_t)(PR * PC + 2 * PR * TW);
    static_assert(L % 2 == 0 && TW % 32 == 0 && kThreads % TW == 0 && TH % NS == 0, "bad tile");
};

template <int L, int TW, int TH>
__global__ void __launch_bounds__(kThreads) afb2d_tile_kernel(const __grid_constant__ AfbParams p) {
    using Cfg = AfbCfg<L, TW, TH>;
    constexpr int PC = Cfg::PC, PR = Cfg::PR, NS = Cfg::NS, RS = Cfg::RS;
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                 // [PR][PC]
    float* mid_lo = smem + PR * PC;      // [PR][TW]
    float* mid_hi = mid_lo + PR * TW;    // [PR][TW]

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tw = bid % p.tiles_w;
    bid /= p.tiles_w;
    const int th = bid % p.tiles_h;
    const int plane = bid / p.tiles_h;

    const int r0 = 2 * th * TH - p.offH;  // source row of patch row 0
    const int c0 = 2 * tw * TW - p.offW;
    const float* __restrict__ xp = p.x + (long long)plane * p.x_ps;

    // ---- stage the tile + halo; the padding mode is an index map
    const bool interior = r0 >= 0 && r0 + PR <= p.H && c0 >= 0 && c0 + PC <= p.W;
    if (interior) {
        const float* __restrict__ src = xp + (long long)r0 * p.x_rs + c0;
#pragma unroll 4
        for (int idx = tid; idx < PR * PC; idx += kThreads) {
            const int r = idx / PC, c = idx - r * PC;
            patch[idx] = __ldg(src + (long long)r * p.x_rs + c);
        }
    } else {
#pragma unroll 2
        for (int idx = tid; idx < PR * PC; idx += kThreads) {
            const int r = idx / PC, c = idx - r * PC;
            const int sr = ext_index(r0 + r, p.H, p.mode);
            const int sc = ext_index(c0 + c, p.W, p.mode);
            float v = 0.f;
            if (sr >= 0 && sc >= 0) v = __ldg(xp + (long long)sr * p.x_rs + sc);
            patch[idx] = v;
        }
    }
    __syncthreads();

    // ---- row pass (along W), decimate by 2: lane <-> output column, 64-bit conflict-free LDS
    {
        const int k = tid % TW;
        for (int r = tid / TW; r < PR; r += NS) {
            const float2* src = reinterpret_cast<const float2*>(patch + r * PC + 2 * k);
            float lo = 0.f, hi = 0.f;
#pragma unroll
            for (int j2 = 0; j2 < L / 2; ++j2) {
                const float2 v = src[j2];
                lo = fmaf(p.t.w_lo[2 * j2], v.x, lo);
                hi = fmaf(p.t.w_hi[2 * j2], v.x, hi);
                lo = fmaf(p.t.w_lo[2 * j2 + 1], v.y, lo);
                hi = fmaf(p.t.w_hi[2 * j2 + 1], v.y, hi);
            }
            mid_lo[r * TW + k] = lo;
            mid_hi[r * TW + k] = hi;
        }
    }
    __syncthreads();

    // ---- column pass (along H), decimate by 2; each thread owns RS output rows of one column
    {
        const int k = tid % TW;
        const int s = tid / TW;
        float acc[RS][4];
#pragma unroll
        for (int i = 0; i < RS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const float* plo = mid_lo + (2 * s * RS) * TW + k;
        const float* phi = mid_hi + (2 * s * RS) * TW + k;
#pragma unroll
        for (int rr = 0; rr < 2 * RS + L - 2; ++rr) {
            const float vlo = plo[rr * TW];
            const float vhi = phi[rr * TW];
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                const int j = rr - 2 * i;
                if (j >= 0 && j < L) {
                    acc[i][0] = fmaf(p.t.h_lo[j], vlo, acc[i][0]);  // LL
                    acc[i][1] = fmaf(p.t.h_hi[j], vlo, acc[i][1]);  // LH: W-lo, H-hi
                    acc[i][2] = fmaf(p.t.h_lo[j], vhi, acc[i][2]);  // HL: W-hi, H-lo
                    acc[i][3] = fmaf(p.t.h_hi[j], vhi, acc[i][3]);  // HH
                }
            }
        }
        const int kk = tw * TW + k;
        if (kk < p.Wo) {
            const size_t band = (size_t)p.Ho * p.Wo;
            float* lowp = p.low + (size_t)plane * band + kk;
            float* hip = p.highs + (size_t)plane * 3 * band + kk;
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                const int row = th * TH + s * RS + i;
                if (row < p.Ho) {
                    const size_t o = (size_t)row * p.Wo;
                    lowp[o] = acc[i][0];
                    hip[o] = acc[i][1];
                    hip[band + o] = acc[i][2];
                    hip[2 * band + o] = acc[i][3];
                }
            }
        }
    }
}

// analysis, direct (any tap counts up to kMaxTaps, odd or mixed lengths): one thread per output
// position, no staging.  Fallback and on-device cross-check of the tiled kernel.
__global__ void __launch_bounds__(kThreads) afb2d_direct_kernel(const __grid_constant__ AfbParams p) {
    const size_t band = (size_t)p.Ho * p.Wo;
    const size_t total = band * p.planes;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % p.Wo);
        const int i = (int)((idx / p.Wo) % p.Ho);
        const int plane = (int)(idx / band);
        const float* __restrict__ xp = p.x + (long long)plane * p.x_ps;
        float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
        for (int jh = 0; jh < p.Lh; ++jh) {
            const int sr = ext_index(2 * i + jh - p.offH, p.H, p.mode);
            if (sr < 0) continue;
            float lo = 0.f, hi = 0.f;
            for (int jw = 0; jw < p.Lw; ++jw) {
                const int sc = ext_index(2 * k + jw - p.offW, p.W, p.mode);
                if (sc < 0) continue;
                const float v = __ldg(xp + (long long)sr * p.x_rs + sc);
                lo = fmaf(p.t.w_lo[jw], v, lo);
                hi = fmaf(p.t.w_hi[jw], v, hi);
            }
            ll = fmaf(p.t.h_lo[jh], lo, ll);
            lh = fmaf(p.t.h_hi[jh], lo, lh);
            hl = fmaf(p.t.h_lo[jh], hi, hl);
            hh = fmaf(p.t.h_hi[jh], hi, hh);
        }
        const size_t o = (size_t)i * p.Wo + k;
        p.low[(size_t)plane * band + o] = ll;
        float* hip = p.highs + (size_t)plane * 3 * band + o;
        hip[0] = lh;
        hip[band] = hl;
        hip[2 * band] = hh;
    }
}

// ------------------------------------------------------------------------------------------------
// synthesis, tiled.  Works in "A-space": a = n + off, so that the polyphase split (which taps an
// output uses) depends only on the parity of the tile-local coordinate.  Output tile TH x TW.
// W synthesis first (on the KH coefficient rows), then H synthesis; by separability this equals the
// reference's H-then-W order (pw/dwt/lowlevel.py:677-679) up to fp32 rounding.
// ------------------------------------------------------------------------------------------------
template <int L, int TW, int TH>
struct SfbCfg {
    static constexpr int H2 = L / 2;
    static constexpr int KW = TW / 2 + H2 - 1;  // coefficient columns staged
    static constexpr int KH = TH / 2 + H2 - 1;  // coefficient rows staged
    static constexpr int NS = kThreads / TW;
    static constexpr int RS = TH / NS;  // output rows per thread in the H pass (even)
    static constexpr size_t smem = sizeof(float) * (size_t)(4 * KH * KW + 2 * KH * TW);
    static_assert(L % 2 == 0 && TW % 32 == 0 && kThreads % TW == 0 && TH % NS == 0 && RS % 2 == 0, "bad tile");
};

template <int L, int TW, int TH>
__global__ void __launch_bounds__(kThreads) sfb2d_tile_kernel(const __grid_constant__ SfbParams p) {
    using Cfg = SfbCfg<L, TW, TH>;
    constexpr int H2 = Cfg::H2, KW = Cfg::KW, KH = Cfg::KH, NS = Cfg::NS, RS = Cfg::RS;
    extern __shared__ __align__(16) float smem[];
    float* sub = smem;                   // [4][KH][KW]  LL, LH, HL, HH
    float* u_lo = smem + 4 * KH * KW;    // [KH][TW]  W-synthesised, to be combined with h_lo
    float* u_hi = u_lo + KH * TW;        // [KH][TW]  ... with h_hi

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tw = bid % p.tiles_w;
    bid /= p.tiles_w;
    const int th = bid % p.tiles_h;
    const int plane = bid / p.tiles_h;

    const int aW = p.a0W + tw * TW;  // even
    const int aH = p.a0H + th * TH;  // even
    const int kW0 = aW / 2 - (H2 - 1);
    const int kH0 = aH / 2 - (H2 - 1);
    const size_t band = (size_t)p.h * p.w;
    const float* __restrict__ lowp = p.low + (long long)plane * p.low_ps;
    const float* __restrict__ hip = p.highs ? p.highs + (size_t)plane * 3 * band : nullptr;

    // ---- stage the four sub-band patches
    const bool interior = kH0 >= 0 && kH0 + KH <= p.h && kW0 >= 0 && kW0 + KW <= p.w;
    if (interior) {
#pragma unroll 4
        for (int idx = tid; idx < KH * KW; idx += kThreads) {
            const int r = idx / KW, c = idx - r * KW;
            sub[idx] = __ldg(lowp + (long long)(kH0 + r) * p.low_rs + kW0 + c);
            if (hip) {
                const float* q = hip + (size_t)(kH0 + r) * p.w + kW0 + c;
                sub[KH * KW + idx] = __ldg(q);
                sub[2 * KH * KW + idx] = __ldg(q + band);
                sub[3 * KH * KW + idx] = __ldg(q + 2 * band);
            } else {
                sub[KH * KW + idx] = 0.f;
                sub[2 * KH * KW + idx] = 0.f;
                sub[3 * KH * KW + idx] = 0.f;
            }
        }
    } else {
        for (int idx = tid; idx < KH * KW; idx += kThreads) {
            const int r = idx / KW, c = idx - r * KW;
            const int kr = coef_index(kH0 + r, p.h, p.periodic);
            const int kc = coef_index(kW0 + c, p.w, p.periodic);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            if (kr >= 0 && kc >= 0) {
                v0 = __ldg(lowp + (long long)kr * p.low_rs + kc);
                if (hip) {
                    const float* q = hip + (size_t)kr * p.w + kc;
                    v1 = __ldg(q);
                    v2 = __ldg(q + band);
                    v3 = __ldg(q + 2 * band);
                }
            }
            sub[idx] = v0;
            sub[KH * KW + idx] = v1;
            sub[2 * KH * KW + idx] = v2;
            sub[3 * KH * KW + idx] = v3;
        }
    }
    __syncthreads();

    // ---- W synthesis: each item makes the even/odd output pair (a = 2q, 2q+1) of one coefficient row
    for (int item = tid; item < KH * (TW / 2); item += kThreads) {
        const int q = item % (TW / 2);
        const int r = item / (TW / 2);
        const float* s0 = sub + r * KW + q + H2 - 1;  // LL  (k index decreases with u)
        const float* s1 = s0 + KH * KW;               // LH
        const float* s2 = s1 + KH * KW;               // HL
        const float* s3 = s2 + KH * KW;               // HH
        float lo_e = 0.f, lo_o = 0.f, hi_e = 0.f, hi_o = 0.f;
#pragma unroll
        for (int u = 0; u < H2; ++u) {
            const float ll = s0[-u], lh = s1[-u], hl = s2[-u], hh = s3[-u];
            lo_e = fmaf(ll, p.t.w_lo[2 * u], lo_e);
            lo_e = fmaf(hl, p.t.w_hi[2 * u], lo_e);
            lo_o = fmaf(ll, p.t.w_lo[2 * u + 1], lo_o);
            lo_o = fmaf(hl, p.t.w_hi[2 * u + 1], lo_o);
            hi_e = fmaf(lh, p.t.w_lo[2 * u], hi_e);
            hi_e = fmaf(hh, p.t.w_hi[2 * u], hi_e);
            hi_o = fmaf(lh, p.t.w_lo[2 * u + 1], hi_o);
            hi_o = fmaf(hh, p.t.w_hi[2 * u + 1], hi_o);
        }
        *reinterpret_cast<float2*>(u_lo + r * TW + 2 * q) = make_float2(lo_e, lo_o);
        *reinterpret_cast<float2*>(u_hi + r * TW + 2 * q) = make_float2(hi_e, hi_o);
    }
    __syncthreads();

    // ---- H synthesis: thread = one output column, RS consecutive output rows
    {
        const int c = tid % TW;
        const int s = tid / TW;
        constexpr int NR = RS / 2 + H2 - 1;  // coefficient rows feeding RS outputs
        float vlo[NR], vhi[NR];
        const float* plo = u_lo + (s * (RS / 2)) * TW + c;
        const float* phi = u_hi + (s * (RS / 2)) * TW + c;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            vlo[r] = plo[r * TW];
            vhi[r] = phi[r * TW];
        }
        const int nW = aW + c - p.offW;
        const bool col_ok = nW >= 0 && nW < p.out_w;
        float* yp = p.y + (size_t)plane * p.out_h * p.out_w + nW;
#pragma unroll
        for (int i = 0; i < RS / 2; ++i) {
            float ye = 0.f, yo = 0.f;
#pragma unroll
            for (int u = 0; u < H2; ++u) {
                const int r = i - u + H2 - 1;  // local coefficient row
                ye = fmaf(vlo[r], p.t.h_lo[2 * u], ye);
                ye = fmaf(vhi[r], p.t.h_hi[2 * u], ye);
                yo = fmaf(vlo[r], p.t.h_lo[2 * u + 1], yo);
                yo = fmaf(vhi[r], p.t.h_hi[2 * u + 1], yo);
            }
            const int nH = aH + s * RS + 2 * i - p.offH;
            if (col_ok) {
                if (nH >= 0 && nH < p.out_h) yp[(size_t)nH * p.out_w] = ye;
                if (nH + 1 >= 0 && nH + 1 < p.out_h) yp[(size_t)(nH + 1) * p.out_w] = yo;
            }
        }
    }
}

// synthesis, direct: one thread per output sample; any tap count.
__global__ void __launch_bounds__(kThreads) sfb2d_direct_kernel(const __grid_constant__ SfbParams p) {
    const size_t total = (size_t)p.planes * p.out_h * p.out_w;
    const size_t band = (size_t)p.h * p.w;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int nW = (int)(idx % p.out_w);
        const int nH = (int)((idx / p.out_w) % p.out_h);
        const int plane = (int)(idx / ((size_t)p.out_h * p.out_w));
        const float* __restrict__ lowp = p.low + (long long)plane * p.low_ps;
        const float* __restrict__ hip = p.highs ? p.highs + (size_t)plane * 3 * band : nullptr;
        const int AH = nH + p.offH, AW = nW + p.offW;
        float y = 0.f;
        for (int tH = AH & 1; tH < p.Lh; tH += 2) {
            const int kr = coef_index((AH - tH) / 2, p.h, p.periodic);
            if (kr < 0) continue;
            float lo = 0.f, hi = 0.f;  // W-synthesised values to be combined with h_lo / h_hi
            for (int tW = AW & 1; tW < p.Lw; tW += 2) {
                const int kc = coef_index((AW - tW) / 2, p.w, p.periodic);
                if (kc < 0) continue;
                const float ll = __ldg(lowp + (long long)kr * p.low_rs + kc);
                lo = fmaf(ll, p.t.w_lo[tW], lo);
                if (hip) {
                    const float* q = hip + (size_t)kr * p.w + kc;
                    hi = fmaf(__ldg(q), p.t.w_lo[tW], hi);               // LH
                    lo = fmaf(__ldg(q + band), p.t.w_hi[tW], lo);        // HL
                    hi = fmaf(__ldg(q + 2 * band), p.t.w_hi[tW], hi);    // HH
                }
            }
            y = fmaf(lo, p.t.h_lo[tH], y);
            y = fmaf(hi, p.t.h_hi[tH], y);
        }
        p.y[idx] = y;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool force_direct() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_FORCE_DIRECT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static bool mode_supported(int mode) {
    return mode == B200W_MODE_ZERO || mode == B200W_MODE_SYMMETRIC || mode == B200W_MODE_PERIODIZATION ||
           mode == B200W_MODE_REFLECT || mode == B200W_MODE_PERIODIC;
}

static int fill_taps(Taps& t, const float* w_lo, const float* w_hi, int Lw, const float* h_lo, const float* h_hi,
                     int Lh) {
    if (Lw < 1 || Lw > kMaxTaps || Lh < 1 || Lh > kMaxTaps || !w_lo || !w_hi || !h_lo || !h_hi)
        return B200W_ERR_BAD_TAPS;
    for (int i = 0; i < kMaxTaps; ++i) {
        t.w_lo[i] = i < Lw ? w_lo[i] : 0.f;
        t.w_hi[i] = i < Lw ? w_hi[i] : 0.f;
        t.h_lo[i] = i < Lh ? h_lo[i] : 0.f;
        t.h_hi[i] = i < Lh ? h_hi[i] : 0.f;
    }
    return B200W_OK;
}

static int coeff_len(int n, int l, int mode) { return mode == B200W_MODE_PERIODIZATION ? (n + 1) / 2 : (n + l - 1) / 2; }

// left padding of the analysis bank along one axis; also validates the axis
static int analysis_offset(int n, int l, int mode, int* off) {
    if (mode == B200W_MODE_PERIODIZATION) {
        if (n + (n & 1) < l) return B200W_ERR_PER_TOO_SHORT;
        *off = l - 1 - l / 2;
        return B200W_OK;
    }
    const int m = coeff_len(n, l, mode);
    const int p = 2 * (m - 1) - n + l;
    if (mode == B200W_MODE_REFLECT && p > 0 && (p + 1) / 2 >= n) return B200W_ERR_REFLECT_PAD;
    *off = p / 2;
    return B200W_OK;
}

template <typename K, typename P>
static int launch(K kernel, const P& p, size_t grid, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_last_cuda_error(e);
    }
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <int L>
static int launch_afb_tiled(AfbParams& p, cudaStream_t st) {
    constexpr int TW = 32, TH = 32;
    p.tiles_w = (p.Wo + TW - 1) / TW;
    p.tiles_h = (p.Ho + TH - 1) / TH;
    const size_t grid = (size_t)p.tiles_w * p.tiles_h * p.planes;
    return launch(afb2d_tile_kernel<L, TW, TH>, p, grid, AfbCfg<L, TW, TH>::smem, st);
}

template <int L>
static int launch_sfb_tiled(SfbParams& p, cudaStream_t st) {
    constexpr int TW = 64, TH = 64;
    p.a0W = p.offW & ~1;
    p.a0H = p.offH & ~1;
    p.tiles_w = (p.offW + p.out_w - p.a0W + TW - 1) / TW;
    p.tiles_h = (p.offH + p.out_h - p.a0H + TH - 1) / TH;
    const size_t grid = (size_t)p.tiles_w * p.tiles_h * p.planes;
    return launch(sfb2d_tile_kernel<L, TW, TH>, p, grid, SfbCfg<L, TW, TH>::smem, st);
}

static size_t direct_grid(size_t total) {
    size_t g = (total + kThreads - 1) / kThreads;
    const size_t cap = 148 * 16;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_dwt_coeff_len(int n, int l, int mode) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (n < 1 || l < 1) return B200W_ERR_BAD_SHAPE;
    return coeff_len(n, l, mode);
}

extern "C" int b200w_idwt_len(int m, int l, int mode) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (m < 1 || l < 1) return B200W_ERR_BAD_SHAPE;
    return mode == B200W_MODE_PERIODIZATION ? 2 * m : 2 * m - l + 2;
}

extern "C" int b200w_afb2d_f32(const float* x, int64_t x_plane_stride, int64_t x_row_stride, int planes, int H,
                               int W, const float* w_lo, const float* w_hi, int Lw, const float* h_lo,
                               const float* h_hi, int Lh, int mode, float* low, float* highs, void* stream) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (!x || !low || !highs) return B200W_ERR_NULL_POINTER;
    if (planes < 1 || H < 1 || W < 1) return B200W_ERR_BAD_SHAPE;
    AfbParams p;
    int rc = fill_taps(p.t, w_lo, w_hi, Lw, h_lo, h_hi, Lh);
    if (rc) return rc;
    if ((rc = analysis_offset(W, Lw, mode, &p.offW))) return rc;
    if ((rc = analysis_offset(H, Lh, mode, &p.offH))) return rc;
    p.x = x;
    p.low = low;
    p.highs = highs;
    p.x_ps = x_plane_stride;
    p.x_rs = x_row_stride;
    p.planes = planes;
    p.H = H;
    p.W = W;
    p.Ho = coeff_len(H, Lh, mode);
    p.Wo = coeff_len(W, Lw, mode);
    p.mode = mode;
    p.Lw = Lw;
    p.Lh = Lh;
    p.tiles_w = p.tiles_h = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (Lw == Lh && !force_direct()) {
        switch (Lw) {
            case 2: return launch_afb_tiled<2>(p, st);
            case 4: return launch_afb_tiled<4>(p, st);
            case 6: return launch_afb_tiled<6>(p, st);
            case 8: return launch_afb_tiled<8>(p, st);
            case 10: return launch_afb_tiled<10>(p, st);
            case 12: return launch_afb_tiled<12>(p, st);
            case 14: return launch_afb_tiled<14>(p, st);
            case 16: return launch_afb_tiled<16>(p, st);
            default: break;
        }
    }
    const size_t total = (size_t)planes * p.Ho * p.Wo;
    return launch(afb2d_direct_kernel, p, direct_grid(total), 0, st);
}

extern "C" int b200w_sfb2d_f32(const float* low, int64_t low_plane_stride, int64_t low_row_stride,
                               const float* highs, int planes, int h, int w, const float* w_lo, const float* w_hi,
                               int Lw, const float* h_lo, const float* h_hi, int Lh, int mode, float* y, int out_h,
                               int out_w, void* stream) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (!low || !y) return B200W_ERR_NULL_POINTER;
    if (planes < 1 || h < 1 || w < 1 || out_h < 1 || out_w < 1) return B200W_ERR_BAD_SHAPE;
    SfbParams p;
    int rc = fill_taps(p.t, w_lo, w_hi, Lw, h_lo, h_hi, Lh);
    if (rc) return rc;
    const bool per = mode == B200W_MODE_PERIODIZATION;
    if (per && (2 * h < Lh || 2 * w < Lw)) return B200W_ERR_PER_TOO_SHORT;
    const int full_h = per ? 2 * h : 2 * h - Lh + 2;
    const int full_w = per ? 2 * w : 2 * w - Lw + 2;
    if (out_h > full_h || out_w > full_w) return B200W_ERR_BAD_SHAPE;
    p.low = low;
    p.highs = highs;
    p.y = y;
    p.low_ps = low_plane_stride;
    p.low_rs = low_row_stride;
    p.planes = planes;
    p.h = h;
    p.w = w;
    p.out_h = out_h;
    p.out_w = out_w;
    p.periodic = per ? 1 : 0;
    p.offW = per ? Lw / 2 - 1 : Lw - 2;
    p.offH = per ? Lh / 2 - 1 : Lh - 2;
    p.Lw = Lw;
    p.Lh = Lh;
    p.a0W = p.a0H = p.tiles_w = p.tiles_h = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (Lw == Lh && !force_direct()) {
        switch (Lw) {
            case 2: return launch_sfb_tiled<2>(p, st);
            case 4: return launch_sfb_tiled<4>(p, st);
            case 6: return launch_sfb_tiled<6>(p, st);
            case 8: return launch_sfb_tiled<8>(p, st);
            case 10: return launch_sfb_tiled<10>(p, st);
            case 12: return launch_sfb_tiled<12>(p, st);
            case 14: return launch_sfb_tiled<14>(p, st);
            case 16: return launch_sfb_tiled<16>(p, st);
            default: break;
        }
    }
    const size_t total = (size_t)planes * out_h * out_w;
    return launch(sfb2d_direct_kernel, p, direct_grid(total), 0, st);
}
