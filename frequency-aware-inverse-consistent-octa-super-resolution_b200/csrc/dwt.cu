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
//
// These are HBM-bound stencils (8 B of traffic per pixel for 2L FMAs): the kernels are organised to keep
// the instruction count per pixel low -- 128-bit shared-memory loads feeding register-blocked FMAs whose
// coefficients come from the constant bank, 64-bit coalesced global stores, and a staging loop in which
// every thread owns one vector column of the tile so that addresses advance by a constant.
#include "common.cuh"

namespace b200w {

// Persistent tile schedule: CTA b handles tiles b, b+G, b+2G, ... of the (plane, tile_h, tile_w) space, w fastest,
// so that CTAs running at the same time work on neighbouring tiles (halo re-reads hit L2).  The step G is
// pre-decomposed on the host so the per-tile update needs no division.
struct TileSched {
    long long total;
    int tiles_w, tiles_h;
    int d_w, d_h, d_p;  // G = (d_p * tiles_h + d_h) * tiles_w + d_w
};

struct TileIter {
    long long tile;
    int tw, th, plane;
    __device__ __forceinline__ void init(const TileSched& s) {
        tile = blockIdx.x;
        tw = (int)(tile % s.tiles_w);
        const long long t2 = tile / s.tiles_w;
        th = (int)(t2 % s.tiles_h);
        plane = (int)(t2 / s.tiles_h);
    }
    __device__ __forceinline__ void next(const TileSched& s) {
        tile += gridDim.x;
        tw += s.d_w;
        if (tw >= s.tiles_w) { tw -= s.tiles_w; ++th; }
        th += s.d_h;
        if (th >= s.tiles_h) { th -= s.tiles_h; ++plane; }
        plane += s.d_p;
    }
};

struct AfbParams {
    const float* x;
    float* low;
    float* highs;
    long long x_ps, x_rs;
    int planes, H, W, Ho, Wo;
    int mode, offW, offH, Lw, Lh;
    int in_vec;     // widest aligned vector (1, 2 or 4 floats) usable for staging copies
    int out_vec2;   // 64-bit stores allowed (Wo even, 8 B aligned bases)
    TileSched sched;
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
    int in_vec2;   // 64-bit staging copies allowed
    int out_vec2;  // 64-bit stores allowed
    TileSched sched;
    Taps t;
};

__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------
// analysis, tiled + persistent.  Output tile TH x TW per iteration (x4 sub-bands), NT threads.
// ------------------------------------------------------------------------------------------------
template <int L, int TW, int TH, int NT>
struct AfbCfg {
    static constexpr int PC = 2 * TW + L - 2;     // staged patch columns actually needed
    static constexpr int PCP = (PC + 3) & ~3;     // row pitch (multiple of 4 floats: 128-bit LDS)
    static constexpr int PR = 2 * TH + L - 2;     // staged patch rows
    static constexpr int NP = TW / 2;             // output pairs per row (row pass)
    static constexpr int RSTEP = NT / NP;         // patch rows advanced per row-pass iteration
    static constexpr int NV = (L + 2 + 3) / 4;    // float4 loads per row-pass item
    static constexpr int CP = TW / 2;             // column pairs (column pass)
    static constexpr int NS = NT / CP;            // row strips in the column pass
    static constexpr int RS = TH / NS;            // output rows per thread in the column pass
    static constexpr size_t smem = sizeof(float) * (size_t)(2 * PR * PCP + 2 * PR * TW);  // 2 patches + mid
    static_assert(L % 2 == 0 && TW % 4 == 0 && NT % CP == 0 && TH % NS == 0 && NT % NP == 0, "bad tile");
    static_assert(2 * TW - 4 + 4 * NV <= PCP, "row pass would read past the patch row");
    static_assert(PCP <= NT, "staging needs one thread per scalar column");
};

// Stage ROWS x PITCH floats asynchronously.  Every thread owns one vector column (V floats) and walks down
// the rows, so both addresses advance by constants.  (r0, c0) = source coordinates of patch element (0,0);
// the padding mode is applied as an index map, "zero" = zero fill.
template <int V, int ROWS, int PITCH, int NT>
__device__ __forceinline__ void stage_analysis(unsigned patch_s, const float* __restrict__ xp, long long rs, int r0,
                                               int c0, int H, int W, int mode, int nrows, int ncols, int tid) {
    constexpr int NVC = PITCH / V;   // vector columns per row
    constexpr int NRG = NT / NVC;    // row groups
    if (tid >= NVC * NRG) return;
    const int cv = tid % NVC;
    const int rg = tid / NVC;
    if (V * cv >= ncols) return;     // columns that feed no valid output of an edge tile are not staged
    const int sc0 = c0 + V * cv;
    const bool col_in = sc0 >= 0 && sc0 + V <= W;
    unsigned dst = patch_s + (unsigned)((rg * PITCH + V * cv) * 4);
    if (col_in && r0 >= 0 && r0 + nrows <= H) {
        const float* src = xp + (long long)(r0 + rg) * rs + sc0;
        const long long step = (long long)NRG * rs;
#pragma unroll 4
        for (int r = rg; r < nrows; r += NRG) {
            cp_async<V>(dst, src);
            dst += NRG * PITCH * 4;
            src += step;
        }
        return;
    }
    int ci[V];
#pragma unroll
    for (int e = 0; e < V; ++e) ci[e] = ext_index(sc0 + e, W, mode);
    for (int r = rg; r < nrows; r += NRG, dst += NRG * PITCH * 4) {
        const int sr = ext_index(r0 + r, H, mode);
        if (sr < 0) {
            cp_async_zero<V>(dst, xp);
            continue;
        }
        const float* rowp = xp + (long long)sr * rs;
        if (col_in) {
            cp_async<V>(dst, rowp + sc0);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) cp_async4_if(dst + 4 * e, ci[e] >= 0 ? rowp + ci[e] : xp, ci[e] >= 0);
        }
    }
}

template <int L, int TW, int TH, int NT>
__device__ __forceinline__ void afb_issue(unsigned patch_s, const AfbParams& p, const TileIter& it, int tid) {
    using Cfg = AfbCfg<L, TW, TH, NT>;
    const int r0 = 2 * it.th * TH - p.offH;  // source row of patch row 0
    const int c0 = 2 * it.tw * TW - p.offW;
    const float* xp = p.x + (long long)it.plane * p.x_ps;
    // an edge tile only needs the rows / columns its valid outputs read: 2*(n_valid-1) + L of them
    const int nrows = min(Cfg::PR, 2 * (min(TH, p.Ho - it.th * TH) - 1) + L);
    const int ncols = min(Cfg::PCP, 2 * (min(TW, p.Wo - it.tw * TW) - 1) + L);
    if (p.in_vec == 4) stage_analysis<4, Cfg::PR, Cfg::PCP, NT>(patch_s, xp, p.x_rs, r0, c0, p.H, p.W, p.mode, nrows, ncols, tid);
    else if (p.in_vec == 2) stage_analysis<2, Cfg::PR, Cfg::PCP, NT>(patch_s, xp, p.x_rs, r0, c0, p.H, p.W, p.mode, nrows, ncols, tid);
    else stage_analysis<1, Cfg::PR, Cfg::PCP, NT>(patch_s, xp, p.x_rs, r0, c0, p.H, p.W, p.mode, nrows, ncols, tid);
}

template <int L, int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) afb2d_tile_kernel(const __grid_constant__ AfbParams p) {
    using Cfg = AfbCfg<L, TW, TH, NT>;
    constexpr int PCP = Cfg::PCP, PR = Cfg::PR, NP = Cfg::NP, NV = Cfg::NV, CP = Cfg::CP, RS = Cfg::RS,
                  RSTEP = Cfg::RSTEP;
    extern __shared__ __align__(16) float smem[];
    float* mid_lo = smem + 2 * PR * PCP;  // [PR][TW]   (two patch buffers [PR][PCP] come first)
    float* mid_hi = mid_lo + PR * TW;     // [PR][TW]
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);

    const int tid = threadIdx.x;
    TileIter it, nx;
    it.init(p.sched);
    nx = it;
    afb_issue<L, TW, TH, NT>(smem_s, p, it, tid);  // prologue: start fetching the first tile
    cp_async_commit();

    for (int buf = 0; it.tile < p.sched.total; buf ^= 1, it = nx) {
        const float* patch = smem + buf * (PR * PCP);
        // prefetch this CTA's next tile into the other buffer while the current one is processed
        nx.next(p.sched);
        if (nx.tile < p.sched.total) afb_issue<L, TW, TH, NT>(smem_s + (unsigned)((buf ^ 1) * (PR * PCP * 4)), p, nx, tid);
        cp_async_commit();
        cp_async_wait<1>();  // this thread's copies of the current tile have landed ...
        __syncthreads();     // ... and everybody else's; also orders the previous tile's column pass before mid_* is rewritten

        // ---- row pass (along W), decimate by 2: each item makes two adjacent outputs of one patch row from
        //      NV 128-bit shared loads (conflict free: consecutive lanes read consecutive float4)
        {
            const int kk = tid % NP;
            int r = tid / NP;
            const float* src = patch + r * PCP + 4 * kk;
            float* dlo = mid_lo + r * TW + 2 * kk;
#pragma unroll 2
            for (; r < PR; r += RSTEP, src += RSTEP * PCP, dlo += RSTEP * TW) {
                float v[4 * NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const float4 t = reinterpret_cast<const float4*>(src)[q];
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                float lo0 = 0.f, hi0 = 0.f, lo1 = 0.f, hi1 = 0.f;
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    lo0 = fmaf(p.t.w_lo[j], v[j], lo0);
                    hi0 = fmaf(p.t.w_hi[j], v[j], hi0);
                    lo1 = fmaf(p.t.w_lo[j], v[j + 2], lo1);
                    hi1 = fmaf(p.t.w_hi[j], v[j + 2], hi1);
                }
                *reinterpret_cast<float2*>(dlo) = make_float2(lo0, lo1);
                *reinterpret_cast<float2*>(dlo + PR * TW) = make_float2(hi0, hi1);
            }
        }
        __syncthreads();

        // ---- column pass (along H), decimate by 2; each thread owns RS output rows of two adjacent columns
        {
            const int cp = tid % CP;
            const int s = tid / CP;
            float2 acc[RS][4];
#pragma unroll
            for (int i = 0; i < RS; ++i)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[i][b] = make_float2(0.f, 0.f);
            const float2* plo = reinterpret_cast<const float2*>(mid_lo + (2 * s * RS) * TW + 2 * cp);
            const float2* phi = reinterpret_cast<const float2*>(mid_hi + (2 * s * RS) * TW + 2 * cp);
#pragma unroll
            for (int rr = 0; rr < 2 * RS + L - 2; ++rr) {
                const float2 vlo = plo[rr * (TW / 2)];
                const float2 vhi = phi[rr * (TW / 2)];
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    const int j = rr - 2 * i;
                    if (j >= 0 && j < L) {
                        const float a = p.t.h_lo[j], b = p.t.h_hi[j];
                        acc[i][0].x = fmaf(a, vlo.x, acc[i][0].x); acc[i][0].y = fmaf(a, vlo.y, acc[i][0].y);  // LL
                        acc[i][1].x = fmaf(b, vlo.x, acc[i][1].x); acc[i][1].y = fmaf(b, vlo.y, acc[i][1].y);  // LH
                        acc[i][2].x = fmaf(a, vhi.x, acc[i][2].x); acc[i][2].y = fmaf(a, vhi.y, acc[i][2].y);  // HL
                        acc[i][3].x = fmaf(b, vhi.x, acc[i][3].x); acc[i][3].y = fmaf(b, vhi.y, acc[i][3].y);  // HH
                    }
                }
            }
            const int kk = it.tw * TW + 2 * cp;
            const int row0 = it.th * TH + s * RS;
            const size_t band = (size_t)p.Ho * p.Wo;
            float* q0 = p.low + (size_t)it.plane * band + (size_t)row0 * p.Wo + kk;
            float* q1 = p.highs + (size_t)it.plane * 3 * band + (size_t)row0 * p.Wo + kk;
            float* q2 = q1 + band;
            float* q3 = q2 + band;
            if (p.out_vec2 && row0 + RS <= p.Ho && kk + 1 < p.Wo) {  // whole strip inside, 64-bit stores
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    *reinterpret_cast<float2*>(q0) = acc[i][0];
                    *reinterpret_cast<float2*>(q1) = acc[i][1];
                    *reinterpret_cast<float2*>(q2) = acc[i][2];
                    *reinterpret_cast<float2*>(q3) = acc[i][3];
                    q0 += p.Wo; q1 += p.Wo; q2 += p.Wo; q3 += p.Wo;
                }
            } else if (kk < p.Wo) {
                const bool second = kk + 1 < p.Wo;
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    if (row0 + i < p.Ho) {
                        q0[0] = acc[i][0].x; q1[0] = acc[i][1].x; q2[0] = acc[i][2].x; q3[0] = acc[i][3].x;
                        if (second) { q0[1] = acc[i][0].y; q1[1] = acc[i][1].y; q2[1] = acc[i][2].y; q3[1] = acc[i][3].y; }
                    }
                    q0 += p.Wo; q1 += p.Wo; q2 += p.Wo; q3 += p.Wo;
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
// synthesis, tiled + persistent.  Works in "A-space": a = n + off, so that the polyphase split (which taps
// an output uses) depends only on the parity of the tile-local coordinate.  Output tile TH x TW.
// W synthesis first (on the KH coefficient rows), then H synthesis; by separability this equals the
// reference's H-then-W order (pw/dwt/lowlevel.py:677-679) up to fp32 rounding.
// ------------------------------------------------------------------------------------------------
template <int L, int TW, int TH, int NT>
struct SfbCfg {
    static constexpr int H2 = L / 2;
    static constexpr int NV2 = (H2 + 2) / 2;            // float2 loads per band per W-synthesis item
    static constexpr int KWP = TW / 2 - 2 + 2 * NV2;    // staged coefficient columns (even, >= TW/2 + H2 - 1)
    static constexpr int KH = TH / 2 + H2 - 1;          // staged coefficient rows
    static constexpr int PB = KH * KWP;                 // one band's patch
    static constexpr int NQ = TW / 4;                   // W-synthesis items per coefficient row (4 outputs each)
    static constexpr int QSTEP = NT / NQ;               // coefficient rows advanced per W-synthesis iteration
    static constexpr int CP = TW / 2;                   // output column pairs (H pass)
    static constexpr int NS = NT / CP;
    static constexpr int RS = TH / NS;                  // output rows per thread in the H pass (even)
    static constexpr int NR = RS / 2 + H2 - 1;          // coefficient rows feeding RS outputs
    static constexpr size_t smem = sizeof(float) * (size_t)(2 * 4 * PB + 2 * KH * TW);  // 2 x 4 patches + u
    static_assert(L % 2 == 0 && TW % 4 == 0 && NT % CP == 0 && TH % NS == 0 && RS % 2 == 0 && NT % NQ == 0, "bad tile");
    static_assert(KWP >= TW / 2 + H2 - 1 && KWP % 2 == 0 && KWP <= NT, "bad KWP");
};

// Stage the 4 sub-band patches [4][KH][KWP] asynchronously; thread = one vector column, walking down the rows.
// Coefficients outside the arrays are zero (or wrap for periodization); missing `highs` = zeros.
template <int V, int KH, int KWP, int NT>
__device__ __forceinline__ void stage_synthesis(unsigned sub_s, const float* __restrict__ lowp, long long low_rs,
                                                const float* __restrict__ hip, size_t band, int kH0, int kW0, int h,
                                                int w, bool periodic, int nrows, int ncols, int tid) {
    constexpr int NVC = KWP / V;
    constexpr int NRG = NT / NVC;
    constexpr unsigned PBB = KH * KWP * 4;  // bytes of one band's patch
    if (tid >= NVC * NRG) return;
    const int cv = tid % NVC;
    const int rg = tid / NVC;
    if (V * cv >= ncols) return;
    const int kc0 = kW0 + V * cv;
    const bool col_in = kc0 >= 0 && kc0 + V <= w;
    unsigned dst = sub_s + (unsigned)((rg * KWP + V * cv) * 4);
    if (col_in && kH0 >= 0 && kH0 + nrows <= h && hip != nullptr) {
        const float* lp = lowp + (long long)(kH0 + rg) * low_rs + kc0;
        const float* hp = hip + (size_t)(kH0 + rg) * w + kc0;
        const long long lstep = (long long)NRG * low_rs;
        const size_t hstep = (size_t)NRG * w;
#pragma unroll 2
        for (int r = rg; r < nrows; r += NRG) {
            cp_async<V>(dst, lp);
            cp_async<V>(dst + PBB, hp);
            cp_async<V>(dst + 2 * PBB, hp + band);
            cp_async<V>(dst + 3 * PBB, hp + 2 * band);
            dst += NRG * KWP * 4;
            lp += lstep;
            hp += hstep;
        }
        return;
    }
    int ci[V];
#pragma unroll
    for (int e = 0; e < V; ++e) ci[e] = coef_index(kc0 + e, w, periodic);
    for (int r = rg; r < nrows; r += NRG, dst += NRG * KWP * 4) {
        const int kr = coef_index(kH0 + r, h, periodic);
        if (kr < 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b) cp_async_zero<V>(dst + b * PBB, lowp);
            continue;
        }
        const float* lp = lowp + (long long)kr * low_rs;
        const float* hp = hip ? hip + (size_t)kr * w : nullptr;
        if (col_in) {
            cp_async<V>(dst, lp + kc0);
            if (hp) {
                cp_async<V>(dst + PBB, hp + kc0);
                cp_async<V>(dst + 2 * PBB, hp + band + kc0);
                cp_async<V>(dst + 3 * PBB, hp + 2 * band + kc0);
            } else {
#pragma unroll
                for (int b = 1; b < 4; ++b) cp_async_zero<V>(dst + b * PBB, lowp);
            }
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool ok = ci[e] >= 0;
                const int k = ok ? ci[e] : 0;
                cp_async4_if(dst + 4 * e, lp + k, ok);
                cp_async4_if(dst + PBB + 4 * e, hp ? hp + k : lowp, ok && hp);
                cp_async4_if(dst + 2 * PBB + 4 * e, hp ? hp + band + k : lowp, ok && hp);
                cp_async4_if(dst + 3 * PBB + 4 * e, hp ? hp + 2 * band + k : lowp, ok && hp);
            }
        }
    }
}

template <int L, int TW, int TH, int NT>
__device__ __forceinline__ void sfb_issue(unsigned sub_s, const SfbParams& p, const TileIter& it, int tid) {
    using Cfg = SfbCfg<L, TW, TH, NT>;
    const int kW0 = (p.a0W + it.tw * TW) / 2 - (Cfg::H2 - 1);
    const int kH0 = (p.a0H + it.th * TH) / 2 - (Cfg::H2 - 1);
    const size_t band = (size_t)p.h * p.w;
    const float* lowp = p.low + (long long)it.plane * p.low_ps;
    const float* hip = p.highs ? p.highs + (size_t)it.plane * 3 * band : nullptr;
    // coefficient rows / columns feeding the valid outputs of an edge tile: local a <= a_max -> k_local <= a_max/2 + H2-1
    const int amax_h = min(TH - 1, p.offH + p.out_h - 1 - (p.a0H + it.th * TH));
    const int amax_w = min(TW - 1, p.offW + p.out_w - 1 - (p.a0W + it.tw * TW));
    const int nrows = min(Cfg::KH, amax_h / 2 + Cfg::H2);
    const int ncols = min(Cfg::KWP, amax_w / 2 + Cfg::H2);
    // kW0 is even whenever in_vec2 is set (checked on the host)
    if (p.in_vec2) stage_synthesis<2, Cfg::KH, Cfg::KWP, NT>(sub_s, lowp, p.low_rs, hip, band, kH0, kW0, p.h, p.w, p.periodic, nrows, ncols, tid);
    else stage_synthesis<1, Cfg::KH, Cfg::KWP, NT>(sub_s, lowp, p.low_rs, hip, band, kH0, kW0, p.h, p.w, p.periodic, nrows, ncols, tid);
}

template <int L, int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) sfb2d_tile_kernel(const __grid_constant__ SfbParams p) {
    using Cfg = SfbCfg<L, TW, TH, NT>;
    constexpr int H2 = Cfg::H2, NV2 = Cfg::NV2, KWP = Cfg::KWP, KH = Cfg::KH, PB = Cfg::PB, NQ = Cfg::NQ,
                  QSTEP = Cfg::QSTEP, CP = Cfg::CP, RS = Cfg::RS, NR = Cfg::NR;
    extern __shared__ __align__(16) float smem[];
    float* u_lo = smem + 2 * 4 * PB;   // [KH][TW]  W-synthesised, to be combined with h_lo (2 patch sets first)
    float* u_hi = u_lo + KH * TW;      // [KH][TW]  ... with h_hi
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);

    const int tid = threadIdx.x;
    TileIter it, nx;
    it.init(p.sched);
    nx = it;
    sfb_issue<L, TW, TH, NT>(smem_s, p, it, tid);
    cp_async_commit();

    for (int buf = 0; it.tile < p.sched.total; buf ^= 1, it = nx) {
        const float* sub = smem + buf * (4 * PB);  // [4][KH][KWP]  LL, LH, HL, HH
        nx.next(p.sched);
        if (nx.tile < p.sched.total) sfb_issue<L, TW, TH, NT>(smem_s + (unsigned)((buf ^ 1) * (4 * PB * 4)), p, nx, tid);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        // ---- W synthesis: each item makes four consecutive outputs (a = 4qq .. 4qq+3) of one coefficient row,
        //      for both the h_lo branch (LL, HL) and the h_hi branch (LH, HH)
        {
            const int qq = tid % NQ;
            int r = tid / NQ;
            const float* s0 = sub + r * KWP + 2 * qq;
            float* d = u_lo + r * TW + 4 * qq;
            for (; r < KH; r += QSTEP, s0 += QSTEP * KWP, d += QSTEP * TW) {
                float c[4][2 * NV2];  // local coefficients 2qq .. 2qq+2NV2-1 of LL, LH, HL, HH
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int q = 0; q < NV2; ++q) {
                        const float2 t = reinterpret_cast<const float2*>(s0 + b * PB)[q];
                        c[b][2 * q] = t.x;
                        c[b][2 * q + 1] = t.y;
                    }
                float lo[4], hi[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qo = e >> 1, par = e & 1;  // output a = 2(2qq+qo) + par uses k_local = qo + H2-1-u
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int u = 0; u < H2; ++u) {
                        const int k = qo + H2 - 1 - u;
                        a = fmaf(c[0][k], p.t.w_lo[par + 2 * u], a);
                        a = fmaf(c[2][k], p.t.w_hi[par + 2 * u], a);
                        b = fmaf(c[1][k], p.t.w_lo[par + 2 * u], b);
                        b = fmaf(c[3][k], p.t.w_hi[par + 2 * u], b);
                    }
                    lo[e] = a;
                    hi[e] = b;
                }
                *reinterpret_cast<float4*>(d) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(d + KH * TW) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
        }
        __syncthreads();

        // ---- H synthesis: thread = two adjacent output columns, RS consecutive output rows
        {
            const int cp = tid % CP;
            const int s = tid / CP;
            float2 vlo[NR], vhi[NR];
            const float2* plo = reinterpret_cast<const float2*>(u_lo + (s * (RS / 2)) * TW + 2 * cp);
            const float2* phi = reinterpret_cast<const float2*>(u_hi + (s * (RS / 2)) * TW + 2 * cp);
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                vlo[r] = plo[r * (TW / 2)];
                vhi[r] = phi[r * (TW / 2)];
            }
            float2 y[RS];
#pragma unroll
            for (int i = 0; i < RS / 2; ++i) {
                float2 ye = make_float2(0.f, 0.f), yo = ye;
#pragma unroll
                for (int u = 0; u < H2; ++u) {
                    const int r = i - u + H2 - 1;  // local coefficient row
                    const float a0 = p.t.h_lo[2 * u], b0 = p.t.h_hi[2 * u], a1 = p.t.h_lo[2 * u + 1], b1 = p.t.h_hi[2 * u + 1];
                    ye.x = fmaf(vlo[r].x, a0, ye.x); ye.y = fmaf(vlo[r].y, a0, ye.y);
                    ye.x = fmaf(vhi[r].x, b0, ye.x); ye.y = fmaf(vhi[r].y, b0, ye.y);
                    yo.x = fmaf(vlo[r].x, a1, yo.x); yo.y = fmaf(vlo[r].y, a1, yo.y);
                    yo.x = fmaf(vhi[r].x, b1, yo.x); yo.y = fmaf(vhi[r].y, b1, yo.y);
                }
                y[2 * i] = ye;
                y[2 * i + 1] = yo;
            }
            const int nW = p.a0W + it.tw * TW + 2 * cp - p.offW;
            const int nH = p.a0H + it.th * TH + s * RS - p.offH;
            float* q = p.y + ((size_t)it.plane * p.out_h + nH) * p.out_w + nW;  // only dereferenced where valid
            if (p.out_vec2 && nH >= 0 && nH + RS <= p.out_h && nW >= 0 && nW + 1 < p.out_w) {
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    *reinterpret_cast<float2*>(q) = y[i];
                    q += p.out_w;
                }
            } else {
                const bool ok0 = nW >= 0 && nW < p.out_w;
                const bool ok1 = nW + 1 >= 0 && nW + 1 < p.out_w;
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    const int row = nH + i;
                    if (row >= 0 && row < p.out_h) {
                        if (ok0) q[0] = y[i].x;
                        if (ok1) q[1] = y[i].y;
                    }
                    q += p.out_w;
                }
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

static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

constexpr int kMaxDevices = 64;

struct DeviceInfo {
    int dev;
    int sms;
};

static DeviceInfo device_info() {
    static int sms[kMaxDevices] = {0};
    DeviceInfo d{0, 148};
    if (cudaGetDevice(&d.dev) != cudaSuccess || d.dev < 0 || d.dev >= kMaxDevices) {
        d.dev = 0;
        return d;
    }
    if (sms[d.dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d.dev) != cudaSuccess || n <= 0) n = 148;
        sms[d.dev] = n;
    }
    d.sms = sms[d.dev];
    return d;
}

// Persistent launch: one CTA per resident slot (SMs x occupancy), each looping over tiles.
// `occ` is the caller's per-kernel, per-device cache of the occupancy (0 = not yet queried); the query also
// raises the kernel's dynamic shared memory limit once.  Function attributes and occupancy are immutable
// facts about (kernel, device), so caching them keeps the library re-entrant.
template <typename K, typename P>
static int launch_tiles(K kernel, P& p, int tiles_w, int tiles_h, int threads, size_t smem, int* occ_cache,
                        cudaStream_t st) {
    const DeviceInfo di = device_info();
    int& occ = occ_cache[di.dev];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, smem);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        occ = o > 0 ? o : 1;
    }
    const long long total = (long long)tiles_w * tiles_h * p.planes;
    long long grid = (long long)di.sms * occ;
    if (grid > total) grid = total;
    p.sched.total = total;
    p.sched.tiles_w = tiles_w;
    p.sched.tiles_h = tiles_h;
    p.sched.d_w = (int)(grid % tiles_w);
    p.sched.d_h = (int)((grid / tiles_w) % tiles_h);
    p.sched.d_p = (int)(grid / ((long long)tiles_w * tiles_h));
    kernel<<<(unsigned)grid, threads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <typename K, typename P>
static int launch_flat(K kernel, const P& p, size_t grid, cudaStream_t st) {
    kernel<<<(unsigned)grid, kThreads, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

template <int L>
static int launch_afb_tiled(AfbParams& p, cudaStream_t st) {
    // two tile shapes: 32x32 outputs / 128 threads, or 64 wide x 32 / 256 threads (less halo); take the wide
    // one unless it wastes noticeably more of the padded output area
    const long long a32 = (long long)ceil_div(p.Wo, 32) * 32;
    const long long a64 = (long long)ceil_div(p.Wo, 64) * 64;
    if (p.Wo >= 128 && a64 * 100 <= a32 * 103) {
        constexpr int TW = 64, TH = 32, NT = 256;
        static int occ[kMaxDevices] = {0};
        return launch_tiles(afb2d_tile_kernel<L, TW, TH, NT>, p, ceil_div(p.Wo, TW), ceil_div(p.Ho, TH), NT,
                            AfbCfg<L, TW, TH, NT>::smem, occ, st);
    }
    constexpr int TW = 32, TH = 32, NT = 128;
    static int occ[kMaxDevices] = {0};
    return launch_tiles(afb2d_tile_kernel<L, TW, TH, NT>, p, ceil_div(p.Wo, TW), ceil_div(p.Ho, TH), NT,
                        AfbCfg<L, TW, TH, NT>::smem, occ, st);
}

template <int L>
static int launch_sfb_tiled(SfbParams& p, cudaStream_t st) {
    constexpr int TW = 64, TH = 64, NT = 256;
    static int occ[kMaxDevices] = {0};
    p.a0W = p.offW & ~1;
    p.a0H = p.offH & ~1;
    const int tiles_w = ceil_div(p.offW + p.out_w - p.a0W, TW);
    const int tiles_h = ceil_div(p.offH + p.out_h - p.a0H, TH);
    // tile 0 starts at coefficient column a0W/2 - (L/2-1): even for every non-periodization mode
    const int kW0 = p.a0W / 2 - (L / 2 - 1);
    if ((kW0 & 1) != 0) p.in_vec2 = 0;
    if ((p.offW & 1) != 0) p.out_vec2 = 0;
    return launch_tiles(sfb2d_tile_kernel<L, TW, TH, NT>, p, tiles_w, tiles_h, NT, SfbCfg<L, TW, TH, NT>::smem, occ,
                        st);
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
    // staging vector width: the first staged column of every tile is 2*TW*tw - offW
    p.in_vec = 1;
    if ((p.offW % 2) == 0 && (x_row_stride % 2) == 0 && (x_plane_stride % 2) == 0 && aligned_to(x, 8)) p.in_vec = 2;
    if (p.in_vec == 2 && (p.offW % 4) == 0 && (x_row_stride % 4) == 0 && (x_plane_stride % 4) == 0 && aligned_to(x, 16))
        p.in_vec = 4;
    p.out_vec2 = ((p.Wo % 2) == 0 && aligned_to(low, 8) && aligned_to(highs, 8)) ? 1 : 0;
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
    return launch_flat(afb2d_direct_kernel, p, direct_grid(total), st);
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
    p.a0W = p.a0H = 0;
    p.in_vec2 = ((w % 2) == 0 && (low_row_stride % 2) == 0 && (low_plane_stride % 2) == 0 && aligned_to(low, 8) &&
                 (!highs || aligned_to(highs, 8))) ? 1 : 0;
    p.out_vec2 = ((out_w % 2) == 0 && aligned_to(y, 8)) ? 1 : 0;
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
    return launch_flat(sfb2d_direct_kernel, p, direct_grid(total), st);
}
