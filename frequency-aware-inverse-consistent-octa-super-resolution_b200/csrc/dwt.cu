// 2-D DWT analysis / synthesis filter banks for sm_100a: one persistent kernel per multi-level transform.
//
// What it replaces.  Per level the reference runs pad-gather + 2x F.conv2d + reshape + 2x .contiguous()
// (AFB2D, pw/dwt/lowlevel.py:336-347) resp. 6x F.conv_transpose2d + 3 adds (SFB2D, :671-680), and
// DWTForward / DWTInverse loop over the J levels in Python (pw/dwt/transform2d.py:66-74, 134-148).
// Here a J-level transform is ONE launch: a persistent grid walks an ordered list of tiles (all tiles of the
// first level, then the next level, ...); a tile of level j+1 of image plane p starts as soon as every tile of
// level j of that plane has been written (per-plane completion counters, release/acquire), so levels overlap,
// there are no launch gaps between the small coarse levels, and the LL intermediates are consumed out of L2.
//
// Per tile, analysis: the input patch (+halo) is staged in shared memory with cp.async -- the padding mode is
// an index map applied to the source address -- double buffered so the next tile's fetch overlaps this tile's
// arithmetic; row (W) pass with stride-2 decimation into shared memory; column (H) pass; LL goes to `low`,
// LH/HL/HH straight into `highs[:, :, 0..2]`.  Synthesis is the polyphase mirror image (upsample + filter +
// accumulate of all four sub-bands fused).
//
// Closed forms (validated against the reference to 1e-15 by oracle/dwt_oracle.py):
//   analysis   y_c[k] = sum_j w_c[j] * x_ext[2k + j - off],  off = p//2 with p = 2(M-1) - N + L,
//              periodization: off = L-1 - L//2 on the (even-extended) N'-periodic signal
//   synthesis  y[n]   = sum_k lo[k] g0[t] + hi[k] g1[t],  t = n + off - 2k in [0,L),
//              off = L-2 (coefficients outside [0,M) are zero), periodization: off = L//2 - 1 and the
//              coefficient sequence is M-periodic
//
// These are HBM-bound stencils (8 B of traffic per pixel for 2L FMAs): the code keeps the instruction count
// per pixel low -- 128-bit shared loads feeding register-blocked FMAs whose coefficients come from the
// constant bank, 64-bit coalesced global stores, staging loops whose addresses advance by constants.
#include "dwt_levels.cuh"
#include "dwt_tma.cuh"

namespace b200w {

__global__ void zero_words_kernel(unsigned* w, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = 0u;
}

int zero_sync_words(unsigned* words, size_t n, cudaStream_t st) {
    zero_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(words, (unsigned)n);
    note_launch("zero_words_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

struct TileRef {
    int level, plane, th, tw;
};

template <class P>
__device__ __forceinline__ TileRef decode_tile(const P& p, long long tile) {
    int level = 0;
    for (int j = 1; j < p.J; ++j)
        if (tile >= p.lv[j].tile_base) level = j;
    const unsigned local = (unsigned)(tile - p.lv[level].tile_base);
    const unsigned ntw = (unsigned)p.lv[level].tiles_w;
    const unsigned tpp = ntw * (unsigned)p.lv[level].tiles_h;
    TileRef t;
    t.level = level;
    t.plane = (int)(local / tpp);
    const unsigned rem = local - (unsigned)t.plane * tpp;
    t.th = (int)(rem / ntw);
    t.tw = (int)(rem - (unsigned)t.th * ntw);
    return t;
}

// has every tile of the previous level of this plane been written?
template <class P>
__device__ __forceinline__ bool tile_ready(const P& p, const TileRef& t) {
    if (t.level == 0) return true;
    const unsigned need = (unsigned)(p.lv[t.level - 1].tiles_w * p.lv[t.level - 1].tiles_h);
    return ld_acquire_u32(p.done + (size_t)(t.level - 1) * p.planes + t.plane) >= need;
}

// Persistent, double-buffered walk over the tile list.  Op provides the tile body:
//   issue(p, tile, shared-window address of the staging buffer, tid)   cp.async the tile's inputs
//   pass1(p, tile, staging buffer, scratch, tid)                        first 1-D pass, shared -> shared
//   pass2(p, tile, scratch, tid)                                        second 1-D pass + global stores
template <class Op>
__global__ void __launch_bounds__(Op::NT) chain_kernel(const __grid_constant__ typename Op::Params p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_pf_ok;
    float* scratch = smem + 2 * Op::BUF;   // two staging buffers come first
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x;
    const long long G = gridDim.x;

    long long tile = blockIdx.x;
    if (tile >= p.total) return;
    TileRef cur = decode_tile(p, tile), nxt = cur;
    if (cur.level > 0) {   // only when the grid is larger than the first level
        if (tid == 0)
            while (!tile_ready(p, cur)) __nanosleep(64);
        __syncthreads();
    }
    Op::issue(p, cur, smem_s, tid);
    cp_async_commit();
    // prefetch decision for the following tile; later ones are made by thread 0 inside the loop
    int pf_ok = (tile + G < p.total) && (tile + G < (p.J > 1 ? p.lv[1].tile_base : p.total));
    int pend_level = -1, pend_plane = 0;   // tile whose completion still has to be published

    for (int buf = 0;; buf ^= 1) {
        cp_async_wait<0>();   // this thread's copies of the current tile have landed ...
        __syncthreads();      // ... and everybody else's; the previous tile's stores are ordered before this barrier
        // publish the previous tile BEFORE this thread has copies in flight again (the fence waits for them)
        if (tid == 0 && pend_level >= 0) signal_done(p.done + (size_t)pend_level * p.planes + pend_plane);
        const long long ntile = tile + G;
        const bool have_next = ntile < p.total;
        bool prefetched = false;
        if (have_next) {
            nxt = decode_tile(p, ntile);
            if (pf_ok) {   // the next tile's inputs are complete: fetch them while this tile is processed
                Op::issue(p, nxt, smem_s + (unsigned)((buf ^ 1) * Op::BUF * 4), tid);
                prefetched = true;
            }
        }
        cp_async_commit();
        // thread 0 probes whether the tile after next will be fetchable; the load stays in flight across pass1
        unsigned dep_have = 1, dep_need = 0;
        if (tid == 0) {
            const long long nn = ntile + G;
            if (nn >= p.total) {
                dep_have = 0;
                dep_need = 1;
            } else {
                const TileRef t2 = decode_tile(p, nn);
                if (t2.level > 0) {
                    dep_need = (unsigned)(p.lv[t2.level - 1].tiles_w * p.lv[t2.level - 1].tiles_h);
                    dep_have = ld_acquire_u32(p.done + (size_t)(t2.level - 1) * p.planes + t2.plane);
                }
            }
        }
        Op::pass1(p, cur, smem + buf * Op::BUF, scratch, tid);
        if (tid == 0) s_pf_ok = dep_have >= dep_need ? 1 : 0;
        __syncthreads();
        Op::pass2(p, cur, scratch, tid);
        pf_ok = s_pf_ok;
        const bool has_consumer = cur.level + 1 < p.J;
        pend_level = has_consumer ? cur.level : -1;
        pend_plane = cur.plane;
        if (!have_next) break;
        if (!prefetched) {
            // the next tile's inputs were not complete when we looked: publish our own result first (it may be
            // the missing piece), then wait for the producers -- all of them own earlier tiles and are resident
            __syncthreads();
            if (tid == 0) {
                if (pend_level >= 0) signal_done(p.done + (size_t)pend_level * p.planes + pend_plane);
                while (!tile_ready(p, nxt)) __nanosleep(64);
            }
            pend_level = -1;
            __syncthreads();
            Op::issue(p, nxt, smem_s + (unsigned)((buf ^ 1) * Op::BUF * 4), tid);
            cp_async_commit();
        }
        tile = ntile;
        cur = nxt;
    }
    if (pend_level >= 0) {
        __syncthreads();
        if (tid == 0) signal_done(p.done + (size_t)pend_level * p.planes + pend_plane);
    }
}

// ------------------------------------------------------------------------------------------------
// analysis tile.  Output tile TH x TW (x4 sub-bands), NT threads.
// ------------------------------------------------------------------------------------------------
// Stage ROWS x PITCH floats asynchronously.  Every thread owns one vector column (V floats) and walks down
// the rows, so both addresses advance by constants.  (r0, c0) = source coordinates of patch element (0,0);
// the padding mode is applied as an index map, "zero" = zero fill (cp.async with src-size 0).
template <int V, int ROWS, int PITCH, int NT>
__device__ __forceinline__ void stage_analysis(unsigned patch_s, const float* __restrict__ xp, long long rs, int r0,
                                               int c0, int H, int W, int Hreal, int Wreal, int mode, int nrows,
                                               int ncols, int tid) {
    constexpr int NVC = PITCH / V;   // vector columns per row
    constexpr int NRG = NT / NVC;    // row groups
    if (tid >= NVC * NRG) return;
    const int cv = tid % NVC;
    const int rg = tid / NVC;
    if (V * cv >= ncols) return;     // columns that feed no valid output of an edge tile are not staged
    const int sc0 = c0 + V * cv;
    const bool col_in = sc0 >= 0 && sc0 + V <= Wreal;
    unsigned dst = patch_s + (unsigned)((rg * PITCH + V * cv) * 4);
    if (col_in && r0 >= 0 && r0 + nrows <= Hreal) {
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
    for (int e = 0; e < V; ++e) {
        ci[e] = ext_index(sc0 + e, W, mode);
        if (ci[e] >= Wreal) ci[e] = -1;
    }
    for (int r = rg; r < nrows; r += NRG, dst += NRG * PITCH * 4) {
        const int sr = ext_index(r0 + r, H, mode);
        if (sr < 0 || sr >= Hreal) {
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

template <int L_, int TW_, int TH_, int NT_>
struct AfbOp {
    using Params = AfbParams;
    static constexpr const char* name = "chain_kernel<AfbOp>";
    static constexpr int L = L_, TW = TW_, TH = TH_, NT = NT_;
    static constexpr int PC = 2 * TW + L - 2;     // staged patch columns actually needed
    static constexpr int PCP = (PC + 3) & ~3;     // row pitch (multiple of 4 floats: 128-bit LDS)
    static constexpr int PR = 2 * TH + L - 2;     // staged patch rows
    static constexpr int BUF = PR * PCP;          // floats per staging buffer
    static constexpr int NP = TW / 2;             // output pairs per row (row pass)
    static constexpr int RSTEP = NT / NP;         // patch rows advanced per row-pass iteration
    static constexpr int NV = (L + 2 + 3) / 4;    // float4 loads per row-pass item
    static constexpr int CP = TW / 2;             // column pairs (column pass)
    static constexpr int NS = NT / CP;            // row strips in the column pass
    static constexpr int RS = TH / NS;            // output rows per thread in the column pass
    static constexpr size_t smem = sizeof(float) * (size_t)(2 * BUF + 2 * PR * TW);  // 2 patches + mid_lo/mid_hi
    static_assert(L % 2 == 0 && TW % 4 == 0 && NT % CP == 0 && TH % NS == 0 && NT % NP == 0, "bad tile");
    static_assert(2 * TW - 4 + 4 * NV <= PCP, "row pass would read past the patch row");
    static_assert(PCP <= NT, "staging needs one thread per scalar column");

    static __device__ __forceinline__ void issue(const AfbParams& p, const TileRef& t, unsigned patch_s, int tid) {
        const AfbLevel& lv = p.lv[t.level];
        const int r0 = 2 * t.th * TH - lv.offH;  // source row of patch row 0
        const int c0 = 2 * t.tw * TW - lv.offW;
        const float* xp = lv.x + (long long)t.plane * lv.x_ps;
        // an edge tile only needs the rows / columns its valid outputs read: 2*(n_valid-1) + L of them
        const int nrows = min(PR, 2 * (min(TH, lv.Ho - t.th * TH) - 1) + L);
        const int ncols = min(PCP, 2 * (min(TW, lv.Wo - t.tw * TW) - 1) + L);
        if (lv.in_vec == 4)
            stage_analysis<4, PR, PCP, NT>(patch_s, xp, lv.x_rs, r0, c0, lv.H, lv.W, lv.Hreal, lv.Wreal, p.mode, nrows, ncols, tid);
        else if (lv.in_vec == 2)
            stage_analysis<2, PR, PCP, NT>(patch_s, xp, lv.x_rs, r0, c0, lv.H, lv.W, lv.Hreal, lv.Wreal, p.mode, nrows, ncols, tid);
        else
            stage_analysis<1, PR, PCP, NT>(patch_s, xp, lv.x_rs, r0, c0, lv.H, lv.W, lv.Hreal, lv.Wreal, p.mode, nrows, ncols, tid);
    }

    // row pass (along W), decimate by 2: each item makes two adjacent outputs of one patch row from NV 128-bit
    // shared loads (conflict free: consecutive lanes read consecutive float4)
    static __device__ __forceinline__ void pass1(const AfbParams& p, const TileRef&, const float* patch, float* mid,
                                                 int tid) {
        const int kk = tid % NP;
        int r = tid / NP;
        const float* src = patch + r * PCP + 4 * kk;
        float* dlo = mid + r * TW + 2 * kk;
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

    // column pass (along H), decimate by 2; each thread owns RS output rows of two adjacent columns
    static __device__ __forceinline__ void pass2(const AfbParams& p, const TileRef& t, const float* mid, int tid) {
        const AfbLevel& lv = p.lv[t.level];
        const int cp = tid % CP;
        const int s = tid / CP;
        float2 acc[RS][4];
#pragma unroll
        for (int i = 0; i < RS; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[i][b] = make_float2(0.f, 0.f);
        const float2* plo = reinterpret_cast<const float2*>(mid + (2 * s * RS) * TW + 2 * cp);
        const float2* phi = reinterpret_cast<const float2*>(mid + PR * TW + (2 * s * RS) * TW + 2 * cp);
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
                    acc[i][1].x = fmaf(b, vlo.x, acc[i][1].x); acc[i][1].y = fmaf(b, vlo.y, acc[i][1].y);  // LH: W-lo, H-hi
                    acc[i][2].x = fmaf(a, vhi.x, acc[i][2].x); acc[i][2].y = fmaf(a, vhi.y, acc[i][2].y);  // HL: W-hi, H-lo
                    acc[i][3].x = fmaf(b, vhi.x, acc[i][3].x); acc[i][3].y = fmaf(b, vhi.y, acc[i][3].y);  // HH
                }
            }
        }
        const int Ho = lv.Ho, Wo = lv.Wo;
        const int kk = t.tw * TW + 2 * cp;
        const int row0 = t.th * TH + s * RS;
        const size_t band = (size_t)Ho * Wo;
        const long long lrs = lv.low_rs;
        float* q0 = lv.low + (long long)t.plane * lv.low_ps + (long long)row0 * lrs + kk;
        float* q1 = lv.highs + (size_t)t.plane * 3 * band + (size_t)row0 * Wo + kk;
        float* q2 = q1 + band;
        float* q3 = q2 + band;
        if (lv.out_vec2 && lv.low_vec2 && row0 + RS <= Ho && kk + 1 < Wo) {  // whole strip inside: 64-bit stores
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                *reinterpret_cast<float2*>(q0) = acc[i][0];
                *reinterpret_cast<float2*>(q1) = acc[i][1];
                *reinterpret_cast<float2*>(q2) = acc[i][2];
                *reinterpret_cast<float2*>(q3) = acc[i][3];
                q0 += lrs; q1 += Wo; q2 += Wo; q3 += Wo;
            }
        } else if (kk < Wo) {
            const bool second = kk + 1 < Wo;
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                if (row0 + i < Ho) {
                    q0[0] = acc[i][0].x; q1[0] = acc[i][1].x; q2[0] = acc[i][2].x; q3[0] = acc[i][3].x;
                    if (second) { q0[1] = acc[i][0].y; q1[1] = acc[i][1].y; q2[1] = acc[i][2].y; q3[1] = acc[i][3].y; }
                }
                q0 += lrs; q1 += Wo; q2 += Wo; q3 += Wo;
            }
        }
    }
};

// analysis, direct (any tap counts up to kMaxTaps, odd or mixed lengths): one thread per output position of
// ONE level, no staging.  Fallback and on-device cross-check of the tiled kernel.
struct AfbDirectParams {
    AfbLevel lv;
    int planes, mode, Lw, Lh;
    Taps t;
};

__global__ void __launch_bounds__(kThreads) afb2d_direct_kernel(const __grid_constant__ AfbDirectParams p) {
    const AfbLevel& lv = p.lv;
    const size_t band = (size_t)lv.Ho * lv.Wo;
    const size_t total = band * p.planes;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % lv.Wo);
        const int i = (int)((idx / lv.Wo) % lv.Ho);
        const int plane = (int)(idx / band);
        const float* __restrict__ xp = lv.x + (long long)plane * lv.x_ps;
        float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
        for (int jh = 0; jh < p.Lh; ++jh) {
            const int sr = ext_index(2 * i + jh - lv.offH, lv.H, p.mode);
            if (sr < 0 || sr >= lv.Hreal) continue;
            float lo = 0.f, hi = 0.f;
            for (int jw = 0; jw < p.Lw; ++jw) {
                const int sc = ext_index(2 * k + jw - lv.offW, lv.W, p.mode);
                if (sc < 0 || sc >= lv.Wreal) continue;
                const float v = __ldg(xp + (long long)sr * lv.x_rs + sc);
                lo = fmaf(p.t.w_lo[jw], v, lo);
                hi = fmaf(p.t.w_hi[jw], v, hi);
            }
            ll = fmaf(p.t.h_lo[jh], lo, ll);
            lh = fmaf(p.t.h_hi[jh], lo, lh);
            hl = fmaf(p.t.h_lo[jh], hi, hl);
            hh = fmaf(p.t.h_hi[jh], hi, hh);
        }
        const size_t o = (size_t)i * lv.Wo + k;
        if (lv.st_low) lv.low[(long long)plane * lv.low_ps + (long long)i * lv.low_rs + k] = ll;
        if (lv.st_hi) {
            float* hip = lv.highs + (size_t)plane * 3 * band + o;
            hip[0] = fmaf(lh, lv.hi_scale, lv.hi_shift);
            hip[band] = fmaf(hl, lv.hi_scale, lv.hi_shift);
            hip[2 * band] = fmaf(hh, lv.hi_scale, lv.hi_shift);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// synthesis tile.  Works in "A-space": a = n + off, so that the polyphase split (which taps an output uses)
// depends only on the parity of the tile-local coordinate.  Output tile TH x TW.  W synthesis first (on the KH
// coefficient rows), then H synthesis; by separability this equals the reference's H-then-W order
// (pw/dwt/lowlevel.py:677-679) up to fp32 rounding.
// ------------------------------------------------------------------------------------------------
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

template <int L_, int TW_, int TH_, int NT_>
struct SfbOp {
    using Params = SfbParams;
    static constexpr const char* name = "chain_kernel<SfbOp>";
    static constexpr int L = L_, TW = TW_, TH = TH_, NT = NT_;
    static constexpr int H2 = L / 2;
    static constexpr int NV2 = (H2 + 2) / 2;            // float2 loads per band per W-synthesis item
    static constexpr int KWP = TW / 2 - 2 + 2 * NV2;    // staged coefficient columns (even, >= TW/2 + H2 - 1)
    static constexpr int KH = TH / 2 + H2 - 1;          // staged coefficient rows
    static constexpr int PB = KH * KWP;                 // one band's patch
    static constexpr int BUF = 4 * PB;                  // floats per staging buffer (LL, LH, HL, HH)
    static constexpr int NQ = TW / 4;                   // W-synthesis items per coefficient row (4 outputs each)
    static constexpr int QSTEP = NT / NQ;               // coefficient rows advanced per W-synthesis iteration
    static constexpr int CP = TW / 2;                   // output column pairs (H pass)
    static constexpr int NS = NT / CP;
    static constexpr int RS = TH / NS;                  // output rows per thread in the H pass (even)
    static constexpr int NR = RS / 2 + H2 - 1;          // coefficient rows feeding RS outputs
    static constexpr size_t smem = sizeof(float) * (size_t)(2 * BUF + 2 * KH * TW);  // 2 x 4 patches + u_lo/u_hi
    static_assert(L % 2 == 0 && TW % 4 == 0 && NT % CP == 0 && TH % NS == 0 && RS % 2 == 0 && NT % NQ == 0, "bad tile");
    static_assert(KWP >= TW / 2 + H2 - 1 && KWP % 2 == 0 && KWP <= NT, "bad KWP");

    static __device__ __forceinline__ void issue(const SfbParams& p, const TileRef& t, unsigned sub_s, int tid) {
        const SfbLevel& lv = p.lv[t.level];
        const int aW = lv.a0W + t.tw * TW, aH = lv.a0H + t.th * TH;
        const int kW0 = aW / 2 - (H2 - 1);
        const int kH0 = aH / 2 - (H2 - 1);
        const size_t band = (size_t)lv.h * lv.w;
        const float* lowp = lv.low + (long long)t.plane * lv.low_ps;
        const float* hip = lv.highs ? lv.highs + (size_t)t.plane * 3 * band : nullptr;
        // coefficient rows / columns feeding the valid outputs of an edge tile: a <= a_max -> k_local <= a_max/2 + H2-1
        const int amax_h = min(TH - 1, lv.offH + lv.out_h - 1 - aH);
        const int amax_w = min(TW - 1, lv.offW + lv.out_w - 1 - aW);
        const int nrows = min(KH, amax_h / 2 + H2);
        const int ncols = min(KWP, amax_w / 2 + H2);
        // kW0 is even whenever in_vec2 is set (checked on the host)
        if (lv.in_vec2)
            stage_synthesis<2, KH, KWP, NT>(sub_s, lowp, lv.low_rs, hip, band, kH0, kW0, lv.h, lv.w, p.periodic, nrows, ncols, tid);
        else
            stage_synthesis<1, KH, KWP, NT>(sub_s, lowp, lv.low_rs, hip, band, kH0, kW0, lv.h, lv.w, p.periodic, nrows, ncols, tid);
    }

    // W synthesis: each item makes four consecutive outputs (a = 4qq .. 4qq+3) of one coefficient row, for both
    // the h_lo branch (LL, HL) and the h_hi branch (LH, HH)
    static __device__ __forceinline__ void pass1(const SfbParams& p, const TileRef&, const float* sub, float* u, int tid) {
        const int qq = tid % NQ;
        int r = tid / NQ;
        const float* s0 = sub + r * KWP + 2 * qq;
        float* d = u + r * TW + 4 * qq;
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
                for (int uu = 0; uu < H2; ++uu) {
                    const int k = qo + H2 - 1 - uu;
                    a = fmaf(c[0][k], p.t.w_lo[par + 2 * uu], a);
                    a = fmaf(c[2][k], p.t.w_hi[par + 2 * uu], a);
                    b = fmaf(c[1][k], p.t.w_lo[par + 2 * uu], b);
                    b = fmaf(c[3][k], p.t.w_hi[par + 2 * uu], b);
                }
                lo[e] = a;
                hi[e] = b;
            }
            *reinterpret_cast<float4*>(d) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(d + KH * TW) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
    }

    // H synthesis: thread = two adjacent output columns, RS consecutive output rows
    static __device__ __forceinline__ void pass2(const SfbParams& p, const TileRef& t, const float* u, int tid) {
        const SfbLevel& lv = p.lv[t.level];
        const int cp = tid % CP;
        const int s = tid / CP;
        float2 vlo[NR], vhi[NR];
        const float2* plo = reinterpret_cast<const float2*>(u + (s * (RS / 2)) * TW + 2 * cp);
        const float2* phi = reinterpret_cast<const float2*>(u + KH * TW + (s * (RS / 2)) * TW + 2 * cp);
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
            for (int uu = 0; uu < H2; ++uu) {
                const int r = i - uu + H2 - 1;  // local coefficient row
                const float a0 = p.t.h_lo[2 * uu], b0 = p.t.h_hi[2 * uu], a1 = p.t.h_lo[2 * uu + 1], b1 = p.t.h_hi[2 * uu + 1];
                ye.x = fmaf(vlo[r].x, a0, ye.x); ye.y = fmaf(vlo[r].y, a0, ye.y);
                ye.x = fmaf(vhi[r].x, b0, ye.x); ye.y = fmaf(vhi[r].y, b0, ye.y);
                yo.x = fmaf(vlo[r].x, a1, yo.x); yo.y = fmaf(vlo[r].y, a1, yo.y);
                yo.x = fmaf(vhi[r].x, b1, yo.x); yo.y = fmaf(vhi[r].y, b1, yo.y);
            }
            y[2 * i] = ye;
            y[2 * i + 1] = yo;
        }
        const int out_h = lv.out_h, out_w = lv.out_w;
        const int nW = lv.a0W + t.tw * TW + 2 * cp - lv.offW;
        const int nH = lv.a0H + t.th * TH + s * RS - lv.offH;
        const long long yrs = lv.y_rs;
        float* q = lv.y + (long long)t.plane * lv.y_ps + (long long)nH * yrs + nW;  // only dereferenced where valid
        if (lv.out_vec2 && nH >= 0 && nH + RS <= out_h && nW >= 0 && nW + 1 < out_w) {
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                *reinterpret_cast<float2*>(q) = y[i];
                q += yrs;
            }
        } else {
            const bool ok0 = nW >= 0 && nW < out_w;
            const bool ok1 = nW + 1 >= 0 && nW + 1 < out_w;
#pragma unroll
            for (int i = 0; i < RS; ++i) {
                const int row = nH + i;
                if (row >= 0 && row < out_h) {
                    if (ok0) q[0] = y[i].x;
                    if (ok1) q[1] = y[i].y;
                }
                q += yrs;
            }
        }
    }
};

// synthesis, direct: one thread per output sample of ONE level; any tap count.
struct SfbDirectParams {
    SfbLevel lv;
    int planes, periodic, Lw, Lh;
    Taps t;
};

__global__ void __launch_bounds__(kThreads) sfb2d_direct_kernel(const __grid_constant__ SfbDirectParams p) {
    const SfbLevel& lv = p.lv;
    const size_t total = (size_t)p.planes * lv.out_h * lv.out_w;
    const size_t band = (size_t)lv.h * lv.w;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int nW = (int)(idx % lv.out_w);
        const int nH = (int)((idx / lv.out_w) % lv.out_h);
        const int plane = (int)(idx / ((size_t)lv.out_h * lv.out_w));
        const float* __restrict__ lowp = lv.low + (long long)plane * lv.low_ps;
        const float* __restrict__ hip = lv.highs ? lv.highs + (size_t)plane * 3 * band : nullptr;
        const int AH = nH + lv.offH, AW = nW + lv.offW;
        float y = 0.f;
        for (int tH = AH & 1; tH < p.Lh; tH += 2) {
            const int kr = coef_index((AH - tH) / 2, lv.h, p.periodic);
            if (kr < 0) continue;
            float lo = 0.f, hi = 0.f;  // W-synthesised values to be combined with h_lo / h_hi
            for (int tW = AW & 1; tW < p.Lw; tW += 2) {
                const int kc = coef_index((AW - tW) / 2, lv.w, p.periodic);
                if (kc < 0) continue;
                const float ll = __ldg(lowp + (long long)kr * lv.low_rs + kc);
                lo = fmaf(ll, p.t.w_lo[tW], lo);
                if (hip) {
                    const float* q = hip + (size_t)kr * lv.w + kc;
                    hi = fmaf(__ldg(q), p.t.w_lo[tW], hi);               // LH
                    lo = fmaf(__ldg(q + band), p.t.w_hi[tW], lo);        // HL
                    hi = fmaf(__ldg(q + 2 * band), p.t.w_hi[tW], hi);    // HH
                }
            }
            y = fmaf(lo, p.t.h_lo[tH], y);
            y = fmaf(hi, p.t.h_hi[tH], y);
        }
        lv.y[(long long)plane * lv.y_ps + (long long)nH * lv.y_rs + nW] = y;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// B200W_FORCE_DIRECT=1: one-thread-per-output kernels; B200W_FORCE_TILED=1: shared-memory tile kernels instead of
// the streaming kernels (both are on-device cross-checks of the default path, used by the tests)
static int env_flag(const char* name) {
    const char* e = getenv(name);
    return (e && e[0] == '1') ? 1 : 0;
}
static bool force_direct() {
    static int v = -1;
    if (v < 0) v = env_flag("B200W_FORCE_DIRECT");
    return v == 1;
}
static bool force_tiled() {
    static int v = -1;
    if (v < 0) v = env_flag("B200W_FORCE_TILED");
    return v == 1;
}
// B200W_OWNER=0 switches the owner kernels off (chains of small planes then take the ticketed chain kernels);
// B200W_OWNER_J0=n makes them start no earlier than level n (the levels before run as a chain launch).
// B200W_OWNER=2 uses them whenever the shapes fit, however few planes there are (tests).
static int owner_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_OWNER");
        v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return v;
}
static int owner_j0_min() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_OWNER_J0");
        v = (e && *e) ? atoi(e) : 0;
        if (v < 0) v = 0;
    }
    return v;
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
    for (int i = 0; i < kMaxTemplTaps; ++i) {
        t.h_lo2[i] = make_float2(t.h_lo[i], t.h_lo[i]);
        t.h_hi2[i] = make_float2(t.h_hi[i], t.h_hi[i]);
    }
    return B200W_OK;
}

static int coeff_len(int n, int l, int mode) { return mode == B200W_MODE_PERIODIZATION ? (n + 1) / 2 : (n + l - 1) / 2; }
static int idwt_len(int m, int l, int mode) { return mode == B200W_MODE_PERIODIZATION ? 2 * m : 2 * m - l + 2; }

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

// Workspace layout of a J-level chain: [ticket, done[J][planes]] (u32), then the intermediate low-pass images
// (planes x rows x pitch floats, pitch = columns rounded up to 4 so that every row is 16-byte aligned), each
// region rounded up to 256 bytes.  The caller's pointer must be 256-byte aligned (torch allocations are).
static size_t round256(size_t n) { return (n + 255) & ~(size_t)255; }
static size_t sync_bytes(int planes, int J) { return J > 1 ? round256(sizeof(unsigned) * ((size_t)J * planes + 1)) : 0; }
static int pitch4(int w) { return (w + 3) & ~3; }
static size_t scratch_bytes(int planes, int rows, int cols) {
    return round256(sizeof(float) * (size_t)planes * rows * pitch4(cols));
}

// Persistent launch of a tile chain: grid = min(tiles, SMs x occupancy).  A multi-level chain spins on completion
// counters, so its CTAs must all be resident: it is launched cooperatively (the runtime refuses the launch
// otherwise) and the counters are cleared on the stream first.  `occ_cache` is the caller's per-kernel,
// per-device cache of the occupancy (0 = not queried yet; the query also raises the kernel's dynamic
// shared-memory limit).  Both are immutable facts about (kernel, device), so caching them keeps the library
// re-entrant.
template <class Op>
static int launch_chain(typename Op::Params& p, int* occ_cache, cudaStream_t st) {
    auto kernel = chain_kernel<Op>;
    const DeviceInfo di = device_info();
    int& occ = occ_cache[di.dev];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Op::smem);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, Op::NT, Op::smem);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        occ = o > 0 ? o : 1;
    }
    long long grid = (long long)di.sms * occ;
    if (grid > p.total) grid = p.total;
    cudaError_t e;
    if (p.J > 1) {
        const int rc = zero_sync_words(p.ticket, (size_t)p.J * p.planes + 1, st);
        if (rc) return rc;
        void* args[] = {(void*)&p};
        e = cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3(Op::NT), args, Op::smem, st);
    } else {
        kernel<<<(unsigned)grid, Op::NT, Op::smem, st>>>(p);
        e = cudaGetLastError();
    }
    note_launch(Op::name);
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <int TW, int TH>
static void afb_layout(AfbParams& p) {
    long long base = 0;
    for (int j = 0; j < p.J; ++j) {
        p.lv[j].tiles_w = ceil_div(p.lv[j].Wo, TW);
        p.lv[j].tiles_h = ceil_div(p.lv[j].Ho, TH);
        p.lv[j].tile_base = base;
        base += (long long)p.lv[j].tiles_w * p.lv[j].tiles_h * p.planes;
    }
    p.total = base;
}

template <int L>
static int launch_afb_chain(AfbParams& p, cudaStream_t st) {
    // two tile shapes: 32x32 outputs / 128 threads, or 64 wide x 32 / 256 threads (less halo); take the wide
    // one for wide single-level problems unless it wastes noticeably more of the padded output area
    const int Wo = p.lv[0].Wo;
    const long long a32 = (long long)ceil_div(Wo, 32) * 32;
    const long long a64 = (long long)ceil_div(Wo, 64) * 64;
    if (p.J == 1 && Wo >= 128 && a64 * 100 <= a32 * 103) {
        using Op = AfbOp<L, 64, 32, 256>;
        static int occ[kMaxDevices] = {0};
        afb_layout<Op::TW, Op::TH>(p);
        return launch_chain<Op>(p, occ, st);
    }
    using Op = AfbOp<L, 32, 32, 128>;
    static int occ[kMaxDevices] = {0};
    afb_layout<Op::TW, Op::TH>(p);
    return launch_chain<Op>(p, occ, st);
}

template <int L>
static int launch_sfb_chain(SfbParams& p, cudaStream_t st) {
    using Op = SfbOp<L, 64, 64, 256>;
    static int occ[kMaxDevices] = {0};
    long long base = 0;
    for (int j = 0; j < p.J; ++j) {
        SfbLevel& lv = p.lv[j];
        lv.a0W = lv.offW & ~1;
        lv.a0H = lv.offH & ~1;
        lv.tiles_w = ceil_div(lv.offW + lv.out_w - lv.a0W, Op::TW);
        lv.tiles_h = ceil_div(lv.offH + lv.out_h - lv.a0H, Op::TH);
        // tile 0 starts at coefficient column a0W/2 - (L/2-1): even for every non-periodization mode
        const int kW0 = lv.a0W / 2 - (L / 2 - 1);
        if ((kW0 & 1) != 0) lv.in_vec2 = 0;
        if ((lv.offW & 1) != 0) lv.out_vec2 = 0;
        lv.tile_base = base;
        base += (long long)lv.tiles_w * lv.tiles_h * p.planes;
    }
    p.total = base;
    return launch_chain<Op>(p, occ, st);
}

static size_t direct_grid(size_t total) {
    size_t g = (total + kThreads - 1) / kThreads;
    const size_t cap = 148 * 16;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

static bool templated_taps(int Lw, int Lh) {
    return Lw == Lh && (Lw % 2) == 0 && Lw >= 2 && Lw <= 16 && !force_direct();
}

// ---- analysis chain --------------------------------------------------------------------------------
static int run_afb_big_levels(AfbParams& p, int L, cudaStream_t st) {
    if (!force_tiled() && afb_stream_supported(p, L)) return launch_afb_stream(p, L, device_info().sms, st);
    switch (L) {
        case 2: return launch_afb_chain<2>(p, st);
        case 4: return launch_afb_chain<4>(p, st);
        case 6: return launch_afb_chain<6>(p, st);
        case 8: return launch_afb_chain<8>(p, st);
        case 10: return launch_afb_chain<10>(p, st);
        case 12: return launch_afb_chain<12>(p, st);
        case 14: return launch_afb_chain<14>(p, st);
        default: return launch_afb_chain<16>(p, st);
    }
}

// level sizes of the analysis chain; returns a status
static int afb_dims(int H, int W, int Lw, int Lh, int mode, int J, const int* pad_hw, int* Ho, int* Wo) {
    int h = H, w = W;
    for (int j = 0; j < J; ++j) {
        const int ph = (pad_hw && j > 0) ? pad_hw[2 * j] : 0, pw = (pad_hw && j > 0) ? pad_hw[2 * j + 1] : 0;
        if (ph < 0 || ph > 1 || pw < 0 || pw > 1) return B200W_ERR_BAD_SHAPE;
        h = coeff_len(h + ph, Lh, mode);
        w = coeff_len(w + pw, Lw, mode);
        if (h < 1 || w < 1) return B200W_ERR_BAD_SHAPE;
        Ho[j] = h;
        Wo[j] = w;
    }
    return B200W_OK;
}

static size_t afb_workspace_bytes(int planes, int H, int W, int Lw, int Lh, int mode, int J, const int* pad_hw) {
    if (planes < 1 || H < 1 || W < 1 || J < 1 || J > kMaxLevels || Lw < 1 || Lh < 1 || !mode_supported(mode)) return 0;
    int Ho[kMaxLevels], Wo[kMaxLevels];
    if (afb_dims(H, W, Lw, Lh, mode, J, pad_hw, Ho, Wo)) return 0;
    size_t n = sync_bytes(planes, J);
    for (int j = 0; j + 1 < J; ++j) n += scratch_bytes(planes, Ho[j], Wo[j]);
    return n;
}

static int run_afb_chain(const float* x, int64_t x_ps, int64_t x_rs, int planes, int H, int W, const float* w_lo,
                         const float* w_hi, int Lw, const float* h_lo, const float* h_hi, int Lh, int mode, int J,
                         const int* pad_hw, float* yl, float* const* highs, void* workspace, size_t workspace_bytes,
                         cudaStream_t st, float hi_scale = 1.f, float hi_shift = 0.f, bool allow_skip = false) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (J < 1 || J > kMaxLevels) return B200W_ERR_BAD_SHAPE;
    // the filter_wavelet entry point (single level) may leave out the low-pass or the detail output
    if (!x || !highs || (!allow_skip && !yl) || (allow_skip && !yl && !highs[0])) return B200W_ERR_NULL_POINTER;
    const bool plain = !allow_skip && hi_scale == 1.f && hi_shift == 0.f;
    if (planes < 1 || H < 1 || W < 1) return B200W_ERR_BAD_SHAPE;
    AfbParams p;
    int rc = fill_taps(p.t, w_lo, w_hi, Lw, h_lo, h_hi, Lh);
    if (rc) return rc;
    int Ho[kMaxLevels], Wo[kMaxLevels];
    if ((rc = afb_dims(H, W, Lw, Lh, mode, J, pad_hw, Ho, Wo))) return rc;
    if (J > 1) {
        if (!workspace || workspace_bytes < afb_workspace_bytes(planes, H, W, Lw, Lh, mode, J, pad_hw) ||
            !aligned_to(workspace, 256))
            return B200W_ERR_WORKSPACE;
    }
    p.J = J;
    p.planes = planes;
    p.mode = mode;
    p.ticket = J > 1 ? (unsigned*)workspace : nullptr;
    p.done = J > 1 ? p.ticket + 1 : nullptr;
    char* scratch = (char*)workspace + sync_bytes(planes, J);
    int h = H, w = W;  // real size of the level input
    for (int j = 0; j < J; ++j) {
        AfbLevel& lv = p.lv[j];
        if (!highs[j] && !allow_skip) return B200W_ERR_NULL_POINTER;
        const int ph = (pad_hw && j > 0) ? pad_hw[2 * j] : 0, pw = (pad_hw && j > 0) ? pad_hw[2 * j + 1] : 0;
        lv.st_low = (j < J - 1 || yl != nullptr) ? 1 : 0;
        lv.st_hi = highs[j] != nullptr ? 1 : 0;
        lv.hi_scale = hi_scale;
        lv.hi_shift = hi_shift;
        if (j == 0) {
            lv.x = x;
            lv.x_ps = x_ps;
            lv.x_rs = x_rs;
        } else {  // the previous level's low-pass image in the workspace
            lv.x = p.lv[j - 1].low;
            lv.x_ps = p.lv[j - 1].low_ps;
            lv.x_rs = p.lv[j - 1].low_rs;
        }
        lv.Hreal = h;
        lv.Wreal = w;
        lv.H = h + ph;
        lv.W = w + pw;
        if ((rc = analysis_offset(lv.W, Lw, mode, &lv.offW))) return rc;
        if ((rc = analysis_offset(lv.H, Lh, mode, &lv.offH))) return rc;
        lv.Ho = Ho[j];
        lv.Wo = Wo[j];
        if (j == J - 1) {
            lv.low = yl;
            lv.low_rs = lv.Wo;
            lv.low_ps = (long long)lv.Ho * lv.Wo;
        } else {
            lv.low = (float*)scratch;
            lv.low_rs = pitch4(lv.Wo);
            lv.low_ps = (long long)lv.Ho * lv.low_rs;
            scratch += scratch_bytes(planes, lv.Ho, lv.Wo);
        }
        lv.highs = highs[j];
        // staging vector width of the tile kernel: the first staged column of every tile is 2*TW*tw - offW
        lv.in_vec = 1;
        if ((lv.offW % 2) == 0 && (lv.x_rs % 2) == 0 && (lv.x_ps % 2) == 0 && aligned_to(lv.x, 8)) lv.in_vec = 2;
        if (lv.in_vec == 2 && (lv.offW % 4) == 0 && (lv.x_rs % 4) == 0 && (lv.x_ps % 4) == 0 && aligned_to(lv.x, 16))
            lv.in_vec = 4;
        lv.out_vec2 = ((lv.Wo % 2) == 0 && lv.highs && aligned_to(lv.highs, 8)) ? 1 : 0;
        lv.low_vec2 = ((lv.low_rs % 2) == 0 && (lv.low_ps % 2) == 0 && lv.low && aligned_to(lv.low, 8)) ? 1 : 0;
        lv.tile_base = lv.cta_base = 0;
        lv.tiles_h = lv.tiles_w = lv.R = lv.ncp = lv.cpp = lv.cp0A = lv.ncpA = lv.itemsA = lv.cppA = lv.RB = lv.itemsB = 0;
        h = lv.Ho;
        w = lv.Wo;
    }
    // the store epilogue lives in the stream and the direct kernels: other inputs (unaligned rows) take the direct one
    if (templated_taps(Lw, Lh) && (plain || (!force_tiled() && afb_stream_supported(p, Lw)))) {
        // chains of small planes: one CTA owns a part of a plane for all levels (TMA-staged owner kernel first, the
        // cp.async owner kernel for the shapes it declines); everything else goes through the ticketed stream chain
        // (or the tile chain when rows are unaligned)
        if (J > 1 && plain && !force_tiled() && owner_mode() != 0) {
            AfbTmaParams tp;
            if (afb_tma_plan(p, Lw, device_info().sms, owner_mode() == 2, tp)) return launch_afb_tma(tp, Lw, st);
        }
        if (J > 1 && !force_tiled() && owner_mode() != 0) {
            AfbOwnerParams op;
            if (afb_owner_plan(p, Lw, device_info().sms, owner_j0_min(), owner_mode() == 2, op)) {
                if (op.j0 > 0) {
                    AfbParams head = p;
                    head.J = op.j0;
                    rc = run_afb_big_levels(head, Lw, st);
                    if (rc) return rc;
                }
                return launch_afb_owner(op, Lw, st);
            }
        }
        return run_afb_big_levels(p, Lw, st);
    }
    for (int j = 0; j < J; ++j) {  // level by level with the direct kernel
        AfbDirectParams d;
        d.lv = p.lv[j];
        d.planes = planes;
        d.mode = mode;
        d.Lw = Lw;
        d.Lh = Lh;
        d.t = p.t;
        afb2d_direct_kernel<<<(unsigned)direct_grid((size_t)planes * d.lv.Ho * d.lv.Wo), kThreads, 0, st>>>(d);
        note_launch("afb2d_direct_kernel");
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_last_cuda_error(e);
    }
    return B200W_OK;
}

// ---- synthesis chain -------------------------------------------------------------------------------
static int run_sfb_big_levels(SfbParams& p, int L, cudaStream_t st) {
    if (!force_tiled() && sfb_stream_supported(p, L)) return launch_sfb_stream(p, L, device_info().sms, st);
    switch (L) {
        case 2: return launch_sfb_chain<2>(p, st);
        case 4: return launch_sfb_chain<4>(p, st);
        case 6: return launch_sfb_chain<6>(p, st);
        case 8: return launch_sfb_chain<8>(p, st);
        case 10: return launch_sfb_chain<10>(p, st);
        case 12: return launch_sfb_chain<12>(p, st);
        case 14: return launch_sfb_chain<14>(p, st);
        default: return launch_sfb_chain<16>(p, st);
    }
}

static int sfb_check_dims(int planes, const int* hs, const int* ws, int Lw, int Lh, int mode, int J, const int* out_hs,
                          const int* out_ws) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (J < 1 || J > kMaxLevels) return B200W_ERR_BAD_SHAPE;
    if (!hs || !ws || !out_hs || !out_ws) return B200W_ERR_NULL_POINTER;
    if (planes < 1 || Lw < 1 || Lh < 1) return B200W_ERR_BAD_SHAPE;
    const bool per = mode == B200W_MODE_PERIODIZATION;
    for (int j = 0; j < J; ++j) {
        if (hs[j] < 1 || ws[j] < 1 || out_hs[j] < 1 || out_ws[j] < 1) return B200W_ERR_BAD_SHAPE;
        if (per && (2 * hs[j] < Lh || 2 * ws[j] < Lw)) return B200W_ERR_PER_TOO_SHORT;
        if (out_hs[j] > idwt_len(hs[j], Lh, mode) || out_ws[j] > idwt_len(ws[j], Lw, mode)) return B200W_ERR_BAD_SHAPE;
        // level j < J-1 reads the top-left h[j] x w[j] block of the previous output ('unpad')
        if (j + 1 < J && (out_hs[j + 1] < hs[j] || out_ws[j + 1] < ws[j])) return B200W_ERR_BAD_SHAPE;
    }
    return B200W_OK;
}

static size_t sfb_workspace_bytes(int planes, int J, const int* out_hs, const int* out_ws) {
    size_t n = sync_bytes(planes, J);
    for (int j = 1; j < J; ++j) n += scratch_bytes(planes, out_hs[j], out_ws[j]);
    return n;
}

// levels are given finest first (index j like yh[j]); the chain runs j = J-1 .. 0
static int run_sfb_chain(const float* yl, int64_t yl_ps, int64_t yl_rs, const float* const* highs, int planes,
                         const int* hs, const int* ws, const float* w_lo, const float* w_hi, int Lw, const float* h_lo,
                         const float* h_hi, int Lh, int mode, int J, const int* out_hs, const int* out_ws, float* y,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
    int rc = sfb_check_dims(planes, hs, ws, Lw, Lh, mode, J, out_hs, out_ws);
    if (rc) return rc;
    if (!yl || !y) return B200W_ERR_NULL_POINTER;
    SfbParams p;
    if ((rc = fill_taps(p.t, w_lo, w_hi, Lw, h_lo, h_hi, Lh))) return rc;
    if (J > 1) {
        if (!workspace || workspace_bytes < sfb_workspace_bytes(planes, J, out_hs, out_ws) || !aligned_to(workspace, 256))
            return B200W_ERR_WORKSPACE;
    }
    const bool per = mode == B200W_MODE_PERIODIZATION;
    p.J = J;
    p.planes = planes;
    p.periodic = per ? 1 : 0;
    p.ticket = J > 1 ? (unsigned*)workspace : nullptr;
    p.done = J > 1 ? p.ticket + 1 : nullptr;
    char* scratch = (char*)workspace + sync_bytes(planes, J);
    for (int c = 0; c < J; ++c) {  // chain position c handles level j = J-1-c
        const int j = J - 1 - c;
        SfbLevel& lv = p.lv[c];
        if (c == 0) {
            lv.low = yl;
            lv.low_ps = yl_ps;
            lv.low_rs = yl_rs;
        } else {  // the previous chain output, of which the top-left h x w block is used ('unpad')
            lv.low = p.lv[c - 1].y;
            lv.low_ps = p.lv[c - 1].y_ps;
            lv.low_rs = p.lv[c - 1].y_rs;
        }
        lv.highs = highs ? highs[j] : nullptr;
        lv.h = hs[j];
        lv.w = ws[j];
        lv.out_h = out_hs[j];
        lv.out_w = out_ws[j];
        if (j == 0) {
            lv.y = y;
            lv.y_rs = lv.out_w;
            lv.y_ps = (long long)lv.out_h * lv.out_w;
        } else {
            lv.y = (float*)scratch;
            lv.y_rs = pitch4(lv.out_w);
            lv.y_ps = (long long)lv.out_h * lv.y_rs;
            scratch += scratch_bytes(planes, lv.out_h, lv.out_w);
        }
        lv.offW = per ? Lw / 2 - 1 : Lw - 2;
        lv.offH = per ? Lh / 2 - 1 : Lh - 2;
        lv.a0W = lv.a0H = 0;
        lv.in_vec2 = ((lv.w % 2) == 0 && (lv.low_rs % 2) == 0 && (lv.low_ps % 2) == 0 && aligned_to(lv.low, 8) &&
                      (!lv.highs || aligned_to(lv.highs, 8))) ? 1 : 0;
        lv.out_vec2 = ((lv.y_rs % 2) == 0 && (lv.y_ps % 2) == 0 && aligned_to(lv.y, 8)) ? 1 : 0;
        lv.tile_base = lv.cta_base = 0;
        lv.tiles_h = lv.tiles_w = lv.Rp = lv.cpp = lv.tA0 = lv.ntA = lv.itemsA = lv.cppA = 0;
        lv.nA0 = lv.nA1 = lv.itemsB = lv.n0_off = lv.kb_off = lv.m_lo = lv.vec2 = 0;
        lv.y_vec = 1;
    }
    if (templated_taps(Lw, Lh)) {
        if (J > 1 && !force_tiled() && owner_mode() != 0) {
            SfbTmaParams tp;
            if (sfb_tma_plan(p, Lw, device_info().sms, owner_mode() == 2, tp)) return launch_sfb_tma(tp, Lw, st);
        }
        if (J > 1 && !force_tiled() && owner_mode() != 0) {
            SfbOwnerParams op;
            if (sfb_owner_plan(p, Lw, device_info().sms, owner_mode() == 2, op)) return launch_sfb_owner(op, Lw, st);
        }
        return run_sfb_big_levels(p, Lw, st);
    }
    for (int c = 0; c < J; ++c) {
        SfbDirectParams d;
        d.lv = p.lv[c];
        d.planes = planes;
        d.periodic = p.periodic;
        d.Lw = Lw;
        d.Lh = Lh;
        d.t = p.t;
        sfb2d_direct_kernel<<<(unsigned)direct_grid((size_t)planes * d.lv.out_h * d.lv.out_w), kThreads, 0, st>>>(d);
        note_launch("sfb2d_direct_kernel");
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_last_cuda_error(e);
    }
    return B200W_OK;
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
    return idwt_len(m, l, mode);
}

extern "C" size_t b200w_dwt2_workspace_bytes(int planes, int H, int W, int Lw, int Lh, int mode, int J,
                                             const int* pad_hw) {
    return afb_workspace_bytes(planes, H, W, Lw, Lh, mode, J, pad_hw);
}

extern "C" size_t b200w_idwt2_workspace_bytes(int planes, int J, const int* out_h, const int* out_w) {
    if (planes < 1 || J < 1 || J > kMaxLevels || !out_h || !out_w) return 0;
    for (int j = 0; j < J; ++j)
        if (out_h[j] < 1 || out_w[j] < 1) return 0;
    return sfb_workspace_bytes(planes, J, out_h, out_w);
}

extern "C" int b200w_afb2d_f32(const float* x, int64_t x_plane_stride, int64_t x_row_stride, int planes, int H,
                               int W, const float* w_lo, const float* w_hi, int Lw, const float* h_lo,
                               const float* h_hi, int Lh, int mode, float* low, float* highs, void* stream) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (!low || !highs) return B200W_ERR_NULL_POINTER;
    float* his[1] = {highs};
    return run_afb_chain(x, x_plane_stride, x_row_stride, planes, H, W, w_lo, w_hi, Lw, h_lo, h_hi, Lh, mode, 1,
                         nullptr, low, his, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int b200w_afb2d_ex_f32(const float* x, int64_t x_plane_stride, int64_t x_row_stride, int planes, int H,
                                  int W, const float* w_lo, const float* w_hi, int Lw, const float* h_lo,
                                  const float* h_hi, int Lh, int mode, float* low, float* highs, float hi_scale,
                                  float hi_shift, void* stream) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    float* his[1] = {highs};
    return run_afb_chain(x, x_plane_stride, x_row_stride, planes, H, W, w_lo, w_hi, Lw, h_lo, h_hi, Lh, mode, 1,
                         nullptr, low, his, nullptr, 0, (cudaStream_t)stream, hi_scale, hi_shift, true);
}

extern "C" int b200w_dwt2_f32(const float* x, int64_t x_plane_stride, int64_t x_row_stride, int planes, int H, int W,
                              const float* w_lo, const float* w_hi, int Lw, const float* h_lo, const float* h_hi,
                              int Lh, int mode, int J, const int* pad_hw, float* yl, float* const* highs,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return run_afb_chain(x, x_plane_stride, x_row_stride, planes, H, W, w_lo, w_hi, Lw, h_lo, h_hi, Lh, mode, J,
                         pad_hw, yl, highs, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int b200w_sfb2d_f32(const float* low, int64_t low_plane_stride, int64_t low_row_stride,
                               const float* highs, int planes, int h, int w, const float* w_lo, const float* w_hi,
                               int Lw, const float* h_lo, const float* h_hi, int Lh, int mode, float* y, int out_h,
                               int out_w, void* stream) {
    if (!mode_supported(mode)) return B200W_ERR_BAD_MODE;
    if (!low || !y) return B200W_ERR_NULL_POINTER;
    if (planes < 1 || h < 1 || w < 1 || out_h < 1 || out_w < 1) return B200W_ERR_BAD_SHAPE;
    const float* his[1] = {highs};
    return run_sfb_chain(low, low_plane_stride, low_row_stride, his, planes, &h, &w, w_lo, w_hi, Lw, h_lo, h_hi, Lh,
                         mode, 1, &out_h, &out_w, y, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int b200w_idwt2_f32(const float* yl, int64_t yl_plane_stride, int64_t yl_row_stride,
                               const float* const* highs, int planes, const int* h, const int* w, const float* w_lo,
                               const float* w_hi, int Lw, const float* h_lo, const float* h_hi, int Lh, int mode,
                               int J, const int* out_h, const int* out_w, float* y, void* workspace,
                               size_t workspace_bytes, void* stream) {
    return run_sfb_chain(yl, yl_plane_stride, yl_row_stride, highs, planes, h, w, w_lo, w_hi, Lw, h_lo, h_hi, Lh, mode,
                         J, out_h, out_w, y, workspace, workspace_bytes, (cudaStream_t)stream);
}
