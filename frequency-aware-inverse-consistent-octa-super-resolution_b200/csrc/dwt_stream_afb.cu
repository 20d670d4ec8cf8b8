// Analysis filter bank, streaming register-blocked kernel (the fast path for 16-byte aligned rows).
//
// Replaces afb1d(dim=3) + afb1d(dim=2) + reshape + 2x .contiguous() of AFB2D.forward
// (pw/dwt/lowlevel.py:336-347, 91-172) -- and the J-level loop around it (pw/dwt/transform2d.py:66-74).
//
// A thread owns ONE PAIR of adjacent output columns and marches down R output rows of one image plane:
//   * row pass in registers: per input row the 4*NV floats its two outputs need give lo/hi for both columns,
//     4L FMAs with the taps as constant-bank operands,
//   * column pass "accumulate forward": each pair of input rows is scattered into the L/2 output rows it
//     contributes to (8 accumulators each); the oldest one is complete, is stored with 64-bit coalesced stores
//     (LL into `low`, LH/HL/HH straight into `highs[:, :, 0..2]`) and its registers are recycled.
// ~2L FMAs + a handful of memory instructions per input pixel, no block-wide barriers.  The loop is unrolled by
// L/2 row pairs so that the accumulator ring has static register names.
//
// Input staging.  Column pairs whose window lies inside the image ("interior": all but 2..8 per row) are served by
// a per-WARP ring in shared memory: every lane cp.async-copies its own 16 bytes of each input row (a warp = 512
// contiguous bytes per row; the last lane of a run adds the NV-1 trailing float4), D row pairs deep, so ~D KB per
// warp are in flight without holding registers; a lane then reads its window (own float4 + the neighbours')
// with NV 128-bit shared loads.  Only __syncwarp is needed.  Rows outside the image are remapped per row
// (symmetric / reflect / periodic / periodization) or cleared with shared stores.
// The few border columns of each row run in separate CTAs, one thread per output position, through the row /
// column maps of the padding mode.
// Segments overlap by L-2 input rows (re-read through L2).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include "dwt_levels.cuh"

namespace b200w {

constexpr int afb_off(int L, bool per) { return per ? L - 1 - L / 2 : L - 2; }
// the first loaded column is 4*cp - (off + S): S pads the window start down to a multiple of 4
constexpr int afb_shift(int L, bool per) { return (4 - afb_off(L, per) % 4) % 4; }

// ring geometry: per warp, D stages of one input row pair; a row holds the 32 lanes' own float4 plus NV-1 extra
// float4 for each run of lanes that sits on one image row (a warp may straddle up to kMaxRuns segment rows)
constexpr int kMaxRuns = 5;
constexpr int kMinColPairs = 8;   // interior pairs per row needed for the ring path (=> at most kMaxRuns runs)
template <int L, int S>
struct AfbStreamCfg {
    static constexpr int H2 = L / 2;
    static constexpr int NV = (S + L + 2 + 3) / 4;   // float4 per window
    static constexpr int NE = 4 * NV;
    static constexpr int RP = 32 + kMaxRuns * (NV - 1);   // float4 per ring row
    static constexpr int D = L <= 8 ? 6 : 4;         // ring depth (row pairs)
    static constexpr int STAGE = 2 * RP;             // float4 per stage
    // resident CTAs per SM the ring path is compiled for (caps the registers; the rare border path may spill)
    static constexpr int MINB = L <= 8 ? 5 : (L <= 12 ? 4 : 3);
    static constexpr size_t smem = sizeof(float4) * (size_t)(kStreamNT / 32) * D * STAGE;
};

// source row of input row r: r itself inside the image, else the padding mode's map; -1 = zero row
__device__ __forceinline__ int afb_src_row(int r, int H, int Hreal, int mode) {
    if ((unsigned)r < (unsigned)Hreal) return r;
    const int sr = ext_index_far(r, H, mode);
    return sr >= Hreal ? -1 : sr;
}

// one pair of input rows: row pass for both, then scatter into the accumulator ring (ph = pair index mod L/2).
// The ring keeps each sub-band's two columns as one float2, so the column pass runs on the packed FFMA2: the two
// columns share the tap, which comes as a (t, t) pair from uniform registers.
// With kRotate the ring is kept in age order instead (slot 0 = oldest) and shifted by the caller after the store:
// L/2-1 register moves per row pair, but the loop needs no unrolling by L/2 -- for long filters the unrolled
// body would not fit the instruction cache.
template <int L, int S, int NE>
__device__ __forceinline__ void afb_pair_fma(const Taps& t, const float (&v)[2][NE], float2 (&acc)[L / 2][4], int ph) {
    constexpr int H2 = L / 2;
    constexpr bool kRotate = L >= 10;
    float2 rl[2], rh[2];   // [row of the pair] = (column 0, column 1)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) {
            lo0 = fmaf(t.w_lo[j], v[e][S + j], lo0);
            hi0 = fmaf(t.w_hi[j], v[e][S + j], hi0);
            lo1 = fmaf(t.w_lo[j], v[e][S + j + 2], lo1);
            hi1 = fmaf(t.w_hi[j], v[e][S + j + 2], hi1);
        }
        rl[e] = make_float2(lo0, lo1);
        rh[e] = make_float2(hi0, hi1);
    }
    // this pair carries taps (2u, 2u+1) of output row q - u
#pragma unroll
    for (int u = 0; u < H2; ++u) {
        const int sl = kRotate ? H2 - 1 - u : (ph - u + H2) % H2;
        const float2 a = t.h_lo2[2 * u], b = t.h_hi2[2 * u];
        const float2 c = t.h_lo2[2 * u + 1], d = t.h_hi2[2 * u + 1];
        float2* s = acc[sl];
        if (u == 0) {   // first contribution: start the accumulators
            s[0] = fmul2(a, rl[0]);   // LL: W-lo, H-lo
            s[1] = fmul2(b, rl[0]);   // LH: W-lo, H-hi
            s[2] = fmul2(a, rh[0]);   // HL: W-hi, H-lo
            s[3] = fmul2(b, rh[0]);   // HH
        } else {
            s[0] = ffma2(a, rl[0], s[0]);
            s[1] = ffma2(b, rl[0], s[1]);
            s[2] = ffma2(a, rh[0], s[2]);
            s[3] = ffma2(b, rh[0], s[3]);
        }
        s[0] = ffma2(c, rl[1], s[0]);
        s[1] = ffma2(d, rl[1], s[1]);
        s[2] = ffma2(c, rh[1], s[2]);
        s[3] = ffma2(d, rh[1], s[3]);
    }
}

#ifdef B200W_TIMELINE
// debug build only (B200W_TIMELINE=1 python -m b200wave._build): per-CTA timestamps of the last chained launch
__device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TL_MARK(slot) do { if (g_timeline && threadIdx.x == 0) g_timeline[(size_t)item * 16 + (slot)] = ((slot) == 0 || (slot) == 5) ? gtime() : (unsigned long long)clock64(); } while (0)
#else
#define TL_MARK(slot) do { } while (0)
#endif

// block until every CTA item of the previous level of this plane has been published (called by all threads,
// after their index arithmetic so that the set-up overlaps the wait)
__device__ __forceinline__ void chain_wait(const unsigned* ctr, unsigned need) {
    if (ctr != nullptr) {
        if (threadIdx.x == 0)
            while (ld_acquire_u32(ctr) < need) __nanosleep(20);
        __syncthreads();
    }
}

// ---- interior column pairs: per-warp cp.async ring ---------------------------------------------------
template <int L, int S>
__device__ __forceinline__ void afb_ring_cta(const AfbParams& p, const AfbLevel& lv, int plane, int cta, float4* ring_all,
                                             const unsigned* wait_ctr, unsigned wait_need, unsigned item) {
    using C = AfbStreamCfg<L, S>;
    constexpr int H2 = C::H2, NV = C::NV, NE = C::NE, D = C::D;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int it = cta * kStreamNT + tid;
    const bool active = it < lv.itemsA;
    const int ncpA = lv.ncpA;
    const int itc = active ? it : lv.itemsA - 1;     // inactive lanes shadow the last item (no copies, no stores)
    const int seg = itc / ncpA;
    const int cpl = itc - seg * ncpA;
    const int cp = lv.cp0A + cpl;
    const int i0 = seg * lv.R;                       // first output row of the segment
    const int nout = min(lv.R, lv.Ho - i0);
    const int npairs = active ? nout + H2 - 1 : 0;   // input row pairs feeding them
    const int r0 = 2 * i0 - lv.offH;                 // first input row
    const int H = lv.H, Hreal = lv.Hreal, mode = p.mode;
    const int cb = 4 * cp - (lv.offW + S);           // first column of the window: >= 0, multiple of 4
    // position in the ring row: own float4 at slot lane + (NV-1)*run; the last lane of a run adds the NV-1 extras
    const int seg_first = __shfl_sync(0xffffffffu, seg, 0);
    const int slot = lane + (NV - 1) * (seg - seg_first);
    const bool run_last = NV > 1 && active && (lane == 31 || cpl == ncpA - 1);
    const long long rs = lv.x_rs;
    const float* xcol = lv.x + (long long)plane * lv.x_ps + cb;   // column cb of row 0
    float4* ring = ring_all + (size_t)(tid >> 5) * D * C::STAGE;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)slot * 16u;

    // stage one input row pair (pair index q of this lane's segment) into ring stage `st`.  Rows inside the image
    // take the plain 16-byte cp.async (the zero-filling form costs three padding instructions each); zero rows
    // are cleared with shared stores.
    auto issue = [&](int q, int st) {
        // nothing is staged beyond this lane's segment: the lanes that would read those slots (same run, same
        // segment) are past their last pair too, and an idle lane must not touch slots that belong to others
        if (q < npairs) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int sr = afb_src_row(r0 + 2 * q + e, H, Hreal, mode);
                const unsigned dst = ring_s + (unsigned)((st * 2 + e) * C::RP) * 16u;
                if (sr >= 0) {
                    const float* src = xcol + (long long)sr * rs;
                    cp_async<4>(dst, src);
                    if (run_last) {
#pragma unroll
                        for (int k = 1; k < NV; ++k) cp_async<4>(dst + 16u * k, src + 4 * k);
                    }
                } else {
                    float4* d = ring + (st * 2 + e) * C::RP + slot;
                    d[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (run_last) {
#pragma unroll
                        for (int k = 1; k < NV; ++k) d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
    };

    const int Wo = lv.Wo;
    const size_t band = (size_t)lv.Ho * Wo;
    const bool v2lo = lv.low_vec2 != 0, v2hi = lv.out_vec2 != 0;
    const bool st_low = lv.st_low != 0, st_hi = lv.st_hi != 0;
    const float2 hsc = make_float2(lv.hi_scale, lv.hi_scale), hsh = make_float2(lv.hi_shift, lv.hi_shift);
    const long long low_rs = lv.low_rs;
    float* q0 = lv.low + (long long)plane * lv.low_ps + (long long)i0 * low_rs + 2 * cp;   // only dereferenced if st_low
    float* q1 = lv.highs + (size_t)plane * 3 * band + (size_t)i0 * Wo + 2 * cp;            // ... if st_hi

    // warp-uniform trip count (lanes of a warp may sit in segments of different length)
    int npw = npairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npw = max(npw, __shfl_xor_sync(0xffffffffu, npw, o));

    chain_wait(wait_ctr, wait_need);
    TL_MARK(1);
#pragma unroll 1
    for (int s = 0; s < D - 1; ++s) {
        issue(s, s);
        cp_async_commit();
    }
    TL_MARK(14);
    float2 acc[H2][4];   // ring of pending output rows: LL, LH, HL, HH, each (column 0, column 1)
    int st_r = 0, st_w = D - 1;
    constexpr bool kRotate = L >= 10;
    constexpr int UQ = kRotate ? 1 : H2;
    for (int qb = 0; qb < npw; qb += UQ) {
#pragma unroll
        for (int ph = 0; ph < UQ; ++ph) {
            const int q = qb + ph;
            if (q < npw) {   // warp-uniform
                cp_async_wait<D - 2>();   // this lane's copies of pair q have landed ...
                __syncwarp();             // ... and everybody's; all lanes are done reading pair q-1
                if (q < 4) TL_MARK(6 + 2 * q);
                issue(q + D - 1, st_w);   // refill the stage pair q-1 was read from
                cp_async_commit();
                st_w = st_w + 1 == D ? 0 : st_w + 1;
                {
                    float v[2][NE];
                    const float4* src = ring + (st_r * 2) * C::RP + slot;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int k = 0; k < NV; ++k) {
                            const float4 t = src[e * C::RP + k];
                            v[e][4 * k] = t.x; v[e][4 * k + 1] = t.y; v[e][4 * k + 2] = t.z; v[e][4 * k + 3] = t.w;
                        }
                    afb_pair_fma<L, S, NE>(p.t, v, acc, ph);
                    // output row q - (H2-1) has now seen all its L input rows
                    if (q >= H2 - 1 && q < npairs) {
                        const float2* s = acc[kRotate ? 0 : (ph + 1) % H2];
                        if (st_low) {
                            if (v2lo) {
                                *reinterpret_cast<float2*>(q0) = s[0];
                            } else {
                                q0[0] = s[0].x; q0[1] = s[0].y;
                            }
                        }
                        if (st_hi) {
                            const float2 lh = ffma2(s[1], hsc, hsh), hl = ffma2(s[2], hsc, hsh), hh = ffma2(s[3], hsc, hsh);
                            if (v2hi) {
                                *reinterpret_cast<float2*>(q1) = lh;
                                *reinterpret_cast<float2*>(q1 + band) = hl;
                                *reinterpret_cast<float2*>(q1 + 2 * band) = hh;
                            } else {
                                q1[0] = lh.x; q1[band] = hl.x; q1[2 * band] = hh.x;
                                q1[1] = lh.y; q1[band + 1] = hl.y; q1[2 * band + 1] = hh.y;
                            }
                        }
                        q0 += low_rs;
                        q1 += Wo;
                    }
                    if (kRotate) {
#pragma unroll
                        for (int k = 0; k + 1 < H2; ++k)
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[k][i] = acc[k + 1][i];
                    }
                }
                st_r = st_r + 1 == D ? 0 : st_r + 1;
                if (q < 4) TL_MARK(7 + 2 * q);
            }
        }
    }
    cp_async_wait<0>();
}

// ---- border columns: one thread per output position -----------------------------------------------------
// The 2..8 column pairs per row whose window leaves the image are evaluated directly: thread = (output row,
// border column); its L x L input samples are independent loads through the row / column maps (one memory
// latency instead of a dependent march), at the price of recomputing the row pass L/2 times -- for < 3 % of
// the outputs.
template <int L, int S>
__device__ __forceinline__ void afb_border_item(const AfbParams& p, const AfbLevel& lv, int plane, int it, bool active,
                                                const unsigned* wait_ctr, unsigned wait_need) {
    if (!active) it = 0;                             // idle threads shadow item 0 (they only take part in the wait)
    const int ncB = lv.Wo - 2 * lv.ncpA;             // border columns per output row
    const int i = it / ncB;
    const int e0 = it - i * ncB;
    const int k = e0 < 2 * lv.cp0A ? e0 : e0 + 2 * lv.ncpA;   // left border columns, then right border columns
    const int mode = p.mode;
    int cidx[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
        int c = 2 * k + j - lv.offW;
        if ((unsigned)c >= (unsigned)lv.Wreal) {
            c = ext_index_far(c, lv.W, mode);
            if (c >= lv.Wreal) c = -1;
        }
        cidx[j] = c;
    }
    int srow[L];
#pragma unroll
    for (int jh = 0; jh < L; ++jh) srow[jh] = afb_src_row(2 * i + jh - lv.offH, lv.H, lv.Hreal, mode);
    chain_wait(wait_ctr, wait_need);                 // index arithmetic above overlaps the wait
    if (!active) return;
    const float* xp = lv.x + (long long)plane * lv.x_ps;
    // all L x L samples are loaded before any arithmetic (clamped address + select instead of branches), so
    // the thread pays one memory round trip; rows are done in chunks of <= 6 to bound the registers
    float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
    constexpr int CH = L < 6 ? L : 6;
#pragma unroll
    for (int j0 = 0; j0 < L; j0 += CH) {
        float v[CH][L];
        bool rok[CH];
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
            const int jh = j0 + jj;
            const int sr = jh < L ? srow[jh < L ? jh : 0] : -1;
            rok[jj] = sr >= 0;
            const float* rowp = xp + (long long)max(sr, 0) * lv.x_rs;
#pragma unroll
            for (int j = 0; j < L; ++j) v[jj][j] = rowp[max(cidx[j], 0)];
        }
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
            const int jh = j0 + jj;
            if (jh < L) {
                float lo = 0.f, hi = 0.f;
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    const float x = (rok[jj] && cidx[j] >= 0) ? v[jj][j] : 0.f;
                    lo = fmaf(p.t.w_lo[j], x, lo);
                    hi = fmaf(p.t.w_hi[j], x, hi);
                }
                ll = fmaf(p.t.h_lo[jh], lo, ll);
                lh = fmaf(p.t.h_hi[jh], lo, lh);
                hl = fmaf(p.t.h_lo[jh], hi, hl);
                hh = fmaf(p.t.h_hi[jh], hi, hh);
            }
        }
    }
    const size_t band = (size_t)lv.Ho * lv.Wo;
    const size_t o = (size_t)i * lv.Wo + k;
    if (lv.st_low) lv.low[(long long)plane * lv.low_ps + (long long)i * lv.low_rs + k] = ll;
    if (lv.st_hi) {
        float* hip = lv.highs + (size_t)plane * 3 * band + o;
        hip[0] = fmaf(lh, lv.hi_scale, lv.hi_shift);
        hip[band] = fmaf(hl, lv.hi_scale, lv.hi_shift);
        hip[2 * band] = fmaf(hh, lv.hi_scale, lv.hi_shift);
    }
}

template <int L, int S>
__global__ void __launch_bounds__(kStreamNT, AfbStreamCfg<L, S>::MINB) afb_stream_kernel(const __grid_constant__ AfbParams p) {
    extern __shared__ float4 ring_all[];
    __shared__ unsigned s_item;
    const int tid = threadIdx.x;
    unsigned item = blockIdx.x;
    if (p.J > 1) {   // work items are handed out in list order: an item only waits for earlier, running ones
        if (tid == 0) s_item = atomicAdd(p.ticket, 1u);
        __syncthreads();
        item = s_item;
    }
    int level = 0;
    for (int j = 1; j < p.J; ++j)
        if ((long long)item >= p.lv[j].cta_base) level = j;
    const AfbLevel& lv = p.lv[level];
    const unsigned local = item - (unsigned)lv.cta_base;
    const int plane = (int)(local / (unsigned)lv.cpp);
    const int cta = (int)(local - (unsigned)plane * (unsigned)lv.cpp);
    TL_MARK(0);
    TL_MARK(15);
#ifdef B200W_TIMELINE
    if (g_timeline && tid == 0) {
        g_timeline[(size_t)item * 16 + 3] = ((unsigned long long)level << 48) | ((unsigned long long)plane << 24) | (unsigned)cta;
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_timeline[(size_t)item * 16 + 4] = smid | ((unsigned long long)(cta < lv.cppA ? 0 : 1) << 32);
    }
#endif
    // the previous level of this plane must be complete before its low-pass image is read
    const unsigned* wait_ctr = level > 0 ? p.done + (size_t)(level - 1) * p.planes + plane : nullptr;
    const unsigned wait_need = level > 0 ? (unsigned)p.lv[level - 1].cpp : 0u;
    if (cta < lv.cppA) {
        afb_ring_cta<L, S>(p, lv, plane, cta, ring_all, wait_ctr, wait_need, item);
    } else {
        const int it = (cta - lv.cppA) * kStreamNT + tid;
        afb_border_item<L, S>(p, lv, plane, it, it < lv.itemsB, wait_ctr, wait_need);
    }
    TL_MARK(2);
    if (level + 1 < p.J) {
        __syncthreads();
        if (tid == 0) signal_done(p.done + (size_t)level * p.planes + plane);
    }
    TL_MARK(5);
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const int v = atoi(e);
    return v > 0 ? v : dflt;
}
// tuning knob (read once): rows per segment override
static int stream_rows_override() {
    static int v = -1;
    if (v < 0) v = env_int("B200W_STREAM_ROWS", 0);
    return v;
}

bool afb_stream_supported(const AfbParams& p, int L) {
    if (L < 2 || L > 16 || (L & 1)) return false;
    const bool per = p.mode == B200W_MODE_PERIODIZATION;
    for (int j = 0; j < p.J; ++j) {
        const AfbLevel& lv = p.lv[j];
        if ((lv.x_rs & 3) || (lv.x_ps & 3) || !aligned_to(lv.x, 16)) return false;
        if (lv.offW != afb_off(L, per) || lv.offH != afb_off(L, per)) return false;
    }
    return true;
}

template <int L, int S>
static int launch_afb_stream_t(AfbParams& p, int sms, cudaStream_t st) {
    using C = AfbStreamCfg<L, S>;
    constexpr int H2 = L / 2;
    // Segments of ~16 output rows (longer for long filters, to amortise the L-2 warm-up rows): many short work
    // items balance better than one wave of long ones (measured, profiles/r01_sweep_rows.log).  A level that
    // cannot fill the resident CTA slots anyway is latency-bound (per-CTA set-up + one dependent row pair after
    // the other): it gets shorter segments, down to 2 rows for the small dependent levels of a chain.
    const int rpref = std::max(16, 4 * (H2 - 1));
    const long long slots = (long long)sms * C::MINB;
    long long base = 0;
    for (int j = 0; j < p.J; ++j) {
        AfbLevel& lv = p.lv[j];
        lv.ncp = ceil_div(lv.Wo, 2);
        // interior column pairs: window [4cp - (offW+S), +NE) inside [0, Wreal) and both output columns valid
        lv.cp0A = (lv.offW + S) / 4;
        int cpR = (lv.Wreal + lv.offW + S - C::NE) / 4 + 1;   // first pair whose window leaves the image
        if (lv.Wreal + lv.offW + S - C::NE < 0) cpR = 0;
        if (cpR > lv.Wo / 2) cpR = lv.Wo / 2;
        lv.ncpA = cpR - lv.cp0A;
        if (lv.ncpA < kMinColPairs) lv.ncpA = 0;              // too narrow for the ring: all columns take the border path
        // the first level keeps the preferred length (measured best even when it leaves CTA slots empty; an
        // SM-balancing search over R was tried and was not better, profiles/r01_notes.md); the dependent levels of a
        // chain shrink until they fill the resident slots
        int R = rpref;
        if (j > 0) {
            const int rmin = std::max(2, H2 - 1);
            while (R > rmin && (long long)p.planes * ceil_div(ceil_div(lv.Ho, R) * std::max(lv.ncpA, 1), kStreamNT) < slots)
                R = std::max(rmin, R - 2);
        }
        if (stream_rows_override() > 0) R = stream_rows_override();
        if (R > lv.Ho) R = lv.Ho;
        lv.R = R;
        const int nseg = ceil_div(lv.Ho, R);
        lv.itemsA = nseg * lv.ncpA;
        lv.RB = 1;
        lv.itemsB = lv.Ho * (lv.Wo - 2 * lv.ncpA);   // border columns: one thread per output position
        lv.cppA = ceil_div(lv.itemsA, kStreamNT);
        lv.cpp = lv.cppA + ceil_div(lv.itemsB, kStreamNT);
        lv.cta_base = base;
        base += (long long)lv.cpp * p.planes;
    }
    p.total = base;
    if (base > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    if (p.J > 1) {
        const int rc = zero_sync_words(p.ticket, (size_t)p.J * p.planes + 1, st);
        if (rc) return rc;
    }
#ifdef B200W_TIMELINE
    static unsigned long long* tl = nullptr;
    const char* tl_path = getenv("B200W_TIMELINE_FILE");
    if (tl_path) {
        if (!tl) cudaMalloc(&tl, sizeof(unsigned long long) * 16 * 65536);
        cudaMemsetAsync(tl, 0, sizeof(unsigned long long) * 16 * 65536, st);
        cudaMemcpyToSymbolAsync(g_timeline, &tl, sizeof(tl), 0, cudaMemcpyHostToDevice, st);
    }
#endif
    afb_stream_kernel<L, S><<<(unsigned)base, kStreamNT, C::smem, st>>>(p);
    const cudaError_t e = cudaGetLastError();
#ifdef B200W_TIMELINE
    if (tl_path && base <= 65536) {
        cudaStreamSynchronize(st);
        static unsigned long long host[16 * 65536];
        cudaMemcpy(host, tl, sizeof(unsigned long long) * 16 * (size_t)base, cudaMemcpyDeviceToHost);
        FILE* f = fopen(tl_path, "wb");
        if (f) { fwrite(host, sizeof(unsigned long long) * 16, (size_t)base, f); fclose(f); }
    }
#endif
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <int L>
static int launch_afb_stream_l(AfbParams& p, int sms, cudaStream_t st) {
    if (p.mode == B200W_MODE_PERIODIZATION) return launch_afb_stream_t<L, afb_shift(L, true)>(p, sms, st);
    return launch_afb_stream_t<L, afb_shift(L, false)>(p, sms, st);
}

int launch_afb_stream(AfbParams& p, int L, int sms, cudaStream_t st) {
    switch (L) {
        case 2: return launch_afb_stream_l<2>(p, sms, st);
        case 4: return launch_afb_stream_l<4>(p, sms, st);
        case 6: return launch_afb_stream_l<6>(p, sms, st);
        case 8: return launch_afb_stream_l<8>(p, sms, st);
        case 10: return launch_afb_stream_l<10>(p, sms, st);
        case 12: return launch_afb_stream_l<12>(p, sms, st);
        case 14: return launch_afb_stream_l<14>(p, sms, st);
        case 16: return launch_afb_stream_l<16>(p, sms, st);
        default: return B200W_ERR_BAD_TAPS;
    }
}

}  // namespace b200w
