// Analysis filter bank, streaming register-blocked kernel (the fast path for 16-byte aligned rows).
//
// Replaces afb1d(dim=3) + afb1d(dim=2) + reshape + 2x .contiguous() of AFB2D.forward
// (pw/dwt/lowlevel.py:336-347, 91-172) -- and the J-level loop around it (pw/dwt/transform2d.py:66-74).
//
// A thread owns ONE PAIR of adjacent output columns and marches down R output rows of one image plane:
//   * per input row it loads the 4*NV floats its two outputs need with NV aligned 128-bit loads (a warp reads
//     512 contiguous bytes; the half-window shared with the neighbouring lane is an L1 hit),
//   * row pass in registers: lo/hi for both columns, 4L FMAs with the taps as constant-bank operands,
//   * column pass "accumulate forward": each pair of input rows is scattered into the L/2 output rows it
//     contributes to (8 accumulators each); the oldest one is complete, is stored with 64-bit coalesced stores
//     (LL into `low`, LH/HL/HH straight into `highs[:, :, 0..2]`) and its registers are recycled.
// No shared memory, no barriers, ~2L FMAs + ~1 load/store instruction per input pixel.  The loop is unrolled by
// lcm(2, L/2) row pairs so that the accumulator ring and the load double buffer have static register names;
// the loads of the next row pair are issued before the arithmetic of the current one.
// Segments overlap by L-2 input rows (re-read through L2).  Padding: rows are remapped per row (only in
// segments that touch the border); lanes whose window leaves [0, W) load element-wise through a precomputed
// column map (symmetric / reflect / periodic / periodization), "zero" = no load.
#include "dwt_levels.cuh"

namespace b200w {

constexpr int afb_off(int L, bool per) { return per ? L - 1 - L / 2 : L - 2; }
// the first loaded column is 4*cp - (off + S): S pads the window start down to a multiple of 4
constexpr int afb_shift(int L, bool per) { return (4 - afb_off(L, per) % 4) % 4; }

template <int NE>
__device__ __forceinline__ void afb_load_row(float (&v)[NE], const float* xp, long long rs, int r, int H, int Hreal,
                                             int mode, bool rows_in, bool lane_in, int cb, const int (&cidx)[NE]) {
    int sr = r;
    if (!rows_in) {
        sr = ext_index(r, H, mode);
        if (sr >= Hreal) sr = -1;
    }
    if (sr < 0) {
#pragma unroll
        for (int e = 0; e < NE; ++e) v[e] = 0.f;
        return;
    }
    const float* rowp = xp + (long long)sr * rs;
    if (lane_in) {
        const float4* q = reinterpret_cast<const float4*>(rowp + cb);
#pragma unroll
        for (int i = 0; i < NE / 4; ++i) {
            const float4 t = q[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) v[e] = cidx[e] >= 0 ? rowp[cidx[e]] : 0.f;
    }
}

template <int L, int S>
__device__ __forceinline__ void afb_stream_item(const AfbParams& p, const AfbLevel& lv, int plane, int it) {
    constexpr int H2 = L / 2;
    constexpr int NV = (S + L + 2 + 3) / 4;
    constexpr int NE = 4 * NV;
    constexpr int U = (H2 % 2) ? 2 * H2 : H2;   // lcm(2, H2)

    const int seg = it / lv.ncp;
    const int cp = it - seg * lv.ncp;
    const int i0 = seg * lv.R;                       // first output row of the segment
    const int nout = min(lv.R, lv.Ho - i0);
    const int npairs = nout + H2 - 1;                // input row pairs feeding them
    const int r0 = 2 * i0 - lv.offH;                 // first input row
    const int H = lv.H, Hreal = lv.Hreal, mode = p.mode;
    const bool rows_in = r0 >= 0 && r0 + 2 * npairs <= Hreal;
    const int cb = 4 * cp - (lv.offW + S);           // first loaded column (multiple of 4)
    const bool lane_in = cb >= 0 && cb + NE <= lv.Wreal;
    int cidx[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) cidx[e] = 0;
    if (!lane_in) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int c = ext_index(cb + e, lv.W, mode);
            cidx[e] = c >= lv.Wreal ? -1 : c;
        }
    }
    const float* xp = lv.x + (long long)plane * lv.x_ps;
    const long long rs = lv.x_rs;

    const int Wo = lv.Wo;
    const size_t band = (size_t)lv.Ho * Wo;
    const int k0 = 2 * cp;
    const bool c1ok = k0 + 1 < Wo;
    const bool v2lo = lv.low_vec2 && c1ok, v2hi = lv.out_vec2 && c1ok;
    const long long low_rs = lv.low_rs;
    float* q0 = lv.low + (long long)plane * lv.low_ps + (long long)i0 * low_rs + k0;
    float* q1 = lv.highs + (size_t)plane * 3 * band + (size_t)i0 * Wo + k0;

    float v[2][2][NE];   // [double buffer][row of the pair][window element]
    float acc[H2][8];    // ring of pending output rows: LL.x LL.y LH.x LH.y HL.x HL.y HH.x HH.y
    afb_load_row<NE>(v[0][0], xp, rs, r0, H, Hreal, mode, rows_in, lane_in, cb, cidx);
    afb_load_row<NE>(v[0][1], xp, rs, r0 + 1, H, Hreal, mode, rows_in, lane_in, cb, cidx);

    for (int qb = 0; qb < npairs; qb += U) {
#pragma unroll
        for (int uq = 0; uq < U; ++uq) {
            const int q = qb + uq;
            if (q < npairs) {
                const int cur = uq & 1;
                const int ph = uq % H2;
                if (q + 1 < npairs) {   // next pair's loads go out before this pair's arithmetic
                    afb_load_row<NE>(v[cur ^ 1][0], xp, rs, r0 + 2 * q + 2, H, Hreal, mode, rows_in, lane_in, cb, cidx);
                    afb_load_row<NE>(v[cur ^ 1][1], xp, rs, r0 + 2 * q + 3, H, Hreal, mode, rows_in, lane_in, cb, cidx);
                }
                // row pass (along W, decimated): outputs k0 and k0+1 of both rows of the pair
                float rl[2][2], rh[2][2];   // [row of the pair][column]
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        lo0 = fmaf(p.t.w_lo[j], v[cur][e][S + j], lo0);
                        hi0 = fmaf(p.t.w_hi[j], v[cur][e][S + j], hi0);
                        lo1 = fmaf(p.t.w_lo[j], v[cur][e][S + j + 2], lo1);
                        hi1 = fmaf(p.t.w_hi[j], v[cur][e][S + j + 2], hi1);
                    }
                    rl[e][0] = lo0; rl[e][1] = lo1; rh[e][0] = hi0; rh[e][1] = hi1;
                }
                // column pass: this pair carries taps (2u, 2u+1) of output row q - u
#pragma unroll
                for (int u = 0; u < H2; ++u) {
                    const int slot = (ph - u + H2) % H2;
                    const float a = p.t.h_lo[2 * u], b = p.t.h_hi[2 * u];
                    const float c = p.t.h_lo[2 * u + 1], d = p.t.h_hi[2 * u + 1];
                    float* s = acc[slot];
                    if (u == 0) {   // first contribution: start the accumulators
                        s[0] = a * rl[0][0]; s[1] = a * rl[0][1];
                        s[2] = b * rl[0][0]; s[3] = b * rl[0][1];
                        s[4] = a * rh[0][0]; s[5] = a * rh[0][1];
                        s[6] = b * rh[0][0]; s[7] = b * rh[0][1];
                    } else {
                        s[0] = fmaf(a, rl[0][0], s[0]); s[1] = fmaf(a, rl[0][1], s[1]);
                        s[2] = fmaf(b, rl[0][0], s[2]); s[3] = fmaf(b, rl[0][1], s[3]);
                        s[4] = fmaf(a, rh[0][0], s[4]); s[5] = fmaf(a, rh[0][1], s[5]);
                        s[6] = fmaf(b, rh[0][0], s[6]); s[7] = fmaf(b, rh[0][1], s[7]);
                    }
                    s[0] = fmaf(c, rl[1][0], s[0]); s[1] = fmaf(c, rl[1][1], s[1]);   // LL: W-lo, H-lo
                    s[2] = fmaf(d, rl[1][0], s[2]); s[3] = fmaf(d, rl[1][1], s[3]);   // LH: W-lo, H-hi
                    s[4] = fmaf(c, rh[1][0], s[4]); s[5] = fmaf(c, rh[1][1], s[5]);   // HL: W-hi, H-lo
                    s[6] = fmaf(d, rh[1][0], s[6]); s[7] = fmaf(d, rh[1][1], s[7]);   // HH
                }
                // output row q - (H2-1) has now seen all its L input rows
                if (q >= H2 - 1) {
                    const float* s = acc[(ph + 1) % H2];
                    if (v2lo) {
                        *reinterpret_cast<float2*>(q0) = make_float2(s[0], s[1]);
                    } else {
                        q0[0] = s[0];
                        if (c1ok) q0[1] = s[1];
                    }
                    if (v2hi) {
                        *reinterpret_cast<float2*>(q1) = make_float2(s[2], s[3]);
                        *reinterpret_cast<float2*>(q1 + band) = make_float2(s[4], s[5]);
                        *reinterpret_cast<float2*>(q1 + 2 * band) = make_float2(s[6], s[7]);
                    } else {
                        q1[0] = s[2]; q1[band] = s[4]; q1[2 * band] = s[6];
                        if (c1ok) { q1[1] = s[3]; q1[band + 1] = s[5]; q1[2 * band + 1] = s[7]; }
                    }
                    q0 += low_rs;
                    q1 += Wo;
                }
            }
        }
    }
}

template <int L, int S>
__global__ void __launch_bounds__(kStreamNT) afb_stream_kernel(const __grid_constant__ AfbParams p) {
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
    const int c = (int)(local - (unsigned)plane * (unsigned)lv.cpp);
    if (level > 0) {   // the previous level of this plane must be complete
        if (tid == 0) {
            const unsigned need = (unsigned)p.lv[level - 1].cpp;
            const unsigned* ctr = p.done + (size_t)(level - 1) * p.planes + plane;
            while (ld_acquire_u32(ctr) < need) __nanosleep(100);
        }
        __syncthreads();
    }
    const int it = c * kStreamNT + tid;
    if (it < lv.items) afb_stream_item<L, S>(p, lv, plane, it);
    if (level + 1 < p.J) {
        __syncthreads();
        if (tid == 0) signal_done(p.done + (size_t)level * p.planes + plane);
    }
}

static int stream_rows_override() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_STREAM_ROWS");
        v = e ? atoi(e) : 0;
        if (v < 0) v = 0;
    }
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
    constexpr int H2 = L / 2;
    // enough thread items for ~12 warps per SM, but segments long enough to amortise the L-2 warm-up rows
    const long long target = (long long)sms * 32 * 12;
    const int rmin = H2 > 1 ? 4 * (H2 - 1) : 4;
    long long base = 0;
    for (int j = 0; j < p.J; ++j) {
        AfbLevel& lv = p.lv[j];
        lv.ncp = ceil_div(lv.Wo, 2);
        const long long rowitems = (long long)p.planes * lv.ncp;
        long long nseg_want = (target + rowitems - 1) / rowitems;
        if (nseg_want < 1) nseg_want = 1;
        int R = (int)((lv.Ho + nseg_want - 1) / nseg_want);
        if (R < rmin) R = rmin;
        if (stream_rows_override() > 0) R = stream_rows_override();
        if (R > lv.Ho) R = lv.Ho;
        lv.R = R;
        lv.items = ceil_div(lv.Ho, R) * lv.ncp;
        lv.cpp = ceil_div(lv.items, kStreamNT);
        lv.cta_base = base;
        base += (long long)lv.cpp * p.planes;
    }
    p.total = base;
    if (base > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    if (p.J > 1) {
        cudaError_t e = cudaMemsetAsync(p.ticket, 0, sizeof(unsigned) * ((size_t)p.J * p.planes + 1), st);
        if (e != cudaSuccess) return set_last_cuda_error(e);
    }
    afb_stream_kernel<L, S><<<(unsigned)base, kStreamNT, 0, st>>>(p);
    const cudaError_t e = cudaGetLastError();
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
