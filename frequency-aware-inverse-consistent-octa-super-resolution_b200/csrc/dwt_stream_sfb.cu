// Synthesis filter bank, streaming register-blocked kernel (fused upsample + filter + accumulate).
//
// Replaces the six F.conv_transpose2d + three adds of SFB2D.forward (pw/dwt/lowlevel.py:671-680, 226-271), the
// J-level loop of DWTInverse.forward incl. the 'unpad' crop (pw/dwt/transform2d.py:134-148), and -- with the
// analysis taps and a crop -- AFB2D.backward (pw/dwt/lowlevel.py:349-365).
//
// "A-space": a = n + off (off = L-2, periodization L/2-1), so output a uses taps of parity a&1 only:
//     y[a] = sum_u c[(a>>1) - u] * g[(a&1) + 2u],  u = 0 .. L/2-1            (the polyphase form of sfb1d)
// A thread owns FOUR adjacent output columns (two coefficient pairs) and marches down Rp output row pairs:
//   * per coefficient row it loads L/2+1 coefficients of each of the four sub-bands (64-bit loads when the
//     rows are 8-byte aligned, else 32-bit),
//   * W synthesis in registers: 4 outputs for the h_lo branch (LL, HL) and 4 for the h_hi branch (LH, HH),
//   * H synthesis "accumulate forward": the row is scattered into the L/2 output row pairs it contributes to
//     (8 accumulators each); the oldest pair is complete and is written with 128-bit coalesced stores.
// No shared memory, no barriers.  Coefficients outside the sub-band are zero (periodization: wrap around); a
// null `highs` means zeros (transform2d.py:137-139); out_h / out_w smaller than the natural size crop.
#include "dwt_levels.cuh"

namespace b200w {

constexpr int sfb_off(int L, bool per) { return per ? L / 2 - 1 : L - 2; }
constexpr int sfb_a0_off(int L, bool per) { return sfb_off(L, per) & ~1; }          // A-space start of thread 0 (even)
constexpr int sfb_q0_off(int L, bool per) { return sfb_a0_off(L, per) / 2; }
constexpr int sfb_n0_off(int L, bool per) { return sfb_a0_off(L, per) - sfb_off(L, per); }   // 0 or -1
constexpr int sfb_ks_off(int L, bool per) { return sfb_q0_off(L, per) - L / 2 + 1; }        // first needed coefficient
constexpr int sfb_shift2(int L, bool per) { return ((sfb_ks_off(L, per) % 2) + 2) % 2; }    // pad down to even

template <int V, int NCF>
__device__ __forceinline__ void sfb_load_row(float (&c)[4][NCF], const float* lowp, long long low_rs, const float* hip,
                                             size_t band, int kr, int h, int w, bool periodic, bool rows_in,
                                             bool lane_in, int kb, const int (&cidx)[NCF]) {
    int sr = kr;
    if (!rows_in) sr = coef_index(kr, h, periodic);
    if (sr < 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < NCF; ++e) c[b][e] = 0.f;
        return;
    }
    const float* lp = lowp + (long long)sr * low_rs;
    const float* hp = hip ? hip + (size_t)sr * w : nullptr;
    if (lane_in) {
        if (V == 2) {
            const float2* q = reinterpret_cast<const float2*>(lp + kb);
#pragma unroll
            for (int i = 0; i < NCF / 2; ++i) { const float2 t = q[i]; c[0][2 * i] = t.x; c[0][2 * i + 1] = t.y; }
        } else {
#pragma unroll
            for (int e = 0; e < NCF; ++e) c[0][e] = lp[kb + e];
        }
        if (hp) {
#pragma unroll
            for (int b = 1; b < 4; ++b) {
                const float* src = hp + (size_t)(b - 1) * band + kb;
                if (V == 2) {
                    const float2* q = reinterpret_cast<const float2*>(src);
#pragma unroll
                    for (int i = 0; i < NCF / 2; ++i) { const float2 t = q[i]; c[b][2 * i] = t.x; c[b][2 * i + 1] = t.y; }
                } else {
#pragma unroll
                    for (int e = 0; e < NCF; ++e) c[b][e] = src[e];
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < NCF; ++e) c[0][e] = cidx[e] >= 0 ? lp[cidx[e]] : 0.f;
        if (hp) {
#pragma unroll
            for (int b = 1; b < 4; ++b)
#pragma unroll
                for (int e = 0; e < NCF; ++e) c[b][e] = cidx[e] >= 0 ? hp[(size_t)(b - 1) * band + cidx[e]] : 0.f;
        }
    }
    if (!hp) {
#pragma unroll
        for (int b = 1; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < NCF; ++e) c[b][e] = 0.f;
    }
}

template <int L, int V, int S2>
__device__ __forceinline__ void sfb_stream_item(const SfbParams& p, const SfbLevel& lv, int plane, int it) {
    constexpr int H2 = L / 2;
    constexpr int NCF = V == 2 ? 2 * ((S2 + H2 + 2) / 2) : H2 + 1;   // coefficients loaded per band and row
    constexpr int U = (H2 % 2) ? 2 * H2 : H2;                        // lcm(2, H2)

    const int seg = it / lv.nt;
    const int t = it - seg * lv.nt;
    const int m0 = lv.m_lo + seg * lv.Rp;                            // first output row pair (A-space)
    const int m_end = ((lv.offH + lv.out_h - 1) >> 1) + 1;
    const int nm = min(lv.Rp, m_end - m0);
    const int nrows = nm + H2 - 1;                                   // coefficient rows feeding them
    const int kr0 = m0 - (H2 - 1);
    const int h = lv.h, w = lv.w;
    const bool periodic = p.periodic != 0;
    const bool rows_in = kr0 >= 0 && kr0 + nrows <= h;
    const int kb = 2 * t + lv.kb_off;                                // first loaded coefficient column
    const bool lane_in = kb >= 0 && kb + NCF <= w;
    int cidx[NCF];
#pragma unroll
    for (int e = 0; e < NCF; ++e) cidx[e] = 0;
    if (!lane_in) {
#pragma unroll
        for (int e = 0; e < NCF; ++e) cidx[e] = coef_index(kb + e, w, periodic);
    }
    const size_t band = (size_t)h * w;
    const float* lowp = lv.low + (long long)plane * lv.low_ps;
    const long long low_rs = lv.low_rs;
    const float* hip = lv.highs ? lv.highs + (size_t)plane * 3 * band : nullptr;

    const int out_h = lv.out_h, out_w = lv.out_w;
    const int n0 = 4 * t + lv.n0_off;                                // first output column
    const long long y_rs = lv.y_rs;
    int nrow = 2 * m0 - lv.offH;                                     // output row of the even row of pair m0
    float* yq = lv.y + (long long)plane * lv.y_ps + (long long)nrow * y_rs + n0;   // dereferenced only where valid
    const bool full4 = n0 >= 0 && n0 + 3 < out_w;
    const int yv = full4 ? lv.y_vec : 1;

    float c[2][4][NCF];   // [double buffer][LL, LH, HL, HH][window element]
    float acc[H2][8];     // ring of pending output row pairs: even row x4 columns, odd row x4 columns
    sfb_load_row<V, NCF>(c[0], lowp, low_rs, hip, band, kr0, h, w, periodic, rows_in, lane_in, kb, cidx);

    for (int qb = 0; qb < nrows; qb += U) {
#pragma unroll
        for (int uq = 0; uq < U; ++uq) {
            const int q = qb + uq;
            if (q < nrows) {
                const int cur = uq & 1;
                const int ph = uq % H2;
                if (q + 1 < nrows)
                    sfb_load_row<V, NCF>(c[cur ^ 1], lowp, low_rs, hip, band, kr0 + q + 1, h, w, periodic, rows_in,
                                         lane_in, kb, cidx);
                // W synthesis of this coefficient row: lo = h_lo branch (LL, HL), hi = h_hi branch (LH, HH)
                float lo[4], hi[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qo = e >> 1, par = e & 1;
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int u = 0; u < H2; ++u) {
                        const int kl = S2 + qo + H2 - 1 - u;
                        a = fmaf(c[cur][0][kl], p.t.w_lo[par + 2 * u], a);
                        a = fmaf(c[cur][2][kl], p.t.w_hi[par + 2 * u], a);
                        b = fmaf(c[cur][1][kl], p.t.w_lo[par + 2 * u], b);
                        b = fmaf(c[cur][3][kl], p.t.w_hi[par + 2 * u], b);
                    }
                    lo[e] = a;
                    hi[e] = b;
                }
                // H synthesis: coefficient row q carries taps (2u, 2u+1) of output row pair q - (H2-1) + u
#pragma unroll
                for (int u = H2 - 1; u >= 0; --u) {
                    const int slot = (ph + u + 1) % H2;
                    const float a0 = p.t.h_lo[2 * u], b0 = p.t.h_hi[2 * u];
                    const float a1 = p.t.h_lo[2 * u + 1], b1 = p.t.h_hi[2 * u + 1];
                    float* s = acc[slot];
                    if (u == H2 - 1) {   // first contribution to that pair
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s[e] = lo[e] * a0;
                            s[4 + e] = lo[e] * a1;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s[e] = fmaf(lo[e], a0, s[e]);
                            s[4 + e] = fmaf(lo[e], a1, s[4 + e]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s[e] = fmaf(hi[e], b0, s[e]);
                        s[4 + e] = fmaf(hi[e], b1, s[4 + e]);
                    }
                }
                // pair q - (H2-1) is complete
                if (q >= H2 - 1) {
                    const float* s = acc[(ph + 1) % H2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = nrow + r;
                        if (row >= 0 && row < out_h) {
                            float* d = yq + r * y_rs;
                            if (yv == 4) {
                                *reinterpret_cast<float4*>(d) = make_float4(s[4 * r], s[4 * r + 1], s[4 * r + 2], s[4 * r + 3]);
                            } else if (yv == 2) {
                                *reinterpret_cast<float2*>(d) = make_float2(s[4 * r], s[4 * r + 1]);
                                *reinterpret_cast<float2*>(d + 2) = make_float2(s[4 * r + 2], s[4 * r + 3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (n0 + e >= 0 && n0 + e < out_w) d[e] = s[4 * r + e];
                            }
                        }
                    }
                    nrow += 2;
                    yq += 2 * y_rs;
                }
            }
        }
    }
}

template <int L, int V, int S2>
__global__ void __launch_bounds__(kStreamNT) sfb_stream_kernel(const __grid_constant__ SfbParams p) {
    __shared__ unsigned s_item;
    const int tid = threadIdx.x;
    unsigned item = blockIdx.x;
    if (p.J > 1) {
        if (tid == 0) s_item = atomicAdd(p.ticket, 1u);
        __syncthreads();
        item = s_item;
    }
    int level = 0;
    for (int j = 1; j < p.J; ++j)
        if ((long long)item >= p.lv[j].cta_base) level = j;
    const SfbLevel& lv = p.lv[level];
    const unsigned local = item - (unsigned)lv.cta_base;
    const int plane = (int)(local / (unsigned)lv.cpp);
    const int c = (int)(local - (unsigned)plane * (unsigned)lv.cpp);
    if (level > 0) {
        if (tid == 0) {
            const unsigned need = (unsigned)p.lv[level - 1].cpp;
            const unsigned* ctr = p.done + (size_t)(level - 1) * p.planes + plane;
            while (ld_acquire_u32(ctr) < need) __nanosleep(100);
        }
        __syncthreads();
    }
    const int it = c * kStreamNT + tid;
    if (it < lv.items) sfb_stream_item<L, V, S2>(p, lv, plane, it);
    if (level + 1 < p.J) {
        __syncthreads();
        if (tid == 0) signal_done(p.done + (size_t)level * p.planes + plane);
    }
}

static int stream_pairs_override() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_STREAM_ROWS");
        v = e ? atoi(e) : 0;
        if (v < 0) v = 0;
    }
    return v;
}

bool sfb_stream_supported(const SfbParams& p, int L) {
    if (L < 2 || L > 16 || (L & 1)) return false;
    const bool per = p.periodic != 0;
    for (int j = 0; j < p.J; ++j)
        if (p.lv[j].offW != sfb_off(L, per) || p.lv[j].offH != sfb_off(L, per)) return false;
    return true;
}

// can every level read its coefficient rows with 64-bit loads?
static bool sfb_rows_vec2(const SfbParams& p) {
    for (int j = 0; j < p.J; ++j) {
        const SfbLevel& lv = p.lv[j];
        if ((lv.low_rs & 1) || (lv.low_ps & 1) || !aligned_to(lv.low, 8)) return false;
        if (lv.highs && ((lv.w & 1) || !aligned_to(lv.highs, 8))) return false;
    }
    return true;
}

template <int L, int V, int S2>
static int launch_sfb_stream_t(SfbParams& p, int sms, cudaStream_t st) {
    constexpr int H2 = L / 2;
    const bool per = p.periodic != 0;
    const long long target = (long long)sms * 32 * 12;
    const int rmin = H2 > 1 ? 4 * (H2 - 1) : 4;
    long long base = 0;
    for (int j = 0; j < p.J; ++j) {
        SfbLevel& lv = p.lv[j];
        lv.q0_off = per ? sfb_q0_off(L, true) : sfb_q0_off(L, false);
        lv.n0_off = per ? sfb_n0_off(L, true) : sfb_n0_off(L, false);
        const int ks_off = per ? sfb_ks_off(L, true) : sfb_ks_off(L, false);
        lv.kb_off = V == 2 ? ks_off - S2 : ks_off;
        lv.m_lo = lv.offH >> 1;
        lv.nt = ceil_div(lv.out_w - lv.n0_off, 4);
        const int npairs = ((lv.offH + lv.out_h - 1) >> 1) + 1 - lv.m_lo;
        const long long rowitems = (long long)p.planes * lv.nt;
        long long nseg_want = (target + rowitems - 1) / rowitems;
        if (nseg_want < 1) nseg_want = 1;
        int Rp = (int)((npairs + nseg_want - 1) / nseg_want);
        if (Rp < rmin) Rp = rmin;
        if (stream_pairs_override() > 0) Rp = stream_pairs_override();
        if (Rp > npairs) Rp = npairs;
        lv.Rp = Rp;
        lv.items = ceil_div(npairs, Rp) * lv.nt;
        lv.cpp = ceil_div(lv.items, kStreamNT);
        lv.y_vec = 1;
        if (lv.n0_off == 0 && !(lv.y_rs & 1) && !(lv.y_ps & 1) && aligned_to(lv.y, 8)) lv.y_vec = 2;
        if (lv.y_vec == 2 && !(lv.y_rs & 3) && !(lv.y_ps & 3) && aligned_to(lv.y, 16)) lv.y_vec = 4;
        lv.cta_base = base;
        base += (long long)lv.cpp * p.planes;
    }
    p.total = base;
    if (base > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    if (p.J > 1) {
        const int rc = zero_sync_words(p.ticket, (size_t)p.J * p.planes + 1, st);
        if (rc) return rc;
    }
    sfb_stream_kernel<L, V, S2><<<(unsigned)base, kStreamNT, 0, st>>>(p);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <int L>
static int launch_sfb_stream_l(SfbParams& p, int sms, cudaStream_t st) {
    const bool v2 = sfb_rows_vec2(p);
    if (!v2) return launch_sfb_stream_t<L, 1, 0>(p, sms, st);
    if (p.periodic) return launch_sfb_stream_t<L, 2, sfb_shift2(L, true)>(p, sms, st);
    return launch_sfb_stream_t<L, 2, sfb_shift2(L, false)>(p, sms, st);
}

int launch_sfb_stream(SfbParams& p, int L, int sms, cudaStream_t st) {
    switch (L) {
        case 2: return launch_sfb_stream_l<2>(p, sms, st);
        case 4: return launch_sfb_stream_l<4>(p, sms, st);
        case 6: return launch_sfb_stream_l<6>(p, sms, st);
        case 8: return launch_sfb_stream_l<8>(p, sms, st);
        case 10: return launch_sfb_stream_l<10>(p, sms, st);
        case 12: return launch_sfb_stream_l<12>(p, sms, st);
        case 14: return launch_sfb_stream_l<14>(p, sms, st);
        case 16: return launch_sfb_stream_l<16>(p, sms, st);
        default: return B200W_ERR_BAD_TAPS;
    }
}

}  // namespace b200w
