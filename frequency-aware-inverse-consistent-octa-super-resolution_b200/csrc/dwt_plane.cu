// Plane-resident multi-level DWT / IDWT: one CTA per image plane, the small levels entirely in shared memory.
//
// Once a level's input plane fits in shared memory (about 156 x 156 samples) the remaining levels of a transform
// are latency hops, not bandwidth problems: a chain of CTAs that hands the low-pass image through global memory
// pays a publish / acquire round trip, CTA set-up and a dependent march per level.  Here one CTA of 512 threads
// loads the plane once and runs all remaining levels with __syncthreads between the passes:
//   analysis  (the levels of DWTForward below the first big ones, pw/dwt/transform2d.py:66-74; each level is
//             afb1d along W then H, pw/dwt/lowlevel.py:336-347): row pass smem -> smem (lo | hi), column pass
//             smem -> LL in smem (next level's input) and LH/HL/HH straight to `highs` in global memory;
//   synthesis (the coarse levels of DWTInverse, transform2d.py:134-148 incl. the 'unpad' crop; sfb1d x3,
//             lowlevel.py:671-680): sub-bands staged in smem, W synthesis smem -> smem, H synthesis -> the
//             next level's low-pass image in smem, the last one to global memory.
// Padding modes are index maps on the shared-memory coordinates; rows / columns beyond the real extent of a
// zero-extended level read as zero.
#include <algorithm>
#include "dwt_levels.cuh"

namespace b200w {

constexpr int kPlaneNT = 1024;
constexpr size_t kPlaneSmemMax = 216 * 1024;   // dynamic part; the index maps are static
constexpr int kPlaneMapMax = 640;              // entries of a row / column index map

static __host__ __device__ inline int pitch4i(int w) { return (w + 3) & ~3; }
static __host__ __device__ inline int pitch2i(int w) { return (w + 1) & ~1; }

// Stage `rows` rows of `cols` floats (row stride `srs`) into shared memory (row pitch `dpitch`, a multiple of the
// vector width) with cp.async: no register round trip, so every copy of a thread is in flight at once.  The
// widest vector the source alignment allows is used; a 16-byte copy may read up to 3 floats past `cols` inside
// the source row (the caller guarantees srs >= dpitch in that case).  Ends with commit; the caller waits.
__device__ __forceinline__ void stage_plane(float* dst, int dpitch, const float* src, long long srs, int rows, int cols,
                                            int tid) {
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
    const bool a16 = (srs & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (dpitch & 3) == 0 &&
                     srs >= ((cols + 3) & ~3);
    const bool a8 = (srs & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0 && (dpitch & 1) == 0 && (cols & 1) == 0;
    if (a16) {
        const int nv = (cols + 3) >> 2;
        for (int idx = tid; idx < rows * nv; idx += kPlaneNT) {
            const int r = idx / nv, c = idx - r * nv;
            cp_async<4>(d0 + (unsigned)(r * dpitch + 4 * c) * 4u, src + (long long)r * srs + 4 * c);
        }
    } else if (a8) {
        const int nv = cols >> 1;
        for (int idx = tid; idx < rows * nv; idx += kPlaneNT) {
            const int r = idx / nv, c = idx - r * nv;
            cp_async<2>(d0 + (unsigned)(r * dpitch + 2 * c) * 4u, src + (long long)r * srs + 2 * c);
        }
    } else {
        for (int idx = tid; idx < rows * cols; idx += kPlaneNT) {
            const int r = idx / cols, c = idx - r * cols;
            cp_async<1>(d0 + (unsigned)(r * dpitch + c) * 4u, src + (long long)r * srs + c);
        }
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------------
// analysis
// ------------------------------------------------------------------------------------------------
// shared-memory need (floats) of running analysis levels [first, J) of p plane-resident; 0 = does not apply
static size_t afb_plane_floats(const AfbParams& p, int first) {
    const AfbLevel& l0 = p.lv[first];
    size_t a = (size_t)l0.Hreal * pitch4i(l0.Wreal);             // level input
    size_t mid = 2 * (size_t)l0.Hreal * pitch2i(l0.Wo);          // row-pass result (lo | hi)
    size_t b = (size_t)l0.Ho * pitch4i(l0.Wo);                   // LL = next level's input
    return a + mid + b;
}

__device__ __forceinline__ int afb_map(int c, int n, int nreal, int mode) {
    if ((unsigned)c < (unsigned)nreal) return c;
    const int m = ext_index_far(c, n, mode);
    return m >= nreal ? -1 : m;
}

template <int L, int S>
__global__ void __launch_bounds__(kPlaneNT, 1) afb_plane_kernel(const __grid_constant__ AfbParams p, int first) {
    constexpr int NV = (S + L + 2 + 3) / 4;
    constexpr int NE = 4 * NV;
    extern __shared__ __align__(16) float smem_f[];
    __shared__ int s_cmap[kPlaneMapMax], s_rmap[kPlaneMapMax];
    const int tid = threadIdx.x;
    const int plane = blockIdx.x;
    const int mode = p.mode;
    const AfbLevel& l0 = p.lv[first];
    float* bufA = smem_f;
    float* mid = bufA + (size_t)l0.Hreal * pitch4i(l0.Wreal);
    float* bufB = mid + 2 * (size_t)l0.Hreal * pitch2i(l0.Wo);

    // ---- load the plane ----------------------------------------------------------------------------
    stage_plane(bufA, pitch4i(l0.Wreal), l0.x + (long long)plane * l0.x_ps, l0.x_rs, l0.Hreal, l0.Wreal, tid);
    cp_async_wait<0>();
    __syncthreads();

    float* in = bufA;
    float* out = bufB;
    for (int j = first; j < p.J; ++j) {
        const AfbLevel& lv = p.lv[j];
        const int Hreal = lv.Hreal, Wreal = lv.Wreal, H = lv.H, W = lv.W;
        const int Ho = lv.Ho, Wo = lv.Wo, offH = lv.offH, offW = lv.offW;
        const int P = pitch4i(Wreal);        // input pitch
        const int MP = pitch2i(Wo);          // mid pitch
        const int OP = pitch4i(Wo);          // output (next input) pitch
        const int ncp = (Wo + 1) >> 1;
        float* mlo = mid;
        float* mhi = mid + Hreal * MP;
        // index maps of the padding mode, once per level: columns -(offW+S) .. and rows -offH ..
        const int ncm = 4 * ncp + NE, nrm = 2 * Ho + L;
        for (int k = tid; k < ncm; k += kPlaneNT) s_cmap[k] = afb_map(k - (offW + S), W, Wreal, mode);
        for (int k = tid; k < nrm; k += kPlaneNT) s_rmap[k] = afb_map(k - offH, H, Hreal, mode);
        __syncthreads();
        // ---- row pass.  Interior column pairs: a thread keeps ONE pair and walks down the rows (constant window
        // offset, pointer increments only); the 2-8 border pairs of a row go through the column map afterwards.
        const int cp0 = (offW + S) >> 2;
        int cpR = Wreal + offW + S - NE >= 0 ? (Wreal + offW + S - NE) / 4 + 1 : 0;
        if (cpR > ncp) cpR = ncp;
        const int ncpI = cpR > cp0 ? cpR - cp0 : 0;
        if (ncpI > 0) {
            const int RG = kPlaneNT / ncpI > 0 ? kPlaneNT / ncpI : 1;
            for (int t0 = tid; t0 < RG * ncpI; t0 += kPlaneNT) {   // one trip unless the row is wider than the CTA
                const int rg = t0 / ncpI, cp = cp0 + (t0 - rg * ncpI);
                const float* row = in + rg * P + (4 * cp - (offW + S));
                float* dlo = mlo + rg * MP + 2 * cp;
                float* dhi = mhi + rg * MP + 2 * cp;
                for (int r = rg; r < Hreal; r += RG, row += RG * P, dlo += RG * MP, dhi += RG * MP) {
                    float v[NE];
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        const float4 t = reinterpret_cast<const float4*>(row)[q];
                        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                    }
                    float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
                    for (int t = 0; t < L; ++t) {
                        lo0 = fmaf(p.t.w_lo[t], v[S + t], lo0);
                        hi0 = fmaf(p.t.w_hi[t], v[S + t], hi0);
                        lo1 = fmaf(p.t.w_lo[t], v[S + t + 2], lo1);
                        hi1 = fmaf(p.t.w_hi[t], v[S + t + 2], hi1);
                    }
                    *reinterpret_cast<float2*>(dlo) = make_float2(lo0, lo1);
                    *reinterpret_cast<float2*>(dhi) = make_float2(hi0, hi1);
                }
            }
        }
        {
            const int nE = ncp - ncpI;
            for (int it = tid; it < nE * Hreal; it += kPlaneNT) {
                const int r = it / nE, e = it - r * nE;
                const int cp = e < cp0 || ncpI == 0 ? e : e - cp0 + cpR;
                const float* row = in + r * P;
                float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
                for (int t = 0; t < L + 2; ++t) {
                    const int c = s_cmap[4 * cp + S + t];
                    const float x = c >= 0 ? row[c] : 0.f;
                    if (t < L) { lo0 = fmaf(p.t.w_lo[t], x, lo0); hi0 = fmaf(p.t.w_hi[t], x, hi0); }
                    if (t >= 2) { lo1 = fmaf(p.t.w_lo[t - 2], x, lo1); hi1 = fmaf(p.t.w_hi[t - 2], x, hi1); }
                }
                *reinterpret_cast<float2*>(mlo + r * MP + 2 * cp) = make_float2(lo0, lo1);
                *reinterpret_cast<float2*>(mhi + r * MP + 2 * cp) = make_float2(hi0, hi1);
            }
        }
        __syncthreads();
        // ---- column pass: a thread keeps one column pair and walks down the output rows -------------------
        const int band = Ho * Wo;
        float* lowg = lv.low + (long long)plane * lv.low_ps;
        float* hig = lv.highs + (size_t)plane * 3 * (size_t)band;
        const bool last = j + 1 == p.J;
        const bool v2hi = lv.out_vec2 != 0, v2lo = lv.low_vec2 != 0;
        const int low_rs = (int)lv.low_rs;
        const int RGc = kPlaneNT / ncp > 0 ? kPlaneNT / ncp : 1;
        for (int t0 = tid; t0 < RGc * ncp; t0 += kPlaneNT) {
            const int ig = t0 / ncp, cp = t0 - ig * ncp;
            const int k0 = 2 * cp;
            const bool c1 = k0 + 1 < Wo;
            for (int i = ig; i < Ho; i += RGc) {
                float2 ll = make_float2(0.f, 0.f), lh = ll, hl = ll, hh = ll;
                const int rf = 2 * i - offH;
                if (rf >= 0 && rf + L <= Hreal) {   // all L source rows inside: straight strided reads
                    const float* a0 = mlo + rf * MP + k0;
                    const float* b0 = mhi + rf * MP + k0;
#pragma unroll
                    for (int t = 0; t < L; ++t) {
                        const float2 a = *reinterpret_cast<const float2*>(a0 + t * MP);
                        const float2 b = *reinterpret_cast<const float2*>(b0 + t * MP);
                        const float gl = p.t.h_lo[t], gh = p.t.h_hi[t];
                        ll.x = fmaf(gl, a.x, ll.x); ll.y = fmaf(gl, a.y, ll.y);
                        lh.x = fmaf(gh, a.x, lh.x); lh.y = fmaf(gh, a.y, lh.y);
                        hl.x = fmaf(gl, b.x, hl.x); hl.y = fmaf(gl, b.y, hl.y);
                        hh.x = fmaf(gh, b.x, hh.x); hh.y = fmaf(gh, b.y, hh.y);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < L; ++t) {
                        const int sr = s_rmap[2 * i + t];
                        if (sr >= 0) {
                            const float2 a = *reinterpret_cast<const float2*>(mlo + sr * MP + k0);
                            const float2 b = *reinterpret_cast<const float2*>(mhi + sr * MP + k0);
                            const float gl = p.t.h_lo[t], gh = p.t.h_hi[t];
                            ll.x = fmaf(gl, a.x, ll.x); ll.y = fmaf(gl, a.y, ll.y);
                            lh.x = fmaf(gh, a.x, lh.x); lh.y = fmaf(gh, a.y, lh.y);
                            hl.x = fmaf(gl, b.x, hl.x); hl.y = fmaf(gl, b.y, hl.y);
                            hh.x = fmaf(gh, b.x, hh.x); hh.y = fmaf(gh, b.y, hh.y);
                        }
                    }
                }
                if (!last) {   // LL feeds the next level from shared memory
                    float* o = out + i * OP + k0;
                    o[0] = ll.x;
                    if (c1) o[1] = ll.y;
                } else {       // the final low-pass image is an output
                    float* q = lowg + i * low_rs + k0;
                    if (v2lo && c1) *reinterpret_cast<float2*>(q) = ll;
                    else { q[0] = ll.x; if (c1) q[1] = ll.y; }
                }
                float* q = hig + i * Wo + k0;
                if (v2hi && c1) {
                    *reinterpret_cast<float2*>(q) = lh;
                    *reinterpret_cast<float2*>(q + band) = hl;
                    *reinterpret_cast<float2*>(q + 2 * band) = hh;
                } else {
                    q[0] = lh.x; q[band] = hl.x; q[2 * band] = hh.x;
                    if (c1) { q[1] = lh.y; q[band + 1] = hl.y; q[2 * band + 1] = hh.y; }
                }
            }
        }
        __syncthreads();
        float* tmp = in; in = out; out = tmp;
    }
}

constexpr int afb_plane_off(int L, bool per) { return per ? L - 1 - L / 2 : L - 2; }
constexpr int afb_plane_shift(int L, bool per) { return (4 - afb_plane_off(L, per) % 4) % 4; }

int afb_plane_first(const AfbParams& p, int L) {
    if (L < 2 || L > 16 || (L & 1)) return p.J;
    const bool per = p.mode == B200W_MODE_PERIODIZATION;
    for (int j = 0; j < p.J; ++j)
        if (p.lv[j].offW != afb_plane_off(L, per) || p.lv[j].offH != afb_plane_off(L, per)) return p.J;
    for (int j = 0; j < p.J; ++j)
        if (afb_plane_floats(p, j) * sizeof(float) <= kPlaneSmemMax && 2 * p.lv[j].Wo + 32 <= kPlaneMapMax &&
            2 * p.lv[j].Ho + L <= kPlaneMapMax)
            return j;
    return p.J;
}

template <int L, int S>
static int launch_afb_plane_t(const AfbParams& p, int first, cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kernel = afb_plane_kernel<L, S>;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        attr_set[dev] = true;
    }
    const size_t smem = afb_plane_floats(p, first) * sizeof(float);
    kernel<<<(unsigned)p.planes, kPlaneNT, smem, st>>>(p, first);
    note_launch("afb_plane_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

template <int L>
static int launch_afb_plane_l(const AfbParams& p, int first, cudaStream_t st) {
    if (p.mode == B200W_MODE_PERIODIZATION) return launch_afb_plane_t<L, afb_plane_shift(L, true)>(p, first, st);
    return launch_afb_plane_t<L, afb_plane_shift(L, false)>(p, first, st);
}

int launch_afb_plane(const AfbParams& p, int L, int first, cudaStream_t st) {
    switch (L) {
        case 2: return launch_afb_plane_l<2>(p, first, st);
        case 4: return launch_afb_plane_l<4>(p, first, st);
        case 6: return launch_afb_plane_l<6>(p, first, st);
        case 8: return launch_afb_plane_l<8>(p, first, st);
        case 10: return launch_afb_plane_l<10>(p, first, st);
        case 12: return launch_afb_plane_l<12>(p, first, st);
        case 14: return launch_afb_plane_l<14>(p, first, st);
        case 16: return launch_afb_plane_l<16>(p, first, st);
        default: return B200W_ERR_BAD_TAPS;
    }
}

// ------------------------------------------------------------------------------------------------
// synthesis
// ------------------------------------------------------------------------------------------------
// A-space geometry of one level: output n <-> a = n + off; column pairs q = a >> 1 in [q_lo, q_hi]
struct SfbGeom {
    int q_lo, nq;      // first column pair, number of column pairs (W)
    int m_lo, nm;      // first row pair, number of row pairs (H)
};
static __host__ __device__ inline SfbGeom sfb_geom(const SfbLevel& lv) {
    SfbGeom g;
    g.q_lo = lv.offW >> 1;
    g.nq = ((lv.offW + lv.out_w - 1) >> 1) - g.q_lo + 1;
    g.m_lo = lv.offH >> 1;
    g.nm = ((lv.offH + lv.out_h - 1) >> 1) - g.m_lo + 1;
    return g;
}

// shared-memory layout (floats) of chain position c: [low][3 sub-bands][mid lo|hi][out]; the head keeps every
// output but the last in shared memory
struct SfbPlaneLayout {
    size_t low, hi, mid, out, total;
};
static SfbPlaneLayout sfb_plane_layout(const SfbParams& p, int c, bool out_in_smem) {
    const SfbLevel& lv = p.lv[c];
    const SfbGeom g = sfb_geom(lv);
    SfbPlaneLayout l;
    // position 0 stages yl; later positions read the previous output (kept in the `out` region of that position)
    l.low = c == 0 ? (size_t)lv.h * pitch2i(lv.w) : 0;
    l.hi = lv.highs ? 3 * (size_t)lv.h * pitch2i(lv.w) : 0;
    l.mid = 2 * (size_t)lv.h * (2 * (size_t)g.nq);
    l.out = out_in_smem ? (size_t)lv.out_h * pitch2i(lv.out_w) : 0;
    l.total = l.low + l.hi + l.mid + l.out;
    return l;
}

// number of leading chain positions that run plane-resident (0 = none).  Region plan: two ping-pong `out`
// regions sized for the largest kept output, plus low/hi/mid sized for the largest position.
struct SfbPlanePlan {
    int count;
    size_t out_sz, low_sz, hi_sz, mid_sz;
    size_t total() const { return 2 * out_sz + low_sz + hi_sz + mid_sz; }
};
static SfbPlanePlan sfb_plane_plan(const SfbParams& p, int count) {
    SfbPlanePlan pl{count, 0, 0, 0, 0};
    for (int c = 0; c < count; ++c) {
        const SfbPlaneLayout l = sfb_plane_layout(p, c, c + 1 < count);
        pl.out_sz = std::max(pl.out_sz, l.out);
        pl.low_sz = std::max(pl.low_sz, l.low);
        pl.hi_sz = std::max(pl.hi_sz, l.hi);
        pl.mid_sz = std::max(pl.mid_sz, l.mid);
    }
    pl.out_sz = (pl.out_sz + 3) & ~(size_t)3;
    pl.low_sz = (pl.low_sz + 3) & ~(size_t)3;
    pl.hi_sz = (pl.hi_sz + 3) & ~(size_t)3;
    return pl;
}

__device__ __forceinline__ int sfb_map(int k, int m, int periodic) {
    if ((unsigned)k < (unsigned)m) return k;
    return periodic ? coef_index_far(k, m, periodic) : -1;
}

template <int L>
__global__ void __launch_bounds__(kPlaneNT, 1) sfb_plane_kernel(const __grid_constant__ SfbParams p, int count, int out_sz,
                                                                int low_sz, int hi_sz) {
    constexpr int H2 = L / 2;
    extern __shared__ __align__(16) float smem_f[];
    const int tid = threadIdx.x;
    const int plane = blockIdx.x;
    const int periodic = p.periodic;
    float* outbuf[2] = {smem_f, smem_f + out_sz};
    float* lowbuf = smem_f + 2 * (size_t)out_sz;
    float* hibuf = lowbuf + low_sz;
    float* mid = hibuf + hi_sz;

    const float* low_s = nullptr;   // previous output in shared memory
    int low_pitch = 0;
    for (int c = 0; c < count; ++c) {
        const SfbLevel& lv = p.lv[c];
        const int h = lv.h, w = lv.w, out_h = lv.out_h, out_w = lv.out_w;
        const int offH = lv.offH, offW = lv.offW;
        const int WP = pitch2i(w);
        const SfbGeom g = sfb_geom(lv);
        const size_t band = (size_t)h * w;
        // ---- stage the inputs ----------------------------------------------------------------------
        if (c == 0) {
            stage_plane(lowbuf, WP, lv.low + (long long)plane * lv.low_ps, lv.low_rs, h, w, tid);
            low_s = lowbuf;
            low_pitch = WP;
        }
        const bool has_hi = lv.highs != nullptr;
        if (has_hi)   // the three sub-bands are 3h dense rows of w
            stage_plane(hibuf, WP, lv.highs + (size_t)plane * 3 * band, w, 3 * h, w, tid);
        cp_async_wait<0>();
        __syncthreads();
        // ---- W synthesis: every coefficient row, both outputs of a column pair, both H branches ------
        const int MP = 2 * g.nq;               // mid pitch (A-space columns 2*q_lo .. )
        float* mlo = mid;
        float* mhi = mid + (size_t)h * MP;
        for (int it = tid; it < h * g.nq; it += kPlaneNT) {
            const int kr = it / g.nq, qi = it - kr * g.nq;
            const int q = g.q_lo + qi;
            float le = 0.f, lo_ = 0.f, he = 0.f, ho = 0.f;   // h_lo branch even/odd column, h_hi branch even/odd
#pragma unroll
            for (int u = 0; u < H2; ++u) {
                const int k = sfb_map(q - u, w, periodic);
                if (k >= 0) {
                    const float cl = low_s[(size_t)kr * low_pitch + k];
                    float c1 = 0.f, c2 = 0.f, c3 = 0.f;
                    if (has_hi) {
                        c1 = hibuf[((size_t)0 * h + kr) * WP + k];
                        c2 = hibuf[((size_t)1 * h + kr) * WP + k];
                        c3 = hibuf[((size_t)2 * h + kr) * WP + k];
                    }
                    const float g0e = p.t.w_lo[2 * u], g0o = p.t.w_lo[2 * u + 1];
                    const float g1e = p.t.w_hi[2 * u], g1o = p.t.w_hi[2 * u + 1];
                    le = fmaf(cl, g0e, le); le = fmaf(c2, g1e, le);     // LL, HL -> h_lo branch
                    lo_ = fmaf(cl, g0o, lo_); lo_ = fmaf(c2, g1o, lo_);
                    he = fmaf(c1, g0e, he); he = fmaf(c3, g1e, he);     // LH, HH -> h_hi branch
                    ho = fmaf(c1, g0o, ho); ho = fmaf(c3, g1o, ho);
                }
            }
            *reinterpret_cast<float2*>(mlo + (size_t)kr * MP + 2 * qi) = make_float2(le, lo_);
            *reinterpret_cast<float2*>(mhi + (size_t)kr * MP + 2 * qi) = make_float2(he, ho);
        }
        __syncthreads();
        // ---- H synthesis: one output row pair x one column pair per item ---------------------------------
        const bool keep = c + 1 < count;
        float* ob = outbuf[c & 1];
        const int OP = pitch2i(out_w);
        float* yg = lv.y + (long long)plane * lv.y_ps;
        for (int it = tid; it < g.nm * g.nq; it += kPlaneNT) {
            const int mi = it / g.nq, qi = it - mi * g.nq;
            const int m = g.m_lo + mi;
            float2 ye = make_float2(0.f, 0.f), yo = ye;   // even / odd output row
#pragma unroll
            for (int u = 0; u < H2; ++u) {
                const int kr = sfb_map(m - u, h, periodic);
                if (kr >= 0) {
                    const float2 a = *reinterpret_cast<const float2*>(mlo + (size_t)kr * MP + 2 * qi);
                    const float2 b = *reinterpret_cast<const float2*>(mhi + (size_t)kr * MP + 2 * qi);
                    const float g0e = p.t.h_lo[2 * u], g0o = p.t.h_lo[2 * u + 1];
                    const float g1e = p.t.h_hi[2 * u], g1o = p.t.h_hi[2 * u + 1];
                    ye.x = fmaf(a.x, g0e, ye.x); ye.y = fmaf(a.y, g0e, ye.y);
                    ye.x = fmaf(b.x, g1e, ye.x); ye.y = fmaf(b.y, g1e, ye.y);
                    yo.x = fmaf(a.x, g0o, yo.x); yo.y = fmaf(a.y, g0o, yo.y);
                    yo.x = fmaf(b.x, g1o, yo.x); yo.y = fmaf(b.y, g1o, yo.y);
                }
            }
            const int n0 = 2 * (g.q_lo + qi) - offW;     // output column of the even A-space column
            const int r0 = 2 * m - offH;                 // output row of the even A-space row
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = r0 + rr;
                if (row < 0 || row >= out_h) continue;
                const float2 v = rr == 0 ? ye : yo;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = n0 + cc;
                    if (col < 0 || col >= out_w) continue;
                    const float val = cc == 0 ? v.x : v.y;
                    if (keep) ob[(size_t)row * OP + col] = val;
                    else yg[(long long)row * lv.y_rs + col] = val;
                }
            }
        }
        __syncthreads();
        low_s = ob;
        low_pitch = OP;
    }
}

int sfb_plane_count(const SfbParams& p, int L) {
    if (L < 2 || L > 16 || (L & 1)) return 0;
    const bool per = p.periodic != 0;
    const int off = per ? L / 2 - 1 : L - 2;
    for (int j = 0; j < p.J; ++j)
        if (p.lv[j].offW != off || p.lv[j].offH != off) return 0;
    int best = 0;
    for (int count = 1; count <= p.J; ++count) {
        if (sfb_plane_plan(p, count).total() * sizeof(float) > kPlaneSmemMax) break;
        best = count;
    }
    return best;
}

template <int L>
static int launch_sfb_plane_t(const SfbParams& p, int count, cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kernel = sfb_plane_kernel<L>;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
        if (e != cudaSuccess) return set_last_cuda_error(e);
        attr_set[dev] = true;
    }
    const SfbPlanePlan pl = sfb_plane_plan(p, count);
    kernel<<<(unsigned)p.planes, kPlaneNT, pl.total() * sizeof(float), st>>>(p, count, (int)pl.out_sz, (int)pl.low_sz,
                                                                             (int)pl.hi_sz);
    note_launch("sfb_plane_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

int launch_sfb_plane(const SfbParams& p, int L, int count, cudaStream_t st) {
    switch (L) {
        case 2: return launch_sfb_plane_t<2>(p, count, st);
        case 4: return launch_sfb_plane_t<4>(p, count, st);
        case 6: return launch_sfb_plane_t<6>(p, count, st);
        case 8: return launch_sfb_plane_t<8>(p, count, st);
        case 10: return launch_sfb_plane_t<10>(p, count, st);
        case 12: return launch_sfb_plane_t<12>(p, count, st);
        case 14: return launch_sfb_plane_t<14>(p, count, st);
        case 16: return launch_sfb_plane_t<16>(p, count, st);
        default: return B200W_ERR_BAD_TAPS;
    }
}

}  // namespace b200w
