/*
 * CPU port of the reference's wavelet + SSIM hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain C (fp32, OpenMP over image planes) restating, pass by pass, what the reference does on the CPU:
 *   afb1d   pytorch_wavelets/pytorch_wavelets/dwt/lowlevel.py:91-172   pad (materialised) + stride-2 correlation
 *   sfb1d   .../dwt/lowlevel.py:226-271                                 transposed conv, crop / wrap-add + roll
 *   AFB2D   .../dwt/lowlevel.py:336-365    SFB2D .../dwt/lowlevel.py:671-694
 *   _ssim   ssim.py:17-37 (dense 11x11 window, five blurs) and the closed form of its autograd backward
 *
 * It exists for two reasons only: (1) it is cross-checked against oracle/dwt_oracle.py / ssim_oracle.py
 * (which are pinned to the reference's own outputs) and (2) bench.py times it on the GPU box's host cores
 * as the CPU baseline ("kind": "port").  The product never links or calls it.
 *
 * Build: make -C oracle   ->  oracle/_build/libref_port.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { MODE_ZERO = 0, MODE_SYMMETRIC = 1, MODE_PER = 2, MODE_REFLECT = 4, MODE_PERIODIC = 6 };

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for the cores it may use explicitly */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int ext_index(int s, int n, int mode) {
    if (s >= 0 && s < n) return s;
    int p, m;
    switch (mode) {
        case MODE_SYMMETRIC: p = 2 * n; m = s % p; if (m < 0) m += p; return m < n ? m : p - 1 - m;
        case MODE_REFLECT: if (n == 1) return 0; p = 2 * n - 2; m = s % p; if (m < 0) m += p; return m < n ? m : p - m;
        case MODE_PERIODIC: m = s % n; if (m < 0) m += n; return m;
        case MODE_PER: p = n + (n & 1); m = s % p; if (m < 0) m += p; return m < n ? m : n - 1;
        default: return -1;
    }
}

static int coeff_len(int n, int l, int mode) { return mode == MODE_PER ? (n + 1) / 2 : (n + l - 1) / 2; }

/* one 1-D analysis pass along a strided axis: src has `n` samples at stride `ss`; writes m lo / hi samples */
static void afb1d_line(const float* src, int n, long ss, const float* lo_t, const float* hi_t, int L, int mode,
                       float* ext, float* lo, float* hi, long ds, int m) {
    int off;
    if (mode == MODE_PER) off = L - 1 - L / 2;
    else { int p = 2 * (m - 1) - n + L; off = p / 2; }
    const int ne = 2 * (m - 1) + L;              /* padded length the correlation reads */
    for (int i = 0; i < ne; ++i) {               /* the pad is materialised, like mypad / F.pad */
        int s = ext_index(i - off, n, mode);
        ext[i] = s < 0 ? 0.f : src[s * ss];
    }
    for (int k = 0; k < m; ++k) {
        float a = 0.f, b = 0.f;
        const float* e = ext + 2 * k;
        for (int j = 0; j < L; ++j) { a += lo_t[j] * e[j]; b += hi_t[j] * e[j]; }
        lo[k * ds] = a;
        hi[k * ds] = b;
    }
}

/* x (planes,H,W) -> low (planes,Ho,Wo), highs (planes,3,Ho,Wo) */
void ref_afb2d(const float* x, int planes, int H, int W, const float* w_lo, const float* w_hi, int Lw,
               const float* h_lo, const float* h_hi, int Lh, int mode, float* low, float* highs) {
    const int Ho = coeff_len(H, Lh, mode), Wo = coeff_len(W, Lw, mode);
#pragma omp parallel
    {
        float* rowlo = (float*)malloc(sizeof(float) * (size_t)H * Wo);
        float* rowhi = (float*)malloc(sizeof(float) * (size_t)H * Wo);
        float* ext = (float*)malloc(sizeof(float) * (size_t)(2 * (H > W ? H : W) + 4 * (Lw > Lh ? Lw : Lh) + 8));
#pragma omp for schedule(static)
        for (int p = 0; p < planes; ++p) {
            const float* xp = x + (size_t)p * H * W;
            for (int r = 0; r < H; ++r)          /* row pass (dim 3) */
                afb1d_line(xp + (size_t)r * W, W, 1, w_lo, w_hi, Lw, mode, ext, rowlo + (size_t)r * Wo,
                           rowhi + (size_t)r * Wo, 1, Wo);
            float* ll = low + (size_t)p * Ho * Wo;
            float* lh = highs + (size_t)p * 3 * Ho * Wo;
            float* hl = lh + (size_t)Ho * Wo;
            float* hh = hl + (size_t)Ho * Wo;
            for (int c = 0; c < Wo; ++c) {       /* column pass (dim 2) */
                afb1d_line(rowlo + c, H, Wo, h_lo, h_hi, Lh, mode, ext, ll + c, lh + c, Wo, Ho);
                afb1d_line(rowhi + c, H, Wo, h_lo, h_hi, Lh, mode, ext, hl + c, hh + c, Wo, Ho);
            }
        }
        free(rowlo); free(rowhi); free(ext);
    }
}

/* one 1-D synthesis line: m coefficients each of lo / hi -> nout samples (first `nout` of the natural output) */
static void sfb1d_line(const float* lo, const float* hi, int m, long ss, const float* g0, const float* g1, int L,
                       int mode, float* full, float* dst, long ds, int nout) {
    const int nf = 2 * (m - 1) + L;
    for (int i = 0; i < nf; ++i) full[i] = 0.f;
    for (int k = 0; k < m; ++k) {                /* conv_transpose, stride 2 */
        const float a = lo[k * ss], b = hi ? hi[k * ss] : 0.f;
        float* f = full + 2 * k;
        for (int j = 0; j < L; ++j) f[j] += a * g0[j] + b * g1[j];
    }
    if (mode == MODE_PER) {
        const int N = 2 * m;
        for (int i = 0; i < L - 2; ++i) full[i] += full[N + i];
        const int shift = L / 2 - 1;             /* roll(y, 1 - L/2) */
        for (int i = 0; i < nout; ++i) dst[i * ds] = full[(i + shift) % N];
    } else {
        for (int i = 0; i < nout; ++i) dst[i * ds] = full[i + L - 2];
    }
}

/* low (planes,h,w), highs (planes,3,h,w) or NULL -> y (planes,out_h,out_w) */
void ref_sfb2d(const float* low, const float* highs, int planes, int h, int w, const float* w_lo, const float* w_hi,
               int Lw, const float* h_lo, const float* h_hi, int Lh, int mode, float* y, int out_h, int out_w) {
    const int full_h = mode == MODE_PER ? 2 * h : 2 * h - Lh + 2;
#pragma omp parallel
    {
        float* lo = (float*)malloc(sizeof(float) * (size_t)full_h * w);
        float* hi = (float*)malloc(sizeof(float) * (size_t)full_h * w);
        float* full = (float*)malloc(sizeof(float) * (size_t)(2 * (h > w ? h : w) + 2 * (Lw > Lh ? Lw : Lh) + 8));
#pragma omp for schedule(static)
        for (int p = 0; p < planes; ++p) {
            const float* ll = low + (size_t)p * h * w;
            const float* lh = highs ? highs + (size_t)p * 3 * h * w : NULL;
            const float* hl = lh ? lh + (size_t)h * w : NULL;
            const float* hh = hl ? hl + (size_t)h * w : NULL;
            for (int c = 0; c < w; ++c) {        /* column synthesis (dim 2): lo = (ll, lh), hi = (hl, hh) */
                sfb1d_line(ll + c, lh ? lh + c : NULL, h, w, h_lo, h_hi, Lh, mode, full, lo + c, w, out_h);
                if (hl) sfb1d_line(hl + c, hh + c, h, w, h_lo, h_hi, Lh, mode, full, hi + c, w, out_h);
                else for (int r = 0; r < out_h; ++r) hi[(size_t)r * w + c] = 0.f;
            }
            float* yp = y + (size_t)p * out_h * out_w;
            for (int r = 0; r < out_h; ++r)      /* row synthesis (dim 3) */
                sfb1d_line(lo + (size_t)r * w, hi + (size_t)r * w, w, 1, w_lo, w_hi, Lw, mode, full,
                           yp + (size_t)r * out_w, 1, out_w);
        }
        free(lo); free(hi); free(full);
    }
}

/* dense ws x ws zero-padded "same" correlation of one plane (F.conv2d(..., padding=ws/2, groups=C)) */
static void blur2d(const float* src, int H, int W, const float* win2d, int ws, float* dst) {
    const int p = ws / 2;
    for (int r = 0; r < H; ++r) {
        float* d = dst + (size_t)r * W;
        for (int c = 0; c < W; ++c) d[c] = 0.f;
        for (int i = 0; i < ws; ++i) {
            const int sr = r + i - p;
            if (sr < 0 || sr >= H) continue;
            const float* s = src + (size_t)sr * W;
            for (int j = 0; j < ws; ++j) {
                const float wv = win2d[i * ws + j];
                const int sh = j - p;
                const int c0 = sh < 0 ? -sh : 0, c1 = sh > 0 ? W - sh : W;
                for (int c = c0; c < c1; ++c) d[c] += wv * s[c + sh];
            }
        }
    }
}

/*
 * SSIM forward (+ optional gradient w.r.t. img1 and img2).  out: 1 float (size_average) or N floats.
 * grad_out: 1 or N floats (may be NULL = ones).  d1/d2 may be NULL.
 */
void ref_ssim(const float* img1, const float* img2, int N, int C, int H, int W, const float* win2d, int ws,
              int size_average, float* out, const float* grad_out, float* d1, float* d2) {
    const float C1 = 0.0001f, C2 = 0.0009f;
    const size_t hw = (size_t)H * W;
    const int planes = N * C;
    double* psum = (double*)calloc((size_t)planes, sizeof(double));
#pragma omp parallel
    {
        float* buf = (float*)malloc(sizeof(float) * hw * 12);
        float *mu1 = buf, *mu2 = buf + hw, *e11 = buf + 2 * hw, *e22 = buf + 3 * hw, *e12 = buf + 4 * hw,
              *tmp = buf + 5 * hw, *m0 = buf + 6 * hw, *m1 = buf + 7 * hw, *m2 = buf + 8 * hw, *m3 = buf + 9 * hw,
              *b0 = buf + 10 * hw, *b1 = buf + 11 * hw;
#pragma omp for schedule(static)
        for (int p = 0; p < planes; ++p) {
            const float* a = img1 + (size_t)p * hw;
            const float* b = img2 + (size_t)p * hw;
            blur2d(a, H, W, win2d, ws, mu1);
            blur2d(b, H, W, win2d, ws, mu2);
            for (size_t i = 0; i < hw; ++i) tmp[i] = a[i] * a[i];
            blur2d(tmp, H, W, win2d, ws, e11);
            for (size_t i = 0; i < hw; ++i) tmp[i] = b[i] * b[i];
            blur2d(tmp, H, W, win2d, ws, e22);
            for (size_t i = 0; i < hw; ++i) tmp[i] = a[i] * b[i];
            blur2d(tmp, H, W, win2d, ws, e12);
            double s = 0.0;
            const int n = p / C;
            float g = 1.f;
            if (grad_out) g = size_average ? grad_out[0] : grad_out[n];
            g /= size_average ? (float)((double)N * C * hw) : (float)((double)C * hw);
            for (size_t i = 0; i < hw; ++i) {
                const float u1 = mu1[i], u2 = mu2[i];
                const float s11 = e11[i] - u1 * u1, s22 = e22[i] - u2 * u2, s12 = e12[i] - u1 * u2;
                const float A1 = 2.f * u1 * u2 + C1, A2 = 2.f * s12 + C2;
                const float B1 = u1 * u1 + u2 * u2 + C1, B2 = s11 + s22 + C2;
                const float S = (A1 * A2) / (B1 * B2);
                s += S;
                if (d1 || d2) {
                    const float k = 2.f * (A2 - A1) / (B1 * B2), e = 2.f * S * (1.f / B1 - 1.f / B2);
                    m0[i] = g * (u2 * k - u1 * e);
                    m3[i] = g * (u1 * k - u2 * e);
                    m1[i] = g * (-S / B2);
                    m2[i] = g * (2.f * A1 / (B1 * B2));
                }
            }
            psum[p] = s;
            if (d1 || d2) {
                blur2d(m1, H, W, win2d, ws, b0);
                blur2d(m2, H, W, win2d, ws, b1);
                if (d1) {
                    blur2d(m0, H, W, win2d, ws, tmp);
                    float* o = d1 + (size_t)p * hw;
                    for (size_t i = 0; i < hw; ++i) o[i] = tmp[i] + 2.f * a[i] * b0[i] + b[i] * b1[i];
                }
                if (d2) {
                    blur2d(m3, H, W, win2d, ws, tmp);
                    float* o = d2 + (size_t)p * hw;
                    for (size_t i = 0; i < hw; ++i) o[i] = tmp[i] + 2.f * b[i] * b0[i] + a[i] * b1[i];
                }
            }
        }
        free(buf);
    }
    if (size_average) {
        double t = 0.0;
        for (int p = 0; p < planes; ++p) t += psum[p];
        out[0] = (float)(t / ((double)planes * hw));
    } else {
        for (int n = 0; n < N; ++n) {
            double t = 0.0;
            for (int c = 0; c < C; ++c) t += psum[n * C + c];
            out[n] = (float)(t / ((double)C * hw));
        }
    }
    free(psum);
}
