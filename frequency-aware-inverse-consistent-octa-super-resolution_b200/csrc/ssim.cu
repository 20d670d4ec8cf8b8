// Gaussian-window SSIM, forward and backward, for sm_100a.
//
// Forward (replaces ssim.py:17-37: five dense 11x11 F.conv2d + ~15 pointwise kernels + mean): one
// streaming kernel.  A CTA owns a strip of TW output columns of one image plane and marches down it in
// chunks of 32 rows.  Per chunk: stage the two input row blocks (zero outside the image, i.e. the
// conv's zero padding), horizontal 11-tap pass of the five moments x1, x2, x1^2, x2^2, x1*x2 (products
// formed in registers) into a shared-memory window of 42 rows, vertical 11-tap pass, SSIM map, running
// sum.  The last 10 horizontally-blurred rows are carried to the next chunk, so no row is filtered
// twice.  The 2-D window of ssim.py:11-15 is the outer product of the 1-D Gaussian, applied separably.
// Optionally the forward also stores the 3 (or 4) derivative maps the backward needs, so the backward
// is one more streaming kernel (3-4 blurs + combine) instead of autograd through ~25 kernels.
//
// Closed-form backward (oracle/ssim_oracle.py):
//   dx1 = W*(g M0) + 2 x1 W*(g M1) + x2 W*(g M2),   dx2 = W*(g M3) + 2 x2 W*(g M1) + x1 W*(g M2)
//   M0 = dS/dmu1, M1 = dS/dE[x1^2] = dS/dE[x2^2] = -S/B2, M2 = dS/dE[x1 x2] = 2 A1/(B1 B2), M3 = dS/dmu2
#include "common.cuh"

namespace b200w {

constexpr int kWin = B200W_SSIM_MAX_WINDOW;  // 11
constexpr int kHalo = kWin / 2;              // 5
constexpr int kTW = 64;                      // output columns per CTA
constexpr int kCR = 32;                      // rows per chunk
constexpr int kPW = kTW + 16;                // staged columns: image cols [C0-8, C0+72)
constexpr int kHR = kCR + 2 * kHalo;         // 42 horizontally-blurred rows resident

struct SsimParams {
    const float* in[4];   // fwd: img1, img2 ; bwd: maps 0..3
    const float* img1;    // bwd epilogue
    const float* img2;
    const float* grad_out;
    float* out0;          // fwd: maps base (or null) ; bwd: d1
    float* out1;          // bwd: d2 (or null)
    float* partials;      // fwd: one float per CTA
    long long plane_elems;   // H*W
    long long map_stride;    // planes*H*W
    int planes, C, H, W;
    int strips, nseg, seg_rows;
    int n_maps, size_average;
    int vec_ok;           // 16 B staging copies allowed (W % 4 == 0 and every base 16 B aligned)
    float inv_count;      // 1/(N*C*H*W) or 1/(C*H*W)
    float win[kWin];
};

// Stage rows [row0, row0+nrows) x cols [C0-8, C0+72) of one plane into dst[nrows][kPW] with cp.async, zero
// outside the image (= the conv's zero padding).  Thread = one vector column, walking down the rows.
template <int V>
__device__ __forceinline__ void stage_rows(unsigned dst_s, const float* __restrict__ src, int row0, int nrows, int C0,
                                           int H, int W, int tid) {
    constexpr int NVC = kPW / V;
    constexpr int NRG = kThreads / NVC;
    if (tid >= NVC * NRG) return;
    const int cv = tid % NVC;
    const int rg = tid / NVC;
    const int gc = C0 - 8 + V * cv;
    const bool col_in = gc >= 0 && gc + V <= W;   // V == 4 needs W % 4 == 0: a chunk is entirely in or out
    unsigned dst = dst_s + (unsigned)((rg * kPW + V * cv) * 4);
    const float* p = src + (long long)(row0 + rg) * W + gc;
    const long long step = (long long)NRG * W;
    for (int r = rg; r < nrows; r += NRG, dst += NRG * kPW * 4, p += step) {
        const int gr = row0 + r;
        if (col_in && gr >= 0 && gr < H) cp_async<V>(dst, p);
        else cp_async_zero<V>(dst, src);
    }
}

// FWD: NC = 5 channels from NI = 2 staged inputs.  BWD: NC = NI = 3 or 4 staged maps.
template <bool FWD, int NC>
__global__ void __launch_bounds__(kThreads) ssim_stream_kernel(const __grid_constant__ SsimParams p) {
    constexpr int NI = FWD ? 2 : NC;
    constexpr int SB = NI * kCR * kPW;         // floats per staging buffer
    extern __shared__ __align__(16) float smem[];
    float* hbuf = smem + 2 * SB;               // [NC][kHR][kTW]   (two staging buffers [NI][kCR][kPW] first)
    __shared__ float red[kThreads / 32];
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int strip = bid % p.strips;
    bid /= p.strips;
    const int seg = bid % p.nseg;
    const int plane = bid / p.nseg;
    const int C0 = strip * kTW;
    const int S0 = seg * p.seg_rows;
    const int S1 = min(S0 + p.seg_rows, p.H);
    const int H = p.H, W = p.W;

    const float* src[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) src[i] = p.in[i] + (long long)plane * p.plane_elems;

    float g = 0.f;
    if (!FWD) g = __ldg(p.grad_out + (p.size_average ? 0 : plane / p.C)) * p.inv_count;
    float local_sum = 0.f;

    // chunk q covers horizontally-blurred rows [S0 + 32q + 5, +32); chunk -1 is the prologue [S0-5, S0+5)
    const int niter = (S1 - S0 + kCR - 1) / kCR;
    auto issue = [&](int q, int buf) {
        const int nrows = q < 0 ? 2 * kHalo : kCR;
        const int in_row0 = q < 0 ? S0 - kHalo : S0 + kCR * q + kHalo;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const unsigned d = smem_s + (unsigned)((buf * SB + i * kCR * kPW) * 4);
            if (p.vec_ok) stage_rows<4>(d, src[i], in_row0, nrows, C0, H, W, tid);
            else stage_rows<1>(d, src[i], in_row0, nrows, C0, H, W, tid);
        }
    };
    issue(-1, 0);
    cp_async_commit();

    for (int q = -1; q < niter; ++q) {
        const int buf = (q + 1) & 1;
        if (q + 1 < niter) issue(q + 1, buf ^ 1);   // prefetch the next chunk while this one is filtered
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* stage = smem + buf * SB;
        const int nrows = q < 0 ? 2 * kHalo : kCR;
        const int hrow0 = q < 0 ? 0 : 2 * kHalo;   // hbuf row receiving staged row 0

        // ---- horizontal pass: item = (row, quad of 4 output columns); 128-bit shared loads and stores
        {
            const int cq = tid % (kTW / 4);
            int r = tid / (kTW / 4);
            const float* sp0 = stage + r * kPW + 4 * cq;
            float* hp = hbuf + (hrow0 + r) * kTW + 4 * cq;
            constexpr int RSTEP = kThreads / (kTW / 4);
            for (; r < nrows; r += RSTEP, sp0 += RSTEP * kPW, hp += RSTEP * kTW) {
                float acc[NC][4];
#pragma unroll
                for (int m = 0; m < NC; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
                float v[NI][20];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const float4 t = reinterpret_cast<const float4*>(sp0 + i * kCR * kPW)[c];
                        v[i][4 * c] = t.x; v[i][4 * c + 1] = t.y; v[i][4 * c + 2] = t.z; v[i][4 * c + 3] = t.w;
                    }
                }
                // output o (0..3) of this quad is image col C0+4cq+o = staged col 4cq+o+8; its window is staged
                // cols 4cq+o+3 .. 4cq+o+13, i.e. v[o+3+d], d = 0..10
#pragma unroll
                for (int i = 3; i < 17; ++i) {
                    float ch[NC];
                    if (FWD) {
                        const float x1 = v[0][i], x2 = v[1][i];
                        ch[0] = x1; ch[1] = x2; ch[2] = x1 * x1; ch[3] = x2 * x2; ch[4 % NC] = x1 * x2;
                    } else {
#pragma unroll
                        for (int m = 0; m < NC; ++m) ch[m] = v[m][i];
                    }
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int d = i - 3 - o;
                        if (d >= 0 && d < kWin) {
#pragma unroll
                            for (int m = 0; m < NC; ++m) acc[m][o] = fmaf(p.win[d], ch[m], acc[m][o]);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < NC; ++m)
                    *reinterpret_cast<float4*>(hp + m * kHR * kTW) = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
            }
        }
        __syncthreads();
        if (q < 0) continue;

        // ---- vertical pass: thread = (column, strip of 8 output rows)
        {
            const int c = tid % kTW;
            const int s = tid / kTW;  // 0..3
            float acc[NC][8];
#pragma unroll
            for (int m = 0; m < NC; ++m)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[m][i] = 0.f;
            const float* hb = hbuf + (8 * s) * kTW + c;
#pragma unroll
            for (int rr = 0; rr < 8 + kWin - 1; ++rr) {
#pragma unroll
                for (int m = 0; m < NC; ++m) {
                    const float t = hb[(m * kHR + rr) * kTW];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int d = rr - i;
                        if (d >= 0 && d < kWin) acc[m][i] = fmaf(p.win[d], t, acc[m][i]);
                    }
                }
            }
            const int col = C0 + c;
            const int row0 = S0 + kCR * q + 8 * s;
            if (col < W && row0 < S1) {
                const long long off = (long long)plane * p.plane_elems + (long long)row0 * W + col;
                if (FWD) {
                    float* m0 = p.out0 + off;
                    const long long ms = p.map_stride;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (row0 + i < S1) {
                            const float C1 = 0.0001f, C2 = 0.0009f;
                            const float mu1 = acc[0][i], mu2 = acc[1][i];
                            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
                            const float s11 = acc[2][i] - mu1_sq, s22 = acc[3][i] - mu2_sq, s12 = acc[4 % NC][i] - mu12;
                            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
                            const float B1 = mu1_sq + mu2_sq + C1, B2 = s11 + s22 + C2;
                            const float inv = __fdividef(1.f, B1 * B2);   // B1, B2 >= ~1e-4: one MUFU.RCP, 2 ulp
                            const float S = (A1 * A2) * inv;
                            local_sum += S;
                            if (p.n_maps) {
                                const float k = 2.f * (A2 - A1) * inv;            // common factor of dS/dmu
                                const float e = 2.f * S * (B2 - B1) * inv;        // 2 S (1/B1 - 1/B2)
                                m0[0] = mu2 * k - mu1 * e;                        // M0 = dS/dmu1
                                m0[ms] = -S * B1 * inv;                           // M1 = -S / B2
                                m0[2 * ms] = 2.f * A1 * inv;                      // M2
                                if (p.n_maps == 4) m0[3 * ms] = mu1 * k - mu2 * e;  // M3 = dS/dmu2
                            }
                        }
                        m0 += W;
                    }
                } else {
                    const float* x1p = p.img1 + off;
                    const float* x2p = p.img2 + off;
                    float* d1 = p.out0 + off;
                    float* d2 = (NC == 4 && p.out1) ? p.out1 + off : nullptr;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (row0 + i < S1) {
                            const float x1 = __ldg(x1p), x2 = __ldg(x2p);
                            *d1 = g * (acc[0][i] + 2.f * x1 * acc[1][i] + x2 * acc[2][i]);
                            if (NC == 4 && d2) *d2 = g * (acc[NC - 1][i] + 2.f * x2 * acc[1][i] + x1 * acc[2][i]);
                        }
                        x1p += W; x2p += W; d1 += W;
                        if (NC == 4 && d2) d2 += W;
                    }
                }
            }
        }
        __syncthreads();
        // ---- carry the last 10 blurred rows to the top of the window (ordered before the next vertical pass by
        //      the barrier that follows the next cp.async wait)
        if (q + 1 < niter) {
            constexpr int N4 = NC * 2 * kHalo * kTW / 4;
            for (int idx = tid; idx < N4; idx += kThreads) {
                const int m = idx / (2 * kHalo * kTW / 4);
                const int rem = idx - m * (2 * kHalo * kTW / 4);
                float4* base = reinterpret_cast<float4*>(hbuf + m * kHR * kTW);
                base[rem] = base[kCR * kTW / 4 + rem];
            }
        }
    }

    if (FWD) {
        float s = warp_sum(local_sum);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid < 32) {
            float t = tid < kThreads / 32 ? red[tid] : 0.f;
            t = warp_sum(t);
            if (tid == 0) p.partials[blockIdx.x] = t;
        }
    }
}

// deterministic final reduction: out[n] = (sum of the partials of sample n) * inv_count, in double
__global__ void __launch_bounds__(kThreads) ssim_finalize_kernel(const float* __restrict__ partials,
                                                                  int per_out, double inv_count,
                                                                  float* __restrict__ out) {
    __shared__ double red[kThreads / 32];
    const float* src = partials + (size_t)blockIdx.x * per_out;
    double s = 0.0;
    for (int i = threadIdx.x; i < per_out; i += kThreads) s += (double)src[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
        out[blockIdx.x] = (float)(t * inv_count);
    }
}

static void plan_grid(int planes, int H, int W, int* strips, int* nseg, int* seg_rows) {
    *strips = (W + kTW - 1) / kTW;
    const int max_seg = (H + kCR - 1) / kCR;
    const long long base = (long long)planes * *strips;
    const long long target = 148LL * 3 * 4;  // a few waves at 3 CTAs/SM
    int n = (int)((target + base - 1) / base);
    if (n < 1) n = 1;
    if (n > max_seg) n = max_seg;
    int rows = (H + n - 1) / n;
    rows = (rows + kCR - 1) / kCR * kCR;
    *seg_rows = rows;
    *nseg = (H + rows - 1) / rows;
}

template <bool FWD, int NC>
static int launch_ssim(const SsimParams& p, cudaStream_t st) {
    constexpr int NI = FWD ? 2 : NC;
    const size_t smem = sizeof(float) * (size_t)(2 * NI * kCR * kPW + NC * kHR * kTW);
    auto kern = ssim_stream_kernel<FWD, NC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_last_cuda_error(e);
    const size_t grid = (size_t)p.planes * p.nseg * p.strips;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(p);
    note_launch(FWD ? "ssim_stream_kernel<fwd>" : "ssim_stream_kernel<bwd>");
    e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

static int fill_window(SsimParams& p, const float* win, int ws) {
    if (!win || ws < 1 || ws > kWin || (ws % 2) == 0) return B200W_ERR_BAD_WINDOW;
    for (int i = 0; i < kWin; ++i) p.win[i] = 0.f;
    const int shift = (kWin - ws) / 2;  // a smaller odd window is the same "same" conv with zero outer taps
    for (int i = 0; i < ws; ++i) p.win[shift + i] = win[i];
    return B200W_OK;
}

}  // namespace b200w

using namespace b200w;

extern "C" size_t b200w_ssim_workspace_bytes(int N, int C, int H, int W) {
    if (N < 1 || C < 1 || H < 1 || W < 1) return 0;
    int strips, nseg, seg_rows;
    plan_grid(N * C, H, W, &strips, &nseg, &seg_rows);
    return sizeof(float) * (size_t)N * C * strips * nseg;
}

extern "C" int b200w_ssim_fwd_f32(const float* img1, const float* img2, int N, int C, int H, int W,
                                  const float* win, int ws, int size_average, int n_maps, float* maps, float* out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!img1 || !img2 || !out) return B200W_ERR_NULL_POINTER;
    if (N < 1 || C < 1 || H < 1 || W < 1) return B200W_ERR_BAD_SHAPE;
    if (!(n_maps == 0 || n_maps == 3 || n_maps == 4)) return B200W_ERR_BAD_SHAPE;
    if (n_maps && !maps) return B200W_ERR_NULL_POINTER;
    SsimParams p = {};
    int rc = fill_window(p, win, ws);
    if (rc) return rc;
    p.planes = N * C;
    p.C = C;
    p.H = H;
    p.W = W;
    plan_grid(p.planes, H, W, &p.strips, &p.nseg, &p.seg_rows);
    const size_t need = sizeof(float) * (size_t)p.planes * p.strips * p.nseg;
    if (!workspace || workspace_bytes < need) return B200W_ERR_WORKSPACE;
    p.in[0] = img1;
    p.in[1] = img2;
    p.out0 = maps;
    p.partials = (float*)workspace;
    p.plane_elems = (long long)H * W;
    p.map_stride = (long long)p.planes * H * W;
    p.n_maps = n_maps;
    p.size_average = size_average ? 1 : 0;
    p.vec_ok = ((W % 4) == 0 && aligned16(img1) && aligned16(img2)) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_ssim<true, 5>(p, st);
    if (rc) return rc;
    const int per_plane = p.strips * p.nseg;
    const int nout = size_average ? 1 : N;
    const int per_out = size_average ? p.planes * per_plane : C * per_plane;
    const double inv = size_average ? 1.0 / ((double)N * C * H * W) : 1.0 / ((double)C * H * W);
    ssim_finalize_kernel<<<nout, kThreads, 0, st>>>(p.partials, per_out, inv, out);
    note_launch("ssim_finalize_kernel");
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

extern "C" int b200w_ssim_bwd_f32(const float* img1, const float* img2, const float* maps, int n_maps,
                                  const float* grad_out, int N, int C, int H, int W, const float* win, int ws,
                                  int size_average, float* d1, float* d2, void* stream) {
    if (!img1 || !img2 || !maps || !grad_out || !d1) return B200W_ERR_NULL_POINTER;
    if (N < 1 || C < 1 || H < 1 || W < 1) return B200W_ERR_BAD_SHAPE;
    if (!(n_maps == 3 || n_maps == 4) || (d2 && n_maps != 4)) return B200W_ERR_BAD_SHAPE;
    SsimParams p = {};
    int rc = fill_window(p, win, ws);
    if (rc) return rc;
    p.planes = N * C;
    p.C = C;
    p.H = H;
    p.W = W;
    plan_grid(p.planes, H, W, &p.strips, &p.nseg, &p.seg_rows);
    p.plane_elems = (long long)H * W;
    p.map_stride = (long long)p.planes * H * W;
    for (int i = 0; i < n_maps; ++i) p.in[i] = maps + (size_t)i * p.map_stride;
    p.img1 = img1;
    p.img2 = img2;
    p.grad_out = grad_out;
    p.out0 = d1;
    p.out1 = d2;
    p.n_maps = n_maps;
    p.size_average = size_average ? 1 : 0;
    p.inv_count = size_average ? (float)(1.0 / ((double)N * C * H * W)) : (float)(1.0 / ((double)C * H * W));
    p.vec_ok = ((W % 4) == 0 && aligned16(maps)) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d2) return launch_ssim<false, 4>(p, st);
    return launch_ssim<false, 3>(p, st);
}
