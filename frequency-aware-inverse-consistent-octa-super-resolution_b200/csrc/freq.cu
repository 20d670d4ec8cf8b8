// Fourier-domain Gaussian frequency split: the pointwise kernels around cuFFT.
//
// Replaces utils.py:71-117 of the reference: guais_low_pass / guais_high_pass build a rows x cols mask with two
// Python loops (one exp per pixel, on the host, every call), high_pass / low_pass do fft2 -> fftshift -> * mask ->
// ifftshift -> ifft2 -> abs on ONE image.  Here the mask is never materialised and nothing is shifted: in the
// unshifted layout the shifted index i holds frequency d = i - rows/2, so the mask at FFT bin (u, v) is
//     m = exp(-0.5 (du^2 + dv^2) / r^2),  du = u < rows - rows/2 ? u : u - rows   (same for dv)
// (1 - m for the high pass).  The mask is real and even, so the filtered spectrum of a real image stays Hermitian:
// the half spectrum of rfft2 / irfft2 is enough (half the bytes of the reference's complex fft2), batched over all
// planes.  The FFTs themselves are cuFFT (library); these kernels are the HBM-bound pointwise passes:
//   freq_mask   in place on the (planes, rows, cols/2+1) complex half spectrum: 16 B of traffic per bin
//   abs_sign    y = sign * |x| (low_pass returns -|.|, utils.py:117), optionally keeping sgn(x) for backward
//   sign_mul    backward of abs_sign: g * sign * sgn(x)
#include "common.cuh"

namespace b200w {

// flat grid-stride loop over the bins of all planes; the mask costs one expf and two small integer divisions per
// bin (an (u, v)-outer / plane-inner variant that evaluates the mask once per bin was slower: its plane-strided
// accesses and 129-bin rows used the memory system worse -- profiles/r01_notes.md)
__global__ void __launch_bounds__(256) freq_mask_kernel(float2* __restrict__ spec, int planes, int rows, int cols, int wh,
                                                        float inv2r2, int highpass) {
    const unsigned per_plane = (unsigned)rows * (unsigned)wh;
    const size_t total = (size_t)per_plane * planes;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const unsigned rem = (unsigned)(idx % per_plane);
        const int u = (int)(rem / (unsigned)wh), v = (int)(rem - (unsigned)u * (unsigned)wh);
        const int du = u < rows - rows / 2 ? u : u - rows;
        const int dv = v < cols - cols / 2 ? v : v - cols;
        float m = __expf(-(float)(du * du + dv * dv) * inv2r2);
        if (highpass) m = 1.f - m;
        float2 z = spec[idx];
        z.x *= m;
        z.y *= m;
        spec[idx] = z;
    }
}

__global__ void __launch_bounds__(256) abs_sign_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4,
                                                       const float* __restrict__ xt, float* __restrict__ yt, size_t n,
                                                       float sign) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = x[i];
        y[i] = make_float4(sign * fabsf(a.x), sign * fabsf(a.y), sign * fabsf(a.z), sign * fabsf(a.w));
    }
    if (blockIdx.x == 0)
        for (size_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) yt[i] = sign * fabsf(xt[i]);
}

__global__ void __launch_bounds__(256) sign_mul_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                       float* __restrict__ out, size_t n, float sign) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        out[i] = v > 0.f ? sign * g[i] : (v < 0.f ? -sign * g[i] : 0.f);
    }
}

static unsigned pointwise_grid(size_t n) {
    size_t g = (n + 255) / 256;
    const size_t cap = 148 * 8;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_freq_mask_c64(void* spec, int planes, int rows, int cols, float radius, int highpass, void* stream) {
    if (!spec) return B200W_ERR_NULL_POINTER;
    if (planes < 1 || rows < 1 || cols < 1 || !(radius > 0.f)) return B200W_ERR_BAD_SHAPE;
    const int wh = cols / 2 + 1;
    if ((long long)rows * wh > 0x7fffffffLL) return B200W_ERR_BAD_SHAPE;
    const size_t total = (size_t)planes * rows * wh;
    freq_mask_kernel<<<pointwise_grid(total), 256, 0, (cudaStream_t)stream>>>((float2*)spec, planes, rows, cols, wh,
                                                                              0.5f / (radius * radius), highpass ? 1 : 0);
    b200w::note_launch("freq_mask_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

extern "C" int b200w_abs_sign_f32(const float* x, float* y, size_t n, float sign, void* stream) {
    if (!x || !y) return B200W_ERR_NULL_POINTER;
    if (n == 0) return B200W_OK;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    abs_sign_kernel<<<pointwise_grid(n4 ? n4 : n), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, n4, x, y, n,
                                                                                  sign);
    b200w::note_launch("abs_sign_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}

extern "C" int b200w_sign_mul_f32(const float* g, const float* x, float* out, size_t n, float sign, void* stream) {
    if (!g || !x || !out) return B200W_ERR_NULL_POINTER;
    if (n == 0) return B200W_OK;
    sign_mul_kernel<<<pointwise_grid(n), 256, 0, (cudaStream_t)stream>>>(g, x, out, n, sign);
    b200w::note_launch("sign_mul_kernel");
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200W_OK : set_last_cuda_error(e);
}
