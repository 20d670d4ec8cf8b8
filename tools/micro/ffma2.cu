// Micro-benchmark: issue / pipe throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 b = make_float2(s, s * 0.5f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {   // 16 scalar FFMA (8 independent chains x 2)
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b.x), "f"(b.y));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(b.x), "f"(b.y));
                } else if (MODE == 1) {   // 8 FFMA2
                    a[i] = ffma2(a[i], b, b);
                } else {   // mixed: 1 FFMA2 + 2 FFMA
                    a[i] = ffma2(a[i], b, b);
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 7].x) : "f"(b.x), "f"(b.y));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 7].y) : "f"(b.x), "f"(b.y));
                }
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&clk, 8 * 148);
    const int iters = 2000;
    for (int nt : {32, 128, 256, 512, 1024}) {
        long long h[3];
        for (int m = 0; m < 3; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<1, nt>>>(out, clk, iters, 1.0001f);
                if (m == 1) k<1><<<1, nt>>>(out, clk, iters, 1.0001f);
                if (m == 2) k<2><<<1, nt>>>(out, clk, iters, 1.0001f);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h[m], clk, 8, cudaMemcpyDeviceToHost);
        }
        const double warps_per_smsp = nt / 32.0 / 4.0;
        // per iteration and warp: MODE 0: 64 FFMA; MODE 1: 32 FFMA2 (= 64 FMA lanes-ops x2); MODE 2: 32 FFMA2 + 64 FFMA
        printf("threads %4d (%.2f warps/SMSP): FFMA %.2f clk/instr/SMSP | FFMA2 %.2f clk/instr/SMSP | mixed(1 FFMA2+2 FFMA) %.2f clk per triple/SMSP\n",
               nt, warps_per_smsp, h[0] / (double)iters / 64 / (warps_per_smsp < 1 ? 1 : warps_per_smsp),
               h[1] / (double)iters / 32 / (warps_per_smsp < 1 ? 1 : warps_per_smsp),
               h[2] / (double)iters / 32 / (warps_per_smsp < 1 ? 1 : warps_per_smsp));
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
