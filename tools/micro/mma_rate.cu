// Microbenchmark: issue rate of legacy warp-level mma.sync (m16n8k8 tf32, m16n8k16 bf16) and of packed fp32 FMA on one SM of a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void k_mma_tf32(float* out, int iters, long long* cyc) {
    float c[6][4] = {};
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f810000u, 0x3f820000u, 0x3f830000u}, b[2] = {0x3f800000u, 0x3f900000u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_mma_bf16(float* out, int iters, long long* cyc) {
    float c[6][4] = {};
    uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f813f80u, 0x3f823f80u, 0x3f833f80u}, b[2] = {0x3f803f80u, 0x3f903f80u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2(float* out, int iters, long long* cyc) {
    float2 c[12]; for (int j = 0; j < 12; ++j) c[j] = make_float2(threadIdx.x, j);
    float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.5f, 0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 12; ++j) c[j] = __ffma2_rn(a, c[j], b);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 12; ++j) s += c[j].x + c[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma(float* out, int iters, long long* cyc) {
    float c[12]; for (int j = 0; j < 12; ++j) c[j] = threadIdx.x + j;
    float a = 1.0001f, b = 0.5f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 12; ++j) c[j] = fmaf(a, c[j], b);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 12; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        long long h;
        k_mma_tf32<<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d mma.m16n8k8.tf32 : %.2f cycles per mma per SM-subcore-warp, %.1f MAC/clk/SM\n", warps, (double)h / (iters * 6.0), 1024.0 * iters * 6 * warps / h);
        k_mma_bf16<<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d mma.m16n8k16.bf16: %.2f cycles per mma, %.1f MAC/clk/SM\n", warps, (double)h / (iters * 6.0), 2048.0 * iters * 6 * warps / h);
        k_ffma2<<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d ffma2            : %.2f cycles per instr, %.1f FMA/clk/SM\n", warps, (double)h / (iters * 12.0), 64.0 * iters * 12 * warps / h);
        k_ffma<<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d ffma             : %.2f cycles per instr, %.1f FMA/clk/SM\n", warps, (double)h / (iters * 12.0), 32.0 * iters * 12 * warps / h);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
