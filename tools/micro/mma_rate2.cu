// Microbenchmark 2: mma.sync issue rate with DISTINCT operand registers (as in a real kernel), 3 or 6 accumulator chains,
// tf32 m16n8k8 / bf16 m16n8k16 / fp16 m16n8k16, 1 or 2 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int KIND, int NACC>
__global__ void k_mma(float* out, int iters, long long* cyc, const uint32_t* src) {
    float c[NACC][4] = {};
    uint32_t a[8][4], b[24][2];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) a[i][j] = src[(i * 4 + j) * 32 + (threadIdx.x & 31)];
    for (int i = 0; i < 24; ++i) for (int j = 0; j < 2; ++j) b[i][j] = src[1024 + (i * 2 + j) * 32 + (threadIdx.x & 31)];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int acc = (NACC == 3) ? j : (j + 3 * (s & 1));
                if (KIND == 0)
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[acc][0]), "+f"(c[acc][1]), "+f"(c[acc][2]), "+f"(c[acc][3]) : "r"(a[s][0]), "r"(a[s][1]), "r"(a[s][2]), "r"(a[s][3]), "r"(b[s * 3 + j][0]), "r"(b[s * 3 + j][1]));
                else if (KIND == 1)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[acc][0]), "+f"(c[acc][1]), "+f"(c[acc][2]), "+f"(c[acc][3]) : "r"(a[s][0]), "r"(a[s][1]), "r"(a[s][2]), "r"(a[s][3]), "r"(b[s * 3 + j][0]), "r"(b[s * 3 + j][1]));
                else
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[acc][0]), "+f"(c[acc][1]), "+f"(c[acc][2]), "+f"(c[acc][3]) : "r"(a[s][0]), "r"(a[s][1]), "r"(a[s][2]), "r"(a[s][3]), "r"(b[s * 3 + j][0]), "r"(b[s * 3 + j][1]));
            }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < NACC; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int KIND, int NACC>
void run(const char* name, float* out, long long* cyc, const uint32_t* src) {
    const int iters = 500;
    for (int warps : {4, 8, 16}) {
        long long h;
        k_mma<KIND, NACC><<<1, warps * 32>>>(out, iters, cyc, src); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (iters * 24.0);
        printf("%-6s acc=%d warps=%2d: %.2f clk per mma per warp, %.2f clk per mma per sub-partition\n", name, NACC, warps, per, per / (warps / 4));
    }
}
int main() {
    float* out; long long* cyc; uint32_t* src; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&src, 1 << 16);
    cudaMemset(src, 0x3c, 1 << 16);
    run<0, 3>("tf32", out, cyc, src); run<0, 6>("tf32", out, cyc, src);
    run<1, 3>("bf16", out, cyc, src); run<1, 6>("bf16", out, cyc, src);
    run<2, 3>("fp16", out, cyc, src);
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
