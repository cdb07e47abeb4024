// Microbenchmark: MUFU.EX2 issue rate for f32 / f16 / bf16 operands on one SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned *out, long long *cyc, unsigned seed)
{
    unsigned v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 8 + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
            if (MODE == 1) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %0; ex2.approx.f16 l, l; mov.b32 %0, {l,h};}" : "+r"(v[i]));
            if (MODE == 2) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %0; ex2.approx.ftz.bf16 l, l; mov.b32 %0, {l,h};}" : "+r"(v[i]));
            if (MODE == 3) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
            if (MODE == 4) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(v[i]));
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    unsigned *o; long long *c, h;
    cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 8);
    const char *names[] = {"ex2.f32", "ex2.f16", "ex2.bf16", "ex2.f16x2", "tanh.f32"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 5; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<1, 32 * warps>>>(o, c, 1);
                if (m == 1) k<1><<<1, 32 * warps>>>(o, c, 1);
                if (m == 2) k<2><<<1, 32 * warps>>>(o, c, 1);
                if (m == 3) k<3><<<1, 32 * warps>>>(o, c, 1);
                if (m == 4) k<4><<<1, 32 * warps>>>(o, c, 1);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            // per SM sub-partition: warps/4 warps x 2048 instructions each
            printf("%-10s warps/SMSP=%d  cycles=%lld  cycles per warp-instruction per SMSP=%.2f\n", names[m],
                   warps / 4, h, (double)h / (2048.0 * warps / 4));
        }
    return 0;
}
