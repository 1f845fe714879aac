// Throughput of the integer multiply-add forms on sm_100a, per SM sub-partition (SMSP): IMAD, IDP.4A (dp4a), PRMT, and the
// IMAD + PRMT mix of a byte-wise dot product.  One CTA of 32 warps per SM (8 warps per SMSP), 8 independent chains per thread.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_pipe_bench int_pipe_bench.cu ; run: ./int_pipe_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned *out, unsigned seed, int iters, long long *cycles)
{
    unsigned a[8], b = seed + threadIdx.x, c = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = a[i] * b + c;                                                     // IMAD
            if (MODE == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));      // IDP.4A
            if (MODE == 2) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));          // PRMT
            if (MODE == 3) {                                                                        // PRMT then IMAD (byte MAC)
                unsigned e;
                asm volatile("prmt.b32 %0, %1, %2, 0x4440;" : "=r"(e) : "r"(b + i), "r"(0u));
                a[i] = e * c + a[i];
            }
            if (MODE == 4) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 5) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int per_iter)
{
    unsigned *out;
    long long *cyc, h[4];
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    k<MODE><<<148, 1024>>>(out, 12345u, 16, cyc);
    k<MODE><<<148, 1024>>>(out, 12345u, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    // warp instructions per SMSP = 8 warps x iters x 8 x per_iter
    const double wi = 8.0 * iters * 8 * per_iter;
    printf("%-28s %8lld cycles  -> %.2f cycles per warp instruction per SMSP (%s)\n", name, h[0], h[0] / wi, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0>("IMAD", 1);
    run<1>("IDP.4A u8.u8", 1);
    run<4>("IDP.4A u8.s8", 1);
    run<5>("IDP.2A", 1);
    run<2>("PRMT", 1);
    run<3>("PRMT + IMAD", 2);
    return 0;
}
