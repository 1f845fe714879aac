// Microbenchmark (B200): how fast can a CTA push zeros / shared-memory rows to global memory?
//   mode 0: st.global.v4 from registers            mode 1: cp.async.bulk S2G, ops issued by the lanes of warp 0
//   mode 2: cp.async.bulk S2G, ops spread over the warps (one lane each)   mode 3: as 2, lanes 0..3 of each warp
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_store_bench tma_store_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256) fill_kernel(float *out, size_t bytes_per_cta, int op_bytes, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    char *dst = reinterpret_cast<char *>(out) + (size_t)blockIdx.x * bytes_per_cta;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (mode == 0) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (size_t e = tid; e < bytes_per_cta / 16; e += 256) d4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    for (int e = tid; e < op_bytes / 16; e += 256) reinterpret_cast<float4 *>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int nops = (int)(bytes_per_cta / op_bytes);
    if (mode == 1) {
        if (warp == 0)
            for (int o = lane; o < nops; o += 32) bulk_s2g(dst + (size_t)o * op_bytes, smem, op_bytes);
    } else if (mode == 2) {
        if (lane == 0)
            for (int o = warp; o < nops; o += 8) bulk_s2g(dst + (size_t)o * op_bytes, smem, op_bytes);
    } else {
        if (lane < 4)
            for (int o = warp * 4 + lane; o < nops; o += 32) bulk_s2g(dst + (size_t)o * op_bytes, smem, op_bytes);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main()
{
    const size_t total = 64ull * 3 * 224 * 224 * 4;          // cfg2's gx: 38.5 MB
    const int nbuf = 6;
    std::vector<float *> bufs(nbuf);
    for (auto &b : bufs) cudaMalloc(&b, total);
    cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grids[] = {148, 592, 1184};
    const int ops[] = {896, 1792, 3584, 7168, 14336, 28672};
    for (int g : grids)
        for (int mode = 0; mode < 4; ++mode)
            for (int op : ops) {
                if (mode == 0 && op != ops[0]) continue;
                size_t per = total / g / op * op;
                if (per == 0) continue;
                for (int w = 0; w < 3; ++w) fill_kernel<<<g, 256, op>>>(bufs[w % nbuf], per, op, mode);
                cudaEventRecord(e0);
                const int reps = 30;
                for (int r = 0; r < reps; ++r) fill_kernel<<<g, 256, op>>>(bufs[r % nbuf], per, op, mode);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaError_t err = cudaGetLastError();
                printf("grid %4d mode %d op %5d B : %7.2f us per 38.5 MB fill  (%6.0f GB/s)%s\n", g, mode, op, ms * 1e3 / reps,
                       (double)per * g / (ms * 1e-3 / reps) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
            }
    return 0;
}
