// Micro-benchmark: cost of a grid-wide barrier (one atomic arrival per CTA + acquire polling) on all SMs,
// the building block a persistent per-step kernel would use instead of dependent launches.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/grid_barrier_bench.cu -o tools/_tmp/grid_barrier_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void barrier_loop(unsigned* cnt, int iters) {
    unsigned target = 0;
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(cnt, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned* cnt;
    cudaMalloc(&cnt, 4);
    for (int threads : {256, 512}) {
        for (int per_sm : {1, 2}) {
            const int iters = 2000;
            cudaMemset(cnt, 0, 4);
            barrier_loop<<<sms * per_sm, threads>>>(cnt, 10);   // warm-up
            cudaDeviceSynchronize();
            cudaMemset(cnt, 0, 4);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            barrier_loop<<<sms * per_sm, threads>>>(cnt, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%d SMs x %d CTAs of %d threads: %.2f us per grid barrier (%s)\n", sms, per_sm, threads,
                   1e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
