// Micro-benchmark of grid-wide barrier variants on one GPU (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

struct __align__(128) Slot { double a, b, c; int pos, aux; unsigned flag; unsigned pad[23]; };

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void fence_acq() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// mode 0: atomic counter, __threadfence; 1: red.release + ld.acquire counter; 2: flags padded (thread i polls i)
// 3: cg grid.sync; 4: flags, polling with ld.acquire (no fence); 5: flags, warp 0 only polls (5 flags per lane)
// 6: counter, no fences at all (lower bound); 7: mode 2 + each thread writes one double before (store drain)
__global__ void __launch_bounds__(1024, 1) k_bar(int mode, int iters, unsigned *counter, Slot *slots, double *junk, long long *cyc)
{
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, tid = threadIdx.x, cta = blockIdx.x;
    unsigned epoch = 0, seq = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (mode == 0) {
            __syncthreads(); epoch += G;
            if (tid == 0) { __threadfence(); atomicAdd(counter, 1u); while (*((volatile unsigned *)counter) < epoch) {} __threadfence(); }
            __syncthreads();
        } else if (mode == 1) {
            __syncthreads(); epoch += G;
            if (tid == 0) { red_release(counter); while (ld_acquire(counter) < epoch) {} }
            __syncthreads();
        } else if (mode == 2 || mode == 7) {
            if (mode == 7) junk[(size_t)(cta * 1024 + tid) * 16 + (it & 15)] = it;
            __syncthreads(); seq++;
            Slot *ring = slots + (seq & 3) * 160;
            if (tid == 0) st_release(&ring[cta].flag, seq);
            if (tid < G) { while (ld_relaxed(&ring[tid].flag) < seq) {} fence_acq(); }
            __syncthreads();
        } else if (mode == 3) {
            grid.sync();
        } else if (mode == 4) {
            __syncthreads(); seq++;
            Slot *ring = slots + (seq & 3) * 160;
            if (tid == 0) st_release(&ring[cta].flag, seq);
            if (tid < G) { while (ld_acquire(&ring[tid].flag) < seq) {} }
            __syncthreads();
        } else if (mode == 5) {
            __syncthreads(); seq++;
            Slot *ring = slots + (seq & 3) * 160;
            if (tid == 0) st_release(&ring[cta].flag, seq);
            if (tid < 32) { for (int i = tid; i < G; i += 32) while (ld_relaxed(&ring[i].flag) < seq) {} fence_acq(); }
            __syncthreads();
        } else if (mode == 6) {
            __syncthreads(); epoch += G;
            if (tid == 0) { atomicAdd(counter, 1u); while (*((volatile unsigned *)counter) < epoch) {} }
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (cta == 0 && tid == 0) cyc[0] = t1 - t0;
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    unsigned *counter; Slot *slots; long long *cyc; double *junk;
    cudaMalloc(&counter, 4); cudaMalloc(&slots, sizeof(Slot) * 160 * 4); cudaMalloc(&cyc, 8);
    cudaMalloc(&junk, (size_t)160 * 1024 * 16 * 8);
    printf("SMs %d clock %d kHz\n", sms, khz);
    const int iters = 2000;
    int grids[] = {sms, sms / 2, 32, 8, 2};
    for (int mode = 0; mode <= 7; mode++)
        for (int gi = 0; gi < 5; gi++) {
            int G = grids[gi];
            cudaMemset(counter, 0, 4); cudaMemset(slots, 0, sizeof(Slot) * 160 * 4);
            int it = iters;
            void *args[] = {&mode, &it, &counter, &slots, &junk, &cyc};
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_bar, dim3(G), dim3(1024), args, 0, 0);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("mode %d G %3d: %.3f us/barrier (event) %.0f cycles/barrier  %s %s\n", mode, G, ms * 1000.0 / iters, (double)c / iters,
                   e == cudaSuccess ? "" : cudaGetErrorString(e), e2 == cudaSuccess ? "" : cudaGetErrorString(e2));
        }
    return 0;
}
