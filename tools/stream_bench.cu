// Micro-benchmark (development aid): y = T v over a k x k fp64 matrix, 148 CTAs x 1024 threads,
// (a) column-major T, every CTA owns ~k/148 rows of every column (strided 350-byte pieces);
// (b) panel-major T (16-row panels, each panel contiguous), every CTA owns whole panels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k_colmajor(const double *T, size_t ldt, int k, const double *v, double *y)
{
    __shared__ double red[32][65];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RPC = (((k + G - 1) / G) + 3) & ~3;
    const int q0 = min(k, cta * RPC), q1 = min(k, q0 + RPC);
    if (q1 <= q0) return;
    const int ba = q0 + 2 * lane;
    const bool in0 = ba < q1;
    const double *Tb = T + (in0 ? ba : q0);
    double ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0;
    int e = warp;
    if (in0)
        for (; e + 7 * 32 < k; e += 8 * 32) {
            double2 t[8];
#pragma unroll
            for (int x = 0; x < 8; x++) t[x] = __ldcg((const double2 *)(Tb + (size_t)(e + x * 32) * ldt));
#pragma unroll
            for (int x = 0; x < 8; x += 2) {
                double v0 = v[e + x * 32], v1 = v[e + (x + 1) * 32];
                ax0 += t[x].x * v0; ay0 += t[x].y * v0; ax1 += t[x + 1].x * v1; ay1 += t[x + 1].y * v1;
            }
        }
    red[warp][2 * lane] = ax0 + ax1; red[warp][2 * lane + 1] = ay0 + ay1;
    __syncthreads();
    if (tid < q1 - q0) { double s = 0; for (int w = 0; w < 32; w++) s += red[w][tid]; y[q0 + tid] = s; }
}

// panel p: 16 rows x k columns contiguous: element (r, c) at p*16*ldt + c*16 + r
__global__ void __launch_bounds__(1024, 1) k_panel(const double *T, size_t ldt, int k, const double *v, double *y)
{
    __shared__ double red[32][17];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int np = (k + 15) >> 4;
    for (int p = cta; p < np; p += G) {
        const double *Tp = T + (size_t)p * 16 * ldt;
        const int r2 = lane & 7, sub = lane >> 3;          // 8 lanes x double2 = 16 rows, 4 columns per warp instruction
        double ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0;
        int e = sub + 4 * warp;                             // column
        for (; e + 7 * 128 < k; e += 8 * 128) {
            double2 t[8];
#pragma unroll
            for (int x = 0; x < 8; x++) t[x] = __ldcg((const double2 *)(Tp + (size_t)(e + x * 128) * 16 + 2 * r2));
#pragma unroll
            for (int x = 0; x < 8; x += 2) {
                double v0 = v[e + x * 128], v1 = v[e + (x + 1) * 128];
                ax0 += t[x].x * v0; ay0 += t[x].y * v0; ax1 += t[x + 1].x * v1; ay1 += t[x + 1].y * v1;
            }
        }
        double ax = ax0 + ax1, ay = ay0 + ay1;
        ax += __shfl_xor_sync(0xffffffffu, ax, 8); ay += __shfl_xor_sync(0xffffffffu, ay, 8);
        ax += __shfl_xor_sync(0xffffffffu, ax, 16); ay += __shfl_xor_sync(0xffffffffu, ay, 16);
        __syncthreads();
        if (sub == 0) { red[warp][2 * r2] = ax; red[warp][2 * r2 + 1] = ay; }
        __syncthreads();
        if (tid < 16) { double s = 0; for (int w = 0; w < 32; w++) s += red[w][tid]; y[p * 16 + tid] = s; }
    }
}

int main()
{
    const int ks[] = {2560, 4096, 6353};
    const size_t ldt = 16384;
    double *T, *v, *y;
    cudaMalloc(&T, ldt * ldt * 8); cudaMalloc(&v, ldt * 8); cudaMalloc(&y, ldt * 8);
    cudaMemset(T, 0, ldt * ldt * 8); cudaMemset(v, 0, ldt * 8);
    char *flush; cudaMalloc(&flush, 256 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k : ks)
        for (int mode = 0; mode < 2; mode++) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; rep++) {
                cudaMemset(flush, rep, 256 << 20);
                cudaEventRecord(e0);
                if (mode == 0) k_colmajor<<<148, 1024>>>(T, ldt, k, v, y); else k_panel<<<148, 1024>>>(T, ldt, k, v, y);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("k %5d %-12s %8.1f us  %7.1f GB/s  %s\n", k, mode ? "panel-major" : "column-major", best * 1000, 8.0 * k * k / best / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
