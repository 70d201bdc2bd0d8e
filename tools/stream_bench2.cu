// Micro-benchmark (development aid), second generation: the dense stream y = T v (8 k^2 bytes) as
//   (a) register-staged 16-byte loads, column-major T (the engine's round-1/2 path)
//   (b) 2-D tensor-map TMA: every CTA owns RPC rows, one cp.async.bulk.tensor.2d brings an RPC x BC box
//       (BC columns of its row range) into a ring of shared-memory stages, full/empty mbarriers
//   (c) ceiling: every CTA reads one contiguous piece of the same size with 16-byte loads
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench2 stream_bench2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

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

// contiguous ceiling: CTA reads RPC*k doubles from one contiguous piece
__global__ void __launch_bounds__(1024, 1) k_contig(const double *T, size_t ldt, int k, const double *v, double *y)
{
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    const int RPC = (((k + G - 1) / G) + 3) & ~3;
    const int q0 = min(k, cta * RPC), q1 = min(k, q0 + RPC);
    if (q1 <= q0) return;
    const size_t n2 = (size_t)(q1 - q0) * k / 2;
    const double2 *p = (const double2 *)(T + (size_t)q0 * ldt);     // not the same bytes, the same amount
    double a0 = 0, a1 = 0;
    size_t e = tid;
    for (; e + 7 * 1024 < n2; e += 8 * 1024) {
        double2 t[8];
#pragma unroll
        for (int x = 0; x < 8; x++) t[x] = __ldcg(p + e + x * 1024);
#pragma unroll
        for (int x = 0; x < 8; x++) { a0 += t[x].x; a1 += t[x].y; }
    }
    if (a0 + a1 == 123.456) y[q0] = a0;
}

// (b) 2-D tensor-map TMA.  stage layout [BC][BR] doubles (rows of the box contiguous)
template <int BC, int ST>
__global__ void __launch_bounds__(1024, 1) k_tma2d(const CUtensorMap *maps, int k, const double *v, double *y)
{
    extern __shared__ __align__(128) double dyn[];
    __shared__ __align__(8) unsigned long long full_bar[ST], empty_bar[ST];
    __shared__ double red[32][65];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RPC = (((k + G - 1) / G) + 3) & ~3;               // <= 64
    const int q0 = min(k, cta * RPC), q1 = min(k, q0 + RPC);
    const int rows = q1 - q0;
    if (rows <= 0) return;
    const int BR = (rows + 3) & ~3;                              // box height: map BR/4 - 1
    const CUtensorMap *map = maps + (BR / 4 - 1);
    double *vs = dyn;                                            // [k rounded to 16]
    double *stage = dyn + ((k + 15) & ~15);                      // ST x BC x BR (128-byte aligned: BR*BC*8 multiple of 1024)
    for (int e = tid; e < k; e += 1024) vs[e] = v[e];
    const int ntile = (k + BC - 1) / BC;
    const unsigned int stage_bytes = (unsigned int)BR * BC * 8u;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(&full_bar[s], 1u); mbar_init(&empty_bar[s], 32u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < min(ST, ntile); t++) {
            mbar_expect_tx(&full_bar[t], stage_bytes);
            tma_2d(stage + (size_t)t * BC * BR, map, q0, t * BC, &full_bar[t]);
        }
    }
    // consumers: thread = (row pair rp = tid & 31, column group cg = tid >> 5): a warp per column group, lanes = row pairs
    const int rp = lane, cg = warp;
    const bool act = 2 * rp < rows;
    double ax = 0.0, ay = 0.0;
    for (int tile = 0; tile < ntile; tile++) {
        const int s = tile % ST;
        const unsigned int par = (unsigned int)((tile / ST) & 1);
        mbar_wait(&full_bar[s], par);
        const int c0 = tile * BC;
        const double *src = stage + (size_t)s * BC * BR + 2 * rp;
        if (act) {
#pragma unroll
            for (int c = cg; c < BC; c += 32) {
                if (c0 + c < k) {
                    const double2 t = *(const double2 *)(src + (size_t)c * BR);
                    const double vv = vs[c0 + c];
                    ax += t.x * vv; ay += t.y * vv;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (tid == 0 && tile + ST < ntile) {
            mbar_wait(&empty_bar[s], par);
            mbar_expect_tx(&full_bar[s], stage_bytes);
            tma_2d(stage + (size_t)s * BC * BR, map, q0, (tile + ST) * BC, &full_bar[s]);
        }
    }
    red[cg][2 * rp] = ax; red[cg][2 * rp + 1] = ay;
    __syncthreads();
    if (tid < rows) { double s = 0; for (int w = 0; w < 32; w++) s += red[w][tid]; y[q0 + tid] = s; }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BC, int ST>
static void run_tma(const CUtensorMap *dmaps, int k, const double *v, double *y, char *flush, cudaEvent_t e0, cudaEvent_t e1, const double *yref, double *yh)
{
    const int RPC = (((k + 147) / 148) + 3) & ~3;
    if (RPC > 64) { printf("k %5d tma2d BC=%d: RPC %d > 64, skipped\n", k, BC, RPC); return; }
    const size_t smem = (size_t)(((k + 15) & ~15) + ST * BC * 64) * 8;
    cudaFuncSetAttribute(k_tma2d<BC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e9f;
    for (int rep = 0; rep < 6; rep++) {
        cudaMemset(flush, rep, 256 << 20);
        cudaEventRecord(e0);
        k_tma2d<BC, ST><<<148, 1024, smem>>>(dmaps, k, v, y);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaMemcpy(yh, y, (size_t)k * 8, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < k; i++) { double d = yh[i] - yref[i]; if (d < 0) d = -d; if (d > err) err = d; }
    printf("k %5d tma2d BC=%2d ST=%d smem %6zu  %8.1f us  %7.1f GB/s  maxerr %.2e %s\n", k, BC, ST, smem, best * 1000, 8.0 * k * k / best / 1e6, err,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int ks[] = {1024, 2560, 4096, 6353, 9000};
    const size_t ldt = 16384;
    double *T, *v, *y;
    cudaMalloc(&T, ldt * ldt * 8); cudaMalloc(&v, ldt * 8); cudaMalloc(&y, ldt * 8);
    // T[r, c] = ((r * 7 + c * 3) % 11) - 5, v[c] = (c % 5) - 2: exact in fp64
    {
        double *h = (double *)malloc(ldt * 8);
        for (size_t c = 0; c < ldt; c++) {
            for (size_t r = 0; r < ldt; r++) h[r] = (double)((r * 7 + c * 3) % 11) - 5.0;
            cudaMemcpy(T + c * ldt, h, ldt * 8, cudaMemcpyHostToDevice);
        }
        for (size_t c = 0; c < ldt; c++) h[c] = (double)(c % 5) - 2.0;
        cudaMemcpy(v, h, ldt * 8, cudaMemcpyHostToDevice);
        free(h);
    }
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap hmaps[3][16];
    const int bcs[3] = {16, 32, 64};
    for (int b = 0; b < 3; b++)
        for (int i = 0; i < 16; i++) {
            cuuint64_t dims[2] = {ldt, ldt}, strides[1] = {ldt * 8};
            cuuint32_t box[2] = {(cuuint32_t)(4 * (i + 1)), (cuuint32_t)bcs[b]}, estr[2] = {1, 1};
            CUresult r = encode(&hmaps[b][i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, T, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d (box %d x %d)\n", (int)r, 4 * (i + 1), bcs[b]); return 1; }
        }
    CUtensorMap *dmaps;
    cudaMalloc(&dmaps, sizeof(hmaps));
    cudaMemcpy(dmaps, hmaps, sizeof(hmaps), cudaMemcpyHostToDevice);
    char *flush; cudaMalloc(&flush, 256 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double *yref = (double *)malloc(ldt * 8), *yh = (double *)malloc(ldt * 8);
    for (int k : ks) {
        for (int mode = 0; mode < 2; mode++) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; rep++) {
                cudaMemset(flush, rep, 256 << 20);
                cudaEventRecord(e0);
                if (mode == 0) k_colmajor<<<148, 1024>>>(T, ldt, k, v, y); else k_contig<<<148, 1024>>>(T, ldt, k, v, y);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            if (mode == 0) cudaMemcpy(yref, y, (size_t)k * 8, cudaMemcpyDeviceToHost);
            printf("k %5d %-22s %8.1f us  %7.1f GB/s  %s\n", k, mode ? "contiguous (ceiling)" : "column-major LDG.128", best * 1000,
                   8.0 * k * k / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        run_tma<16, 8>(dmaps + 0, k, v, y, flush, e0, e1, yref, yh);
        run_tma<32, 4>(dmaps + 16, k, v, y, flush, e0, e1, yref, yh);
        run_tma<32, 6>(dmaps + 16, k, v, y, flush, e0, e1, yref, yh);
        run_tma<64, 3>(dmaps + 32, k, v, y, flush, e0, e1, yref, yh);
    }
    return 0;
}
