// Microbenchmark: fp64 mma.sync m8n8k4 fed from shared memory (one LDS.64 per DMMA, as in gp_predict_kernel) at
// 8 / 12 / 16 warps per SM with 64 / 32 accumulators per lane.  Question: how much of the DMMA issue rate is lost
// with only 2 warps per SM sub-partition?    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lds_mix dmma_lds_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void __launch_bounds__(NACC > 16 ? 256 : 512) k(double* out, int iters) {
    __shared__ double B[16][260];
    for (int i = threadIdx.x; i < 16 * 260; i += blockDim.x) (&B[0][0])[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
    double acc[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; j++) acc[j][0] = acc[j][1] = 0.0;
    double a = 1.0 + lane * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
#pragma unroll
            for (int j = 0; j < NACC; j++) {
                const double b = B[k4 * 4 + c][(j % 32) * 8 + r];
                dmma(acc[j][0], acc[j][1], a, b);
            }
        }
        a += 1e-12;
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += acc[j][0] + acc[j][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
void run(int warps, const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000 * 32 / NACC;
    k<NACC><<<sms, warps * 32>>>(out, 100);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<NACC><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = 2.0 * 256 * 4.0 * NACC * iters * warps * sms;
    printf("%-28s warps/SM=%2d acc/lane=%2d  %8.3f ms  %6.2f TFLOP/s\n", name, warps, NACC, best, flops / (best * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    run<32>(8, "DMMA+LDS 8 rows x 256 cols");   // the current kernel's warp tile (64 acc doubles = 32 pairs)
    run<16>(8, "DMMA+LDS 8 rows x 128 cols");
    run<16>(12, "DMMA+LDS 8 rows x 128 cols");
    run<16>(16, "DMMA+LDS 8 rows x 128 cols");
    run<32>(4, "DMMA+LDS 8 rows x 256 cols");
    return 0;
}
