// Microbenchmark: fp64 tensor (DMMA, mma.sync f64) vs fp64 vector (DFMA) issue rates on sm_100a.
// Answers: (1) peak DMMA rate per shape, (2) whether DMMA and DFMA share a pipe, (3) fp64 exp cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma_16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma_1688(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma_1684(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void dmma_884(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
}

// MODE: 0=m16n8k16 1=m16n8k8 2=m16n8k4 3=m8n8k4 ; NACC independent accumulator tiles per warp
// DF = number of independent DFMA chains interleaved per MMA group (0 = none)
template <int MODE, int NACC, int DF>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double seed) {
    double c[NACC][4];
    double a[8], b[4];
    double f[DF > 0 ? DF : 1];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) c[j][i] = 0.0;
#pragma unroll
    for (int i = 0; i < (DF > 0 ? DF : 1); i++) f[i] = seed + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (MODE == 0) dmma_16816(c[j], a, b);
            if (MODE == 1) dmma_1688(c[j], a, b);
            if (MODE == 2) dmma_1684(c[j], a, b);
            if (MODE == 3) dmma_884(c[j], a, b);
        }
#pragma unroll
        for (int i = 0; i < DF; i++) f[i] = fma(f[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) s += c[j][i];
#pragma unroll
    for (int i = 0; i < (DF > 0 ? DF : 1); i++) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int NCH>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double seed) {
    double f[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) f[i] = seed + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) f[i] = fma(f[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int NCH>
__global__ void __launch_bounds__(1024) k_exp(double* out, int iters, double seed) {
    double f[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) f[i] = -(seed + i * 0.01 + threadIdx.x * 1e-3);
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) { s += exp(f[i]); f[i] -= 1e-6; }
    }
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_it(F launch, int reps = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s SMs=%d clock(kHz)=%d\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    int warps_list[] = {4, 8, 16, 32};
#define RUN_MMA(MODE, NACC, DF, FMA_PER, NAME) \
    for (int wi = 0; wi < 4; wi++) { int w = warps_list[wi]; \
        float ms = time_it([&] { k_dmma<MODE, NACC, DF><<<sms, w * 32>>>(out, iters, 1.0); }); \
        double fma_mma = (double)sms * w * iters * NACC * FMA_PER; \
        double fma_v = (double)sms * w * 32 * iters * DF; \
        printf("%-22s warps/SM=%2d  %8.3f ms  MMA %.2f TFLOP/s  + DFMA %.2f TFLOP/s  (%.1f MMA-FMA/clk/SM @1.965GHz)\n", NAME, w, ms, \
               2 * fma_mma / ms * 1e-9, 2 * fma_v / ms * 1e-9, fma_mma / (ms * 1e-3) / sms / 1.965e9); }
    RUN_MMA(0, 8, 0, 2048, "m16n8k16 x8acc")
    RUN_MMA(1, 8, 0, 1024, "m16n8k8  x8acc")
    RUN_MMA(2, 8, 0, 512, "m16n8k4  x8acc")
    RUN_MMA(3, 8, 0, 256, "m8n8k4   x8acc")
    RUN_MMA(0, 16, 0, 2048, "m16n8k16 x16acc")
    RUN_MMA(0, 8, 8, 2048, "m16n8k16 x8 + 8 DFMA")
    RUN_MMA(0, 8, 32, 2048, "m16n8k16 x8 + 32 DFMA")
    RUN_MMA(0, 8, 128, 2048, "m16n8k16 x8 + 128 DFMA")
    for (int wi = 0; wi < 4; wi++) { int w = warps_list[wi];
        float ms = time_it([&] { k_dfma<16><<<sms, w * 32>>>(out, iters, 1.0); });
        double fl = 2.0 * sms * w * 32 * (double)iters * 16;
        printf("%-22s warps/SM=%2d  %8.3f ms  %.2f TFLOP/s\n", "DFMA x16 chains", w, ms, fl / ms * 1e-9); }
    for (int wi = 0; wi < 4; wi++) { int w = warps_list[wi];
        float ms = time_it([&] { k_exp<8><<<sms, w * 32>>>(out, iters / 10, 1.0); });
        double n = (double)sms * w * 32 * (iters / 10) * 8;
        printf("%-22s warps/SM=%2d  %8.3f ms  %.2f Gexp/s  (%.1f clk/exp/SM-lane... %.2f exp/clk/SM @1.965GHz)\n", "exp(double) x8", w, ms, n / ms * 1e-6,
               0.0, n / (ms * 1e-3) / sms / 1.965e9); }
    return 0;
}
