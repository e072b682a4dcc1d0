// Microbenchmark: shared-memory wavefronts of the K* generators' 16-byte stores into the UMMA canonical (no-swizzle, K-major)
// A tile of observe_tf32_kernel: byte address(row, kcm) = (row >> 3) * 512 + kcm * 128 + (row & 7) * 16.
// ncu of the real kernel counts 1.8x the ideal number of store wavefronts (profiles/ncu_observe_f16x2_kernel_r02.txt);
// this isolates the address pattern from the tensor core / TMA traffic.  Patterns (one warp-wide STS.128 each):
//   0  lane -> row (32 consecutive rows, one kcm)             the kernel's mapping
//   1  lane * 16 bytes, fully contiguous                       reference: 4 wavefronts
//   2  lane -> (row & 7, kcm = lane >> 3): one 8-row group, all four K core matrices (512 contiguous bytes)
//   3  pattern 0 as two STS.64
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sts_pattern sts_pattern.cu ; run under
// ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum
#include <cstdio>
#include <cuda_runtime.h>
template <int PAT>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    __shared__ __align__(1024) unsigned char A[4][8192];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = (warp & 3) * 32 + lane, khalf = warp >> 2;
    uint4 v = make_uint4(tid, tid + 1, tid + 2, tid + 3);
    for (int it = 0; it < iters; it++) {
        unsigned char* base = A[it & 3];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int kcm = khalf * 2 + q;
            if (PAT == 0) *reinterpret_cast<uint4*>(base + (row >> 3) * 512 + kcm * 128 + (row & 7) * 16) = v;
            if (PAT == 1) *reinterpret_cast<uint4*>(base + (warp * 2 + q) * 512 + lane * 16) = v;
            if (PAT == 2) *reinterpret_cast<uint4*>(base + ((warp * 2 + q) & 15) * 512 + (lane >> 3) * 128 + (lane & 7) * 16) = v;
            if (PAT == 3) {
                unsigned char* p = base + (row >> 3) * 512 + kcm * 128 + (row & 7) * 16;
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v.x), "r"(v.y) : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p + 8)), "r"(v.z), "r"(v.w) : "memory");
            }
        }
        v.x += it;
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = reinterpret_cast<float*>(A[0])[iters & 1023];
}
int main() {
    float* out;
    cudaMalloc(&out, 4096);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float ms[4];
#define RUN(P)                                  \
    k<P><<<148, 256>>>(out, iters);             \
    cudaEventRecord(e0);                        \
    k<P><<<148, 256>>>(out, iters);             \
    cudaEventRecord(e1);                        \
    cudaEventSynchronize(e1);                   \
    cudaEventElapsedTime(&ms[P], e0, e1);
    RUN(0) RUN(1) RUN(2) RUN(3)
    for (int p = 0; p < 4; p++) printf("pattern %d: %.3f ms, %.2f clk per warp-wide 16-byte store per SM (ideal 4)\n", p, ms[p],
                                       ms[p] * 1e-3 * 1.965e9 / (iters * 2.0 * 8));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
