// b1 AND-popc mma.sync probe (north_star: "the b1 AND-popc mma.sync path adopted only if ncu shows it beating __popc").
// mma.sync m16n8k256 b1 and.popc: A 16x256 bits (row-major), B 256x8 bits (col-major), C 16x8 s32 = 128 pair-popc256 per mma.
// Hamming(a,b) = popc(a) + popc(b) - 2 popc(a & b), so one mma yields 128 descriptor-pair distances.
// Prints throughput for 1, 2 and 4 independent accumulator sets per warp (ILP) at 64 warps/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o b1_mma_probe b1_mma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
template <int ILP>
__global__ void k_mma(const unsigned* A, const unsigned* B, int* C, int iters) {
    int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    unsigned b0 = B[g * 8 + t], b1 = B[g * 8 + t + 4];
    int c[ILP][4];
#pragma unroll
    for (int j = 0; j < ILP; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j)
            asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s0 += c[j][0]; s1 += c[j][1]; s2 += c[j][2]; s3 += c[j][3]; }
    if (w == 0) { C[g * 8 + 2 * t] = s0 / ILP; C[g * 8 + 2 * t + 1] = s1 / ILP; C[(g + 8) * 8 + 2 * t] = s2 / ILP; C[(g + 8) * 8 + 2 * t + 1] = s3 / ILP; }
    else if (s0 == 0x7fffffff) C[0] = s0;
}
template <int ILP>
static void run(const unsigned* dA, const unsigned* dB, int* dC) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        int iters = 20000 / ILP, blocks = 148 * 8, threads = 256;
        cudaEventRecord(e0); k_mma<ILP><<<blocks, threads>>>(dA, dB, dC, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = (double)blocks * (threads / 32) * iters * ILP;
        printf("ILP %d: %.3f ms: %.3e mma/s, %.3e pair-popc256/s (64 warps/SM)\n", ILP, ms, mmas / (ms * 1e-3), mmas * 128 / (ms * 1e-3));
    }
}
int main() {
    unsigned hA[16 * 8], hB[8 * 8]; int hC[16 * 8];
    srand(1); for (auto& x : hA) x = rand() * 65537u + rand(); for (auto& x : hB) x = rand() * 65537u + rand();
    unsigned *dA, *dB; int* dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    k_mma<1><<<1, 32>>>(dA, dB, dC, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) { int s = 0; for (int k = 0; k < 8; ++k) s += __builtin_popcount(hA[r * 8 + k] & hB[c * 8 + k]); if (s != hC[r * 8 + c]) ++bad; }
    printf("mismatches: %d\n", bad);
    run<1>(dA, dB, dC); run<2>(dA, dB, dC); run<4>(dA, dB, dC);
    return 0;
}
