#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
// mma.sync m16n8k256 b1 and.popc: A 16x256 bits (row-major), B 256x8 bits (col-major), C 16x8 s32
__global__ void k_mma(const unsigned* A, const unsigned* B, int* C, int iters) {
    int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // A fragment: a0:(row g, k 0..31 of chunk t), a1:(row g+8, chunk t), a2:(row g, chunk t+4), a3:(row g+8, chunk t+4)
    unsigned a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    // B fragment: b0:(col g, chunk t), b1:(col g, chunk t+4)
    unsigned b0 = B[g * 8 + t], b1 = B[g * 8 + t + 4];
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w == 0) { C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3; }
    else if (c0 == 0x7fffffff) C[0] = c0;
}
int main() {
    unsigned hA[16 * 8], hB[8 * 8]; int hC[16 * 8];
    srand(1); for (auto& x : hA) x = rand() * 65537u + rand(); for (auto& x : hB) x = rand() * 65537u + rand();
    unsigned *dA, *dB; int* dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    k_mma<<<1, 32>>>(dA, dB, dC, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) { int s = 0; for (int k = 0; k < 8; ++k) s += __builtin_popcount(hA[r * 8 + k] & hB[c * 8 + k]); if (s != hC[r * 8 + c]) ++bad; }
    printf("mismatches: %d\n", bad);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        int iters = 20000, blocks = 148 * 8, threads = 256;
        cudaEventRecord(e0); k_mma<<<blocks, threads>>>(dA, dB, dC, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = (double)blocks * (threads / 32) * iters;
        printf("%.3f ms: %.3e mma/s, %.3e pair-popc256/s (dependent chain per warp, %d warps/SM)\n", ms, mmas / (ms * 1e-3), mmas * 128 / (ms * 1e-3), 8 * 8);
    }
    return 0;
}
