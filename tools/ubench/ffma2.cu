// Microbenchmark: does packed FP32 (FFMA2, sm_100) relieve issue pressure?  nvcc -arch=sm_100a -O3 ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters, unsigned m0) {
    float x[16];
    unsigned z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = a0 + threadIdx.x + i; z[i] = threadIdx.x * 7u + i; }
    const float b = b0, c = a0;
    const float2 b2 = make_float2(b0, b0), c2 = make_float2(a0, a0);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0) {  // 16 FFMA
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], b, c);
            } else if (MODE == 1) {  // 8 FFMA2 (same flops)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), b2, c2);
                    x[i] = v.x; x[i + 1] = v.y;
                }
            } else if (MODE == 2) {  // 16 FFMA + 16 ALU (LOP3)
#pragma unroll
                for (int i = 0; i < 16; i++) { x[i] = fmaf(x[i], b, c); z[i] = (z[i] ^ m0) + (z[i] >> 3); }
            } else {  // 8 FFMA2 + 16 ALU
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), b2, c2);
                    x[i] = v.x; x[i + 1] = v.y;
                    z[i] = (z[i] ^ m0) + (z[i] >> 3); z[i + 1] = (z[i + 1] ^ m0) + (z[i + 1] >> 3);
                }
            }
        }
    }
    float s = 0; unsigned q = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { s += x[i]; q ^= z[i]; }
    if (s == 12345.678f || q == 0xdeadbeefu) out[blockIdx.x * blockDim.x + threadIdx.x] = s + q;
}
template <int MODE> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4 * 256 * sms * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 1.0000001f, 0.9999999f, 64, 5u); cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, 1.0000001f, 0.9999999f, iters, 5u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fmas = 16.0 * 8 * iters * 256.0 * grid;
    printf("%-22s %.3f ms  %.1f TFLOP/s (FMA flops)\n", name, best, 2 * fmas / best / 1e9);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); run<2>("FFMA + ALU 1:1"); run<3>("FFMA2 + ALU (same work)"); return 0; }
