// fma_peak.cu -- FP32 / FP64 FMA-issue microbenchmark: the measured roofline denominator the
// driver does not provide (MEASURED_PEAKS.json has only HBM and bf16 GEMM).
// 16 independent FMA chains per thread, register operands, every SM filled with 2048 threads.
#include <cuda_runtime.h>
#include <cstdio>

template <typename T>
__global__ void __launch_bounds__(256) fma_chain(T* out, T a0, T b0, int iters) {
    T x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a0 + (T)(threadIdx.x + i);
    const T b = b0, c = a0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fma(x[i], b, c);
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == (T)12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chain live
}

template <typename T>
static double run(int device, int iters) {
    cudaSetDevice(device);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    T* out = nullptr;
    cudaMalloc(&out, sizeof(T) * 256 * sms * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 8;
    fma_chain<T><<<grid, 256>>>(out, (T)1.0000001, (T)0.9999999, iters / 8 + 1);  // warm-up
    cudaDeviceSynchronize();
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        fma_chain<T><<<grid, 256>>>(out, (T)1.0000001, (T)0.9999999, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16 * 8 * (double)iters * 256.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

extern "C" double b200_fma_peak_tflops(int device, int is_f64, int iters) {
    return is_f64 ? run<double>(device, iters) : run<float>(device, iters);
}
