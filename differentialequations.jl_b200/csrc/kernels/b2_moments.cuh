// b2_moments.cuh -- on-device ensemble reduction: per-(save point, component) sum and
// sum of squares over the successful (Success or Terminated) trajectories of an ensemble, so that summary statistics
// (SciMLBase.EnsembleAnalysis timestep_mean / timestep_meanvar, /root/reference/test/qa/qa.jl:211) never ship
// the [N][n_save][n_state] output to the host.  HBM-bound streaming read of out_u (once), coalesced: thread x
// owns one column of the trajectory row, blockIdx.y strides over trajectories; double accumulators, one
// atomicAdd per (block, column).  B2M_F64 selects the element type.
#pragma once
#if B2M_F64
typedef double b2m_real;
#else
typedef float b2m_real;
#endif

extern "C" __global__ void __launch_bounds__(128) b2_moments_kernel(const b2m_real* __restrict__ out_u,
                                                                     const int* __restrict__ retcode, long long N,
                                                                     int row_len, double* __restrict__ sum,
                                                                     double* __restrict__ sumsq,
                                                                     unsigned long long* __restrict__ count) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0, q = 0.0;
    unsigned long long c = 0;
    for (long long i = blockIdx.y; i < N; i += gridDim.y) {
        // successful_retcode: Success (1) and Terminated (2, a callback called terminate!; the kernel filled the remaining
        // save slots with the terminal state) -- what upstream's EnsembleSummary / timestep_meanvar include.  Block-uniform.
        const int rc = __ldg(retcode + i);
        if (rc != 1 && rc != 2) continue;
        if (col < row_len) {
            const double v = (double)__ldg(out_u + i * (long long)row_len + col);
            s += v;
            q = fma(v, v, q);
        }
        c++;
    }
    if (col < row_len) {
        atomicAdd(sum + col, s);
        atomicAdd(sumsq + col, q);
    }
    if (col == 0) atomicAdd(count, c);
}
