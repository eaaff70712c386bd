// b2_split.cuh -- shared-memory exchange for the SPLIT kernels: one trajectory spread over the 4 warps of a CTA.
//
// Large systems do not fit one thread: Vern7 with n = 16 in Float64 keeps 14 stage vectors = 448 32-bit registers
// alive (config 5 of BASELINE.json).  The one-thread-per-trajectory kernel then either spills to local memory
// (DRAM-bound, 135 GB of spill traffic per 100k trajectories) or keeps the stage vectors in shared memory at 4 warps
// per SM (latency-bound, ~23 cycles per issued instruction) -- profiles/README.md.  In split mode a CTA of 4 warps
// integrates 32 trajectories: lane = trajectory, warp g owns components [g*NL, (g+1)*NL) of every state / stage
// vector (NL = ceil(n/4)), so a thread holds a QUARTER of the registers, nothing spills and 3x more warps are
// resident.  All per-trajectory control (t, dt, controller, accept/reject, saveat index, event search) is replicated
// in the four warps from identical data, hence bit-identical and CTA-uniform; the only communication is
//   * b2_split_rhs: publish the owned block of a stage argument, ONE barrier, read the components the owned outputs
//     need (the generated b2_rhs is inlined once per warp role and dead outputs are pruned by the compiler);
//   * b2_split_gather: the same exchange for the error-norm ratios / event data (every warp sums all n terms in the
//     oracle's order, so the norm is bit-identical to the one-thread kernel).
// Every component goes through exactly the expression tree of the one-thread kernel: results are bit-identical.
#pragma once
#include "b2_common.cuh"

#define B2_SPLIT_G 4
#define B2_NL ((B2_N + B2_SPLIT_G - 1) / B2_SPLIT_G)
#define B2_NP (B2_NL * B2_SPLIT_G)

struct B2Xchg {
    real* base;     // double-buffered exchange area in shared memory: real [2][B2_NP][32]
    int lane;       // trajectory slot = lane index (column of the exchange area)
    int phase;      // which half the next exchange uses
    int g;          // warp role (0..3), warp-uniform
    const real* pcol; // this lane's column of the parameter mirror in shared memory: real [B2_NPA][32]
};
__host__ __device__ constexpr int b2_popc_c(unsigned x) {
    int c = 0;
    while (x) {
        c += (int)(x & 1u);
        x >>= 1;
    }
    return c;
}

// Publish NL values of this warp's block, barrier, return a pointer to the lane's column of the full vector
// (component i at [i*32]).  Double-buffered: the buffer written now is next overwritten two exchanges later, after
// every warp has passed the following barrier, so one barrier per exchange is enough.
template <class T>
__device__ __forceinline__ const T* b2_split_publish(B2Xchg& xc, const T (&x)[B2_NL]) {
    T* b = reinterpret_cast<T*>(xc.base + (size_t)xc.phase * (B2_NP * 32)) + xc.lane;
    xc.phase ^= 1;
#pragma unroll
    for (int j = 0; j < B2_NL; j++) b[(xc.g * B2_NL + j) * 32] = x[j];
    __syncthreads();
    return b;
}

// The owned block of f(X, p, t) for warp role G: the generated b2_rhs is inlined once per role and the outputs other
// roles own are dead code.  ONE out-of-line copy per kernel (not one per call site): a Vern7 step with its event search
// evaluates the RHS at 18 places, and with the four role bodies inlined everywhere the loop body was 172 KB of SASS --
// ncu showed the kernel stalled on instruction fetch (no_instruction: 9 of 13 stalled warps per issue).  Arguments
// travel through shared memory (the stage argument just published, the trajectory's parameters), the result in
// registers.
struct B2Blk {
    real v[B2_NL];
};
#define B2_SPLIT_CASE(G)                                                                   \
    case G: {                                                                              \
        real X_[B2_N], F_[B2_N];                                                           \
        _Pragma("unroll") for (int i = 0; i < B2_N; i++) X_[i] = xb[i * 32];               \
        b2_rhs(F_, X_, P_, t);                                                             \
        _Pragma("unroll") for (int j = 0; j < B2_NL; j++)                                  \
            r.v[j] = (G * B2_NL + j < B2_N) ? F_[(G * B2_NL + j < B2_N) ? G * B2_NL + j : 0] : (real)0; \
    } break;

__device__ __noinline__ B2Blk b2_split_rhs_role(int g, const real* xb, const real* pb, real t) {
    B2Blk r;
    real P_[B2_NPA];
#pragma unroll
    for (int i = 0; i < B2_NPARAM; i++) P_[i] = pb[i * 32];
    switch (g) {
        B2_SPLIT_CASE(0)
        B2_SPLIT_CASE(1)
        B2_SPLIT_CASE(2)
        default:
        B2_SPLIT_CASE(3)
    }
    return r;
}

// out = the owned block of f(x_full, p, t)
template <class Alg>
__device__ __forceinline__ void b2_split_rhs(Alg& alg, real (&out)[B2_NL], const real (&x)[B2_NL], const real (&p)[B2_NPA],
                                             real t) {
    (void)p;   // the out-of-line role body reads the parameters from shared memory (alg.xc.pcol)
    const real* b = b2_split_publish(alg.xc, x);
    const B2Blk r = b2_split_rhs_role(alg.xc.g, b, alg.xc.pcol, t);
#pragma unroll
    for (int j = 0; j < B2_NL; j++) out[j] = r.v[j];
}
