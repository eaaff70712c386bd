// aot_check.cu -- offline nvcc build of every stepper kernel for sm_100a with the Lorenz
// model (the same headers NVRTC compiles at run time).  Used by `make aot_check` and
// __graft_entry__.build() as the "does it build" gate and for -Xptxas -v / cuobjdump checks.
#ifndef AOT_ALG
#define AOT_ALG 1
#endif
#ifndef AOT_F64
#define AOT_F64 0
#endif
#define B2_F64 AOT_F64
#define B2_NSTATE 3
#define B2_NPARAM 4
#define B2_ALG AOT_ALG
#define B2_HAS_JAC 1
#define B2_HAS_TGRAD 0
#define B2_HAS_NOISE 1
#define B2_HAS_EVENT 0
#define B2_HAS_DEVENT 0
#define B2_BLOCK 128
#ifdef AOT_SPLIT
#define B2_SPLIT 1   // one trajectory per lane of a 4-warp CTA (kernels/b2_split.cuh); n = 3 -> one component per warp, one padded
#endif
#ifdef AOT_SDE_ADAPT
#define B2_SDE_ADAPT 1   // adaptive SRIW1 / SOSRA (kernels/b2_sde_adaptive.cuh)
#endif
#ifndef B2_MINBLOCKS
#define B2_MINBLOCKS 1
#endif
#include "b2_common.cuh"

__device__ __forceinline__ void b2_rhs(real* __restrict__ du, const real* __restrict__ u, const real* __restrict__ p, const real t) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
}
__device__ __forceinline__ void b2_jac(real* __restrict__ J, const real* __restrict__ u, const real* __restrict__ p, const real t) {
    J[0] = -p[0]; J[1] = p[0]; J[2] = 0;
    J[3] = p[1] - u[2]; J[4] = -1; J[5] = -u[0];
    J[6] = u[1]; J[7] = u[0]; J[8] = -p[2];
}
__device__ __forceinline__ void b2_noise(real* __restrict__ g, const real* __restrict__ u, const real* __restrict__ p, const real t) {
    g[0] = p[3]; g[1] = p[3]; g[2] = p[3];
}
#include "b2_entry.cuh"
