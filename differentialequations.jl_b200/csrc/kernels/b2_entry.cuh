// b2_entry.cuh -- the one __global__ entry point of a compiled model: picks the stepper for
// B2_ALG (include/b200ens.h enum b200ens_alg) and runs the persistent ensemble driver.
#pragma once
#include "b2_common.cuh"

#ifndef B2_SPLIT
#define B2_SPLIT 0   // 1: one trajectory per lane of a 4-warp CTA, components split over the warps (large systems)
#endif
#if B2_ALG == 1 || B2_ALG == 2
#if B2_SPLIT
#include "b2_split.cuh"
#endif
#include "b2_erk.cuh"
#if B2_SPLIT
#include "b2_ode_driver_split.cuh"
#else
#include "b2_ode_driver.cuh"
#endif
#elif B2_ALG == 3 || B2_ALG == 4 || B2_ALG == 5 || B2_ALG == 8
#include "b2_rosenbrock.cuh"
#include "b2_ode_driver.cuh"
#elif B2_ALG == 10
#include "b2_bdf.cuh"
#include "b2_ode_driver.cuh"
#elif B2_ALG == 6 || B2_ALG == 7 || B2_ALG == 9
#include "b2_sde.cuh"
#ifdef B2_SDE_ADAPT
#include "b2_sde_adaptive.cuh"   // adaptive SRIW1 / SOSRA with rejection sampling with memory (B200ENS_MODEL_SDE_ADAPTIVE)
#endif
#else
#error "unknown B2_ALG"
#endif

#if !(B2_ALG == 6 || B2_ALG == 7 || B2_ALG == 9)
#include "b2_work.cuh"   // expected-work ordering of the trajectory queue (adaptive ODE steppers)
#endif

#if (B2_ALG == 1 || B2_ALG == 2) && !B2_SPLIT
// the common explicit case (adaptive, saveat interpolated, caller-supplied dt) folded at compile time
extern "C" __global__ void __launch_bounds__(B2_BLOCK, B2_MINBLOCKS) b2_ensemble_kernel_adaptive(const __grid_constant__ B2Args a) {
#if B2_ALG == 1
    b2_ode_driver<B2Tsit5, 1, 0, 0, 0>(a);
#else
    b2_ode_driver<B2Vern7, 1, 0, 0, 0>(a);
#endif
}
#endif

#if (B2_ALG == 1 || B2_ALG == 2) && B2_SPLIT
// split kernels: the same compile-time folding (adaptive, interpolated saveat, caller-supplied dt)
extern "C" __global__ void __launch_bounds__(B2_BLOCK, B2_MINBLOCKS) b2_ensemble_kernel_adaptive(const __grid_constant__ B2Args a) {
#if B2_ALG == 1
    b2_ode_driver_split<B2Tsit5, 1, 0, 0>(a);
#else
    b2_ode_driver_split<B2Vern7, 1, 0, 0>(a);
#endif
}
#endif

#if B2_ALG == 3 || B2_ALG == 4 || B2_ALG == 5 || B2_ALG == 8
// Rosenbrock methods: adaptive, caller-supplied dt, direct stores folded at compile time; saveat-as-tstops stays a
// run-time flag (it is the default for the Rodas family, off for Rosenbrock23)
extern "C" __global__ void __launch_bounds__(B2_BLOCK, B2_MINBLOCKS) b2_ensemble_kernel_adaptive(const __grid_constant__ B2Args a) {
#if B2_ALG == 3
    b2_ode_driver<B2Ros23, 1, -1, 0, 0>(a);
#else
    b2_ode_driver<B2Rodas, 1, -1, 0, 0>(a);
#endif
}
#endif

#if B2_ALG == 10
// FBDF: adaptive, caller-supplied dt, direct stores, saveat through the Hermite interpolant
extern "C" __global__ void __launch_bounds__(B2_BLOCK, B2_MINBLOCKS) b2_ensemble_kernel_adaptive(const __grid_constant__ B2Args a) {
    b2_ode_driver<B2Fbdf, 1, 0, 0, 0>(a);
}
#endif

extern "C" __global__ void __launch_bounds__(B2_BLOCK, B2_MINBLOCKS) b2_ensemble_kernel(const __grid_constant__ B2Args a) {
#if B2_ALG == 1 && B2_SPLIT
    b2_ode_driver_split<B2Tsit5>(a);
#elif B2_ALG == 2 && B2_SPLIT
    b2_ode_driver_split<B2Vern7>(a);
#elif B2_ALG == 1
    b2_ode_driver<B2Tsit5>(a);
#elif B2_ALG == 2
    b2_ode_driver<B2Vern7>(a);
#elif B2_ALG == 3
    b2_ode_driver<B2Ros23>(a);
#elif B2_ALG == 4 || B2_ALG == 5 || B2_ALG == 8
    b2_ode_driver<B2Rodas>(a);
#elif B2_ALG == 10
    b2_ode_driver<B2Fbdf>(a);
#elif defined(B2_SDE_ADAPT)
    b2_sde_adaptive_driver<B2_ALG>(a);
#else
    b2_sde_driver<B2_ALG>(a);
#endif
}
