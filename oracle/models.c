/* models.c -- hand-written model functions for the oracle (TEST INFRASTRUCTURE).
 * Independent of the sympy->CUDA-C emitter, so parity tests also cover the emitter.
 *   lorenz        /root/reference test/core.jl:22-30
 *   robertson     /root/reference test/core.jl:39-46, src/DifferentialEquations.jl:14-23
 *   linear        /root/reference test/core.jl:10-13   (u' = p0*u)
 *   gbm           SURVEY.md 8(d) cfg 4:  du = mu u dt + sigma u dW
 *   lorenz_additive  Lorenz drift + additive diagonal noise g = p[3]
 *   net16         SURVEY.md 8(d) cfg 5 (builder-defined 16-species mass-action network)
 */
#include "oracle.h"
#include <string.h>

#define DEFMODELS(T, S)                                                                                   \
    static void lorenz_rhs_##S(T* du, const T* u, const T* p, T t) {                                      \
        (void)t;                                                                                          \
        du[0] = p[0] * (u[1] - u[0]);                                                                     \
        du[1] = u[0] * (p[1] - u[2]) - u[1];                                                              \
        du[2] = u[0] * u[1] - p[2] * u[2];                                                                \
    }                                                                                                     \
    static void lorenz_jac_##S(T* J, const T* u, const T* p, T t) {                                       \
        (void)t;                                                                                          \
        J[0] = -p[0]; J[1] = p[0]; J[2] = 0;                                                              \
        J[3] = p[1] - u[2]; J[4] = -1; J[5] = -u[0];                                                      \
        J[6] = u[1]; J[7] = u[0]; J[8] = -p[2];                                                           \
    }                                                                                                     \
    static void lorenz_noise_##S(T* g, const T* u, const T* p, T t) {                                     \
        (void)u; (void)t;                                                                                 \
        g[0] = p[3]; g[1] = p[3]; g[2] = p[3];                                                            \
    }                                                                                                     \
    static void robertson_rhs_##S(T* du, const T* u, const T* p, T t) {                                   \
        (void)t;                                                                                          \
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];                                                        \
        du[1] = p[0] * u[0] - p[1] * u[1] * u[1] - p[2] * u[1] * u[2];                                    \
        du[2] = p[1] * u[1] * u[1];                                                                       \
    }                                                                                                     \
    static void robertson_jac_##S(T* J, const T* u, const T* p, T t) {                                    \
        (void)t;                                                                                          \
        J[0] = -p[0]; J[1] = p[2] * u[2]; J[2] = p[2] * u[1];                                             \
        J[3] = p[0]; J[4] = -2 * p[1] * u[1] - p[2] * u[2]; J[5] = -p[2] * u[1];                          \
        J[6] = 0; J[7] = 2 * p[1] * u[1]; J[8] = 0;                                                       \
    }                                                                                                     \
    static void linear_rhs_##S(T* du, const T* u, const T* p, T t) { (void)t; du[0] = p[0] * u[0]; }       \
    static void linear_jac_##S(T* J, const T* u, const T* p, T t) { (void)t; (void)u; J[0] = p[0]; }       \
    static void gbm_rhs_##S(T* du, const T* u, const T* p, T t) { (void)t; du[0] = p[0] * u[0]; }          \
    static void gbm_noise_##S(T* g, const T* u, const T* p, T t) { (void)t; g[0] = p[1] * u[0]; }          \
    /* net16: reversible chain X_i <-> X_{i+1} (kf = p[0]*w_i, kr = p[1]*v_i) plus couplings          \
     * X_i + X_{i+1} -> X_{i+2} (rate p[2]*z_i), first-order decay of X_0 at p[3], i = 0..; the        \
     * fixed weights w,v,z are the constants below (seed-2 logU(0.1,10), frozen). */                    \
    static void net16_rhs_##S(T* du, const T* u, const T* p, T t) {                                       \
        (void)t;                                                                                          \
        for (int i = 0; i < 16; i++) du[i] = 0;                                                           \
        for (int i = 0; i < 15; i++) {                                                                    \
            const T fl = p[0] * (T)net16_w[i] * u[i] - p[1] * (T)net16_v[i] * u[i + 1];                   \
            du[i] -= fl;                                                                                  \
            du[i + 1] += fl;                                                                              \
        }                                                                                                 \
        for (int i = 0; i < 14; i++) {                                                                    \
            const T r = p[2] * (T)net16_z[i] * u[i] * u[i + 1];                                           \
            du[i] -= r;                                                                                   \
            du[i + 1] -= r;                                                                               \
            du[i + 2] += r;                                                                               \
        }                                                                                                 \
        du[0] -= p[3] * u[0];                                                                             \
    }                                                                                                     \
    static T net16_cond_##S(const T* u, const T* p, T t) { (void)t; return u[0] - p[4]; }                 \
    static void net16_affect_##S(T* u, const T* p, T t) { (void)t; u[0] += p[5]; }                        \
    static T tcross_cond_##S(const T* u, const T* p, T t) { (void)u; return t - p[1]; }                   \
    static void noop_affect_##S(T* u, const T* p, T t) { (void)u; (void)p; (void)t; }                      \
    /* test/core.jl:76: DiscreteCallback((u,t,integrator) -> t >= 0.5, affect!) ; here affect! halves u */ \
    static int tge_dcond_##S(const T* u, const T* p, T t) { (void)u; return t >= p[1]; }                  \
    static void halve_affect_##S(T* u, const T* p, T t) { (void)p; (void)t; u[0] = u[0] * (T)0.5; }

const double net16_w[15] = {1.3, 0.42, 6.1, 0.17, 2.9, 0.88, 4.4, 0.23, 7.7, 1.9, 0.35, 3.3, 0.61, 5.2, 1.1};
const double net16_v[15] = {0.7, 2.4, 0.19, 3.8, 0.52, 1.6, 0.11, 8.3, 0.93, 0.27, 4.9, 0.44, 2.2, 0.15, 6.6};
const double net16_z[14] = {0.9, 0.31, 2.7, 0.14, 1.8, 0.66, 3.9, 0.21, 5.5, 0.48, 1.2, 0.12, 2.1, 0.77};

DEFMODELS(double, f64)
DEFMODELS(float, f32)

#define PICK(name) (is_f64 ? (void*)name##_f64 : (void*)name##_f32)
void* orc_model_fn(const char* model, const char* which, int is_f64) {
    if (!strcmp(model, "lorenz") || !strcmp(model, "lorenz_additive")) {
        if (!strcmp(which, "rhs")) return PICK(lorenz_rhs);
        if (!strcmp(which, "jac")) return PICK(lorenz_jac);
        if (!strcmp(which, "noise")) return PICK(lorenz_noise);
    } else if (!strcmp(model, "robertson")) {
        if (!strcmp(which, "rhs")) return PICK(robertson_rhs);
        if (!strcmp(which, "jac")) return PICK(robertson_jac);
    } else if (!strcmp(model, "linear")) {
        if (!strcmp(which, "rhs")) return PICK(linear_rhs);
        if (!strcmp(which, "jac")) return PICK(linear_jac);
        if (!strcmp(which, "cond")) return PICK(tcross_cond);
        if (!strcmp(which, "affect")) return PICK(noop_affect);
        if (!strcmp(which, "dcond")) return PICK(tge_dcond);
        if (!strcmp(which, "daffect")) return PICK(halve_affect);
    } else if (!strcmp(model, "gbm")) {
        if (!strcmp(which, "rhs")) return PICK(gbm_rhs);
        if (!strcmp(which, "noise")) return PICK(gbm_noise);
    } else if (!strcmp(model, "net16")) {
        if (!strcmp(which, "rhs")) return PICK(net16_rhs);
        if (!strcmp(which, "cond")) return PICK(net16_cond);
        if (!strcmp(which, "affect")) return PICK(net16_affect);
    }
    return 0;
}
