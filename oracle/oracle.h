/* oracle.h -- CPU ORACLE for the B200 ensemble ODE/SDE hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libb200ens.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (/root/reference, DifferentialEquations.jl v8.0.3) is a
 * 32-line re-export metapackage (src/DifferentialEquations.jl:8-9); the arithmetic lives
 * in un-vendored OrdinaryDiffEq 7 / SciMLBase 3 (Project.toml:7,10,13,18) and Julia is not
 * installed here.  This file restates the upstream algorithm as specified in SURVEY.md
 * Appendix A (integrator loop, PI controller, norm, saveat, callbacks, SDE steps) with the
 * coefficient tables of Appendix B.  It is anchored on what the reference's own tests pin
 * (test/core.jl:15,18,34,93-95: Success retcodes, first saved value == u0, 11 saveat
 * points) and on independent truths (closed forms, order conditions, scipy/mpmath
 * solutions under tests/golden/).
 */
#ifndef B2_ORACLE_H
#define B2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NMAX 32
#define ORC_PMAX 64

enum { ORC_TSIT5 = 1, ORC_VERN7 = 2, ORC_ROSENBROCK23 = 3, ORC_RODAS5 = 4, ORC_RODAS5P = 5,
       ORC_EM = 6, ORC_SOSRA = 7, ORC_RODAS4 = 8, ORC_SRIW1 = 9, ORC_FBDF = 10 };
enum { ORC_RC_DEFAULT = 0, ORC_RC_SUCCESS = 1, ORC_RC_TERMINATED = 2, ORC_RC_MAXITERS = 3,
       ORC_RC_DTLESSTHANMIN = 4, ORC_RC_UNSTABLE = 5, ORC_RC_DTNAN = 6, ORC_RC_FAILURE = 7 };

typedef struct {
    int32_t naccept, nreject, nf, nevents;
} orc_stats;

/* model functions; the void* members are cast to the f32 or f64 signature:
 *   rhs   (T* du, const T* u, const T* p, T t)
 *   jac   (T* J /+ row-major n x n +/, const T* u, const T* p, T t)
 *   tgrad (T* dT, const T* u, const T* p, T t)             optional (NULL = autonomous)
 *   noise (T* g, const T* u, const T* p, T t)              diagonal noise
 *   cond  T (const T* u, const T* p, T t)
 *   affect(T* u, const T* p, T t)
 */
typedef struct {
    int32_t alg, n_state, n_param, adaptive;
    double t0, t1, dt, abstol, reltol;
    double dtmin, dtmax, qmin, qmax, gamma, beta1, beta2, qoldinit; /* <0 or NaN => default */
    int64_t maxiters;
    int32_t n_save;
    int32_t noise_injected;   /* 1: dW given [N][nsteps][nvec][n]; 0: Philox4x32-10 */
    uint64_t seed;
    uint64_t traj_offset;     /* global index of trajectory 0 (Philox counter base) */
    int32_t has_event, event_terminate, interp_points;
    int32_t save_tstops;      /* 1: saveat points are tstops (steps clipped, no interpolation) */
    void *rhs, *jac, *tgrad, *noise, *cond, *affect;
    void *dcond, *daffect;    /* DiscreteCallback: int dcond(u,p,t), daffect(u,p,t); NULL = none */
    int32_t devent_terminate, pad_;
    const double *abstol_vec, *reltol_vec;   /* NULL, or n_state per-component tolerances (override abstol / reltol) */
    void *vcond, *vaffect;    /* VectorContinuousCallback: vcond(g,u,p,t) fills ncond values, vaffect(u,p,t,idx); has_event=1 */
    int32_t ncond;
    uint32_t vterm_mask;      /* bit k: event index k terminates the trajectory */
    const double* mass;       /* NULL, or constant mass matrix [n_state][n_state] (M u' = f; Rodas4/5/5P only) */
    void* every_t;            /* save_everystep: step times [N][n_save] of the state type; n_save = capacity, saveat ignored */
                              /* sde_adaptive: NULL, or receives W(t_end) of the accepted Brownian path, [N][n_state] */
    int32_t save_everystep;
    int32_t sde_adaptive;     /* 1: adaptive SRIW1 / SOSRA with rejection sampling with memory (RSwM1); `adaptive` stays unused for SDE algs */
    int64_t noise_stream_len; /* sde_adaptive && noise_injected: dW is [N][noise_stream_len] STANDARD NORMALS, consumed in order
                                 instead of the trajectory's Philox stream (running out of them -> ORC_RC_FAILURE) */
    int32_t spec_arith;       /* 1: step control to the letter of SURVEY.md A.4 / A.5 -- error norm in the working precision with
                                 IEEE division and sqrt, EEst^beta1 / qold^beta2 with libm pow, dt / q -- instead of the kernels'
                                 contract (Float32 norm with a Newton reciprocal, log-domain controller with polynomial
                                 log2 / exp2).  ODE steppers only.  Exists to MEASURE what the contract changes
                                 (tests/test_spec_arith.py); the kernels are compared bit for bit against spec_arith = 0 */
    int32_t n_tstops;         /* solve(...; tstops): number of entries of tstops */
    int32_t event_dir;        /* ContinuousCallback / VectorContinuousCallback direction (an upcrossing triggers affect!, a
                                 downcrossing affect_neg!): 0 both, +1 upcrossings only (affect_neg! = nothing), -1 downcrossings only
                                 (affect! = nothing).  The direction is the sign of the condition at the step start */
    int32_t pad2_;
    void* affect_neg;         /* NULL: downcrossings run `affect` (upstream's default affect_neg! = affect!) */
    const double* tstops;     /* ascending times the integrator must hit exactly (handle_tstop!); compared in the
                                 state type; entries outside (t0, t1) are ignored.  ODE steppers only */
} orc_opts;

/* u0 [N][n], p [N][m], saveat [n_save], out_u [N][n_save][n], retcode [N], stats [N] or NULL.
 * nthreads: OpenMP threads (<=0: all).  Returns 0 or a negative error. */
int orc_solve_f64(const orc_opts* o, int64_t N, const double* u0, const double* p, const double* saveat,
                  const double* dW, double* out_u, int32_t* retcode, orc_stats* stats, int nthreads);
int orc_solve_f32(const orc_opts* o, int64_t N, const float* u0, const float* p, const float* saveat,
                  const float* dW, float* out_u, int32_t* retcode, orc_stats* stats, int nthreads);

/* built-in hand-written models (oracle/models.c): "lorenz", "robertson", "linear",
 * "gbm", "lorenz_additive", "net16".  which: "rhs","jac","noise","cond","affect". */
void* orc_model_fn(const char* model, const char* which, int is_f64);

/* deterministic pow used by the PI controller (restates upstream's approximate FastPower) */
float orc_fastpow(float x, float y);
float orc_fastlog2(float x);
float orc_fastexp2(float y);
/* Float32 reciprocal of the error norm: exponent-flip guess + two Newton steps (relative error <= 6.6e-6) */
float orc_rcp_nr(float s);
/* Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* normals exactly as the kernels draw them: Philox block `block` of trajectory `traj`'s normal stream */
void orc_normals_f32(uint64_t seed, uint64_t traj, uint64_t block, float z[4]);
void orc_normals_f64(uint64_t seed, uint64_t traj, uint64_t block, double z[2]);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
