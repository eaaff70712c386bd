/* b200ens.h -- C ABI of libb200ens.so, the B200-native ensemble ODE/SDE back-end.
 *
 * This is the drop-in boundary for the hot path
 *     solve(EnsembleProblem(prob; prob_func), alg, EnsembleB200();
 *           trajectories=N, saveat, dt, abstol, reltol)
 * of DifferentialEquations.jl.  The reference defines no FFI for this path (it is a
 * 32-line re-export metapackage, /root/reference/src/DifferentialEquations.jl:8-9); what
 * each entry point replaces is therefore cited by the reference-visible NAME it stands
 * behind (the exported-name contract /root/reference/test/qa/qa.jl:3-217) and by the call
 * sites in /root/reference/test/core.jl.  The Julia binding a maintainer would add is
 * shown in INTEGRATION.md (julia/EnsembleB200.jl).
 *
 * Rules of the ABI: plain pointers and sizes, caller owns every buffer, no exceptions or
 * longjmp cross it, no callbacks into the host language (model functions arrive as CUDA-C
 * SOURCE, JIT-compiled with NVRTC for sm_100a).  There is NO CPU fallback: every solve
 * entry fails with B200ENS_E_NODEVICE when no CUDA device is usable.
 */
#ifndef B200ENS_H
#define B200ENS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 7: alg B200ENS_FBDF; ContinuousCallback directions through B2_EVENT_DIR / b2_affect_neg and event_terminate bit 2;
 *    save_tstops = -1 (auto) now interpolates for the Rodas family too (derived dense output, DAEs included) */
#define B200ENS_ABI_VERSION 7

/* scalar type of u, p, t (Julia eltype(u0)) */
enum b200ens_dtype { B200ENS_F32 = 0, B200ENS_F64 = 1 };

/* algorithm ids <- OrdinaryDiffEq / StochasticDiffEq algorithm types
 * (Tsit5 qa.jl:119, Vern7 qa.jl:126, Rosenbrock23 qa.jl:98, Rodas5P qa.jl:97;
 *  Rodas5/Rodas4 live in OrdinaryDiffEqRosenbrock, EM/SOSRA in StochasticDiffEq) */
enum b200ens_alg {
    B200ENS_TSIT5 = 1, B200ENS_VERN7 = 2, B200ENS_ROSENBROCK23 = 3, B200ENS_RODAS5 = 4,
    B200ENS_RODAS5P = 5, B200ENS_EM = 6, B200ENS_SOSRA = 7, B200ENS_RODAS4 = 8,
    B200ENS_SRIW1 = 9,  /* Roessler SRI W1: strong order 1.5 for diagonal noise (fixed dt; dW and dZ per step) */
    B200ENS_FBDF = 10   /* FBDF (qa.jl:57): variable-order (1..5) fixed-leading-coefficient BDF, Newton corrector with the analytic
                           Jacobian (jac_src required), Hermite dense output; a callback that fires restarts the history at order 1; constant mass matrix as for the Rodas family */
};

/* per-trajectory return codes <- SciMLBase.ReturnCode (qa.jl:213); the Julia glue maps by name */
enum b200ens_retcode {
    B200ENS_RC_DEFAULT = 0, B200ENS_RC_SUCCESS = 1, B200ENS_RC_TERMINATED = 2, B200ENS_RC_MAXITERS = 3,
    B200ENS_RC_DTLESSTHANMIN = 4, B200ENS_RC_UNSTABLE = 5, B200ENS_RC_DTNAN = 6, B200ENS_RC_FAILURE = 7
};

/* library error codes (negative returns) */
enum b200ens_error {
    B200ENS_OK = 0, B200ENS_E_INVALID = -1, B200ENS_E_COMPILE = -2, B200ENS_E_NODEVICE = -3,
    B200ENS_E_CUDA = -4, B200ENS_E_NOMEM = -5, B200ENS_E_UNSUPPORTED = -6
};

/* model flags */
#define B200ENS_MODEL_FAST_MATH 1u /* let NVRTC contract a*b+c in MODEL code (breaks bitwise oracle parity) */
/* bit 2u: reserved (ABI <= 4: B200ENS_MODEL_PACKED_X2, the two-trajectories-per-thread kernel; removed -- the Float32
   steppers now pack component PAIRS of one trajectory into FFMA2/FMUL2, which is faster and bit-identical) */
#define B200ENS_MODEL_KSMEM 4u     /* force ERK stage vectors into shared memory (default: automatic when the register variant spills > 4 KB) */
#define B200ENS_MODEL_SPLIT 8u     /* force the split kernel (one trajectory per lane of a 4-warp CTA, components split over the
                                      warps; Tsit5 / Vern7, ContinuousCallback or VectorContinuousCallback but no DiscreteCallback; default: automatic when
                                      the one-thread variant spills > 1 KB (Vern7) / 4 KB (Tsit5)) */
#define B200ENS_MODEL_NOSPLIT 16u  /* never use the split kernel */
#define B200ENS_MODEL_SDE_ADAPTIVE 32u /* SRIW1 / SOSRA only: compile the ADAPTIVE stepper (embedded error estimate, PI controller
                                      with the strong order 3/2, qmax default 1.125, rejection sampling with memory RSwM1).
                                      opts.dt is the initial step, abstol / reltol apply; noise_injected = 1 injects the
                                      stream of standard normals (opts.noise_stream_len) instead of Brownian increments */

/* What a problem looks like to the library: ODEProblem / SDEProblem (qa.jl:86,103) with f,
 * jac, tgrad, g and one ContinuousCallback (qa.jl:26; test/core.jl:69-72) given as CUDA-C
 * source defining these device functions (`real` is float or double per dtype):
 *   __device__ void b2_rhs      (real* du, const real* u, const real* p, real t);
 *   __device__ void b2_jac      (real* J,  const real* u, const real* p, real t);   row-major n x n
 *   __device__ void b2_tgrad    (real* dT, const real* u, const real* p, real t);   optional
 *   __device__ void b2_noise    (real* g,  const real* u, const real* p, real t);   diagonal noise
 *   __device__ real b2_condition(const real* u, const real* p, real t);             ContinuousCallback
 *   __device__ void b2_affect   (real* u,  const real* p, real t);
 *   direction (upstream: an upcrossing runs affect!, a downcrossing affect_neg!; `nothing` ignores that direction):
 *   condition_src may carry `#undef B2_EVENT_DIR` + `#define B2_EVENT_DIR +1` (upcrossings only) or `-1`
 *   (downcrossings only; b2_affect is then the downcrossing affect); a two-sided callback whose affect_neg! differs
 *   from affect! adds `#undef B2_HAS_AFFECT_NEG` + `#define B2_HAS_AFFECT_NEG 1` and
 *   __device__ void b2_affect_neg(real* u, const real* p, real t);                to affect_src
 *   VectorContinuousCallback (qa.jl:124): condition_src carries `#define B2_NCOND <len>` and defines
 *   __device__ void b2_vcondition(real* g, const real* u, const real* p, real t);   affect_src defines
 *   __device__ void b2_vaffect  (real* u,  const real* p, real t, int idx);         idx = 0-based event index
 *   constant mass matrix (M u' = f, Rodas4/5/5P and FBDF): rhs_src carries `#define B2_HAS_MASS 1` and
 *   `static constexpr double B2_MASS_[n*n]` (row-major)
 *   __device__ bool b2_dcondition(const real* u, const real* p, real t);            DiscreteCallback (qa.jl:36,
 *   __device__ void b2_daffect  (real* u,  const real* p, real t);                   test/core.jl:76-77)
 * NULL = absent.  One ContinuousCallback and one DiscreteCallback may be combined (CallbackSet, qa.jl:24):
 * the continuous event is handled first, the discrete condition is then tested on the accepted state. */
typedef struct b200ens_model_desc {
    uint32_t struct_size;   /* sizeof(b200ens_model_desc) */
    int32_t n_state;        /* length(u0), 1..32 */
    int32_t n_param;        /* length(p), 0..64 */
    int32_t dtype;          /* enum b200ens_dtype */
    int32_t alg;            /* enum b200ens_alg */
    uint32_t flags;
    const char* rhs_src;
    const char* jac_src;
    const char* tgrad_src;
    const char* noise_src;
    const char* condition_src;
    const char* affect_src;
    const char* name;       /* label for logs / cache, may be NULL */
    const char* dcondition_src;
    const char* daffect_src;
    const int32_t* save_idxs; /* solve(...; save_idxs = [...]): NULL, or the 0-based state components that are saved, in this
                                 order (distinct, each < n_state).  out_u rows, the moments and save_everystep rows then
                                 have n_save_idxs entries instead of n_state: the kernels store (and the host link carries)
                                 only what the caller asked for */
    int32_t n_save_idxs;      /* number of entries of save_idxs; 0 = every component */
    int32_t reserved0;        /* must be 0 */
} b200ens_model_desc;

/* solve keyword arguments (test/core.jl:14,54,72,93: reltol, abstol, dense/saveat, callback).
 * Doubles set to NaN or <0 take the upstream default. */
typedef struct b200ens_opts {
    uint32_t struct_size;   /* sizeof(b200ens_opts) */
    int32_t adaptive;       /* 1 adaptive (default for ODE algs), 0 fixed dt */
    double t0, t1;          /* tspan */
    double dt;              /* initial dt (adaptive; 0 = automatic per-trajectory initial step) or the fixed dt */
    double abstol, reltol;  /* defaults 1e-6 / 1e-3 */
    double dtmin, dtmax;    /* defaults 0 (+ eps(t) floor) / t1-t0 */
    double qmin, qmax, gamma, beta1, beta2, qoldinit; /* PI controller; defaults 1/5, 10, 9/10, 7/(10k), 2/(5k), 1e-4 */
    int64_t maxiters;       /* <=0: 100000 */
    uint64_t seed;          /* Philox key */
    uint64_t traj_offset;   /* global index of trajectory 0 of this call (Philox counter base) */
    int32_t noise_injected; /* 1: dW holds the Brownian increments; 0: Philox4x32-10 on device */
    int32_t event_terminate;/* bit 0: the ContinuousCallback's affect! terminates the trajectory (terminate!); bit 1: the DiscreteCallback
                               does; bit 2: the ContinuousCallback's affect_neg! does (read only when affect_src defines b2_affect_neg) */
    int32_t interp_points;  /* ContinuousCallback interp_points, <=0: 10 */
    int32_t save_tstops;    /* -1 auto (interpolate; tstops for FBDF on a mass-matrix problem), 0 interpolate through the stepper's dense output,
                               1 saveat points are tstops (steps are clipped to them) */
    uint32_t device_mask;   /* bit g set: use CUDA device g; 0: all visible devices */
    int32_t refill_threshold; /* lanes of a warp that must be idle before it fetches new trajectories; <=0 auto */
    int32_t block_threads;  /* <=0 auto */
    int32_t stage_outputs;  /* -1 auto, 0 direct global stores, 1 stage saveat outputs in shared memory */
    int32_t work_order;     /* -1 auto (on for adaptive ODE solves of >= 32768 trajectories per device chunk), 0 caller's
                               order, 1 integrate trajectories in descending expected-work order (device-side counting
                               sort on the initial-step proxy; scheduling only, results are bit-identical).  The order is
                               established inside windows of consecutive trajectories (sized so that the output rows in
                               flight stay within L2's reach; one window when the rows are whole 32-byte sectors);
                               a value > 1 is an explicit window size in trajectories (rounded up to 4096) */
    int32_t save_everystep; /* 1: save every accepted step instead of the saveat grid (upstream default without saveat):
                               n_save is then the CAPACITY per trajectory, saveat is ignored (may be NULL), out_u is
                               [N][n_save][n_state] and out_t (required) is [N][n_save] of the state type: slot 0 = (t0, u0),
                               slot k = after the k-th accepted step, unused slots NaN; a trajectory with more than
                               n_save - 1 accepted steps keeps integrating, the surplus is dropped (stats.naccept tells).
                               ODE steppers, one-thread kernels; b200ens_solve only */
    const double* abstol_vec; /* NULL, or n_state per-component absolute tolerances (solve(...; abstol = [...])); overrides abstol */
    const double* reltol_vec; /* NULL, or n_state per-component relative tolerances; overrides reltol */
    int64_t noise_stream_len; /* B200ENS_MODEL_SDE_ADAPTIVE with noise_injected = 1: dW holds STANDARD NORMALS,
                                 [N][noise_stream_len] of the state type, which every trajectory consumes in order in place
                                 of its Philox stream (2 n_state per fresh step, bridge draw or rejection); a trajectory that
                                 runs out of them ends with B200ENS_RC_FAILURE.  Ignored otherwise */
    int32_t shard_blocks;     /* b200ens_solve on G > 1 devices: how the trajectories are dealt.  0 auto (= 8), k >= 1: the
                                 ensemble is cut into G*k contiguous blocks which go to the devices in boustrophedon order
                                 0,1,..,G-1,G-1,..,1,0,0,1,.. -- every device gets the same mix of an ORDERED parameter sweep
                                 (work varies ~10x along the Lorenz rho-sweep; a linear trend cancels exactly), results are
                                 bit-identical to k = 1 (contiguous ranges [g N/G, (g+1) N/G)) */
    int32_t n_tstops;         /* number of entries of tstops (0: none) */
    const double* tstops;     /* solve(...; tstops = [...]) (SURVEY A.1 handle_tstop!): ascending times the integrator must hit
                                 exactly -- a step that would pass the next one is clipped to end on it, e.g. so that a
                                 DiscreteCallback with condition t == 4.0 sees that time (the classic dosing example).  Entries
                                 outside (t0, t1) are ignored.  ODE steppers only (a fixed-step SDE solve has its dt grid) */
} b200ens_opts;

typedef struct b200ens_stats {
    int32_t naccept, nreject, nf, nevents;
} b200ens_stats;

typedef struct b200ens_timing {
    double h2d_ms, kernel_ms, d2h_ms, total_ms; /* device-event times; max over devices */
    int32_t n_devices, launches;
    int32_t grid, block, smem_bytes, regs;      /* of the last launch on the first device */
    double kernel_ms_min;                       /* min over devices of the summed kernel time (kernel_ms is the max): their
                                                   ratio is the work balance of a multi-device solve */
} b200ens_timing;

typedef struct b200ens_model b200ens_model;

int b200ens_abi_version(void);
/* number of usable CUDA devices; <0 on CUDA initialisation failure */
int b200ens_device_count(void);
/* thread-local description of the last error */
const char* b200ens_last_error(void);

/* "<major>.<minor> <path>" of the NVRTC the library JIT-compiles with (dlopen'ed explicitly: $B200ENS_NVRTC, then the
 * CUDA toolkit's libnvrtc.so.12, then the default loader path); "" when none could be loaded. */
const char* b200ens_nvrtc_info(void);

/* Fill o with defaults (everything "auto"). */
void b200ens_opts_init(b200ens_opts* o);

/* JIT-compile (model source x algorithm kernel x dtype) for sm_100a.  Needs no GPU.  The
 * NVRTC/ptxas log (register count, spills) is copied into log (may be NULL). */
int b200ens_compile(const b200ens_model_desc* d, b200ens_model** out, char* log, size_t log_len);
void b200ens_free(b200ens_model* m);
/* cubin size, registers/thread, static smem, local (spill) bytes as parsed from the ptxas log; any may be NULL */
int b200ens_model_info(const b200ens_model* m, int64_t* cubin_bytes, int32_t* regs, int32_t* smem, int32_t* lmem);

/* The ensemble solve behind  solve(::EnsembleProblem, alg, ::EnsembleB200; trajectories=N, ...)
 * (replaces SciMLBase.__solve / solve_batch / batch_func of EnsembleThreads, qa.jl:56,192).
 * HOST buffers, trajectory-major:
 *   u0 [N][n_state], p [N][n_param], saveat [n_save] (ascending, within [t0,t1]; checked on the host,
 *   B200ENS_E_INVALID otherwise -- b200ens_solve_device trusts its device-resident grid),
 *   dW  NULL or [N][nsteps][nvec][n_state]  (nvec = 1 EM, 2 SOSRA / SRIW1: dW then dZ; nsteps = ceil((t1-t0)/dt));
 *       adaptive SDE models: [N][noise_stream_len] standard normals,
 *   out_u [N][n_save][n_state], out_t [n_save] or NULL, retcode [N], stats [N] or NULL.
 * Trajectory ranges are sharded over the devices in device_mask; blocking. */
int b200ens_solve(b200ens_model* m, const b200ens_opts* o, int64_t N, const void* u0, const void* p,
                  const void* saveat, int32_t n_save, const void* dW, void* out_u, void* out_t,
                  int32_t* retcode, b200ens_stats* stats, b200ens_timing* timing);

/* Same solve with every buffer already resident in the HBM of `device`, launched on
 * `stream` (a cudaStream_t, NULL = default stream).  Blocks until the kernel has finished
 * unless timing == NULL, in which case it only enqueues. */
int b200ens_solve_device(b200ens_model* m, const b200ens_opts* o, int32_t device, void* stream, int64_t N,
                         const void* d_u0, const void* d_p, const void* d_saveat, int32_t n_save,
                         const void* d_dW, void* d_out_u, int32_t* d_retcode, b200ens_stats* d_stats,
                         b200ens_timing* timing);

/* Ensemble summary statistics without shipping the trajectories to the host (SciMLBase.EnsembleAnalysis
 * timestep_mean / timestep_meanvar, qa.jl:211).  Same inputs as b200ens_solve; instead of
 * out_u the call returns, for every (save point, component), the sum and the sum of squares over the trajectories
 * that finished with B200ENS_RC_SUCCESS or B200ENS_RC_TERMINATED (SciMLBase.successful_retcode), and their number:
 *   sum [n_save][n_state], sumsq [n_save][n_state] (double), count (int64).  mean = sum/count,
 *   var = (sumsq - sum^2/count)/(count-1).  Partial results of several GPUs / ranks add up.
 * Two device paths, same results: the solve writes out_u into HBM chunk buffers and a streaming second pass reduces
 * them (default: HBM absorbs the rows faster than L2 adds them), or -- FUSED, for rows of >= 4 MB per trajectory or with
 * B200ENS_FUSE_MOMENTS=1 in the environment -- the ODE kernels add every saved value to the accumulators themselves and
 * out_u is never allocated (a chunk in which a trajectory fails is recomputed through out_u, so failures never count). */
int b200ens_solve_moments(b200ens_model* m, const b200ens_opts* o, int64_t N, const void* u0, const void* p,
                          const void* saveat, int32_t n_save, const void* dW, double* sum, double* sumsq,
                          int64_t* count, int32_t* retcode, b200ens_timing* timing);

/* pinned host memory for callers that want zero-staging transfers */
void* b200ens_host_alloc(size_t bytes);
void b200ens_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif
