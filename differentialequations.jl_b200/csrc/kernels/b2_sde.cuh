// b2_sde.cuh -- fixed-step SDE ensemble kernel: Euler-Maruyama (diagonal noise), SRIW1 (diagonal noise, strong
// order 1.5) and SOSRA (additive noise, strong order 1.5) with counter-based Philox4x32-10 noise generated on the
// device, or Brownian increments injected by the caller for pathwise parity.
// Reference names: SDEProblem /root/reference/test/qa/qa.jl:103 (EM/SOSRA live in
// StochasticDiffEq, outside the dep closure; step forms SURVEY.md A.9, table B.8, RNG B.9).
#pragma once
#include "b2_common.cuh"
#include "tableaus_gen.cuh"

#define B2_PHILOX_M0 0xD2511F53u
#define B2_PHILOX_M1 0xCD9E8D57u
#define B2_PHILOX_W0 0x9E3779B9u
#define B2_PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ uint4 b2_philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(B2_PHILOX_M0, c.x), lo0 = B2_PHILOX_M0 * c.x;
        const unsigned hi1 = __umulhi(B2_PHILOX_M1, c.z), lo1 = B2_PHILOX_M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += B2_PHILOX_W0;
        k.y += B2_PHILOX_W1;
    }
    return c;
}

#define B2_TWO_PI 6.283185307179586476925
#if B2_F64
#define B2_NORMALS_PER_CALL 2
// Each trajectory consumes one continuous stream of normals: normal q = element q % K of the Philox block with
// counter = (traj_lo, traj_hi, b_lo, b_hi), b = q / K, key = (seed_lo, seed_hi); no generated normal is discarded.
// f64: two 53-bit uniforms -> one Box-Muller pair (K = 2)
__device__ __forceinline__ void b2_normals(unsigned long long seed, unsigned long long traj, unsigned long long block,
                                           real* z) {
    const uint4 r = b2_philox4x32_10(make_uint4((unsigned)traj, (unsigned)(traj >> 32), (unsigned)block, (unsigned)(block >> 32)),
                                     make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned long long a = ((unsigned long long)r.x << 32) | r.y, b = ((unsigned long long)r.z << 32) | r.w;
    const double u1 = ((double)(a >> 11) + 0.5) * 1.1102230246251565404e-16;
    const double u2 = ((double)(b >> 11) + 0.5) * 1.1102230246251565404e-16;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);   // angle 2 pi u2: exact argument (2 u2), one shared range reduction
    z[0] = rad * cs;
    z[1] = rad * sn;
}
#else
#define B2_NORMALS_PER_CALL 4
// f32: four u32 -> two Box-Muller pairs (K = 4); uniforms ((x>>8)+0.5)*2^-24 in (0,1)
__device__ __forceinline__ void b2_normals(unsigned long long seed, unsigned long long traj, unsigned long long block,
                                           real* z) {
    const uint4 r = b2_philox4x32_10(make_uint4((unsigned)traj, (unsigned)(traj >> 32), (unsigned)block, (unsigned)(block >> 32)),
                                     make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float u1 = ((float)(w[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-8f;
        const float u2 = ((float)(w[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);   // angle 2 pi u2 with an exact argument and ONE range reduction: ~30 issue slots
                                          // fewer per pair than sinf + cosf of a rounded 2 pi u2, and closer to the true value
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
    }
}
#endif

#define SO(x) ((real)(B2T_SOSRA_##x))
#define B2_INV_SQRT3 0.57735026918962576451

// one step of stepper ALG from (t, up) with the increments dW, dZ over dt -> u.  ERR: also form the two parts of the
// embedded error estimate (adaptive driver, b2_sde_adaptive.cuh); the fixed-step driver compiles them out.
template <int ALG, bool ERR>
__device__ __forceinline__ void b2_sde_step(real* __restrict__ u, const real* __restrict__ up, const real* __restrict__ dW,
                                            const real* __restrict__ dZ, const real* __restrict__ p, const real t, const real dt,
                                            real* __restrict__ E1, real* __restrict__ E2) {
    real k1[B2_N], g1[B2_N];
    if (ALG == 6) {  // Euler-Maruyama: u += f dt + g dW
        b2_rhs(k1, up, p, t);
        b2_noise(g1, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++) u[i] = b2_fma(g1[i], dW[i], b2_fma(dt, k1[i], up[i]));
    } else if (ALG == 9) {
        // SRIW1 (Roessler SRI W1): strong order 1.5 for diagonal noise; same expression tree as the oracle
        real chi2[B2_N], i11[B2_N], i111[B2_N], k2[B2_N], g2[B2_N], g3[B2_N], g4[B2_N], H[B2_N];
        const real sq = b2_sqrt(dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            chi2[i] = (real)0.5 * b2_fma(dZ[i], (real)B2_INV_SQRT3, dW[i]);
            i11[i] = (real)0.5 * b2_fma(dW[i], dW[i], -dt) / sq;
            i111[i] = dW[i] * b2_fma(dW[i], dW[i], (real)-3 * dt) / ((real)6 * dt);
        }
        b2_rhs(k1, up, p, t);
        b2_noise(g1, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++) H[i] = b2_fma(chi2[i], (real)1.5 * g1[i], b2_fma(dt, (real)0.75 * k1[i], up[i]));
        b2_rhs(k2, H, p, t + (real)0.75 * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) H[i] = b2_fma(sq, (real)0.5 * g1[i], b2_fma(dt, (real)0.25 * k1[i], up[i]));
        b2_noise(g2, H, p, t + (real)0.25 * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) H[i] = b2_fma(sq, -g1[i], b2_fma(dt, k1[i], up[i]));
        b2_noise(g3, H, p, t + dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real bb = b2_fma((real)0.5, g3[i], b2_fma((real)3, g2[i], (real)-5 * g1[i]));
            H[i] = b2_fma(sq, bb, b2_fma(dt, (real)0.25 * k1[i], up[i]));
        }
        b2_noise(g4, H, p, t + (real)0.25 * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real d = b2_fma((real)(2.0 / 3.0), k2[i], (real)(1.0 / 3.0) * k1[i]);
            real c1 = -dW[i];
            c1 = b2_fma((real)-1, i11[i], c1);
            c1 = b2_fma((real)2, chi2[i], c1);
            c1 = b2_fma((real)-2, i111[i], c1);
            real c2 = (real)(4.0 / 3.0) * dW[i];
            c2 = b2_fma((real)(4.0 / 3.0), i11[i], c2);
            c2 = b2_fma((real)(-4.0 / 3.0), chi2[i], c2);
            c2 = b2_fma((real)(5.0 / 3.0), i111[i], c2);
            real c3 = (real)(2.0 / 3.0) * dW[i];
            c3 = b2_fma((real)(-1.0 / 3.0), i11[i], c3);
            c3 = b2_fma((real)(-2.0 / 3.0), chi2[i], c3);
            c3 = b2_fma((real)(-2.0 / 3.0), i111[i], c3);
            u[i] = b2_fma(i111[i], g4[i], b2_fma(c3, g3[i], b2_fma(c2, g2[i], b2_fma(c1, g1[i], b2_fma(dt, d, up[i])))));
            if (ERR) {  // embedded estimate of the adaptive driver: E1 = h (f1 + f2), E2 = the chi2 and I111 terms of the update
                real e2a = (real)2 * g1[i];
                e2a = b2_fma((real)(-4.0 / 3.0), g2[i], e2a);
                e2a = b2_fma((real)(-2.0 / 3.0), g3[i], e2a);
                real e2b = (real)-2 * g1[i];
                e2b = b2_fma((real)(5.0 / 3.0), g2[i], e2b);
                e2b = b2_fma((real)(-2.0 / 3.0), g3[i], e2b);
                e2b = e2b + g4[i];
                E1[i] = dt * (k1[i] + k2[i]);
                E2[i] = b2_fma(i111[i], e2b, chi2[i] * e2a);
            }
        }
    } else {  // SOSRA
        real chi2[B2_N], g2[B2_N], g3[B2_N], k2[B2_N], k3[B2_N], H[B2_N];
#pragma unroll
        for (int i = 0; i < B2_N; i++) chi2[i] = (real)0.5 * b2_fma(dZ[i], (real)B2_INV_SQRT3, dW[i]);
        b2_noise(g1, up, p, t + SO(c11) * dt);
        b2_noise(g2, up, p, t + SO(c12) * dt);
        b2_noise(g3, up, p, t + SO(c13) * dt);
        b2_rhs(k1, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++)
            H[i] = b2_fma(chi2[i], SO(B021) * g1[i], b2_fma(dt, SO(A021) * k1[i], up[i]));
        b2_rhs(k2, H, p, t + SO(c02) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real aa = b2_fma(SO(A032), k2[i], SO(A031) * k1[i]);
            const real bb = b2_fma(SO(B032), g2[i], SO(B031) * g1[i]);
            H[i] = b2_fma(chi2[i], bb, b2_fma(dt, aa, up[i]));
        }
        b2_rhs(k3, H, p, t + SO(c03) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real d = SO(alpha1) * k1[i];
            d = b2_fma(SO(alpha2), k2[i], d);
            d = b2_fma(SO(alpha3), k3[i], d);
            real e1 = SO(beta11) * g1[i];
            e1 = b2_fma(SO(beta12), g2[i], e1);
            e1 = b2_fma(SO(beta13), g3[i], e1);
            real e2 = SO(beta21) * g1[i];
            e2 = b2_fma(SO(beta22), g2[i], e2);
            e2 = b2_fma(SO(beta23), g3[i], e2);
            u[i] = b2_fma(chi2[i], e2, b2_fma(dW[i], e1, b2_fma(dt, d, up[i])));
            if (ERR) {  // SRA estimate: E1 = h (f1 + f2 + f3), E2 = chi2 * sum beta2_i g_i
                E1[i] = dt * ((k1[i] + k2[i]) + k3[i]);
                E2[i] = chi2[i] * e2;
            }
        }
    }
}

template <bool B>
struct B2Bool {
    static constexpr bool v = B;
};
__host__ __device__ constexpr int b2_gcd_c(int a, int b) { return b == 0 ? a : b2_gcd_c(b, a % b); }

template <int ALG>
__device__ __forceinline__ void b2_sde_driver(const B2Args& a) {
    extern __shared__ __align__(16) unsigned char b2_smem[];
    const unsigned lane = threadIdx.x & 31u;
    const int warp_in_block = threadIdx.x >> 5;
    const int stride = a.stage_stride;
    real* const warp_stage = reinterpret_cast<real*>(b2_smem) + (size_t)warp_in_block * 32 * stride;
    real* const gout = reinterpret_cast<real*>(a.out_u);
    const real* const gu0 = reinterpret_cast<const real*>(a.u0);
    const real* const gp = reinterpret_cast<const real*>(a.p);
    const real* const gsave = reinterpret_cast<const real*>(a.saveat);
    const real* const gdW = reinterpret_cast<const real*>(a.dW);
    const int n_save = a.n_save;
    const int out_per_traj = n_save * B2_NOUT;
    constexpr int NVEC = (ALG == 7 || ALG == 9) ? 2 : 1;
    const real t0 = B2_ARG(a, t0), t1 = B2_ARG(a, t1), dt_user = B2_ARG(a, dt);

    B2Sink sink;
    sink.stage = stride ? warp_stage + (size_t)lane * stride : nullptr;
    sink.gout = gout;
    sink.margs = nullptr;   // fused moments: ODE kernels only
    bool exhausted = false;
    while (!exhausted) {
        // uniform work per path: hand out whole warps of consecutive paths
        long long idx = b2_fetch(B2_FULL, a.work_counter, a.N, lane, exhausted);
        const bool active = idx >= 0;
        if (__ballot_sync(B2_FULL, active) == 0u) break;
        if (active) {
            real u[B2_N], p[B2_NPA];
            sink.base = idx * (long long)out_per_traj;
#pragma unroll
            for (int i = 0; i < B2_N; i++) u[i] = gu0[idx * B2_N + i];
#pragma unroll
            for (int i = 0; i < B2_NPARAM; i++) p[i] = gp[idx * B2_NPARAM + i];
            int si = 0, rc = 0;
            int step = 0;
            while (si < n_save && __ldg(gsave + si) <= t0) {
                sink.put(si, u);
                si++;
            }
            const unsigned long long traj = a.traj_offset + (unsigned long long)idx;
            real zbuf[B2_NORMALS_PER_CALL];
            unsigned long long zblock = 0;
            // The time grid is driven by the INTEGER step index: t_k = fma(k, dt, t0), k < nsteps = ceil((t1 - t0)/dt)
            // (computed once in double by the host), the last step ends exactly at t1.  Accumulating t += dt in the
            // state type drifts (Float32, dt = 0.0057, t1 = 100: 17546 steps against an injected-increment buffer of
            // 17544), which read past the trajectory's block of dW.
            const int nsteps = (int)a.nsteps_noise;
            // Per-step control kept out of the loop body: the 64-bit maxiters bound as an int, sqrt(dt) of the regular
            // step hoisted (the IEEE square root is ~12 issue slots; only the clipped last step has another dt), the next
            // save time cached in a register (was a global load + compare per step).  The steps run in GROUPS of
            // U = K / gcd(K, NEED): a group consumes a whole number of K-normal Philox blocks, so inside the unrolled
            // group every normal's position in its block is a compile-time constant (no availability counter, no
            // select chain / indexed branch per normal).  Same stream layout -- normal q of a path is element q % K of
            // its block q / K -- and the same bits as the step-at-a-time loop.
            constexpr int NEED = NVEC * B2_N, KN = B2_NORMALS_PER_CALL;
            constexpr int GC = b2_gcd_c(KN, NEED), U = KN / GC;
            const int maxit = a.maxiters > 0x7fffffffLL ? 0x7fffffff : (int)a.maxiters;
            const real sq_user = b2_sqrt(dt_user);
            real tau_next = si < n_save ? __ldg(gsave + si) : (real)__int_as_float(0x7f800000);
            bool stop = false;
            // one group of U steps starting at step0.  TAIL = false: the caller guarantees that the group contains
            // neither the last step nor a step past maxiters, so those tests (and the clipped step's own dt and
            // square root) are compiled out of the main loop
            auto run_group = [&](const int step0, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::v;
#pragma unroll
                for (int s = 0; s < U; s++) {
                    step = step0 + s;
                    if (TAIL) {
                        if (step >= nsteps) break;
                        if (step >= maxit) {
                            rc = B2_RC_MAXITERS;
                            stop = true;
                            break;
                        }
                    }
                    const bool last = TAIL && step == nsteps - 1;
                    const real t = b2_fma((real)step, dt_user, t0);
                    const real dt = last ? t1 - t : dt_user;
                    real up[B2_N], dW[B2_N], dZ[B2_N];
#pragma unroll
                    for (int i = 0; i < B2_N; i++) up[i] = u[i];
                    if (a.noise_injected) {
                        const real* src = gdW + ((size_t)idx * a.nsteps_noise + (size_t)step) * (NVEC * B2_N);
#pragma unroll
                        for (int i = 0; i < B2_N; i++) {
                            dW[i] = src[i];
                            dZ[i] = NVEC > 1 ? src[B2_N + i] : (real)0;
                        }
                    } else {
                        real z[NEED];
                        const real sq = last ? b2_sqrt(dt) : sq_user;
#pragma unroll
                        for (int j = 0; j < NEED; j++) {
                            const int q = s * NEED + j;   // position in the group's stream: a constant after unrolling
                            if (q % KN == 0) {
                                b2_normals(a.seed, traj, zblock, zbuf);
                                zblock++;
                            }
                            z[j] = zbuf[q % KN];
                        }
#pragma unroll
                        for (int i = 0; i < B2_N; i++) {
                            dW[i] = sq * z[i];
                            dZ[i] = NVEC > 1 ? sq * z[(NVEC > 1 ? B2_N : 0) + i] : (real)0;
                        }
                    }
                    b2_sde_step<ALG, false>(u, up, dW, dZ, p, t, dt, nullptr, nullptr);
                    bool bad = false;
#pragma unroll
                    for (int i = 0; i < B2_N; i++) bad |= b2_isnan(u[i]);
                    if (bad) {
                        rc = B2_RC_UNSTABLE;
                        stop = true;
                        break;
                    }
                    const real tprev = t;
                    const real tnew = last ? t1 : b2_fma((real)(step + 1), dt_user, t0);
                    while (tau_next <= tnew) {  // linear interpolation between grid points
                        const real tau = tau_next;
                        if (tau == tnew) {
                            sink.put(si, u);
                        } else {
                            const real th = (tau - tprev) / dt;
                            real w[B2_N];
#pragma unroll
                            for (int i = 0; i < B2_N; i++) w[i] = b2_fma(th, u[i] - up[i], up[i]);
                            sink.put(si, w);
                        }
                        si++;
                        tau_next = si < n_save ? __ldg(gsave + si) : (real)__int_as_float(0x7f800000);
                    }
                    step = step0 + s + 1;   // completed steps (the stats' naccept)
                }
            };
            const int nmain = (nsteps - 1 < maxit ? nsteps - 1 : maxit) / U * U;   // whole groups before the last step / maxiters
            int step0 = 0;
            for (; step0 < nmain && !stop; step0 += U) run_group(step0, B2Bool<false>());
            for (; step0 < nsteps && !stop; step0 += U) run_group(step0, B2Bool<true>());
            if (rc == 0) rc = B2_RC_SUCCESS;
            else sink.fill(si, n_save, (real)__int_as_float(0x7fc00000));
            a.retcode[idx] = rc;
            if (a.stats) {
                B2Stats s;
                s.naccept = step;
                s.nreject = 0;
                s.nf = 0;
                s.nevents = 0;
                a.stats[idx] = s;
            }
        }
        if (stride) b2_flush(__ballot_sync(B2_FULL, active), warp_stage, stride, gout, idx, out_per_traj, lane);
    }
}
#undef SO
