// b2_work.cuh -- expected-work ordering of the trajectory queue ("longest expected first").
//
// Why: adaptive step counts differ ~10x between the trajectories of a parameter sweep (SURVEY.md section 6:
// 15..117 accepted steps for the random Lorenz sweep).  The persistent kernel hands trajectories to idle lanes
// in queue order, so with the caller's order (a) the kernel ends with a drain tail as long as the LONGEST
// trajectory that happened to start last (measured: 516 lane-steps of makespan against an ideal 427), and
// (b) the 32 lanes of a warp hold trajectories of unrelated length.  Sorting the queue by descending expected
// work fixes both (list scheduling, longest-processing-time-first): the last trajectories to start are the
// shortest, and neighbours in the queue take similar numbers of steps.
//
// The work proxy is the quantity the Hairer-Norsett-Wanner initial-step estimate is built from:
//   d1 = ||f(u0)/sk||,  d2 = ||(f(u0 + dt0 f0) - f0)/sk|| / dt0,   proxy = max(d1, d2)
// (2 RHS evaluations per trajectory against ~340 for the solve).  Spearman rank correlation with the true
// step count on the random Lorenz sweep: 0.965 (tools/work_proxy_check.py).  The order only decides WHEN a
// trajectory is integrated, never what is computed: results are bit-identical with and without it
// (tests/test_gpu_work_order.py), and Philox streams are keyed by the global trajectory index.
//
// Two kernels (counting sort on a 1024-bucket logarithmic key, 16 buckets per octave):
//   b2_work_keys    : proxy -> key[i], global histogram
//   b2_work_scatter : every block scans the histogram itself, ranks its tile in shared memory and reserves
//                     one range per non-empty bucket with a single global atomicAdd -> perm[position] = i
//
// WINDOWS.  The order is established inside windows of `window` consecutive trajectories (a multiple of the scatter
// tile; window >= N = one global sort), window after window.  A global permutation scatters the trajectories that are
// in flight at any moment over the WHOLE output array: every 32-byte sector of out_u then sits half-written in L2
// for the whole kernel, the array (132 MB for 1M Float32 Lorenz trajectories) does not fit, and sectors are evicted
// partially written and fetched back (ncu round 1/2: DRAM traffic 2.1x the algorithmic bytes).  With windows the
// rows being written at any time span one or two windows (tens of MB), sectors complete in L2 and travel to DRAM
// once.  Scheduling keeps what it needs: neighbours in the queue still take similar step counts, and the last
// trajectories to start are the shortest of the last window.
#pragma once
#include "b2_common.cuh"

#define B2_WB 1024          // buckets
#define B2_WTILE 4          // trajectories per thread in the scatter kernel (block = 1024 threads)

__device__ __forceinline__ unsigned b2_work_bucket(float proxy) {
    // descending order: bucket 0 = largest proxy (NaN / inf / negative garbage sort first: they fail fast)
    if (!(proxy >= 0.0f) || proxy > 3.0e38f) return 0u;
    const int q = (int)(__float_as_uint(proxy) >> 19) - ((127 - 32) << 4);   // 4 mantissa bits, octaves 2^-32..2^32
    const int k = q < 0 ? 0 : (q > B2_WB - 1 ? B2_WB - 1 : q);
    return (unsigned)(B2_WB - 1 - k);
}

// hist: [n_windows][B2_WB].  A block's 256 trajectories of one grid-stride iteration are consecutive, hence in ONE
// window (window is a multiple of 256): the block keeps the histogram of its current window in shared memory and
// flushes it when the window changes (block-uniform) and at the end.
extern "C" __global__ void __launch_bounds__(256) b2_work_keys(const __grid_constant__ B2Args a, unsigned short* __restrict__ keys,
                                                                  unsigned* __restrict__ hist, const long long window) {
    __shared__ unsigned sh[B2_WB];
    for (int i = threadIdx.x; i < B2_WB; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const sreal* const gu0 = reinterpret_cast<const sreal*>(a.u0);
    const sreal* const gp = reinterpret_cast<const sreal*>(a.p);
    const sreal t0 = B2_ARG(a, t0);
    long long cur_w = -1;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < a.N; base += (long long)gridDim.x * blockDim.x) {
        const long long w = base / window;
        if (w != cur_w) {   // block-uniform
            if (cur_w >= 0) {
                __syncthreads();
                for (int i = threadIdx.x; i < B2_WB; i += blockDim.x) {
                    if (sh[i]) atomicAdd(&hist[cur_w * B2_WB + i], sh[i]);
                    sh[i] = 0u;
                }
                __syncthreads();
            }
            cur_w = w;
        }
        const long long i = base + threadIdx.x;
        if (i >= a.N) continue;
        sreal u[B2_N], p[B2_NPA], f0[B2_N], f1[B2_N], u1[B2_N];
#pragma unroll
        for (int j = 0; j < B2_N; j++) u[j] = gu0[i * B2_N + j];
#pragma unroll
        for (int j = 0; j < B2_NPARAM; j++) p[j] = gp[i * B2_NPARAM + j];
        b2_rhs(f0, u, p, t0);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        float isk[B2_N];
#pragma unroll
        for (int j = 0; j < B2_N; j++) {
            isk[j] = __fdividef(1.0f, (float)b2_fma(b2_abs(u[j]), (sreal)B2_RTOL(a, j), (sreal)B2_ATOL(a, j)));
            const float r0 = (float)u[j] * isk[j], r1 = (float)f0[j] * isk[j];
            a0 = fmaf(r0, r0, a0);
            a1 = fmaf(r1, r1, a1);
        }
        const float d0 = sqrtf(a0 * (1.0f / B2_N)), d1 = sqrtf(a1 * (1.0f / B2_N));
        const float dt0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * __fdividef(d0, d1);
#pragma unroll
        for (int j = 0; j < B2_N; j++) u1[j] = b2_fma((sreal)dt0, f0[j], u[j]);
        b2_rhs(f1, u1, p, t0 + (sreal)dt0);
#pragma unroll
        for (int j = 0; j < B2_N; j++) {
            const float r2 = (float)(f1[j] - f0[j]) * isk[j];
            a2 = fmaf(r2, r2, a2);
        }
        const float d2 = __fdividef(sqrtf(a2 * (1.0f / B2_N)), dt0);
        const unsigned k = b2_work_bucket(fmaxf(d1, d2));
        keys[i] = (unsigned short)k;
        atomicAdd(&sh[k], 1u);
    }
    __syncthreads();
    if (cur_w >= 0)
        for (int i = threadIdx.x; i < B2_WB; i += blockDim.x)
            if (sh[i]) atomicAdd(&hist[cur_w * B2_WB + i], sh[i]);
}

// hist / cursor: [n_windows][B2_WB]; a tile (B2_WB * B2_WTILE consecutive trajectories) lies in one window (window is a
// multiple of the tile) and is ranked into that window's range of perm.
extern "C" __global__ void __launch_bounds__(B2_WB) b2_work_scatter(long long N, const unsigned short* __restrict__ keys,
                                                                      const unsigned* __restrict__ hist_all, unsigned* __restrict__ cursor_all,
                                                                      unsigned* __restrict__ perm, const long long window) {
    const long long win = ((long long)blockIdx.x * (B2_WB * B2_WTILE)) / window;
    const unsigned* const hist = hist_all + win * B2_WB;
    unsigned* const cursor = cursor_all + win * B2_WB;
    const unsigned wbase = (unsigned)(win * window);
    __shared__ unsigned offs[B2_WB];    // exclusive scan of the global histogram, then + this block's reserved base
    __shared__ unsigned cnt[B2_WB];     // this tile's histogram
    __shared__ unsigned wsum[32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    // block-wide exclusive scan of hist[0..1023] (one entry per thread)
    const unsigned h = hist[tid];
    unsigned x = h;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(B2_FULL, x, d);
        if (lane >= (unsigned)d) x += y;
    }
    if (lane == 31u) wsum[w] = x;
    cnt[tid] = 0u;
    __syncthreads();
    if (w == 0) {
        unsigned s = wsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(B2_FULL, s, d);
            if (lane >= (unsigned)d) s += y;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    const unsigned excl = x - h + (w ? wsum[w - 1] : 0u);
    // rank this block's tile inside shared memory
    const long long base_i = (long long)blockIdx.x * (B2_WB * B2_WTILE);
    unsigned k[B2_WTILE], r[B2_WTILE];
#pragma unroll
    for (int j = 0; j < B2_WTILE; j++) {
        const long long i = base_i + (long long)j * B2_WB + tid;
        k[j] = 0xffffffffu;
        if (i < N) {
            k[j] = keys[i];
            r[j] = atomicAdd(&cnt[k[j]], 1u);
        }
    }
    __syncthreads();
    // one global atomic per non-empty bucket reserves this tile's range inside the bucket
    const unsigned c = cnt[tid];
    offs[tid] = wbase + excl + (c ? atomicAdd(&cursor[tid], c) : 0u);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < B2_WTILE; j++) {
        const long long i = base_i + (long long)j * B2_WB + tid;
        if (i < N) perm[offs[k[j]] + r[j]] = (unsigned)i;
    }
}
