// Adaptive Dormand-Prince 5(4) on the device: stage combination, error norm, step-size controller, dense output.
//
// Replaces odeint(func, z, [t0, t1], method="dopri5", rtol, atol) at flow_matching.py:79-84 (the default sampling
// path), :50-57 (solver="adaptive") and :172-179 (the adaptive tail of solver="fixadp").  torchdiffeq is not vendored
// by the reference; its algorithm is restated (see oracle/uvit_oracle.py::odeint_dopri5 for the CPU statement this is
// tested against): Hairer's starting step, rms error norm over the WHOLE state (the batch shares one step size),
// factor = min(10, max(0.9 / ratio^(1/5), ratio < 1 ? 1 : 0.2)), no clipping of the last step - the result is the
// 4th-order dense-output polynomial evaluated at t1.  Time is fp64, the state fp32, stage times are formed in fp32.
//
// All of it is HBM / latency work on a few MB (B * 16 KiB per buffer): plain grid-stride kernels, fp64 block partials
// reduced in a fixed order (deterministic, no atomics).  The velocity evaluations in between dominate by > 1000x.
#include "common.cuh"
#include "kernels.h"

namespace usp {
namespace {

// Butcher tableaus of torchdiffeq's adaptive explicit methods (k[0] = f(t0, y0); stage i evaluates at
// t0 + alpha[i-1] dt, y0 + dt sum_j beta[i-1][j] k[j]).  fsal: the last stage's state IS the solution (c_sol[:-1] ==
// beta[-1]); otherwise y1 = y0 + dt sum c_sol[j] k[j] is formed separately - and, as in torchdiffeq, the derivative
// carried into the next step is still the last stage's (for adaptive_heun that is f(t1, y0 + dt k0), not f(t1, y1)).
struct Tableau {
    int stages;    // evaluations per attempted step
    int order;     // step-size exponent 1 / order; the starting step uses the same exponent
    int fsal;
    float alpha[6];
    float beta[6][6];
    float c_sol[7];
    float c_err[7];
    float c_mid[7];   // y(t0 + dt/2) ~ y0 + dt sum c_mid[j] k[j], for the 4th-order Hermite-type dense output
};
#define D_(x) static_cast<float>(x)
__constant__ Tableau c_tab[3] = {
    // dopri5: Dormand-Prince 5(4) with Shampine's companion weights
    {6, 5, 1,
     {1.f / 5, 3.f / 10, 4.f / 5, 8.f / 9, 1.f, 1.f},
     {{1.f / 5, 0, 0, 0, 0, 0},
      {3.f / 40, 9.f / 40, 0, 0, 0, 0},
      {44.f / 45, -56.f / 15, 32.f / 9, 0, 0, 0},
      {19372.f / 6561, -25360.f / 2187, 64448.f / 6561, -212.f / 729, 0, 0},
      {9017.f / 3168, -355.f / 33, 46732.f / 5247, 49.f / 176, -5103.f / 18656, 0},
      {35.f / 384, 0.f, 500.f / 1113, 125.f / 192, -2187.f / 6784, 11.f / 84}},
     {35.f / 384, 0.f, 500.f / 1113, 125.f / 192, -2187.f / 6784, 11.f / 84, 0.f},
     {D_(35.0 / 384 - 1951.0 / 21600), 0.f, D_(500.0 / 1113 - 22642.0 / 50085), D_(125.0 / 192 - 451.0 / 720),
      D_(-2187.0 / 6784 + 12231.0 / 42400), D_(11.0 / 84 - 649.0 / 6300), D_(-1.0 / 60)},
     {D_(6025192743.0 / 30085553152.0 / 2), 0.f, D_(51252292925.0 / 65400821598.0 / 2),
      D_(-2691868925.0 / 45128329728.0 / 2), D_(187940372067.0 / 1594534317056.0 / 2),
      D_(-1776094331.0 / 19743644256.0 / 2), D_(11237099.0 / 235043384.0 / 2)}},
    // bosh3: Bogacki-Shampine 3(2)
    {3, 3, 1,
     {1.f / 2, 3.f / 4, 1.f, 0, 0, 0},
     {{1.f / 2, 0, 0, 0, 0, 0}, {0.f, 3.f / 4, 0, 0, 0, 0}, {2.f / 9, 1.f / 3, 4.f / 9, 0, 0, 0}, {0}, {0}, {0}},
     {2.f / 9, 1.f / 3, 4.f / 9, 0.f, 0, 0, 0},
     {D_(2.0 / 9 - 7.0 / 24), D_(1.0 / 3 - 1.0 / 4), D_(4.0 / 9 - 1.0 / 3), D_(-1.0 / 8), 0, 0, 0},
     {0.f, 0.5f, 0.f, 0.f, 0, 0, 0}},
    // adaptive_heun: Heun 2(1)
    {1, 2, 0,
     {1.f, 0, 0, 0, 0, 0},
     {{1.f, 0, 0, 0, 0, 0}, {0}, {0}, {0}, {0}, {0}},
     {0.5f, 0.5f, 0, 0, 0, 0, 0},
     {0.5f, -0.5f, 0, 0, 0, 0, 0},
     {0.5f, 0.f, 0, 0, 0, 0, 0}},
};
#undef D_

constexpr int RK_THREADS = 256;

inline int rk_blocks(long long n) {
    long long b = (n + RK_THREADS - 1) / RK_THREADS;
    if (b > RK_MAX_PARTIALS) b = RK_MAX_PARTIALS;
    return static_cast<int>(b < 1 ? 1 : b);
}

// The stage's model time and what the "%.2f"-keyed hooks do at it (libs/dissection.py:21-26, tools/utils_t2i.py:284).
// rint(double(t) * 100) is the digit python prints: fp32 values are never within 1e-12 of a .5 tie unless exactly on it.
__device__ void set_stage_time(const RkArgs& a, float s_stage) {
    const RkState* rs = a.rs;
    const float t = rs->sign * s_stage;
    StepState* st = a.st;
    st->t = t;
    st->dt = 0.f;
    const int idx = static_cast<int>(rint(static_cast<double>(t) * 100.0));
    const bool in = idx >= 0 && idx < RK_DIGITS;
    st->didx = in && idx < rs->n_rows ? idx : 0;
    if (a.eval_times != nullptr) {
        // "read": this evaluation dumps into its own trace row (the host names the files by "%.2f" of the time, in
        // evaluation order, so that later evaluations overwrite earlier ones like the reference's np.save does)
        const int c = *a.eval_count;
        st->didx = c < a.eval_cap ? c : -1;      // -1: no room, nothing is written; the host reports the overflow
        if (c < a.eval_cap) a.eval_times[c] = t;
        *a.eval_count = c + 1;
    }
    st->edit = (in && idx < rs->n_rows && a.emask[idx]) ? rs->write_scale : 0.f;
    st->attn_on = in ? a.amask[idx] : 0;
}

__global__ void __launch_bounds__(RK_THREADS) rk_stage_kernel(RkArgs a, int stage) {
    const RkState* rs = a.rs;
    const float s0f = static_cast<float>(rs->s0);
    float dtf, s_stage;
    const Tableau& T = c_tab[a.method];
    const bool solution = stage == RK_SOLUTION;    // y1 of a non-FSAL method: a combination only, no evaluation
    if (solution) {
        dtf = static_cast<float>(rs->dt);
        s_stage = 0.f;
    } else if (stage > 0) {
        dtf = static_cast<float>(rs->dt);
        // t1 = (t0 + dt) rounded once; inner stages t0 + alpha * dt in the state's precision
        s_stage = T.alpha[stage - 1] == 1.f ? static_cast<float>(rs->s0 + rs->dt)
                                            : __fadd_rn(s0f, __fmul_rn(T.alpha[stage - 1], dtf));
    } else if (stage == 0) {
        dtf = 0.f;
        s_stage = s0f;
    } else {
        dtf = rs->h0;
        s_stage = static_cast<float>(rs->s0 + static_cast<double>(rs->h0));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && !solution) set_stage_time(a, s_stage);
    if (stage == 0) return;
    float w[RK_STAGES];
    const int nk = solution ? T.stages + 1 : (stage > 0 ? stage : 1);
#pragma unroll
    for (int j = 0; j < RK_STAGES; ++j) {
        if (solution) w[j] = __fmul_rn(T.c_sol[j], dtf);
        else if (stage > 0) w[j] = j < 6 ? __fmul_rn(T.beta[stage - 1][j], dtf) : 0.f;
        else w[j] = dtf;
    }
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < RK_STAGES; ++j)
            if (j < nk) acc = fmaf(a.k[j * a.n + i], w[j], acc);
        a.ytmp[i] = a.y0[i] + acc;
    }
}

__device__ double block_sum(double v) {
    __shared__ double sh[RK_THREADS / 32];
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < RK_THREADS / 32; ++i) t += sh[i];
    return t;
}

// per-block partial sums of squares; what as in launch_rk_control
__global__ void __launch_bounds__(RK_THREADS) rk_norm_kernel(RkArgs a, int what) {
    const RkState* rs = a.rs;
    const float rtol = static_cast<float>(rs->rtol), atol = static_cast<float>(rs->atol);
    const float dtf = static_cast<float>(rs->dt);
    const Tableau& T = c_tab[a.method];
    double p0 = 0.0, p1 = 0.0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const float y0 = a.y0[i];
        if (what == 0) {
            const float scale = atol + fabsf(y0) * rtol;
            const float u = y0 / scale, v = a.k[i] / scale;
            p0 += static_cast<double>(u) * u;
            p1 += static_cast<double>(v) * v;
        } else if (what == 1) {
            const float scale = atol + fabsf(y0) * rtol;
            const float u = (a.k[a.n + i] - a.k[i]) / scale;
            p0 += static_cast<double>(u) * u;
        } else {
            float e = 0.f;
#pragma unroll
            for (int j = 0; j < RK_STAGES; ++j)
                if (j <= T.stages) e = fmaf(a.k[j * a.n + i], __fmul_rn(T.c_err[j], dtf), e);
            const float tol = atol + rtol * fmaxf(fabsf(y0), fabsf(a.ytmp[i]));
            const float u = e / tol;
            p0 += static_cast<double>(u) * u;
        }
    }
    p0 = block_sum(p0);
    p1 = block_sum(p1);
    if (threadIdx.x == 0) {
        a.partials[blockIdx.x] = p0;
        a.partials[RK_MAX_PARTIALS + blockIdx.x] = p1;
    }
}

__global__ void __launch_bounds__(RK_THREADS) rk_control_kernel(RkArgs a, int what, int nparts) {
    double p0 = 0.0, p1 = 0.0;
    for (int i = threadIdx.x; i < nparts; i += RK_THREADS) {
        p0 += a.partials[i];
        p1 += a.partials[RK_MAX_PARTIALS + i];
    }
    p0 = block_sum(p0);
    p1 = block_sum(p1);
    if (threadIdx.x != 0) return;
    RkState* rs = a.rs;
    const Tableau& T = c_tab[a.method];
    const double inv_n = 1.0 / static_cast<double>(a.n);
    if (what == 0) {
        const float d0 = static_cast<float>(sqrt(p0 * inv_n)), d1 = static_cast<float>(sqrt(p1 * inv_n));
        rs->d0 = d0;
        rs->d1 = d1;
        rs->h0 = fabsf((d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1);
        rs->nfe = 1;
    } else if (what == 1) {
        const float h0 = rs->h0, d1 = rs->d1;
        const float d2 = fabsf(static_cast<float>(sqrt(p0 * inv_n)) / h0);
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
        else h1 = powf(0.01f / fmaxf(d1, d2), 1.f / static_cast<float>(T.order));
        rs->dt = static_cast<double>(fminf(100.f * h0, fabsf(h1)));
        rs->nfe = 2;
    } else {
        const float ratio = static_cast<float>(sqrt(p0 * inv_n));
        const bool accept = ratio <= 1.f;     // a NaN ratio rejects, and the NaN factor below poisons dt: the host stops
        const double dt = rs->dt, s0 = rs->s0;
        double factor;
        if (ratio == 0.f) factor = 10.0;
        else {
            const double dfac = ratio < 1.f ? 1.0 : 0.2;
            factor = fmin(10.0, fmax(0.9 / pow(static_cast<double>(ratio), 1.0 / T.order), dfac));
        }
        rs->ratio = ratio;
        rs->accept = accept ? 1 : 0;
        rs->s0_prev = s0;
        rs->dt_prev = dt;
        rs->nfe += T.stages;
        if (accept) {
            rs->n_accept += 1;
            if (s0 + dt >= rs->s_end) rs->done = 1;
            rs->s0 = s0 + dt;
        } else {
            rs->n_reject += 1;
        }
        rs->dt = dt * factor;
    }
}

__global__ void __launch_bounds__(RK_THREADS) rk_commit_kernel(RkArgs a) {
    const RkState* rs = a.rs;
    if (!rs->accept) return;
    const Tableau& T = c_tab[a.method];
    const long long last = static_cast<long long>(T.stages) * a.n;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (!rs->done) {
        for (long long i = i0; i < a.n; i += stride) {
            a.y0[i] = a.ytmp[i];
            a.k[i] = a.k[last + i];
        }
        return;
    }
    const float dtf = static_cast<float>(rs->dt_prev);
    const float x = static_cast<float>((rs->s_end - rs->s0_prev) / rs->dt_prev);
    for (long long i = i0; i < a.n; i += stride) {
        const float y0 = a.y0[i], y1 = a.ytmp[i], f0 = a.k[i], f1 = a.k[last + i];
        float m = 0.f;
#pragma unroll
        for (int j = 0; j < RK_STAGES; ++j)
            if (j <= T.stages) m = fmaf(a.k[j * a.n + i], __fmul_rn(T.c_mid[j], dtf), m);
        const float ym = y0 + m;
        const float ca = 2.f * dtf * (f1 - f0) - 8.f * (y1 + y0) + 16.f * ym;
        const float cb = dtf * (5.f * f0 - 3.f * f1) + 18.f * y0 + 14.f * y1 - 32.f * ym;
        const float cc = dtf * (f1 - 4.f * f0) - 11.f * y0 - 5.f * y1 + 16.f * ym;
        const float cd = dtf * f0;
        float tot = y0 + x * cd;
        float xp = x * x;
        tot += xp * cc;
        xp *= x;
        tot += xp * cb;
        xp *= x;
        tot += xp * ca;
        a.out[i] = tot;
    }
}

}  // namespace

int rk_stages(int method) {
    static const int n[3] = {6, 3, 1};
    return method >= 0 && method < 3 ? n[method] : 0;
}
bool rk_fsal(int method) { return method != 2; }

cudaError_t launch_rk_stage(const RkArgs& a, int stage, cudaStream_t s) {
    if (a.method < 0 || a.method > 2) return cudaErrorInvalidValue;
    if (stage != RK_SOLUTION && (stage < -1 || stage > rk_stages(a.method))) return cudaErrorInvalidValue;
    rk_stage_kernel<<<stage == 0 ? 1 : rk_blocks(a.n), RK_THREADS, 0, s>>>(a, stage);
    return cudaGetLastError();
}

cudaError_t launch_rk_control(const RkArgs& a, int what, cudaStream_t s) {
    if (what < 0 || what > 2) return cudaErrorInvalidValue;
    const int nb = rk_blocks(a.n);
    rk_norm_kernel<<<nb, RK_THREADS, 0, s>>>(a, what);
    rk_control_kernel<<<1, RK_THREADS, 0, s>>>(a, what, nb);
    return cudaGetLastError();
}

cudaError_t launch_rk_commit(const RkArgs& a, cudaStream_t s) {
    rk_commit_kernel<<<rk_blocks(a.n), RK_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace usp
