// HMM parameter construction on the device, in float64:
//   particle x (unconstrained)  ->  (t, c, rho)            reference: src/phlash/params.py:94-131
//   (t, c, rho, theta)          ->  b, d, u, v, emis0, emis1, pi
//                                                           reference: src/phlash/params.py:32-55,
//                                                           src/phlash/transition.py:9-85,
//                                                           src/phlash/size_history.py:17-22, 123-138, 170-193
// and the vector-Jacobian product that takes d l / d log(theta) [7, M] back to d l / d x.
//
// The reference builds the dense M x M matrix (an O(M^3) masked product for the upper triangle,
// transition.py:69-83) and then reads three diagonals and the first row off it
// (params.py:45-50).  Only those O(M) entries are computed here, and only ROW 0 of the cumulative
// 3x3 products (transition.py:51-56) is carried, so the state per thread is O(1).
//
// Differentiation: forward mode with ONE tangent direction per thread (Dual: value + derivative
// w.r.t. x[dir]); a launch over (particle, direction) gives the whole Jacobian, contracted on the
// fly with the incoming cotangent.  Selections (where / clip / isclose guards) pass the tangent of
// the selected branch and zero where a value is clipped, like JAX.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace phb {

struct Dual {
    double v, d;
};
__device__ __forceinline__ Dual mk(double v, double d = 0.0) { return Dual{v, d}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    const double q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual operator+(Dual a, double b) { return {a.v + b, a.d}; }
__device__ __forceinline__ Dual operator-(Dual a, double b) { return {a.v - b, a.d}; }
__device__ __forceinline__ Dual operator-(double a, Dual b) { return {a - b.v, -b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, double b) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ Dual operator*(double a, Dual b) { return {a * b.v, a * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, double b) { return {a.v / b, a.d / b}; }
__device__ __forceinline__ Dual operator/(double a, Dual b) { return mk(a) / b; }
__device__ __forceinline__ Dual dexp(Dual a) {
    const double e = exp(a.v);
    return {e, e * a.d};
}
__device__ __forceinline__ Dual dexpm1(Dual a) { return {expm1(a.v), exp(a.v) * a.d}; }
__device__ __forceinline__ Dual dsqrt(Dual a) {
    const double s = sqrt(a.v);
    return {s, a.d / (2.0 * s)};
}
__device__ __forceinline__ Dual dclip(Dual a, double lo, double hi) {
    if (a.v < lo) return {lo, 0.0};
    if (a.v > hi) return {hi, 0.0};
    return a;
}
__device__ __forceinline__ Dual dmax(Dual a, double lo) { return a.v >= lo ? a : mk(lo); }
// numpy.isclose(x, 0) with default tolerances: |x| <= 1e-8
__device__ __forceinline__ bool close_to_zero(double x) { return fabs(x) <= 1e-8; }

// 1 / expm1(x), stable for large x (size_history.py:17-22)
__device__ __forceinline__ Dual expm1inv(Dual x) {
    if (x.v > 10.0) return -dexp(-x) / dexpm1(-x);
    return 1.0 / dexpm1(x);
}

// row0 <- row0 * expQ(r, c, n = 2)   (transition.py:9-34; only the product with a row is needed)
__device__ __forceinline__ void apply_expQ(Dual (&row)[3], Dual r, Dual c) {
    const double n = 2.0;
    const Dual disc = dsqrt((c * n) * (c * n) - 2.0 * c * (n - 2.0) * r + r * r) / 2.0;
    const Dual mean = (r + c * n) / 2.0;
    const Dual half = (r - c * n) / 2.0;
    const Dual t1 = (dexp(disc - mean) + dexp(-(disc + mean))) / 2.0;
    Dual t2;
    if (disc.v < 1e-6) {
        // as written in the reference: the series term uses the substituted u_safe = 1
        t2 = dexp(-mean) * (1.0 + 1.0 / 6.0);
    } else {
        t2 = (dexp(disc - mean) - dexp(-(disc + mean))) / 2.0 / disc;
    }
    const Dual p11 = t1 - half * t2, p12 = r * t2, p21 = c * t2, p22 = t1 + half * t2;
    const Dual p13 = 1.0 - p11 - p12, p23 = 1.0 - p21 - p22;
    const Dual a = row[0], b = row[1];
    row[0] = a * p11 + b * p21;
    row[1] = a * p12 + b * p22;
    row[2] = a * p13 + b * p23 + row[2];
}

constexpr int kMaxM = 64;

// Everything for one particle and one tangent direction.  out[g * M + m] (value and tangent).
// widths: epoch widths of the PSMC pattern (sum = M); x: [2 + n_epochs + 1]; dir < 0: no tangent.
__device__ inline void particle_to_params(const double *x, int dir, const int *widths, int n_epochs, int M,
                                          double theta, Dual *out /* [7 * M] */) {
    auto in = [&](int i) { return mk(x[i], i == dir ? 1.0 : 0.0); };
    // ---- MCMCParams.to_dm (params.py:94-131)
    const Dual t1 = dexp(in(0));
    const Dual tM = t1 + dexp(in(1));
    const Dual log_ratio = mk(log(tM.v / t1.v), tM.d / tM.v - t1.d / t1.v);
    Dual t[kMaxM], c[kMaxM];
    t[0] = mk(0.0);
    for (int i = 1; i < M; ++i) {
        // geomspace(t1, tM, M - 1)[i - 1]; end points exact like numpy
        if (i == 1) t[i] = t1;
        else if (i == M - 1) t[i] = tM;
        else t[i] = t1 * dexp(log_ratio * (double(i - 1) / double(M - 2)));
    }
    {
        int m = 0;
        for (int e = 0; e < n_epochs; ++e) {
            const Dual z = in(2 + e);
            // softplus(z) = log1p(exp(z)), derivative sigmoid(z)
            const double sp = z.v > 30.0 ? z.v : log1p(exp(z.v));
            const Dual ce = mk(sp, z.d / (1.0 + exp(-z.v)));
            for (int w = 0; w < widths[e]; ++w) c[m++] = ce;
        }
    }
    Dual rho;
    {
        const Dual z = in(2 + n_epochs);
        const double sg = 1.0 / (1.0 + exp(-z.v));
        rho = mk(theta * (0.1 + 9.9 * sg), theta * 9.9 * sg * (1.0 - sg) * z.d);
    }
    // ---- SizeHistory.ect (size_history.py:170-193)
    Dual ect[kMaxM];
    for (int k = 0; k < M - 1; ++k) {
        const Dual dt = t[k + 1] - t[k];
        if (close_to_zero(c[k].v)) ect[k] = (t[k] + t[k + 1]) / 2.0;
        else if (isinf(c[k].v) || c[k].v > 100.0) ect[k] = t[k];
        else ect[k] = 1.0 / c[k] + t[k] - dt * expm1inv(c[k] * dt);
    }
    ect[M - 1] = t[M - 1] + 1.0 / c[M - 1];
    for (int k = 0; k < M; ++k) ect[k] = dmax(ect[k], 1e-20);
    // ---- emissions and pi (params.py:36-43, size_history.py:123-138)
    const double lo = 1e-20, hi = 1.0 - 1e-20;
    for (int k = 0; k < M; ++k) {
        const Dual ue = theta * ect[k];
        out[4 * M + k] = dclip(dexp(-ue), lo, hi);
        out[5 * M + k] = dclip(-dexpm1(-ue), lo, hi);
    }
    {
        // surv = [S_0, ..., S_{M-2}, 0] with S_k = exp(-sum_{i<=k} c_i dt_i);  pi[i] = surv[i-1] - surv[i]
        // for i >= 1 and pi[0] = 1 - sum of the others
        Dual hazard = mk(0.0), prev = mk(0.0), rest = mk(0.0);
        for (int k = 0; k < M - 1; ++k) {
            hazard = hazard + c[k] * (t[k + 1] - t[k]);
            const Dual s_k = dexp(-hazard);
            if (k >= 1) {
                out[6 * M + k] = prev - s_k;
                rest = rest + out[6 * M + k];
            }
            prev = s_k;
        }
        out[6 * M + (M - 1)] = prev;  // S_{M-2} - 0
        rest = rest + prev;
        out[6 * M + 0] = 1.0 - rest;
        for (int k = 0; k < M; ++k) out[6 * M + k] = dclip(out[6 * M + k], lo, hi);
    }
    // ---- transition bands (transition.py:37-85 restricted to what params.py:45-50 reads)
    Dual row[3] = {mk(1.0), mk(0.0), mk(0.0)};
    Dual sub[kMaxM], diag[kMaxM], p_float[kMaxM], p_pass[kMaxM], p_coal[kMaxM];
    Dual at_t2 = mk(0.0);  // P_t[k][0, 2]
    for (int k = 0; k < M; ++k) {
        // first half of interval k: t_k -> ect_k at rate c_k
        {
            const Dual step = ect[k] - t[k];
            if (!close_to_zero(step.v)) apply_expQ(row, 2.0 * step * rho, step * c[k]);
        }
        const Dual e0 = row[0], e1 = row[1], e2 = row[2];
        Dual back, stay;
        if (k < M - 1) {
            const Dual left = (t[k + 1] - ect[k]) * c[k];
            back = -dexpm1(-left);
            stay = dexp(-left);
        } else {
            back = mk(1.0);
            stay = mk(0.0);
        }
        diag[k] = e0 + e1 * back + e2 - at_t2;
        p_float[k] = dclip(e1 * stay, 1e-8, 1.0 - 1e-8);
        if (k < M - 1) {
            const Dual dtc = (t[k + 1] - t[k]) * c[k];
            p_pass[k] = dclip(dexp(-dtc), 1e-8, 1.0 - 1e-8);
            p_coal[k] = dclip(-dexpm1(-dtc), 1e-8, 1.0 - 1e-8);
            // second half: ect_k -> t_{k+1}
            const Dual step = t[k + 1] - ect[k];
            if (!close_to_zero(step.v)) apply_expQ(row, 2.0 * step * rho, step * c[k]);
            sub[k] = row[2] - at_t2;  // every entry below the diagonal in column k
            at_t2 = row[2];
        } else {
            p_pass[k] = dclip(mk(0.0), 1e-8, 1.0 - 1e-8);
            p_coal[k] = dclip(mk(1.0), 1e-8, 1.0 - 1e-8);
            // absorbing step: P_t[M][0, 2] = total mass of the row
            sub[k] = (row[0] + row[1] + row[2]) - at_t2;
        }
    }
    // ---- PSMCParams.from_dm (params.py:44-55): clip A, then b, d, u, v
    for (int k = 0; k < M; ++k) {
        out[0 * M + k] = k < M - 1 ? dclip(sub[k], lo, hi) : mk(0.0);
        out[1 * M + k] = dclip(diag[k], lo, hi);
    }
    // first row of A above the diagonal: A[0, j] = p_float[0] * prod_{0<l<j} p_pass[l] * p_coal[j]
    Dual run = p_float[0];
    const Dual a01 = dclip(run * p_coal[1], lo, hi);
    out[3 * M + 0] = mk(0.0);
    out[2 * M + (M - 1)] = mk(0.0);
    for (int j = 1; j < M; ++j) {
        const Dual a0j = dclip(run * p_coal[j], lo, hi);
        out[3 * M + j] = a0j / a01;  // v[j]
        run = run * p_pass[j];
    }
    for (int i = 0; i < M - 1; ++i) {
        const Dual sup = dclip(p_float[i] * p_coal[i + 1], lo, hi);  // A[i, i+1]
        out[2 * M + i] = sup / out[3 * M + i + 1];                   // u[i] = A[i, i+1] / v[i+1]
    }
}

struct ParamsArgs {
    const double *x;     // [B, P]
    int64_t B;
    int P;               // 2 + n_epochs + 1
    int n_epochs;
    int M;
    double theta;
    int widths[kMaxM];
    void *params7;       // [B, 7, M] FLOAT out (forward kernel)
    const void *cotangent;  // [B, 7, M] FLOAT: d l / d log(theta) (vjp kernel)
    double *grad_x;      // [B, P] out (vjp kernel)
    int out_double;      // FLOAT is double?
    // vjp kernel, whole-term entry: the cotangent is read as DOUBLE from per-particle sums
    // [B, cot_stride] (cot_stride > 0; element 0 of a row is the log-likelihood sum) and the result is
    // multiplied by `scale`
    int64_t cot_stride;
    double scale;
};

// one thread per particle
__global__ void psmc_params_forward_kernel(const ParamsArgs a) {
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    Dual out[7 * kMaxM];
    particle_to_params(a.x + b * a.P, -1, a.widths, a.n_epochs, a.M, a.theta, out);
    const int n = 7 * a.M;
    if (a.out_double) {
        double *dst = static_cast<double *>(a.params7) + b * n;
        for (int i = 0; i < n; ++i) dst[i] = out[i].v;
    } else {
        float *dst = static_cast<float *>(a.params7) + b * n;
        for (int i = 0; i < n; ++i) dst[i] = float(out[i].v);
    }
}

// one thread per (particle, direction): grad_x[b, p] = sum_{g,m} cot[g,m] * d log(theta[g,m]) / d x_p
__global__ void psmc_params_vjp_kernel(const ParamsArgs a) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.P) return;
    const int64_t b = idx / a.P;
    const int dir = int(idx % a.P);
    Dual out[7 * kMaxM];
    particle_to_params(a.x + b * a.P, dir, a.widths, a.n_epochs, a.M, a.theta, out);
    const int n = 7 * a.M;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double cot;
        if (a.cot_stride > 0) cot = static_cast<const double *>(a.cotangent)[b * a.cot_stride + 1 + i];
        else cot = a.out_double ? static_cast<const double *>(a.cotangent)[b * n + i]
                                : double(static_cast<const float *>(a.cotangent)[b * n + i]);
        if (out[i].v != 0.0) acc += cot * out[i].d / out[i].v;
    }
    a.grad_x[idx] = a.cot_stride > 0 ? a.scale * acc : acc;
}

// Per-particle sums over the chunks of a minibatch (the HMM term is additive over chunks once the
// warm-up is fused): sums[b, 0] = sum_s ll[b, s], sums[b, 1 + c] = sum_s dlog[b, s, c], c < 7 M.
// One CTA per particle, one thread per column; coalesced over c.
template <typename F>
__global__ void sum_over_chunks_kernel(const double *__restrict__ ll, const F *__restrict__ dlog, int64_t S, int C,
                                       double *__restrict__ sums) {
    const int64_t b = blockIdx.x;
    for (int c = threadIdx.x; c <= C; c += blockDim.x) {
        double acc = 0.0;
        if (c == 0) {
            for (int64_t s = 0; s < S; ++s) acc += ll[b * S + s];
        } else if (dlog != nullptr) {
            const F *col = dlog + (b * S) * C + (c - 1);
            for (int64_t s = 0; s < S; ++s) acc += double(col[s * C]);
        }
        sums[b * (C + 1) + c] = acc;
    }
}

// value[b] = scale * sums[b, 0]
__global__ void scaled_first_column_kernel(const double *__restrict__ sums, int64_t B, int64_t stride, double scale,
                                           double *__restrict__ value) {
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b < B) value[b] = scale * sums[b * stride];
}

}  // namespace phb
