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
// 3x3 products (transition.py:51-56) is carried.
//
// Differentiation: forward mode with ONE tangent direction per warp (Dual: value + derivative
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

// expQ(r, c, n = 2) (transition.py:9-34): the four entries that are not implied by unit row sums
struct ExpQ {
    Dual p11, p12, p21, p22;
};
__device__ __forceinline__ ExpQ identity_expQ() { return ExpQ{mk(1.0), mk(0.0), mk(0.0), mk(1.0)}; }
__device__ __forceinline__ ExpQ expQ_matrix(Dual r, Dual c) {
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
    return ExpQ{t1 - half * t2, r * t2, c * t2, t1 + half * t2};
}
// row <- row * P  (the third row of P is (0, 0, 1))
__device__ __forceinline__ void apply_expQ(Dual (&row)[3], const ExpQ &m) {
    const Dual p13 = 1.0 - m.p11 - m.p12, p23 = 1.0 - m.p21 - m.p22;
    const Dual a = row[0], b = row[1];
    row[0] = a * m.p11 + b * m.p21;
    row[1] = a * m.p12 + b * m.p22;
    row[2] = a * p13 + b * p23 + row[2];
}

constexpr int kMaxM = 64;

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

// One group of LPI = min(32, M rounded up to a power of two) lanes per particle (forward) or per (particle, tangent
// direction) (VJP) - a whole warp from M = 32 on, two items per warp at M = 16; lane l of the group owns states
// l, l + 32.
//
// The expensive part of the construction - two 3x3 matrix exponentials per state, each a square root and four
// double-precision exponentials on dual numbers - does not depend on the other states: the lanes compute them side by
// side.  What is sequential (the running row of the cumulative 3x3 products, the cumulative hazard, the running
// product of the pass probabilities) is a walk over M states with a handful of multiply-adds each; every lane walks
// it redundantly from shared memory.  The first version ran the whole chain in one thread per particle / direction:
// 0.10 + 0.13 ms of every likelihood step, a fifth of the step on eight GPUs (profiles/r02_launches_step_S1.csv).
constexpr int kParamArrays = 22;   // Dual arrays of M entries per warp, see the offsets below
constexpr int kStatesPerLane = kMaxM / 32;
inline int params_lanes_per_item(int M) { return M >= 32 ? 32 : M > 8 ? 16 : M > 4 ? 8 : 4; }
inline int params_warps_per_cta(int M) { return M > 32 ? 2 : 4; }
inline int params_items_per_cta(int M) { return params_warps_per_cta(M) * (32 / params_lanes_per_item(M)); }
inline size_t params_smem_bytes(int M) { return size_t(params_items_per_cta(M)) * kParamArrays * M * sizeof(Dual); }

template <bool VJP> __global__ void psmc_params_warp_kernel(const ParamsArgs a) {
    extern __shared__ __align__(16) unsigned char params_smem[];
    const int M = a.M, n_epochs = a.n_epochs;
    const int lpi = M >= 32 ? 32 : M > 8 ? 16 : M > 4 ? 8 : 4;  // lanes per item
    const int lane = threadIdx.x % lpi, slot = threadIdx.x / lpi;
    const int64_t n_items = VJP ? a.B * a.P : a.B;
    const int64_t item_raw = int64_t(blockIdx.x) * (blockDim.x / lpi) + slot;
    const bool valid = item_raw < n_items;
    const int64_t item = valid ? item_raw : n_items - 1;  // idle groups shadow the last item (the warp synchronises as a whole)
    const int64_t b = VJP ? item / a.P : item;
    const int dir = VJP ? int(item % a.P) : -1;
    const double theta = a.theta;
    const double *x = a.x + b * a.P;
    auto in = [&](int i) { return mk(x[i], i == dir ? 1.0 : 0.0); };
    Dual *sm = reinterpret_cast<Dual *>(params_smem) + size_t(slot) * kParamArrays * M;
    Dual *T_ = sm, *C_ = sm + M, *HZ = sm + 2 * M, *PP = sm + 3 * M, *PC = sm + 4 * M, *PPFX = sm + 5 * M, *E0 = sm + 6 * M,
         *E1 = sm + 7 * M, *E2 = sm + 8 * M, *R2 = sm + 9 * M, *PF = sm + 10 * M, *SV = sm + 11 * M, *V_ = sm + 12 * M,
         *P1 = sm + 13 * M /* [4][M] */, *P2 = sm + 17 * M /* [4][M] */;
    // (22 = 13 + 4 + 4 + 1 spare)

    // ---- MCMCParams.to_dm (params.py:94-131): every lane computes the scalars, its own states of t and c
    const Dual t1 = dexp(in(0));
    const Dual tM = t1 + dexp(in(1));
    const Dual log_ratio = mk(log(tM.v / t1.v), tM.d / tM.v - t1.d / t1.v);
    Dual rho;
    {
        const Dual z = in(2 + n_epochs);
        const double sg = 1.0 / (1.0 + exp(-z.v));
        rho = mk(theta * (0.1 + 9.9 * sg), theta * 9.9 * sg * (1.0 - sg) * z.d);
    }
    for (int k = lane; k < M; k += 32) {
        // geomspace(t1, tM, M - 1)[k - 1]; end points exact like numpy
        Dual tk;
        if (k == 0) tk = mk(0.0);
        else if (k == 1) tk = t1;
        else if (k == M - 1) tk = tM;
        else tk = t1 * dexp(log_ratio * (double(k - 1) / double(M - 2)));
        T_[k] = tk;
        int e = 0, upto = a.widths[0];
        while (k >= upto && e + 1 < n_epochs) upto += a.widths[++e];
        const Dual z = in(2 + e);
        // softplus(z) = log1p(exp(z)), derivative sigmoid(z)
        const double sp = z.v > 30.0 ? z.v : log1p(exp(z.v));
        C_[k] = mk(sp, z.d / (1.0 + exp(-z.v)));
    }
    __syncwarp();

    // ---- per state: ect (size_history.py:170-193), emissions (params.py:36-43), the two half-interval
    //      exponentials and the pass / coalescence probabilities of the interval (transition.py:37-85)
    const double lo = 1e-20, hi = 1.0 - 1e-20;
    Dual out[kStatesPerLane][7];
    Dual back[kStatesPerLane], stay[kStatesPerLane];
#pragma unroll
    for (int s = 0; s < kStatesPerLane; ++s) {
        const int k = lane + 32 * s;
        if (k >= M) continue;
        const Dual tk = T_[k], ck = C_[k];
        const bool last = k == M - 1;
        const Dual tk1 = last ? tk : T_[k + 1];
        Dual ect;
        if (last) {
            ect = tk + 1.0 / ck;
        } else {
            const Dual dt = tk1 - tk;
            if (close_to_zero(ck.v)) ect = (tk + tk1) / 2.0;
            else if (isinf(ck.v) || ck.v > 100.0) ect = tk;
            else ect = 1.0 / ck + tk - dt * expm1inv(ck * dt);
        }
        ect = dmax(ect, 1e-20);
        const Dual ue = theta * ect;
        out[s][4] = dclip(dexp(-ue), lo, hi);
        out[s][5] = dclip(-dexpm1(-ue), lo, hi);
        // first half of the interval: t_k -> ect_k at rate c_k
        ExpQ m1 = identity_expQ();
        {
            const Dual step = ect - tk;
            if (!close_to_zero(step.v)) m1 = expQ_matrix(2.0 * step * rho, step * ck);
        }
        P1[0 * M + k] = m1.p11, P1[1 * M + k] = m1.p12, P1[2 * M + k] = m1.p21, P1[3 * M + k] = m1.p22;
        ExpQ m2 = identity_expQ();
        if (!last) {
            const Dual left = (tk1 - ect) * ck;
            back[s] = -dexpm1(-left);
            stay[s] = dexp(-left);
            const Dual dtc = (tk1 - tk) * ck;
            HZ[k] = dtc;
            PP[k] = dclip(dexp(-dtc), 1e-8, 1.0 - 1e-8);
            PC[k] = dclip(-dexpm1(-dtc), 1e-8, 1.0 - 1e-8);
            // second half: ect_k -> t_{k+1}
            const Dual step = tk1 - ect;
            if (!close_to_zero(step.v)) m2 = expQ_matrix(2.0 * step * rho, step * ck);
        } else {
            back[s] = mk(1.0);
            stay[s] = mk(0.0);
            HZ[k] = mk(0.0);
            PP[k] = dclip(mk(0.0), 1e-8, 1.0 - 1e-8);
            PC[k] = dclip(mk(1.0), 1e-8, 1.0 - 1e-8);
        }
        P2[0 * M + k] = m2.p11, P2[1 * M + k] = m2.p12, P2[2 * M + k] = m2.p21, P2[3 * M + k] = m2.p22;
    }
    __syncwarp();

    // ---- the sequential part, walked by every lane: row 0 of the cumulative products (E0..E2 after the first half
    //      of interval k, R2 = its third entry after the second half), the cumulative hazard, the running product of
    //      the pass probabilities (PPFX[j] = prod_{0 < l < j} p_pass[l])
    {
        Dual row[3] = {mk(1.0), mk(0.0), mk(0.0)};
        Dual hazard = mk(0.0), run = mk(1.0);
        for (int k = 0; k < M; ++k) {
            apply_expQ(row, ExpQ{P1[0 * M + k], P1[1 * M + k], P1[2 * M + k], P1[3 * M + k]});
            const Dual r0 = row[0], r1 = row[1], r2 = row[2];
            if (k < M - 1) apply_expQ(row, ExpQ{P2[0 * M + k], P2[1 * M + k], P2[2 * M + k], P2[3 * M + k]});
            hazard = hazard + HZ[k];
            const Dual run_k = run;
            if (k >= 1) run = run * PP[k];
            if (lane == 0) {
                E0[k] = r0, E1[k] = r1, E2[k] = r2;
                R2[k] = row[2];
                SV[k] = hazard;  // the survival function is exponentiated below, in parallel
                PPFX[k] = run_k;
            }
        }
    }
    __syncwarp();

    // ---- per state again: transition bands, survival
#pragma unroll
    for (int s = 0; s < kStatesPerLane; ++s) {
        const int k = lane + 32 * s;
        if (k >= M) continue;
        const Dual e0 = E0[k], e1 = E1[k], e2 = E2[k];
        const Dual at_t2 = k > 0 ? R2[k - 1] : mk(0.0);  // P_t[k][0, 2]
        const Dual diag = e0 + e1 * back[s] + e2 - at_t2;
        PF[k] = dclip(e1 * stay[s], 1e-8, 1.0 - 1e-8);
        // every entry below the diagonal in column k; absorbing last step: the total mass of the row
        const Dual sub = k < M - 1 ? R2[k] - at_t2 : (e0 + e1 + e2) - at_t2;
        // ---- PSMCParams.from_dm (params.py:44-55): clip A, then b, d, u, v
        out[s][0] = k < M - 1 ? dclip(sub, lo, hi) : mk(0.0);
        out[s][1] = dclip(diag, lo, hi);
    }
    // survival S_k = exp(-hazard up to interval k) in place (each entry is read and written by its own lane)
    for (int k = lane; k < M - 1; k += 32) SV[k] = dexp(-SV[k]);
    __syncwarp();

    // ---- pi (size_history.py:123-138): pi[i] = S_{i-1} - S_i, pi[M-1] = S_{M-2}, pi[0] = 1 - sum of the others
    Dual rest = mk(0.0);
    for (int k = 1; k < M - 1; ++k) rest = rest + (SV[k - 1] - SV[k]);
    rest = rest + SV[M - 2];
    // first row of A above the diagonal: A[0, j] = p_float[0] * prod_{0<l<j} p_pass[l] * p_coal[j]
    const Dual a01 = dclip(PF[0] * PC[1], lo, hi);
#pragma unroll
    for (int s = 0; s < kStatesPerLane; ++s) {
        const int k = lane + 32 * s;
        if (k >= M) continue;
        Dual pi;
        if (k == 0) pi = 1.0 - rest;
        else if (k == M - 1) pi = SV[M - 2];
        else pi = SV[k - 1] - SV[k];
        out[s][6] = dclip(pi, lo, hi);
        Dual v = mk(0.0);
        if (k >= 1) v = dclip(PF[0] * PPFX[k] * PC[k], lo, hi) / a01;
        out[s][3] = v;
        V_[k] = v;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kStatesPerLane; ++s) {
        const int k = lane + 32 * s;
        if (k >= M) continue;
        // u[i] = A[i, i+1] / v[i+1]
        out[s][2] = k < M - 1 ? dclip(PF[k] * PC[k + 1], lo, hi) / V_[k + 1] : mk(0.0);
    }

    // ---- results
    if constexpr (!VJP) {
#pragma unroll
        for (int s = 0; s < kStatesPerLane; ++s) {
            const int k = lane + 32 * s;
            if (k >= M || !valid) continue;
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                if (a.out_double) static_cast<double *>(a.params7)[(b * 7 + g) * M + k] = out[s][g].v;
                else static_cast<float *>(a.params7)[(b * 7 + g) * M + k] = float(out[s][g].v);
            }
        }
    } else {
        // grad_x[b, dir] = sum_{g,m} cot[g,m] * d log(theta[g,m]) / d x_dir
        double acc = 0.0;
        const int n = 7 * M;
#pragma unroll
        for (int s = 0; s < kStatesPerLane; ++s) {
            const int k = lane + 32 * s;
            if (k >= M) continue;
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                const int i = g * M + k;
                double cot;
                if (a.cot_stride > 0) cot = static_cast<const double *>(a.cotangent)[b * a.cot_stride + 1 + i];
                else cot = a.out_double ? static_cast<const double *>(a.cotangent)[b * n + i]
                                        : double(static_cast<const float *>(a.cotangent)[b * n + i]);
                if (out[s][g].v != 0.0) acc += cot * out[s][g].d / out[s][g].v;
            }
        }
        for (int o = lpi / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && valid) a.grad_x[item] = a.cot_stride > 0 ? a.scale * acc : acc;
    }
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
