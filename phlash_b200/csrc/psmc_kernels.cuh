// PSMC coalescent-HMM log-likelihood + gradient for sm_100a.
//
// What is computed (reference semantics): the scaled forward recursion of
// src/phlash/hmm.py:68-82 with the O(M) structured transition of hmm.py:52-65 / gpu.py:504-522,
// and the gradient contract of the reference's loglik_grad kernel (gpu.py:575-692, host roll
// :303-313).  How it is computed is new:
//
//  * thread-per-pair (T = 1) or T lanes per pair, each lane holding MT = M / T consecutive
//    states of the forward vector in registers.  The prefix / suffix sums of the structured
//    transition are serial FMA chains inside a lane, stitched across the T lanes of a pair with
//    log2(T) shuffles.  The lanes of a warp are different chunks of (normally) the same
//    particle, so the parameter block is read from shared memory as broadcast 128-bit loads.
//  * the gradient is the adjoint (backward) recursion, O(M) per site instead of the reference's
//    O(7 M^2) forward-mode sensitivities.  The forward vectors it needs are not kept for the whole
//    chunk: pass 1 stores a checkpoint every K sites to HBM (4*M/K bytes per site and pair);
//    pass 2 walks the segments backwards, re-runs the K forward steps of a segment into a
//    shared-memory ring private to the thread, then runs the K adjoint steps, accumulating
//    d ll / d (b, d, u, v, emis0, emis1) in registers.
//  * persistent grid: one CTA slot per resident CTA, looping over groups of pairs; checkpoint
//    scratch is indexed by CTA slot, so its size depends on the GPU, not on the problem.
//
// Per site and pair (M = 16): ~7.5 M forward + 7.5 M recompute + 16 M adjoint FMA-pipe
// instructions, 1 B of observations, 2 * 4 * M / K B of checkpoint traffic.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace phb {

constexpr int kThreads = 128;  // threads per CTA

struct KernelArgs {
    const int8_t *data;  // [N, pitch]
    int64_t pitch;       // bytes per row, multiple of 16
    int64_t n_rows;      // N
    int64_t L;           // sites per row
    const int64_t *inds; // [S] row of every chunk of the minibatch
    int64_t B, S;        // pairs are (b, s), s fastest
    const void *params6; // rows b,d,u,v,e0,e1 of pair (b,s) at + b*pstride_b + s*pstride_s
    int64_t pstride_b, pstride_s;
    const void *pi;      // [M] of pair (b,s) at + b*pistride_b + s*pistride_s
    int64_t pistride_b, pistride_s;
    double *ll;          // [B, S]
    void *dlog;          // [B, S, 7, M] or nullptr
    void *alpha_out;     // [B, S, M] filtered distribution after the last site, or nullptr
    void *ckpt;          // checkpoint scratch: gridDim.x * n_seg * M/T... see ckpt_elems()
    int64_t n_groups;    // ceil(B*S / pairs-per-CTA)
    int n_slots;         // parameter slots in shared memory per CTA
    int *err_flag;       // bit 0: index out of range, bit 1: non-finite result
};

template <typename F> struct Vec;
template <> struct Vec<float> {
    using type = float4;
    static constexpr int W = 4;
};
template <> struct Vec<double> {
    using type = double2;
    static constexpr int W = 2;
};

template <typename F> __device__ __forceinline__ void unpack(const float4 &v, F *o) {
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <typename F> __device__ __forceinline__ void unpack(const double2 &v, F *o) {
    o[0] = v.x; o[1] = v.y;
}
__device__ __forceinline__ float4 pack(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack(const double *o) { return make_double2(o[0], o[1]); }

// ---- sums across the T lanes that share a pair (lane index inside the pair = sub) ----
template <typename F, int T> __device__ __forceinline__ F lanes_before(F mine, int sub) {
    if constexpr (T == 1) {
        return F(0);
    } else {
        F incl = mine;
#pragma unroll
        for (int o = 1; o < T; o <<= 1) {
            F t = __shfl_up_sync(0xffffffffu, incl, o, T);
            if (sub >= o) incl += t;
        }
        F ex = __shfl_up_sync(0xffffffffu, incl, 1, T);
        return sub == 0 ? F(0) : ex;
    }
}
template <typename F, int T> __device__ __forceinline__ F lanes_after(F mine, int sub) {
    if constexpr (T == 1) {
        return F(0);
    } else {
        F incl = mine;
#pragma unroll
        for (int o = 1; o < T; o <<= 1) {
            F t = __shfl_down_sync(0xffffffffu, incl, o, T);
            if (sub + o < T) incl += t;
        }
        F ex = __shfl_down_sync(0xffffffffu, incl, 1, T);
        return sub == T - 1 ? F(0) : ex;
    }
}
template <typename F, int T> __device__ __forceinline__ F lanes_total(F mine) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o, T);
    return mine;
}

template <typename F> __device__ __forceinline__ F fast_rcp(F x);
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { return __frcp_rn(x); }
template <> __device__ __forceinline__ double fast_rcp<double>(double x) { return 1.0 / x; }
template <typename F> __device__ __forceinline__ F log2_of(F x);
template <> __device__ __forceinline__ float log2_of<float>(float x) { return log2f(x); }
template <> __device__ __forceinline__ double log2_of<double>(double x) { return log2(x); }

// Shared-memory image of one parameter slot: rows b, d, u, v, emis0, emis1, ones (7 * M values,
// the last three double as the emission look-up table indexed by the observation), padded so that
// different slots start in different banks.
template <int M> struct Slot {
    static constexpr int kRowB = 0, kRowD = M, kRowU = 2 * M, kRowV = 3 * M, kRowE = 4 * M;
    static constexpr int kStride = 7 * M + 4;
};

// One forward step for this lane's MT states: x <- (x A) .* emis(ob), returns sum over the pair.
template <typename F, int MT, int T>
__device__ __forceinline__ F forward_site(F (&x)[MT], const F *__restrict__ prm, int ob_row, int sub) {
    constexpr int M = MT * T;
    using V = typename Vec<F>::type;
    constexpr int W = Vec<F>::W;
    F pre[MT], suf[MT];
    F run = F(0);
#pragma unroll
    for (int q = 0; q < MT / W; ++q) {
        F u[W];
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowU + q * W), u);
#pragma unroll
        for (int r = 0; r < W; ++r) {
            pre[q * W + r] = run;
            run = fma(u[r], x[q * W + r], run);
        }
    }
    F tail = F(0);
#pragma unroll
    for (int k = MT - 1; k >= 0; --k) {
        suf[k] = tail;
        tail += x[k];
    }
    const F pre_off = lanes_before<F, T>(run, sub);
    const F suf_off = lanes_after<F, T>(tail, sub);
    F part[4] = {F(0), F(0), F(0), F(0)};
#pragma unroll
    for (int q = 0; q < MT / W; ++q) {
        F d[W], v[W], b[W], e[W];
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowD + q * W), d);
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowV + q * W), v);
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowB + q * W), b);
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowE + ob_row * M + q * W), e);
#pragma unroll
        for (int r = 0; r < W; ++r) {
            const int k = q * W + r;
            F o = d[r] * x[k];
            if constexpr (T == 1) {
                o = fma(v[r], pre[k], o);
                o = fma(b[r], suf[k], o);
            } else {
                o = fma(v[r], pre[k] + pre_off, o);
                o = fma(b[r], suf[k] + suf_off, o);
            }
            o *= e[r];
            x[k] = o;
            part[k & 3] += o;
        }
    }
    return lanes_total<F, T>((part[0] + part[1]) + (part[2] + part[3]));
}

// Gradient accumulators of one lane (d ll / d theta, not yet multiplied by theta, for b, d, u, v;
// already d / d log for the two emission rows).
template <typename F, int MT> struct Grad {
    F b[MT], d[MT], u[MT], v[MT], e0[MT], e1[MT];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < MT; ++k) b[k] = d[k] = u[k] = v[k] = e0[k] = e1[k] = F(0);
    }
};

// g.e{0,1} += x .* beta for the row selected by ob (nothing for a missing observation).
template <typename F, int MT>
__device__ __forceinline__ void posterior_to_emission(const F (&beta)[MT], const F (&x)[MT], int ob, Grad<F, MT> &g) {
    const F is0 = ob == 0 ? F(1) : F(0);
    const F is1 = ob == 1 ? F(1) : F(0);
#pragma unroll
    for (int k = 0; k < MT; ++k) {
        const F gam = x[k] * beta[k];
        g.e0[k] = fma(gam, is0, g.e0[k]);
        g.e1[k] = fma(gam, is1, g.e1[k]);
    }
}

// One adjoint step for site t.  On entry beta is the adjoint vector after site t (normalised so
// that beta . alpha_t == 1); x = forward vector before the site (alpha_{t-1}); inv_c = 1 / (forward
// normaliser of the site).  On exit beta is the adjoint vector before the site (beta . x == 1), and
// the posterior x .* beta of site t-1 has been added to the emission row of ob_prev (the
// observation at site t-1; pass -1 when there is none).
template <typename F, int MT, int T>
__device__ __forceinline__ void backward_site(F (&beta)[MT], const F (&x)[MT], F inv_c, int ob, int ob_prev,
                                              const F *__restrict__ prm, int sub, Grad<F, MT> &g) {
    constexpr int M = MT * T;
    using V = typename Vec<F>::type;
    constexpr int W = Vec<F>::W;
    const int ob_row = ob < 0 ? 2 : ob;
    F w[MT];
#pragma unroll
    for (int q = 0; q < MT / W; ++q) {
        F e[W];
        unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowE + ob_row * M + q * W), e);
#pragma unroll
        for (int r = 0; r < W; ++r) w[q * W + r] = (e[r] * inv_c) * beta[q * W + r];
    }
    if constexpr (T == 1) {
        // descending sweep: tails  Q_k = sum_{j>k} v_j w_j  and  S_k = sum_{j>k} x_j
        F q_run = F(0), s_run = F(0);
#pragma unroll
        for (int q = MT / W - 1; q >= 0; --q) {
            F v[W], u[W];
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowV + q * W), v);
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowU + q * W), u);
#pragma unroll
            for (int r = W - 1; r >= 0; --r) {
                const int k = q * W + r;
                beta[k] = u[r] * q_run;
                g.u[k] = fma(x[k], q_run, g.u[k]);
                g.b[k] = fma(s_run, w[k], g.b[k]);
                q_run = fma(v[r], w[k], q_run);
                s_run += x[k];
            }
        }
        // ascending sweep: heads  Pb_k = sum_{j<k} b_j w_j  and  Px_k = sum_{j<k} u_j x_j
        F b_run = F(0), x_run = F(0);
#pragma unroll
        for (int q = 0; q < MT / W; ++q) {
            F b[W], d[W], u[W];
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowB + q * W), b);
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowD + q * W), d);
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowU + q * W), u);
#pragma unroll
            for (int r = 0; r < W; ++r) {
                const int k = q * W + r;
                beta[k] = fma(d[r], w[k], beta[k] + b_run);
                g.d[k] = fma(x[k], w[k], g.d[k]);
                g.v[k] = fma(x_run, w[k], g.v[k]);
                b_run = fma(b[r], w[k], b_run);
                x_run = fma(u[r], x[k], x_run);
            }
        }
    } else {
        F qs[MT], ss[MT], bs[MT], xs[MT];
        F q_run = F(0), s_run = F(0), b_run = F(0), x_run = F(0);
#pragma unroll
        for (int q = MT / W - 1; q >= 0; --q) {
            F v[W];
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowV + q * W), v);
#pragma unroll
            for (int r = W - 1; r >= 0; --r) {
                const int k = q * W + r;
                qs[k] = q_run;
                ss[k] = s_run;
                q_run = fma(v[r], w[k], q_run);
                s_run += x[k];
            }
        }
#pragma unroll
        for (int q = 0; q < MT / W; ++q) {
            F b[W], u[W];
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowB + q * W), b);
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowU + q * W), u);
#pragma unroll
            for (int r = 0; r < W; ++r) {
                const int k = q * W + r;
                bs[k] = b_run;
                xs[k] = x_run;
                b_run = fma(b[r], w[k], b_run);
                x_run = fma(u[r], x[k], x_run);
            }
        }
        const F q_off = lanes_after<F, T>(q_run, sub);
        const F s_off = lanes_after<F, T>(s_run, sub);
        const F b_off = lanes_before<F, T>(b_run, sub);
        const F x_off = lanes_before<F, T>(x_run, sub);
#pragma unroll
        for (int q = 0; q < MT / W; ++q) {
            F d[W], u[W];
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowD + q * W), d);
            unpack<F>(*reinterpret_cast<const V *>(prm + Slot<M>::kRowU + q * W), u);
#pragma unroll
            for (int r = 0; r < W; ++r) {
                const int k = q * W + r;
                const F qk = qs[k] + q_off;
                beta[k] = fma(u[r], qk, fma(d[r], w[k], bs[k] + b_off));
                g.u[k] = fma(x[k], qk, g.u[k]);
                g.b[k] = fma(ss[k] + s_off, w[k], g.b[k]);
                g.d[k] = fma(x[k], w[k], g.d[k]);
                g.v[k] = fma(xs[k] + x_off, w[k], g.v[k]);
            }
        }
    }
    posterior_to_emission<F, MT>(beta, x, ob_prev, g);
}

// Observations of one K-site segment packed in 64-bit words (K = 8 or 16).
template <int K> struct ObsWords {
    uint64_t w[K / 8];
    __device__ __forceinline__ void load(const int8_t *row, int64_t site0) {
#pragma unroll
        for (int i = 0; i < K / 8; ++i) w[i] = __ldg(reinterpret_cast<const unsigned long long *>(row + site0) + i);
    }
    __device__ __forceinline__ int at(int k) const {
        uint64_t word = w[0];
        if constexpr (K == 16) word = (k & 8) ? w[1] : w[0];
        return static_cast<int>(static_cast<int8_t>((word >> ((k & 7) * 8)) & 0xff));
    }
};

template <typename F, int MT, int T, int K> constexpr size_t smem_bytes(int n_slots) {
    return sizeof(F) * (size_t(n_slots) * Slot<MT * T>::kStride + size_t(K) * MT * kThreads + size_t(K) * kThreads);
}
// checkpoint scratch elements (of F) per CTA slot
template <int MT, int K> constexpr int64_t ckpt_elems_per_cta(int64_t L) {
    return ((L + K - 1) / K) * int64_t(MT) * kThreads;
}

template <typename F, int MT, int T, int K, bool GRAD, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) psmc_loglik_kernel(const KernelArgs a) {
    constexpr int M = MT * T;
    constexpr int PB = kThreads / T;  // pairs per CTA
    using V = typename Vec<F>::type;
    constexpr int W = Vec<F>::W;
    constexpr int QN = MT / W;
    static_assert(MT % 4 == 0 && K % 8 == 0, "layout assumptions");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *prm_s = reinterpret_cast<F *>(smem_raw);
    V *seg_s = reinterpret_cast<V *>(prm_s + a.n_slots * Slot<M>::kStride);  // [K][QN][kThreads]
    F *invc_s = reinterpret_cast<F *>(seg_s + K * QN * kThreads);            // [K][kThreads]

    const int tid = threadIdx.x;
    const int sub = tid % T;
    const int lp = tid / T;
    const int64_t n_pairs = a.B * a.S;
    const int64_t n_seg = (a.L + K - 1) / K;
    const bool shared_params = a.pstride_s == 0;
    const F *params6 = static_cast<const F *>(a.params6);
    const F *pi_g = static_cast<const F *>(a.pi);
    V *ck = GRAD ? reinterpret_cast<V *>(static_cast<F *>(a.ckpt) + int64_t(blockIdx.x) * n_seg * MT * kThreads) : nullptr;

    for (int64_t grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
        const int64_t p_first = grp * PB;
        const int64_t p_last = min(p_first + PB, n_pairs) - 1;
        const int64_t b_first = p_first / a.S;
        __syncthreads();  // everyone is done with the previous group's slots
        {
            const int n_used = shared_params ? int(p_last / a.S - b_first) + 1 : int(p_last - p_first) + 1;
            for (int i = tid; i < n_used * 7 * M; i += kThreads) {
                const int slot = i / (7 * M), r = i % (7 * M);
                F val = F(1);  // row 6 = emission of a missing observation
                if (r < 6 * M) {
                    const int64_t pb = shared_params ? (b_first + slot) : (p_first + slot) / a.S;
                    const int64_t ps = shared_params ? 0 : (p_first + slot) % a.S;
                    val = params6[pb * a.pstride_b + ps * a.pstride_s + r];
                }
                prm_s[slot * Slot<M>::kStride + r] = val;
            }
        }
        __syncthreads();

        const int64_t pair = min(p_first + lp, n_pairs - 1);
        const bool writer = (p_first + lp) < n_pairs;
        const int64_t pb = pair / a.S, ps = pair % a.S;
        const int slot = shared_params ? int(pb - b_first) : int(pair - p_first);
        const F *prm = prm_s + slot * Slot<M>::kStride + sub * MT;
        int64_t row = a.inds[ps];
        if (row < 0 || row >= a.n_rows) {
            if (sub == 0) atomicOr(a.err_flag, 1);
            row = 0;
        }
        const int8_t *obs = a.data + row * a.pitch;
        const F *pi_p = pi_g + pb * a.pistride_b + ps * a.pistride_s + sub * MT;

        // ------------------------------------------------------------------ pass 1: forward
        F x[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) x[k] = pi_p[k];
        double ll = 0.0;
        for (int64_t seg = 0; seg < n_seg; ++seg) {
            if (GRAD && seg > 0) {
#pragma unroll
                for (int q = 0; q < QN; ++q) ck[(seg * QN + q) * kThreads + tid] = pack(&x[q * W]);
            }
            ObsWords<K> ow;
            ow.load(obs, seg * K);
            const int len = int(min(int64_t(K), a.L - seg * K));
            F acc = F(0);
#pragma unroll 2
            for (int k = 0; k < len; ++k) {
                const int ob = ow.at(k);
                const F tot = forward_site<F, MT, T>(x, prm, ob < 0 ? 2 : ob, sub);
                const F inv = fast_rcp<F>(tot);
#pragma unroll
                for (int j = 0; j < MT; ++j) x[j] *= inv;
                acc += log2_of<F>(tot);
            }
            ll += double(acc);
        }
        ll *= 0.69314718055994530942;
        if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
            if (sub == 0) atomicOr(a.err_flag, 2);
        }
        if (writer && sub == 0) a.ll[pair] = ll;
        if (writer && a.alpha_out != nullptr) {
            F *ao = static_cast<F *>(a.alpha_out) + pair * M + sub * MT;
#pragma unroll
            for (int k = 0; k < MT; ++k) ao[k] = x[k];
        }

        if constexpr (GRAD) {
            // -------------------------------------------------------------- pass 2: adjoint
            Grad<F, MT> g;
            g.clear();
            F beta[MT];
            {
                // after the last site: beta = 1 / sum(x) so that beta . x == 1, and the posterior of
                // the last site is x .* beta
                F tot = F(0);
#pragma unroll
                for (int k = 0; k < MT; ++k) tot += x[k];
                tot = fast_rcp<F>(lanes_total<F, T>(tot));
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] = tot;
                posterior_to_emission<F, MT>(beta, x, int(obs[a.L - 1]), g);
            }
            ObsWords<K> ow;
            ow.load(obs, (n_seg - 1) * K);
            for (int64_t seg = n_seg - 1; seg >= 0; --seg) {
                ObsWords<K> ow_prev = ow;
                if (seg > 0) ow_prev.load(obs, (seg - 1) * K);
                const int len = int(min(int64_t(K), a.L - seg * K));
                // re-run the forward steps of this segment, keeping every input vector
                F xs[MT];
                if (seg == 0) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) xs[k] = pi_p[k];
                } else {
#pragma unroll
                    for (int q = 0; q < QN; ++q) unpack<F>(ck[(seg * QN + q) * kThreads + tid], &xs[q * W]);
                }
#pragma unroll 2
                for (int k = 0; k < len; ++k) {
#pragma unroll
                    for (int q = 0; q < QN; ++q) seg_s[(k * QN + q) * kThreads + tid] = pack(&xs[q * W]);
                    const int ob = ow.at(k);
                    const F tot = forward_site<F, MT, T>(xs, prm, ob < 0 ? 2 : ob, sub);
                    const F inv = fast_rcp<F>(tot);
                    invc_s[k * kThreads + tid] = inv;
#pragma unroll
                    for (int j = 0; j < MT; ++j) xs[j] *= inv;
                }
                // xs is now the forward vector after the segment: re-impose beta . xs == 1
                // (controls round-off drift of the adjoint scaling)
                {
                    F dot = F(0);
#pragma unroll
                    for (int k = 0; k < MT; ++k) dot = fma(xs[k], beta[k], dot);
                    dot = fast_rcp<F>(lanes_total<F, T>(dot));
#pragma unroll
                    for (int k = 0; k < MT; ++k) beta[k] *= dot;
                }
#pragma unroll 2
                for (int k = len - 1; k >= 0; --k) {
                    F xin[MT];
#pragma unroll
                    for (int q = 0; q < QN; ++q) unpack<F>(seg_s[(k * QN + q) * kThreads + tid], &xin[q * W]);
                    const F inv_c = invc_s[k * kThreads + tid];
                    const int ob_prev = k > 0 ? ow.at(k - 1) : (seg > 0 ? ow_prev.at(K - 1) : -1);
                    backward_site<F, MT, T>(beta, xin, inv_c, ow.at(k), ob_prev, prm, sub, g);
                }
                ow = ow_prev;
            }
            if (writer) {
                F *out = static_cast<F *>(a.dlog) + pair * 7 * M + sub * MT;
#pragma unroll
                for (int k = 0; k < MT; ++k) {
                    out[0 * M + k] = g.b[k] * prm[Slot<M>::kRowB + k];
                    out[1 * M + k] = g.d[k] * prm[Slot<M>::kRowD + k];
                    out[2 * M + k] = g.u[k] * prm[Slot<M>::kRowU + k];
                    out[3 * M + k] = g.v[k] * prm[Slot<M>::kRowV + k];
                    out[4 * M + k] = g.e0[k];
                    out[5 * M + k] = g.e1[k];
                    out[6 * M + k] = beta[k] * pi_p[k];
                }
            }
        }
    }
}

}  // namespace phb
