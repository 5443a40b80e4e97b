// Thread-per-pair throughput kernel at M = 16 (float) on the RESCALED recursion: 4 fused multiply-adds per state
// and site in the forward step instead of 6 instructions.  OPT-IN (PHB_SFORM=1): correct, 15 % fewer executed
// instructions, but measured SLOWER than psmc_loglik_kernel - kept as the record of that experiment (tested by
// tests/test_gpu_sform.py).
//
// The reference recursion (src/phlash/hmm.py:68-82) is  x_t = e(ob_t) .* (x_{t-1} A)  with the structured
// A = strictLower(1 b^T) + diag(d) + strictUpper(u v^T)  (hmm.py:52-65).  psmc_loglik_kernel evaluates it literally:
// per state two chain steps, three products for the combination, one for the emission.  Here the carried
// vector is  y = x / (e0 .* s),  s = (1, v_1, ..., v_{M-1}),  i.e. the vector divided by the ob = 0 emission and by
// the column factor of the rank-one upper triangle.  One site then is
//
//     z_k = beta_k S_k + delta_k y_k + P_k,     S_k = sum_{i>k} sigma_i y_i,     P_k = sum_{i<k} pi_i y_i
//     y'  = z                       for ob = 0 (9 sites in 10)
//     y'  = z .* (e(ob) / e0)       otherwise (ratio rows from a per-thread shared-memory table)
//
// with the coefficients  beta = b / s,  sigma = e0 s,  delta = d e0,  pi = u e0 s  in registers (64, like b, d, u, v):
// two chain FMAs + two combining FMAs per state, no separate emission multiply, and the mass  sum_k x_k = S_{-1}
// falls out of the S chain, so the rescaling needs no extra sum.  The adjoint step has the same FMA count as before
// but needs neither the emission look-up nor the posterior accumulation: every posterior is the sum of the three
// ways of arriving in a state, so the total posterior mass of a state is  dlog b + dlog d + dlog v  and only the
// sites with ob != 0 (where the ratio branch runs anyway) are booked separately:
//     dlog emis1 = Gamma_1,     dlog emis0 = (dlog b + dlog d + dlog v) - Gamma_1 - Gamma_missing.
// A full block of four ob = 0 sites (3 blocks in 4; the lanes of a warp share the chunk, so the test is uniform) runs
// as one straight-line basic block without any ratio code; other blocks go through a compact per-site loop.
//
// Measured on B200 at the benchmark shape (profiles/r02_ncu_sform_summary.md): 394 executed instructions per pair and
// site (psmc_loglik_kernel: 466; forward 84, recompute 88, adjoint 206) but 311 ms against 285 ms: the adjoint pass
// issues at 0.9 instructions per scheduler and clock, the two lean forward passes at 0.4 - with the emission multiply,
// the separate d * x and the suffix adds gone there is nothing left to issue between the steps of the two 16-deep FMA
// chains, and the instructions that were removed had been filling exactly those slots.  Cutting the chains in halves
// (depth 8, 16 more instructions per site) made it slower still (4.2e10 against 4.7e10 site-transitions/s); the same
// step in transfer_rows_kernel (three warps per scheduler) lost 4 % as well.
//
// Accuracy.  delta = fl(d e0) carries a SYSTEMATIC relative error of up to 6e-8 on the stay weight of a state, which
// the gradient amplifies by two orders of magnitude (tests/test_scaled_recursion_math.py); measured against the
// fp64 CPU restatement at 50 000 bins: ll 1.4e-7, gradient 1.2e-5 - inside the 1e-5 / 1e-4 bars, where psmc_loglik_kernel is at
// 1e-7 / 1.4e-6.
//
// Domain: v_k > 0 for k >= 1 and emis0_k > 0 (PSMCParams.from_dm clips everything to [1e-20, 1 - 1e-20],
// params.py:44-47).  A violation sets bit 2 of the error flag and makes the pair's ll NaN; phb_sync reports it.
//
// Everything else - persistent grid, chunk-major pair enumeration, checkpoints every K sites + recompute into a
// per-lane shared-memory ring, 1024-site fp32 windows flushed into fp64 slots, segment mode - is psmc_loglik_kernel's.
#pragma once

#include "psmc_kernels.cuh"

namespace phb {

constexpr int kSM = 16;  // hidden states of this kernel

struct SCoef {
    float beta[kSM], sigma[kSM], delta[kSM], pic[kSM];
};

// one transition on the rescaled vector (no emission: folded into the coefficients); returns sum_k x_k of the INPUT
// (cutting each chain in halves that run side by side - depth 8 instead of 16 at 16 more instructions - was
// measured slower: 4.2e10 against 4.7e10 site-transitions/s)
__device__ __forceinline__ float sform_step(float (&y)[kSM], const SCoef &c) {
    float part[kSM], suf[kSM];
    float P = 0.f, S = 0.f;
#pragma unroll
    for (int i = 0; i < kSM; ++i) {
        const int k = i, j = kSM - 1 - i;
        part[k] = fmaf(c.delta[k], y[k], P);
        P = fmaf(c.pic[k], y[k], P);
        suf[j] = S;
        S = fmaf(c.sigma[j], y[j], S);
    }
#pragma unroll
    for (int k = 0; k < kSM; ++k) y[k] = fmaf(c.beta[k], suf[k], part[k]);
    return S;
}

// Per-thread shared-memory rows: ratio rows emis1 / emis0 and 1 / emis0 (applied at sites with ob != 0), and the
// posterior accumulators of those sites; [row][QN][NT] 128-bit words like EmisTable, `base` = this thread's column.
template <int NT> struct RatioRows {
    static constexpr int QN = kSM / 4;
    uint32_t base;
    __device__ __forceinline__ void store(int row, const float (&v)[kSM]) const {
#pragma unroll
        for (int q = 0; q < QN; ++q) sts_word(base + (row * QN + q) * NT * 16, &v[q * 4]);
    }
    __device__ __forceinline__ void load(int row, float (&v)[kSM]) const {
#pragma unroll
        for (int q = 0; q < QN; ++q) lds_word(base + (row * QN + q) * NT * 16, &v[q * 4]);
    }
};

// one site of the forward passes: transition, then the emission ratio of the site's observation (ob != 0 only);
// returns the mass of the vector that ENTERED the transition
template <int NT> __device__ __forceinline__ float sform_site(float (&y)[kSM], const SCoef &c, const RatioRows<NT> &ratio, int ob) {
    const float tot = sform_step(y, c);
    if (ob != 0) {
        float r[kSM];
        ratio.load(ob < 0 ? 1 : 0, r);
#pragma unroll
        for (int k = 0; k < kSM; ++k) y[k] *= r[k];
    }
    return tot;
}

template <int K, int NT, bool GRAD> constexpr size_t sform_smem_bytes() {
    // ratio rows [2] | ring [K] | block scales [K / kNorm] scalars | posterior accumulators of ob = 1 / missing [2]
    return sizeof(float) * (size_t(2) * kSM * NT + (GRAD ? size_t(K) * kSM * NT + size_t(K / kNorm) * NT + size_t(2) * kSM * NT : 0));
}

template <int K, bool GRAD, int NT, int MINB, bool SEG = false>
__global__ void __maxnreg__(max_regs(NT, MINB)) psmc_sform_kernel(const KernelArgs a) {
    using F = float;
    constexpr int MT = kSM, M = kSM, W = 4, QN = MT / W, kWarps = NT / 32, PW = 32;
    using V = float4;
    static_assert(K % 8 == 0 && K % kNorm == 0, "layout assumptions");
    static_assert(!SEG || GRAD, "segment mode: gradient kernel");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // shared-memory map in 128-bit words: ratio rows [2][QN][NT] | ring per warp [K][QN][32] | scales | accumulators [2][QN][NT]
    constexpr int kRatioWords = 2 * QN * NT;
    constexpr int kRingWords = GRAD ? K * QN * NT : 0;
    constexpr int kScaleWords = GRAD ? (K / kNorm) * NT / W : 0;
    const uint32_t smem0 = smem_base_addr();
    RatioRows<NT> ratio, gam;
    ratio.base = smem0 + threadIdx.x * 16;
    const uint32_t seg_a = smem0 + (kRatioWords + warp * (K * QN * 32) + lane) * 16;
    const uint32_t scale_a = smem0 + (kRatioWords + kRingWords) * 16 + (warp * (K / kNorm * 32) + lane) * 4;
    gam.base = smem0 + (kRatioWords + kRingWords + kScaleWords + threadIdx.x) * 16;

    const int64_t s_eff = listed_chunks(a);
    const int64_t n_pairs = (SEG ? a.S : s_eff) * a.B;
    const int64_t n_groups = (!SEG && a.s_list) ? (n_pairs + kWarps * PW - 1) / (kWarps * PW) : a.n_groups;
    const int64_t L_max = SEG ? a.seg_len : a.L;
    const int64_t warp_slot = int64_t(blockIdx.x) * kWarps + warp;
    const F *params6 = static_cast<const F *>(a.params6);
    const F *pi_g = static_cast<const F *>(a.pi);
    V *ck = GRAD ? reinterpret_cast<V *>(static_cast<char *>(a.ckpt) + warp_slot * ckpt_bytes_per_warp<F, MT, K>(L_max)) + lane : nullptr;
    constexpr int kFlushSegs = kFlushSites / K;
    double *const gacc_base = a.gacc + int64_t(blockIdx.x) * NT + threadIdx.x;
    const int64_t gacc_stride = int64_t(gridDim.x) * NT;

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t pseg = SEG ? a.seg_first + grp / a.seg_ctas : 0;
        const int64_t L = SEG ? min(a.seg_len, a.L - pseg * a.seg_len) : a.L;
        const int64_t n_seg = (L + K - 1) / K;
        const int64_t pair_raw = ((SEG ? grp % a.seg_ctas : grp) * kWarps + warp) * PW + lane;
        const bool writer = pair_raw < n_pairs;
        const int64_t pair_idx = writer ? pair_raw : n_pairs - 1;  // idle lanes shadow the last pair
        const int64_t slot = pair_idx / a.B;
        const int64_t pb = pair_idx - slot * a.B;
        const int64_t ps = (!SEG && a.s_list) ? int64_t(a.s_list[slot]) : slot;
        const int64_t chunk_pair = pb * a.S + ps;
        int64_t pair = chunk_pair;
        if constexpr (SEG) pair = chunk_pair * (a.seg_local ? a.seg_local : a.seg_count) + (pseg - a.seg_first);

        // ---- coefficients of this pair
        const F *par = params6 + pb * a.pstride_b + ps * a.pstride_s;
        SCoef c;
        bool bad_domain = false;
        {
            float r1[MT], rm[MT];
#pragma unroll
            for (int k = 0; k < MT; ++k) {
                const F b = par[0 * M + k], d = par[1 * M + k], u = par[2 * M + k], v = par[3 * M + k];
                const F e0 = par[4 * M + k], e1 = par[5 * M + k];
                const F s = k == 0 ? F(1) : v;
                bad_domain |= !(s > F(0)) || !(e0 > F(0));
                c.beta[k] = b / s;
                c.sigma[k] = e0 * s;
                c.delta[k] = d * e0;
                c.pic[k] = u * e0 * s;
                rm[k] = F(1) / e0;
                r1[k] = e1 * rm[k];
            }
            ratio.store(0, r1);
            ratio.store(1, rm);
        }
        int64_t row = a.inds[ps];
        const bool bad_row = row < 0 || row >= a.n_rows;
        if (bad_row) {
            atomicOr(a.err_flag, 1);
            row = 0;
        }
        if (bad_domain) atomicOr(a.err_flag, 4);
        const int8_t *obs = a.data + row * a.pitch + (SEG ? pseg * a.seg_len : 0);
        const F *x_in = SEG ? static_cast<const F *>(a.bnd_alpha) + (chunk_pair * (a.seg_count + 1) + pseg) * M
                            : pi_g + pb * a.pistride_b + ps * a.pistride_s;

        // ------------------------------------------------------------------ pass 1: forward
        // y = x / sigma: the vector entering the first site in the units of this kernel
        F y[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) y[k] = x_in[k] / c.sigma[k];
        double ll = 0.0;
        ObsWords<K> ow_next;
        ow_next.load(obs, 0);
        for (int64_t seg = 0; seg < n_seg; ++seg) {
            if (GRAD && seg > 0) {
#pragma unroll
                for (int q = 0; q < QN; ++q) ck[(seg * QN + q) * 32] = pack(&y[q * W]);
            }
            const ObsWords<K> ow = ow_next;
            if (seg + 1 < n_seg) ow_next.load(obs, (seg + 1) * K);
            const int len = int(min(int64_t(K), L - seg * K));
            F acc = F(0);
            for (int kb = 0; kb < len; kb += kNorm) {
                const uint64_t blk = ow.block(kb);
                F tot = F(1);
                // ONE decision per block of kNorm sites: a full block of ob = 0 sites (3 blocks in 4; the lanes of a warp
                // share the chunk, so the decision is uniform) is one straight-line basic block of 4 x 64 FMAs in which
                // the scheduler overlaps neighbouring sites; blocks with a het / missing site and the ragged tail of a
                // chunk go through a compact per-site loop
                if (kb + kNorm <= len && uint32_t(blk) == 0u) {
#pragma unroll
                    for (int j = 0; j < kNorm; ++j) tot = sform_step(y, c);
                } else {
#pragma unroll 1
                    for (int j = 0; j < kNorm && kb + j < len; ++j) tot = sform_site<NT>(y, c, ratio, ObsWords<K>::byte_of(blk, j));
                }
                // lazy rescaling with the mass that entered the block's last transition (one site stale: any
                // positive factor does, as long as the same one is booked in ll and replayed in pass 2)
                const F inv = fast_rcp<F>(tot);
#pragma unroll
                for (int k = 0; k < MT; ++k) y[k] *= inv;
                acc += log2_of<F>(tot);
            }
            ll += double(acc);
        }
        // the mass after the last site
        F total = F(0);
#pragma unroll
        for (int k = 0; k < MT; ++k) total = fmaf(c.sigma[k], y[k], total);
        ll = (ll + double(log2_of<F>(total))) * 0.69314718055994530942;
        if (!(ll == ll) || ll > 1e300 || ll < -1e300) atomicOr(a.err_flag, 2);
        if (bad_row || bad_domain) ll = __longlong_as_double(0x7ff8000000000000LL);
        if constexpr (!SEG) {
            if (writer) a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
            if (writer && a.alpha_out != nullptr) {
                F *ao = static_cast<F *>(a.alpha_out) + pair * M;
#pragma unroll
                for (int k = 0; k < MT; ++k) ao[k] = c.sigma[k] * y[k] / total;
            }
        }

        if constexpr (GRAD) {
            // -------------------------------------------------------------- pass 2: adjoint
            // om = adjoint of y, kept at om . y == 1.  Behind the last site: sigma .* (adjoint of x), where the adjoint
            // of x is all ones (end of a chunk) or the boundary vector of the segment.
            F om[MT];
            {
                F dot = F(0);
#pragma unroll
                for (int k = 0; k < MT; ++k) {
                    F bx = F(1);
                    if constexpr (SEG) bx = static_cast<const F *>(a.bnd_beta)[(chunk_pair * (a.seg_count + 1) + pseg + 1) * M + k];
                    om[k] = c.sigma[k] * bx;
                    dot = fmaf(om[k], y[k], dot);
                }
                dot = F(1) / dot;
#pragma unroll
                for (int k = 0; k < MT; ++k) om[k] *= dot;
            }
            F Ab[MT], Ad[MT], Au[MT], Av[MT];
#pragma unroll
            for (int k = 0; k < MT; ++k) Ab[k] = Ad[k] = Au[k] = Av[k] = F(0);
            {
                F z[MT];
#pragma unroll
                for (int k = 0; k < MT; ++k) z[k] = F(0);
                gam.store(0, z);
                gam.store(1, z);
            }
#pragma unroll 1
            for (int i = 0; i < 6 * MT; ++i) gacc_base[int64_t(i) * gacc_stride] = 0.0;
            ObsWords<K> ow_ahead;
            ow_ahead.load(obs, (n_seg - 1) * K);
            for (int64_t seg = n_seg - 1; seg >= 0; --seg) {
                const ObsWords<K> ow = ow_ahead;
                if (seg > 0) {
                    ow_ahead.load(obs, (seg - 1) * K);
                    if (seg > 1) prefetch_l2(&ck[(seg - 1) * QN * 32]);
                }
                const int len = int(min(int64_t(K), L - seg * K));
                // re-run the forward steps of this segment, keeping the INPUT vector of every site
                F ys[MT];
                if (seg == 0) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) ys[k] = x_in[k] / c.sigma[k];
                } else {
#pragma unroll
                    for (int q = 0; q < QN; ++q) unpack<F>(ck[(seg * QN + q) * 32], &ys[q * W]);
                }
                for (int kb = 0; kb < len; kb += kNorm) {
                    const uint64_t blk = ow.block(kb);
                    F tot = F(1);
                    if (kb + kNorm <= len && uint32_t(blk) == 0u) {
#pragma unroll
                        for (int j = 0; j < kNorm; ++j) {
#pragma unroll
                            for (int q = 0; q < QN; ++q) sts_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &ys[q * W]);
                            tot = sform_step(ys, c);
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < kNorm && kb + j < len; ++j) {
#pragma unroll
                            for (int q = 0; q < QN; ++q) sts_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &ys[q * W]);
                            tot = sform_site<NT>(ys, c, ratio, ObsWords<K>::byte_of(blk, j));
                        }
                    }
                    const F inv = fast_rcp<F>(tot);
                    sts_scalar(scale_a + (kb / kNorm) * 32 * 4, inv);
#pragma unroll
                    for (int k = 0; k < MT; ++k) ys[k] *= inv;
                }
                // ys is the vector behind the segment: re-impose om . ys == 1 (round-off drift of the adjoint scaling)
                {
                    F dot = F(0);
#pragma unroll
                    for (int k = 0; k < MT; ++k) dot = fmaf(ys[k], om[k], dot);
                    dot = fast_rcp<F>(dot);
#pragma unroll
                    for (int k = 0; k < MT; ++k) om[k] *= dot;
                    // posterior of the segment's LAST site (= the vector it handed on .* its adjoint) while that vector is
                    // at hand; the other sites find theirs in the ring (the input of the next site)
                    const int ob_last = ow.at(len - 1);
                    if (ob_last != 0) {
                        F g[MT];
                        gam.load(ob_last < 0 ? 1 : 0, g);
#pragma unroll
                        for (int k = 0; k < MT; ++k) g[k] = fmaf(ys[k], om[k], g[k]);
                        gam.store(ob_last < 0 ? 1 : 0, g);
                    }
                }
                for (int kb = ((len - 1) / kNorm) * kNorm; kb >= 0; kb -= kNorm) {
                    const uint64_t blk = ow.block(kb);
                    const F scale = lds_scalar_f(scale_a + (kb / kNorm) * 32 * 4, F(0));
                    const int last = min(kNorm, len - kb) - 1;  // the block's last site: its output was rescaled in pass 1
                    auto adjoint_site = [&](int j, bool all_zero) {
                        F yin[MT];
#pragma unroll
                        for (int q = 0; q < QN; ++q) lds_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &yin[q * W]);
                        const int ob = all_zero ? 0 : ObsWords<K>::byte_of(blk, j);
                        if (ob != 0 && kb + j + 1 < len) {
                            // posterior of this site = (the vector it handed on) .* (its adjoint); the vector is the INPUT of
                            // the next site, in the ring (the segment's last site was booked above)
                            F yn[MT], g[MT];
#pragma unroll
                            for (int q = 0; q < QN; ++q) lds_word(seg_a + ((kb + j + 1) * QN + q) * 32 * 16, &yn[q * W]);
                            gam.load(ob < 0 ? 1 : 0, g);
#pragma unroll
                            for (int k = 0; k < MT; ++k) g[k] = fmaf(yn[k], om[k], g[k]);
                            gam.store(ob < 0 ? 1 : 0, g);
                        }
                        // the factor pass 1 applied behind the block's last site, carried on om keeps om . y == 1
                        if (j == last) {
#pragma unroll
                            for (int k = 0; k < MT; ++k) om[k] *= scale;
                        }
                        if (ob != 0) {
                            F r[MT];
                            ratio.load(ob < 0 ? 1 : 0, r);
#pragma unroll
                            for (int k = 0; k < MT; ++k) om[k] *= r[k];  // -> adjoint of the transition's output
                        }
                        // adjoint of the transition: chains over the incoming adjoint (Bp ascending, Q descending) and over the
                        // site's input (P ascending, S descending), walked from both ends in one loop so that their latencies overlap
                        F Bp = F(0), Q = F(0), P = F(0), S = F(0);
                        F tailq[MT], head[MT];
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const int k = i, jj = MT - 1 - i;
                            head[k] = fmaf(c.delta[k], om[k], c.sigma[k] * Bp);
                            Ad[k] = fmaf(yin[k], om[k], Ad[k]);
                            Av[k] = fmaf(P, om[k], Av[k]);
                            Bp = fmaf(c.beta[k], om[k], Bp);
                            P = fmaf(c.pic[k], yin[k], P);
                            tailq[jj] = Q;
                            Au[jj] = fmaf(yin[jj], Q, Au[jj]);
                            Ab[jj] = fmaf(S, om[jj], Ab[jj]);
                            Q += om[jj];
                            S = fmaf(c.sigma[jj], yin[jj], S);
                        }
#pragma unroll
                        for (int k = 0; k < MT; ++k) om[k] = fmaf(c.pic[k], tailq[k], head[k]);
                    };
                    if (kb + kNorm <= len && uint32_t(blk) == 0u) {
#pragma unroll
                        for (int j = kNorm - 1; j >= 0; --j) adjoint_site(j, true);  // (ob == 0 at compile time: no ratio code)
                    } else {
#pragma unroll 1
                        for (int j = last; j >= 0; --j) adjoint_site(j, false);
                    }
                }
                if ((seg & (kFlushSegs - 1)) == 0) {
                    flush_row<F, MT>(Ab, 0, gacc_base, gacc_stride);
                    flush_row<F, MT>(Ad, 1, gacc_base, gacc_stride);
                    flush_row<F, MT>(Au, 2, gacc_base, gacc_stride);
                    flush_row<F, MT>(Av, 3, gacc_base, gacc_stride);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        F g[MT];
                        gam.load(r, g);
                        flush_row<F, MT>(g, 4 + r, gacc_base, gacc_stride);
                        gam.store(r, g);  // (flush_row cleared it)
                    }
                }
            }
            if (writer) {
                F *out = static_cast<F *>(SEG ? a.seg_dlog : a.dlog) + pair * 7 * M;
#pragma unroll
                for (int k = 0; k < MT; ++k) {
                    const double gb = gacc_base[int64_t(0 * MT + k) * gacc_stride] * double(c.beta[k]);
                    const double gd = gacc_base[int64_t(1 * MT + k) * gacc_stride] * double(c.delta[k]);
                    const double gu = gacc_base[int64_t(2 * MT + k) * gacc_stride] * double(c.pic[k]);
                    const double gv = k == 0 ? 0.0 : gacc_base[int64_t(3 * MT + k) * gacc_stride];
                    const double g1 = gacc_base[int64_t(4 * MT + k) * gacc_stride];
                    const double gm = gacc_base[int64_t(5 * MT + k) * gacc_stride];
                    F val[7];
                    val[0] = F(gb);
                    val[1] = F(gd);
                    val[2] = F(gu);
                    val[3] = F(gv);
                    // every posterior is the sum of the three ways of arriving in the state
                    val[4] = F(gb + gd + gv - g1 - gm);
                    val[5] = F(g1);
                    val[6] = om[k] * (x_in[k] / c.sigma[k]);
#pragma unroll
                    for (int r = 0; r < 7; ++r) out[r * M + k] = (!SEG && a.out_mode) ? out[r * M + k] - val[r] : val[r];
                }
            }
        }
    }
}

}  // namespace phb
