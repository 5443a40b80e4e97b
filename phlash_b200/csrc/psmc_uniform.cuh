// Forward-only throughput kernel for large minibatches at M = 16 (float) whose parameter rows are shared by the
// chunks of a particle - how the reference calls the kernel, model.py:55.
//
// Layout: a single-warp CTA scores 32 chunks of ONE particle (psmc_loglik_kernel scores consecutive particles of
// one chunk).  With one particle per warp
//   * the emission rows sit in ONE 192-byte table per warp instead of 192 bytes per thread,
//   * rows b and v come from __constant__ memory through a warp-uniform slot index (psmc_loglik_kernel reads all
//     rows from global memory into per-lane registers),
//   * the kernel needs 149 registers and 192 bytes of shared memory per warp: 13 warps per SM.
// The recursion is psmc_loglik_kernel's forward pass (same site function, same rescaling): the log-likelihoods
// are bit-identical.  Measured on B200 (profiles/r02_probe_uniform_kernel.log, 124 particles x 576 chunks x
// 50 000 bins): 17.2 ms against 20.7 ms (+21 %).
//
// What was tried around it (round 2) and is NOT used:
//   * all six rows in the constant bank as FMA operands of the uniform datapath (FFMA R, R, UR, R - ptxas does
//     this by itself once the slot index is provably uniform): a site-like loop issues 0.79 instead of 0.72 FMA
//     per scheduler and clock that way (profiles/r02_microbench3_b200.json), and the forward-only build gained
//     another 5 % (16.3 ms).  But ptxas keeps at most ~32 of the 64 transition values resident in uniform
//     registers (with all 64 it re-loads them through LDCU at every site), whether it uses uniform or vector
//     registers for them at all depends on the register pressure of the build, and 384 bytes per particle limit a
//     launch to 128 particles, which serialises launches that do not fill the GPU on their own;
//   * the GRADIENT build of the same layout: the dispatch stalls halve (0.62 -> 0.29 per issue) but at 2 warps per
//     scheduler the warps then wait on their own dependency chains (short scoreboard 0.10 -> 0.51, fixed-latency
//     waits 0.21 -> 0.51): 77.2 ms against 71.6 ms;
//   * the segment transfer operators in this layout: no gain at S = 1 (3.21 against 3.20 ms) and a second,
//     nearly empty round of single-warp CTAs for the ELPD shape (119 against 88 ms).
// The constant bank holds kUniformSlots particles; larger batches are scored in several launches (the host copies
// each batch's rows device-to-device into the bank in between).
#pragma once

#include "psmc_kernels.cuh"

namespace phb {

constexpr int kUniformM = 16;
// Rows b and v, 128 B per particle, so that the reference's 500 particles fit in ONE launch (504 x 128 B = 63 KB
// of the 64 KB bank).
constexpr int kUniformSlots = 504;
struct UniformParams {
    float b[kUniformM], v[kUniformM];
};
__constant__ UniformParams c_uniform_params[kUniformSlots];

struct UniformArgs {
    KernelArgs k;        // data, inds, B, S, pi layout, ll, err_flag, out_mode, s_list / s_count
    int64_t first_b;     // particle of constant slot 0
    int64_t n_b;         // particles of this launch (<= kUniformSlots)
    int64_t n_tasks;     // n_b * ceil(chunks / 32) upper bound (with a sub-list the kernel recomputes it)
};

// staging of the constant bank: out[b] = rows b and v of particle b, contiguous
__global__ void pack_uniform_params_kernel(const float *__restrict__ params6, int64_t pstride_b, int64_t B, UniformParams *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B * 2 * kUniformM) return;
    const int64_t b = i / (2 * kUniformM), r = i % (2 * kUniformM);
    reinterpret_cast<float *>(out + b)[r] = params6[b * pstride_b + (r < kUniformM ? r : 2 * kUniformM + r)];  // row 0 (b), row 3 (v)
}

constexpr size_t uniform_smem_bytes() { return 3 * kUniformM * sizeof(float); }  // the warp's emission table [3][16]

template <int K> __global__ void __maxnreg__(168) psmc_uniform_forward_kernel(const UniformArgs ua) {
    using F = float;
    constexpr int MT = kUniformM, T = 1, W = 4, QN = MT / W;
    const KernelArgs &a = ua.k;
    const int lane = threadIdx.x;
    const uint32_t smem0 = smem_base_addr();
    EmisTable<F, MT, 1> et;  // column stride 1: every lane addresses the same table
    et.ones = smem0;
    et.base = smem0;
    PartnerCoef<F, MT, T, false> pc;

    const int64_t s_eff = listed_chunks(a);
    const int64_t wpp = (s_eff + 31) / 32;  // warp tasks per particle
    const int64_t n_tasks = ua.n_b * wpp;
    const float *pi_g = static_cast<const float *>(a.pi);
    const int64_t L = a.L;
    const int64_t n_seg = (L + K - 1) / K;

    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int slot = int(task / wpp);  // warp-uniform by construction (blockIdx, loop counter)
        const int64_t blk = task - int64_t(slot) * wpp;
        const UniformParams &up = c_uniform_params[slot];
        const int64_t pb = ua.first_b + slot;
        // rows b and v from the constant bank, d and u (and the emission rows below) from global memory
        const float *par = static_cast<const float *>(a.params6) + pb * a.pstride_b;
        Params<F, MT> p;
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            p.b[k] = up.b[k];
            p.v[k] = up.v[k];
            p.d[k] = par[1 * MT + k];
            p.u[k] = par[2 * MT + k];
        }
        const int64_t j_raw = blk * 32 + lane;
        const bool writer = j_raw < s_eff;
        const int64_t j = writer ? j_raw : s_eff - 1;  // idle lanes shadow the last chunk
        const int64_t ps = a.s_list ? int64_t(a.s_list[j]) : j;
        const int64_t pair = pb * a.S + ps;
        // the warp's emission table: rows emis0, emis1, ones
        __syncwarp();
        if (lane < 3 * QN) {
            const int r = lane / QN, q = lane % QN;
            float tmp[W];
#pragma unroll
            for (int i = 0; i < W; ++i) tmp[i] = r < 2 ? par[(4 + r) * MT + q * W + i] : 1.f;
            sts_word(smem0 + lane * 16, tmp);
        }
        __syncwarp();
        int64_t row = a.inds[ps];
        const bool bad_row = row < 0 || row >= a.n_rows;
        if (bad_row) {
            atomicOr(a.err_flag, 1);
            row = 0;
        }
        const int8_t *obs = a.data + row * a.pitch;
        const float *pi_p = pi_g + pb * a.pistride_b + ps * a.pistride_s;

        // ------------------------------------------------------------------ the forward recursion
        F x[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) x[k] = pi_p[k];
        double ll = 0.0;
        ObsWords<K> ow_next;
        ow_next.load(obs, 0);
        for (int64_t seg = 0; seg < n_seg; ++seg) {
            const ObsWords<K> ow = ow_next;
            if (seg + 1 < n_seg) ow_next.load(obs, (seg + 1) * K);
            const int len = int(min(int64_t(K), L - seg * K));
            F acc = F(0);
            for (int kb = 0; kb < len; kb += kNorm) {
                const uint64_t blkw = ow.block(kb);
                if (kb + kNorm <= len) {
#pragma unroll
                    for (int jj = 0; jj < kNorm; ++jj) forward_site<F, MT, T, false, 1>(x, p, pc, et, ObsWords<K>::byte_of(blkw, jj), 0);
                } else {
#pragma unroll 1
                    for (int jj = 0; kb + jj < len; ++jj) forward_site<F, MT, T, false, 1>(x, p, pc, et, ObsWords<K>::byte_of(blkw, jj), 0);
                }
                const F tot = pair_sum<F, MT, T>(x);
                const F inv = fast_rcp<F>(tot);
#pragma unroll
                for (int jj = 0; jj < MT; ++jj) x[jj] *= inv;
                acc += log2_of<F>(tot);
            }
            ll += double(acc);
        }
        ll *= 0.69314718055994530942;
        if (!(ll == ll) || ll > 1e300 || ll < -1e300) atomicOr(a.err_flag, 2);
        if (bad_row) ll = __longlong_as_double(0x7ff8000000000000LL);
        if (writer) a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;

    }
}

}  // namespace phb
