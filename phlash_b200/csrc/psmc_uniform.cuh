// Forward-only throughput kernel with the parameters in UNIFORM REGISTERS (M = 16, float, large minibatches whose
// six parameter rows are shared by the chunks of a particle - how the reference calls the kernel, model.py:55).
//
// Why: psmc_loglik_kernel keeps b, d, u, v (64 values) in vector registers, and almost every FMA of the site
// step reads three distinct vector registers.  That operand pattern is what bounds the kernel: a loop of the
// same shape issues 0.72 FMA per scheduler and clock whatever the occupancy, 0.79 when the coefficients come
// from the constant bank / uniform registers (profiles/r02_microbench3_b200.json).  Here a WARP scores 32
// chunks of ONE particle, the particle's rows live in __constant__ memory, and because the slot index is
// provably warp-uniform (it derives from blockIdx and the loop counter: a CTA is one warp) ptxas loads them
// once per task with LDCU into uniform registers and uses them as the third operand of the FMAs
// (FFMA R, R, UR, R).  64 vector registers and the per-thread emission table are gone; the emission rows sit in
// one 192-byte table per warp.
//
// The algorithm is exactly psmc_loglik_kernel's forward pass (same site function, same rescaling: the results are
// bit-identical); see psmc_kernels.cuh.  __constant__ memory holds kUniformSlots particles, so a minibatch is
// scored in batches of particles (the host copies each batch's rows device-to-device into the bank between
// launches).  Measured on B200 (profiles/r02_probe_uniform_kernel.log, 124 particles x 576 chunks x 50 000
// bins): forward only 16.3 ms against 20.8 ms (+28 %).  The GRADIENT build of the same idea was measured too and
// is NOT used: with the adjoint's working set the kernel stays at 2 warps per scheduler, the dispatch stalls
// halve (0.62 -> 0.29 per issue) but the warps then wait on their own dependency chains instead (short
// scoreboard 0.10 -> 0.51, fixed-latency waits 0.21 -> 0.51): 77.2 ms against 71.6 ms.
#pragma once

#include "psmc_kernels.cuh"

namespace phb {

constexpr int kUniformM = 16;
constexpr int kUniformSlots = 128;  // 128 x 384 B = 48 KB of the 64 KB constant bank
struct UniformParams {
    float b[kUniformM], d[kUniformM], u[kUniformM], v[kUniformM], e0[kUniformM], e1[kUniformM];
};
__constant__ UniformParams c_uniform_params[kUniformSlots];

struct UniformArgs {
    KernelArgs k;        // data, inds, B, S, pi layout, ll, err_flag, out_mode, s_list / s_count
    int64_t first_b;     // particle of constant slot 0
    int64_t n_b;         // particles of this launch (<= kUniformSlots)
    int64_t n_tasks;     // n_b * ceil(chunks / 32) upper bound (with a sub-list the kernel recomputes it)
};

// staging of the constant bank: out[b] = rows b, d, u, v, emis0, emis1 of particle b, contiguous
__global__ void pack_uniform_params_kernel(const float *__restrict__ params6, int64_t pstride_b, int64_t B, UniformParams *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B * 6 * kUniformM) return;
    const int64_t b = i / (6 * kUniformM), r = i % (6 * kUniformM);
    reinterpret_cast<float *>(out + b)[r] = params6[b * pstride_b + r];
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
        const int slot = int(task / wpp);                 // warp-uniform by construction (blockIdx, loop counter)
        const int64_t blk = task - int64_t(slot) * wpp;
        const UniformParams &up = c_uniform_params[slot];
        // rows b and v stay in the constant bank / uniform registers (FMA operands); d and u are moved into
        // vector registers (opaque moves: ptxas would otherwise re-load all 64 values through LDCU at every
        // site instead of keeping any of them resident - there are not enough uniform registers for 64)
        Params<F, MT> p;
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            p.b[k] = up.b[k];
            p.v[k] = up.v[k];
            asm volatile("mov.f32 %0, %1;" : "=f"(p.d[k]) : "f"(up.d[k]));
            asm volatile("mov.f32 %0, %1;" : "=f"(p.u[k]) : "f"(up.u[k]));
        }
        const int64_t pb = ua.first_b + slot;
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
            for (int i = 0; i < W; ++i) tmp[i] = r == 0 ? up.e0[q * W + i] : (r == 1 ? up.e1[q * W + i] : 1.f);
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
