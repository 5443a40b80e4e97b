// PSMC coalescent-HMM log-likelihood + gradient for sm_100a.
//
// What is computed (reference semantics): the scaled forward recursion of
// src/phlash/hmm.py:68-82 with the O(M) structured transition of hmm.py:52-65 / gpu.py:504-522,
// and the gradient contract of the reference's loglik_grad kernel (gpu.py:575-692, host roll
// :303-313).  How it is computed is new:
//
//  * T lanes cooperate on one (chunk, particle) pair; each lane owns MT = M / T consecutive hidden
//    states and keeps its slice of the forward vector AND of the six parameter rows in registers
//    for the whole chunk.  The prefix / suffix sums of the structured transition are serial FMA
//    chains inside a lane; lane totals are combined across the T lanes with a few shuffles
//    ("totals first": reduce, shuffle, then run the chains starting from the offsets).
//    (v1 of this kernel streamed the parameters from shared memory with broadcast LDS.128; on
//    B200 that is bound by the ~0.4 LDS.128/clk/SM register-fill rate - see profiles/.)
//  * the gradient is the adjoint (backward) recursion, O(M) per site instead of the reference's
//    O(7 M^2) forward-mode sensitivities.  The forward vectors it needs are not kept for the whole
//    chunk: pass 1 stores a checkpoint every K sites to HBM (4*M/K bytes per site and pair);
//    pass 2 walks the segments backwards, re-runs the K forward steps of a segment into a
//    shared-memory ring private to the lane (conflict-free 128-bit accesses), then runs the K
//    adjoint steps, accumulating d ll / d (b, d, u, v, emis0, emis1) in registers.
//  * persistent grid: every resident CTA loops over groups of 128 / T pairs; the warps of a CTA
//    never synchronise (no block-level barrier anywhere).  Checkpoint scratch is indexed by warp slot, so its
//    size depends on the GPU, not on the problem.
//
// Per site and pair: ~9 M FMA-pipe instructions forward (twice with the recompute) and ~20 M for
// the adjoint step, 1 B of observations, 2 * 4 * M / K B of checkpoint traffic,
// 2 * 4 * (M + 1) B of shared-memory traffic.
//
// Kernels in this file:
//   psmc_loglik_kernel            the throughput kernel described above (loglik, or loglik + gradient);
//                                 SEG = true: the same passes over the segments of a chunk (optionally with the
//                                 checkpoints left by the forward sweep, and with the warm-up term of the
//                                 whole-term entries as one more segment)
//   psmc_loglik_storeall_kernel   gradient of small minibatches: every forward vector kept in HBM
//   transfer_rows_kernel,         parallel in time for FEW pairs: segment transfer operators, chained in
//   chain_transfer_kernel,        float64 to the log-likelihood (forward only) or to the forward / adjoint
//   chain_boundaries_kernel,      vectors at the segment boundaries (gradient); chain_product_kernel /
//   chain_product_kernel, ..._sharded_kernel   the same with the segments sharded over processes
//   boundary_sweep_kernel         the same boundary vectors from two sequential sweeps side by side: the latency
//                                 path (one warp per scheduler), with its own lane layouts per direction,
//                                 low-latency site functions and hand-pipelined block loops
//   sum_segments_kernel           adds the partial gradients of the segments (and subtracts the warm-up term)
//   flag_long_runs_kernel,        precision escalation: rows with long runs of identical observations
//   split_minibatch_kernel        are scored in double
//   chunk_het_kernel, validate_params_kernel   data chunking / input validation
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace phb {

// Threads per CTA are a template parameter (NT) of the kernel: the warps of a CTA never
// synchronise, so NT only sets the register budget (65536 / NT per thread at one CTA per SM).

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
    void *ckpt;          // checkpoint scratch, see ckpt_bytes_per_warp()
    double *gacc;        // fp64 gradient accumulators: [6 * MT][gridDim.x * NT]
    void *xall;          // store-all kernel: every forward input vector, [warp slot][site][MT / W][32] words
    void *sall;          // store-all kernel: block scale factors, [warp slot][ceil(L / kNorm)][32]
    int64_t n_groups;    // ceil(B*S / pairs-per-CTA)
    int *err_flag;       // bit 0: index out of range, bit 1: non-finite result
    int out_mode;        // 0: ll / dlog are written;  1: the results are SUBTRACTED from what is there
                         // (second launch of the fused warm-up evaluation, see phb_loglik_warmup_*)
    // Optional restriction to a SUBSET of the minibatch (precision escalation, see
    // flag_long_runs_kernel): the launch scores pairs (b, s_list[j]) for j < *s_count only; both live
    // in device memory, so the split needs no host round trip.  nullptr = all S chunks.
    const int32_t *s_list;
    const int32_t *s_count;
    // Parallel-in-time paths score the WHOLE minibatch (their grids are sized on the host) but leave the
    // outputs of pairs whose row is marked in skip_flag ([N], see flag_long_runs_kernel) untouched: those
    // pairs belong to the double-arithmetic launch that follows.  nullptr = write everything.
    const uint8_t *skip_flag;
    // Segment mode of the gradient kernel (parallel-in-time gradient, see chain_boundaries_kernel): the
    // launch scores the G segments of every chunk as independent short "pairs" (L = sites per chunk),
    // started from bnd_alpha and closed with bnd_beta; partial gradients go to seg_dlog.
    int64_t seg_count;      // G, segments per chunk
    int64_t seg_first;      // this launch scores segments [seg_first, seg_first + seg_local) only (time-axis
    int64_t seg_local;      // sharding over processes; seg_local == 0 means all G)
    int64_t seg_len;        // sites per segment (multiple of 16); the last one has L - (G - 1) seg_len
    int64_t seg_ctas;       // groups per segment: group g scores segment g / seg_ctas, so that all warps of a CTA
                            // share the segment length (the loop bounds stay uniform for the shuffles)
    int64_t sweep_fwd_ctas; // boundary_sweep_kernel: CTAs [0, sweep_fwd_ctas) sweep forwards, the rest backwards
    const void *bnd_alpha;  // [B * S][G + 1][M] FLOAT: forward vector entering segment g (sum 1)
    const void *bnd_beta;   // [B * S][G + 1][M] FLOAT: adjoint vector behind segment g - 1 (any scale)
    void *seg_dlog;         // [B * S][segments of this launch][7][M] FLOAT
    // Two-sweep path: the forward sweep also leaves the forward vector entering every K-site group (K = the
    // checkpoint spacing of the gradient kernel that runs the segment passes), so that those start with the adjoint
    // pass right away instead of re-running the forward recursion for their checkpoints.  nullptr = they make their own.
    void *ext_ck;           // [B * S][ext_ck_count][M] FLOAT, record m = vector after site K m - 1 (record 0 unused)
    int64_t ext_ck_count;   // ceil(L / K) + 1
    int ext_ck_blocks;      // K / kNorm
    // Fused warm-up term (phb_loglik_warmup_*: LL of the whole row minus LL of its first warm_len sites): the segment
    // passes score the first warm_len sites as ONE MORE segment per chunk - started from pi itself, closed with a
    // vector of ones, its partial gradient in the slot behind the real segments - and sum_segments_kernel subtracts it.
    // The 500 dependent sites of the warm-up then cost no launch of their own.  0 = no warm-up term.
    int64_t warm_len;
    double *warm_ll;        // [B * S] log-likelihood of the first warm_len sites
};

// Pair enumeration of the store-all kernel: b major, position in the (sub-)list minor.  (psmc_loglik_kernel
// enumerates chunk major, see there.)
struct PairIndex {
    int64_t b, s, out;  // particle, position in the full minibatch, index into ll / dlog ([B, S])
};
__device__ __forceinline__ int64_t listed_chunks(const KernelArgs &a) { return a.s_list ? int64_t(*a.s_count) : a.S; }
// does another launch own the outputs of the pairs on chunk position ps?
__device__ __forceinline__ bool outputs_skipped(const KernelArgs &a, int64_t ps) {
    if (a.skip_flag == nullptr) return false;
    const int64_t row = a.inds[ps];
    return row >= 0 && row < a.n_rows && a.skip_flag[row] != 0;
}
__device__ __forceinline__ PairIndex pair_index(const KernelArgs &a, int64_t pair, int64_t s_eff) {
    PairIndex r;
    r.b = pair / s_eff;
    const int64_t j = pair % s_eff;
    r.s = a.s_list ? int64_t(a.s_list[j]) : j;
    r.out = r.b * a.S + r.s;
    return r;
}

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
// Butterfly: at level l every lane holds the sum of its aligned block of 2^l lanes and fetches the
// sibling block's sum with one shfl.xor; a sibling with a smaller lane index contributes to the
// "before" sum, one with a larger index to the "after" sum.  log2(T) dependent shuffles.
// sum of `mine` over the lanes with a smaller sub
template <typename F, int T> __device__ __forceinline__ F lanes_before(F mine, int sub) {
    F block = mine, out = F(0);
#pragma unroll
    for (int o = 1; o < T; o <<= 1) {
        const F sibling = __shfl_xor_sync(0xffffffffu, block, o, T);
        if (sub & o) out += sibling;
        block += sibling;
    }
    return out;
}
// sum of `mine` over the lanes with a larger sub
template <typename F, int T> __device__ __forceinline__ F lanes_after(F mine, int sub) {
    F block = mine, out = F(0);
#pragma unroll
    for (int o = 1; o < T; o <<= 1) {
        const F sibling = __shfl_xor_sync(0xffffffffu, block, o, T);
        if (!(sub & o)) out += sibling;
        block += sibling;
    }
    return out;
}
template <typename F, int T> __device__ __forceinline__ F lanes_total(F mine) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o, T);
    return mine;
}

template <typename F> __device__ __forceinline__ F fast_rcp(F x);
// approximate reciprocal (MUFU.RCP, <= 1 ulp): the rescaling factor only has to be applied
// consistently (the forward pass records it for the adjoint pass), not to be exactly 1 / sum
template <> __device__ __forceinline__ float fast_rcp<float>(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <> __device__ __forceinline__ double fast_rcp<double>(double x) { return 1.0 / x; }
template <typename F> __device__ __forceinline__ F log2_of(F x);
// MUFU.LG2: absolute error <= 2^-22 per call, one call per rescaling block; over a 50 000-bin chunk
// that is < 1e-7 of the log-likelihood even if every error had the same sign
template <> __device__ __forceinline__ float log2_of<float>(float x) { return __log2f(x); }
template <> __device__ __forceinline__ double log2_of<double>(double x) { return log2(x); }

// Branch-free selection by bit arithmetic (LOP3 / SEL on the integer pipe).  The obvious ternary
// on per-lane data is compiled into a divergent branch, which serialises the warp: the lanes of a
// warp hold different chunks.
__device__ __forceinline__ float bit_select(uint32_t mask, float a, float b) {
    return __uint_as_float((__float_as_uint(a) & mask) | (__float_as_uint(b) & ~mask));
}
__device__ __forceinline__ double bit_select(uint32_t mask, double a, double b) {
    const uint64_t m = (uint64_t(mask) << 32) | mask;
    return __longlong_as_double((__double_as_longlong(a) & m) | (__double_as_longlong(b) & ~m));
}

// This lane's slice of the four transition rows, register resident.  The two emission rows live
// in a per-lane shared-memory table (see EmisTable): selecting between registers costs two
// selects per state and site, the table costs one 128-bit shared load per four states.
template <typename F, int MT> struct Params {
    F b[MT], d[MT], u[MT], v[MT];
    template <typename IO> __device__ __forceinline__ void load(const IO *__restrict__ src, int M) {
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            b[k] = F(src[0 * M + k]);
            d[k] = F(src[1 * M + k]);
            u[k] = F(src[2 * M + k]);
            v[k] = F(src[3 * M + k]);
        }
    }
};

// All dynamic shared memory of the kernel.  It is addressed with INTEGER offsets (in 128-bit words)
// rather than through pointers: a pointer into shared memory stored in a struct is a generic
// 64-bit address, and the compiler then re-derives the shared-window offset (S2R SR_CgaCtaId, LEA,
// 64-bit IADD3) inside the hot loops - about 8 % of the stall samples in the profile of the first
// thread-per-pair kernel.
extern __shared__ __align__(16) unsigned char phb_smem[];

// Explicit shared-state-space accesses with 32-bit addresses (word index * 16 bytes from the start
// of the dynamic shared memory).  volatile + "memory": a ring slot is written by the recompute pass
// and read back by the adjoint pass, so these must not be reordered or merged by the compiler.
__device__ __forceinline__ uint32_t smem_base_addr() { return static_cast<uint32_t>(__cvta_generic_to_shared(phb_smem)); }
__device__ __forceinline__ void lds_word(uint32_t addr, float *o) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds_word(uint32_t addr, double *o) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_word(uint32_t addr, const float *v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void sts_word(uint32_t addr, const double *v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ float lds_scalar_f(uint32_t addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_scalar_f(uint32_t addr, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_scalar(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_scalar(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// Per-lane emission table in shared memory: rows emis0, emis1 of this lane's MT states, laid out
// [row][MT / W][NT] in 128-bit words so that a warp's access is conflict free.  `base` already
// points at this thread's column.
template <int MT> __host__ __device__ constexpr bool emis_acc_in_smem() { return MT >= 16; }

template <typename F, int MT, int NT> struct EmisTable {
    using V = typename Vec<F>::type;
    static constexpr int W = Vec<F>::W;
    static constexpr int QN = MT / W;
    // MT >= 16 layouts keep a third per-thread row of ones (the emission of a missing observation):
    // the row is then chosen by ONE address computation and the loads use immediate offsets.  The
    // other layouts have no shared memory to spare and redirect each load to one shared word of ones.
    static constexpr bool kOnesRow = emis_acc_in_smem<MT>();
    static constexpr int kRows = kOnesRow ? 3 : 2;
    uint32_t base;  // shared address of this thread's column
    uint32_t ones;  // shared address of the one 128-bit word of 1.0 shared by the CTA (!kOnesRow)
    template <typename IO> __device__ __forceinline__ void fill(const IO *__restrict__ src, int M) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
            for (int q = 0; q < QN; ++q) {
                F tmp[W];
#pragma unroll
                for (int i = 0; i < W; ++i) tmp[i] = r < 2 ? F(src[(4 + r) * M + q * W + i]) : F(1);
                sts_word(base + (r * QN + q) * NT * 16, tmp);
            }
        }
    }
    // emission probabilities of observation `ob` for this lane's states (missing -> 1)
    __device__ __forceinline__ void get(int ob, F (&e)[MT]) const {
        if constexpr (kOnesRow) {
            const uint32_t row_base = base + (ob < 0 ? 2 : ob) * (QN * NT * 16);
#pragma unroll
            for (int q = 0; q < QN; ++q) lds_word(row_base + q * NT * 16, &e[q * W]);
        } else {
            const uint32_t row_base = base + (ob == 1 ? QN * NT * 16 : 0);
#pragma unroll
            for (int q = 0; q < QN; ++q) lds_word(ob < 0 ? ones : row_base + q * NT * 16, &e[q * W]);
        }
    }
};

// T == 2 only.  The lower lane of a pair needs the upper lane's totals of (x, v.*w) as suffix
// offsets; the upper lane needs the lower lane's totals of (u.*x, b.*w) as prefix offsets.  Each
// lane therefore only has to PRODUCE the two totals its partner consumes: one FMA chain per vector
// with lane-dependent coefficients, one shuffle each (instead of two chains and two shuffles).
template <typename F, int MT, int T, bool GRAD> struct PartnerCoef {
    __device__ __forceinline__ void init(const Params<F, MT> &, int) {}
};
template <typename F, int MT> struct PartnerCoef<F, MT, 2, false> {
    F cx[MT];  // sub 0: u (partner wants sum u.*x);  sub 1: 1 (partner wants sum x)
    __device__ __forceinline__ void init(const Params<F, MT> &p, int sub) {
#pragma unroll
        for (int k = 0; k < MT; ++k) cx[k] = sub == 0 ? p.u[k] : F(1);
    }
};
template <typename F, int MT> struct PartnerCoef<F, MT, 2, true> {
    F cx[MT];  // as above
    F cw[MT];  // sub 0: b (partner wants sum b.*w);  sub 1: v (partner wants sum v.*w)
    __device__ __forceinline__ void init(const Params<F, MT> &p, int sub) {
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            cx[k] = sub == 0 ? p.u[k] : F(1);
            cw[k] = sub == 0 ? p.b[k] : p.v[k];
        }
    }
};

// One forward step for this lane's MT states: x <- (x A) .* emis(ob).  No rescaling here: the
// callers renormalise every kNorm sites (lazy scaling), which is exact for the log-likelihood and
// the gradient as long as the bookkeeping uses the same factors.
template <typename F, int MT, int T, bool GRAD, int NT>
__device__ __forceinline__ void forward_site(F (&x)[MT], const Params<F, MT> &p, const PartnerCoef<F, MT, T, GRAD> &pc,
                                             const EmisTable<F, MT, NT> &et, int ob, int sub) {
    F e[MT];
    et.get(ob, e);
    F pre_run = F(0), suf_run = F(0);
    if constexpr (T == 2) {
        F t[2] = {F(0), F(0)};
#pragma unroll
        for (int k = 0; k < MT; ++k) t[k & 1] = fma(pc.cx[k], x[k], t[k & 1]);
        const F other = __shfl_xor_sync(0xffffffffu, t[0] + t[1], 1, 2);
        pre_run = sub == 0 ? F(0) : other;
        suf_run = sub == 0 ? other : F(0);
    } else if constexpr (T > 2) {
        F tu[2] = {F(0), F(0)}, tx[2] = {F(0), F(0)};
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            tu[k & 1] = fma(p.u[k], x[k], tu[k & 1]);
            tx[k & 1] += x[k];
        }
        pre_run = lanes_before<F, T>(tu[0] + tu[1], sub);
        suf_run = lanes_after<F, T>(tx[0] + tx[1], sub);
    }
    // The ascending chain (prefix of u.*x) and the descending chain (suffix of x) are independent:
    // walk them from both ends in the same loop so that their latencies overlap (each is MT
    // dependent operations long), then combine.
    F part[MT], suf[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int k = i, j = MT - 1 - i;
        part[k] = fma(p.v[k], pre_run, p.d[k] * x[k]);
        pre_run = fma(p.u[k], x[k], pre_run);
        suf[j] = suf_run;
        suf_run += x[j];
    }
#pragma unroll
    for (int k = 0; k < MT; ++k) x[k] = fma(p.b[k], suf[k], part[k]) * e[k];
}

// sum of x over the whole pair
template <typename F, int MT, int T> __device__ __forceinline__ F pair_sum(const F (&x)[MT]) {
    F part[4] = {F(0), F(0), F(0), F(0)};
#pragma unroll
    for (int k = 0; k < MT; ++k) part[k & 3] += x[k];
    return lanes_total<F, T>((part[0] + part[1]) + (part[2] + part[3]));
}

// Gradient accumulators of one lane (d ll / d theta, not yet multiplied by theta, for b, d, u, v;
// already d / d log for the two emission rows).  They are fp32 registers (for F = float) that
// only ever hold the sum over a window of kFlushSites sites; every window is added into a
// per-thread fp64 slot in global memory (L2 resident).  Without this, increments smaller than
// half an ulp of a 50 000-site running sum are lost and the gradient is off by up to 5e-4.
constexpr int kFlushSites = 1024;

// ESM = true: the two emission rows are accumulated in shared memory instead (thread-per-pair
// variant: 3 instructions per state become one 128-bit load + 4 FMA + one 128-bit store per 4
// states, and 2*MT registers are freed).
template <typename F, int MT, bool ESM> struct Grad;

template <typename F, int MT> struct Grad<F, MT, false> {
    F b[MT], d[MT], u[MT], v[MT], e0[MT], e1[MT];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < MT; ++k) b[k] = d[k] = u[k] = v[k] = e0[k] = e1[k] = F(0);
    }
};
// (no accumulator for d: every posterior sums the three ways of arriving in a state, so
//     d ll / d log d_k = sum_t posterior_t(k) - d ll / d log b_k - d ll / d log v_k,
// and the posterior sums are accumulated anyway for the emission rows - 16 FMAs per site and 16 registers less)
template <typename F, int MT> struct Grad<F, MT, true> {
    F b[MT], u[MT], v[MT];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < MT; ++k) b[k] = u[k] = v[k] = F(0);
    }
};

// acc[i * stride] += window sum, then clear: fire-and-forget reduction, no loaded value kept live
template <typename F, int MT>
__device__ __forceinline__ void flush_row(F (&row)[MT], int r, double *acc, int64_t stride) {
#pragma unroll
    for (int k = 0; k < MT; ++k) {
        atomicAdd(acc + int64_t(r * MT + k) * stride, double(row[k]));
        row[k] = F(0);
    }
}

// Shared-memory accumulators for the emission rows: [3][MT / W][NT] 128-bit words per CTA (row 2
// swallows the contributions of missing observations so that no branch is needed).
template <typename F, int MT, int NT> struct EmisAcc {
    using V = typename Vec<F>::type;
    static constexpr int W = Vec<F>::W;
    static constexpr int QN = MT / W;
    uint32_t base;  // shared address of this thread's column
    __device__ __forceinline__ void clear() {
        F z[W];
#pragma unroll
        for (int i = 0; i < W; ++i) z[i] = F(0);
#pragma unroll
        for (int i = 0; i < 3 * QN; ++i) sts_word(base + i * NT * 16, z);
    }
    // row(ob) += x .* beta
    __device__ __forceinline__ void add(int ob, const F (&beta)[MT], const F (&x)[MT]) {
        const int row = ob < 0 ? 2 : ob;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
            F a[W];
            lds_word(base + (row * QN + q) * NT * 16, a);
#pragma unroll
            for (int i = 0; i < W; ++i) a[i] = fma(x[q * W + i], beta[q * W + i], a[i]);
            sts_word(base + (row * QN + q) * NT * 16, a);
        }
    }
    // rows emis0 / emis1 go to slots 4 / 5; the posterior mass of MISSING observations (row 2) goes to
    // slot 1, which has no accumulator of its own (see Grad<F, MT, true>)
    __device__ __forceinline__ void flush(double *acc, int64_t stride) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int slot = r < 2 ? 4 + r : 1;
#pragma unroll
            for (int q = 0; q < QN; ++q) {
                F a[W], z[W];
                lds_word(base + (r * QN + q) * NT * 16, a);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    atomicAdd(acc + int64_t(slot * MT + q * W + i) * stride, double(a[i]));
                    z[i] = F(0);
                }
                sts_word(base + (r * QN + q) * NT * 16, z);
            }
        }
    }
};

// add the posterior x .* beta of a site to the emission row of its observation
template <typename F, int MT, int NT, bool ESM>
__device__ __forceinline__ void posterior_to_emission(const F (&beta)[MT], const F (&x)[MT], int ob, Grad<F, MT, ESM> &g,
                                                      EmisAcc<F, MT, NT> &ea) {
    if constexpr (ESM) {
        ea.add(ob, beta, x);
    } else {
        const F is0 = ob == 0 ? F(1) : F(0);
        const F is1 = ob == 1 ? F(1) : F(0);
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            const F gam = x[k] * beta[k];
            g.e0[k] = fma(gam, is0, g.e0[k]);
            g.e1[k] = fma(gam, is1, g.e1[k]);
        }
    }
}

// One adjoint step for site t.  Invariant: beta . alpha == 1 at every site, where alpha is the
// (lazily scaled) forward vector the forward passes actually carried.  On entry beta belongs to
// "after site t" (the caller has already multiplied it by the factor the forward pass applied to
// its vector right after this site, if any); x = forward vector before the site.  On exit beta
// belongs to "before site t" (beta . x == 1), and the posterior x .* beta of site t-1 has been added
// to the emission row of ob_prev (the observation at site t-1; -1 when there is none).
//
// With w = emis(ob) .* beta:  beta'_i = sum_{j<i} b_j w_j + d_i w_i + u_i sum_{j>i} v_j w_j
//   d ll/d b_j += (sum_{i>j} x_i) w_j      d ll/d d_j += x_j w_j
//   d ll/d u_i += x_i sum_{j>i} v_j w_j    d ll/d v_j += (sum_{i<j} u_i x_i) w_j
// (TNT: threads per column of the emission table - NT for the per-thread tables, 1 for a table shared by
// the warp, see psmc_uniform.cuh)
template <typename F, int MT, int T, int NT, bool ESM, int TNT = NT>
__device__ __forceinline__ void backward_site(F (&beta)[MT], const F (&x)[MT], int ob, int ob_prev,
                                              const Params<F, MT> &p, const PartnerCoef<F, MT, T, true> &pc,
                                              const EmisTable<F, MT, TNT> &et, int sub, Grad<F, MT, ESM> &g,
                                              EmisAcc<F, MT, NT> &ea) {
    F w[MT];
    et.get(ob, w);
#pragma unroll
    for (int k = 0; k < MT; ++k) w[k] *= beta[k];
    F q_run = F(0), s_run = F(0), b_run = F(0), x_run = F(0);
    if constexpr (T == 2) {
        F tw[2] = {F(0), F(0)}, tx[2] = {F(0), F(0)};
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            tw[k & 1] = fma(pc.cw[k], w[k], tw[k & 1]);
            tx[k & 1] = fma(pc.cx[k], x[k], tx[k & 1]);
        }
        const F ow = __shfl_xor_sync(0xffffffffu, tw[0] + tw[1], 1, 2);
        const F ox = __shfl_xor_sync(0xffffffffu, tx[0] + tx[1], 1, 2);
        q_run = sub == 0 ? ow : F(0);
        s_run = sub == 0 ? ox : F(0);
        b_run = sub == 0 ? F(0) : ow;
        x_run = sub == 0 ? F(0) : ox;
    } else if constexpr (T > 2) {
        F tq = F(0), ts = F(0), tb = F(0), tx = F(0);
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            tq = fma(p.v[k], w[k], tq);
            ts += x[k];
            tb = fma(p.b[k], w[k], tb);
            tx = fma(p.u[k], x[k], tx);
        }
        q_run = lanes_after<F, T>(tq, sub);
        s_run = lanes_after<F, T>(ts, sub);
        b_run = lanes_before<F, T>(tb, sub);
        x_run = lanes_before<F, T>(tx, sub);
    }
    // Ascending heads  Pb_k = sum_{j<k} b_j w_j,  Px_k = sum_{j<k} u_j x_j  and descending tails
    // Q_k = sum_{j>k} v_j w_j,  S_k = sum_{j>k} x_j  are independent chains: walk them from both ends
    // in one loop so that their latencies overlap.  (beta is dead once w has been formed, so it
    // collects the new vector: head part first, tail part added where the sweeps have crossed.)
    F tailq[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int k = i, j = MT - 1 - i;
        beta[k] = fma(p.d[k], w[k], b_run);
        if constexpr (!ESM) g.d[k] = fma(x[k], w[k], g.d[k]);
        g.v[k] = fma(x_run, w[k], g.v[k]);
        b_run = fma(p.b[k], w[k], b_run);
        x_run = fma(p.u[k], x[k], x_run);
        tailq[j] = q_run;
        g.u[j] = fma(x[j], q_run, g.u[j]);
        g.b[j] = fma(s_run, w[j], g.b[j]);
        q_run = fma(p.v[j], w[j], q_run);
        s_run += x[j];
    }
#pragma unroll
    for (int k = 0; k < MT; ++k) beta[k] = fma(p.u[k], tailq[k], beta[k]);
    posterior_to_emission<F, MT, NT, ESM>(beta, x, ob_prev, g, ea);
}

constexpr int kNorm = 4;  // the forward vector is rescaled after every kNorm-th site of a segment (8 measured slower: the unrolled blocks outgrow the instruction cache)

// With ~227 KB of shared memory per SM in use there is practically no L1, so every segment's
// observation / checkpoint load would see L2 or DRAM latency; a hint one segment ahead costs no
// registers.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Observations of one K-site segment packed in 64-bit words (K = 8 or 16).
template <int K> struct ObsWords {
    uint64_t w[K / 8];
    __device__ __forceinline__ void load(const int8_t *row, int64_t site0) {
#pragma unroll
        for (int i = 0; i < K / 8; ++i) w[i] = __ldg(reinterpret_cast<const unsigned long long *>(row + site0) + i);
    }
    // the kNorm (4 or 8) observations of sites [k0, k0 + kNorm), k0 a multiple of kNorm, in the low bytes
    __device__ __forceinline__ uint64_t block(int k0) const {
        uint64_t word = w[0];
        if constexpr (K == 16) word = (k0 & 8) ? w[1] : w[0];
        if constexpr (kNorm == 4) word >>= (k0 & 4) * 8;
        return word;
    }
    __device__ __forceinline__ int at(int k) const { return byte_of(block(k & ~(kNorm - 1)), k & (kNorm - 1)); }
    static __device__ __forceinline__ int byte_of(uint64_t blk, int j) {
        return static_cast<int>(static_cast<int8_t>((blk >> (8 * j)) & 0xffu));
    }
};

// emission table always; the ring of forward vectors and block scales only for the gradient kernel
// thread-per-pair layouts (MT >= 16) keep the emission-row accumulators in shared memory
template <typename F, int MT, int K, int NT, bool GRAD> constexpr size_t smem_bytes() {
    return (emis_acc_in_smem<MT>() ? 0 : 16) + sizeof(F) * (size_t(emis_acc_in_smem<MT>() ? 3 : 2) * MT * NT +
                             (GRAD ? size_t(K) * MT * NT + size_t(K / kNorm) * NT + (emis_acc_in_smem<MT>() ? size_t(3) * MT * NT : 0) : 0));
}
// checkpoint scratch bytes per resident warp
template <typename F, int MT, int K> __host__ __device__ constexpr int64_t ckpt_bytes_per_warp(int64_t L) {
    return ((L + K - 1) / K) * int64_t(MT) * 32 * int64_t(sizeof(F));
}

// Register budget: the register file is 16 K registers per SM sub-partition, so what matters is
// the number of warps per sub-partition w = ceil(NT * MINB / 128): 255 registers at w <= 2, 168 at
// w = 3, 128 at w = 4 (allocation granularity: 8 registers per thread).
constexpr int max_regs(int nt, int minb) {
    const int w = (nt * minb + 127) / 128;
    const int r = (16384 / w) / 32 / 8 * 8;
    return r > 255 ? 255 : r;
}

// F is the arithmetic type, IO the type of the parameter / gradient buffers (IO = float with
// F = double is the precision-escalation variant of a single-precision kernel object).
// SEG = segment mode of the gradient kernel (parallel-in-time gradient of larger minibatches, see
// boundary_sweep_kernel): group grp scores segment grp / seg_ctas of the chunks it covers, started from
// bnd_alpha, closed with bnd_beta, partial gradient to seg_dlog.
template <typename F, int MT, int T, int K, bool GRAD, int NT, int MINB, typename IO = F, bool SEG = false>
__global__ void __maxnreg__(max_regs(NT, MINB)) psmc_loglik_kernel(const KernelArgs a) {
    constexpr int kThreads = NT;
    constexpr int kWarps = NT / 32;
    static_assert(NT % 32 == 0, "whole warps");
    constexpr int M = MT * T;
    constexpr int PW = 32 / T;  // pairs per warp
    using V = typename Vec<F>::type;
    constexpr int W = Vec<F>::W;
    constexpr int QN = MT / W;
    static_assert(MT % 4 == 0 && K % 8 == 0 && K % kNorm == 0 && (kNorm == 4 || kNorm == 8) && T <= 32, "layout assumptions");
    // one decision per block of kNorm sites instead of one per site (see pass 1)
    constexpr bool kBlockBranch = T == 1 && !SEG;  // (the segment-mode build has no register to spare)

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // Shared-memory map, in 128-bit words (all indices are ints, see phb_smem):
    //   [0]                         one word of ones (emission of a missing observation)
    //   emission table              [2][QN][NT]
    //   ring of forward vectors     per warp [K][QN][32]                       (gradient kernel)
    //   block scale factors         per warp [K / kNorm][32] scalars           (gradient kernel)
    //   emission-row accumulators   [3][QN][NT]                                 (MT >= 16 gradient kernel)
    constexpr bool ESM = GRAD && emis_acc_in_smem<MT>();
    constexpr int kOnesWords = EmisTable<F, MT, NT>::kOnesRow ? 0 : 1;
    constexpr int kTabWords = kOnesWords + EmisTable<F, MT, NT>::kRows * QN * NT;
    constexpr int kRingWords = K * QN * NT;
    constexpr int kScaleWords = (K / kNorm) * NT / W;  // scalars packed W per word
    const uint32_t smem0 = smem_base_addr();
    EmisTable<F, MT, NT> et;
    et.ones = smem0;
    et.base = smem0 + (kOnesWords + threadIdx.x) * 16;
    const uint32_t seg_a = smem0 + (kTabWords + warp * (K * QN * 32) + lane) * 16;  // + (k * QN + q) * 32 * 16
    const uint32_t scale_a = smem0 + (kTabWords + kRingWords) * 16 + (warp * (K / kNorm * 32) + lane) * uint32_t(sizeof(F));
    EmisAcc<F, MT, NT> ea;
    ea.base = smem0 + (kTabWords + kRingWords + kScaleWords + threadIdx.x) * 16;
    if constexpr (kOnesWords == 1) {
        if (threadIdx.x == 0) {
            F one[W];
#pragma unroll
            for (int i = 0; i < W; ++i) one[i] = F(1);
            sts_word(smem0, one);
        }
        __syncthreads();  // the only block-level barrier: publishes the word of ones
    }

    const int sub = lane % T;
    const int lp = lane / T;
    static_assert(!SEG || (GRAD && sizeof(IO) == sizeof(F)), "segment mode: gradient kernel, plain buffers");
    // Work enumeration: CHUNK major, particle minor (pair p -> chunk p / B, particle p % B), so that the lanes of
    // a warp score consecutive particles of the SAME chunk (a warp straddles two chunks only where 32 / T does
    // not divide B): their observation requests carry one address and are served as one transaction, and all
    // warps that score a chunk share its lines in L1 / L2.  (Rounding B up to whole warps per chunk would make
    // the observations provably warp-uniform, but costs padding lanes and, at the benchmark shape, a ninth round
    // of groups on the persistent grid: measured 7 % slower.)
    const int64_t s_eff = listed_chunks(a);
    const int64_t n_pairs = (SEG ? a.S : s_eff) * a.B;
    const int64_t n_groups = (!SEG && a.s_list) ? (n_pairs + kWarps * PW - 1) / (kWarps * PW) : a.n_groups;
    const int64_t L_max = SEG ? a.seg_len : a.L;
    const int64_t warp_slot = int64_t(blockIdx.x) * kWarps + warp;
    const IO *params6 = static_cast<const IO *>(a.params6);
    const IO *pi_g = static_cast<const IO *>(a.pi);
    V *ck = GRAD ? reinterpret_cast<V *>(static_cast<char *>(a.ckpt) + warp_slot * ckpt_bytes_per_warp<F, MT, K>(L_max)) + lane
                 : nullptr;
    constexpr int kFlushSegs = kFlushSites / K;
    static_assert((kFlushSegs & (kFlushSegs - 1)) == 0, "flush cadence must be a power of two");
    // fp64 gradient slots of this thread: a.gacc[i * gacc_stride + gacc_index]; recomputed where
    // needed instead of being carried in registers through the hot loops
#define PHB_GACC_STRIDE (int64_t(gridDim.x) * kThreads)
#define PHB_GACC_BASE (a.gacc + int64_t(blockIdx.x) * kThreads + threadIdx.x)

    // the work list is walked per CTA (not per warp) so that every loop bound below is provably
    // uniform and the shuffles need no reconvergence guards
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        // SEG: group -> (segment of the chunk, group within the segment); L = sites of that segment
        const int64_t n_real = a.seg_local ? a.seg_local : a.seg_count;        // segments scored by this launch
        const bool warm = SEG && a.warm_len > 0 && grp / a.seg_ctas == n_real;  // the groups of the warm-up term
        const int64_t pseg = SEG ? (warm ? 0 : a.seg_first + grp / a.seg_ctas) : 0;
        const int64_t L = SEG ? (warm ? a.warm_len : min(a.seg_len, a.L - pseg * a.seg_len)) : a.L;
        const int64_t n_seg = (L + K - 1) / K;
        const int64_t pair_raw = ((SEG ? grp % a.seg_ctas : grp) * kWarps + warp) * PW + lp;
        const bool writer = pair_raw < n_pairs;
        const int64_t pair_idx = writer ? pair_raw : n_pairs - 1;  // idle lanes shadow the last pair
        const int64_t slot = pair_idx / a.B;                        // position in the (sub-)list of chunks
        const int64_t pb = pair_idx - slot * a.B;
        const int64_t ps = (!SEG && a.s_list) ? int64_t(a.s_list[slot]) : slot;
        const int64_t chunk_pair = pb * a.S + ps;  // index into ll / dlog ([B, S]) and the boundary vectors
        int64_t pair = chunk_pair;
        if constexpr (SEG) pair = chunk_pair * (n_real + (a.warm_len > 0 ? 1 : 0)) + grp / a.seg_ctas;  // slot in seg_dlog
        Params<F, MT> p;
        p.load(params6 + pb * a.pstride_b + ps * a.pstride_s + sub * MT, M);
        et.fill(params6 + pb * a.pstride_b + ps * a.pstride_s + sub * MT, M);
        PartnerCoef<F, MT, T, GRAD> pc;
        pc.init(p, sub);
        int64_t row = a.inds[ps];
        const bool bad_row = row < 0 || row >= a.n_rows;
        if (bad_row) {
            if (sub == 0) atomicOr(a.err_flag, 1);  // reported by phb_sync(); this pair's ll becomes NaN
            row = 0;
        }
        const int8_t *obs = a.data + row * a.pitch + (SEG ? pseg * a.seg_len : 0);
        const IO *pi_p = (SEG && !warm) ? static_cast<const IO *>(a.bnd_alpha) + (chunk_pair * (a.seg_count + 1) + pseg) * M + sub * MT
                                        : pi_g + pb * a.pistride_b + ps * a.pistride_s + sub * MT;

        // ------------------------------------------------------------------ pass 1: forward
        // (segment mode after the sweeps: the checkpoints and the vector behind the segment are already there)
        const IO *ext_ck = nullptr;
        if constexpr (SEG) {
            if (a.ext_ck != nullptr && !warm)
                ext_ck = static_cast<const IO *>(a.ext_ck) + (chunk_pair * a.ext_ck_count + pseg * (a.seg_len / K)) * M + sub * MT;
        }
        F x[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) x[k] = F((ext_ck != nullptr ? pi_p + M : pi_p)[k]);
        double ll = 0.0;
        ObsWords<K> ow_next;
        if (ext_ck == nullptr) ow_next.load(obs, 0);
        for (int64_t seg = 0; seg < (ext_ck != nullptr ? 0 : n_seg); ++seg) {
            if (GRAD && seg > 0) {
#pragma unroll
                for (int q = 0; q < QN; ++q) ck[(seg * QN + q) * 32] = pack(&x[q * W]);
            }
            const ObsWords<K> ow = ow_next;
            if (seg + 1 < n_seg) ow_next.load(obs, (seg + 1) * K);  // one segment ahead
            const int len = int(min(int64_t(K), L - seg * K));
            F acc = F(0);
            for (int kb = 0; kb < len; kb += kNorm) {
                const uint64_t blk = ow.block(kb);
                // one decision per block instead of one per site: the sites of a full block form one
                // basic block (the scheduler can overlap neighbouring sites), the ragged tail of a
                // chunk goes through a compact loop
                // (thread-per-pair layouts only: measured +3 % at M = 16; the multi-lane layouts of M = 32 / 64
                // run out of registers with the second code path and keep one predicated site per branch)
                if (kBlockBranch && kb + kNorm <= len) {
#pragma unroll
                    for (int j = 0; j < kNorm; ++j) forward_site<F, MT, T, GRAD, NT>(x, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                } else if constexpr (kBlockBranch) {
#pragma unroll 1
                    for (int j = 0; kb + j < len; ++j) forward_site<F, MT, T, GRAD, NT>(x, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                } else {
#pragma unroll
                    for (int j = 0; j < kNorm; ++j)
                        if (kb + j < len) forward_site<F, MT, T, GRAD, NT>(x, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                }
                const F tot = pair_sum<F, MT, T>(x);
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
        if (bad_row) ll = __longlong_as_double(0x7ff8000000000000LL);
        if constexpr (!SEG) {
            if (writer && sub == 0) a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
        } else {
            if (warm && writer && sub == 0) a.warm_ll[chunk_pair] = ll;
        }
        if (!SEG && writer && a.alpha_out != nullptr) {
            IO *ao = static_cast<IO *>(a.alpha_out) + pair * M + sub * MT;
#pragma unroll
            for (int k = 0; k < MT; ++k) ao[k] = IO(x[k]);
        }

        if constexpr (GRAD) {
            // -------------------------------------------------------------- pass 2: adjoint
            Grad<F, MT, ESM> g;
            g.clear();
            if constexpr (ESM) ea.clear();
            F beta[MT];
            if constexpr (SEG) {
                // the adjoint vector behind this segment, scaled so that beta . x == 1
                const F *bb = static_cast<const F *>(a.bnd_beta) + (chunk_pair * (a.seg_count + 1) + pseg + 1) * M + sub * MT;
                F dot = F(0);
#pragma unroll
                for (int k = 0; k < MT; ++k) {
                    beta[k] = warm ? F(1) : bb[k];  // (the warm-up term ends after its last site: a vector of ones)
                    dot = fma(beta[k], x[k], dot);
                }
                dot = F(1) / lanes_total<F, T>(dot);
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] *= dot;
                posterior_to_emission<F, MT, NT, ESM>(beta, x, int(obs[L - 1]), g, ea);
            } else {
                // after the last site: beta = 1 / sum(x) so that beta . x == 1, and the posterior of
                // the last site is x .* beta
                const F tot = fast_rcp<F>(pair_sum<F, MT, T>(x));
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] = tot;
                posterior_to_emission<F, MT, NT, ESM>(beta, x, int(obs[L - 1]), g, ea);
            }
#pragma unroll 1
            for (int i = 0; i < 6 * MT; ++i) PHB_GACC_BASE[int64_t(i) * PHB_GACC_STRIDE] = 0.0;
            ObsWords<K> ow_ahead;  // observations of the next segment to process, one segment ahead
            ow_ahead.load(obs, (n_seg - 1) * K);
            for (int64_t seg = n_seg - 1; seg >= 0; --seg) {
                const ObsWords<K> ow = ow_ahead;
                if (seg > 0) {
                    ow_ahead.load(obs, (seg - 1) * K);
                    if (seg > 1) {
                        if (ext_ck != nullptr) {
                            // this thread's record of the next group: MT values, one 32-byte sector per 8
#pragma unroll
                            for (int q = 0; q < (MT * int(sizeof(F)) + 31) / 32; ++q)
                                prefetch_l1(reinterpret_cast<const char *>(ext_ck + (seg - 1) * M) + q * 32);
                        } else {
                            prefetch_l2(&ck[(seg - 1) * QN * 32]);
                        }
                    }
                }
                const int len = int(min(int64_t(K), L - seg * K));
                // re-run the forward steps of this segment, keeping every input vector
                F xs[MT];
                if (seg == 0) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) xs[k] = F(pi_p[k]);
                } else if (ext_ck != nullptr) {
                    // (16-byte aligned: records are M values long, a lane's part MT, both multiples of 4)
#pragma unroll
                    for (int q = 0; q < QN; ++q) unpack<F>(reinterpret_cast<const V *>(ext_ck + seg * M)[q], &xs[q * W]);
                } else {
#pragma unroll
                    for (int q = 0; q < QN; ++q) unpack<F>(ck[(seg * QN + q) * 32], &xs[q * W]);
                }
                for (int kb = 0; kb < len; kb += kNorm) {
                    const uint64_t blk = ow.block(kb);
                    if (kBlockBranch && kb + kNorm <= len) {
#pragma unroll
                        for (int j = 0; j < kNorm; ++j) {
#pragma unroll
                            for (int q = 0; q < QN; ++q) sts_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xs[q * W]);
                            forward_site<F, MT, T, GRAD, NT>(xs, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                        }
                    } else if constexpr (kBlockBranch) {
#pragma unroll 1
                        for (int j = 0; kb + j < len; ++j) {
#pragma unroll
                            for (int q = 0; q < QN; ++q) sts_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xs[q * W]);
                            forward_site<F, MT, T, GRAD, NT>(xs, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < kNorm; ++j) {
                            if (kb + j < len) {
#pragma unroll
                                for (int q = 0; q < QN; ++q) sts_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xs[q * W]);
                                forward_site<F, MT, T, GRAD, NT>(xs, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
                            }
                        }
                    }
                    const F inv = fast_rcp<F>(pair_sum<F, MT, T>(xs));
                    sts_scalar(scale_a + (kb / kNorm) * 32 * uint32_t(sizeof(F)), inv);
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
                for (int kb = ((len - 1) / kNorm) * kNorm; kb >= 0; kb -= kNorm) {
                    const uint64_t blk = ow.block(kb);
                    // the observation just before this block; for the first block of the segment it is
                    // the last one of the previous segment, whose words are already on their way
                    const int ob_before = kb > 0 ? ow.at(kb - 1) : (seg > 0 ? ow_ahead.at(K - 1) : -1);
                    // the forward pass multiplied its vector by `scale` after the last site of this
                    // block; carrying the same factor on beta keeps beta . alpha == 1
                    const F scale = lds_scalar_f(scale_a + (kb / kNorm) * 32 * uint32_t(sizeof(F)), F(0));
#pragma unroll
                    for (int k = 0; k < MT; ++k) beta[k] *= scale;
                    if (kBlockBranch && kb + kNorm <= len) {
#pragma unroll
                        for (int j = kNorm - 1; j >= 0; --j) {
                            F xin[MT];
#pragma unroll
                            for (int q = 0; q < QN; ++q) lds_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xin[q * W]);
                            const int ob_prev = j > 0 ? ObsWords<K>::byte_of(blk, j - 1) : ob_before;
                            backward_site<F, MT, T, NT, ESM>(beta, xin, ObsWords<K>::byte_of(blk, j), ob_prev, p, pc, et, sub, g, ea);
                        }
                    } else if constexpr (!kBlockBranch) {
#pragma unroll
                        for (int j = kNorm - 1; j >= 0; --j) {
                            if (kb + j < len) {
                                F xin[MT];
#pragma unroll
                                for (int q = 0; q < QN; ++q) lds_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xin[q * W]);
                                const int ob_prev = j > 0 ? ObsWords<K>::byte_of(blk, j - 1) : ob_before;
                                backward_site<F, MT, T, NT, ESM>(beta, xin, ObsWords<K>::byte_of(blk, j), ob_prev, p, pc, et, sub, g, ea);
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int j = len - 1 - kb; j >= 0; --j) {
                            F xin[MT];
#pragma unroll
                            for (int q = 0; q < QN; ++q) lds_word(seg_a + ((kb + j) * QN + q) * 32 * 16, &xin[q * W]);
                            const int ob_prev = j > 0 ? ObsWords<K>::byte_of(blk, j - 1) : ob_before;
                            backward_site<F, MT, T, NT, ESM>(beta, xin, ObsWords<K>::byte_of(blk, j), ob_prev, p, pc, et, sub, g, ea);
                        }
                    }
                }
                if ((seg & (kFlushSegs - 1)) == 0) {
                    double *acc = PHB_GACC_BASE;
                    const int64_t stride = PHB_GACC_STRIDE;
                    flush_row<F, MT>(g.b, 0, acc, stride);
                    if constexpr (!ESM) flush_row<F, MT>(g.d, 1, acc, stride);
                    flush_row<F, MT>(g.u, 2, acc, stride);
                    flush_row<F, MT>(g.v, 3, acc, stride);
                    if constexpr (ESM) {
                        ea.flush(acc, stride);
                    } else {
                        flush_row<F, MT>(g.e0, 4, acc, stride);
                        flush_row<F, MT>(g.e1, 5, acc, stride);
                    }
                }
            }
            if (writer) {
                IO *out = static_cast<IO *>(SEG ? a.seg_dlog : a.dlog) + pair * 7 * M + sub * MT;
                const double *gacc = PHB_GACC_BASE;
                const int64_t gacc_stride = PHB_GACC_STRIDE;
#pragma unroll
                for (int k = 0; k < MT; ++k) {
                    F val[7];
                    const double gb = gacc[int64_t(0 * MT + k) * gacc_stride] * double(p.b[k]);
                    const double gv = gacc[int64_t(3 * MT + k) * gacc_stride] * double(p.v[k]);
                    const double ge0 = gacc[int64_t(4 * MT + k) * gacc_stride], ge1 = gacc[int64_t(5 * MT + k) * gacc_stride];
                    val[0] = F(gb);
                    val[2] = F(gacc[int64_t(2 * MT + k) * gacc_stride] * double(p.u[k]));
                    val[3] = F(gv);
                    val[4] = F(ge0);
                    val[5] = F(ge1);
                    val[6] = beta[k] * F(pi_p[k]);
                    if constexpr (ESM) {
                        // slot 1 holds the posterior mass of the missing observations plus the posterior of the
                        // START vector (the first adjoint step books it under "no observation"), which is val[6]
                        const double arrivals = ge0 + ge1 + gacc[int64_t(1 * MT + k) * gacc_stride] - double(val[6]);
                        val[1] = p.d[k] == F(0) ? F(0) : F(arrivals - gb - gv);
                    } else {
                        val[1] = F(gacc[int64_t(1 * MT + k) * gacc_stride] * double(p.d[k]));
                    }
#pragma unroll
                    for (int r = 0; r < 7; ++r) out[r * M + k] = IO((!SEG && a.out_mode) ? F(out[r * M + k]) - val[r] : val[r]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// "Store-all" variant of the gradient kernel for SMALL minibatches (the reference's default is
// S <= 5 chunks, mcmc.py:119-121, i.e. a few thousand pairs).  Those runs are bound by the serial
// depth of the recursion, not by throughput: three passes (forward, recompute, adjoint) of L
// dependent site steps.  With few pairs every forward input vector fits in HBM
// (pairs * L * M * 4 B: 8 GB for 2 500 pairs of 50 500 bins), so pass 1 stores them all and the
// adjoint pass streams them back with the loads of the next block issued one block ahead - two
// passes instead of three, nothing recomputed.  One CTA handles exactly one group of pairs (the
// host launches enough CTAs), so the scratch is indexed by the global warp index.
template <typename F, int MT, int T, int NT, int MINB>
__global__ void __maxnreg__(max_regs(NT, MINB)) psmc_loglik_storeall_kernel(const KernelArgs a) {
    constexpr int M = MT * T;
    constexpr int PW = 32 / T;
    constexpr int kWarps = NT / 32;
    using V = typename Vec<F>::type;
    constexpr int W = Vec<F>::W;
    constexpr int QN = MT / W;
    constexpr int K = 8;  // only used to size the observation words
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t smem0 = smem_base_addr();
    EmisTable<F, MT, NT> et;
    et.ones = smem0;
    et.base = smem0 + (1 + threadIdx.x) * 16;
    EmisAcc<F, MT, NT> ea;  // unused (register accumulators), needed by the shared site functions
    ea.base = 0;
    if (threadIdx.x == 0) {
        F one[W];
#pragma unroll
        for (int i = 0; i < W; ++i) one[i] = F(1);
        sts_word(smem0, one);
    }
    __syncthreads();

    const int sub = lane % T;
    const int lp = lane / T;
    const int64_t s_eff = listed_chunks(a);
    const int64_t n_pairs = a.B * s_eff;
    const int64_t cta = blockIdx.x;
    const int64_t L = a.L;
    // the grid is sized for the whole minibatch; with a sub-list the surplus warps have nothing to do
    // (no block-level barrier follows, so whole warps may leave)
    if ((cta * kWarps + warp) * PW >= n_pairs) return;
    const int64_t n_blocks = (L + kNorm - 1) / kNorm;
    const int64_t warp_slot = int64_t(blockIdx.x) * kWarps + warp;
    const F *params6 = static_cast<const F *>(a.params6);
    const F *pi_g = static_cast<const F *>(a.pi);
    V *xall = reinterpret_cast<V *>(a.xall) + warp_slot * L * QN * 32 + lane;  // + (t * QN + q) * 32
    F *sall = static_cast<F *>(a.sall) + warp_slot * n_blocks * 32 + lane;     // + block * 32

    const int64_t pair_raw = (cta * kWarps + warp) * PW + lp;
    const bool writer = pair_raw < n_pairs;
    const PairIndex pidx = pair_index(a, writer ? pair_raw : n_pairs - 1, s_eff);
    const int64_t pb = pidx.b, ps = pidx.s, pair = pidx.out;
    Params<F, MT> p;
    p.load(params6 + pb * a.pstride_b + ps * a.pstride_s + sub * MT, M);
    et.fill(params6 + pb * a.pstride_b + ps * a.pstride_s + sub * MT, M);
    PartnerCoef<F, MT, T, true> pc;
    pc.init(p, sub);
    int64_t row = a.inds[ps];
    const bool bad_row = row < 0 || row >= a.n_rows;
    if (bad_row) {
        if (sub == 0) atomicOr(a.err_flag, 1);
        row = 0;
    }
    const int8_t *obs = a.data + row * a.pitch;
    const F *pi_p = pi_g + pb * a.pistride_b + ps * a.pistride_s + sub * MT;

    // ---------------------------------------------------------------- pass 1: forward, keep everything
    F x[MT];
#pragma unroll
    for (int k = 0; k < MT; ++k) x[k] = pi_p[k];
    double ll = 0.0;
    F acc = F(0);
    // kNorm (= 4) observation bytes per block, requested one block ahead (rows are padded to a
    // multiple of 16 bytes, so reading the word that straddles the end of the row is safe)
    uint32_t blk_next = __ldg(reinterpret_cast<const unsigned int *>(obs));
    for (int64_t blk_i = 0; blk_i < n_blocks; ++blk_i) {
        const int64_t t0 = blk_i * kNorm;
        const uint32_t blk = blk_next;
        if (blk_i + 1 < n_blocks) blk_next = __ldg(reinterpret_cast<const unsigned int *>(obs + t0 + kNorm));
        const int len = int(min(int64_t(kNorm), L - t0));
        // (a block-level decision instead of one per site was measured SLOWER here: 18.0 vs 17.0 ms at S = 16)
#pragma unroll
        for (int j = 0; j < kNorm; ++j) {
            if (j < len) {
#pragma unroll
                for (int q = 0; q < QN; ++q) xall[((t0 + j) * QN + q) * 32] = pack(&x[q * W]);
                forward_site<F, MT, T, true, NT>(x, p, pc, et, ObsWords<K>::byte_of(blk, j), sub);
            }
        }
        const F tot = pair_sum<F, MT, T>(x);
        const F inv = fast_rcp<F>(tot);
        sall[blk_i * 32] = inv;
#pragma unroll
        for (int j = 0; j < MT; ++j) x[j] *= inv;
        acc += log2_of<F>(tot);
        if ((blk_i & 3) == 3) {
            ll += double(acc);
            acc = F(0);
        }
    }
    ll = (ll + double(acc)) * 0.69314718055994530942;
    if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
        if (sub == 0) atomicOr(a.err_flag, 2);
    }
    if (bad_row) ll = __longlong_as_double(0x7ff8000000000000LL);
    if (writer && sub == 0) a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;

    // ---------------------------------------------------------------- pass 2: adjoint, streaming the vectors back
    Grad<F, MT, false> g;
    g.clear();
    F beta[MT];
    {
        const F tot = fast_rcp<F>(pair_sum<F, MT, T>(x));
#pragma unroll
        for (int k = 0; k < MT; ++k) beta[k] = tot;
        posterior_to_emission<F, MT, NT, false>(beta, x, int(obs[L - 1]), g, ea);
    }
    double *gacc_base = a.gacc + int64_t(blockIdx.x) * NT + threadIdx.x;
    const int64_t gacc_stride = int64_t(gridDim.x) * NT;
#pragma unroll 1
    for (int i = 0; i < 6 * MT; ++i) gacc_base[int64_t(i) * gacc_stride] = 0.0;
    // vectors of the block being processed and of the next one (loaded one block ahead)
    F xb[kNorm][MT], xn[kNorm][MT];
    auto load_block = [&](int64_t blk_i, F(&dst)[kNorm][MT]) {
        const int64_t t0 = blk_i * kNorm;
#pragma unroll
        for (int j = 0; j < kNorm; ++j) {
            if (t0 + j < L) {
#pragma unroll
                for (int q = 0; q < QN; ++q) unpack<F>(xall[((t0 + j) * QN + q) * 32], &dst[j][q * W]);
            }
        }
    };
    load_block(n_blocks - 1, xn);
    uint32_t obs_next = __ldg(reinterpret_cast<const unsigned int *>(obs + (n_blocks - 1) * kNorm));
    F scale_next = sall[(n_blocks - 1) * 32];
    F post[MT];  // forward vector after the most recently processed site (for the drift control)
#pragma unroll
    for (int k = 0; k < MT; ++k) post[k] = x[k];
    constexpr int kAhead = 8;  // blocks of lead for the L2 prefetch hints (DRAM latency >> one block)
    for (int64_t blk_i = n_blocks - 1; blk_i >= 0; --blk_i) {
        const int64_t t0 = blk_i * kNorm;
#pragma unroll
        for (int j = 0; j < kNorm; ++j) {
#pragma unroll
            for (int k = 0; k < MT; ++k) xb[j][k] = xn[j][k];
        }
        const uint32_t blk = obs_next;
        const F scale = scale_next;
        if (blk_i > 0) {
            // everything the next block needs is requested now, one block ahead
            load_block(blk_i - 1, xn);
            obs_next = __ldg(reinterpret_cast<const unsigned int *>(obs + t0 - kNorm));
            scale_next = sall[(blk_i - 1) * 32];
        }
        if (blk_i >= kAhead) {
#pragma unroll
            for (int j = 0; j < kNorm; ++j) {
#pragma unroll
                for (int q = 0; q < QN; ++q) prefetch_l2(&xall[(((blk_i - kAhead) * kNorm + j) * QN + q) * 32]);
            }
        }
        // the observation just before this block is the last byte of the word that is on its way
        const int ob_before = t0 > 0 ? ObsWords<K>::byte_of(obs_next, kNorm - 1) : -1;
        const int len = int(min(int64_t(kNorm), L - t0));
        if ((blk_i & 3) == 3) {
            // every 16 sites: re-impose beta . alpha == 1 against round-off drift
            F dot = F(0);
#pragma unroll
            for (int k = 0; k < MT; ++k) dot = fma(post[k], beta[k], dot);
            dot = fast_rcp<F>(lanes_total<F, T>(dot));
#pragma unroll
            for (int k = 0; k < MT; ++k) beta[k] *= dot;
        }
#pragma unroll
        for (int k = 0; k < MT; ++k) beta[k] *= scale;
#pragma unroll
        for (int j = kNorm - 1; j >= 0; --j) {
            if (j < len) {
                const int ob_prev = j > 0 ? ObsWords<K>::byte_of(blk, j - 1) : ob_before;
                backward_site<F, MT, T, NT, false>(beta, xb[j], ObsWords<K>::byte_of(blk, j), ob_prev, p, pc, et, sub, g, ea);
            }
        }
#pragma unroll
        for (int k = 0; k < MT; ++k) post[k] = xb[0][k];
        if ((blk_i & (kFlushSites / kNorm - 1)) == 0) {
            flush_row<F, MT>(g.b, 0, gacc_base, gacc_stride);
            flush_row<F, MT>(g.d, 1, gacc_base, gacc_stride);
            flush_row<F, MT>(g.u, 2, gacc_base, gacc_stride);
            flush_row<F, MT>(g.v, 3, gacc_base, gacc_stride);
            flush_row<F, MT>(g.e0, 4, gacc_base, gacc_stride);
            flush_row<F, MT>(g.e1, 5, gacc_base, gacc_stride);
        }
    }
    if (writer) {
        F *out = static_cast<F *>(a.dlog) + pair * 7 * M + sub * MT;
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            F val[7];
            val[0] = F(gacc_base[int64_t(0 * MT + k) * gacc_stride] * double(p.b[k]));
            val[1] = F(gacc_base[int64_t(1 * MT + k) * gacc_stride] * double(p.d[k]));
            val[2] = F(gacc_base[int64_t(2 * MT + k) * gacc_stride] * double(p.u[k]));
            val[3] = F(gacc_base[int64_t(3 * MT + k) * gacc_stride] * double(p.v[k]));
            val[4] = F(gacc_base[int64_t(4 * MT + k) * gacc_stride]);
            val[5] = F(gacc_base[int64_t(5 * MT + k) * gacc_stride]);
            val[6] = beta[k] * pi_p[k];
#pragma unroll
            for (int r = 0; r < 7; ++r) out[r * M + k] = a.out_mode ? out[r * M + k] - val[r] : val[r];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Parallel-in-time FORWARD evaluation of FEW, LONG pairs: the reference's ELPD (mcmc.py:213-238) scores
// whole un-chunked test contigs (millions of bins) for every particle every 10th iteration - a few
// hundred pairs, which cannot fill the GPU, so the launch lasts L x (latency of one dependent site
// step), ~80 ns per site whatever the lane layout.  The recursion is linear in the forward vector:
// cut the sequence into G segments and propagate, for every segment, the M unit vectors instead of
// the (unknown) incoming vector.  That yields the segment's transfer operator
//     T_g[i, :] = e_i A diag(emis(ob_t)) ... (all sites of the segment),
// M x more arithmetic, but M x G x as many independent recursions - enough to fill the machine - and
// the G operators are chained afterwards in float64:  alpha_{g+1} = alpha_g T_g  (M^2 per segment).
// Exact (no mixing / burn-in approximation); every row of T_g is an ordinary scaled forward
// recursion, so the accuracy is that of the sequential kernel.
//
// transfer_rows_kernel: one thread per (pair, segment, unit vector) - the M rows of one operator are
// adjacent lanes, so they read the same observations and parameters - running the same site step
// as psmc_loglik_kernel (T = 1: thread-private state, no shuffles).  Output per row: the final
// vector rescaled to sum 1 and the log2 of the total scale.
struct TransferArgs {
    KernelArgs k;          // data, inds, B, S, params / pi layout, ll, err_flag, out_mode of the evaluation
    int64_t n_seg;         // G
    int64_t seg_len;       // sites per segment (multiple of 16; the last segment takes the rest)
    // Operator blocks are stored SEGMENT-MAJOR in slots of segs_per_slot segments: block (pair, g) lives in slot
    // g / segs_per_slot at [g % segs_per_slot][pair][M rows][M].  One process: a single slot holding all G
    // segments.  Time-axis sharding: every process fills its own slot and an all-gather makes all of them
    // visible everywhere (slot = rows then row_log2 of that process, so ONE collective moves both).
    float *rows;
    double *row_log2;      // same addressing, [M rows] per block
    int64_t segs_per_slot;
    int64_t slot_stride_rows;  // floats between consecutive slots of `rows`
    int64_t slot_stride_log;   // doubles between consecutive slots of `row_log2`
    int64_t seg_first;     // transfer_rows_kernel computes segments [seg_first, seg_first + n_seg_local)
    int64_t n_seg_local;
};
__device__ __forceinline__ int64_t op_block(const TransferArgs &ta, int64_t pair, int64_t g, int64_t n_pairs, int64_t *slot) {
    *slot = g / ta.segs_per_slot;
    return (g - *slot * ta.segs_per_slot) * n_pairs + pair;
}
__device__ __forceinline__ const float *op_rows(const TransferArgs &ta, int64_t pair, int64_t g, int64_t n_pairs, int M) {
    int64_t slot;
    const int64_t blk = op_block(ta, pair, g, n_pairs, &slot);
    return ta.rows + slot * ta.slot_stride_rows + blk * M * M;
}
__device__ __forceinline__ const double *op_log2(const TransferArgs &ta, int64_t pair, int64_t g, int64_t n_pairs, int M) {
    int64_t slot;
    const int64_t blk = op_block(ta, pair, g, n_pairs, &slot);
    return ta.row_log2 + slot * ta.slot_stride_log + blk * M;
}

template <typename F, int M, int NT> __global__ void __maxnreg__(max_regs(NT, 3)) transfer_rows_kernel(const TransferArgs ta) {
    constexpr int MT = M, T = 1, K = 8;
    const KernelArgs &a = ta.k;
    const uint32_t smem0 = smem_base_addr();
    EmisTable<F, MT, NT> et;
    et.ones = smem0;
    et.base = smem0 + ((EmisTable<F, MT, NT>::kOnesRow ? 0 : 1) + threadIdx.x) * 16;
    if constexpr (!EmisTable<F, MT, NT>::kOnesRow) {
        if (threadIdx.x == 0) {
            F one[Vec<F>::W];
#pragma unroll
            for (int i = 0; i < Vec<F>::W; ++i) one[i] = F(1);
            sts_word(smem0, one);
        }
        __syncthreads();
    }
    const int64_t n_virtual = a.B * a.S * ta.n_seg_local * M;
    const int64_t vraw = int64_t(blockIdx.x) * NT + threadIdx.x;
    const bool writer = vraw < n_virtual;
    const int64_t v = writer ? vraw : n_virtual - 1;
    const int unit = int(v % M);
    const int64_t seg = ta.seg_first + (v / M) % ta.n_seg_local;
    const int64_t pair = v / (M * ta.n_seg_local);
    const int64_t pb = pair / a.S, ps = pair % a.S;
    const F *par = static_cast<const F *>(a.params6) + pb * a.pstride_b + ps * a.pstride_s;
    Params<F, MT> p;
    p.load(par, M);
    et.fill(par, M);
    PartnerCoef<F, MT, T, false> pc;
    int64_t row = a.inds[ps];
    if (row < 0 || row >= a.n_rows) row = 0;  // reported (and the result made NaN) by chain_transfer_kernel
    const int64_t site0 = seg * ta.seg_len;
    const int64_t len = min(ta.seg_len, a.L - site0);
    const int8_t *obs = a.data + row * a.pitch + site0;

    F x[MT];
#pragma unroll
    for (int k = 0; k < MT; ++k) x[k] = k == unit ? F(1) : F(0);
    double log2_scale = 0.0;
    const int64_t n_blk = (len + K - 1) / K;
    ObsWords<K> ow_next;
    ow_next.load(obs, 0);
    for (int64_t blk = 0; blk < n_blk; ++blk) {
        const ObsWords<K> ow = ow_next;
        if (blk + 1 < n_blk) ow_next.load(obs, (blk + 1) * K);
        const int n = int(min(int64_t(K), len - blk * K));
        F acc = F(0);
        for (int kb = 0; kb < n; kb += kNorm) {
            const uint64_t word = ow.block(kb);
            if (kb + kNorm <= n) {
#pragma unroll
                for (int j = 0; j < kNorm; ++j) forward_site<F, MT, T, false, NT>(x, p, pc, et, ObsWords<K>::byte_of(word, j), 0);
            } else {
#pragma unroll 1
                for (int j = 0; kb + j < n; ++j) forward_site<F, MT, T, false, NT>(x, p, pc, et, ObsWords<K>::byte_of(word, j), 0);
            }
            const F tot = pair_sum<F, MT, T>(x);
            const F inv = fast_rcp<F>(tot);
#pragma unroll
            for (int j = 0; j < MT; ++j) x[j] *= inv;
            acc += log2_of<F>(tot);
        }
        log2_scale += double(acc);
    }
    if (writer) {
        // exact normalisation of what is handed on (the reciprocal above is approximate)
        const F tot = pair_sum<F, MT, T>(x);
        float *out = const_cast<float *>(op_rows(ta, pair, seg, a.B * a.S, M)) + unit * M;
#pragma unroll
        for (int k = 0; k < MT; ++k) out[k] = float(x[k] / tot);
        const_cast<double *>(op_log2(ta, pair, seg, a.B * a.S, M))[unit] = log2_scale + double(log2_of<F>(tot));
    }
}

// chain_transfer_kernel: one thread per pair;  alpha <- alpha T_g  in float64, g = 0 .. G-1, with a common
// scale per segment taken from the largest contributing row.
template <typename F, int M> __global__ void chain_transfer_kernel(const TransferArgs ta) {
    const KernelArgs &a = ta.k;
    const int64_t pair = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pair >= a.B * a.S) return;
    const int64_t pb = pair / a.S, ps = pair % a.S;
    const F *pi_p = static_cast<const F *>(a.pi) + pb * a.pistride_b + ps * a.pistride_s;
    double alpha[M];
    double tot = 0.0;
    for (int k = 0; k < M; ++k) tot += (alpha[k] = double(pi_p[k]));
    double ll2 = log2(tot);  // log2 of everything divided out so far
    for (int k = 0; k < M; ++k) alpha[k] /= tot;
    for (int64_t g = 0; g < ta.n_seg; ++g) {
        const float *rows = op_rows(ta, pair, g, a.B * a.S, M);
        const double *lg = op_log2(ta, pair, g, a.B * a.S, M);
        double top = -1e300;
        for (int i = 0; i < M; ++i)
            if (alpha[i] > 0.0 && lg[i] > top) top = lg[i];
        double next[M];
        for (int k = 0; k < M; ++k) next[k] = 0.0;
        for (int i = 0; i < M; ++i) {
            if (!(alpha[i] > 0.0)) continue;
            const double w = alpha[i] * exp2(lg[i] - top);
            for (int k = 0; k < M; ++k) next[k] += w * double(rows[i * M + k]);
        }
        tot = 0.0;
        for (int k = 0; k < M; ++k) tot += next[k];
        ll2 += top + log2(tot);
        for (int k = 0; k < M; ++k) alpha[k] = next[k] / tot;
    }
    double ll = ll2 * 0.69314718055994530942;
    const int64_t row = a.inds[ps];
    if (row < 0 || row >= a.n_rows) {
        atomicOr(a.err_flag, 1);
        ll = __longlong_as_double(0x7ff8000000000000LL);
    } else if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
        atomicOr(a.err_flag, 2);
    }
    a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
}

// The same operators serve the GRADIENT of few pairs (the reference's default minibatch is ONE chunk
// for a single genome, mcmc.py:119-121: 500 pairs, 14 ms of dependent site steps): chained forwards
// they give the forward vector entering every segment, chained backwards (beta_g = T_g beta_{g+1},
// beta_G = 1) the adjoint vector behind it.  With both boundary vectors known the segments are
// independent: the gradient kernel runs over every segment as if it were a short chunk (SEG mode of
// psmc_loglik_kernel: started from alpha_g, closed with beta_{g+1} rescaled to beta . alpha == 1), and
// the partial gradients are added up.
//
// chain_boundaries_kernel: M lanes per pair (lane k owns component k); writes ll and the G + 1 boundary
// vectors of both kinds.  Launch with 128 threads per CTA.
template <int M> __device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int o = M / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o, M));
    return v;
}
template <int M> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = M / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, M);
    return v;
}
template <typename F, int M> __global__ void chain_boundaries_kernel(const TransferArgs ta, F *bnd_alpha, F *bnd_beta) {
    const KernelArgs &a = ta.k;
    const int64_t n_pairs = a.B * a.S;
    const int64_t pair_raw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / M;
    const bool writer = pair_raw < n_pairs;
    const int64_t pair = writer ? pair_raw : n_pairs - 1;  // idle groups shadow the last pair (whole warps shuffle)
    const int k = threadIdx.x % M;
    const int64_t pb = pair / a.S, ps = pair % a.S;
    const int64_t G = ta.n_seg;
    const F *pi_p = static_cast<const F *>(a.pi) + pb * a.pistride_b + ps * a.pistride_s;
    F *al = bnd_alpha + pair * (G + 1) * M;
    F *be = bnd_beta + pair * (G + 1) * M;
    double v = double(pi_p[k]);
    double tot = group_sum<M>(v);
    double ll2 = log2(tot);
    v /= tot;
    if (writer) al[k] = F(v);
    // (the chain is sequential over the segments: the operator of the NEXT segment is requested while this one is
    // applied, otherwise every step waits for L2)
    float col[M];  // column k of the segment's operator
    double lg_col = 0.0;
    {
        const float *rows = op_rows(ta, pair, 0, n_pairs, M);
#pragma unroll
        for (int i = 0; i < M; ++i) col[i] = rows[i * M + k];
        lg_col = op_log2(ta, pair, 0, n_pairs, M)[k];
    }
    for (int64_t g = 0; g < G; ++g) {
        float cur[M];
#pragma unroll
        for (int i = 0; i < M; ++i) cur[i] = col[i];
        const double lg = lg_col;
        if (g + 1 < G) {
            const float *rows = op_rows(ta, pair, g + 1, n_pairs, M);
#pragma unroll
            for (int i = 0; i < M; ++i) col[i] = rows[i * M + k];
            lg_col = op_log2(ta, pair, g + 1, n_pairs, M)[k];
        }
        const double top = group_max<M>(v > 0.0 ? lg : -1e300);
        const double w = v > 0.0 ? v * exp2(lg - top) : 0.0;
        double next = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) next += __shfl_sync(0xffffffffu, w, i, M) * double(cur[i]);
        tot = group_sum<M>(next);
        ll2 += top + log2(tot);
        v = next / tot;
        if (writer) al[(g + 1) * M + k] = F(v);
    }
    // adjoint boundaries: beta_g[i] = 2^lg_i sum_j T_g[i, j] beta_{g+1}[j], kept at maximum 1 (lane k owns row k)
    v = 1.0;
    if (writer) be[G * M + k] = F(1);
    {
        const float *rows = op_rows(ta, pair, G - 1, n_pairs, M);
#pragma unroll
        for (int j = 0; j < M; ++j) col[j] = rows[k * M + j];  // (row k now)
        lg_col = op_log2(ta, pair, G - 1, n_pairs, M)[k];
    }
    for (int64_t g = G - 1; g >= 0; --g) {
        float cur[M];
#pragma unroll
        for (int j = 0; j < M; ++j) cur[j] = col[j];
        const double lg = lg_col;
        if (g > 0) {
            const float *rows = op_rows(ta, pair, g - 1, n_pairs, M);
#pragma unroll
            for (int j = 0; j < M; ++j) col[j] = rows[k * M + j];
            lg_col = op_log2(ta, pair, g - 1, n_pairs, M)[k];
        }
        const double top = group_max<M>(lg);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) acc += double(cur[j]) * __shfl_sync(0xffffffffu, v, j, M);
        const double next = acc * exp2(lg - top);
        v = next / group_max<M>(next);
        if (writer) be[g * M + k] = F(v);
    }
    if (writer && k == 0 && !outputs_skipped(a, ps)) {
        double ll = ll2 * 0.69314718055994530942;
        const int64_t row = a.inds[ps];
        if (row < 0 || row >= a.n_rows) {
            atomicOr(a.err_flag, 1);
            ll = __longlong_as_double(0x7ff8000000000000LL);
        } else if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
            atomicOr(a.err_flag, 2);
        }
        a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
    }
}

// ---- time-axis sharding over processes: every process multiplies ITS segment operators into one operator per
// pair (chain_product_kernel), only those (world operators per pair instead of G) are all-gathered, and
// chain_boundaries_sharded_kernel chains them to the vectors entering / leaving the process's slice, then the
// process's own segment operators to the boundary vectors inside it.
//
// One forward chain step for the M lanes of a group (lane k owns component k):  v <- normalised (v T),  ll2 += log2 of
// what was divided out.  rows / lg: the operator's [M][M] rows and [M] log2 row scales.
template <int M> __device__ __forceinline__ void chain_forward_step(double &v, double &ll2, const float *__restrict__ rows,
                                                                    const double *__restrict__ lg_rows, int k) {
    const double lg = lg_rows[k];
    const double top = group_max<M>(v > 0.0 ? lg : -1e300);
    const double w = v > 0.0 ? v * exp2(lg - top) : 0.0;
    double next = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) next += __shfl_sync(0xffffffffu, w, i, M) * double(rows[i * M + k]);
    const double tot = group_sum<M>(next);
    ll2 += top + log2(tot);
    v = next / tot;
}
// One adjoint chain step:  v <- (T v) / max  (lane k owns row k of T)
template <int M> __device__ __forceinline__ void chain_backward_step(double &v, const float *__restrict__ rows,
                                                                     const double *__restrict__ lg_rows, int k) {
    const double lg = lg_rows[k];
    const double top = group_max<M>(lg);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) acc += double(rows[k * M + j]) * __shfl_sync(0xffffffffu, v, j, M);
    const double next = acc * exp2(lg - top);
    v = next / group_max<M>(next);
}

// chain_product_kernel: M x M lanes per pair; the M lanes of group r own ROW r of the running product - a
// distribution over the M states with its own log2 scale, the same representation as a segment operator - and push
// it through the next operator exactly like a forward chain step.  Launch with a multiple of M threads per CTA.
template <int M> __global__ void chain_product_kernel(const TransferArgs ta, float *__restrict__ out_rows, double *__restrict__ out_log2) {
    const KernelArgs &a = ta.k;
    const int64_t n_pairs = a.B * a.S;
    const int64_t t_raw = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool writer = t_raw < n_pairs * M * M;
    const int64_t t = writer ? t_raw : n_pairs * M * M - 1;  // idle lanes shadow the last entry (whole groups shuffle)
    const int64_t pair = t / (M * M);
    const int r = int((t / M) % M), k = int(t % M);
    double v = r == k ? 1.0 : 0.0, lg_acc = 0.0;  // a process without segments contributes the identity
    if (ta.n_seg_local > 0) {
        v = double(op_rows(ta, pair, ta.seg_first, n_pairs, M)[r * M + k]);
        lg_acc = op_log2(ta, pair, ta.seg_first, n_pairs, M)[r];
    }
    for (int64_t g = ta.seg_first + 1; g < ta.seg_first + ta.n_seg_local; ++g)
        chain_forward_step<M>(v, lg_acc, op_rows(ta, pair, g, n_pairs, M), op_log2(ta, pair, g, n_pairs, M), k);
    if (writer) {
        out_rows[(pair * M + r) * M + k] = float(v);
        if (k == 0) out_log2[pair * M + r] = lg_acc;
    }
}

// rank_rows / rank_log2: the all-gathered operators of the processes, process r at + r * rank_stride_* ([pairs][M][M]
// floats, [pairs][M] doubles); ta: THIS process's segment operators (single slot, segment-major).  Writes the
// boundary vectors of the segments [ta.seg_first, ta.seg_first + ta.n_seg_local] and - on every process - ll.
// Launch with 128 threads per CTA.
template <typename F, int M>
__global__ void chain_boundaries_sharded_kernel(const TransferArgs ta, const float *__restrict__ rank_rows, const double *__restrict__ rank_log2,
                                                int64_t rank_stride_rows, int64_t rank_stride_log, int rank, int world,
                                                F *__restrict__ bnd_alpha, F *__restrict__ bnd_beta) {
    const KernelArgs &a = ta.k;
    const int64_t n_pairs = a.B * a.S;
    const int64_t pair_raw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / M;
    const bool writer = pair_raw < n_pairs;
    const int64_t pair = writer ? pair_raw : n_pairs - 1;
    const int k = threadIdx.x % M;
    const int64_t pb = pair / a.S, ps = pair % a.S;
    const int64_t G = ta.n_seg;
    const F *pi_p = static_cast<const F *>(a.pi) + pb * a.pistride_b + ps * a.pistride_s;
    F *al = bnd_alpha + pair * (G + 1) * M;
    F *be = bnd_beta + pair * (G + 1) * M;
    const int64_t g_lo = ta.seg_first, g_hi = ta.seg_first + ta.n_seg_local;
    // forward over the processes: the vector entering this process's slice, and the total log-likelihood
    double v = double(pi_p[k]);
    double tot = group_sum<M>(v);
    double ll2 = log2(tot);
    v /= tot;
    double v_in = v;
    for (int r = 0; r < world; ++r) {
        if (r == rank) v_in = v;
        chain_forward_step<M>(v, ll2, rank_rows + r * rank_stride_rows + pair * M * M, rank_log2 + r * rank_stride_log + pair * M, k);
    }
    // backward over the processes behind this one: the adjoint vector leaving this process's slice
    double w_out = 1.0;
    for (int r = world - 1; r > rank; --r)
        chain_backward_step<M>(w_out, rank_rows + r * rank_stride_rows + pair * M * M, rank_log2 + r * rank_stride_log + pair * M, k);
    // inside the slice: this process's own segment operators
    v = v_in;
    double unused = 0.0;
    if (writer) al[g_lo * M + k] = F(v);
    for (int64_t g = g_lo; g < g_hi; ++g) {
        chain_forward_step<M>(v, unused, op_rows(ta, pair, g, n_pairs, M), op_log2(ta, pair, g, n_pairs, M), k);
        if (writer) al[(g + 1) * M + k] = F(v);
    }
    v = w_out;
    if (writer) be[g_hi * M + k] = F(v);
    for (int64_t g = g_hi - 1; g >= g_lo; --g) {
        chain_backward_step<M>(v, op_rows(ta, pair, g, n_pairs, M), op_log2(ta, pair, g, n_pairs, M), k);
        if (writer) be[g * M + k] = F(v);
    }
    if (writer && k == 0 && !outputs_skipped(a, ps)) {
        double ll = ll2 * 0.69314718055994530942;
        const int64_t row = a.inds[ps];
        if (row < 0 || row >= a.n_rows) {
            atomicOr(a.err_flag, 1);
            ll = __longlong_as_double(0x7ff8000000000000LL);
        } else if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
            atomicOr(a.err_flag, 2);
        }
        a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
    }
}

// dlog[pair] = sum over the segments of their partial gradients (rows b .. emis1); the pi row is the
// one of the first segment.  One thread per output entry.
template <typename F>
__global__ void sum_segments_kernel(const F *__restrict__ seg_dlog, int64_t n_pairs, int64_t G, int M, F *__restrict__ dlog,
                                    int out_mode, const KernelArgs a) {
    const int64_t n = n_pairs * 7 * M;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t pair = i / (7 * M);
    const int64_t rm = i % (7 * M);
    if (outputs_skipped(a, pair % a.S)) return;
    const bool warm = a.warm_len > 0;  // one more slot per pair: the warm-up term, to be subtracted
    const F *src = seg_dlog + pair * (G + (warm ? 1 : 0)) * 7 * M + rm;
    // (G = segments of this launch; the pi row belongs to the chunk's first segment: zero from a launch that
    // does not hold it, i.e. under time-axis sharding)
    double acc = (rm < 6 * M || a.seg_first == 0) ? double(src[0]) : 0.0;
    if (rm < 6 * M)
        for (int64_t g = 1; g < G; ++g) acc += double(src[g * 7 * M]);
    if (warm) {
        acc -= double(src[G * 7 * M]);
        if (rm == 0) a.ll[pair] -= a.warm_ll[pair];
    }
    dlog[i] = out_mode ? F(double(dlog[i]) - acc) : F(acc);
}

// For MORE pairs than the operators pay for (the reference's default minibatch of 5 chunks: 2 500
// pairs, still far too few to fill the GPU) the boundary vectors come from two SEQUENTIAL sweeps that run
// side by side instead: the first CTAs run the plain forward recursion and leave the forward vector at every
// segment boundary (and the log-likelihood), the others run the adjoint recursion without any gradient
// bookkeeping, beta <- A (emis .* beta), from the end of the chunk and leave the adjoint vectors.  Each
// is one dependent pass of the cheap kind (65 ns per site with the low-latency site functions below, ~110 ns with the generic ones); the expensive gradient passes then run over
// all segments at once (psmc_loglik_kernel, SEG mode) as in the operator variant.
template <typename F, int MT, int T, int NT>
__device__ __forceinline__ void adjoint_only_site(F (&beta)[MT], const Params<F, MT> &p, const EmisTable<F, MT, NT> &et, int ob, int sub) {
    F w[MT];
    et.get(ob, w);
#pragma unroll
    for (int k = 0; k < MT; ++k) w[k] *= beta[k];
    F q_run = F(0), b_run = F(0);
    if constexpr (T > 1) {
        F tq[2] = {F(0), F(0)}, tb[2] = {F(0), F(0)};
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            tq[k & 1] = fma(p.v[k], w[k], tq[k & 1]);
            tb[k & 1] = fma(p.b[k], w[k], tb[k & 1]);
        }
        q_run = lanes_after<F, T>(tq[0] + tq[1], sub);
        b_run = lanes_before<F, T>(tb[0] + tb[1], sub);
    }
    F tailq[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int k = i, j = MT - 1 - i;
        beta[k] = fma(p.d[k], w[k], b_run);
        b_run = fma(p.b[k], w[k], b_run);
        tailq[j] = q_run;
        q_run = fma(p.v[j], w[j], q_run);
    }
#pragma unroll
    for (int k = 0; k < MT; ++k) beta[k] = fma(p.u[k], tailq[k], beta[k]);
}

// ---- low-latency site functions for the sweeps (LL = true, T >= 4) ----
// A sweep runs with about one warp per scheduler: what counts is the LENGTH of the dependency chain from
// x(t) to x(t + 1), not the instruction count.  The generic site functions scan across the T lanes of a pair
// with a butterfly (log2 T DEPENDENT shuffles, ~28 cycles each) and only then walk the local chains.  Here
//   * every lane fetches the local totals of all T - 1 partners with INDEPENDENT xor shuffles (one shuffle
//     latency) and weighs them with 0 / 1 lane masks,
//   * everything that does not need the partners' totals - the local prefix / suffix chains, the diagonal
//     term, the products with the emission row - is computed in the shadow of the shuffles,
//   * behind the shuffles remain one short multiply-add chain for the offsets and two FMAs per state,
//   * the rescaling factor of a block is applied two sites LATER (the recursion is linear: when the factor is
//     applied does not matter), so the sum / reciprocal never sits on the chain either.
// ~60 dependent cycles per site instead of ~120.
//   * the emission row is addressed by (observation & 3) * row bytes - rows emis0, emis1, ones, ones, so that -1
//     (0xff) lands on a row of ones - instead of a compare / select chain: integer instructions issue at half
//     rate, and a lone warp feels every one of them.
template <typename F, int MT, int NT> struct EmisTable4 {
    static constexpr int W = Vec<F>::W;
    static constexpr int QN = MT / W;
    static constexpr uint32_t kRowBytes = QN * NT * 16;
    static constexpr size_t kBytes = size_t(4) * kRowBytes;
    uint32_t base;  // shared address of this thread's column
    template <typename IO> __device__ __forceinline__ void fill(const IO *__restrict__ src, int M) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int q = 0; q < QN; ++q) {
                F tmp[W];
#pragma unroll
                for (int i = 0; i < W; ++i) tmp[i] = r < 2 ? F(src[(4 + r) * M + q * W + i]) : F(1);
                sts_word(base + r * kRowBytes + q * NT * 16, tmp);
            }
        }
    }
    // byte offset of the row of site j of a block of four observations
    static __device__ __forceinline__ uint32_t row_of(uint32_t blk, int j) { return ((blk >> (8 * j)) & 3u) * kRowBytes; }
    __device__ __forceinline__ void get(uint32_t row_off, F (&e)[MT]) const {
#pragma unroll
        for (int q = 0; q < QN; ++q) lds_word(base + row_off + q * NT * 16, &e[q * W]);
    }
};
// MUFU.LG2 without the denormal detour of __log2f (the totals are far from denormal)
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double lg2_fast(double x) { return log2(x); }

template <typename F, int T> struct LaneMasks {
    F before[T - 1], after[T - 1];  // partner sub ^ (o + 1) has a smaller / larger lane index than this lane
    __device__ __forceinline__ void init(int sub) {
#pragma unroll
        for (int o = 1; o < T; ++o) {
            before[o - 1] = (sub ^ o) < sub ? F(1) : F(0);
            after[o - 1] = (sub ^ o) > sub ? F(1) : F(0);
        }
    }
};
// sum of `mine` over all T lanes of the pair, independent shuffles
template <typename F, int T> __device__ __forceinline__ F lanes_total_flat(F mine) {
    F s[2] = {mine, F(0)};
#pragma unroll
    for (int o = 1; o < T; ++o) s[o & 1] += __shfl_xor_sync(0xffffffffu, mine, o, T);
    return s[0] + s[1];
}
template <typename F, int MT, int T> __device__ __forceinline__ F pair_sum_flat(const F (&x)[MT]) {
    F part[2] = {F(0), F(0)};
#pragma unroll
    for (int k = 0; k < MT; ++k) part[k & 1] += x[k];
    return lanes_total_flat<F, T>(part[0] + part[1]);
}

// The forward sweep carries the PREDICTED vector z(t) = alpha(t - 1) A (before the emission of site t) instead of
// alpha(t): the step z <- (emis .* z) A then has the shape of the adjoint step - one product with the emission row up
// front, nothing but two FMAs per state behind the shuffles - and needs a quarter fewer instructions than a step that
// multiplies by the emission row last.  w = emis .* z = alpha(t) is a by-product (the segment boundaries store it).
template <typename F, int MT, int T, int NT>
// (w comes in holding the emission row of the site - the caller loads it, one site ahead across block boundaries)
__device__ __forceinline__ void forward_site_ll(F (&z)[MT], F (&w)[MT], const Params<F, MT> &p, const LaneMasks<F, T> &lm) {
#pragma unroll
    for (int k = 0; k < MT; ++k) w[k] *= z[k];
    // the lane's own prefix / suffix chains; their ends are what the partners need from this lane (a lone warp is
    // bound by its instruction count: separate, shorter chains for the totals cost 2 MT instructions more)
    F lp[MT + 1], ls[MT + 1];
    lp[0] = F(0);
    ls[MT] = F(0);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int k = i, j = MT - 1 - i;
        lp[k + 1] = fma(p.u[k], w[k], lp[k]);
        ls[j] = ls[j + 1] + w[j];
    }
    const F my_u = lp[MT], my_x = ls[0];
    F su[T - 1], sx[T - 1];
#pragma unroll
    for (int o = 1; o < T; ++o) {
        su[o - 1] = __shfl_xor_sync(0xffffffffu, my_u, o, T);
        sx[o - 1] = __shfl_xor_sync(0xffffffffu, my_x, o, T);
    }
    // in the shadow of the shuffles: the lane's own part
    F loc[MT];
#pragma unroll
    for (int k = 0; k < MT; ++k) loc[k] = fma(p.b[k], ls[k + 1], fma(p.v[k], lp[k], p.d[k] * w[k]));
    F pre = lm.before[0] * su[0], suf_off = lm.after[0] * sx[0];
#pragma unroll
    for (int o = 2; o < T; ++o) {
        pre = fma(lm.before[o - 1], su[o - 1], pre);
        suf_off = fma(lm.after[o - 1], sx[o - 1], suf_off);
    }
#pragma unroll
    for (int k = 0; k < MT; ++k) z[k] = fma(p.v[k], pre, fma(p.b[k], suf_off, loc[k]));
}

template <typename F, int MT, int T, int NT>
// (w comes in holding the emission row of the site)
__device__ __forceinline__ void adjoint_only_site_ll(F (&beta)[MT], F (&w)[MT], const Params<F, MT> &p, const LaneMasks<F, T> &lm) {
#pragma unroll
    for (int k = 0; k < MT; ++k) w[k] *= beta[k];
    F lb[MT + 1], lq[MT + 1];
    lb[0] = F(0);
    lq[MT] = F(0);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int k = i, j = MT - 1 - i;
        lb[k + 1] = fma(p.b[k], w[k], lb[k]);
        lq[j] = fma(p.v[j], w[j], lq[j + 1]);
    }
    const F my_q = lq[0], my_b = lb[MT];
    F sq[T - 1], sb[T - 1];
#pragma unroll
    for (int o = 1; o < T; ++o) {
        sq[o - 1] = __shfl_xor_sync(0xffffffffu, my_q, o, T);
        sb[o - 1] = __shfl_xor_sync(0xffffffffu, my_b, o, T);
    }
    F loc[MT];
#pragma unroll
    for (int k = 0; k < MT; ++k) loc[k] = fma(p.u[k], lq[k + 1], fma(p.d[k], w[k], lb[k]));
    F off_q = lm.after[0] * sq[0], off_b = lm.before[0] * sb[0];
#pragma unroll
    for (int o = 2; o < T; ++o) {
        off_q = fma(lm.after[o - 1], sq[o - 1], off_q);
        off_b = fma(lm.before[o - 1], sb[o - 1], off_b);
    }
#pragma unroll
    for (int k = 0; k < MT; ++k) beta[k] = fma(p.u[k], off_q, loc[k] + off_b);
}

// Shared memory of a sweep CTA: the larger of the two emission tables.
template <typename F, int MT, int NT> constexpr size_t sweep_smem_bytes() {
    return smem_bytes<F, MT, 8, NT, false>() > EmisTable4<F, MT, NT>::kBytes ? smem_bytes<F, MT, 8, NT, false>() : EmisTable4<F, MT, NT>::kBytes;
}

// One step of either kind through a common interface: Site<LL>::forward / ::adjoint.
template <typename F, int MT, int T, int NT, bool LL> struct SweepSite;
template <typename F, int MT, int T, int NT> struct SweepSite<F, MT, T, NT, true> {
    EmisTable4<F, MT, NT> et;
    LaneMasks<F, T> lm;
    __device__ __forceinline__ void init(uint32_t smem0, const F *par, int M, const Params<F, MT> &, int sub) {
        et.base = smem0 + threadIdx.x * 16;
        et.fill(par, M);
        lm.init(sub);
    }
    // (z-form: x is the predicted vector, w receives alpha of this site)
    __device__ __forceinline__ void row(uint32_t blk, int j, F (&e)[MT]) const { et.get(EmisTable4<F, MT, NT>::row_of(blk, j), e); }
    __device__ __forceinline__ void forward(F (&x)[MT], F (&w)[MT], const Params<F, MT> &p, uint32_t blk, int j, int) const {
        row(blk, j, w);
        forward_site_ll<F, MT, T, NT>(x, w, p, lm);
    }
    __device__ __forceinline__ void adjoint(F (&beta)[MT], const Params<F, MT> &p, uint32_t blk, int j, int) const {
        F w[MT];
        row(blk, j, w);
        adjoint_only_site_ll<F, MT, T, NT>(beta, w, p, lm);
    }
    static __device__ __forceinline__ F total(const F (&x)[MT]) { return pair_sum_flat<F, MT, T>(x); }
};

// The sum of a vector over the pair, split in two: issue() leaves the lane's own part and the partners' parts (the
// shuffles are in flight), finish() adds them up.  The sweeps issue at the end of a block and finish after the first
// site of the NEXT block: a lone warp issues in order, and the shuffle -> add -> reciprocal -> logarithm -> double
// chain at the end of the block body stalled it for ~45 cycles per block.
template <typename F, int MT, int T> struct PendingTotal {
    F own, part[T > 1 ? T - 1 : 1];
    __device__ __forceinline__ void one() {
        own = F(1);
#pragma unroll
        for (int o = 1; o < T; ++o) part[o - 1] = F(0);
    }
    __device__ __forceinline__ void issue(const F (&x)[MT]) {
        F a[2] = {F(0), F(0)};
#pragma unroll
        for (int k = 0; k < MT; ++k) a[k & 1] += x[k];
        own = a[0] + a[1];
#pragma unroll
        for (int o = 1; o < T; ++o) part[o - 1] = __shfl_xor_sync(0xffffffffu, own, o, T);
    }
    __device__ __forceinline__ F finish() const {
        F a[2] = {own, F(0)};
#pragma unroll
        for (int o = 1; o < T; ++o) a[o & 1] += part[o - 1];
        return a[0] + a[1];
    }
};
template <typename F, int MT, int T, int NT> struct SweepSite<F, MT, T, NT, false> {
    EmisTable<F, MT, NT> et;
    PartnerCoef<F, MT, T, false> pc;
    __device__ __forceinline__ void init(uint32_t smem0, const F *par, int M, const Params<F, MT> &p, int sub) {
        constexpr int W = Vec<F>::W;
        et.ones = smem0;
        et.base = smem0 + ((EmisTable<F, MT, NT>::kOnesRow ? 0 : 1) + threadIdx.x) * 16;
        if constexpr (!EmisTable<F, MT, NT>::kOnesRow) {
            if (threadIdx.x == 0) {
                F one[W];
#pragma unroll
                for (int i = 0; i < W; ++i) one[i] = F(1);
                sts_word(smem0, one);
            }
            __syncthreads();
        }
        et.fill(par, M);
        pc.init(p, sub);
    }
    // (x is alpha itself; w is a copy for the common interface)
    __device__ __forceinline__ void forward(F (&x)[MT], F (&w)[MT], const Params<F, MT> &p, uint32_t blk, int j, int sub) const {
        forward_site<F, MT, T, false, NT>(x, p, pc, et, ObsWords<8>::byte_of(blk, j), sub);
#pragma unroll
        for (int k = 0; k < MT; ++k) w[k] = x[k];
    }
    __device__ __forceinline__ void adjoint(F (&beta)[MT], const Params<F, MT> &p, uint32_t blk, int j, int sub) const {
        adjoint_only_site<F, MT, T, NT>(beta, p, et, ObsWords<8>::byte_of(blk, j), sub);
    }
    static __device__ __forceinline__ F total(const F (&x)[MT]) { return pair_sum<F, MT, T>(x); }
};

// DIR: 0 = forward sweep, 1 = adjoint sweep (compile-time, so that a kernel holding two lane layouts carries one
// direction of each).  The block loops count in 32 bits and carry as little bookkeeping as possible: about a
// quarter of the time of the first version went into 64-bit index arithmetic, the predicated fp64 accumulation of
// the log-likelihood and the select chain of the emission row (profiles/r02_ncu_sweep_summary.md).
template <typename F, int MT, int T, int NT, bool LL, int DIR>
__device__ __forceinline__ void boundary_sweep_body(const KernelArgs &a, const int64_t cta) {
    static_assert(!LL || T >= 2, "the low-latency site functions need partner lanes");
    static_assert(kNorm == 4, "one 32-bit word of observations per block");
    constexpr int M = MT * T;
    constexpr int PW = 32 / T;
    constexpr int kWarps = NT / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % T;
    const int lp = lane / T;
    constexpr bool backwards = DIR != 0;
    const int64_t n_pairs = a.B * a.S;
    const int64_t pair_raw = (cta * kWarps + warp) * PW + lp;
    const bool writer = pair_raw < n_pairs;
    const int64_t pair = writer ? pair_raw : n_pairs - 1;
    const int64_t pb = pair / a.S, ps = pair % a.S;
    const F *par = static_cast<const F *>(a.params6) + pb * a.pstride_b + ps * a.pstride_s + sub * MT;
    Params<F, MT> p;
    p.load(par, M);
    SweepSite<F, MT, T, NT, LL> site;
    site.init(smem_base_addr(), par, M, p, sub);
    int64_t row = a.inds[ps];
    const bool bad_row = row < 0 || row >= a.n_rows;
    if (bad_row) row = 0;
    // rows start on 16-byte boundaries and are padded to whole 16-byte words
    const uint32_t *obs = reinterpret_cast<const uint32_t *>(a.data + row * a.pitch);
    const int G = int(a.seg_count);
    const int L = int(a.L);              // (the host takes this path for L < 2^31 only)
    const int full_blocks = L / kNorm;   // the ragged tail, if any, is block number full_blocks
    const int tail = L - full_blocks * kNorm;
    const int seg_blocks = int(a.seg_len) / kNorm;  // segment lengths are multiples of 16
    F *bnd = static_cast<F *>(const_cast<void *>(backwards ? a.bnd_beta : a.bnd_alpha)) + pair * (G + 1) * M + sub * MT;

    if constexpr (!backwards) {
        const F *pi_p = static_cast<const F *>(a.pi) + pb * a.pistride_b + ps * a.pistride_s + sub * MT;
        F x[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) x[k] = pi_p[k];
        double ll;
        {
            const F tot = pair_sum<F, MT, T>(x);
            ll = double(log2_of<F>(tot));
#pragma unroll
            for (int k = 0; k < MT; ++k) {
                x[k] = x[k] / tot;
                if (writer) bnd[k] = x[k];
            }
        }
        // LL: the factor of a block is applied in the middle of the NEXT one - its sum and reciprocal have two sites
        // to arrive; otherwise at once
        F inv_pending = F(1);
        F w[MT];  // alpha of the last site
        if constexpr (LL) site.forward(x, w, p, 0xffffffffu, 0, sub);  // z(0) = pi A: a step with the emission row of ones
        // observation words are requested TWO blocks ahead: one block (~500 cycles) does not cover a miss in L1
        uint32_t blk_next = __ldg(obs);
        uint32_t blk_next2 = __ldg(obs + min(1, (L - 1) / kNorm));
        int to_boundary = seg_blocks;  // blocks until the next segment boundary
        // checkpoints for the segment passes: the vector after every K-th site (any scale), see KernelArgs::ext_ck
        // (the pointer is made opaque so that it stays in registers: re-deriving it from the kernel arguments put a
        // constant-bank load and its dependants - ~40 exposed cycles for a lone warp - in front of every store)
        F *ckp = static_cast<F *>(a.ext_ck) + (pair * a.ext_ck_count + 1) * M + sub * MT;
        asm volatile("" : "+l"(ckp));
        const bool ck_on = a.ext_ck != nullptr && writer;
        int to_ck = a.ext_ck_blocks;  // (0 without checkpoints: the counter below never comes back to zero)
        const int last_word = (L - 1) / kNorm;
        // LL: software pipelining by hand across the block boundary (the compiler does not move work over the loop's
        // back edge, and a lone warp waits in order): the emission row of a block's first site is loaded at the end of
        // the block before, and the sum over the pair started at the end of a block is finished after the first
        // site of the next one (PendingTotal)
        PendingTotal<F, MT, T> pend;
        pend.one();
        F e_first[MT];
        if constexpr (LL) site.row(blk_next, 0, e_first);
        for (int bi = 0; bi < full_blocks; ++bi) {
            const uint32_t blk = blk_next;
            blk_next = blk_next2;
            blk_next2 = __ldg(obs + min(bi + 2, last_word));
            if constexpr (LL) {
#pragma unroll
                for (int k = 0; k < MT; ++k) w[k] = e_first[k];
                forward_site_ll<F, MT, T, NT>(x, w, p, site.lm);
                const F tot = pend.finish();
                inv_pending = fast_rcp<F>(tot);
                const F lg = lg2_fast(tot);
                site.forward(x, w, p, blk, 1, sub);
#pragma unroll
                for (int k = 0; k < MT; ++k) x[k] *= inv_pending;
                site.forward(x, w, p, blk, 2, sub);
                site.forward(x, w, p, blk, 3, sub);
                ll += double(lg);
                pend.issue(x);
                site.row(blk_next, 0, e_first);
            } else {
#pragma unroll
                for (int j = 0; j < kNorm; ++j) site.forward(x, w, p, blk, j, sub);
                const F tot = site.total(x);
                inv_pending = fast_rcp<F>(tot);
#pragma unroll
                for (int k = 0; k < MT; ++k) x[k] *= inv_pending;
                ll += double(lg2_fast(tot));
            }
            if (--to_ck == 0) {
                to_ck = a.ext_ck_blocks;
                if (ck_on) {
#pragma unroll
                    for (int q = 0; q < MT / Vec<F>::W; ++q)
                        *reinterpret_cast<typename Vec<F>::type *>(ckp + q * Vec<F>::W) = pack(&w[q * Vec<F>::W]);
                }
                ckp += M;
            }
            if (--to_boundary == 0) {
                // the vector entering the next segment (approximately normalised - in LL mode short of this
                // block's factor, the decay over four sites: good enough, the segment kernel rescales)
                bnd += M;
                to_boundary = seg_blocks;
                if ((bi + 1) * kNorm < L && writer) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) bnd[k] = w[k];
                }
            }
        }
        if constexpr (LL) {
            const F tot = pend.finish();
            inv_pending = fast_rcp<F>(tot);
            ll += double(lg2_fast(tot));
#pragma unroll
            for (int k = 0; k < MT; ++k) x[k] *= inv_pending;
        }
        for (int j = 0; j < tail; ++j) site.forward(x, w, p, blk_next, j, sub);
        if (writer) {
            // behind the last segment: the vector after the last site (the segment passes that take their checkpoints
            // from this sweep close their adjoint vector against it)
            F *last = static_cast<F *>(const_cast<void *>(a.bnd_alpha)) + (pair * (G + 1) + G) * M + sub * MT;
#pragma unroll
            for (int k = 0; k < MT; ++k) last[k] = w[k];
        }
        // the reciprocals are approximate: the last total is not exactly 1.  (LL: sum z(L) = sum alpha(L - 1) because
        // the rows of A sum to one - to within 6e-8 in fp32, which is 6e-8 ABSOLUTE on the log-likelihood.)
        ll = (ll + double(log2_of<F>(pair_sum<F, MT, T>(x)))) * 0.69314718055994530942;
        if (!(ll == ll) || ll > 1e300 || ll < -1e300) {
            if (sub == 0) atomicOr(a.err_flag, 2);
        }
        if (bad_row) {
            if (sub == 0) atomicOr(a.err_flag, 1);
            ll = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (writer && sub == 0 && !outputs_skipped(a, ps)) a.ll[pair] = a.out_mode ? a.ll[pair] - ll : ll;
    } else {
        F beta[MT];
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            beta[k] = F(1);
            if (writer) bnd[int64_t(G) * M + k] = F(1);  // behind the last segment: the end of the chunk
        }
        bnd += int64_t(G) * M;
        F inv_pending = F(1);
        PendingTotal<F, MT, T> pend;
        pend.one();
        if (tail > 0) {
            const uint32_t blk = __ldg(obs + full_blocks);
            for (int j = tail - 1; j >= 0; --j) site.adjoint(beta, p, blk, j, sub);
            if constexpr (LL) {
                pend.issue(beta);
            } else {
                inv_pending = fast_rcp<F>(site.total(beta));
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] *= inv_pending;
            }
        }
        // blocks until the start of the last segment, counted from the last full block
        int to_boundary = full_blocks - (G - 1) * seg_blocks;
        if (to_boundary == 0) {  // the last segment is the ragged tail alone
            bnd -= M;
            to_boundary = seg_blocks;
            if (full_blocks > 0 && writer) {
#pragma unroll
                for (int k = 0; k < MT; ++k) bnd[k] = beta[k];
            }
        }
        uint32_t blk_next = full_blocks > 0 ? __ldg(obs + full_blocks - 1) : 0u;
        uint32_t blk_next2 = __ldg(obs + max(full_blocks - 2, 0));
        F e_first[MT];  // emission row of the first site (number 3) of the next block, see the forward sweep
        if constexpr (LL) site.row(blk_next, kNorm - 1, e_first);
        for (int bi = full_blocks - 1; bi >= 0; --bi) {
            const uint32_t blk = blk_next;
            blk_next = blk_next2;
            blk_next2 = __ldg(obs + max(bi - 2, 0));
            if constexpr (LL) {
                adjoint_only_site_ll<F, MT, T, NT>(beta, e_first, p, site.lm);
                inv_pending = fast_rcp<F>(pend.finish());
                site.adjoint(beta, p, blk, 2, sub);
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] *= inv_pending;
                site.adjoint(beta, p, blk, 1, sub);
                site.adjoint(beta, p, blk, 0, sub);
                pend.issue(beta);
                site.row(blk_next, kNorm - 1, e_first);
            } else {
#pragma unroll
                for (int j = kNorm - 1; j >= 0; --j) site.adjoint(beta, p, blk, j, sub);
                inv_pending = fast_rcp<F>(site.total(beta));
#pragma unroll
                for (int k = 0; k < MT; ++k) beta[k] *= inv_pending;
            }
            if (--to_boundary == 0) {
                // the adjoint vector behind the previous segment (any scale)
                bnd -= M;
                to_boundary = seg_blocks;
                if (bi > 0 && writer) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) bnd[k] = beta[k];
                }
            }
        }
    }
}

// Both sweeps in one launch: CTAs [0, sweep_fwd_ctas) run the forward sweep with TF lanes per pair, the others the
// adjoint sweep with TB lanes per pair.  The layouts are chosen on the host so that, whenever possible, every warp
// has a scheduler of its own (4 x 148 of them): two warps on one scheduler are issue bound and take twice as long
// per site, and the launch ends with its slowest warp.
template <typename F, int M, int TF, int TB, int NT, int MINB, bool LL>
__global__ void __maxnreg__(max_regs(NT, MINB)) boundary_sweep_kernel(const KernelArgs a) {
    if (int64_t(blockIdx.x) < a.sweep_fwd_ctas)
        boundary_sweep_body<F, M / TF, TF, NT, LL && (TF >= 2), 0>(a, blockIdx.x);
    else
        boundary_sweep_body<F, M / TB, TB, NT, LL && (TB >= 2), 1>(a, int64_t(blockIdx.x) - a.sweep_fwd_ctas);
}

// ---------------------------------------------------------------------------------------------
// Precision escalation for single-precision kernel objects.
//
// Through a long run of IDENTICAL observations (a masked centromere, the -1 padding of a contig's
// last chunk, a run of homozygosity) the forward and adjoint vectors converge to a fixed point of
// the site operator.  Near it the true change per site is smaller than half an ulp of the fp32
// components, the fp32 iteration stops moving, and the vectors are off by up to ~run length x 6e-8
// relative; a 30 000-site run cost 4e-4 on the transition rows of the gradient
// (profiles/r01_padded_chunk_accuracy.log).  Rows that contain such a run are rare, so they are simply
// evaluated with double arithmetic (same float buffers): flag_long_runs_kernel marks them once, when
// the data is made resident, and split_minibatch_kernel splits every minibatch into the two lists
// the two launches work through.
//
// A row is flagged when one of its ALIGNED windows of kRunWindow sites holds a single value; every
// constant run of >= 2 * kRunWindow - 1 sites contains such a window.  Un-flagged rows therefore
// have runs < 2047 sites: <= 6e-5 relative on any gradient entry.
constexpr int kRunWindow = 1024;

// one warp per (row, window); lane l compares bytes [32 l, 32 l + 32) of the window with its first byte
__global__ void flag_long_runs_kernel(const int8_t *__restrict__ data, int64_t n_rows, int64_t L, int64_t pitch,
                                      uint8_t *__restrict__ row_flag) {
    const int64_t n_win = L / kRunWindow;  // whole windows only
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < n_rows * n_win; i += warps) {
        const int64_t row = i / n_win, win = i % n_win;
        const uint4 *src = reinterpret_cast<const uint4 *>(data + row * pitch + win * kRunWindow) + 2 * lane;
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        const uint32_t first = __shfl_sync(0xffffffffu, lo.x, 0) & 0xffu;
        const uint32_t pat = first * 0x01010101u;
        const bool same = lo.x == pat && lo.y == pat && lo.z == pat && lo.w == pat && hi.x == pat && hi.y == pat &&
                          hi.z == pat && hi.w == pat;
        if (__all_sync(0xffffffffu, same) && lane == 0) row_flag[row] = 1;
    }
}

// One CTA.  lists = [2][S]: positions s of the minibatch whose row is un-flagged (list 0) / flagged
// (list 1); counts = [2].  Out-of-range rows go to list 0, whose kernel reports them.
__global__ void split_minibatch_kernel(const int64_t *__restrict__ inds, int64_t S, const uint8_t *__restrict__ row_flag,
                                       int64_t n_rows, int32_t *__restrict__ lists, int32_t *__restrict__ counts) {
    __shared__ int32_t n[2];
    if (threadIdx.x < 2) n[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t s = threadIdx.x; s < S; s += blockDim.x) {
        const int64_t r = inds[s];
        const int f = (r >= 0 && r < n_rows && row_flag[r]) ? 1 : 0;
        lists[f * S + atomicAdd(&n[f], 1)] = int32_t(s);
    }
    __syncthreads();
    if (threadIdx.x < 2) counts[threadIdx.x] = n[threadIdx.x];
}

// Device-side chunking (reference: _chunk_het_matrix, data.py:37-61): out[n * n_chunks + k][j] =
// clip(het[n][k * chunk_size + j], -1, 1) for j < W = chunk_size + overlap, -1 beyond the end of the
// row (also fills the pitch padding).  One thread per output byte.
__global__ void chunk_het_kernel(const int8_t *__restrict__ het, int64_t n_rows, int64_t length, int64_t chunk_size,
                                 int64_t width, int64_t n_chunks, int8_t *__restrict__ out, int64_t pitch) {
    const int64_t total = n_rows * n_chunks * pitch;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t row = i / pitch, j = i % pitch;
        const int64_t n = row / n_chunks, k = row % n_chunks;
        const int64_t src = k * chunk_size + j;
        int8_t v = -1;
        if (j < width && src < length) {
            v = het[n * length + src];
            v = v > 1 ? int8_t(1) : v;  // (values < -1 are left for fixup_rows_kernel to report)
        }
        out[i] = v;
    }
}

// Validation of an uploaded [B, S, 7, M] parameter block (host entry): bit 0 of *flags is set when
// any value is non-finite, bit 1 when rows b..emis1 of some pair differ from those of pair (b, 0)
// (i.e. the block is NOT "shared across the chunks of a particle").  Replaces two host passes over
// the block (133 MB at the benchmark shape).
template <typename F>
__global__ void validate_params_kernel(const F *__restrict__ pa, int64_t B, int64_t S, int M, int *flags) {
    const int64_t n = B * S * 7 * M;
    const int64_t blk = int64_t(7) * M;
    int bad = 0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const F v = pa[i];
        if (!(v - v == F(0))) bad |= 1;  // NaN or +-inf
        const int64_t pair = i / blk, r = i % blk;
        if (r < int64_t(6) * M && (pair % S) != 0) {
            const F first = pa[(pair - pair % S) * blk + r];
            if (!(first == v)) bad |= 2;
        }
    }
    if (bad) atomicOr(flags, bad);
}

#undef PHB_GACC_STRIDE
#undef PHB_GACC_BASE

}  // namespace phb
