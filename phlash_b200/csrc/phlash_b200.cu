// C-ABI of phlash_b200 (see include/phlash_b200.h): owns the device copy of the observation
// matrix, scratch buffers and stream, validates inputs the way the reference's host class does
// (src/phlash/gpu.py:104-151, 182-237) and dispatches to the sm_100a kernels in
// psmc_kernels.cuh.  No CPU fallback: every evaluation is a CUDA kernel launch or an error.
#include "../../include/phlash_b200.h"
#include "psmc_kernels.cuh"
#include "psmc_params.cuh"
#include "psmc_sform.cuh"
#include "psmc_support.cuh"
#include "psmc_uniform.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define PHB_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(e_ == cudaErrorMemoryAllocation ? PHB_E_NOMEM : PHB_E_CUDA, "%s: %s",    \
                        #call, cudaGetErrorString(e_));                                          \
    } while (0)

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t cap = 0;
    int64_t *counter = nullptr;  // incremented on every (re)allocation
    int reserve(size_t bytes) {
        if (bytes <= cap) return PHB_OK;
        if (counter) ++*counter;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        // grow geometrically so that alternating call sizes do not reallocate every time
        size_t want = std::max(bytes, size_t(256));
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            ptr = nullptr;
            cudaGetLastError();
            return fail(PHB_E_NOMEM, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return PHB_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct Variant {
    int M, T, K, NT;
    bool dbl, grad;
    const void *func;
    size_t smem;                                  // dynamic shared memory per CTA
    int64_t (*ckpt_bytes_per_warp)(int64_t L);    // checkpoint scratch per resident warp
};

}  // namespace

struct phb_kernel {
    int M = 0;
    int dbl = 0;
    int device = 0;
    int64_t N = 0, L = 0, pitch = 0;
    int8_t *d_data = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int *d_err = nullptr;
    int *d_flags = nullptr;  // scratch word for validate_params_kernel
    int force_T = 0;
    int num_sms = 0;
    int64_t launches = 0;
    char last_name[128] = "";
    std::unordered_map<const void *, int> occupancy;  // per kernel function: attribute set, CTAs per SM
    DeviceBuffer params, inds, ll, dlog, ckpt, gacc, xall, sall, split;
    DeviceBuffer term_params, term_ll, term_dlog, term_sums;  // scratch of the whole-term entries
    DeviceBuffer term_io;                                     // device copies of the host entry's buffers
    DeviceBuffer transfer_rows, transfer_log;                 // parallel-in-time forward evaluation
    DeviceBuffer bnd_alpha, bnd_beta, seg_dlog;               // ... and gradient
    DeviceBuffer sweep_ckpt;                                  // checkpoints left by the forward sweep for the segment passes
    DeviceBuffer warm_ll;                                     // log-likelihood of the fused warm-up term
    DeviceBuffer uniform_stage;                               // parameter rows packed for the constant bank (psmc_uniform.cuh)
    int parallel_in_time = -1;  // -1 auto, 0 never, 1 whenever possible
    int store_all_mode = -1;  // -1 auto, 0 never, 1 whenever a store-all variant exists
    // precision escalation (float objects): rows holding a long run of identical observations are
    // scored with double arithmetic, see flag_long_runs_kernel
    uint8_t *d_rowflag = nullptr;  // [N]
    int64_t n_flagged = 0;
    int escalate = 1;
    // experiment knobs, read from the environment ONCE when the object is created (never per call):
    // PHB_NT, PHB_STORE_ALL, PHB_PARALLEL_IN_TIME, PHB_PIT_SEGMENTS
    int env_nt = 0, env_store_all = -2, env_pit = -2, env_pit_segments = 0, env_sweep_tf = 0, env_sweep_tb = 0, env_uniform = 1, env_sform = 0, env_sweep_ll = 1, env_sweep_ckpt = 1, env_fuse_warmup = 1;
    bool warm_fused = false;  // the last parallel-in-time gradient launch scored the warm-up term too (KernelArgs::warm_len)
    // phb_reserve(): the dispatch runs with dry = true - every scratch buffer is sized and every kernel
    // attribute set exactly as a real call would, but nothing is launched
    bool dry = false;
    unsigned long long *d_iteration = nullptr;  // minibatch counter of phb_sample_minibatch_device
    int64_t allocations = 0;                    // device allocations made by the scratch buffers (tests: no growth after reserve)
    size_t elem() const { return dbl ? sizeof(double) : sizeof(float); }
};

namespace {

template <typename F, int MT, int T, int K, bool GRAD, int NT, int MINB, typename IO = F, bool SEG = false> Variant make_variant() {
    Variant v;
    v.M = MT * T;
    v.T = T;
    v.K = K;
    v.NT = NT;
    v.dbl = sizeof(F) == 8;
    v.grad = GRAD;
    v.func = reinterpret_cast<const void *>(&phb::psmc_loglik_kernel<F, MT, T, K, GRAD, NT, MINB, IO, SEG>);
    v.smem = phb::smem_bytes<F, MT, K, NT, GRAD>();
    v.ckpt_bytes_per_warp = [](int64_t L) { return phb::ckpt_bytes_per_warp<F, MT, K>(L); };
    return v;
}

// Every (precision, M, threads-per-pair) combination that is compiled.  Within one
// (precision, M, grad) the entries are ordered by increasing T; the dispatcher takes the first
// one that fills the GPU.  The gradient kernel keeps 4*MT parameters and 4*MT (MT = 16: the emission
// rows live in shared memory) or 6*MT accumulators in registers: MT = 16 at 255 registers, MT = 8 at 168.
const std::vector<Variant> &variants() {
    static const std::vector<Variant> table = [] {
        std::vector<Variant> t;
#define PHB_GRAD(F, MT, T, K, NT, MINB) t.push_back(make_variant<F, MT, T, K, true, NT, MINB>());
#define PHB_FWD(F, MT, T, K, NT, MINB) t.push_back(make_variant<F, MT, T, K, false, NT, MINB>());
        // ---- float, loglik + grad
        PHB_GRAD(float, 4, 1, 16, 128, 4)   // M = 4
        PHB_GRAD(float, 8, 1, 16, 128, 3)   // M = 8
        PHB_GRAD(float, 4, 2, 16, 128, 4)
        PHB_GRAD(float, 16, 1, 8, 128, 2)   // M = 16: thread per pair, 255 registers, 2 warps per sub-partition
        PHB_GRAD(float, 8, 2, 16, 128, 3)
        PHB_GRAD(float, 4, 4, 16, 128, 4)
        PHB_GRAD(float, 16, 2, 8, 128, 2)   // M = 32
        PHB_GRAD(float, 8, 4, 16, 128, 3)
        PHB_GRAD(float, 4, 8, 16, 128, 4)
        PHB_GRAD(float, 16, 4, 8, 128, 2)   // M = 64
        PHB_GRAD(float, 8, 8, 16, 128, 3)
        PHB_GRAD(float, 4, 16, 16, 128, 4)
        // ---- float, forward only
        PHB_FWD(float, 4, 1, 16, 128, 4)    // M = 4
        PHB_FWD(float, 8, 1, 16, 128, 4)    // M = 8
        PHB_FWD(float, 4, 2, 16, 128, 4)
        PHB_FWD(float, 16, 1, 8, 128, 3)    // M = 16
        PHB_FWD(float, 8, 2, 16, 128, 4)
        PHB_FWD(float, 4, 4, 16, 128, 4)
        PHB_FWD(float, 16, 2, 8, 128, 3)    // M = 32
        PHB_FWD(float, 8, 4, 16, 128, 4)
        PHB_FWD(float, 4, 8, 16, 128, 4)
        PHB_FWD(float, 16, 4, 8, 128, 3)    // M = 64
        PHB_FWD(float, 8, 8, 16, 128, 4)
        PHB_FWD(float, 4, 16, 16, 128, 4)
        // ---- double, loglik + grad
        PHB_GRAD(double, 4, 1, 8, 128, 2)   // M = 4
        PHB_GRAD(double, 4, 2, 8, 128, 2)   // M = 8
        PHB_GRAD(double, 4, 4, 8, 128, 2)   // M = 16
        PHB_GRAD(double, 4, 8, 8, 128, 2)   // M = 32
        PHB_GRAD(double, 4, 16, 8, 128, 2)  // M = 64
        // ---- double, forward only
        PHB_FWD(double, 4, 1, 8, 128, 3)    // M = 4
        PHB_FWD(double, 8, 1, 8, 128, 3)    // M = 8
        PHB_FWD(double, 4, 2, 8, 128, 3)
        PHB_FWD(double, 8, 2, 8, 128, 3)    // M = 16
        PHB_FWD(double, 4, 4, 8, 128, 3)
        PHB_FWD(double, 8, 4, 8, 128, 3)    // M = 32
        PHB_FWD(double, 4, 8, 8, 128, 3)
        PHB_FWD(double, 8, 8, 8, 128, 3)    // M = 64
        PHB_FWD(double, 4, 16, 8, 128, 3)
#undef PHB_GRAD
#undef PHB_FWD
        return t;
    }();
    return table;
}

// Thread-per-pair kernels at M = 16 (float) on the rescaled recursion (psmc_sform.cuh): 15 % fewer executed
// instructions, but measured SLOWER than psmc_loglik_kernel<float, 16, 1, ...> (see the header of that file), so they
// are only selected with PHB_SFORM=1.
const Variant *sform_variant(bool grad, bool seg) {
    auto make = [](const void *func, bool g, size_t smem) {
        Variant v;
        v.M = 16;
        v.T = 1;
        v.K = 8;
        v.NT = 128;
        v.dbl = false;
        v.grad = g;
        v.func = func;
        v.smem = smem;
        v.ckpt_bytes_per_warp = [](int64_t L) { return phb::ckpt_bytes_per_warp<float, 16, 8>(L); };
        return v;
    };
    static const Variant grad_v = make(reinterpret_cast<const void *>(&phb::psmc_sform_kernel<8, true, 128, 2, false>), true,
                                       phb::sform_smem_bytes<8, 128, true>());
    static const Variant seg_v = make(reinterpret_cast<const void *>(&phb::psmc_sform_kernel<8, true, 128, 2, true>), true,
                                      phb::sform_smem_bytes<8, 128, true>());
    static const Variant fwd_v = make(reinterpret_cast<const void *>(&phb::psmc_sform_kernel<8, false, 128, 3, false>), false,
                                      phb::sform_smem_bytes<8, 128, false>());
    return seg ? &seg_v : (grad ? &grad_v : &fwd_v);
}
bool is_sform(const Variant *v) { return v == sform_variant(true, false) || v == sform_variant(true, true) || v == sform_variant(false, false); }

// Precision-escalation kernels of float objects: double arithmetic on float buffers (gradient path).
const Variant *escalation_variant(int M) {
    static const std::vector<Variant> table = {
        make_variant<double, 4, 1, 8, true, 128, 2, float>(),   // M = 4
        make_variant<double, 4, 2, 8, true, 128, 2, float>(),   // M = 8
        make_variant<double, 4, 4, 8, true, 128, 2, float>(),   // M = 16
        make_variant<double, 4, 8, 8, true, 128, 2, float>(),   // M = 32
        make_variant<double, 4, 16, 8, true, 128, 2, float>(),  // M = 64
    };
    for (const Variant &v : table)
        if (v.M == M) return &v;
    return nullptr;
}

// Segment-mode builds of the throughput gradient kernel (segment passes of the two-sweep gradient once
// the segments of all pairs fill the GPU): the fastest lane layout of every M.
const Variant *segment_variant_generic(int M) {
    static const std::vector<Variant> table = {
        make_variant<float, 4, 1, 16, true, 128, 4, float, true>(),   // M = 4
        make_variant<float, 8, 1, 16, true, 128, 3, float, true>(),   // M = 8
        make_variant<float, 16, 1, 8, true, 128, 2, float, true>(),   // M = 16
        make_variant<float, 16, 2, 8, true, 128, 2, float, true>(),   // M = 32
        make_variant<float, 16, 4, 8, true, 128, 2, float, true>(),   // M = 64
    };
    for (const Variant &v : table)
        if (v.M == M) return &v;
    return nullptr;
}

const Variant *segment_variant(const phb_kernel *k) {
    if (k->M == 16 && !k->dbl && k->env_sform) return sform_variant(true, true);
    return segment_variant_generic(k->M);
}

// Store-all gradient kernels (small minibatches, see psmc_kernels.cuh): float only.
struct StoreAllVariant {
    int M, T, MT, NT;
    const void *func;
    size_t smem;
};
template <int MT, int T, int NT, int MINB> StoreAllVariant make_storeall() {
    StoreAllVariant v;
    v.M = MT * T;
    v.T = T;
    v.MT = MT;
    v.NT = NT;
    v.func = reinterpret_cast<const void *>(&phb::psmc_loglik_storeall_kernel<float, MT, T, NT, MINB>);
    v.smem = phb::smem_bytes<float, MT, 8, NT, false>();
    return v;
}
const std::vector<StoreAllVariant> &storeall_variants() {
    static const std::vector<StoreAllVariant> table = {
        // (MT = 8 layouts were measured too: same latency at 255 registers, slower at 168)
        make_storeall<4, 1, 128, 4>(),   // M = 4
        make_storeall<4, 2, 128, 4>(),   // M = 8
        make_storeall<4, 4, 128, 4>(),   // M = 16
        make_storeall<4, 8, 128, 4>(),   // M = 32
        make_storeall<4, 16, 128, 4>(),  // M = 64
    };
    return table;
}

// Lane layouts of the boundary sweeps.  Latency regime: the launch ends with its slowest warp, a warp that has a
// scheduler to itself takes ~1.2 cycles per instruction, two warps on one scheduler take twice as long.  So the
// layout is the widest one (most lanes per pair = fewest instructions per lane and site) that still gives every
// warp a scheduler of its own, i.e. at most one 4-warp CTA per SM; forward and adjoint sweep may differ.
// PHB_SWEEP_TF / PHB_SWEEP_TB force a layout, PHB_SWEEP_LL=0 selects the generic site functions.
struct SweepVariant {
    int M, TF, TB, NT;
    bool ll;
    const void *func;
    size_t smem;
};
template <int M, int TF, int TB, int MINB, bool LL> SweepVariant make_sweep() {
    constexpr int NT = 128;
    return SweepVariant{M, TF, TB, NT, LL, reinterpret_cast<const void *>(&phb::boundary_sweep_kernel<float, M, TF, TB, NT, MINB, LL>),
                        std::max(phb::sweep_smem_bytes<float, M / TF, NT>(), phb::sweep_smem_bytes<float, M / TB, NT>())};
}
const std::vector<SweepVariant> &sweep_variants() {
    // in order of preference for each M (widest first); MINB = 2: up to 255 registers, one CTA per SM is the aim
    static const std::vector<SweepVariant> table = {
        make_sweep<4, 1, 1, 2, false>(),
        make_sweep<8, 2, 2, 2, true>(),    make_sweep<8, 2, 2, 2, false>(),
        make_sweep<16, 4, 4, 2, true>(),   make_sweep<16, 4, 2, 2, true>(),  make_sweep<16, 2, 2, 2, true>(),
        make_sweep<16, 4, 4, 2, false>(),  make_sweep<16, 1, 1, 2, false>(),
        make_sweep<32, 8, 8, 2, true>(),   make_sweep<32, 8, 4, 2, true>(),  make_sweep<32, 4, 4, 2, true>(),
        make_sweep<32, 4, 2, 2, true>(),   make_sweep<32, 2, 2, 2, true>(),  make_sweep<32, 8, 8, 2, false>(),
        make_sweep<64, 16, 16, 2, false>(),
    };
    return table;
}
inline int64_t sweep_ctas(const SweepVariant &v, int64_t n_pairs) {
    const int64_t pf = v.NT / v.TF, pb = v.NT / v.TB;
    return (n_pairs + pf - 1) / pf + (n_pairs + pb - 1) / pb;
}
const SweepVariant *pick_sweep(const phb_kernel *k, int64_t n_pairs) {
    // PHB_SWEEP_LL: 0 = generic site functions only, 1 = low-latency ones wherever they exist (default)
    const bool forced = k->env_sweep_tf > 0;
    const int want_tb = k->env_sweep_tb > 0 ? k->env_sweep_tb : k->env_sweep_tf;
    const bool want_ll = k->env_sweep_ll != 0;
    const SweepVariant *first = nullptr;
    for (const SweepVariant &v : sweep_variants()) {
        if (v.M != k->M) continue;
        if (forced) {
            if (v.TF == k->env_sweep_tf && v.TB == want_tb && v.ll == (k->env_sweep_ll != 0 && v.TF + v.TB > 2)) return &v;
            continue;
        }
        const bool has_ll_form = v.TF + v.TB > 2 && v.M != 64;
        if (has_ll_form && v.ll != want_ll) continue;
        if (!first) first = &v;
        if (sweep_ctas(v, n_pairs) <= k->num_sms) return &v;
    }
    return first;
}

// Parallel-in-time forward evaluation (few, long pairs; see transfer_rows_kernel): float, M <= 16.
struct TransferVariant {
    int M;
    const void *rows_func, *chain_func, *boundaries_func, *product_func, *sharded_boundaries_func;
    size_t smem;
};
template <int M> TransferVariant make_transfer() {
    return TransferVariant{M, reinterpret_cast<const void *>(&phb::transfer_rows_kernel<float, M, 128>),
                           reinterpret_cast<const void *>(&phb::chain_transfer_kernel<float, M>),
                           reinterpret_cast<const void *>(&phb::chain_boundaries_kernel<float, M>),
                           reinterpret_cast<const void *>(&phb::chain_product_kernel<M>),
                           reinterpret_cast<const void *>(&phb::chain_boundaries_sharded_kernel<float, M>),
                           phb::smem_bytes<float, M, 8, 128, false>()};
}
const TransferVariant *transfer_variant(int M) {
    static const std::vector<TransferVariant> table = {make_transfer<4>(), make_transfer<8>(), make_transfer<16>()};
    for (const TransferVariant &v : table)
        if (v.M == M) return &v;
    return nullptr;
}

// Does the store-all scratch (every forward vector + block scales) fit comfortably?  The driver is only
// asked when the buffers have to grow: cudaMemGetInfo costs host milliseconds in a process with many
// allocations, which a 3 ms evaluation must not pay on every call.
bool storeall_scratch_fits(const phb_kernel *k, size_t x_bytes, size_t s_bytes) {
    if (x_bytes <= k->xall.cap && s_bytes <= k->sall.cap) return true;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return x_bytes + s_bytes <= (free_b + k->xall.cap + k->sall.cap) / 2;
}

// Groups (CTAs' worth of pairs) of psmc_loglik_kernel
int64_t chunk_major_groups(const Variant *v, int64_t B, int64_t S) {
    const int64_t pairs_per_cta = v->NT / v->T;
    return (B * S + pairs_per_cta - 1) / pairs_per_cta;
}

const Variant *pick_variant(const phb_kernel *k, bool grad, int64_t n_pairs) {
    const Variant *last = nullptr, *forced = nullptr, *first_fill = nullptr;
    // Lane layouts are ordered by increasing T (fewer lanes per pair = fewer instructions per pair).
    // Measured on B200 (profiles/r01_probe_small_minibatch.log): below ~64 threads per SM the run is
    // bound by the per-site dependency latency (~20 ms per 50 000-bin chunk) and more lanes per pair
    // help a little; above it the thread-per-pair layout is already the fastest.
    const int64_t fill = int64_t(k->num_sms) * 64;
    // tuning knob: PHB_NT=<threads per CTA> picks among variants that differ only in CTA size
    const int want_nt = k->env_nt;
    for (const Variant &v : variants()) {
        if (v.M != k->M || v.dbl != (k->dbl != 0) || v.grad != grad) continue;
        if (last && last->T == v.T && v.NT != want_nt) continue;  // same layout, other CTA size
        if (last && last->T == v.T && v.NT == want_nt) {
            if (forced == last) forced = &v;
            if (first_fill == last) first_fill = &v;
            last = &v;
            continue;
        }
        if (v.T == k->force_T) forced = &v;
        last = &v;
        if (!first_fill && n_pairs * v.T >= fill) first_fill = &v;
    }
    const Variant *chosen = forced ? forced : (first_fill ? first_fill : last);  // (a forced T that is not compiled falls back to auto)
    if (chosen && chosen->M == 16 && chosen->T == 1 && !chosen->dbl && k->env_sform) return sform_variant(grad, false);
    return chosen;
}

int check_handle(const phb_kernel *k) {
    if (!k) return fail(PHB_E_INVALID, "kernel handle is NULL");
    return PHB_OK;
}

// Dispatch.  A call is served by the first of four paths that applies; kNotTaken = "this path does not
// apply, try the next one".
constexpr int kNotTaken = 1;

// The constant bank is ONE per device: evaluations of different kernel objects / streams that use it are ordered
// through a per-device event (not inside a stream capture, where an event recorded outside cannot be waited on: a
// captured step must not run concurrently with another constant-bank evaluation on the same device).
int constant_bank_acquire(phb_kernel *k, cudaStream_t stream, bool release) {
    static cudaEvent_t bank_free[64] = {};
    if (k->device >= 64) return PHB_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) cudaGetLastError();
    if (cs != cudaStreamCaptureStatusNone) return PHB_OK;
    if (release) {
        if (bank_free[k->device] == nullptr) PHB_CUDA(cudaEventCreateWithFlags(&bank_free[k->device], cudaEventDisableTiming));
        PHB_CUDA(cudaEventRecord(bank_free[k->device], stream));
    } else if (bank_free[k->device] != nullptr) {
        PHB_CUDA(cudaStreamWaitEvent(stream, bank_free[k->device], 0));
    }
    return PHB_OK;
}

// Segment transfer operators (transfer_rows_kernel) of the segments [ta.seg_first, ta.seg_first + ta.n_seg_local)
int launch_transfer_rows(phb_kernel *k, const TransferVariant *tv, const phb::TransferArgs &ta, cudaStream_t stream) {
    const phb::KernelArgs &a = ta.k;
    const int M = k->M;
    if (k->occupancy.find(tv->rows_func) == k->occupancy.end()) {
        PHB_CUDA(cudaFuncSetAttribute(tv->rows_func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tv->smem)));
        k->occupancy.emplace(tv->rows_func, 1);
    }
    if (k->dry) return PHB_OK;
    const int64_t n_virtual = a.B * a.S * ta.n_seg_local * M;
    phb::TransferArgs copy = ta;
    void *kargs[] = {&copy};
    PHB_CUDA(cudaLaunchKernel(tv->rows_func, dim3(unsigned((n_virtual + 127) / 128)), dim3(128), kargs, tv->smem, stream));
    k->launches += 1;
    return PHB_OK;
}

// Does the warm-up term (a.warm_len sites, see KernelArgs) ride along with the segment passes?  Its groups must be
// RESIDENT next to the real ones: as a second round they would add 500 dependent sites at the throughput kernel's
// 0.8 us per site, more than the 0.14 ms launch they replace.  With many segments one is given up for that (each
// of the others grows by 1 / n_seg); with few (S = 5: 14) the separate launch stays.
static bool plan_fused_warmup(const phb_kernel *k, const phb::KernelArgs &a, const Variant *v, int64_t seg_ctas, int64_t resident,
                              int64_t *n_seg) {
    if (a.warm_len <= 0 || k->env_fuse_warmup == 0 || a.skip_flag != nullptr || a.s_list != nullptr || a.out_mode != 0 || is_sform(v))
        return false;
    if ((*n_seg + 1) * seg_ctas <= resident) return true;
    if (*n_seg >= 20 && *n_seg * seg_ctas <= resident) {
        *n_seg -= 1;
        return true;
    }
    return false;
}

// (1) Gradient of FEW pairs (the reference's default minibatch for one genome is a single chunk: 500
// pairs): parallel in time.  Segment transfer operators give the forward / adjoint vectors at the
// segment boundaries (chain_boundaries_kernel), after which the segments are independent short
// chunks for the store-all kernel.  Worth it while the operators (M x the forward work, at full
// throughput) cost less than the dependent site steps they remove: pairs * M below 0.3 of the
// resident threads (measured, profiles/r01_probe_parallel_in_time.log: B = 500, L = 50 000, S = 1 / 2 / 3
// chunks 3.4 / 6.5 / 13.6 ms against 13.6 ms sequential).
int try_parallel_in_time_gradient(phb_kernel *k, const phb::KernelArgs &a, cudaStream_t stream, int pit_mode) {
    const int64_t n_pairs = a.B * a.S;
    const TransferVariant *tv = transfer_variant(k->M);
    const Variant *gv = segment_variant(k);  // segment passes: throughput kernel in SEG mode
    if (!tv || !gv) return kNotTaken;
    const int M = k->M;
    const int64_t capacity = int64_t(k->num_sms) * 384;  // resident threads of the row kernel
    // (break-even against the two sweeps, 3.3 ms at 50 500 sites: 1.8 ms of operator work per 500 pairs at M = 16,
    // profiles/r02_probe_sweep_layouts_v7.log: 1 000 pairs take 5.0 ms through the operators, 4.4 ms through the sweeps)
    if (pit_mode != 1 && n_pairs * M * 4 > capacity) return kNotTaken;
    int occ = 0;
    {
        auto it = k->occupancy.find(gv->func);
        if (it == k->occupancy.end()) {
            PHB_CUDA(cudaFuncSetAttribute(gv->func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gv->smem)));
            PHB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gv->func, gv->NT, gv->smem));
            k->occupancy.emplace(gv->func, occ);
        } else {
            occ = it->second;
        }
    }
    if (occ < 1) return kNotTaken;
    const int64_t resident = int64_t(occ) * k->num_sms;
    const int64_t seg_ctas = chunk_major_groups(gv, a.B, a.S);  // groups per segment
    // (segments of >= 512 sites: at S = 1 that is 73 segments + the warm-up term = every resident group slot; with
    // the 1024 of round 1 a third of the slots stayed empty: 2.88 -> 2.67 ms)
    const int64_t min_seg = pit_mode == 1 ? 64 : 512;
    // as many segments as keep every group of the segment passes resident at once
    int64_t n_seg = std::min(resident / seg_ctas, a.L / min_seg);
    if (pit_mode == 1) n_seg = std::max<int64_t>(n_seg, std::min<int64_t>(3, a.L / min_seg));
    if (k->env_pit_segments > 0) n_seg = std::min<int64_t>(k->env_pit_segments, a.L / 64);  // experiments
    if (n_seg < 3) return kNotTaken;
    bool warm = plan_fused_warmup(k, a, gv, seg_ctas, resident, &n_seg);
    const int64_t seg_len = ((a.L + n_seg - 1) / n_seg + 15) / 16 * 16;
    n_seg = (a.L + seg_len - 1) / seg_len;
    warm = warm && a.warm_len <= seg_len;
    const int64_t n_slots = n_seg + (warm ? 1 : 0);  // partial gradients per pair
    const int64_t n_rows_virtual = n_pairs * n_seg * M;
    const int64_t n_groups = seg_ctas * n_slots;
    const int64_t grid = std::min<int64_t>(n_groups, resident);
    int rc;
    if ((rc = k->transfer_rows.reserve(size_t(n_rows_virtual) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->transfer_log.reserve(size_t(n_rows_virtual) * sizeof(double))) != PHB_OK) return rc;
    if ((rc = k->bnd_alpha.reserve(size_t(n_pairs) * (n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->bnd_beta.reserve(size_t(n_pairs) * (n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->seg_dlog.reserve(size_t(n_pairs) * n_slots * 7 * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->ckpt.reserve(size_t(grid) * (gv->NT / 32) * size_t(gv->ckpt_bytes_per_warp(seg_len)))) != PHB_OK) return rc;
    if ((rc = k->gacc.reserve(size_t(grid) * gv->NT * 6 * (gv->M / gv->T) * sizeof(double))) != PHB_OK) return rc;
    if (warm && (rc = k->warm_ll.reserve(size_t(n_pairs) * sizeof(double))) != PHB_OK) return rc;
    k->warm_fused = warm;
    phb::TransferArgs ta{};
    ta.k = a;
    ta.k.err_flag = k->d_err;
    ta.n_seg = n_seg;
    ta.seg_len = seg_len;
    ta.rows = static_cast<float *>(k->transfer_rows.ptr);
    ta.row_log2 = static_cast<double *>(k->transfer_log.ptr);
    ta.segs_per_slot = n_seg;  // one process: a single slot with all segments
    ta.n_seg_local = n_seg;
    if ((rc = launch_transfer_rows(k, tv, ta, stream)) != PHB_OK) return rc;  // (sizes its own staging in a dry run)
    if (k->dry) return PHB_OK;
    void *bnd_a = k->bnd_alpha.ptr, *bnd_b = k->bnd_beta.ptr;
    {
        void *bargs[] = {&ta, &bnd_a, &bnd_b};
        PHB_CUDA(cudaLaunchKernel(tv->boundaries_func, dim3(unsigned((n_pairs * M + 127) / 128)), dim3(128), bargs, 0, stream));
    }
    phb::KernelArgs sa = a;
    sa.err_flag = k->d_err;
    sa.ckpt = k->ckpt.ptr;
    sa.gacc = static_cast<double *>(k->gacc.ptr);
    sa.seg_count = n_seg;
    sa.seg_len = seg_len;
    sa.bnd_alpha = bnd_a;
    sa.bnd_beta = bnd_b;
    sa.seg_dlog = k->seg_dlog.ptr;
    sa.seg_ctas = seg_ctas;
    sa.n_groups = n_groups;
    sa.warm_len = warm ? a.warm_len : 0;
    sa.warm_ll = warm ? static_cast<double *>(k->warm_ll.ptr) : nullptr;
    {
        void *kargs[] = {&sa};
        PHB_CUDA(cudaLaunchKernel(gv->func, dim3(unsigned(grid)), dim3(gv->NT), kargs, gv->smem, stream));
    }
    const int64_t n_out = n_pairs * 7 * M;
    phb::sum_segments_kernel<float><<<unsigned((n_out + 255) / 256), 256, 0, stream>>>(
        static_cast<const float *>(k->seg_dlog.ptr), n_pairs, n_seg, M, static_cast<float *>(a.dlog), a.out_mode, sa);
    PHB_CUDA(cudaGetLastError());
    k->launches += 3;
    snprintf(k->last_name, sizeof k->last_name, "transfer_rows_kernel<float,M=%d> + psmc_loglik_kernel<SEG,MT=%d,T=%d> x %lld segments%s", M,
             gv->M / gv->T, gv->T, (long long)n_seg, warm ? " + warm-up term" : "");
    return PHB_OK;
}

// (1b) Gradient of a small minibatch with too many pairs for the operators (the reference's S = 5: 2 500
// pairs): the boundary vectors come from two sequential sweeps side by side (boundary_sweep_kernel), then
// the same segment passes.  Any M.  Latency regime only (the bound of the store-all path).
int try_two_sweep_gradient(phb_kernel *k, const phb::KernelArgs &a, cudaStream_t stream, int pit_mode) {
    const int64_t n_pairs = a.B * a.S;
    const StoreAllVariant *sv = nullptr;
    for (const StoreAllVariant &c : storeall_variants())
        if (c.M == k->M) sv = &c;
    const Variant *tv = segment_variant(k);
    if (!sv || !tv) return kNotTaken;
    if (pit_mode != 2 && n_pairs * sv->T > int64_t(k->num_sms) * 330) return kNotTaken;
    const int M = k->M;
    // Segment passes on the throughput kernel (thread per pair at M = 16: 1.7 x the throughput of the
    // store-all layout): as many segments as keep every group resident at once.
    int occ = 0;
    {
        auto it = k->occupancy.find(tv->func);
        if (it == k->occupancy.end()) {
            PHB_CUDA(cudaFuncSetAttribute(tv->func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tv->smem)));
            PHB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tv->func, tv->NT, tv->smem));
            k->occupancy.emplace(tv->func, occ);
        } else {
            occ = it->second;
        }
    }
    if (occ < 1) return kNotTaken;
    const int64_t resident = int64_t(occ) * k->num_sms;
    const int64_t seg_ctas = chunk_major_groups(tv, a.B, a.S);  // groups per segment
    // (short rows - the 500-site warm-up launch of the whole-term entries - stay with the store-all kernel: cut into
    // segments of 64+ sites they take 0.09 instead of 0.14 ms, which did not show in the step time)
    const int64_t min_seg = pit_mode == 2 ? 64 : 1024;
    int64_t n_seg = std::min(std::max<int64_t>(resident / seg_ctas, 2), a.L / min_seg);
    if (k->env_pit_segments > 0) n_seg = std::min<int64_t>(k->env_pit_segments, a.L / 64);  // experiments
    // measured at B = 500, L = 50 000, M = 16 (profiles/r01_probe_parallel_in_time.log)
    if (n_seg < (pit_mode == 2 ? 2 : 6)) return kNotTaken;
    bool warm = plan_fused_warmup(k, a, tv, seg_ctas, resident, &n_seg);
    const int64_t seg_len = ((a.L + n_seg - 1) / n_seg + 15) / 16 * 16;
    n_seg = (a.L + seg_len - 1) / seg_len;
    if (n_seg < 2) return kNotTaken;
    warm = warm && a.warm_len <= seg_len;
    const int64_t n_slots = n_seg + (warm ? 1 : 0);  // partial gradients per pair
    const int64_t n_groups = seg_ctas * n_slots;
    const int64_t grid = std::min<int64_t>(n_groups, resident);
    const SweepVariant *sw = pick_sweep(k, n_pairs);
    if (!sw || a.L >= (int64_t(1) << 31) - 16) return kNotTaken;  // (the sweeps count sites in 32 bits)
    const void *sweep_func = sw->func;
    const size_t sweep_smem = sw->smem;
    const int64_t fwd_ctas = (n_pairs + sw->NT / sw->TF - 1) / (sw->NT / sw->TF);
    int rc;
    if ((rc = k->bnd_alpha.reserve(size_t(n_pairs) * (n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->bnd_beta.reserve(size_t(n_pairs) * (n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->seg_dlog.reserve(size_t(n_pairs) * n_slots * 7 * M * sizeof(float))) != PHB_OK) return rc;
    if (warm && (rc = k->warm_ll.reserve(size_t(n_pairs) * sizeof(double))) != PHB_OK) return rc;
    if ((rc = k->ckpt.reserve(size_t(grid) * (tv->NT / 32) * size_t(tv->ckpt_bytes_per_warp(seg_len)))) != PHB_OK) return rc;
    if ((rc = k->gacc.reserve(size_t(grid) * tv->NT * 6 * (tv->M / tv->T) * sizeof(double))) != PHB_OK) return rc;
    // checkpoints of the segment passes, written by the forward sweep (64 B per pair and 8 sites: 1 GB for 2 500
    // pairs of 50 500 sites); PHB_SWEEP_CKPT=0 or a budget of 8 GB leaves the segment passes to make their own
    const int64_t ck_count = (a.L + tv->K - 1) / tv->K + 1;
    const size_t ck_bytes = size_t(n_pairs) * ck_count * M * sizeof(float);
    const bool ext_ck = k->env_sweep_ckpt != 0 && !is_sform(tv) && ck_bytes <= (size_t(8) << 30);
    if (ext_ck && (rc = k->sweep_ckpt.reserve(ck_bytes)) != PHB_OK) return rc;
    if (k->occupancy.find(sweep_func) == k->occupancy.end()) {
        PHB_CUDA(cudaFuncSetAttribute(sweep_func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sweep_smem)));
        k->occupancy.emplace(sweep_func, 1);
    }
    k->warm_fused = warm;
    if (k->dry) return PHB_OK;
    phb::KernelArgs sa = a;
    sa.err_flag = k->d_err;
    sa.ckpt = k->ckpt.ptr;
    sa.gacc = static_cast<double *>(k->gacc.ptr);
    sa.seg_count = n_seg;
    sa.seg_len = seg_len;
    sa.bnd_alpha = k->bnd_alpha.ptr;
    sa.bnd_beta = k->bnd_beta.ptr;
    sa.seg_dlog = k->seg_dlog.ptr;
    sa.seg_ctas = seg_ctas;
    sa.n_groups = n_groups;
    sa.sweep_fwd_ctas = fwd_ctas;
    sa.ext_ck = ext_ck ? k->sweep_ckpt.ptr : nullptr;
    sa.ext_ck_count = ck_count;
    sa.ext_ck_blocks = tv->K / 4;
    sa.warm_len = warm ? a.warm_len : 0;
    sa.warm_ll = warm ? static_cast<double *>(k->warm_ll.ptr) : nullptr;
    void *kargs[] = {&sa};
    PHB_CUDA(cudaLaunchKernel(sweep_func, dim3(unsigned(sweep_ctas(*sw, n_pairs))), dim3(sw->NT), kargs, sweep_smem, stream));
    PHB_CUDA(cudaLaunchKernel(tv->func, dim3(unsigned(grid)), dim3(tv->NT), kargs, tv->smem, stream));
    const int64_t n_out = n_pairs * 7 * M;
    phb::sum_segments_kernel<float><<<unsigned((n_out + 255) / 256), 256, 0, stream>>>(
        static_cast<const float *>(k->seg_dlog.ptr), n_pairs, n_seg, M, static_cast<float *>(a.dlog), a.out_mode, sa);
    PHB_CUDA(cudaGetLastError());
    k->launches += 3;
    snprintf(k->last_name, sizeof k->last_name, "boundary_sweep_kernel<float,TF=%d,TB=%d%s> + psmc_loglik_kernel<SEG,MT=%d,T=%d> x %lld segments%s",
             sw->TF, sw->TB, sw->ll ? ",LL" : "", tv->M / tv->T, tv->T, (long long)n_seg, warm ? " + warm-up term" : "");
    return PHB_OK;
}

// (2) Gradient of a small minibatch: the store-all kernel (two dependent passes instead of three).
int try_store_all(phb_kernel *k, phb::KernelArgs a, cudaStream_t stream, int sa_mode) {
    const int64_t n_pairs = a.B * a.S;  // upper bound when a sub-list is given
    for (const StoreAllVariant &sv : storeall_variants()) {
        if (sv.M != k->M) continue;
        const int pairs_per_cta = sv.NT / sv.T;
        const int64_t grid = (n_pairs + pairs_per_cta - 1) / pairs_per_cta;
        const int64_t warps = grid * (sv.NT / 32);
        const size_t x_bytes = size_t(warps) * size_t(a.L) * sv.MT * 32 * sizeof(float);
        const size_t s_bytes = size_t(warps) * size_t((a.L + phb::kNorm - 1) / phb::kNorm) * 32 * sizeof(float);
        bool use = sa_mode == 1;
        if (sa_mode < 0 && n_pairs * sv.T <= int64_t(k->num_sms) * 330) {
            // latency-bound regime and the scratch fits comfortably.  Measured on B200 at M = 16,
            // B = 500 (profiles/r01_probe_small_minibatch.log): store-all wins up to ~12 000 pairs
            // (S = 24: 23.4 vs 27.2 ms) and loses from ~16 000 pairs on.
            use = storeall_scratch_fits(k, x_bytes, s_bytes);
        }
        if (!use) break;
        int rc;
        if ((rc = k->xall.reserve(x_bytes)) != PHB_OK) return rc;
        if ((rc = k->sall.reserve(s_bytes)) != PHB_OK) return rc;
        if ((rc = k->gacc.reserve(size_t(grid) * sv.NT * 6 * sv.MT * sizeof(double))) != PHB_OK) return rc;
        a.xall = k->xall.ptr;
        a.sall = k->sall.ptr;
        a.gacc = static_cast<double *>(k->gacc.ptr);
        a.n_groups = grid;
        a.err_flag = k->d_err;
        if (k->occupancy.find(sv.func) == k->occupancy.end()) {
            PHB_CUDA(cudaFuncSetAttribute(sv.func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sv.smem)));
            k->occupancy.emplace(sv.func, 1);
        }
        if (k->dry) return PHB_OK;
        void *kargs[] = {&a};
        PHB_CUDA(cudaLaunchKernel(sv.func, dim3(unsigned(grid)), dim3(sv.NT), kargs, sv.smem, stream));
        k->launches += 1;
        snprintf(k->last_name, sizeof k->last_name, "psmc_loglik_storeall_kernel<float,MT=%d,T=%d,NT=%d>", sv.MT, sv.T, sv.NT);
        return PHB_OK;
    }

    return kNotTaken;
}

// (3) Forward-only, few long pairs: segment transfer operators chained afterwards.  The sequential
// kernel needs ~80 ns per site whatever the number of pairs (measured, profiles/r01_elpd_shape_probe.log);
// M times the work at full throughput is faster while pairs * M is below a quarter of the
// resident threads.
int try_parallel_in_time_forward(phb_kernel *k, const phb::KernelArgs &a, cudaStream_t stream, int pit_mode) {
    const int64_t n_pairs = a.B * a.S;
    if (const TransferVariant *tv = transfer_variant(k->M)) {
        const int64_t capacity = int64_t(k->num_sms) * 384;  // resident threads of the row kernel
        const int64_t min_seg = pit_mode == 1 ? 64 : 4096;   // sites; shorter segments are all overhead
        int64_t n_seg = std::min(capacity / (n_pairs * tv->M), a.L / min_seg);
        if (pit_mode == 1) n_seg = std::max<int64_t>(n_seg, std::min<int64_t>(3, a.L / min_seg));
        // (measured at M = 16, L = 2.5 M: 500 pairs -> 7 segments 91 ms vs 209 ms sequential; 1000 pairs ->
        // 3 segments 213 ms, no gain any more; profiles/r01_elpd_shape_probe.log)
        if (n_seg >= (pit_mode == 1 ? 3 : 4)) {
            int64_t seg_len = ((a.L + n_seg - 1) / n_seg + 15) / 16 * 16;
            n_seg = (a.L + seg_len - 1) / seg_len;
            const int64_t n_virtual = n_pairs * n_seg * tv->M;
            int rc;
            if ((rc = k->transfer_rows.reserve(size_t(n_virtual) * tv->M * sizeof(float))) != PHB_OK) return rc;
            if ((rc = k->transfer_log.reserve(size_t(n_virtual) * sizeof(double))) != PHB_OK) return rc;
            phb::TransferArgs ta{};
            ta.k = a;
            ta.k.err_flag = k->d_err;
            ta.n_seg = n_seg;
            ta.seg_len = seg_len;
            ta.rows = static_cast<float *>(k->transfer_rows.ptr);
            ta.row_log2 = static_cast<double *>(k->transfer_log.ptr);
            ta.segs_per_slot = n_seg;
            ta.n_seg_local = n_seg;
            if ((rc = launch_transfer_rows(k, tv, ta, stream)) != PHB_OK) return rc;
            if (k->dry) return PHB_OK;
            void *kargs[] = {&ta};
            PHB_CUDA(cudaLaunchKernel(tv->chain_func, dim3(unsigned((n_pairs + 63) / 64)), dim3(64), kargs, 0, stream));
            k->launches += 1;
            snprintf(k->last_name, sizeof k->last_name, "transfer_rows_kernel<float,M=%d> x %lld segments + chain_transfer_kernel",
                     tv->M, (long long)n_seg);
            return PHB_OK;
        }
    }

    return kNotTaken;
}

// ---- time-axis sharding of a SMALL minibatch over several processes (one per GPU) ------------------------
// With S < world chunks there is nothing to shard on the chunk axis (the reference splits S <= 5 indices over
// its devices and leaves the others idle, gpu.py:398-400).  The segments of the parallel-in-time gradient ARE
// independent once the boundary vectors exist: every process computes the transfer operators of its own
// slice of the segments and multiplies them into ONE operator per pair (float64), a small all-gather makes
// those `world` operators per pair visible everywhere, every process chains them to the vectors entering and
// leaving its slice, chains its own operators to the boundary vectors inside the slice and runs the gradient
// passes over the slice; the partial per-particle sums join the all-reduce that the chunk-sharded path uses.
struct ShardPlan {
    int64_t n_seg = 0, seg_len = 0, seg_ctas = 0, per_rank = 0;
    size_t rows_bytes = 0, slot_bytes = 0;  // per process
};
bool shard_plan(phb_kernel *k, int64_t B, int64_t S, int64_t L, int world, ShardPlan &p) {
    p = ShardPlan{};
    const TransferVariant *tv = transfer_variant(k->M);
    const Variant *gv = segment_variant(k);
    if (!tv || !gv || k->dbl || world < 2) return false;
    const int64_t n_pairs = B * S;
    const int M = k->M;
    const int64_t capacity = int64_t(k->num_sms) * 384 * world;  // resident threads of the row kernel, all GPUs
    if (n_pairs * M * 10 > capacity * 3) return false;           // the single-GPU rule, scaled
    int occ = 0;
    auto it = k->occupancy.find(gv->func);
    if (it == k->occupancy.end()) {
        if (cudaFuncSetAttribute(gv->func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gv->smem)) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gv->func, gv->NT, gv->smem) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        k->occupancy.emplace(gv->func, occ);
    } else {
        occ = it->second;
    }
    if (occ < 1) return false;
    p.seg_ctas = chunk_major_groups(gv, B, S);
    // as many segments per process as keep its gradient groups resident at once; segments >= 512 sites
    int64_t per_rank = std::min<int64_t>(int64_t(occ) * k->num_sms / p.seg_ctas, (L / 512) / world);
    if (k->env_pit_segments > 0) per_rank = std::max<int64_t>(1, k->env_pit_segments / world);
    if (per_rank < 1) return false;
    int64_t n_seg = per_rank * world;
    p.seg_len = ((L + n_seg - 1) / n_seg + 15) / 16 * 16;
    n_seg = (L + p.seg_len - 1) / p.seg_len;
    p.per_rank = (n_seg + world - 1) / world;  // the last process may hold fewer
    p.n_seg = n_seg;
    if (n_seg < 2 * world && n_seg < 3) return false;
    // what a process contributes to the all-gather: ONE operator per pair, the product of its segment operators
    p.rows_bytes = size_t(n_pairs) * M * M * sizeof(float);
    p.slot_bytes = p.rows_bytes + size_t(n_pairs) * M * sizeof(double);
    return true;
}

// (4a) FORWARD-ONLY evaluation of large minibatches at M = 16 whose parameter rows are shared by the chunks of a
// particle: the same recursion in the one-particle-per-warp layout of psmc_uniform.cuh (+21 % measured), scored
// in batches of kUniformSlots particles (the constant bank holds that many).
int try_uniform_throughput(phb_kernel *k, const phb::KernelArgs &a, bool grad, cudaStream_t stream) {
    if (grad || k->dbl || k->M != phb::kUniformM || !k->env_uniform || a.pstride_s != 0 || a.S < 64 || a.alpha_out != nullptr) return kNotTaken;
    {
        // (inside a stream capture the register-parameter kernel is used: the launches below are several, and a
        // captured forward-only evaluation of a large minibatch is not a case worth the constant-bank ordering)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) cudaGetLastError();
        if (cs != cudaStreamCaptureStatusNone) return kNotTaken;
    }
    constexpr int K = 8;
    const void *func = reinterpret_cast<const void *>(&phb::psmc_uniform_forward_kernel<K>);
    const size_t smem = phb::uniform_smem_bytes();
    int occ = 0;
    {
        auto it = k->occupancy.find(func);
        if (it == k->occupancy.end()) {
            PHB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, 32, smem));
            k->occupancy.emplace(func, occ);
        } else {
            occ = it->second;
        }
    }
    if (occ < 1) return kNotTaken;
    const int64_t resident = int64_t(occ) * k->num_sms;
    const int64_t wpp = (a.S + 31) / 32;  // warp tasks per particle (upper bound with a sub-list)
    int rc;
    if ((rc = k->uniform_stage.reserve(size_t(a.B) * sizeof(phb::UniformParams))) != PHB_OK) return rc;
    if (k->dry) return PHB_OK;
    phb::UniformParams *stage = static_cast<phb::UniformParams *>(k->uniform_stage.ptr);
    {
        const int64_t n = a.B * 2 * phb::kUniformM;
        phb::pack_uniform_params_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(static_cast<const float *>(a.params6), a.pstride_b, a.B, stage);
        PHB_CUDA(cudaGetLastError());
        k->launches += 1;
    }
    if ((rc = constant_bank_acquire(k, stream, false)) != PHB_OK) return rc;
    phb::UniformArgs ua{};
    ua.k = a;
    ua.k.err_flag = k->d_err;
    for (int64_t first = 0; first < a.B; first += phb::kUniformSlots) {
        const int64_t nb = std::min<int64_t>(phb::kUniformSlots, a.B - first);
        PHB_CUDA(cudaMemcpyToSymbolAsync(phb::c_uniform_params, stage + first, size_t(nb) * sizeof(phb::UniformParams), 0,
                                         cudaMemcpyDeviceToDevice, stream));
        ua.first_b = first;
        ua.n_b = nb;
        ua.n_tasks = nb * wpp;
        const int64_t grid = std::min<int64_t>(ua.n_tasks, resident);
        void *kargs[] = {&ua};
        PHB_CUDA(cudaLaunchKernel(func, dim3(unsigned(grid)), dim3(32), kargs, smem, stream));
        k->launches += 1;
    }
    if ((rc = constant_bank_acquire(k, stream, true)) != PHB_OK) return rc;
    snprintf(k->last_name, sizeof k->last_name, "psmc_uniform_forward_kernel<float,M=16,K=%d,fwd> (one particle per warp)", K);
    return PHB_OK;
}

// (4) The throughput kernel: persistent grid, checkpoints + recompute for the gradient.  `fixed` pins the
// kernel variant (precision escalation), nullptr lets pick_variant choose.
int launch_throughput_kernel(phb_kernel *k, phb::KernelArgs a, bool grad, cudaStream_t stream, const Variant *fixed) {
    const int64_t n_pairs = a.B * a.S;  // upper bound when a sub-list is given
    if (!fixed && k->force_T == 0) {
        const int rc = try_uniform_throughput(k, a, grad, stream);
        if (rc != kNotTaken) return rc;
    }
    const Variant *v = fixed ? fixed : pick_variant(k, grad, n_pairs);
    if (!v) return fail(PHB_E_INVALID, "no kernel variant for M=%d, threads_per_pair=%d", k->M, k->force_T);
    a.n_groups = chunk_major_groups(v, a.B, a.S);
    const size_t smem = v->smem;
    int occ = 0;
    {
        auto it = k->occupancy.find(v->func);
        if (it == k->occupancy.end()) {
            PHB_CUDA(cudaFuncSetAttribute(v->func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            PHB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v->func, v->NT, smem));
            k->occupancy.emplace(v->func, occ);
        } else {
            occ = it->second;
        }
    }
    if (occ < 1) return fail(PHB_E_CUDA, "kernel does not fit on an SM (smem %zu bytes)", smem);
    const int64_t grid = std::min<int64_t>(a.n_groups, int64_t(occ) * k->num_sms);
    if (grad) {
        const size_t need = size_t(grid) * (v->NT / 32) * size_t(v->ckpt_bytes_per_warp(a.L));
        int rc = k->ckpt.reserve(need);
        if (rc != PHB_OK) return rc;
        a.ckpt = k->ckpt.ptr;
        const int mt = v->M / v->T;
        rc = k->gacc.reserve(size_t(grid) * v->NT * 6 * mt * sizeof(double));
        if (rc != PHB_OK) return rc;
        a.gacc = static_cast<double *>(k->gacc.ptr);
    }
    a.err_flag = k->d_err;
    if (k->dry) return PHB_OK;
    void *kargs[] = {&a};
    PHB_CUDA(cudaLaunchKernel(v->func, dim3(unsigned(grid)), dim3(v->NT), kargs, smem, stream));
    k->launches += 1;
    if (!fixed)
        snprintf(k->last_name, sizeof k->last_name, "%s<%s,MT=%d,T=%d,K=%d,%s,NT=%d>", is_sform(v) ? "psmc_sform_kernel" : "psmc_loglik_kernel",
                 v->dbl ? "double" : "float", v->M / v->T, v->T, v->K, v->grad ? "grad" : "fwd", v->NT);
    return PHB_OK;
}

// One evaluation on `stream` over the whole minibatch or over the sub-list in `a`.  pit_only: try the
// parallel-in-time gradient paths only (they honour a.skip_flag) and report kNotTaken otherwise.
int launch_one(phb_kernel *k, phb::KernelArgs a, bool grad, cudaStream_t stream, const Variant *fixed, bool pit_only = false) {
    // experiment knobs PHB_STORE_ALL / PHB_PARALLEL_IN_TIME (read at creation) override the modes set through the API
    const int sa_mode = k->env_store_all != -2 ? k->env_store_all : k->store_all_mode;
    const int pit_mode = k->env_pit != -2 ? k->env_pit : k->parallel_in_time;
    const bool free_choice = !fixed && !k->dbl && k->force_T == 0;
    // (a store-all kernel asked for by name wins over the automatic choice of a parallel-in-time path)
    const bool pit_allowed = pit_mode != 0 && sa_mode != 0 && !(sa_mode == 1 && pit_mode != 1 && pit_mode != 2);
    int rc = kNotTaken;
    if (free_choice && grad && pit_allowed && pit_mode != 2 && a.s_list == nullptr)
        rc = try_parallel_in_time_gradient(k, a, stream, pit_mode);
    if (rc == kNotTaken && free_choice && grad && pit_allowed && pit_mode != 1 && a.s_list == nullptr)
        rc = try_two_sweep_gradient(k, a, stream, pit_mode);
    if (pit_only) return rc;
    if (rc == kNotTaken && free_choice && grad && sa_mode != 0) rc = try_store_all(k, a, stream, sa_mode);
    if (rc == kNotTaken && free_choice && !grad && pit_mode != 0 && a.s_list == nullptr)
        rc = try_parallel_in_time_forward(k, a, stream, pit_mode);
    if (rc == kNotTaken) rc = launch_throughput_kernel(k, a, grad, stream, fixed);
    return rc;
}


// Launch on `stream`; all pointers in `a` are device pointers except the ones filled in here.
int launch(phb_kernel *k, phb::KernelArgs a, bool grad, cudaStream_t stream) {
    if (a.B * a.S == 0) return PHB_OK;
    const Variant *esc = (grad && !k->dbl && k->escalate && k->n_flagged > 0) ? escalation_variant(k->M) : nullptr;
    // the timing events are skipped inside a stream capture (a captured event cannot be queried) and in a dry run
    bool timing = !k->dry;
    if (timing) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) cudaGetLastError();
        timing = cs == cudaStreamCaptureStatusNone;
    }
    if (timing) PHB_CUDA(cudaEventRecord(k->ev0, stream));
    int rc;
    if (!esc) {
        rc = launch_one(k, a, grad, stream, nullptr);
    } else {
        // split the minibatch on the device into un-flagged / flagged chunks; one launch for each
        if (a.S > int64_t(INT32_MAX)) return fail(PHB_E_INVALID, "minibatch of %lld chunks is too large", (long long)a.S);
        if ((rc = k->split.reserve((size_t(2) * a.S + 2) * sizeof(int32_t))) != PHB_OK) return rc;
        int32_t *lists = static_cast<int32_t *>(k->split.ptr);
        int32_t *counts = lists + 2 * a.S;
        if (!k->dry) {
            phb::split_minibatch_kernel<<<1, 1024, 0, stream>>>(a.inds, a.S, k->d_rowflag, k->N, lists, counts);
            PHB_CUDA(cudaGetLastError());
            k->launches += 1;
        }
        // The decision is per CALL, not per object: a small minibatch goes through the parallel-in-time paths
        // whether or not it holds a marked row (they score everything and leave the outputs of marked rows
        // to the double launch below); larger ones run the ordinary kernels over the un-marked sub-list.
        phb::KernelArgs whole = a;
        whole.skip_flag = k->d_rowflag;
        rc = launch_one(k, whole, grad, stream, nullptr, /*pit_only=*/true);
        phb::KernelArgs part = a;
        part.s_list = lists;
        part.s_count = counts;
        if (rc == kNotTaken) rc = launch_one(k, part, grad, stream, nullptr);
        if (rc == PHB_OK) {
            part.s_list = lists + a.S;
            part.s_count = counts + 1;
            rc = launch_one(k, part, grad, stream, esc);
        }
    }
    if (rc != PHB_OK) return rc;
    if (timing) {
        PHB_CUDA(cudaEventRecord(k->ev1, stream));
        k->timed = true;
    } else if (!k->dry) {
        k->timed = false;
    }
    return PHB_OK;
}

template <typename F> bool all_finite(const F *p, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(p[i])) return false;
    return true;
}

}  // namespace

extern "C" {

int phb_abi_version(void) { return 3; }

const char *phb_last_error(void) { return g_err.c_str(); }

int phb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PHB_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

// Shared part of the constructors: device checks, handle, data buffer [N, pitch], stream, events.
static int new_kernel(int M, int64_t N, int64_t L, int double_precision, int device, phb_kernel **out) {
    *out = nullptr;
    if (!(M == 4 || M == 8 || M == 16 || M == 32 || M == 64))
        return fail(PHB_E_INVALID, "M=%d is not supported (4, 8, 16, 32 or 64)", M);
    int ndev = 0;
    PHB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PHB_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    PHB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PHB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PHB_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    phb_kernel *k = new (std::nothrow) phb_kernel();
    if (!k) return fail(PHB_E_NOMEM, "host allocation failed");
    k->M = M;
    k->dbl = double_precision ? 1 : 0;
    k->device = device;
    k->N = N;
    k->L = L;
    k->pitch = (L + 15) / 16 * 16;
    k->num_sms = prop.multiProcessorCount;
    if (const char *v = getenv("PHB_NT")) k->env_nt = atoi(v);
    if (const char *v = getenv("PHB_STORE_ALL")) k->env_store_all = atoi(v);
    if (const char *v = getenv("PHB_PARALLEL_IN_TIME")) k->env_pit = atoi(v);
    if (const char *v = getenv("PHB_PIT_SEGMENTS")) k->env_pit_segments = atoi(v);
    if (const char *v = getenv("PHB_SWEEP_TF")) k->env_sweep_tf = atoi(v);
    if (const char *v = getenv("PHB_SWEEP_TB")) k->env_sweep_tb = atoi(v);
    if (const char *v = getenv("PHB_SWEEP_LL")) k->env_sweep_ll = atoi(v);
    if (const char *v = getenv("PHB_SWEEP_CKPT")) k->env_sweep_ckpt = atoi(v);
    if (const char *v = getenv("PHB_FUSE_WARMUP")) k->env_fuse_warmup = atoi(v);
    if (const char *v = getenv("PHB_UNIFORM")) k->env_uniform = atoi(v);
    if (const char *v = getenv("PHB_SFORM")) k->env_sform = atoi(v);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&k->d_data), size_t(N) * size_t(k->pitch));
    if (e != cudaSuccess) {
        cudaGetLastError();
        phb_destroy(k);
        return fail(PHB_E_NOMEM, "While trying to allocate %lld bytes on GPU: %s", (long long)(N * k->pitch), cudaGetErrorString(e));
    }
    for (DeviceBuffer *b : {&k->params, &k->inds, &k->ll, &k->dlog, &k->ckpt, &k->gacc, &k->xall, &k->sall, &k->split, &k->term_params,
                            &k->term_ll, &k->term_dlog, &k->term_sums, &k->term_io, &k->transfer_rows, &k->transfer_log, &k->bnd_alpha,
                            &k->bnd_beta, &k->seg_dlog, &k->uniform_stage, &k->sweep_ckpt, &k->warm_ll})
        b->counter = &k->allocations;
    if ((e = cudaStreamCreateWithFlags(&k->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&k->d_iteration), sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(k->d_iteration, 0, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaEventCreate(&k->ev0)) != cudaSuccess || (e = cudaEventCreate(&k->ev1)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&k->d_err), sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&k->d_flags), sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(k->d_err, 0, sizeof(int))) != cudaSuccess) {
        phb_destroy(k);
        return fail(PHB_E_CUDA, "stream/event setup: %s", cudaGetErrorString(e));
    }
    *out = k;
    return PHB_OK;
}

// Marks the rows that hold a long run of identical observations (float objects; see
// flag_long_runs_kernel).  The data must be resident.
static int flag_rows(phb_kernel *k) {
    if (k->dbl) return PHB_OK;
    PHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&k->d_rowflag), size_t(k->N)));
    PHB_CUDA(cudaMemsetAsync(k->d_rowflag, 0, size_t(k->N), k->stream));
    const int64_t n_win = k->L / phb::kRunWindow;
    if (n_win > 0) {
        const int threads = 256;
        const int64_t want = (k->N * n_win * 32 + threads - 1) / threads;
        phb::flag_long_runs_kernel<<<unsigned(std::min<int64_t>(want, int64_t(k->num_sms) * 32)), threads, 0, k->stream>>>(
            k->d_data, k->N, k->L, k->pitch, k->d_rowflag);
        PHB_CUDA(cudaGetLastError());
        k->launches += 1;
    }
    std::vector<uint8_t> flags(size_t(k->N));
    PHB_CUDA(cudaMemcpyAsync(flags.data(), k->d_rowflag, size_t(k->N), cudaMemcpyDeviceToHost, k->stream));
    PHB_CUDA(cudaStreamSynchronize(k->stream));
    k->n_flagged = 0;
    for (uint8_t f : flags) k->n_flagged += f;
    return PHB_OK;
}

}  // extern "C" (the helpers below are C++: a template and lambdas)

// Host rows [n_rows, row_bytes] (contiguous) -> device rows of `dst_pitch` bytes, through two pinned
// staging buffers filled by a few host threads while the previous slab is on the wire, so that neither a
// second host copy of a 50 GB matrix nor a single-threaded pass over it is needed.  after_slab(first_row,
// n_rows) is called once the slab's copy has been enqueued on k->stream.
template <typename AfterSlab>
static int staged_upload(phb_kernel *k, const int8_t *src, int64_t n_rows, int64_t row_bytes, int8_t *dst, int64_t dst_pitch,
                         AfterSlab after_slab) {
    constexpr size_t kSlabBytes = size_t(128) << 20;
    const int64_t slab_rows = std::max<int64_t>(1, std::min<int64_t>(n_rows, int64_t(kSlabBytes / size_t(dst_pitch))));
    const size_t slab_bytes = size_t(slab_rows) * size_t(dst_pitch);
    int8_t *pinned[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    int rc = PHB_OK;
    for (int i = 0; i < 2 && rc == PHB_OK; ++i) {
        if (cudaMallocHost(reinterpret_cast<void **>(&pinned[i]), slab_bytes) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(PHB_E_NOMEM, "pinned staging buffer of %zu bytes", slab_bytes);
        }
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_threads = int(std::min<unsigned>(16u, hw));  // (8 threads staged 25 GB/s on the 16-core box: half the PCIe rate)
    int64_t slab = 0;
    for (int64_t r0 = 0; r0 < n_rows && rc == PHB_OK; r0 += slab_rows, ++slab) {
        const int64_t nr = std::min(slab_rows, n_rows - r0);
        const int b = int(slab & 1);
        if (slab >= 2 && cudaEventSynchronize(done[b]) != cudaSuccess) {
            rc = fail(PHB_E_CUDA, "staging: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        auto fill = [&](int t) {
            const int64_t lo = nr * t / n_threads, hi = nr * (t + 1) / n_threads;
            if (dst_pitch == row_bytes) {
                memcpy(pinned[b] + lo * dst_pitch, src + (r0 + lo) * row_bytes, size_t(hi - lo) * size_t(row_bytes));
            } else {
                for (int64_t i = lo; i < hi; ++i) memcpy(pinned[b] + i * dst_pitch, src + (r0 + i) * row_bytes, size_t(row_bytes));
            }
        };
        if (n_threads > 1 && size_t(nr) * size_t(row_bytes) >= (size_t(8) << 20)) {
            std::vector<std::thread> pool;
            for (int t = 1; t < n_threads; ++t) pool.emplace_back(fill, t);
            fill(0);
            for (std::thread &th : pool) th.join();
        } else {
            for (int t = 0; t < n_threads; ++t) fill(t);
        }
        cudaError_t e = cudaMemcpyAsync(dst + r0 * dst_pitch, pinned[b], size_t(nr) * size_t(dst_pitch), cudaMemcpyHostToDevice, k->stream);
        if (e == cudaSuccess) e = cudaEventRecord(done[b], k->stream);
        if (e != cudaSuccess) {
            rc = fail(PHB_E_CUDA, "cudaMemcpyAsync(data): %s", cudaGetErrorString(e));
            break;
        }
        rc = after_slab(r0, nr);
    }
    if (cudaStreamSynchronize(k->stream) != cudaSuccess && rc == PHB_OK)
        rc = fail(PHB_E_CUDA, "upload: %s", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 2; ++i) {
        if (pinned[i]) cudaFreeHost(pinned[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    return rc;
}

// The reference's constructor checks (gpu.py:106-113), evaluated on the DEVICE over the resident rows
// [first_row, first_row + n_rows): clip to <= 1, pad the pitch with -1, note values < -1 and rows whose
// columns [check_from, L) hold no observation.  check_state = {first bad linear index, row flags}.
struct CheckState {
    unsigned long long *d_first_bad = nullptr;  // [2]: smallest linear index of a value < -1; first row without an observation
    uint8_t *d_row_observed = nullptr;          // [N]
};
static int check_alloc(phb_kernel *k, CheckState &cs) {
    PHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&cs.d_first_bad), 2 * sizeof(unsigned long long)));
    PHB_CUDA(cudaMemsetAsync(cs.d_first_bad, 0xff, 2 * sizeof(unsigned long long), k->stream));
    PHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&cs.d_row_observed), size_t(k->N)));
    return PHB_OK;
}
static void check_free(CheckState &cs) {
    if (cs.d_first_bad) cudaFree(cs.d_first_bad);
    if (cs.d_row_observed) cudaFree(cs.d_row_observed);
    cs = CheckState{};
}
static int check_rows(phb_kernel *k, CheckState &cs, int64_t first_row, int64_t n_rows, int64_t check_from) {
    const unsigned grid = unsigned(std::min<int64_t>(n_rows, int64_t(k->num_sms) * 8));
    phb::fixup_rows_kernel<<<grid, 256, 0, k->stream>>>(k->d_data + first_row * k->pitch, n_rows, k->L, k->pitch, check_from, first_row,
                                                         cs.d_first_bad, cs.d_row_observed);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}
static int check_verdict(phb_kernel *k, CheckState &cs, const char *what) {
    phb::first_zero_kernel<<<unsigned(std::min<int64_t>((k->N + 255) / 256, int64_t(k->num_sms) * 8)), 256, 0, k->stream>>>(
        cs.d_row_observed, k->N, cs.d_first_bad + 1);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    unsigned long long res[2];
    PHB_CUDA(cudaMemcpyAsync(res, cs.d_first_bad, sizeof res, cudaMemcpyDeviceToHost, k->stream));
    PHB_CUDA(cudaStreamSynchronize(k->stream));
    if (res[0] != ~0ull)
        return fail(PHB_E_DATA, "%s[%lld, %lld] < -1", what, (long long)(res[0] / (unsigned long long)k->L), (long long)(res[0] % (unsigned long long)k->L));
    if (res[1] != ~0ull) return fail(PHB_E_DATA, "data contains observations with all missing values (row %lld)", (long long)res[1]);
    return PHB_OK;
}

static int create_from_rows(int M, const int8_t *data, int64_t N, int64_t L, int64_t check_from, int double_precision, int device,
                            phb_kernel **out) {
    if (!out) return fail(PHB_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!(M == 4 || M == 8 || M == 16 || M == 32 || M == 64))
        return fail(PHB_E_INVALID, "M=%d is not supported (4, 8, 16, 32 or 64)", M);
    if (!data || N <= 0 || L <= 0) return fail(PHB_E_INVALID, "data must be a non-empty [N, L] int8 matrix");
    if (check_from < 0 || check_from >= L) return fail(PHB_E_INVALID, "overlap %lld not in [0, row length %lld)", (long long)check_from, (long long)L);
    phb_kernel *k = nullptr;
    if (int rc = new_kernel(M, N, L, double_precision, device, &k)) return rc;
    CheckState cs;
    int rc = check_alloc(k, cs);
    if (rc == PHB_OK)
        rc = staged_upload(k, data, N, L, k->d_data, k->pitch, [&](int64_t r0, int64_t nr) { return check_rows(k, cs, r0, nr, check_from); });
    if (rc == PHB_OK) rc = check_verdict(k, cs, "data");
    check_free(cs);
    if (rc == PHB_OK) rc = flag_rows(k);
    if (rc != PHB_OK) {
        const std::string msg = g_err;  // phb_destroy must not clobber the message
        phb_destroy(k);
        g_err = msg;
        return rc;
    }
    *out = k;
    return PHB_OK;
}

extern "C" {

int phb_create(int M, const int8_t *data, int64_t N, int64_t L, int double_precision, int device, phb_kernel **out) {
    return create_from_rows(M, data, N, L, 0, double_precision, device, out);
}

int phb_create_chunks(int M, const int8_t *chunks, int64_t N, int64_t W, int64_t overlap, int double_precision, int device,
                      phb_kernel **out) {
    return create_from_rows(M, chunks, N, W, overlap, double_precision, device, out);
}

int phb_create_from_contig(int M, const int8_t *het, int64_t n_rows, int64_t length, int64_t overlap,
                           int64_t chunk_size, int double_precision, int device, phb_kernel **out) {
    if (!out) return fail(PHB_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!het || n_rows <= 0 || length <= 0) return fail(PHB_E_INVALID, "het must be a non-empty [n_rows, length] int8 matrix");
    if (overlap < 0 || chunk_size <= 0) return fail(PHB_E_INVALID, "need overlap >= 0 and chunk_size > 0");
    const int64_t width = chunk_size + overlap;
    const int64_t n_chunks = (length + width - 1) / width;
    phb_kernel *k = nullptr;
    if (int rc = new_kernel(M, n_rows * n_chunks, width, double_precision, device, &k)) return rc;
    int8_t *d_het = nullptr;
    CheckState cs;
    int rc = check_alloc(k, cs);
    if (rc == PHB_OK && cudaMalloc(reinterpret_cast<void **>(&d_het), size_t(n_rows) * size_t(length)) != cudaSuccess) {
        cudaGetLastError();
        rc = fail(PHB_E_NOMEM, "While trying to allocate %lld bytes on GPU for the binned contig", (long long)(n_rows * length));
    }
    // the contig as ONE row of n_rows * length bytes cut into slabs (any slab boundary will do)
    if (rc == PHB_OK) {
        const int64_t total = n_rows * length, piece = int64_t(1) << 20;
        const int64_t whole = total / piece;
        if (whole > 0) rc = staged_upload(k, het, whole, piece, d_het, piece, [](int64_t, int64_t) { return PHB_OK; });
        if (rc == PHB_OK && total > whole * piece)
            rc = staged_upload(k, het + whole * piece, 1, total - whole * piece, d_het + whole * piece, total - whole * piece,
                               [](int64_t, int64_t) { return PHB_OK; });
    }
    if (rc == PHB_OK) {
        const int64_t total = k->N * k->pitch;
        const int threads = 256;
        const int blocks = int(std::min<int64_t>((total + threads - 1) / threads, int64_t(k->num_sms) * 16));
        phb::chunk_het_kernel<<<blocks, threads, 0, k->stream>>>(d_het, n_rows, length, chunk_size, width, n_chunks, k->d_data, k->pitch);
        k->launches += 1;
        if (cudaGetLastError() != cudaSuccess) rc = fail(PHB_E_CUDA, "chunking on the device failed");
    }
    // the reference's checks on the chunk matrix (values >= -1; the data part of every chunk holds an
    // observation: mcmc.py:203 splits the warm-up columns off before gpu.py:111-113 looks)
    if (rc == PHB_OK) rc = check_rows(k, cs, 0, k->N, overlap);
    if (rc == PHB_OK) rc = check_verdict(k, cs, "chunks");
    if (d_het) cudaFree(d_het);
    check_free(cs);
    if (rc == PHB_OK) rc = flag_rows(k);
    if (rc != PHB_OK) {
        const std::string msg = g_err;
        phb_destroy(k);
        g_err = msg;
        return rc;
    }
    *out = k;
    return PHB_OK;
}

int phb_download_data(const phb_kernel *k, int8_t *out) {
    if (int rc = check_handle(k)) return rc;
    if (!out) return fail(PHB_E_INVALID, "out is NULL");
    PHB_CUDA(cudaSetDevice(k->device));
    PHB_CUDA(cudaMemcpy2D(out, size_t(k->L), k->d_data, size_t(k->pitch), size_t(k->L), size_t(k->N), cudaMemcpyDeviceToHost));
    return PHB_OK;
}

void phb_destroy(phb_kernel *k) {
    if (!k) return;
    cudaSetDevice(k->device);
    if (k->stream) cudaStreamSynchronize(k->stream);
    k->params.release();
    k->inds.release();
    k->ll.release();
    k->dlog.release();
    k->ckpt.release();
    k->gacc.release();
    k->xall.release();
    k->sall.release();
    k->split.release();
    k->term_params.release();
    k->term_ll.release();
    k->term_dlog.release();
    k->term_sums.release();
    k->term_io.release();
    k->transfer_rows.release();
    k->transfer_log.release();
    k->bnd_alpha.release();
    k->bnd_beta.release();
    k->seg_dlog.release();
    k->uniform_stage.release();
    k->sweep_ckpt.release();
    k->warm_ll.release();
    if (k->d_rowflag) cudaFree(k->d_rowflag);
    if (k->d_data) cudaFree(k->d_data);
    if (k->d_err) cudaFree(k->d_err);
    if (k->d_flags) cudaFree(k->d_flags);
    if (k->d_iteration) cudaFree(k->d_iteration);
    if (k->ev0) cudaEventDestroy(k->ev0);
    if (k->ev1) cudaEventDestroy(k->ev1);
    if (k->stream) cudaStreamDestroy(k->stream);
    delete k;
}

int phb_M(const phb_kernel *k) { return k ? k->M : 0; }
int phb_double_precision(const phb_kernel *k) { return k ? k->dbl : 0; }
int64_t phb_num_rows(const phb_kernel *k) { return k ? k->N : 0; }
int64_t phb_row_length(const phb_kernel *k) { return k ? k->L : 0; }
int phb_device(const phb_kernel *k) { return k ? k->device : -1; }
int64_t phb_launch_count(const phb_kernel *k) { return k ? k->launches : 0; }
const char *phb_last_kernel_name(const phb_kernel *k) { return k ? k->last_name : ""; }

const int8_t *phb_device_data(const phb_kernel *k, int64_t *pitch) {
    if (!k) return nullptr;
    if (pitch) *pitch = k->pitch;
    return k->d_data;
}

int phb_set_store_all(phb_kernel *k, int mode) {
    if (int rc = check_handle(k)) return rc;
    if (mode < -1 || mode > 1) return fail(PHB_E_INVALID, "store-all mode must be -1 (auto), 0 (off) or 1 (on)");
    k->store_all_mode = mode;
    return PHB_OK;
}

int phb_set_precision_escalation(phb_kernel *k, int enabled) {
    if (int rc = check_handle(k)) return rc;
    k->escalate = enabled ? 1 : 0;
    return PHB_OK;
}

int64_t phb_num_escalated_rows(const phb_kernel *k) { return k ? k->n_flagged : 0; }

int phb_set_parallel_in_time(phb_kernel *k, int mode) {
    if (int rc = check_handle(k)) return rc;
    if (mode < -1 || mode > 2)
        return fail(PHB_E_INVALID, "parallel-in-time mode must be -1 (auto), 0 (off), 1 (operators) or 2 (sweeps)");
    k->parallel_in_time = mode;
    return PHB_OK;
}

int phb_set_threads_per_pair(phb_kernel *k, int threads_per_pair) {
    if (int rc = check_handle(k)) return rc;
    if (threads_per_pair != 0) {
        bool ok = false;
        for (const Variant &v : variants()) ok |= (v.M == k->M && v.dbl == (k->dbl != 0) && v.T == threads_per_pair);
        if (!ok) return fail(PHB_E_INVALID, "threads_per_pair=%d is not compiled for M=%d", threads_per_pair, k->M);
    }
    k->force_T = threads_per_pair;
    return PHB_OK;
}

static int device_eval(phb_kernel *k, const void *params6, int64_t params_stride_b, int64_t params_stride_s,
                       const void *pi, int64_t pi_stride_b, int64_t pi_stride_s, const int64_t *inds, int64_t B,
                       int64_t S, int want_grad, double *ll, void *dlog, cudaStream_t stream, int64_t n_sites,
                       int out_mode, int64_t warm_len = 0) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (B == 0 || S == 0) return PHB_OK;
    if (!params6 || !pi || !inds || !ll) return fail(PHB_E_INVALID, "NULL device pointer");
    if (want_grad && !dlog) return fail(PHB_E_INVALID, "want_grad is set but dlog is NULL");
    if (n_sites <= 0 || n_sites > k->L) return fail(PHB_E_INVALID, "number of sites %lld not in [1, %lld]", (long long)n_sites, (long long)k->L);
    PHB_CUDA(cudaSetDevice(k->device));
    phb::KernelArgs a{};
    a.data = k->d_data;
    a.pitch = k->pitch;
    a.n_rows = k->N;
    a.L = n_sites;
    a.inds = inds;
    a.B = B;
    a.S = S;
    a.params6 = params6;
    a.pstride_b = params_stride_b;
    a.pstride_s = params_stride_s;
    a.pi = pi;
    a.pistride_b = pi_stride_b;
    a.pistride_s = pi_stride_s;
    a.ll = ll;
    a.dlog = want_grad ? dlog : nullptr;
    a.alpha_out = nullptr;
    a.out_mode = out_mode;
    a.warm_len = want_grad ? warm_len : 0;  // (a request: a parallel-in-time gradient path may honour it, see warm_fused)
    k->warm_fused = false;
    return launch(k, a, want_grad != 0, stream);
}

int phb_loglik_device(phb_kernel *k, const void *params6, int64_t params_stride_b, int64_t params_stride_s,
                      const void *pi, int64_t pi_stride_b, int64_t pi_stride_s, const int64_t *inds,
                      int64_t B, int64_t S, int want_grad, double *ll, void *dlog, void *stream) {
    if (int rc = check_handle(k)) return rc;
    return device_eval(k, params6, params_stride_b, params_stride_s, pi, pi_stride_b, pi_stride_s, inds, B, S,
                       want_grad, ll, dlog, static_cast<cudaStream_t>(stream), k->L, 0);
}

int phb_loglik_warmup_device(phb_kernel *k, const void *params7, const int64_t *inds, int64_t B, int64_t S,
                             int64_t overlap, int want_grad, double *ll, void *dlog, void *stream) {
    if (int rc = check_handle(k)) return rc;
    if (overlap < 0 || overlap >= k->L)
        return fail(PHB_E_INVALID, "overlap %lld not in [0, row length %lld)", (long long)overlap, (long long)k->L);
    const int M = k->M;
    const char *base = static_cast<const char *>(params7);
    const void *pi = base + size_t(6) * M * k->elem();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // LL over warm-up + chunk, started from the particle's stationary pi ...
    int rc = device_eval(k, params7, 7 * M, 0, pi, 7 * M, 0, inds, B, S, want_grad, ll, dlog, st, k->L, 0, overlap);
    if (rc != PHB_OK || overlap == 0) return rc;
    // ... minus LL over the warm-up bins alone: either the first call scored them as one more segment of its
    // segment passes (small minibatches, see KernelArgs::warm_len) or a second launch on the same stream does
    if (k->warm_fused) return PHB_OK;
    return device_eval(k, params7, 7 * M, 0, pi, 7 * M, 0, inds, B, S, want_grad, ll, dlog, st, overlap, 1);
}

int phb_reserve(phb_kernel *k, int64_t B, int64_t S_max, int64_t overlap, int want_grad) {
    if (int rc = check_handle(k)) return rc;
    if (B <= 0 || S_max <= 0) return fail(PHB_E_INVALID, "need B > 0 and S_max > 0");
    if (overlap < 0 || overlap >= k->L) return fail(PHB_E_INVALID, "overlap %lld not in [0, row length %lld)", (long long)overlap, (long long)k->L);
    PHB_CUDA(cudaSetDevice(k->device));
    // Dry run of the dispatcher for every minibatch size a caller may use: the path (and with it the scratch)
    // depends on S, not monotonically.  Small sizes are walked one by one, then powers of two up to S_max.
    std::vector<int64_t> sizes;
    for (int64_t S = 1; S <= std::min<int64_t>(S_max, 64); ++S) sizes.push_back(S);
    for (int64_t S = 128; S < S_max; S *= 2) sizes.push_back(S);
    sizes.push_back(S_max);
    std::vector<int32_t> widths(size_t(k->M), 1);
    void *dummy = k->d_err;  // any non-NULL device address: nothing is dereferenced in a dry run
    int rc = PHB_OK;
    k->dry = true;
    for (int64_t S : sizes) {
        // the one-call term (particles -> parameters -> fused warm-up evaluation -> sums -> VJP) ...
        rc = phb_hmm_term_device(k, static_cast<const double *>(dummy), B, widths.data(), k->M, 1.0, static_cast<const int64_t *>(dummy), S,
                                 overlap, 1.0, static_cast<double *>(dummy), want_grad ? static_cast<double *>(dummy) : nullptr, k->stream);
        // ... and the plain kernel call the reference's interface makes (gpu.py:182-325)
        if (rc == PHB_OK)
            rc = phb_loglik_device(k, dummy, 6 * k->M, 0, dummy, k->M, 0, static_cast<const int64_t *>(dummy), B, S, want_grad,
                                   static_cast<double *>(dummy), dummy, k->stream);
        if (rc != PHB_OK) break;
    }
    // host-entry staging of the term
    if (rc == PHB_OK) {
        const int64_t P = 2 + int64_t(k->M) + 1;
        rc = k->term_io.reserve((size_t(B) * P * 2 + size_t(B)) * sizeof(double) + size_t(S_max) * sizeof(int64_t));
    }
    k->dry = false;
    return rc;
}

int64_t phb_allocation_count(const phb_kernel *k) { return k ? k->allocations : 0; }

void phb_minibatch_indices(uint64_t seed, uint64_t iteration, int64_t N, int64_t S, int64_t *inds) {
    for (int64_t s = 0; s < S; ++s) inds[s] = phb::minibatch_index(seed, iteration, uint64_t(s), N);
}

int phb_set_iteration(phb_kernel *k, uint64_t iteration, void *stream) {
    if (int rc = check_handle(k)) return rc;
    PHB_CUDA(cudaSetDevice(k->device));
    // (a kernel-argument store, not a memcpy from pageable host memory: legal inside a stream capture)
    phb::set_counter_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(k->d_iteration, (unsigned long long)iteration);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_sample_minibatch_device(phb_kernel *k, uint64_t seed, int64_t S, int64_t *inds, void *stream) {
    if (int rc = check_handle(k)) return rc;
    if (S < 0 || (S > 0 && !inds)) return fail(PHB_E_INVALID, "bad minibatch buffer");
    if (S == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    phb::sample_minibatch_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, k->d_iteration, k->N, S, inds);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_measure_fp32_peak(int device, double *independent_tflops, double *accumulate_tflops) {
    if (!independent_tflops) return fail(PHB_E_INVALID, "NULL pointer");
    PHB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PHB_CUDA(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, ctas = prop.multiProcessorCount * 8;
    float *out = nullptr, *in = nullptr;
    PHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&out), sizeof(float) * threads * ctas));
    PHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&in), sizeof(float) * 64));
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 1e-3f * i;
    PHB_CUDA(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    PHB_CUDA(cudaEventCreate(&e0));
    PHB_CUDA(cudaEventCreate(&e1));
    const double flop = 2.0 * double(threads) * ctas * phb::kPeakIters * phb::kPeakChains;
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 13; ++rep) {  // 3 warm-ups, best of 10
            cudaEventRecord(e0, nullptr);
            if (which == 0)
                phb::ffma_peak_kernel<<<ctas, threads>>>(out, 0.999f, 1e-3f);
            else
                phb::ffma_accumulate_kernel<<<ctas, threads>>>(out, in);
            cudaEventRecord(e1, nullptr);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 3 && ms > 0.f) best[which] = std::max(best[which], flop / (double(ms) * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(in);
    PHB_CUDA(cudaGetLastError());
    *independent_tflops = best[0];
    if (accumulate_tflops) *accumulate_tflops = best[1];
    return PHB_OK;
}

void *phb_stream(const phb_kernel *k) { return k ? static_cast<void *>(k->stream) : nullptr; }

int phb_sync(phb_kernel *k) {
    if (int rc = check_handle(k)) return rc;
    PHB_CUDA(cudaSetDevice(k->device));
    PHB_CUDA(cudaStreamSynchronize(k->stream));
    if (k->timed) PHB_CUDA(cudaEventSynchronize(k->ev1));  // last launch, whatever stream it used
    int flag = 0;
    PHB_CUDA(cudaMemcpy(&flag, k->d_err, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) PHB_CUDA(cudaMemset(k->d_err, 0, sizeof(int)));
    if (flag & 1) return fail(PHB_E_INVALID, "index out of range: need 0 <= inds < N=%lld", (long long)k->N);
    // (reference: assert np.isfinite on the parameters, gpu.py:214; a non-finite parameter gives a non-finite ll)
    if (flag & 4)
        return fail(PHB_E_INVALID, "parameters outside the domain of the rescaled kernel: need v[k] > 0 for k >= 1 and emis0[k] > 0 "
                                   "(PSMCParams.from_dm clips to [1e-20, 1 - 1e-20]; PHB_SFORM=0 selects the unscaled kernel)");
    if (flag & 2) return fail(PHB_E_INVALID, "not all parameters / results finite");
    return PHB_OK;
}

float phb_last_kernel_ms(phb_kernel *k) {
    if (!k || !k->timed) return -1.f;
    cudaSetDevice(k->device);
    if (cudaEventSynchronize(k->ev1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, k->ev0, k->ev1) != cudaSuccess) return -1.f;
    return ms;
}

static int host_eval(phb_kernel *k, const void *params, size_t params_elems, int64_t pstride_b, int64_t pstride_s,
                     int64_t pi_offset, int64_t pistride_b, int64_t pistride_s, const void *pi_host,
                     size_t pi_elems, const int64_t *inds, int64_t B, int64_t S, int want_grad, double *ll,
                     void *dlog) {
    const size_t es = k->elem();
    const int M = k->M;
    for (int64_t s = 0; s < S; ++s)
        if (inds[s] < 0 || inds[s] >= k->N)
            return fail(PHB_E_INVALID, "0 <= inds[%lld]=%lld < N=%lld violated", (long long)s, (long long)inds[s], (long long)k->N);
    const bool fin = k->dbl ? (all_finite(static_cast<const double *>(params), params_elems) &&
                               (!pi_host || all_finite(static_cast<const double *>(pi_host), pi_elems)))
                            : (all_finite(static_cast<const float *>(params), params_elems) &&
                               (!pi_host || all_finite(static_cast<const float *>(pi_host), pi_elems)));
    if (!fin) return fail(PHB_E_INVALID, "not all parameters finite");
    PHB_CUDA(cudaSetDevice(k->device));
    int rc;
    if ((rc = k->params.reserve((params_elems + pi_elems) * es)) != PHB_OK) return rc;
    if ((rc = k->inds.reserve(size_t(S) * sizeof(int64_t))) != PHB_OK) return rc;
    if ((rc = k->ll.reserve(size_t(B) * S * sizeof(double))) != PHB_OK) return rc;
    if (want_grad && (rc = k->dlog.reserve(size_t(B) * S * 7 * M * es)) != PHB_OK) return rc;
    char *d_params = static_cast<char *>(k->params.ptr);
    PHB_CUDA(cudaMemcpyAsync(d_params, params, params_elems * es, cudaMemcpyHostToDevice, k->stream));
    const char *d_pi = d_params + size_t(pi_offset) * es;
    if (pi_host) {
        PHB_CUDA(cudaMemcpyAsync(d_params + params_elems * es, pi_host, pi_elems * es, cudaMemcpyHostToDevice, k->stream));
        d_pi = d_params + params_elems * es;
    }
    PHB_CUDA(cudaMemcpyAsync(k->inds.ptr, inds, size_t(S) * sizeof(int64_t), cudaMemcpyHostToDevice, k->stream));
    rc = phb_loglik_device(k, d_params, pstride_b, pstride_s, d_pi, pistride_b, pistride_s,
                           static_cast<const int64_t *>(k->inds.ptr), B, S, want_grad,
                           static_cast<double *>(k->ll.ptr), want_grad ? k->dlog.ptr : nullptr, k->stream);
    if (rc != PHB_OK) return rc;
    PHB_CUDA(cudaMemcpyAsync(ll, k->ll.ptr, size_t(B) * S * sizeof(double), cudaMemcpyDeviceToHost, k->stream));
    if (want_grad)
        PHB_CUDA(cudaMemcpyAsync(dlog, k->dlog.ptr, size_t(B) * S * 7 * M * es, cudaMemcpyDeviceToHost, k->stream));
    return phb_sync(k);
}

int phb_loglik_host(phb_kernel *k, const void *params, const int64_t *inds, int64_t B, int64_t S, int want_grad,
                    double *ll, void *dlog) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (B == 0 || S == 0) return PHB_OK;
    if (!params || !inds || !ll) return fail(PHB_E_INVALID, "NULL pointer");
    if (want_grad && !dlog) return fail(PHB_E_INVALID, "want_grad is set but dlog is NULL");
    const int M = k->M;
    const size_t es = k->elem();
    const int64_t blk = 7 * M;
    const size_t n_par = size_t(B) * S * blk;
    for (int64_t s = 0; s < S; ++s)
        if (inds[s] < 0 || inds[s] >= k->N)
            return fail(PHB_E_INVALID, "0 <= inds[%lld]=%lld < N=%lld violated", (long long)s, (long long)inds[s], (long long)k->N);
    PHB_CUDA(cudaSetDevice(k->device));
    int rc;
    if ((rc = k->params.reserve(n_par * es)) != PHB_OK) return rc;
    if ((rc = k->inds.reserve(size_t(S) * sizeof(int64_t))) != PHB_OK) return rc;
    if ((rc = k->ll.reserve(size_t(B) * S * sizeof(double))) != PHB_OK) return rc;
    if (want_grad && (rc = k->dlog.reserve(size_t(B) * S * blk * es)) != PHB_OK) return rc;
    PHB_CUDA(cudaMemcpyAsync(k->params.ptr, params, n_par * es, cudaMemcpyHostToDevice, k->stream));
    PHB_CUDA(cudaMemcpyAsync(k->inds.ptr, inds, size_t(S) * sizeof(int64_t), cudaMemcpyHostToDevice, k->stream));
    // On the device: are all parameters finite (gpu.py:214), and are rows b..emis1 identical across
    // the S chunks of every particle (the way the reference builds its argument, model.py:55)?  Then
    // the kernel reads one copy per particle.
    PHB_CUDA(cudaMemsetAsync(k->d_flags, 0, sizeof(int), k->stream));
    {
        const int threads = 256;
        const int blocks = int(std::min<int64_t>((int64_t(n_par) + threads - 1) / threads, int64_t(k->num_sms) * 8));
        if (k->dbl)
            phb::validate_params_kernel<double><<<blocks, threads, 0, k->stream>>>(static_cast<const double *>(k->params.ptr), B, S, M, k->d_flags);
        else
            phb::validate_params_kernel<float><<<blocks, threads, 0, k->stream>>>(static_cast<const float *>(k->params.ptr), B, S, M, k->d_flags);
        PHB_CUDA(cudaGetLastError());
        k->launches += 1;
    }
    int flags = 0;
    PHB_CUDA(cudaMemcpyAsync(&flags, k->d_flags, sizeof(int), cudaMemcpyDeviceToHost, k->stream));
    PHB_CUDA(cudaStreamSynchronize(k->stream));
    if (flags & 1) return fail(PHB_E_INVALID, "not all parameters finite");
    const bool shared = (flags & 2) == 0;
    const char *d_params = static_cast<const char *>(k->params.ptr);
    rc = phb_loglik_device(k, d_params, S * blk, shared ? 0 : blk, d_params + size_t(6) * M * es, S * blk, blk,
                           static_cast<const int64_t *>(k->inds.ptr), B, S, want_grad, static_cast<double *>(k->ll.ptr),
                           want_grad ? k->dlog.ptr : nullptr, k->stream);
    if (rc != PHB_OK) return rc;
    PHB_CUDA(cudaMemcpyAsync(ll, k->ll.ptr, size_t(B) * S * sizeof(double), cudaMemcpyDeviceToHost, k->stream));
    if (want_grad)
        PHB_CUDA(cudaMemcpyAsync(dlog, k->dlog.ptr, size_t(B) * S * blk * es, cudaMemcpyDeviceToHost, k->stream));
    return phb_sync(k);
}

int phb_loglik_shared_host(phb_kernel *k, const void *params6, const void *pi, int pi_per_pair,
                           const int64_t *inds, int64_t B, int64_t S, int want_grad, double *ll, void *dlog) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (B == 0 || S == 0) return PHB_OK;
    if (!params6 || !pi || !inds || !ll) return fail(PHB_E_INVALID, "NULL pointer");
    if (want_grad && !dlog) return fail(PHB_E_INVALID, "want_grad is set but dlog is NULL");
    const int M = k->M;
    const size_t pi_elems = pi_per_pair ? size_t(B) * S * M : size_t(B) * M;
    return host_eval(k, params6, size_t(B) * 6 * M, 6 * M, 0, 0, pi_per_pair ? S * M : M, pi_per_pair ? M : 0, pi,
                     pi_elems, inds, B, S, want_grad, ll, dlog);
}

int phb_loglik_warmup_host(phb_kernel *k, const void *params7, const int64_t *inds, int64_t B, int64_t S,
                           int64_t overlap, int want_grad, double *ll, void *dlog) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (B == 0 || S == 0) return PHB_OK;
    if (!params7 || !inds || !ll) return fail(PHB_E_INVALID, "NULL pointer");
    if (want_grad && !dlog) return fail(PHB_E_INVALID, "want_grad is set but dlog is NULL");
    const int M = k->M;
    const size_t es = k->elem();
    const size_t n_par = size_t(B) * 7 * M;
    for (int64_t s = 0; s < S; ++s)
        if (inds[s] < 0 || inds[s] >= k->N)
            return fail(PHB_E_INVALID, "0 <= inds[%lld]=%lld < N=%lld violated", (long long)s, (long long)inds[s], (long long)k->N);
    const bool fin = k->dbl ? all_finite(static_cast<const double *>(params7), n_par)
                            : all_finite(static_cast<const float *>(params7), n_par);
    if (!fin) return fail(PHB_E_INVALID, "not all parameters finite");
    PHB_CUDA(cudaSetDevice(k->device));
    int rc;
    if ((rc = k->params.reserve(n_par * es)) != PHB_OK) return rc;
    if ((rc = k->inds.reserve(size_t(S) * sizeof(int64_t))) != PHB_OK) return rc;
    if ((rc = k->ll.reserve(size_t(B) * S * sizeof(double))) != PHB_OK) return rc;
    if (want_grad && (rc = k->dlog.reserve(size_t(B) * S * 7 * M * es)) != PHB_OK) return rc;
    PHB_CUDA(cudaMemcpyAsync(k->params.ptr, params7, n_par * es, cudaMemcpyHostToDevice, k->stream));
    PHB_CUDA(cudaMemcpyAsync(k->inds.ptr, inds, size_t(S) * sizeof(int64_t), cudaMemcpyHostToDevice, k->stream));
    rc = phb_loglik_warmup_device(k, k->params.ptr, static_cast<const int64_t *>(k->inds.ptr), B, S, overlap, want_grad,
                                  static_cast<double *>(k->ll.ptr), want_grad ? k->dlog.ptr : nullptr, k->stream);
    if (rc != PHB_OK) return rc;
    PHB_CUDA(cudaMemcpyAsync(ll, k->ll.ptr, size_t(B) * S * sizeof(double), cudaMemcpyDeviceToHost, k->stream));
    if (want_grad)
        PHB_CUDA(cudaMemcpyAsync(dlog, k->dlog.ptr, size_t(B) * S * 7 * M * es, cudaMemcpyDeviceToHost, k->stream));
    return phb_sync(k);
}

}  // extern "C"
// one warp per particle (forward) or per (particle, direction) (VJP), see psmc_params.cuh
template <bool VJP> static void launch_params_kernel(const phb::ParamsArgs &a, int64_t n_items, cudaStream_t stream) {
    const int per_cta = phb::params_items_per_cta(a.M);
    phb::psmc_params_warp_kernel<VJP><<<unsigned((n_items + per_cta - 1) / per_cta), phb::params_warps_per_cta(a.M) * 32, phb::params_smem_bytes(a.M), stream>>>(a);
}
extern "C" {

static int fill_params_args(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                            double theta, phb::ParamsArgs &a) {
    if (int rc = check_handle(k)) return rc;
    if (!x || !epoch_widths || B < 0) return fail(PHB_E_INVALID, "NULL pointer or negative batch");
    if (n_epochs < 1 || n_epochs > phb::kMaxM) return fail(PHB_E_INVALID, "n_epochs=%d out of range", n_epochs);
    int total = 0;
    for (int e = 0; e < n_epochs; ++e) {
        if (epoch_widths[e] <= 0) return fail(PHB_E_INVALID, "epochs must be positive");
        total += epoch_widths[e];
        a.widths[e] = epoch_widths[e];
    }
    if (total != k->M) return fail(PHB_E_INVALID, "pattern has %d states, kernel was built for M=%d", total, k->M);
    if (!(theta > 0.0) || !std::isfinite(theta)) return fail(PHB_E_INVALID, "theta must be positive and finite");
    a.x = x;
    a.B = B;
    a.n_epochs = n_epochs;
    a.P = 2 + n_epochs + 1;
    a.M = k->M;
    a.theta = theta;
    a.out_double = k->dbl;
    return PHB_OK;
}

int phb_params_from_particles(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                              double theta, void *params7, void *stream) {
    phb::ParamsArgs a{};
    if (int rc = fill_params_args(k, x, B, epoch_widths, n_epochs, theta, a)) return rc;
    if (!params7) return fail(PHB_E_INVALID, "params7 is NULL");
    if (B == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    a.params7 = params7;
    if (k->dry) return PHB_OK;
    launch_params_kernel<false>(a, B, static_cast<cudaStream_t>(stream));
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_params_vjp(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs, double theta,
                   const void *cotangent, double *grad_x, void *stream) {
    phb::ParamsArgs a{};
    if (int rc = fill_params_args(k, x, B, epoch_widths, n_epochs, theta, a)) return rc;
    if (!cotangent || !grad_x) return fail(PHB_E_INVALID, "NULL pointer");
    if (B == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    a.cotangent = cotangent;
    a.grad_x = grad_x;
    if (k->dry) return PHB_OK;
    launch_params_kernel<true>(a, B * a.P, static_cast<cudaStream_t>(stream));
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_hmm_term_sums_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                             double theta, const int64_t *inds, int64_t S, int64_t overlap, int want_grad, double *sums,
                             void *stream) {
    phb::ParamsArgs pa{};
    if (int rc = fill_params_args(k, x, B, epoch_widths, n_epochs, theta, pa)) return rc;
    if (S < 0) return fail(PHB_E_INVALID, "negative minibatch size");
    if (!sums || (S > 0 && !inds)) return fail(PHB_E_INVALID, "NULL pointer");
    if (B == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    const int C = 7 * k->M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if ((rc = k->term_params.reserve(size_t(B) * C * k->elem())) != PHB_OK) return rc;
    if ((rc = k->term_ll.reserve(size_t(B) * std::max<int64_t>(S, 1) * sizeof(double))) != PHB_OK) return rc;
    if (want_grad && (rc = k->term_dlog.reserve(size_t(B) * std::max<int64_t>(S, 1) * C * k->elem())) != PHB_OK) return rc;
    if (S > 0) {
        if ((rc = phb_params_from_particles(k, x, B, epoch_widths, n_epochs, theta, k->term_params.ptr, stream)) != PHB_OK) return rc;
        if ((rc = phb_loglik_warmup_device(k, k->term_params.ptr, inds, B, S, overlap, want_grad, static_cast<double *>(k->term_ll.ptr),
                                           want_grad ? k->term_dlog.ptr : nullptr, stream)) != PHB_OK)
            return rc;
    }
    if (k->dry) return PHB_OK;
    const int threads = 128;
    if (k->dbl)
        phb::sum_over_chunks_kernel<double><<<unsigned(B), threads, 0, st>>>(static_cast<const double *>(k->term_ll.ptr),
                                                                            want_grad ? static_cast<const double *>(k->term_dlog.ptr) : nullptr, S, C, sums);
    else
        phb::sum_over_chunks_kernel<float><<<unsigned(B), threads, 0, st>>>(static_cast<const double *>(k->term_ll.ptr),
                                                                           want_grad ? static_cast<const float *>(k->term_dlog.ptr) : nullptr, S, C, sums);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_hmm_term_sharded_plan(phb_kernel *k, int64_t B, int64_t S, int64_t overlap, int world, int64_t *n_segments, int64_t *slot_bytes) {
    if (int rc = check_handle(k)) return rc;
    if (!n_segments || !slot_bytes) return fail(PHB_E_INVALID, "NULL pointer");
    (void)overlap;
    *n_segments = 0;
    *slot_bytes = 0;
    if (B <= 0 || S <= 0 || world < 1) return fail(PHB_E_INVALID, "need B, S > 0 and world >= 1");
    PHB_CUDA(cudaSetDevice(k->device));
    ShardPlan p;
    if (shard_plan(k, B, S, k->L, world, p)) {
        *n_segments = p.n_seg;
        *slot_bytes = int64_t(p.slot_bytes);
    }
    return PHB_OK;
}

int phb_hmm_term_sharded_begin(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs, double theta,
                               const int64_t *inds, int64_t S, int64_t overlap, int rank, int world, void *gather, void *stream) {
    phb::ParamsArgs pa{};
    if (int rc = fill_params_args(k, x, B, epoch_widths, n_epochs, theta, pa)) return rc;
    if (!inds || !gather || S <= 0 || rank < 0 || rank >= world) return fail(PHB_E_INVALID, "bad argument");
    if (overlap < 0 || overlap >= k->L) return fail(PHB_E_INVALID, "overlap %lld not in [0, row length %lld)", (long long)overlap, (long long)k->L);
    PHB_CUDA(cudaSetDevice(k->device));
    ShardPlan p;
    if (!shard_plan(k, B, S, k->L, world, p)) return fail(PHB_E_INVALID, "time-axis sharding does not apply to this call (see phb_hmm_term_sharded_plan)");
    const TransferVariant *tv = transfer_variant(k->M);
    const int M = k->M, C = 7 * M;
    int rc;
    if ((rc = k->term_params.reserve(size_t(B) * C * k->elem())) != PHB_OK) return rc;
    if ((rc = phb_params_from_particles(k, x, B, epoch_widths, n_epochs, theta, k->term_params.ptr, stream)) != PHB_OK) return rc;
    const int64_t seg_lo = std::min<int64_t>(p.n_seg, int64_t(rank) * p.per_rank);
    const int64_t seg_hi = std::min<int64_t>(p.n_seg, seg_lo + p.per_rank);
    const int64_t n_local = std::max<int64_t>(seg_hi - seg_lo, 0);
    if ((rc = k->transfer_rows.reserve(size_t(std::max<int64_t>(n_local, 1)) * B * S * M * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->transfer_log.reserve(size_t(std::max<int64_t>(n_local, 1)) * B * S * M * sizeof(double))) != PHB_OK) return rc;
    phb::TransferArgs ta{};
    phb::KernelArgs &a = ta.k;
    a.data = k->d_data;
    a.pitch = k->pitch;
    a.n_rows = k->N;
    a.L = k->L;
    a.inds = inds;
    a.B = B;
    a.S = S;
    a.params6 = k->term_params.ptr;
    a.pstride_b = C;
    a.pi = static_cast<const char *>(k->term_params.ptr) + size_t(6) * M * k->elem();
    a.pistride_b = C;
    a.err_flag = k->d_err;
    ta.n_seg = p.n_seg;
    ta.seg_len = p.seg_len;
    // this process's operators stay local: segment g lives at [g - seg_lo] (slots of per_rank segments that all
    // alias the one local buffer: g / per_rank == rank for every g of the slice)
    ta.rows = static_cast<float *>(k->transfer_rows.ptr);
    ta.row_log2 = static_cast<double *>(k->transfer_log.ptr);
    ta.segs_per_slot = p.per_rank;
    ta.seg_first = seg_lo;
    ta.n_seg_local = n_local;
    if (n_local > 0 && (rc = launch_transfer_rows(k, tv, ta, static_cast<cudaStream_t>(stream))) != PHB_OK) return rc;
    if (k->dry) return PHB_OK;
    // ... and their product goes into this process's slot of the all-gather buffer
    char *slot = static_cast<char *>(gather) + size_t(rank) * p.slot_bytes;
    float *out_rows = reinterpret_cast<float *>(slot);
    double *out_log2 = reinterpret_cast<double *>(slot + p.rows_bytes);
    void *pargs[] = {&ta, &out_rows, &out_log2};
    PHB_CUDA(cudaLaunchKernel(tv->product_func, dim3(unsigned((B * S * M * M + 127) / 128)), dim3(128), pargs, 0, static_cast<cudaStream_t>(stream)));
    k->launches += 1;
    return PHB_OK;
}

int phb_hmm_term_sharded_end(phb_kernel *k, const int64_t *inds, int64_t B, int64_t S, int64_t overlap, int rank, int world,
                             const void *gather, double *sums, void *stream) {
    if (int rc = check_handle(k)) return rc;
    if (!inds || !gather || !sums || B <= 0 || S <= 0 || rank < 0 || rank >= world) return fail(PHB_E_INVALID, "bad argument");
    if (overlap < 0 || overlap >= k->L) return fail(PHB_E_INVALID, "overlap %lld not in [0, row length %lld)", (long long)overlap, (long long)k->L);
    PHB_CUDA(cudaSetDevice(k->device));
    ShardPlan p;
    if (!shard_plan(k, B, S, k->L, world, p)) return fail(PHB_E_INVALID, "time-axis sharding does not apply to this call");
    const TransferVariant *tv = transfer_variant(k->M);
    const Variant *gv = segment_variant(k);
    const int M = k->M, C = 7 * M;
    const int64_t n_pairs = B * S;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t seg_lo = std::min<int64_t>(p.n_seg, int64_t(rank) * p.per_rank);
    const int64_t seg_hi = std::min<int64_t>(p.n_seg, seg_lo + p.per_rank);
    const int64_t n_local = seg_hi - seg_lo;
    const int64_t occ = k->occupancy[gv->func];
    const bool marked = !k->dbl && k->escalate && k->n_flagged > 0;
    // process 0 scores the warm-up term as one more segment of its slice when its groups stay resident (KernelArgs::warm_len)
    const bool warm = rank == 0 && overlap > 0 && n_local > 0 && !marked && k->env_fuse_warmup != 0 && !is_sform(gv) &&
                      overlap <= p.seg_len && p.seg_ctas * (n_local + 1) <= occ * k->num_sms;
    const int64_t n_slots = n_local + (warm ? 1 : 0);
    const int64_t n_groups = p.seg_ctas * n_slots;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(n_groups, occ * k->num_sms));
    int rc;
    if ((rc = k->term_ll.reserve(size_t(n_pairs) * sizeof(double))) != PHB_OK) return rc;
    if ((rc = k->term_dlog.reserve(size_t(n_pairs) * C * k->elem())) != PHB_OK) return rc;
    if ((rc = k->bnd_alpha.reserve(size_t(n_pairs) * (p.n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->bnd_beta.reserve(size_t(n_pairs) * (p.n_seg + 1) * M * sizeof(float))) != PHB_OK) return rc;
    if ((rc = k->seg_dlog.reserve(size_t(n_pairs) * std::max<int64_t>(n_slots, 1) * C * sizeof(float))) != PHB_OK) return rc;
    if (warm && (rc = k->warm_ll.reserve(size_t(n_pairs) * sizeof(double))) != PHB_OK) return rc;
    if ((rc = k->ckpt.reserve(size_t(grid) * (gv->NT / 32) * size_t(gv->ckpt_bytes_per_warp(p.seg_len)))) != PHB_OK) return rc;
    if ((rc = k->gacc.reserve(size_t(grid) * gv->NT * 6 * (gv->M / gv->T) * sizeof(double))) != PHB_OK) return rc;
    if (marked && (rc = k->split.reserve((size_t(2) * S + 2) * sizeof(int32_t))) != PHB_OK) return rc;
    if (!k->dry) {
        // every process contributes zero where it has nothing to say (ll: process 0 only; marked rows: process 0's
        // double launch only)
        PHB_CUDA(cudaMemsetAsync(k->term_ll.ptr, 0, size_t(n_pairs) * sizeof(double), st));
        PHB_CUDA(cudaMemsetAsync(k->term_dlog.ptr, 0, size_t(n_pairs) * C * k->elem(), st));
        phb::TransferArgs ta{};
        phb::KernelArgs &a = ta.k;
        a.data = k->d_data;
        a.pitch = k->pitch;
        a.n_rows = k->N;
        a.L = k->L;
        a.inds = inds;
        a.B = B;
        a.S = S;
        a.params6 = k->term_params.ptr;
        a.pstride_b = C;
        a.pi = static_cast<const char *>(k->term_params.ptr) + size_t(6) * M * k->elem();
        a.pistride_b = C;
        a.ll = static_cast<double *>(k->term_ll.ptr);
        a.dlog = k->term_dlog.ptr;
        a.err_flag = k->d_err;
        a.skip_flag = marked ? k->d_rowflag : nullptr;
        ta.n_seg = p.n_seg;
        ta.seg_len = p.seg_len;
        ta.rows = static_cast<float *>(k->transfer_rows.ptr);   // this process's own operators (phb_hmm_term_sharded_begin)
        ta.row_log2 = static_cast<double *>(k->transfer_log.ptr);
        ta.segs_per_slot = p.per_rank;
        ta.seg_first = seg_lo;
        ta.n_seg_local = n_local;
        const char *base = static_cast<const char *>(gather);
        const float *rank_rows = reinterpret_cast<const float *>(base);
        const double *rank_log2 = reinterpret_cast<const double *>(base + p.rows_bytes);
        int64_t stride_rows = int64_t(p.slot_bytes / sizeof(float)), stride_log = int64_t(p.slot_bytes / sizeof(double));
        int rank_arg = rank, world_arg = world;
        void *bnd_a = k->bnd_alpha.ptr, *bnd_b = k->bnd_beta.ptr;
        void *bargs[] = {&ta, &rank_rows, &rank_log2, &stride_rows, &stride_log, &rank_arg, &world_arg, &bnd_a, &bnd_b};
        PHB_CUDA(cudaLaunchKernel(tv->sharded_boundaries_func, dim3(unsigned((n_pairs * M + 127) / 128)), dim3(128), bargs, 0, st));
        k->launches += 1;
        if (n_local > 0) {
            phb::KernelArgs sa = a;
            sa.ckpt = k->ckpt.ptr;
            sa.gacc = static_cast<double *>(k->gacc.ptr);
            sa.seg_count = p.n_seg;
            sa.seg_len = p.seg_len;
            sa.seg_first = seg_lo;
            sa.seg_local = n_local;
            sa.bnd_alpha = bnd_a;
            sa.bnd_beta = bnd_b;
            sa.seg_dlog = k->seg_dlog.ptr;
            sa.seg_ctas = p.seg_ctas;
            sa.n_groups = n_groups;
            sa.warm_len = warm ? overlap : 0;
            sa.warm_ll = warm ? static_cast<double *>(k->warm_ll.ptr) : nullptr;
            void *kargs[] = {&sa};
            PHB_CUDA(cudaLaunchKernel(gv->func, dim3(unsigned(grid)), dim3(gv->NT), kargs, gv->smem, st));
            const int64_t n_out = n_pairs * C;
            phb::sum_segments_kernel<float><<<unsigned((n_out + 255) / 256), 256, 0, st>>>(
                static_cast<const float *>(k->seg_dlog.ptr), n_pairs, n_local, M, static_cast<float *>(k->term_dlog.ptr), 0, sa);
            PHB_CUDA(cudaGetLastError());
            k->launches += 2;
        }
        snprintf(k->last_name, sizeof k->last_name, "time-sharded %d/%d: transfer_rows_kernel<float,M=%d> + psmc_loglik_kernel<SEG> x %lld of %lld segments%s",
                 rank, world, M, (long long)n_local, (long long)p.n_seg, warm ? " + warm-up term" : "");
    }
    if (rank == 0) {
        // process 0 owns: the log-likelihood (chain_boundaries_kernel wrote it everywhere: the others drop theirs below),
        // the pairs on marked rows (double arithmetic over the whole chunk) and the warm-up term that is subtracted
        phb::KernelArgs a{};
        a.data = k->d_data;
        a.pitch = k->pitch;
        a.n_rows = k->N;
        a.inds = inds;
        a.B = B;
        a.S = S;
        a.params6 = k->term_params.ptr;
        a.pstride_b = C;
        a.pi = static_cast<const char *>(k->term_params.ptr) + size_t(6) * M * k->elem();
        a.pistride_b = C;
        a.ll = static_cast<double *>(k->term_ll.ptr);
        a.dlog = k->term_dlog.ptr;
        if (marked) {
            const Variant *esc = escalation_variant(M);
            int32_t *lists = static_cast<int32_t *>(k->split.ptr);
            int32_t *counts = lists + 2 * S;
            if (!k->dry) {
                phb::split_minibatch_kernel<<<1, 1024, 0, st>>>(inds, S, k->d_rowflag, k->N, lists, counts);
                PHB_CUDA(cudaGetLastError());
                k->launches += 1;
            }
            phb::KernelArgs part = a;
            part.L = k->L;
            part.s_list = lists + S;
            part.s_count = counts + 1;
            if ((rc = launch_one(k, part, true, st, esc)) != PHB_OK) return rc;
        }
        if (overlap > 0 && !warm) {
            a.L = overlap;
            a.out_mode = 1;
            if ((rc = launch(k, a, true, st)) != PHB_OK) return rc;
        }
    }
    if (k->dry) return PHB_OK;
    // per-particle sums; the log-likelihood column comes from process 0 alone
    if (rank != 0) PHB_CUDA(cudaMemsetAsync(k->term_ll.ptr, 0, size_t(n_pairs) * sizeof(double), st));
    phb::sum_over_chunks_kernel<float><<<unsigned(B), 128, 0, st>>>(static_cast<const double *>(k->term_ll.ptr),
                                                                      static_cast<const float *>(k->term_dlog.ptr), S, C, sums);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_sum_over_chunks_device(phb_kernel *k, const double *ll, const void *dlog, int64_t B, int64_t S, double *sums, void *stream) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (!sums || (S > 0 && !ll)) return fail(PHB_E_INVALID, "NULL pointer");
    if (B == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    const int C = 7 * k->M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (k->dbl)
        phb::sum_over_chunks_kernel<double><<<unsigned(B), 128, 0, st>>>(ll, static_cast<const double *>(dlog), S, C, sums);
    else
        phb::sum_over_chunks_kernel<float><<<unsigned(B), 128, 0, st>>>(ll, static_cast<const float *>(dlog), S, C, sums);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    return PHB_OK;
}

int phb_hmm_term_finish_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                               double theta, const double *sums, double weight, double *value, double *grad_x, void *stream) {
    phb::ParamsArgs pa{};
    if (int rc = fill_params_args(k, x, B, epoch_widths, n_epochs, theta, pa)) return rc;
    if (!sums || !value) return fail(PHB_E_INVALID, "NULL pointer");
    if (B == 0) return PHB_OK;
    PHB_CUDA(cudaSetDevice(k->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t stride = 1 + 7 * k->M;
    if (k->dry) return PHB_OK;
    const int threads = 64;
    phb::scaled_first_column_kernel<<<unsigned((B + threads - 1) / threads), threads, 0, st>>>(sums, B, stride, weight, value);
    PHB_CUDA(cudaGetLastError());
    k->launches += 1;
    if (grad_x) {
        pa.cotangent = sums;
        pa.cot_stride = stride;
        pa.scale = weight;
        pa.grad_x = grad_x;
        launch_params_kernel<true>(pa, B * pa.P, st);
        PHB_CUDA(cudaGetLastError());
        k->launches += 1;
    }
    return PHB_OK;
}

int phb_hmm_term_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs, double theta,
                        const int64_t *inds, int64_t S, int64_t overlap, double weight, double *value, double *grad_x,
                        void *stream) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0) return fail(PHB_E_INVALID, "negative batch");
    if (int rc = k->term_sums.reserve(size_t(std::max<int64_t>(B, 1)) * (1 + 7 * k->M) * sizeof(double))) return rc;
    double *sums = static_cast<double *>(k->term_sums.ptr);
    if (int rc = phb_hmm_term_sums_device(k, x, B, epoch_widths, n_epochs, theta, inds, S, overlap, grad_x != nullptr, sums, stream))
        return rc;
    return phb_hmm_term_finish_device(k, x, B, epoch_widths, n_epochs, theta, sums, weight, value, grad_x, stream);
}

int phb_hmm_term_host(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs, double theta,
                      const int64_t *inds, int64_t S, int64_t overlap, double weight, double *value, double *grad_x) {
    if (int rc = check_handle(k)) return rc;
    if (B < 0 || S < 0) return fail(PHB_E_INVALID, "negative batch shape");
    if (B == 0) return PHB_OK;
    if (!x || !value || (S > 0 && !inds)) return fail(PHB_E_INVALID, "NULL pointer");
    for (int64_t s = 0; s < S; ++s)
        if (inds[s] < 0 || inds[s] >= k->N)
            return fail(PHB_E_INVALID, "0 <= inds[%lld]=%lld < N=%lld violated", (long long)s, (long long)inds[s], (long long)k->N);
    const int64_t P = 2 + int64_t(n_epochs) + 1;
    for (int64_t i = 0; i < B * P; ++i)
        if (!std::isfinite(x[i])) return fail(PHB_E_INVALID, "not all particle coordinates finite");
    PHB_CUDA(cudaSetDevice(k->device));
    // device copies: x [B, P] | value [B] | grad_x [B, P] (doubles), then inds [S]
    const size_t n_dbl = size_t(B) * P * 2 + size_t(B);
    if (int rc = k->term_io.reserve(n_dbl * sizeof(double) + size_t(std::max<int64_t>(S, 1)) * sizeof(int64_t))) return rc;
    double *d_x = static_cast<double *>(k->term_io.ptr);
    double *d_value = d_x + B * P;
    double *d_grad = d_value + B;
    int64_t *d_inds = reinterpret_cast<int64_t *>(d_grad + B * P);
    PHB_CUDA(cudaMemcpyAsync(d_x, x, size_t(B) * P * sizeof(double), cudaMemcpyHostToDevice, k->stream));
    if (S > 0) PHB_CUDA(cudaMemcpyAsync(d_inds, inds, size_t(S) * sizeof(int64_t), cudaMemcpyHostToDevice, k->stream));
    if (int rc = phb_hmm_term_device(k, d_x, B, epoch_widths, n_epochs, theta, d_inds, S, overlap, weight, d_value,
                                     grad_x ? d_grad : nullptr, k->stream))
        return rc;
    PHB_CUDA(cudaMemcpyAsync(value, d_value, size_t(B) * sizeof(double), cudaMemcpyDeviceToHost, k->stream));
    if (grad_x) PHB_CUDA(cudaMemcpyAsync(grad_x, d_grad, size_t(B) * P * sizeof(double), cudaMemcpyDeviceToHost, k->stream));
    return phb_sync(k);
}

}  // extern "C"
