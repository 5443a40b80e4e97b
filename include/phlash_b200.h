/* phlash_b200 - B200-native (sm_100a) replacement for the data-parallel hot path of
 * jthlab/phlash: the PSMC coalescent-HMM log-likelihood and its gradient for every
 * (chunk x SVGD particle) pair.
 *
 * This header is the drop-in boundary: a plain C ABI (pointers and sizes, no C++ / torch / jax
 * types).  Each entry point names the reference interface it replaces, relative to the
 * reference tree (jthlab/phlash, src/phlash/...).  INTEGRATION.md shows the binding a phlash
 * maintainer would add on the Python/JAX side.
 *
 * Conventions
 *   M       number of hidden states (time intervals); supported: 4, 8, 16, 32, 64.
 *   FLOAT   float, or double when the kernel was created with double_precision != 0
 *           (reference: `typedef ... FLOAT`, gpu.py:132-136).
 *   params  7 rows of M values in the order b, d, u, v, emis0, emis1, pi
 *           (reference: PSMCParams, params.py:16-23; device layout gpu.py:483-502).
 *   dlog    same 7 x M shape: d ll / d log(theta) for rows b, d, u, v, emis0, emis1 and
 *           pi * d ll / d pi for the pi row; exactly zero where the parameter is zero.  The v row
 *           is returned in its FINAL position (entry j belongs to v[j]): the reference kernel
 *           stores it shifted by one and rolls it back on the host (gpu.py:303-313); that roll is
 *           already applied here.
 *   ll      always double (reference: gpu.py:216, 535, 583).
 *
 * Threading: a kernel object is NOT thread-safe (it owns one stream and one set of scratch buffers);
 * use it from one thread at a time (the reference is driven from XLA's single host-callback
 * thread) or create one object per thread / per device.  Different objects are independent.
 *
 * Error handling: every function returns PHB_OK (0) or a negative PHB_E_* code and never
 * throws; phb_last_error() returns a thread-local message for the last failure
 * (reference: CudaError / AssertionError / MemoryError raised from gpu.py:23-46, 106-124, 197-214).
 */
#ifndef PHLASH_B200_H
#define PHLASH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHB_OK 0
#define PHB_E_INVALID (-1) /* bad argument: shape, M, index out of range, non-finite parameter */
#define PHB_E_CUDA (-2)    /* a CUDA runtime call failed (message has the CUDA error string) */
#define PHB_E_NOMEM (-3)   /* device allocation failed (reference: MemoryError, gpu.py:118-124) */
#define PHB_E_DATA (-4)    /* data matrix violates the reference's checks (gpu.py:106-113) */

typedef struct phb_kernel phb_kernel; /* opaque; owns the device copy of the data, stream, scratch */

/* ABI version of this header (bumped on any signature change). */
int phb_abi_version(void);

/* Thread-local description of the most recent error in this thread ("" if none). */
const char *phb_last_error(void);

/* Number of visible CUDA devices (<0: error code).  Replaces cuDeviceGetCount, gpu.py:373. */
int phb_device_count(void);

/* Create a kernel object on `device` holding its own device copy of the observation matrix.
 * Replaces _PSMCKernelBase.__init__ (gpu.py:104-151): `data` is host int8 [N, L] row-major with
 * values >= -1 (-1 = missing); values > 1 are clipped to 1; a row with no non-missing entry is
 * rejected with PHB_E_DATA. */
int phb_create(int M, const int8_t *data, int64_t N, int64_t L, int double_precision, int device,
               phb_kernel **out);

/* The same for FULL chunks [N, W = overlap + L] as init_mcmc_data returns them (data.py:506-558), i.e. for
 * the fused warm-up entries below.  The reference splits the warm-up columns off BEFORE it builds its kernel
 * object (mcmc.py:203-209), so its "every row has an observation" check (gpu.py:111-113) sees the data part
 * only; here the check covers columns [overlap, W) likewise.  All checks (values >= -1, clipping to <= 1,
 * observed rows) run on the device after the copy, which is staged through pinned buffers by a few host
 * threads: the host never walks the matrix (BASELINE config 5 is 50 GB). */
int phb_create_chunks(int M, const int8_t *chunks, int64_t N, int64_t W, int64_t overlap, int double_precision,
                      int device, phb_kernel **out);

/* Create a kernel object from BINNED CONTIGS instead of ready-made chunks: het is host int8
 * [n_rows, length] (one row per diploid of one contig); the overlapping windows of
 * _chunk_het_matrix (data.py:37-61: ceil(length / W) windows of W = chunk_size + overlap bins per
 * row, starting every chunk_size bins, padded with -1, values clipped to [-1, 1]) are cut on the
 * device.  The object then holds N = n_rows * ceil(length / W) rows of W bins - what
 * Contig.to_chunked + np.concatenate produce (data.py:102-112, 558) - ready for
 * phb_loglik_warmup_* with the same `overlap`.  Windows without a single observation are rejected
 * like in phb_create (gpu.py:111-113). */
int phb_create_from_contig(int M, const int8_t *het, int64_t n_rows, int64_t length, int64_t overlap,
                           int64_t chunk_size, int double_precision, int device, phb_kernel **out);

/* Copy the resident observation matrix back to the host as [N, L] (tests; the reference keeps its
 * chunks on the host). */
int phb_download_data(const phb_kernel *k, int8_t *out);

/* Replaces _PSMCKernelBase.__del__ (gpu.py:153-174).  NULL is allowed. */
void phb_destroy(phb_kernel *k);

/* Introspection (reference: PSMCKernel.M / .double_precision / .float_type, gpu.py:343-357). */
int phb_M(const phb_kernel *k);
int phb_double_precision(const phb_kernel *k);
int64_t phb_num_rows(const phb_kernel *k);
int64_t phb_row_length(const phb_kernel *k);
int phb_device(const phb_kernel *k);

/* Tuning knob, mainly for tests: force the number of threads that cooperate on one
 * (chunk, particle) pair (0 = choose automatically from the number of pairs). */
int phb_set_threads_per_pair(phb_kernel *k, int threads_per_pair);

/* Small minibatches (the reference default is S <= 5 chunks) are bound by the serial depth of the
 * recursion; for them a "store-all" gradient kernel keeps every forward vector in HBM instead of
 * recomputing it (two dependent passes over the chunk instead of three).  mode: -1 = automatic
 * (few pairs and the scratch fits), 0 = never, 1 = whenever available (float kernels; ignored while
 * threads_per_pair is forced). */
int phb_set_store_all(phb_kernel *k, int mode);

/* FEW pairs cannot fill the GPU with one recursion each; such a call lasts L x the latency of one
 * dependent site step whatever the lane layout.  Two cases of the reference are of this kind: the ELPD
 * evaluation (whole un-chunked test contigs for every particle, forward only, mcmc.py:213-238) and the
 * default minibatch for a single genome (S = min(5, N / niter) = 1 chunk, mcmc.py:119-121).  They are
 * evaluated parallel in time: the sequence is cut into segments, the M unit vectors are propagated
 * through every segment (its transfer operator: M x the forward arithmetic, M x segments x the
 * parallelism) and the operators are chained in float64 - forwards for the log-likelihood and the
 * forward vectors at the segment boundaries, backwards for the adjoint vectors there, after which the
 * gradient passes run over all segments independently.  Exact - no burn-in approximation.  For somewhat
 * larger minibatches (the reference's S = 5: 2 500 pairs) the operators cost more than they save; there
 * the boundary vectors come from two plain sequential sweeps running side by side (forward recursion,
 * adjoint recursion without gradient bookkeeping), followed by the same segment passes (any M).
 * mode: -1 = automatic (float objects; operators for M <= 16: forward-only while pairs * M is below a
 * quarter of the resident threads and segments stay >= 4096 sites, gradient while pairs * M is below 0.3
 * of them; else sweeps for gradients while they still split into >= 6 segments of >= 1024 sites),
 * 0 = never, 1 = operators whenever possible, 2 = sweeps whenever possible (ignored while
 * threads_per_pair is forced). */
int phb_set_parallel_in_time(phb_kernel *k, int mode);

/* PRECISION ESCALATION (single-precision objects, gradient path).  Through a long run of identical
 * observations (masked centromere, the -1 padding of a contig's last chunk, a run of homozygosity)
 * fp32 forward / adjoint vectors stagnate at a floating-point fixed point and the transition rows of
 * that chunk's gradient lose up to ~(run length) x 6e-8 relative - a property of any fp32
 * implementation of the recursion, the reference's included.  At creation every row that contains an
 * aligned window of 1024 identical observations is marked; pairs on marked rows are evaluated with
 * double arithmetic (same float buffers, second launch on the same stream).  enabled: 1 (default) / 0.
 * phb_num_escalated_rows reports how many rows are marked (0 for double-precision objects). */
int phb_set_precision_escalation(phb_kernel *k, int enabled);
int64_t phb_num_escalated_rows(const phb_kernel *k);

/* HOST-buffer evaluation; blocking.  Replaces _PSMCKernelBase.__call__ (gpu.py:182-325) for
 * pa of shape [B, S, 7, M]: pair (b, s) scores data row inds[s] with parameter block
 * params[b, s].  ll is [B, S]; dlog is [B, S, 7, M] and may be NULL together with
 * want_grad == 0, which selects the forward-only path (reference: the `loglik` kernel,
 * gpu.py:529-573, but with every pair using its own parameter block).
 * Checks 0 <= inds[s] < N and that all parameters are finite (gpu.py:197-199, 214). */
int phb_loglik_host(phb_kernel *k, const void *params, const int64_t *inds, int64_t B, int64_t S,
                    int want_grad, double *ll, void *dlog);

/* Same evaluation when the 6 rows b, d, u, v, emis0, emis1 are shared by all S chunks of a
 * particle and only pi differs per pair - how the reference actually calls the kernel
 * (model.py:52-57: pps = pp._replace(pi=pis)).  params6 is [B, 6, M], pi is [B, S, M] (or
 * [B, M] when pi_per_pair == 0).  Outputs as above. */
int phb_loglik_shared_host(phb_kernel *k, const void *params6, const void *pi, int pi_per_pair,
                           const int64_t *inds, int64_t B, int64_t S, int want_grad, double *ll,
                           void *dlog);

/* ALLOCATION-FREE, CAPTURABLE STEPS.  phb_reserve sizes every scratch buffer and sets every kernel
 * attribute that an evaluation with `B` particles and up to `S_max` chunks per minibatch can need (both the
 * plain call and the fused warm-up / whole-term entries with this `overlap`), by running the dispatcher
 * "dry" for each minibatch size.  Afterwards such calls neither allocate nor free device memory nor query
 * the driver, so a caller may record them into a CUDA graph (cudaStreamBeginCapture on the stream it passes,
 * or XLA's command buffers); phb_allocation_count reports how many device allocations the object's scratch
 * buffers have made so far (tests: unchanged across calls after phb_reserve). */
int phb_reserve(phb_kernel *k, int64_t B, int64_t S_max, int64_t overlap, int want_grad);
int64_t phb_allocation_count(const phb_kernel *k);

/* MINIBATCH SAMPLING ON THE DEVICE (replaces `inds = jax.random.choice(subkey, N, (S,))` and the host
 * gather `warmup_chunks[inds]`, mcmc.py:277-278: the warm-up bins are resident with the chunks, so the
 * indices are all an iteration needs).  inds[s] = phb_minibatch_indices(seed, it, N, S)[s], S draws WITH
 * replacement from [0, N), where `it` is the object's iteration counter, which the call then advances -
 * a captured graph therefore draws a new minibatch at every replay.  phb_minibatch_indices is the same
 * counter-based generator on the host (pure function: reproducibility, tests). */
int phb_sample_minibatch_device(phb_kernel *k, uint64_t seed, int64_t S, int64_t *inds, void *stream);
int phb_set_iteration(phb_kernel *k, uint64_t iteration, void *stream);
void phb_minibatch_indices(uint64_t seed, uint64_t iteration, int64_t N, int64_t S, int64_t *inds);

/* FP32 FMA throughput of `device` measured now (TFLOP/s, best of 10): independent FFMA chains - the roofline
 * denominator bench.py quotes - and, optionally, the accumulate pattern acc += x * y with three distinct
 * register operands, which is what gradient sums look like to the register file. */
int phb_measure_fp32_peak(int device, double *independent_tflops, double *accumulate_tflops);

/* The kernel object's own (non-blocking) stream, as a cudaStream_t.  The host entries run on it. */
void *phb_stream(const phb_kernel *k);

/* DEVICE-buffer evaluation, asynchronous on `stream` (a cudaStream_t used as given: NULL is the
 * CUDA default stream, as everywhere in CUDA): the entry an XLA FFI custom call binds (replaces the
 * jax.pure_callback round trip, gpu.py:441-465).  All pointers are device pointers valid on
 * the kernel object's device.  params_stride_s == 0 selects the shared-parameter fast path:
 *   params6 + b * params_stride_b + s * params_stride_s  -> [6, M] block of pair (b, s)
 *   pi      + b * pi_stride_b     + s * pi_stride_s      -> [M]    initial distribution
 * (strides in elements of FLOAT).  Index and finiteness checks are done on the device:
 * a violation makes the affected ll entries NaN and is reported by phb_sync(). */
int phb_loglik_device(phb_kernel *k, const void *params6, int64_t params_stride_b,
                      int64_t params_stride_s, const void *pi, int64_t pi_stride_b,
                      int64_t pi_stride_s, const int64_t *inds, int64_t B, int64_t S,
                      int want_grad, double *ll, void *dlog, void *stream);

/* FUSED WARM-UP evaluation (replaces the native-JAX warm-up scan and the per-chunk pi plumbing of
 * model.py:50-57).  The kernel object must have been created on FULL chunks [N, overlap + L]
 * (mcmc.py:203 before the split).  For every pair (b, s):
 *     ll[b, s]   = log p(chunk bins | pi_s)      with pi_s = the filtered distribution after the
 *                  first `overlap` bins started from params7[b, 6, :] (the stationary pi),
 *     dlog[b, s] = its gradient w.r.t. log of the PARTICLE's 7 rows (through the warm-up as well).
 * Both are additive over s, so sum_s is d l2 / d log(theta_b) of model.py:57.
 * Computed as LL(all bins) - LL(first `overlap` bins) by two launches of the same kernel.
 * params7 is [B, 7, M] (device pointer for _device, host pointer for _host). */
int phb_loglik_warmup_device(phb_kernel *k, const void *params7, const int64_t *inds, int64_t B,
                             int64_t S, int64_t overlap, int want_grad, double *ll, void *dlog,
                             void *stream);
int phb_loglik_warmup_host(phb_kernel *k, const void *params7, const int64_t *inds, int64_t B,
                           int64_t S, int64_t overlap, int want_grad, double *ll, void *dlog);

/* HMM PARAMETER CONSTRUCTION on the device, float64 arithmetic (replaces MCMCParams.to_dm,
 * params.py:94-131, and PSMCParams.from_dm, params.py:32-55, with transition.py:9-85 and
 * size_history.py:123-193 behind it).  x is a device pointer to [B, P] doubles, P = 2 + n_epochs + 1:
 * the flattened particle (t_tr[2], c_tr[n_epochs], rho_over_theta_tr; params.py:58-66).
 * epoch_widths (host pointer, n_epochs ints summing to M) is the parsed PSMC pattern
 * (util.py:8-37), theta the static mutation rate.  params7 receives [B, 7, M] FLOAT. */
int phb_params_from_particles(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths,
                              int n_epochs, double theta, void *params7, void *stream);

/* Its vector-Jacobian product: cotangent [B, 7, M] FLOAT holds d l / d log(theta) (e.g. the
 * per-particle sum of dlog over the minibatch); grad_x [B, P] doubles receives d l / d x - what
 * JAX's reverse pass through jnp.log, from_dm and to_dm produces (gpu.py:361, model.py:50-51). */
int phb_params_vjp(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                   double theta, const void *cotangent, double *grad_x, void *stream);

/* THE WHOLE HMM TERM of log_density for all particles (replaces model.py:50-57 with the weights of
 * model.py:71-72 / mcmc.py:240-247, and JAX's reverse pass through it), asynchronous on `stream`, device
 * pointers throughout, no host round trip:
 *     value[b]     = weight * sum_s log p(chunk inds[s] | particle b, warm-up fused)
 *     grad_x[b, :] = d value[b] / d x[b, :]            (x as in phb_params_from_particles)
 * i.e. phb_params_from_particles -> phb_loglik_warmup_device -> sum over the minibatch -> phb_params_vjp
 * on internal scratch.  The kernel object must hold FULL chunks [N, overlap + L].  grad_x == NULL selects
 * the forward-only evaluation (the reference's ELPD, mcmc.py:221-236, is this with weight 1, overlap 1
 * and the whole test contigs as "chunks").  A JAX binding is ONE custom_vjp over the flattened particle:
 * forward = this call, backward = g[b] * grad_x[b, :].
 *
 * For one process per GPU the term is split around the collective: _sums_ leaves the per-particle sums
 * [B, 1 + 7 M] (double: log-likelihood, then d / d log(theta) in params order) of this rank's part of the
 * minibatch in `sums`; the caller all-reduces them (one NCCL call) and _finish_ maps the total to
 * value / grad_x.  S == 0 is allowed (a rank without work contributes zeros). */
int phb_hmm_term_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                        double theta, const int64_t *inds, int64_t S, int64_t overlap, double weight,
                        double *value, double *grad_x, void *stream);
/* The same with HOST buffers, blocking (x, value, grad_x are a few tens of kilobytes: this is the entry a
 * jax.pure_callback binds when no FFI handler is built - the host round trip that costs the reference
 * 2 x 1.1 MB per call, gpu.py:441-465, shrinks to B * (2 P + 1) doubles).  Checks 0 <= inds[s] < N and that
 * x is finite. */
int phb_hmm_term_host(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                      double theta, const int64_t *inds, int64_t S, int64_t overlap, double weight,
                      double *value, double *grad_x);
int phb_hmm_term_sums_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                             double theta, const int64_t *inds, int64_t S, int64_t overlap, int want_grad,
                             double *sums, void *stream);
int phb_hmm_term_finish_device(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths,
                               int n_epochs, double theta, const double *sums, double weight, double *value,
                               double *grad_x, void *stream);

/* TIME-AXIS SHARDING of the whole HMM term over `world` processes (one per GPU) for SMALL minibatches: with
 * S < world chunks the chunk axis cannot occupy every GPU (the reference splits its S <= 5 indices over the
 * devices, gpu.py:398-400, and leaves the rest idle).  The parallel-in-time gradient cuts every chunk into
 * segments that are independent once the boundary vectors are known, so the SEGMENTS are sharded instead:
 *   _plan   n_segments / slot_bytes for this call shape; n_segments == 0 means "does not apply" (double
 *           precision, M > 16, or too many pairs for the operators to pay) - shard the chunks instead;
 *   _begin  particles -> parameters, then this process's slice of the segment transfer operators (kept in the
 *           object) and their PRODUCT, one operator per pair, written into ITS slot (rank * slot_bytes) of
 *           `gather` (world * slot_bytes bytes of device memory; 1 KB per pair and process at M = 16);
 *   -- the caller all-gathers `gather` in place (one NCCL call) --
 *   _end    chains the processes' operators (float64) to the vectors entering and leaving this process's slice,
 *           its own segment operators to the boundary vectors inside the slice, runs the gradient passes over
 *           those segments and leaves its PARTIAL per-particle sums [B, 1 + 7 M] in `sums`; process 0 adds the
 *           log-likelihood, the pairs on marked rows (precision escalation) and subtracts the warm-up term;
 *   -- the caller all-reduces `sums` and calls phb_hmm_term_finish_device, as in the chunk-sharded form.
 * The result equals phb_hmm_term_sums_device on one process up to floating-point summation order. */
int phb_hmm_term_sharded_plan(phb_kernel *k, int64_t B, int64_t S, int64_t overlap, int world, int64_t *n_segments,
                              int64_t *slot_bytes);
int phb_hmm_term_sharded_begin(phb_kernel *k, const double *x, int64_t B, const int32_t *epoch_widths, int n_epochs,
                               double theta, const int64_t *inds, int64_t S, int64_t overlap, int rank, int world,
                               void *gather, void *stream);
int phb_hmm_term_sharded_end(phb_kernel *k, const int64_t *inds, int64_t B, int64_t S, int64_t overlap, int rank, int world,
                             const void *gather, double *sums, void *stream);

/* Per-particle sums over the chunk axis of an evaluation's outputs: ll [B, S] and dlog [B, S, 7, M] (or NULL)
 * -> sums [B, 1 + 7 M] doubles, the buffer one process per GPU all-reduces per step (the summing stage of
 * phb_hmm_term_sums_device on its own; the reference sums in XLA, model.py:57).  Device pointers. */
int phb_sum_over_chunks_device(phb_kernel *k, const double *ll, const void *dlog, int64_t B, int64_t S, double *sums,
                               void *stream);

/* Wait for the kernel object's own stream AND for the stream of the most recent
 * phb_loglik_device call, then report deferred device-side errors. */
int phb_sync(phb_kernel *k);

/* Device pointer to the resident observation matrix and its row pitch in bytes (read-only;
 * for callers that want to fuse their own kernels, and for tests). */
const int8_t *phb_device_data(const phb_kernel *k, int64_t *pitch);

/* Timing of the most recent evaluation on this object, measured with CUDA events on the
 * launching stream around the compute kernel only (milliseconds; <0 if unavailable).
 * Blocks until that kernel has finished. */
float phb_last_kernel_ms(phb_kernel *k);

/* Name of the kernel variant of the most recent evaluation, e.g.
 * "psmc_loglik_kernel<float,MT=16,T=1,K=8,grad,NT=128>" ("" before the first one). */
const char *phb_last_kernel_name(const phb_kernel *k);

/* Number of kernels this library has launched on this object since creation. */
int64_t phb_launch_count(const phb_kernel *k);

#ifdef __cplusplus
}
#endif
#endif /* PHLASH_B200_H */
