// Support kernels around the likelihood kernels:
//   fixup_rows_kernel          validation + clipping of an uploaded slab of the observation matrix
//                              (reference: the constructor checks of src/phlash/gpu.py:106-113, done on the
//                              host there; here on the device, after the copy, so that a 50 GB matrix is
//                              not walked by the host)
//   sample_minibatch_kernel    inds ~ choice(N, (S,)) WITH replacement on the device, counter based
//                              (reference: jax.random.choice on the host every iteration, mcmc.py:277)
//   ffma_peak_kernel*          independent / accumulating FFMA chains: the FP32 roofline denominator,
//                              measured in the same process as the benchmark (tools/microbench.cu has the
//                              full set of pipe-rate probes)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace phb {

// ---- observation matrix: validate, clip, pad --------------------------------------------------
// One CTA per row (grid-stride).  For row r of the slab, columns [0, L): a value < -1 records the
// smallest offending linear index (global_row * L + column) in *first_bad; values > 1 become 1
// (gpu.py:108-110); row_observed[global_row] = 1 when some column >= check_from holds an observation
// (gpu.py:111-113; check_from = overlap for full chunks, whose data part is what the reference checks,
// mcmc.py:203).  Columns [L, pitch) are set to -1.
__global__ void fixup_rows_kernel(int8_t *__restrict__ data, int64_t n_rows, int64_t L, int64_t pitch, int64_t check_from,
                                  int64_t first_row, unsigned long long *__restrict__ first_bad, uint8_t *__restrict__ row_observed) {
    __shared__ int any_obs;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        if (threadIdx.x == 0) any_obs = 0;
        __syncthreads();
        uint4 *row = reinterpret_cast<uint4 *>(data + r * pitch);
        bool seen = false;
        for (int64_t q = threadIdx.x; q * 16 < pitch; q += blockDim.x) {
            uint4 w = row[q];
            const int64_t c0 = q * 16;
            uint32_t *words = reinterpret_cast<uint32_t *>(&w);
            const bool interior = c0 >= check_from && c0 + 16 <= L;
            if (interior) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t x = words[i];
                    if (__vcmplts4(x, 0xffffffffu)) {  // some byte < -1
                        for (int j = 0; j < 4; ++j)
                            if (int8_t(x >> (8 * j)) < -1)
                                atomicMin(first_bad, (unsigned long long)((first_row + r) * L + c0 + 4 * i + j));
                    }
                    seen |= __vcmpges4(x, 0u) != 0;   // some byte >= 0
                    words[i] = __vmins4(x, 0x01010101u);
                }
            } else {
                int8_t *bytes = reinterpret_cast<int8_t *>(&w);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int64_t c = c0 + j;
                    if (c >= L) {
                        bytes[j] = -1;
                    } else {
                        if (bytes[j] < -1) atomicMin(first_bad, (unsigned long long)((first_row + r) * L + c));
                        if (bytes[j] > 1) bytes[j] = 1;
                        seen |= c >= check_from && bytes[j] > -1;
                    }
                }
            }
            row[q] = w;
        }
        if (seen) any_obs = 1;  // benign race: all writers store 1
        __syncthreads();
        if (threadIdx.x == 0) row_observed[first_row + r] = uint8_t(any_obs);
        __syncthreads();
    }
}

// first row whose flag is 0 (n if none) -> *out, by atomicMin; launch with enough threads to cover n
__global__ void first_zero_kernel(const uint8_t *__restrict__ flags, int64_t n, unsigned long long *__restrict__ out) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        if (flags[i] == 0) atomicMin(out, (unsigned long long)i);
}

// number of non-zero flags -> *out
__global__ void count_flags_kernel(const uint8_t *__restrict__ flags, int64_t n, unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) c += flags[i] != 0;
    if (c) atomicAdd(out, c);
}

// ---- minibatch sampling -------------------------------------------------------------------------
// Counter-based generator: index s of iteration `it` under `seed` is a pure function (splitmix64
// finaliser of a Weyl sequence), mapped to [0, N) by the high half of a 64 x 64-bit product.  The same
// function runs on the host (phb_minibatch_indices) so that a run is reproducible from (seed, it).
__host__ __device__ inline uint64_t minibatch_hash(uint64_t seed, uint64_t it, uint64_t s) {
    uint64_t z = (seed ^ (it * 0xD1342543DE82EF95ull)) + (s + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline int64_t minibatch_index(uint64_t seed, uint64_t it, uint64_t s, int64_t n) {
    const uint64_t h = minibatch_hash(seed, it, s);
#ifdef __CUDA_ARCH__
    return int64_t(__umul64hi(h, uint64_t(n)));
#else
    return int64_t((unsigned __int128)h * (unsigned __int128)uint64_t(n) >> 64);
#endif
}
// One CTA.  Reads the iteration counter, writes inds[0 .. S), then advances the counter - so that a
// captured CUDA graph draws a NEW minibatch at every replay.
__global__ void sample_minibatch_kernel(uint64_t seed, unsigned long long *__restrict__ iteration, int64_t n, int64_t S,
                                        int64_t *__restrict__ inds) {
    const unsigned long long it = *iteration;
    for (int64_t s = threadIdx.x; s < S; s += blockDim.x) inds[s] = minibatch_index(seed, it, uint64_t(s), n);
    __syncthreads();
    if (threadIdx.x == 0) *iteration = it + 1;
}

__global__ void set_counter_kernel(unsigned long long *counter, unsigned long long value) { *counter = value; }

// ---- FP32 FMA peak ----------------------------------------------------------------------------------
constexpr int kPeakIters = 4096;
constexpr int kPeakChains = 16;
// a_i = a_i * b + c: 16 independent chains per thread, two operands shared (the most the pipe gives)
__global__ void ffma_peak_kernel(float *out, float b, float c) {
    float a[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// acc_i += x_i * y_i with three distinct register operands: the operand pattern of gradient sums
__global__ void ffma_accumulate_kernel(float *out, const float *in) {
    float acc[kPeakChains], x[kPeakChains], y[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) {
        acc[i] = 0.f;
        x[i] = in[i] + threadIdx.x * 1e-6f;
        y[i] = in[kPeakChains + i];
    }
    for (int it = 0; it < kPeakIters / 2; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) x[i] = fmaf(y[i], 0.999f, x[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += acc[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace phb
