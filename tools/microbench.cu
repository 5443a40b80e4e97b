// Pipe-rate microbenchmarks for B200 (sm_100a): the FP32 FMA peak this project's roofline is
// quoted against is not in MEASURED_PEAKS.json, so it is measured here, together with the rates
// that decide the lane layout of the HMM kernel (warp shuffles, broadcast shared-memory loads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Output: one JSON object on stdout.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x)                                                                  \
    do {                                                                          \
        cudaError_t e = (x);                                                      \
        if (e != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));               \
            return 1;                                                             \
        }                                                                         \
    } while (0)

__device__ __forceinline__ float4 lds128(const void *p) {
    float4 v;
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

constexpr int kIters = 4096;
constexpr int kChains = 16;

// a_i = a_i * b + c : b, c shared by all chains (2 of 3 operands repeat; reuse cache can help)
__global__ void ffma_shared_operands(float *out, float b, float c) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// a_i = a_i * b_i + c_i : three distinct registers per FMA
__global__ void ffma_distinct_operands(float *out, const float *in) {
    float a[kChains], b[kChains], c[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 1e-3f + i;
        b[i] = in[i];
        c[i] = in[kChains + i];
    }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b[i], c[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// acc_i += x_i * y_i with x, y changing: the accumulate pattern of the gradient sums
__global__ void ffma_accumulate(float *out, const float *in) {
    float acc[kChains], x[kChains], y[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        acc[i] = 0.f;
        x[i] = in[i] + threadIdx.x * 1e-6f;
        y[i] = in[kChains + i];
    }
    for (int it = 0; it < kIters / 2; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
#pragma unroll
        for (int i = 0; i < kChains; ++i) x[i] = fmaf(y[i], 0.999f, x[i]);  // immediate-operand form
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed f32x2 FMA (Blackwell FFMA2): 16 independent float2 chains
__global__ void ffma2_shared_operands(float *out, float b, float c) {
    float2 a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 bb = make_float2(b, b * 0.5f), cc = make_float2(c, c * 2.f);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = __ffma2_rn(a[i], bb, cc);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same number of FMAs as ffma2_shared_operands but mixed 1:1 with scalar FFMA chains:
// does FFMA2 leave issue slots for other FMA-pipe instructions?
__global__ void ffma2_plus_scalar(float *out, float b, float c) {
    float2 a[8];
    float s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        s1[i] = threadIdx.x * 3e-3f + i;
    }
    const float2 bb = make_float2(b, b * 0.5f), cc = make_float2(c, c * 2.f);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = __ffma2_rn(a[i], bb, cc);
            s1[i] = fmaf(s1[i], b, c);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + s1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- single-warp probes (one warp on an otherwise idle SM), timed with clock64(): what bounds a
// latency-bound kernel with at most one warp per SM sub-partition (small minibatches, ELPD)
template <int MODE> __global__ void single_warp_probe(float *out, long long *cycles, float b, float c) {
    float a[kChains];
    float2 a2[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 1e-3f + i;
        a2[i] = make_float2(a[i], a[i] * 0.5f);
    }
    const float2 bb = make_float2(b, b * 0.5f), cc = make_float2(c, c * 2.f);
    int ia = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {  // 16 independent FFMA
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {  // 16 independent FFMA2
#pragma unroll
            for (int i = 0; i < kChains; ++i) a2[i] = __ffma2_rn(a2[i], bb, cc);
        } else if (MODE == 2) {  // 16 dependent FFMA
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[0] = fmaf(a[0], b, c);
        } else if (MODE == 3) {  // 16 dependent SHFL
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[0] = __shfl_xor_sync(0xffffffffu, a[0], 1);
        } else if (MODE == 4) {  // 16 dependent (SHFL + FFMA)
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[0] = fmaf(__shfl_xor_sync(0xffffffffu, a[0], 1), b, c);
        } else if (MODE == 5) {  // 8 independent FFMA interleaved with 8 independent integer ops (ALU pipe)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fmaf(a[i], b, c);
                ia = (ia ^ (ia << 1)) + i;
            }
        } else if (MODE == 6) {  // 16 dependent FFMA2
#pragma unroll
            for (int i = 0; i < kChains; ++i) a2[0] = __ffma2_rn(a2[0], bb, cc);
        } else if (MODE == 7) {  // 16 dependent MUFU.RCP
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
        }
    }
    const long long t1 = clock64();
    float s = ia;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + a2[i].x + a2[i].y;
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[MODE] = t1 - t0;
}

__global__ void shfl_rate(float *out) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS.128 (all lanes read the same 16 bytes) feeding FMAs
__global__ void lds_broadcast_rate(float *out) {
    __shared__ float4 tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = make_float4(1.f, 0.5f, 0.25f, 0.125f);
    __syncthreads();
    float acc[4] = {0, 0, 0, 0};
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 v = lds128(&tab[(it + i) & 63]);
            acc[0] += v.x;
            acc[1] += v.y;
            acc[2] += v.z;
            acc[3] += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

// per-lane LDS.128 (each lane its own 16 bytes, conflict-free)
__global__ void lds_private_rate(float *out) {
    extern __shared__ float4 buf[];
    for (int i = threadIdx.x; i < 8 * blockDim.x; i += blockDim.x) buf[i] = make_float4(1.f, 0.5f, 0.25f, 0.125f);
    __syncthreads();
    float acc[4] = {0, 0, 0, 0};
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = lds128(&buf[i * blockDim.x + threadIdx.x]);
            acc[0] += v.x;
            acc[1] += v.y;
            acc[2] += v.z;
            acc[3] += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <typename Launch> float time_ms(Launch launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int threads = 256, ctas = sms * 8;
    float *out, *in;
    CHECK(cudaMalloc(&out, sizeof(float) * threads * ctas));
    CHECK(cudaMalloc(&in, sizeof(float) * 64));
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 1e-3f * i;
    CHECK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
    const double lanes = double(threads) * ctas;
    const double fma_ops = lanes * kIters * kChains;
    float t1 = time_ms([&] { ffma_shared_operands<<<ctas, threads>>>(out, 0.999f, 1e-3f); }, 10);
    float t2 = time_ms([&] { ffma_distinct_operands<<<ctas, threads>>>(out, in); }, 10);
    float t3 = time_ms([&] { ffma_accumulate<<<ctas, threads>>>(out, in); }, 10);
    float t7 = time_ms([&] { ffma2_shared_operands<<<ctas, threads>>>(out, 0.999f, 1e-3f); }, 10);
    float t8 = time_ms([&] { ffma2_plus_scalar<<<ctas, threads>>>(out, 0.999f, 1e-3f); }, 10);
    float t4 = time_ms([&] { shfl_rate<<<ctas, threads>>>(out); }, 10);
    float t5 = time_ms([&] { lds_broadcast_rate<<<ctas, threads>>>(out); }, 10);
    CHECK(cudaFuncSetAttribute(lds_private_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * threads * 16));
    float t6 = time_ms([&] { lds_private_rate<<<ctas, threads, 8 * threads * 16>>>(out); }, 10);
    long long *cyc;
    CHECK(cudaMalloc(&cyc, sizeof(long long) * 8));
    single_warp_probe<0><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<1><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<2><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<3><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<4><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<5><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<6><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    single_warp_probe<7><<<1, 32>>>(out, cyc, 0.999f, 1e-3f);
    long long hc[8];
    CHECK(cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost));
    const double per = 1.0 / (double(kIters) * kChains);
    CHECK(cudaGetLastError());
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d,\n", prop.name, sms, clock_khz);
    printf(" \"ffma_shared_operands_tflops\": %.2f,\n", 2 * fma_ops / (t1 * 1e-3) / 1e12);
    printf(" \"ffma_distinct_operands_tflops\": %.2f,\n", 2 * fma_ops / (t2 * 1e-3) / 1e12);
    printf(" \"ffma_accumulate_mix_tflops\": %.2f,\n", 2 * fma_ops / (t3 * 1e-3) / 1e12);
    printf(" \"ffma2_packed_tflops\": %.2f,\n", 2 * 2 * fma_ops / (t7 * 1e-3) / 1e12);
    printf(" \"ffma2_plus_scalar_tflops\": %.2f,\n", 2 * (lanes * kIters * 8 * 3) / (t8 * 1e-3) / 1e12);
    printf(" \"shfl_warp_instr_per_sec\": %.4g,\n", lanes / 32 * kIters * 8 / (t4 * 1e-3));
    printf(" \"shfl_warp_instr_per_clk_per_sm_at_1965MHz\": %.3f,\n", lanes / 32 * kIters * 8 / (t4 * 1e-3) / sms / 1.965e9);
    printf(" \"lds128_broadcast_warp_instr_per_clk_per_sm_at_1965MHz\": %.3f,\n", lanes / 32 * kIters * 16 / (t5 * 1e-3) / sms / 1.965e9);
    printf(" \"lds128_private_warp_instr_per_clk_per_sm_at_1965MHz\": %.3f,\n", lanes / 32 * kIters * 8 / (t6 * 1e-3) / sms / 1.965e9);
    printf(" \"single_warp_cycles_per_instr\": {\"ffma_independent\": %.2f, \"ffma2_independent\": %.2f, \"ffma_dependent\": %.2f, "
           "\"shfl_dependent\": %.2f, \"shfl_plus_ffma_dependent\": %.2f, \"ffma_and_alu_interleaved_per_pair\": %.2f, "
           "\"ffma2_dependent\": %.2f, \"mufu_rcp_dependent\": %.2f}}\n",
           hc[0] * per, hc[1] * per, hc[2] * per, hc[3] * per, hc[4] * per, hc[5] * per * 2, hc[6] * per, hc[7] * per);
    return 0;
}
