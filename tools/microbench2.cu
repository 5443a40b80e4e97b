// Register-operand patterns of packed (f32x2) vs scalar FMA on B200: does FFMA2 relieve the register-read
// limit of accumulate-heavy code?  (tools/microbench.cu measured 71.7 TF/s for scalar FFMA with shared
// operands, 63.7 with three distinct registers, 51.1 for the accumulate pattern, 65.4 for FFMA2 with shared
// operands.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 16;

// scalar: acc_i += x_i * y_i ; x_i += y_i * c   (the accumulate mix of microbench.cu)
__global__ void scalar_accumulate(float *out, const float *in) {
    float acc[kChains], x[kChains], y[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { acc[i] = 0.f; x[i] = in[i] + threadIdx.x * 1e-6f; y[i] = in[kChains + i]; }
    for (int it = 0; it < kIters / 2; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
#pragma unroll
        for (int i = 0; i < kChains; ++i) x[i] = fmaf(y[i], 0.999f, x[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed: the same arithmetic on 8 float2 chains (same number of scalar FMAs)
__global__ void packed_accumulate(float *out, const float *in) {
    float2 acc[kChains / 2], x[kChains / 2], y[kChains / 2];
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) {
        acc[i] = make_float2(0.f, 0.f);
        x[i] = make_float2(in[2 * i] + threadIdx.x * 1e-6f, in[2 * i + 1] + threadIdx.x * 1e-6f);
        y[i] = make_float2(in[kChains + 2 * i], in[kChains + 2 * i + 1]);
    }
    const float2 c = make_float2(0.999f, 0.999f);
    for (int it = 0; it < kIters / 2; ++it) {
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) x[i] = __ffma2_rn(y[i], c, x[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) s += acc[i].x + acc[i].y + x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// three distinct register operands, scalar vs packed:  a_i = a_i * b_i + c_i
__global__ void scalar_distinct(float *out, const float *in) {
    float a[kChains], b[kChains], c[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = in[i]; c[i] = in[kChains + i]; }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b[i], c[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packed_distinct(float *out, const float *in) {
    float2 a[kChains / 2], b[kChains / 2], c[kChains / 2];
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) {
        a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        b[i] = make_float2(in[2 * i], in[2 * i + 1]);
        c[i] = make_float2(in[kChains + 2 * i], in[kChains + 2 * i + 1]);
    }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) a[i] = __ffma2_rn(a[i], b[i], c[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the shape of one adjoint site step: 4 serial chains of 16 (scalar, dependent) + 6 x 16 independent
// element-wise FMAs, the element-wise part scalar (MODE 0) or packed (MODE 1)
template <int MODE> __global__ void site_like(float *out, const float *in) {
    float p[4][16], x[16], w[16], acc[4][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        x[k] = in[k] + threadIdx.x * 1e-6f;
        w[k] = in[16 + k];
#pragma unroll
        for (int r = 0; r < 4; ++r) { p[r][k] = in[32 + k] * (0.5f + 0.1f * r); acc[r][k] = 0.f; }
    }
    for (int it = 0; it < kIters / 8; ++it) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
        float t0[16], t1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = i, j = 15 - i;
            t0[k] = c0;
            c0 = fmaf(p[0][k], w[k], c0);
            c1 = fmaf(p[1][k], x[k], c1);
            t1[j] = c2;
            c2 = fmaf(p[2][j], w[j], c2);
            c3 += x[j];
        }
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                acc[0][k] = fmaf(x[k], w[k], acc[0][k]);
                acc[1][k] = fmaf(t0[k], w[k], acc[1][k]);
                acc[2][k] = fmaf(x[k], t1[k], acc[2][k]);
                acc[3][k] = fmaf(t1[k], w[k], acc[3][k]);
                const float nw = fmaf(p[3][k], t1[k], fmaf(p[1][k], w[k], t0[k]));
                x[k] = fmaf(nw, 1e-3f, x[k]);
                w[k] = nw * 0.25f + (c1 + c3) * 1e-9f;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const float2 x2 = make_float2(x[k], x[k + 1]), w2 = make_float2(w[k], w[k + 1]);
                const float2 t02 = make_float2(t0[k], t0[k + 1]), t12 = make_float2(t1[k], t1[k + 1]);
                float2 a;
                a = __ffma2_rn(x2, w2, make_float2(acc[0][k], acc[0][k + 1])); acc[0][k] = a.x; acc[0][k + 1] = a.y;
                a = __ffma2_rn(t02, w2, make_float2(acc[1][k], acc[1][k + 1])); acc[1][k] = a.x; acc[1][k + 1] = a.y;
                a = __ffma2_rn(x2, t12, make_float2(acc[2][k], acc[2][k + 1])); acc[2][k] = a.x; acc[2][k + 1] = a.y;
                a = __ffma2_rn(t12, w2, make_float2(acc[3][k], acc[3][k + 1])); acc[3][k] = a.x; acc[3][k + 1] = a.y;
                const float2 nw = __ffma2_rn(make_float2(p[3][k], p[3][k + 1]), t12,
                                             __ffma2_rn(make_float2(p[1][k], p[1][k + 1]), w2, t02));
                const float2 nx = __ffma2_rn(nw, make_float2(1e-3f, 1e-3f), x2);
                const float s = (c1 + c3) * 1e-9f;
                const float2 nw2 = __ffma2_rn(nw, make_float2(0.25f, 0.25f), make_float2(s, s));
                x[k] = nx.x; x[k + 1] = nx.y;
                w[k] = nw2.x; w[k + 1] = nw2.y;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k] + w[k] + acc[0][k] + acc[1][k] + acc[2][k] + acc[3][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float *out, *in;
    cudaMalloc(&out, sizeof(float) * 256 * sms * 8);
    cudaMalloc(&in, sizeof(float) * 64);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 1e-3f * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int threads = 256, ctas = sms * 8;
    const double fma_ops = double(threads) * ctas * kIters * kChains;
    const float t1 = time_ms([&] { scalar_accumulate<<<ctas, threads>>>(out, in); }, 10);
    const float t2 = time_ms([&] { packed_accumulate<<<ctas, threads>>>(out, in); }, 10);
    const float t3 = time_ms([&] { scalar_distinct<<<ctas, threads>>>(out, in); }, 10);
    const float t4 = time_ms([&] { packed_distinct<<<ctas, threads>>>(out, in); }, 10);
    // site-like: 2 CTAs of 128 threads per SM (the occupancy of the gradient kernel), FMA-pipe instructions
    // per iteration in scalar form: chains 4 * 16 + element-wise 8 * 16 = 192
    const int st = 128, sc = sms * 2;
    const double site_ops = double(st) * sc * (kIters / 8) * 192.0;
    const float t5 = time_ms([&] { site_like<0><<<sc, st>>>(out, in); }, 10);
    const float t6 = time_ms([&] { site_like<1><<<sc, st>>>(out, in); }, 10);
    printf("{\"scalar_accumulate_tflops\": %.2f, \"packed_accumulate_tflops\": %.2f, \"scalar_distinct_tflops\": %.2f, "
           "\"packed_distinct_tflops\": %.2f, \"site_like_scalar_gfma_per_s\": %.1f, \"site_like_packed_gfma_per_s\": %.1f, "
           "\"site_like_scalar_ms\": %.3f, \"site_like_packed_ms\": %.3f}\n",
           2 * fma_ops / (t1 * 1e-3) / 1e12, 2 * fma_ops / (t2 * 1e-3) / 1e12, 2 * fma_ops / (t3 * 1e-3) / 1e12,
           2 * fma_ops / (t4 * 1e-3) / 1e12, site_ops / (t5 * 1e-3) / 1e9, site_ops / (t6 * 1e-3) / 1e9, t5, t6);
    return 0;
}
