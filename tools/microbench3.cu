// Issue rate of FP32 FMA streams at LOW occupancy on B200: how many warps per scheduler does it take to keep
// the FMA pipe busy?  The gradient kernel holds 254 registers per thread, i.e. 2 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

// MODE 0: 16 independent chains a_i = a_i * b + c (two operands shared)
// MODE 1: 16 independent chains with three distinct registers a_i = a_i * b_i + c_i
// MODE 2: accumulate pattern acc_i += x_i * y_i ; x_i += y_i * c
// MODE 3: one adjoint-site-like body: 4 serial chains of 16 + 96 element-wise FMAs that depend on them
template <int MODE> __global__ void probe(float *out, const float *in, float b, float c) {
    float a[16], p[16], q[16], r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = threadIdx.x * 1e-3f + i;
        p[i] = in[i];
        q[i] = in[16 + i];
        r[i] = 0.f;
    }
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], p[i], q[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = fmaf(a[i], p[i], r[i]);
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(p[i], 0.999f, a[i]);
        } else {
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, t0[16], t1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = i, j = 15 - i;
                t0[k] = c0;
                c0 = fmaf(p[k], a[k], c0);
                c1 = fmaf(q[k], a[k], c1);
                t1[j] = c2;
                c2 = fmaf(p[j], a[j], c2);
                c3 += a[j];
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                r[k] = fmaf(a[k], t0[k], r[k]);
                const float nw = fmaf(q[k], t1[k], fmaf(p[k], a[k], t0[k]));
                a[k] = fmaf(nw, 1e-3f, a[k]) + (c1 + c3) * 1e-9f;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 3 with the 32 loop-invariant coefficients in the KERNEL-PARAMETER constant bank instead of registers:
// the FMAs then read two registers and one constant operand
struct Coef {
    float p[16], q[16];
};
__global__ void probe_const(float *out, const float *in, const __grid_constant__ Coef cf) {
    float a[16], r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = threadIdx.x * 1e-3f + i + in[i];
        r[i] = 0.f;
    }
    for (int it = 0; it < kIters; ++it) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, t0[16], t1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = i, j = 15 - i;
            t0[k] = c0;
            c0 = fmaf(cf.p[k], a[k], c0);
            c1 = fmaf(cf.q[k], a[k], c1);
            t1[j] = c2;
            c2 = fmaf(cf.p[j], a[j], c2);
            c3 += a[j];
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            r[k] = fmaf(a[k], t0[k], r[k]);
            const float nw = fmaf(cf.q[k], t1[k], fmaf(cf.p[k], a[k], t0[k]));
            a[k] = fmaf(nw, 1e-3f, a[k]) + (c1 + c3) * 1e-9f;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ... and in __constant__ memory indexed by a CTA-uniform slot: ptxas keeps them in UNIFORM registers
// (LDCU ... c[0x3][UR + imm] hoisted out of the loop, FFMA R, R, UR, R inside)
struct Coef4 {
    float p[16], q[16], r[16], s[16];
};
__constant__ Coef4 g_coef[64];
__global__ void probe_uniform(float *out, const float *in) {
    const Coef4 &cf = g_coef[blockIdx.x & 63];
    float a[16], r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = threadIdx.x * 1e-3f + i + in[i];
        r[i] = 0.f;
    }
    for (int it = 0; it < kIters; ++it) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, t0[16], t1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = i, j = 15 - i;
            t0[k] = c0;
            c0 = fmaf(cf.p[k], a[k], c0);
            c1 = fmaf(cf.q[k], a[k], c1);
            t1[j] = c2;
            c2 = fmaf(cf.r[j], a[j], c2);
            c3 += a[j];
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            r[k] = fmaf(a[k], t0[k], r[k]);
            const float nw = fmaf(cf.s[k], t1[k], fmaf(cf.p[k], a[k], t0[k]));
            a[k] = fmaf(nw, 1e-3f, a[k]) + (c1 + c3) * 1e-9f;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
double run_uniform(int warps_per_scheduler, int sms, float *out, const float *in) {
    const int ctas_per_sm = warps_per_scheduler;
    const size_t smem = (size_t(227) * 1024) / ctas_per_sm - 2048;
    cudaFuncSetAttribute(probe_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    static Coef4 h[64];
    for (int c = 0; c < 64; ++c)
        for (int i = 0; i < 16; ++i) { h[c].p[i] = 0.5f + 1e-3f * i; h[c].q[i] = 0.516f + 1e-3f * i; h[c].r[i] = 0.5f + 2e-3f * i; h[c].s[i] = 0.51f; }
    cudaMemcpyToSymbol(g_coef, h, sizeof h);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        probe_uniform<<<sms * ctas_per_sm, 128, smem>>>(out, in);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    const double warp_instr = double(sms) * ctas_per_sm * 4 * kIters * 128.0;
    return warp_instr / (best * 1e-3) / (double(sms) * 4 * 1.965e9);
}

double run_const(int warps_per_scheduler, int sms, float *out, const float *in) {
    const int ctas_per_sm = warps_per_scheduler;
    const size_t smem = (size_t(227) * 1024) / ctas_per_sm - 2048;
    cudaFuncSetAttribute(probe_const, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    Coef cf;
    for (int i = 0; i < 16; ++i) { cf.p[i] = 0.5f + 1e-3f * i; cf.q[i] = 0.516f + 1e-3f * i; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        probe_const<<<sms * ctas_per_sm, 128, smem>>>(out, in, cf);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    const double warp_instr = double(sms) * ctas_per_sm * 4 * kIters * 128.0;
    return warp_instr / (best * 1e-3) / (double(sms) * 4 * 1.965e9);
}

template <int MODE> double run(int warps_per_scheduler, int sms, float *out, const float *in, int fma_per_iter) {
    // one CTA of 128 threads = 1 warp per scheduler; occupancy is set by the number of CTAs per SM, pinned with
    // dynamic shared memory
    const int ctas_per_sm = warps_per_scheduler;
    const size_t smem = (size_t(227) * 1024) / ctas_per_sm - 2048;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<sms * ctas_per_sm, 128, smem>>>(out, in, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    const double warp_instr = double(sms) * ctas_per_sm * 4 * kIters * fma_per_iter;
    return warp_instr / (best * 1e-3) / (double(sms) * 4 * 1.965e9);  // FMA-pipe instructions per scheduler per clock
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float *out, *in;
    cudaMalloc(&out, sizeof(float) * 128 * sms * 8);
    cudaMalloc(&in, sizeof(float) * 64);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 1e-3f * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    printf("{\"unit\": \"FMA-pipe warp-instructions per scheduler per clock at 1965 MHz\",\n");
    const char *names[4] = {"independent_shared_operands", "independent_distinct_operands", "accumulate_pattern", "site_like"};
    for (int mode = 0; mode < 4; ++mode) {
        printf(" \"%s\": {", names[mode]);
        for (int w = 1; w <= 6; ++w) {
            double v = 0;
            if (mode == 0) v = run<0>(w, sms, out, in, 16);
            if (mode == 1) v = run<1>(w, sms, out, in, 16);
            if (mode == 2) v = run<2>(w, sms, out, in, 32);
            if (mode == 3) v = run<3>(w, sms, out, in, 16 * 4 + 16 * 4);
            printf("\"%d_warps\": %.3f%s", w, v, w < 6 ? ", " : "");
        }
        printf("},\n");
    }
    printf(" \"site_like_constant_bank_coefficients\": {");
    for (int w = 1; w <= 6; ++w) printf("\"%d_warps\": %.3f%s", w, run_const(w, sms, out, in), w < 6 ? ", " : "");
    printf("},\n \"site_like_uniform_register_coefficients\": {");
    for (int w = 1; w <= 6; ++w) printf("\"%d_warps\": %.3f%s", w, run_uniform(w, sms, out, in), w < 6 ? ", " : "");
    printf("}}\n");
    return 0;
}
