// microbench.cu -- issue-rate probes on B200 for the instruction mix of the PairHMM kernel.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Each kernel runs ITER iterations of an unrolled body of independent chains; reports warp-instr/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 16384
#define CHK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e));return 1;}}while(0)

template <int ILP> __global__ void k_ffma(float *out, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct register sources per FFMA (x = x*y + z with y,z registers varying per chain)
template <int ILP> __global__ void k_ffma3(float *out, const float *in) {
    float x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = in[i] + threadIdx.x; y[i] = in[i + 32] + threadIdx.x * 1e-9f; z[i] = in[i + 64] + threadIdx.x * 1e-9f; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], y[i], z[(i + 1) % ILP]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2 register sources + 1 constant-bank operand (kernel parameter): x = x*c + z
template <int ILP> __global__ void k_ffma2r1c(float *out, const float *in, float c0, float c1, float c2, float c3) {
    float x[ILP], z[ILP];
    const float cs[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = in[i] + threadIdx.x; z[i] = in[i + 64] + threadIdx.x * 1e-9f; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], cs[i & 3], z[(i + 1) % ILP]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the PairHMM cell mix with constant-bank coefficients: per "cell" FMUL(c,r) FFMA(c,r,r) FFMA(c,r,r) FMUL(r,r) FFMA(c,r,r) FFMA(c,r,r)
template <int ILP> __global__ void k_cellmix_const(float *out, const float *in, float ca, float cb, float cc, float cd, float cg) {
    float M[ILP], I[ILP], D[ILP], pr[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { M[i] = in[i] + threadIdx.x; I[i] = in[i + 8]; D[i] = in[i + 16] + threadIdx.x * 1e-9f; pr[i] = in[i + 32] + threadIdx.x * 1e-9f; }
    for (int it = 0; it < ITER; ++it) {
        float Mn[ILP];
#pragma unroll
        for (int i = 1; i < ILP; ++i) { float u = cc * D[i - 1]; u = fmaf(cb, I[i - 1], u); u = fmaf(ca, M[i - 1], u); Mn[i] = pr[i] * u; }
        { float u = cc * D[ILP - 1]; u = fmaf(cb, I[ILP - 1], u); u = fmaf(ca, M[ILP - 1], u); Mn[0] = pr[0] * u; }
#pragma unroll
        for (int i = 0; i < ILP; ++i) D[i] = fmaf(cd, D[i], M[i]);
#pragma unroll
        for (int i = 1; i < ILP; ++i) I[i] = fmaf(cg, I[i - 1], Mn[i - 1]);
        I[0] = fmaf(cg, I[ILP - 1], Mn[ILP - 1]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) M[i] = Mn[i];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += M[i] + I[i] + D[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same mix with per-thread register coefficients (what the current kernel does)
template <int ILP> __global__ void k_cellmix_reg(float *out, const float *in) {
    float M[ILP], I[ILP], D[ILP], pr[ILP], ca[ILP], cb[ILP], cc[ILP], cd[ILP], cg[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { M[i] = in[i] + threadIdx.x; I[i] = in[i + 8]; D[i] = in[i + 16] + threadIdx.x * 1e-9f; pr[i] = in[i + 32] + threadIdx.x * 1e-9f;
        ca[i] = in[i + 40] + threadIdx.x * 1e-9f; cb[i] = in[i + 48] + threadIdx.x * 1e-9f; cc[i] = in[i + 56] + threadIdx.x * 1e-9f; cd[i] = in[i + 64] + threadIdx.x * 1e-9f; cg[i] = in[i + 72] + threadIdx.x * 1e-9f; }
    for (int it = 0; it < ITER; ++it) {
        float Mn[ILP];
#pragma unroll
        for (int i = 1; i < ILP; ++i) { float u = cc[i] * D[i - 1]; u = fmaf(cb[i], I[i - 1], u); u = fmaf(ca[i], M[i - 1], u); Mn[i] = pr[i] * u; }
        { float u = cc[0] * D[ILP - 1]; u = fmaf(cb[0], I[ILP - 1], u); u = fmaf(ca[0], M[ILP - 1], u); Mn[0] = pr[0] * u; }
#pragma unroll
        for (int i = 0; i < ILP; ++i) D[i] = fmaf(cd[i], D[i], M[i]);
#pragma unroll
        for (int i = 1; i < ILP; ++i) I[i] = fmaf(cg[i], I[i - 1], Mn[i - 1]);
        I[0] = fmaf(cg[0], I[ILP - 1], Mn[ILP - 1]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) M[i] = Mn[i];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += M[i] + I[i] + D[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_ffma2(float *out, const float *in) {
    unsigned long long x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 a = make_float2(in[i] + threadIdx.x, in[i + 1]), b = make_float2(in[i + 32] + threadIdx.x * 1e-9f, in[i + 33]), c = make_float2(in[i + 64] + threadIdx.x * 1e-9f, in[i + 65]);
        x[i] = *reinterpret_cast<unsigned long long *>(&a); y[i] = *reinterpret_cast<unsigned long long *>(&b); z[i] = *reinterpret_cast<unsigned long long *>(&c);
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y[i]), "l"(z[(i + 1) % ILP]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 v = *reinterpret_cast<float2 *>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA2 interleaved with scalar ALU work (integer adds) to see whether the freed issue slots are usable
template <int ILP> __global__ void k_ffma2_alu(float *out, const float *in) {
    unsigned long long x[ILP], y[ILP], z[ILP];
    int c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 a = make_float2(in[i] + threadIdx.x, in[i + 1]), b = make_float2(in[i + 32] + threadIdx.x * 1e-9f, in[i + 33]), cc = make_float2(in[i + 64] + threadIdx.x * 1e-9f, in[i + 65]);
        x[i] = *reinterpret_cast<unsigned long long *>(&a); y[i] = *reinterpret_cast<unsigned long long *>(&b); z[i] = *reinterpret_cast<unsigned long long *>(&cc);
        c[i] = threadIdx.x + i;
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y[i]), "l"(z[(i + 1) % ILP]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c[i]) : "r"(c[(i + 1) % ILP]), "r"(it));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 v = *reinterpret_cast<float2 *>(&x[i]); s += v.x + v.y + c[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_ffma_alu(float *out, const float *in) {
    float x[ILP], y[ILP], z[ILP];
    int c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = in[i] + threadIdx.x; y[i] = in[i + 32] + threadIdx.x * 1e-9f; z[i] = in[i + 64] + threadIdx.x * 1e-9f; c[i] = threadIdx.x + i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            x[i] = fmaf(x[i], y[i], z[(i + 1) % ILP]);
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c[i]) : "r"(c[(i + 1) % ILP]), "r"(it));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_shfl(float *out, const float *in) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = in[i] + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_lds128(float *out, const float *in) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(in[i % 96], 1, 2, 3);
    __syncthreads();
    float4 acc[ILP];
    int idx = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = make_float4(0, 0, 0, 0);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float4 v = sm[(idx + i * 32 + (it & 7) * 64) & 1023];
            acc[i].x += v.x; idx ^= (int)v.w & 0;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_kernel(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    float *out, *in; CHK(cudaMalloc(&out, 1 << 24)); CHK(cudaMalloc(&in, 4096));
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 1e-6f * i; CHK(cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, nominal %d MHz\n", p.name, sms, clk_khz / 1000);
    const int ILP = 8;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 128, blocks = sms * warps_per_sm / 4;
        double winstr = (double)blocks * (threads / 32) * ITER * ILP;  // warp instructions of the probed kind
        auto rep = [&](const char *name, float ms, double per_iter_instr) {
            double clk = ms * 1e-3 * 1.965e9;  // assume max clock; compare ratios
            printf("  w/SM=%2d %-28s %8.3f ms  %6.2f warp-instr/clk/SM (x%.0f instr kinds)\n", warps_per_sm, name, ms, winstr * per_iter_instr / clk / sms, per_iter_instr);
        };
        rep("FFMA imm/const operands", time_kernel([&] { k_ffma<ILP><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 1);
        rep("FFMA 3 distinct regs", time_kernel([&] { k_ffma3<ILP><<<blocks, threads>>>(out, in); }), 1);
        rep("FFMA 2 regs + 1 const", time_kernel([&] { k_ffma2r1c<ILP><<<blocks, threads>>>(out, in, 0.999f, 0.998f, 0.997f, 0.996f); }), 1);
        rep("cell mix, const coeffs (x6)", time_kernel([&] { k_cellmix_const<ILP><<<blocks, threads>>>(out, in, 0.9f, 1e-5f, 1e-5f, 0.1f, 0.1f); }), 6);
        rep("cell mix, reg coeffs (x6)", time_kernel([&] { k_cellmix_reg<ILP><<<blocks, threads>>>(out, in); }), 6);
        rep("FFMA2 (fma.rn.f32x2)", time_kernel([&] { k_ffma2<ILP><<<blocks, threads>>>(out, in); }), 1);
        rep("FFMA + LOP3 interleaved", time_kernel([&] { k_ffma_alu<ILP><<<blocks, threads>>>(out, in); }), 2);
        rep("FFMA2 + LOP3 interleaved", time_kernel([&] { k_ffma2_alu<ILP><<<blocks, threads>>>(out, in); }), 2);
        rep("SHFL.UP", time_kernel([&] { k_shfl<ILP><<<blocks, threads>>>(out, in); }), 1);
        rep("LDS.128 (+FADD)", time_kernel([&] { k_lds128<ILP><<<blocks, threads>>>(out, in); }), 1);
    }
    return 0;
}
