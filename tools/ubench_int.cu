// ubench_int.cu — integer-pipe microbenchmarks on the B200 that the Goldilocks kernels are designed against:
// per-SM issue rates of the instructions the field arithmetic is built from, and the throughput of
// candidate formulations of gl_mul / butterfly. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o ubench_int tools/ubench_int.cu ; run on the GPU box; prints one JSON line per test.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../winterfell_b200/csrc/gl64.cuh"

#define ITERS 4096
#define ILP 8

enum { T_IADD3, T_IADD_CC, T_IMAD, T_IMADHI, T_IMADWIDE, T_SHF, T_LOP3, T_MIX_ADD_MAD, T_MIX_ADD_WIDE, T_SHFL, T_GLMUL, T_GLMUL_NOCANON,
       T_GLBFLY, T_GLSUB, T_GLADD, T_SHIFTMUL24, T_LDS64, T_LDS128, T_GLMUL_WIDE, T_COUNT };
static const char* NAMES[] = {"iadd3", "iadd_cc_chain(2 instr)", "imad.lo", "imad.hi", "imad.wide", "shf", "lop3", "mix iadd3+imad (2 instr)",
                              "mix iadd3+imad.wide (2 instr)", "shfl.bfly", "gl_mul", "gl_mul (no canon)", "gl_butterfly (2 elem)", "gl_sub",
                              "gl_add", "gl_mul_2exp<24>", "lds.64", "lds.128", "gl_mul via mad.wide"};

__device__ __forceinline__ u64 gl_mul_nocanon(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, c0, c1, c2, c3, o0, o1, o2, k, m;\n\t"
        ".reg .s32 adj;\n\t"
        ".reg .u64 t;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "mul.lo.u32 c0, a0, b0;\n\t"
        "mul.hi.u32 c1, a0, b0;\n\t"
        "mul.lo.u32 c2, a1, b1;\n\t"
        "mul.hi.u32 c3, a1, b1;\n\t"
        "mul.lo.u32 o0, a0, b1;\n\t"
        "mul.hi.u32 o1, a0, b1;\n\t"
        "mad.lo.cc.u32 o0, a1, b0, o0;\n\t"
        "madc.hi.cc.u32 o1, a1, b0, o1;\n\t"
        "addc.u32 o2, 0, 0;\n\t"
        "add.cc.u32 c1, c1, o0;\n\t"
        "addc.cc.u32 c2, c2, o1;\n\t"
        "addc.u32 c3, c3, o2;\n\t"
        "mad.lo.cc.u32 c0, c2, 0xffffffff, c0;\n\t"
        "madc.hi.cc.u32 c1, c2, 0xffffffff, c1;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "sub.cc.u32 c0, c0, c3;\n\t"
        "subc.cc.u32 c1, c1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "add.s32 adj, k, m;\n\t"
        "mov.b64 t, {c0, c1};\n\t"
        "mad.wide.s32 t, adj, -1, t;\n\t"
        "mov.b64 {c0, c1}, t;\n\t"
        "add.u32 c1, c1, adj;\n\t"
        "mov.b64 %0, {c0, c1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
// 64x64 product through four mad.wide.u32 with 64-bit accumulation (no overflow: see the derivation in DESIGN.md)
__device__ __forceinline__ u64 gl_mul_wide(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 p00 = (u64)a0 * b0;
    u64 t = (u64)a0 * b1 + (p00 >> 32);
    u64 u = (u64)a1 * b0 + (u32)t;
    u64 v = (u64)a1 * b1 + (t >> 32) + (u >> 32);
    u64 lo = (u64)(u32)p00 | (u << 32);
    return gl_reduce128(lo, v);
}

template <int T>
__global__ void __launch_bounds__(256) k(u64* out, u64 seed) {
    __shared__ __align__(16) u64 sm[256 * 4];
    u64 x[ILP], y[ILP];
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        x[i] = seed * (u64)(tid * ILP + i + 1) + blockIdx.x;
        y[i] = (seed ^ 0x9e3779b97f4a7c15ULL) * (u64)(tid + i + 3);
        if (T == T_GLMUL || T == T_GLBFLY || T == T_GLSUB || T == T_GLADD || T == T_SHIFTMUL24 || T == T_GLMUL_WIDE) { x[i] %= GL_P; y[i] %= GL_P; }
    }
    sm[tid * 4] = x[0]; sm[tid * 4 + 1] = x[1]; sm[tid * 4 + 2] = x[2]; sm[tid * 4 + 3] = x[3];
    __syncthreads();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            u32 a = (u32)x[i], b = (u32)(x[i] >> 32), c = (u32)y[i], d = (u32)(y[i] >> 32);
            if (T == T_IADD3) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(d)); }
            if (T == T_IADD_CC) { asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a), "+r"(b) : "r"(c), "r"(d)); }
            if (T == T_IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(d), "r"(c)); }
            if (T == T_IMADHI) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d)); asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(d), "r"(c)); }
            if (T == T_IMADWIDE) { u64 t = x[i]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(c), "r"(d)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(d), "r"(c)); a = (u32)t; b = (u32)(t >> 32); }
            if (T == T_SHF) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 9;" : "+r"(b) : "r"(a)); }
            if (T == T_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b) : "r"(d), "r"(c)); }
            if (T == T_MIX_ADD_MAD) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(d), "r"(c)); }
            if (T == T_MIX_ADD_WIDE) { u64 t = y[i]; asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(b), "r"(d)); y[i] = t; }
            if (T == T_SHFL) { a = __shfl_xor_sync(0xffffffffu, a, 1); b = __shfl_xor_sync(0xffffffffu, b, 2); }
            x[i] = (u64)a | ((u64)b << 32);
            if (T == T_GLMUL) x[i] = gl_mul(x[i], y[i]);
            if (T == T_GLMUL_NOCANON) x[i] = gl_mul_nocanon(x[i], y[i]);
            if (T == T_GLMUL_WIDE) x[i] = gl_mul_wide(x[i], y[i]);
            if (T == T_GLBFLY) gl_butterfly(x[i], y[i]);
            if (T == T_GLSUB) x[i] = gl_sub(x[i], y[i]);
            if (T == T_GLADD) x[i] = gl_add(x[i], y[i]);
            if (T == T_SHIFTMUL24) x[i] = gl_mul_2exp<24>(x[i]);
            if (T == T_LDS64) { x[i] ^= sm[((u32)x[i] & 255) * 4 + (i & 3)]; }
            if (T == T_LDS128) { ulonglong2 v = *reinterpret_cast<ulonglong2*>(&sm[(((u32)x[i] & 255) * 4) + (i & 1) * 2]); x[i] ^= v.x + v.y; }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= x[i] ^ y[i];
    out[(size_t)blockIdx.x * blockDim.x + tid] = acc;
}

template <int T>
void run(u64* d_out, int sms, double clk_ghz, int ops_per_item) {
    const int blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<T><<<blocks, 256>>>(d_out, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<T><<<blocks, 256>>>(d_out, 12345);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double items = (double)blocks * 256 * ITERS * ILP;  // per-thread "items" (an item = ops_per_item instructions / one field op)
    double per_s = items / (ms * 1e-3);
    printf("{\"test\": \"%s\", \"ms\": %.4f, \"Gitems_per_s\": %.2f, \"items_per_clk_per_sm\": %.2f, \"instr_per_item\": %d}\n", NAMES[T], ms,
           per_s / 1e9, per_s / (clk_ghz * 1e9) / sms, ops_per_item);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clk = clk_khz / 1e6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, clk);
    u64* d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    const int s = p.multiProcessorCount;
    run<T_IADD3>(d, s, clk, 2); run<T_IADD_CC>(d, s, clk, 2); run<T_IMAD>(d, s, clk, 2); run<T_IMADHI>(d, s, clk, 2);
    run<T_IMADWIDE>(d, s, clk, 2); run<T_SHF>(d, s, clk, 2); run<T_LOP3>(d, s, clk, 2); run<T_MIX_ADD_MAD>(d, s, clk, 2);
    run<T_MIX_ADD_WIDE>(d, s, clk, 2); run<T_SHFL>(d, s, clk, 2); run<T_GLMUL>(d, s, clk, 1); run<T_GLMUL_NOCANON>(d, s, clk, 1);
    run<T_GLMUL_WIDE>(d, s, clk, 1); run<T_GLBFLY>(d, s, clk, 1); run<T_GLSUB>(d, s, clk, 1); run<T_GLADD>(d, s, clk, 1);
    run<T_SHIFTMUL24>(d, s, clk, 1); run<T_LDS64>(d, s, clk, 1); run<T_LDS128>(d, s, clk, 1);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
