// rp64.cuh — Rescue Prime Rp64_256 permutation and sponge for device kernels and the host
// transcript. Follows crypto/src/hash/rescue/rp64_256/mod.rs: state width 12, capacity = elements
// 0..4, rate = 4..12, digest = 4..8, 7 rounds of (x^7, MDS, +ARK1, x^(1/7), MDS, +ARK2)
// (apply_round :309-321, apply_inv_sbox :351-385), sponge hash_elements :224-257, merge :181-192,
// merge_with_int :198-218. The MDS is the circulant with first row [7,23,8,26,13,10,9,7,6,22,21,8]
// (:390); the reference multiplies by it through an integer FFT (mds_f64_12x12.rs:41) — here the
// row has entries < 32, so the product is accumulated from 32-bit halves without any reduction
// until the end (sum < 2^73).
#pragma once
#include "gl64.cuh"

#ifdef __CUDACC__
#define RP64_CONST_QUAL static __device__ __constant__ const
#include "rp64_constants.inc"
#undef RP64_CONST_QUAL
#endif
namespace rp64_host {
#define RP64_CONST_QUAL static const
#include "rp64_constants.inc"
#undef RP64_CONST_QUAL
}  // namespace rp64_host

#ifdef __CUDA_ARCH__
#define RP64_TAB(name) name
#else
#define RP64_TAB(name) rp64_host::name
#endif

GL_HD u64 rp64_exp7(u64 x) {  // f64/mod.rs:96
    u64 x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}

// x^(1/7) = x^10540996611094048183, addition chain of apply_inv_sbox (:351-385), per element
GL_HD u64 rp64_inv7(u64 x) {
    u64 t1 = gl_sqr(x);            // x^(0b10)
    u64 t2 = gl_sqr(t1);           // x^(0b100)
    u64 t3 = t2;                   // exp_acc<3>(t2, t2)
#pragma unroll
    for (int i = 0; i < 3; i++) t3 = gl_sqr(t3);
    t3 = gl_mul(t3, t2);
    u64 t4 = t3;                   // exp_acc<6>(t3, t3)
#pragma unroll
    for (int i = 0; i < 6; i++) t4 = gl_sqr(t4);
    t4 = gl_mul(t4, t3);
    u64 t5 = t4;                   // exp_acc<12>(t4, t4)
#pragma unroll
    for (int i = 0; i < 12; i++) t5 = gl_sqr(t5);
    t5 = gl_mul(t5, t4);
    u64 t6 = t5;                   // exp_acc<6>(t5, t3)
#pragma unroll
    for (int i = 0; i < 6; i++) t6 = gl_sqr(t6);
    t6 = gl_mul(t6, t3);
    u64 t7 = t6;                   // exp_acc<31>(t6, t6)
#pragma unroll 1
    for (int i = 0; i < 31; i++) t7 = gl_sqr(t7);
    t7 = gl_mul(t7, t6);
    u64 a = gl_sqr(gl_sqr(gl_mul(gl_sqr(t7), t6)));
    u64 b = gl_mul(gl_mul(t1, t2), x);
    return gl_mul(a, b);
}

GL_HD void rp64_mds(u64 s[12]) {
    u64 r[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 lo = 0, hi = 0;  // sums of 32-bit halves times coefficients < 32: each < 2^41
#pragma unroll
        for (int j = 0; j < 12; j++) {
            u64 c = RP64_TAB(RP64_MDS_ROW0)[(j + 12 - i) % 12];
            lo += (s[j] & GL_EPS) * c;
            hi += (s[j] >> 32) * c;
        }
        // value = lo + hi * 2^32
        u64 l = lo + (hi << 32);
        u64 h = (hi >> 32) + (l < lo ? 1 : 0);
        r[i] = gl_reduce128(l, h);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = r[i];
}

GL_HD void rp64_permute(u64 s[12]) {
#pragma unroll 1
    for (int r = 0; r < 7; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = rp64_exp7(s[i]);
        rp64_mds(s);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], RP64_TAB(RP64_ARK1)[r][i]);
#pragma unroll 1
        for (int i = 0; i < 12; i++) s[i] = rp64_inv7(s[i]);
        rp64_mds(s);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], RP64_TAB(RP64_ARK2)[r][i]);
    }
}

// merge of two digests given as 8 canonical elements (mod.rs:181-192)
GL_HD void rp64_merge(const u64 in[8], u64 out[4]) {
    u64 s[12];
    s[0] = 8; s[1] = 0; s[2] = 0; s[3] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s[4 + i] = in[i];
    rp64_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = s[4 + i];
}

// host sponge (mod.rs:224-257) for the transcript
static inline void rp64_host_hash_elements(const u64* e, size_t n, u64 out[4]) {
    u64 s[12] = {0};
    s[0] = (u64)n % GL_P;
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        s[4 + i] = gl_add(s[4 + i], e[k]);
        if (++i == 8) { rp64_permute(s); i = 0; }
    }
    if (i > 0) rp64_permute(s);
    for (int k = 0; k < 4; k++) out[k] = s[4 + k];
}
static inline void rp64_host_merge_with_int(const u64 seed[4], u64 value, u64 out[4]) {  // mod.rs:198-218
    u64 s[12] = {0};
    for (int k = 0; k < 4; k++) s[4 + k] = seed[k];
    if (value < GL_P) { s[8] = value; s[0] = 5; }
    else { s[8] = value - GL_P; s[9] = value / GL_P; s[0] = 6; }
    rp64_permute(s);
    for (int k = 0; k < 4; k++) out[k] = s[4 + k];
}
