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

// The S-box arithmetic runs on words that are only reduced below 2^64 (gl_mul_weak): every chain ends in the MDS product,
// which takes any 64-bit words and returns canonical ones.
#ifndef RP64_SQR_WIDE
#define RP64_SQR_WIDE 0
#endif
GL_HD u64 rp64_sqr(u64 x) {
#if RP64_SQR_WIDE
    return gl_sqr_weak(x);
#else
    return gl_mul_weak(x, x);
#endif
}
GL_HD u64 rp64_exp7(u64 x) {  // f64/mod.rs:96
    u64 x2 = rp64_sqr(x), x4 = rp64_sqr(x2), x3 = gl_mul_weak(x2, x);
    return gl_mul_weak(x3, x4);
}

// x^(1/7) = x^10540996611094048183 for G state elements at once, the addition chain of apply_inv_sbox (:351-385):
//   t1 = x^2, t2 = t1^2, t3 = t2^(2^3) t2, t4 = t3^(2^6) t3, t5 = t4^(2^12) t4, t6 = t5^(2^6) t3, t7 = t6^(2^31) t6,
//   result = ((t7^2 t6)^2)^2 * (t1 t2 x).
// The G chains are independent: the element loops are unrolled inside the (rolled) squaring loops, so G multiplications are in
// flight per thread (the first version walked one element's 72 dependent squarings at a time and ran at 0.4 instructions per
// clock and scheduler). Live values: four per element (the t1 t2 x product, t3, the saved factor, the running power).
#ifndef RP64_GROUP
#define RP64_GROUP 6
#endif
template <int G>
GL_HD void rp64_inv7_group(u64* s) {
    u64 b[G], t3[G], p[G], acc[G];
#pragma unroll
    for (int e = 0; e < G; e++) {
        const u64 x = s[e];
        const u64 t1 = rp64_sqr(x);
        const u64 t2 = rp64_sqr(t1);
        b[e] = gl_mul_weak(gl_mul_weak(t1, t2), x);
        p[e] = t2;
        acc[e] = t2;
    }
#define RP64_SQ(n)                                              \
    _Pragma("unroll 1") for (int i = 0; i < (n); i++) {         \
        _Pragma("unroll") for (int e = 0; e < G; e++) acc[e] = rp64_sqr(acc[e]); \
    }
    RP64_SQ(3)
#pragma unroll
    for (int e = 0; e < G; e++) { acc[e] = gl_mul_weak(acc[e], p[e]); t3[e] = acc[e]; }             // t3
    RP64_SQ(6)
#pragma unroll
    for (int e = 0; e < G; e++) { acc[e] = gl_mul_weak(acc[e], t3[e]); p[e] = acc[e]; }             // t4
    RP64_SQ(12)
#pragma unroll
    for (int e = 0; e < G; e++) acc[e] = gl_mul_weak(acc[e], p[e]);                                 // t5
    RP64_SQ(6)
#pragma unroll
    for (int e = 0; e < G; e++) { acc[e] = gl_mul_weak(acc[e], t3[e]); p[e] = acc[e]; }             // t6
    RP64_SQ(31)
#pragma unroll
    for (int e = 0; e < G; e++) {
        u64 a = gl_mul_weak(acc[e], p[e]);                                                          // t7
        a = gl_mul_weak(rp64_sqr(a), p[e]);
        a = rp64_sqr(a);
        a = rp64_sqr(a);
        s[e] = gl_mul_weak(a, b[e]);
    }
#undef RP64_SQ
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
        for (int g = 0; g < 12; g += RP64_GROUP) rp64_inv7_group<RP64_GROUP>(s + g);
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
