// minidft.cuh — register-resident DFTs of size 2..16 over Goldilocks whose twiddles are powers of two.
#pragma once
#include "gl64.cuh"

// -------------------------------------------------------------------------------------------------
// Register mini-DFTs (decimation in frequency, in place, bit-reversed output). All internal
// twiddles are powers of two: w_8 = 2^24, w_4 = 2^48, w_2 = -1 (gl64.cuh gl_mul_2exp).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf2(u64& a, u64& b) { gl_butterfly(a, b); }
template <int R>
__device__ __forceinline__ void mini_dft(u64* x);
template <>
__device__ __forceinline__ void mini_dft<1>(u64* x) { bf2(x[0], x[1]); }
template <>
__device__ __forceinline__ void mini_dft<2>(u64* x) {
    bf2(x[0], x[2]);
    bf2(x[1], x[3]);
    x[3] = gl_mul_2exp<48>(x[3]);
    bf2(x[0], x[1]);
    bf2(x[2], x[3]);
}
template <>
__device__ __forceinline__ void mini_dft<3>(u64* x);
template <>
__device__ __forceinline__ void mini_dft<3>(u64* x) {
    bf2(x[0], x[4]);
    bf2(x[1], x[5]);
    bf2(x[2], x[6]);
    bf2(x[3], x[7]);
    x[5] = gl_mul_2exp<24>(x[5]);
    x[6] = gl_mul_2exp<48>(x[6]);
    x[7] = gl_mul_2exp<72>(x[7]);
    mini_dft<2>(x);
    mini_dft<2>(x + 4);
}
template <>
__device__ __forceinline__ void mini_dft<4>(u64* x) {
    bf2(x[0], x[8]);
    bf2(x[1], x[9]);
    bf2(x[2], x[10]);
    bf2(x[3], x[11]);
    bf2(x[4], x[12]);
    bf2(x[5], x[13]);
    bf2(x[6], x[14]);
    bf2(x[7], x[15]);
    x[9] = gl_mul_2exp<12>(x[9]);
    x[10] = gl_mul_2exp<24>(x[10]);
    x[11] = gl_mul_2exp<36>(x[11]);
    x[12] = gl_mul_2exp<48>(x[12]);
    x[13] = gl_mul_2exp<60>(x[13]);
    x[14] = gl_mul_2exp<72>(x[14]);
    x[15] = gl_mul_2exp<84>(x[15]);
    mini_dft<3>(x);
    mini_dft<3>(x + 8);
}
template <>
__device__ __forceinline__ void mini_dft<5>(u64* x) {
#pragma unroll
    for (int q = 0; q < 16; q++) bf2(x[q], x[q + 16]);
    // w_32 = 2^6: x[16 + q] *= 2^(6q)
    x[17] = gl_mul_2exp<6>(x[17]);
    x[18] = gl_mul_2exp<12>(x[18]);
    x[19] = gl_mul_2exp<18>(x[19]);
    x[20] = gl_mul_2exp<24>(x[20]);
    x[21] = gl_mul_2exp<30>(x[21]);
    x[22] = gl_mul_2exp<36>(x[22]);
    x[23] = gl_mul_2exp<42>(x[23]);
    x[24] = gl_mul_2exp<48>(x[24]);
    x[25] = gl_mul_2exp<54>(x[25]);
    x[26] = gl_mul_2exp<60>(x[26]);
    x[27] = gl_mul_2exp<66>(x[27]);
    x[28] = gl_mul_2exp<72>(x[28]);
    x[29] = gl_mul_2exp<78>(x[29]);
    x[30] = gl_mul_2exp<84>(x[30]);
    x[31] = gl_mul_2exp<90>(x[31]);
    mini_dft<4>(x);
    mini_dft<4>(x + 16);
}
__device__ __forceinline__ u32 brev(u32 v, int bits) { return __brev(v) >> (32 - bits); }

// w_S^e for e in [0, S) from the half table (w^(e + S/2) = -w^e)
__device__ __forceinline__ u64 tw_lookup(const u64* tab, u32 e, u32 half) {
    u64 t = tab[e & (half - 1)];
    return (e & half) ? gl_neg(t) : t;
}

