// blake3.cuh — BLAKE3 compression for device kernels (one hash per thread, state and message in
// registers) and for the host-side Fiat-Shamir transcript of the product.
//
// The reference gets BLAKE3 from the third-party `blake3 = "1.8"` crate (crypto/Cargo.toml:34) and
// wraps it as Blake3_256 (crypto/src/hash/blake/mod.rs): hash_elements :52-65 = BLAKE3 of the
// canonical little-endian bytes (no length prefix), merge :33 = BLAKE3 of 64 bytes,
// merge_with_int :41-46 = BLAKE3 of seed || u64 LE. This file implements the public BLAKE3 spec.
#pragma once
#include <stdint.h>
#include <string.h>

#include "gl64.cuh"

#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_PARENT 4u
#define B3_ROOT 8u

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u

GL_HD u32 b3_rotr(u32 x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

// The G function's six additions are written as x = y * one + x with `one` a RUNTIME 1 (the kernels
// derive it from blockDim so ptxas cannot fold it): they become IMADs on the FMA pipe, leaving only
// the four XORs and four rotates on the ALU pipe. ncu on the plain version showed the ALU pipe 91 %
// busy and the FMA pipe 9 % (IADD3 / LOP3 / SHF all issue on the ALU pipe at half rate), so this
// balances the two pipes: 8 ALU + 6 FMA instead of 12 ALU per G.
#define B3_G(a, b, c, d, mx, my)      \
    a = b * one + a;                  \
    a = (mx) * one + a;               \
    d = b3_rotr(d ^ a, 16);           \
    c = d * one + c;                  \
    b = b3_rotr(b ^ c, 12);           \
    a = b * one + a;                  \
    a = (my) * one + a;               \
    d = b3_rotr(d ^ a, 8);            \
    c = d * one + c;                  \
    b = b3_rotr(b ^ c, 7);

// One round with the message words addressed through the round's schedule (compile-time indices
// once unrolled, so m[] stays in registers).
#define B3_ROUND(S0, S1, S2, S3, S4, S5, S6, S7, S8, S9, S10, S11, S12, S13, S14, S15) \
    B3_G(s0, s4, s8, s12, m[S0], m[S1])                                                  \
    B3_G(s1, s5, s9, s13, m[S2], m[S3])                                                  \
    B3_G(s2, s6, s10, s14, m[S4], m[S5])                                                 \
    B3_G(s3, s7, s11, s15, m[S6], m[S7])                                                 \
    B3_G(s0, s5, s10, s15, m[S8], m[S9])                                                 \
    B3_G(s1, s6, s11, s12, m[S10], m[S11])                                               \
    B3_G(s2, s7, s8, s13, m[S12], m[S13])                                                \
    B3_G(s3, s4, s9, s14, m[S14], m[S15])

// cv[8] <- first 8 words of compress(cv, m, counter, block_len, flags)
GL_HD void b3_compress(u32 cv[8], const u32 m[16], u64 counter, u32 block_len, u32 flags, u32 one = 1) {
    u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3;
    u32 s12 = (u32)counter, s13 = (u32)(counter >> 32), s14 = block_len, s15 = flags;
    B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8;
    cv[1] = s1 ^ s9;
    cv[2] = s2 ^ s10;
    cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12;
    cv[5] = s5 ^ s13;
    cv[6] = s6 ^ s14;
    cv[7] = s7 ^ s15;
}

GL_HD void b3_iv(u32 cv[8]) {
    cv[0] = B3_IV0; cv[1] = B3_IV1; cv[2] = B3_IV2; cv[3] = B3_IV3;
    cv[4] = B3_IV4; cv[5] = B3_IV5; cv[6] = B3_IV6; cv[7] = B3_IV7;
}

// 64-byte input (two digests, or one 8-element row): single block, CHUNK_START|CHUNK_END|ROOT.
GL_HD void b3_hash64(const u32 m[16], u32 out[8], u32 one = 1) {
    b3_iv(out);
    b3_compress(out, m, 0, 64, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT, one);
}
// merge of two digests of DW 32-bit words each (8: Blake3_256, blake/mod.rs:33; 6: Blake3_192, :85-88 — the 48 digest bytes
// back to back), result truncated to DW words, the rest of the slot zero
template <int DW>
GL_HD void b3_merge_words(const u32 a[8], const u32 b[8], u32 out[8], u32 one = 1) {
    u32 m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = k < DW ? a[k] : (k < 2 * DW ? b[k - DW] : 0);
    b3_iv(out);
    b3_compress(out, m, 0, 8 * DW, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT, one);
#pragma unroll
    for (int k = DW; k < 8; k++) out[k] = 0;
}
// merge_with_int (blake/mod.rs:41-46, :95-102): seed digest bytes then the u64 little-endian
template <int DW>
GL_HD void b3_merge_with_int_words(const u32 seed[8], u64 value, u32 out[8], u32 one = 1) {
    u32 m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = k < DW ? seed[k] : 0;
    m[DW] = (u32)value;
    m[DW + 1] = (u32)(value >> 32);
    b3_iv(out);
    b3_compress(out, m, 0, 4 * DW + 8, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT, one);
#pragma unroll
    for (int k = DW; k < 8; k++) out[k] = 0;
}
// runtime 1 for the device kernels (see B3_G)
#ifdef __CUDACC__
__device__ __forceinline__ u32 b3_runtime_one() { return blockDim.y; }
#endif

// -------------------------------------------------------------------------------------------------
// Host-side general hasher (any length; used by the transcript: seeds, OOD frames, remainders).
// Iterative chunk loop with a chaining-value stack (BLAKE3 spec section 5.1.2).
// -------------------------------------------------------------------------------------------------
static inline void b3_host_chunk(const u8* data, size_t len, u64 chunk_idx, bool root, u32 cv[8]) {
    b3_iv(cv);
    size_t nblk = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblk; b++) {
        u32 m[16] = {0};
        size_t bl = len == 0 ? 0 : (len - b * 64 < 64 ? len - b * 64 : 64);
        memcpy(m, data + b * 64, bl);
        u32 fl = (b == 0 ? B3_CHUNK_START : 0) | (b == nblk - 1 ? (B3_CHUNK_END | (root ? B3_ROOT : 0)) : 0);
        b3_compress(cv, m, chunk_idx, (u32)bl, fl);
    }
}
static inline void b3_host_hash(const u8* data, size_t len, u8 out[32]) {
    u32 stack[64][8];
    int sp = 0;
    size_t nchunks = len <= 1024 ? 1 : (len + 1023) / 1024;
    u32 cv[8];
    for (size_t c = 0; c < nchunks; c++) {
        size_t cl = (c == nchunks - 1) ? len - c * 1024 : 1024;
        b3_host_chunk(data + c * 1024, cl, c, nchunks == 1, cv);
        if (c == nchunks - 1) break;
        // merge completed subtrees: one merge per trailing 1 bit of the chunk count so far
        size_t total = c + 1;
        while ((total & 1) == 0) {
            u32 m[16];
            memcpy(m, stack[--sp], 32);
            memcpy(m + 8, cv, 32);
            b3_iv(cv);
            b3_compress(cv, m, 0, 64, B3_PARENT);
            total >>= 1;
        }
        memcpy(stack[sp++], cv, 32);
    }
    // fold the stack right-to-left; the last merge carries ROOT
    while (sp > 0) {
        u32 m[16];
        memcpy(m, stack[--sp], 32);
        memcpy(m + 8, cv, 32);
        b3_iv(cv);
        b3_compress(cv, m, 0, 64, B3_PARENT | (sp == 0 ? B3_ROOT : 0));
    }
    memcpy(out, cv, 32);
}
