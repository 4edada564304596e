// sha3.cuh — SHA3-256 (FIPS 202: Keccak-f[1600], rate 136 bytes, domain byte 0x06) for device kernels and the host
// transcript. The reference takes it from the third-party `sha3 = "0.10"` crate (crypto/Cargo.toml) and wraps it as Sha3_256
// (crypto/src/hash/sha/mod.rs:19-60): hash_elements = SHA3 of the canonical little-endian element bytes, merge = SHA3 of the
// 64 digest bytes, merge_many = SHA3 of all digest bytes, merge_with_int = SHA3 of seed || u64 LE. Every message on this path
// is a whole number of 64-bit little-endian words, so the sponge absorbs one word (= one lane) at a time.
#pragma once
#include "gl64.cuh"

#ifdef __CUDACC__
static __device__ __constant__ const u64 SHA3_RC_DEV[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#endif
static const u64 SHA3_RC_HOST[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#ifdef __CUDA_ARCH__
#define SHA3_RC(i) SHA3_RC_DEV[i]
#else
#define SHA3_RC(i) SHA3_RC_HOST[i]
#endif

GL_HD u64 sha3_rotl(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600]: state lane (x, y) at a[x + 5 y]
GL_HD void keccak_f1600(u64 a[25]) {
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        u64 c[5], d[5];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];               // theta
#pragma unroll
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ sha3_rotl(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        u64 b[25];                                                                                              // rho + pi
        b[0] = a[0];
        b[10] = sha3_rotl(a[1], 1);   b[20] = sha3_rotl(a[2], 62);  b[5] = sha3_rotl(a[3], 28);   b[15] = sha3_rotl(a[4], 27);
        b[16] = sha3_rotl(a[5], 36);  b[1] = sha3_rotl(a[6], 44);   b[11] = sha3_rotl(a[7], 6);   b[21] = sha3_rotl(a[8], 55);
        b[6] = sha3_rotl(a[9], 20);   b[7] = sha3_rotl(a[10], 3);   b[17] = sha3_rotl(a[11], 10); b[2] = sha3_rotl(a[12], 43);
        b[12] = sha3_rotl(a[13], 25); b[22] = sha3_rotl(a[14], 39); b[23] = sha3_rotl(a[15], 41); b[8] = sha3_rotl(a[16], 45);
        b[18] = sha3_rotl(a[17], 15); b[3] = sha3_rotl(a[18], 21);  b[13] = sha3_rotl(a[19], 8);  b[14] = sha3_rotl(a[20], 18);
        b[24] = sha3_rotl(a[21], 2);  b[9] = sha3_rotl(a[22], 61);  b[19] = sha3_rotl(a[23], 56); b[4] = sha3_rotl(a[24], 14);
#pragma unroll
        for (int y = 0; y < 25; y += 5)                                                                         // chi
#pragma unroll
            for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        a[0] ^= SHA3_RC(round);                                                                                 // iota
    }
}

// SHA3-256 of a message delivered one 64-bit little-endian word at a time
struct Sha3Sponge {
    u64 st[25];
    u32 i;
    GL_HD void init(size_t /*n: unused, the padding encodes the length*/) {
#pragma unroll
        for (int k = 0; k < 25; k++) st[k] = 0;
        i = 0;
    }
    GL_HD void absorb(u64 w) {
        // the lane index is kept out of dynamic addressing of st[] where it matters (kernels unroll over the row)
        st[i] ^= w;
        if (++i == 17) { keccak_f1600(st); i = 0; }
    }
    GL_HD void finish(u64 out[4]) {
        st[i] ^= 0x06ULL;                       // domain separation + first padding bit at the next byte
        st[16] ^= 0x8000000000000000ULL;        // last padding bit: byte 135 of the rate
        keccak_f1600(st);
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = st[k];
    }
};
static inline void sha3_host_words(const u64* w, size_t n, u64 out[4]) {
    Sha3Sponge sp;
    sp.init(n);
    for (size_t k = 0; k < n; k++) sp.absorb(w[k]);
    sp.finish(out);
}
