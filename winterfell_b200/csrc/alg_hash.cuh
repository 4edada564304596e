// alg_hash.cuh — the hashers that absorb 64-bit words (the two arithmetization-friendly ones and SHA3-256) behind one compile-time interface, so that every kernel that
// hashes field elements (row leaves, FRI leaves, Merkle merges, the device coin, grinding) is written once:
//   AlgSponge<H>::init(n) / absorb(x) / finish(out)   ElementHasher::hash_elements
//   alg_merge<H>(in[8], out[4])                        Hasher::merge
//   alg_merge_with_int<H>(seed[4], value, out[4])      Hasher::merge_with_int
// H = WF_HASH_RP64_256 (crypto/src/hash/rescue/rp64_256/mod.rs), WF_HASH_RPJIVE64_256 (rp64_256_jive/mod.rs) or WF_HASH_SHA3_256
// (crypto/src/hash/sha/mod.rs: every message is a whole number of little-endian u64 words — elements, digests, seed + u64).
// merge_many of both = hash_elements over the digests' elements (rp64_256/mod.rs:194, rp64_256_jive/mod.rs:198).
#pragma once
#include "commit.cuh"
#include "rp64.cuh"
#include "rpjive.cuh"
#include "sha3.cuh"

template <int HASH>
struct AlgSponge;
template <>
struct AlgSponge<WF_HASH_RP64_256> {  // rp64_256/mod.rs:224-257: state[0] = length, rate 8, no padding
    u64 s[12];
    u32 i;
    GL_HD void init(size_t n) {
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = 0;
        s[0] = (u64)n;
        i = 0;
    }
    GL_HD void absorb(u64 x) {
        s[4 + i] = gl_add(s[4 + i], x);
        if (++i == 8) { rp64_permute(s); i = 0; }
    }
    GL_HD void finish(u64 out[4]) {
        if (i > 0) rp64_permute(s);
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = s[4 + k];
    }
};
template <>
struct AlgSponge<WF_HASH_RPJIVE64_256> : RpjSponge {};
template <>
struct AlgSponge<WF_HASH_SHA3_256> : Sha3Sponge {};

template <int HASH>
GL_HD void alg_merge(const u64 in[8], u64 out[4]) {
    if (HASH == WF_HASH_RP64_256) rp64_merge(in, out);
    else if (HASH == WF_HASH_SHA3_256) {  // sha/mod.rs:30-32: SHA3 of the 64 digest bytes
        Sha3Sponge sp;
        sp.init(8);
#pragma unroll
        for (int k = 0; k < 8; k++) { sp.st[k] ^= in[k]; }
        sp.i = 8;
        sp.finish(out);
    } else rpj_merge(in, out);
}
template <int HASH>
GL_HD void alg_merge_with_int(const u64 seed[4], u64 value, u64 out[4]) {
    if (HASH == WF_HASH_RP64_256) {  // rp64_256/mod.rs:198-218
        u64 s[12];
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) s[4 + k] = seed[k];
        if (value < GL_P) { s[8] = value; s[0] = 5; }
        else { s[8] = value - GL_P; s[9] = 1; s[0] = 6; }
        rp64_permute(s);
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = s[4 + k];
    } else if (HASH == WF_HASH_SHA3_256) {  // sha/mod.rs:38-43: seed || u64 LE, 40 bytes
        Sha3Sponge sp;
        sp.init(5);
#pragma unroll
        for (int k = 0; k < 4; k++) sp.st[k] ^= seed[k];
        sp.st[4] ^= value;
        sp.i = 5;
        sp.finish(out);
    } else {
        rpj_merge_with_int(seed, value, out);
    }
}
