// fri.cu — FRI commit-phase kernels (K11 in SURVEY.md §2.1): leaf hashing of the transposed layer
// and the degree-respecting projection. See fri.cuh for the reference citations.
#include "fri.cuh"

#include "blake3.cuh"
#include "commit.cuh"
#include "minidft.cuh"
#include "rp64.cuh"

__global__ void __launch_bounds__(256) fri_hash_blake3_kernel(const u64* __restrict__ ev, size_t m, int d, int ld, int nf,
                                                              uint4* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u32 ne = (u32)(nf * d);       // <= 48 elements: a single chunk
    const u32 nblk = (ne + 7) / 8;
    u32 cv[8];
    b3_iv(cv);
    for (u32 b = 0; b < nblk; b++) {
        u32 msg[16];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 e = b * 8 + k;
            u64 v = 0;
            if (e < ne) {
                u32 kk = e / d, comp = e % d;
                v = ev[(i + (size_t)kk * m) * ld + comp];
            }
            msg[2 * k] = (u32)v;
            msg[2 * k + 1] = (u32)(v >> 32);
        }
        u32 bl = min(64u, (ne - b * 8) * 8);
        u32 fl = (b == 0 ? B3_CHUNK_START : 0) | (b == nblk - 1 ? (B3_CHUNK_END | B3_ROOT) : 0);
        b3_compress(cv, msg, 0, bl, fl, b3_runtime_one());
    }
    digests[2 * i] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    digests[2 * i + 1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

__global__ void __launch_bounds__(128) fri_hash_rp64_kernel(const u64* __restrict__ ev, size_t m, int d, int ld, int nf,
                                                            u64* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u32 ne = (u32)(nf * d);
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    s[0] = ne;
    u32 r = 0;
    for (u32 e = 0; e < ne; e++) {
        u32 kk = e / d, comp = e % d;
        s[4 + r] = gl_add(s[4 + r], ev[(i + (size_t)kk * m) * ld + comp]);
        if (++r == 8) { rp64_permute(s); r = 0; }
    }
    if (r > 0) rp64_permute(s);
#pragma unroll
    for (int k = 0; k < 4; k++) digests[i * 4 + k] = s[4 + k];
}

__host__ __device__ constexpr u32 cbrev(u32 v, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <int D, int LOGNF>
__global__ void __launch_bounds__(256) fri_fold_kernel(const u64* __restrict__ ev, size_t m, int ld, GlExt<D> alpha,
                                                       const u64* __restrict__ master, u32 logL, u64* __restrict__ next,
                                                       int next_ld) {
    constexpr int NF = 1 << LOGNF;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 x[D][NF];
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
        for (int c = 0; c < D; c++) x[c][k] = ev[(i + (size_t)k * m) * ld + c];
#pragma unroll
    for (int c = 0; c < D; c++) mini_dft<LOGNF>(x[c]);
    // x[c][pos] = F[bitrev(pos)], F = forward size-NF DFT; inverse coefficient j = F[(NF - j) % NF] / NF
    const u32 L = 1u << logL;
    u32 e = (L - (u32)i) & (L - 1);
    u64 winv = (e & (L >> 1)) ? gl_neg(master[e & ((L >> 1) - 1)]) : master[e & ((L >> 1) - 1)];
    u64 xinv = gl_mul(winv, 2635249152773512046ULL);  // 7^-1 mod p
    GlExt<D> beta = ext_mul_base(alpha, xinv);
    GlExt<D> acc;
    {
        constexpr u32 pos = cbrev(1u, LOGNF);  // j = NF-1 -> forward index 1
#pragma unroll
        for (int c = 0; c < D; c++) acc.v[c] = x[c][pos];
    }
#pragma unroll
    for (int j = NF - 2; j >= 0; j--) {
        const u32 pos = cbrev((u32)((NF - j) % NF), LOGNF);
        GlExt<D> cj;
#pragma unroll
        for (int c = 0; c < D; c++) cj.v[c] = x[c][pos];
        acc = ext_add(ext_mul(acc, beta), cj);
    }
    const u64 inv_nf = GL_P - ((GL_P - 1) >> LOGNF);  // (2^LOGNF)^-1 mod p
    acc = ext_mul_base(acc, inv_nf);
#pragma unroll
    for (int c = 0; c < D; c++) next[i * next_ld + c] = acc.v[c];
}

cudaError_t fri_hash_layer(int hash_id, const u64* evals, size_t len, int d, int ld, int nf, u64* digests,
                           cudaStream_t st) {
    size_t m = len / nf;
    if (hash_id == WF_HASH_BLAKE3_256)
        fri_hash_blake3_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(evals, m, d, ld, nf,
                                                                            reinterpret_cast<uint4*>(digests));
    else
        fri_hash_rp64_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(evals, m, d, ld, nf, digests);
    return cudaGetLastError();
}

template <int D>
static cudaError_t fold_dispatch(const u64* evals, size_t len, int ld, int nf, const u64* alpha, const u64* master,
                                 u64* next, int next_ld, cudaStream_t st) {
    size_t m = len / nf;
    u32 logL = 0;
    while (((size_t)1 << logL) < len) logL++;
    GlExt<D> a;
    for (int c = 0; c < D; c++) a.v[c] = alpha[c];
    unsigned blocks = (unsigned)((m + 255) / 256);
    switch (nf) {
        case 2: fri_fold_kernel<D, 1><<<blocks, 256, 0, st>>>(evals, m, ld, a, master, logL, next, next_ld); break;
        case 4: fri_fold_kernel<D, 2><<<blocks, 256, 0, st>>>(evals, m, ld, a, master, logL, next, next_ld); break;
        case 8: fri_fold_kernel<D, 3><<<blocks, 256, 0, st>>>(evals, m, ld, a, master, logL, next, next_ld); break;
        case 16: fri_fold_kernel<D, 4><<<blocks, 256, 0, st>>>(evals, m, ld, a, master, logL, next, next_ld); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t fri_fold_layer(const u64* evals, size_t len, int d, int ld, int nf, const u64* alpha, const u64* master,
                           u64* next, int next_ld, cudaStream_t st) {
    switch (d) {
        case 1: return fold_dispatch<1>(evals, len, ld, nf, alpha, master, next, next_ld, st);
        case 2: return fold_dispatch<2>(evals, len, ld, nf, alpha, master, next, next_ld, st);
        case 3: return fold_dispatch<3>(evals, len, ld, nf, alpha, master, next, next_ld, st);
        default: return cudaErrorInvalidValue;
    }
}
