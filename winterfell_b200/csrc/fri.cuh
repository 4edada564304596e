// fri.cuh — FRI layer leaf hashing and degree-respecting projection on device (fri.cu).
#pragma once
#include <cuda_runtime.h>

#include "gl64.cuh"

// Layer evaluations: `len` extension elements of degree d in natural order, element i at
// evals[i * ld + 0..d) (ld >= d words per element).
// Leaf i (i < len/nf) = hash_elements([v[i], v[i + m], ..., v[i + (nf-1) m]]), m = len / nf
// (utils/core/src/lib.rs:166-185 transpose_slice + fri/src/prover/mod.rs:321-336).
cudaError_t fri_hash_layer(int hash_id, const u64* evals, size_t len, int d, int ld, int nf, u64* digests,
                           cudaStream_t st);
// next[i] = apply_drp(row i) (fri/src/folding/mod.rs:86-118) with domain offset 7:
// size-nf inverse DFT of the row, coefficient k scaled by (7 w_len^i)^-k / nf, evaluated at alpha.
// inv_master: w_len^i for i < len/2 (forward root; negative powers are taken through the index).
// alpha: d host words, or nullptr with d_alpha = device pointer to the alpha drawn by fri_coin_step.
// Row-sharded layers (multi-GPU): `evals` holds rows [i0, i0 + len/nf) of each of the nf strided pieces of a layer of
// 2^logL_global points, piece after piece; master = w^i table of that global size.
cudaError_t fri_fold_layer(const u64* evals, size_t len, int d, int ld, int nf, const u64* alpha,
                           const u64* master, u64* next, int next_ld, cudaStream_t st, const u64* d_alpha = nullptr,
                           size_t i0 = 0, u32 logL_global = 0);
// One step of the device copy of the public coin: state[0..4) = seed (in/out), state[4] = counter (out);
// reseed with root[4], draw alpha (d words) into alpha_out[3]; log_entry[8] = root[4], alpha[3], ok flag.
cudaError_t fri_coin_step(int hash_id, u64* state, const u64* root, int d, u64* alpha_out, u64* log_entry, cudaStream_t st);
