// oracle/wf_oracle.h — CPU restatement of the winterfell v0.13.1 STARK hot path (TEST INFRASTRUCTURE).
//
// This is the parity oracle: a plain C++ restatement of the reference's algorithms, each function
// citing the reference file:line it follows. It is NOT part of the product. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Parity pinning: every known-answer vector the reference's own tests hold for this path is
// checked in tests/test_oracle_kats.py (Rp64_256 Sage vector, f64 quad/cubic products, Merkle
// fixtures, transpose/fold_positions, NTT == naive evaluation, LDE == naive evaluation).
// BLAKE3 has no golden digests in the reference ("parity unpinned" w.r.t. reference fixtures);
// it is pinned against the BLAKE3 spec via the Python `blake3` package.
//
// All field elements cross this interface as CANONICAL u64 (value in [0, p)), p = 2^64 - 2^32 + 1.
// Extension elements are d consecutive base elements (math/src/field/extensions/cubic.rs:117-121).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { WFO_HASH_BLAKE3_256 = 0, WFO_HASH_RP64_256 = 1, WFO_HASH_RPJIVE64_256 = 2, WFO_HASH_BLAKE3_192 = 3, WFO_HASH_SHA3_256 = 4 };

void wfo_set_threads(int n);  // number of OpenMP threads for the `concurrent`-style loops
int wfo_get_threads(void);

// ---- field (math/src/field/f64/mod.rs) ----
uint64_t wfo_add(uint64_t a, uint64_t b);
uint64_t wfo_sub(uint64_t a, uint64_t b);
uint64_t wfo_mul(uint64_t a, uint64_t b);
uint64_t wfo_inv(uint64_t a);
uint64_t wfo_exp(uint64_t a, uint64_t e);
uint64_t wfo_to_mont(uint64_t a);    // canonical -> in-memory Montgomery word (x * 2^64 mod p)
uint64_t wfo_from_mont(uint64_t a);  // Montgomery word -> canonical
uint64_t wfo_root_of_unity(uint32_t log_n);
void wfo_ext_mul(int d, const uint64_t* a, const uint64_t* b, uint64_t* out);
void wfo_ext_inv(int d, const uint64_t* a, uint64_t* out);

// ---- fft (math/src/fft) — arrays of n elements of extension degree d (n*d words) ----
void wfo_get_twiddles(size_t n, uint64_t* out);      // n/2 words, bit-reversed order
void wfo_get_inv_twiddles(size_t n, uint64_t* out);  // n/2 words, bit-reversed order
void wfo_evaluate_poly(uint64_t* p, size_t n, int d);
void wfo_interpolate_poly(uint64_t* v, size_t n, int d);
void wfo_evaluate_poly_with_offset(const uint64_t* p, size_t n, int d, uint64_t offset,
                                   size_t blowup, uint64_t* out);
void wfo_interpolate_poly_with_offset(uint64_t* v, size_t n, int d, uint64_t offset);
void wfo_eval_poly_at(const uint64_t* p, size_t n, int dp, const uint64_t* x, int dx, uint64_t* out);

// ---- matrices (prover/src/matrix) ----
// cols: c columns, each n elements of degree d, column-major ([c][n*d]); in place.
void wfo_interpolate_columns(uint64_t* cols, size_t c, size_t n, int d);
// polys [c][n*d] -> row-major LDE out[N][c*d], N = n*blowup, row i <-> point 7*w_N^i.
void wfo_lde_rows(const uint64_t* polys, size_t c, size_t n, int d, size_t blowup, uint64_t* out);

// ---- hashing (crypto/src/hash) ----
void wfo_blake3(const uint8_t* data, size_t len, uint8_t out[32]);
void wfo_rp64_permute(uint64_t state[12]);
void wfo_rpjive_permute(uint64_t state[8]);  // rp64_256_jive/mod.rs:313-319
void wfo_hash_elements(int hash_id, const uint64_t* elems, size_t n, uint8_t out[32]);
void wfo_merge(int hash_id, const uint8_t two[64], uint8_t out[32]);
void wfo_merge_many(int hash_id, const uint8_t* digests, size_t n, uint8_t out[32]);
void wfo_merge_with_int(int hash_id, const uint8_t seed[32], uint64_t value, uint8_t out[32]);
// row hashing (prover/src/matrix/row_matrix.rs:184-228); partition_size == row_elems => no partitions
void wfo_hash_rows(int hash_id, const uint64_t* rows, size_t nrows, size_t row_elems,
                   size_t partition_size, uint8_t* digests);

// ---- Merkle (crypto/src/merkle) ----
// nodes: nleaves digests; nodes[0] = zeros, nodes[1] = root (mod.rs:344-368)
void wfo_merkle_nodes(int hash_id, const uint8_t* leaves, size_t nleaves, uint8_t* nodes);
// batch proof (mod.rs:217-272 + proofs.rs:390-401 serialization). Returns number of bytes written
// to `out` (capacity out_cap), or -1 on error. `leaves_out` receives k digests in index order.
long wfo_merkle_prove_batch(const uint8_t* leaves, const uint8_t* nodes, size_t nleaves,
                            const uint64_t* indexes, size_t k, uint8_t* leaves_out,
                            uint8_t* out, size_t out_cap);
long wfo_merkle_prove_batch_h(int hash_id, const uint8_t* leaves, const uint8_t* nodes, size_t nleaves, const uint64_t* idx, size_t k,
                              uint8_t* leaves_out, uint8_t* out, size_t cap);

// ---- FRI (fri/src) ----
void wfo_transpose_slice(const uint64_t* src, size_t len, int d, size_t folding, uint64_t* dst);
void wfo_apply_drp(const uint64_t* transposed, size_t rows, int d, size_t folding, uint64_t offset,
                   const uint64_t* alpha, uint64_t* out);
size_t wfo_fold_positions(const uint64_t* pos, size_t k, size_t source_domain, size_t folding,
                          uint64_t* out);
size_t wfo_fri_num_layers(size_t domain_size, size_t folding, size_t remainder_max_degree, size_t blowup);
// Commit phase with a DefaultProverChannel-style transcript seeded with hash_elements([])
// (fri/src/prover/channel.rs). roots: (num_layers+1) x 32 bytes (last = remainder commitment);
// remainder: reversed coefficients (domain/blowup elements). Returns num_layers.
size_t wfo_fri_build_layers(int hash_id, const uint64_t* evals, size_t len, int d, size_t folding,
                            size_t remainder_max_degree, size_t blowup, uint8_t* roots,
                            uint64_t* remainder, size_t* remainder_len, uint64_t* alphas);

// ---- random coin (crypto/src/random/default.rs) ----
typedef struct { uint8_t seed[32]; uint64_t counter; int hash_id; } wfo_coin;
void wfo_coin_new(wfo_coin* c, int hash_id, const uint64_t* seed_elems, size_t n);
void wfo_coin_reseed(wfo_coin* c, const uint8_t data[32]);
int wfo_coin_draw(wfo_coin* c, int d, uint64_t* out);  // 0 ok, -1 failed after 1000 tries
uint32_t wfo_coin_leading_zeros(const wfo_coin* c, uint64_t value);
int wfo_coin_draw_integers(wfo_coin* c, size_t num, size_t domain, uint64_t nonce, uint64_t* out);

#ifdef __cplusplus
}
#endif
