// oracle/wf_oracle.cpp — CPU restatement of the winterfell v0.13.1 STARK hot path.
// TEST INFRASTRUCTURE ONLY (see wf_oracle.h). Every function cites the reference file:line it
// follows. Loops that the reference runs through rayon under its `concurrent` feature are OpenMP
// loops here with the same decomposition, so that the library doubles as the CPU baseline
// ("C++ restatement of winterfell v0.13.1, T threads" — never "reference Rust").
#include "wf_oracle.h"

#include <omp.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <vector>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;

static const u64 P = 0xffffffff00000001ULL;  // math/src/field/f64/mod.rs:46
static const u64 EPS = 0xffffffffULL;        // 2^64 mod p

// =================================================================================================
// FIELD  (math/src/field/f64/mod.rs)
// =================================================================================================
// The reference keeps elements in Montgomery form in memory (mod.rs:57-64) and reduces with
// mont_red_cst (mod.rs:714). Montgomery form is an internal representation: the canonical value of
// every result is the same as plain arithmetic mod p, which is what is restated here.

static inline u64 f_add(u64 a, u64 b) {  // mod.rs:319
    u64 r = a + b;
    if (r < a || r >= P) r -= P;  // a,b < p so a+b < 2p
    return r;
}
static inline u64 f_sub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }  // mod.rs:339
static inline u64 f_neg(u64 a) { return a ? P - a : 0; }
static inline u64 f_red128(u128 x) {
    // x = lo + 2^64*(hl + 2^32*hh); 2^64 = 2^32-1, 2^96 = -1 (mod p)
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & EPS;
    u64 t = lo - hh;
    if (lo < hh) t -= EPS;
    u64 m = hl * EPS;
    u64 r = t + m;
    if (r < m) r += EPS;
    if (r >= P) r -= P;
    return r;
}
static inline u64 f_mul(u64 a, u64 b) { return f_red128((u128)a * b); }  // mod.rs:357
static inline u64 f_dbl(u64 a) { return f_add(a, a); }
static u64 f_exp(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = f_mul(r, a);
        a = f_mul(a, a);
        e >>= 1;
    }
    return r;
}
static u64 f_inv(u64 a) { return a ? f_exp(a, P - 2) : 0; }  // mod.rs:157 (inv(0) = 0)

static const u64 TWO_ADIC_ROOT = 7277203076849721926ULL;  // mod.rs:267
static const u64 GENERATOR = 7;                           // mod.rs:251
// math/src/field/traits.rs get_root_of_unity (f64/mod.rs:258-263)
static u64 root_of_unity(u32 log_n) { return f_exp(TWO_ADIC_ROOT, 1ULL << (32 - log_n)); }

// --- extensions (f64/mod.rs:401-435 quadratic x^2 - x + 2; :443-499 cubic x^3 - x - 1) ---
static inline void e2_mul(const u64* a, const u64* b, u64* o) {  // mod.rs:403-409
    u64 a0b0 = f_mul(a[0], b[0]);
    u64 o0 = f_sub(a0b0, f_dbl(f_mul(a[1], b[1])));
    u64 o1 = f_sub(f_mul(f_add(a[0], a[1]), f_add(b[0], b[1])), a0b0);
    o[0] = o0; o[1] = o1;
}
static inline void e3_mul(const u64* a, const u64* b, u64* o) {  // mod.rs:445-466
    u64 a0b0 = f_mul(a[0], b[0]), a1b1 = f_mul(a[1], b[1]), a2b2 = f_mul(a[2], b[2]);
    u64 s01 = f_mul(f_add(a[0], a[1]), f_add(b[0], b[1]));
    u64 s02 = f_mul(f_add(a[0], a[2]), f_add(b[0], b[2]));
    u64 s12 = f_mul(f_add(a[1], a[2]), f_add(b[1], b[2]));
    u64 a0b0_m_a1b1 = f_sub(a0b0, a1b1);
    u64 o0 = f_sub(f_add(s12, a0b0_m_a1b1), a2b2);
    u64 o1 = f_sub(f_sub(f_add(s01, s12), f_dbl(a1b1)), a0b0);
    u64 o2 = f_sub(s02, a0b0_m_a1b1);
    o[0] = o0; o[1] = o1; o[2] = o2;
}
static inline void e_mul(int d, const u64* a, const u64* b, u64* o) {
    if (d == 1) o[0] = f_mul(a[0], b[0]);
    else if (d == 2) e2_mul(a, b, o);
    else e3_mul(a, b, o);
}
static inline void e_add(int d, const u64* a, const u64* b, u64* o) { for (int i = 0; i < d; i++) o[i] = f_add(a[i], b[i]); }
static inline void e_sub(int d, const u64* a, const u64* b, u64* o) { for (int i = 0; i < d; i++) o[i] = f_sub(a[i], b[i]); }
static inline void e_mul_base(int d, const u64* a, u64 b, u64* o) { for (int i = 0; i < d; i++) o[i] = f_mul(a[i], b); }
static void e2_frob(const u64* x, u64* o) { o[0] = f_add(x[0], x[1]); o[1] = f_neg(x[1]); }  // mod.rs:431
static void e3_frob(const u64* x, u64* o) {  // mod.rs:490-498
    u64 o0 = f_add(x[0], f_add(f_mul(10615703402128488253ULL, x[1]), f_mul(6700183068485440220ULL, x[2])));
    u64 o1 = f_add(f_mul(10050274602728160328ULL, x[1]), f_mul(14531223735771536287ULL, x[2]));
    u64 o2 = f_add(f_mul(11746561000929144102ULL, x[1]), f_mul(8396469466686423992ULL, x[2]));
    o[0] = o0; o[1] = o1; o[2] = o2;
}
static void e_inv(int d, const u64* a, u64* o) {
    if (d == 1) { o[0] = f_inv(a[0]); return; }
    bool zero = true;
    for (int i = 0; i < d; i++) zero = zero && a[i] == 0;
    if (zero) { for (int i = 0; i < d; i++) o[i] = 0; return; }
    if (d == 2) {  // extensions/quadratic.rs:81-94
        u64 num[2], norm[2];
        e2_frob(a, num);
        e2_mul(a, num, norm);
        u64 di = f_inv(norm[0]);
        o[0] = f_mul(num[0], di); o[1] = f_mul(num[1], di);
    } else {  // extensions/cubic.rs:81-97
        u64 c1[3], c2[3], num[3], norm[3];
        e3_frob(a, c1);
        e3_frob(c1, c2);
        e3_mul(c1, c2, num);
        e3_mul(a, num, norm);
        u64 di = f_inv(norm[0]);
        for (int i = 0; i < 3; i++) o[i] = f_mul(num[i], di);
    }
}

// =================================================================================================
// FFT  (math/src/fft)
// =================================================================================================
static inline size_t permute_index(size_t size, size_t index) {  // fft/mod.rs:570-578
    u32 bits = (u32)__builtin_ctzll(size);
    if (bits == 0) return 0;
    u64 r = index;  // index.reverse_bits() >> (usize::BITS - bits)
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    r = ((r >> 2) & 0x3333333333333333ULL) | ((r & 0x3333333333333333ULL) << 2);
    r = ((r >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((r & 0x0f0f0f0f0f0f0f0fULL) << 4);
    r = __builtin_bswap64(r);
    return (size_t)(r >> (64 - bits));
}
static std::vector<u64> power_series(u64 b, size_t n) {  // math/src/utils/mod.rs:36
    std::vector<u64> r(n);
    u64 x = 1;
    for (size_t i = 0; i < n; i++) { r[i] = x; x = f_mul(x, b); }
    return r;
}
// threads available to one of `outer_tasks` concurrently running transforms (rayon's work stealing lets the
// reference's nested par_iter use every core; OpenMP needs the split stated: outer x inner <= threads)
static int inner_threads(size_t outer_tasks) {
    if (omp_in_parallel()) return 1;
    size_t t = (size_t)omp_get_max_threads();
    return (int)std::max<size_t>(1, t / std::max<size_t>(outer_tasks, 1));
}
// fft_inputs.rs permute (swap i <-> bitrev(i)); with nt > 1 the batches of concurrent.rs:103-124 (disjoint
// index ranges; a swap happens only from its smaller index, so no two threads touch the same pair)
static void permute_words(u64* v, size_t n, int d, int nt = 1) {
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1 && n >= 1024)
    for (size_t i = 0; i < n; i++) {
        size_t j = permute_index(n, i);
        if (j > i) for (int k = 0; k < d; k++) std::swap(v[i * d + k], v[j * d + k]);
    }
}
static std::vector<u64> get_twiddles(size_t n) {  // fft/mod.rs:455-468
    if (n < 2) return {};
    std::vector<u64> t = power_series(root_of_unity((u32)__builtin_ctzll(n)), n / 2);
    permute_words(t.data(), n / 2, 1);
    return t;
}
static std::vector<u64> get_inv_twiddles(size_t n) {  // fft/mod.rs:491-505
    if (n < 2) return {};
    u64 root = root_of_unity((u32)__builtin_ctzll(n));
    u64 inv_root = f_exp(root, n - 1);
    std::vector<u64> t = power_series(inv_root, n / 2);
    permute_words(t.data(), n / 2, 1);
    return t;
}
// In-place FFT over bit-reversed twiddles: natural-order input, bit-reversed output. Restates the
// recursion of fft_inputs.rs:215-252 with the butterflies of fft_inputs.rs:107-125
// (butterfly: (a, b) -> (a+b, a-b); butterfly_twiddle: b *= w first). Extension elements use
// mul_base(twiddle), i.e. the same network on every base component.
static void ref_fft_rec(u64* v, size_t vlen, int d, const u64* tw, size_t count, size_t stride, size_t offset) {
    // fft_inputs.rs:215-252
    const size_t MAX_LOOP = 256;
    size_t size = vlen / stride;
    if (size > 2) {
        if (stride == count && count < MAX_LOOP) {
            ref_fft_rec(v, vlen, d, tw, 2 * count, 2 * stride, offset);
        } else {
            ref_fft_rec(v, vlen, d, tw, count, 2 * stride, offset);
            ref_fft_rec(v, vlen, d, tw, count, 2 * stride, offset + stride);
        }
    }
    // butterflies without twiddle (fft_inputs.rs:108-115): (i, j=i+stride) -> (a+b, a-b)
    for (size_t o = offset; o < offset + count; o++) {
        size_t i = o, j = o + stride;
        for (int k = 0; k < d; k++) {
            u64 a = v[i * d + k], b = v[j * d + k];
            v[i * d + k] = f_add(a, b);
            v[j * d + k] = f_sub(a, b);
        }
    }
    size_t last_offset = offset + size * stride;
    size_t idx = 0;
    for (size_t o = offset; o < last_offset; o += 2 * stride, idx++) {
        if (idx == 0) continue;  // .skip(1)
        u64 w = tw[idx];
        for (size_t jj = o; jj < o + count; jj++) {
            size_t i = jj, j = jj + stride;
            for (int k = 0; k < d; k++) {
                u64 a = v[i * d + k];
                u64 b = f_mul(v[j * d + k], w);
                v[i * d + k] = f_add(a, b);
                v[j * d + k] = f_sub(a, b);
            }
        }
    }
}
static void ref_fft_in_place(u64* v, size_t n, int d, const u64* tw) {
    if (n < 2) return;
    ref_fft_rec(v, n, d, tw, 1, 1, 0);  // fft_inputs.rs:103-105: fft_in_place(self, twiddles, 1, 1, 0)
}

// fft/concurrent.rs:131-171 split_radix_fft: the `concurrent` build's transform for >= 1024 points — an
// inner x outer decomposition whose row transforms run in parallel (par_chunks_mut), joined by two serial
// square transpositions (:176-218). Same permuted output as fft_in_place. `tw` is the bit-reversed twiddle
// table of the full size (its prefixes serve the row transforms), nt the threads this transform may use.
static void transpose_square_stretch(u64* m, size_t size, size_t stretch, int d) {  // concurrent.rs:176-218
    const size_t it = stretch * (size_t)d;  // words per transposed item
    for (size_t row = 0; row < size; row++)
        for (size_t col = row + 1; col < size; col++) {
            u64 *a = m + (row * size + col) * it, *b = m + (col * size + row) * it;
            for (size_t k = 0; k < it; k++) std::swap(a[k], b[k]);
        }
}
static void split_radix_fft(u64* v, size_t n, int d, const u64* tw, int nt) {
    const u32 log_n = (u32)__builtin_ctzll(n);
    const u64 g = tw[(n / 2) / 2];                 // "generator of the domain should be in the middle of twiddles"
    const size_t inner_len = (size_t)1 << (log_n / 2), outer_len = n / inner_len, stretch = outer_len / inner_len;
    transpose_square_stretch(v, inner_len, stretch, d);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t r = 0; r < inner_len; r++)          // row.fft_in_place_raw(twiddles, stretch, stretch, 0)
        ref_fft_rec(v + r * outer_len * d, outer_len, d, tw, stretch, stretch, 0);
    transpose_square_stretch(v, inner_len, stretch, d);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t i = 0; i < inner_len; i++) {
        u64* row = v + i * outer_len * d;
        if (i > 0) {
            u64 inner_twiddle = f_exp(g, permute_index(inner_len, i)), outer_twiddle = inner_twiddle;
            for (size_t e = 1; e < outer_len; e++) {
                for (int k = 0; k < d; k++) row[e * d + k] = f_mul(row[e * d + k], outer_twiddle);
                outer_twiddle = f_mul(outer_twiddle, inner_twiddle);
            }
        }
        ref_fft_rec(row, outer_len, d, tw, 1, 1, 0);  // row.fft_in_place(twiddles)
    }
}
// fft/mod.rs:95-110 (and :180, :275, :362): the concurrent implementation is chosen for >= MIN_CONCURRENT_SIZE
// (1024) points when the `concurrent` feature is on — here: when this transform has more than one thread
static void fft_in_place_auto(u64* v, size_t n, int d, const u64* tw, int nt) {
    if (nt > 1 && n >= 1024) split_radix_fft(v, n, d, tw, nt);
    else ref_fft_in_place(v, n, d, tw);
}

static void evaluate_poly(u64* p, size_t n, int d, const u64* tw) {  // fft/serial.rs:18-25, concurrent.rs:17-20
    const int nt = inner_threads(1);
    fft_in_place_auto(p, n, d, tw, nt);
    permute_words(p, n, d, nt);
}
static void interpolate_poly(u64* v, size_t n, int d, const u64* inv_tw) {  // fft/serial.rs:66-76
    u64 inv_len = f_inv((u64)n % P);
    const int nt = inner_threads(1);                 // concurrent.rs:59-70
    fft_in_place_auto(v, n, d, inv_tw, nt);
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1 && n >= 1024)
    for (size_t i = 0; i < n * d; i++) v[i] = f_mul(v[i], inv_len);  // shift_by
    permute_words(v, n, d, nt);
}
static void interpolate_poly_with_offset(u64* v, size_t n, int d, const u64* inv_tw, u64 domain_offset) {
    // fft/serial.rs:84-101; concurrent.rs:77-98 (batches, each starting from its own power of the offset)
    const int nt = inner_threads(1);
    fft_in_place_auto(v, n, d, inv_tw, nt);
    permute_words(v, n, d, nt);
    const u64 inc = f_inv(domain_offset), inv_len = f_inv((u64)n % P);
    const size_t nb = (nt > 1 && n >= 1024) ? (size_t)nt : 1, bs = (n + nb - 1) / nb;
#pragma omp parallel for schedule(static) num_threads(nt) if (nb > 1)
    for (size_t bi = 0; bi < nb; bi++) {
        const size_t lo = bi * bs, hi = std::min(n, lo + bs);
        u64 off = f_mul(f_exp(inc, lo), inv_len);
        for (size_t i = lo; i < hi; i++) {  // shift_by_series(offset, increment)
            for (int k = 0; k < d; k++) v[i * d + k] = f_mul(v[i * d + k], off);
            off = f_mul(off, inc);
        }
    }
}
static void evaluate_poly_with_offset(const u64* p, size_t n, int d, const u64* tw, u64 domain_offset,
                                      size_t blowup, u64* result) {
    // fft/serial.rs:29-56 (chunks are independent: concurrent.rs:26-49 runs them via rayon)
    size_t domain = n * blowup;
    u64 g = root_of_unity((u32)__builtin_ctzll(domain));
    const int total = inner_threads(1), nt_in = inner_threads(blowup), nt_out = (int)std::min<size_t>(blowup, (size_t)total);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt_out)
    for (size_t i = 0; i < blowup; i++) {
        u64* chunk = result + i * n * d;
        u64 idx = permute_index(blowup, i);
        u64 offset = f_mul(f_exp(g, idx), domain_offset);
        u64 factor = 1;
        for (size_t j = 0; j < n; j++) {  // clone_and_shift (concurrent.rs:223-236)
            for (int k = 0; k < d; k++) chunk[j * d + k] = f_mul(p[j * d + k], factor);
            factor = f_mul(factor, offset);
        }
        fft_in_place_auto(chunk, n, d, tw, nt_in);
    }
    permute_words(result, domain, d, total);
}

// polynom::eval (math/src/polynom/mod.rs:55-62): Horner, coefficients of degree dp, point of degree dx
static void eval_poly_at(const u64* p, size_t n, int dp, const u64* x, int dx, u64* out) {
    int d = std::max(dp, dx);
    u64 acc[3] = {0, 0, 0}, xe[3] = {0, 0, 0};
    for (int k = 0; k < dx; k++) xe[k] = x[k];
    for (size_t i = n; i-- > 0;) {
        u64 t[3];
        e_mul(d, acc, xe, t);
        u64 c[3] = {0, 0, 0};
        for (int k = 0; k < dp; k++) c[k] = p[i * dp + k];
        e_add(d, t, c, acc);
    }
    for (int k = 0; k < d; k++) out[k] = acc[k];
}

// =================================================================================================
// MATRICES  (prover/src/matrix)
// =================================================================================================
static void interpolate_columns(u64* cols, size_t c, size_t n, int d) {  // col_matrix.rs:192-202
    std::vector<u64> inv_tw = get_inv_twiddles(n);
    // iter_mut!(columns) over fft::interpolate_poly, itself concurrent (nested rayon): columns x inner threads
    const int total = inner_threads(1), nt_in = inner_threads(c), nt_out = (int)std::min<size_t>(c, (size_t)total);
    const u64 inv_len = f_inv((u64)n % P);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt_out)
    for (size_t j = 0; j < c; j++) {
        u64* v = cols + j * n * d;
        fft_in_place_auto(v, n, d, inv_tw.data(), nt_in);
        for (size_t i = 0; i < n * d; i++) v[i] = f_mul(v[i], inv_len);
        permute_words(v, n, d, nt_in);
    }
}

// row_matrix.rs:84-100 evaluate_polys_over::<8> + get_evaluation_offsets :238-271 +
// segments.rs:96-158 (per coset: scale coefficients by offset powers, FFT of the [[B;8]] batch,
// then one bit-reversal over the whole buffer) + transpose :298-343. The reference batches base
// columns 8 at a time purely for cache behaviour; element (row i, base column q) of the result is
// P_q(7 * w_N^i) regardless of batching, and the padding columns of the last segment are never
// hashed (row_matrix.rs:162-166), so the oracle emits the unpadded row width c*d.
static void lde_rows(const u64* polys, size_t c, size_t n, int d, size_t blowup, u64* out) {
    size_t N = n * blowup, w = c * d;
    std::vector<u64> tw = get_twiddles(n);
    u64 g = root_of_unity((u32)__builtin_ctzll(N));
    // offsets[k][j] = (g^bitrev_b(k) * 7)^j   (row_matrix.rs:238-271)
    std::vector<u64> seg_off(blowup);
    for (size_t k = 0; k < blowup; k++) seg_off[k] = f_mul(f_exp(g, permute_index(blowup, k)), GENERATOR);
    size_t num_seg = (w + 7) / 8;
    // segments.rs:127-141 (`concurrent`): par_chunks_mut over the cosets, split_radix_fft inside each
    const int total = inner_threads(1), nt_in = inner_threads(num_seg * blowup), nt_out = (int)std::min<size_t>(num_seg * blowup, (size_t)total);
#pragma omp parallel for collapse(2) schedule(dynamic, 1) num_threads(nt_out)
    for (size_t s = 0; s < num_seg; s++) {
        for (size_t k = 0; k < blowup; k++) {  // segments.rs:130 par_chunks_mut(poly_size)
            size_t q0 = s * 8, q1 = std::min(w, q0 + 8), nq = q1 - q0;
            std::vector<u64> buf(n * nq);
            u64 factor = 1;
            for (size_t j = 0; j < n; j++) {  // copy_polys segments.rs:178-190
                for (size_t q = q0; q < q1; q++) {
                    // base column q = component (q % d) of column (q / d)
                    u64 coef = polys[(q / d) * n * d + j * d + (q % d)];
                    buf[j * nq + (q - q0)] = f_mul(coef, factor);
                }
                factor = f_mul(factor, seg_off[k]);
            }
            fft_in_place_auto(buf.data(), n, (int)nq, tw.data(), nt_in);
            // whole-buffer bit reversal (segments.rs:279-301) maps (coset k, bitrev position) to
            // natural row order: row = permute_index(N, k*n + pos)
            for (size_t pos = 0; pos < n; pos++) {
                size_t row = permute_index(N, k * n + pos);
                for (size_t q = q0; q < q1; q++) out[row * w + q] = buf[pos * nq + (q - q0)];
            }
        }
    }
}

// =================================================================================================
// BLAKE3  (third-party crate `blake3 = "1.8"`, crypto/Cargo.toml:34; public spec restated)
// =================================================================================================
static const u32 B3_IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                             0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const int B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };
static inline u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void b3_g(u32* s, int a, int b, int c, int d, u32 mx, u32 my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr32(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr32(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr32(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr32(s[b] ^ s[c], 7);
}
static void b3_compress(const u32 cv[8], const u32 block[16], u64 counter, u32 block_len, u32 flags, u32 out[16]) {
    u32 s[16], m[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = B3_IV[i];
    s[12] = (u32)counter; s[13] = (u32)(counter >> 32); s[14] = block_len; s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        b3_g(s, 0, 4, 8, 12, m[0], m[1]);   b3_g(s, 1, 5, 9, 13, m[2], m[3]);
        b3_g(s, 2, 6, 10, 14, m[4], m[5]);  b3_g(s, 3, 7, 11, 15, m[6], m[7]);
        b3_g(s, 0, 5, 10, 15, m[8], m[9]);  b3_g(s, 1, 6, 11, 12, m[10], m[11]);
        b3_g(s, 2, 7, 8, 13, m[12], m[13]); b3_g(s, 3, 4, 9, 14, m[14], m[15]);
        u32 t[16];
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
}
// chaining value of one chunk (<= 1024 bytes); if is_root, ROOT is set on the last block
static void b3_chunk_cv(const u8* data, size_t len, u64 chunk_counter, bool is_root, u32 out_cv[8]) {
    u32 cv[8];
    memcpy(cv, B3_IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        u8 blk[64] = {0};
        size_t bl = std::min((size_t)64, len - b * 64);
        if (len) memcpy(blk, data + b * 64, bl); else bl = 0;
        u32 m[16];
        memcpy(m, blk, 64);
        u32 flags = 0;
        if (b == 0) flags |= B3_CHUNK_START;
        if (b == nblocks - 1) { flags |= B3_CHUNK_END; if (is_root) flags |= B3_ROOT; }
        u32 o[16];
        b3_compress(cv, m, chunk_counter, (u32)bl, flags, o);
        memcpy(cv, o, 32);
    }
    memcpy(out_cv, cv, 32);
}
static void b3_parent_cv(const u32 l[8], const u32 r[8], bool is_root, u32 out_cv[8]) {
    u32 m[16], o[16];
    memcpy(m, l, 32); memcpy(m + 8, r, 32);
    b3_compress(B3_IV, m, 0, 64, B3_PARENT | (is_root ? B3_ROOT : 0), o);
    memcpy(out_cv, o, 32);
}
static void b3_subtree(const u8* data, size_t len, u64 chunk0, bool is_root, u32 out_cv[8]) {
    if (len <= 1024) { b3_chunk_cv(data, len, chunk0, is_root, out_cv); return; }
    // left subtree = largest power-of-two number of chunks strictly less than the total
    size_t chunks = (len + 1023) / 1024;
    size_t left = 1;
    while (left * 2 < chunks) left *= 2;
    u32 l[8], r[8];
    b3_subtree(data, left * 1024, chunk0, false, l);
    b3_subtree(data + left * 1024, len - left * 1024, chunk0 + left, false, r);
    b3_parent_cv(l, r, is_root, out_cv);
}
static void blake3_hash(const u8* data, size_t len, u8 out[32]) {
    u32 cv[8];
    b3_subtree(data, len, 0, true, cv);
    memcpy(out, cv, 32);  // little-endian words
}

// =================================================================================================
// RESCUE PRIME Rp64_256  (crypto/src/hash/rescue/rp64_256/mod.rs)
// =================================================================================================
#include "rp64_constants.inc"
static inline u64 f_exp7(u64 x) { u64 x2 = f_mul(x, x), x4 = f_mul(x2, x2), x3 = f_mul(x2, x); return f_mul(x3, x4); }  // f64/mod.rs:96
static void rp_mds(u64 s[12]) {
    // mds_f64_12x12.rs:41 computes the same circulant product with an integer-FFT trick
    // (rp64_256/tests.rs:205-222 proves it equal to the naive product, restated here).
    u64 r[12];
    for (int i = 0; i < 12; i++) {
        u128 acc = 0;
        for (int j = 0; j < 12; j++) acc += (u128)RP64_MDS_ROW0[(j + 12 - i) % 12] * s[j];
        r[i] = f_red128(acc % ((u128)P << 32));  // acc < 12*26*2^64; fold safely
    }
    memcpy(s, r, sizeof(r));
}
static void rp_exp_acc(const u64 base[12], const u64 tail[12], int m, u64 out[12]) {  // rescue/mod.rs:20-28
    for (int i = 0; i < 12; i++) {
        u64 r = base[i];
        for (int k = 0; k < m; k++) r = f_mul(r, r);
        out[i] = f_mul(r, tail[i]);
    }
}
static void rp_inv_sbox(u64 s[12]) {  // mod.rs:351-385
    u64 t1[12], t2[12], t3[12], t4[12], t5[12], t6[12], t7[12];
    for (int i = 0; i < 12; i++) { t1[i] = f_mul(s[i], s[i]); t2[i] = f_mul(t1[i], t1[i]); }
    rp_exp_acc(t2, t2, 3, t3);
    rp_exp_acc(t3, t3, 6, t4);
    rp_exp_acc(t4, t4, 12, t5);
    rp_exp_acc(t5, t3, 6, t6);
    rp_exp_acc(t6, t6, 31, t7);
    for (int i = 0; i < 12; i++) {
        u64 a = f_mul(f_mul(t7[i], t7[i]), t6[i]);
        a = f_mul(a, a); a = f_mul(a, a);
        u64 b = f_mul(f_mul(t1[i], t2[i]), s[i]);
        s[i] = f_mul(a, b);
    }
}
static void rp_permute(u64 s[12]) {  // mod.rs:299-321
    for (int r = 0; r < 7; r++) {
        for (int i = 0; i < 12; i++) s[i] = f_exp7(s[i]);
        rp_mds(s);
        for (int i = 0; i < 12; i++) s[i] = f_add(s[i], RP64_ARK1[r][i]);
        rp_inv_sbox(s);
        rp_mds(s);
        for (int i = 0; i < 12; i++) s[i] = f_add(s[i], RP64_ARK2[r][i]);
    }
}
static void rp_digest_bytes(const u64 s[12], u8 out[32]) { memcpy(out, s + 4, 32); }  // digest.rs:36-45 (LE host)
static void rp_hash_elements(const u64* e, size_t n, u8 out[32]) {  // mod.rs:224-257
    u64 s[12] = {0};
    s[0] = (u64)n % P;
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        s[4 + i] = f_add(s[4 + i], e[k]);
        i++;
        if (i % 8 == 0) { rp_permute(s); i = 0; }
    }
    if (i > 0) rp_permute(s);
    rp_digest_bytes(s, out);
}
static void rp_merge(const u8 two[64], u8 out[32]) {  // mod.rs:181-192
    u64 s[12] = {0};
    memcpy(s + 4, two, 64);
    s[0] = 8;
    rp_permute(s);
    rp_digest_bytes(s, out);
}
static void rp_merge_with_int(const u8 seed[32], u64 value, u8 out[32]) {  // mod.rs:198-218
    u64 s[12] = {0};
    memcpy(s + 4, seed, 32);
    if (value < P) { s[8] = value; s[0] = 5; }
    else { s[8] = value - P; /* BaseElement::new(value) reduces */ s[9] = value / P; s[0] = 6; }
    rp_permute(s);
    rp_digest_bytes(s, out);
}

// =================================================================================================
// RESCUE PRIME RpJive64_256  (crypto/src/hash/rescue/rp64_256_jive/mod.rs): 8-element state, Jive compression
// =================================================================================================
#include "rpjive_constants.inc"
static void rpj_mds(u64 s[8]) {  // mds_f64_8x8.rs: circulant with first row RPJ_MDS_ROW0 (naive product, as for Rp64 above)
    u64 r[8];
    for (int i = 0; i < 8; i++) {
        u128 acc = 0;
        for (int j = 0; j < 8; j++) acc += (u128)RPJ_MDS_ROW0[(j + 8 - i) % 8] * s[j];
        r[i] = f_red128(acc % ((u128)P << 32));
    }
    memcpy(s, r, sizeof(r));
}
static u64 rpj_inv7(u64 x) {  // mod.rs:378-412: x^10540996611094048183 by plain square-and-multiply (the reference uses a chain)
    u64 e = 10540996611094048183ULL, r = 1, b = x;
    while (e) { if (e & 1) r = f_mul(r, b); b = f_mul(b, b); e >>= 1; }
    return r;
}
static void rpj_permute(u64 s[8]) {  // mod.rs:313-333
    for (int r = 0; r < 7; r++) {
        for (int i = 0; i < 8; i++) s[i] = f_exp7(s[i]);
        rpj_mds(s);
        for (int i = 0; i < 8; i++) s[i] = f_add(s[i], RPJ_ARK1[r][i]);
        for (int i = 0; i < 8; i++) s[i] = rpj_inv7(s[i]);
        rpj_mds(s);
        for (int i = 0; i < 8; i++) s[i] = f_add(s[i], RPJ_ARK2[r][i]);
    }
}
static void rpj_jive(const u64 init[8], u8 out[32]) {  // permutation + apply_jive_summation (mod.rs:337-350)
    u64 s[8], d[4];
    memcpy(s, init, sizeof(s));
    rpj_permute(s);
    for (int i = 0; i < 4; i++) d[i] = f_add(f_add(init[i], init[4 + i]), f_add(s[i], s[4 + i]));
    memcpy(out, d, 32);
}
static void rpj_hash_elements(const u64* e, size_t n, u8 out[32]) {  // mod.rs:240-282
    u64 s[8] = {0};
    if (n % 4 != 0) s[0] = 1;
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        s[4 + i] = f_add(s[4 + i], e[k]);
        i++;
        if (i % 4 == 0) { rpj_permute(s); i = 0; }
    }
    if (i > 0) {
        s[4 + i] = 1;
        i++;
        while (i != 4) { s[4 + i] = 0; i++; }
        rpj_permute(s);
    }
    memcpy(out, s + 4, 32);
}
static void rpj_merge(const u8 two[64], u8 out[32]) {  // mod.rs:186-196
    u64 s[8];
    memcpy(s, two, 64);
    rpj_jive(s, out);
}
static void rpj_merge_with_int(const u8 seed[32], u64 value, u8 out[32]) {  // mod.rs:206-229
    u64 s[8] = {0};
    memcpy(s, seed, 32);
    s[4] = value % P;
    if (value < P) s[7] = 5;
    else { s[5] = value / P; s[7] = 6; }
    rpj_jive(s, out);
}

// =================================================================================================
// HASHER DISPATCH  (crypto/src/hash/mod.rs:31-64)
// =================================================================================================
// =================================================================================================
// SHA3-256  (third-party crate `sha3 = "0.10"`, crypto/Cargo.toml; FIPS 202 restated, byte-oriented)
// =================================================================================================
static void keccak_f(u64 st[25]) {
    static const u64 RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
                               0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
                               0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
                               0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                               0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int r = 0; r < 24; r++) {
        u64 bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            u64 t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        u64 t = st[1];
        for (int i = 0; i < 24; i++) { int j = PIL[i]; u64 b = st[j]; st[j] = (t << ROT[i]) | (t >> (64 - ROT[i])); t = b; }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[r];
    }
}
static void sha3_256(const u8* data, size_t len, u8 out[32]) {
    u64 st[25] = {0};
    u8* sb = (u8*)st;  // little-endian host
    size_t i = 0;
    for (size_t k = 0; k < len; k++) { sb[i++] ^= data[k]; if (i == 136) { keccak_f(st); i = 0; } }
    sb[i] ^= 0x06;
    sb[135] ^= 0x80;
    keccak_f(st);
    memcpy(out, st, 32);
}

// Digests live in 32-byte slots everywhere in this file; Blake3_192 (blake/mod.rs:73-123) keeps the first 24 bytes of the BLAKE3
// output (ByteDigest<24>; the slot's last 8 bytes are zero, as ByteDigest::as_bytes pads them) and hashes / serializes 24
// bytes per digest.
static size_t digest_len(int h) { return h == WFO_HASH_BLAKE3_192 ? 24 : 32; }
static void b3_192(const u8* data, size_t len, u8 out[32]) { blake3_hash(data, len, out); memset(out + 24, 0, 8); }
static void hash_elements(int h, const u64* e, size_t n, u8 out[32]) {
    if (h == WFO_HASH_BLAKE3_256) blake3_hash((const u8*)e, n * 8, out);  // blake/mod.rs:52-65: canonical LE bytes
    else if (h == WFO_HASH_BLAKE3_192) b3_192((const u8*)e, n * 8, out);  // :108-115
    else if (h == WFO_HASH_SHA3_256) sha3_256((const u8*)e, n * 8, out);  // sha/mod.rs:49-55
    else if (h == WFO_HASH_RP64_256) rp_hash_elements(e, n, out);
    else rpj_hash_elements(e, n, out);
}
static void merge_many(int h, const u8* dg, size_t n, u8 out[32]);
static void merge(int h, const u8 two[64], u8 out[32]) {
    if (h == WFO_HASH_BLAKE3_256) blake3_hash(two, 64, out);  // blake/mod.rs:33
    else if (h == WFO_HASH_BLAKE3_192) merge_many(h, two, 2, out);  // :85-88: BLAKE3 of the 48 digest bytes
    else if (h == WFO_HASH_SHA3_256) sha3_256(two, 64, out);         // sha/mod.rs:30-32
    else if (h == WFO_HASH_RP64_256) rp_merge(two, out);
    else rpj_merge(two, out);
}
static void merge_many(int h, const u8* dg, size_t n, u8 out[32]) {
    if (h == WFO_HASH_BLAKE3_256) blake3_hash(dg, n * 32, out);  // blake/mod.rs:37
    else if (h == WFO_HASH_BLAKE3_192) {                          // :90-93 digests_as_bytes: 24 bytes each, back to back
        std::vector<u8> cat(n * 24);
        for (size_t i = 0; i < n; i++) memcpy(cat.data() + 24 * i, dg + 32 * i, 24);
        b3_192(cat.data(), cat.size(), out);
    }
    else if (h == WFO_HASH_SHA3_256) sha3_256(dg, n * 32, out);                     // sha/mod.rs:34-36
    else if (h == WFO_HASH_RP64_256) rp_hash_elements((const u64*)dg, n * 4, out);  // rp64_256/mod.rs:194
    else rpj_hash_elements((const u64*)dg, n * 4, out);                             // rp64_256_jive/mod.rs:198-200
}
static void merge_with_int(int h, const u8 seed[32], u64 value, u8 out[32]) {
    if (h == WFO_HASH_BLAKE3_256) {  // blake/mod.rs:41-46
        u8 data[40];
        memcpy(data, seed, 32);
        memcpy(data + 32, &value, 8);
        blake3_hash(data, 40, out);
    } else if (h == WFO_HASH_BLAKE3_192) {  // :95-102: 24 seed bytes + 8 value bytes
        u8 data[32];
        memcpy(data, seed, 24);
        memcpy(data + 24, &value, 8);
        b3_192(data, 32, out);
    } else if (h == WFO_HASH_SHA3_256) {  // sha/mod.rs:38-43
        u8 data[40];
        memcpy(data, seed, 32);
        memcpy(data + 32, &value, 8);
        sha3_256(data, 40, out);
    } else if (h == WFO_HASH_RP64_256) rp_merge_with_int(seed, value, out);
    else rpj_merge_with_int(seed, value, out);
}

static void hash_rows(int h, const u64* rows, size_t nrows, size_t w, size_t part, u8* digests) {
    // row_matrix.rs:184-228 (batch_iter_mut! with min batch 128 -> static row batches)
    if (part == w || part == 0) {  // partition_size == num_cols (row_matrix.rs:193); a partition size ABOVE the width (hash_rate
                                    // larger than the row) takes the other branch with one partition: merge_many of one digest
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nrows; i++) hash_elements(h, rows + i * w, w, digests + i * 32);
    } else {
        size_t np = (w + part - 1) / part;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nrows; i++) {
            std::vector<u8> buf(np * 32);
            for (size_t j = 0; j < np; j++) {
                size_t q0 = j * part, q1 = std::min(w, q0 + part);
                hash_elements(h, rows + i * w + q0, q1 - q0, buf.data() + j * 32);
            }
            merge_many(h, buf.data(), np, digests + i * 32);
        }
    }
}

// =================================================================================================
// MERKLE  (crypto/src/merkle)
// =================================================================================================
static void merkle_nodes(int h, const u8* leaves, size_t nleaves, u8* nodes) {
    // mod.rs:344-368 serial; concurrent.rs:26-75 splits into subtrees (same nodes)
    size_t n = nleaves / 2;
    memset(nodes, 0, 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) merge(h, leaves + i * 64, nodes + (n + i) * 32);
    for (size_t lvl = n / 2; lvl >= 1; lvl >>= 1) {
#pragma omp parallel for schedule(static) if (lvl >= 256)
        for (size_t i = lvl; i < 2 * lvl; i++) merge(h, nodes + 2 * i * 32, nodes + i * 32);
    }
}

static void write_vint64(std::vector<u8>& o, u64 v) {  // utils/core/src/serde/byte_writer.rs:77-92
    int lz = v == 0 ? 64 : __builtin_clzll(v);
    int len = 9 - std::min((lz > 0 ? lz - 1 : 0) / 7, 8);  // usize_encoded_len :145-149 (saturating_sub)
    if (len == 9) {
        o.push_back(0);
        for (int i = 0; i < 8; i++) o.push_back((u8)(v >> (8 * i)));
    } else {
        u64 enc = ((v << 1) | 1) << (len - 1);
        for (int i = 0; i < len; i++) o.push_back((u8)(enc >> (8 * i)));
    }
}

static long merkle_prove_batch(const u8* leaves, const u8* nodes, size_t nleaves, const u64* indexes,
                               size_t k, u8* leaves_out, u8* out, size_t cap, size_t dlen = 32) {
    // mod.rs:217-272
    if (k == 0) return -1;
    size_t depth = (size_t)__builtin_ctzll(nleaves);
    std::map<size_t, size_t> index_map;  // map_indexes :371-388
    for (size_t i = 0; i < k; i++) {
        if (indexes[i] >= nleaves) return -1;
        index_map[indexes[i]] = i;
    }
    if (index_map.size() != k) return -1;
    std::set<size_t> norm;  // normalize_indexes :390-396
    for (size_t i = 0; i < k; i++) norm.insert(indexes[i] - (indexes[i] & 1));
    std::vector<std::vector<const u8*>> pn;
    std::vector<size_t> next;
    for (size_t index : norm) {
        std::vector<const u8*> missing;
        for (size_t i = index; i < index + 2; i++) {
            auto it = index_map.find(i);
            if (it != index_map.end()) memcpy(leaves_out + it->second * 32, leaves + i * 32, 32);
            else missing.push_back(leaves + i * 32);
        }
        pn.push_back(missing);
        next.push_back((index + nleaves) >> 1);
    }
    for (size_t lvl = 1; lvl < depth; lvl++) {
        std::vector<size_t> idx = next;
        next.clear();
        size_t i = 0;
        while (i < idx.size()) {
            size_t sib = idx[i] ^ 1;
            if (i + 1 < idx.size() && idx[i + 1] == sib) i += 1;
            else pn[i].push_back(nodes + sib * 32);
            next.push_back(sib >> 1);
            i += 1;
        }
    }
    // proofs.rs:390-401: u8 depth, usize(vint64) #vectors, each Vec<Digest> = vint64 len + digests
    std::vector<u8> o;
    o.push_back((u8)depth);
    write_vint64(o, pn.size());
    for (auto& v : pn) {
        write_vint64(o, v.size());
        for (const u8* dg : v) o.insert(o.end(), dg, dg + dlen);   // digests serialize to their own length (ByteDigest<N>)
    }
    if (o.size() > cap) return -1;
    memcpy(out, o.data(), o.size());
    return (long)o.size();
}

// =================================================================================================
// RANDOM COIN  (crypto/src/random/default.rs)
// =================================================================================================
static void coin_next(wfo_coin* c, u8 out[32]) {  // :82-85
    c->counter += 1;
    merge_with_int(c->hash_id, c->seed, c->counter, out);
}
static int coin_draw(wfo_coin* c, int d, u64* out) {  // :156-170; f64/mod.rs:576-600 try_from rejects >= p
    for (int t = 0; t < 1000; t++) {
        u8 dg[32];
        coin_next(c, dg);
        u64 w[3];
        memcpy(w, dg, 8 * d);
        bool ok = true;
        for (int k = 0; k < d; k++) ok = ok && w[k] < P;
        if (ok) { for (int k = 0; k < d; k++) out[k] = w[k]; return 0; }
    }
    return -1;
}

// =================================================================================================
// FRI  (fri/src)
// =================================================================================================
static void transpose_slice(const u64* src, size_t len, int d, size_t nf, u64* dst) {
    // utils/core/src/lib.rs:166-185: result[i][j] = source[i + j * row_count]
    size_t rows = len / nf;
#pragma omp parallel for schedule(static) if (rows >= 1024)
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < nf; j++)
            for (int k = 0; k < d; k++) dst[(i * nf + j) * d + k] = src[(i + j * rows) * d + k];
}
static void apply_drp(const u64* tv, size_t rows, int d, size_t nf, u64 domain_offset, const u64* alpha, u64* out) {
    // folding/mod.rs:86-118
    size_t n = rows * nf;
    u64 g = root_of_unity((u32)__builtin_ctzll(n));
    u64 ginv = f_inv(g), oinv = f_inv(domain_offset);
    std::vector<u64> inv_tw = get_inv_twiddles(nf);
    u64 len_offset = f_inv((u64)nf);
#pragma omp parallel
    {
        std::vector<u64> poly(nf * d);
#pragma omp for schedule(static)
        for (size_t i = 0; i < rows; i++) {
            // get_inv_offsets :181-188: offset^-1 * g^-i
            u64 xinv = f_mul(oinv, f_exp(ginv, i));
            memcpy(poly.data(), tv + i * nf * d, nf * d * 8);
            ref_fft_in_place(poly.data(), nf, d, inv_tw.data());  // serial_fft = fft_in_place + permute
            permute_words(poly.data(), nf, d);
            u64 off = len_offset;
            for (size_t j = 0; j < nf; j++) {
                for (int k = 0; k < d; k++) poly[j * d + k] = f_mul(poly[j * d + k], off);
                off = f_mul(off, xinv);
            }
            eval_poly_at(poly.data(), nf, d, alpha, d, out + i * d);
        }
    }
}
static size_t fold_positions(const u64* pos, size_t k, size_t source, size_t nf, u64* out) {  // folding/mod.rs:159-176
    size_t target = source / nf, cnt = 0;
    for (size_t i = 0; i < k; i++) {
        u64 p = pos[i] % target;
        bool dup = false;
        for (size_t j = 0; j < cnt; j++) dup = dup || out[j] == p;
        if (!dup) out[cnt++] = p;
    }
    return cnt;
}
static size_t fri_num_layers(size_t domain, size_t nf, size_t rem_max_deg, size_t blowup) {  // fri/src/options.rs:85-93
    size_t r = 0, max_rem = (rem_max_deg + 1) * blowup;
    while (domain > max_rem) { domain /= nf; r++; }
    return r;
}

// =================================================================================================
// C API
// =================================================================================================
extern "C" {
void wfo_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); omp_set_max_active_levels(2); }
// split_radix_fft (fft/concurrent.rs:131-171) with nt threads, for the parity test against the serial network
void wfo_split_radix_fft(uint64_t* v, size_t n, int d, int inverse, int nt) {
    omp_set_max_active_levels(2);
    auto t = inverse ? get_inv_twiddles(n) : get_twiddles(n);
    split_radix_fft(v, n, d, t.data(), nt);
}
void wfo_fft_in_place(uint64_t* v, size_t n, int d, int inverse) {
    auto t = inverse ? get_inv_twiddles(n) : get_twiddles(n);
    ref_fft_in_place(v, n, d, t.data());
}
int wfo_get_threads(void) { return omp_get_max_threads(); }
uint64_t wfo_add(uint64_t a, uint64_t b) { return f_add(a, b); }
uint64_t wfo_sub(uint64_t a, uint64_t b) { return f_sub(a, b); }
uint64_t wfo_mul(uint64_t a, uint64_t b) { return f_mul(a, b); }
uint64_t wfo_inv(uint64_t a) { return f_inv(a); }
uint64_t wfo_exp(uint64_t a, uint64_t e) { return f_exp(a, e); }
uint64_t wfo_to_mont(uint64_t a) { return f_red128((u128)a << 64); }                 // f64/mod.rs:72-76 (x * R mod p)
uint64_t wfo_from_mont(uint64_t a) { return f_mul(a, f_inv(f_red128((u128)1 << 64))); }  // mont_to_int :731
uint64_t wfo_root_of_unity(uint32_t log_n) { return root_of_unity(log_n); }
void wfo_ext_mul(int d, const uint64_t* a, const uint64_t* b, uint64_t* out) { u64 t[3]; e_mul(d, a, b, t); memcpy(out, t, 8 * d); }
void wfo_ext_inv(int d, const uint64_t* a, uint64_t* out) { u64 t[3]; e_inv(d, a, t); memcpy(out, t, 8 * d); }

void wfo_get_twiddles(size_t n, uint64_t* out) { auto t = get_twiddles(n); memcpy(out, t.data(), t.size() * 8); }
void wfo_get_inv_twiddles(size_t n, uint64_t* out) { auto t = get_inv_twiddles(n); memcpy(out, t.data(), t.size() * 8); }
void wfo_evaluate_poly(uint64_t* p, size_t n, int d) { auto t = get_twiddles(n); evaluate_poly(p, n, d, t.data()); }
void wfo_interpolate_poly(uint64_t* v, size_t n, int d) { auto t = get_inv_twiddles(n); interpolate_poly(v, n, d, t.data()); }
void wfo_evaluate_poly_with_offset(const uint64_t* p, size_t n, int d, uint64_t offset, size_t blowup, uint64_t* out) {
    auto t = get_twiddles(n);
    evaluate_poly_with_offset(p, n, d, t.data(), offset, blowup, out);
}
void wfo_interpolate_poly_with_offset(uint64_t* v, size_t n, int d, uint64_t offset) {
    auto t = get_inv_twiddles(n);
    interpolate_poly_with_offset(v, n, d, t.data(), offset);
}
void wfo_eval_poly_at(const uint64_t* p, size_t n, int dp, const uint64_t* x, int dx, uint64_t* out) { eval_poly_at(p, n, dp, x, dx, out); }
void wfo_interpolate_columns(uint64_t* cols, size_t c, size_t n, int d) { interpolate_columns(cols, c, n, d); }
void wfo_lde_rows(const uint64_t* polys, size_t c, size_t n, int d, size_t blowup, uint64_t* out) { lde_rows(polys, c, n, d, blowup, out); }

void wfo_blake3(const uint8_t* data, size_t len, uint8_t out[32]) { blake3_hash(data, len, out); }
void wfo_rp64_permute(uint64_t state[12]) { rp_permute(state); }
void wfo_rpjive_permute(uint64_t state[8]) { rpj_permute(state); }
void wfo_hash_elements(int h, const uint64_t* e, size_t n, uint8_t out[32]) { hash_elements(h, e, n, out); }
void wfo_merge(int h, const uint8_t two[64], uint8_t out[32]) { merge(h, two, out); }
void wfo_merge_many(int h, const uint8_t* dg, size_t n, uint8_t out[32]) { merge_many(h, dg, n, out); }
void wfo_merge_with_int(int h, const uint8_t seed[32], uint64_t v, uint8_t out[32]) { merge_with_int(h, seed, v, out); }
void wfo_hash_rows(int h, const uint64_t* rows, size_t nrows, size_t w, size_t part, uint8_t* dg) { hash_rows(h, rows, nrows, w, part, dg); }
void wfo_merkle_nodes(int h, const uint8_t* leaves, size_t nleaves, uint8_t* nodes) { merkle_nodes(h, leaves, nleaves, nodes); }
long wfo_merkle_prove_batch(const uint8_t* leaves, const uint8_t* nodes, size_t nleaves, const uint64_t* idx, size_t k,
                            uint8_t* leaves_out, uint8_t* out, size_t cap) {
    return merkle_prove_batch(leaves, nodes, nleaves, idx, k, leaves_out, out, cap);
}
// same for a hasher whose digests serialize to fewer than 32 bytes (Blake3_192: 24)
long wfo_merkle_prove_batch_h(int hash_id, const uint8_t* leaves, const uint8_t* nodes, size_t nleaves, const uint64_t* idx, size_t k,
                              uint8_t* leaves_out, uint8_t* out, size_t cap) {
    return merkle_prove_batch(leaves, nodes, nleaves, idx, k, leaves_out, out, cap, digest_len(hash_id));
}

void wfo_transpose_slice(const uint64_t* src, size_t len, int d, size_t nf, uint64_t* dst) { transpose_slice(src, len, d, nf, dst); }
void wfo_apply_drp(const uint64_t* tv, size_t rows, int d, size_t nf, uint64_t off, const uint64_t* alpha, uint64_t* out) {
    apply_drp(tv, rows, d, nf, off, alpha, out);
}
size_t wfo_fold_positions(const uint64_t* pos, size_t k, size_t src, size_t nf, uint64_t* out) { return fold_positions(pos, k, src, nf, out); }
size_t wfo_fri_num_layers(size_t domain, size_t nf, size_t rem_max_deg, size_t blowup) { return fri_num_layers(domain, nf, rem_max_deg, blowup); }

void wfo_coin_new(wfo_coin* c, int h, const uint64_t* seed, size_t n) { c->hash_id = h; c->counter = 0; hash_elements(h, seed, n, c->seed); }
void wfo_coin_reseed(wfo_coin* c, const uint8_t data[32]) {  // default.rs:131-134
    u8 two[64];
    memcpy(two, c->seed, 32); memcpy(two + 32, data, 32);
    merge(c->hash_id, two, c->seed);
    c->counter = 0;
}
int wfo_coin_draw(wfo_coin* c, int d, uint64_t* out) { return coin_draw(c, d, out); }
uint32_t wfo_coin_leading_zeros(const wfo_coin* c, uint64_t value) {  // default.rs:141-146
    u8 dg[32];
    merge_with_int(c->hash_id, c->seed, value, dg);
    u64 head;
    memcpy(&head, dg, 8);
    return head == 0 ? 64 : (uint32_t)__builtin_ctzll(head);
}
int wfo_coin_draw_integers(wfo_coin* c, size_t num, size_t domain, uint64_t nonce, uint64_t* out) {  // default.rs:210-247
    u8 s[32];
    merge_with_int(c->hash_id, c->seed, nonce, s);
    memcpy(c->seed, s, 32);
    c->counter = 0;
    u64 mask = (u64)domain - 1;
    size_t cnt = 0;
    for (int t = 0; t < 1000 && cnt < num; t++) {
        u8 dg[32];
        coin_next(c, dg);
        u64 v;
        memcpy(&v, dg, 8);
        out[cnt++] = v & mask;
    }
    return cnt == num ? 0 : -1;
}

size_t wfo_fri_build_layers(int h, const uint64_t* evals, size_t len, int d, size_t nf, size_t rem_max_deg,
                            size_t blowup, uint8_t* roots, uint64_t* remainder, size_t* remainder_len,
                            uint64_t* alphas) {
    // fri/src/prover/mod.rs:179-239 with DefaultProverChannel (fri/src/prover/channel.rs): coin = new(&[])
    wfo_coin coin;
    wfo_coin_new(&coin, h, nullptr, 0);
    std::vector<u64> cur(evals, evals + len * d);
    size_t cur_len = len, nl = 0;
    size_t max_rem = (rem_max_deg + 1) * blowup;  // fri/src/options.rs:85-93
    while (cur_len > max_rem) {
        size_t rows = cur_len / nf;
        std::vector<u64> tv(cur_len * d);
        transpose_slice(cur.data(), cur_len, d, nf, tv.data());
        std::vector<u8> dg(rows * 32), nodes(rows * 32);
        hash_rows(h, tv.data(), rows, nf * d, nf * d, dg.data());  // build_layer_commitment :321-336
        merkle_nodes(h, dg.data(), rows, nodes.data());
        memcpy(roots + nl * 32, nodes.data() + 32, 32);
        wfo_coin_reseed(&coin, nodes.data() + 32);  // commit_fri_layer
        u64 alpha[3];
        coin_draw(&coin, d, alpha);  // draw_fri_alpha
        if (alphas) memcpy(alphas + nl * d, alpha, 8 * d);
        std::vector<u64> nxt(rows * d);
        apply_drp(tv.data(), rows, d, nf, GENERATOR, alpha, nxt.data());
        cur.swap(nxt);
        cur_len = rows;
        nl++;
    }
    // set_remainder :230-239
    auto inv_tw = get_inv_twiddles(cur_len);
    interpolate_poly_with_offset(cur.data(), cur_len, d, inv_tw.data(), GENERATOR);
    size_t rsize = cur_len / blowup;
    for (size_t i = 0; i < rsize; i++)
        for (int k = 0; k < d; k++) remainder[i * d + k] = cur[(rsize - 1 - i) * d + k];
    hash_elements(h, remainder, rsize * d, roots + nl * 32);
    *remainder_len = rsize;
    return nl;
}
}  // extern "C"
