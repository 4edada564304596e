// winterfell_b200.hpp — C++ host mirror of the reference's prover-side plugin interface, written ONLY
// against the C ABI of winterfell_b200.h (no CUDA headers, builds with plain g++).
//
// The reference selects its hot path through associated types on `Prover` (prover/src/lib.rs:125-223):
// TraceLde, ConstraintEvaluator, ConstraintCommitment; FRI, the DEEP composer and the channel are
// concrete types used by the provided method `generate_proof` (lib.rs:282-492). A Rust shim crate would
// implement exactly these types over the C ABI; no Rust toolchain exists in this image,
// so this header plays that role in C++ with the same type and method names, the same argument meaning
// and the same order of calls, and `generate_proof` below is the call-for-call double of the reference's
// method. tests/shim/generate_proof_main.cpp drives it; tests/test_gpu_shim.py checks that it emits
// byte-identical proofs to the one-call entry points (wf_prove_air / wf_prove_air_aux) and the oracle.
//
// Elements cross this interface as canonical u64 words (the C ABI's `mont` flag is exposed where the
// reference hands over `&[BaseElement]` memory: TraceLde::new and set_aux_trace).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "winterfell_b200.h"

namespace winterfell_b200 {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
// the reference's traits are infallible and panic on misuse (trace_lde/default/mod.rs:150-158): a
// non-zero status becomes an exception carrying wf_last_error
inline void check(wf_ctx* ctx, int rc) {
    if (rc != WF_OK) throw Error(rc, ctx ? wf_last_error(ctx) : "winterfell_b200: call failed");
}

// ---- field elements of E (math/src/field/f64/mod.rs:401-499), host side: only what the channel needs ----
static const u64 P = 0xFFFFFFFF00000001ULL;
inline u64 f_add(u64 a, u64 b) { u64 s = a + b; return (s < a || s >= P) ? s - P : s; }
inline u64 f_sub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
inline u64 f_mul(u64 a, u64 b) { return wf_host_mul(a, b); }
struct Elem {
    u64 v[3] = {0, 0, 0};
};
inline Elem e_mul(int d, const Elem& a, const Elem& b) {
    Elem r;
    if (d == 1) { r.v[0] = f_mul(a.v[0], b.v[0]); return r; }
    if (d == 2) {  // x^2 = x - 2 (f64/mod.rs:401-435)
        u64 t0 = f_mul(a.v[0], b.v[0]), t1 = f_mul(a.v[1], b.v[1]);
        r.v[0] = f_sub(t0, f_add(t1, t1));
        r.v[1] = f_add(f_add(f_mul(a.v[0], b.v[1]), f_mul(a.v[1], b.v[0])), t1);
        return r;
    }
    // x^3 = x + 1 (f64/mod.rs:443-499)
    u64 t0 = f_mul(a.v[0], b.v[0]);
    u64 t1 = f_add(f_mul(a.v[0], b.v[1]), f_mul(a.v[1], b.v[0]));
    u64 t2 = f_add(f_add(f_mul(a.v[0], b.v[2]), f_mul(a.v[1], b.v[1])), f_mul(a.v[2], b.v[0]));
    u64 t3 = f_add(f_mul(a.v[1], b.v[2]), f_mul(a.v[2], b.v[1]));
    u64 t4 = f_mul(a.v[2], b.v[2]);
    r.v[0] = f_add(t0, t3);
    r.v[1] = f_add(f_add(t1, t3), t4);
    r.v[2] = f_add(t2, t4);
    return r;
}

// ---- air::ProofOptions (air/src/options.rs:132) ----
enum class BatchingMethod : u32 { Linear = 0, Algebraic = 1, Horner = 2 };
// air::PartitionOptions (air/src/options.rs:405-445)
struct PartitionOptions {
    u32 num_partitions = 1, hash_rate = 1;
    // partition_size::<E>(num_columns), E of extension degree `ext_degree`, in columns of E (:428-438)
    u32 partition_size(u32 ext_degree, u32 num_columns) const {
        if (num_partitions == 1) return num_columns;
        const u32 min_partition_size = hash_rate / ext_degree, per = (num_columns + num_partitions - 1) / num_partitions;
        return per > min_partition_size ? per : min_partition_size;
    }
};
struct ProofOptions {
    u32 num_queries = 28, blowup_factor = 8, grinding_factor = 0, field_extension = 1, fri_folding_factor = 4,
        fri_remainder_max_degree = 7;
    BatchingMethod batching_constraints = BatchingMethod::Linear, batching_deep = BatchingMethod::Linear;
    int hash_id = WF_HASH_BLAKE3_256;
    PartitionOptions partition_options;  // ProofOptions::with_partitions (options.rs:193-200)
};

// ---- Air (air/src/air/mod.rs:174): the flat description the device evaluator consumes ----
struct TransitionConstraintDegree {  // air/src/air/transition/degree.rs
    u32 base;
    std::vector<u32> cycles;
};
struct Air {
    std::vector<u64> desc;  // as documented at wf_prove_air / wf_prove_air_aux
    u32 trace_width = 0, aux_width = 0, num_aux_rands = 0, num_exemptions = 1;
    std::vector<TransitionConstraintDegree> degrees, aux_degrees;
    size_t num_assertions = 0, num_aux_assertions = 0, num_periodic = 0;
    std::vector<u64> pub_inputs;

    // parses the counts the host side needs out of the flat description
    explicit Air(const std::vector<u64>& d) : desc(d) {
        size_t p = 0;
        auto rd = [&]() { if (p >= desc.size()) throw Error(WF_ERR_INVALID, "truncated AIR description"); return desc[p++]; };
        auto degs = [&](std::vector<TransitionConstraintDegree>& out) {
            u64 cnt = rd();
            for (u64 i = 0; i < cnt; i++) {
                TransitionConstraintDegree t;
                t.base = (u32)rd();
                u64 nc = rd();
                for (u64 j = 0; j < nc; j++) t.cycles.push_back((u32)rd());
                out.push_back(t);
            }
        };
        trace_width = (u32)rd();
        degs(degrees);
        num_periodic = rd();
        for (size_t i = 0; i < num_periodic; i++) { u64 ln = rd(); p += ln; }
        { u64 nc = rd(); p += nc; }
        rd();  // num_regs
        { u64 ni = rd(); p += 4 * ni; }
        num_assertions = rd();
        for (size_t i = 0; i < num_assertions; i++) { p += 3; u64 nv = rd(); p += nv; }  // column, first_step, stride, nvals, values
        { u64 np = rd(); for (u64 i = 0; i < np; i++) pub_inputs.push_back(rd()); }
        num_exemptions = (u32)rd();
        if (p == desc.size()) return;
        aux_width = (u32)rd();
        num_aux_rands = (u32)rd();
        degs(aux_degrees);
        rd();
        { u64 ni = rd(); p += 4 * ni; }
        num_aux_assertions = rd();
        for (size_t i = 0; i < num_aux_assertions; i++) { p += 3; u64 nv = rd(); p += 3 * nv; }
        if (p != desc.size()) throw Error(WF_ERR_INVALID, "malformed AIR description");
    }
    size_t num_transition_constraints() const { return degrees.size() + aux_degrees.size(); }  // context.rs:205
    size_t num_all_assertions() const { return num_assertions + num_aux_assertions; }         // context.rs:223
    u32 ce_blowup_factor() const {  // context.rs:87-100
        u32 r = 2;
        auto one = [&](const TransitionConstraintDegree& t) {
            u32 bound = t.base + (u32)t.cycles.size() - 1, p2 = 1;
            while (p2 < bound) p2 <<= 1;
            r = std::max(r, std::max(p2, 2u));
        };
        for (auto& t : degrees) one(t);
        for (auto& t : aux_degrees) one(t);
        return r;
    }
    u32 num_constraint_composition_columns(size_t n) const {  // context.rs:265-285
        size_t hi = 0;
        auto one = [&](const TransitionConstraintDegree& t) {
            size_t e = (size_t)t.base * (n - 1);
            for (u32 c : t.cycles) e += (n / c) * (c - 1);
            hi = std::max(hi, e);
        };
        for (auto& t : degrees) one(t);
        for (auto& t : aux_degrees) one(t);
        size_t div = n - num_exemptions;
        return (u32)std::max((hi - div + n - 1) / n, (size_t)1);
    }
};

// ---- utils::ByteWriter (utils/core/src/serde/byte_writer.rs) ----
struct ByteWriter {
    std::vector<u8> v;
    void write_u8(u8 x) { v.push_back(x); }
    void write_u16(uint16_t x) { for (int i = 0; i < 2; i++) v.push_back((u8)(x >> (8 * i))); }
    void write_u64(u64 x) { for (int i = 0; i < 8; i++) v.push_back((u8)(x >> (8 * i))); }
    void write_bytes(const void* p, size_t n) { const u8* q = (const u8*)p; v.insert(v.end(), q, q + n); }
    void write_usize(u64 value) {  // vint64 (:77-92)
        int zeros = value == 0 ? 64 : __builtin_clzll(value);
        int len = zeros == 0 ? 9 : std::max(1, 9 - (zeros - 1) / 7);
        if (len >= 9) { write_u8(0); write_u64(value); return; }
        u64 enc = ((value << 1) | 1) << (len - 1);
        for (int i = 0; i < len; i++) v.push_back((u8)(enc >> (8 * i)));
    }
    void write_elems(int d, const std::vector<Elem>& e) { for (auto& x : e) for (int q = 0; q < d; q++) write_u64(x.v[q]); }
};

// ---- crypto::DefaultRandomCoin (crypto/src/random/default.rs:95-247) over the ABI's host hashers ----
struct RandomCoin {
    int hash_id;
    u8 seed[32];
    u64 counter = 0;
    RandomCoin(int h, const std::vector<u64>& seed_elems) : hash_id(h) { wf_host_hash_elements(h, seed_elems.data(), seed_elems.size(), seed); }
    void reseed(const u8 data[32]) {
        u8 ab[64];
        memcpy(ab, seed, 32);
        memcpy(ab + 32, data, 32);
        wf_host_merge(hash_id, ab, seed);
        counter = 0;
    }
    void next(u8 out[32]) { counter += 1; wf_host_merge_with_int(hash_id, seed, counter, out); }
    Elem draw(int d) {
        for (int t = 0; t < 1000; t++) {
            u8 b[32];
            next(b);
            Elem e;
            memcpy(e.v, b, 8 * d);
            bool ok = true;
            for (int k = 0; k < d; k++) ok = ok && e.v[k] < P;
            if (ok) return e;
        }
        throw Error(WF_ERR_STATE, "RandomCoinError::FailedToDrawFieldElement");
    }
    u32 check_leading_zeros(u64 value) const {
        u8 b[32];
        wf_host_merge_with_int(hash_id, seed, value, b);
        u64 head;
        memcpy(&head, b, 8);
        return head == 0 ? 64 : (u32)__builtin_ctzll(head);
    }
    std::vector<u64> draw_integers(size_t num, size_t domain, u64 nonce) {
        u8 s[32];
        wf_host_merge_with_int(hash_id, seed, nonce, s);
        memcpy(seed, s, 32);
        counter = 0;
        std::vector<u64> out;
        for (int t = 0; t < 1000 && out.size() < num; t++) {
            u8 b[32];
            next(b);
            u64 x;
            memcpy(&x, b, 8);
            out.push_back(x & ((u64)domain - 1));
        }
        if (out.size() != num) throw Error(WF_ERR_STATE, "RandomCoinError::FailedToDrawIntegers");
        return out;
    }
};

// ---- air::proof::Queries (air/src/proof/queries.rs:51-78,138-146) ----
struct Queries {
    std::vector<u8> values, opening_proof;
    void write_into(ByteWriter& w) const {
        w.write_usize(values.size()); w.write_bytes(values.data(), values.size());
        w.write_usize(opening_proof.size()); w.write_bytes(opening_proof.data(), opening_proof.size());
    }
};
inline Queries build_queries(wf_ctx* ctx, const wf_mat* lde, const wf_tree* tree, const std::vector<u64>& positions) {
    Queries q;
    const size_t k = positions.size(), cols = wf_mat_cols(lde);
    q.values.resize(k * cols * 8);
    check(ctx, wf_mat_read_rows(ctx, lde, positions.data(), k, (u64*)q.values.data(), 0));
    std::vector<u8> leaves(k * 32);
    q.opening_proof.resize(64 + k * 40 * 33);
    size_t len = q.opening_proof.size();
    check(ctx, wf_tree_open_many(ctx, tree, positions.data(), k, leaves.data(), q.opening_proof.data(), &len));
    q.opening_proof.resize(len);
    return q;
}

// ---- TraceLde (prover/src/trace/trace_lde/mod.rs:26-76; DefaultTraceLde trace_lde/default/mod.rs) ----
struct TracePolyTable {  // prover/src/trace/poly_table.rs
    wf_mat* main_polys = nullptr;
    wf_mat* aux_polys = nullptr;
};
class TraceLde {
  public:
    wf_ctx* ctx;
    int hash_id;
    u32 log_n, blowup_factor, ext;
    PartitionOptions partition_options;
    wf_mat *main_lde = nullptr, *aux_lde = nullptr;
    wf_tree *main_tree = nullptr, *aux_tree = nullptr;
    TracePolyTable polys;

    // DefaultTraceLde::new (:63-100): interpolate, extend, commit the main segment
    TraceLde(wf_ctx* c, int h, const u64* const* main_trace_cols, u32 width, u32 log_n_, u32 blowup, u32 ext_, int mont,
             PartitionOptions partition_options_ = PartitionOptions())
        : ctx(c), hash_id(h), log_n(log_n_), blowup_factor(blowup), ext(ext_), partition_options(partition_options_) {
        check(ctx, wf_trace_lde_from_host(ctx, main_trace_cols, width, (size_t)1 << log_n, mont, log2(blowup), &polys.main_polys, &main_lde));
        // build_trace_commitment::<E, E::BaseField, ..> (:71): the main segment is a matrix over the base field
        check(ctx, wf_commit_rows_partitioned(ctx, hash_id, main_lde, partition_options.partition_size(1, width), &main_tree));
    }
    ~TraceLde() {
        for (wf_mat* m : {main_lde, aux_lde, polys.main_polys, polys.aux_polys}) if (m) wf_mat_free(ctx, m);
        for (wf_tree* t : {main_tree, aux_tree}) if (t) wf_tree_free(ctx, t);
    }
    TraceLde(const TraceLde&) = delete;
    void get_main_trace_commitment(u8 root[32]) const { check(ctx, wf_tree_root(ctx, main_tree, root)); }
    // set_aux_trace (:140-166): aux columns over E, [aux_width][n][ext] words; returns the commitment
    void set_aux_trace(const u64* aux_cols, u32 aux_width, int mont, u8 root[32]) {
        if (aux_lde) throw Error(WF_ERR_STATE, "the auxiliary trace has already been added");
        const size_t n = (size_t)1 << log_n;
        std::vector<const u64*> cols(aux_width);
        for (u32 j = 0; j < aux_width; j++) cols[j] = aux_cols + (size_t)j * n * ext;
        wf_mat* trace;
        check(ctx, wf_mat_from_host_columns(ctx, cols.data(), aux_width, n, (int)ext, mont, &trace));
        check(ctx, wf_mat_interpolate(ctx, trace, &polys.aux_polys));
        wf_mat_free(ctx, trace);
        check(ctx, wf_mat_lde(ctx, polys.aux_polys, log2(blowup_factor), &aux_lde));
        check(ctx, wf_commit_rows_partitioned(ctx, hash_id, aux_lde, partition_options.partition_size(ext, aux_width) * ext, &aux_tree));
        check(ctx, wf_tree_root(ctx, aux_tree, root));
    }
    // read_main_trace_frame_into (:169-180): rows lde_step and (lde_step + blowup) mod N
    void read_main_trace_frame_into(size_t lde_step, std::vector<u64>& current, std::vector<u64>& next) const {
        const size_t w = wf_mat_cols(main_lde);
        u64 pos[2] = {lde_step, (lde_step + blowup_factor) % trace_len()};
        std::vector<u64> rows(2 * w);
        check(ctx, wf_mat_read_rows(ctx, main_lde, pos, 2, rows.data(), 0));
        current.assign(rows.begin(), rows.begin() + w);
        next.assign(rows.begin() + w, rows.end());
    }
    std::vector<Queries> query(const std::vector<u64>& positions) const {  // :199-218
        std::vector<Queries> r;
        r.push_back(build_queries(ctx, main_lde, main_tree, positions));
        if (aux_lde) r.push_back(build_queries(ctx, aux_lde, aux_tree, positions));
        return r;
    }
    size_t trace_len() const { return (size_t)1 << (log_n + log2(blowup_factor)); }
    u32 blowup() const { return blowup_factor; }
    static u32 log2(u32 x) { u32 l = 0; while ((1u << l) < x) l++; return l; }
};

// ---- ConstraintEvaluator (prover/src/constraints/evaluator/mod.rs:28-42) ----
struct ConstraintCompositionCoefficients {  // air/src/air/coefficients.rs:72
    std::vector<Elem> transition, boundary;
};
struct CompositionPolyTrace {  // prover/src/constraints/composition_poly.rs:23
    wf_ctx* ctx;
    wf_mat* evaluations;
};
class ConstraintEvaluator {
  public:
    const Air& air;
    std::vector<Elem> aux_rand_elements;
    ConstraintCompositionCoefficients coefficients;
    // Prover::new_evaluator (lib.rs:195-202)
    ConstraintEvaluator(const Air& a, const std::vector<Elem>& aux_rand, const ConstraintCompositionCoefficients& cc)
        : air(a), aux_rand_elements(aux_rand), coefficients(cc) {}
    CompositionPolyTrace evaluate(const TraceLde& trace) const {  // default.rs:60-118
        const int d = (int)trace.ext;
        std::vector<u64> cc, rnd;
        for (auto& e : coefficients.transition) for (int q = 0; q < d; q++) cc.push_back(e.v[q]);
        for (auto& e : coefficients.boundary) for (int q = 0; q < d; q++) cc.push_back(e.v[q]);
        for (auto& e : aux_rand_elements) for (int q = 0; q < d; q++) rnd.push_back(e.v[q]);
        CompositionPolyTrace t{trace.ctx, nullptr};
        check(trace.ctx, wf_eval_constraints(trace.ctx, air.desc.data(), air.desc.size(), trace.log_n, trace.blowup_factor, trace.ext,
                                             trace.main_lde, trace.aux_lde, cc.data(), rnd.empty() ? nullptr : rnd.data(), &t.evaluations));
        return t;
    }
};

// ---- ConstraintCommitment + CompositionPoly (prover/src/constraints/commitment/mod.rs:24-37) ----
class ConstraintCommitment {
  public:
    wf_ctx* ctx;
    wf_mat *composition_poly = nullptr, *lde = nullptr;  // CompositionPoly columns; their LDE
    wf_tree* tree = nullptr;
    // Prover::build_constraint_commitment (lib.rs:215-223); consumes the composition trace
    ConstraintCommitment(CompositionPolyTrace trace, int hash_id, u32 num_columns, u32 log_n, u32 blowup, u32 ext,
                         PartitionOptions partition_options = PartitionOptions())
        : ctx(trace.ctx) {
        int rc = wf_composition_commit_partitioned(ctx, hash_id, trace.evaluations, log_n, blowup, ext, num_columns,
                                                   partition_options.partition_size(ext, num_columns) * ext, &composition_poly, &lde, &tree);
        wf_mat_free(ctx, trace.evaluations);
        check(ctx, rc);
    }
    ~ConstraintCommitment() {
        if (composition_poly) wf_mat_free(ctx, composition_poly);
        if (lde) wf_mat_free(ctx, lde);
        if (tree) wf_tree_free(ctx, tree);
    }
    ConstraintCommitment(const ConstraintCommitment&) = delete;
    void commitment(u8 root[32]) const { check(ctx, wf_tree_root(ctx, tree, root)); }
    Queries query(const std::vector<u64>& positions) const { return build_queries(ctx, lde, tree, positions); }
};

// ---- ProverChannel (prover/src/channel.rs:57-210) ----
class ProverChannel {
  public:
    const Air& air;
    ProofOptions options;
    u32 log_n;
    RandomCoin public_coin;
    ByteWriter commitments;  // air/src/proof/commitments.rs: trace roots, constraint root, FRI roots
    ByteWriter ood_trace_states, ood_quotient_states;
    u64 pow_nonce = 0;

    static std::vector<u64> context_elements(const Air& air, const ProofOptions& o, u32 log_n) {
        // Context::to_elements (air/src/proof/context.rs:119-136), TraceInfo (trace_info.rs:209-238),
        // ProofOptions (options.rs:294-305)
        const u64 w = air.trace_width;
        std::vector<u64> e;
        e.push_back(air.aux_width ? (((((w << 8) | 1) << 8) | air.aux_width) << 8) | air.num_aux_rands : (w << 8));
        e.push_back((u64)1 << log_n);
        e.push_back(1);
        e.push_back(0xFFFFFFFFULL);
        e.push_back(air.num_transition_constraints() + air.num_all_assertions());
        e.push_back(((u64)o.field_extension << 24) | ((u64)o.fri_folding_factor << 16) | ((u64)o.fri_remainder_max_degree << 8) | o.blowup_factor);
        e.push_back(o.grinding_factor);
        e.push_back(o.num_queries);
        for (u64 v : air.pub_inputs) e.push_back(v);
        return e;
    }
    ProverChannel(const Air& a, const ProofOptions& o, u32 log_n_)
        : air(a), options(o), log_n(log_n_), public_coin(o.hash_id, context_elements(a, o, log_n_)) {}

    int d() const { return (int)options.field_extension; }
    // digests cross the ABI in 32-byte slots; they serialize to the hasher's digest size (24 bytes for Blake3_192)
    size_t digest_bytes() const { return options.hash_id == WF_HASH_BLAKE3_192 ? 24 : 32; }
    void commit_trace(const u8 root[32]) { commitments.write_bytes(root, digest_bytes()); public_coin.reseed(root); }        // :88-91
    void commit_constraints(const u8 root[32]) { commitments.write_bytes(root, digest_bytes()); public_coin.reseed(root); }  // :94-97
    void commit_fri_layer(const u8 root[32]) { commitments.write_bytes(root, digest_bytes()); public_coin.reseed(root); }    // :215-219
    Elem draw_fri_alpha() { return public_coin.draw(d()); }                                                      // :222-224
    std::vector<Elem> get_aux_rand_elements() {  // Air::get_aux_rand_elements (air/src/air/mod.rs:292-306)
        std::vector<Elem> r;
        for (u32 i = 0; i < air.num_aux_rands; i++) r.push_back(public_coin.draw(d()));
        return r;
    }
    std::vector<Elem> draw_coefficients(BatchingMethod m, size_t n) {  // air/src/air/coefficients.rs:201-218
        std::vector<Elem> r;
        if (m == BatchingMethod::Linear) { for (size_t i = 0; i < n; i++) r.push_back(public_coin.draw(d())); return r; }
        Elem alpha = public_coin.draw(d()), x;
        x.v[0] = 1;
        for (size_t i = 0; i < n; i++) { r.push_back(x); x = e_mul(d(), x, alpha); }
        if (m == BatchingMethod::Horner) std::reverse(r.begin(), r.end());
        return r;
    }
    ConstraintCompositionCoefficients get_constraint_composition_coeffs() {  // :118-122
        std::vector<Elem> all = draw_coefficients(options.batching_constraints, air.num_transition_constraints() + air.num_all_assertions());
        ConstraintCompositionCoefficients cc;
        cc.transition.assign(all.begin(), all.begin() + air.num_transition_constraints());
        cc.boundary.assign(all.begin() + air.num_transition_constraints(), all.end());
        return cc;
    }
    Elem get_ood_point() { return public_coin.draw(d()); }  // :127-129
    // send_ood_evaluations (:102-113): frames into the proof, merged evaluations into the coin
    void send_ood_evaluations(const std::vector<Elem>& t_cur, const std::vector<Elem>& t_next, const std::vector<Elem>& q_cur,
                              const std::vector<Elem>& q_next) {
        ood_trace_states.write_u8(2); ood_trace_states.write_elems(d(), t_cur); ood_trace_states.write_elems(d(), t_next);
        ood_quotient_states.write_u8(2); ood_quotient_states.write_elems(d(), q_cur); ood_quotient_states.write_elems(d(), q_next);
        ByteWriter m;  // merge_ood_evaluations (air/src/proof/ood_frame.rs:335-349)
        m.write_elems(d(), t_cur); m.write_elems(d(), q_cur); m.write_elems(d(), t_next); m.write_elems(d(), q_next);
        u8 dg[32];
        wf_host_hash_elements(options.hash_id, (const u64*)m.v.data(), m.v.size() / 8, dg);
        public_coin.reseed(dg);
    }
    std::vector<Elem> get_deep_composition_coeffs(size_t trace_width, size_t num_quotients) {  // :134-138
        return draw_coefficients(options.batching_deep, trace_width + num_quotients);
    }
    void grind_query_seed(wf_ctx* ctx) {  // :169-184 (serial semantics: smallest nonce)
        check(ctx, wf_grind(ctx, options.hash_id, public_coin.seed, options.grinding_factor, &pow_nonce));
    }
    std::vector<u64> get_query_positions() {  // :151-166
        std::vector<u64> pos = public_coin.draw_integers(options.num_queries, ((size_t)1 << log_n) * options.blowup_factor, pow_nonce);
        std::sort(pos.begin(), pos.end());
        pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
        return pos;
    }
    // build_proof (:187-210) + Proof::write_into (air/src/proof/mod.rs:189-200)
    std::vector<u8> build_proof(const std::vector<Queries>& trace_queries, const Queries& constraint_queries,
                                const std::vector<u8>& fri_proof, size_t num_unique_queries) {
        ByteWriter w;
        const ProofOptions& o = options;
        w.write_u8((u8)air.trace_width); w.write_u8((u8)air.aux_width); w.write_u8((u8)air.num_aux_rands); w.write_u8((u8)log_n);
        w.write_u16(0);
        w.write_u8(8); w.write_u64(P);
        w.write_u8((u8)o.num_queries); w.write_u8((u8)o.blowup_factor); w.write_u8((u8)o.grinding_factor); w.write_u8((u8)o.field_extension);
        w.write_u8((u8)o.fri_folding_factor); w.write_u8((u8)o.fri_remainder_max_degree); w.write_u8((u8)o.batching_constraints);
        w.write_u8((u8)o.batching_deep); w.write_u8((u8)o.partition_options.num_partitions); w.write_u8((u8)o.partition_options.hash_rate);
        w.write_usize(air.num_transition_constraints() + air.num_all_assertions());
        w.write_u8((u8)num_unique_queries);
        w.write_u16((uint16_t)commitments.v.size()); w.write_bytes(commitments.v.data(), commitments.v.size());
        for (auto& q : trace_queries) q.write_into(w);
        constraint_queries.write_into(w);
        w.write_u16((uint16_t)ood_trace_states.v.size()); w.write_bytes(ood_trace_states.v.data(), ood_trace_states.v.size());
        w.write_u16((uint16_t)ood_quotient_states.v.size()); w.write_bytes(ood_quotient_states.v.data(), ood_quotient_states.v.size());
        w.write_bytes(fri_proof.data(), fri_proof.size());
        w.write_u64(pow_nonce);
        return w.v;
    }
};

// ---- DeepCompositionPoly (prover/src/composer/mod.rs) ----
struct DeepCompositionPoly {
    wf_ctx* ctx;
    wf_mat* evaluations;  // DeepCompositionPoly::evaluate (:171): N x ext
};

// ---- FriProver (fri/src/prover/mod.rs:179-319) ----
class FriProver {
  public:
    wf_ctx* ctx;
    wf_fri* fri = nullptr;
    explicit FriProver(wf_ctx* c) : ctx(c) {}
    ~FriProver() { if (fri) wf_fri_free(ctx, fri); }
    void build_layers(ProverChannel& channel, DeepCompositionPoly evaluations) {
        const ProofOptions& o = channel.options;
        auto commit = [](void* u, const u8 root[32]) { ((ProverChannel*)u)->commit_fri_layer(root); };
        auto draw = [](void* u, u64* alpha) {
            ProverChannel* ch = (ProverChannel*)u;
            Elem a = ch->draw_fri_alpha();
            for (int q = 0; q < ch->d(); q++) alpha[q] = a.v[q];
        };
        int rc = wf_fri_build_layers(ctx, o.hash_id, evaluations.evaluations, (int)o.field_extension, o.fri_folding_factor,
                                     o.fri_remainder_max_degree, o.blowup_factor, commit, draw, &channel, &fri);
        wf_mat_free(ctx, evaluations.evaluations);
        check(ctx, rc);
    }
    std::vector<u8> build_proof(const std::vector<u64>& positions) {  // serialized FriProof (fri/src/proof.rs)
        std::vector<u8> out((size_t)1 << 22);
        size_t len = out.size();
        check(ctx, wf_fri_build_proof(ctx, fri, positions.data(), positions.size(), out.data(), &len));
        out.resize(len);
        return out;
    }
};

// Prover::build_aux_trace (prover/src/lib.rs:236-247): aux columns [aux_width][n][ext] from the random elements
typedef std::function<std::vector<u64>(const std::vector<Elem>& aux_rand_elements)> AuxTraceBuilder;

// ---- Prover::generate_proof (prover/src/lib.rs:282-492), call for call ----
inline std::vector<u8> generate_proof(wf_ctx* ctx, const Air& air, const u64* const* main_trace_cols, u32 log_n, const ProofOptions& options,
                                      int mont = 0, AuxTraceBuilder build_aux_trace = nullptr) {
    const u32 d = options.field_extension;
    const size_t n = (size_t)1 << log_n;
    ProverChannel channel(air, options, log_n);                                            // :296-297
    // 1 ----- commit to the execution trace (:304-349)
    TraceLde trace_lde(ctx, options.hash_id, main_trace_cols, air.trace_width, log_n, options.blowup_factor, d, mont, options.partition_options);
    u8 root[32];
    trace_lde.get_main_trace_commitment(root);
    channel.commit_trace(root);
    std::vector<Elem> aux_rand_elements;
    if (air.aux_width) {
        if (!build_aux_trace) throw Error(WF_ERR_INVALID, "multi-segment AIR needs build_aux_trace");
        aux_rand_elements = channel.get_aux_rand_elements();
        std::vector<u64> aux_trace = build_aux_trace(aux_rand_elements);
        if (aux_trace.size() != (size_t)air.aux_width * n * d) throw Error(WF_ERR_INVALID, "aux trace of the wrong shape");
        trace_lde.set_aux_trace(aux_trace.data(), air.aux_width, mont, root);
        channel.commit_trace(root);
    }
    // 2 ----- evaluate constraints (:366-379)
    ConstraintEvaluator evaluator(air, aux_rand_elements, channel.get_constraint_composition_coeffs());
    CompositionPolyTrace composition_poly_trace = evaluator.evaluate(trace_lde);
    // 3 ----- commit to constraint evaluations (:381-384, :527-552)
    const u32 num_quotients = air.num_constraint_composition_columns(n);
    ConstraintCommitment constraint_commitment(composition_poly_trace, options.hash_id, num_quotients, log_n, options.blowup_factor, d,
                                               options.partition_options);
    constraint_commitment.commitment(root);
    channel.commit_constraints(root);
    // 4 ----- build DEEP composition polynomial (:386-440)
    Elem z = channel.get_ood_point();
    Elem g;
    g.v[0] = 1;
    {   // trace domain generator: TWO_ADIC_ROOT^(2^(32 - log_n)) (math/src/field/f64/mod.rs:255-262)
        u64 r = 7277203076849721926ULL;
        for (u32 i = log_n; i < 32; i++) r = f_mul(r, r);
        g.v[0] = r;
    }
    Elem zg = e_mul((int)d, z, g);
    const size_t cw = air.trace_width, aw = air.aux_width, ct = cw + aw;
    std::vector<u64> m_cur(cw * d), m_next(cw * d), a_cur(aw * d), a_next(aw * d), q_cur(num_quotients * d), q_next(num_quotients * d);
    // TracePolyTable::get_ood_frame (poly_table.rs:68-76), CompositionPoly::get_ood_frame (composition_poly.rs:101-108)
    check(ctx, wf_mat_evaluate_at(ctx, trace_lde.polys.main_polys, d, 1, z.v, zg.v, m_cur.data(), m_next.data()));
    if (aw) check(ctx, wf_mat_evaluate_at(ctx, trace_lde.polys.aux_polys, d, d, z.v, zg.v, a_cur.data(), a_next.data()));
    check(ctx, wf_mat_evaluate_at(ctx, constraint_commitment.composition_poly, d, d, z.v, zg.v, q_cur.data(), q_next.data()));
    auto to_elems = [&](const std::vector<u64>& a, const std::vector<u64>& b) {
        std::vector<Elem> r((a.size() + b.size()) / d);
        for (size_t i = 0; i < r.size(); i++)
            for (u32 q = 0; q < d; q++) r[i].v[q] = i * d + q < a.size() ? a[i * d + q] : b[i * d + q - a.size()];
        return r;
    };
    std::vector<Elem> t_cur = to_elems(m_cur, a_cur), t_next = to_elems(m_next, a_next);
    std::vector<Elem> h_cur = to_elems(q_cur, {}), h_next = to_elems(q_next, {});
    channel.send_ood_evaluations(t_cur, t_next, h_cur, h_next);
    std::vector<Elem> deep_coefficients = channel.get_deep_composition_coeffs(ct, num_quotients);
    DeepCompositionPoly deep{ctx, nullptr};
    {
        std::vector<u64> cc, oc, on;
        for (auto& e : deep_coefficients) for (u32 q = 0; q < d; q++) cc.push_back(e.v[q]);
        for (auto* v : {&t_cur, &h_cur}) for (auto& e : *v) for (u32 q = 0; q < d; q++) oc.push_back(e.v[q]);
        for (auto* v : {&t_next, &h_next}) for (auto& e : *v) for (u32 q = 0; q < d; q++) on.push_back(e.v[q]);
        check(ctx, wf_deep_compose(ctx, d, trace_lde.main_lde, trace_lde.aux_lde, constraint_commitment.lde, log_n, z.v, cc.data(), oc.data(),
                                   on.data(), &deep.evaluations));
    }
    // 5 ----- compute FRI layers for the composition polynomial (:442-448)
    FriProver fri_prover(ctx);
    fri_prover.build_layers(channel, deep);
    // 6 ----- determine query positions (:450-462)
    channel.grind_query_seed(ctx);
    std::vector<u64> query_positions = channel.get_query_positions();
    // 7 ----- build proof object (:464-489)
    std::vector<u8> fri_proof = fri_prover.build_proof(query_positions);
    std::vector<Queries> trace_queries = trace_lde.query(query_positions);
    Queries constraint_queries = constraint_commitment.query(query_positions);
    return channel.build_proof(trace_queries, constraint_queries, fri_proof, query_positions.size());
}

}  // namespace winterfell_b200
