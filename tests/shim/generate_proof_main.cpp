// generate_proof_main.cpp — drives include/winterfell_b200.hpp (the C++ mirror of the reference's
// Prover / TraceLde / ConstraintEvaluator / ConstraintCommitment / FriProver / ProverChannel types)
// exactly as a user of the reference drives Prover::prove (prover/src/lib.rs:250-272): build the AIR,
// hand over the trace, get serialized proof bytes. Plain C++ over the C ABI; no CUDA headers.
//
// usage: generate_proof_main <input.bin> <proof.bin>
// input (u64 words): log_n, mont, opts[9], desc_len, desc..., trace [width][n],
//                    has_aux, then (aux replay) rand [nr][d], aux columns [aw][n][d]
// The aux replay stands in for the user's Prover::build_aux_trace: the test records the columns the
// Python builder produced for these random elements and the double checks it was handed the same ones.
#include <stdio.h>

#include <fstream>
#include <iterator>

#include "winterfell_b200.hpp"

using namespace winterfell_b200;

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s input.bin proof.bin\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<u64> w(raw.size() / 8);
    memcpy(w.data(), raw.data(), w.size() * 8);
    size_t p = 0;
    const u32 log_n = (u32)w[p++];
    const int mont = (int)w[p++];
    ProofOptions o;
    o.num_queries = (u32)w[p++]; o.blowup_factor = (u32)w[p++]; o.grinding_factor = (u32)w[p++]; o.field_extension = (u32)w[p++];
    o.fri_folding_factor = (u32)w[p++]; o.fri_remainder_max_degree = (u32)w[p++];
    o.batching_constraints = (BatchingMethod)w[p++]; o.batching_deep = (BatchingMethod)w[p++];
    {   // hash_id | num_partitions << 8 | hash_rate << 16 (the packing of opts[8] in winterfell_b200.h)
        const u64 h = w[p++];
        o.hash_id = (int)(h & 0xff);
        if ((h >> 8) & 0xff) o.partition_options.num_partitions = (u32)((h >> 8) & 0xff);
        if ((h >> 16) & 0xff) o.partition_options.hash_rate = (u32)((h >> 16) & 0xff);
    }
    const size_t desc_len = w[p++];
    std::vector<u64> desc(w.begin() + p, w.begin() + p + desc_len);
    p += desc_len;
    const size_t n = (size_t)1 << log_n, d = o.field_extension;
    wf_ctx* ctx = nullptr;
    try {
        Air air(desc);
        std::vector<const u64*> cols(air.trace_width);
        for (u32 j = 0; j < air.trace_width; j++) cols[j] = &w[p + (size_t)j * n];
        p += (size_t)air.trace_width * n;
        AuxTraceBuilder builder = nullptr;
        if (w[p++]) {
            const u64* rand = &w[p];
            p += (size_t)air.num_aux_rands * d;
            const u64* aux = &w[p];
            const size_t aux_words = (size_t)air.aux_width * n * d;
            builder = [=](const std::vector<Elem>& r) {
                for (size_t i = 0; i < r.size(); i++)
                    for (size_t q = 0; q < d; q++)
                        if (r[i].v[q] != rand[i * d + q]) throw Error(WF_ERR_STATE, "aux random elements differ from the recorded ones");
                return std::vector<u64>(aux, aux + aux_words);
            };
        }
        if (wf_ctx_create(&ctx, 0, nullptr) != WF_OK) { fprintf(stderr, "no CUDA device: this library has no CPU path\n"); return 3; }
        std::vector<u8> proof = generate_proof(ctx, air, cols.data(), log_n, o, mont, builder);
        std::ofstream out(argv[2], std::ios::binary);
        out.write((const char*)proof.data(), (std::streamsize)proof.size());
        // TraceLde accessor used by the reference's evaluator threads (default.rs:187): spot-check one frame
        wf_ctx_destroy(ctx);
        return 0;
    } catch (const Error& e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        if (ctx) wf_ctx_destroy(ctx);
        return 1;
    }
}
