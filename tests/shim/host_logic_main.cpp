// host_logic_main.cpp — the host-side pieces of include/winterfell_b200.hpp that need no GPU (ByteWriter,
// RandomCoin, Air parsing, context elements): prints them for tests/test_host_abi.py to compare with the
// oracle. usage: host_logic_main <air_desc.bin> <hash_id> <ext>
#include <stdio.h>

#include <fstream>
#include <iterator>

#include "winterfell_b200.hpp"

using namespace winterfell_b200;

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<u64> desc(raw.size() / 8);
    memcpy(desc.data(), raw.data(), desc.size() * 8);
    const int hash_id = atoi(argv[2]), d = atoi(argv[3]);
    Air air(desc);
    ProofOptions o;
    o.hash_id = hash_id; o.field_extension = (u32)d; o.num_queries = 30; o.blowup_factor = 8; o.grinding_factor = 20;
    o.fri_folding_factor = 8; o.fri_remainder_max_degree = 127;
    printf("air %u %u %u %zu %zu %u %u\n", air.trace_width, air.aux_width, air.num_aux_rands, air.num_transition_constraints(),
           air.num_all_assertions(), air.ce_blowup_factor(), air.num_constraint_composition_columns(64));
    printf("ctx");
    for (u64 v : ProverChannel::context_elements(air, o, 6)) printf(" %llu", (unsigned long long)v);
    printf("\n");
    ByteWriter w;
    for (u64 v : {0ULL, 1ULL, 255ULL, 234567ULL, ~0ULL}) w.write_usize(v);
    printf("usize");
    for (u8 b : w.v) printf(" %u", b);
    printf("\n");
    ProverChannel ch(air, o, 6);
    u8 root[32];
    for (int i = 0; i < 32; i++) root[i] = (u8)i;
    ch.commit_trace(root);
    printf("draw");
    for (auto& e : ch.draw_coefficients(BatchingMethod::Linear, 5)) for (int q = 0; q < d; q++) printf(" %llu", (unsigned long long)e.v[q]);
    printf("\nalg");
    for (auto& e : ch.draw_coefficients(BatchingMethod::Horner, 4)) for (int q = 0; q < d; q++) printf(" %llu", (unsigned long long)e.v[q]);
    printf("\n");
    return 0;
}
