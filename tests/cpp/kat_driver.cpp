// kat_driver.cpp -- the reference's own known-answer tests for the query path, run from C++
// through include/msbwt_gpu.hpp exactly as a Rust caller would go through the C ABI.
// Built and run by tests/test_gpu_cpp_driver.py on a GPU box.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "msbwt_gpu.hpp"

using msbwt::BWTRange;
using msbwt::convert_stoi;

static int failures = 0;
#define EXPECT(cond)                                                         \
    do {                                                                     \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { printf("usage: kat_driver two_string.npy\n"); return 2; }
    // test_data/two_string.npy: rle_bwt.rs:76-79, dynamic_bwt.rs:783-793, README.md:62-70
    msbwt::RleBWT two;
    two.load_numpy_file(argv[1]);
    EXPECT(two.get_total_size() == 10);
    EXPECT(two.count_kmer(convert_stoi("ACGT")) == 1);
    EXPECT(two.count_kmer(convert_stoi("TGCA")) == 1);
    EXPECT(two.count_kmer(convert_stoi("$")) == 2);

    // bwt "GTN$$ACCC$G" -> [11,13,12,16,9,26,8,11] (bwt_converter.rs:246-256); rle_bwt.rs:601-675
    const std::vector<uint8_t> rle = {11, 13, 12, 16, 9, 26, 8, 11};
    const std::vector<uint8_t> text = convert_stoi("GTN$$ACCC$G");
    msbwt::RleBWT b;
    b.load_vector(rle);
    const uint64_t expected_totals[6] = {3, 1, 3, 2, 1, 1};
    uint64_t start[7] = {0};
    for (int s = 0; s < 6; s++) { EXPECT(b.get_symbol_count(s) == expected_totals[s]); start[s + 1] = start[s] + expected_totals[s]; }
    for (uint8_t sym = 0; sym < 6; sym++) {
        EXPECT((b.constrain_range(sym, BWTRange{0, 11}) == BWTRange{start[sym], start[sym + 1]}));
        uint64_t cnt = 0;
        for (uint64_t ind = 0; ind <= 11; ind++) {
            EXPECT((b.constrain_range(sym, BWTRange{0, ind}) == BWTRange{start[sym], start[sym] + cnt}));
            EXPECT((b.constrain_range(sym, BWTRange{ind, 11}) == BWTRange{start[sym] + cnt, start[sym + 1]}));
            if (ind < 11 && text[ind] == sym) cnt++;
        }
    }
    // doc-tests msbwt_core.rs:110-122: BWT "TG$$CAGCCG" -> bytes T G 2$ C A G 2C G
    msbwt::RleBWT d;
    d.load_vector({13, 11, 16, 10, 9, 11, 18, 11});
    EXPECT(d.get_total_size() == 10);
    EXPECT(d.get_symbol_count(0) == 2);
    EXPECT(d.count_kmer({1, 2, 3, 5}) == 1);
    EXPECT(d.count_kmer({2, 3}) == 2);
    EXPECT(d.count_kmer({}) == 10);
    auto batch = d.count_kmers({{1, 2, 3, 5}, {2, 3}, {}, {5}});
    EXPECT(batch.size() == 4 && batch[0] == 1 && batch[1] == 2 && batch[2] == 10 && batch[3] == 1);
    // the reference asserts on symbols >= 6 (msbwt_core.rs:127)
    bool threw = false;
    try { d.count_kmer({1, 6}); } catch (const msbwt::Panic &) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { msbwt::RleBWT x; x.load_numpy_file("/nonexistent/file.npy"); } catch (const msbwt::IoError &) { threw = true; }
    EXPECT(threw);
    // bwt_converter.rs:246-256 (convert_to_vec KAT), then the same index through the packed-integer entry points
    EXPECT(msbwt::convert_to_vec("GTN$$ACCC$G") == rle);
    threw = false;
    try { msbwt::convert_to_vec("ACGX"); } catch (const msbwt::Panic &) { threw = true; }
    EXPECT(threw);
    // ACG = 0b000110, CC = 0b0101, T = 0b11 (first symbol most significant): counts in "CCGT$ N$ ACG$"
    const std::vector<uint64_t> km3 = {0b000110u};
    EXPECT(b.count_kmers_u64(km3, 3)[0] == b.count_kmer(convert_stoi("ACG")));
    EXPECT(b.count_kmers_u64({0b0101u}, 2)[0] == b.count_kmer(convert_stoi("CC")));
    EXPECT(b.count_kmers_u64_u32({0b11u, 0b00u}, 1)[0] == 1 && b.count_kmers_u64_u32({0b11u, 0b00u}, 1)[1] == 1);
    printf(failures ? "kat_driver: %d failure(s)\n" : "kat_driver: all passed (%d)\n", failures);
    return failures ? 1 : 0;
}
