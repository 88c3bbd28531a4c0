// Host-side check of csrc/codec.cuh (g++, PTX carry primitives emulated) against vectors written by the Python oracle.
// File: u64 n1, then n1 x { 32 B encoding, 8 u64 expected affine (Montgomery), u64 valid };
//       u64 n2, then n2 x { 64 B encoding, 16 u64 expected affine, u64 valid, u64 check_subgroup }.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include "../../distributed_groth16_b200/csrc/codec.cuh"

using namespace b200zk;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int fails = 0, total = 0;
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) return 2;
    for (uint64_t i = 0; i < n; ++i, ++total) {
        uint8_t enc[32]; uint64_t exp[8], valid;
        if (fread(enc, 1, 32, f) != 32 || fread(exp, 8, 8, f) != 8 || fread(&valid, 8, 1, f) != 1) return 2;
        affine_t<Fq> p;
        bool ok = g1_decode(enc, &p);
        if (ok != (valid != 0)) { ++fails; printf("g1 %llu: validity %d, want %llu\n", (unsigned long long)i, ok, (unsigned long long)valid); continue; }
        if (!ok) { if (!p.is_inf()) { ++fails; printf("g1 %llu: invalid slot not infinity\n", (unsigned long long)i); } continue; }
        if (memcmp(&p, exp, 64)) { ++fails; printf("g1 %llu: point mismatch\n", (unsigned long long)i); }
        uint8_t back[32];
        g1_encode(p, back);
        if (memcmp(back, enc, 32)) { ++fails; printf("g1 %llu: re-encoding mismatch\n", (unsigned long long)i); }
    }
    if (fread(&n, 8, 1, f) != 1) return 2;
    for (uint64_t i = 0; i < n; ++i, ++total) {
        uint8_t enc[64]; uint64_t exp[16], valid, sub;
        if (fread(enc, 1, 64, f) != 64 || fread(exp, 8, 16, f) != 16 || fread(&valid, 8, 1, f) != 1 || fread(&sub, 8, 1, f) != 1) return 2;
        affine_t<Fq2> p;
        bool ok = g2_decode(enc, sub != 0, &p);
        if (ok != (valid != 0)) { ++fails; printf("g2 %llu: validity %d, want %llu\n", (unsigned long long)i, ok, (unsigned long long)valid); continue; }
        if (!ok) continue;
        if (memcmp(&p, exp, 128)) { ++fails; printf("g2 %llu: point mismatch\n", (unsigned long long)i); }
        uint8_t back[64];
        g2_encode(p, back);
        if (memcmp(back, enc, 64)) { ++fails; printf("g2 %llu: re-encoding mismatch\n", (unsigned long long)i); }
    }
    fclose(f);
    printf(fails ? "FAILED %d of %d\n" : "ALL OK %d\n", fails ? fails : total, total);
    return fails ? 1 : 0;
}
