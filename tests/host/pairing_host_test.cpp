// Host-side check of csrc/pairing.cuh (g++, PTX carry primitives emulated) against vectors the Python oracle wrote:
// file = u64 count, then per case P (8 u64), Q (16 u64), expected e(P, Q) as 6 x Fq2 tower coefficients (48 u64),
// all Montgomery limbs.  Prints ALL OK on success.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../../distributed_groth16_b200/csrc/pairing.cuh"

using namespace b200zk;

int main(int argc, char** argv) {
    if (argc < 2) { printf("usage: pairing_host_test vectors.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { printf("cannot open %s\n", argv[1]); return 2; }
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) return 2;
    int fails = 0;
    Fq12 prod = Fq12::one();
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t buf[8 + 16 + 48];
        if (fread(buf, 8, 72, f) != 72) { printf("short read\n"); return 2; }
        affine_t<Fq> P; affine_t<Fq2> Q; Fq12 exp;
        memcpy(&P, buf, 64); memcpy(&Q, buf + 8, 128); memcpy(&exp, buf + 24, 384);
        static_assert(sizeof(Fq12) == 384, "Fq12 layout");
        Fq12 ml = miller_loop(P, Q);
        Fq12 got = final_exponentiation(ml);
        if (!(got == exp)) { ++fails; printf("case %llu: pairing mismatch\n", (unsigned long long)i); }
        prod = Fq12::mul(prod, ml);
        // inverse / Frobenius sanity on the Miller value
        if (!(Fq12::mul(ml, Fq12::inv(ml)) == Fq12::one())) { ++fails; printf("case %llu: inv\n", (unsigned long long)i); }
    }
    // the file's cases are built so that the product of all pairings is 1 (e(aP,Q) e(-P,aQ) ...): one shared final exp
    uint64_t want_prod_one = 0;
    if (fread(&want_prod_one, 8, 1, f) == 1 && want_prod_one) {
        if (!(final_exponentiation(prod) == Fq12::one())) { ++fails; printf("product of pairings != 1\n"); }
    }
    fclose(f);
    printf(fails ? "FAILED %d\n" : "ALL OK %d\n", fails ? fails : (int)n);
    return fails ? 1 : 0;
}
