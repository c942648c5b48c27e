// Host-only harness for tests/test_ff_tc.py: prints the tensor-core reduction table tc_build_table() (ff_tc.cuh) makes
// for a modulus given as 8 little-endian hex limbs on the command line.  No device code runs.
#include <cstdio>
#include <cstdlib>
#include "ff_tc.cuh"
int main(int argc, char** argv) {
    if (argc != 9) return 2;
    uint32_t mod[8];
    for (int i = 0; i < 8; ++i) mod[i] = (uint32_t)strtoul(argv[1 + i], nullptr, 16);
    zke::dev::TcTable t;
    zke::dev::tc_build_table(mod, &t);
    printf("%08x\n", t.mu);
    for (int lane = 0; lane < 32; ++lane) {
        for (int k = 0; k < 8; ++k) printf("%08x ", t.bfrag[lane][k]);
        printf("\n");
    }
    return 0;
}
