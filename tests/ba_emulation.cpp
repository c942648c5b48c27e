// CPU harness for zk-email-verify_b200/csrc/ba.cuh (see tests/test_ba_emulation.py): one "thread" reduces a list of
// sign-tagged table entries with LEVELS batched-affine tree levels followed by XYZZ accumulation, exactly as
// ba_chunk_sum_kernel does per thread, with the block-wide inversion replaced by a direct one.
#define ZKE_FF_EMULATE
#include "ba.cuh"
#include <cstring>
#include <vector>
using namespace zke::dev;
extern "C" {
void ba_set_consts(const uint32_t* mod, const uint32_t* r, const uint32_t* r2, uint32_t inv) {
    memcpy(FQ_C.mod, mod, 32); memcpy(FQ_C.r, r, 32); memcpy(FQ_C.r2, r2, 32); FQ_C.inv = inv;
}
// points: n_points affine Montgomery (64 B each); entries: cnt words; out: XYZZ (128 B)
void ba_reduce(const uint8_t* points, const uint32_t* entries, int cnt, int levels, uint8_t* out) {
    // two "chunks" per thread (the list is cut in two) share one running product and one inversion per level,
    // with a stride of 3 elements in the scratch arrays, as the kernel's [row][chunk] layout has
    const size_t S = 3;
    const int cnts[2] = {cnt / 2, cnt - cnt / 2};
    const uint32_t* ent[2] = {entries, entries + cnts[0]};
    std::vector<Fq> pref[2];
    std::vector<Affine<Fq>> bufa[2], bufb[2];
    for (int c = 0; c < 2; ++c) { pref[c].resize(S * (cnt / 2 + 2)); bufa[c].resize(S * (cnt + 2)); bufb[c].resize(S * (cnt + 2)); }
    int n[2] = {cnts[0], cnts[1]};
    for (int lev = 0; lev < levels; ++lev) {
        Fq run = Fq::one(), before[2];
        for (int c = 0; c < 2; ++c) {
            before[c] = run;
            const int np = n[c] / 2;
            if (lev == 0) ba_phase_a(BaTableSource<Fq>{points, ent[c]}, np, pref[c].data(), S, run);
            else ba_phase_a(BaBufferSource<Fq>{(lev & 1) ? bufa[c].data() : bufb[c].data(), S}, np, pref[c].data(), S, run);
        }
        Fq v = run.inv();
        for (int c = 1; c >= 0; --c) {
            const int np = n[c] / 2;
            Affine<Fq>* dst = (lev & 1) ? bufb[c].data() : bufa[c].data();
            if (lev == 0) {
                BaTableSource<Fq> src{points, ent[c]};
                ba_phase_b(src, np, pref[c].data(), S, before[c], v, dst, S);
                if (n[c] & 1) src.get(n[c] - 1).store(dst + S * np);
            } else {
                BaBufferSource<Fq> src{(lev & 1) ? bufa[c].data() : bufb[c].data(), S};
                ba_phase_b(src, np, pref[c].data(), S, before[c], v, dst, S);
                if (n[c] & 1) src.get(n[c] - 1).store(dst + S * np);
            }
            n[c] = np + (n[c] & 1);
        }
    }
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < n[c]; ++i) {
            Affine<Fq> p = levels == 0 ? BaTableSource<Fq>{points, ent[c]}.get(i)
                                       : BaBufferSource<Fq>{(levels & 1) ? bufa[c].data() : bufb[c].data(), S}.get(i);
            acc.madd(p, false);
        }
    memcpy(out, &acc, sizeof(acc));
}
}
