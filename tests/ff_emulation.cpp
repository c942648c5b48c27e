#define ZKE_FF_EMULATE
#include "ff.cuh"
#include <cstring>
using namespace zke::dev;
extern "C" {
void ff_set_consts(const uint32_t* mod, const uint32_t* r, const uint32_t* r2, uint32_t inv) {
    memcpy(FR_C.mod, mod, 32); memcpy(FR_C.r, r, 32); memcpy(FR_C.r2, r2, 32); FR_C.inv = inv;
    uint64_t borrow = 0;   // nmod = 2^256 - mod
    for (int i = 0; i < 8; ++i) { uint64_t d = (uint64_t)0 - mod[i] - borrow; FR_C.nmod[i] = (uint32_t)d; borrow = (d >> 32) & 1; }
}
// fixed-operand product: out = a * w mod p for w given as (standard form, floor(w 2^256 / p))
void ff_shoup(const uint32_t* a, const uint32_t* w, const uint32_t* wq, uint32_t* out, int n) {
    for (int i = 0; i < n; ++i) {
        Fr x; memcpy(x.v, a + 8 * i, 32);
        Fr z = Fr::mul_shoup(x, w + 8 * i, wq + 8 * i);
        memcpy(out + 8 * i, z.v, 32);
    }
}
static void run(int which, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fr x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    switch (which) {
        case 0: z = Fr::mul_cios(x, y); break;
        case 1: z = Fr::mul_sos(x, y); break;
        case 2: z = Fr::mul_sos_plain(x, y); break;
        case 3: { uint32_t T[16]; Fr::sqr_wide(T, x.v); z = Fr::redc_wide(T); break; }
        case 4: z = x + y; break;
        case 5: z = x - y; break;
        case 8: z = Fr::sqr_cios(x); break;
        case 6: z = Fr::mul_add2(x, y, y, x + y); break;          // x*y + y*(x+y)
        case 7: z = Fr::mul_sub2(x, y, y, x + y); break;          // x*y - y*(x+y)
        default: z = x;
    }
    memcpy(out, z.v, 32);
}
void ff_op(int which, const uint32_t* a, const uint32_t* b, uint32_t* out, int n) { for (int i = 0; i < n; ++i) run(which, a + 8 * i, b + 8 * i, out + 8 * i); }
void ff_wide(int which, const uint32_t* a, const uint32_t* b, uint32_t* out16, int n) {
    for (int i = 0; i < n; ++i) {
        if (which == 0) Fr::mul_wide<8>(out16 + 16 * i, a + 8 * i, b + 8 * i);
        else if (which == 1) Fr::mul_wide_karatsuba(out16 + 16 * i, a + 8 * i, b + 8 * i);
        else Fr::sqr_wide(out16 + 16 * i, a + 8 * i);
    }
}
}
