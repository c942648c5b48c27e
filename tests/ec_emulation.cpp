// CPU harness for zk-email-verify_b200/csrc/ec.cuh under ZKE_FF_EMULATE (tests/test_ec_emulation.py): the XYZZ
// formulas the bucket kernels run - mixed addition, full addition, doubling - on arbitrary operand pairs.
#define ZKE_FF_EMULATE
#include "ec.cuh"
#include <cstring>
using namespace zke::dev;
extern "C" {
void ec_set_consts(const uint32_t* mod, const uint32_t* r, const uint32_t* r2, uint32_t inv) {
    memcpy(FQ_C.mod, mod, 32); memcpy(FQ_C.r, r, 32); memcpy(FQ_C.r2, r2, 32); FQ_C.inv = inv;
}
// acc (XYZZ, 128 B, in/out) op= point (affine 64 B or XYZZ 128 B): which = 0 madd(+), 1 madd(-), 2 add, 3 dbl
void ec_op(int which, uint8_t* acc_bytes, const uint8_t* operand) {
    XYZZ<Fq> acc;
    memcpy(&acc, acc_bytes, sizeof(acc));
    if (which == 0 || which == 1) { Affine<Fq> p; memcpy(&p, operand, sizeof(p)); acc.madd(p, which == 1); }
    else if (which == 2) { XYZZ<Fq> o; memcpy(&o, operand, sizeof(o)); acc.add(o); }
    else acc.dbl();
    memcpy(acc_bytes, &acc, sizeof(acc));
}
}
