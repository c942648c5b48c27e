// Internal definitions behind the opaque C-ABI handles (include/zkemail_b200.h).
#pragma once
#include "circuit.hpp"
#include <string>

namespace zke {
void set_err(char* err, size_t cap, const std::string& msg);
void random_scalar(U256& out);   // uniform in [0, r), from /dev/urandom
}

struct zke_circuit {
    zke::Circuit c;
};

// window widths (bits) of the signed-digit Pippenger: witness-scalar MSMs (mostly tiny scalars) and the H MSM
#define ZKE_MSM_C_WITNESS 12
#define ZKE_MSM_C_H 16
// per-email result block on the device: A, B1, C, H (G1 XYZZ, 128 bytes each) then B2 (G2 XYZZ, 256 bytes)
#define ZKE_RESULT_STRIDE 768
