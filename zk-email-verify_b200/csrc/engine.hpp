// Internal definitions behind the opaque C-ABI handles (include/zkemail_b200.h).
#pragma once
#include "circuit.hpp"
#include <string>

namespace zke {
void set_err(char* err, size_t cap, const std::string& msg);
}

struct zke_circuit {
    zke::Circuit c;
};
