// Device-side data structures and kernel launchers shared by the .cu files.
#pragma once
#include "ec.cuh"
#include <cstddef>
#include <cstdint>

namespace zke {
namespace dev {

static const int WITNESS_THREADS = 256;

// number of kernels launched by this library since load (reported by bench.py as gpu_launches)
extern unsigned long long g_kernel_launches;
#define ZKE_COUNT_LAUNCH(n) (::zke::dev::g_kernel_launches += (n))

// Witness program resident in HBM (built once per circuit).
struct DevProgram {
    const uint4* ops;            // {dst, a, b, c | code << 28}, sorted by level
    const uint32_t* level_ptr;   // n_levels + 1
    const uint32_t* lc_ptr;      // LC pool CSR
    const uint2* lc_terms;       // {var, coef index}
    const uint32_t* aux;
    const uint8_t* coef_r;       // [n_coefs][32]: coefficient * R mod r  (Montgomery-scaled: (cR) (x) w = c*w)
    const uint8_t* small_inv;    // [n_small_inv][32]: x^-1 mod r in standard form, entry 0 unused
    uint32_t n_small_inv;
    uint32_t n_levels, n_ops, n_vars, n_temps, n_outputs, n_inputs;
};

// R1CS matrices resident in HBM.
struct DevR1cs {
    const uint32_t *a_ptr, *b_ptr, *c_ptr;
    const uint2 *a_terms, *b_terms, *c_terms;   // {var, coef index}
    const uint8_t* coef_r;                      // same table as DevProgram::coef_r
    uint32_t n_constraints, n_public, n_vars;
};

void upload_field_constants();

void launch_witness(const DevProgram& P, uint8_t* w_all, size_t stride_elems, const uint8_t* inputs, uint32_t batch, cudaStream_t st);

// a[i] = <A_i, w>, b[i] = <B_i, w> in Montgomery form for i < n_constraints, the n_public + 1 extra rows of the
// Groth16 QAP (a = w_j, b = 0), zero padding up to n; also checks <A,w><B,w> = <C,w> and atomically records the
// smallest violated row in *first_bad (initialised to 0xffffffff by the caller).
void launch_build_ab(const DevR1cs& R, const uint8_t* w, uint8_t* a_out, uint8_t* b_out, uint32_t n, uint32_t* first_bad, cudaStream_t st);

}  // namespace dev
}  // namespace zke
