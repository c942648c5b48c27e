// Sparse matrix-vector products of the Groth16 prover ("buildABC" in snarkjs' groth16_prove)
// plus the circom `===` check: one thread per R1CS row evaluates <A_i,w>, <B_i,w>, <C_i,w>, verifies
// <A_i,w><B_i,w> = <C_i,w>, and writes a_i, b_i in Montgomery form for the NTTs.
// HBM-bound gather: nnz * 8 B of (var, coef) pairs + one 32-byte witness read per non-zero, 2 * 32 * N written.
#include "device_engine.cuh"

namespace zke {
namespace dev {

__device__ __forceinline__ Fr row_dot(const uint32_t* ptr, const uint2* terms, const uint8_t* coef_r, const uint8_t* w, uint32_t row) {
    Fr acc = Fr::zero();
    const uint32_t beg = ptr[row], end = ptr[row + 1];
    for (uint32_t k = beg; k < end; ++k) {
        const uint2 t = terms[k];
        Fr x = Fr::load(w + 32ull * t.x);
        if (t.y == 0) acc = acc + x;
        else if (t.y == 1) acc = acc - x;
        else acc = acc + Fr::load(coef_r + 32ull * t.y) * x;
    }
    return acc;
}

__global__ void __launch_bounds__(256)
build_ab_kernel(DevR1cs R, const uint8_t* __restrict__ w, uint8_t* __restrict__ a_out, uint8_t* __restrict__ b_out,
                uint32_t n, uint32_t* first_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr a = Fr::zero(), b = Fr::zero();
    if (i < R.n_constraints) {
        a = row_dot(R.a_ptr, R.a_terms, R.coef_r, w, i);
        b = row_dot(R.b_ptr, R.b_terms, R.coef_r, w, i);
        Fr c = row_dot(R.c_ptr, R.c_terms, R.coef_r, w, i);
        Fr ab = (a * b) * Fr::r2();   // standard form product
        if (ab != c) atomicMin(first_bad, i);
        a = a.to_mont();
        b = b.to_mont();
    } else if (i <= R.n_constraints + R.n_public) {
        // extra rows that make the public-input polynomials independent: a = w_j, b = 0
        a = Fr::load(w + 32ull * (i - R.n_constraints)).to_mont();
    }
    a.store(a_out + 32ull * i);
    b.store(b_out + 32ull * i);
}

void launch_build_ab(const DevR1cs& R, const uint8_t* w, uint8_t* a_out, uint8_t* b_out, uint32_t n, uint32_t* first_bad, cudaStream_t st) {
    build_ab_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, w, a_out, b_out, n, first_bad);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
