// Sparse matrix-vector products of the Groth16 prover ("buildABC" in snarkjs' groth16_prove)
// plus the circom `===` check: one thread per R1CS row evaluates <A_i,w>, <B_i,w>, <C_i,w>, verifies
// <A_i,w><B_i,w> = <C_i,w>, and writes a_i, b_i in Montgomery form for the NTTs.
// HBM-bound gather: nnz * 8 B of (var, coef) pairs + one 32-byte witness read per non-zero, 2 * 32 * N written.
#include "device_engine.cuh"
#include "lc_term.cuh"
#include <algorithm>

namespace zke {
namespace dev {

__device__ __forceinline__ Fr row_dot(const uint32_t* ptr, const uint2* terms, const uint8_t* coef_r, const uint8_t* w, uint32_t row) {
    Fr acc = Fr::zero();
    const uint32_t beg = ptr[row], end = ptr[row + 1];
    for (uint32_t k = beg; k < end; ++k) {
        const uint2 t = terms[k];
        const TermVal tv = term_value(coef_r, t, Fr::load(w + 32ull * t.x));   // +-1 / +-2^k coefficients: no product
        acc = tv.neg ? acc - tv.v : acc + tv.v;
    }
    return acc;
}

// c_out (optional): c_i = a_i b_i in Montgomery form (the third vector of the quotient computation; snarkjs derives it
// from a and b rather than from the C matrix) - fused here so that no separate Hadamard pass is needed.
__global__ void __launch_bounds__(256)
build_ab_kernel(DevR1cs R, const uint8_t* __restrict__ w, uint8_t* __restrict__ a_out, uint8_t* __restrict__ b_out,
                uint8_t* __restrict__ c_out, uint32_t n, uint32_t* first_bad, RowMap map) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    // sharded proving: thread t works on row (block, column) of the column range this GPU holds
    const uint32_t i = map.log_cols < 0 ? t : (((t >> map.log_cols) << map.log_m) + map.col0 + (t & ((1u << map.log_cols) - 1)));
    Fr a = Fr::zero(), b = Fr::zero(), c = Fr::zero();
    if (i < R.n_constraints) {
        a = row_dot(R.a_ptr, R.a_terms, R.coef_r, w, i);
        b = row_dot(R.b_ptr, R.b_terms, R.coef_r, w, i);
        a = a.to_mont();
        if (R.c_ptr) {   // contexts opened from a `.zkey` alone have no C matrix (snarkjs does not check either)
            const Fr c_lc = row_dot(R.c_ptr, R.c_terms, R.coef_r, w, i);
            if ((a * b) != c_lc) atomicMin(first_bad, i);   // (aR) (x) b = a b in standard form
        }
        b = b.to_mont();
        if (c_out) c = a * b;
    } else if (i <= R.n_constraints + R.n_public) {
        // extra rows that make the public-input polynomials independent: a = w_j, b = 0
        a = Fr::load(w + 32ull * (i - R.n_constraints)).to_mont();
    }
    a.store(a_out + 32ull * i);
    b.store(b_out + 32ull * i);
    if (c_out) c.store(c_out + 32ull * i);
}

// The circom `===` check alone, for a whole batch in one launch (calculateWitness + checkConstraints without proving:
// zke_witness): no a / b vectors are written - build_ab_kernel stores 2 x 32 x N bytes per email that only the
// transforms need - and blockIdx.y walks the emails.  first_bad[e] = smallest violated row of email e.
// Measured on B200, batch 64, default circuit: 40 ms against 61 ms for 64 build_ab launches (configs[1]: 897 -> 1,270
// emails/s); a variant in which a thread checks its row for four emails at once (terms read once, four gathers in
// flight) needs 142 registers and is slower (75 ms).
__global__ void __launch_bounds__(256)
check_rows_kernel(DevR1cs R, const uint8_t* __restrict__ w_all, size_t stride_elems, uint32_t* first_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n_constraints) return;
    const uint8_t* w = w_all + 32ull * stride_elems * blockIdx.y;
    const Fr a = row_dot(R.a_ptr, R.a_terms, R.coef_r, w, i);
    const Fr b = row_dot(R.b_ptr, R.b_terms, R.coef_r, w, i);
    const Fr c = row_dot(R.c_ptr, R.c_terms, R.coef_r, w, i);
    if ((a.to_mont() * b) != c) atomicMin(first_bad + blockIdx.y, i);   // (aR) (x) b = a b in standard form
}
void launch_check_rows(const DevR1cs& R, const uint8_t* w_all, size_t stride_elems, uint32_t batch, uint32_t* first_bad, cudaStream_t st) {
    if (!R.c_ptr || batch == 0) return;
    const uint32_t MAXY = 32768;
    for (uint32_t e0 = 0; e0 < batch; e0 += MAXY) {
        const dim3 grid((R.n_constraints + 255) / 256, std::min(MAXY, batch - e0), 1);
        check_rows_kernel<<<grid, 256, 0, st>>>(R, w_all + 32ull * stride_elems * e0, stride_elems, first_bad + e0);
        ZKE_COUNT_LAUNCH(1);
    }
}

// Validation of externally supplied witnesses (zke_load_witness): every value canonical (< r) and w[0] == 1.
__global__ void check_witness_kernel(const uint8_t* __restrict__ w_all, size_t stride_elems, uint32_t n_vars, uint32_t batch, uint32_t* flag) {
    const size_t total = (size_t)n_vars * batch;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t e = (uint32_t)(t / n_vars), i = (uint32_t)(t % n_vars);
        Fr x = Fr::load(w_all + 32ull * (stride_elems * e + i));
        Fr y = x;
        y.reduce_once();
        if (y != x) flag[0] = 1;
        if (i == 0) { Fr one = Fr::zero(); one.v[0] = 1; if (x != one) flag[1] = 1; }
    }
}
void launch_check_witness(const uint8_t* w_all, size_t stride_elems, uint32_t n_vars, uint32_t batch, uint32_t* flag, cudaStream_t st) {
    check_witness_kernel<<<148 * 8, 256, 0, st>>>(w_all, stride_elems, n_vars, batch, flag);
    ZKE_COUNT_LAUNCH(1);
}

void launch_build_ab(const DevR1cs& R, const uint8_t* w, uint8_t* a_out, uint8_t* b_out, uint8_t* c_out, uint32_t n, uint32_t* first_bad, cudaStream_t st,
                     const RowMap* map) {
    RowMap m;
    if (map) m = *map;
    build_ab_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, w, a_out, b_out, c_out, n, first_bad, m);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
