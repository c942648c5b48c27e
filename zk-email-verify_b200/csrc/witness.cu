// Batched witness generation on the GPU.
//
// Replaces the circom-generated WASM witness calculator that snarkjs' fullProve runs single-threaded
// (/root/reference/packages/helpers/src/chunked-zkey.ts:80-84 -> wtns.calculate).
//
// Mapping: ONE CTA PER EMAIL.  The witness program is levelised by the front-end (circuit.cpp): all ops of a level
// depend only on earlier levels, so the CTA's threads evaluate a level's ops in parallel and meet at a
// __syncthreads(); the long dependency chains of the circuit (40 chained SHA-256 compressions, 17 chained
// 2048-bit modular multiplications, the 1025-step regex automaton) cost one barrier per level instead of one
// kernel launch per level, and emails are independent so the batch fills the SMs (148 SMs -> size the batch in
// multiples of 148 x CTAs/SM for full occupancy).  Witness values live in HBM as 32-byte standard-form
// little-endian words, [email][signal]; the bounding resource is dependent-load latency, not bandwidth
// (32*m bytes written per email).
#include "device_engine.cuh"
#include "bigdiv.hpp"

namespace zke {
namespace dev {

__device__ __forceinline__ Fr lc_term(const DevProgram& P, const Fr& acc, const uint2& term, const Fr& x) {
    if (term.y == 0) return acc + x;
    if (term.y == 1) return acc - x;
    return acc + Fr::load(P.coef_r + 32ull * term.y) * x;   // (c*R) (x) -> c*x, standard form
}

// sum of coef * w over the LC's terms.  Terms are fetched four at a time: the four descriptor loads and then the four
// witness loads are independent of each other, so a thread has up to four gathers in flight instead of one dependent
// load after another (the level time of the witness program is the latency of its longest LC).
__device__ __forceinline__ Fr eval_lc(const DevProgram& P, const uint8_t* w, uint32_t id) {
    Fr acc = Fr::zero();
    uint32_t k = P.lc_ptr[id];
    const uint32_t end = P.lc_ptr[id + 1];
    for (; k + 4 <= end; k += 4) {
        const uint2 t0 = P.lc_terms[k], t1 = P.lc_terms[k + 1], t2 = P.lc_terms[k + 2], t3 = P.lc_terms[k + 3];
        const Fr x0 = Fr::load(w + 32ull * t0.x), x1 = Fr::load(w + 32ull * t1.x);
        const Fr x2 = Fr::load(w + 32ull * t2.x), x3 = Fr::load(w + 32ull * t3.x);
        acc = lc_term(P, acc, t0, x0);
        acc = lc_term(P, acc, t1, x1);
        acc = lc_term(P, acc, t2, x2);
        acc = lc_term(P, acc, t3, x3);
    }
    for (; k < end; ++k) {
        const uint2 t = P.lc_terms[k];
        acc = lc_term(P, acc, t, Fr::load(w + 32ull * t.x));
    }
    return acc;
}

__device__ __forceinline__ bool is_small(const Fr& x, uint32_t bound) {
    return (x.v[1] | x.v[2] | x.v[3] | x.v[4] | x.v[5] | x.v[6] | x.v[7]) == 0 && x.v[0] < bound;
}

__device__ Fr invz(const DevProgram& P, const Fr& x) {   // circomlib IsZero: inv <-- in != 0 ? 1/in : 0
    if (x.is_zero()) return x;
    if (is_small(x, P.n_small_inv)) return Fr::load(P.small_inv + 32ull * x.v[0]);
    Fr n = Fr::zero() - x;
    if (is_small(n, P.n_small_inv)) return Fr::zero() - Fr::load(P.small_inv + 32ull * n.v[0]);
    return x.to_mont().inv().from_mont();
}

__device__ __forceinline__ Fr shrand(const Fr& x, uint32_t shift, uint32_t nbits) {
    Fr o = Fr::zero();
    if (shift < 256) {
        const uint32_t ws = shift >> 5, bs = shift & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t lo = (j + ws < 8) ? x.v[(j + ws) & 7] : 0;
            uint32_t hi = (j + ws + 1 < 8) ? x.v[(j + ws + 1) & 7] : 0;
            o.v[j] = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
        }
    }
    if (nbits && nbits < 256) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (32u * j >= nbits) o.v[j] = 0;
            else if (32u * (j + 1) > nbits) o.v[j] &= (1u << (nbits - 32u * j)) - 1;
        }
    }
    return o;
}

__device__ void fpmul_hint_dev(const DevProgram& P, uint8_t* w, uint32_t aux_off, uint32_t dst) {
    const uint32_t* ax = P.aux + aux_off;
    const uint32_t n = ax[0], k = ax[1];
    uint32_t a[8 * 20], b[8 * 20], p[8 * 20], q[8 * 20], r[8 * 20];
    if (k > 20) return;
    for (uint32_t i = 0; i < k; ++i) {
        Fr x = Fr::load(w + 32ull * ax[2 + i]), y = Fr::load(w + 32ull * ax[2 + k + i]), z = Fr::load(w + 32ull * ax[2 + 2 * k + i]);
        for (int j = 0; j < 8; ++j) { a[8 * i + j] = x.v[j]; b[8 * i + j] = y.v[j]; p[8 * i + j] = z.v[j]; }
    }
    if (fpmul_hint_words(n, k, a, b, p, q, r) != 0) {
        for (uint32_t i = 0; i < 8 * k; ++i) { q[i] = 0; r[i] = 0; }
    }
    for (uint32_t i = 0; i < k; ++i) {
        Fr x, y;
        for (int j = 0; j < 8; ++j) { x.v[j] = q[8 * i + j]; y.v[j] = r[8 * i + j]; }
        x.store(w + 32ull * (dst + i));
        y.store(w + 32ull * (dst + k + i));
    }
}

__global__ void __launch_bounds__(WITNESS_THREADS)
witness_kernel(DevProgram P, uint8_t* __restrict__ w_all, size_t stride_elems, const uint8_t* __restrict__ inputs, uint32_t batch) {
    const uint32_t email = blockIdx.x;
    if (email >= batch) return;
    uint8_t* w = w_all + 32ull * stride_elems * email;
    const uint32_t tid = threadIdx.x;

    // level 0: constant one, inputs, zero the outputs (assigned by ops later)
    if (tid == 0) { Fr one = Fr::zero(); one.v[0] = 1; one.store(w); }
    const uint8_t* in = inputs + 32ull * P.n_inputs * email;
    for (uint32_t i = tid; i < P.n_inputs; i += blockDim.x) Fr::load(in + 32ull * i).store(w + 32ull * (1 + P.n_outputs + i));
    __syncthreads();

    for (uint32_t lvl = 0; lvl < P.n_levels; ++lvl) {
        const uint32_t beg = P.level_ptr[lvl], end = P.level_ptr[lvl + 1];
        for (uint32_t i = beg + tid; i < end; i += blockDim.x) {
            const uint4 op = P.ops[i];          // {dst, a, b, c | code << 28}
            const uint32_t code = op.w >> 28, c = op.w & 0x0FFFFFFFu;
            switch (code) {
                case 0:  // OP_LIN
                    eval_lc(P, w, op.y).store(w + 32ull * op.x);
                    break;
                case 1: {  // OP_QUAD: dst = A*B + C  (standard-form in/out: two Montgomery products)
                    Fr x = eval_lc(P, w, op.y), y = eval_lc(P, w, op.z), z = eval_lc(P, w, c);
                    ((x * y) * Fr::r2() + z).store(w + 32ull * op.x);
                    break;
                }
                case 2:  // OP_SHRAND
                    shrand(Fr::load(w + 32ull * op.y), op.z, c).store(w + 32ull * op.x);
                    break;
                case 3:  // OP_INVZ
                    invz(P, Fr::load(w + 32ull * op.y)).store(w + 32ull * op.x);
                    break;
                case 4:  // OP_FPMUL
                    fpmul_hint_dev(P, w, op.y, op.x);
                    break;
                default: break;
            }
        }
        __syncthreads();
    }
}

void launch_witness(const DevProgram& P, uint8_t* w_all, size_t stride_elems, const uint8_t* inputs, uint32_t batch, cudaStream_t st) {
    witness_kernel<<<batch, WITNESS_THREADS, 0, st>>>(P, w_all, stride_elems, inputs, batch);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
