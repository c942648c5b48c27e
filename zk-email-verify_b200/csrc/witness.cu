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
#include "lc_term.cuh"
#include "bigdiv.hpp"
#include <cstdlib>

namespace zke {
namespace dev {

// Evaluates up to three LCs whose terms are laid out back to back [A | B | C] at `t` (shared memory, or global for an
// oversized iteration).  The gathers of four consecutive terms are issued together - across the LC boundaries - so
// an 8-term `a*b + c` costs two dependent memory round trips instead of one per LC.
__device__ __forceinline__ void eval_lcs(const DevProgram& P, const uint8_t* w, const uint2* t, uint32_t nA, uint32_t nB,
                                         uint32_t nC, Fr& xa, Fr& xb, Fr& xc) {
    xa = Fr::zero(); xb = Fr::zero(); xc = Fr::zero();
    const uint32_t eA = nA, eB = nA + nB, total = nA + nB + nC;
    for (uint32_t j = 0; j < total; j += 4) {
        uint2 tt[4];
        Fr x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j + u < total) tt[u] = t[j + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j + u < total) x[u] = Fr::load(w + 32ull * tt[u].x);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j + u < total) {
                const TermVal tv = term_value(P.coef_r, tt[u], x[u]);
                const uint32_t idx = j + u;
                if (idx < eA) xa = tv.neg ? xa - tv.v : xa + tv.v;
                else if (idx < eB) xb = tv.neg ? xb - tv.v : xb + tv.v;
                else xc = tv.neg ? xc - tv.v : xc + tv.v;
            }
        }
    }
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stages the term block of one iteration into shared memory (16-byte chunks, coalesced); blocks larger than the
// buffer are not staged - the ops of such an iteration read their terms from global memory instead
__device__ __forceinline__ void stage_terms(const DevProgram& P, uint2* buf, const uint2& hdr) {
    if (hdr.y > WITNESS_TERM_BUF) return;
    const uint4* src = reinterpret_cast<const uint4*>(P.terms + hdr.x);
    uint4* dst = reinterpret_cast<uint4*>(buf);
    for (uint32_t i = threadIdx.x; i < hdr.y / 2; i += WITNESS_THREADS) cp_async16(dst + i, src + i);
}

// MINB = 2 caps the kernel at 64 registers per thread (half of an SM's register file per CTA instead of all of it), so
// that the proving kernels of the previous batch can share the SM with a witness CTA when batches are pipelined
// (zke_fullprove_submit): a witness CTA is latency-bound and leaves the multiplier pipe idle.
template <int MINB>
__global__ void __launch_bounds__(WITNESS_THREADS, MINB)
witness_kernel(DevProgram P, uint8_t* __restrict__ w_all, size_t stride_elems, const uint8_t* __restrict__ inputs, uint32_t batch) {
    extern __shared__ uint4 witness_smem[];
    uint2* const term_buf = reinterpret_cast<uint2*>(witness_smem);   // 2 x WITNESS_TERM_BUF
    const uint32_t email = blockIdx.x;
    if (email >= batch) return;
    uint8_t* w = w_all + 32ull * stride_elems * email;
    const uint32_t tid = threadIdx.x;

    // constant one and inputs
    if (tid == 0) { Fr one = Fr::zero(); one.v[0] = 1; one.store(w); }
    const uint8_t* in = inputs + 32ull * P.n_inputs * email;
    for (uint32_t i = tid; i < P.n_inputs; i += blockDim.x) {
        // inputs cross the ABI as raw 32-byte integers: reduce them mod r the way snarkjs' witness calculator does
        // (2^256 < 6 r, so at most five subtractions); everything downstream assumes canonical values
        Fr x = Fr::load(in + 32ull * i);
#pragma unroll 1
        for (int k = 0; k < 5; ++k) x.reduce_once();
        x.store(w + 32ull * (1 + P.n_outputs + i));
    }
    if (P.n_iters == 0) return;

    // software pipeline: op records one iteration ahead (registers), term blocks one iteration ahead (cp.async into
    // the other shared-memory buffer), iteration headers two ahead
    uint2 hdr = P.iter_hdr[0], hdr_next = P.iter_hdr[1];
    uint4 op = P.ops[tid];
    stage_terms(P, term_buf, hdr);
    cp_async_wait_all();
    __syncthreads();

    for (uint32_t k = 0; k < P.n_iters; ++k) {
        const uint2 hdr_next2 = P.iter_hdr[k + 2];      // the table has two sentinel entries
        uint4 op_next = make_uint4(0, WOP_NOP, 0, 0);
        if (k + 1 < P.n_iters) {
            op_next = P.ops[(size_t)(k + 1) * WITNESS_THREADS + tid];
            stage_terms(P, term_buf + ((k + 1) & 1) * WITNESS_TERM_BUF, hdr_next);
        }
        const uint32_t code = op.y & 0xffu;
        if (code <= 1 || code == 5) {   // OP_LIN: dst = A ; OP_QUAD: dst = A*B + C ; OP_SHRLC: dst = (A >> shift) & mask
            const uint32_t nA = (op.y >> 8) & 31u, nB = (op.y >> 13) & 31u, nC = (op.y >> 18) & 31u;
            const uint2* t = hdr.y > WITNESS_TERM_BUF ? P.terms + op.z : term_buf + (k & 1) * WITNESS_TERM_BUF + (op.z - hdr.x);
            Fr xa, xb, xc;
            eval_lcs(P, w, t, nA, nB, nC, xa, xb, xc);
            if (code == 1) {
                // bits and bytes (most of SHA-256 / the regex automaton): the product fits 64 bits, no reduction
                const bool tiny = ((xa.v[1] | xa.v[2] | xa.v[3] | xa.v[4] | xa.v[5] | xa.v[6] | xa.v[7] |
                                    xb.v[1] | xb.v[2] | xb.v[3] | xb.v[4] | xb.v[5] | xb.v[6] | xb.v[7]) == 0);
                if (tiny) {
                    const unsigned long long pr = (unsigned long long)xa.v[0] * xb.v[0];
                    Fr q = Fr::zero();
                    q.v[0] = (uint32_t)pr; q.v[1] = (uint32_t)(pr >> 32);
                    xa = q + xc;
                } else {
                    xa = (xa * xb) * Fr::r2() + xc;
                }
            }
            if (code == 5) xa = shrand(xa, op.w & 0xffffu, op.w >> 16);
            xa.store(w + 32ull * op.x);
        } else if (code == 2) {   // OP_SHRAND
            shrand(Fr::load(w + 32ull * op.z), op.w & 0xffffu, op.w >> 16).store(w + 32ull * op.x);
        } else if (code == 3) {   // OP_INVZ
            invz(P, Fr::load(w + 32ull * op.z)).store(w + 32ull * op.x);
        } else if (code == 4) {   // OP_FPMUL
            fpmul_hint_dev(P, w, op.z, op.x);
        }
        cp_async_wait_all();
        __syncthreads();     // level barrier and hand-over of the staged term block
        if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[k] = clock64();
        op = op_next; hdr = hdr_next; hdr_next = hdr_next2;
    }
}

static const size_t WITNESS_SMEM = 2 * (size_t)WITNESS_TERM_BUF * sizeof(uint2);   // 128 KB: above the 48 KB default

// The opt-in to > 48 KB of dynamic shared memory is a per-device (per-context) function attribute: the engine calls
// this from select_device() for every device it touches, and checks the result.
cudaError_t configure_witness_kernel() {
    cudaError_t e = cudaFuncSetAttribute(witness_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WITNESS_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(witness_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WITNESS_SMEM);
}

void launch_witness(const DevProgram& P, uint8_t* w_all, size_t stride_elems, const uint8_t* inputs, uint32_t batch, cudaStream_t st) {
    static const bool slim = getenv("ZKE_WITNESS_SLIM") && atoi(getenv("ZKE_WITNESS_SLIM")) != 0;
    if (slim) witness_kernel<2><<<batch, WITNESS_THREADS, WITNESS_SMEM, st>>>(P, w_all, stride_elems, inputs, batch);
    else witness_kernel<1><<<batch, WITNESS_THREADS, WITNESS_SMEM, st>>>(P, w_all, stride_elems, inputs, batch);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
