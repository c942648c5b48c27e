#pragma once
#include "ec.cuh"

namespace zke {
namespace dev {

struct MsmConfig {
    int c = 16;            // signed window width in bits
    uint32_t chunk = 256;  // max bucket entries accumulated by one thread
    uint32_t group = 16;   // buckets per running-sum thread
    bool classify = false; // true: drop zero scalars / infinite points and sum unit scalars outside the buckets
    uint32_t extra_passes = 0;  // further fan-in-32 reductions of the per-chunk partial sums (skewed buckets)
    // true: `points` is a table [n_windows][n] with level j holding 2^(c j) * P_i (fixed-base precomputation, legal
    // because the proving key is reused for every proof): all windows then share ONE bucket set, so the window
    // count no longer multiplies the bucket count, c can grow (fewer additions per point) and no Horner tail is left
    bool precomputed = false;
    // > 0: the first ba_levels levels of every chunk's addition tree use batched-affine additions (ba.cuh)
    int ba_levels = 0;
};
MsmConfig msm_config_witness();          // witness-scalar MSMs (mostly 0 / 1 / byte-sized scalars)
MsmConfig msm_config_full(uint32_t n, bool precomputed);   // full-width scalars (the H MSM)
inline int msm_windows(const MsmConfig& c) { return (255 + c.c - 1) / c.c; }

// Result block of one MSM (device or host memory), XYZZ points:
//   slots [0, MSM_ONES_SLOTS)                 partial sums of the unit-scalar points (infinity-padded)
//   slots [MSM_ONES_SLOTS, +n_windows)        per-window bucket sums S_j; the MSM value is
//                                             sum(ones) + sum_j 2^(c j) S_j  - the short serial tail (Horner over the
//                                             windows, ~255 doublings) is left to the host, which does it ~20x faster
//                                             than a single GPU thread.
static const int MSM_ONES_SLOTS = 32;
static const int MSM_MAX_WINDOWS = 32;
static const int MSM_RESULT_SLOTS = MSM_ONES_SLOTS + MSM_MAX_WINDOWS;

// Multi-scalar multiplication sum_i scalars[i] * points[i] over G1 (F = Fq) or G2 (F = Fq2).
//   points  : n affine points (Montgomery coordinates), (0,0) = infinity
//   scalars : n x 32-byte standard-form little-endian integers < r
//   result  : MSM_RESULT_SLOTS XYZZ points (device memory), layout above
template <class F>
struct MsmPlan {
    static size_t workspace_bytes(uint32_t n, const MsmConfig& cfg);
    // ev (optional): two events recorded immediately before / after the bucket-accumulation kernel (profiling)
    // heavy (optional): the bucket-accumulation kernel - the one kernel of an MSM that saturates the integer pipe - is
    // launched on heavy->st instead of `st` (the two streams are ordered by heavy->before / heavy->after), so that a
    // caller can keep the saturating kernels of all proofs on low-priority streams and let the latency-bound rest of
    // other proofs slip in between their blocks.
    struct Heavy { cudaStream_t st; cudaEvent_t before, after; };
    static void run(const uint8_t* points, const uint8_t* scalars, uint32_t n, const MsmConfig& cfg, uint8_t* workspace,
                    uint8_t* result, cudaStream_t st, cudaEvent_t* ev = nullptr, const Heavy* heavy = nullptr);
};

}  // namespace dev
}  // namespace zke
