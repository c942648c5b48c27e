#pragma once
#include "ec.cuh"

namespace zke {
namespace dev {

// Multi-scalar multiplication sum_i scalars[i] * points[i] over G1 (F = Fq) or G2 (F = Fq2).
//   points  : n affine points (Montgomery coordinates), (0,0) = infinity
//   scalars : n x 32-byte standard-form little-endian integers < r
//   c       : window width in bits (signed digits)
//   classify: true -> zero scalars / infinite points are skipped and unit scalars are summed outside the bucket
//             machinery (witness MSMs); false -> every scalar goes through Pippenger (the H MSM)
//   result  : one XYZZ point (device memory)
template <class F>
struct MsmPlan {
    static size_t workspace_bytes(uint32_t n, int c);
    // ev (optional): two events recorded immediately before / after the bucket-accumulation kernel (profiling)
    static void run(const uint8_t* points, const uint8_t* scalars, uint32_t n, int c, bool classify, uint8_t* workspace,
                    uint8_t* result, cudaStream_t st, cudaEvent_t* ev = nullptr);
};

}  // namespace dev
}  // namespace zke
