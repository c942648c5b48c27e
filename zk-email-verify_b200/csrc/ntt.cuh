#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zke {
namespace dev {

struct NttTables {
    const uint8_t* tw_fwd;   // [N/2] omega^k: shoup ? {w (standard form), floor(w 2^256 / r)} (64 B) : w in Montgomery form (32 B)
    const uint8_t* tw_inv;   // [N/2] omega^-k
    int log_n;
    bool shoup;              // table format: fixed-operand pairs (64-byte entries) or Montgomery form (32-byte entries)
};

// In-place inverse transform, natural order in, BIT-REVERSED order out, not scaled by 1/N.  If scale_bitrev is
// non-null, output position p is multiplied by scale_bitrev[p] (used to fuse the coset shift g^j / N); the scale table
// holds fixed-operand pairs like the twiddle tables ([N][2][32]).
void launch_intt_dif(uint8_t* data, const NttTables& T, const uint8_t* scale_bitrev, cudaStream_t st);
// In-place forward transform, bit-reversed order in, natural order out.
void launch_ntt_dit(uint8_t* data, const NttTables& T, cudaStream_t st);
// Pieces of a transform whose 2^log_g blocks of 2^(log_n - log_g) elements live on different GPUs:
// the block-local stages (twiddles of the full transform) and the cross-block stages on a range of columns.
void launch_intt_dif_block(uint8_t* block, const NttTables& T, int log_local, const uint8_t* scale_block, cudaStream_t st);
void launch_ntt_dit_block(uint8_t* block, const NttTables& T, int log_local, cudaStream_t st);
void launch_intt_cross(uint8_t* data, const NttTables& T, int log_g, uint32_t col0, uint32_t n_cols, cudaStream_t st);
void launch_ntt_cross(uint8_t* data, const NttTables& T, int log_g, uint32_t col0, uint32_t n_cols, cudaStream_t st);
void launch_quotient_cols(const uint8_t* a, const uint8_t* b, const uint8_t* c, uint8_t* d, int log_n, int log_g, uint32_t col0, uint32_t n_cols_log, cudaStream_t st);
void launch_hadamard(const uint8_t* a, const uint8_t* b, uint8_t* c, uint32_t n, cudaStream_t st);
void launch_quotient(const uint8_t* a, const uint8_t* b, const uint8_t* c, uint8_t* d, uint32_t n, cudaStream_t st);

}  // namespace dev
}  // namespace zke
