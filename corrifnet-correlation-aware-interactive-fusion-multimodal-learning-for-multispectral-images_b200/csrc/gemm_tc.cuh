// Pieces shared by the tcgen05 TF32 GEMM kernels (gemm_tf32.cu: single-CTA tiles; gemm_tf32_pair.cu:
// CTA-pair tiles): tile constants, shared-memory / instruction descriptors, kernel arguments.
#pragma once
#include <stdlib.h>
#include "gemm.cuh"
#include "tc05.cuh"

namespace corrif {
namespace tc {
using namespace tc05;

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;              // tf32: 32 bytes of K per instruction
constexpr int ROW_BYTES = BK * 4;      // 128
constexpr int A_BYTES = BM * ROW_BYTES;
constexpr int NUM_THREADS = 192;

// Shared-memory matrix descriptor (sm_100 "version 1").
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   bits [46,48) = 1, bits [61,64) layout type.
//   K-major : SWIZZLE_128B (type 2): rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (type 1; 32-B chunks swizzled over
//             4-row atoms, TMA mode 128B_ATOM_32B): [k][32 floats] boxes, 4-k-row atoms 512 B apart
//             (SBO), the next 32 MN elements one box (BK*128 B) further (LBO).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t lbo = MN_MAJOR ? (uint64_t)(BK * ROW_BYTES) >> 4 : 1;
  constexpr uint64_t sbo = MN_MAJOR ? (512 >> 4) : (1024 >> 4);
  constexpr uint64_t layout = MN_MAJOR ? 1ull : 2ull;
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Instruction descriptor, kind::tf32, fp32 accumulate:
//   [4,6) c_format = 1 (F32), [7,10) a_format = 2 (TF32), [10,13) b_format = 2, bit 15 a_major,
//   bit 16 b_major (1 = MN-major), [17,23) N >> 3, [24,29) M >> 4.
template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct KernelArgs {
  EpiArgs epi;
  DropArgs drop;
  int K;
  int batch_inner, split_k;
  // TMA start coordinates per batch index: c0 is the contiguous dim of the operand in memory
  int64_t a_bo, a_bi, b_bo, b_bi, d_bo, d_bi;
  int64_t lda, ldb;
  int64_t bias_bo;           // bias element offset per outer batch index
};

}  // namespace tc

// CTA-pair (cta_group::2) kernel: true if it can run this problem, and its launcher
bool gemm_tf32_pair_supported(const corrif_gemm_desc& g);
int gemm_tf32_pair_launch(const corrif_gemm_desc& g, cudaStream_t stream);

}  // namespace corrif
