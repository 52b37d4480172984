// Internal host-callable operations of libvaecap (one declaration per kernel family).
#pragma once
#include "host_util.h"

namespace vc {

// gemm_api.cu ------------------------------------------------------------------------------
int gemm_store(cudaStream_t stream, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M,
               int N, int K, const EpiStore& epi, int bn, int splits);

}  // namespace vc
