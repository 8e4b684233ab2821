// Internal shared declarations for libb200_train (not part of the public C ABI; see include/b200_train.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_DRIVER (-3)
#define B200_ERR_UNSUPPORTED (-4)
#define B200_ERR_NCCL (-5)

// GEMM operand layouts (row-major storage; see gemm_tcgen05.cu header)
#define GEMM_NT 0  // C = A[M,K] * B[N,K]^T   (forward)
#define GEMM_NN 1  // C = A[M,K] * B[K,N]     (dgrad)
#define GEMM_TN 2  // C = A[K,M]^T * B[K,N]   (wgrad)

#define GEMM_FLAG_RESIDUAL 1          // C = acc + R
#define GEMM_FLAG_ROUND_BEFORE_ADD 2  // C = bf16(acc) + R   (matches the reference's two materialised bf16 ops)
#define GEMM_FLAG_SWIGLU 4            // NT only, B = [gate; up] rows (N = 2F): C = [gate | up] as usual AND R[M, F] = silu(gate) * up (R is an OUTPUT)

namespace b200 {

int set_error(int code, const char* fmt, ...);

int gemm_bf16_tcgen05(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, const void* R, int ldr,
                      int M, int N, int K, int flags, int group_m, int max_ctas, cudaStream_t stream);
int gemm_bf16_cublaslt(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);

#define B200_CHECK_LAUNCH(what)                                                                 \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return b200::set_error(B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace b200
