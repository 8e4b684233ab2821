// NVLink / NVSwitch peer-memory data path of the sharded step (sm_100a, one process per GPU).
//
// Replaces FSDP2's NCCL reduce_scatter_tensor / all_gather_into_tensor
// (/opt/.../torch/distributed/fsdp/_fully_shard/_fsdp_collectives.py:448-664, 237-291 as driven by
//  /root/reference/nemo_automodel/components/distributed/parallelizer.py:858-872) with:
//   * reduce-scatter = ONE pull kernel: every rank reads its 1/N slice of each peer's bf16 gradient buffer directly over
//     NVLink (peer-mapped pointers, 16-byte loads), accumulates in fp32 in rank order (single rounding, deterministic),
//     writes the bf16 shard in place and emits the shard's sum of squares for the global grad-norm in the same pass
//     (the reference runs a separate 291-tensor norm loop, components/training/utils.py:122-141);
//     wire bytes = 2 B/param*(N-1)/N, half of the reference's fp32 reduce-scatter;
//   * all-gather = copy-engine pushes (cudaMemcpyAsync peer-to-peer) of the updated shard into every peer's parameter buffer:
//     zero SMs, so the overlapped GEMMs keep the whole chip.
// Buffers that peers touch are cudaMalloc'ed here (one slab per rank) and exported with CUDA IPC.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

#define CUDA_TRY(expr, what)                                                                         \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) return set_error(B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

int mem_alloc(void** ptr, size_t bytes) {
  CUDA_TRY(cudaMalloc(ptr, bytes), "cudaMalloc");
  return 0;
}
int mem_free(void* ptr) {
  CUDA_TRY(cudaFree(ptr), "cudaFree");
  return 0;
}
int ipc_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  CUDA_TRY(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr), "cudaIpcGetMemHandle");
  return 0;
}
int ipc_import(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return 0;
}
int ipc_close(void* ptr) {
  CUDA_TRY(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
  return 0;
}
int copy_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st), "cudaMemcpyAsync(peer)");
  return 0;
}

struct PeerPtrs {
  const uint4* p[8];
};

__device__ __forceinline__ uint4 ld_peer(const uint4* p) {  // peer (NVLink) or local read-once data
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void acc8(float (&a)[8], const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    a[2 * i] += t.x;
    a[2 * i + 1] += t.y;
  }
}

// dst[i] = bf16( sum_j src_j[i] ) over nsrc sources (src 0 = this rank's own slice == dst), fp32 accumulate in fixed order.
// partial[blockIdx] = sum of squares of the rounded results.  UNROLL independent 16-byte loads per source keep NVLink busy.
template <int UNROLL>
__global__ void __launch_bounds__(512) reduce_slices_kernel(uint4* __restrict__ dst, PeerPtrs src, int nsrc, int64_t nvec,
                                                          float* __restrict__ partial) {
  __shared__ float s_w[16];
  float ss = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
    float acc[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
    // two sources per round: 2*UNROLL independent 16-byte NVLink loads in flight per thread before the first use
    for (int j = 0; j < nsrc; j += 2) {
      uint4 v0[UNROLL], v1[UNROLL];
      const bool two = j + 1 < nsrc;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        v0[u] = i < nvec ? ld_peer(src.p[j] + i) : make_uint4(0, 0, 0, 0);
        v1[u] = (two && i < nvec) ? ld_peer(src.p[j + 1] + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        acc8(acc[u], v0[u]);
        acc8(acc[u], v1[u]);   // zeros when the source count is odd: the sum order stays source order
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < nvec) {
        uint4 o;
        o.x = pack_bf16x2(acc[u][0], acc[u][1]);
        o.y = pack_bf16x2(acc[u][2], acc[u][3]);
        o.z = pack_bf16x2(acc[u][4], acc[u][5]);
        o.w = pack_bf16x2(acc[u][6], acc[u][7]);
        dst[i] = o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float r = bf16_round(acc[u][e]);
          ss += r * r;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_w[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, int accumulate) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s;
}

int reduce_scatter_pull_workspace_floats() { return 256; }

// srcs[0..nsrc): device pointers (own slice first, then the peers' views of the SAME slice), n bf16 elements (multiple of 8).
int reduce_scatter_pull_bf16(void* dst, const void* const* srcs, int nsrc, int64_t n, float* norm_sq, int accumulate_norm, float* ws,
                             int ctas, cudaStream_t st) {
  if (nsrc < 1 || nsrc > 8) return set_error(B200_ERR_ARG, "reduce_scatter_pull: 1..8 sources supported, got %d", nsrc);
  if (n % 8) return set_error(B200_ERR_ARG, "reduce_scatter_pull: n %% 8 != 0");
  PeerPtrs p;
  for (int j = 0; j < 8; ++j) p.p[j] = static_cast<const uint4*>(j < nsrc ? srcs[j] : nullptr);
  if (ctas <= 0) ctas = 64;
  if (ctas > 256) ctas = 256;
  reduce_slices_kernel<4><<<ctas, 512, 0, st>>>(static_cast<uint4*>(dst), p, nsrc, n / 8, ws);
  B200_CHECK_LAUNCH("reduce_slices");
  if (norm_sq) {
    sum_partials_kernel<<<1, 32, 0, st>>>(ws, ctas, norm_sq, accumulate_norm);
    B200_CHECK_LAUNCH("sum_partials");
  }
  return 0;
}

}  // namespace b200
