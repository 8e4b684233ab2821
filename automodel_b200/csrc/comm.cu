// NVLink 5 / NVSwitch data path of the sharded step (sm_100a, one process per GPU): the per-unit gradient reduce-scatter and parameter
// all-gather as OUR kernels on symmetric (peer-mapped + NVLS multicast-mapped) memory, behind the b200_ctx of include/b200_train.h.
//
// Replaces FSDP2's NCCL reduce_scatter_tensor / all_gather_into_tensor
// (torch/distributed/fsdp/_fully_shard/_fsdp_collectives.py:448-664, 237-291 as driven by
//  /root/reference/nemo_automodel/components/distributed/parallelizer.py:858-872 and the MixedPrecisionPolicy of
//  components/distributed/config.py:121-132: bf16 parameters, fp32 gradient reduction):
//   * reduce-scatter: every rank reduces ITS 1/N slice of the unit's flat bf16 gradient buffer across all ranks
//       - NVLS:  multimem.ld_reduce .add .acc::f32 .bf16x2 on the multicast mapping - the NVSwitch sums the N copies with fp32
//                accumulation and returns one rounded bf16 value (one rounding, as the reference's fp32 reduce followed by the cast to
//                the bf16 gradient); 16 B per request, the SMs only move the reduced 1/N;
//       - P2P:   16-byte loads from every peer's mapping, fp32 accumulation in rank order, one rounding (no multicast support);
//     the result is written in place into the rank's own slice; wire bytes 2 B/param*(N-1)/N = half of the reference's fp32 reduce-scatter;
//   * all-gather: NVLS multimem.st of the rank's updated slice (the switch replicates it into every GPU's buffer), or P2P pulls;
//   * cross-rank ordering lives INSIDE the kernels: CTA b of every rank meets CTA b of every other rank on a symmetric signal pad
//     (release/acquire CAS flags over NVLink) before the first remote access and after the last one - no NCCL call, no host sync.
// Memory comes from torch's symmetric-memory allocator (plumbing: cuMemCreate + fabric/fd exchange + cuMulticastBindMem); this file
// only sees raw pointers.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <new>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kMaxWorld = 8;
constexpr int kMaxBuffers = 4;
constexpr int kMaxCommCtas = 64;        // barrier channels per kernel family
constexpr int kPadWordsPerFamily = kMaxCommCtas * kMaxWorld;   // uint32 flags [cta][peer]

struct SymBuffer {
  char* peer[kMaxWorld];   // this rank's mapping of every rank's buffer (peer[rank] = the local allocation)
  char* mc;                // NVLS multicast mapping of the same buffer (nullptr: not supported)
  size_t bytes;
};

constexpr int kPadFamilies = 3;            // reduce-scatter, all-gather, scalar all-reduce
constexpr int kMaxScalars = 16;
constexpr size_t kFlagBytes = kPadFamilies * kPadWordsPerFamily * sizeof(uint32_t);
constexpr size_t kPadBytes = kFlagBytes + 2 * kMaxScalars * kMaxWorld * sizeof(float);   // + exchange area [parity][slot][rank]

struct CommCtx {
  int rank, world;
  unsigned scalar_calls;
  SymBuffer buf[kMaxBuffers];
  uint32_t* pad[kMaxWorld];   // symmetric signal pad: kPadFamilies x kPadWordsPerFamily flag words + the scalar exchange area, zero-initialised
  size_t pad_bytes;
  long long timeout_ns;
};

struct Peers {
  char* data[kMaxWorld];
  uint32_t* pad[kMaxWorld];
  char* mc;
  int rank, world;
  long long timeout_ns;
};

// ---------------------------------------------------------------- device-side cross-rank barrier
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// CTA `chan` of this rank meets CTA `chan` of every peer.  Flags reset themselves (0 -> 1 by the signaller, 1 -> 0 by the waiter), so
// the same channel can be reused by back-to-back barriers without a sequence number.  A peer that never arrives (crashed rank) traps
// after timeout_ns instead of hanging the GPU.
__device__ __forceinline__ void cross_rank_barrier(const Peers& c, int family, int chan) {
  __syncthreads();   // every thread's earlier accesses are ordered before the release below (bar.sync is cumulative in the PTX model)
  const int p = threadIdx.x;
  if (p < c.world && p != c.rank) {
    const int base = family * kPadWordsPerFamily + chan * kMaxWorld;
    uint32_t* remote = c.pad[p] + base + c.rank;
    uint32_t* local = c.pad[c.rank] + base + p;
    const long long t0 = globaltimer_ns();
    while (cas_release_sys(remote, 0u, 1u) != 0u) {
      if (globaltimer_ns() - t0 > c.timeout_ns) __trap();
    }
    while (cas_acquire_sys(local, 1u, 0u) != 1u) {
      if (globaltimer_ns() - t0 > c.timeout_ns) __trap();
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------- NVLS primitives
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {   // sum over all ranks, fp32 accumulation, one rounding
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_16B(void* mc_addr, const uint4& v) {    // replicated into every rank's buffer by the switch
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_16B(const void* p) {   // peer (NVLink) or local read-once data
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_relaxed_sys_16B(const void* p) {   // peer data that another GPU has just written (no non-coherent path)
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
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

// ---------------------------------------------------------------- reduce-scatter
// 512 threads x <= 64 registers, no shared memory: a communication CTA fits on an SM NEXT TO a persistent GEMM CTA (256 threads x 112
// registers, ~200 KB smem), so the collectives never wait for - or displace - the tensor-core kernels they overlap.
// Unit = n_shard * world bf16 elements at byte offset `off` of the symmetric buffer; this rank owns elements [rank*n_shard, (rank+1)*n_shard).
template <bool NVLS, int UNROLL>
__global__ void __launch_bounds__(512, 2) reduce_scatter_kernel(Peers c, size_t off, int64_t n_shard) {
  cross_rank_barrier(c, 0, blockIdx.x);   // every rank's gradients of this unit are complete (each rank launches this kernel behind them)
  const size_t slice = off + static_cast<size_t>(c.rank) * n_shard * 2;
  uint4* dst = reinterpret_cast<uint4*>(c.data[c.rank] + slice);
  const int64_t nvec = n_shard / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
    if (NVLS) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < nvec) v[u] = multimem_ld_reduce_bf16x8(c.mc + slice + i * 16);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < nvec) dst[i] = v[u];
      }
    } else {
      float acc[UNROLL][8];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
      for (int j = 0; j < c.world; ++j) {   // rank order 0..N-1 on every rank: the sum does not depend on who computes it
        const uint4* src = reinterpret_cast<const uint4*>(c.data[j] + slice);
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int64_t i = i0 + u * stride;
          v[u] = i < nvec ? ld_relaxed_sys_16B(src + i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc8(acc[u], v[u]);
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
        }
      }
    }
  }
  // nobody may overwrite its gradient buffer (the next backward) while a peer still reads it
  cross_rank_barrier(c, 0, blockIdx.x);
}

// ---------------------------------------------------------------- all-gather
template <bool NVLS, int UNROLL>
__global__ void __launch_bounds__(512, 2) all_gather_kernel(Peers c, size_t off, int64_t n_shard) {
  // every rank's slice holds the updated parameters (each rank launches this kernel behind its AdamW) and nobody still reads the old
  // values of the slices about to be overwritten (the previous step's backward precedes the grad-norm all-reduce that AdamW waits for)
  cross_rank_barrier(c, 1, blockIdx.x);
  const int64_t nvec = n_shard / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (NVLS) {
    const size_t slice = off + static_cast<size_t>(c.rank) * n_shard * 2;
    const uint4* src = reinterpret_cast<const uint4*>(c.data[c.rank] + slice);
    for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < nvec) v[u] = ld_stream_16B(src + i);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < nvec) multimem_st_16B(c.mc + slice + i * 16, v[u]);
      }
    }
  } else {
    for (int k = 1; k < c.world; ++k) {
      const int j = (c.rank + k) % c.world;   // staggered: at any moment the N ranks read from N different peers
      const size_t slice = off + static_cast<size_t>(j) * n_shard * 2;
      const uint4* src = reinterpret_cast<const uint4*>(c.data[j] + slice);
      uint4* dst = reinterpret_cast<uint4*>(c.data[c.rank] + slice);
      for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int64_t i = i0 + u * stride;
          if (i < nvec) v[u] = ld_relaxed_sys_16B(src + i);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int64_t i = i0 + u * stride;
          if (i < nvec) dst[i] = v[u];
        }
      }
    }
  }
  __threadfence_system();   // the multicast / local stores above are performed before the release in the barrier below
  // NVLS: every peer's stores into MY buffer have landed when its CTA signals; P2P: nobody updates a slice a peer is still pulling
  cross_rank_barrier(c, 1, blockIdx.x);
}

// ---------------------------------------------------------------- scalar all-reduce (grad-norm^2, reported loss)
// vals[0..n) := sum over ranks, in rank order on every rank (bit-identical results everywhere).  One warp: every rank stores its values
// into slot [parity][j][rank] of EVERY rank's exchange area (remote stores over NVLink), the ranks meet, and each sums its local copy.
// Keeps NCCL out of the step: with these kernels, the reduce-scatters and the all-gathers on ONE stream, every cross-rank wait of the
// step is part of a single sequence that is identical on all ranks, so no launch-queue aliasing with a second spinning kernel family
// (NCCL's) can close a cycle.
__global__ void __launch_bounds__(32) allreduce_scalars_kernel(Peers c, float* __restrict__ vals, int n, int parity) {
  const int j = threadIdx.x;
  const size_t xoff = kFlagBytes + static_cast<size_t>(parity) * kMaxScalars * kMaxWorld * sizeof(float);
  if (j < n) {
    const float v = vals[j];
    for (int p = 0; p < c.world; ++p) {
      float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(c.pad[p]) + xoff) + j * kMaxWorld + c.rank;
      asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
    }
  }
  cross_rank_barrier(c, 2, 0);
  if (j < n) {
    const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.pad[c.rank]) + xoff) + j * kMaxWorld;
    float s = 0.f;
    for (int r = 0; r < c.world; ++r) {
      float t;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(t) : "l"(src + r) : "memory");
      s += t;
    }
    vals[j] = s;
  }
}

// ---------------------------------------------------------------- host side
static Peers make_peers(const CommCtx* c, int slot) {
  Peers p;
  for (int j = 0; j < kMaxWorld; ++j) {
    p.data[j] = j < c->world ? c->buf[slot].peer[j] : nullptr;
    p.pad[j] = j < c->world ? c->pad[j] : nullptr;
  }
  p.mc = c->buf[slot].mc;
  p.rank = c->rank;
  p.world = c->world;
  p.timeout_ns = c->timeout_ns;
  return p;
}

int ctx_create(CommCtx** out, int rank, int world) {
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return set_error(B200_ERR_ARG, "b200_ctx_create: rank %d / world %d (1..%d ranks of one NVSwitch box)", rank, world, kMaxWorld);
  CommCtx* c = new (std::nothrow) CommCtx();
  if (!c) return set_error(B200_ERR_ARG, "b200_ctx_create: out of host memory");
  c->rank = rank;
  c->world = world;
  c->timeout_ns = 60LL * 1000 * 1000 * 1000;
  c->scalar_calls = 0;
  *out = c;
  return 0;
}
int ctx_destroy(CommCtx* c) {
  delete c;
  return 0;
}
int ctx_set_timeout_ms(CommCtx* c, int64_t ms) {
  if (!c || ms <= 0) return set_error(B200_ERR_ARG, "b200_ctx_set_timeout_ms: bad argument");
  c->timeout_ns = ms * 1000000LL;
  return 0;
}
int ctx_set_signal_pad(CommCtx* c, void* const* pads, size_t bytes) {
  if (!c || !pads) return set_error(B200_ERR_ARG, "b200_ctx_set_signal_pad: null argument");
  if (bytes < kPadBytes) return set_error(B200_ERR_ARG, "b200_ctx_set_signal_pad: %zu bytes, need %zu", bytes, kPadBytes);
  for (int j = 0; j < c->world; ++j) {
    if (!pads[j]) return set_error(B200_ERR_ARG, "b200_ctx_set_signal_pad: rank %d pad is null", j);
    c->pad[j] = static_cast<uint32_t*>(pads[j]);
  }
  c->pad_bytes = bytes;
  return 0;
}
int ctx_register_buffer(CommCtx* c, int slot, void* const* peer_ptrs, void* mc_ptr, size_t bytes) {
  if (!c || !peer_ptrs || slot < 0 || slot >= kMaxBuffers) return set_error(B200_ERR_ARG, "b200_ctx_register_buffer: bad slot %d", slot);
  for (int j = 0; j < c->world; ++j) {
    if (!peer_ptrs[j] || (reinterpret_cast<uintptr_t>(peer_ptrs[j]) & 15)) return set_error(B200_ERR_ARG, "b200_ctx_register_buffer: rank %d pointer null / not 16-byte aligned", j);
    c->buf[slot].peer[j] = static_cast<char*>(peer_ptrs[j]);
  }
  c->buf[slot].mc = static_cast<char*>(mc_ptr);
  c->buf[slot].bytes = bytes;
  return 0;
}
size_t ctx_signal_pad_bytes() { return kPadBytes; }
int ctx_has_multicast(const CommCtx* c, int slot) { return c && slot >= 0 && slot < kMaxBuffers && c->buf[slot].mc != nullptr; }

static int check_unit(const CommCtx* c, int slot, size_t off, int64_t n_shard, int ctas, const char* who) {
  if (!c || slot < 0 || slot >= kMaxBuffers || !c->buf[slot].peer[c->rank]) return set_error(B200_ERR_ARG, "%s: buffer slot %d is not registered", who, slot);
  if (!c->pad[c->rank]) return set_error(B200_ERR_ARG, "%s: no signal pad registered", who);
  if (n_shard <= 0 || n_shard % 8 || off % 16) return set_error(B200_ERR_ARG, "%s: shard of %lld elements at byte offset %zu (need a multiple of 8 elements, 16-byte aligned)", who, (long long)n_shard, off);
  if (off + static_cast<size_t>(n_shard) * c->world * 2 > c->buf[slot].bytes) return set_error(B200_ERR_ARG, "%s: unit exceeds the registered buffer", who);
  if (ctas < 1 || ctas > kMaxCommCtas) return set_error(B200_ERR_ARG, "%s: ctas %d outside 1..%d", who, ctas, kMaxCommCtas);
  return 0;
}

// mode: 0 = NVLS if the buffer has a multicast mapping else P2P, 1 = force P2P
int reducescatter_layer(CommCtx* c, int slot, size_t off, int64_t n_shard, int mode, int ctas, cudaStream_t st) {
  if (int rc = check_unit(c, slot, off, n_shard, ctas, "b200_reducescatter_layer")) return rc;
  const Peers p = make_peers(c, slot);
  if (p.mc && mode == 0)
    reduce_scatter_kernel<true, 4><<<ctas, 512, 0, st>>>(p, off, n_shard);
  else
    reduce_scatter_kernel<false, 4><<<ctas, 512, 0, st>>>(p, off, n_shard);
  B200_CHECK_LAUNCH("reduce_scatter_kernel");
  return 0;
}
int allgather_layer(CommCtx* c, int slot, size_t off, int64_t n_shard, int mode, int ctas, cudaStream_t st) {
  if (int rc = check_unit(c, slot, off, n_shard, ctas, "b200_allgather_layer")) return rc;
  const Peers p = make_peers(c, slot);
  if (p.mc && mode == 0)
    all_gather_kernel<true, 4><<<ctas, 512, 0, st>>>(p, off, n_shard);
  else
    all_gather_kernel<false, 4><<<ctas, 512, 0, st>>>(p, off, n_shard);
  B200_CHECK_LAUNCH("all_gather_kernel");
  return 0;
}

// vals: n <= 16 fp32 values in this rank's device memory, replaced by their sum over all ranks (rank order, identical on every rank)
int allreduce_scalars(CommCtx* c, float* vals, int n, cudaStream_t st) {
  if (!c || !vals || n < 1 || n > kMaxScalars) return set_error(B200_ERR_ARG, "b200_allreduce_scalars: n %d outside 1..%d", n, kMaxScalars);
  if (!c->pad[c->rank]) return set_error(B200_ERR_ARG, "b200_allreduce_scalars: no signal pad registered");
  Peers p;
  for (int j = 0; j < kMaxWorld; ++j) {
    p.data[j] = nullptr;
    p.pad[j] = j < c->world ? c->pad[j] : nullptr;
  }
  p.mc = nullptr;
  p.rank = c->rank;
  p.world = c->world;
  p.timeout_ns = c->timeout_ns;
  allreduce_scalars_kernel<<<1, 32, 0, st>>>(p, vals, n, static_cast<int>(c->scalar_calls++ & 1));
  B200_CHECK_LAUNCH("allreduce_scalars_kernel");
  return 0;
}

}  // namespace b200
