// bf16 GEMM, CTA-pair version: tcgen05.mma.cta_group::2 (M = 256 per pair), cluster of two CTAs on one TPC.
// Same contract and operand layouts as gemm_tcgen05.cu (NT / NN / TN without transposes).  Why a pair:
//   a 128x256 single-CTA tile pulls 48 KB of operands from L2 per 64-wide k-block (96 B/clk/SM at tensor peak: the
//   measured limiter); a pair computes a 256x256 tile and each CTA loads its 128 rows of A plus only HALF of B
//   (the MMA reads B from both CTAs' shared memory): 32 KB per CTA per k-block, one third less L2 traffic and
//   one third less shared-memory fill per FLOP.
// Roles per CTA: warp0 TMA producer (own A rows + own B half, signalling the LEADER's full barrier), warp1 of the leader
// issues the MMAs and multicasts the commits (smem-slot release, accumulator-ready) to both CTAs, warp2 TMEM allocator
// (cta_group::2, both CTAs), warps4-7 epilogue of the CTA's own 128 accumulator rows (TMEM lanes) -> TMA store.
// Optional SwiGLU epilogue (GEMM_FLAG_SWIGLU, x @ [W_gate; W_up]^T): for column tile j the leader stages W rows j*128.. (gate) and the peer
// rows F + j*128.. (up) as their halves of B, so the 256 accumulator columns are gate and up of the SAME 128 features side by side; the
// epilogue stores both (the backward needs them) and also a = bf16(bf16(silu(g)) * u) computed from the rounded values - bit-identical to
// swiglu_fwd_kernel, without re-reading gu from HBM and with no permutation of the weight rows.
// Tile scheduling, two modes (b200_set_option("gemm_sched", 0|1)):
//   0  static persistent: one cluster per SM pair, cluster c walks tiles c, c + #clusters, ...
//   1  cluster launch control (Blackwell CLC): the grid has one cluster per tile; a running cluster that finishes a tile cancels a
//      not-yet-launched cluster with clusterlaunchcontrol.try_cancel and computes that cluster's tile itself.  Tiles therefore flow to
//      whichever SM pairs are actually available: SMs held by a co-running kernel (NCCL channels, attention CTAs that cover only part
//      of the GPU, the other stream's GEMM) simply take fewer tiles, and SM pairs freed by another kernel start taking tiles at once
//      (the hardware launches pending clusters there).  With static striding a cluster whose SM pair is busy elsewhere still owns
//      1/#clusters of the tiles and finishes them alone at the end.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"
#include "common.h"

#pragma nv_diag_suppress 128   // "loop is not reachable": the plain epilogue loop in the SWIGLU instantiation (that branch ends in `continue`)

namespace b200 {

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                      uint32_t box_rows);

namespace pair {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;   // columns per pair tile; each CTA stages BN/2 rows of B
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int THREADS = 256;
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 128;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB (this CTA's half)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_OFF = STAGES * STAGE_BYTES;
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;
constexpr int BAR_OFF = EPI_OFF + EPI_BYTES;
constexpr int CLC_SLOTS = 4;                  // in-flight "next tile" responses (the producer asks one tile ahead; the epilogue lags <= 3 tiles)
constexpr int NUM_BARS = 2 * STAGES + 4 + EPI_WARPS + 2 * CLC_SLOTS;
constexpr int CLC_RESP_OFF = BAR_OFF + NUM_BARS * 8 + 16;      // 16-byte aligned (BAR_OFF is, NUM_BARS is even)
constexpr int DYN_BYTES = CLC_RESP_OFF + CLC_SLOTS * 16 + 1024;
static_assert(NUM_BARS % 2 == 0, "CLC response slots must stay 16-byte aligned");
// consumers of one CLC response: leader {producer, MMA thread, 4 epilogue warps} + peer {producer, 4 epilogue warps}
constexpr int CLC_CONSUMERS = (2 + EPI_WARPS) + (1 + EPI_WARPS);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of the pair into ITS OWN shared memory; the transaction bytes are credited to the barrier at the
// same offset in the LEADER CTA (peer bit of the barrier address cleared, cf. CUTLASS SM100_TMA_2SM_LOAD).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive (once all prior MMAs of this thread are done) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// arm (arrive + expect `bytes`) the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint64_t* bar, uint32_t cta, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [ra], %2;\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta), "r"(bytes)
      : "memory");
}
// Ask the hardware for a not-yet-launched cluster of this grid.  The 16-byte response is written to `resp` in EVERY CTA of the
// cluster and completes 16 transaction bytes on the barrier at `bar`'s offset in every CTA.
__device__ __forceinline__ void clc_try_cancel_multicast(void* resp, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];" ::"r"(
                   smem_u32(resp)),
               "r"(smem_u32(bar))
               : "memory");
}
// Decode a response: linear tile index (= cluster index = first ctaid.x / 2) of the cancelled cluster, or -1 when nothing was left.
__device__ __forceinline__ int clc_decode(const void* resp) {
  uint32_t x, valid;
  asm volatile(
      "{\n"
      ".reg .pred p1;\n"
      ".reg .b128 r;\n"
      "ld.shared.b128 r, [%2];\n"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
      "selp.u32 %1, 1, 0, p1;\n"
      "mov.u32 %0, 0;\n"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %0, r;\n"
      "}"
      : "=r"(x), "=r"(valid)
      : "r"(smem_u32(resp))
      : "memory");
  fence_proxy_async_smem();   // this generic-proxy read is ordered before the async-proxy write of the slot's next response
  return valid ? static_cast<int>(x >> 1) : -1;
}

__device__ __forceinline__ float rcp_fast_(float x) {  // same MUFU reciprocal as elementwise.cu's SwiGLU kernel
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Sched {
  int tiles_m, tiles_n, total, group_m;
  __device__ __forceinline__ void coords(int t, int& tm, int& tn) const {
    const int per_group = group_m * tiles_n;
    const int g = t / per_group;
    const int first_m = g * group_m;
    const int gm = min(group_m, tiles_m - first_m);
    const int r = t - g * per_group;
    tm = first_m + (r % gm);
    tn = r / gm;
  }
};

template <bool A_MN, bool B_MN, bool SWIGLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmR, int M, int N, int K, int flags, int group_m, int use_clc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* epi_bar = bars + 2 * STAGES + 4;
  uint64_t* clc_full = bars + 2 * STAGES + 4 + EPI_WARPS;                // per CTA: response k has landed in this CTA's slot
  uint64_t* clc_empty = clc_full + CLC_SLOTS;                           // (used in the leader) every consumer of both CTAs has read it
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  uint8_t* clc_resp = smem + CLC_RESP_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  Sched sched;
  sched.tiles_m = (M + 2 * BM - 1) / (2 * BM);
  sched.tiles_n = (N + BN - 1) / BN;
  sched.total = sched.tiles_m * sched.tiles_n;
  sched.group_m = group_m;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (SWIGLU || (flags & GEMM_FLAG_RESIDUAL)) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // the leader's arrive.expect_tx; the bytes of BOTH CTAs' loads are credited here
      mbar_init(&empty_bar[i], 1);  // multicast commit from the leader's MMA thread
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);               // multicast commit
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);  // (used in the leader) epilogue warps of both CTAs
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&epi_bar[i], 1);
    for (int i = 0; i < CLC_SLOTS; ++i) {
      mbar_init(&clc_full[i], 1);               // armed by the leader's producer (arrive.expect_tx 16, local and remote)
      mbar_init(&clc_empty[i], CLC_CONSUMERS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_holder, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / multicast commit / 2-SM TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // Next tile of this cluster after its k-th one.  Static: stride by the number of clusters.  CLC: every consumer waits for response k in
  // its own CTA, decodes it and hands the slot back to the leader (whose producer re-uses it for response k + CLC_SLOTS).
  auto next_tile = [&](int t, int k) -> int {
    if (!use_clc) {
      t += ncl;
      return t < sched.total ? t : -1;
    }
    const int slot = k % CLC_SLOTS;
    mbar_wait(&clc_full[slot], (k / CLC_SLOTS) & 1);
    const int nt = clc_decode(clc_resp + slot * 16);
    mbar_arrive_remote(&clc_empty[slot], 0);
    return nt;
  };

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs); the leader's also drives the tile scheduler
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int kt = 0;
      for (int t = cid; t >= 0; t = next_tile(t, kt++)) {
        if (use_clc && rank == 0) {
          // ask for the tile after this one now, so the answer is there when this tile's loads have been issued; asking only one tile
          // ahead keeps a cluster from hoarding tiles it will not reach before others go idle
          const int slot = kt % CLC_SLOTS;
          mbar_wait(&clc_empty[slot], ((kt / CLC_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&clc_full[slot], 16);
          mbar_arrive_expect_tx_remote(&clc_full[slot], 1, 16);
          clc_try_cancel_multicast(clc_resp + slot * 16, &clc_full[slot]);
        }
        int tm, tn;
        sched.coords(t, tm, tn);
        const int m0 = tm * 2 * BM + rank * BM;        // this CTA's 128 rows of the 256-row tile
        // this CTA's half of the B rows; SwiGLU: the leader's half is the gate block of feature tile tn, the peer's the matching up block
        const int n0 = SWIGLU ? tn * (BN / 2) + rank * (N / 2) : tn * BN + rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const int k0 = kb * BK;
          if constexpr (!A_MN) {
            tma_load_2d_pair(sA, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + j * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair(sB, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_pair(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? 16 * 128 : 32, B_KSTEP = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = cid; t >= 0; t = next_tile(t, iter++)) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_base + k * A_KSTEP, A_LBO, 1024);
            const uint64_t db = make_smem_desc_sw128(b_base + k * B_KSTEP, B_LBO, 1024);
            umma_bf16_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tmem_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue warps (both CTAs: own 128 rows)
    const int w = warp - 4;
    uint8_t* stg_base = smem + EPI_OFF + w * 2 * EPI_BUF_BYTES;
    const bool resid = (flags & GEMM_FLAG_RESIDUAL) != 0;
    const bool round_first = (flags & GEMM_FLAG_ROUND_BEFORE_ADD) != 0;
    uint32_t epi_phase = 0;
    int chunk_ctr = 0;
    int iter = 0;
    for (int t = cid; t >= 0; ++iter) {
      int tm, tn;
      sched.coords(t, tm, tn);
      const int m0 = tm * 2 * BM + rank * BM + w * 32, n0 = tn * BN;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(w * 32) << 16) + acc * BN;
      if constexpr (SWIGLU) {
        // columns [0,128) = gate, [128,256) = up of features tn*128 .. tn*128+127.  Per 64-feature chunk: three 32x64 bf16 tiles leave
        // through the two staging buffers (gate -> C[:, f], up -> C[:, F + f], a -> R[:, f]).
        const int F = N / 2, f0 = tn * (BN / 2);
        const bool live = m0 < M;
        auto put_tile = [&](const uint32_t (&pk)[32], const CUtensorMap* tm, int col) {
          uint8_t* stg = stg_base + (chunk_ctr & 1) * EPI_BUF_BYTES;
          ++chunk_ctr;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          if (live) {
            uint8_t* row_ptr = stg + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 o;
              o.x = pk[4 * j]; o.y = pk[4 * j + 1]; o.z = pk[4 * j + 2]; o.w = pk[4 * j + 3];
              *reinterpret_cast<uint4*>(row_ptr + ((j ^ (lane & 7)) << 4)) = o;
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (live && lane == 0) tma_store_2d(tm, stg, col, m0);
          if (lane == 0) tma_store_commit();
        };
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[64], gp[32], up[32], ap[32];
          tmem_ld_32x32b_x32(t_row + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32b_x32(t_row + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) gp[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          put_tile(gp, &tmC, f0 + c * 64);
          tmem_ld_32x32b_x32(t_row + BN / 2 + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32b_x32(t_row + BN / 2 + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) up[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          put_tile(up, &tmC, F + f0 + c * 64);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gp[e]));
            const float2 u = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&up[e]));
            const float s0 = bf16_round(g.x * rcp_fast_(1.f + __expf(-g.x))), s1 = bf16_round(g.y * rcp_fast_(1.f + __expf(-g.y)));
            ap[e] = pack_bf16x2(s0 * u.x, s1 * u.y);
          }
          put_tile(ap, &tmR, f0 + c * 64);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
        int nt = 0;
        if (lane == 0) nt = next_tile(t, iter);
        t = __shfl_sync(0xffffffffu, nt, 0);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c, ++chunk_ctr) {
        uint8_t* stg = stg_base + (chunk_ctr & 1) * EPI_BUF_BYTES;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const bool live = (m0 < M) && (n0 + c * 64 < N);
        if (resid && live && lane == 0) {
          mbar_arrive_expect_tx(&epi_bar[w], EPI_BUF_BYTES);
          tma_load_2d(stg, &tmR, &epi_bar[w], n0 + c * 64, m0);
        }
        uint32_t v[64];
        tmem_ld_32x32b_x32(t_row + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32b_x32(t_row + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        if (resid && live) {
          mbar_wait(&epi_bar[w], epi_phase);
          epi_phase ^= 1;
        }
        if (live) {
          uint8_t* row_ptr = stg + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4* p = reinterpret_cast<uint4*>(row_ptr + ((j ^ (lane & 7)) << 4));
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
            if (resid) {
              const uint4 r = *p;
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[e]));
                float a0 = f[2 * e], a1 = f[2 * e + 1];
                if (round_first) {
                  a0 = bf16_round(a0);
                  a1 = bf16_round(a1);
                }
                f[2 * e] = a0 + rf.x;
                f[2 * e + 1] = a1 + rf.y;
              }
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            *p = o;
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (live && lane == 0) tma_store_2d(&tmC, stg, n0 + c * 64, m0);
        if (lane == 0) tma_store_commit();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);  // the leader's MMA thread owns the accumulator hand-back
      int nt = 0;
      if (lane == 0) nt = next_tile(t, iter);
      t = __shfl_sync(0xffffffffu, nt, 0);
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA may retire (or free TMEM) while its peer can still touch its barriers / shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

static int g_sms = 0;
int sched_mode = 0;   // b200_set_option("gemm_sched", 0 = static persistent (default) | 1 = cluster launch control)

template <bool A_MN, bool B_MN, bool SWIGLU = false>
static int launch(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tR, int M, int N, int K, int flags,
                  int group_m, int max_ctas, cudaStream_t stream) {
  auto kern = gemm_pair_kernel<A_MN, B_MN, SWIGLU>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e));
    configured = true;
  }
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  int clusters = g_sms / 2;
  if (max_ctas > 0 && max_ctas / 2 < clusters) clusters = max_ctas / 2;
  if (tiles < clusters) clusters = tiles;
  if (clusters < 1) clusters = 1;
  const int use_clc = sched_mode == 1 && max_ctas <= 0;
  if (use_clc) clusters = tiles;   // one cluster per tile; running clusters cancel and absorb the pending ones
  kern<<<2 * clusters, THREADS, DYN_BYTES, stream>>>(tA, tB, tC, tR, M, N, K, flags, group_m, use_clc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "gemm(pair) launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace pair

int gemm_bf16_tcgen05_pair(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, const void* R, int ldr, int M, int N,
                           int K, int flags, int group_m, int max_ctas, cudaStream_t stream) {
  using namespace pair;
  CUtensorMap tA, tB, tC, tR;
  int rc;
  const bool a_mn = (kind == GEMM_TN);
  const bool b_mn = (kind == GEMM_NN || kind == GEMM_TN);
  rc = a_mn ? make_tmap_2d_bf16(&tA, A, K, M, lda, 64, BK) : make_tmap_2d_bf16(&tA, A, M, K, lda, BK, BM);
  if (rc) return rc;
  rc = b_mn ? make_tmap_2d_bf16(&tB, B, K, N, ldb, 64, BK) : make_tmap_2d_bf16(&tB, B, N, K, ldb, BK, BN / 2);
  if (rc) return rc;
  if ((rc = make_tmap_2d_bf16(&tC, C, M, N, ldc, 64, 32))) return rc;
  if (flags & GEMM_FLAG_SWIGLU) {
    if ((rc = make_tmap_2d_bf16(&tR, R, M, N / 2, ldr, 64, 32))) return rc;    // the activation output a [M, F]
  } else if (flags & GEMM_FLAG_RESIDUAL) {
    if ((rc = make_tmap_2d_bf16(&tR, R, M, N, ldr, 64, 32))) return rc;
  } else {
    tR = tC;
  }
  if (group_m <= 0) {
    // L2 rasterisation (in 256-row tiles), from the sweep in profiles/r1_gemm_groupm_sweep.md: wide outputs (more column tiles than
    // row tiles: gate/up, lm_head, da) walk all row tiles of a column panel first (the B panel is shared by the whole wave); square or
    // tall outputs walk column tiles first.
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN;
    group_m = tiles_n > tiles_m ? (tiles_m < 16 ? tiles_m : 16) : 1;
  }
  if (kind == GEMM_NT && (flags & GEMM_FLAG_SWIGLU)) return launch<false, false, true>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  if (kind == GEMM_NT) return launch<false, false>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  if (kind == GEMM_NN) return launch<false, true>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  if (kind == GEMM_TN) return launch<true, true>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  return set_error(B200_ERR_ARG, "unknown gemm kind %d", kind);
}

}  // namespace b200
