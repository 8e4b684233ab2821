// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed), sm_100a only.
//
//   C[M,N] = op(A) * op(B)  (+ R)          fp32 accumulate, bf16 in/out
//
// Replaces the nn.Linear fwd / dgrad / wgrad cuBLAS calls of the reference's transformer block
// (/root/reference/nemo_automodel/components/models/llama/model.py:113-115,151,170,511 and their autograd):
//   fwd   : A = X  [M,K] (K contiguous)      B = W  [N,K] (K contiguous)     -> kind NT  (A K-major, B K-major)
//   dgrad : A = dY [M,K] (K contiguous)      B = W  [K,N] (N contiguous)     -> kind NN  (A K-major, B MN-major)
//   wgrad : A = dY [K,M] (M contiguous)      B = X  [K,N] (N contiguous)     -> kind TN  (A MN-major, B MN-major)
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: global -> 128B-swizzled smem ring (STAGES deep), mbarrier complete_tx
//   warp 1      MMA issuer: one thread issues tcgen05.mma (128 x BN x 16) reading smem descriptors,
//               accumulating into one of two TMEM accumulator buffers; tcgen05.commit frees smem slots
//   warp 2      TMEM allocator
//   warps 4..7  epilogue: tcgen05.ld (each thread = one accumulator row) -> bf16 -> swizzled smem -> TMA store;
//               optional residual tile R is TMA-loaded into the same staging buffer and added in place.
//   The epilogue of tile i overlaps the main loop of tile i+1 (double-buffered TMEM, 2 x BN fp32 columns).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <unordered_map>

#include "ptx.cuh"
#include "common.h"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 128;  // one warp: 32 rows x 64 bf16

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + EPI_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + EPI_WARPS;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

struct TileSched {
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

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M, int N,
                         int K, int flags, int group_m) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* epi_bar = bars + 2 * STAGES + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // allocation must be a power of two >= 32

  TileSched sched;
  sched.tiles_m = (M + BM - 1) / BM;
  sched.tiles_n = (N + BN - 1) / BN;
  sched.total = sched.tiles_m * sched.tiles_n;
  sched.group_m = group_m;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (flags & GEMM_FLAG_RESIDUAL) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&epi_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < sched.total; t += gridDim.x) {
        int tm, tn;
        sched.coords(t, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          const int k0 = kb * BK;
          if constexpr (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);  // box (64 k, 128 m)
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + j * 64, k0);  // box (64 m, 64 k)
          }
          if constexpr (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);  // box (64 k, BN n)
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + j * 64, k0);  // box (64 n, 64 k)
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major  SW128: 8-row groups 1024B apart (SBO); one swizzle atom along K (LBO unused, set to 16B)
      // MN-major SW128: 64-element MN atoms BK*128 B apart (LBO); 8-k-row groups 1024B apart (SBO)
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? 16 * 128 : 32, B_KSTEP = B_MN ? 16 * 128 : 32;  // bytes per UMMA_K=16
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < sched.total; t += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_base = a_base + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_base + k * A_KSTEP, A_LBO, 1024);
            const uint64_t db = make_smem_desc_sw128(b_base + k * B_KSTEP, B_LBO, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue warps
    const int w = warp - 4;  // == warp % 4: TMEM lanes [32w, 32w+32)
    uint8_t* stg_base = smem + L::EPI_OFF + w * 2 * EPI_BUF_BYTES;
    const bool resid = (flags & GEMM_FLAG_RESIDUAL) != 0;
    const bool round_first = (flags & GEMM_FLAG_ROUND_BEFORE_ADD) != 0;
    uint32_t epi_phase = 0;
    int chunk_ctr = 0;
    int iter = 0;
    for (int t = blockIdx.x; t < sched.total; t += gridDim.x, ++iter) {
      int tm, tn;
      sched.coords(t, tm, tn);
      const int m0 = tm * BM + w * 32, n0 = tn * BN;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(w * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c, ++chunk_ctr) {
        uint8_t* stg = stg_base + (chunk_ctr & 1) * EPI_BUF_BYTES;
        if (lane == 0) tma_store_wait_read<1>();  // the store issued from this buffer two chunks ago has drained
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
        if (live && lane == 0) {
          tma_store_2d(&tmC, stg, n0 + c * 64, m0);
        }
        if (lane == 0) tma_store_commit();
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2D row-major tensor [rows, cols] with leading dimension ld (elements); box = (box_cols, box_rows), 128B swizzle.
static int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, CUtensorMapDataType dt, int esize);
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                      uint32_t box_rows) {
  return make_tmap_2d(out, ptr, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}
int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                     uint32_t box_rows) {
  return make_tmap_2d(out, ptr, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}
static int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, CUtensorMapDataType dt, int esize) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * esize) & 15))
    return set_error(B200_ERR_ARG, "TMA operand must be 16B aligned with a 16B-multiple row pitch");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tR, int M,
                       int N, int K, int flags, int group_m, int max_ctas, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (tiles < grid) grid = tiles;
  kern<<<grid, GEMM_THREADS, L::DYN_BYTES, stream>>>(tA, tB, tC, tR, M, N, K, flags, group_m);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  return 0;
}

template <int BN, int STAGES>
static int gemm_dispatch_kind(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, const void* R,
                              int ldr, int M, int N, int K, int flags, int group_m, int max_ctas, cudaStream_t stream) {
  CUtensorMap tA, tB, tC, tR;
  int rc;
  const bool a_mn = (kind == GEMM_TN);
  const bool b_mn = (kind == GEMM_NN || kind == GEMM_TN);
  // A: K-major -> tensor [M rows, K cols], box (64, 128);  MN-major -> tensor [K rows, M cols], box (64, 64)
  rc = a_mn ? make_tmap_2d_bf16(&tA, A, K, M, lda, 64, BK) : make_tmap_2d_bf16(&tA, A, M, K, lda, BK, BM);
  if (rc) return rc;
  rc = b_mn ? make_tmap_2d_bf16(&tB, B, K, N, ldb, 64, BK) : make_tmap_2d_bf16(&tB, B, N, K, ldb, BK, BN);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tC, C, M, N, ldc, 64, 32);
  if (rc) return rc;
  if (flags & GEMM_FLAG_RESIDUAL) {
    rc = make_tmap_2d_bf16(&tR, R, M, N, ldr, 64, 32);
    if (rc) return rc;
  } else {
    tR = tC;
  }
  if (kind == GEMM_NT) return launch_gemm<BN, STAGES, false, false>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  if (kind == GEMM_NN) return launch_gemm<BN, STAGES, false, true>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  if (kind == GEMM_TN) return launch_gemm<BN, STAGES, true, true>(tA, tB, tC, tR, M, N, K, flags, group_m, max_ctas, stream);
  return set_error(B200_ERR_ARG, "unknown gemm kind %d", kind);
}

int forced_bn = 0;  // debug: b200_set_option("gemm_bn", 128|192|256) pins the tile width
int use_pair = 1;   // b200_set_option("gemm_2cta", 0|1): CTA-pair kernel (gemm_tcgen05_2cta.cu) for problems with M, N >= 256
int gemm_bf16_tcgen05_pair(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, const void* R, int ldr, int M, int N,
                           int K, int flags, int group_m, int max_ctas, cudaStream_t stream);

int gemm_bf16_tcgen05(int kind, const void* A, int lda, const void* B, int ldb, void* C, int ldc, const void* R, int ldr,
                      int M, int N, int K, int flags, int group_m, int max_ctas, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(B200_ERR_ARG, "gemm: empty problem %dx%dx%d", M, N, K);
  if ((flags & GEMM_FLAG_RESIDUAL) && !R) return set_error(B200_ERR_ARG, "gemm: residual flag without R");
  if (flags & GEMM_FLAG_SWIGLU) {
    if (kind != GEMM_NT || (flags & GEMM_FLAG_RESIDUAL) || !R || N % 256 != 0 || M < 256 || !use_pair || forced_bn != 0)
      return set_error(B200_ERR_UNSUPPORTED, "gemm: the SwiGLU epilogue needs the CTA-pair kernel, kind NT, M >= 256, N = 2F with F %% 128 == 0, R = the [M, F] output");
  }
  if (use_pair && M >= 256 && N >= 256 && forced_bn == 0)
    return gemm_bf16_tcgen05_pair(kind, A, lda, B, ldb, C, ldc, R, ldr, M, N, K, flags, group_m, max_ctas, stream);
  if (group_m <= 0) group_m = 8;
  // 128x256 tiles whenever N allows (measured: 128x192 / 128x128 tiles lose more in per-tile efficiency than they win back in
  // wave quantisation on every Llama-3-8B shape, profiles/r1_gemm_tile_sweep.md)
  int bn = ((N % 256 == 0) || N >= 1024) ? 256 : 128;
  if (forced_bn == 128 || forced_bn == 192 || forced_bn == 256) bn = forced_bn;
  if (bn == 256) return gemm_dispatch_kind<256, 4>(kind, A, lda, B, ldb, C, ldc, R, ldr, M, N, K, flags, group_m, max_ctas, stream);
  if (bn == 192) return gemm_dispatch_kind<192, 4>(kind, A, lda, B, ldb, C, ldc, R, ldr, M, N, K, flags, group_m, max_ctas, stream);
  return gemm_dispatch_kind<128, 6>(kind, A, lda, B, ldb, C, ldc, R, ldr, M, N, K, flags, group_m, max_ctas, stream);
}

}  // namespace b200
