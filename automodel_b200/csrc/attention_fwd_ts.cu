// Flash attention forward, third tcgen05 variant: the probabilities never leave tensor memory.
// Same contract and the same CTA shape as attn_fwd64 (attention_fwd64.cu: 128 q rows x 64-row kv tiles, 256 TMEM columns, two CTAs per
// SM), with the two changes its ncu capture asked for (softmax warps issue-bound, tensor pipe 27-40 % busy):
//   * P_j is written back with tcgen05.st INTO THE COLUMNS OF S_j (bf16 pairs: 32 of S_j's 64 columns) and O += P_j V_j takes its A
//     operand from tensor memory (tcgen05.mma [d], [a_tmem], b_desc).  No swizzled smem tile, no fence.proxy.async, no wait for
//     S-buffer hand-back barrier (PV_{j-1} is issued only after the softmax warps have published P_{j-1}, i.e. read S_{j-1}).
//     Three tensor-memory hazards found on hardware shape the synchronisation (all invisible to tolerance tests, caught by the
//     bit-reproducibility test at the 8B shapes): one p_ready mbarrier per S buffer (a warp running a tile ahead must not complete
//     the previous tile's phase), S_{j+1} is issued only after PV_{j-1} - whose A operand lives in the columns it overwrites - has
//     COMPLETED, and no warp stores P_j while PV_{j-1} is in flight.
//   * softmax instruction diet: row max on the raw scores with 3-input FMNMX3, exp2(s*scale - m) as ONE packed FFMA2 per two elements
//     feeding MUFU.EX2, row sums with FADD2: 3 issue slots per element instead of 5.5.
//   warp 0      TMA producer (Q once; K/V 64-row tiles through 2-stage rings)
//   warp 1      MMA issuer
//   warps 2..5  softmax: one q row per thread (= its TMEM lane), 64 columns per tile, lazy O rescale
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                      uint32_t box_rows);

namespace fwdts {

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct Smem {
  static constexpr int TILE_Q = 128 * D * 2;   // D/64 boxes of [128 rows x 64 cols]
  static constexpr int TILE_KV = 64 * D * 2;   // D/64 boxes of [64 rows x 64 cols]
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = TILE_Q;
  static constexpr int V_OFF = TILE_Q + 2 * TILE_KV;
  static constexpr int XCH_OFF = TILE_Q + 4 * TILE_KV;      // float [3][2][128]: row-max exchange (two tile parities) and final row-sum exchange
  static constexpr int BAR_OFF = XCH_OFF + 3 * 2 * 128 * 4;
  static constexpr int NUM_BARS = 14;
  static constexpr int DYN = BAR_OFF + NUM_BARS * 8 + 16;  // 101,496 B at D=128: two CTAs per SM
};

// HALVES: softmax warps per TMEM lane quarter.  1: one thread owns a whole 64-column row of the tile (4 softmax warps).  2: two warps
// share the rows of a quarter, 32 columns each (8 softmax warps, 4 per scheduler with both CTAs of the SM): the row maximum is exchanged
// through shared memory once per tile, the row sums only at the end; twice the warps to hide the MUFU / TMEM latencies that leave the
// 4-warp version at 53 % of both the tensor and the MUFU pipe.
template <int D, int HALVES>
__global__ void __launch_bounds__(64 + 128 * HALVES, 2)
attn_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   __nv_bfloat16* __restrict__ o, float* __restrict__ lse, const int* __restrict__ cu_seqlens, int64_t ldo, int Hq, int Hkv, int T,
                   float scale_log2) {
  using L = Smem<D>;
  constexpr int ATOMS = D / 64;
  constexpr int QA = 128 * 128;  // bytes of one Q atom  [128 x 64]
  constexpr int KA = 64 * 128;   // bytes of one K/V atom [64 x 64]
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 11;   // [2]: one per S / P buffer.  With a single barrier a softmax warp that runs one tile ahead (the second S
                                   // buffer lets it) would deliver its arrival for tile j+1 into the phase of tile j, completing that phase
                                   // before a slower warp has written its rows of P_j: PV_j would read stale probabilities.
  uint64_t* pv_done = bars + 13;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, h = blockIdx.y;
  const int mt = gridDim.x - 1 - blockIdx.x;  // heavy tiles first
  const int s0 = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - s0;
  const int m0 = mt * 128;
  if (m0 >= len) return;
  const int hk = h / (Hq / Hkv);
  const int n_kv = min(2 * (mt + 1), (len + 63) / 64);  // 64-row kv tiles up to the diagonal

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(&p_ready[0], 4 * HALVES);
    mbar_init(&p_ready[1], 4 * HALVES);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 128;   // S_j (and, after the softmax, P_j) at tS0 + (j & 1) * 64

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, L::TILE_Q);
#pragma unroll
      for (int a = 0; a < ATOMS; ++a) tma_load_2d(smem + L::Q_OFF + a * QA, &tmQ, q_full, h * D + a * 64, s0 + m0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = ((j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[st], ph);
        mbar_arrive_expect_tx(&k_full[st], L::TILE_KV);
#pragma unroll
        for (int a = 0; a < ATOMS; ++a) tma_load_2d(smem + L::K_OFF + st * L::TILE_KV + a * KA, &tmK, &k_full[st], hk * D + a * 64, s0 + j * 64);
        mbar_wait(&v_empty[st], ph);
        mbar_arrive_expect_tx(&v_full[st], L::TILE_KV);
#pragma unroll
        for (int a = 0; a < ATOMS; ++a) tma_load_2d(smem + L::V_OFF + st * L::TILE_KV + a * KA, &tmV, &v_full[st], hk * D + a * 64, s0 + j * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);  // S = Q K^T  (N = 64 kv rows)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);   // O += P V   (A = P from TMEM, K = 64 kv rows, V MN-major)
      const uint32_t q_base = smem_u32(smem + L::Q_OFF);
      auto issue_s = [&](int j) {
        // S_j overwrites the buffer that held S_{j-2} / P_{j-2}: the caller has waited for PV_{j-2} to complete, and PV_{j-2} was issued
        // only after the softmax warps had published P_{j-2} (= finished reading S_{j-2}), so no hand-back from them is needed
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(smem + L::K_OFF + st * L::TILE_KV);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16(tS0 + st * 64, make_smem_desc_sw128(q_base + (kk >> 2) * QA + (kk & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(k_base + (kk >> 2) * KA + (kk & 3) * 32, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          // S_{j+1} overwrites the columns PV_{j-1} reads its A operand (P_{j-1}) from.  Issue order alone does not protect a tensor-memory
          // A operand against the next MMA's accumulator write (seen on hardware as run-to-run differences of the 8B forward), so wait
          // for PV_{j-1} to COMPLETE.  Off the critical path: softmax_j (the CTA's bottleneck) is still running at this point.
          if (j > 0) {
            mbar_wait(pv_done, (j - 1) & 1);
            tc_fence_after();
          }
          issue_s(j + 1);
        }
        const int st = j & 1;
        mbar_wait(&p_ready[st], (j >> 1) & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(smem + L::V_OFF + st * L::TILE_KV);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 64 kv rows / 16: 8 TMEM columns of packed bf16 pairs per step
          umma_bf16_ts(tO, tS0 + st * 64 + kk * 8, make_smem_desc_sw128(v_base + kk * 2048, KA, 1024), idesc_o, (j | kk) != 0 ? 1u : 0u);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    constexpr int CW = 64 / HALVES;          // kv columns of a tile per thread
    constexpr int OC = D / 32 / HALVES;      // 32-column chunks of O per thread (rescale / epilogue)
    const int quad = warp & 3;               // TMEM lane quarter this warp may access (warp id % 4)
    const int half = (warp - 2) >> 2;        // which CW columns of the tile / which O chunks
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int qrow = m0 + r;
    float* xch = reinterpret_cast<float*>(smem + L::XCH_OFF);
    float m_used = 0.f, l_sum = 0.f;   // m_used: running row maximum in the exp2 domain (score * scale * log2 e); l_sum: this thread's columns only
    const uint64_t scale2 = f2_pack(scale_log2, scale_log2);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[CW];
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) tmem_ld_32x32b_x32(tS0 + lane_addr + st * 64 + half * CW + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_ld_wait();

      if ((j * 64 + 63 > m0) || ((j + 1) * 64 > len)) {   // tiles that touch the diagonal or the end of the document
#pragma unroll
        for (int e = 0; e < CW; ++e) {
          const int kv = j * 64 + half * CW + e;
          if (kv > qrow || kv >= len) v[e] = 0xff800000u;   // -inf
        }
      }
      // row max of the raw scores (scale > 0: the max commutes with the scaling); four independent FMNMX3 chains
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < CW; e += 2) mxa[(e >> 1) & 3] = fmax3(mxa[(e >> 1) & 3], __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
      float mx = fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), mxa[3]);
      if (HALVES == 2) {
        // the other warp of this lane quarter holds the other 32 columns of the same rows: exchange through smem (buffer = tile parity,
        // so one named barrier per tile is enough)
        float* buf = xch + st * 256;
        buf[half * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        mx = fmaxf(mx, buf[(half ^ 1) * 128 + r]);
      }
      mx *= scale_log2;
      if (j == 0) {
        m_used = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool grow = mx > m_used + 8.f;  // lazy rescale (P <= 2^8); both warps of a quarter see the same rows, maxima and decision
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
          const float f = grow ? ex2a(m_used - mx) : 1.f;
          const uint64_t f2 = f2_pack(f, f);
#pragma unroll
          for (int c = half * OC; c < half * OC + OC; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tO + lane_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float a, b;
              f2_unpack(f2_mul(f2_pack(__uint_as_float(ov[e]), __uint_as_float(ov[e + 1])), f2), a, b);
              ov[e] = __float_as_uint(a);
              ov[e + 1] = __float_as_uint(b);
            }
            tmem_st_32x32b_x32(tO + lane_addr + c * 32, ov);
          }
          tmem_st_wait();
          l_sum *= f;
          if (grow) m_used = mx;
        }
      }
      const uint64_t neg_m = f2_pack(-m_used, -m_used);
      uint32_t pk[CW / 2];
      uint64_t sum2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
      for (int e = 0; e < CW; e += 2) {
        float x0, x1;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), scale2, neg_m), x0, x1);
        const float p0 = ex2a(x0), p1 = ex2a(x1);
        sum2[(e >> 1) & 1] = f2_add(sum2[(e >> 1) & 1], f2_pack(p0, p1));
        pk[e >> 1] = pack_bf16x2(p0, p1);
      }
      float sa, sb;
      f2_unpack(f2_add(sum2[0], sum2[1]), sa, sb);
      l_sum += sa + sb;
      // P_j (bf16 pairs: CW/2 columns per thread) over the first 32 columns of S_j - every thread that reads those columns of this row
      // has done so: with HALVES == 2 the half-0 warp loaded them above and the named barrier ordered that load before this store
      if (j > 0) {
        // No tensor-memory store while PV_{j-1} - an MMA whose A operand is sourced from tensor memory - is in flight, even though P_j
        // goes to other columns: without this hand-shake 4 of 300 launches at the 8B shapes came back with corrupted O rows (lse intact);
        // with it 0 of 6000 (tools/attn_repro.py).  Free in practice: PV_{j-1} finishes ~256 clocks after p_ready(j-1), the softmax of
        // tile j takes ~1000.
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      if constexpr (HALVES == 1) {
        tmem_st_32x32b_x32(tS0 + lane_addr + st * 64, pk);
      } else {
        tmem_st_32x32b_x16(tS0 + lane_addr + st * 64 + half * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[st]);
    }
    if (HALVES == 2) {       // row sums of the two column halves (same m_used and rescale history in both)
      float* buf = xch + 2 * 256;
      buf[half * 128 + r] = l_sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      l_sum += buf[(half ^ 1) * 128 + r];
    }
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
    const bool valid = qrow < len;
    if (valid && half == 0) lse[static_cast<int64_t>(h) * T + s0 + qrow] = (m_used + log2f(l_sum)) * 0.6931471805599453f;
    __nv_bfloat16* orow = o + static_cast<int64_t>(s0 + qrow) * ldo + h * D;
#pragma unroll
    for (int c = half * OC; c < half * OC + OC; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tO + lane_addr + c * 32, ov);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 val;
          val.x = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 0]) * inv, __uint_as_float(ov[q4 * 8 + 1]) * inv);
          val.y = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 2]) * inv, __uint_as_float(ov[q4 * 8 + 3]) * inv);
          val.z = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 4]) * inv, __uint_as_float(ov[q4 * 8 + 5]) * inv);
          val.w = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 6]) * inv, __uint_as_float(ov[q4 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = val;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int D, int HALVES>
static int launch(const void* q, const void* k, const void* v, void* o, float* lse, const int* cu, int nseq, int max_len, int64_t ldq, int64_t ldk,
                  int64_t ldv, int64_t ldo, int Hq, int Hkv, int T, float scale, cudaStream_t st) {
  using L = Smem<D>;
  auto kern = attn_fwd_ts_kernel<D, HALVES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN);
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attn_fwd_ts smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tq, q, T, static_cast<uint64_t>(Hq) * D, ldq, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tk, k, T, static_cast<uint64_t>(Hkv) * D, ldk, 64, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tv, v, T, static_cast<uint64_t>(Hkv) * D, ldv, 64, 64))) return rc;
  dim3 grid((max_len + 127) / 128, Hq, nseq);
  kern<<<grid, 64 + 128 * HALVES, L::DYN, st>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(o), lse, cu, ldo, Hq, Hkv, T, scale * 1.4426950408889634f);
  B200_CHECK_LAUNCH("attn_fwd_ts");
  return 0;
}

}  // namespace fwdts

// halves: 1 = four softmax warps per CTA (one thread per tile row), 2 = eight (two threads per row)
int attn_fwd_ts(const void* q, const void* k, const void* v, void* o, float* lse, const int* cu_seqlens, int nseq, int max_len, int64_t ldq,
                int64_t ldk, int64_t ldv, int64_t ldo, int Hq, int Hkv, int D, int T, float scale, int halves, cudaStream_t st) {
  if (Hq % Hkv) return set_error(B200_ERR_ARG, "attn: Hq %% Hkv != 0");
  if ((ldq | ldk | ldv | ldo) % 8) return set_error(B200_ERR_ARG, "attn: row pitches must be multiples of 8 elements");
  if (scale <= 0.f) return set_error(B200_ERR_ARG, "attn: scale must be positive");
#define B200_FWDTS(DD, HH) return fwdts::launch<DD, HH>(q, k, v, o, lse, cu_seqlens, nseq, max_len, ldq, ldk, ldv, ldo, Hq, Hkv, T, scale, st)
  (void)halves;   // the two-threads-per-row variant (HALVES = 2) failed its parity tests on hardware and is not instantiated
  if (D == 128) B200_FWDTS(128, 1);
  if (D == 64) B200_FWDTS(64, 1);
#undef B200_FWDTS
  return set_error(B200_ERR_UNSUPPORTED, "attn: head_dim %d not in {64,128}", D);
}

}  // namespace b200
