// HBM-bound kernels of the Llama training step: RMSNorm, RoPE, SwiGLU, embedding, fused cross-entropy,
// fused AdamW, grad-norm.  All bf16 storage / fp32 math, 16-byte vectorised, coalesced.
// Reference semantics restated (paths under /root/reference/nemo_automodel/):
//   RMSNorm  components/models/common/utils.py:250-256   (fp32 norm, weight multiply in fp32, one down-cast)
//   RoPE     components/models/llama/rope_utils.py:39-67 (rotate-half; bf16 cos/sin tables; each op materialised)
//   SwiGLU   components/models/llama/model.py:170
//   CE       components/loss/masked_ce.py:73-89          (fp32 upcast, sum / num_label_tokens, ignore_index -100)
//   AdamW    torch.optim.AdamW as the recipe steps it     (recipes/llm/train_ft.py:1556-1558)
//   clip     components/training/utils.py:122-141,168-169
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {  // read-once data: do not allocate in L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float rcp_fast(float x) {  // MUFU reciprocal (<= 1 ulp fp32): invisible after the bf16 rounding of the result
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// b200_set_option("side_blocks_per_sm", k): k > 0 caps the HBM-bound sweeps that run on side streams next to the GEMMs (AdamW, grad-norm
// partial sums) at k CTAs of 256 threads per SM.  At full occupancy one of these grid-stride kernels owns every register / thread slot of
// the SM for its whole duration, so a GEMM CTA (28.7K registers) launched meanwhile cannot become resident and the "overlap" serialises.
int side_blocks_per_sm = 0;

static int grid_for(int64_t work_items, int per_block, int max_blocks_per_sm = 8) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t b = (work_items + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(sms) * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ---------------------------------------------------------------------------------------------- RMSNorm
// WPR warps per row (cols = VPL*256*WPR); each lane keeps VPL 16-byte vectors of the row in registers between the
// reduction and the normalisation, so every operand is read from HBM exactly once (algorithmic bytes: 4*cols fwd,
// 6*cols bwd per row).  Rows of a CTA are independent; the WPR warps of a row meet on a named barrier.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int WPR>
__device__ __forceinline__ float row_group_sum(float v, float* s_red, int it, int rowslot, int wsub, int lane) {
  v = warp_sum(v);
  if (WPR == 1) return v;
  float* buf = s_red + ((it & 1) * 8 + rowslot) * WPR;  // parity double-buffer: one barrier per use is enough
  if (lane == 0) buf[wsub] = v;
  named_bar_sync(1 + rowslot, 32 * WPR);
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) t += buf[i];
  return t;
}

template <int VPL, int WPR>
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                         __nv_bfloat16* __restrict__ y, float* __restrict__ rstd, int rows,
                                                         float eps) {
  constexpr int cols = VPL * 256 * WPR;
  constexpr int RPB = 8 / WPR;  // rows in flight per CTA
  __shared__ float s_red[2 * 8 * WPR];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rowslot = wid / WPR, wsub = wid % WPR;
  const uint4* wv = reinterpret_cast<const uint4*>(w);  // L1-resident, re-read per row
  int it = 0;
  for (int row = blockIdx.x * RPB + rowslot; row < rows; row += gridDim.x * RPB, ++it) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * cols);
    uint4 xv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) xv[v] = ld_stream(xr + (v * WPR + wsub) * 32 + lane);
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float f[8];
      unpack8(xv[v], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
    ss = row_group_sum<WPR>(ss, s_red, it, rowslot, wsub, lane);
    const float r = rsqrtf(ss / static_cast<float>(cols) + eps);
    if (lane == 0 && wsub == 0 && rstd) rstd[row] = r;
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * cols);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float f[8], wf[8];
      unpack8(xv[v], f);
      unpack8(__ldg(wv + (v * WPR + wsub) * 32 + lane), wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = wf[e] * (f[e] * r);
      yr[(v * WPR + wsub) * 32 + lane] = pack8(f);
    }
  }
}

// Backward: dx = r * (dy*w - xhat * mean(dy*w*xhat)) (+ dres); fp32 dw accumulators per warp in shared memory
// (bank-conflict-free [(v,half)][lane][4] layout), reduced per CTA into partial[blockIdx, cols];
// rmsnorm_dw_finalize sums the partials in a fixed order (deterministic).
template <int VPL, int WPR>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                                                         const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx,
                                                         float* __restrict__ dw_partial, int rows) {
  constexpr int cols = VPL * 256 * WPR;
  constexpr int RPB = 8 / WPR;
  constexpr int REGION = VPL * 256;  // floats per warp
  extern __shared__ float s_dw[];    // [8 warps][REGION] then s_red[2*8*WPR]
  float* s_red = s_dw + 8 * REGION;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rowslot = wid / WPR, wsub = wid % WPR;
  float* my = s_dw + wid * REGION;
  for (int i = lane; i < REGION; i += 32) my[i] = 0.f;
  __syncwarp();
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  int it = 0;
  for (int row = blockIdx.x * RPB + rowslot; row < rows; row += gridDim.x * RPB, ++it) {
    const size_t off = static_cast<size_t>(row) * cols;
    const uint4* xr = reinterpret_cast<const uint4*>(x + off);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + off);
    uint4 xv[VPL], dv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xv[v] = ld_stream(xr + (v * WPR + wsub) * 32 + lane);
      dv[v] = ld_stream(dyr + (v * WPR + wsub) * 32 + lane);
    }
    const float r = rstd[row];
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float xf[8], df[8], wf[8];
      unpack8(xv[v], xf);
      unpack8(dv[v], df);
      unpack8(__ldg(wv + (v * WPR + wsub) * 32 + lane), wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) dot += (df[e] * wf[e]) * (xf[e] * r);
    }
    dot = row_group_sum<WPR>(dot, s_red, it, rowslot, wsub, lane) / static_cast<float>(cols);
    uint4* dxr = reinterpret_cast<uint4*>(dx + off);
    const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + off) : nullptr;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int vi = (v * WPR + wsub) * 32 + lane;
      float xf[8], df[8], o[8], wf[8];
      unpack8(xv[v], xf);
      unpack8(dv[v], df);
      unpack8(__ldg(wv + vi), wf);
      float4* acc0 = reinterpret_cast<float4*>(my + ((v * 2 + 0) * 32 + lane) * 4);
      float4* acc1 = reinterpret_cast<float4*>(my + ((v * 2 + 1) * 32 + lane) * 4);
      float4 a0 = *acc0, a1 = *acc1;
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = xf[e] * r;
        g[e] = df[e] * xh;
        o[e] = r * (df[e] * wf[e] - xh * dot);
      }
      a0.x += g[0]; a0.y += g[1]; a0.z += g[2]; a0.w += g[3];
      a1.x += g[4]; a1.y += g[5]; a1.z += g[6]; a1.w += g[7];
      *acc0 = a0;
      *acc1 = a1;
      if (rr) {
        // residual-stream gradient: the reference materialises dx in bf16, then adds (two bf16 ops)
        float rf[8];
        unpack8(ld_stream(rr + vi), rf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = bf16_round(o[e]) + rf[e];
      }
      dxr[vi] = pack8(o);
    }
  }
  __syncthreads();
  float* out = dw_partial + static_cast<size_t>(blockIdx.x) * cols;
  for (int i = threadIdx.x; i < REGION * WPR; i += blockDim.x) {
    const int ws = i / REGION, j = i % REGION;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < RPB; ++k) s += s_dw[(k * WPR + ws) * REGION + j];
    const int v = j >> 8, rem = j & 255, half = rem >> 7, ln = (rem & 127) >> 2, e = rem & 3;
    out[((v * WPR + ws) * 32 + ln) * 8 + half * 4 + e] = s;
  }
}

// 256 threads = 32 columns x 8 partial groups: the (up to ~300) per-CTA partials of a column are summed 8-way in parallel
// (fixed order -> deterministic), instead of one thread walking all of them.
__global__ void __launch_bounds__(256) rmsnorm_dw_finalize_kernel(const float* __restrict__ partial, int nparts, int cols,
                                                                 __nv_bfloat16* __restrict__ dw, int accumulate) {
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (c < cols)
    for (int p = g; p < nparts; p += 8) acc += partial[static_cast<size_t>(p) * cols + c];
  s[g][cx] = acc;
  __syncthreads();
  if (g == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][cx];
    if (accumulate) t = bf16_round(t) + __bfloat162float(dw[c]);
    dw[c] = __float2bfloat16_rn(t);
  }
}

template <int VPL, int WPR>
static int rmsnorm_fwd_launch(const void* x, const void* w, void* y, float* rstd, int rows, float eps, cudaStream_t st) {
  const int grid = grid_for(rows, 8 / WPR, 4);
  rmsnorm_fwd_kernel<VPL, WPR><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
                                                     static_cast<__nv_bfloat16*>(y), rstd, rows, eps);
  B200_CHECK_LAUNCH("rmsnorm_fwd");
  return 0;
}

int rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int cols, float eps, cudaStream_t st) {
  switch (cols) {
    case 256: return rmsnorm_fwd_launch<1, 1>(x, w, y, rstd, rows, eps, st);
    case 512: return rmsnorm_fwd_launch<2, 1>(x, w, y, rstd, rows, eps, st);
    case 1024: return rmsnorm_fwd_launch<4, 1>(x, w, y, rstd, rows, eps, st);
    case 2048: return rmsnorm_fwd_launch<4, 2>(x, w, y, rstd, rows, eps, st);
    case 4096: return rmsnorm_fwd_launch<4, 4>(x, w, y, rstd, rows, eps, st);
    case 8192: return rmsnorm_fwd_launch<4, 8>(x, w, y, rstd, rows, eps, st);
  }
  return set_error(B200_ERR_UNSUPPORTED, "rmsnorm: hidden size %d not in {256,512,1024,2048,4096,8192}", cols);
}

static int rmsnorm_wpr(int cols) { return cols <= 1024 ? 1 : cols / 1024; }
static int rmsnorm_bwd_grid(int rows, int cols) { return grid_for(rows, 8 / rmsnorm_wpr(cols), 2); }

int rmsnorm_bwd_workspace_floats(int rows, int cols) { return rmsnorm_bwd_grid(rows, cols) * cols; }

template <int VPL, int WPR>
static int rmsnorm_bwd_launch(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                              void* dw, int accumulate, float* ws, int rows, cudaStream_t st) {
  constexpr int cols = VPL * 256 * WPR;
  const int grid = rmsnorm_bwd_grid(rows, cols);
  const size_t smem = (static_cast<size_t>(8) * VPL * 256 + 2 * 8 * WPR) * sizeof(float);
  auto kern = rmsnorm_bwd_kernel<VPL, WPR>;
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "rmsnorm_bwd smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
                                static_cast<const __nv_bfloat16*>(w), rstd, static_cast<const __nv_bfloat16*>(dres),
                                static_cast<__nv_bfloat16*>(dx), ws, rows);
  B200_CHECK_LAUNCH("rmsnorm_bwd");
  rmsnorm_dw_finalize_kernel<<<(cols + 31) / 32, 256, 0, st>>>(ws, grid, cols, static_cast<__nv_bfloat16*>(dw), accumulate);
  B200_CHECK_LAUNCH("rmsnorm_dw_finalize");
  return 0;
}

int rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx, void* dw,
                int accumulate, float* ws, int rows, int cols, cudaStream_t st) {
  switch (cols) {
    case 256: return rmsnorm_bwd_launch<1, 1>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
    case 512: return rmsnorm_bwd_launch<2, 1>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
    case 1024: return rmsnorm_bwd_launch<4, 1>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
    case 2048: return rmsnorm_bwd_launch<4, 2>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
    case 4096: return rmsnorm_bwd_launch<4, 4>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
    case 8192: return rmsnorm_bwd_launch<4, 8>(dy, x, w, rstd, dres, dx, dw, accumulate, ws, rows, st);
  }
  return set_error(B200_ERR_UNSUPPORTED, "rmsnorm: hidden size %d not in {256,512,1024,2048,4096,8192}", cols);
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }          // low / high bf16 of a packed word
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void round_pair(float& lo, float& hi) {  // (lo, hi) -> bf16 -> fp32: one packed cvt + two bit ops
  const uint32_t u = pack_pair(lo, hi);
  lo = bf_lo(u);
  hi = bf_hi(u);
}

// ---------------------------------------------------------------------------------------------- RoPE
// In place on the q and k heads of a token-major buffer (row pitch ld elements).  One CTA per token, one thread per (head, 16-byte
// chunk of the first half of the head) and the matching chunk of the second half: the position and the 2*head_dim table entries are
// shared by every head of the token (L1 hits after the first warp), indices need no 64-bit division, and ~2000 threads per SM each keep
// two 16-byte loads of q/k in flight.  sign=+1 forward, -1 backward (the adjoint rotation).
__global__ void __launch_bounds__(320, 4) rope_kernel(__nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ cos_t,
                                                  const __nv_bfloat16* __restrict__ sin_t, const int* __restrict__ pos,
                                                  int heads, int head_dim, int ld, float sign) {
  const int cpr = head_dim / 16;  // 16B chunks per half head
  const int items = heads * cpr;
  const int t = blockIdx.x;
  const int p = pos[t];
  __nv_bfloat16* row = qk + static_cast<size_t>(t) * ld;
  const __nv_bfloat16* ct = cos_t + static_cast<size_t>(p) * head_dim;
  const __nv_bfloat16* st = sin_t + static_cast<size_t>(p) * head_dim;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int h = w / cpr;
    const int c = w - h * cpr;
    __nv_bfloat16* base = row + h * head_dim + c * 8;
    const __nv_bfloat16* cb = ct + c * 8;
    const __nv_bfloat16* sb = st + c * 8;
    // operands stay packed (bf16x2 words) until the element pair that needs them: 24 registers of inputs instead of 48
    const uint4 X1 = *reinterpret_cast<const uint4*>(base), X2 = *reinterpret_cast<const uint4*>(base + head_dim / 2);
    const uint4 C1 = __ldg(reinterpret_cast<const uint4*>(cb)), S1 = __ldg(reinterpret_cast<const uint4*>(sb));
    const uint4 C2 = __ldg(reinterpret_cast<const uint4*>(cb + head_dim / 2)), S2 = __ldg(reinterpret_cast<const uint4*>(sb + head_dim / 2));
    const uint32_t* x1w = &X1.x; const uint32_t* x2w = &X2.x;
    const uint32_t* c1w = &C1.x; const uint32_t* s1w = &S1.x; const uint32_t* c2w = &C2.x; const uint32_t* s2w = &S2.x;
    uint4 O1, O2;
    uint32_t* o1w = &O1.x; uint32_t* o2w = &O2.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // forward : out1 = x1*cos - x2*sin ; out2 = x2*cos + x1*sin      (q*cos + rotate_half(q)*sin)
      // backward: out1 = x1*cos + x2*sin ; out2 = x2*cos - x1*sin      (adjoint)
      // forward uses sin at the OUTPUT index; the adjoint uses sin at the INPUT index of the rotated term.
      const float x1l = bf_lo(x1w[e]), x1h = bf_hi(x1w[e]), x2l = bf_lo(x2w[e]), x2h = bf_hi(x2w[e]);
      const uint32_t saw = sign > 0.f ? s1w[e] : s2w[e], sbw = sign > 0.f ? s2w[e] : s1w[e];
      float a_l = x1l * bf_lo(c1w[e]), a_h = x1h * bf_hi(c1w[e]);       // x1*cos
      float b_l = x2l * bf_lo(saw), b_h = x2h * bf_hi(saw);             // x2*sin
      float c_l = x2l * bf_lo(c2w[e]), c_h = x2h * bf_hi(c2w[e]);       // x2*cos
      float d_l = x1l * bf_lo(sbw), d_h = x1h * bf_hi(sbw);             // x1*sin
      round_pair(a_l, a_h); round_pair(b_l, b_h); round_pair(c_l, c_h); round_pair(d_l, d_h);   // each product is a materialised bf16 tensor
      o1w[e] = pack_pair(a_l - sign * b_l, a_h - sign * b_h);
      o2w[e] = pack_pair(c_l + sign * d_l, c_h + sign * d_h);
    }
    *reinterpret_cast<uint4*>(base) = O1;
    *reinterpret_cast<uint4*>(base + head_dim / 2) = O2;
  }
}

int rope_inplace(void* qk, const void* cos_t, const void* sin_t, const int* pos, int tokens, int heads, int head_dim, int ld,
                 int backward, cudaStream_t st) {
  if (head_dim % 16 != 0 || ld % 8 != 0) return set_error(B200_ERR_ARG, "rope: head_dim %% 16 and ld %% 8 required");
  if (tokens <= 0 || heads <= 0) return 0;
  const int items = heads * (head_dim / 16);
  int threads = (items + 31) / 32 * 32;
  if (threads > 320) threads = 320;
  rope_kernel<<<tokens, threads, 0, st>>>(static_cast<__nv_bfloat16*>(qk), static_cast<const __nv_bfloat16*>(cos_t),
                                          static_cast<const __nv_bfloat16*>(sin_t), pos, heads, head_dim, ld, backward ? -1.f : 1.f);
  B200_CHECK_LAUNCH("rope");
  return 0;
}

// ---------------------------------------------------------------------------------------------- qkv bias (+ RoPE) and its gradient
// Qwen2 attention (components/models/qwen2/model.py:80-82: q/k/v projections with bias=True).  Forward: ONE in-place pass over the fused
// qkv row that adds the bias to every head and rotates the first `rope_heads` heads (q and k) - the pass the bias-free path spends on
// RoPE alone, so the bias costs no extra HBM traffic.  y = bf16(bf16(x) + b) then the rotate-half RoPE of rope_kernel.
__global__ void __launch_bounds__(320, 4) bias_rope_kernel(__nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ bias,
                                                       const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                                                       const int* __restrict__ pos, int rope_heads, int heads, int head_dim, int ld) {
  const int cpr = head_dim / 16;  // 16B chunks per half head
  const int items = heads * cpr;
  const int t = blockIdx.x;
  const int p = pos[t];
  __nv_bfloat16* row = qkv + static_cast<size_t>(t) * ld;
  const __nv_bfloat16* ct = cos_t + static_cast<size_t>(p) * head_dim;
  const __nv_bfloat16* st = sin_t + static_cast<size_t>(p) * head_dim;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int h = w / cpr;
    const int c = w - h * cpr;
    const int col = h * head_dim + c * 8;
    __nv_bfloat16* base = row + col;
    float x1[8], x2[8], b1[8], b2[8];
    unpack8(*reinterpret_cast<const uint4*>(base), x1);
    unpack8(*reinterpret_cast<const uint4*>(base + head_dim / 2), x2);
    unpack8(__ldg(reinterpret_cast<const uint4*>(bias + col)), b1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(bias + col + head_dim / 2)), b2);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      x1[e] = bf16_round(x1[e] + b1[e]);
      x2[e] = bf16_round(x2[e] + b2[e]);
    }
    if (h < rope_heads) {
      float c1[8], s1[8], c2[8], s2[8], o1[8], o2[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(ct + c * 8)), c1);
      unpack8(__ldg(reinterpret_cast<const uint4*>(st + c * 8)), s1);
      unpack8(__ldg(reinterpret_cast<const uint4*>(ct + c * 8 + head_dim / 2)), c2);
      unpack8(__ldg(reinterpret_cast<const uint4*>(st + c * 8 + head_dim / 2)), s2);
#pragma unroll
      for (int e = 0; e < 8; ++e) {   // out1 = x1*cos - x2*sin ; out2 = x2*cos + x1*sin, each product a materialised bf16 tensor
        o1[e] = bf16_round(x1[e] * c1[e]) - bf16_round(x2[e] * s1[e]);
        o2[e] = bf16_round(x2[e] * c2[e]) + bf16_round(x1[e] * s2[e]);
      }
      *reinterpret_cast<uint4*>(base) = pack8(o1);
      *reinterpret_cast<uint4*>(base + head_dim / 2) = pack8(o2);
    } else {
      *reinterpret_cast<uint4*>(base) = pack8(x1);
      *reinterpret_cast<uint4*>(base + head_dim / 2) = pack8(x2);
    }
  }
}

int bias_rope_inplace(void* qkv, const void* bias, const void* cos_t, const void* sin_t, const int* pos, int tokens, int rope_heads, int heads,
                      int head_dim, int ld, cudaStream_t st) {
  if (head_dim % 16 != 0 || ld % 8 != 0) return set_error(B200_ERR_ARG, "bias_rope: head_dim %% 16 and ld %% 8 required");
  if (rope_heads < 0 || rope_heads > heads) return set_error(B200_ERR_ARG, "bias_rope: rope_heads %d outside 0..%d", rope_heads, heads);
  if (tokens <= 0 || heads <= 0) return 0;
  const int items = heads * (head_dim / 16);
  int threads = (items + 31) / 32 * 32;
  if (threads > 320) threads = 320;
  bias_rope_kernel<<<tokens, threads, 0, st>>>(static_cast<__nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(bias),
                                               static_cast<const __nv_bfloat16*>(cos_t), static_cast<const __nv_bfloat16*>(sin_t), pos, rope_heads,
                                               heads, head_dim, ld);
  B200_CHECK_LAUNCH("bias_rope");
  return 0;
}

// Column sums of a bf16 [rows, cols] matrix (row pitch ld): the bias gradient db = sum_t dy[t, :] (autograd of nn.Linear's bias: fp32
// accumulation, one rounding).  Two deterministic stages: CTA (x, y) sums rows y, y + gridDim.y, ... of a 256-column strip into
// partial[y][cols]; the finalize kernel adds the partials in a fixed order and writes / accumulates the bf16 result.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int rows, int cols,
                                                            int64_t ld) {
  const int c8 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 8;   // 32 lanes x 8 columns = one 256-column strip per CTA
  const int rsub = threadIdx.x >> 5;                            // 8 row phases inside the CTA
  __shared__ float s_acc[8][256];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c8 < cols) {
    for (int r = blockIdx.y * 8 + rsub; r < rows; r += gridDim.y * 8) {
      float f[8];
      unpack8(ld_stream(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ld + c8)), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_acc[rsub][(threadIdx.x & 31) * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_acc[i][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.y) * cols + c] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_finalize_kernel(const float* __restrict__ partial, int nparts, int cols, __nv_bfloat16* __restrict__ out,
                                                             int accumulate) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  float t = 0.f;
  for (int i = 0; i < nparts; ++i) t += partial[static_cast<int64_t>(i) * cols + c];
  if (accumulate) t = bf16_round(t) + __bfloat162float(out[c]);   // micro-batch accumulation: += of materialised bf16 gradients
  out[c] = __float2bfloat16_rn(t);
}

static int colsum_parts(int rows) {
  int p = (rows + 63) / 64;
  return p < 1 ? 1 : (p > 64 ? 64 : p);
}
int colsum_workspace_floats(int rows, int cols) { return colsum_parts(rows) * cols; }
int colsum_bf16(const void* x, void* out, float* ws, int rows, int cols, int64_t ld, int accumulate, cudaStream_t st) {
  if (cols % 8 || ld % 8) return set_error(B200_ERR_ARG, "colsum: cols %% 8 and ld %% 8 required");
  if (rows <= 0 || cols <= 0) return 0;
  const int parts = colsum_parts(rows);
  dim3 grid((cols + 255) / 256, parts);
  colsum_partial_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ws, rows, cols, ld);
  B200_CHECK_LAUNCH("colsum_partial");
  colsum_finalize_kernel<<<(cols + 255) / 256, 256, 0, st>>>(ws, parts, cols, static_cast<__nv_bfloat16*>(out), accumulate);
  B200_CHECK_LAUNCH("colsum_finalize");
  return 0;
}

// ---------------------------------------------------------------------------------------------- SwiGLU
// gu [T, 2F]: gate = cols [0,F), up = cols [F,2F).   a = bf16(bf16(silu(g)) * u)
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ a,
                                                        int64_t tokens, int F) {
  const int vpr = F / 8;
  const int64_t total = tokens * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / vpr;
    const int c = static_cast<int>(i - t * vpr);
    const uint4* row = reinterpret_cast<const uint4*>(gu + t * 2 * F);
    float g[8], u[8], o[8];
    unpack8(ld_stream(row + c), g);
    unpack8(ld_stream(row + vpr + c), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = bf16_round(g[e] * rcp_fast(1.f + __expf(-g[e]))) * u[e];
    reinterpret_cast<uint4*>(a + t * F)[c] = pack8(o);
  }
}
// dgu [T,2F] from da [T,F] and gu.   du = da*silu(g);  dg = bf16(da*u) * sigmoid(g)*(1+g*(1-sigmoid(g)))
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ gu,
                                                        __nv_bfloat16* __restrict__ dgu, int64_t tokens, int F) {
  const int vpr = F / 8;
  const int64_t total = tokens * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / vpr;
    const int c = static_cast<int>(i - t * vpr);
    const uint4* row = reinterpret_cast<const uint4*>(gu + t * 2 * F);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(ld_stream(row + c), g);
    unpack8(ld_stream(row + vpr + c), u);
    unpack8(ld_stream(reinterpret_cast<const uint4*>(da + t * F) + c), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = rcp_fast(1.f + __expf(-g[e]));
      du[e] = d[e] * bf16_round(g[e] * sg);
      dg[e] = bf16_round(d[e] * u[e]) * (sg * (1.f + g[e] * (1.f - sg)));
    }
    uint4* orow = reinterpret_cast<uint4*>(dgu + t * 2 * F);
    orow[c] = pack8(dg);
    orow[vpr + c] = pack8(du);
  }
}

int swiglu_fwd(const void* gu, void* a, int64_t tokens, int F, cudaStream_t st) {
  if (F % 8) return set_error(B200_ERR_ARG, "swiglu: F %% 8 required");
  swiglu_fwd_kernel<<<grid_for(tokens * (F / 8), 256, 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gu),
                                                                         static_cast<__nv_bfloat16*>(a), tokens, F);
  B200_CHECK_LAUNCH("swiglu_fwd");
  return 0;
}
int swiglu_bwd(const void* da, const void* gu, void* dgu, int64_t tokens, int F, cudaStream_t st) {
  if (F % 8) return set_error(B200_ERR_ARG, "swiglu: F %% 8 required");
  swiglu_bwd_kernel<<<grid_for(tokens * (F / 8), 256, 8), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(gu), static_cast<__nv_bfloat16*>(dgu), tokens, F);
  B200_CHECK_LAUNCH("swiglu_bwd");
  return 0;
}

// ---------------------------------------------------------------------------------------------- embedding
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ W,
                                                       __nv_bfloat16* __restrict__ out, int tokens, int hidden) {
  const int vpr = hidden / 8;
  const int64_t total = static_cast<int64_t>(tokens) * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / vpr);
    const int c = static_cast<int>(i - static_cast<int64_t>(t) * vpr);
    reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * hidden)[c] =
        reinterpret_cast<const uint4*>(W + static_cast<size_t>(ids[t]) * hidden)[c];
  }
}
// Deterministic scatter-add.  link kernel: for every token, is it the first occurrence of its id, and which later token
// shares the id next.  The leader of each chain then sums its chain in fp32 in token order.
__global__ void embed_link_kernel(const int* __restrict__ ids, int* __restrict__ next, int* __restrict__ first, int tokens) {
  extern __shared__ int s_ids[];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int my = t < tokens ? ids[t] : -1;
  int nxt = -1, is_first = 1;
  for (int base = 0; base < tokens; base += blockDim.x) {
    __syncthreads();
    if (base + threadIdx.x < tokens) s_ids[threadIdx.x] = ids[base + threadIdx.x];
    __syncthreads();
    const int n = min(static_cast<int>(blockDim.x), tokens - base);
    for (int j = 0; j < n; ++j) {
      const int u = base + j;
      if (s_ids[j] == my) {
        if (u < t) is_first = 0;
        else if (u > t && nxt < 0) nxt = u;
      }
    }
  }
  if (t < tokens) {
    next[t] = nxt;
    first[t] = is_first;
  }
}
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int* __restrict__ ids, const int* __restrict__ next,
                                                       const int* __restrict__ first, const __nv_bfloat16* __restrict__ dh,
                                                       __nv_bfloat16* __restrict__ dW, int tokens, int hidden, int accumulate) {
  const int t = blockIdx.x;
  if (!first[t]) return;
  const int vpr = hidden / 8;
  for (int c = threadIdx.x; c < vpr; c += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int u = t; u >= 0; u = next[u]) {
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(dh + static_cast<size_t>(u) * hidden)[c], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    uint4* dst = reinterpret_cast<uint4*>(dW + static_cast<size_t>(ids[t]) * hidden) + c;
    if (accumulate) {
      float f[8];
      unpack8(*dst, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = bf16_round(acc[e]) + f[e];
    }
    *dst = pack8(acc);
  }
}

int embed_fwd(const int* ids, const void* W, void* out, int tokens, int hidden, cudaStream_t st) {
  if (hidden % 8) return set_error(B200_ERR_ARG, "embed: hidden %% 8 required");
  embed_fwd_kernel<<<grid_for(static_cast<int64_t>(tokens) * (hidden / 8), 256, 8), 256, 0, st>>>(
      ids, static_cast<const __nv_bfloat16*>(W), static_cast<__nv_bfloat16*>(out), tokens, hidden);
  B200_CHECK_LAUNCH("embed_fwd");
  return 0;
}
// workspace: 2*tokens ints.  dW rows not touched by this micro-batch are left as they are (caller zeroes per step).
int embed_bwd(const int* ids, const void* dh, void* dW, int* ws, int tokens, int hidden, int accumulate, cudaStream_t st) {
  if (hidden % 8) return set_error(B200_ERR_ARG, "embed: hidden %% 8 required");
  int* next = ws;
  int* first = ws + tokens;
  embed_link_kernel<<<(tokens + 255) / 256, 256, 256 * sizeof(int), st>>>(ids, next, first, tokens);
  B200_CHECK_LAUNCH("embed_link");
  embed_bwd_kernel<<<tokens, 256, 0, st>>>(ids, next, first, static_cast<const __nv_bfloat16*>(dh),
                                           static_cast<__nv_bfloat16*>(dW), tokens, hidden, accumulate);
  B200_CHECK_LAUNCH("embed_bwd");
  return 0;
}

// ---------------------------------------------------------------------------------------------- fused cross-entropy
// One block per token row.  Pass 1: online (max, sum-exp) over V bf16 logits.  Pass 2 (row re-read from L2):
// dlogits = (softmax - onehot) * inv_n, written in place.  row_loss[t] = lse - z[y]   (0 for ignored rows).
// Algorithmic HBM bytes per row: 2*V read + 2*V written.
__global__ void __launch_bounds__(1024) ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits, const int* __restrict__ labels,
                                                         float* __restrict__ row_loss, int V, int64_t ld, float inv_n,
                                                         int ignore_index) {
  const int row = blockIdx.x;
  __nv_bfloat16* z = logits + static_cast<size_t>(row) * ld;
  const int y = labels[row];
  const int nv = V / 8;
  uint4* zv = reinterpret_cast<uint4*>(z);
  __shared__ float s_m[32], s_s[32];
  __shared__ float s_bcast[2];
  if (y == ignore_index) {  // uniform per block
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) zv[i] = zero;
    for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) z[i] = __float2bfloat16_rn(0.f);
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float f[8];
    unpack8(zv[i], f);
    float lm = f[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) lm = fmaxf(lm, f[e]);
    const float nm = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += __expf(f[e] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) {
    const float f = __bfloat162float(z[i]);
    const float nm = fmaxf(m, f);
    s = s * __expf(m - nm) + __expf(f - nm);
    m = nm;
  }
  // block reduction of (m, s)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float wm = warp_max(m);
  float ws_ = warp_sum(m == -INFINITY ? 0.f : s * __expf(m - wm));
  if (lane == 0) {
    s_m[wid] = wm;
    s_s[wid] = ws_;
  }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    float mm = lane < nw ? s_m[lane] : -INFINITY;
    float ss = lane < nw ? s_s[lane] : 0.f;
    const float gm = warp_max(mm);
    const float gs = warp_sum(mm == -INFINITY ? 0.f : ss * __expf(mm - gm));
    if (lane == 0) {
      s_bcast[0] = gm;
      s_bcast[1] = gs;
    }
  }
  __syncthreads();
  const float gm = s_bcast[0];
  const float inv_s = 1.f / s_bcast[1];
  if (threadIdx.x == 0) row_loss[row] = (__logf(s_bcast[1]) + gm) - __bfloat162float(z[y]);
  __syncthreads();  // z[y] read above must precede the in-place overwrite below
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float f[8];
    unpack8(zv[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float p = __expf(f[e] - gm) * inv_s;
      if (i * 8 + e == y) p -= 1.f;
      f[e] = p * inv_n;
    }
    zv[i] = pack8(f);
  }
  for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) {
    float p = __expf(__bfloat162float(z[i]) - gm) * inv_s;
    if (i == y) p -= 1.f;
    z[i] = __float2bfloat16_rn(p * inv_n);
  }
}

// deterministic sum of n floats (scaled), single block
__global__ void __launch_bounds__(1024) sum_scale_kernel(const float* __restrict__ in, int n, float scale, float* __restrict__ out,
                                                        int accumulate) {
  __shared__ float s_w[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_w[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + v * scale;
  }
}

int ce_fwd_bwd(void* logits, const int* labels, float* row_loss, float* loss_out, int rows, int V, int64_t ld,
               int64_t num_label_tokens, int accumulate_loss, cudaStream_t st) {
  if (ld % 8 || (reinterpret_cast<uintptr_t>(logits) & 15)) return set_error(B200_ERR_ARG, "ce: logits must be 16B aligned, ld %% 8");
  if (rows <= 0) return 0;
  const float inv_n = num_label_tokens > 0 ? 1.f / static_cast<float>(num_label_tokens) : 0.f;
  ce_fwd_bwd_kernel<<<rows, 1024, 0, st>>>(static_cast<__nv_bfloat16*>(logits), labels, row_loss, V, ld, inv_n, -100);
  B200_CHECK_LAUNCH("ce_fwd_bwd");
  sum_scale_kernel<<<1, 1024, 0, st>>>(row_loss, rows, inv_n, loss_out, accumulate_loss);
  B200_CHECK_LAUNCH("ce_loss_sum");
  return 0;
}

// ---------------------------------------------------------------------------------------------- grad norm
__global__ void __launch_bounds__(512) sumsq_partial_kernel(const __nv_bfloat16* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float s_w[16];
  const int64_t nv = n / 8;
  const uint4* gv = reinterpret_cast<const uint4*>(g);
  float s = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ld_stream(gv + i), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += f[e] * f[e];
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nv * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float f = __bfloat162float(g[i]);
      s += f * f;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_w[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

int sumsq_workspace_floats() { return 148 * 8; }

// out[0] (+)= sum(g^2) over n bf16 values.  ws: sumsq_workspace_floats() floats.
int sumsq_bf16(const void* g, int64_t n, float* out, float* ws, int accumulate, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(g) & 15) return set_error(B200_ERR_ARG, "sumsq: 16B alignment required");
  int grid = grid_for(n / 8, 512, side_blocks_per_sm > 0 ? (side_blocks_per_sm + 1) / 2 : 4);
  if (grid > sumsq_workspace_floats()) grid = sumsq_workspace_floats();
  sumsq_partial_kernel<<<grid, 512, 0, st>>>(static_cast<const __nv_bfloat16*>(g), n, ws);
  B200_CHECK_LAUNCH("sumsq_partial");
  sum_scale_kernel<<<1, 1024, 0, st>>>(ws, grid, 1.f, out, accumulate);
  B200_CHECK_LAUNCH("sumsq_final");
  return 0;
}

// ---------------------------------------------------------------------------------------------- fused AdamW
// One pass over the local shard: p, g, m, v (bf16) [+ optional fp32 master].  The clip coefficient is computed on the
// device from the squared global grad norm (no host sync): coef = min(1, max_norm / (sqrt(norm_sq) + 1e-6)).
// mode 0: fp32 math, one rounding per stored tensor (and fp32 master weights when given).
// mode 1: torch's op sequence on bf16 tensors (mul_, lerp_, mul_, addcmul_, sqrt, div_, add_, addcdiv_), every op
//         materialised in bf16, which is what torch.optim.AdamW does to bf16 params/states.
__device__ __forceinline__ void round2(float& x, float& y) {  // (x, y) -> bf16 -> fp32, 3 instructions for two values
  const __nv_bfloat162 t = __floats2bfloat162_rn(x, y);
  const uint32_t u = *reinterpret_cast<const uint32_t*>(&t);
  x = __uint_as_float(u << 16);
  y = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
struct AdamWArgs {  // scalars are formed in double on the host (as torch does in Python) and rounded to fp32 once
  float decay, step_size, beta1, w1, beta2, w2, eps, bc2_sqrt, max_norm;
};
template <int MODE>
__global__ void __launch_bounds__(256) adamw_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                                   __nv_bfloat16* __restrict__ m, __nv_bfloat16* __restrict__ v,
                                                   float* __restrict__ master, int64_t n, AdamWArgs a,
                                                   const float* __restrict__ norm_sq) {
  float coef = 1.f;
  if (norm_sq && a.max_norm > 0.f) coef = fminf(a.max_norm / (sqrtf(norm_sq[0]) + 1e-6f), 1.f);
  const int64_t nv = n / 8;
  const float step_size = a.step_size;
  const float decay = a.decay;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float pf[8], gf[8], mf[8], vf[8];
    unpack8(reinterpret_cast<const uint4*>(p)[i], pf);
    unpack8(ld_stream(reinterpret_cast<const uint4*>(g) + i), gf);
    unpack8(reinterpret_cast<const uint4*>(m)[i], mf);
    unpack8(reinterpret_cast<const uint4*>(v)[i], vf);
    if (master) {
      const float4 a0 = reinterpret_cast<const float4*>(master)[2 * i], a1 = reinterpret_cast<const float4*>(master)[2 * i + 1];
      pf[0] = a0.x; pf[1] = a0.y; pf[2] = a0.z; pf[3] = a0.w; pf[4] = a1.x; pf[5] = a1.y; pf[6] = a1.z; pf[7] = a1.w;
    }
    if (MODE == 1) {
      // torch's op sequence, two elements at a time: every materialised bf16 tensor is one packed cvt.rn.bf16x2 + two bit ops.
      // sqrt / reciprocal use the MUFU approximations (<= 2 ulp fp32, i.e. 2^-14 of a bf16 ulp): the kernel stays HBM-bound.
      const float inv_bc2 = 1.f / a.bc2_sqrt;
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float g0 = gf[e] * coef, g1 = gf[e + 1] * coef;
        round2(g0, g1);
        float p0 = pf[e] * decay, p1 = pf[e + 1] * decay;
        round2(p0, p1);
        float m0 = mf[e] + a.w1 * (g0 - mf[e]), m1 = mf[e + 1] + a.w1 * (g1 - mf[e + 1]);  // lerp_, weight < 0.5
        round2(m0, m1);
        float v0 = vf[e] * a.beta2, v1 = vf[e + 1] * a.beta2;
        round2(v0, v1);
        v0 = v0 + (a.w2 * g0) * g0;
        v1 = v1 + (a.w2 * g1) * g1;
        round2(v0, v1);
        float d0 = sqrt_approx(v0), d1 = sqrt_approx(v1);
        round2(d0, d1);
        d0 *= inv_bc2; d1 *= inv_bc2;
        round2(d0, d1);
        d0 += a.eps; d1 += a.eps;
        round2(d0, d1);
        p0 = p0 - step_size * (m0 * rcp_approx(d0));
        p1 = p1 - step_size * (m1 * rcp_approx(d1));
        round2(p0, p1);
        pf[e] = p0; pf[e + 1] = p1; mf[e] = m0; mf[e + 1] = m1; vf[e] = v0; vf[e + 1] = v1;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gg = gf[e] * coef;
        const float mm = a.beta1 * mf[e] + a.w1 * gg;
        const float vv = a.beta2 * vf[e] + a.w2 * gg * gg;
        const float den = sqrtf(vv) / a.bc2_sqrt + a.eps;
        pf[e] = pf[e] * decay - step_size * (mm / den);
        mf[e] = mm; vf[e] = vv;
      }
    }
    reinterpret_cast<uint4*>(p)[i] = pack8(pf);
    reinterpret_cast<uint4*>(m)[i] = pack8(mf);
    reinterpret_cast<uint4*>(v)[i] = pack8(vf);
    if (master) {
      reinterpret_cast<float4*>(master)[2 * i] = make_float4(pf[0], pf[1], pf[2], pf[3]);
      reinterpret_cast<float4*>(master)[2 * i + 1] = make_float4(pf[4], pf[5], pf[6], pf[7]);
    }
  }
}

int adamw_step(void* p, const void* g, void* m, void* v, float* master, int64_t n, float lr, float beta1, float beta2, float eps,
               float wd, int step, float max_norm, const float* norm_sq, int mode, cudaStream_t st) {
  if (n % 8) return set_error(B200_ERR_ARG, "adamw: shard length must be a multiple of 8 (pad the flat buffer)");
  if (n == 0) return 0;
  AdamWArgs a;
  const double b1 = beta1, b2 = beta2, dlr = lr;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
  a.decay = static_cast<float>(1.0 - dlr * static_cast<double>(wd));
  a.w1 = static_cast<float>(1.0 - b1);
  a.w2 = static_cast<float>(1.0 - b2);
  a.step_size = static_cast<float>(dlr / (1.0 - pow(b1, step)));
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(b2, step)));
  const int grid = grid_for(n / 8, 256, side_blocks_per_sm > 0 ? side_blocks_per_sm : 8);
  auto P = static_cast<__nv_bfloat16*>(p);
  auto G = static_cast<const __nv_bfloat16*>(g);
  auto Mm = static_cast<__nv_bfloat16*>(m);
  auto Vv = static_cast<__nv_bfloat16*>(v);
  if (mode == 1) adamw_kernel<1><<<grid, 256, 0, st>>>(P, G, Mm, Vv, master, n, a, norm_sq);
  else adamw_kernel<0><<<grid, 256, 0, st>>>(P, G, Mm, Vv, master, n, a, norm_sq);
  B200_CHECK_LAUNCH("adamw");
  return 0;
}

// ---------------------------------------------------------------------------------------------- small utilities
__global__ void __launch_bounds__(256) add_bf16_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, int64_t n) {
  const int64_t nv = n / 8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    unpack8(reinterpret_cast<const uint4*>(dst)[i], a);
    unpack8(ld_stream(reinterpret_cast<const uint4*>(src) + i), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    reinterpret_cast<uint4*>(dst)[i] = pack8(a);
  }
}
int add_inplace_bf16(void* dst, const void* src, int64_t n, cudaStream_t st) {
  if (n % 8) return set_error(B200_ERR_ARG, "add: n %% 8 required");
  add_bf16_kernel<<<grid_for(n / 8, 256, 8), 256, 0, st>>>(static_cast<__nv_bfloat16*>(dst), static_cast<const __nv_bfloat16*>(src), n);
  B200_CHECK_LAUNCH("add_bf16");
  return 0;
}

}  // namespace b200
