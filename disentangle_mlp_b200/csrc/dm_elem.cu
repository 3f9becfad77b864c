// HBM-bound kernels of the training step: layout conversion, BatchNorm (training mode) forward/backward
// with folded ReLU / LeakyReLU, bias+activation, reparameterisation, the N=1 discriminator head, loss
// reductions.  All are vectorised (8 channels = 16 B of bf16 per thread per access), channel-innermost
// (NHWC) so a warp touches contiguous memory, and reduce with per-thread accumulators -> shared memory ->
// one atomicAdd per (block, channel).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <mutex>
#include <cstdint>
#include <cstdlib>

#include "../../include/dm_b200.h"
#include "dm_common.h"
#include "dm_bn_fin.cuh"

namespace dm {
extern std::atomic<long long> g_launch_count;

#define DM_LAUNCHED(name)                                   \
  do {                                                      \
    g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    return check_launch(name);                              \
  } while (0)

// ------------------------------------------------------------------------------------------ helpers
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  bf16x8 x = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(x.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  bf16x8 x;
#pragma unroll
  for (int i = 0; i < 4; ++i) x.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(p) = x;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  if (act == 1) return z > 0.f ? z : 0.f;
  if (act == 2) return z > 0.f ? z : z * slope;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return z > 0.f ? 1.f : slope;
  return 1.f;
}

// Thread layout shared by all [rows, C] channel-innermost kernels: TX threads cover TX*8 channels,
// TY = 256/TX rows in flight; grid.x tiles channels, grid.y tiles rows.
struct RowLayout {
  int tx, ty, gx, gy;
  long long rows_per_block;
};
static RowLayout make_row_layout(long long rows, int c, bool streaming = false, int groups = 1) {
  RowLayout l;
  int cv = c / 8;
  l.tx = 1;
  while (l.tx < cv && l.tx < 256) l.tx <<= 1;
  if (l.tx > cv) l.tx >>= 1;  // cv not a power of two: fall back to the largest pow2 below (never for this model)
  if (l.tx < 1) l.tx = 1;
  l.ty = 256 / l.tx;
  l.gx = (cv + l.tx - 1) / l.tx;
  // blocks per SM: bytes in flight vs fixed cost per block.  3 = what the register budget of these kernels lets an SM hold
  // (one full wave).  Reducing kernels and pure streaming kernels (apply) have separate knobs.
  static const int per_sm_red = [] { const char* e = getenv("DM_BN_BLOCKS_PER_SM"); return e ? std::max(1, atoi(e)) : 3; }();
  static const int per_sm_str = [] { const char* e = getenv("DM_BN_APPLY_BLOCKS_PER_SM"); return e ? std::max(1, atoi(e)) : 4; }();
  const int per_sm = streaming ? per_sm_str : per_sm_red;
  // (stacked passes multiply the grid by `groups`: DM_BN_GROUP_AWARE=1 divides the per-group block count accordingly)
  static const bool group_aware = [] { const char* e = getenv("DM_BN_GROUP_AWARE"); return e && e[0] == '1'; }();
  const int gdiv = group_aware ? std::max(1, groups) : 1;
  long long want = std::max<long long>(1, (148ll * per_sm) / (l.gx * gdiv));  // blocks per SM: bytes in flight vs partial vectors
  long long rpb = (rows + want - 1) / want;
  // DM_BN_MIN_SWEEPS > 1 gives small tensors fewer, fatter blocks (fewer partial vectors for the finalize kernels).
  // Measured (batch 64): 1 is best -- the reduce / apply kernels lose more than the finalize kernels gain.
  static const int mult = [] { const char* e = getenv("DM_BN_MIN_SWEEPS"); return e ? std::max(1, atoi(e)) : 1; }();
  rpb = std::max<long long>(rpb, static_cast<long long>(mult) * l.ty);
  rpb = std::max<long long>(l.ty, (rpb + l.ty - 1) / l.ty * l.ty);
  l.rows_per_block = rpb;
  l.gy = static_cast<int>((rows + rpb - 1) / rpb);
  return l;
}

// ------------------------------------------------------------------------------------------ BN statistics
// Row reduction skeleton: every block reduces its row range and writes ONE partial vector
//   out[(blockIdx.y * NACC + a) * c + channel]        (no atomics: the consumer sums the gridDim.y partials)
template <typename T, int NACC, typename F>
__device__ __forceinline__ void rows_reduce(long long rows, int c, long long rows_per_block, float* out, F&& body) {
  extern __shared__ float red[];  // [ty][tx*8*NACC]
  const int tx = blockDim.x, ty = blockDim.y;
  const int cv = blockIdx.x * tx + threadIdx.x;
  const bool live = cv * 8 < c;
  float acc[NACC][8];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  if (live) {
#pragma unroll 4
    for (long long r = r0 + threadIdx.y; r < r1; r += ty) body(r, cv * 8, acc);
  }
  float* mine = red + (threadIdx.y * tx + threadIdx.x) * (8 * NACC);
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) mine[a * 8 + i] = acc[a][i];
  __syncthreads();
  for (int j = threadIdx.y; j < 8 * NACC; j += ty) {
    float s = 0.f;
    for (int y = 0; y < ty; ++y) s += red[(y * tx + threadIdx.x) * (8 * NACC) + j];
    if (live) out[(static_cast<long long>(blockIdx.y) * NACC + (j / 8)) * c + cv * 8 + (j % 8)] = s;
  }
}

// Block-cooperative sum over partial vectors: blockDim = (32, 32); thread (x, y) accumulates partials
// y, y+32, ... of element i = blockIdx.x*32 + x for NV consecutive groups (stride `gstride` elements), then the
// 32 y-lanes are reduced through shared memory.  Result valid in threads with y == 0.
template <int NV>
__device__ __forceinline__ void sum_partials_block(const float* __restrict__ partials, int nparts, long long n,
                                                   int i, bool live, int gstride, float (&out)[NV]) {
  __shared__ float red[NV][32][33];
  float acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.f;
  if (live) {
#pragma unroll 4
    for (int p = threadIdx.y; p < nparts; p += 32) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] += partials[static_cast<long long>(p) * n + v * gstride + i];
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) red[v][threadIdx.y][threadIdx.x] = acc[v];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float s = 0.f;
      for (int y = 0; y < 32; ++y) s += red[v][y][threadIdx.x];
      out[v] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------ BatchNorm (training mode)
// Two launches per application and direction, no finalize kernels (dm_bn_fin.cuh has the scheme): the producer of the
// per-channel sums (the tcgen05 GEMM that writes the pre-BatchNorm tensor, in its epilogue -- dm_gemm.cu; or
// bn_stats_kernel / bn_bwd_reduce_kernel here) adds its partial sums to a slot scratch with fp32 red.add; its LAST
// block finalizes (constants for the consumer, running statistics / dgamma, dbeta) and re-zeroes the scratch; the
// consumers are plain streaming kernels.  Tensors with at most kBn1dMaxRows rows (BatchNorm1d behind the Linear
// layers: rows = batch) take a single-launch kernel with an exact two-pass variance.
constexpr int kBn1dMaxRows = 256;
constexpr int kBnUnroll = 8;  // rows in flight per thread of the streaming BatchNorm kernels (16-byte loads each)

// Row reduction into slots: every block reduces its row range, then adds its NACC x (tx*8) totals to slot
// (blockIdx.y % kBnSlots) of `slots` = [kBnSlots][NACC][c].
template <typename T, int NACC, typename F>
__device__ __forceinline__ void rows_reduce_slots(long long rows, int c, long long rows_per_block, float* slots, F&& body) {
  extern __shared__ float red[];  // [ty][tx*8*NACC]
  const int tx = blockDim.x, ty = blockDim.y;
  const int cv = blockIdx.x * tx + threadIdx.x;
  const bool live = cv * 8 < c;
  float acc[NACC][8];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  if (live) {
#pragma unroll kBnUnroll
    for (long long r = r0 + threadIdx.y; r < r1; r += ty) body(r, cv * 8, acc);
  }
  float* mine = red + (threadIdx.y * tx + threadIdx.x) * (8 * NACC);
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) mine[a * 8 + i] = acc[a][i];
  __syncthreads();
  float* dst = slots + static_cast<long long>(blockIdx.y % kBnSlots) * NACC * c;
  for (int j = threadIdx.y; j < 8 * NACC; j += ty) {
    float s = 0.f;
    for (int y = 0; y < ty; ++y) s += red[(y * tx + threadIdx.x) * (8 * NACC) + j];
    if (live) atomicAdd(dst + static_cast<long long>(j / 8) * c + cv * 8 + (j % 8), s);
  }
}

// Ticket: call when this block's partial sums have been added.  Returns true in every thread of the LAST block of the
// grid to get here: all other blocks' red.adds are then visible (each published them with __threadfence first).
__device__ __forceinline__ bool take_ticket(unsigned int* counter) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int n = gridDim.x * gridDim.y * gridDim.z;
    s_last = (atomicAdd(counter, 1u) == n - 1u) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  return s_last != 0;
}

// Forward producer for tensors no GEMM epilogue covers: shifted sums (k = running_mean) + finalize by the last block.
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ y, int c, long long rows_per_block,
                                                       const dm_bn_fuse f) {
  pdl_sync();
  // blockIdx.z = group: `gridDim.z` independent batches stacked along rows, each with its own slots
  const long long rows = f.rows;
  y += static_cast<long long>(blockIdx.z) * rows * c;
  float* slots = f.scratch + static_cast<long long>(blockIdx.z) * kBnSlots * 2 * c;
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  float k[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (f.running_mean && cv * 8 < c) {  // (plain loads: the last block of this launch rewrites running_mean)
    const float4 u = __ldcg(reinterpret_cast<const float4*>(f.running_mean + cv * 8));
    const float4 w = __ldcg(reinterpret_cast<const float4*>(f.running_mean + cv * 8) + 1);
    k[0] = u.x; k[1] = u.y; k[2] = u.z; k[3] = u.w; k[4] = w.x; k[5] = w.y; k[6] = w.z; k[7] = w.w;
  }
  rows_reduce_slots<T, 2>(rows, c, rows_per_block, slots, [&](long long r, int ch, float(&acc)[2][8]) {
    float v[8];
    load8(y + r * c + ch, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - k[i];
      acc[0][i] += d;
      acc[1][i] += d * d;
    }
  });
  if (!take_ticket(bn_ticket(f.scratch, c, f.groups))) return;
  bn_forward_finalize(f, c, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y);
}

// out = act(y * scale + shift) with given constants; blockIdx.z = group
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_act_kernel(const T* __restrict__ y, long long rows, int c,
                                                           long long rows_per_block,
                                                           const float* __restrict__ scale_shift, int act,
                                                           float slope, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  if (cv * 8 >= c) return;
  y += static_cast<long long>(blockIdx.z) * rows * c;
  out += static_cast<long long>(blockIdx.z) * rows * c;
  scale_shift += static_cast<long long>(blockIdx.z) * 2 * c;
  float sc[8], sh[8];
  load8(scale_shift + cv * 8, sc);
  load8(scale_shift + c + cv * 8, sh);
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
#pragma unroll kBnUnroll
  for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float f[8];
    load8(y + r * c + cv * 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = act_fwd(f[i] * sc[i] + sh[i], act, slope);
    store8(out + r * c + cv * 8, f);
  }
}

// Backward producer: slots[.][0][c] += sum dz, slots[.][1][c] += sum dz * xhat, dz = dout * act'(z); the last block
// writes the per-group sums for bn_bwd_apply_kernel, accumulates dgamma / dbeta and re-zeroes the slots.
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout,
                                                            const T* __restrict__ y, long long rows, int c,
                                                            long long rows_per_block,
                                                            const float* __restrict__ scale_shift,
                                                            const float* __restrict__ mean_invstd, int act,
                                                            float slope, float* scratch, float* dgamma, float* dbeta) {
  pdl_sync();
  const int groups = gridDim.z;
  float* slots = scratch + static_cast<long long>(blockIdx.z) * kBnSlots * 2 * c;
  {
    const long long z = blockIdx.z;  // group
    dout += z * rows * c;
    y += z * rows * c;
    scale_shift += z * 2 * c;
    mean_invstd += z * 2 * c;
  }
  const int cv0 = blockIdx.x * blockDim.x + threadIdx.x;
  float sc[8], sh[8], mu[8], is[8];
  if (cv0 * 8 < c) {
    load8(scale_shift + cv0 * 8, sc);
    load8(scale_shift + c + cv0 * 8, sh);
    load8(mean_invstd + cv0 * 8, mu);
    load8(mean_invstd + c + cv0 * 8, is);
  }
  rows_reduce_slots<T, 2>(rows, c, rows_per_block, slots, [&](long long r, int ch, float(&acc)[2][8]) {
    float f[8], g[8];
    load8(y + r * c + ch, f);
    load8(dout + r * c + ch, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dz = g[i] * act_grad(f[i] * sc[i] + sh[i], act, slope);
      acc[0][i] += dz;
      acc[1][i] += dz * (f[i] - mu[i]) * is[i];
    }
  });
  if (!take_ticket(bn_ticket(scratch, c, groups))) return;
  bn_backward_finalize(scratch, c, groups, dgamma, dbeta, threadIdx.y * blockDim.x + threadIdx.x,
                       blockDim.x * blockDim.y);
}

// Backward consumer: dy = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat)); sums = [groups][2][c]
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout,
                                                           const T* __restrict__ y, long long rows, int c,
                                                           long long rows_per_block,
                                                           const float* __restrict__ scale_shift,
                                                           const float* __restrict__ mean_invstd,
                                                           const float* __restrict__ sums, int act, float slope,
                                                           __nv_bfloat16* __restrict__ dy) {
  pdl_sync();
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  if (cv * 8 >= c) return;
  {
    const long long z = blockIdx.z;  // group
    dout += z * rows * c;
    y += z * rows * c;
    dy += z * rows * c;
    scale_shift += z * 2 * c;
    mean_invstd += z * 2 * c;
    sums += z * 2 * c;
  }
  float sc[8], sh[8], mu[8], is[8], s0[8], s1[8];
  load8(scale_shift + cv * 8, sc);
  load8(scale_shift + c + cv * 8, sh);
  load8(mean_invstd + cv * 8, mu);
  load8(mean_invstd + c + cv * 8, is);
  load8(sums + cv * 8, s0);
  load8(sums + c + cv * 8, s1);
  const float inv_n = 1.f / static_cast<float>(rows);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s0[i] *= inv_n;
    s1[i] *= inv_n;
  }
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
#pragma unroll kBnUnroll
  for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float f[8], g[8];
    load8(y + r * c + cv * 8, f);
    load8(dout + r * c + cv * 8, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dz = g[i] * act_grad(f[i] * sc[i] + sh[i], act, slope);
      const float xh = (f[i] - mu[i]) * is[i];
      g[i] = sc[i] * (dz - s0[i] - xh * s1[i]);  // sc = gamma * invstd
    }
    store8(dy + r * c + cv * 8, g);
  }
}

// Backward in ONE launch: reduce -> grid-wide barrier -> apply.  The barrier is the ticket of the two-launch scheme
// plus a flag the finalizing block raises; every block of the grid must be resident at once, so the host caps the
// grid at two blocks per SM (dm_bn_backward falls back to the two-launch form otherwise).  Waiting blocks only wait for
// blocks of this same kernel, which depend on nothing but earlier kernels: other kernels sharing the SMs can delay the
// barrier but not deadlock it.  The apply phase re-reads this block's own rows (L2 hits for the tensors of this model)
// -- no second launch, no second ramp-up.  Scratch words after the slots: [0] ticket, [1] flag, [2] exit ticket.
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dout,
                                                           const T* __restrict__ y, long long rows, int c,
                                                           long long rows_per_block,
                                                           const float* __restrict__ scale_shift,
                                                           const float* __restrict__ mean_invstd, int act,
                                                           float slope, float* scratch, float* dgamma, float* dbeta,
                                                           __nv_bfloat16* __restrict__ dy) {
  pdl_sync();
  const int groups = gridDim.z;
  float* slots = scratch + static_cast<long long>(blockIdx.z) * kBnSlots * 2 * c;
  const float* sums = scratch + bn_slot_floats(c, groups) + 4 + static_cast<long long>(blockIdx.z) * 2 * c;
  {
    const long long z = blockIdx.z;  // group
    dout += z * rows * c;
    y += z * rows * c;
    dy += z * rows * c;
    scale_shift += z * 2 * c;
    mean_invstd += z * 2 * c;
  }
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = cv * 8 < c;
  float sc[8], sh[8], mu[8], is[8];
  if (live) {
    load8(scale_shift + cv * 8, sc);
    load8(scale_shift + c + cv * 8, sh);
    load8(mean_invstd + cv * 8, mu);
    load8(mean_invstd + c + cv * 8, is);
  }
  rows_reduce_slots<T, 2>(rows, c, rows_per_block, slots, [&](long long r, int ch, float(&acc)[2][8]) {
    float f[8], g[8];
    load8(y + r * c + ch, f);
    load8(dout + r * c + ch, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dz = g[i] * act_grad(f[i] * sc[i] + sh[i], act, slope);
      acc[0][i] += dz;
      acc[1][i] += dz * (f[i] - mu[i]) * is[i];
    }
  });
  unsigned int* words = bn_ticket(scratch, c, groups);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (take_ticket(words)) {
    bn_backward_finalize(scratch, c, groups, dgamma, dbeta, tid, blockDim.x * blockDim.y);
    __threadfence();  // every finalizing thread publishes its sums ...
    __syncthreads();
    if (tid == 0) atomicExch(words + 1, 1u);  // ... before the flag goes up
  }
  if (tid == 0) {
    unsigned int f;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(words + 1) : "memory");
    } while (f == 0u);
  }
  __syncthreads();
  if (live) {
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0[i] = __ldcg(sums + cv * 8 + i);
      s1[i] = __ldcg(sums + c + cv * 8 + i);
    }
    const float inv_n = 1.f / static_cast<float>(rows);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0[i] *= inv_n;
      s1[i] *= inv_n;
    }
    const long long r0 = blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
#pragma unroll kBnUnroll
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      float f[8], g[8];
      load8(y + r * c + cv * 8, f);
      load8(dout + r * c + cv * 8, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dz = g[i] * act_grad(f[i] * sc[i] + sh[i], act, slope);
        const float xh = (f[i] - mu[i]) * is[i];
        g[i] = sc[i] * (dz - s0[i] - xh * s1[i]);  // sc = gamma * invstd
      }
      store8(dy + r * c + cv * 8, g);
    }
  }
  // the last block past the barrier lowers the flag for the next launch
  __syncthreads();
  if (tid == 0) {
    const unsigned int n = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(words + 2, 1u) == n - 1u) {
      words[2] = 0u;
      __threadfence();
      atomicExch(words + 1, 0u);
    }
  }
}

// ---- small-row BatchNorm (rows <= kBn1dMaxRows; BatchNorm1d behind the Linear layers, model.py:462,468,492):
// ONE launch, one block = 32 channels x all rows, exact two-pass variance (what torch computes), running statistics,
// normalise + activation.  blockDim (4, 64): thread (x, y) holds rows y, y+64, y+128, y+192 of channels 8x .. 8x+7.
// Groups (stacked passes) are processed in order by the same block.
__device__ __forceinline__ void bn1d_block_sum(float (&v)[8], float (*red)[33], float* stat) {
  // sum over the 64 y-threads of each of the block's 32 channels; result broadcast through stat[32]
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x * 8 + i] = v[i];
  __syncthreads();
  const int tid = threadIdx.y * 4 + threadIdx.x;
  if (tid < 32) {
    float s = 0.f;
#pragma unroll 8
    for (int yy = 0; yy < 64; ++yy) s += red[yy][tid];
    stat[tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = stat[threadIdx.x * 8 + i];
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) bn1d_fwd_kernel(const T* __restrict__ y, int rows, int c, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* running_mean,
                                                       float* running_var, long long* num_batches_tracked, float momentum,
                                                       float eps, int act, float slope, float* __restrict__ scale_shift,
                                                       float* __restrict__ mean_invstd, __nv_bfloat16* __restrict__ out,
                                                       int groups, long long ld_y, const float* __restrict__ pre_bias) {
  // ld_y: row stride of y (>= c: y may be a column block of a wider matrix, e.g. one head of a fused two-head GEMM);
  // pre_bias: the Linear bias the producing GEMM did NOT add: normalised as if y + pre_bias had been stored, while the
  // saved constants (mean, shift) are expressed for the stored y, which is what the backward pass reads
  pdl_sync();
  __shared__ float red[64][33];
  __shared__ float stat[32];
  const int ch = (blockIdx.x * 4 + threadIdx.x) * 8;  // c is a multiple of 32: every thread is live
  float pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pre_bias) load8(pre_bias + ch, pb);
  for (int g = 0; g < groups; ++g) {
    const T* yg = y + static_cast<long long>(g) * rows * ld_y;
    __nv_bfloat16* og = out + static_cast<long long>(g) * rows * c;
    float v[4][8];
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = threadIdx.y + 64 * j;
      if (r < rows) {
        load8(yg + static_cast<long long>(r) * ld_y + ch, v[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += v[j][i];
      }
    }
    bn1d_block_sum(s, red, stat);
    float mean[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = s[i] / static_cast<float>(rows);
      q[i] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (threadIdx.y + 64 * j < rows) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[j][i] - mean[i];
          q[i] += d * d;
        }
      }
    bn1d_block_sum(q, red, stat);
    float sc[8], sh[8], is[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float var = q[i] / static_cast<float>(rows);
      is[i] = rsqrtf(var + eps);
      sc[i] = gamma[ch + i] * is[i];
      sh[i] = beta[ch + i] - mean[i] * sc[i];  // (mean of the STORED y: the bias cancels in the normalisation)
      if (threadIdx.y == 0 && running_mean) {
        const float unbiased = rows > 1 ? var * static_cast<float>(rows) / static_cast<float>(rows - 1) : var;
        running_mean[ch + i] = (1.f - momentum) * running_mean[ch + i] + momentum * (mean[i] + pb[i]);
        running_var[ch + i] = (1.f - momentum) * running_var[ch + i] + momentum * unbiased;
      }
    }
    if (threadIdx.y == 0) {
      float* ss = scale_shift + static_cast<long long>(g) * 2 * c;
      float* mi = mean_invstd + static_cast<long long>(g) * 2 * c;
      store8(ss + ch, sc);
      store8(ss + c + ch, sh);
      store8(mi + ch, mean);
      store8(mi + c + ch, is);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = threadIdx.y + 64 * j;
      if (r < rows) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = act_fwd(v[j][i] * sc[i] + sh[i], act, slope);
        store8(og + static_cast<long long>(r) * c + ch, f);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && num_batches_tracked) *num_batches_tracked += groups;
}

template <typename T>
__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const T* __restrict__ y,
                                                       int rows, int c, const float* __restrict__ scale_shift,
                                                       const float* __restrict__ mean_invstd, int act, float slope,
                                                       __nv_bfloat16* __restrict__ dy, float* dgamma, float* dbeta,
                                                       int groups, long long ld_y, long long ld_dy) {
  // ld_y / ld_dy: row strides of y and dy (>= c: column blocks of wider matrices); dout is dense [rows, c]
  pdl_sync();
  __shared__ float red[64][33];
  __shared__ float stat[32];
  const int ch = (blockIdx.x * 4 + threadIdx.x) * 8;
  float gsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int g = 0; g < groups; ++g) {
    const long long off = static_cast<long long>(g) * rows * c;
    const float* ss = scale_shift + static_cast<long long>(g) * 2 * c;
    const float* mi = mean_invstd + static_cast<long long>(g) * 2 * c;
    float sc[8], sh[8], mu[8], is[8];
    load8(ss + ch, sc);
    load8(ss + c + ch, sh);
    load8(mi + ch, mu);
    load8(mi + c + ch, is);
    float dz[4][8], xh[4][8];
    float a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = threadIdx.y + 64 * j;
      if (r < rows) {
        float f[8], gg[8];
        load8(y + static_cast<long long>(g) * rows * ld_y + static_cast<long long>(r) * ld_y + ch, f);
        load8(dout + off + static_cast<long long>(r) * c + ch, gg);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dz[j][i] = gg[i] * act_grad(f[i] * sc[i] + sh[i], act, slope);
          xh[j][i] = (f[i] - mu[i]) * is[i];
          a0[i] += dz[j][i];
          a1[i] += dz[j][i] * xh[j][i];
        }
      }
    }
    bn1d_block_sum(a0, red, stat);
    bn1d_block_sum(a1, red, stat);
    const float inv_n = 1.f / static_cast<float>(rows);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gsum[i] += a1[i];
      bsum[i] += a0[i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = threadIdx.y + 64 * j;
      if (r < rows) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sc[i] * (dz[j][i] - a0[i] * inv_n - xh[j][i] * a1[i] * inv_n);
        store8(dy + static_cast<long long>(g) * rows * ld_dy + static_cast<long long>(r) * ld_dy + ch, o);
      }
    }
  }
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (dgamma) dgamma[ch + i] += gsum[i];
      if (dbeta) dbeta[ch + i] += bsum[i];
    }
  }
}

// out[i] (+)= sum_p partials[p][i]
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int nparts, int n,
                                                                int accumulate, float* __restrict__ out) {
  pdl_sync();
  const int i = blockIdx.x * 32 + threadIdx.x;
  float sum[1];
  sum_partials_block<1>(partials, nparts, n, i, i < n, 0, sum);
  if (threadIdx.y != 0 || i >= n) return;
  out[i] = accumulate ? out[i] + sum[0] : sum[0];
}

// ------------------------------------------------------------------------------------------ bias + activation
// out = act(acc + bias); writes fp32 and/or bf16 copies.  Backward: dpre = dout * act'(pre) (+ column sums).
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ acc, long long rows, int c,
                                                       long long rows_per_block, const float* __restrict__ bias,
                                                       int act, float slope, float* __restrict__ out_f32,
                                                       __nv_bfloat16* __restrict__ out_bf16) {
  pdl_sync();
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  if (cv * 8 >= c) return;
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = 0.f;
  if (bias) load8(bias + cv * 8, b);
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
#pragma unroll 4
  for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float f[8];
    load8(acc + r * c + cv * 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = act_fwd(f[i] + b[i], act, slope);
    if (out_f32) store8(out_f32 + r * c + cv * 8, f);
    if (out_bf16) store8(out_bf16 + r * c + cv * 8, f);
  }
}

// dpre[r][c] = dout[r][c] * act'(out[r][c]) (sign of out == sign of pre for relu/leaky); colsum[c] += sum_r dpre
__global__ void __launch_bounds__(256) act_bwd_colsum_kernel(const float* __restrict__ dout,
                                                             const float* __restrict__ out, long long rows, int c,
                                                             long long rows_per_block, int act, float slope,
                                                             __nv_bfloat16* __restrict__ dpre,
                                                             float* __restrict__ colsum) {
  pdl_sync();
  rows_reduce<float, 1>(rows, c, rows_per_block, colsum, [&](long long r, int ch, float(&acc)[1][8]) {
    float g[8], o[8];
    load8(dout + r * c + ch, g);
    load8(out + r * c + ch, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] *= act_grad(o[i], act, slope);
      acc[0][i] += g[i];
    }
    store8(dpre + r * c + ch, g);
  });
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long rows, int c,
                                                     long long rows_per_block, float* __restrict__ colsum) {
  pdl_sync();
  rows_reduce<T, 1>(rows, c, rows_per_block, colsum, [&](long long r, int ch, float(&acc)[1][8]) {
    float f[8];
    load8(x + r * c + ch, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] += f[i];
  });
}

// ------------------------------------------------------------------------------------------ layout kernels
// fp32 NCHW 3-channel image -> bf16 im2col matrix [batch*oh*ow, 80]; column = c*25 + kh*5 + kw (75 valid, 5 zero).
// The GEMMs read it with K boxes of 32/64 columns: columns >= 80 are TMA out-of-bounds zero fill, never stored.
// One block = one output row of one image: the 5 input rows x 3 channels it needs are staged in shared memory
// (zero halo of 2 on each side), then the 160-byte rows of the matrix are written as coalesced 16-byte pieces.
__global__ void __launch_bounds__(256) im2col3_kernel(const float* __restrict__ x, int batch, int h, int w,
                                                      int stride, __nv_bfloat16* __restrict__ col) {
  pdl_sync();
  extern __shared__ float patch[];  // [3 ch][5 kh][w + 4]
  __shared__ int koff[80];          // column k -> offset of its tap in the patch (-1: zero column)
  const int oh = h / stride, ow = w / stride;
  const int pw = w + 4;
  if (threadIdx.x < 80) {
    const int k = threadIdx.x;
    const int ch = k / 25, t = k - ch * 25, kh = t / 5, kw = t - kh * 5;
    koff[k] = (k < 75) ? (ch * 5 + kh) * pw + kw : -1;
  }
  for (long long row = blockIdx.x; row < static_cast<long long>(batch) * oh; row += gridDim.x) {
    const int n = static_cast<int>(row / oh);
    const int y0 = static_cast<int>(row - static_cast<long long>(n) * oh);
    __syncthreads();  // previous row's readers are done (also orders the koff writes the first time)
    for (int i = threadIdx.x; i < 15 * pw; i += blockDim.x) {
      const int r = i / pw, px = i - r * pw;  // r = ch*5 + kh
      const int ch = r / 5, kh = r - ch * 5;
      const int iy = y0 * stride + kh - 2, ix = px - 2;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(x + ((static_cast<long long>(n) * 3 + ch) * h + iy) * w + ix);
      patch[i] = v;
    }
    __syncthreads();
    __nv_bfloat16* dst = col + row * ow * 80;
    for (int item = threadIdx.x; item < ow * 10; item += blockDim.x) {
      const int x0 = item / 10, grp = item - x0 * 10;
      const int base = x0 * stride;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = koff[grp * 8 + i];
        f[i] = (o >= 0) ? patch[o + base] : 0.f;
      }
      store8(dst + item * 8, f);
    }
  }
}

// ---- padded image ("pim"): the 3-channel image side of the three image-facing layers as a TMA-friendly operand.
// bf16 [batch][kPimH = 68][kPimW = 72][4]: pixel (h, w) at [h + 2][w + 2]; channel 3, the 2-pixel border and 4 spare
// pixels per row are zero.  A 5x5 filter ROW of an output pixel is then 5 pixels x 4 channels = 20 contiguous elements.
// TMA strides must be multiples of 16 bytes = TWO pixels, so the implicit GEMM (dm_gemm.cu: dm_conv3_*) reads a window
// of 8 pixels (32 elements, 64 B) starting at an EVEN pixel per output-pixel PAIR (stride 1) or per output pixel
// (stride 2), over a tensor map whose position stride (16 B) is smaller than the box: overlapping windows, no im2col.
constexpr int kPimH = 68, kPimW = 72;

__device__ __forceinline__ void pim_store(__nv_bfloat16* pim, long long n, int h, int w, float r, float g, float b) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(r, g), hi = __floats2bfloat162_rn(b, 0.f);
  uint2 v;
  v.x = *reinterpret_cast<const uint32_t*>(&lo);
  v.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(pim + ((n * kPimH + h + 2) * kPimW + w + 2) * 4) = v;
}

// zero border of one padded image: rows 0,1,66,67 entirely, pixels 0,1 and 66..71 of the other rows (16 B = 2 pixels)
__device__ __forceinline__ void pim_zero_border(__nv_bfloat16* pim, long long n, int tid, int nt) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint4* base = reinterpret_cast<uint4*>(pim + n * kPimH * kPimW * 4);
  constexpr int kRow16 = kPimW / 2;  // 16-byte pieces per row
  for (int i = tid; i < 4 * kRow16; i += nt) {
    const int r = i / kRow16, col = i - r * kRow16;
    base[(r < 2 ? r : 64 + r) * kRow16 + col] = z;
  }
  for (int i = tid; i < 64 * 4; i += nt) {
    const int r = i >> 2, j = i & 3;
    base[(r + 2) * kRow16 + (j == 0 ? 0 : 32 + j)] = z;  // pixels 0-1, 66-67, 68-69, 70-71
  }
}

// image -> pim.  src: fp32 NCHW [b,3,64,64] in [-1,1] (src_u8 == 0), or uint8 NHWC [b,64,64,3] (src_u8 != 0): the
// reference's input pipeline ToTensor() + Normalize(.5,.5) (dataloader/dataset.py:37-43) = (u/255 - .5)/.5, fused here;
// then dst_nchw (may be NULL) also receives the normalised fp32 NCHW image the losses read.  One block per image row.
__global__ void __launch_bounds__(256) pad_image3_kernel(const void* __restrict__ src, int src_u8, int batch,
                                                         __nv_bfloat16* __restrict__ pim, float* __restrict__ dst_nchw) {
  pdl_sync();
  for (long long row = blockIdx.x; row < static_cast<long long>(batch) * 64; row += gridDim.x) {
    const long long n = row >> 6;
    const int h = static_cast<int>(row & 63);
    if (h == 0) pim_zero_border(pim, n, threadIdx.x, blockDim.x);
    if (threadIdx.x < 64) {
      const int w = threadIdx.x;
      float v[3];
      if (src_u8) {
        const unsigned char* u = static_cast<const unsigned char*>(src) + ((n * 64 + h) * 64 + w) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = (static_cast<float>(u[ch]) / 255.f - 0.5f) / 0.5f;
        if (dst_nchw) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) dst_nchw[((n * 3 + ch) * 64 + h) * 64 + w] = v[ch];
        }
      } else {
        const float* f = static_cast<const float*>(src);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = f[((n * 3 + ch) * 64 + h) * 64 + w];
      }
      pim_store(pim, n, h, w, v[0], v[1], v[2]);
    }
  }
}

// fp32 NHWC(3) -> fp32 NCHW, optionally through tanh (decoder output, models/model.py:509,565); pim (may be NULL, hw = 64*64
// only) also receives the result as a padded bf16 image: the discriminator's input operand
__global__ void __launch_bounds__(256) nhwc3_to_nchw_kernel(const float* __restrict__ src, long long batch, int hw,
                                                            int apply_tanh, float* __restrict__ dst,
                                                            __nv_bfloat16* __restrict__ pim) {
  pdl_sync();
  const long long total = batch * hw;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / hw;
    const int q = static_cast<int>(p - n * hw);
    float v[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      v[ch] = src[p * 3 + ch];
      if (apply_tanh) v[ch] = tanhf(v[ch]);
      dst[(n * 3 + ch) * hw + q] = v[ch];
    }
    if (pim) pim_store(pim, n, q >> 6, q & 63, v[0], v[1], v[2]);
  }
  if (pim) {  // borders: one block per image, strided
    for (long long n = blockIdx.x; n < batch; n += gridDim.x) pim_zero_border(pim, n, threadIdx.x, blockDim.x);
  }
}

// dy = dout * (1 - out^2), fp32 NCHW in; bias_grad[c] += sum over batch and pixels of dy.  dy goes to fp32 NCHW (dy, may
// be NULL) and / or to a padded bf16 image (pim, may be NULL): the operand of deconv4's input- and weight-gradient GEMMs.
// One thread per pixel (3 channels).
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                       long long batch, int hw, float* __restrict__ dy,
                                                       float* __restrict__ bias_grad, __nv_bfloat16* __restrict__ pim) {
  pdl_sync();
  __shared__ float red[3][8];
  float acc[3] = {0.f, 0.f, 0.f};
  const long long total = batch * hw;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / hw;
    const int q = static_cast<int>(p - n * hw);
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const long long i = (n * 3 + ch) * hw + q;
      const float o = out[i];
      g[ch] = dout[i] * (1.f - o * o);
      if (dy) dy[i] = g[ch];
      acc[ch] += g[ch];
    }
    if (pim) pim_store(pim, n, q >> 6, q & 63, g[0], g[1], g[2]);
  }
  if (pim) {
    for (long long n = blockIdx.x; n < batch; n += gridDim.x) pim_zero_border(pim, n, threadIdx.x, blockDim.x);
  }
  if (bias_grad == nullptr) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = acc[ch];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[ch][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[threadIdx.x][wv];
    atomicAdd(bias_grad + threadIdx.x, v);
  }
}

// fp32 weight [cs][3][5][5] -> bf16 window pack: the K-major B operand matching a 32-element pim window (8 pixels x 4
// channels, element j*4 + c).  stride 2 (window starts at the output pixel's own first tap): w_win[kh][cs][32], kw = j.
// stride 1 (one window per output-pixel PAIR, starting at the even pixel): w_win[kh][pw*cs + n][32], kw = j - pw.
__global__ void __launch_bounds__(256) pack_win_kernel(const float* __restrict__ w, int cs, int stride,
                                                       __nv_bfloat16* __restrict__ w_win) {
  pdl_sync();
  const int nn = stride == 1 ? 2 * cs : cs;
  const int total = 5 * nn * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 31, row = (i >> 5) % nn, kh = i / (32 * nn);
    const int pw = row / cs, n = row - pw * cs;
    const int j = e >> 2, c = e & 3, kw = j - pw;
    float v = 0.f;
    if (kw >= 0 && kw < 5 && c < 3) v = w[((n * 3 + c) * 5 + kh) * 5 + kw];
    w_win[i] = __float2bfloat16_rn(v);
  }
}

// window-layout weight gradient scratch [kh][nn][64] (fp32; nn = cs, or 2*cs = (pw, n) for stride 1; element j*4 + c of
// a 16-pixel window) -> dw[cs][3][5][5] += ; the scratch is re-zeroed
__global__ void __launch_bounds__(256) unpack_win_grad_kernel(float* __restrict__ scratch, int cs, int stride,
                                                              float* __restrict__ dw) {
  pdl_sync();
  const int nn = stride == 1 ? 2 * cs : cs;
  const int total = 5 * nn * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 63, row = (i >> 6) % nn, kh = i / (64 * nn);
    const int pw = row / cs, n = row - pw * cs;
    const int j = e >> 2, c = e & 3, kw = j - pw;
    const float v = scratch[i];
    scratch[i] = 0.f;
    if (kw >= 0 && kw < 5 && c < 3) atomicAdd(dw + ((n * 3 + c) * 5 + kh) * 5 + kw, v);  // (two pw rows share a dw entry)
  }
}

// bf16 [b][r][c] -> [b][c][r] through a 32x32 (+1 pad) shared-memory tile
__global__ void __launch_bounds__(256) transpose_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols,
                                                        __nv_bfloat16* __restrict__ dst) {
  pdl_sync();
  __shared__ __nv_bfloat16 tile[32][33];
  const long long base = static_cast<long long>(blockIdx.z) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[base + static_cast<long long>(r) * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[base + static_cast<long long>(c) * rows + r] = tile[threadIdx.x][j];
  }
}

// fp32 [cs][cb][25] -> bf16 down [25][cs][cb], up [25][cb_pad][cs] (zero rows for cb >= cb), col [cs][128]
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ w, int cs, int cb, int cb_pad,
                                                        __nv_bfloat16* __restrict__ w_down,
                                                        __nv_bfloat16* __restrict__ w_up,
                                                        __nv_bfloat16* __restrict__ w_col) {
  pdl_sync();
  const long long n_down = 25ll * cs * cb, n_up = 25ll * cb_pad * cs, n_col = w_col ? 128ll * cs : 0;
  const long long total = (w_down ? n_down : 0) + (w_up ? n_up : 0) + n_col;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long j = i;
    if (w_down) {
      if (j < n_down) {
        const int b = static_cast<int>(j % cb), s = static_cast<int>((j / cb) % cs), t = static_cast<int>(j / (static_cast<long long>(cb) * cs));
        w_down[j] = __float2bfloat16_rn(w[(static_cast<long long>(s) * cb + b) * 25 + t]);
        continue;
      }
      j -= n_down;
    }
    if (w_up) {
      if (j < n_up) {
        const int s = static_cast<int>(j % cs), b = static_cast<int>((j / cs) % cb_pad), t = static_cast<int>(j / (static_cast<long long>(cs) * cb_pad));
        float v = 0.f;
        if (cb == 3) {  // folded layout [kh][kw*3 + cb][cs] (t = kh < 5, b = kw*3 + cb < 15): see dm_conv_up
          if (t < 5 && b < 15) v = w[(static_cast<long long>(s) * 3 + (b % 3)) * 25 + t * 5 + b / 3];
        } else if (b < cb) {
          v = w[(static_cast<long long>(s) * cb + b) * 25 + t];
        }
        w_up[j] = __float2bfloat16_rn(v);
        continue;
      }
      j -= n_up;
    }
    {
      const int k = static_cast<int>(j % 128), s = static_cast<int>(j / 128);
      w_col[j] = __float2bfloat16_rn(k < cb * 25 ? w[static_cast<long long>(s) * cb * 25 + k] : 0.f);
    }
  }
}

// Stride-2 transposed convolution with few output channels (cb = 32): the four sub-pixel phases share one GEMM.
// w_upm[t = dhi*3 + dwi][n = (ph*2 + pw)*cb + c][cs] = w_up[kh*5 + kw][c][cs] with kh = ph + 2 - 2*dh, kw = pw + 2 - 2*dw
// (dh = 1 - dhi, dw = 1 - dwi) when that filter tap exists, else 0: 9 input taps x N = 4*cb instead of 25 taps x N = cb.
__global__ void __launch_bounds__(256) pack_up_merged_kernel(const __nv_bfloat16* __restrict__ w_up, int cs, int cb,
                                                             __nv_bfloat16* __restrict__ w_upm) {
  pdl_sync();
  const long long total = 9ll * 4 * cb * cs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % cs);
    long long r = i / cs;
    const int c = static_cast<int>(r % cb);
    r /= cb;
    const int phase = static_cast<int>(r % 4);
    const int t = static_cast<int>(r / 4);
    const int dh = 1 - t / 3, dw = 1 - t % 3, ph = phase >> 1, pw = phase & 1;
    const int kh = ph + 2 - 2 * dh, kw = pw + 2 - 2 * dw;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (kh >= 0 && kh < 5 && kw >= 0 && kw < 5) v = w_up[(static_cast<long long>(kh * 5 + kw) * cb + c) * cs + k];
    w_upm[i] = v;
  }
}

// w_pair[kh*3 + j][n][p*cb + c] = w_down[kh*5 + 2j + p][n][c] (zero where 2j + p = 5): the K = 2*cb operand rows of
// dm_conv_down_paired (two filter columns per k-block)
__global__ void __launch_bounds__(256) pack_down_pairs_kernel(const __nv_bfloat16* __restrict__ w_down, int cs, int cb,
                                                              __nv_bfloat16* __restrict__ w_pair) {
  pdl_sync();
  const long long total = 15ll * cs * 2 * cb;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % (2 * cb));
    long long r = i / (2 * cb);
    const int n = static_cast<int>(r % cs);
    const int t = static_cast<int>(r / cs);
    const int kh = t / 3, kw = 2 * (t % 3) + k / cb, c = k % cb;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (kw < 5) v = w_down[(static_cast<long long>(kh * 5 + kw) * cs + n) * cb + c];
    w_pair[i] = v;
  }
}

// tap-major packed conv gradient [25][n] -> master layout [n][25] (n = cs*cb); dw (+)= ; the packed buffer is
// re-zeroed so that the next backward pass can accumulate into it again
__global__ void __launch_bounds__(256) unpack_conv_grad_kernel(float* __restrict__ packed, long long n, int accumulate,
                                                               float* __restrict__ dw) {
  pdl_sync();
  constexpr int W = 128;  // elements of n per block
  __shared__ float tile[25][W + 1];
  const long long i0 = static_cast<long long>(blockIdx.x) * W;
  for (int j = threadIdx.x; j < 25 * W; j += 256) {
    const int t = j / W, ii = j - t * W;
    float v = 0.f;
    if (i0 + ii < n) {
      v = packed[t * n + i0 + ii];
      packed[t * n + i0 + ii] = 0.f;
    }
    tile[t][ii] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 25 * W; j += 256) {
    const int ii = j / 25, t = j - ii * 25;
    if (i0 + ii < n) {
      float* dst = dw + (i0 + ii) * 25 + t;
      *dst = accumulate ? *dst + tile[t][ii] : tile[t][ii];
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, long long n,
                                                        __nv_bfloat16* __restrict__ dst) {
  pdl_sync();
  const long long nv = n / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    load8(src + i * 8, f);
    store8(dst + i * 8, f);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nv * 8) dst[nv * 8 + threadIdx.x] = __float2bfloat16_rn(src[nv * 8 + threadIdx.x]);
}

// ------------------------------------------------------------------------------------------ small heads
// z = mu + eps * exp(0.5 * logvar)   (models/model.py:532-535); writes fp32 and bf16 copies
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, long long n, float* __restrict__ z_f32,
                                   __nv_bfloat16* __restrict__ z_bf16) {
  pdl_sync();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float z = mu[i] + eps[i] * expf(0.5f * logvar[i]);
  if (z_f32) z_f32[i] = z;
  if (z_bf16) z_bf16[i] = __float2bfloat16_rn(z);
}
// dmu = dz (+ dmu_ext), dlogvar = dz * eps * 0.5 * exp(0.5 logvar) (+ dlogvar_ext); bf16 copies for the GEMMs
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, const float* __restrict__ dmu_ext,
                                   const float* __restrict__ dlogvar_ext, long long n,
                                   __nv_bfloat16* __restrict__ dmu, __nv_bfloat16* __restrict__ dlogvar,
                                   float* __restrict__ dmu_f32, float* __restrict__ dlogvar_f32) {
  pdl_sync();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float g = dz ? dz[i] : 0.f;
  float a = g, b = g * eps[i] * 0.5f * expf(0.5f * logvar[i]);
  if (dmu_ext) a += dmu_ext[i];
  if (dlogvar_ext) b += dlogvar_ext[i];
  dmu[i] = __float2bfloat16_rn(a);
  dlogvar[i] = __float2bfloat16_rn(b);
  if (dmu_f32) dmu_f32[i] = a;
  if (dlogvar_f32) dlogvar_f32[i] = b;
}

// The SECOND Linear of the encoder's two heads (model.py:464,470: Linear(2048, 128) of x_to_mu and x_to_logvar) for
// BOTH heads in one launch.  These are 17 MFLOP problems: a tensor-core launch per head costs its fixed ~10 us (plus
// a zero fill, a bf16 cast, a column-sum pair ...), a SIMT kernel over both heads a few.  PairPtrs: [0] = mu, [1] = logvar.
struct PairPtrs {
  const void* x[2];   // bf16 [rows][k]   inputs of the Linear (h1)
  const void* w[2];   // bf16 [n][k]      weights (optimizer's bf16 shadow)
  const float* b[2];  // fp32 [n]         biases (forward)
  const float* d[2];  // fp32 [rows][n]   output gradients (backward)
  float* out[2];      // fp32 [rows][n]   forward outputs
  void* dx[2];        // bf16 [rows][k]   input gradients
  float* dw[2];       // fp32 [n][k]      weight gradients (accumulated)
  float* db[2];       // fp32 [n]         bias gradients (accumulated)
};

// out[r][nn] = sum_j x[r][j] w[nn][j] + b[nn]: block = 8 rows x 8 output columns (one column per warp), lanes split k.
// All nine 16-byte loads of an iteration are issued before the first FMA (guarded loads interleaved with their FMAs
// serialise on the L2 latency: 118 us instead of a few).
__global__ void __launch_bounds__(256) linear_pair_fwd_kernel(const PairPtrs p, int rows, int n, int k) {
  pdl_sync();
  const int head = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nn = blockIdx.x * 8 + warp;
  const int r0 = blockIdx.y * 8;
  if (nn >= n) return;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x[head]);
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(p.w[head]) + static_cast<long long>(nn) * k;
  float* out = p.out[head];
  const float bias = p.b[head] ? p.b[head][nn] : 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = lane * 8; j < k; j += 256) {
    float wf[8], xf[8][8];
    load8(w + j, wf);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = min(r0 + rr, rows - 1);  // (clamped: rows past the end are computed and not stored)
      load8(x + static_cast<long long>(r) * k + j, xf[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[rr] = fmaf(xf[rr][i], wf[i], acc[rr]);
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    float v = acc[rr];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && r0 + rr < rows) out[static_cast<long long>(r0 + rr) * n + nn] = v + bias;
  }
}

// dx[r][j] = sum_kk d[r][kk] w[kk][j]: block = 8 rows x 256 columns, thread = one column j; n <= 128
__global__ void __launch_bounds__(256) linear_pair_dgrad_kernel(const PairPtrs p, int rows, int n, int k) {
  pdl_sync();
  __shared__ float d_s[8][128];
  const int head = blockIdx.z;
  const int r0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < 8 * n; i += 256) {
    const int rr = i / n, kk = i - rr * n;
    d_s[rr][kk] = (r0 + rr < rows) ? p.d[head][static_cast<long long>(r0 + rr) * n + kk] : 0.f;
  }
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= k) return;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(p.w[head]);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int kk = 0; kk < n; ++kk) {
    const float wv = __bfloat162float(w[static_cast<long long>(kk) * k + j]);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) acc[rr] = fmaf(d_s[rr][kk], wv, acc[rr]);
  }
  __nv_bfloat16* dx = static_cast<__nv_bfloat16*>(p.dx[head]);
#pragma unroll
  for (int rr = 0; rr < 8; ++rr)
    if (r0 + rr < rows) dx[static_cast<long long>(r0 + rr) * k + j] = __float2bfloat16_rn(acc[rr]);
}

// dw[kk][j] += sum_r d[r][kk] x[r][j]; db[kk] += sum_r d[r][kk]: block = 16 output features x 256 columns; rows <= 256
__global__ void __launch_bounds__(256) linear_pair_wgrad_kernel(const PairPtrs p, int rows, int n, int k) {
  pdl_sync();
  __shared__ float d_s[256][16];
  const int head = blockIdx.z;
  const int k0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < rows * 16; i += 256) {
    const int r = i >> 4, kk = i & 15;
    d_s[r][kk] = (k0 + kk < n) ? p.d[head][static_cast<long long>(r) * n + k0 + kk] : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 16 && k0 + threadIdx.x < n && p.db[head]) {
    float sb = 0.f;
    for (int r = 0; r < rows; ++r) sb += d_s[r][threadIdx.x];
    p.db[head][k0 + threadIdx.x] += sb;
  }
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= k) return;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x[head]);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    const float xv = __bfloat162float(x[static_cast<long long>(r) * k + j]);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(d_s[r][i], xv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (k0 + i < n) p.dw[head][static_cast<long long>(k0 + i) * k + j] += acc[i];
}

// Discriminator head Linear(k,1)+Sigmoid (models/model.py:406-408): one warp per row
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ feat, int rows, int k,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ prob) {
  pdl_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int i = (threadIdx.x & 31) * 4; i < k; i += 128) {
    const float4 a = *reinterpret_cast<const float4*>(feat + static_cast<long long>(row) * k + i);
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + i));
    s += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) prob[row] = 1.f / (1.f + expf(-(s + b[0])));
}
// dlogit = dprob * p * (1-p); dfeat[r][:] = dfeat_ext[r][:] + dlogit[r] * w;  dw += sum_r dlogit[r] feat[r][:]; db += sum dlogit
// grid (k / 256 column blocks, row chunks of kHeadRows): the weight / bias gradients are combined with fp32 atomics
constexpr int kHeadRows = 16;
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dprob, const float* __restrict__ prob,
                                                       const float* __restrict__ feat, const float* __restrict__ dfeat_ext,
                                                       int rows, int k, const float* __restrict__ w,
                                                       float* __restrict__ dfeat, float* __restrict__ dw,
                                                       float* __restrict__ db) {
  pdl_sync();
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * kHeadRows, r1 = min(rows, r0 + kHeadRows);
  float accw = 0.f, accb = 0.f;
  const float wv = col < k ? w[col] : 0.f;
  for (int r = r0; r < r1; ++r) {
    const float p = prob[r];
    const float dl = dprob[r] * p * (1.f - p);
    accb += dl;
    if (col < k) {
      const long long o = static_cast<long long>(r) * k + col;
      dfeat[o] = (dfeat_ext ? dfeat_ext[o] : 0.f) + dl * wv;
      accw += dl * feat[o];
    }
  }
  if (col < k && dw) atomicAdd(dw + col, accw);
  if (col == 0 && db) atomicAdd(db, accb);
}

// ------------------------------------------------------------------------------------------ loss reductions
template <int NT>
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[NT / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < NT / 32) s = red[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  return s;
}

// loss[0] += wloss * sum (a-b)^2 ; if grad: grad (+)= wgrad * 2 (a-b)      (F.mse_loss(reduction='sum'),
// experiments/new_betavaegan.py:67-75; new_vae.py:40)
__global__ void __launch_bounds__(256) mse_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      long long n, float wloss, float* __restrict__ loss,
                                                      float wgrad, int grad_accumulate, float* __restrict__ grad) {
  pdl_sync();
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    acc += d * d;
    if (grad) grad[i] = (grad_accumulate ? grad[i] : 0.f) + wgrad * 2.f * d;
  }
  const float s = block_sum<256>(acc);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, wloss * s);
}

// KL = -0.5 sum(1 + logvar - mu^2 - exp(logvar)) (new_betavaegan.py:64-65); loss += w * KL;
// dmu (+)= w * mu, dlogvar (+)= w * 0.5 (exp(logvar) - 1)
__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                 long long n, float w, float* __restrict__ loss,
                                                 int grad_accumulate, float* __restrict__ dmu,
                                                 float* __restrict__ dlogvar) {
  pdl_sync();
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float m = mu[i], lv = logvar[i], e = expf(lv);
    acc += -0.5f * (1.f + lv - m * m - e);
    if (dmu) dmu[i] = (grad_accumulate ? dmu[i] : 0.f) + w * m;
    if (dlogvar) dlogvar[i] = (grad_accumulate ? dlogvar[i] : 0.f) + w * 0.5f * (e - 1.f);
  }
  const float s = block_sum<256>(acc);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, w * s);
}

// nn.BCELoss(mean) against a constant target t (new_betavaegan.py:53,97-101): loss += w * mean(-(t log p + (1-t) log(1-p)))
// with log clamped at -100; dprob (+)= w/n_total * (-(t/p) + (1-t)/(1-p)) with the clamp's zero-gradient region respected;
// stat[0] += sum p (for D_x logging)
__global__ void __launch_bounds__(256) bce_const_kernel(const float* __restrict__ p, int n, float n_total, float target,
                                                        const float* __restrict__ target_dev, float w,
                                                        float* __restrict__ loss, int grad_accumulate,
                                                        float* __restrict__ dprob, float* __restrict__ stat) {
  pdl_sync();
  if (target_dev) target = *target_dev;  // CUDA-graph mode: the per-step label lives in device memory
  float acc = 0.f, accp = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float q = p[i];
    const float lp = fmaxf(logf(q), -100.f), l1p = fmaxf(logf(1.f - q), -100.f);
    acc += -(target * lp + (1.f - target) * l1p);
    accp += q;
    if (dprob) {
      // torch: grad = (p - t) / max((1-p) p, 1e-12) / n
      const float g = (q - target) / fmaxf((1.f - q) * q, 1e-12f) * (w / n_total);
      dprob[i] = (grad_accumulate ? dprob[i] : 0.f) + g;
    }
  }
  const float s = block_sum<256>(acc);
  const float sp = block_sum<256>(accp);
  if (threadIdx.x == 0) {
    if (loss) atomicAdd(loss, w * s / n_total);
    if (stat) atomicAdd(stat, sp);
  }
}

// ------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (betas, eps, no weight decay, no amsgrad; experiments/new_betavaegan.py:49-50) on a flat buffer,
// optionally refreshing the bf16 shadow copy the GEMMs read.  28 B/param (+2 B shadow).
__global__ void adam_count_kernel(int* step) {
  pdl_sync(); *step += 1; }

__device__ __forceinline__ float4 load_grad4(const float* g, long long i) { return reinterpret_cast<const float4*>(g)[i]; }
__device__ __forceinline__ float4 load_grad4(const __nv_bfloat16* g, long long i) {
  const uint2 u = reinterpret_cast<const uint2*>(g)[i];
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float load_grad1(const float* g, long long i) { return g[i]; }
__device__ __forceinline__ float load_grad1(const __nv_bfloat16* g, long long i) { return __bfloat162float(g[i]); }

template <typename GT>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const GT* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float step_size, float beta1, float beta2, float omb1, float omb2,
                                                   float eps, float bc2_sqrt, float grad_scale, __nv_bfloat16* __restrict__ shadow,
                                                   const int* __restrict__ step_dev, double lr_d, double beta1_d,
                                                   double beta2_d, const int* __restrict__ enable, int step_offset) {
  pdl_sync();
  if (enable && *enable == 0) return;  // a deferred update whose gradient has not been produced yet (or was applied)
  if (step_dev) {  // CUDA-graph mode: bias corrections from the device-side step counter (same double arithmetic)
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const int st = *step_dev + step_offset;
      sh[0] = static_cast<float>(lr_d / (1.0 - pow(beta1_d, static_cast<double>(st))));
      sh[1] = static_cast<float>(sqrt(1.0 - pow(beta2_d, static_cast<double>(st))));
    }
    __syncthreads();
    step_size = sh[0];
    bc2_sqrt = sh[1];
  }
  const long long nv = n / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = load_grad4(g, i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pf = &pp.x; float* gf = &gg.x; float* mf = &mm.x; float* vf = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gf[j] * grad_scale;
      mf[j] = beta1 * mf[j] + omb1 * gr;
      vf[j] = beta2 * vf[j] + omb2 * gr * gr;
      const float denom = sqrtf(vf[j]) / bc2_sqrt + eps;
      pf[j] -= step_size * (mf[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
  // tail (n not a multiple of 4)
  const long long t = nv * 4 + threadIdx.x;
  if (blockIdx.x == 0 && t < n) {
    const float gr = load_grad1(g, t) * grad_scale;
    m[t] = beta1 * m[t] + omb1 * gr;
    v[t] = beta2 * v[t] + omb2 * gr * gr;
    p[t] -= step_size * (m[t] / (sqrtf(v[t]) / bc2_sqrt + eps));
    if (shadow) shadow[t] = __float2bfloat16_rn(p[t]);
  }
}

static int grid_for(long long work_items, int threads = 256, int max_blocks = 148 * 16) {
  long long b = (work_items + threads - 1) / threads;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(b, max_blocks)));
}

}  // namespace dm

using namespace dm;
typedef __nv_bfloat16 bf16;

#define DM_CHECK_C8(c, who) DM_REQUIRE((c) % 8 == 0, who ": channel count %d must be a multiple of 8", (c))

// rows of the [parts][c] partial-sum scratch handed to dm_act_backward / dm_colsum (one per row block)
extern "C" int dm_bn_parts(long long rows, int c) { return make_row_layout(rows, c).gy + 1; }

extern "C" int dm_bn_slots(void) { return kBnSlots; }
extern "C" long long dm_bn_scratch_floats(int c, int groups) { return bn_scratch_floats(c, groups); }

extern "C" int dm_bn_stats(const void* y, int y_f32, int c, const dm_bn_fuse* f, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_bn_stats");
  DM_REQUIRE(f != nullptr && f->scratch != nullptr && f->groups >= 1 && f->rows > 0 && f->gamma && f->beta &&
                 f->scale_shift && f->mean_invstd,
             "dm_bn_stats: incomplete dm_bn_fuse");
  RowLayout l = make_row_layout(f->rows, c, false, f->groups);
  const size_t sm = sizeof(float) * 256 * 16;
  if (y_f32)
    launch_pdl(bn_stats_kernel<float>, dim3(l.gx, l.gy, f->groups), dim3(l.tx, l.ty), sm, s, static_cast<const float*>(y), c, l.rows_per_block, *f);
  else
    launch_pdl(bn_stats_kernel<bf16>, dim3(l.gx, l.gy, f->groups), dim3(l.tx, l.ty), sm, s, static_cast<const bf16*>(y), c, l.rows_per_block, *f);
  DM_LAUNCHED("dm_bn_stats");
}

extern "C" int dm_bn_apply_act(const void* y, int y_f32, long long rows, int c, const float* scale_shift, int act,
                               float slope, void* out_bf16, int groups, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_bn_apply_act");
  DM_REQUIRE(groups >= 1, "dm_bn_apply_act: groups must be >= 1");
  RowLayout l = make_row_layout(rows, c, true, groups);
  if (y_f32)
    launch_pdl(bn_apply_act_kernel<float>, dim3(l.gx, l.gy, groups), dim3(l.tx, l.ty), 0, s, static_cast<const float*>(y), rows, c, l.rows_per_block, scale_shift, act, slope, static_cast<bf16*>(out_bf16));
  else
    launch_pdl(bn_apply_act_kernel<bf16>, dim3(l.gx, l.gy, groups), dim3(l.tx, l.ty), 0, s, static_cast<const bf16*>(y), rows, c, l.rows_per_block, scale_shift, act, slope, static_cast<bf16*>(out_bf16));
  DM_LAUNCHED("dm_bn_apply_act");
}

static bool bn1d_ok(long long rows, int c) {
  static const bool on = [] { const char* e = getenv("DM_BN1D"); return !e || e[0] != '0'; }();
  return on && rows <= kBn1dMaxRows && c % 32 == 0;
}

// BatchNorm forward (training mode) + activation: statistics + normalisation + running-stat update.
//   rows <= 256: ONE launch (bn1d_fwd_kernel, exact two-pass variance; `scratch` unused, may be NULL);
//   otherwise:   bn_stats_kernel (slot partial sums, shift = running_mean, finalize by its last block) + bn_apply_act.
extern "C" int dm_bn_forward(const void* y, int y_f32, long long rows, int c, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                             float eps, int act, float slope, float* scratch, float* scale_shift, float* mean_invstd,
                             void* out_bf16, int groups, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_bn_forward");
  DM_REQUIRE(groups >= 1, "dm_bn_forward: groups must be >= 1");
  if (bn1d_ok(rows, c)) {
    if (y_f32)
      launch_pdl(bn1d_fwd_kernel<float>, c / 32, dim3(4, 64), 0, s, static_cast<const float*>(y), static_cast<int>(rows), c, gamma, beta,
                                                           running_mean, running_var, num_batches_tracked, momentum, eps, act,
                                                           slope, scale_shift, mean_invstd, static_cast<bf16*>(out_bf16), groups,
                                                           static_cast<long long>(c), static_cast<const float*>(nullptr));
    else
      launch_pdl(bn1d_fwd_kernel<bf16>, c / 32, dim3(4, 64), 0, s, static_cast<const bf16*>(y), static_cast<int>(rows), c, gamma, beta,
                                                          running_mean, running_var, num_batches_tracked, momentum, eps, act,
                                                          slope, scale_shift, mean_invstd, static_cast<bf16*>(out_bf16), groups,
                                                          static_cast<long long>(c), static_cast<const float*>(nullptr));
    DM_LAUNCHED("dm_bn_forward(1d)");
  }
  DM_REQUIRE(scratch != nullptr, "dm_bn_forward: scratch required for rows > %d", kBn1dMaxRows);
  dm_bn_fuse f;
  f.scratch = scratch; f.groups = groups; f.rows = rows; f.gamma = gamma; f.beta = beta;
  f.running_mean = running_mean; f.running_var = running_var; f.num_batches_tracked = num_batches_tracked;
  f.momentum = momentum; f.eps = eps; f.scale_shift = scale_shift; f.mean_invstd = mean_invstd;
  if (int rc = dm_bn_stats(y, y_f32, c, &f, stream_)) return rc;
  return dm_bn_apply_act(y, y_f32, rows, c, scale_shift, act, slope, out_bf16, groups, stream_);
}

// Small-row BatchNorm (rows <= 256) on a COLUMN BLOCK of a wider fp32 matrix: y = base + column offset, row stride ld_y.
// Used for the encoder's two heads computed by one N = 4096 GEMM (model.py:460-471): pre_bias = the head's Linear bias,
// which that GEMM does not add.  One pass (groups = 1).
extern "C" int dm_bn1d_forward(const float* y, long long ld_y, int rows, int c, const float* pre_bias, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float momentum, float eps, int act, float slope,
                               float* scale_shift, float* mean_invstd, void* out_bf16, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(bn1d_ok(rows, c) && ld_y >= c && ld_y % 8 == 0, "dm_bn1d_forward: rows %d (<= %d), c %d (%% 32), ld %lld", rows,
             kBn1dMaxRows, c, ld_y);
  launch_pdl(bn1d_fwd_kernel<float>, c / 32, dim3(4, 64), 0, s, y, rows, c, gamma, beta, running_mean, running_var,
             num_batches_tracked, momentum, eps, act, slope, scale_shift, mean_invstd, static_cast<bf16*>(out_bf16), 1, ld_y,
             pre_bias);
  DM_LAUNCHED("dm_bn1d_forward");
}

extern "C" int dm_bn1d_backward(const void* dout_bf16, const float* y, long long ld_y, int rows, int c,
                                const float* scale_shift, const float* mean_invstd, int act, float slope, void* dy_bf16,
                                long long ld_dy, float* dgamma, float* dbeta, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(bn1d_ok(rows, c) && ld_y >= c && ld_dy >= c && ld_y % 8 == 0 && ld_dy % 8 == 0,
             "dm_bn1d_backward: rows %d (<= %d), c %d (%% 32), ld %lld / %lld", rows, kBn1dMaxRows, c, ld_y, ld_dy);
  launch_pdl(bn1d_bwd_kernel<float>, c / 32, dim3(4, 64), 0, s, static_cast<const bf16*>(dout_bf16), y, rows, c, scale_shift,
             mean_invstd, act, slope, static_cast<bf16*>(dy_bf16), dgamma, dbeta, 1, ld_y, ld_dy);
  DM_LAUNCHED("dm_bn1d_backward");
}

extern "C" int dm_bn_backward(const void* dout_bf16, const void* y, int y_f32, long long rows, int c,
                              const float* scale_shift, const float* mean_invstd, int act, float slope,
                              float* scratch, void* dy_bf16, float* dgamma, float* dbeta, int groups,
                              void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_bn_backward");
  DM_REQUIRE(groups >= 1, "dm_bn_backward: groups must be >= 1");
  const bf16* d = static_cast<const bf16*>(dout_bf16);
  if (bn1d_ok(rows, c)) {
    if (y_f32)
      launch_pdl(bn1d_bwd_kernel<float>, c / 32, dim3(4, 64), 0, s, d, static_cast<const float*>(y), static_cast<int>(rows), c, scale_shift,
                                                           mean_invstd, act, slope, static_cast<bf16*>(dy_bf16), dgamma, dbeta, groups,
                                                           static_cast<long long>(c), static_cast<long long>(c));
    else
      launch_pdl(bn1d_bwd_kernel<bf16>, c / 32, dim3(4, 64), 0, s, d, static_cast<const bf16*>(y), static_cast<int>(rows), c, scale_shift,
                                                          mean_invstd, act, slope, static_cast<bf16*>(dy_bf16), dgamma, dbeta, groups,
                                                          static_cast<long long>(c), static_cast<long long>(c));
    DM_LAUNCHED("dm_bn_backward(1d)");
  }
  DM_REQUIRE(scratch != nullptr, "dm_bn_backward: scratch required for rows > %d", kBn1dMaxRows);
  RowLayout l = make_row_layout(rows, c, false, groups);
  const RowLayout la = make_row_layout(rows, c, true, groups);
  const size_t sm = sizeof(float) * 256 * 16;
  {
    // single-launch form: the whole grid must be resident (grid barrier) -> at most two blocks per SM, all groups included
    // OFF by default: measured 4.90 vs 4.61 ms per batch-64 step (profiles/r02b_ab_bn_bwd_fused.txt) -- the barrier makes
    // every block wait for the last one to get an SM slot, and the side-stream weight-gradient GEMMs hold those slots
    static const bool fused_on = [] { const char* e = getenv("DM_BN_BWD_FUSED"); return e && e[0] == '1'; }();
    static const int occ = [sm] {
      int a = 0, b = 0, sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, bn_bwd_fused_kernel<bf16>, 256, sm);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, bn_bwd_fused_kernel<float>, 256, sm);
      return std::min(a, b) >= 2 ? 2 * sms : 0;
    }();
    const long long per_group = occ / std::max(1, l.gx * groups);
    if (fused_on && per_group >= 1) {
      RowLayout f = l;
      long long rpb = (rows + per_group - 1) / per_group;
      rpb = std::max<long long>(f.ty, (rpb + f.ty - 1) / f.ty * f.ty);
      f.rows_per_block = rpb;
      f.gy = static_cast<int>((rows + rpb - 1) / rpb);
      if (static_cast<long long>(f.gx) * f.gy * groups <= occ) {
        if (y_f32)
          launch_pdl(bn_bwd_fused_kernel<float>, dim3(f.gx, f.gy, groups), dim3(f.tx, f.ty), sm, s, d, static_cast<const float*>(y), rows, c, f.rows_per_block, scale_shift, mean_invstd, act, slope, scratch, dgamma, dbeta, static_cast<bf16*>(dy_bf16));
        else
          launch_pdl(bn_bwd_fused_kernel<bf16>, dim3(f.gx, f.gy, groups), dim3(f.tx, f.ty), sm, s, d, static_cast<const bf16*>(y), rows, c, f.rows_per_block, scale_shift, mean_invstd, act, slope, scratch, dgamma, dbeta, static_cast<bf16*>(dy_bf16));
        DM_LAUNCHED("dm_bn_backward(fused)");
      }
    }
  }
  const float* sums = scratch + bn_slot_floats(c, groups) + 4;
  if (y_f32)
    launch_pdl(bn_bwd_reduce_kernel<float>, dim3(l.gx, l.gy, groups), dim3(l.tx, l.ty), sm, s, d, static_cast<const float*>(y), rows, c, l.rows_per_block, scale_shift, mean_invstd, act, slope, scratch, dgamma, dbeta);
  else
    launch_pdl(bn_bwd_reduce_kernel<bf16>, dim3(l.gx, l.gy, groups), dim3(l.tx, l.ty), sm, s, d, static_cast<const bf16*>(y), rows, c, l.rows_per_block, scale_shift, mean_invstd, act, slope, scratch, dgamma, dbeta);
  if (y_f32)
    launch_pdl(bn_bwd_apply_kernel<float>, dim3(la.gx, la.gy, groups), dim3(la.tx, la.ty), 0, s, d, static_cast<const float*>(y), rows, c, la.rows_per_block, scale_shift, mean_invstd, sums, act, slope, static_cast<bf16*>(dy_bf16));
  else
    launch_pdl(bn_bwd_apply_kernel<bf16>, dim3(la.gx, la.gy, groups), dim3(la.tx, la.ty), 0, s, d, static_cast<const bf16*>(y), rows, c, la.rows_per_block, scale_shift, mean_invstd, sums, act, slope, static_cast<bf16*>(dy_bf16));
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return check_launch("dm_bn_backward");
}

extern "C" int dm_bias_act(const float* acc, long long rows, int c, const float* bias, int act, float slope,
                           float* out_f32, void* out_bf16, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_bias_act");
  RowLayout l = make_row_layout(rows, c);
  launch_pdl(bias_act_kernel, dim3(l.gx, l.gy), dim3(l.tx, l.ty), 0, s, acc, rows, c, l.rows_per_block, bias, act, slope, out_f32, static_cast<bf16*>(out_bf16));
  DM_LAUNCHED("dm_bias_act");
}

extern "C" int dm_act_backward(const float* dout, const float* out, long long rows, int c, int act, float slope,
                               void* dpre_bf16, float* partials, float* colsum, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_act_backward");
  RowLayout l = make_row_layout(rows, c);
  launch_pdl(act_bwd_colsum_kernel, dim3(l.gx, l.gy), dim3(l.tx, l.ty), sizeof(float) * 256 * 8, s, dout, out, rows, c, l.rows_per_block, act, slope, static_cast<bf16*>(dpre_bf16), partials);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (colsum) {
    launch_pdl(reduce_partials_kernel, (c + 31) / 32, dim3(32, 32), 0, s, partials, l.gy, c, 1, colsum);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  return check_launch("dm_act_backward");
}

extern "C" int dm_colsum(const void* x, int x_f32, long long rows, int c, float* partials, float* colsum, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_CHECK_C8(c, "dm_colsum");
  RowLayout l = make_row_layout(rows, c);
  const size_t sm = sizeof(float) * 256 * 8;
  if (x_f32)
    launch_pdl(colsum_kernel<float>, dim3(l.gx, l.gy), dim3(l.tx, l.ty), sm, s, static_cast<const float*>(x), rows, c, l.rows_per_block, partials);
  else
    launch_pdl(colsum_kernel<bf16>, dim3(l.gx, l.gy), dim3(l.tx, l.ty), sm, s, static_cast<const bf16*>(x), rows, c, l.rows_per_block, partials);
  launch_pdl(reduce_partials_kernel, (c + 31) / 32, dim3(32, 32), 0, s, partials, l.gy, c, 1, colsum);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return check_launch("dm_colsum");
}

extern "C" int dm_im2col3(const float* x_nchw, int batch, int h, int w, int stride, void* col_bf16, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(stride == 1 || stride == 2, "dm_im2col3: stride must be 1 or 2");
  const long long items = static_cast<long long>(batch) * (h / stride) * (w / stride) * 10;
  (void)items;
  const long long out_rows = static_cast<long long>(batch) * (h / stride);
  const int blocks = static_cast<int>(std::min<long long>(out_rows, 148 * 16));
  launch_pdl(im2col3_kernel, blocks, 256, 15 * (w + 4) * sizeof(float), s, x_nchw, batch, h, w, stride, static_cast<bf16*>(col_bf16));
  DM_LAUNCHED("dm_im2col3");
}

extern "C" int dm_nhwc3_to_nchw(const float* src, long long batch, int hw, int apply_tanh, float* dst, void* pim_bf16,
                                void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(pim_bf16 == nullptr || hw == 64 * 64, "dm_nhwc3_to_nchw: the padded-image output needs 64x64 images");
  launch_pdl(nhwc3_to_nchw_kernel, grid_for(batch * hw), 256, 0, s, src, batch, hw, apply_tanh, dst, static_cast<bf16*>(pim_bf16));
  DM_LAUNCHED("dm_nhwc3_to_nchw");
}

extern "C" int dm_tanh_backward(const float* dout, const float* out, long long batch, int hw, float* dy,
                                float* bias_grad, void* pim_bf16, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(pim_bf16 == nullptr || hw == 64 * 64, "dm_tanh_backward: the padded-image output needs 64x64 images");
  launch_pdl(tanh_bwd_kernel, grid_for(batch * hw, 256, 148 * 4), 256, 0, s, dout, out, batch, hw, dy, bias_grad, static_cast<bf16*>(pim_bf16));
  DM_LAUNCHED("dm_tanh_backward");
}

// elements of a padded-image buffer for `batch` images: [batch][68][72][4] plus 64 elements of slack (the 16-pixel windows
// of the weight-gradient GEMM read up to 48 bytes past the last row of the last image; what they read is discarded)
extern "C" long long dm_pim_elems(int batch) { return static_cast<long long>(batch) * kPimH * kPimW * 4 + 64; }

extern "C" int dm_pad_image3(const void* src, int src_u8, int batch, void* pim_bf16, float* dst_nchw, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(batch > 0 && src != nullptr && pim_bf16 != nullptr, "dm_pad_image3: bad arguments");
  DM_REQUIRE((reinterpret_cast<uintptr_t>(pim_bf16) & 127) == 0, "dm_pad_image3: pim must be 128-byte aligned");
  const int blocks = static_cast<int>(std::min<long long>(static_cast<long long>(batch) * 64, 148 * 16));
  launch_pdl(pad_image3_kernel, blocks, 256, 0, s, src, src_u8, batch, static_cast<bf16*>(pim_bf16), dst_nchw);
  DM_LAUNCHED("dm_pad_image3");
}

extern "C" int dm_pack_conv3_weights(const float* w, int cs, int stride, void* w_win, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(stride == 1 || stride == 2, "dm_pack_conv3_weights: stride must be 1 or 2");
  launch_pdl(pack_win_kernel, grid_for(5ll * cs * 64), 256, 0, s, w, cs, stride, static_cast<bf16*>(w_win));
  DM_LAUNCHED("dm_pack_conv3_weights");
}

extern "C" int dm_unpack_conv3_grad(float* scratch, int cs, int stride, float* dw, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(stride == 1 || stride == 2, "dm_unpack_conv3_grad: stride must be 1 or 2");
  launch_pdl(unpack_win_grad_kernel, grid_for(5ll * cs * 128), 256, 0, s, scratch, cs, stride, dw);
  DM_LAUNCHED("dm_unpack_conv3_grad");
}

extern "C" int dm_transpose_bf16(const void* src, int batch, int rows, int cols, void* dst, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(transpose_kernel, dim3((cols + 31) / 32, (rows + 31) / 32, batch), dim3(32, 8), 0, s, static_cast<const bf16*>(src), rows, cols, static_cast<bf16*>(dst));
  DM_LAUNCHED("dm_transpose_bf16");
}

extern "C" int dm_pack_conv_weights(const float* w, int cs, int cb, void* w_down, void* w_up, void* w_col,
                                    void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(w_col == nullptr || cb * 25 <= 128, "dm_pack_conv_weights: col form needs cb*25 <= 128");
  const int cb_pad = std::max(16, (cb + 15) / 16 * 16);
  const long long total = 25ll * cs * cb + 25ll * cb_pad * cs + 128ll * cs;
  launch_pdl(pack_conv_kernel, grid_for(total), 256, 0, s, w, cs, cb, cb_pad, static_cast<bf16*>(w_down), static_cast<bf16*>(w_up), static_cast<bf16*>(w_col));
  DM_LAUNCHED("dm_pack_conv_weights");
}

extern "C" int dm_pack_up_merged(const void* w_up, int cs, int cb, void* w_upm, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(cb % 16 == 0, "dm_pack_up_merged: cb %d must be a multiple of 16", cb);
  launch_pdl(pack_up_merged_kernel, grid_for(36ll * cb * cs), 256, 0, s, static_cast<const bf16*>(w_up), cs, cb, static_cast<bf16*>(w_upm));
  DM_LAUNCHED("dm_pack_up_merged");
}

extern "C" int dm_pack_down_pairs(const void* w_down, int cs, int cb, void* w_pair, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(cb == 32, "dm_pack_down_pairs: cb %d must be 32 (64-element paired rows)", cb);
  launch_pdl(pack_down_pairs_kernel, grid_for(30ll * cs * cb), 256, 0, s, static_cast<const bf16*>(w_down), cs, cb, static_cast<bf16*>(w_pair));
  DM_LAUNCHED("dm_pack_down_pairs");
}

extern "C" int dm_unpack_conv_grad(float* packed, int cs, int cb, int accumulate, float* dw, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  const long long n = static_cast<long long>(cs) * cb;
  launch_pdl(unpack_conv_grad_kernel, static_cast<unsigned>((n + 127) / 128), 256, 0, s, packed, n, accumulate, dw);
  DM_LAUNCHED("dm_unpack_conv_grad");
}

extern "C" int dm_cast_bf16(const float* src, long long n, void* dst, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dm_cast_bf16: pointers must be 16-byte aligned");
  launch_pdl(cast_bf16_kernel, grid_for(n / 8 + 1), 256, 0, s, src, n, static_cast<bf16*>(dst));
  DM_LAUNCHED("dm_cast_bf16");
}

extern "C" int dm_reparam_forward(const float* mu, const float* logvar, const float* eps, long long n, float* z_f32,
                                  void* z_bf16, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(reparam_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, s, mu, logvar, eps, n, z_f32, static_cast<bf16*>(z_bf16));
  DM_LAUNCHED("dm_reparam_forward");
}

extern "C" int dm_reparam_backward(const float* dz, const float* logvar, const float* eps, const float* dmu_ext,
                                   const float* dlogvar_ext, long long n, void* dmu_bf16, void* dlogvar_bf16,
                                   float* dmu_f32, float* dlogvar_f32, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(reparam_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, s, dz, logvar, eps, dmu_ext, dlogvar_ext, n, static_cast<bf16*>(dmu_bf16), static_cast<bf16*>(dlogvar_bf16), dmu_f32, dlogvar_f32);
  DM_LAUNCHED("dm_reparam_backward");
}

extern "C" int dm_head_forward(const float* feat, int rows, int k, const float* w, const float* b, float* prob,
                               void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(k % 4 == 0, "dm_head_forward: k must be a multiple of 4");
  launch_pdl(head_fwd_kernel, (rows + 7) / 8, 256, 0, s, feat, rows, k, w, b, prob);
  DM_LAUNCHED("dm_head_forward");
}

extern "C" int dm_linear_pair_forward(const void* x0, const void* x1, const void* w0, const void* w1, const float* b0,
                                      const float* b1, int rows, int n, int k, float* out0, float* out1, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(k % 256 == 0 && n % 8 == 0 && rows >= 1, "dm_linear_pair_forward: k %d (%% 256), n %d (%% 8)", k, n);
  PairPtrs p = {};
  p.x[0] = x0; p.x[1] = x1; p.w[0] = w0; p.w[1] = w1; p.b[0] = b0; p.b[1] = b1; p.out[0] = out0; p.out[1] = out1;
  launch_pdl(linear_pair_fwd_kernel, dim3(n / 8, (rows + 7) / 8, 2), 256, 0, s, p, rows, n, k);
  DM_LAUNCHED("dm_linear_pair_forward");
}

extern "C" int dm_linear_pair_backward(const float* d0, const float* d1, const void* x0, const void* x1, const void* w0,
                                       const void* w1, int rows, int n, int k, void* dx0, void* dx1, float* dw0,
                                       float* dw1, float* db0, float* db1, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(k % 256 == 0 && n <= 128 && n % 16 == 0 && rows >= 1 && rows <= 256,
             "dm_linear_pair_backward: k %d (%% 256), n %d (<= 128, %% 16), rows %d (<= 256)", k, n, rows);
  PairPtrs p = {};
  p.d[0] = d0; p.d[1] = d1; p.x[0] = x0; p.x[1] = x1; p.w[0] = w0; p.w[1] = w1;
  p.dx[0] = dx0; p.dx[1] = dx1; p.dw[0] = dw0; p.dw[1] = dw1; p.db[0] = db0; p.db[1] = db1;
  launch_pdl(linear_pair_dgrad_kernel, dim3(k / 256, (rows + 7) / 8, 2), 256, 0, s, p, rows, n, k);
  if (dw0 && dw1) {
    launch_pdl(linear_pair_wgrad_kernel, dim3(k / 256, n / 16, 2), 256, 0, s, p, rows, n, k);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  DM_LAUNCHED("dm_linear_pair_backward");
}

extern "C" int dm_head_backward(const float* dprob, const float* prob, const float* feat, const float* dfeat_ext,
                                int rows, int k, const float* w, float* dfeat, float* dw, float* db, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(head_bwd_kernel, dim3((k + 255) / 256, (rows + kHeadRows - 1) / kHeadRows), 256, 0, s, dprob, prob, feat, dfeat_ext, rows, k, w, dfeat, dw, db);
  DM_LAUNCHED("dm_head_backward");
}

extern "C" int dm_mse_sum(const float* a, const float* b, long long n, float wloss, float* loss, float wgrad,
                          int grad_accumulate, float* grad, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(mse_sum_kernel, grid_for(n, 256, 148 * 4), 256, 0, s, a, b, n, wloss, loss, wgrad, grad_accumulate, grad);
  DM_LAUNCHED("dm_mse_sum");
}

extern "C" int dm_kl(const float* mu, const float* logvar, long long n, float w, float* loss, int grad_accumulate,
                     float* dmu, float* dlogvar, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(kl_kernel, grid_for(n, 256, 148), 256, 0, s, mu, logvar, n, w, loss, grad_accumulate, dmu, dlogvar);
  DM_LAUNCHED("dm_kl");
}

extern "C" int dm_bce_const(const float* p, int n, float n_total, float target, const float* target_dev, float w,
                            float* loss, int grad_accumulate, float* dprob, float* stat, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  launch_pdl(bce_const_kernel, grid_for(n, 256, 8), 256, 0, s, p, n, n_total, target, target_dev, w, loss, grad_accumulate, dprob, stat);
  DM_LAUNCHED("dm_bce_const");
}

static int adam_impl(float* p, const void* g, int g_bf16, float* m, float* v, long long n, double lr,
                     double beta1, double beta2, double eps, int step, int* step_dev, int count_step,
                     float grad_scale, void* shadow_bf16, const int* enable_dev, int step_offset, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(step >= 1 || step_dev != nullptr, "dm_adam_step: step must be >= 1 (or a device counter given)");
  // scalar arithmetic in double, then rounded to float once -- as torch.optim.Adam does with Python floats
  float step_size = 0.f, bc2s = 1.f;
  if (step_dev) {
    if (count_step) {  // *step_dev += 1, then the update reads it; later segments of the same step reuse the value
      launch_pdl(adam_count_kernel, 1, 1, 0, s, step_dev);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
  } else {
    const double bc1 = 1.0 - pow(beta1, step);
    const double bc2 = 1.0 - pow(beta2, step);
    step_size = static_cast<float>(lr / bc1);
    bc2s = static_cast<float>(sqrt(bc2));
  }
  const int grid = grid_for(n / 4 + 1, 256, 148 * 8);
  if (g_bf16)
    launch_pdl(adam_kernel<bf16>, grid, 256, 0, s, p, static_cast<const bf16*>(g), m, v, n, step_size, static_cast<float>(beta1),
                                           static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                           static_cast<float>(1.0 - beta2), static_cast<float>(eps), bc2s, grad_scale,
                                           static_cast<bf16*>(shadow_bf16), step_dev, lr, beta1, beta2, enable_dev, step_offset);
  else
    launch_pdl(adam_kernel<float>, grid, 256, 0, s, p, static_cast<const float*>(g), m, v, n, step_size, static_cast<float>(beta1),
                                            static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                            static_cast<float>(1.0 - beta2), static_cast<float>(eps), bc2s, grad_scale,
                                            static_cast<bf16*>(shadow_bf16), step_dev, lr, beta1, beta2, enable_dev, step_offset);
  DM_LAUNCHED("dm_adam_step");
}

extern "C" int dm_adam_step_ex(float* p, const void* g, int g_bf16, float* m, float* v, long long n, double lr,
                               double beta1, double beta2, double eps, int step, int* step_dev, int count_step,
                               float grad_scale, void* shadow_bf16, void* stream_) {
  return adam_impl(p, g, g_bf16, m, v, n, lr, beta1, beta2, eps, step, step_dev, count_step, grad_scale, shadow_bf16,
                   nullptr, 0, stream_);
}

// Same, gated by a device-side flag: the launch is a no-op when *enable_dev == 0.  Lets a CUDA graph contain the DEFERRED
// update of a tensor (applied at the start of the next step, under work that does not read it): the flag says whether
// the gradient buffer holds an unapplied gradient.
extern "C" int dm_adam_step_gated(float* p, const void* g, int g_bf16, float* m, float* v, long long n, double lr,
                                  double beta1, double beta2, double eps, int* step_dev, int step_offset,
                                  float grad_scale, void* shadow_bf16, const int* enable_dev, void* stream_) {
  DM_REQUIRE(step_dev != nullptr, "dm_adam_step_gated: device step counter required");
  return adam_impl(p, g, g_bf16, m, v, n, lr, beta1, beta2, eps, 0, step_dev, 0, grad_scale, shadow_bf16, enable_dev,
                   step_offset, stream_);
}

extern "C" int dm_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                            double beta2, double eps, int step, int* step_dev, float grad_scale, void* shadow_bf16,
                            void* stream_) {
  return dm_adam_step_ex(p, g, 0, m, v, n, lr, beta1, beta2, eps, step, step_dev, 1, grad_scale, shadow_bf16, stream_);
}
