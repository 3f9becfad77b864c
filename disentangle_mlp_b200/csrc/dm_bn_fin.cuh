// BatchNorm finalize step, executed by the LAST block / CTA of the kernel that produced the per-channel partial sums
// (dm_elem.cu: bn_stats_kernel, bn_bwd_reduce_kernel; dm_gemm.cu: the epilogue of the GEMM that writes the
// pre-BatchNorm tensor).  Producers add their partial sums to a small slot scratch [groups][kBnSlots][2][c] with fp32
// red.add, publish them (__threadfence) and take a ticket; whoever draws the last ticket sums the slots, writes the
// constants the consumer kernel needs, updates the running statistics and re-zeroes the slots and the ticket: the
// scratch is zero on entry and zero on exit, and no separate finalize kernel is launched.
#pragma once
#include <cuda_runtime.h>

#include "../../include/dm_b200.h"
#include "dm_common.h"

namespace dm {

// floats of a call site's scratch: slots | ticket (+3 pad) | backward sums [groups][2][c]
__host__ __device__ inline long long bn_slot_floats(int c, int groups) {
  return static_cast<long long>(groups) * kBnSlots * 2 * c;
}
__host__ __device__ inline long long bn_scratch_floats(int c, int groups) {
  return bn_slot_floats(c, groups) + 4 + 2ll * groups * c;
}
__device__ __forceinline__ unsigned int* bn_ticket(float* scratch, int c, int groups) {
  return reinterpret_cast<unsigned int*>(scratch + bn_slot_floats(c, groups));
}

// Forward finalize (training-mode nn.BatchNorm1d/2d; models/model.py:451-457,390-399): for every channel and every
// stacked pass, in pass order: mean / biased variance from the SHIFTED sums (k = running_mean before this call),
// scale = gamma * invstd, shift = beta - mean * scale, running stats with momentum and the unbiased variance.
// Called by all `nt` threads (tid in [0, nt)) of the last producer block.  fp32 arithmetic: s2/n - dmean^2 only
// cancels by (mean - k)^2 / var, and the running mean tracks the batch mean.
__device__ __forceinline__ void bn_forward_finalize(const dm_bn_fuse& f, int c, int tid, int nt) {
  const float inv_n = 1.f / static_cast<float>(f.rows);
  const float unb = f.rows > 1 ? static_cast<float>(f.rows) / static_cast<float>(f.rows - 1) : 1.f;
  for (int ch = tid; ch < c; ch += nt) {
    const float k = f.running_mean ? f.running_mean[ch] : 0.f;
    float rm = k, rv = f.running_var ? f.running_var[ch] : 0.f;
    const float ga = f.gamma[ch], be = f.beta[ch];
    for (int g = 0; g < f.groups; ++g) {
      float* sl = f.scratch + static_cast<long long>(g) * kBnSlots * 2 * c + ch;
      float v1[kBnSlots], v2[kBnSlots];
#pragma unroll
      for (int p = 0; p < kBnSlots; ++p) {  // 16 independent L2 loads in flight
        v1[p] = __ldcg(sl + (2ll * p) * c);
        v2[p] = __ldcg(sl + (2ll * p + 1) * c);
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int p = 0; p < kBnSlots; ++p) {
        s1 += v1[p];
        s2 += v2[p];
        sl[(2ll * p) * c] = 0.f;  // this thread read it, this thread clears it
        sl[(2ll * p + 1) * c] = 0.f;
      }
      const float dmean = s1 * inv_n;
      const float mean = k + dmean;
      const float var = fmaxf(fmaf(-dmean, dmean, s2 * inv_n), 0.f);
      const float invstd = 1.f / sqrtf(var + f.eps);
      const float sc = ga * invstd;
      float* ss = f.scale_shift + static_cast<long long>(g) * 2 * c;
      float* mi = f.mean_invstd + static_cast<long long>(g) * 2 * c;
      ss[ch] = sc;
      ss[c + ch] = be - mean * sc;
      mi[ch] = mean;
      mi[c + ch] = invstd;
      rm = (1.f - f.momentum) * rm + f.momentum * mean;
      rv = (1.f - f.momentum) * rv + f.momentum * var * unb;
    }
    if (f.running_mean) {
      f.running_mean[ch] = rm;
      f.running_var[ch] = rv;
    }
  }
  if (tid == 0) {
    if (f.num_batches_tracked) *f.num_batches_tracked += f.groups;
    *bn_ticket(f.scratch, c, f.groups) = 0u;
  }
}

// Backward finalize: sums[g][0][c] = sum dz, sums[g][1][c] = sum dz * xhat (read by bn_bwd_apply_kernel);
// dgamma += sum over passes of sum dz*xhat, dbeta += sum dz (accumulating, like autograd's AccumulateGrad).
__device__ __forceinline__ void bn_backward_finalize(float* scratch, int c, int groups, float* dgamma, float* dbeta,
                                                     int tid, int nt) {
  float* sums = scratch + bn_slot_floats(c, groups) + 4;
  for (int ch = tid; ch < c; ch += nt) {
    float g0 = 0.f, g1 = 0.f;
    for (int g = 0; g < groups; ++g) {
      float* sl = scratch + static_cast<long long>(g) * kBnSlots * 2 * c + ch;
      float v1[kBnSlots], v2[kBnSlots];
#pragma unroll
      for (int p = 0; p < kBnSlots; ++p) {
        v1[p] = __ldcg(sl + (2ll * p) * c);
        v2[p] = __ldcg(sl + (2ll * p + 1) * c);
      }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int p = 0; p < kBnSlots; ++p) {
        s0 += v1[p];
        s1 += v2[p];
        sl[(2ll * p) * c] = 0.f;
        sl[(2ll * p + 1) * c] = 0.f;
      }
      sums[static_cast<long long>(g) * 2 * c + ch] = s0;
      sums[static_cast<long long>(g) * 2 * c + c + ch] = s1;
      g0 += s0;
      g1 += s1;
    }
    if (dgamma) dgamma[ch] += g1;
    if (dbeta) dbeta[ch] += g0;
  }
  if (tid == 0) *bn_ticket(scratch, c, groups) = 0u;
}

}  // namespace dm
