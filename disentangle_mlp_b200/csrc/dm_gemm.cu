// Tap-GEMM: one tcgen05/TMEM/TMA kernel behind every dense contraction of the training step.
//
//   D[128 x BN tile] = sum over k-blocks  A_kb * B_kb^T      (bf16 x bf16 -> fp32 in TMEM)
//
// MODE_FWD   : M = rows of a pixel tile (or plain matrix rows); the K loop walks (filter tap, channel
//              chunk).  The A k-block of tap (kh,kw) is ONE 5-D TMA box of the NHWC activation tensor,
//              shifted by the tap; padding comes from TMA out-of-bounds zero fill, stride 2 from a
//              parity-split view (c,w,p,h,n) = (2C, W/2, 2, H/2, B) of the same memory.  Implicit GEMM:
//              no im2col buffer exists.  B = packed weights [tap][N][K] (K-major) or [K][N] (MN-major).
// MODE_WGRAD : M and N are channels, the K loop walks pixel tiles (K = batch*h*w).  Both operands are
//              MN-major boxes of NHWC tensors (channels contiguous), B shifted by one fixed tap.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 =
// epilogue (TMEM -> registers -> global).  smem ring of `stages` k-blocks guarded by full/empty mbarriers.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/dm_b200.h"
#include "dm_common.h"
#include "dm_bn_fin.cuh"
#include "dm_ptx.cuh"

namespace dm {

extern std::atomic<long long> g_launch_count;

struct Tap {
  int16_t dc;      // offset along dim 0 (channels / parity-merged channels)
  int8_t dw, dp, dh;
  uint8_t wt;      // weight tap index (B coordinate 2 in MODE_FWD)
  int16_t nvalid;  // MODE_WGRAD: valid columns of this tap unit
  int16_t mvalid;  // MODE_WGRAD: valid rows of this tap unit
  int16_t out_tap; // MODE_WGRAD, TMA epilogue: tap coordinate of this unit in the packed gradient [25][cs][cb]
  int32_t out_off; // MODE_WGRAD: element offset of this tap unit in the output
};

enum { MODE_FWD = 0, MODE_WGRAD = 1 };

struct alignas(64) GemmParams {
  CUtensorMap map_a;
  CUtensorMap map_b;
  int mode;
  int a_mn, b_mn;  // 0 = K-major, 1 = MN-major
  int tf32;        // 1: fp32 operands read as TF32 (kind::tf32): kc = 32 elements (128 B), MN-major atoms 32 x 32
  int kc;          // K elements per k-block: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B, K-major only)
  int bn;          // N tile (multiple of 16, <= 256)
  int stages;
  int tmem_cols;
  int num_splits;
  int num_m_tiles, num_n_tiles;
  int num_units;    // MODE_WGRAD: tap units
  int total_tiles;  // persistent: CTAs grid-stride over [0, total_tiles)
  int acc_stride;   // TMEM columns between the two accumulator buffers
  int cluster;      // 1, or 2 = CTA pairs on adjacent M tiles (cg2)
  int cg2;          // cluster == 2, FWD: ONE tcgen05.mma.cta_group::2 (M = 256) per CTA pair; each CTA holds its 128
                    // A rows and half of the B rows, the leader CTA issues, both read their own TMEM half
  int cpt;     // MODE_FWD: channel chunks per tap
  int num_kb;  // MODE_WGRAD: total pixel-tile k-blocks
  // pixel-tile decode: tile j -> (w0, h0, n0); used for the M tile (FWD) or the K tile (WGRAD)
  int tw_step, tpi, th_step, tn_step;
  int bw, bh;  // FWD row r -> w = w0 + r % bw, h = h0 + (r / bw) % bh, n = n0 + r / (bw*bh)
  int phase_tap_start[5];
  Tap taps[28];
  void* out;
  const float* bias;
  int out_f32, out_atomic;
  int fold_kw;  // FWD, stride-1 transposed conv to 3 channels: accumulator column = kw*3 + cb, epilogue sums the
                // five horizontally shifted partial results (out[w] = sum_kw acc[w + 2 - kw][kw])
  // FWD epilogue: element offset of row (w,h,n) and column j
  long long os_w, os_h, os_n, os_col;
  long long phase_out_off[4];
  int w_lim, n_lim;  // row validity
  int n_valid;       // global column validity
  // WGRAD epilogue: off = (m % mmod)*os_m + (m / mmod)*os_m2 + (n % nmod)*os_n1 + (n / nmod)*os_n2 + tap.out_off
  long long os_m, os_m2, os_n1, os_n2;
  int mmod, nmod;
  int m_valid;
  int wgrad_direct;    // host-only: conv wgrad writes the parameter layout directly
  int wgrad_tap_on_a;  // MODE_WGRAD: the tap shift applies to operand A (conv wgrad) instead of B
  int wgrad_win;       // MODE_WGRAD over padded-image windows: unit u = filter rows (2u, 2u+1); the two 64-row A atoms
                       // are the 64-element windows of taps[2u] and taps[2u+1] (same channel origin, different rows)
  // TMA epilogue (epi_tma != 0): each epilogue warp stages its 32 rows in smem and issues bulk tensor stores /
  // reductions through map_out; out-of-bounds rows and columns are clipped by the TMA unit.
  //   1 = rows: out[row][col], col contiguous (activations NHWC, plain matrices); box = {epi_cols, 32 rows as a
  //       (epi_bw, epi_bh, 32/(epi_bw*epi_bh)) pixel sub-box}; bf16 or fp32; store or fp32 reduce-add (split-K)
  //   2 = transposed reduce (conv weight gradient): packed dW[tap][n][m] += acc[m][n], box = {32 m, 32 n, 1 tap}
  CUtensorMap map_out;
  int epi_tma, epi_reduce, epi_cols, epi_swz, epi_bufs;
  int epi_bw, epi_bh;    // mode 1: rows of the tile are pixels (w fastest, then h, then n) of an epi_bw x epi_bh x . box
  int epi_merge;         // mode 1, phase-merged transposed conv: 64-column box j goes to row parity j, channel 0
  int bias_mod;          // != 0 (a power of two): bias index = column % bias_mod (phase-merged columns repeat the channels)
  int epi_row_step;      // mode 1: coordinate-1 origin of tile m (plain matrices: 128), 0 for pixel tiles
  int phase_c[4], phase_p[4];  // mode 1, transposed-conv phases: channel-coordinate offset (pw * cb) and row parity ph
  // BatchNorm statistics fused into the FWD / epi_tma == 1 epilogue: per-channel shifted sums of the tile's VALID rows
  // (fp32 accumulator + bias, before the bf16 rounding), added to slot scratch [group][kBnSlots][2][stat_c]; the last
  // CTA to finish finalizes (dm_bn_fin.cuh)
  float* stat_out;           // nullptr = off (= stat_fin.scratch)
  const float* stat_shift;   // k[stat_c] (the layer's running_mean) or nullptr
  int stat_c;                // channels, a power of two; column j is channel j & (stat_c - 1) (phase-merged columns wrap)
  int stat_tiles_per_group;  // M tiles per stacked pass
  int stat_groups;
  dm_bn_fuse stat_fin;
#ifdef DM_STAMPS
  unsigned long long* stamps;  // debug build only: [cta][16] %globaltimer stamps of the kernel's phases
#endif
};

#ifdef DM_STAMPS
#define DM_STAMP(i)                                                      \
  do {                                                                   \
    if (p.stamps) {                                                      \
      unsigned long long t_;                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));             \
      p.stamps[blockIdx.x * 16 + (i)] = t_;                              \
    }                                                                    \
  } while (0)
#else
#define DM_STAMP(i)
#endif

constexpr int kThreads = 192;
constexpr int kAtomBytes = 8192;  // one MN-major atom: 64 k-rows x 128 B

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// v[i] of lane l = element (row l, column i) of a 32 x 32 block -> returns in v[0] of lane l the sum of COLUMN l over
// the 32 rows: a transpose-reduce butterfly, 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool hi = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = hi ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (hi ? v[i + s] : v[i]) + recv;
    }
  }
}

// One unit of work: a 128 x BN output tile and the k-block range that feeds it.
struct TileWork {
  int m_tile, n_tile, phase, split, unit_tap, tap_begin, kb0, kb1;
};

// t indexes tiles (cluster == 1) or tile PAIRS (cluster == 2; `rank` = this CTA's rank in the pair).
__device__ __forceinline__ TileWork decode_tile(const GemmParams& p, int t, int rank) {
  TileWork w;
  int total_kb;
  if (p.mode == MODE_FWD) {
    // order: m fastest, then n, then (phase, split): concurrently resident CTAs share one weight tile in L2
    const int m_slots = (p.cluster == 2) ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
    const int ms = t % m_slots;
    w.m_tile = (p.cluster == 2) ? 2 * ms + rank : ms;  // may be a phantom tile (== num_m_tiles): rows are masked
    int r = t / m_slots;
    w.n_tile = r % p.num_n_tiles;
    const int z = r / p.num_n_tiles;
    w.phase = z / p.num_splits;
    w.split = z - w.phase * p.num_splits;
    w.unit_tap = 0;
    w.tap_begin = p.phase_tap_start[w.phase];
    total_kb = (p.phase_tap_start[w.phase + 1] - w.tap_begin) * p.cpt;
  } else {
    // order: (tap unit, n) fastest, then m, then split: concurrent CTAs share the pixel tiles, write disjoint outputs
    const int u_slots = (p.cluster == 2) ? (p.num_units + 1) / 2 : p.num_units;
    const int units_n = u_slots * p.num_n_tiles;
    const int u = t % units_n;
    const int r = t / units_n;
    const int us = u / p.num_n_tiles;
    w.unit_tap = (p.cluster == 2) ? 2 * us + rank : us;  // may be a phantom unit (== num_units): stores are skipped
    w.n_tile = u - us * p.num_n_tiles;
    w.m_tile = r % p.num_m_tiles;
    w.split = r / p.num_m_tiles;
    w.phase = 0;
    w.tap_begin = 0;
    total_kb = p.num_kb;
  }
  const int kb_per_split = (total_kb + p.num_splits - 1) / p.num_splits;
  w.kb0 = w.split * kb_per_split;
  w.kb1 = min(total_kb, w.kb0 + kb_per_split);
  return w;
}

// kCG2 = true: CTA-pair instantiation (contains cta_group::2 instructions, so it MUST be launched with an even
// cluster dimension); kCG2 = false: single-CTA MMA (optionally with multicast pairs).
// kTF32 = true: fp32 operands read as TF32 (tcgen05.mma.kind::tf32, K = 8 per instruction): every k-block is 32 elements =
// 128 bytes per row, so the K-major byte layout is identical to bf16's; MN-major atoms are 32 elements x 32 k-rows (4 KB).
template <bool kCG2, bool kTF32 = false>
__global__ void __launch_bounds__(kThreads, 1) dm_tapgemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0) DM_STAMP(0);
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SWIZZLE_128B boxes need 1024-byte aligned stages

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = (p.cluster == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int t_begin = (p.cluster == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int t_step = (p.cluster == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  constexpr int kEsz = kTF32 ? 4 : 2;               // operand element size
  constexpr int kAtomW = kTF32 ? 32 : 64;            // MN elements of one 128-byte MN-major atom row
  constexpr int kAtomB = kTF32 ? 4096 : kAtomBytes;  // bytes of one MN-major atom (32 / 64 k-rows x 128 B)
  const int a_bytes = p.a_mn ? (128 / kAtomW) * kAtomB : 128 * p.kc * kEsz;
  const int b_bytes = p.b_mn ? max(1, p.bn / kAtomW) * kAtomB : (kCG2 ? (p.bn >> 1) : p.bn) * p.kc * kEsz;
  const int stage_bytes = a_bytes + b_bytes;

  uint8_t* staging = smem + p.stages * stage_bytes;  // epi_bufs x 4 warps x 4 KB, TMA epilogue only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (p.epi_tma ? p.epi_bufs * 16384 : 0));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* fold_buf = reinterpret_cast<float*>(tmem_ptr_smem + 4);  // [128][17] fp32, fold_kw epilogue only

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
  }
  if (kCG2) cluster_sync_all();  // both CTAs of the pair are resident before the paired TMEM allocation
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        // cg2: the leader's full barrier collects both CTAs' producers; every CTA's empty barrier is released by
        // the leader's multicast commit.  multicast pairs: each CTA's MMA thread releases the stage in both CTAs.
        mbar_init(&full_bar[s], kCG2 ? 2u : 1u);
        mbar_init(&empty_bar[s], kCG2 ? 1u : static_cast<uint32_t>(p.cluster));
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full_bar[b], 1);
        mbar_init(&tmem_empty_bar[b], kCG2 ? 256u : 128u);  // every epilogue thread (of both CTAs) arrives
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (kCG2)
      tmem_alloc_2sm(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
    else
      tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
  }
  tc_fence_before();
  if (p.cluster == 2)
    cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast at them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) DM_STAMP(1);
  // everything above (barriers, TMEM, descriptor prefetch) ran while the previous kernel of the stream was still
  // finishing; from here on its results are needed, and the next kernel may start its own prologue
  pdl_sync();
  if (threadIdx.x == 0) DM_STAMP(2);

  // Both single-thread roles below are ISSUE-bound if careless: one lane retires roughly one dependent instruction
  // every 4-6 cycles, and a 128x128x64 k-block is only 256 tensor-core cycles.  Everything that does not change per
  // k-block is therefore hoisted into registers (barrier addresses, descriptor templates, tap coordinates), the
  // k-block index is never divided, and the mode switches are resolved per tile, not per k-block.
  const uint32_t full0 = smem_u32(full_bar);
  const uint32_t empty0 = smem_u32(empty_bar);
  const uint32_t smem0 = smem_u32(smem);
  const int nstages = p.stages;

  if (warp == 0) {
    // =========================================================== TMA producer
    // The whole warp runs the (warp-uniform) loops so that addresses and coordinates live in uniform registers; one
    // elected lane issues the barrier arrivals and the TMA loads.
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t parity = 1;            // parity to wait for on the EMPTY barrier of `stage` (fresh barriers pass)
      uint32_t sa = smem0;            // smem address of the current stage
      const uint64_t map_a = reinterpret_cast<uint64_t>(&p.map_a);
      const uint64_t map_b = reinterpret_cast<uint64_t>(&p.map_b);
      const uint32_t tx_bytes = static_cast<uint32_t>(kCG2 ? 2 * stage_bytes : stage_bytes);
      // the barrier that receives the complete_tx of this CTA's loads: its own, or (cta_group::2) the leader's
      const uint32_t full_tx0 = kCG2 ? (full0 & kPeerBitMask) : full0;
      const int kc = p.kc;
#ifdef DM_STAMPS
      bool stamped_first = false;
#endif
      for (int t = t_begin; t < p.total_tiles; t += t_step) {
        const TileWork w = decode_tile(p, t, crank);
        int nkb = w.kb1 - w.kb0;
        if (nkb <= 0) continue;
        if (p.mode == MODE_FWD) {
          const int w0 = w.m_tile * p.tw_step;
          const int h0 = (w.m_tile % p.tpi) * p.th_step;
          const int n0 = (w.m_tile / p.tpi) * p.tn_step;
          const int ncol0 = w.n_tile * p.bn + (kCG2 ? crank * (p.bn >> 1) : 0);
          const int cpt = p.cpt;
          int tt = w.kb0 / cpt;            // per tile, not per k-block
          int chunk = w.kb0 - tt * cpt;
          const int b_atoms = p.b_mn ? (p.bn / kAtomW) : 0;
          while (nkb > 0) {
            const Tap tap = p.taps[w.tap_begin + tt];
            const int cw = w0 + tap.dw, cp = tap.dp, ch = h0 + tap.dh, wt = tap.wt;
            int c0 = tap.dc + chunk * kc;
            int kcol = chunk * kc;
            const int nch = min(cpt - chunk, nkb);
            for (int i = 0; i < nch; ++i) {
              const uint32_t fb = full_tx0 + 8u * stage;
              mbar_wait_u32(empty0 + 8u * stage, parity);
              if (leader) {
                if constexpr (kCG2) {
                  if (crank == 0)
                    mbar_arrive_expect_tx_u32(full0 + 8u * stage, tx_bytes);
                  else
                    mbar_arrive_remote_u32(full0 + 8u * stage, 0);
                  tma_load_5d_2sm_u32(sa, map_a, fb, c0, cw, cp, ch, n0);
                  tma_load_3d_2sm_u32(sa + a_bytes, map_b, fb, kcol, ncol0, wt);
#ifdef DM_STAMPS
                  if (!stamped_first) { DM_STAMP(3); stamped_first = true; }
#endif
                } else {
                  mbar_arrive_expect_tx_u32(fb, tx_bytes);
                  tma_load_5d_u32(sa, map_a, fb, c0, cw, cp, ch, n0);
                  if (b_atoms == 0) {
                    tma_load_3d_u32(sa + a_bytes, map_b, fb, kcol, ncol0, wt);
                  } else {
                    for (int a = 0; a < b_atoms; ++a)
                      tma_load_3d_u32(sa + a_bytes + a * kAtomB, map_b, fb, ncol0 + a * kAtomW, kcol, wt);
                  }
                }
#ifdef DM_STAMPS
                if (!stamped_first) { DM_STAMP(3); stamped_first = true; }
#endif
              }
              c0 += kc;
              kcol += kc;
              sa += stage_bytes;
              if (++stage == nstages) {
                stage = 0;
                parity ^= 1u;
                sa = smem0;
              }
            }
            nkb -= nch;
            chunk = 0;
            ++tt;
          }
        } else {
          const int ut = min(w.unit_tap, p.num_units - 1);
          const Tap tap = p.taps[p.wgrad_win ? 2 * ut : ut];
          const Tap tap2 = p.taps[p.wgrad_win ? 2 * ut + 1 : ut];  // second A atom (window mode only)
          const int mch0 = w.m_tile * 128;
          const int nch0 = w.n_tile * p.bn;
          const int b_atoms = max(1, p.bn / kAtomW);  // (bf16, bn = 32: one 64-wide box, upper half out of bounds = zeros)
          // pixel-tile coordinates of k-block kb: (w0, h0, n0) = (kb * tw_step, (kb % tpi) * th_step, (kb / tpi) *
          // tn_step), advanced incrementally
          int w0 = w.kb0 * p.tw_step;
          int ti = w.kb0 % p.tpi;
          int h0 = ti * p.th_step;
          int n0 = (w.kb0 / p.tpi) * p.tn_step;
          // the tap shift applies to A (conv wgrad) or to B (dense TN GEMM has a zero tap)
          const int adc = p.wgrad_tap_on_a ? tap.dc : 0, adw = p.wgrad_tap_on_a ? tap.dw : 0;
          const int adp = p.wgrad_tap_on_a ? tap.dp : 0, adh = p.wgrad_tap_on_a ? tap.dh : 0;
          const int bdc = p.wgrad_tap_on_a ? 0 : tap.dc, bdw = p.wgrad_tap_on_a ? 0 : tap.dw;
          const int bdp = p.wgrad_tap_on_a ? 0 : tap.dp, bdh = p.wgrad_tap_on_a ? 0 : tap.dh;
          for (; nkb > 0; --nkb) {
            const uint32_t fb = full0 + 8u * stage;
            mbar_wait_u32(empty0 + 8u * stage, parity);
            if (leader) {
              mbar_arrive_expect_tx_u32(fb, tx_bytes);
              tma_load_5d_u32(sa, map_a, fb, mch0 + adc, w0 + adw, adp, h0 + adh, n0);
              if (p.wgrad_win) {
                tma_load_5d_u32(sa + kAtomB, map_a, fb, tap2.dc, w0 + tap2.dw, tap2.dp, h0 + tap2.dh, n0);
              } else {
#pragma unroll
                for (int a = 1; a < 128 / kAtomW; ++a)
                  tma_load_5d_u32(sa + a * kAtomB, map_a, fb, mch0 + a * kAtomW + adc, w0 + adw, adp, h0 + adh, n0);
              }
              for (int a = 0; a < b_atoms; ++a)
                tma_load_5d_u32(sa + a_bytes + a * kAtomB, map_b, fb, nch0 + a * kAtomW + bdc, w0 + bdw, bdp,
                                h0 + bdh, n0);
#ifdef DM_STAMPS
              if (!stamped_first) { DM_STAMP(3); stamped_first = true; }
#endif
            }
            w0 += p.tw_step;
            h0 += p.th_step;
            if (++ti == p.tpi) {
              ti = 0;
              h0 = 0;
              n0 += p.tn_step;
            }
            sa += stage_bytes;
            if (++stage == nstages) {
              stage = 0;
              parity ^= 1u;
              sa = smem0;
            }
          }
        }
      }
      if (leader) DM_STAMP(11);
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (whole warp loops, one elected lane issues)
    if (!kCG2 || crank == 0) {  // cta_group::2: the leader CTA issues for the pair
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(kCG2 ? 256u : 128u, static_cast<uint32_t>(p.bn), p.a_mn, p.b_mn, kTF32);
      const uint32_t k_layout = (p.kc * kEsz == 128) ? 2u : 4u;       // SWIZZLE_128B : SWIZZLE_64B
      const uint32_t k_sbo = static_cast<uint32_t>(8 * p.kc * kEsz);  // 8 rows of one swizzle atom
      // descriptor templates (everything but the start address) and the per-UMMA_K advance of the start-address
      // field (address >> 4; smem addresses stay below 2^18, so the 14-bit field never carries)
      // MN-major: bf16 = SWIZZLE_128B atoms of 64 MN x 8 k-rows (SBO = 1024 B between 8-row groups); tf32 = the only
      // layout 32-bit MN-major operands may use, SWIZZLE_128B_BASE32B (layout type 1; TMA: SWIZZLE_128B_ATOM_32B):
      // atoms of 32 MN x 4 k-rows, 32-byte chunks XOR-ed with (row & 3) -> SBO = 512 B.  LBO = next MN atom (one box).
      const uint64_t da_t = p.a_mn ? make_smem_desc(0u, kAtomB, kTF32 ? 512u : 1024u, kTF32 ? 1u : 2u)
                                   : make_smem_desc(0u, 0u, k_sbo, k_layout);
      const uint64_t db_t = p.b_mn ? make_smem_desc(0u, kAtomB, kTF32 ? 512u : 1024u, kTF32 ? 1u : 2u)
                                   : make_smem_desc(0u, 0u, k_sbo, k_layout);
      // per UMMA_K (16 bf16 / 8 tf32 elements of K): 32 bytes along a K-major row, 16 / 8 k-rows of an MN-major atom
      const uint32_t a_step = (p.a_mn ? (kTF32 ? 1024u : 2048u) : 32u) >> 4;
      const uint32_t b_step = (p.b_mn ? (kTF32 ? 1024u : 2048u) : 32u) >> 4;
      const bool four = (p.kc * kEsz == 128);
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint32_t a16 = static_cast<uint32_t>(a_bytes) >> 4;
      const uint32_t smem16 = smem0 >> 4;
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      int stage = 0;
      uint32_t parity = 0;
      uint32_t s16 = smem16;
      int it = 0;
      for (int t = t_begin; t < p.total_tiles; t += t_step) {
        const TileWork w = decode_tile(p, t, crank);
        int nkb = w.kb1 - w.kb0;
        if (nkb <= 0) continue;
        const int buf = it & 1;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * p.acc_stride);
        mbar_wait_u32(tempty0 + 8u * buf, ((it >> 1) & 1) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        uint32_t acc = 0;
        for (; nkb > 0; --nkb) {
          mbar_wait_u32(full0 + 8u * stage, parity);
#ifdef DM_STAMPS
          if (leader && it == 0 && acc == 0) DM_STAMP(4);
#endif
          const uint64_t da = da_t | s16;
          const uint64_t db = db_t | (s16 + a16);
          if (!leader) {
          } else if constexpr (kCG2) {
            umma_bf16_2sm(tmem_d, da, db, idesc, acc);
            umma_bf16_2sm(tmem_d, da + a_step, db + b_step, idesc, 1u);
            if (four) {
              umma_bf16_2sm(tmem_d, da + 2 * a_step, db + 2 * b_step, idesc, 1u);
              umma_bf16_2sm(tmem_d, da + 3 * a_step, db + 3 * b_step, idesc, 1u);
            }
            umma_commit_2sm_mc_u32(empty0 + 8u * stage, 0x3);  // releases this stage in both CTAs of the pair
          } else if constexpr (kTF32) {
            umma_tf32(tmem_d, da, db, idesc, acc);
            umma_tf32(tmem_d, da + a_step, db + b_step, idesc, 1u);
            umma_tf32(tmem_d, da + 2 * a_step, db + 2 * b_step, idesc, 1u);
            umma_tf32(tmem_d, da + 3 * a_step, db + 3 * b_step, idesc, 1u);
            umma_commit_u32(empty0 + 8u * stage);
          } else {
            umma_bf16(tmem_d, da, db, idesc, acc);
            umma_bf16(tmem_d, da + a_step, db + b_step, idesc, 1u);
            if (four) {
              umma_bf16(tmem_d, da + 2 * a_step, db + 2 * b_step, idesc, 1u);
              umma_bf16(tmem_d, da + 3 * a_step, db + 3 * b_step, idesc, 1u);
            }
            umma_commit_u32(empty0 + 8u * stage);  // frees this smem stage once the MMAs above retire
          }
          acc = 1u;
          s16 += stage16;
          if (++stage == nstages) {
            stage = 0;
            parity ^= 1u;
            s16 = smem16;
          }
        }
        if (!leader) {
        } else if constexpr (kCG2) {
          umma_commit_2sm_mc_u32(tfull0 + 8u * buf, 0x3);  // each CTA's epilogue drains its own 128 TMEM lanes
        } else {
          umma_commit_u32(tfull0 + 8u * buf);  // accumulator complete -> epilogue
        }
        ++it;
      }
      if (leader) DM_STAMP(5);
    }
  } else {
    // =========================================================== epilogue (4 warps, 128 TMEM lanes)
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;  // row of the 128-row tile
    float* outf = reinterpret_cast<float*>(p.out);
    __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(p.out);
    int it = 0;
    int ebuf = 0;  // TMA epilogue: staging buffer of the next box
    // fused BatchNorm statistics: lane l holds this warp's running sums of column (32 * chunk + l) of the current
    // (group, N tile) run in registers (bn <= 128: at most 4 chunks, selected by predication -- no indexed register
    // array, no shared memory: the smem budget decides how many CTAs share an SM); flushed with one coalesced red.add
    // per 32 columns when the run ends
    float st_s0 = 0.f, st_s1 = 0.f, st_s2 = 0.f, st_s3 = 0.f, st_q0 = 0.f, st_q1 = 0.f, st_q2 = 0.f, st_q3 = 0.f;
    int stat_key = -1;
    auto stat_flush = [&](int key) {
      const int g = key / p.num_n_tiles, nt = key - g * p.num_n_tiles;
      float* dst = p.stat_out + (static_cast<long long>(g) * kBnSlots + ((blockIdx.x * 4 + q) % kBnSlots)) * 2 * p.stat_c;
      for (int ci = 0; ci * 32 < p.bn; ++ci) {
        const int col = nt * p.bn + ci * 32 + lane;
        const float vs = ci == 0 ? st_s0 : (ci == 1 ? st_s1 : (ci == 2 ? st_s2 : st_s3));
        const float vq = ci == 0 ? st_q0 : (ci == 1 ? st_q1 : (ci == 2 ? st_q2 : st_q3));
        if (col < p.n_valid) {
          atomicAdd(dst + (col & (p.stat_c - 1)), vs);
          atomicAdd(dst + p.stat_c + (col & (p.stat_c - 1)), vq);
        }
      }
      st_s0 = st_s1 = st_s2 = st_s3 = st_q0 = st_q1 = st_q2 = st_q3 = 0.f;
    };
    for (int t = t_begin; t < p.total_tiles; t += t_step) {
      const TileWork w = decode_tile(p, t, crank);
      if (w.kb0 >= w.kb1) continue;
      const int buf = it & 1;
      mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1);
      tc_fence_after();
#ifdef DM_STAMPS
      if (q == 0 && lane == 0) {
        if (it == 0) DM_STAMP(6);
        DM_STAMP(7);
      }
#endif
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * p.acc_stride);
      const int m_tile = w.m_tile, n_tile = w.n_tile;

      if (p.epi_tma == 1) {
        // ---- rows: this warp's 32 rows -> smem (swizzled like the store map) -> one bulk tensor store per box
        const uint64_t map_out = reinterpret_cast<uint64_t>(&p.map_out);
        const int rq = q * 32;
        const int cw = m_tile * p.epi_row_step + ((p.mode == MODE_FWD) ? m_tile * p.tw_step : 0) + rq % p.epi_bw;
        const int chh = ((p.mode == MODE_FWD) ? (m_tile % p.tpi) * p.th_step : 0) + (rq / p.epi_bw) % p.epi_bh;
        const int cn = ((p.mode == MODE_FWD) ? (m_tile / p.tpi) * p.tn_step : 0) + rq / (p.epi_bw * p.epi_bh);
        const int cph = (p.mode == MODE_FWD) ? p.phase_p[w.phase] : 0;
        const int cc0 = ((p.mode == MODE_FWD) ? p.phase_c[w.phase] : 0) + n_tile * p.bn;
        const bool add_bias = (p.bias != nullptr) && (w.split == 0);
        const int bias_mask = p.bias_mod ? p.bias_mod - 1 : 0x7fffffff;  // bias_mod is a power of two
        const uint32_t stg0 = smem_u32(staging) + static_cast<uint32_t>(q) * 4096u;
        const int row_bytes = p.epi_cols * (p.out_f32 ? 4 : 2);          // 128 or 64
        const uint32_t xr = (row_bytes == 128) ? (lane & 7) : ((lane >> 1) & 3);  // SWIZZLE_128B : SWIZZLE_64B
        const uint32_t row_addr = static_cast<uint32_t>(lane * row_bytes);
        int piece = 0;  // 16-byte piece index within the staging row
        int boxc = 0;   // first column (within the tile) of the box being filled
        bool stat_row_ok = false;
        if (p.stat_out) {
          const int g = min(m_tile / p.stat_tiles_per_group, p.stat_groups - 1);
          const int key = g * p.num_n_tiles + n_tile;
          if (key != stat_key) {
            if (stat_key >= 0) stat_flush(stat_key);
            stat_key = key;
          }
          const int rr = rq + lane;  // this thread's row of the tile
          const int wx = m_tile * p.tw_step + rr % p.bw;
          const int nn = (m_tile / p.tpi) * p.tn_step + rr / (p.bw * p.bh);
          stat_row_ok = (wx < p.w_lim) && (nn < p.n_lim);
        }
#ifdef DM_STAMPS
        long long ld_clk = 0, wait_clk = 0, pack_clk = 0, epi_t0 = clock64();
#endif
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t v[32];
          __syncwarp();  // tcgen05.ld is warp-collective
#ifdef DM_STAMPS
          const long long t_ld = clock64();
#endif
          tmem_ld_32x32(lane_addr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
#ifdef DM_STAMPS
          ld_clk += clock64() - t_ld;
#endif
          if (c0 + 32 >= p.bn) {
            // all of this thread's TMEM reads of the tile are done: hand the accumulator back before the stores
            tc_fence_before();
            if (kCG2)
              mbar_arrive_remote(&tmem_empty_bar[buf], 0);
            else
              mbar_arrive(&tmem_empty_bar[buf]);
          }
          if (add_bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n_tile * p.bn + c0 + i < p.n_valid) {
                const int bi = n_tile * p.bn + c0 + i;
                v[i] = __float_as_uint(__uint_as_float(v[i]) + __ldg(p.bias + (bi & bias_mask)));
              }
          }
          if (p.stat_out) {
            float d[32], d2[32];
            const int smask = p.stat_c - 1;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = n_tile * p.bn + c0 + i;
              const float kk = p.stat_shift ? __ldg(p.stat_shift + (col & smask)) : 0.f;
              const float dv = (stat_row_ok && col < p.n_valid) ? __uint_as_float(v[i]) - kk : 0.f;
              d[i] = dv;
              d2[i] = dv * dv;
            }
            warp_column_sums(d, lane);
            warp_column_sums(d2, lane);
            const int ci = c0 >> 5;
            st_s0 += ci == 0 ? d[0] : 0.f; st_q0 += ci == 0 ? d2[0] : 0.f;
            st_s1 += ci == 1 ? d[0] : 0.f; st_q1 += ci == 1 ? d2[0] : 0.f;
            st_s2 += ci == 2 ? d[0] : 0.f; st_q2 += ci == 2 ? d2[0] : 0.f;
            st_s3 += ci == 3 ? d[0] : 0.f; st_q3 += ci == 3 ? d2[0] : 0.f;
          }
#ifdef DM_STAMPS
          const long long t_w = clock64();
#endif
          if (piece == 0) {
            // the staging buffer about to be filled must have been read by the store issued epi_bufs boxes ago
            if (lane == 0) {
              if (p.epi_bufs == 4) bulk_wait_group_read<3>();
              else if (p.epi_bufs == 2) bulk_wait_group_read<1>();
              else bulk_wait_group_read<0>();
            }
            __syncwarp();
          }
#ifdef DM_STAMPS
          wait_clk += clock64() - t_w;
          const long long t_p = clock64();
#endif
          const uint32_t stg = stg0 + static_cast<uint32_t>(ebuf) * 16384u + row_addr;
          const int ncol = min(32, p.bn - c0);
          if (p.out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < ncol)
                st_shared_v4(stg + (((piece + j) ^ xr) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            piece += ncol >> 2;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (8 * j < ncol)
                st_shared_v4(stg + (((piece + j) ^ xr) << 4),
                             pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
            piece += ncol >> 3;
          }
#ifdef DM_STAMPS
          pack_clk += clock64() - t_p;
#endif
          if (piece * 16 >= row_bytes || c0 + 32 >= p.bn) {
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the bulk (async proxy) store
            __syncwarp();
            if (lane == 0) {
              const uint32_t src = stg0 + static_cast<uint32_t>(ebuf) * 16384u;
              if (p.epi_reduce)
                tma_reduce_add_5d(map_out, src, cc0 + boxc, cw, cph, chh, cn);
              else if (p.epi_merge)
                tma_store_5d(map_out, src, 0, cw, boxc >> 6, chh, cn);
              else
                tma_store_5d(map_out, src, cc0 + boxc, cw, cph, chh, cn);
              bulk_commit_group();
            }
            piece = 0;
            boxc = c0 + 32;
            if (++ebuf == p.epi_bufs) ebuf = 0;
          }
        }
#ifdef DM_STAMPS
        if (q == 0 && lane == 0 && p.stamps) {
          p.stamps[blockIdx.x * 16 + 13] = static_cast<unsigned long long>(ld_clk);
          p.stamps[blockIdx.x * 16 + 15] = static_cast<unsigned long long>(wait_clk) * 100000ull + static_cast<unsigned long long>(pack_clk);
          p.stamps[blockIdx.x * 16 + 14] = static_cast<unsigned long long>(clock64() - epi_t0);
        }
#endif
      } else if (p.epi_tma == 2) {
        // ---- conv weight gradient: acc[m = cb row][n = cs col] -> packed dW[tap][n][m]; staging [32 n][32 m] fp32
        const uint64_t map_out = reinterpret_cast<uint64_t>(&p.map_out);
        const Tap tap = p.taps[p.wgrad_win ? 2 * min(w.unit_tap, p.num_units - 1) : min(w.unit_tap, p.num_units - 1)];
        const bool pair = p.mmod == 32;  // rows 0-31 / 32-63 belong to taps out_tap / out_tap + 1
        const bool pair64 = p.mmod == 64;  // window mode: rows 0-63 / 64-127 belong to filter rows out_tap / out_tap + 1
        const int m0 = pair ? 0 : (pair64 ? (q & 1) * 32 : m_tile * 128 + q * 32);
        const int otap = tap.out_tap + (pair ? q : (pair64 ? (q >> 1) : 0));
        const bool warp_ok = (w.unit_tap < p.num_units) && (q * 32 < tap.mvalid) && (m_tile * 128 + q * 32 < p.m_valid);
        const uint32_t stg0 = smem_u32(staging) + static_cast<uint32_t>(q) * 4096u;
        const int ncols = min(p.bn, static_cast<int>(tap.nvalid) - n_tile * p.bn);
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32(lane_addr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (c0 + 32 >= p.bn) {
            // all of this thread's TMEM reads of the tile are done: hand the accumulator back before the stores
            tc_fence_before();
            if (kCG2)
              mbar_arrive_remote(&tmem_empty_bar[buf], 0);
            else
              mbar_arrive(&tmem_empty_bar[buf]);
          }
          if (!warp_ok || c0 >= ncols) continue;
          if (lane == 0) {
            if (p.epi_bufs == 4) bulk_wait_group_read<3>();
            else if (p.epi_bufs == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
          }
          __syncwarp();
          if (p.out_f32) {
            const uint32_t stg = stg0 + static_cast<uint32_t>(ebuf) * 16384u + static_cast<uint32_t>(lane) * 4u;
#pragma unroll
            for (int i = 0; i < 32; ++i) st_shared_b32(stg + i * 128, v[i]);
          } else {  // bf16 output: staging rows of 32 m x 2 bytes
            const uint32_t stg = stg0 + static_cast<uint32_t>(ebuf) * 16384u + static_cast<uint32_t>(lane) * 2u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const __nv_bfloat16 h = __float2bfloat16_rn(__uint_as_float(v[i]));
              asm volatile("st.shared.b16 [%0], %1;" ::"r"(stg + i * 64), "h"(*reinterpret_cast<const uint16_t*>(&h)) : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.epi_reduce)
              tma_reduce_add_3d(map_out, stg0 + static_cast<uint32_t>(ebuf) * 16384u, m0, n_tile * p.bn + c0, otap);
            else
              tma_store_3d(map_out, stg0 + static_cast<uint32_t>(ebuf) * 16384u, m0, n_tile * p.bn + c0, otap);
            bulk_commit_group();
          }
          if (++ebuf == p.epi_bufs) ebuf = 0;
        }
      } else if (p.mode == MODE_FWD && p.fold_kw) {
        // tile = bh full image rows of bw pixels; thread r = pixel (h, w').  acc[w'][kw*3+cb] holds the sum over kh
        // and cs for UNSHIFTED w'; out[h][w][cb] = sum_kw acc[h][w + 2 - kw][kw*3 + cb] (zero outside the row).
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 15; ++i) fold_buf[r * 17 + i] = __uint_as_float(v[i]);
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
        const int wq = r % p.bw, hq = r / p.bw;
        const int h = (m_tile % p.tpi) * p.th_step + hq % p.bh;
        const int n = (m_tile / p.tpi) * p.tn_step + hq / p.bh;
        float o[3];
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) o[cb] = p.bias ? __ldg(p.bias + cb) : 0.f;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const int ws = wq + 2 - kw;
          if (ws >= 0 && ws < p.bw) {
            const float* src = fold_buf + (hq * p.bw + ws) * 17 + kw * 3;
            o[0] += src[0];
            o[1] += src[1];
            o[2] += src[2];
          }
        }
        if (n < p.n_lim) {
          const long long off = wq * p.os_w + h * p.os_h + n * p.os_n;
          outf[off] = o[0];
          outf[off + 1] = o[1];
          outf[off + 2] = o[2];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // fold_buf is reused by the next tile
      } else if (p.mode == MODE_FWD) {
        const int wx = m_tile * p.tw_step + r % p.bw;
        const int h = (m_tile % p.tpi) * p.th_step + (r / p.bw) % p.bh;
        const int n = (m_tile / p.tpi) * p.tn_step + r / (p.bw * p.bh);
        const bool row_ok = (wx < p.w_lim) && (n < p.n_lim);
        const long long row_off = wx * p.os_w + h * p.os_h + n * p.os_n + p.phase_out_off[w.phase];
        const bool add_bias = (p.bias != nullptr) && (w.split == 0);
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t v[32];
          __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the divergent stores
          tmem_ld_32x32(lane_addr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          const int ng = n_tile * p.bn + c0;  // first global column of this chunk
          if (!row_ok || ng >= p.n_valid) continue;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] = __uint_as_float(v[i]);
            if (add_bias && ng + i < p.n_valid) f[i] += __ldg(p.bias + ng + i);
          }
          const bool full_chunk = (ng + 32 <= p.n_valid) && (c0 + 32 <= p.bn) && (p.os_col == 1);
          if (p.out_atomic) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.bn && ng + i < p.n_valid) atomicAdd(outf + row_off + (ng + i) * p.os_col, f[i]);
          } else if (p.out_f32) {
            if (full_chunk && ((row_off + ng) & 3) == 0) {
              float4* dst = reinterpret_cast<float4*>(outf + row_off + ng);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.bn && ng + i < p.n_valid) outf[row_off + (ng + i) * p.os_col] = f[i];
            }
          } else {
            if (full_chunk && ((row_off + ng) & 7) == 0) {
              uint4* dst = reinterpret_cast<uint4*>(outh + row_off + ng);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(pack_bf16x2(f[8 * i], f[8 * i + 1]), pack_bf16x2(f[8 * i + 2], f[8 * i + 3]),
                                    pack_bf16x2(f[8 * i + 4], f[8 * i + 5]), pack_bf16x2(f[8 * i + 6], f[8 * i + 7]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.bn && ng + i < p.n_valid)
                  outh[row_off + (ng + i) * p.os_col] = __float2bfloat16_rn(f[i]);
            }
          }
        }
      } else {
        const Tap tap = p.taps[min(w.unit_tap, p.num_units - 1)];
        const int m = m_tile * 128 + r;
        const bool row_ok = (m < p.m_valid) && (m < tap.mvalid) && (w.unit_tap < p.num_units);
        const long long row_off = (m % p.mmod) * p.os_m + (m / p.mmod) * p.os_m2 + tap.out_off;
        const bool simple_n = p.nmod >= (1 << 30);  // column offset is affine: no div/mod per element
        const int ncols = min(p.bn, static_cast<int>(tap.nvalid) - n_tile * p.bn);
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t v[32];
          __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the divergent stores
          tmem_ld_32x32(lane_addr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (!row_ok || c0 >= ncols) continue;
          if (simple_n) {
            float* dst = outf + row_off + static_cast<long long>(n_tile * p.bn + c0) * p.os_n1;
            const int lim = min(32, ncols - c0);
            if (p.out_atomic) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < lim) atomicAdd(dst + i * p.os_n1, __uint_as_float(v[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < lim) dst[i * p.os_n1] = __uint_as_float(v[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int nl = n_tile * p.bn + c0 + i;  // column within this tap unit
              if (c0 + i < ncols) {
                const long long off = row_off + (nl % p.nmod) * p.os_n1 + (nl / p.nmod) * p.os_n2;
                if (p.out_atomic)
                  atomicAdd(outf + off, __uint_as_float(v[i]));
                else
                  outf[off] = __uint_as_float(v[i]);
              }
            }
          }
        }
      }
      // this thread's TMEM reads of the accumulator are complete: hand the buffer back to the MMA issuer
      // (the TMA epilogues did that already, right after their last tcgen05.ld)
      __syncwarp();
      if (!p.epi_tma) {
        tc_fence_before();
        if (kCG2)
          mbar_arrive_remote(&tmem_empty_bar[buf], 0);  // the leader's MMA thread waits for both CTAs' epilogues
        else
          mbar_arrive(&tmem_empty_bar[buf]);
      }
      ++it;
    }
    if (q == 0 && lane == 0) DM_STAMP(8);
    if (p.stat_out) {
      // publish this CTA's partial sums, take a ticket; the LAST CTA of the grid finalizes the BatchNorm (constants for
      // the consumer kernel, running statistics) and re-zeroes the scratch -- no finalize kernel
      if (stat_key >= 0) stat_flush(stat_key);
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
      if (q == 0 && lane == 0) {
        const unsigned int tk = atomicAdd(bn_ticket(p.stat_out, p.stat_c, p.stat_groups), 1u);
        tmem_ptr_smem[1] = (tk == gridDim.x - 1u) ? 1u : 0u;
        __threadfence();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tmem_ptr_smem[1] != 0u) bn_forward_finalize(p.stat_fin, p.stat_c, q * 32 + lane, 128);
    }
    if (q == 0 && lane == 0) DM_STAMP(9);
    if (p.epi_tma && lane == 0) bulk_wait_group_all();  // smem must outlive the stores' reads; writes complete
    if (q == 0 && lane == 0) DM_STAMP(10);
  }

  tc_fence_before();
  if (p.cluster == 2)
    cluster_sync_all();  // the peer may still multicast into this CTA's smem / arrive on its barriers
  else
    __syncthreads();
  if (warp == 1) {
    if constexpr (kCG2)
      tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else
      tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  if (threadIdx.x == 0) DM_STAMP(12);
}

// ================================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

constexpr int kSwz128Atom32 = 129;

// rank-`rank` bf16 tensor map; dims/box innermost first; strides in BYTES for dims 1..rank-1.
static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle_bytes, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(-2, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
  // swizzle_bytes: 128 / 64 / 0, or kSwz128Atom32 = 128-byte span with 32-byte chunks (MN-major fp32 operands)
  CUtensorMapSwizzle sw = swizzle_bytes == kSwz128Atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                          : swizzle_bytes == 128         ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), gdim,
                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(static_cast<int>(r),
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                     static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                     rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  }
  return 0;
}

// NHWC bf16 activation [b,h,w,c] as the 5-D view (c, w, p, h, n); stride 2 -> parity-split view.
static int encode_act_map(CUtensorMap* m, const void* ptr, int b, int h, int w, int c, int stride, const uint32_t* box,
                          int swizzle_bytes, bool f32 = false) {
  uint64_t dims[5], str[4];
  const uint64_t e = f32 ? 4 : 2;
  if (stride == 1) {
    dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = b;
    str[0] = c * e; str[1] = (uint64_t)w * c * e; str[2] = (uint64_t)w * c * e; str[3] = (uint64_t)h * w * c * e;
  } else {
    dims[0] = 2 * c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = b;
    str[0] = 2 * c * e; str[1] = (uint64_t)w * c * e; str[2] = 2ull * w * c * e; str[3] = (uint64_t)h * w * c * e;
  }
  return encode_map(m, ptr, 5, dims, str, box, swizzle_bytes, f32);
}

// Row-major bf16 (fp32) matrix [rows, cols] (leading dimension ld) as the 5-D view (c=cols, w=rows, 1, 1, 1).
static int encode_mat_map5(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld,
                           uint32_t box_c, uint32_t box_r, int swizzle_bytes, bool f32 = false) {
  uint64_t dims[5] = {(uint64_t)cols, (uint64_t)rows, 1, 1, 1};
  const uint64_t rs = (uint64_t)ld * (f32 ? 4 : 2);
  uint64_t str[4] = {rs, rs * rows, rs * rows, rs * rows};
  uint32_t box[5] = {box_c, box_r, 1, 1, 1};
  return encode_map(m, ptr, 5, dims, str, box, swizzle_bytes, f32);
}

// bf16 (fp32) [t][rows][cols] as 3-D (cols, rows, t)
static int encode_w_map3(CUtensorMap* m, const void* ptr, long long t, long long rows, long long cols, long long ld,
                         uint32_t box_c, uint32_t box_r, int swizzle_bytes, bool f32 = false) {
  const uint64_t e = f32 ? 4 : 2;
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)rows, (uint64_t)t};
  uint64_t str[2] = {(uint64_t)ld * e, (uint64_t)ld * e * rows};
  uint32_t box[3] = {box_c, box_r, 1};
  return encode_map(m, ptr, 3, dims, str, box, swizzle_bytes, f32);
}

#ifdef DM_STAMPS
static unsigned long long* g_stamps = nullptr;
extern "C" int dm_debug_set_stamps(void* dev_buf) {
  g_stamps = static_cast<unsigned long long*>(dev_buf);
  return 0;
}
#endif
static int g_last_grid[3] = {0, 0, 0};
static int g_last_smem = 0, g_last_stages = 0;

// Optional per-launch timing of the GEMM-class kernel (bench.py's roofline): CUDA event pairs around every
// launch on the launching stream, read back (after a device sync) by dm_profile_read().
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops;
  int gx, gy, gz, mode, bn, kc, stages, ctas, a_mn, b_mn, kind;
};
// what the next launch computes (profile records only): 0 dense GEMM, 1 conv_down, 2 conv_up, 3 conv_wgrad,
// 4 = the 3-image-channel layers (window GEMMs over the padded image: bound by TMA rows / HBM, not by the tensor pipe)
static thread_local int t_kind = 0;
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<ProfRec> g_prof_pool;

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// `grid` = logical tile space: x = m tiles, y = n tiles (FWD) or n tiles * tap units (WGRAD), z = phases*splits
// (FWD) or splits (WGRAD).  The kernel is persistent: min(tiles, SMs * CTAs/SM) CTAs grid-stride over the tiles,
// each with a double-buffered TMEM accumulator so that a tile's epilogue overlaps the next tile's main loop.
static int launch(GemmParams& p, dim3 grid, cudaStream_t stream, double flops, int kb_per_cta = 1 << 30,
                  int cluster = 1) {
  if (cluster != 2 || p.mode != MODE_FWD || p.b_mn) p.cg2 = 0;
  if (p.tf32) p.cg2 = 0;
  const int esz = p.tf32 ? 4 : 2, atom_w = p.tf32 ? 32 : 64, atom_b = p.tf32 ? 4096 : kAtomBytes;
  const int a_bytes = p.a_mn ? (128 / atom_w) * atom_b : 128 * p.kc * esz;
  const int b_bytes = p.b_mn ? std::max(1, p.bn / atom_w) * atom_b : (p.cg2 ? (p.bn >> 1) : p.bn) * p.kc * esz;
  const int stage_bytes = a_bytes + b_bytes;
  p.cluster = cluster;
  p.num_m_tiles = grid.x;
  if (p.mode == MODE_WGRAD) p.num_units = grid.y / p.num_n_tiles;
  // work items: tiles, or tile pairs when CTA pairs share the B operand
  if (cluster == 2) {
    if (p.mode == MODE_FWD)
      p.total_tiles = static_cast<int>(((grid.x + 1) / 2) * grid.y * grid.z);
    else
      p.total_tiles = static_cast<int>(grid.x * ((p.num_units + 1) / 2) * p.num_n_tiles * grid.z);
  } else {
    p.total_tiles = static_cast<int>(grid.x * grid.y * grid.z);
  }
  p.acc_stride = (p.bn + 31) / 32 * 32;
  p.tmem_cols = pow2_cols(2 * p.acc_stride);
  const int sms = num_sms();
  // two CTAs per SM only if both their TMEM (2 x <=256 columns) and their smem rings fit
  // (and only if there are more tiles than SMs: otherwise one CTA per SM with a deeper ring hides more latency)
  const bool two_per_sm = p.tmem_cols <= 256 && stage_bytes <= 32768 && env_int("DM_ONE_CTA", 0) == 0 &&
                          p.total_tiles * cluster > sms;
  // per-CTA ring budget: two CTAs per SM share 228 KB (1 KB of each is reserved); the TMA epilogue stages 16 KB (x2)
  const int budget = two_per_sm ? env_int("DM_SMEM_BUDGET_SMALL", 98304)
                                : env_int("DM_SMEM_BUDGET_BIG", 196608) - (p.epi_tma ? 16384 : 0);
  const int slots = sms * (two_per_sm ? 2 : 1) / cluster;  // concurrently resident CTAs (or CTA pairs)
  const int tiles_per_cta = (p.total_tiles + slots - 1) / slots;
  const long long kb_stream = static_cast<long long>(std::min(kb_per_cta, 1 << 20)) * tiles_per_cta;
  int stages = std::max(2, std::min(8, budget / stage_bytes));
  stages = static_cast<int>(std::max<long long>(1, std::min<long long>(stages, kb_stream)));
  p.stages = stages;
  const int fixed = (2 * stages + 4) * 8 + 16 + (p.fold_kw ? 128 * 17 * 4 : 0);
  p.epi_bufs = 0;
  if (p.epi_tma) {
    // per-CTA smem: 227 KB alone, or half of (228 KB - 2 x 1 KB reserved) when two CTAs share the SM
    const int cap = two_per_sm ? (228 * 1024 - 2048) / 2 : 227 * 1024;
    p.epi_bufs = 1;
    for (int nb = 4; nb >= 2; nb >>= 1)
      if (stages * stage_bytes + nb * 16384 + fixed <= cap) { p.epi_bufs = nb; break; }
  }
  int smem = stages * stage_bytes + p.epi_bufs * 16384 + fixed;
  // TMEM is 512 columns per SM: keep co-residency at <= 512 / tmem_cols CTAs by padding the smem request
  smem = std::max(smem, (two_per_sm ? 80 : 120) * 1024);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(dm_tapgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dm_tapgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dm_tapgemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int ctas = cluster * std::min(p.total_tiles, slots);
  g_last_grid[0] = grid.x; g_last_grid[1] = grid.y; g_last_grid[2] = grid.z;
  g_last_smem = smem; g_last_stages = stages;
  ProfRec rec;
  bool prof = false;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof = g_prof_on;
    if (prof) {
      if (!g_prof_pool.empty()) {
        rec = g_prof_pool.back();
        g_prof_pool.pop_back();
      } else {
        cudaEventCreate(&rec.e0);
        cudaEventCreate(&rec.e1);
      }
      rec.flops = flops;
      rec.gx = grid.x; rec.gy = grid.y; rec.gz = grid.z; rec.mode = p.mode; rec.bn = p.bn; rec.kc = p.kc;
      rec.stages = stages; rec.ctas = ctas; rec.a_mn = p.a_mn; rec.b_mn = p.b_mn; rec.kind = t_kind;
    }
  }
  if (prof) cudaEventRecord(rec.e0, stream);
#ifdef DM_STAMPS
  p.stamps = g_stamps;
#endif
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (cluster > 1) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = cluster;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  if (pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  cudaError_t le = p.tf32 ? cudaLaunchKernelEx(&cfg, dm_tapgemm_kernel<false, true>, p)
                          : (p.cg2 ? cudaLaunchKernelEx(&cfg, dm_tapgemm_kernel<true>, p)
                                   : cudaLaunchKernelEx(&cfg, dm_tapgemm_kernel<false>, p));
  if (prof) {
    cudaEventRecord(rec.e1, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (le != cudaSuccess) return set_error((int)le, "dm_tapgemm_kernel launch: %s", cudaGetErrorString(le));
  return check_launch("dm_tapgemm_kernel");
}

// pixel tile of P pixels on an (h, w) grid: full-width rows
struct PixTile {
  int bw, bh, bimg, tpi, th_step, tn_step, tiles;
};
static bool make_pix_tile(int P, int batch, int h, int w, PixTile* t) {
  if (w <= 0 || h <= 0 || P % w != 0) return false;
  t->bw = w;
  if (w * h >= P) {
    t->bh = P / w;
    if (h % t->bh != 0) return false;
    t->bimg = 1;
    t->tpi = h / t->bh;
    t->th_step = t->bh;
    t->tn_step = 1;
    t->tiles = batch * t->tpi;
  } else {
    if (P % (w * h) != 0) return false;
    t->bh = h;
    t->bimg = P / (w * h);
    t->tpi = 1;
    t->th_step = 0;
    t->tn_step = t->bimg;
    t->tiles = (batch + t->bimg - 1) / t->bimg;
  }
  return true;
}


// ---- TMA epilogue set-up (GemmParams::epi_tma).  Returns false (legacy thread-store epilogue stays) when the
// output cannot be described by a box store: odd leading dimensions, N tiles that are not whole boxes, ...
static bool epi_common(GemmParams& p, const void* out, bool f32) {
  if (env_int("DM_EPI_TMA", 1) == 0) return false;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return false;
  const int full = f32 ? 32 : 64;           // columns of a 128-byte staging row
  int cols = (p.bn % full == 0) ? full : ((p.bn % (full / 2) == 0 && p.bn == full / 2) ? full / 2 : 0);
  if (cols == 0 || p.bn % 32 != 0) return false;
  p.epi_cols = cols;
  p.epi_swz = cols * (f32 ? 4 : 2);         // 128 or 64
  return true;
}

// sub-box of 32 consecutive tile rows for a tile of bw x bh x bimg pixels (w fastest)
static bool epi_sub_box(int bw, int bh, int bimg, uint32_t* sub) {
  const int sw = std::min(bw, 32);
  if (32 % sw != 0 || bw % sw != 0) return false;
  const int sh = std::min(bh, 32 / sw);
  if ((32 / sw) % sh != 0 || bh % sh != 0) return false;
  const int sn = 32 / (sw * sh);
  if (sn > bimg || bimg % sn != 0) return false;
  sub[0] = sw; sub[1] = sh; sub[2] = sn;
  return true;
}

// plain row-major matrix out[rows][cols] (leading dimension ld elements)
static bool epi_rows_matrix(GemmParams& p, void* out, bool f32, long long rows, long long cols, long long ld,
                            bool reduce, int row_step) {
  if (!epi_common(p, out, f32)) return false;
  const uint64_t es = f32 ? 4 : 2;
  if ((ld * es) % 16 != 0) return false;
  uint64_t dims[5] = {(uint64_t)cols, (uint64_t)rows, 1, 1, 1};
  const uint64_t rs = (uint64_t)ld * es;
  uint64_t str[4] = {rs, rs * rows, rs * rows, rs * rows};
  uint32_t box[5] = {(uint32_t)p.epi_cols, 32, 1, 1, 1};
  if (encode_map(&p.map_out, out, 5, dims, str, box, p.epi_swz, f32) != 0) return false;
  p.epi_tma = 1; p.epi_reduce = reduce ? 1 : 0;
  p.epi_bw = 128; p.epi_bh = 1; p.epi_row_step = row_step;
  return true;
}

// NHWC activation out[b][h][w][c]; stride 2 = the four sub-pixel phases of a transposed convolution write the
// parity-split view (2c, w/2, 2, h/2, b) at channel offset pw*c, parity ph
static bool epi_rows_act(GemmParams& p, void* out, bool f32, int b, int h, int w, int c, int stride, const PixTile& pt) {
  if (!epi_common(p, out, f32)) return false;
  if (c % p.bn != 0) return false;  // a tile's columns must not spill into the other parity's channels
  uint32_t sub[3];
  if (!epi_sub_box(pt.bw, pt.bh, pt.bimg, sub)) return false;
  const uint64_t e = f32 ? 4 : 2;
  if ((c * e) % 16 != 0) return false;
  uint64_t dims[5], str[4];
  if (stride == 1) {
    dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = b;
    str[0] = c * e; str[1] = (uint64_t)w * c * e; str[2] = (uint64_t)w * c * e; str[3] = (uint64_t)h * w * c * e;
  } else {
    dims[0] = 2 * c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = b;
    str[0] = 2 * c * e; str[1] = (uint64_t)w * c * e; str[2] = 2ull * w * c * e; str[3] = (uint64_t)h * w * c * e;
  }
  uint32_t box[5] = {(uint32_t)p.epi_cols, sub[0], 1, sub[1], sub[2]};
  if (encode_map(&p.map_out, out, 5, dims, str, box, p.epi_swz, f32) != 0) return false;
  p.epi_tma = 1; p.epi_reduce = 0;
  p.epi_bw = pt.bw; p.epi_bh = pt.bh; p.epi_row_step = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      p.phase_c[ph * 2 + pw] = (stride == 2) ? pw * c : 0;
      p.phase_p[ph * 2 + pw] = (stride == 2) ? ph : 0;
    }
  return true;
}

// packed conv weight gradient dW[25][cs][cb] fp32 (cb contiguous): transposed reduce-add of [m = cb][n = cs] tiles
static bool epi_wgrad_packed(GemmParams& p, float* dw, int cs, int cb) {
  if (env_int("DM_EPI_TMA", 1) == 0) return false;
  if ((reinterpret_cast<uintptr_t>(dw) & 15) != 0 || cb % 32 != 0 || cs % 32 != 0) return false;
  uint64_t dims[3] = {(uint64_t)cb, (uint64_t)cs, 25};
  uint64_t str[2] = {(uint64_t)cb * 4, (uint64_t)cb * cs * 4};
  uint32_t box[3] = {32, 32, 1};
  if (encode_map(&p.map_out, dw, 3, dims, str, box, 0, true) != 0) return false;
  p.epi_tma = 2; p.epi_reduce = 1;
  return true;
}

// dense out[n][m] fp32 with m contiguous (Linear weight gradient: rows of the accumulator are in-features)
static bool epi_transposed_matrix(GemmParams& p, void* out, bool f32, long long m, long long n, long long ld, bool reduce) {
  if (env_int("DM_EPI_TMA", 1) == 0 && f32) return false;
  const uint64_t es = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0 || (ld * es) % 16 != 0 || p.bn % 32 != 0 || n % 32 != 0) return false;
  if (!f32 && reduce) return false;
  uint64_t dims[3] = {(uint64_t)m, (uint64_t)n, 1};
  uint64_t str[2] = {(uint64_t)ld * es, (uint64_t)ld * es * n};
  uint32_t box[3] = {32, 32, 1};
  if (encode_map(&p.map_out, out, 3, dims, str, box, 0, f32) != 0) return false;
  p.epi_tma = 2; p.epi_reduce = reduce ? 1 : 0;
  return true;
}

// Attach fused BatchNorm statistics (dm_bn_fuse) to a FWD launch whose epilogue is the row-store TMA path.
// m_tiles = real M tiles; rows of a group must fill whole tiles.
static int attach_stats(GemmParams& p, const dm_bn_fuse* bn, int channels, int m_tiles, const char* who) {
  if (bn == nullptr || bn->scratch == nullptr) return 0;
  DM_REQUIRE(bn->rows > 0 && bn->gamma && bn->beta && bn->scale_shift && bn->mean_invstd, "%s: incomplete dm_bn_fuse", who);
  DM_REQUIRE(p.mode == MODE_FWD && p.epi_tma == 1 && p.num_splits == 1 && !p.fold_kw,
             "%s: fused BatchNorm statistics need the row-store epilogue without split-K", who);
  DM_REQUIRE(bn->groups >= 1 && m_tiles % bn->groups == 0, "%s: %d M tiles do not split into %d groups", who, m_tiles,
             bn->groups);
  DM_REQUIRE(channels > 0 && (channels & (channels - 1)) == 0, "%s: fused statistics need a power-of-two channel count", who);
  DM_REQUIRE(p.bn <= 128, "%s: fused statistics support N tiles up to 128 columns", who);
  p.stat_out = bn->scratch;
  p.stat_shift = bn->running_mean;
  p.stat_fin = *bn;
  p.stat_c = channels;
  p.stat_groups = bn->groups;
  p.stat_tiles_per_group = m_tiles / bn->groups;
  return 0;
}

static void init_params(GemmParams& p) {
  memset(&p, 0, sizeof(p));
  p.num_splits = 1;
  p.num_n_tiles = 1;
  p.cpt = 1;
  p.nmod = 1 << 30;
  p.mmod = 1 << 30;
  p.os_col = 1;
  p.tpi = 1 << 30;
}

static int pick_bn(int n, int cap) {
  // largest multiple of 16 <= cap that divides n rounded up to 16
  int n16 = (n + 15) / 16 * 16;
  if (n16 <= cap) return n16;
  for (int b = cap; b >= 16; b -= 16)
    if (n16 % b == 0) return b;
  return cap;
}

// Same, but halve the N tile (down to 64) while the launch would leave a quarter of the SMs without a tile:
// `work` = number of 128-row tiles x phases that share one N tile.
static int pick_bn_fill(int n, int cap, long long work) {
  int bn = pick_bn(n, cap);
  const int n16 = (n + 15) / 16 * 16;
  while (bn >= 128 && (bn / 2) % 16 == 0 && n16 % (bn / 2) == 0 && work * ((n16 + bn - 1) / bn) < (num_sms() * 3) / 4)
    bn /= 2;
  return bn;
}

}  // namespace dm

using namespace dm;

extern "C" int dm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}

// Sum of per-launch durations (ms), algorithmic FLOPs and launch count of the GEMM-class kernel since the
// last read; synchronises the device.  Any output may be NULL.
extern "C" int dm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return set_error((int)e, "dm_profile_read: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, fl = 0.0;
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) ms += t;
    fl += r.flops;
    g_prof_pool.push_back(r);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = static_cast<long long>(g_prof.size());
  g_prof.clear();
  return 0;
}

// Write one CSV row per profiled GEMM-class launch (logical tile grid, tile shape, duration, algorithmic FLOPs)
// and clear the record list; synchronises the device.
extern "C" int dm_profile_dump(const char* path) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return set_error((int)e, "dm_profile_dump: %s", cudaGetErrorString(e));
  FILE* f = fopen(path, "w");
  if (!f) return set_error(-1, "dm_profile_dump: cannot open %s", path);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  fprintf(f, "idx,kind,mode,a_mn,b_mn,tiles_m,tiles_n,tiles_z,bn,kc,stages,ctas,us,gflop\n");
  int i = 0;
  for (auto& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%.3f,%.4f\n", i++, r.kind, r.mode, r.a_mn, r.b_mn, r.gx, r.gy, r.gz, r.bn,
            r.kc, r.stages, r.ctas, t * 1e3, r.flops * 1e-9);
    g_prof_pool.push_back(r);
  }
  g_prof.clear();
  fclose(f);
  return 0;
}

extern "C" int dm_debug_last_plan(int* grid_xyz, int* smem_bytes, int* stages) {
  if (grid_xyz) { grid_xyz[0] = g_last_grid[0]; grid_xyz[1] = g_last_grid[1]; grid_xyz[2] = g_last_grid[2]; }
  if (smem_bytes) *smem_bytes = g_last_smem;
  if (stages) *stages = g_last_stages;
  return 0;
}

// ------------------------------------------------------------------------------------------ dense GEMM
static int gemm_impl(const dm_gemm_desc* g, bool tf32, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DM_REQUIRE(g != nullptr, "dm_gemm_bf16: null descriptor");
  t_kind = 0;
  DM_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0, "dm_gemm_bf16: bad shape %d %d %d", g->m, g->n, g->k);
  const int ld_align = tf32 ? 4 : 8;
  DM_REQUIRE(g->lda % ld_align == 0 && g->ldb % ld_align == 0, "dm_gemm: lda/ldb must be multiples of 16 bytes");
  const int kc_mn = tf32 ? 32 : 64;   // k-rows of an MN-major k-block = elements of a 128-byte K-major row
  const int aw = tf32 ? 32 : 64;      // MN elements of one MN-major atom
  DM_REQUIRE((reinterpret_cast<uintptr_t>(g->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->b) & 15) == 0,
             "dm_gemm_bf16: operands must be 16-byte aligned");
  const int splits = std::max(1, g->splits);
  DM_REQUIRE(splits == 1 || (g->accumulate && g->d_f32), "dm_gemm_bf16: split-K needs fp32 accumulate output");
  DM_REQUIRE(!g->accumulate || g->d_f32, "dm_gemm_bf16: accumulate needs fp32 output");
  GemmParams p;
  init_params(p);
  p.tf32 = tf32 ? 1 : 0;
  p.out = g->d;
  p.bias = g->bias;
  p.out_f32 = g->d_f32;
  p.out_atomic = g->accumulate;
  p.num_splits = splits;
  const int m_store = g->m_store > 0 ? g->m_store : g->m;
  const int n_store = g->n_store > 0 ? g->n_store : g->n;
  dim3 grid;
  int rc;
  if (g->layout == DM_GEMM_NT || g->layout == DM_GEMM_NN) {
    p.mode = MODE_FWD;
    p.a_mn = 0;
    p.b_mn = (g->layout == DM_GEMM_NN);
    p.kc = (g->k % 64 == 0 || p.b_mn) ? 64 : 32;
    if (!p.b_mn && g->k < 64 && g->k % 32 == 0) p.kc = 32;
    if (tf32) p.kc = 32;
    DM_REQUIRE(g->k % p.kc == 0 || g->k > p.kc, "dm_gemm_bf16: unsupported k %d", g->k);
    p.bn = p.b_mn ? (g->n >= 128 ? 128 : 64) : pick_bn(g->n, env_int("DM_BN_CAP", 128));
    p.cpt = (g->k + p.kc - 1) / p.kc;
    p.phase_tap_start[0] = 0;
    p.phase_tap_start[1] = 1;
    p.tw_step = 128; p.tpi = 1 << 30; p.th_step = 0; p.tn_step = 0;
    p.bw = 128; p.bh = 1;
    p.os_w = g->ldd_m; p.os_col = g->ldd_n;
    p.w_lim = m_store; p.n_lim = 1; p.n_valid = n_store;
    const int esz = tf32 ? 4 : 2;
    rc = encode_mat_map5(&p.map_a, g->a, g->m, g->k, g->lda, p.kc, 128, p.kc * esz, tf32);
    if (rc) return rc;
    if (!p.b_mn)
      rc = encode_w_map3(&p.map_b, g->b, 1, g->n, g->k, g->ldb, p.kc, p.bn, p.kc * esz, tf32);
    else
      rc = encode_w_map3(&p.map_b, g->b, 1, g->k, g->n, g->ldb, aw, kc_mn, tf32 ? kSwz128Atom32 : 128, tf32);
    if (rc) return rc;
    p.num_n_tiles = (g->n + p.bn - 1) / p.bn;
    if (g->ldd_n == 1) epi_rows_matrix(p, g->d, g->d_f32 != 0, m_store, n_store, g->ldd_m, g->accumulate != 0, 0);
    DM_REQUIRE(splits <= p.cpt, "dm_gemm_bf16: splits %d > k-blocks %d", splits, p.cpt);
    grid = dim3((g->m + 127) / 128, p.num_n_tiles, splits);
    if (g->bn.scratch) {
      DM_REQUIRE(g->m % 128 == 0 && (g->m / 128) % std::max(1, g->bn.groups) == 0 && ((g->m / std::max(1, g->bn.groups)) % 128) == 0,
                 "dm_gemm_bf16: fused statistics need whole 128-row tiles per group (m %d, groups %d)", g->m, g->bn.groups);
      if (int rc2 = attach_stats(p, &g->bn, n_store, g->m / 128, "dm_gemm_bf16")) return rc2;
    }
  } else if (g->layout == DM_GEMM_TN) {
    DM_REQUIRE(g->d_f32 || (g->ldd_m == 1 && !g->accumulate),
               "dm_gemm_bf16: TN (weight-gradient) output is fp32, or bf16 with unit row stride and no accumulation");
    DM_REQUIRE(g->bias == nullptr, "dm_gemm_bf16: TN has no bias");
    p.mode = MODE_WGRAD;
    p.a_mn = 1; p.b_mn = 1; p.kc = kc_mn;
    p.bn = g->n >= 256 ? env_int("DM_BN_WGRAD", 128) : (g->n >= 128 ? 128 : 64);
    p.num_kb = (g->k + kc_mn - 1) / kc_mn;
    p.tw_step = kc_mn; p.tpi = 1 << 30; p.th_step = 0; p.tn_step = 0;
    p.taps[0].nvalid = static_cast<int16_t>(std::min(n_store, 32767));
    p.taps[0].mvalid = 32767;
    p.os_m = g->ldd_m; p.os_n1 = g->ldd_n; p.os_n2 = 0; p.nmod = 1 << 30;
    p.m_valid = m_store;
    // A stored [k][m]: (c = m, w = k rows); B stored [k][n]
    rc = encode_mat_map5(&p.map_a, g->a, g->k, g->m, g->lda, aw, kc_mn, tf32 ? kSwz128Atom32 : 128, tf32);
    if (rc) return rc;
    rc = encode_mat_map5(&p.map_b, g->b, g->k, g->n, g->ldb, aw, kc_mn, tf32 ? kSwz128Atom32 : 128, tf32);
    if (rc) return rc;
    p.num_n_tiles = (g->n + p.bn - 1) / p.bn;
    if (g->ldd_n == 1)
      epi_rows_matrix(p, g->d, true, m_store, n_store, g->ldd_m, g->accumulate != 0, 128);
    else if (g->ldd_m == 1)
      epi_transposed_matrix(p, g->d, g->d_f32 != 0, m_store, n_store, g->ldd_n, g->accumulate != 0);
    DM_REQUIRE(g->d_f32 || p.epi_tma == 2, "dm_gemm_bf16: bf16 TN output needs the bulk-store epilogue (alignment)");
    if (n_store > 32767) p.taps[0].nvalid = 32767;  // columns are bounded by n_tiles*bn anyway
    DM_REQUIRE(g->n <= 32767 || g->n % p.bn == 0, "dm_gemm_bf16: TN n too large for masked store");
    DM_REQUIRE(splits <= p.num_kb, "dm_gemm_bf16: splits %d > k-blocks %d", splits, p.num_kb);
    grid = dim3((g->m + 127) / 128, p.num_n_tiles, splits);
  } else {
    return set_error(-1, "dm_gemm_bf16: unknown layout %d", g->layout);
  }
  const double k_alg = g->k_alg > 0 ? g->k_alg : g->k;
  const int kb_total = (p.mode == MODE_FWD) ? p.cpt : p.num_kb;
  return launch(p, grid, stream, 2.0 * m_store * n_store * k_alg, (kb_total + splits - 1) / splits);
}

extern "C" int dm_gemm_bf16(const dm_gemm_desc* g, void* stream_) { return gemm_impl(g, false, stream_); }
// fp32 operands and (usually) fp32 D, TF32 tensor-core arithmetic with fp32 accumulation: the TF32 precision mode
extern "C" int dm_gemm_tf32(const dm_gemm_desc* g, void* stream_) {
  DM_REQUIRE(g == nullptr || g->bn.scratch == nullptr, "dm_gemm_tf32: fused BatchNorm statistics are bf16-path only");
  return gemm_impl(g, true, stream_);
}

// ------------------------------------------------------------------------------------------ convolutions
static int check_geom(const dm_conv_geom* g, const char* who) {
  DM_REQUIRE(g != nullptr, "%s: null geometry", who);
  DM_REQUIRE(g->stride == 1 || g->stride == 2, "%s: stride must be 1 or 2", who);
  DM_REQUIRE(g->hb == g->hs * g->stride && g->wb == g->ws * g->stride, "%s: big side must be stride x small side", who);
  DM_REQUIRE(g->batch > 0 && g->cs > 0 && g->cb > 0, "%s: bad sizes", who);
  return 0;
}

// taps of `small = conv(big)` in the (c,w,p,h,n) view of big
static void down_taps(const dm_conv_geom* g, Tap* taps) {
  for (int kh = 0; kh < 5; ++kh)
    for (int kw = 0; kw < 5; ++kw) {
      Tap& t = taps[kh * 5 + kw];
      memset(&t, 0, sizeof(t));
      t.wt = static_cast<uint8_t>(kh * 5 + kw);
      if (g->stride == 1) {
        t.dh = static_cast<int8_t>(kh - 2);
        t.dw = static_cast<int8_t>(kw - 2);
      } else {
        const int eh = kh - 2, ew = kw - 2;
        const int ah = (eh >= 0) ? eh / 2 : -((-eh + 1) / 2);
        const int aw = (ew >= 0) ? ew / 2 : -((-ew + 1) / 2);
        t.dh = static_cast<int8_t>(ah);
        t.dp = static_cast<int8_t>(eh - 2 * ah);
        t.dw = static_cast<int8_t>(aw);
        t.dc = static_cast<int16_t>((ew - 2 * aw) * g->cb);
      }
    }
}

// pixel tiles of a stacked batch split into `groups` passes of whole tiles?
static bool groups_tile_aligned(const PixTile& pt, int batch, int groups) {
  if (groups <= 1) return true;
  if (batch % groups != 0 || pt.tiles % groups != 0) return false;
  return pt.bimg <= 1 || (batch / groups) % pt.bimg == 0;
}

// paired = true (stride 2, cb = 32, bf16): w_down is the PAIRED pack [15][cs][64] of dm_pack_down_pairs -- a k-block is
// one 128-byte row of the parity view (both w-parities x 32 channels) = the two filter columns (2j, 2j+1) of one
// filter row; 15 k-blocks of K = 64 instead of 25 of K = 32.  The TMA unit delivers ~0.7 box rows per clock per SM
// whatever the row length, so 64-byte rows feed the tensor cores at half the rate of 128-byte rows.
static int conv_down_impl(const dm_conv_geom* g, const void* big, const void* w_down, const float* bias,
                          void* out_small, const dm_bn_fuse* bn, bool tf32, void* stream_, bool paired = false) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv_down")) return rc;
  t_kind = 1;
  DM_REQUIRE(g->cb % 32 == 0, "dm_conv_down: cb %d must be a multiple of 32", g->cb);
  DM_REQUIRE(g->cs % 16 == 0, "dm_conv_down: cs %d must be a multiple of 16", g->cs);
  PixTile pt;
  DM_REQUIRE(make_pix_tile(128, g->batch, g->hs, g->ws, &pt), "dm_conv_down: unsupported grid %dx%d", g->hs, g->ws);
  GemmParams p;
  init_params(p);
  p.mode = MODE_FWD;
  p.tf32 = tf32 ? 1 : 0;
  const int esz = tf32 ? 4 : 2;
  if (paired) DM_REQUIRE(g->stride == 2 && g->cb == 32 && !tf32, "dm_conv_down_paired: stride 2, cb = 32, bf16 only");
  p.kc = ((g->cb % 64 == 0 || paired) && !tf32) ? 64 : 32;
  p.bn = pick_bn_fill(g->cs, env_int("DM_BN_CAP", 128), pt.tiles);
  p.cpt = paired ? 1 : g->cb / p.kc;
  p.phase_tap_start[0] = 0;
  p.phase_tap_start[1] = paired ? 15 : 25;
  if (paired) {
    for (int kh = 0; kh < 5; ++kh)
      for (int j = 0; j < 3; ++j) {  // filter columns (2j, 2j+1): input w = 2*ow + 2j - 2 (+1) -> parity 0 (1), w/2 = ow + j - 1
        Tap& t = p.taps[kh * 3 + j];
        memset(&t, 0, sizeof(t));
        t.wt = static_cast<uint8_t>(kh * 3 + j);
        const int eh = kh - 2;
        const int ah = (eh >= 0) ? eh / 2 : -((-eh + 1) / 2);
        t.dh = static_cast<int8_t>(ah);
        t.dp = static_cast<int8_t>(eh - 2 * ah);
        t.dw = static_cast<int8_t>(j - 1);
        t.dc = 0;
      }
  } else {
    down_taps(g, p.taps);
  }
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.bw = pt.bw; p.bh = pt.bh;
  p.out = out_small; p.bias = bias; p.out_f32 = tf32 ? 1 : 0; p.out_atomic = 0;
  p.os_w = g->cs; p.os_h = (long long)g->ws * g->cs; p.os_n = (long long)g->hs * g->ws * g->cs; p.os_col = 1;
  p.w_lim = g->ws; p.n_lim = g->batch; p.n_valid = g->cs;
  uint32_t box[5] = {(uint32_t)p.kc, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, (uint32_t)pt.bimg};
  if (int rc = encode_act_map(&p.map_a, big, g->batch, g->hb, g->wb, g->cb, g->stride, box, p.kc * esz, tf32)) return rc;
  p.cg2 = (!tf32 && env_int("DM_CG2", 1) != 0 && pt.tiles >= 2 && p.bn >= 32) ? 1 : 0;
  const int cluster = p.cg2 ? 2 : 1;
  if (paired) {
    if (int rc = encode_w_map3(&p.map_b, w_down, 15, g->cs, 64, 64, 64, p.bn / cluster, 128, false)) return rc;
  } else {
    if (int rc = encode_w_map3(&p.map_b, w_down, 25, g->cs, g->cb, g->cb, p.kc, p.bn / cluster, p.kc * esz, tf32)) return rc;
  }
  p.num_n_tiles = (g->cs + p.bn - 1) / p.bn;
  epi_rows_act(p, out_small, tf32, g->batch, g->hs, g->ws, g->cs, 1, pt);
  if (bn && bn->scratch) {
    DM_REQUIRE(groups_tile_aligned(pt, g->batch, bn->groups), "dm_conv_down: groups do not fall on tile boundaries");
    if (int rc = attach_stats(p, bn, g->cs, pt.tiles, "dm_conv_down")) return rc;
  }
  return launch(p, dim3(pt.tiles, p.num_n_tiles, 1), stream, 50.0 * g->batch * g->hs * g->ws * g->cs * g->cb, 1 << 30,
                cluster);
}

extern "C" int dm_conv_down(const dm_conv_geom* g, const void* big, const void* w_down, const float* bias,
                            void* out_small, const dm_bn_fuse* bn, void* stream_) {
  return conv_down_impl(g, big, w_down, bias, out_small, bn, false, stream_);
}
// TF32 precision mode: big / w_down / out_small are fp32 (same layouts), tensor-core arithmetic in TF32, fp32 accumulate
extern "C" int dm_conv_down_paired(const dm_conv_geom* g, const void* big, const void* w_pair, const float* bias,
                                  void* out_small, const dm_bn_fuse* bn, void* stream_) {
  return conv_down_impl(g, big, w_pair, bias, out_small, bn, false, stream_, true);
}

extern "C" int dm_conv_down_tf32(const dm_conv_geom* g, const void* big, const void* w_down, const float* bias,
                                 void* out_small, void* stream_) {
  return conv_down_impl(g, big, w_down, bias, out_small, nullptr, true, stream_);
}

static int conv_up_impl(const dm_conv_geom* g, const void* small, const void* w_up, const float* bias, void* out_big,
                        int out_f32, const dm_bn_fuse* bn, bool tf32, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv_up")) return rc;
  t_kind = 2;
  DM_REQUIRE(g->cs % 32 == 0, "dm_conv_up: cs %d must be a multiple of 32", g->cs);
  const int cb_pad = std::max(16, (g->cb + 15) / 16 * 16);
  PixTile pt;
  DM_REQUIRE(make_pix_tile(128, g->batch, g->hs, g->ws, &pt), "dm_conv_up: unsupported grid %dx%d", g->hs, g->ws);
  GemmParams p;
  init_params(p);
  p.mode = MODE_FWD;
  p.tf32 = tf32 ? 1 : 0;
  const int esz = tf32 ? 4 : 2;
  p.kc = (g->cs % 64 == 0 && !tf32) ? 64 : 32;
  p.bn = pick_bn_fill(cb_pad, env_int("DM_BN_CAP", 128), (g->stride == 2 ? 4ll : 1ll) * pt.tiles);
  p.cpt = g->cs / p.kc;
  int nphase = 0, nt = 0;
  const bool fold = (g->stride == 1 && g->cb == 3);
  DM_REQUIRE(!(tf32 && fold), "dm_conv_up_tf32: the 3-channel kw-folded layer is bf16-path only");
  if (tf32) out_f32 = 1;
  DM_REQUIRE(!fold || out_f32, "dm_conv_up: the 3-channel image side is written as fp32");
  if (fold) {
    // 5 k-blocks (one per kh) instead of 25: the A box is shifted vertically only, the 5 horizontal taps live in
    // the N dimension (w_up for this layer is packed as [5][16 = kw*3+cb][32], see dm_pack_conv_weights)
    p.fold_kw = 1;
    p.bn = 16;
    p.phase_tap_start[0] = 0;
    for (int kh = 0; kh < 5; ++kh) {
      Tap& t = p.taps[nt++];
      t.dh = static_cast<int8_t>(2 - kh);
      t.dw = 0;
      t.wt = static_cast<uint8_t>(kh);
    }
    p.phase_tap_start[1] = nt;
    nphase = 1;
    p.os_w = g->cb; p.os_h = (long long)g->wb * g->cb;
  } else if (g->stride == 1) {
    p.phase_tap_start[0] = 0;
    for (int kh = 0; kh < 5; ++kh)
      for (int kw = 0; kw < 5; ++kw) {
        Tap& t = p.taps[nt++];
        t.dh = static_cast<int8_t>(2 - kh);
        t.dw = static_cast<int8_t>(2 - kw);
        t.wt = static_cast<uint8_t>(kh * 5 + kw);
      }
    p.phase_tap_start[1] = nt;
    p.phase_out_off[0] = 0;
    nphase = 1;
    p.os_w = g->cb; p.os_h = (long long)g->wb * g->cb;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        p.phase_tap_start[nphase] = nt;
        for (int kh = ph; kh < 5; kh += 2)
          for (int kw = pw; kw < 5; kw += 2) {
            Tap& t = p.taps[nt++];
            t.dh = static_cast<int8_t>((ph + 2 - kh) / 2);
            t.dw = static_cast<int8_t>((pw + 2 - kw) / 2);
            t.wt = static_cast<uint8_t>(kh * 5 + kw);
          }
        p.phase_out_off[nphase] = ((long long)ph * g->wb + pw) * g->cb;
        ++nphase;
      }
    p.phase_tap_start[nphase] = nt;
    p.os_w = 2ll * g->cb; p.os_h = 2ll * g->wb * g->cb;
  }
  p.os_n = (long long)g->hb * g->wb * g->cb; p.os_col = 1;
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.bw = pt.bw; p.bh = pt.bh;
  p.out = out_big; p.bias = bias; p.out_f32 = out_f32; p.out_atomic = 0;
  p.w_lim = g->ws; p.n_lim = g->batch; p.n_valid = g->cb;
  uint32_t box[5] = {(uint32_t)p.kc, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, (uint32_t)pt.bimg};
  if (int rc = encode_act_map(&p.map_a, small, g->batch, g->hs, g->ws, g->cs, 1, box, p.kc * esz, tf32)) return rc;
  p.cg2 = (!tf32 && !fold && env_int("DM_CG2", 1) != 0 && pt.tiles >= 2 && p.bn >= 32) ? 1 : 0;
  const int cluster = p.cg2 ? 2 : 1;
  if (int rc = encode_w_map3(&p.map_b, w_up, fold ? 5 : 25, cb_pad, g->cs, g->cs, p.kc, p.bn / cluster, p.kc * esz, tf32)) return rc;
  p.num_n_tiles = (cb_pad + p.bn - 1) / p.bn;
  if (!fold) epi_rows_act(p, out_big, out_f32 != 0, g->batch, g->hb, g->wb, g->cb, g->stride, pt);
  if (bn && bn->scratch) {
    DM_REQUIRE(groups_tile_aligned(pt, g->batch, bn->groups), "dm_conv_up: groups do not fall on tile boundaries");
    DM_REQUIRE(cb_pad == g->cb, "dm_conv_up: fused statistics need cb %% 16 == 0");
    if (int rc = attach_stats(p, bn, g->cb, pt.tiles, "dm_conv_up")) return rc;
  }
  return launch(p, dim3(pt.tiles, p.num_n_tiles, nphase), stream, 50.0 * g->batch * g->hs * g->ws * g->cs * g->cb,
                1 << 30, cluster);
}

extern "C" int dm_conv_up(const dm_conv_geom* g, const void* small, const void* w_up, const float* bias, void* out_big,
                          int out_f32, const dm_bn_fuse* bn, void* stream_) {
  return conv_up_impl(g, small, w_up, bias, out_big, out_f32, bn, false, stream_);
}
extern "C" int dm_conv_up_tf32(const dm_conv_geom* g, const void* small, const void* w_up, const float* bias,
                               void* out_big, void* stream_) {
  return conv_up_impl(g, small, w_up, bias, out_big, 1, nullptr, true, stream_);
}

// Phase-merged stride-2 transposed convolution for cb == 32 (see dm_pack_up_merged): ONE GEMM with 9 input taps and
// N = 4 * cb = 128 columns (column = (ph*2 + pw)*cb + c) instead of four phase GEMMs with N = 32.  An A tile is then
// loaded 9 times instead of 25 and feeds 128-wide MMAs: the N = 32 form is bound by L2 -> smem operand traffic
// (measured 15.6 TB/s, 21 % tensor-pipe), this one by the tensor pipe.  31 % of the MMA work multiplies zeros.
extern "C" int dm_conv_up_merged(const dm_conv_geom* g, const void* small, const void* w_upm, const float* bias,
                                 void* out_big, const dm_bn_fuse* bn, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv_up_merged")) return rc;
  t_kind = 2;
  DM_REQUIRE(g->stride == 2 && g->cb == 32, "dm_conv_up_merged: needs stride 2 and cb == 32");
  DM_REQUIRE(g->cs % 64 == 0, "dm_conv_up_merged: cs %d must be a multiple of 64", g->cs);
  PixTile pt;
  DM_REQUIRE(make_pix_tile(128, g->batch, g->hs, g->ws, &pt), "dm_conv_up_merged: unsupported grid %dx%d", g->hs, g->ws);
  GemmParams p;
  init_params(p);
  p.mode = MODE_FWD;
  p.kc = 64;
  p.bn = 4 * g->cb;
  p.cpt = g->cs / p.kc;
  int nt = 0;
  p.phase_tap_start[0] = 0;
  for (int dh = 1; dh >= -1; --dh)
    for (int dw = 1; dw >= -1; --dw) {
      Tap& t = p.taps[nt];
      t.dh = static_cast<int8_t>(dh);
      t.dw = static_cast<int8_t>(dw);
      t.wt = static_cast<uint8_t>(nt);
      ++nt;
    }
  p.phase_tap_start[1] = nt;
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.bw = pt.bw; p.bh = pt.bh;
  p.out = out_big; p.bias = bias; p.out_f32 = 0; p.out_atomic = 0;
  p.bias_mod = g->cb;
  p.w_lim = g->ws; p.n_lim = g->batch; p.n_valid = p.bn;
  uint32_t box[5] = {(uint32_t)p.kc, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, (uint32_t)pt.bimg};
  if (int rc = encode_act_map(&p.map_a, small, g->batch, g->hs, g->ws, g->cs, 1, box, p.kc * 2)) return rc;
  p.cg2 = (env_int("DM_CG2", 1) != 0 && pt.tiles >= 2) ? 1 : 0;
  const int cluster = p.cg2 ? 2 : 1;
  if (int rc = encode_w_map3(&p.map_b, w_upm, 9, p.bn, g->cs, g->cs, p.kc, p.bn / cluster, p.kc * 2)) return rc;
  p.num_n_tiles = 1;
  // epilogue: 64-column box j = row parity ph = j of the parity-split output view (2cb, wb/2, 2, hb/2, b)
  uint32_t sub[3];
  DM_REQUIRE(epi_common(p, out_big, false) && epi_sub_box(pt.bw, pt.bh, pt.bimg, sub),
             "dm_conv_up_merged: output not expressible as box stores");
  {
    const uint64_t e = 2, c = g->cb, w = g->wb, h = g->hb;
    uint64_t dims[5] = {2 * c, w / 2, 2, h / 2, (uint64_t)g->batch};
    uint64_t str[4] = {2 * c * e, w * c * e, 2 * w * c * e, h * w * c * e};
    uint32_t obox[5] = {(uint32_t)p.epi_cols, sub[0], 1, sub[1], sub[2]};
    if (int rc = encode_map(&p.map_out, out_big, 5, dims, str, obox, p.epi_swz, false)) return rc;
  }
  p.epi_tma = 1; p.epi_reduce = 0; p.epi_merge = 1;
  p.epi_bw = pt.bw; p.epi_bh = pt.bh; p.epi_row_step = 0;
  if (bn && bn->scratch) {  // the 128 columns are 4 phases x 32 channels: column j is channel j & 31
    DM_REQUIRE(groups_tile_aligned(pt, g->batch, bn->groups), "dm_conv_up_merged: groups do not fall on tile boundaries");
    if (int rc = attach_stats(p, bn, g->cb, pt.tiles, "dm_conv_up_merged")) return rc;
  }
  return launch(p, dim3(pt.tiles, 1, 1), stream, 50.0 * g->batch * g->hs * g->ws * g->cs * g->cb, 1 << 30, cluster);
}

static int conv_wgrad_impl(const dm_conv_geom* g, const void* small, const void* big, float* dw_packed,
                           int direct_layout, bool tf32, void* stream_) {
  // D_t[m = cb][n = cs] = sum_pixels big_tap_t[pix][cb] * small[pix][cs], accumulated into the tap-major packed
  // gradient dw_packed[25][cs][cb]: a warp's 32 rows (cb) are contiguous floats -> coalesced reductions.
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv_wgrad")) return rc;
  t_kind = 3;
  DM_REQUIRE(g->cs % 64 == 0, "dm_conv_wgrad: cs %d must be a multiple of 64", g->cs);
  const bool pair = (!tf32 && g->cb == 32 && g->stride == 2);
  DM_REQUIRE(g->cb % 64 == 0 || pair || (tf32 && g->cb % 32 == 0),
             "dm_conv_wgrad: cb %d must be a multiple of 64 (or 32 with stride 2)", g->cb);
  const int kpix = tf32 ? 32 : 64;   // pixels (k-rows) per k-block = height of an MN-major atom
  const int aw = tf32 ? 32 : 64;     // channels of one atom
  PixTile pt;
  DM_REQUIRE(make_pix_tile(kpix, g->batch, g->hs, g->ws, &pt), "dm_conv_wgrad: unsupported grid %dx%d", g->hs, g->ws);
  GemmParams p;
  init_params(p);
  p.tf32 = tf32 ? 1 : 0;
  p.wgrad_direct = direct_layout;
  p.mode = MODE_WGRAD;
  p.a_mn = 1; p.b_mn = 1; p.kc = kpix;
  p.num_kb = pt.tiles;
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.out = dw_packed; p.out_f32 = 1; p.out_atomic = 1;
  p.bn = g->cs >= 256 ? env_int("DM_BN_WGRAD", 128) : (g->cs >= 128 ? 128 : 64);
  p.num_n_tiles = g->cs / p.bn;
  // direct = 1: accumulate straight into the parameter layout dw[cs][cb][25] (scattered 4-byte reductions, no
  // unpack pass); direct = 0: tap-major packed layout [25][cs][cb] (coalesced reductions + dm_unpack_conv_grad)
  const bool direct = p.wgrad_direct != 0;
  p.os_n1 = direct ? (long long)g->cb * 25 : g->cb; p.os_n2 = 0; p.nmod = 1 << 30;
  const long long tap_stride = direct ? 1 : (long long)g->cs * g->cb;
  const long long m_stride = direct ? 25 : 1;
  int units = 0, m_tiles = 1;
  if (!pair) {
    m_tiles = (g->cb + 127) / 128;
    down_taps(g, p.taps);
    for (int t = 0; t < 25; ++t) {
      p.taps[t].nvalid = static_cast<int16_t>(g->cs);
      p.taps[t].mvalid = static_cast<int16_t>(g->cb);
      p.taps[t].out_off = static_cast<int32_t>(t * tap_stride);
      p.taps[t].out_tap = static_cast<int16_t>(t);
    }
    p.os_m = m_stride; p.os_m2 = 0; p.mmod = 1 << 30;
    p.m_valid = g->cb;
    units = 25;
  } else {
    // cb == 32, stride 2: one 64-wide box covers both w-parities = filter columns (2aw+2, 2aw+3):
    // row m -> (pw = m / 32, cb = m % 32), tap = kh*5 + 2aw+2 + pw
    for (int kh = 0; kh < 5; ++kh)
      for (int aw = -1; aw <= 1; ++aw) {
        Tap& t = p.taps[units++];
        const int eh = kh - 2;
        const int ah = (eh >= 0) ? eh / 2 : -((-eh + 1) / 2);
        t.dc = 0;
        t.dw = static_cast<int8_t>(aw);
        t.dh = static_cast<int8_t>(ah);
        t.dp = static_cast<int8_t>(eh - 2 * ah);
        t.nvalid = static_cast<int16_t>(g->cs);
        t.mvalid = static_cast<int16_t>(aw == 1 ? 32 : 64);
        t.out_off = static_cast<int32_t>((kh * 5 + 2 * aw + 2) * tap_stride);
        t.out_tap = static_cast<int16_t>(kh * 5 + 2 * aw + 2);
      }
    p.os_m = m_stride; p.mmod = 32; p.os_m2 = tap_stride;
    p.m_valid = 64;
  }
  uint32_t boxa[5] = {(uint32_t)aw, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, (uint32_t)pt.bimg};
  // operand roles: A = big (tap-shifted, M = cb), B = small (N = cs)
  const int swz = tf32 ? kSwz128Atom32 : 128;
  if (int rc = encode_act_map(&p.map_a, big, g->batch, g->hb, g->wb, g->cb, g->stride, boxa, swz, tf32)) return rc;
  if (int rc = encode_act_map(&p.map_b, small, g->batch, g->hs, g->ws, g->cs, 1, boxa, swz, tf32)) return rc;
  p.wgrad_tap_on_a = 1;
  if (!direct) epi_wgrad_packed(p, dw_packed, g->cs, g->cb);
  const int base_ctas = m_tiles * p.num_n_tiles * units;
  int splits = env_int("DM_WGRAD_SPLITS", 0);
  if (splits <= 0) {
    // split-K so that the persistent grid runs in WHOLE waves: cost(s) = waves(s) x (k-blocks per item + a fixed
    // per-item epilogue/reduction cost, in k-block units); 300 items on 296 slots would take two rounds
    const int acc_cols = (p.bn + 31) / 32 * 32;
    const int stage_b = 2 * kAtomBytes + (p.bn >> 6) * kAtomBytes;
    const bool two = pow2_cols(2 * acc_cols) <= 256 && stage_b <= 32768 && env_int("DM_ONE_CTA", 0) == 0;
    const int slots = num_sms() * (two ? 2 : 1);
    const int fixed = env_int("DM_WGRAD_FIXED_KB", 12);
    long long best = -1;
    for (int sp = 1; sp <= std::min(p.num_kb, 64); ++sp) {
      const long long items = static_cast<long long>(base_ctas) * sp;
      const long long waves = (items + slots - 1) / slots;
      const long long cost = waves * ((p.num_kb + sp - 1) / sp + fixed);
      if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
  }
  splits = std::min(splits, p.num_kb);
  p.num_splits = splits;
  const int cluster = 1;
  return launch(p, dim3(m_tiles, p.num_n_tiles * units, splits), stream,
                50.0 * g->batch * g->hs * g->ws * g->cs * g->cb, (p.num_kb + splits - 1) / splits, cluster);
}

extern "C" int dm_conv_wgrad(const dm_conv_geom* g, const void* small, const void* big, float* dw_packed,
                             int direct_layout, void* stream_) {
  return conv_wgrad_impl(g, small, big, dw_packed, direct_layout, false, stream_);
}
// TF32 precision mode: small / big are fp32 NHWC; dw_packed[25][cs][cb] fp32 (tap-major), accumulated
extern "C" int dm_conv_wgrad_tf32(const dm_conv_geom* g, const void* small, const void* big, float* dw_packed,
                                  void* stream_) {
  return conv_wgrad_impl(g, small, big, dw_packed, 0, true, stream_);
}

// ------------------------------------------------------------------------------------------ 3-channel image side
// The three layers that touch the 3-channel image (Discriminator convs.0 / Encoder features.0 forward and weight
// gradient, decoder deconv4 input- and weight-gradient) as implicit GEMMs over the PADDED IMAGE (dm_pad_image3):
// bf16 [b][68][72][4], pixel (h, w) at [h+2][w+2], zero border / 4th channel.  A filter row kh of an output pixel is
// the contiguous run of 5 pixels x 4 channels starting at padded pixel (s*oh + kh, s*ow).  TMA strides are multiples
// of 16 B = two pixels, so the A operand of k-block kh is a box of 32 elements (8 pixels) starting at an EVEN pixel:
//   stride 2: one window per output pixel (2*ow is even): M = pixels, N = cs, weights w_win[kh][cs][32];
//   stride 1: one window per output-pixel PAIR (2*w2, 2*w2+1), the pair's two pixels as 2*cs output COLUMNS
//             (pw*cs + n) with weights shifted by pw pixels: M = pairs, N = 2*cs -- and [pairs][2*cs] IS the NHWC tensor.
// The tensor map's position stride (16 B) is smaller than the box (64 B): overlapping windows.  K = 5 x 32 instead of
// 75, but no im2col matrix: 8 B/pixel of image instead of a 160 B/pixel matrix written and read.
constexpr int kPimH = 68, kPimW = 72;

// window view of the padded image: (e = `inner` window elements, position, row parity, row, n)
static int encode_pim_map(CUtensorMap* m, const void* pim, int batch, int stride, const uint32_t* box, int inner) {
  const uint64_t pos = 16, row = (uint64_t)kPimW * 8, img = (uint64_t)kPimH * row;
  uint64_t dims[5], str[4];
  dims[0] = inner; dims[1] = 32; dims[4] = batch;
  str[0] = pos; str[3] = img;
  if (stride == 1) {
    dims[2] = 1; dims[3] = kPimH;
    str[1] = row; str[2] = row;
  } else {
    dims[2] = 2; dims[3] = kPimH / 2;
    str[1] = row; str[2] = 2 * row;
  }
  return encode_map(m, pim, 5, dims, str, box, inner * 2);  // 64 elements: SWIZZLE_128B, 32: SWIZZLE_64B
}

// taps of filter row kh in the window view: padded row = s*oh + kh
static void win_tap(Tap& t, int kh, int stride) {
  memset(&t, 0, sizeof(t));
  t.dh = static_cast<int8_t>(stride == 1 ? kh : kh / 2);
  t.dp = static_cast<int8_t>(stride == 1 ? 0 : kh % 2);
  t.wt = static_cast<uint8_t>(kh);
}

/* out[b,hs,ws,cs] (bf16 NHWC) = conv5x5(image, W) + bias, stride 1 or 2, from the padded image; w_win from
 * dm_pack_conv3_weights(stride).  g: cb == 3, hb == wb == 64. */
extern "C" int dm_conv3_fwd(const dm_conv_geom* g, const void* pim, const void* w_win, const float* bias, void* out_small,
                            const dm_bn_fuse* bn, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv3_fwd")) return rc;
  t_kind = 4;
  DM_REQUIRE(g->cb == 3 && g->hb == 64 && g->wb == 64, "dm_conv3_fwd: needs a 3-channel 64x64 image side");
  DM_REQUIRE(g->cs == 32 || g->cs == 64, "dm_conv3_fwd: cs %d must be 32 or 64", g->cs);
  const int nn = g->stride == 1 ? 2 * g->cs : g->cs;  // GEMM columns: (pixel of the pair, channel) or channel
  PixTile pt;  // tiles of 128 window positions on the (hs rows) x (32 positions) grid
  DM_REQUIRE(make_pix_tile(128, g->batch, g->hs, 32, &pt) && pt.bimg == 1, "dm_conv3_fwd: unsupported grid");
  GemmParams p;
  init_params(p);
  p.mode = MODE_FWD;
  p.kc = 32;
  p.bn = nn;
  p.cpt = 1;
  p.phase_tap_start[0] = 0;
  p.phase_tap_start[1] = 5;
  for (int kh = 0; kh < 5; ++kh) win_tap(p.taps[kh], kh, g->stride);
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.bw = pt.bw; p.bh = pt.bh;
  p.out = out_small; p.bias = bias; p.out_f32 = 0; p.out_atomic = 0;
  p.bias_mod = g->cs;  // (stride 1: the 2*cs columns repeat the channels)
  p.os_w = nn; p.os_h = 32ll * nn; p.os_n = (long long)g->hs * 32 * nn; p.os_col = 1;
  p.w_lim = 32; p.n_lim = g->batch; p.n_valid = nn;
  uint32_t box[5] = {32, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, 1};
  if (int rc = encode_pim_map(&p.map_a, pim, g->batch, g->stride, box, 32)) return rc;
  p.cg2 = (env_int("DM_CG2", 1) != 0 && pt.tiles >= 2) ? 1 : 0;
  const int cluster = p.cg2 ? 2 : 1;
  if (int rc = encode_w_map3(&p.map_b, w_win, 5, nn, 32, 32, 32, p.bn / cluster, 64)) return rc;
  p.num_n_tiles = 1;
  // the output as the NHWC tensor [b][hs][32 positions][nn]
  DM_REQUIRE(epi_rows_act(p, out_small, false, g->batch, g->hs, 32, nn, 1, pt), "dm_conv3_fwd: output not expressible as box stores");
  if (bn && bn->scratch) {
    DM_REQUIRE(groups_tile_aligned(pt, g->batch, bn->groups), "dm_conv3_fwd: groups do not fall on tile boundaries");
    if (int rc = attach_stats(p, bn, g->cs, pt.tiles, "dm_conv3_fwd")) return rc;
  }
  return launch(p, dim3(pt.tiles, 1, 1), stream, 50.0 * g->batch * g->hs * g->ws * g->cs * g->cb, 1 << 30, cluster);
}

/* dw_win[5][nn][64] (fp32 window layout, nn = cs or 2*cs; dm_unpack_conv3_grad moves it to dw[cs][3][5][5]) +=
 *   sum over window positions of small (bf16 NHWC, viewed [positions][nn]: the output gradient of a Conv2d / the input
 *   of deconv4) x 16-pixel window.  D[m = (filter row pair, window element)][n], K = positions. */
extern "C" int dm_conv3_wgrad(const dm_conv_geom* g, const void* pim, const void* small, float* dw_win, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_geom(g, "dm_conv3_wgrad")) return rc;
  t_kind = 4;
  DM_REQUIRE(g->cb == 3 && g->hb == 64 && g->wb == 64, "dm_conv3_wgrad: needs a 3-channel 64x64 image side");
  DM_REQUIRE((g->cs == 32 && g->stride == 1) || (g->cs == 64 && g->stride == 2),
             "dm_conv3_wgrad: supports cs 32 / stride 1 and cs 64 / stride 2 (64 GEMM columns)");
  const int nn = g->stride == 1 ? 2 * g->cs : g->cs;
  PixTile pt;  // k-blocks of 64 window positions
  DM_REQUIRE(make_pix_tile(64, g->batch, g->hs, 32, &pt) && pt.bimg == 1, "dm_conv3_wgrad: unsupported grid");
  GemmParams p;
  init_params(p);
  p.mode = MODE_WGRAD;
  p.a_mn = 1; p.b_mn = 1; p.kc = 64;
  p.num_kb = pt.tiles;
  p.tw_step = 0; p.tpi = pt.tpi; p.th_step = pt.th_step; p.tn_step = pt.tn_step;
  p.out = dw_win; p.out_f32 = 1; p.out_atomic = 1;
  p.bn = nn;
  p.num_n_tiles = 1;
  p.wgrad_win = 1;
  p.wgrad_tap_on_a = 1;
  for (int u = 0; u < 3; ++u) {
    Tap& t = p.taps[2 * u];
    win_tap(t, 2 * u, g->stride);
    t.nvalid = static_cast<int16_t>(nn);
    t.mvalid = static_cast<int16_t>(u == 2 ? 64 : 128);
    t.out_tap = static_cast<int16_t>(2 * u);
    Tap& t2 = p.taps[2 * u + 1];
    win_tap(t2, 2 * u + 1, g->stride);
    if (u == 2) t2.dh = 100;  // no sixth filter row: the box lies outside the tensor -> zero fill
  }
  p.mmod = 64; p.m_valid = 128;
  uint32_t boxa[5] = {64, (uint32_t)pt.bw, 1, (uint32_t)pt.bh, 1};
  if (int rc = encode_pim_map(&p.map_a, pim, g->batch, g->stride, boxa, 64)) return rc;
  // B = small viewed as [b][hs][32 positions][nn = 64]
  if (int rc = encode_act_map(&p.map_b, small, g->batch, g->hs, 32, nn, 1, boxa, 128)) return rc;
  {  // transposed reduce-add into dw_win[5][nn][64]: box {32 m, 32 n, 1 filter row}
    DM_REQUIRE((reinterpret_cast<uintptr_t>(dw_win) & 15) == 0, "dm_conv3_wgrad: dw_win must be 16-byte aligned");
    uint64_t dims[3] = {64, (uint64_t)nn, 5};
    uint64_t str[2] = {64 * 4, (uint64_t)nn * 64 * 4};
    uint32_t box[3] = {32, 32, 1};
    if (int rc = encode_map(&p.map_out, dw_win, 3, dims, str, box, 0, true)) return rc;
    p.epi_tma = 2; p.epi_reduce = 1;
  }
  int splits = 1;
  {  // split-K over the position tiles: whole waves of the persistent grid (as dm_conv_wgrad)
    const int slots = num_sms() * 2;
    const int fixed = env_int("DM_WGRAD_FIXED_KB", 12);
    long long best = -1;
    for (int sp = 1; sp <= std::min(p.num_kb, 128); ++sp) {
      const long long items = 3ll * sp;
      const long long waves = (items + slots - 1) / slots;
      const long long cost = waves * ((p.num_kb + sp - 1) / sp + fixed);
      if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
  }
  p.num_splits = splits;
  return launch(p, dim3(1, 3, splits), stream, 50.0 * g->batch * g->hs * g->ws * g->cs * g->cb,
                (p.num_kb + splits - 1) / splits, 1);
}
