// FC (Dense) layer kernels:  D[m][n] = act( alpha * sum_k A[m][k] W[n][k] + bias[n] )
//
// Replaces the Keras Dense(+relu) / Dense(linear) layers of the reference's CSI predictor
// (massiveMIMO_CSI_prediction_DNN.py:211-227); BatchNormalization is folded into the next
// layer's W/bias on the host (engine.cu), Dropout is the identity at inference.
//
//   fc_tc_kernel   -- tcgen05 + TMEM + TMA, warp-specialised, persistent over output tiles.
//                     Split-precision operands (schemes.cuh): 3 UMMA passes per k-step
//                     accumulate hi.hi + hi.lo + lo.hi in FP32 in tensor memory.
//   fc_simt_kernel -- exact-FP32 CUDA-core GEMM: the on-device accuracy anchor
//                     (MAMIMO_PREC_FP32_SIMT), also usable when 1e-7-grade agreement is wanted.
//
// Both K-major: A = activations [rows][Kpad], W = weights [N][Kpad] (the transpose of the
// Keras kernel), so each output element is a dot product of two contiguous rows.
//
// Accuracy note (measured on B200, round 1): the tensor core adds every MMA's result into the
// TMEM accumulator with truncation (round toward zero).  A 1024-deep layer is 64..128 k-steps x 3
// passes = 192..384 truncations of a full-magnitude accumulator, a systematic shrink of ~2e-5 --
// outside the 1e-5 budget.  So the accumulation chain is cut: the MMA warp accumulates only
// `kb_per_chunk` k-blocks into one TMEM buffer (correction passes first, while the buffer is still
// small, the dominant hi.hi pass last), the epilogue warps drain that buffer into FP32 registers
// with round-to-nearest adds while the MMA warp fills the other buffer.
#pragma once
#include "ptx.cuh"
#include "schemes.cuh"

namespace mm {

// Role wait-cycle counters of the pair kernel are compiled in only with -DMAMIMO_FC_DEBUG_COUNTERS (they cost
// registers in the 40-register producer/MMA warps); see engine.cu MAMIMO_FC_DEBUG.
#ifdef MAMIMO_FC_DEBUG_COUNTERS
#define MM_DBG(a) ((a).dbg != nullptr)
#else
#define MM_DBG(a) false
#endif

struct FcArgs {
  int M;                 // valid rows
  int N;                 // valid output features
  int num_k_blocks;      // Kpad / kBlockK
  int kb_per_chunk;      // k-blocks accumulated inside the tensor core before a register drain
  int a_plane_rows;      // rows_alloc of the A operand (plane stride, rows)
  int b_plane_rows;      // Npad of the W operand
  const float* bias;     // [N], BN-folded
  float alpha;           // undoes the operand scales: 1 / (a_scale * w_scale)
  int relu;
  // hidden layer: next layer's A operand (split planes).  nullptr for the final layer
  void* out_planes;
  int out_kpad;
  int out_plane_rows;
  float out_scale;
  // final layer: float32 [M][out_ld]
  float* out_f32;
  int out_ld;
  uint32_t* flags;
  // SIMT kernel only (the tcgen05 kernel reads operands through its tensor maps)
  const float* A;
  const float* W;
  int kpad;
  int gather_world;      // fused all-gather (pair kernel, final layer): ranks to store to
  int l2_prefetch;       // pair kernel: prefetch the next tile's activation rows into L2
  unsigned long long* dbg;  // optional [8]: cycles the pair kernel's roles spent waiting (MAMIMO_FC_DEBUG=1)
  // FP16X3 range management (schemes.cuh).  dyn != nullptr: alpha = w_inv_scale / dyn->scale[net][level]; a hidden
  // layer picks its output scale from |out| <= rowsum * amax(level) + bmax (or keeps out_scale when fixed_scale),
  // publishes it as dyn->scale[net][level + 1] and the measured amax of its outputs as dyn->amax[net][level + 1].
  DynState* dyn;
  int net, level, fixed_scale;
  int row_off;           // first row of this call inside the operand buffers (sub-batches of a pipelined step keep
                         // their activations side by side in the same buffers); multiple of the 256-row pair tile
  float w_inv_scale;     // 1 / weight scale
  float rowsum;          // max_n sum_k |W[k][n]| (BN-folded), rounded up
  float bmax;            // max_n |bias[n]|
};

struct FcScales {
  float alpha, out_scale;
};

// every epilogue thread resolves the same two numbers (two broadcast loads from L2); `writer` is one thread of the
// grid, which publishes the output scale and runs the pinned-scale window check
template <int S>
__device__ __forceinline__ FcScales fc_resolve_scales(const FcArgs& a, bool writer) {
  FcScales r{a.alpha, a.out_scale};
  if constexpr (S == kFp16x3) {
    if (a.dyn) {
      const float s_in = a.dyn->scale[a.net][a.level];
      const float amax_in = __uint_as_float(a.dyn->amax[a.net][a.level]);
      r.alpha = a.w_inv_scale / s_in;                     // powers of two: exact
      if (a.out_planes) {
        if (!a.fixed_scale) r.out_scale = pow2_scale_for(fmaf(a.rowsum, amax_in, a.bmax));
        if (writer) a.dyn->scale[a.net][a.level + 1] = r.out_scale;
      }
      // pinned scale: the consumer of a level checks that the level sat inside the accuracy window
      if (a.fixed_scale && writer && amax_in > 0.f && amax_in * s_in < 1.0f)
        atomicOr(a.flags, kFlagUnderflow);
    }
  }
  return r;
}

constexpr int kFcBlockM = 128;
// warpgroup 0: warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps 2-3 idle (donate registers)
// warpgroups 1-2: 8 epilogue warps; warp w drains TMEM lane quarter (w & 3), column half ((w - 4) >> 2)
constexpr int kFcThreads = 384;
constexpr int kFcEpiThreads = 256;
constexpr int kFcSmemBytes = 227 * 1024;

template <int S, int BN>
struct FcTcCfg {
  using Sch = Scheme<S>;
  static constexpr int kABytes = kFcBlockM * 128;            // one A plane tile: 128 rows x 128 B
  static constexpr int kBBytes = BN * 128;                   // one W plane tile
  static constexpr int kStageBytes = Sch::kPlanes * (kABytes + kBBytes);
  static constexpr int kAuxBytes = 2048 + 2 * BN * 4;        // barriers + bias tiles
  static constexpr int kStagesRaw = (kFcSmemBytes - 1024 - kAuxBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN;                   // double-buffered accumulator
  static constexpr int kColsPerThread = BN / 2;              // each epilogue thread: one row, half the columns
  static_assert(kStages >= 2, "need at least a double-buffered operand ring");
  static_assert(kTmemCols == 512 || kTmemCols == 256 || kTmemCols == 128, "power-of-two TMEM allocation");
  static_assert(kColsPerThread == 32 || kColsPerThread % 64 == 0, "drain granularity is 64 columns (32 for BN = 64)");
};

// One 32-column group of one row: bias, activation, then either split planes or float32.
template <int S, bool kTrack = true>
__device__ __forceinline__ void fc_epilogue_chunk(const FcArgs& a, const FcScales& sc, const float (&acc)[32],
                                                  const float* sbias, int row, int n0, bool& ovf, float& amx) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = fmaf(acc[j], sc.alpha, sbias[j]);
    if (a.relu) x = fmaxf(x, 0.0f);
    v[j] = x;
  }
  if (a.out_planes) {
    if (n0 >= a.out_kpad) return;
    if constexpr (S == kFp16x3 && kTrack) {
#pragma unroll
      for (int j = 0; j < 32; ++j) amx = fmaxf(amx, fabsf(v[j]));
    }
    constexpr int kWords = 32 * sizeof(E) / 4;            // 32-bit words per plane per group
    uint32_t pk[Sch::kPlanes][kWords];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      if constexpr (S == kFp16x3) {
        uint32_t w[2];
        Sch::split2(v[j], v[j + 1], sc.out_scale, w, &ovf);
        pk[0][j >> 1] = w[0];
        pk[1][j >> 1] = w[1];
        continue;
      }
      E p0[Sch::kPlanes], p1[Sch::kPlanes];
      Sch::split(v[j], sc.out_scale, p0, &ovf);
      Sch::split(v[j + 1], sc.out_scale, p1, &ovf);
#pragma unroll
      for (int q = 0; q < Sch::kPlanes; ++q) {
        if constexpr (sizeof(E) == 4) {
          pk[q][j] = __float_as_uint(*reinterpret_cast<const float*>(&p0[q]));
          pk[q][j + 1] = __float_as_uint(*reinterpret_cast<const float*>(&p1[q]));
        } else {
          const uint32_t lo = *reinterpret_cast<const uint16_t*>(&p0[q]);
          const uint32_t hi = *reinterpret_cast<const uint16_t*>(&p1[q]);
          pk[q][j >> 1] = lo | (hi << 16);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < Sch::kPlanes; ++q) {
      E* dst = reinterpret_cast<E*>(a.out_planes) +
               (static_cast<size_t>(q) * a.out_plane_rows + a.row_off + row) * a.out_kpad + n0;
      // plane rows are kpad * sizeof(E) apart (a multiple of 128 B) and n0 is a multiple of 32 columns: 32-byte aligned
#pragma unroll
      for (int i = 0; i < kWords / 8; ++i)
        st_global_v8(reinterpret_cast<uint32_t*>(dst) + 8 * i, pk[q][8 * i], pk[q][8 * i + 1], pk[q][8 * i + 2],
                     pk[q][8 * i + 3], pk[q][8 * i + 4], pk[q][8 * i + 5], pk[q][8 * i + 6], pk[q][8 * i + 7]);
    }
  } else {
    float* dst = a.out_f32 + static_cast<size_t>(row) * a.out_ld + n0;
    if (n0 + 32 <= a.N && (a.out_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out_f32) & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_global_v8(dst + 8 * i, __float_as_uint(v[8 * i]), __float_as_uint(v[8 * i + 1]), __float_as_uint(v[8 * i + 2]),
                     __float_as_uint(v[8 * i + 3]), __float_as_uint(v[8 * i + 4]), __float_as_uint(v[8 * i + 5]),
                     __float_as_uint(v[8 * i + 6]), __float_as_uint(v[8 * i + 7]));
    } else if (n0 + 32 <= a.N && (a.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < a.N) dst[j] = v[j];
    }
  }
}

template <int S, int BN>
__global__ void __launch_bounds__(kFcThreads, 1)
fc_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const FcArgs a) {
  using Cfg = FcTcCfg<S, BN>;
  using Sch = Scheme<S>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kPlanes = Sch::kPlanes;
  constexpr int kGroups = Cfg::kColsPerThread / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* aux = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [kStages]
  uint64_t* empty_bar = full_bar + kStages;                         // [kStages]
  uint64_t* tfull_bar = empty_bar + kStages;                        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* cta_abort = tmem_slot + 1;
  float* sbias = reinterpret_cast<float*>(aux + 2048);              // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (a.M + kFcBlockM - 1) / kFcBlockM;
  const int n_tiles = (a.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int kbc = a.kb_per_chunk;
  const int n_chunks = (a.num_k_blocks + kbc - 1) / kbc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, kFcEpiThreads / 32);   // one arrive per epilogue warp
    }
    *cta_abort = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ================= TMA producer =================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
          const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
          for (int kb = 0; kb < a.num_k_blocks; ++kb) {
            if (!mbar_wait(empty_bar + stage, phase ^ 1, cta_abort, a.flags)) { ok = false; break; }
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
#pragma unroll
            for (int p = 0; p < kPlanes; ++p)
              tma_load_2d(st + p * Cfg::kABytes, &tmap_a, full_bar + stage, kb * Sch::kBlockK,
                          p * a.a_plane_rows + a.row_off + m_blk * kFcBlockM);
#pragma unroll
            for (int p = 0; p < kPlanes; ++p)
              tma_load_2d(st + kPlanes * Cfg::kABytes + p * Cfg::kBBytes, &tmap_b, full_bar + stage,
                          kb * Sch::kBlockK, p * a.b_plane_rows + n_blk * BN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer (one thread) =================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc(Sch::kFmt, kFcBlockM, BN);
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        uint32_t unit = 0;                            // one unit = one (tile, k-chunk) accumulation
        for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
          for (int c = 0; c < n_chunks && ok; ++c, ++unit) {
            const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
            if (!mbar_wait(tempty_bar + acc, acc_phase ^ 1, cta_abort, a.flags)) { ok = false; break; }
            tc_fence_after_sync();
            const uint32_t tmem_d = tmem_base + acc * BN;
            const int kb_end = min(a.num_k_blocks, (c + 1) * kbc);
            uint32_t fresh = 1;                       // first MMA of the unit overwrites the buffer
            for (int kb = c * kbc; kb < kb_end; ++kb) {
              if (!mbar_wait(full_bar + stage, phase, cta_abort, a.flags)) { ok = false; break; }
              tc_fence_after_sync();
              const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
              const uint32_t sb = sa + kPlanes * Cfg::kABytes;
              // correction passes (hi.lo, lo.hi) first: they land while the accumulator is small, so the
              // tensor core's truncating add costs nothing; the dominant hi.hi pass closes the k-block
#pragma unroll
              for (int ps = Sch::kPasses - 1; ps >= 0; --ps) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {      // 4 k-steps of 32 bytes inside the 128-byte swizzle row
                  const uint64_t da = umma_desc_sw128(sa + pass_a(ps) * Cfg::kABytes + ks * 32);
                  const uint64_t db = umma_desc_sw128(sb + pass_b(ps) * Cfg::kBBytes + ks * 32);
                  umma_ss<Sch::kTf32>(tmem_d, da, db, idesc, fresh ? 0u : 1u);
                  fresh = 0;
                }
              }
              umma_commit(empty_bar + stage);         // smem slot reusable once these MMAs retire
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (ok) umma_commit(tfull_bar + acc);     // chunk accumulated -> epilogue may drain it
          }
        }
      }
    }
  } else {
    // ================= epilogue warps (TMEM -> FP32 registers (RN) -> global) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int ew = warp - 4;                        // 0..7
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int half = ew >> 2;                       // column half
    const int et = ew * 32 + lane;                  // 0..255 among epilogue threads
    bool ovf = false;
    float amx = 0.f;
    const FcScales sc = fc_resolve_scales<S>(a, blockIdx.x == 0 && threadIdx.x == kFcThreads - kFcEpiThreads);
    uint32_t unit = 0;
    int it = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      float* sb = sbias + (it & 1) * BN;
      for (int i = et; i < BN; i += kFcEpiThreads) {
        const int n = n_blk * BN + i;
        sb[i] = (n < a.N) ? __ldg(a.bias + n) : 0.0f;
      }
      named_bar_sync(1, kFcEpiThreads);
      float sum[kGroups][32];
      for (int c = 0; c < n_chunks; ++c, ++unit) {
        const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
        if (!mbar_wait(tfull_bar + acc, acc_phase, cta_abort, a.flags)) { ok = false; break; }
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * Cfg::kColsPerThread;
        if constexpr (kGroups == 1) {               // 64-column tiles (few-row calls): one 32-column drain per thread
          uint32_t r[32];
          tmem_ld32(taddr, r);
          tmem_ld_wait();
          if (c == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[0][j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[0][j] += __uint_as_float(r[j]);
          }
        } else {
#pragma unroll
          for (int g = 0; g < kGroups; g += 2) {      // 64 columns per TMEM load
            uint32_t r[64];
            tmem_ld64(taddr + g * 32, r);
            tmem_ld_wait();
            if (c == 0) {
#pragma unroll
              for (int j = 0; j < 64; ++j) sum[g + (j >> 5)][j & 31] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) sum[g + (j >> 5)][j & 31] += __uint_as_float(r[j]);
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + acc);   // buffer free: MMA warp may start the next chunk
      }
      if (!ok) break;
      const int row = m_blk * kFcBlockM + q * 32 + lane;
      if (row < a.M) {
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const int col = half * Cfg::kColsPerThread + g * 32;
          fc_epilogue_chunk<S>(a, sc, sum[g], sb + col, row, n_blk * BN + col, ovf, amx);
        }
      }
    }
    if (ovf) atomicOr(a.flags, kFlagRange);
    if (a.out_planes) publish_amax<S>(a.dyn, a.net, a.level + 1, amx);
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 output tile.
// Each CTA stages its own 128 activation rows and HALF of the weight tile (128 of the 256 output
// features); one tcgen05.mma issued by the leader CTA reads both shared memories, so the weight bytes
// each SM pulls from L2 are halved (the 1-CTA kernel is L2->SMEM bound: 62 B/clk/SM vs 42 here) and a
// third pipeline stage fits.  Each CTA keeps the accumulator of its own 128 rows in its own TMEM and
// drains / stores it exactly like the 1-CTA kernel.
// Destination planes of the fused all-gather: one 2-D TMA map per rank, each covering THIS rank's row slot of
// that rank's gathered plane ([rows of this call][d_out] float32, box 128 rows x 32 columns, SWIZZLE_128B).
constexpr int kMaxGatherRanks = 8;
struct GatherMaps {
  CUtensorMap m[kMaxGatherRanks];
};

template <int S, bool kGather = false>
struct FcTc2Cfg {
  using Sch = Scheme<S>;
  static constexpr int BN = 256;
  static constexpr int kABytes = kFcBlockM * 128;            // 128 rows x 128 B (this CTA's rows)
  static constexpr int kBBytes = (BN / 2) * 128;             // this CTA's half of the weight tile
  static constexpr int kStageBytes = Sch::kPlanes * (kABytes + kBBytes);
  static constexpr int kAuxBytes = 2048 + 2 * BN * 4;
  // gather variant: 2 column halves x 2 ping-pong staging tiles of 128 rows x 128 B for the TMA stores
  static constexpr int kStoreTileBytes = kFcBlockM * 128;
  static constexpr int kStoreBytes = kGather ? 4 * kStoreTileBytes : 0;
  static constexpr int kStagesRaw = (kFcSmemBytes - 1024 - kAuxBytes - kStoreBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kColsPerThread = BN / 2;
  static constexpr int kSmemBytes = kStages * kStageBytes + kAuxBytes + kStoreBytes + 1024;
  static_assert(kStages >= 2, "need at least a double-buffered operand ring");
  static_assert((kStages * kStageBytes + kAuxBytes) % 1024 == 0, "staging tiles must stay 1024-byte aligned");
};

// kGather: final layer only.  Besides (optionally) the local float32 output, every finished 128 x 32 block is
// staged in shared memory and TMA-stored into the gathered plane of EVERY rank (peer memory over NVLink):
// the all-gather of H-hat happens inside the kernel that produces it, tile by tile.
template <int S, bool kGather = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFcThreads, 1)
fc_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const __grid_constant__ GatherMaps gm, const FcArgs a) {
  using Cfg = FcTc2Cfg<S, kGather>;
  using Sch = Scheme<S>;
  constexpr int BN = Cfg::BN;
  constexpr int kStages = Cfg::kStages;
  constexpr int kPlanes = Sch::kPlanes;
  constexpr int kGroups = Cfg::kColsPerThread / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* aux = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [kStages]  (leader's copy is the live one)
  uint64_t* empty_bar = full_bar + kStages;                         // [kStages]  per CTA, multicast-committed
  uint64_t* tfull_bar = empty_bar + kStages;                        // [2]        per CTA, multicast-committed
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2]        (leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* cta_abort = tmem_slot + 1;
  float* sbias = reinterpret_cast<float*>(aux + 2048);              // [2][BN]
  uint8_t* store_tiles = aux + Cfg::kAuxBytes;                      // gather only: [half][2][128 rows][128 B]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int m_pairs = (a.M + 2 * kFcBlockM - 1) / (2 * kFcBlockM);
  const int n_tiles = (a.N + BN - 1) / BN;
  const int total_tiles = m_pairs * n_tiles;
  const int kbc = a.kb_per_chunk;
  const int n_chunks = (a.num_k_blocks + kbc - 1) / kbc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);                      // leader's arrive.expect_tx; bytes from both CTAs
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, 2 * (kFcEpiThreads / 32));   // epilogue warps of BOTH CTAs
    }
    *cta_abort = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before_sync();
  cluster_sync_all();                                  // barriers of both CTAs initialised before any remote use
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ================= TMA producer (both CTAs) =================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        long long dbg_wait = 0;
        const long long dbg_t0 = MM_DBG(a) ? clock64() : 0;
        for (int tile = cluster_id; tile < total_tiles && ok; tile += n_clusters) {
          const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
          const int row0 = (m_pair * 2 + static_cast<int>(cta_rank)) * kFcBlockM;        // this CTA's activation rows
          const int wrow0 = n_blk * BN + static_cast<int>(cta_rank) * (BN / 2);           // this CTA's weight rows
          // the next tile of this cluster touches other activation rows (tiles are 74 apart): pull them into L2
          // now, so its loads never see HBM latency (3 stages only cover ~2 stage-times of lookahead)
          const int next_tile = tile + n_clusters;
          const int next_row0 = next_tile < total_tiles
                                    ? ((next_tile / n_tiles) * 2 + static_cast<int>(cta_rank)) * kFcBlockM : -1;
          for (int kb = 0; kb < a.num_k_blocks; ++kb) {
            const long long t_w = MM_DBG(a) ? clock64() : 0;
            if (!mbar_wait(empty_bar + stage, phase ^ 1, cta_abort, a.flags)) { ok = false; break; }
            if (MM_DBG(a)) dbg_wait += clock64() - t_w;
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            if (next_row0 >= 0 && a.l2_prefetch) {
#pragma unroll
              for (int p = 0; p < kPlanes; ++p)
                tma_prefetch_l2_2d(&tmap_a, kb * Sch::kBlockK, p * a.a_plane_rows + a.row_off + next_row0);
            }
            if (leader) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);
#pragma unroll
            for (int p = 0; p < kPlanes; ++p)
              tma_load_2d_pair(st + p * Cfg::kABytes, &tmap_a, full_bar + stage, kb * Sch::kBlockK,
                               p * a.a_plane_rows + a.row_off + row0);
#pragma unroll
            for (int p = 0; p < kPlanes; ++p)
              tma_load_2d_pair(st + kPlanes * Cfg::kABytes + p * Cfg::kBBytes, &tmap_b, full_bar + stage,
                               kb * Sch::kBlockK, p * a.b_plane_rows + wrow0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        if (MM_DBG(a) && leader) {
          atomicAdd(a.dbg + 0, static_cast<unsigned long long>(dbg_wait));              // producer: waiting for a free stage
          atomicAdd(a.dbg + 1, static_cast<unsigned long long>(clock64() - dbg_t0));    // producer: total
        }
      }
    } else if (warp == 1 && leader) {
      // ================= MMA issuer (one thread of the leader CTA) =================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc(Sch::kFmt, 2 * kFcBlockM, BN);
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        uint32_t unit = 0;
        long long w_tempty = 0, w_full = 0;
        const long long dbg_t0 = MM_DBG(a) ? clock64() : 0;
        const unsigned long long dbg_g0 = MM_DBG(a) ? globaltimer_ns() : 0ull;
        for (int tile = cluster_id; tile < total_tiles && ok; tile += n_clusters) {
          for (int c = 0; c < n_chunks && ok; ++c, ++unit) {
            const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
            const long long t_e = MM_DBG(a) ? clock64() : 0;
            if (!mbar_wait(tempty_bar + acc, acc_phase ^ 1, cta_abort, a.flags)) { ok = false; break; }
            if (MM_DBG(a)) w_tempty += clock64() - t_e;
            tc_fence_after_sync();
            const uint32_t tmem_d = tmem_base + acc * BN;
            const int kb_end = min(a.num_k_blocks, (c + 1) * kbc);
            uint32_t fresh = 1;
            for (int kb = c * kbc; kb < kb_end; ++kb) {
              const long long t_f = MM_DBG(a) ? clock64() : 0;
              if (!mbar_wait(full_bar + stage, phase, cta_abort, a.flags)) { ok = false; break; }
              if (MM_DBG(a)) w_full += clock64() - t_f;
              tc_fence_after_sync();
              const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
              const uint32_t sb = sa + kPlanes * Cfg::kABytes;
#pragma unroll
              for (int ps = Sch::kPasses - 1; ps >= 0; --ps) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t da = umma_desc_sw128(sa + pass_a(ps) * Cfg::kABytes + ks * 32);
                  const uint64_t db = umma_desc_sw128(sb + pass_b(ps) * Cfg::kBBytes + ks * 32);
                  umma_ss_pair<Sch::kTf32>(tmem_d, da, db, idesc, fresh ? 0u : 1u);
                  fresh = 0;
                }
              }
              umma_commit_pair(empty_bar + stage, 3);   // frees this stage in both CTAs
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (ok) umma_commit_pair(tfull_bar + acc, 3);   // both CTAs' epilogues may drain
          }
        }
        if (MM_DBG(a)) {
          atomicAdd(a.dbg + 2, static_cast<unsigned long long>(w_tempty));               // MMA: waiting for a drained TMEM buffer
          atomicAdd(a.dbg + 3, static_cast<unsigned long long>(w_full));                 // MMA: waiting for operands
          atomicAdd(a.dbg + 4, static_cast<unsigned long long>(clock64() - dbg_t0));     // MMA: total
          atomicAdd(a.dbg + 5, 1ull);                                                    // clusters counted
          atomicAdd(a.dbg + 6, globaltimer_ns() - dbg_g0);                               // MMA: total in ns (wall) ->
                                                                                         // dbg[4] / dbg[6] = SM clock in GHz
        }
      }
    }
  } else {
    // ================= epilogue warps (each CTA drains its own 128 rows) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int ew = warp - 4;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int et = ew * 32 + lane;
    bool ovf = false;
    float amx = 0.f;
    const FcScales sc = fc_resolve_scales<S>(a, blockIdx.x == 0 && threadIdx.x == kFcThreads - kFcEpiThreads);
    uint32_t unit = 0;
    int it = 0;
    bool ok = true;
    for (int tile = cluster_id; tile < total_tiles && ok; tile += n_clusters, ++it) {
      const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
      float* sb = sbias + (it & 1) * BN;
      for (int i = et; i < BN; i += kFcEpiThreads) {
        const int n = n_blk * BN + i;
        sb[i] = (n < a.N) ? __ldg(a.bias + n) : 0.0f;
      }
      named_bar_sync(1, kFcEpiThreads);
      float sum[kGroups][32];
      for (int c = 0; c < n_chunks; ++c, ++unit) {
        const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
        if (!mbar_wait(tfull_bar + acc, acc_phase, cta_abort, a.flags)) { ok = false; break; }
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * Cfg::kColsPerThread;
#pragma unroll
        for (int g = 0; g < kGroups; g += 2) {
          uint32_t r[64];
          tmem_ld64(taddr + g * 32, r);
          tmem_ld_wait();
          if (c == 0) {
#pragma unroll
            for (int j = 0; j < 64; ++j) sum[g + (j >> 5)][j & 31] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) sum[g + (j >> 5)][j & 31] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar + acc);
      }
      if (!ok) break;
      const int row0 = (m_pair * 2 + static_cast<int>(cta_rank)) * kFcBlockM;
      const int row = row0 + q * 32 + lane;
      if (a.out_planes || a.out_f32) {
        if (row < a.M) {
#pragma unroll
          for (int g = 0; g < kGroups; ++g) {
            const int col = half * Cfg::kColsPerThread + g * 32;
            fc_epilogue_chunk<S, !kGather>(a, sc, sum[g], sb + col, row, n_blk * BN + col, ovf, amx);
          }
        }
      }
      if constexpr (kGather) {
        const bool issuer = (q == 0 && lane == 0);          // one thread per column half owns the bulk groups
        const int r = q * 32 + lane;                        // row inside this CTA's tile
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          uint8_t* tile = store_tiles + (half * 2 + (g & 1)) * Cfg::kStoreTileBytes;
          if (issuer) tma_store_wait_read<1>();             // the store that last read this tile has drained it
          named_bar_sync(2 + half, 128);
          const int col = half * Cfg::kColsPerThread + g * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {                     // 8 x 16 B per row, SWIZZLE_128B placement
            float4 v;
            v.x = fmaf(sum[g][4 * c + 0], sc.alpha, sb[col + 4 * c + 0]);
            v.y = fmaf(sum[g][4 * c + 1], sc.alpha, sb[col + 4 * c + 1]);
            v.z = fmaf(sum[g][4 * c + 2], sc.alpha, sb[col + 4 * c + 2]);
            v.w = fmaf(sum[g][4 * c + 3], sc.alpha, sb[col + 4 * c + 3]);
            if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
          }
          fence_proxy_async_smem();
          named_bar_sync(2 + half, 128);
          if (issuer) {
            for (int p = 0; p < a.gather_world; ++p) tma_store_2d(&gm.m[p], tile, n_blk * BN + col, row0);
            tma_store_commit();
          }
        }
      }
    }
    if constexpr (kGather) {
      if (q == 0 && lane == 0) tma_store_wait_all<0>();     // every peer write has left before the CTA retires
    }
    if (ovf) atomicOr(a.flags, kFlagRange);
    if constexpr (!kGather) { if (a.out_planes) publish_amax<S>(a.dyn, a.net, a.level + 1, amx); }
  }

  tc_fence_before_sync();
  cluster_sync_all();                                  // neither CTA may exit while its pair can still touch it
  tc_fence_after_sync();
  if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Push variant of the all-gather (MAMIMO_GATHER_MODE=push): the final layers run the plain (3-stage) kernel into this
// rank's own slot of its gathered planes, and this kernel -- a few CTAs on a side stream -- streams the finished rows
// to the same slot of every peer's plane over NVLink with bulk copies (global -> shared -> peer global), while the
// other SMs compute the next sub-batch.  One thread per CTA drives a ring of kPushBufs buffers: the load of chunk
// i+1 is in flight while the stores of chunk i to all peers drain.
constexpr int kPushChunk = 32 * 1024;
constexpr int kPushBufs = 4;
constexpr int kPushSmem = kPushBufs * kPushChunk + 128;
struct PushArgs {
  const uint8_t* src;                 // first byte of the rows to send (this rank's slot)
  uint8_t* dst[kMaxGatherRanks];      // the same position in every peer's plane
  int n_dst;
  unsigned long long bytes;           // multiple of 16
  uint32_t* flags;
};

__global__ void __launch_bounds__(32) peer_push_kernel(const PushArgs a) {
  extern __shared__ uint8_t push_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(push_raw) + 127) & ~static_cast<uintptr_t>(127));
  __shared__ uint64_t full[kPushBufs];
  __shared__ uint32_t cta_abort;
  if (threadIdx.x != 0) return;
  cta_abort = 0;
  for (int b = 0; b < kPushBufs; ++b) mbar_init(&full[b], 1);
  fence_barrier_init();
  const long long n_chunks = static_cast<long long>((a.bytes + kPushChunk - 1) / kPushChunk);
  auto chunk_bytes = [&](long long c) {
    const unsigned long long off = static_cast<unsigned long long>(c) * kPushChunk;
    return static_cast<uint32_t>(a.bytes - off < static_cast<unsigned long long>(kPushChunk) ? a.bytes - off : kPushChunk);
  };
  long long it = 0, prev = -1;
  for (long long c = blockIdx.x; ; c += gridDim.x, ++it) {
    const bool have = c < n_chunks;
    if (have) {
      const int b = static_cast<int>(it % kPushBufs);
      tma_store_wait_read<kPushBufs - 2>();          // the stores that last read buffer b (4 iterations ago) are done with it
      mbar_arrive_expect_tx(&full[b], chunk_bytes(c));
      bulk_load_1d(buf + b * kPushChunk, a.src + static_cast<unsigned long long>(c) * kPushChunk, chunk_bytes(c), &full[b]);
    }
    if (prev >= 0) {                                  // chunk of the previous iteration: landed -> send to every peer
      const long long pit = it - 1;
      const int b = static_cast<int>(pit % kPushBufs);
      if (!mbar_wait(&full[b], static_cast<uint32_t>((pit / kPushBufs) & 1), &cta_abort, a.flags)) return;
      const unsigned long long off = static_cast<unsigned long long>(prev) * kPushChunk;
      for (int p = 0; p < a.n_dst; ++p) bulk_store_1d(a.dst[p] + off, buf + b * kPushChunk, chunk_bytes(prev));
      tma_store_commit();
    }
    if (!have) break;
    prev = c;
  }
  tma_store_wait_all<0>();
}

// Multicast form of the same step (planes attached with NVSwitch multicast addresses, mamimo_gather_attach): every
// 16-byte vector is read once from this rank's slot and stored ONCE to the multicast address with multimem.st -- the
// switch replicates it into every rank's plane, so a rank's egress link carries its rows once instead of world-1
// times.  Coalesced: a warp covers 512 contiguous bytes per store instruction; kPushMcUnroll vectors per thread in
// flight.
constexpr int kPushMcThreads = 512;
constexpr int kPushMcUnroll = 8;
struct PushMcArgs {
  const float4* src;        // this rank's slot in its own plane
  float4* mc;               // the same position behind the multicast address
  unsigned long long n_vec; // 16-byte vectors
};
__device__ __forceinline__ void multimem_st_v4(float4* mc, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint64_t>(mc)), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__global__ void __launch_bounds__(kPushMcThreads) peer_push_mc_kernel(const PushMcArgs a) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kPushMcThreads;
  unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * kPushMcThreads + threadIdx.x;
  for (; i + (kPushMcUnroll - 1) * stride < a.n_vec; i += kPushMcUnroll * stride) {
    float4 v[kPushMcUnroll];
#pragma unroll
    for (int u = 0; u < kPushMcUnroll; ++u) v[u] = __ldcg(a.src + i + u * stride);
#pragma unroll
    for (int u = 0; u < kPushMcUnroll; ++u) multimem_st_v4(a.mc + i + u * stride, v[u]);
  }
  for (; i < a.n_vec; i += stride) multimem_st_v4(a.mc + i, __ldcg(a.src + i));
}

// ------------------------------------------------------------------------------------------
// Exact FP32 CUDA-core GEMM: 128x128 tile, 256 threads, 8x8 micro-tile, BK = 16.
// A [rows_alloc][kpad] and W [Npad][kpad] are both K-major and zero padded, so no K masking.
template <int S>
__global__ void __launch_bounds__(256) fc_simt_kernel(const FcArgs a) {
  static_assert(S == kFp32Simt, "SIMT path stores single-plane fp32 operands");
  constexpr int BM = 128, BNs = 128, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BNs + 4];
  const int n_tiles = (a.N + BNs - 1) / BNs;
  const int m_blk = blockIdx.x / n_tiles, n_blk = blockIdx.x % n_tiles;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float* Ab = a.A + static_cast<size_t>(m_blk) * BM * a.kpad;
  const float* Wb = a.W + static_cast<size_t>(n_blk) * BNs * a.kpad;
  for (int k0 = 0; k0 < a.kpad; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + i * 256;        // 512 float4 per operand tile
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const float4 va = *reinterpret_cast<const float4*>(Ab + static_cast<size_t>(r) * a.kpad + k0 + kq);
      As[kq + 0][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
      const float4 vw = *reinterpret_cast<const float4*>(Wb + static_cast<size_t>(r) * a.kpad + k0 + kq);
      Ws[kq + 0][r] = vw.x; Ws[kq + 1][r] = vw.y; Ws[kq + 2][r] = vw.z; Ws[kq + 3][r] = vw.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float ar[8], wr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ar[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) wr[j] = Ws[kk][tx * 8 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_blk * BM + ty * 8 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n_blk * BNs + tx * 8 + j;
      float x = 0.f;
      if (n < a.N) {
        x = fmaf(acc[i][j], a.alpha, __ldg(a.bias + n));
        if (a.relu) x = fmaxf(x, 0.f);
      }
      if (a.out_planes) {
        if (n < a.out_kpad)
          reinterpret_cast<float*>(a.out_planes)[static_cast<size_t>(row) * a.out_kpad + n] = x * a.out_scale;
      } else if (n < a.N) {
        a.out_f32[static_cast<size_t>(row) * a.out_ld + n] = x;
      }
    }
  }
}

}  // namespace mm
