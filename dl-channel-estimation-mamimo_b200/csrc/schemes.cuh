// Operand storage schemes of the FC layers.
//
// The FC layers must agree with an FP32 TensorFlow reference to <= 1e-5 relative L2
// (BASELINE.json north_star), but tcgen05 has no FP32 input kind.  Every FP32 operand
// x is therefore stored as a small sum of low-precision "planes" x ~= p0 + p1 and the
// product A.B is accumulated (FP32, in TMEM) as p0.q0 + p0.q1 + p1.q0 -- the dropped
// p1.q1 term is ~2^-22 relative.  A plane is a K-major [rows][Kpad] matrix; planes of one
// operand are stacked along the row axis so a single 2-D TMA map addresses all of them.
//
//   scheme     elem  planes  MMA kind    passes  operand error      range
//   FP32_SIMT  f32   1       (FFMA)      -       exact fp32         fp32
//   TF32X3     f32   2       kind::tf32  3       ~2^-21             fp32 (robust)
//   FP16X3     f16   2       kind::f16   3       ~2^-22 (scaled)    |x|*2^s < 65504, checked
//   BF16X1     bf16  1       kind::f16   1       2^-9  (diagnostic) fp32
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

namespace mm {

enum : int { kFp32Simt = 0, kTf32x3 = 1, kFp16x3 = 2, kBf16x1 = 3 };

template <int S>
struct Scheme;

template <>
struct Scheme<kFp32Simt> {
  using elem = float;
  static constexpr int kPlanes = 1;
  static constexpr int kBlockK = 32;        // K padding granule (elements)
  __host__ __device__ static void split(float x, float scale, elem* p, bool* overflow) {
    (void)overflow;
    p[0] = x * scale;
  }
};

template <>
struct Scheme<kTf32x3> {
  using elem = float;                       // tf32 operands live in 32-bit containers
  static constexpr int kPlanes = 2;
  static constexpr int kBlockK = 32;        // 128-byte swizzle row
  static constexpr int kUmmaK = 8;
  static constexpr uint32_t kFmt = 2;       // F16F32Format::TF32
  static constexpr bool kTf32 = true;
  static constexpr int kPasses = 3;
  __host__ __device__ static float round_tf32(float x) {
    // round-to-nearest (ties away) to 10 explicit mantissa bits, done on the bit pattern so
    // host (weight split) and device (activation split) agree exactly
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(x);
#else
    memcpy(&u, &x, 4);
#endif
    if ((u & 0x7f800000u) == 0x7f800000u) return x;    // inf / nan untouched
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
#ifdef __CUDA_ARCH__
    r = __uint_as_float(u);
#else
    memcpy(&r, &u, 4);
#endif
    return r;
  }
  __host__ __device__ static void split(float x, float scale, elem* p, bool* overflow) {
    (void)overflow;
    const float v = x * scale;
    const float hi = round_tf32(v);
    p[0] = hi;
    p[1] = v - hi;          // exact in fp32; the MMA truncates it to tf32 (error ~2^-23 |v|)
  }
};

template <>
struct Scheme<kFp16x3> {
  using elem = __half;
  static constexpr int kPlanes = 2;
  static constexpr int kBlockK = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kFmt = 0;       // F16
  static constexpr bool kTf32 = false;
  static constexpr int kPasses = 3;
  __host__ __device__ static void split(float x, float scale, elem* p, bool* overflow) {
    const float v = x * scale;              // scale is a power of two: exact
    if (!(fabsf(v) <= 65504.0f)) *overflow = true;
    const __half hi = __float2half_rn(v);
    p[0] = hi;
    p[1] = __float2half_rn(v - __half2float(hi));
  }
#ifdef __CUDACC__
  // two values at once with packed conversions (F2FP / HADD2 on the ALU instead of scalar F2F on the XU pipe,
  // which ncu showed 44 % busy in the LS kernel); same round-to-nearest results as split().
  // w[plane] = (plane value of x0) | (plane value of x1) << 16
  __device__ __forceinline__ static void split2(float x0, float x1, float scale, uint32_t (&w)[2], bool* overflow) {
    const float v0 = x0 * scale, v1 = x1 * scale;
    if (!(fabsf(v0) <= 65504.0f) || !(fabsf(v1) <= 65504.0f)) *overflow = true;
    const __half2 hi = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(hi);
    const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    w[0] = *reinterpret_cast<const uint32_t*>(&hi);
    w[1] = *reinterpret_cast<const uint32_t*>(&lo);
  }
#endif
};

template <>
struct Scheme<kBf16x1> {
  using elem = __nv_bfloat16;
  static constexpr int kPlanes = 1;
  static constexpr int kBlockK = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kFmt = 1;       // BF16
  static constexpr bool kTf32 = false;
  static constexpr int kPasses = 1;
  __host__ __device__ static void split(float x, float scale, elem* p, bool* overflow) {
    (void)overflow;
    p[0] = __float2bfloat16_rn(x * scale);
  }
};

// (A plane, B plane) of pass i -- smallest terms last so the dominant product opens the accumulator
__host__ __device__ constexpr int pass_a(int i) { return i == 2 ? 1 : 0; }
__host__ __device__ constexpr int pass_b(int i) { return i == 1 ? 1 : 0; }

// ---- range management of the FP16X3 scheme ------------------------------------------------------------
// fp16 hi+lo planes carry ~22 bits only while the scaled values sit inside fp16's normal range: the residual
// plane goes subnormal (absolute error 2^-25 in scaled units) once |x|*scale drops below ~2^-2, and anything
// above 65504 overflows.  So the operand scale of every activation level is a power of two chosen ON THE
// DEVICE, per call, from a bound on that level's largest magnitude:
//   level 0 (net input)      : exact amax of the caller's input (one streaming pre-pass) x the LS gain
//   level l+1 (hidden layer) : max_n sum_k |W_l[k][n]| * amax(level l, measured by its producer) + max |b_l|
// which puts the level's amax in [2^7, 2^15) scaled (bound looseness <= K) and leaves >= 10 binades below the
// typical row before a weak row loses relative accuracy.  Powers of two make the scaling itself exact.
// act_scale_log2 != 0 pins one fixed scale instead (no pre-pass); then overflow AND underflow are flagged.
constexpr int kMaxLevels = 9;            // activation levels per net: input of layer 0 .. input of layer 8
struct DynState {
  uint32_t in_amax[2];                   // float bits: amax of the raw input feeding net 0 / net 1 (LS: [0] only; for large
                                         // Y a 1-in-8 cache-line SAMPLE: see ls_resolve_scale for why that is still exact)
  float scale_prov;                      // LS: provisional level-0 scale of pass 0
  uint32_t pad;
  uint32_t amax[2][kMaxLevels];          // float bits: measured amax of the level's (unscaled) activations
  float scale[2][kMaxLevels];            // scale the level's operand planes were written with
};

// largest power of two s <= 2^60 with bound * s < 2^15 (fp16 max is ~2^16: one binade of rounding headroom)
__host__ __device__ inline float pow2_scale_for(float bound) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(bound);
#else
  memcpy(&u, &bound, 4);
#endif
  uint32_t e = (u >> 23) & 0xffu;        // bound in [2^(e-127), 2^(e-126))
  if (e == 0xffu) e = 0xfeu;             // inf / nan: the split flags the overflow
  if (e < 81u) e = 81u;                  // bound < 2^-46 (incl. 0): cap the scale at 2^60
  u = (268u - e) << 23;                  // 2^(141 - e)
  float s;
#ifdef __CUDA_ARCH__
  s = __uint_as_float(u);
#else
  memcpy(&s, &u, 4);
#endif
  return s;
}

#ifdef __CUDACC__
// whole warps only (every thread of the warp calls it once, at the end of the kernel)
template <int S>
__device__ __forceinline__ void publish_amax(DynState* dyn, int net, int level, float amx) {
  if constexpr (S == kFp16x3) {
    if (dyn) {
      const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(amx));   // amx >= 0: uint order == float order
      if ((threadIdx.x & 31) == 0 && m) atomicMax(&dyn->amax[net][level], m);
    }
  }
}
#endif

// A K-major operand: `planes` stacked [plane][rows_alloc][kpad] matrices of elem
struct Operand {
  void* ptr = nullptr;
  int planes = 0;
  int rows_alloc = 0;      // plane stride in rows (multiple of the row tile)
  int kpad = 0;            // row pitch in elements (multiple of kBlockK)
  int elem_bytes = 0;
  size_t bytes() const { return static_cast<size_t>(planes) * rows_alloc * kpad * elem_bytes; }
};

}  // namespace mm
