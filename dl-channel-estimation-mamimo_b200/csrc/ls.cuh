// LS channel estimate + comb-pilot interpolation + operand staging (HBM-bound stage).
//
// Replaces the Nr x Nt interpreted loop of
//   packet_generation/phased_arr/helperMIMOChannelEstimate.m:33-36
//     hD(:,j,i) = rxsym*Puse(:,j)./denom ,  Puse = P', denom = nltf.*ltf(ind)
// for a whole batch of packets in one launch:
//   H[p,i,j,k] = ( sum_n Y[p,i,n,k] * conj(P[j,n]) ) * inv_denom[k],  inv_denom = 1/(nltf * X_pilot)
//
// One CTA owns (packet, rx antenna, tile of pilots).  Phase 1: each thread owns one pilot
// tone, pulls its n_ltf received symbols with coalesced float2/double2 loads (k is the
// contiguous axis of Y), despreads them in registers -- a fast Walsh-Hadamard transform when
// P is Sylvester-Hadamard, a dense complex matvec against P in shared memory otherwise --
// and parks the Nt results in shared memory.  Phase 2: the CTA sweeps (tx, k) with k fastest,
// interpolates between pilots when n_ps > 1 (n_ps == 1 is a pure copy: bit-exact identity),
// and writes (a) the user-visible H_ls and (b) the first FC layer's split operand planes for
// the real and the imaginary net, all with coalesced stores.
#pragma once
#include "ptx.cuh"
#include "schemes.cuh"

namespace mm {

struct LsArgs {
  const void* Y;          // complex [n_pkt][n_rx][n_ltf][n_sc_in]  (float2 or double2)
  const float2* P;        // [n_tx][n_ltf] (dense path only)
  const float2* inv_den;  // [n_pil]  1 / (n_ltf * X_pilot)
  void* H_ls;             // optional complex [n_pkt][n_rx][n_tx][n_sc] (float2 or double2)
  void* planes[2];        // optional: layer-1 operand of net 0 (real part) / net 1 (imag part)
  int plane_rows;         // rows_alloc of those operands
  int kpad;               // row pitch (elements)
  float scale;            // operand scale (power of two)
  int n_pkt, n_rx, n_tx, n_ltf, n_sc, n_ps, n_pil;
  int pil_per_tile;       // pilots per CTA
  int y_double, h_double;
  uint32_t* flags;
  // FP16X3 range management (schemes.cuh): scale resolved on the device from dyn->in_amax[0] * in_gain unless
  // fixed_scale; the kernel publishes the scale it used and the measured amax of both planes as level 0
  DynState* dyn;
  float in_gain;          // bound on |H component| / amax|Y component| (P row sums, 1/|nltf x|, interpolation)
  int fixed_scale;
  int row_off;            // first pair row of this call inside the operand planes (sub-batches of a pipelined step)
  int pass;               // automatic scale only: 0 = provisional pass, 1 = verify (and redo if needed) pass
  int quiet_ovf;          // set by the kernel for pass 0: an overflow of the provisional scale is repaired by pass 1
};

// Level-0 scale of the FP16X3 planes.  Pinned scale: a.scale.  Automatic scale, without a full extra pass over Y:
//   pass 0  uses a PROVISIONAL power of two from a sampled amax of Y (1 cache line in 8, every row of every slab) with
//           one binade of headroom, measures the EXACT amax of the H_ls it emits, and keeps overflow quiet;
//   pass 1  (the same kernel launched again) checks that exact amax against the provisional scale.  Inside the window
//           [2^8, 60000] scaled -- no overflow happened, >= 11 binades above the fp16 residual floor -- every CTA returns
//           at once (a ~3 us launch).  Otherwise the tiles are recomputed with the scale the exact amax calls for.
// Either way the planes that layer 0 reads were written with a scale validated against the exact amax: correct for any
// input, deterministic, and the common case costs 1/8 of a read of Y instead of a full one.
template <int S>
__device__ __forceinline__ float ls_resolve_scale(const LsArgs& a, bool& skip, int& quiet_ovf) {
  skip = false;
  quiet_ovf = 0;
  if constexpr (S == kFp16x3) {
    if (a.dyn && a.planes[0]) {
      const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
      float s = a.scale;
      if (!a.fixed_scale) {
        if (a.pass == 0) {
          s = pow2_scale_for(2.0f * a.in_gain * __uint_as_float(a.dyn->in_amax[0]));
          quiet_ovf = 1;
          if (writer) a.dyn->scale_prov = s;
        } else {
          const float amax = fmaxf(__uint_as_float(a.dyn->amax[0][0]), __uint_as_float(a.dyn->amax[1][0]));
          const float sp = a.dyn->scale_prov;
          const float v = amax * sp;
          if ((v <= 60000.0f && v >= 256.0f) || amax == 0.0f) { skip = true; return sp; }
          s = pow2_scale_for(amax * 1.000001f);
        }
      }
      if (writer) { a.dyn->scale[0][0] = s; a.dyn->scale[1][0] = s; }
      return s;
    }
  }
  if (a.pass != 0) skip = true;           // nothing to verify for the other schemes / pinned scales
  return a.scale;
}

template <int NLTF>
__device__ __forceinline__ void fwht(float2 (&v)[NLTF]) {
#pragma unroll
  for (int h = 1; h < NLTF; h <<= 1) {
#pragma unroll
    for (int i = 0; i < NLTF; ++i) {
      if ((i & h) == 0) {
        const float2 a = v[i], b = v[i + h];
        v[i] = make_float2(a.x + b.x, a.y + b.y);
        v[i + h] = make_float2(a.x - b.x, a.y - b.y);
      }
    }
  }
}

__device__ __forceinline__ float2 ld_y(const void* Y, size_t idx, int is_double) {
  if (is_double) {
    const double2 d = __ldg(reinterpret_cast<const double2*>(Y) + idx);
    return make_float2(static_cast<float>(d.x), static_cast<float>(d.y));
  }
  return __ldg(reinterpret_cast<const float2*>(Y) + idx);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 4 consecutive tones (k .. k+3) of pair row `row`: optional H_ls plus the split operand planes of both nets,
// 8/16/32-byte stores
template <int S>
__device__ __forceinline__ void ls_store4(const LsArgs& a, size_t row, int k, const float (&re)[4], const float (&im)[4],
                                          bool& ovf, float (&amx)[2]) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  if (a.H_ls) {
    if (a.h_double) {
      double2* d = reinterpret_cast<double2*>(a.H_ls) + row * a.n_sc + k;
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = make_double2(re[i], im[i]);
    } else {
      float4* d = reinterpret_cast<float4*>(reinterpret_cast<float2*>(a.H_ls) + row * a.n_sc + k);
      d[0] = make_float4(re[0], im[0], re[1], im[1]);
      d[1] = make_float4(re[2], im[2], re[3], im[3]);
    }
  }
  if (a.planes[0]) {
    if constexpr (S == kFp16x3) {
      amx[0] = fmaxf(amx[0], fmaxf(fmaxf(fabsf(re[0]), fabsf(re[1])), fmaxf(fabsf(re[2]), fabsf(re[3]))));
      amx[1] = fmaxf(amx[1], fmaxf(fmaxf(fabsf(im[0]), fabsf(im[1])), fmaxf(fabsf(im[2]), fabsf(im[3]))));
      uint32_t r01[2], r23[2], i01[2], i23[2];
      Sch::split2(re[0], re[1], a.scale, r01, &ovf);
      Sch::split2(re[2], re[3], a.scale, r23, &ovf);
      Sch::split2(im[0], im[1], a.scale, i01, &ovf);
      Sch::split2(im[2], im[3], a.scale, i23, &ovf);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const size_t off = (static_cast<size_t>(pl) * a.plane_rows + a.row_off + row) * a.kpad + k;
        *reinterpret_cast<uint2*>(reinterpret_cast<E*>(a.planes[0]) + off) = make_uint2(r01[pl], r23[pl]);
        *reinterpret_cast<uint2*>(reinterpret_cast<E*>(a.planes[1]) + off) = make_uint2(i01[pl], i23[pl]);
      }
    } else {
      E pr[4][Sch::kPlanes], pi[4][Sch::kPlanes];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Sch::split(re[i], a.scale, pr[i], &ovf);
        Sch::split(im[i], a.scale, pi[i], &ovf);
      }
#pragma unroll
      for (int pl = 0; pl < Sch::kPlanes; ++pl) {
        const size_t off = (static_cast<size_t>(pl) * a.plane_rows + a.row_off + row) * a.kpad + k;
        E* d0 = reinterpret_cast<E*>(a.planes[0]) + off;
        E* d1 = reinterpret_cast<E*>(a.planes[1]) + off;
        if constexpr (sizeof(E) == 4) {
          *reinterpret_cast<float4*>(d0) = make_float4(pr[0][pl], pr[1][pl], pr[2][pl], pr[3][pl]);
          *reinterpret_cast<float4*>(d1) = make_float4(pi[0][pl], pi[1][pl], pi[2][pl], pi[3][pl]);
        } else {
          auto bits = [](E x) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&x)); };
          *reinterpret_cast<uint2*>(d0) = make_uint2(bits(pr[0][pl]) | (bits(pr[1][pl]) << 16),
                                                     bits(pr[2][pl]) | (bits(pr[3][pl]) << 16));
          *reinterpret_cast<uint2*>(d1) = make_uint2(bits(pi[0][pl]) | (bits(pi[1][pl]) << 16),
                                                     bits(pi[2][pl]) | (bits(pi[3][pl]) << 16));
        }
      }
    }
  }
}

// Phase 2 of the LS kernels: the CTA sweeps (tx, k) of its tile with k fastest, interpolates between pilots when
// n_ps > 1 (n_ps == 1 is a pure copy: bit-exact identity) and writes H_ls and the split operand planes.
template <int S>
__device__ __forceinline__ void ls_emit(const LsArgs& a, const float2* sh, int pitch, int prx, int pil0, int lo,
                                        float (&amx)[2]) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  // ---- phase 2: interpolate + emit ------------------------------------------------------
  const int k0 = pil0 * a.n_ps;
  const int k1 = (pil0 + a.pil_per_tile >= a.n_pil) ? a.n_sc : min(a.n_sc, (pil0 + a.pil_per_tile) * a.n_ps);
  const int nk = k1 - k0;
  bool ovf = false;
  const size_t row0 = static_cast<size_t>(prx) * a.n_tx;
  if (a.n_ps == 1 && (nk & 3) == 0 && (a.n_sc & 3) == 0 && (a.kpad & 3) == 0) {
    // fast path (every reference call site): 4 consecutive tones per thread, 8/16/32-byte stores
    const int nq = nk >> 2;
    const int lg = 31 - __clz(nq);                       // tiles are 128 tones: nq = 32, shifts instead of a division
    const bool pow2 = (nq & (nq - 1)) == 0;
    for (int idx = threadIdx.x; idx < a.n_tx * nq; idx += blockDim.x) {
      const int j = pow2 ? (idx >> lg) : (idx / nq);
      const int kk = (idx - j * nq) << 2;
      const float4 v01 = *reinterpret_cast<const float4*>(sh + j * pitch + kk);
      const float4 v23 = *reinterpret_cast<const float4*>(sh + j * pitch + kk + 2);
      const float re[4] = {v01.x, v01.z, v23.x, v23.z};
      const float im[4] = {v01.y, v01.w, v23.y, v23.w};
      ls_store4<S>(a, row0 + j, k0 + kk, re, im, ovf, amx);
    }
  } else {
    const float inv_nps = 1.0f / static_cast<float>(a.n_ps);
    for (int idx = threadIdx.x; idx < a.n_tx * nk; idx += blockDim.x) {
      const int j = idx / nk;
      const int k = k0 + (idx - j * nk);
      float2 h;
      if (a.n_ps == 1) {
        h = sh[j * pitch + (k - k0)];
      } else if (a.n_pil == 1) {
        h = sh[j * pitch];
      } else {
        const int seg = min(k / a.n_ps, a.n_pil - 2);
        const float w = static_cast<float>(k - seg * a.n_ps) * inv_nps;
        const float2 h0 = sh[j * pitch + (seg - lo)];
        const float2 h1 = sh[j * pitch + (seg + 1 - lo)];
        h = make_float2(h0.x + w * (h1.x - h0.x), h0.y + w * (h1.y - h0.y));
      }
      const size_t row = row0 + j;
      if (a.H_ls) {
        if (a.h_double) reinterpret_cast<double2*>(a.H_ls)[row * a.n_sc + k] = make_double2(h.x, h.y);
        else reinterpret_cast<float2*>(a.H_ls)[row * a.n_sc + k] = h;
      }
      if (a.planes[0]) {
        amx[0] = fmaxf(amx[0], fabsf(h.x));
        amx[1] = fmaxf(amx[1], fabsf(h.y));
        E pr[Sch::kPlanes], pi[Sch::kPlanes];
        Sch::split(h.x, a.scale, pr, &ovf);
        Sch::split(h.y, a.scale, pi, &ovf);
#pragma unroll
        for (int pl = 0; pl < Sch::kPlanes; ++pl) {
          const size_t off = (static_cast<size_t>(pl) * a.plane_rows + a.row_off + row) * a.kpad + k;
          reinterpret_cast<E*>(a.planes[0])[off] = pr[pl];
          reinterpret_cast<E*>(a.planes[1])[off] = pi[pl];
        }
      }
    }
  }
  if (ovf && !a.quiet_ovf) atomicOr(a.flags, kFlagRange);
}

// HAD: P is Sylvester-Hadamard and n_tx == n_ltf == NLTF -> FWHT despread.  Otherwise a dense complex
// matvec against P (shared memory); NLTF > 0 unrolls it over registers, NLTF == 0 is the any-size
// (n_ltf <= 64) fallback.
template <int S, int NLTF, bool HAD>
__global__ void __launch_bounds__(128) ls_kernel(const LsArgs a_in) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  extern __shared__ float2 sm_ls[];
  LsArgs a = a_in;
  bool skip;
  a.scale = ls_resolve_scale<S>(a_in, skip, a.quiet_ovf);
  if (skip) return;
  const int n_tiles = (a.n_pil + a.pil_per_tile - 1) / a.pil_per_tile;
  const int tile = blockIdx.x % n_tiles;
  const int prx = blockIdx.x / n_tiles;                 // pkt * n_rx + rx
  const int pil0 = tile * a.pil_per_tile;
  // pilots held by this CTA: [lo, hi) = tile pilots plus one halo pilot on each side (interp only)
  const int lo = (a.n_ps > 1 && pil0 > 0) ? pil0 - 1 : pil0;
  int hi = min(pil0 + a.pil_per_tile, a.n_pil);
  if (a.n_ps > 1 && hi < a.n_pil) hi += 1;
  const int n_hold = hi - lo;
  const int pitch = a.pil_per_tile + 4;                 // even pitch: rows stay 16-byte aligned for float4 reads
  float2* sh = sm_ls;                                   // [n_tx][pitch]
  float2* sP = sm_ls + static_cast<size_t>(a.n_tx) * pitch;   // dense path: [n_tx][n_ltf]

  if constexpr (!HAD) {
    for (int i = threadIdx.x; i < a.n_tx * a.n_ltf; i += blockDim.x) sP[i] = a.P[i];
    __syncthreads();
  }

  // ---- phase 1: despread at pilot tones -------------------------------------------------
  const size_t y_base = static_cast<size_t>(prx) * a.n_ltf * a.n_sc;
  for (int t = threadIdx.x; t < n_hold; t += blockDim.x) {
    const int pil = lo + t;
    const int k = pil * a.n_ps;
    const float2 inv = __ldg(a.inv_den + pil);
    if constexpr (HAD) {
      // H_NLTF = H_NB (x) H_BLK: 16-point transforms in registers (keeps the kernel at <= 6 CTAs/SM worth
      // of registers), the outer NB-point stage through this thread's own shared-memory column
      constexpr int BLK = NLTF < 16 ? NLTF : 16;
      constexpr int NB = NLTF / BLK;
#pragma unroll 1
      for (int b = 0; b < NB; ++b) {      // not unrolled: 16 loads in flight per thread, ~70 registers
        float2 v[BLK];
#pragma unroll
        for (int n = 0; n < BLK; ++n)
          v[n] = ld_y(a.Y, y_base + static_cast<size_t>(b * BLK + n) * a.n_sc + k, a.y_double);
        fwht<BLK>(v);
#pragma unroll
        for (int j = 0; j < BLK; ++j) sh[(b * BLK + j) * pitch + t] = (NB == 1) ? cmul(v[j], inv) : v[j];
      }
      if constexpr (NB > 1) {
#pragma unroll 4
        for (int j = 0; j < BLK; ++j) {
          float2 u[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) u[b] = sh[(b * BLK + j) * pitch + t];
          fwht<NB>(u);
#pragma unroll
          for (int b = 0; b < NB; ++b) sh[(b * BLK + j) * pitch + t] = cmul(u[b], inv);
        }
      }
    } else if constexpr (NLTF > 0) {
      float2 v[NLTF];
#pragma unroll
      for (int n = 0; n < NLTF; ++n) v[n] = ld_y(a.Y, y_base + static_cast<size_t>(n) * a.n_sc + k, a.y_double);
      for (int j = 0; j < a.n_tx; ++j) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = 0; n < NLTF; ++n) {
          const float2 p = sP[j * NLTF + n];              // y * conj(p)
          acc.x += v[n].x * p.x + v[n].y * p.y;
          acc.y += v[n].y * p.x - v[n].x * p.y;
        }
        sh[j * pitch + t] = cmul(acc, inv);
      }
    } else {
      for (int j = 0; j < a.n_tx; ++j) {
        float2 acc = make_float2(0.f, 0.f);
        for (int n = 0; n < a.n_ltf; ++n) {
          const float2 y = ld_y(a.Y, y_base + static_cast<size_t>(n) * a.n_sc + k, a.y_double);
          const float2 p = sP[j * a.n_ltf + n];
          acc.x += y.x * p.x + y.y * p.y;
          acc.y += y.y * p.x - y.x * p.y;
        }
        sh[j * pitch + t] = cmul(acc, inv);
      }
    }
  }
  __syncthreads();

  float amx[2] = {0.f, 0.f};
  ls_emit<S>(a, sh, pitch, prx, pil0, lo, amx);
  if (a.planes[0]) { publish_amax<S>(a.dyn, 0, 0, amx[0]); publish_amax<S>(a.dyn, 1, 0, amx[1]); }
}

// Hadamard despread with the transform split over threads (n_ps == 1, NLTF = 32 or 64): a CTA owns 64 tones;
// thread (tone t, block b) pulls 16 symbols, does a 16-point FWHT in registers and parks it in shared memory;
// after a barrier the NB-point outer stage runs on the thread's share of the rows.  64 * NB threads, ~17 KB
// (NLTF 32) / 35 KB (NLTF 64) of shared memory and ~50 registers: twice the resident warps of the
// one-thread-per-tone kernel, and every thread has 16 independent loads in flight.
template <int S, int NLTF, int T = 64>
__global__ void __launch_bounds__(T * (NLTF / 16)) ls_had_split_kernel(const LsArgs a) {
  constexpr int BLK = 16, NB = NLTF / BLK;
  extern __shared__ float2 sm_ls[];
  LsArgs a2 = a;
  a2.pil_per_tile = T;
  bool skip;
  a2.scale = ls_resolve_scale<S>(a, skip, a2.quiet_ovf);
  if (skip) return;
  const int n_tiles = (a.n_pil + T - 1) / T;
  const int tile = blockIdx.x % n_tiles;
  const int prx = blockIdx.x / n_tiles;
  const int pil0 = tile * T;
  const int n_here = min(T, a.n_pil - pil0);
  const int pitch = T + 4;
  float2* sh = sm_ls;                                   // [NLTF][pitch]
  const int t = threadIdx.x & (T - 1);
  const int b = threadIdx.x / T;
  const size_t y_base = static_cast<size_t>(prx) * a.n_ltf * a.n_sc + pil0 + t;
  if (t < n_here) {
    float2 v[BLK];
#pragma unroll
    for (int n = 0; n < BLK; ++n) v[n] = ld_y(a.Y, y_base + static_cast<size_t>(b * BLK + n) * a.n_sc, a.y_double);
    fwht<BLK>(v);
#pragma unroll
    for (int j = 0; j < BLK; ++j) sh[(b * BLK + j) * pitch + t] = v[j];
  }
  __syncthreads();
  if (t < n_here) {
    const float2 inv = __ldg(a.inv_den + pil0 + t);
#pragma unroll
    for (int jj = 0; jj < BLK / NB; ++jj) {
      const int j = b + jj * NB;
      float2 u[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) u[c] = sh[(c * BLK + j) * pitch + t];
      fwht<NB>(u);
#pragma unroll
      for (int c = 0; c < NB; ++c) sh[(c * BLK + j) * pitch + t] = cmul(u[c], inv);
    }
  }
  __syncthreads();
  float amx[2] = {0.f, 0.f};
  ls_emit<S>(a2, sh, pitch, prx, pil0, pil0, amx);
  if (a.planes[0]) { publish_amax<S>(a.dyn, 0, 0, amx[0]); publish_amax<S>(a.dyn, 1, 0, amx[1]); }
}

// Persistent TMA-fed variant of ls_had_split_kernel (complex64 Y, n_ps == 1, NLTF = 32 or 64): the [NLTF x 64 tones]
// slab of one (packet, rx, tone tile) is staged into shared memory by ONE cp.async.bulk.tensor.2d per tile
// (SASS: UTMALDG.2D) through an mbarrier ring of STAGES buffers, issued STAGES-1 tiles ahead, so a tile's HBM read
// is in flight while the previous tiles are transformed and emitted (the plain kernel serialises load -> FWHT ->
// emit per CTA and relies on 9 resident CTAs to cover it).  The transform runs IN PLACE in the stage buffer (each
// thread only rewrites the entries it read), so a stage is also the emit source and no separate work buffer is
// needed: 32 KB (NLTF 32) / 64 KB (NLTF 64) per CTA with 2 stages.  CTAs walk the tile list with stride gridDim.x:
// concurrently running CTAs read neighbouring 512-byte segments of the same rows.
// NPS > 1 (comb pilots, the north_star's interpolation stage): the stage rows hold 72 tones, i.e. the tile plus the
// halo pilot the last segment interpolates towards; only the pilot columns are despread (in place), and the emit
// phase interpolates linearly between neighbouring pilots (same arithmetic as the generic kernel / oracle.interp).
// T = tones per tile: 64, or 32 for 64 antennas without comb pilots (16 KB stages instead of 32 KB: 6 resident CTAs per
// SM instead of 3 -- the 64-antenna case was occupancy-limited at 73-78 % of the HBM peak)
template <int NPS, int T = 64>
__host__ __device__ constexpr int ls_tma_row_tones() { return NPS == 1 ? T : T + 8; }
template <int NLTF, int STAGES, int NPS, int T = 64>
constexpr int ls_tma_smem_bytes() { return STAGES * NLTF * ls_tma_row_tones<NPS, T>() * 8 + STAGES * 8 + 128; }

template <int S, int NLTF, int STAGES, int NPS, int T = 64>
__global__ void __launch_bounds__(T * (NLTF / 16)) ls_tma_kernel(const __grid_constant__ CUtensorMap tmap_y, const LsArgs a) {
  static_assert(T == 64 || (T == 32 && NPS == 1), "the comb-pilot emit path is written for 64-tone tiles");
  constexpr int BLK = 16, NB = NLTF / BLK, TW = ls_tma_row_tones<NPS, T>(), PT = T / NPS;
  extern __shared__ uint8_t sm_ls_raw[];
  float2* in = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(sm_ls_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(in + STAGES * NLTF * TW);
  __shared__ uint32_t cta_abort;
  LsArgs a2 = a;
  a2.pil_per_tile = T;
  bool skip;
  a2.scale = ls_resolve_scale<S>(a, skip, a2.quiet_ovf);
  if (skip) return;                                      // verify pass, provisional scale was inside the window
  const int n_tiles = (a.n_sc + T - 1) / T;
  const long long total = static_cast<long long>(a.n_pkt) * a.n_rx * n_tiles;
  auto issue = [&](long long tile, int stage) {
    const int tl = static_cast<int>(tile % n_tiles);
    const long long prx = tile / n_tiles;
    mbar_arrive_expect_tx(&full[stage], NLTF * TW * 8);
    tma_load_2d(in + stage * NLTF * TW, &tmap_y, &full[stage], tl * T * 2, static_cast<int32_t>(prx * NLTF));
  };
  if (threadIdx.x == 0) {
    cta_abort = 0;
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_y);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) {
      const long long tile = blockIdx.x + static_cast<long long>(s) * gridDim.x;
      if (tile < total) issue(tile, s);
    }
  const int t = threadIdx.x & (T - 1);
  const int b = threadIdx.x / T;
  float amx[2] = {0.f, 0.f};
  long long it = 0;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int stage = static_cast<int>(it % STAGES);
    const uint32_t parity = static_cast<uint32_t>((it / STAGES) & 1);
    const int tl = static_cast<int>(tile % n_tiles);
    const int prx = static_cast<int>(tile / n_tiles);
    const int pil0 = tl * PT;                            // first pilot of the tile
    // pilots this thread column handles: NPS == 1 every tone of the tile; else PT pilots + the halo pilot
    const bool active = NPS == 1 ? (t < min(T, a.n_pil - pil0)) : (t <= PT && pil0 + t < a.n_pil);
    const int col = t * NPS;
    if (!mbar_wait(&full[stage], parity, &cta_abort, a.flags)) return;
    float2* sh = in + stage * NLTF * TW;                 // [NLTF][TW], transformed in place
    if (NPS == 1 || active) {
      float2 v[BLK];
#pragma unroll
      for (int n = 0; n < BLK; ++n) v[n] = sh[(b * BLK + n) * TW + col];
      fwht<BLK>(v);
#pragma unroll
      for (int j = 0; j < BLK; ++j) sh[(b * BLK + j) * TW + col] = v[j];
    }
    __syncthreads();
    if (active) {
      const float2 inv = __ldg(a.inv_den + pil0 + t);
#pragma unroll
      for (int jj = 0; jj < BLK / NB; ++jj) {
        const int j = b + jj * NB;
        float2 u[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) u[c] = sh[(c * BLK + j) * TW + col];
        fwht<NB>(u);
#pragma unroll
        for (int c = 0; c < NB; ++c) sh[(c * BLK + j) * TW + col] = cmul(u[c], inv);
      }
    }
    __syncthreads();
    if constexpr (NPS == 1) {
      ls_emit<S>(a2, sh, TW, prx, pil0, pil0, amx);
    } else {
      const int k0 = tl * T;
      const int nk = min(T, a.n_sc - k0);
      const float inv_nps = 1.0f / static_cast<float>(NPS);
      const size_t row0 = static_cast<size_t>(prx) * a.n_tx;
      bool ovf = false;
      for (int idx = threadIdx.x; idx < a.n_tx * (T / 4); idx += blockDim.x) {
        const int j = idx >> 4, kk = (idx & 15) << 2;
        if (kk >= nk) continue;
        float re[4], im[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k0 + kk + i;
          const int seg = min(k / NPS, a.n_pil - 2);
          const float w = static_cast<float>(k - seg * NPS) * inv_nps;
          const float2 h0 = sh[j * TW + (seg * NPS - k0)];
          const float2 h1 = sh[j * TW + (seg * NPS - k0) + NPS];
          re[i] = h0.x + w * (h1.x - h0.x);
          im[i] = h0.y + w * (h1.y - h0.y);
        }
        ls_store4<S>(a2, row0 + j, k0 + kk, re, im, ovf, amx);
      }
      if (ovf && !a2.quiet_ovf) atomicOr(a.flags, kFlagRange);
    }
    __syncthreads();                                     // stage consumed: refill it for the tile STAGES ahead
    if (threadIdx.x == 0) {
      const long long next = tile + static_cast<long long>(STAGES) * gridDim.x;
      if (next < total) {
        fence_proxy_async_smem();                        // generic-proxy writes above vs the async-proxy refill
        issue(next, stage);
      }
    }
  }
  if (a.planes[0]) { publish_amax<S>(a.dyn, 0, 0, amx[0]); publish_amax<S>(a.dyn, 1, 0, amx[1]); }
}

// ---- FP64 LS for the MATLAB-facing surface (complex128 in AND out, no operand planes) ----------------------------
// helperMIMOChannelEstimate.m computes hD in double; a caller that hands over complex double and asks for complex double
// back gets double arithmetic end to end (the FP32 kernels above would return FP32-grade values in a double container).
// One thread per output element (pkt, rx, tx, k): despread of its pilot tone(s) against row tx of P in FP64 -- O(n_ltf)
// per output, the n_tx threads of a tone re-read the same Y values from L1/L2.  Not the hot path (that one ends in FP32
// operand planes anyway); HBM traffic is the algorithmic 2 x 16 B per element.
struct LsF64Args {
  const double2* Y;        // [n_pkt][n_rx][n_ltf][n_sc]
  const double2* P;        // [n_tx][n_ltf]
  const double2* inv_den;  // [n_pil]
  double2* H;              // [n_pkt][n_rx][n_tx][n_sc]
  int n_rx, n_tx, n_ltf, n_sc, n_ps, n_pil;
};

__device__ __forceinline__ double2 ls_f64_despread(const LsF64Args& a, const double2* y_slab, int j, int pil) {
  const int k = pil * a.n_ps;
  double ax = 0.0, ay = 0.0;
  for (int n = 0; n < a.n_ltf; ++n) {
    const double2 y = __ldg(y_slab + static_cast<size_t>(n) * a.n_sc + k);
    const double2 p = __ldg(a.P + j * a.n_ltf + n);                         // y * conj(p)
    ax = fma(y.x, p.x, fma(y.y, p.y, ax));
    ay = fma(y.y, p.x, fma(-y.x, p.y, ay));
  }
  const double2 d = __ldg(a.inv_den + pil);
  return make_double2(ax * d.x - ay * d.y, ax * d.y + ay * d.x);
}

__global__ void __launch_bounds__(128) ls_f64_kernel(const LsF64Args a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_sc) return;
  const size_t prx = blockIdx.y / a.n_tx;                                   // pkt * n_rx + rx
  const int j = blockIdx.y % a.n_tx;
  const double2* y_slab = a.Y + prx * a.n_ltf * static_cast<size_t>(a.n_sc);
  double2 h;
  if (a.n_ps == 1) {
    h = ls_f64_despread(a, y_slab, j, k);
  } else if (a.n_pil == 1) {
    h = ls_f64_despread(a, y_slab, j, 0);
  } else {
    const int seg = min(k / a.n_ps, a.n_pil - 2);
    const double w = static_cast<double>(k - seg * a.n_ps) / a.n_ps;
    const double2 h0 = ls_f64_despread(a, y_slab, j, seg), h1 = ls_f64_despread(a, y_slab, j, seg + 1);
    h = make_double2(h0.x + w * (h1.x - h0.x), h0.y + w * (h1.y - h0.y));
  }
  a.H[(prx * a.n_tx + j) * static_cast<size_t>(a.n_sc) + k] = h;
}

// ---- mode B: caller planes float32 [rows][d_in] -> operand planes (inference.py:29-30) ----
template <int S>
__global__ void stage_planes_kernel(const float* __restrict__ X, void* planes, int64_t rows, int d_in,
                                    int plane_rows, int kpad, float scale, uint32_t* flags, DynState* dyn, int net,
                                    int fixed_scale) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  bool ovf = false;
  float amx = 0.f;
  if constexpr (S == kFp16x3) {
    if (dyn) {      // level 0 of `net`: exact amax of this plane from the pre-pass (schemes.cuh)
      if (!fixed_scale) scale = pow2_scale_for(__uint_as_float(dyn->in_amax[net]));
      if (blockIdx.x == 0 && threadIdx.x == 0) dyn->scale[net][0] = scale;
    }
  }
  const int64_t total = rows * d_in;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / d_in;
    const int c = static_cast<int>(i - r * d_in);
    E p[Sch::kPlanes];
    const float x = __ldg(X + i);
    amx = fmaxf(amx, fabsf(x));
    Sch::split(x, scale, p, &ovf);
#pragma unroll
    for (int pl = 0; pl < Sch::kPlanes; ++pl)
      reinterpret_cast<E*>(planes)[(static_cast<size_t>(pl) * plane_rows + r) * kpad + c] = p[pl];
  }
  if (ovf) atomicOr(flags, kFlagRange);
  publish_amax<S>(dyn, net, 0, amx);
}

// ---- mode A: [time-domain LTF || P(:,iTx)] per pair (massiveMIMO_dataGenerator.py:303-316) ----
// sig float32 [n_pkt][n_rx][len_ltf]; row = (pkt*n_rx + rx)*n_tx + j gets sig[pkt][rx][:] then Re P[j][0..n_tx)
template <int S>
__global__ void stage_time_p_kernel(const float* __restrict__ sig, const float2* __restrict__ P, void* planes,
                                    int64_t n_prx, int n_tx, int n_ltf, int len_ltf, int plane_rows, int kpad,
                                    float scale, uint32_t* flags, DynState* dyn, int net, int fixed_scale) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  bool ovf = false;
  float amx = 0.f;
  if constexpr (S == kFp16x3) {
    if (dyn) {      // in_amax[net] was seeded with max |Re P| before the pre-pass over the signal plane
      if (!fixed_scale) scale = pow2_scale_for(__uint_as_float(dyn->in_amax[net]));
      if (blockIdx.x == 0 && threadIdx.x == 0) dyn->scale[net][0] = scale;
    }
  }
  const int d_in = len_ltf + n_tx;
  const int64_t total = n_prx * n_tx * d_in;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / d_in;
    const int c = static_cast<int>(i - r * d_in);
    const int64_t prx = r / n_tx;
    const int j = static_cast<int>(r - prx * n_tx);
    const float x = (c < len_ltf) ? __ldg(sig + prx * len_ltf + c) : P[j * n_ltf + (c - len_ltf)].x;
    amx = fmaxf(amx, fabsf(x));
    E p[Sch::kPlanes];
    Sch::split(x, scale, p, &ovf);
#pragma unroll
    for (int pl = 0; pl < Sch::kPlanes; ++pl)
      reinterpret_cast<E*>(planes)[(static_cast<size_t>(pl) * plane_rows + r) * kpad + c] = p[pl];
  }
  if (ovf) atomicOr(flags, kFlagRange);
  publish_amax<S>(dyn, net, 0, amx);
}

// ---- mode A, de-duplicated first layer --------------------------------------------------------------------
// The LTF part of the first-layer input is identical for the Nt pairs of one (pkt, rx) -- the reference itself
// stores it once per rx under a hash (create_massiveMIMO_CSIest_dnn_dataset.py:50-59) -- so
//   W1^T [x_ltf || p_j] + b1 = (W1_ltf^T x_ltf) + (W1_p^T p_j + b1) = Z[prx] + T[j].
// Z comes from one GEMM over n_pkt*n_rx rows (32x fewer first-layer FLOPs); this kernel expands it to the
// n_pkt*n_rx*n_tx pair rows: relu(Z[prx] + T[j]) -> split operand planes of the second layer.
template <int S>
__global__ void expand_pairs_kernel(const float* __restrict__ Z, const float* __restrict__ T, void* planes,
                                    int64_t n_prx, int n_tx, int h, int plane_rows, int kpad, float scale,
                                    uint32_t* flags, DynState* dyn, int net, int fixed_scale, float rowsum,
                                    float tmax) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  bool ovf = false;
  float amx = 0.f;
  if constexpr (S == kFp16x3) {
    if (dyn) {      // level 1: |relu(Z + T)| <= rowsum(W1_ltf) * amax(level 0) + max |T|
      if (!fixed_scale) scale = pow2_scale_for(fmaf(rowsum, __uint_as_float(dyn->amax[net][0]), tmax));
      if (blockIdx.x == 0 && threadIdx.x == 0) dyn->scale[net][1] = scale;
    }
  }
  const int hq = kpad >> 2;                                // h % 4 == 0 (checked on the host); pad columns get zeros
  const int64_t total = n_prx * n_tx * hq;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / hq;
    const int n = static_cast<int>(i - r * hq) << 2;
    const int64_t prx = r / n_tx;
    const int j = static_cast<int>(r - prx * n_tx);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f), t = z;
    if (n < h) {
      z = __ldg(reinterpret_cast<const float4*>(Z + prx * h + n));
      t = __ldg(reinterpret_cast<const float4*>(T + static_cast<size_t>(j) * h + n));
    }
    const float v[4] = {fmaxf(z.x + t.x, 0.f), fmaxf(z.y + t.y, 0.f), fmaxf(z.z + t.z, 0.f), fmaxf(z.w + t.w, 0.f)};
    amx = fmaxf(amx, fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])));
    E p[4][Sch::kPlanes];
#pragma unroll
    for (int q = 0; q < 4; ++q) Sch::split(v[q], scale, p[q], &ovf);
#pragma unroll
    for (int pl = 0; pl < Sch::kPlanes; ++pl) {
      E* d = reinterpret_cast<E*>(planes) + (static_cast<size_t>(pl) * plane_rows + r) * kpad + n;
      if constexpr (sizeof(E) == 4) {
        *reinterpret_cast<float4*>(d) = make_float4(p[0][pl], p[1][pl], p[2][pl], p[3][pl]);
      } else {
        auto bits = [](E x) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&x)); };
        *reinterpret_cast<uint2*>(d) = make_uint2(bits(p[0][pl]) | (bits(p[1][pl]) << 16),
                                                  bits(p[2][pl]) | (bits(p[3][pl]) << 16));
      }
    }
  }
  if (ovf) atomicOr(flags, kFlagRange);
  publish_amax<S>(dyn, net, 1, amx);
}

// ---- amax pre-pass of the FP16X3 range management: max |component| of a float / double array ----------------
// SAMPLE: read the first 128-byte line of every 1 KB (1 line in 8) -- only valid where a later stage verifies the
// result against an exact amax (the LS path, ls_resolve_scale); false = every element.
template <typename T, bool SAMPLE = false>
__global__ void __launch_bounds__(256) amax_kernel(const T* __restrict__ p, size_t n, uint32_t* out_bits) {
  constexpr int V = 16 / sizeof(T);                      // elements per 16-byte load
  float m = 0.f;
  const size_t nv = SAMPLE ? (n / V / 64) * 8 : n / V;   // sampled: 8 vectors out of every 64
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < nv; j += stride) {
    const size_t i = SAMPLE ? ((j >> 3) << 6) + (j & 7) : j;
    if constexpr (sizeof(T) == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    } else {
      const double2 v = __ldg(reinterpret_cast<const double2*>(p) + i);
      m = fmaxf(m, fmaxf(fabsf(static_cast<float>(v.x)), fabsf(static_cast<float>(v.y))));
    }
  }
  if (!SAMPLE && blockIdx.x == 0 && threadIdx.x < n - nv * V) m = fmaxf(m, fabsf(static_cast<float>(p[nv * V + threadIdx.x])));
  const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0 && w) atomicMax(out_bits, w);
}

}  // namespace mm
