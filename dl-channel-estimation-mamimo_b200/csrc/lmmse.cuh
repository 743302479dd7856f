// LMMSE smoother: the 1.1 s/packet comparator of the reference.
//
// Replaces, for a whole batch of packets,
//   packet_generation/phased_arr/helperMIMOChannelEstimate.m:37-39
//     hDmmse(:,j,i) = LMMSE_ce(hD(:,j,i), Nsc, Nsc, Nps, tau, SNR(i))
//   packet_generation/phased_arr/LMMSE_ce.m:23-39
//     Rhp = 1./(1 + j2pi_tau_df*(K1 - K2*Nps));  Rpp = 1./(1 + j2pi_tau_df*Nps*(K3 - K4)) + eye/snr
//     H_MMSE = Rhp*inv(Rpp)*H_tilde
//
// The reference rebuilds and inverts the Nsc x Nsc matrix for every (tx, rx) pair; Rpp only depends on the packet
// (tau_rms) and the rx antenna (SNR(i)), so one "slab" = (packet, rx) is ONE Hermitian positive-definite system
// with the Nt LS vectors as right-hand sides.  Everything is FP64 (cond(Rpp) ~ Nsc * snr rules FP32 out):
//
//   M = [ Rpp ; B^H ]   (n_pad + nt_pad) x n_pad, row-major double2, one per slab       lmmse_fill_kernel
//   blocked LEFT-looking Cholesky over 32-column blocks J (each output written once, accumulators in registers):
//     D   = Rpp_JJ - sum_{K<J} L_JK L_JK^H ;  L_JJ = chol(D) ;  Linv_JJ = L_JJ^-1       lmmse_diag_kernel
//     X_J = (M_J - sum_{K<J} L_K L_JK^H) Linv_JJ^H   for all rows below, B^H rows too   lmmse_panel_kernel
//   the appended rows come out as (L^-1 B)^H: the forward substitution is part of the factorisation.
//   Back substitution per right-hand side, Z^H = Y^H L^-1, and the output                lmmse_backsub_kernel
//     Nps == 1:  Rhp == Rpp - I/snr  =>  H_mmse = B - Z/snr        (every reference call site)
//     Nps  > 1:  H_mmse = Rhp Z                                   lmmse_rhp_kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

constexpr int kLmNB = 32;            // block size of the factorisation
constexpr int kLmPanelRows = 64;     // rows per CTA of the panel kernel
constexpr int kLmPitch = kLmNB + 1;  // shared-memory row pitch (double2): conflict-free LDS.128 for row stride 1
constexpr int kLmPanelSmem = (kLmPanelRows + kLmNB) * kLmPitch * 16;
constexpr uint32_t kFlagNotPd = 4u;  // Rpp not positive definite in FP64 (d_flags bit)

struct LmArgs {
  double2* M;              // [n_slab][R][n_pad]
  double2* Dinv;           // [n_slab][nb][32][32]   inverse of every diagonal Cholesky block
  const double2* par;      // [n_slab]  (c = 2 pi tau_rms / Nfft,  s = 1 / snr_linear)
  const void* B;           // H_ls  complex [n_slab][n_tx][n_sc]   (float2 or double2)
  void* out;               // H_mmse, same layout (float2 or double2)
  int b_double, out_double;
  int n, n_pad, nb, n_tx, nt_pad, R, n_ps;
  int J;                   // current block column
  uint32_t* flags;
};

__device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// acc += a * conj(b)
__device__ __forceinline__ void zmac_conj(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
}
// acc += a * b
__device__ __forceinline__ void zmac(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// 1 / (1 + j x)
__device__ __forceinline__ double2 corr(double x) {
  const double d = 1.0 / fma(x, x, 1.0);
  return make_double2(d, -x * d);
}

// ---- M = [Rpp (lower triangle) ; conj(B)^T rows] -----------------------------------------------------------
__global__ void __launch_bounds__(256) lmmse_fill_kernel(const LmArgs a) {
  const int slab = blockIdx.y;
  const double2 cs = a.par[slab];
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const size_t total = static_cast<size_t>(a.R) * a.n_pad;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx / a.n_pad), j = static_cast<int>(idx - static_cast<size_t>(i) * a.n_pad);
    double2 v = make_double2(0.0, 0.0);
    if (i < a.n_pad) {
      if (j > i) continue;                                   // upper triangle is never read
      if (i < a.n && j < a.n) {
        v = corr(cs.x * a.n_ps * static_cast<double>(i - j));  // rf2(i, j)   LMMSE_ce.m:35-36
        if (i == j) v.x += cs.y;                             // + eye / snr   LMMSE_ce.m:38
      } else if (i == j) {
        v.x = 1.0;                                           // padding: identity block, decoupled
      }
    } else {
      const int t = i - a.n_pad;
      if (t < a.n_tx && j < a.n) {
        const size_t g = (static_cast<size_t>(slab) * a.n_tx + t) * a.n + j;
        if (a.b_double) {
          const double2 b = reinterpret_cast<const double2*>(a.B)[g];
          v = make_double2(b.x, -b.y);
        } else {
          const float2 b = reinterpret_cast<const float2*>(a.B)[g];
          v = make_double2(b.x, -static_cast<double>(b.y));
        }
      }
    }
    M[idx] = v;
  }
}

// ---- diagonal block J: Schur update, Cholesky, inverse ----------------------------------------------------------
__global__ void __launch_bounds__(256) lmmse_diag_kernel(const LmArgs a) {
  __shared__ double2 Lt[kLmNB][kLmPitch];       // k-tile of L_J* during the Schur update, then the factor Lo
  __shared__ double2 D[kLmNB][kLmPitch];
  double2 (*Lo)[kLmPitch] = Lt;
  const int slab = blockIdx.x;
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const int Jb = a.J * kLmNB;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;     // outputs (ty + 16 i, tx + 16 j), i, j < 2
  double2 acc[2][2] = {};
  for (int k0 = 0; k0 < Jb; k0 += kLmNB) {
    __syncthreads();
    for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) {
      const int r = e >> 5, kk = e & 31;
      Lt[r][kk] = M[static_cast<size_t>(Jb + r) * a.n_pad + k0 + kk];
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kLmNB; ++kk) {
      const double2 a0 = Lt[ty][kk], a1 = Lt[ty + 16][kk], b0 = Lt[tx][kk], b1 = Lt[tx + 16][kk];
      zmac_conj(acc[0][0], a0, b0);
      zmac_conj(acc[0][1], a0, b1);
      zmac_conj(acc[1][0], a1, b0);
      zmac_conj(acc[1][1], a1, b1);
    }
  }
  __syncthreads();                                            // Lt is free: it becomes Lo
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      double2 v = make_double2(0.0, 0.0);
      if (c <= r) v = zsub(M[static_cast<size_t>(Jb + r) * a.n_pad + Jb + c], acc[i][j]);
      D[r][c] = v;
      Lo[r][c] = make_double2(0.0, 0.0);
    }
  // unblocked right-looking Cholesky of the 32 x 32 block, one barrier per column: updates read the UNscaled
  // column k of D, the scaled column goes to Lo
  bool bad = false;
  for (int k = 0; k < kLmNB; ++k) {
    __syncthreads();
    const double dkk = D[k][k].x;
    if (!(dkk > 0.0)) bad = true;
    const double inv = 1.0 / dkk;
    for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) {
      const int r = e >> 5, c = e & 31;
      if (c > k && r >= c) {
        const double2 lr = D[r][k], lc = D[c][k];
        double2 v = D[r][c];
        v.x -= (lr.x * lc.x + lr.y * lc.y) * inv;
        v.y -= (lr.y * lc.x - lr.x * lc.y) * inv;
        D[r][c] = v;
      }
    }
    if (threadIdx.x < kLmNB && threadIdx.x >= k) {
      const double sq = sqrt(dkk);
      const double2 v = D[threadIdx.x][k];
      Lo[threadIdx.x][k] = (static_cast<int>(threadIdx.x) == k) ? make_double2(sq, 0.0) : make_double2(v.x / sq, v.y / sq);
    }
  }
  __syncthreads();
  if (bad && threadIdx.x == 0) atomicOr(a.flags, kFlagNotPd);
  for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) {
    const int r = e >> 5, c = e & 31;
    if (c <= r) M[static_cast<size_t>(Jb + r) * a.n_pad + Jb + c] = Lo[r][c];
  }
  // Linv = Lo^-1 (lower triangular): thread c solves Lo x = e_c by forward substitution, x kept in column c of D
  if (threadIdx.x < kLmNB) {
    const int c = threadIdx.x;
    for (int r = 0; r < kLmNB; ++r) {
      double2 s = make_double2(r == c ? 1.0 : 0.0, 0.0);
      for (int m = c; m < r; ++m) {                           // x_m = 0 for m < c
        const double2 l = Lo[r][m], x = D[m][c];
        s.x -= l.x * x.x - l.y * x.y;
        s.y -= l.x * x.y + l.y * x.x;
      }
      const double inv = 1.0 / Lo[r][r].x;
      D[r][c] = (r < c) ? make_double2(0.0, 0.0) : make_double2(s.x * inv, s.y * inv);
    }
  }
  __syncthreads();
  double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + a.J) * kLmNB * kLmNB;
  for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) Di[e] = D[e >> 5][e & 31];
}

// ---- panel J: rows below the diagonal block (the B^H rows included) ---------------------------------------------
__global__ void __launch_bounds__(256) lmmse_panel_kernel(const LmArgs a) {
  extern __shared__ double2 lm_smem[];                        // 50.7 KB: over the 48 KB static limit
  double2 (*As)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem);
  double2 (*Bs)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + kLmPanelRows * kLmPitch);
  const int slab = blockIdx.y;
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const int Jb = a.J * kLmNB;
  const int row0 = Jb + kLmNB + blockIdx.x * kLmPanelRows;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;     // outputs (ty + 16 i, tx + 16 j), i < 4, j < 2
  const int lr = threadIdx.x >> 5, lk = threadIdx.x & 31;     // loader: rows lr + 8 q, column lk
  double2 acc[4][2] = {};
  double2 pa[8], pb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = row0 + lr + 8 * q;
      pa[q] = (r < a.R) ? M[static_cast<size_t>(r) * a.n_pad + k0 + lk] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) pb[q] = M[static_cast<size_t>(Jb + lr + 8 * q) * a.n_pad + k0 + lk];
  };
  if (Jb > 0) fetch(0);
  for (int k0 = 0; k0 < Jb; k0 += kLmNB) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) As[lr + 8 * q][lk] = pa[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) Bs[lr + 8 * q][lk] = pb[q];
    __syncthreads();
    if (k0 + kLmNB < Jb) fetch(k0 + kLmNB);                   // next tile in flight while this one is consumed
#pragma unroll 4
    for (int kk = 0; kk < kLmNB; ++kk) {
      double2 av[4], bv[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[ty + 16 * i][kk];
#pragma unroll
      for (int j = 0; j < 2; ++j) bv[j] = Bs[tx + 16 * j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) zmac_conj(acc[i][j], av[i], bv[j]);
    }
  }
  __syncthreads();
  // P = M_J - acc  -> As ;  Linv_JJ -> Bs ;  X = P Linv^H
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = row0 + ty + 16 * i;
      double2 v = make_double2(0.0, 0.0);
      if (r < a.R) v = zsub(M[static_cast<size_t>(r) * a.n_pad + Jb + tx + 16 * j], acc[i][j]);
      As[ty + 16 * i][tx + 16 * j] = v;
    }
  const double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + a.J) * kLmNB * kLmNB;
  for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) Bs[e >> 5][e & 31] = Di[e];
  __syncthreads();
  double2 x[4][2] = {};
#pragma unroll 4
  for (int m = 0; m < kLmNB; ++m) {
    double2 av[4], bv[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = As[ty + 16 * i][m];
#pragma unroll
    for (int j = 0; j < 2; ++j) bv[j] = Bs[tx + 16 * j][m];   // Linv[c][m], zero for m > c
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) zmac_conj(x[i][j], av[i], bv[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = row0 + ty + 16 * i;
      if (r < a.R) M[static_cast<size_t>(r) * a.n_pad + Jb + tx + 16 * j] = x[i][j];
    }
}

// ---- back substitution, one warp per right-hand side: z L = y, from the last block to the first ---------------------
__global__ void __launch_bounds__(256) lmmse_backsub_kernel(const LmArgs a) {
  const int slab = blockIdx.y;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= a.nt_pad) return;
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  double2* z = M + static_cast<size_t>(a.n_pad + t) * a.n_pad;       // y on entry, z on exit (in place)
  const double s = a.par[slab].y;
  for (int J = a.nb - 1; J >= 0; --J) {
    const int Jb = J * kLmNB;
    double2 v0 = z[Jb + lane], v1 = make_double2(0.0, 0.0), v2 = v1, v3 = v1;
    int i = Jb + kLmNB;
    for (; i + 3 < a.n_pad; i += 4) {                                // v -= z[i] * L[i][Jb + lane]
      const double2 z0 = __ldcg(z + i), z1 = __ldcg(z + i + 1), z2 = __ldcg(z + i + 2), z3 = __ldcg(z + i + 3);
      const double2* Lr = M + static_cast<size_t>(i) * a.n_pad + Jb + lane;
      const double2 l0 = Lr[0], l1 = Lr[a.n_pad], l2 = Lr[2 * static_cast<size_t>(a.n_pad)], l3 = Lr[3 * static_cast<size_t>(a.n_pad)];
      zmac(v0, make_double2(-z0.x, -z0.y), l0);
      zmac(v1, make_double2(-z1.x, -z1.y), l1);
      zmac(v2, make_double2(-z2.x, -z2.y), l2);
      zmac(v3, make_double2(-z3.x, -z3.y), l3);
    }
    const double2 v = zadd(zadd(v0, v1), zadd(v2, v3));              // n_pad is a multiple of 32: no remainder loop
    // z_J = v_J Linv_JJ :  z[c] = sum_{m >= c} v[m] Linv[m][c]
    const double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + J) * kLmNB * kLmNB;
    double2 zc = make_double2(0.0, 0.0);
#pragma unroll 8
    for (int m = 0; m < kLmNB; ++m) {
      const double2 vm = make_double2(__shfl_sync(0xffffffffu, v.x, m), __shfl_sync(0xffffffffu, v.y, m));
      zmac(zc, vm, Di[m * kLmNB + lane]);                            // Linv[m][lane] == 0 for m < lane
    }
    __stcg(z + Jb + lane, zc);
    __syncwarp();
    const int k = Jb + lane;
    if (a.n_ps == 1 && t < a.n_tx && k < a.n) {                      // H_mmse = B - conj(z) / snr
      const size_t g = (static_cast<size_t>(slab) * a.n_tx + t) * a.n + k;
      double2 b;
      if (a.b_double) b = reinterpret_cast<const double2*>(a.B)[g];
      else { const float2 f = reinterpret_cast<const float2*>(a.B)[g]; b = make_double2(f.x, f.y); }
      const double2 h = make_double2(b.x - s * zc.x, b.y + s * zc.y);
      if (a.out_double) reinterpret_cast<double2*>(a.out)[g] = h;
      else reinterpret_cast<float2*>(a.out)[g] = make_float2(static_cast<float>(h.x), static_cast<float>(h.y));
    }
  }
}

// ---- Nps > 1 (never used by the reference's call sites): H_mmse[t][k] = sum_b Rhp[k][b] conj(z_t[b]) -----------------
__global__ void __launch_bounds__(256) lmmse_rhp_kernel(const LmArgs a) {
  const int slab = blockIdx.z;
  const int t = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n) return;
  const double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const double2* z = M + static_cast<size_t>(a.n_pad + t) * a.n_pad;
  const double c = a.par[slab].x;
  double2 h = make_double2(0.0, 0.0);
  for (int b = 0; b < a.n; ++b) {
    const double2 r = corr(c * static_cast<double>(k - b * a.n_ps));   // rf(k, b)   LMMSE_ce.m:33-34
    const double2 zb = z[b];
    zmac(h, r, make_double2(zb.x, -zb.y));
  }
  const size_t g = (static_cast<size_t>(slab) * a.n_tx + t) * a.n + k;
  if (a.out_double) reinterpret_cast<double2*>(a.out)[g] = h;
  else reinterpret_cast<float2*>(a.out)[g] = make_float2(static_cast<float>(h.x), static_cast<float>(h.y));
}

}  // namespace mm
