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
// with the Nt LS vectors as right-hand sides.  Everything is FP64 (cond(Rpp) ~ Nsc * snr rules FP32 out).
//
// Two routes share the workspace and the output formula:
//   * Toeplitz route (default; further down: lmmse_schur / linv / solve kernels): Rpp is Hermitian Toeplitz, its
//     Cholesky factor comes from the O(n^2) generalised Schur recursion, then two blocked triangular solves.
//   * dense route (MAMIMO_LMMSE_SCHUR=0; the cross-check, described next): blocked left-looking Cholesky.
//
//   M = [ Rpp ; B^H ]   (n_pad + nt_pad) x n_pad, row-major double2, one per slab; its entries are generated on
//   the fly (lm_elem), the buffer only ever holds the factor L and the solved right-hand sides
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

// ---- element (i, j) of [Rpp ; B^H], generated on the fly (no fill pass, M only ever holds L) ----------------------
__device__ __forceinline__ double2 lm_elem(const LmArgs& a, int slab, double2 cs, int i, int j) {
  double2 v = make_double2(0.0, 0.0);
  if (i < a.n_pad) {
    if (i < a.n && j < a.n) {
      v = corr(cs.x * a.n_ps * static_cast<double>(i - j));    // rf2(i, j)   LMMSE_ce.m:35-36
      if (i == j) v.x += cs.y;                                 // + eye / snr   LMMSE_ce.m:38
    } else if (i == j) {
      v.x = 1.0;                                               // padding: identity block, decoupled
    }
  } else {
    const int t = i - a.n_pad;                                 // appended rows: conj(B)^T
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
  return v;
}

// ---- diagonal block J: Schur update, Cholesky, inverse ----------------------------------------------------------
__global__ void __launch_bounds__(256) lmmse_diag_kernel(const LmArgs a) {
  __shared__ double2 Lt[kLmNB][kLmPitch];       // k-tile of L_J* during the Schur update, then the factor Lo
  __shared__ double2 D[kLmNB][kLmPitch];
  double2 (*Lo)[kLmPitch] = Lt;
  const int slab = blockIdx.x;
  const double2 cs = a.par[slab];
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
      if (c <= r) v = zsub(lm_elem(a, slab, cs, Jb + r, Jb + c), acc[i][j]);
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
  // Linv = Lo^-1 (lower triangular), column c by forward substitution  x_r = (delta_rc - sum_{c<=m<r} Lo[r][m] x_m) / Lo[r][r].
  // 8 lanes share one column (the sum over m is split 8 ways and reduced with shuffles), 4 columns per warp; x is
  // kept in column c of D, which only this column's lanes touch.
  {
    const int c = threadIdx.x >> 3, hlp = threadIdx.x & 7;
    for (int r = 0; r < kLmNB; ++r) {
      double2 sum = make_double2(0.0, 0.0);
      for (int m = c + hlp; m < r; m += 8) zmac(sum, Lo[r][m], D[m][c]);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
      }
      if (hlp == 0) {
        const double inv = 1.0 / Lo[r][r].x;
        D[r][c] = (r < c) ? make_double2(0.0, 0.0)
                          : make_double2(((r == c ? 1.0 : 0.0) - sum.x) * inv, -sum.y * inv);
      }
      __syncwarp();
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
  const double2 cs = a.par[slab];
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
      if (r < a.R) v = zsub(lm_elem(a, slab, cs, r, Jb + tx + 16 * j), acc[i][j]);
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

// ---- back substitution  Z^H = Y^H L^-1  as a blocked GEMM: one CTA = 32 right-hand sides of one slab ----------------
//   for J = last .. 0:   z_J = ( y_J - sum_{I > J} z_I L[I][J] ) Linv_JJ
// The 32 x 32 tiles of z (rows = right-hand sides) and of L are staged in shared memory and shared by all 32
// right-hand sides (the first version, one warp per right-hand side chasing z through L2, took 40 % of the whole
// smoother).  128 threads, 4 x 2 outputs each, next tiles prefetched into registers.
constexpr int kLmBsSmem = 4 * kLmNB * kLmPitch * 16;
__global__ void __launch_bounds__(128) lmmse_backsub_kernel(const LmArgs a) {
  extern __shared__ double2 lm_smem[];
  double2 (*Zs)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem);
  double2 (*Ls)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + kLmNB * kLmPitch);
  double2 (*Vs)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + 2 * kLmNB * kLmPitch);
  double2 (*Li)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + 3 * kLmNB * kLmPitch);
  const int slab = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  double2* Zg = M + static_cast<size_t>(a.n_pad + t0) * a.n_pad;     // rows t0.. : y on entry, z on exit (in place)
  const int n_rows = min(32, a.nt_pad - t0);
  const double s = a.par[slab].y;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;           // outputs (ty + 8 i, tx + 16 j), i < 4, j < 2
  const int lr = threadIdx.x >> 5, lk = threadIdx.x & 31;           // loader: rows lr + 4 q, column lk
  double2 pz[8], pl[8];
  for (int J = a.nb - 1; J >= 0; --J) {
    const int Jb = J * kLmNB;
    double2 acc[4][2] = {};
    auto fetch = [&](int Ib) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = lr + 4 * q;
        pz[q] = (r < n_rows) ? Zg[static_cast<size_t>(r) * a.n_pad + Ib + lk] : make_double2(0.0, 0.0);
        pl[q] = M[static_cast<size_t>(Ib + r) * a.n_pad + Jb + lk];
      }
    };
    if (J + 1 < a.nb) fetch(Jb + kLmNB);
    for (int I = J + 1; I < a.nb; ++I) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q) { Zs[lr + 4 * q][lk] = pz[q]; Ls[lr + 4 * q][lk] = pl[q]; }
      __syncthreads();
      if (I + 1 < a.nb) fetch((I + 1) * kLmNB);
#pragma unroll 4
      for (int ii = 0; ii < kLmNB; ++ii) {
        double2 zv[4], lv[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) zv[i] = Zs[ty + 8 * i][ii];
#pragma unroll
        for (int j = 0; j < 2; ++j) lv[j] = Ls[ii][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) zmac(acc[i][j], zv[i], lv[j]);
      }
    }
    // v = y_J - acc -> Vs ;  Linv_JJ -> Li ;  z_J = v Linv_JJ
    const double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + J) * kLmNB * kLmNB;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = ty + 8 * i, c = tx + 16 * j;
        Vs[r][c] = (r < n_rows) ? zsub(Zg[static_cast<size_t>(r) * a.n_pad + Jb + c], acc[i][j]) : make_double2(0.0, 0.0);
      }
    for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 128) Li[e >> 5][e & 31] = Di[e];
    __syncthreads();
    double2 z[4][2] = {};
#pragma unroll 4
    for (int m = 0; m < kLmNB; ++m) {
      double2 vv[4], lv[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = Vs[ty + 8 * i][m];
#pragma unroll
      for (int j = 0; j < 2; ++j) lv[j] = Li[m][tx + 16 * j];       // Linv[m][c] == 0 for m < c
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) zmac(z[i][j], vv[i], lv[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = ty + 8 * i, k = Jb + tx + 16 * j, t = t0 + r;
        if (r >= n_rows) continue;
        Zg[static_cast<size_t>(r) * a.n_pad + k] = z[i][j];
        if (a.n_ps == 1 && t < a.n_tx && k < a.n) {                  // H_mmse = B - conj(z) / snr
          const size_t g = (static_cast<size_t>(slab) * a.n_tx + t) * a.n + k;
          double2 b;
          if (a.b_double) b = reinterpret_cast<const double2*>(a.B)[g];
          else { const float2 f = reinterpret_cast<const float2*>(a.B)[g]; b = make_double2(f.x, f.y); }
          const double2 h = make_double2(b.x - s * z[i][j].x, b.y + s * z[i][j].y);
          if (a.out_double) reinterpret_cast<double2*>(a.out)[g] = h;
          else reinterpret_cast<float2*>(a.out)[g] = make_float2(static_cast<float>(h.x), static_cast<float>(h.y));
        }
      }
    __syncthreads();                                                 // z_J visible to the next block's tile loads
  }
}

// =====================================================================================================================
// Toeplitz route (default).  Rpp = T + I/snr is HERMITIAN TOEPLITZ (entry (i, j) depends on i - j only), so its
// Cholesky factor follows from the generalised Schur algorithm in O(n^2) instead of n^3/3: with the generator
// u = t / sqrt(t_0), v = u, v_0 = 0 (t = first column), step k emits column k of L = u, shifts u down by one and applies
// the hyperbolic rotation that zeroes v_{k+1} (rho = v_{k+1} / u_{k+1}, |rho| < 1 iff positive definite), in the
// mixed-downdating form  u' = (u - conj(rho) v) / sqrt(1 - |rho|^2),  v' = sqrt(1 - |rho|^2) v - rho u'
// (backward stable for positive-definite Toeplitz matrices, Bojanczyk/Brent/de Hoog/Sweet 1995; checked against
// numpy: ||L L^H - T|| / ||T|| ~ 1e-15 up to cond 4e13).  What remains is the two triangular solves per right-hand
// side: 8 (n^2/... ) i.e. Nt n^2 complex MACs per slab, ~3.4x fewer FLOPs than the blocked Cholesky at 32 x 234.
//   lmmse_schur_kernel : one CTA per slab, u and v in shared memory (the shift is a pointer offset), writes
//                        R = L^H row by row (coalesced) into M[0 .. n_pad)
//   lmmse_linv_kernel  : inverse of every 32 x 32 diagonal block (one CTA per block)
//   lmmse_solve_kernel : forward then backward substitution as blocked GEMMs, one CTA per 32 right-hand sides

__global__ void __launch_bounds__(128) lmmse_schur_kernel(const LmArgs a) {
  extern __shared__ double2 lm_smem[];
  double2* ub = lm_smem;                 // u_k[i] lives at ub[i - k]: the downshift of u is a change of origin
  double2* vb = lm_smem + a.n_pad;       // v[i] at vb[i]
  const int slab = blockIdx.x;
  const double2 cs = a.par[slab];
  double2* R = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const int n = a.n;
  const double t0 = 1.0 + cs.y;
  const double is0 = 1.0 / sqrt(t0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double2 t = corr(cs.x * a.n_ps * static_cast<double>(i));          // first column of rf2   LMMSE_ce.m:35-36
    if (i == 0) t.x += cs.y;                                             // + eye / snr           LMMSE_ce.m:38
    const double2 u = make_double2(t.x * is0, t.y * is0);
    ub[i] = u;
    vb[i] = i == 0 ? make_double2(0.0, 0.0) : u;
  }
  for (int k = n + threadIdx.x; k < a.n_pad; k += blockDim.x)           // padding rows: identity
    for (int i = k; i < a.n_pad; ++i) R[static_cast<size_t>(k) * a.n_pad + i] = make_double2(i == k ? 1.0 : 0.0, 0.0);
  bool bad = false;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    double2* Rk = R + static_cast<size_t>(k) * a.n_pad + k;
    for (int j = threadIdx.x; j < a.n_pad - k; j += blockDim.x) {        // row k of R = conj(column k of L)
      double2 u = make_double2(0.0, 0.0);
      if (j < n - k) { u = ub[j]; u.y = -u.y; }
      Rk[j] = u;
    }
    if (k == n - 1) break;
    const double2 piv = ub[0], vk = vb[k + 1];
    __syncthreads();                                                     // everyone holds the pivot pair
    const double ip = 1.0 / piv.x;                                       // the pivot u_{k+1} = L[k][k] is real
    const double2 rho = make_double2(vk.x * ip, vk.y * ip);
    const double d = 1.0 - (rho.x * rho.x + rho.y * rho.y);
    if (!(d > 0.0)) bad = true;
    const double sq = sqrt(d), isq = 1.0 / sq;
    for (int j = threadIdx.x; j < n - k - 1; j += blockDim.x) {          // element i = j + k + 1
      const double2 u = ub[j], v = vb[j + k + 1];
      // u' = (u - conj(rho) v) / sq ;  v' = sq v - rho u'
      double2 un = make_double2((u.x - (rho.x * v.x + rho.y * v.y)) * isq, (u.y - (rho.x * v.y - rho.y * v.x)) * isq);
      double2 vn = make_double2(sq * v.x - (rho.x * un.x - rho.y * un.y), sq * v.y - (rho.x * un.y + rho.y * un.x));
      ub[j] = un;
      vb[j + k + 1] = vn;
    }
  }
  if (bad && threadIdx.x == 0) atomicOr(a.flags, kFlagNotPd);
}

// Linv_JJ = (R_JJ^H)^-1 for diagonal block J = blockIdx.x of slab blockIdx.y
__global__ void __launch_bounds__(256) lmmse_linv_kernel(const LmArgs a) {
  __shared__ double2 Lo[kLmNB][kLmPitch];
  __shared__ double2 D[kLmNB][kLmPitch];
  const int J = blockIdx.x, slab = blockIdx.y;
  const double2* R = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  const int Jb = J * kLmNB;
  for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) {
    const int r = e >> 5, c = e & 31;                                   // R[Jb + r][Jb + c], upper triangle
    double2 v = make_double2(0.0, 0.0);
    if (c >= r) { v = R[static_cast<size_t>(Jb + r) * a.n_pad + Jb + c]; v.y = -v.y; }
    Lo[c][r] = v;                                                       // Lo = R_JJ^H (lower)
  }
  __syncthreads();
  {
    const int c = threadIdx.x >> 3, hlp = threadIdx.x & 7;
    for (int r = 0; r < kLmNB; ++r) {
      double2 sum = make_double2(0.0, 0.0);
      for (int m = c + hlp; m < r; m += 8) zmac(sum, Lo[r][m], D[m][c]);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
      }
      if (hlp == 0) {
        const double inv = 1.0 / Lo[r][r].x;
        D[r][c] = (r < c) ? make_double2(0.0, 0.0)
                          : make_double2(((r == c ? 1.0 : 0.0) - sum.x) * inv, -sum.y * inv);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + J) * kLmNB * kLmNB;
  for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 256) Di[e] = D[e >> 5][e & 31];
}

// forward (Y^H = B^H L^-H) then backward (Z^H = Y^H L^-1) substitution against R = L^H, one CTA = 32 right-hand sides.
// One pass over the block columns: kBwd = false ascending (forward), true descending (backward).
template <bool kBwd>
__device__ __forceinline__ void lm_solve_pass(const LmArgs& a, int slab, double2 cs, double2* M, double2* Zg, int t0,
                                              int n_rows, double2 (*Zs)[kLmPitch], double2 (*Ls)[kLmPitch],
                                              double2 (*Vs)[kLmPitch], double2 (*Li)[kLmPitch]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;           // outputs (ty + 8 i, tx + 16 j), i < 4, j < 2
  const int lr = threadIdx.x >> 5, lk = threadIdx.x & 31;           // loader: rows lr + 4 q, column lk
  const double s = cs.y;
  double2 pz[8], pl[8];
  for (int jj = 0; jj < a.nb; ++jj) {
    const int J = kBwd ? a.nb - 1 - jj : jj;
    const int Jb = J * kLmNB;
    const int I0 = kBwd ? J + 1 : 0, I1 = kBwd ? a.nb : J;          // blocks already solved in this pass
    double2 acc[4][2] = {};
    auto fetch = [&](int Ib) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = lr + 4 * q;
        pz[q] = (r < n_rows) ? Zg[static_cast<size_t>(r) * a.n_pad + Ib + lk] : make_double2(0.0, 0.0);
        // forward: R[Ib + r][Jb + lk];  backward: L[Ib + ii][Jb + c] = conj(R[Jb + c][Ib + ii]), loaded as (c = r, ii = lk)
        pl[q] = kBwd ? M[static_cast<size_t>(Jb + r) * a.n_pad + Ib + lk] : M[static_cast<size_t>(Ib + r) * a.n_pad + Jb + lk];
      }
    };
    if (I0 < I1) fetch(I0 * kLmNB);
    for (int I = I0; I < I1; ++I) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        Zs[lr + 4 * q][lk] = pz[q];
        if (kBwd) Ls[lk][lr + 4 * q] = make_double2(pl[q].x, -pl[q].y);
        else Ls[lr + 4 * q][lk] = pl[q];
      }
      __syncthreads();
      if (I + 1 < I1) fetch((I + 1) * kLmNB);
#pragma unroll 4
      for (int ii = 0; ii < kLmNB; ++ii) {
        double2 zv[4], lv[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) zv[i] = Zs[ty + 8 * i][ii];
#pragma unroll
        for (int j = 0; j < 2; ++j) lv[j] = Ls[ii][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) zmac(acc[i][j], zv[i], lv[j]);
      }
    }
    // v = rhs_J - acc -> Vs ;  Linv_JJ -> Li ;  forward: y_J = v Linv^H, backward: z_J = v Linv
    const double2* Di = a.Dinv + (static_cast<size_t>(slab) * a.nb + J) * kLmNB * kLmNB;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = ty + 8 * i, c = tx + 16 * j;
        double2 rhs = make_double2(0.0, 0.0);
        if (r < n_rows) rhs = kBwd ? Zg[static_cast<size_t>(r) * a.n_pad + Jb + c] : lm_elem(a, slab, cs, a.n_pad + t0 + r, Jb + c);
        Vs[r][c] = zsub(rhs, acc[i][j]);
      }
    for (int e = threadIdx.x; e < kLmNB * kLmNB; e += 128) {
      const double2 v = Di[e];
      if (kBwd) Li[e >> 5][e & 31] = v;                              // Li[m][c] = Linv[m][c]
      else Li[e & 31][e >> 5] = make_double2(v.x, -v.y);             // Li[m][c] = conj(Linv[c][m])
    }
    __syncthreads();
    double2 z[4][2] = {};
#pragma unroll 4
    for (int m = 0; m < kLmNB; ++m) {
      double2 vv[4], lv[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = Vs[ty + 8 * i][m];
#pragma unroll
      for (int j = 0; j < 2; ++j) lv[j] = Li[m][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) zmac(z[i][j], vv[i], lv[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = ty + 8 * i, k = Jb + tx + 16 * j, t = t0 + r;
        if (r >= n_rows) continue;
        Zg[static_cast<size_t>(r) * a.n_pad + k] = z[i][j];
        if (kBwd && a.n_ps == 1 && t < a.n_tx && k < a.n) {          // H_mmse = B - conj(z) / snr
          const size_t g = (static_cast<size_t>(slab) * a.n_tx + t) * a.n + k;
          double2 b;
          if (a.b_double) b = reinterpret_cast<const double2*>(a.B)[g];
          else { const float2 f = reinterpret_cast<const float2*>(a.B)[g]; b = make_double2(f.x, f.y); }
          const double2 h = make_double2(b.x - s * z[i][j].x, b.y + s * z[i][j].y);
          if (a.out_double) reinterpret_cast<double2*>(a.out)[g] = h;
          else reinterpret_cast<float2*>(a.out)[g] = make_float2(static_cast<float>(h.x), static_cast<float>(h.y));
        }
      }
    __syncthreads();                                                 // this block's solution visible to the next block
  }
}

__global__ void __launch_bounds__(128, 3) lmmse_solve_kernel(const LmArgs a) {
  extern __shared__ double2 lm_smem[];
  double2 (*Zs)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem);
  double2 (*Ls)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + kLmNB * kLmPitch);
  double2 (*Vs)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + 2 * kLmNB * kLmPitch);
  double2 (*Li)[kLmPitch] = reinterpret_cast<double2 (*)[kLmPitch]>(lm_smem + 3 * kLmNB * kLmPitch);
  const int slab = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const double2 cs = a.par[slab];
  double2* M = a.M + static_cast<size_t>(slab) * a.R * a.n_pad;
  double2* Zg = M + static_cast<size_t>(a.n_pad + t0) * a.n_pad;     // this CTA's right-hand sides (Y, then Z)
  const int n_rows = min(32, a.nt_pad - t0);
  lm_solve_pass<false>(a, slab, cs, M, Zg, t0, n_rows, Zs, Ls, Vs, Li);
  lm_solve_pass<true>(a, slab, cs, M, Zg, t0, n_rows, Zs, Ls, Vs, Li);
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
