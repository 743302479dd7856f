// Per-subcarrier SVD of the estimated channel: the first consumer of H-hat.
//
// Replaces, for a whole batch, the decomposition at the top of
//   packet_generation/phased_arr/omphybweights.m:169-176  (getWeightsForSubcarrier, called per subcarrier at :160-163
//   from pg/BER_test_maMIMO_LTF.m:372):
//     H = Hin.';  [~,~,v] = svd(H);  Fopt = v(:,1:Ns);
// Hin = squeeze(hDp(k,:,:)) is [Nt x Nr], so H is [Nr x Nt] with H(i,j) = hD(k,j,i) = H-hat[pkt][i_rx][j_tx][k].
//
// What is computed are the quantities that do NOT depend on LAPACK's choice of basis (v is Nt x Nt, of which Nt - Nr
// columns span the null space in an arbitrary basis; the call site keeps Ns = numSTS of them, 1 as shipped):
//   sigma[r]        the Nr singular values, descending
//   V1[:, r]        the Nr dominant right singular vectors v_r = H^H u_r / sigma_r; each is unique up to a phase, the
//                   projector V1 V1^H onto the row space of H and Fopt Fopt^H restricted to it are unique.
// HBM-bound by design: one thread per (packet, tone); lanes run along the tone axis, so every load / store of the
// [pkt][rx][tx][k] tensor is coalesced.  Pass 1 accumulates the Nr x Nr Gram matrix G = H H^H in FP64 while
// streaming the Nt columns; a cyclic complex Jacobi iteration diagonalises G in registers (Nr <= 4) or in shared memory
// (Nr = 5..8, svd_gram_smem_kernel); pass 2 re-reads H (L1/L2 hit: the same lines, microseconds later) and emits V1.
// Forming G squares the condition number: sigma_r is good to ~1e-16 * (sigma_1/sigma_r)^2 relative in FP64 -- ample for
// the FP32-grade H-hat this engine produces; vectors of singular values below 1e-7 sigma_1 are returned as zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

struct SvdArgs {
  const void* H;       // complex [n_pkt][n_rx][n_tx][n_sc]  (float2 or double2)
  void* sigma;         // real    [n_pkt][n_rx][n_sc]        (float or double), r-th singular value of tone k at [r][k]
  void* V1;            // complex [n_pkt][n_rx][n_tx][n_sc]  (float2 or double2), V1[:, r] of tone k at [r][:][k]; may be NULL
  int h_double, out_double;
  int n_tx, n_sc;
};

__device__ __forceinline__ double2 svd_ld(const void* H, size_t idx, int is_double) {
  if (is_double) return __ldg(reinterpret_cast<const double2*>(H) + idx);
  const float2 v = __ldg(reinterpret_cast<const float2*>(H) + idx);
  return make_double2(v.x, v.y);
}

// HD: H is complex128 (compile-time, so the column loads are straight-line code the compiler can batch)
template <int NR, bool HD>
__global__ void __launch_bounds__(128) svd_gram_kernel(const SvdArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_sc) return;
  const size_t pkt = blockIdx.y;
  const size_t base = pkt * NR * a.n_tx * static_cast<size_t>(a.n_sc) + k;       // H[pkt][0][0][k]
  const size_t rx_stride = static_cast<size_t>(a.n_tx) * a.n_sc;

  // ---- pass 1: G = H H^H (Hermitian, full storage), FP64
  double2 G[NR][NR];
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < NR; ++j) G[i][j] = make_double2(0.0, 0.0);
  // the loads of several columns in flight per thread: at 2 CTAs of 128 threads per SM the kernel is bound by memory
  // latency, not by bandwidth or FP64 rate
#pragma unroll 4
  for (int t = 0; t < a.n_tx; ++t) {
    double2 h[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) h[i] = svd_ld(a.H, base + i * rx_stride + static_cast<size_t>(t) * a.n_sc, HD ? 1 : 0);
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = i; j < NR; ++j) {           // G[i][j] += h_i conj(h_j)
        G[i][j].x = fma(h[i].x, h[j].x, fma(h[i].y, h[j].y, G[i][j].x));
        G[i][j].y = fma(h[i].y, h[j].x, fma(-h[i].x, h[j].y, G[i][j].y));
      }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) G[i][j] = make_double2(G[j][i].x, -G[j][i].y);

  // ---- cyclic Jacobi on the Hermitian G:  G <- J^H G J,  U <- U J,  J = [[c, s e^{i phi}], [-s e^{-i phi}, c]]
  double2 U[NR][NR];
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < NR; ++j) U[i][j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < NR; ++i) tr += G[i][i].x;
  for (int sweep = 0; sweep < 16 && NR > 1; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < NR; ++p)
#pragma unroll
      for (int q = p + 1; q < NR; ++q) off += G[p][q].x * G[p][q].x + G[p][q].y * G[p][q].y;
    if (off <= 1e-30 * tr * tr) break;
#pragma unroll
    for (int p = 0; p < NR; ++p)
#pragma unroll
      for (int q = p + 1; q < NR; ++q) {
        const double b = sqrt(G[p][q].x * G[p][q].x + G[p][q].y * G[p][q].y);
        if (b == 0.0) continue;
        const double phr = G[p][q].x / b, phi = G[p][q].y / b;            // e^{i phi}
        const double tau = (G[q][q].x - G[p][p].x) / (2.0 * b);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = rsqrt(1.0 + t * t), s = t * c;
        const double sr = s * phr, si = s * phi;                           // s e^{i phi}
        // G <- J^H G J touches rows / columns p and q only, and G stays Hermitian: update column entries (m, p), (m, q)
        // for m outside {p, q}, mirror them into rows p and q, and set the 2 x 2 block in closed form
        // (G_pq' = 0, G_pp' = G_pp - t b, G_qq' = G_qq + t b) -- half the multiplies of the two-sided product.
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          if (m != p && m != q) {
            const double2 gp = G[m][p], gq = G[m][q];
            // col p' = c gp - conj(se) gq ; col q' = se gp + c gq
            const double2 np_ = make_double2(c * gp.x - (sr * gq.x + si * gq.y), c * gp.y - (sr * gq.y - si * gq.x));
            const double2 nq = make_double2(sr * gp.x - si * gp.y + c * gq.x, sr * gp.y + si * gp.x + c * gq.y);
            G[m][p] = np_; G[m][q] = nq;
            G[p][m] = make_double2(np_.x, -np_.y);
            G[q][m] = make_double2(nq.x, -nq.y);
          }
          const double2 up = U[m][p], uq = U[m][q];
          U[m][p] = make_double2(c * up.x - (sr * uq.x + si * uq.y), c * up.y - (sr * uq.y - si * uq.x));
          U[m][q] = make_double2(sr * up.x - si * up.y + c * uq.x, sr * up.y + si * up.x + c * uq.y);
        }
        G[p][p].x -= t * b; G[q][q].x += t * b;
        G[p][p].y = 0.0; G[q][q].y = 0.0;
        G[p][q] = make_double2(0.0, 0.0);
        G[q][p] = make_double2(0.0, 0.0);
      }
  }

  // ---- descending order (selection over NR <= 8 values), singular values
  int ord[NR];
  double lam[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) { ord[i] = i; lam[i] = G[i][i].x; }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = i + 1; j < NR; ++j)
      if (lam[j] > lam[i]) { const double tl = lam[i]; lam[i] = lam[j]; lam[j] = tl; const int to = ord[i]; ord[i] = ord[j]; ord[j] = to; }
  double sig[NR], inv_sig[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    sig[r] = sqrt(fmax(lam[r], 0.0));
    inv_sig[r] = sig[r] > 1e-7 * sqrt(fmax(lam[0], 0.0)) && sig[r] > 0.0 ? 1.0 / sig[r] : 0.0;
    const size_t o = (pkt * NR + r) * static_cast<size_t>(a.n_sc) + k;
    if (a.out_double) reinterpret_cast<double*>(a.sigma)[o] = sig[r];
    else reinterpret_cast<float*>(a.sigma)[o] = static_cast<float>(sig[r]);
  }
  if (!a.V1) return;

  // ---- pass 2: V1[t][r] = sum_i conj(H[i][t]) U[i][ord r] / sigma_r
  // the loads of several columns in flight per thread: at 2 CTAs of 128 threads per SM the kernel is bound by memory
  // latency, not by bandwidth or FP64 rate
#pragma unroll 4
  for (int t = 0; t < a.n_tx; ++t) {
    double2 h[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) h[i] = svd_ld(a.H, base + i * rx_stride + static_cast<size_t>(t) * a.n_sc, HD ? 1 : 0);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double vx = 0.0, vy = 0.0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        // the sort above permuted only `ord`; pick column ord[r] of U without dynamic register indexing
        double2 u = make_double2(0.0, 0.0);
#pragma unroll
        for (int cidx = 0; cidx < NR; ++cidx)
          if (cidx == ord[r]) u = U[i][cidx];
        vx = fma(h[i].x, u.x, fma(h[i].y, u.y, vx));                        // conj(h) * u
        vy = fma(h[i].x, u.y, fma(-h[i].y, u.x, vy));
      }
      vx *= inv_sig[r];
      vy *= inv_sig[r];
      const size_t o = ((pkt * NR + r) * a.n_tx + t) * static_cast<size_t>(a.n_sc) + k;
      if (a.out_double) reinterpret_cast<double2*>(a.V1)[o] = make_double2(vx, vy);
      else reinterpret_cast<float2*>(a.V1)[o] = make_float2(static_cast<float>(vx), static_cast<float>(vy));
    }
  }
}

// n_rx = 5..8: the 2 * NR^2 complex doubles of G and U do not fit the register file (ptxas spills 12 KB per thread
// for NR = 8), so they live in shared memory, entry-major [entry][thread] (conflict-free: consecutive threads hit
// consecutive 16-byte slots), 64 threads per CTA.  Same arithmetic as svd_gram_kernel.
constexpr int kSvdSmemThreads = 64;
template <int NR>
constexpr int svd_smem_bytes() { return 2 * NR * NR * kSvdSmemThreads * 16; }

template <int NR>
__global__ void __launch_bounds__(kSvdSmemThreads) svd_gram_smem_kernel(const SvdArgs a) {
  extern __shared__ double2 svd_sm[];
  double2* sG = svd_sm + threadIdx.x;
  double2* sU = svd_sm + NR * NR * kSvdSmemThreads + threadIdx.x;
#define G_(i, j) sG[((i) * NR + (j)) * kSvdSmemThreads]
#define U_(i, j) sU[((i) * NR + (j)) * kSvdSmemThreads]
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_sc) return;
  const size_t pkt = blockIdx.y;
  const size_t base = pkt * NR * a.n_tx * static_cast<size_t>(a.n_sc) + k;
  const size_t rx_stride = static_cast<size_t>(a.n_tx) * a.n_sc;
  for (int i = 0; i < NR; ++i)
    for (int j = 0; j < NR; ++j) { G_(i, j) = make_double2(0.0, 0.0); U_(i, j) = make_double2(i == j ? 1.0 : 0.0, 0.0); }
  for (int t = 0; t < a.n_tx; ++t) {
    double2 h[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) h[i] = svd_ld(a.H, base + i * rx_stride + static_cast<size_t>(t) * a.n_sc, a.h_double);
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = i; j < NR; ++j) {
        double2 g = G_(i, j);
        g.x = fma(h[i].x, h[j].x, fma(h[i].y, h[j].y, g.x));
        g.y = fma(h[i].y, h[j].x, fma(-h[i].x, h[j].y, g.y));
        G_(i, j) = g;
      }
  }
  double tr = 0.0;
  for (int i = 0; i < NR; ++i) {
    tr += G_(i, i).x;
    for (int j = 0; j < i; ++j) { const double2 g = G_(j, i); G_(i, j) = make_double2(g.x, -g.y); }
  }
  for (int sweep = 0; sweep < 20; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < NR; ++p)
      for (int q = p + 1; q < NR; ++q) { const double2 g = G_(p, q); off += g.x * g.x + g.y * g.y; }
    if (off <= 1e-30 * tr * tr) break;
    for (int p = 0; p < NR; ++p)
      for (int q = p + 1; q < NR; ++q) {
        const double2 gpq = G_(p, q);
        const double b = sqrt(gpq.x * gpq.x + gpq.y * gpq.y);
        if (b == 0.0) continue;
        const double phr = gpq.x / b, phi = gpq.y / b;
        const double tau = (G_(q, q).x - G_(p, p).x) / (2.0 * b);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = rsqrt(1.0 + t * t), s = t * c;
        const double sr = s * phr, si = s * phi;
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          const double2 gp = G_(m, p), gq = G_(m, q);
          G_(m, p) = make_double2(c * gp.x - (sr * gq.x + si * gq.y), c * gp.y - (sr * gq.y - si * gq.x));
          G_(m, q) = make_double2(sr * gp.x - si * gp.y + c * gq.x, sr * gp.y + si * gp.x + c * gq.y);
          const double2 up = U_(m, p), uq = U_(m, q);
          U_(m, p) = make_double2(c * up.x - (sr * uq.x + si * uq.y), c * up.y - (sr * uq.y - si * uq.x));
          U_(m, q) = make_double2(sr * up.x - si * up.y + c * uq.x, sr * up.y + si * up.x + c * uq.y);
        }
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          const double2 gp = G_(p, m), gq = G_(q, m);
          G_(p, m) = make_double2(c * gp.x - (sr * gq.x - si * gq.y), c * gp.y - (sr * gq.y + si * gq.x));
          G_(q, m) = make_double2(sr * gp.x + si * gp.y + c * gq.x, sr * gp.y - si * gp.x + c * gq.y);
        }
      }
  }
  int ord[NR];
  double lam[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) { ord[i] = i; lam[i] = G_(i, i).x; }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = i + 1; j < NR; ++j)
      if (lam[j] > lam[i]) { const double tl = lam[i]; lam[i] = lam[j]; lam[j] = tl; const int to = ord[i]; ord[i] = ord[j]; ord[j] = to; }
  double inv_sig[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const double sg = sqrt(fmax(lam[r], 0.0));
    inv_sig[r] = sg > 1e-7 * sqrt(fmax(lam[0], 0.0)) && sg > 0.0 ? 1.0 / sg : 0.0;
    const size_t o = (pkt * NR + r) * static_cast<size_t>(a.n_sc) + k;
    if (a.out_double) reinterpret_cast<double*>(a.sigma)[o] = sg;
    else reinterpret_cast<float*>(a.sigma)[o] = static_cast<float>(sg);
  }
  if (!a.V1) return;
  for (int t = 0; t < a.n_tx; ++t) {
    double2 h[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) h[i] = svd_ld(a.H, base + i * rx_stride + static_cast<size_t>(t) * a.n_sc, a.h_double);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double vx = 0.0, vy = 0.0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const double2 u = U_(i, ord[r]);
        vx = fma(h[i].x, u.x, fma(h[i].y, u.y, vx));
        vy = fma(h[i].x, u.y, fma(-h[i].y, u.x, vy));
      }
      vx *= inv_sig[r];
      vy *= inv_sig[r];
      const size_t o = ((pkt * NR + r) * a.n_tx + t) * static_cast<size_t>(a.n_sc) + k;
      if (a.out_double) reinterpret_cast<double2*>(a.V1)[o] = make_double2(vx, vy);
      else reinterpret_cast<float2*>(a.V1)[o] = make_float2(static_cast<float>(vx), static_cast<float>(vy));
    }
  }
#undef G_
#undef U_
}

}  // namespace mm
