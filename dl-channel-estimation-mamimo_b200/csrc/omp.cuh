// Orthogonal matching pursuit over a steering dictionary: the transmit side of the hybrid-precoder consumer
// (after the per-subcarrier SVD of svd.cuh).
//
// Replaces, for every (packet, tone) of a batch at once,
//   packet_generation/phased_arr/omphybweights.m:178-179  [Fbb,Frf] = ompdecomp(Fopt,At,'MaxSparsity',NtRF);
//                                                         Fbb = sqrt(Ns)*Fbb/norm(Frf*Fbb,'fro');
//   packet_generation/phased_arr/ompdecomp.m:101-121      the greedy loop, identity weight:
//       Psi = At' * Wres;  k = argmax_k sum_s |Psi(k,s)|^2;  refit ALL chosen columns to Fopt by least squares;
//       Wres = (Fopt - atoms*coeff) / ||.||_F;  stop after NtRF columns or when that norm <= eps
// as called once per subcarrier from pg/BER_test_maMIMO_LTF.m:372 with one dictionary per packet batch.
//
// Everything is FP64 like the MATLAB original: the output that matters is an INTEGER (which dictionary column), and it
// must not flip on rounding.  Two kernels per greedy round:
//   omp_corr_kernel   the contraction.  One CTA = 64 tones of one packet against the whole dictionary, 64 rays at a
//                     time: conj(At) tile [n_tx][64] (double-buffered: the next chunk travels through registers while
//                     this one is contracted) and residual tiles [n_tx][64] in shared memory, each thread a 4x4
//                     register block of complex accumulators (16 complex FMAs per 8 shared-memory loads: DFMA-bound),
//                     energies summed over the Ns columns, running first-maximum per tone, one cross-thread reduction.
//                     FLOPs: 8 * n_rays * n_tx * Ns per tone and round -- at 500 rays this is >99 % of the work.
//   omp_refit_kernel  one thread per tone (lanes along k: the [pkt][..][k] tensors stay coalesced): Gram matrix of the
//                     chosen columns, right-hand side, complex Gaussian elimination with partial pivoting (what `\`
//                     does), residual, its norm, the normalised residual for the next round, and Fbb scaled as :179.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

constexpr int kOmpTile = 64;          // tones per CTA and rays per dictionary chunk
constexpr int kOmpThreads = 256;      // 16 x 16 threads, 4 x 4 outputs each
constexpr int kOmpMaxRf = 8;
constexpr int kOmpMaxNs = 8;

struct OmpArgs {
  const void* F;            // Fopt: complex [n_pkt][f_rows][n_tx][n_sc], rows 0..ns-1 used (float2 or double2)
  int f_double, f_rows;
  double2* Wres;            // residual workspace [n_pkt][ns][n_tx][n_sc]
  const double2* AtcT;      // conj(At) transposed: [n_tx][n_rays_pad] (n_rays_pad multiple of 64, zero padded)
  const double2* At;        // At as given: [n_rays][n_tx]
  int32_t* idx;             // [n_pkt][n_rf][n_sc], -1 = not chosen (stopped early)
  float* err;               // [n_pkt][n_rf][n_sc] residual Frobenius norm after each round
  void* Fbb;                // complex [n_pkt][ns][n_rf][n_sc] (float2 or double2)
  int fbb_double;
  uint8_t* active;          // [n_pkt][n_sc]
  int n_tx, n_sc, n_rays, n_rays_pad, ns, n_rf, round;     // round = 0-based index of the column being chosen
};

__device__ __forceinline__ double2 omp_ld(const void* p, size_t i, int is_double) {
  if (is_double) return __ldg(reinterpret_cast<const double2*>(p) + i);
  const float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
  return make_double2(v.x, v.y);
}
__device__ __forceinline__ void cfma(double2& acc, const double2 a, const double2 b) {   // acc += a * b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 cmulc(const double2 a, const double2 b) {             // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

// Shared memory: two conj(At) chunks (the next one is fetched into registers while the current one is contracted) and
// the residual tiles -- all Ns of them when they fit (loaded once per CTA), else one that is reloaded per chunk.
constexpr size_t kOmpSmemBudget = 200 * 1024;
inline size_t omp_tile_bytes(int n_tx) { return static_cast<size_t>(n_tx) * kOmpTile * sizeof(double2); }
inline bool omp_w_resident(int n_tx, int ns) { return (2 + static_cast<size_t>(ns)) * omp_tile_bytes(n_tx) <= kOmpSmemBudget; }
// 16-byte asynchronous copy global -> shared (LDGSTS; no register staging), committed / awaited per thread
__device__ __forceinline__ void omp_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void omp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void omp_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

inline size_t omp_corr_smem(int n_tx, int ns) { return (2 + (omp_w_resident(n_tx, ns) ? ns : 1)) * omp_tile_bytes(n_tx); }
constexpr int kOmpMaxTx = 100;        // 3 tiles of n_tx x 64 complex doubles within the budget

template <int PF>                     // PF = ceil(n_tx * 64 / 256): conj(At) elements each thread prefetches per chunk
__global__ void __launch_bounds__(kOmpThreads) omp_corr_kernel(const OmpArgs a, const int w_resident) {
  extern __shared__ __align__(16) unsigned char omp_smem[];
  const int nt = a.n_tx;
  const int tile = nt * kOmpTile;
  double2* As = reinterpret_cast<double2*>(omp_smem);                 // [2][n_tx][64] conj(At) chunks
  double2* Ws = As + 2 * static_cast<size_t>(tile);                   // [ns or 1][n_tx][64] residual columns of 64 tones
  __shared__ double red_e[16][kOmpTile];
  __shared__ int red_r[16][kOmpTile];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * kOmpTile;
  const size_t pkt = blockIdx.y;
  // the residual of round 0 is Fopt itself (rows of the caller's tensor), later the workspace
  const bool first = a.round == 0;
  const size_t src_rows = first ? a.f_rows : a.ns;
  const void* src = first ? a.F : a.Wres;
  const int src_double = first ? a.f_double : 1;
  auto load_w = [&](int s, double2* dst) {
    for (int i = tid; i < tile; i += kOmpThreads) {
      const int t = i / kOmpTile, k = k0 + i % kOmpTile;
      dst[i] = k < a.n_sc ? omp_ld(src, ((pkt * src_rows + s) * nt + t) * static_cast<size_t>(a.n_sc) + k, src_double)
                          : make_double2(0.0, 0.0);
    }
  };
  if (w_resident)
    for (int s = 0; s < a.ns; ++s) load_w(s, Ws + static_cast<size_t>(s) * tile);
  for (int i = tid; i < tile; i += kOmpThreads)
    As[i] = a.AtcT[static_cast<size_t>(i / kOmpTile) * a.n_rays_pad + i % kOmpTile];
  __syncthreads();

  double best_e[4];
  int best_r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { best_e[j] = -1.0; best_r[j] = 0; }

  int buf = 0;
  for (int c0 = 0; c0 < a.n_rays_pad; c0 += kOmpTile, buf ^= 1) {
    const double2* Ac = As + static_cast<size_t>(buf) * tile;
    const bool more = c0 + kOmpTile < a.n_rays_pad;
    double2 pf[PF];
    if (more) {
#pragma unroll
      for (int q = 0; q < PF; ++q) {
        const int i = tid + q * kOmpThreads;
        if (i < tile) pf[q] = __ldg(a.AtcT + static_cast<size_t>(i / kOmpTile) * a.n_rays_pad + c0 + kOmpTile + i % kOmpTile);
      }
    }
    double E[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) E[i][j] = 0.0;
    for (int s = 0; s < a.ns; ++s) {
      const double2* Wc = Ws + (w_resident ? static_cast<size_t>(s) * tile : 0);
      if (!w_resident) {
        __syncthreads();                // the previous column is still being read
        load_w(s, Ws);
        __syncthreads();
      }
      double2 acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);
#pragma unroll 2
      for (int t = 0; t < nt; ++t) {
        double2 av[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = Ac[t * kOmpTile + ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Wc[t * kOmpTile + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) cfma(acc[i][j], av[i], wv[j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) E[i][j] += acc[i][j].x * acc[i][j].x + acc[i][j].y * acc[i][j].y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // rays ascend with i and with c0: strict > keeps the first maximum
      const int ray = c0 + ty + 16 * i;
      if (ray < a.n_rays) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (E[i][j] > best_e[j]) { best_e[j] = E[i][j]; best_r[j] = ray; }
      }
    }
    if (more) {                                       // the other buffer was last read one chunk ago (barrier below)
      double2* An = As + static_cast<size_t>(buf ^ 1) * tile;
#pragma unroll
      for (int q = 0; q < PF; ++q) {
        const int i = tid + q * kOmpThreads;
        if (i < tile) An[i] = pf[q];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red_e[ty][tx + 16 * j] = best_e[j]; red_r[ty][tx + 16 * j] = best_r[j]; }
  __syncthreads();
  if (tid < kOmpTile) {
    const int k = k0 + tid;
    if (k < a.n_sc) {
      double be = red_e[0][tid];
      int br = red_r[0][tid];
      for (int y = 1; y < 16; ++y) {
        const double e = red_e[y][tid];
        const int r = red_r[y][tid];
        if (e > be || (e == be && r < br)) { be = e; br = r; }
      }
      const bool on = a.active[pkt * a.n_sc + k] != 0;
      a.idx[(pkt * a.n_rf + a.round) * static_cast<size_t>(a.n_sc) + k] = on ? br : -1;
    }
  }
}

// Ns = 1 (the reference's own use: numSTS = 1, pg/generate_maMIMO_LTF.m:23) -- no energy sum over columns, so the
// registers it would take go into a taller block: 8 rays x 4 tones per thread (12 shared-memory loads per 32 complex
// FMAs instead of 8 per 16: a 128-bit shared load costs 4 wavefronts per warp whatever it fetches, and at 4x4 the
// load pipe is as busy as the FP64 pipe).  128 threads, two CTAs per SM; the next conj(At) chunk arrives by cp.async
// (LDGSTS) in a second buffer while the current one is contracted.
constexpr int kOmp1Threads = 128;
inline size_t omp_corr1_smem(int n_tx) { return 3 * omp_tile_bytes(n_tx); }

__global__ void __launch_bounds__(kOmp1Threads, 2) omp_corr1_kernel(const OmpArgs a) {
  extern __shared__ __align__(16) unsigned char omp_smem[];
  const int nt = a.n_tx;
  const int tile = nt * kOmpTile;
  double2* As = reinterpret_cast<double2*>(omp_smem);                 // [2][n_tx][64] conj(At) chunks
  double2* Ws = As + 2 * static_cast<size_t>(tile);                   // [n_tx][64] residual of 64 tones
  __shared__ double red_e[8][kOmpTile];
  __shared__ int red_r[8][kOmpTile];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;          // ty 0..7: rays ty + 8 i
  const int k0 = blockIdx.x * kOmpTile;
  const size_t pkt = blockIdx.y;
  const bool first = a.round == 0;
  const size_t src_rows = first ? a.f_rows : 1;
  const void* src = first ? a.F : a.Wres;
  const int src_double = first ? a.f_double : 1;
  auto fetch = [&](int c0, double2* dst) {
    for (int i = tid; i < tile; i += kOmp1Threads)
      omp_cp_async16(dst + i, a.AtcT + static_cast<size_t>(i / kOmpTile) * a.n_rays_pad + c0 + i % kOmpTile);
    omp_cp_async_commit();
  };
  fetch(0, As);
  for (int i = tid; i < tile; i += kOmp1Threads) {
    const int t = i / kOmpTile, k = k0 + i % kOmpTile;
    Ws[i] = k < a.n_sc ? omp_ld(src, (pkt * src_rows * nt + t) * static_cast<size_t>(a.n_sc) + k, src_double)
                       : make_double2(0.0, 0.0);
  }
  double best_e[4];
  int best_r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { best_e[j] = -1.0; best_r[j] = 0; }
  int buf = 0;
  for (int c0 = 0; c0 < a.n_rays_pad; c0 += kOmpTile, buf ^= 1) {
    omp_cp_async_wait_all();
    __syncthreads();                    // chunk c0 has landed for everyone; nobody still reads the other buffer
    if (c0 + kOmpTile < a.n_rays_pad) fetch(c0 + kOmpTile, As + static_cast<size_t>(buf ^ 1) * tile);
    const double2* Ac = As + static_cast<size_t>(buf) * tile;
    double2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);
#pragma unroll 2
    for (int t = 0; t < nt; ++t) {
      double2 av[8], wv[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = Ac[t * kOmpTile + ty + 8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[t * kOmpTile + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma(acc[i][j], av[i], wv[j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {                     // rays ascend with i and with c0: strict > keeps the first maximum
      const int ray = c0 + ty + 8 * i;
      if (ray < a.n_rays) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double e = acc[i][j].x * acc[i][j].x + acc[i][j].y * acc[i][j].y;
          if (e > best_e[j]) { best_e[j] = e; best_r[j] = ray; }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red_e[ty][tx + 16 * j] = best_e[j]; red_r[ty][tx + 16 * j] = best_r[j]; }
  __syncthreads();
  if (tid < kOmpTile) {
    const int k = k0 + tid;
    if (k < a.n_sc) {
      double be = red_e[0][tid];
      int br = red_r[0][tid];
      for (int y = 1; y < 8; ++y) {
        const double e = red_e[y][tid];
        const int r = red_r[y][tid];
        if (e > be || (e == be && r < br)) { be = e; br = r; }
      }
      const bool on = a.active[pkt * a.n_sc + k] != 0;
      a.idx[(pkt * a.n_rf + a.round) * static_cast<size_t>(a.n_sc) + k] = on ? br : -1;
    }
  }
}

// Least-squares refit, residual and outputs of one round: one thread per (packet, tone).
template <int MM, int NSM>
__global__ void __launch_bounds__(128) omp_refit_kernel(const OmpArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_sc) return;
  const size_t pkt = blockIdx.y;
  const int nt = a.n_tx, ns = a.ns, m = a.round + 1;
  const size_t nsc = a.n_sc;
  const size_t e_at = (pkt * a.n_rf + a.round) * nsc + k;
  if (!a.active[pkt * nsc + k]) {                     // stopped in an earlier round: carry the norm, nothing else changes
    a.err[e_at] = a.err[e_at - nsc];
    return;
  }
  int sel[MM];
#pragma unroll
  for (int i = 0; i < MM; ++i) sel[i] = i < m ? a.idx[(pkt * a.n_rf + i) * nsc + k] : 0;

  // G = A^H A (m x m), R = A^H Fopt (m x ns)
  double2 G[MM][MM], R[MM][NSM];
#pragma unroll
  for (int i = 0; i < MM; ++i) {
#pragma unroll
    for (int j = 0; j < MM; ++j) G[i][j] = make_double2(i == j && i >= m ? 1.0 : 0.0, 0.0);
#pragma unroll
    for (int s = 0; s < NSM; ++s) R[i][s] = make_double2(0.0, 0.0);
  }
  for (int t = 0; t < nt; ++t) {
    double2 at[MM], f[NSM];
#pragma unroll
    for (int i = 0; i < MM; ++i) at[i] = i < m ? a.At[static_cast<size_t>(sel[i]) * nt + t] : make_double2(0.0, 0.0);
#pragma unroll
    for (int s = 0; s < NSM; ++s)
      f[s] = s < ns ? omp_ld(a.F, ((pkt * a.f_rows + s) * nt + t) * nsc + k, a.f_double) : make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < MM; ++i) {
#pragma unroll
      for (int j = 0; j < MM; ++j) {
        const double2 p = cmulc(at[i], at[j]);
        G[i][j].x += p.x; G[i][j].y += p.y;
      }
#pragma unroll
      for (int s = 0; s < NSM; ++s) {
        const double2 p = cmulc(at[i], f[s]);
        R[i][s].x += p.x; R[i][s].y += p.y;
      }
    }
  }
  // X = G \ R: Gaussian elimination with partial pivoting (rows >= m are identity rows with zero right-hand sides)
#pragma unroll
  for (int c = 0; c < MM; ++c) {
    int piv = c;
    double pm = G[c][c].x * G[c][c].x + G[c][c].y * G[c][c].y;
#pragma unroll
    for (int r = c + 1; r < MM; ++r) {
      const double v = G[r][c].x * G[r][c].x + G[r][c].y * G[r][c].y;
      if (v > pm) { pm = v; piv = r; }
    }
#pragma unroll
    for (int r = c + 1; r < MM; ++r) {
      if (r == piv) {
#pragma unroll
        for (int j = 0; j < MM; ++j) { const double2 tmp = G[c][j]; G[c][j] = G[r][j]; G[r][j] = tmp; }
#pragma unroll
        for (int s = 0; s < NSM; ++s) { const double2 tmp = R[c][s]; R[c][s] = R[r][s]; R[r][s] = tmp; }
      }
    }
    const double inv = pm > 0.0 ? 1.0 / pm : 0.0;
    const double2 pinv = make_double2(G[c][c].x * inv, -G[c][c].y * inv);       // 1 / pivot
#pragma unroll
    for (int r = c + 1; r < MM; ++r) {
      const double2 g = G[r][c];
      const double2 l = make_double2(g.x * pinv.x - g.y * pinv.y, g.x * pinv.y + g.y * pinv.x);
      const double2 nl = make_double2(-l.x, -l.y);
#pragma unroll
      for (int j = 0; j < MM; ++j) if (j >= c) cfma(G[r][j], nl, G[c][j]);
#pragma unroll
      for (int s = 0; s < NSM; ++s) cfma(R[r][s], nl, R[c][s]);
    }
  }
#pragma unroll
  for (int c = MM - 1; c >= 0; --c) {
    const double pm = G[c][c].x * G[c][c].x + G[c][c].y * G[c][c].y;
    const double inv = pm > 0.0 ? 1.0 / pm : 0.0;
    const double2 pinv = make_double2(G[c][c].x * inv, -G[c][c].y * inv);
#pragma unroll
    for (int s = 0; s < NSM; ++s) {
      double2 v = R[c][s];
#pragma unroll
      for (int j = 0; j < MM; ++j)
        if (j > c) { const double2 nj = make_double2(-G[c][j].x, -G[c][j].y); cfma(v, nj, R[j][s]); }
      R[c][s] = make_double2(v.x * pinv.x - v.y * pinv.y, v.x * pinv.y + v.y * pinv.x);    // X[c][s]
    }
  }
  // residual norm and the norm of atoms*coeff
  double e2 = 0.0, p2 = 0.0;
  for (int t = 0; t < nt; ++t) {
    double2 at[MM];
#pragma unroll
    for (int i = 0; i < MM; ++i) at[i] = i < m ? a.At[static_cast<size_t>(sel[i]) * nt + t] : make_double2(0.0, 0.0);
#pragma unroll
    for (int s = 0; s < NSM; ++s) {
      if (s < ns) {
        double2 ax = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < MM; ++i) cfma(ax, at[i], R[i][s]);
        const double2 f = omp_ld(a.F, ((pkt * a.f_rows + s) * nt + t) * nsc + k, a.f_double);
        const double dx = f.x - ax.x, dy = f.y - ax.y;
        e2 += dx * dx + dy * dy;
        p2 += ax.x * ax.x + ax.y * ax.y;
      }
    }
  }
  const double en = sqrt(e2);
  a.err[e_at] = static_cast<float>(en);
  const bool last = m == a.n_rf;
  const bool stop = !(en > 2.220446049250313e-16);              // ompdecomp.m:105 `Errnorm > eps`
  if (stop) a.active[pkt * nsc + k] = 0;
  if (!last && !stop) {
    const double ie = 1.0 / en;
    for (int t = 0; t < nt; ++t) {
      double2 at[MM];
#pragma unroll
      for (int i = 0; i < MM; ++i) at[i] = i < m ? a.At[static_cast<size_t>(sel[i]) * nt + t] : make_double2(0.0, 0.0);
#pragma unroll
      for (int s = 0; s < NSM; ++s) {
        if (s < ns) {
          double2 ax = make_double2(0.0, 0.0);
#pragma unroll
          for (int i = 0; i < MM; ++i) cfma(ax, at[i], R[i][s]);
          const double2 f = omp_ld(a.F, ((pkt * a.f_rows + s) * nt + t) * nsc + k, a.f_double);
          a.Wres[((pkt * ns + s) * nt + t) * nsc + k] = make_double2((f.x - ax.x) * ie, (f.y - ax.y) * ie);
        }
      }
    }
  }
  // Fbb = sqrt(Ns) * coeff / ||atoms*coeff||_F, stored as the reference returns it (transposed): [s][i]
  const double sc = p2 > 0.0 ? sqrt(static_cast<double>(ns) / p2) : 0.0;
#pragma unroll
  for (int s = 0; s < NSM; ++s) {
    if (s < ns) {
      for (int i = 0; i < a.n_rf; ++i) {
        double2 v = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < MM; ++q) if (q == i && q < m) v = make_double2(R[q][s].x * sc, R[q][s].y * sc);
        const size_t o = ((pkt * ns + s) * a.n_rf + i) * nsc + k;
        if (a.fbb_double) reinterpret_cast<double2*>(a.Fbb)[o] = v;
        else reinterpret_cast<float2*>(a.Fbb)[o] = make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
      }
    }
  }
}

}  // namespace mm
