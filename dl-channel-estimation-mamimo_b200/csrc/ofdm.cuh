// OFDM demodulation front-end: the step immediately before the LS estimate.
//
// Replaces the toolbox call  rxOFDM = ofdmdemod(inputRXSig, FFT, CP, symOffset, nullIdx, pilotIdx)
//   (packet_generation/phased_arr/generate_maMIMO_LTF.m:336-338) and its numpy mirror
//   massiveMIMO_dataGenerator.py:437-453 (F-order reshape into symbols, CP removal with symbol offset,
//   FFT, fftshift along frequency, removal of null + pilot carriers).
//
// A CTA transforms `syms_per_cta` OFDM symbols of one (packet, rx) stream: coalesced load of each FFT window
// (rotated so the true symbol start comes first, dataGenerator.py:442), radix-4 Stockham autosort FFT in
// shared memory (ping-pong buffers, a trailing radix-2 stage when log2(FFT) is odd, twiddles from a
// host-computed FP64->FP32 table), then a gather of the data carriers written straight into the LS stage's
// layout Y[pkt][rx][sym][k].  HBM-bound: (FFT+CP)*8 B in, Nsc*8 B out per symbol.
#pragma once
#include "ptx.cuh"

namespace mm {

struct OfdmArgs {
  const void* x;            // complex [n_pkt*n_rx][n_sym*(fft+cp)]  float2 or double2
  float2* Y;                // complex64 [n_pkt*n_rx][n_sym][n_sc]
  const float2* twiddle;    // per-stage compact tables, see kernel: [stage][r=1..3][k<ns] then the radix-2 tail [k<ns]
  int n_twiddle;            // entries in that table
  const int* bins;          // [n_sc] natural-order FFT bin of each kept carrier
  int fft_len, cp_len, sym_offset, n_sym, n_sc;
  int syms_per_cta;         // symbols transformed side by side in one CTA
  long long total_syms;     // n_pkt * n_rx * n_sym
  int x_double;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) ofdm_demod_kernel(const OfdmArgs a) {
  extern __shared__ float2 sm_fft[];                 // [syms_per_cta][2][fft_len] then the twiddle tables
  const int N = a.fft_len;
  const int S = a.syms_per_cta;
  const long long sym0 = static_cast<long long>(blockIdx.x) * S;
  const int n_here = static_cast<int>(min(static_cast<long long>(S), a.total_syms - sym0));
  const int sym_len = N + a.cp_len;
  const int lgN = 31 - __clz(N);                      // N is a power of two: shifts, not divisions
  // twiddles: stage ns uses w_r[k] = exp(-2 pi i r k / (4 ns)), k < ns, stored contiguously per (stage, r) so
  // that consecutive butterflies (consecutive k) read consecutive entries: conflict-free, no scattered gathers
  float2* tw = sm_fft + static_cast<size_t>(S) * 2 * N;
  for (int i = threadIdx.x; i < a.n_twiddle; i += blockDim.x) tw[i] = a.twiddle[i];
  // window[i] = x[ix(i)],  ix = [cp, fft+off) ++ [off, cp)   (dataGenerator.py:442).  Symbols of one (pkt,rx)
  // stream are contiguous in x, and streams follow each other, so global symbol g starts at g * sym_len.
  const int first = N + a.sym_offset - a.cp_len;
  for (int idx = threadIdx.x; idx < n_here * N; idx += blockDim.x) {
    const int s = idx >> lgN, i = idx & (N - 1);
    const int src = (i < first) ? (a.cp_len + i) : (a.sym_offset + (i - first));
    const size_t g = static_cast<size_t>(sym0 + s) * sym_len + src;
    float2 v;
    if (a.x_double) {
      const double2 d = __ldg(reinterpret_cast<const double2*>(a.x) + g);
      v = make_float2(static_cast<float>(d.x), static_cast<float>(d.y));
    } else {
      v = __ldg(reinterpret_cast<const float2*>(a.x) + g);
    }
    sm_fft[static_cast<size_t>(s) * 2 * N + i] = v;
  }
  __syncthreads();

  int cur = 0;                                        // which ping-pong half holds the data
  const int quarter = N >> 2;
  int ns = 1;
  int tw_off = 0;
  for (; ns * 4 <= N; tw_off += 3 * ns, ns <<= 2) {   // radix-4 stages
    const float2* w1 = tw + tw_off;
    const float2* w2 = w1 + ns;
    const float2* w3 = w2 + ns;
    for (int idx = threadIdx.x; idx < n_here * quarter; idx += blockDim.x) {
      const int s = idx >> (lgN - 2), j = idx & (quarter - 1);
      const float2* in = sm_fft + static_cast<size_t>(s) * 2 * N + cur * N;
      float2* out = sm_fft + static_cast<size_t>(s) * 2 * N + (cur ^ 1) * N;
      const int k = j & (ns - 1);
      const float2 v0 = in[j];
      const float2 v1 = cmulf(in[j + quarter], w1[k]);
      const float2 v2 = cmulf(in[j + 2 * quarter], w2[k]);
      const float2 v3 = cmulf(in[j + 3 * quarter], w3[k]);
      const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
      const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
      const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));      // (v1 - v3) * (-i)
      const int d = ((j - k) << 2) + k;
      out[d] = make_float2(t0.x + t2.x, t0.y + t2.y);
      out[d + ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
      out[d + 2 * ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
      out[d + 3 * ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
    }
    __syncthreads();
    cur ^= 1;
  }
  if (ns < N) {                                       // one radix-2 stage left (log2 N odd)
    const int half = N >> 1;
    const float2* w1 = tw + tw_off;
    for (int idx = threadIdx.x; idx < n_here * half; idx += blockDim.x) {
      const int s = idx >> (lgN - 1), j = idx & (half - 1);
      const float2* in = sm_fft + static_cast<size_t>(s) * 2 * N + cur * N;
      float2* out = sm_fft + static_cast<size_t>(s) * 2 * N + (cur ^ 1) * N;
      const int k = j & (ns - 1);
      const float2 u = in[j];
      const float2 v = cmulf(in[j + half], w1[k]);
      const int d = ((j - k) << 1) + k;
      out[d] = make_float2(u.x + v.x, u.y + v.y);
      out[d + ns] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
    cur ^= 1;
  }
  // symbols are contiguous in Y too ([stream][sym][k]): global symbol g writes at g * n_sc
  for (int s = 0; s < n_here; ++s) {
    const float2* res = sm_fft + static_cast<size_t>(s) * 2 * N + cur * N;
    float2* out = a.Y + static_cast<size_t>(sym0 + s) * a.n_sc;
    for (int k = threadIdx.x; k < a.n_sc; k += blockDim.x) out[k] = res[__ldg(a.bins + k)];
  }
}

// ------------------------------------------------------------------------------------------------------------
// FFT-256 specialisation (the reference numerology, generate_maMIMO_LTF.m:96-102): 256 = 16 x 16, so the whole
// transform is two radix-16 passes with the 16-point DFT in registers.  Pass 1 reads the window straight from
// global memory, pass 2 writes the kept carriers straight to global memory: one shared-memory buffer, one
// barrier, 16 B of shared-memory traffic per point (the generic kernel moves ~110 B).  16 threads per symbol,
// 8 symbols per 128-thread CTA, 17 KB of shared memory (rows padded 1-in-16: conflict-free both ways).
__host__ __device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)

__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 t0 = caddf(v0, v2), t1 = csubf(v0, v2), t2 = caddf(v1, v3), t3 = cmul_negi(csubf(v1, v3));
  v0 = caddf(t0, t2);
  v1 = caddf(t1, t3);
  v2 = csubf(t0, t2);
  v3 = csubf(t1, t3);
}

// forward DFT-16 in registers, natural order in and out (4 x 4 Cooley-Tukey):
//   X[k1 + 4 k2] = sum_n2 W16^(n2 k1) W4^(n2 k2) ( sum_n1 x[4 n1 + n2] W4^(n1 k1) )
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508978f, h = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // v[4 k1 + n2] = inner[n2][k1]
  v[4 + 1] = cmulf(v[4 + 1], make_float2(c1, -s1));    // W16^(n2 k1)
  v[4 + 2] = cmulf(v[4 + 2], make_float2(h, -h));
  v[4 + 3] = cmulf(v[4 + 3], make_float2(s1, -c1));
  v[8 + 1] = cmulf(v[8 + 1], make_float2(h, -h));
  v[8 + 2] = cmul_negi(v[8 + 2]);
  v[8 + 3] = cmulf(v[8 + 3], make_float2(-h, -h));
  v[12 + 1] = cmulf(v[12 + 1], make_float2(s1, -c1));
  v[12 + 2] = cmulf(v[12 + 2], make_float2(-h, -h));
  v[12 + 3] = cmulf(v[12 + 3], make_float2(-c1, s1));
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4 k1 + k2]
#pragma unroll
  for (int i = 0; i < 4; ++i)           // X[k1 + 4 k2] sits in v[4 k1 + k2]: transpose the 4 x 4 register tile
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const float2 t = v[4 * i + j];
      v[4 * i + j] = v[4 * j + i];
      v[4 * j + i] = t;
    }
}

struct Ofdm256Args {
  const void* x;
  float2* Y;
  const float2* tw2;        // [15][16]  exp(-2 pi i r k / 256), r = 1..15, k < 16 (pass-2 twiddles)
  const int* kmap;          // [256] natural FFT bin -> output column, or -1 when the carrier is dropped
  int cp_len, sym_offset, n_sc;
  long long total_syms;
  int x_double;
};

__global__ void __launch_bounds__(128) ofdm256_kernel(const Ofdm256Args a) {
  constexpr int N = 256, SYMS = 8;
  __shared__ float2 buf[SYMS][N + N / 16];
  __shared__ float2 tw[15 * 16];
  __shared__ int kmap[N];
  for (int i = threadIdx.x; i < 15 * 16; i += blockDim.x) tw[i] = a.tw2[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) kmap[i] = a.kmap[i];
  const int s = threadIdx.x >> 4, j = threadIdx.x & 15;
  const long long g = static_cast<long long>(blockIdx.x) * SYMS + s;
  const bool live = g < a.total_syms;
  float2 v[16];
  if (live) {
    // window[i] = x[ix(i)], ix = [cp, 256+off) ++ [off, cp)   (dataGenerator.py:442); pass 1 takes i = j + 16 r
    const int first = N + a.sym_offset - a.cp_len;
    const size_t base = static_cast<size_t>(g) * (N + a.cp_len);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = j + 16 * r;
      const int src = (i < first) ? (a.cp_len + i) : (a.sym_offset + (i - first));
      if (a.x_double) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(a.x) + base + src);
        v[r] = make_float2(static_cast<float>(d.x), static_cast<float>(d.y));
      } else {
        v[r] = __ldg(reinterpret_cast<const float2*>(a.x) + base + src);
      }
    }
    dft16(v);                                            // pass 1: ns = 1, no twiddles, output index 16 j + r
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[s][17 * j + r] = v[r];   // pad16(16 j + r) = 17 j + r
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = buf[s][pad16(j + 16 * r)];   // pass 2: ns = 16, k = j
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmulf(v[r], tw[(r - 1) * 16 + j]);
    dft16(v);                                            // output bin j + 16 r
    float2* out = a.Y + static_cast<size_t>(g) * a.n_sc;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int k = kmap[j + 16 * r];
      if (k >= 0) out[k] = v[r];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// FFT-512 / 1024 / 2048 / 4096 (the BASELINE.json numerologies): N = 16 x 16 x R3 with R3 = N / 256 in {2,4,8,16}.
// Three register passes (radix 16, radix 16, radix R3), 16 points per thread in every pass, N/16 threads per
// symbol and 4096/N symbols per 256-thread CTA.  Pass 1 reads the window straight from global memory, pass 3
// writes the kept carriers straight to global memory; the two exchanges in between go through ONE padded
// shared-memory buffer (conflict-free both ways), i.e. 32 B of shared-memory traffic per point where the generic
// radix-4 Stockham kernel moves 16 B per point per stage (80-96 B) plus its twiddle gathers.
template <int R>
__device__ __forceinline__ void dft_small(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dft_small<1>(float2 (&)[1]) {}
template <>
__device__ __forceinline__ void dft_small<2>(float2 (&v)[2]) {
  const float2 a = v[0], b = v[1];
  v[0] = caddf(a, b);
  v[1] = csubf(a, b);
}
template <>
__device__ __forceinline__ void dft_small<4>(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
// 8 = 4 x 2:  X[k1 + 4 k2] = sum_n2 W8^(n2 k1) (-1)^(n2 k2) ( sum_n1 x[2 n1 + n2] W4^(n1 k1) )
template <>
__device__ __forceinline__ void dft_small<8>(float2 (&v)[8]) {
  constexpr float h = 0.70710678118654752f;
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  v[3] = cmulf(v[3], make_float2(h, -h));
  v[5] = cmul_negi(v[5]);
  v[7] = cmulf(v[7], make_float2(-h, -h));
  const float2 x0 = caddf(v[0], v[1]), x4 = csubf(v[0], v[1]);
  const float2 x1 = caddf(v[2], v[3]), x5 = csubf(v[2], v[3]);
  const float2 x2 = caddf(v[4], v[5]), x6 = csubf(v[4], v[5]);
  const float2 x3 = caddf(v[6], v[7]), x7 = csubf(v[6], v[7]);
  v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}
template <>
__device__ __forceinline__ void dft_small<16>(float2 (&v)[16]) { dft16(v); }

struct OfdmR16Args {
  const void* x;
  float2* Y;
  const float2* tw2;        // [15][16]      exp(-2 pi i r k / 256),  r = 1..15,   k < 16   (pass 2)
  const float2* tw3;        // [R3-1][256]   exp(-2 pi i r k / N),    r = 1..R3-1, k < 256  (pass 3)
  const int* kmap;          // [N] natural FFT bin -> output column, or -1 when the carrier is dropped
  int cp_len, sym_offset, n_sc;
  long long total_syms;
  int x_double;
  uint32_t* flags;
};

template <int LOG2N>
__global__ void __launch_bounds__(256) ofdm_r16_kernel(const OfdmR16Args a) {
  constexpr int N = 1 << LOG2N;
  constexpr int T = N / 16;            // threads per symbol
  constexpr int SYMS = 256 / T;        // symbols per CTA == radix-R3 butterflies per thread in pass 3
  constexpr int R3 = N / 256;
  static_assert(LOG2N >= 9 && LOG2N <= 12, "three-pass kernel covers 512..4096");
  __shared__ float2 buf[SYMS][N + N / 16];
  __shared__ float2 tw[15 * 16];
  for (int i = threadIdx.x; i < 15 * 16; i += blockDim.x) tw[i] = a.tw2[i];
  const int s = threadIdx.x / T, j = threadIdx.x % T;
  const long long g = static_cast<long long>(blockIdx.x) * SYMS + s;
  const bool live = g < a.total_syms;
  float2 v[16];
  if (live) {
    // window[i] = x[ix(i)], ix = [cp, N+off) ++ [off, cp)   (dataGenerator.py:442); pass 1 takes i = j + T r
    const int first = N + a.sym_offset - a.cp_len;
    const size_t base = static_cast<size_t>(g) * (N + a.cp_len);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = j + T * r;
      const int src = (i < first) ? (a.cp_len + i) : (a.sym_offset + (i - first));
      if (a.x_double) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(a.x) + base + src);
        v[r] = make_float2(static_cast<float>(d.x), static_cast<float>(d.y));
      } else {
        v[r] = __ldg(reinterpret_cast<const float2*>(a.x) + base + src);
      }
    }
    dft16(v);                                            // pass 1: ns = 1, no twiddles, output index 16 j + r
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[s][17 * j + r] = v[r];   // pad16(16 j + r) = 17 j + r
  }
  __syncthreads();
  const int k = j & 15;
  if (live) {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = buf[s][pad16(j + T * r)];   // pass 2: ns = 16, k = j mod 16
  }
  __syncthreads();                                       // every thread holds its inputs: the buffer may be overwritten
  if (live) {
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmulf(v[r], tw[(r - 1) * 16 + k]);
    dft16(v);                                            // output index (j - k) 16 + k + 16 r
    const int o = (j - k) * 16 + k;
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[s][pad16(o + 16 * r)] = v[r];
  }
  __syncthreads();
  if (live) {
    float2* out = a.Y + static_cast<size_t>(g) * a.n_sc;
#pragma unroll
    for (int b = 0; b < SYMS; ++b) {                     // pass 3: ns = 256, butterfly jj = j + T b, k = jj
      const int jj = j + T * b;
      float2 w[R3];
#pragma unroll
      for (int r = 0; r < R3; ++r) w[r] = buf[s][pad16(jj + 256 * r)];
#pragma unroll
      for (int r = 1; r < R3; ++r) w[r] = cmulf(w[r], __ldg(a.tw3 + (r - 1) * 256 + jj));
      dft_small<R3>(w);                                  // output bin jj + 256 r
#pragma unroll
      for (int r = 0; r < R3; ++r) {
        const int col = __ldg(a.kmap + jj + 256 * r);
        if (col >= 0) out[col] = w[r];
      }
    }
  }
}

// Persistent, bulk-copy-fed variant of ofdm_r16_kernel (complex64 input, even cp_len / sym_offset): the FFT windows
// of the NEXT tile (4096/N symbols) are copied global -> shared by cp.async.bulk (SASS: UBLKCP) while the current
// tile runs its three passes.  One stage is enough: pass 1 pulls the whole stage into registers at the start of a
// tile, after which the stage is refilled for the tile gridDim.x ahead.  The cyclic prefix is never read: the copy
// starts at the window (two copies per symbol when sym_offset < cp_len rotates it).
template <int LOG2N>
constexpr int ofdm_r16_tma_smem() { return (1 << LOG2N) * 0 + 4096 * 8 + (4096 + 256) * 8 + 15 * 16 * 8 + 16 + 128; }

template <int LOG2N>
__global__ void __launch_bounds__(256, 3) ofdm_r16_tma_kernel(const OfdmR16Args a) {
  constexpr int N = 1 << LOG2N;
  constexpr int T = N / 16;
  constexpr int SYMS = 256 / T;
  constexpr int R3 = N / 256;
  extern __shared__ uint8_t sm_ofdm_raw[];
  float2* stage = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(sm_ofdm_raw) + 127) & ~static_cast<uintptr_t>(127));
  float2 (*buf)[N + N / 16] = reinterpret_cast<float2 (*)[N + N / 16]>(stage + SYMS * N);
  float2* tw = stage + SYMS * N + SYMS * (N + N / 16);
  uint64_t* full = reinterpret_cast<uint64_t*>(tw + 15 * 16);
  __shared__ uint32_t cta_abort;
  const long long n_tiles = (a.total_syms + SYMS - 1) / SYMS;
  const int first = N + a.sym_offset - a.cp_len;          // window = x[cp, N+off) ++ x[off, cp)  (dataGenerator.py:442)
  auto issue = [&](long long tile) {
    const long long g0 = tile * SYMS;
    const int live = static_cast<int>(min(static_cast<long long>(SYMS), a.total_syms - g0));
    mbar_arrive_expect_tx(full, static_cast<uint32_t>(live) * N * 8);
    const float2* x = reinterpret_cast<const float2*>(a.x);
    for (int s = 0; s < live; ++s) {
      const float2* base = x + static_cast<size_t>(g0 + s) * (N + a.cp_len);
      bulk_load_1d(stage + s * N, base + a.cp_len, static_cast<uint32_t>(first) * 8, full);
      if (first < N) bulk_load_1d(stage + s * N + first, base + a.sym_offset, static_cast<uint32_t>(N - first) * 8, full);
    }
  };
  if (threadIdx.x == 0) {
    cta_abort = 0;
    mbar_init(full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 15 * 16; i += blockDim.x) tw[i] = a.tw2[i];
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x);
  const int s = threadIdx.x / T, j = threadIdx.x % T;
  const int k = j & 15;
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const long long g = tile * SYMS + s;
    const bool live = g < a.total_syms;
    if (!mbar_wait(full, it & 1, &cta_abort, a.flags)) return;
    float2 v[16];
    if (live) {
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = stage[s * N + j + T * r];      // pass 1 inputs
    }
    __syncthreads();                                       // stage consumed (and last tile's pass 3 is done with buf)
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x);
    if (live) {
      dft16(v);                                            // pass 1: ns = 1, output index 16 j + r
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[s][17 * j + r] = v[r];
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = buf[s][pad16(j + T * r)];      // pass 2: ns = 16, k = j mod 16
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = cmulf(v[r], tw[(r - 1) * 16 + k]);
      dft16(v);
      if constexpr (R3 == 1) {                             // N = 256: two passes; output bin j + 16 r (k == j)
        float2* out = a.Y + static_cast<size_t>(g) * a.n_sc;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int col = __ldg(a.kmap + j + 16 * r);
          if (col >= 0) out[col] = v[r];
        }
      } else {
        const int o = (j - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[s][pad16(o + 16 * r)] = v[r];
      }
    }
    if constexpr (R3 > 1) {
      __syncthreads();
      if (live) {
        float2* out = a.Y + static_cast<size_t>(g) * a.n_sc;
#pragma unroll
        for (int b = 0; b < SYMS; ++b) {                   // pass 3: ns = 256, butterfly jj = j + T b
          const int jj = j + T * b;
          float2 w[R3];
#pragma unroll
          for (int r = 0; r < R3; ++r) w[r] = buf[s][pad16(jj + 256 * r)];
#pragma unroll
          for (int r = 1; r < R3; ++r) w[r] = cmulf(w[r], __ldg(a.tw3 + (r - 1) * 256 + jj));
          dft_small<R3>(w);
#pragma unroll
          for (int r = 0; r < R3; ++r) {
            const int col = __ldg(a.kmap + jj + 256 * r);
            if (col >= 0) out[col] = w[r];
          }
        }
      }
    }
  }
}

}  // namespace mm
