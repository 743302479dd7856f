// OFDM demodulation front-end: the step immediately before the LS estimate.
//
// Replaces the toolbox call  rxOFDM = ofdmdemod(inputRXSig, FFT, CP, symOffset, nullIdx, pilotIdx)
//   (packet_generation/phased_arr/generate_maMIMO_LTF.m:336-338) and its numpy mirror
//   massiveMIMO_dataGenerator.py:437-453 (F-order reshape into symbols, CP removal with symbol offset,
//   FFT, fftshift along frequency, removal of null + pilot carriers).
//
// A CTA of 128 threads transforms `syms_per_cta` OFDM symbols: coalesced load of each FFT window (rotated so the
// true symbol start comes first, dataGenerator.py:442), mixed-radix Stockham autosort FFT in shared memory --
// radix-16 passes with the 16-point transform in registers (FFT 256 = two passes, 1024 = 16.16.4), then radix-4 /
// radix-2 passes for what is left -- ping-pong buffers padded by one element per 16 so that every pass reads and
// writes conflict-free, twiddles from per-pass compact tables (host-computed FP64 -> FP32) staged in shared
// memory, then a gather of the data carriers written straight into the LS stage's layout Y[pkt][rx][sym][k].
// HBM-bound: (FFT+CP)*8 B in, Nsc*8 B out per symbol.
#pragma once
#include "ptx.cuh"

namespace mm {

struct OfdmArgs {
  const void* x;            // complex [n_pkt*n_rx][n_sym*(fft+cp)]  float2 or double2
  float2* Y;                // complex64 [n_pkt*n_rx][n_sym][n_sc]
  const float2* twiddle;    // per-pass compact tables: pass (radix R, ns): [r = 1..R-1][k < ns] = exp(-2 pi i r k / (R ns))
  int n_twiddle;            // entries in that table
  const int* bins;          // [n_sc] natural-order FFT bin of each kept carrier
  int fft_len, cp_len, sym_offset, n_sym, n_sc;
  int syms_per_cta;         // symbols transformed side by side in one CTA
  long long total_syms;     // n_pkt * n_rx * n_sym
  int x_double;
};

constexpr int kOfdmThreads = 128;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)
__host__ __device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }                  // 1 pad element per 16

// forward DFT-4 in place: (v0, v1, v2, v3) -> natural order
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 t0 = caddf(v0, v2), t1 = csubf(v0, v2), t2 = caddf(v1, v3), t3 = cmul_negi(csubf(v1, v3));
  v0 = caddf(t0, t2);
  v1 = caddf(t1, t3);
  v2 = csubf(t0, t2);
  v3 = csubf(t1, t3);
}

// forward DFT-16 in registers, natural order in and out: 4 x 4 Cooley-Tukey,
//   X[k1 + 4 k2] = sum_n2 W16^(n2 k1) W4^(n2 k2) ( sum_n1 x[4 n1 + n2] W4^(n1 k1) )
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508978f, h = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // v[4 k1 + n2] = inner[n2][k1]
  // twiddle W16^(n2 k1), n2, k1 in 1..3
  v[4 + 1] = cmulf(v[4 + 1], make_float2(c1, -s1));   // W^1
  v[4 + 2] = cmulf(v[4 + 2], make_float2(h, -h));     // W^2
  v[4 + 3] = cmulf(v[4 + 3], make_float2(s1, -c1));   // W^3
  v[8 + 1] = cmulf(v[8 + 1], make_float2(h, -h));     // W^2
  v[8 + 2] = cmul_negi(v[8 + 2]);                     // W^4 = -i
  v[8 + 3] = cmulf(v[8 + 3], make_float2(-h, -h));    // W^6
  v[12 + 1] = cmulf(v[12 + 1], make_float2(s1, -c1)); // W^3
  v[12 + 2] = cmulf(v[12 + 2], make_float2(-h, -h));  // W^6
  v[12 + 3] = cmulf(v[12 + 3], make_float2(-c1, s1)); // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4 k1 + k2]
  // result X[k1 + 4 k2] sits in v[4 k1 + k2]: transpose the 4 x 4 register tile
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const float2 t = v[4 * i + j];
      v[4 * i + j] = v[4 * j + i];
      v[4 * j + i] = t;
    }
}

__global__ void __launch_bounds__(kOfdmThreads) ofdm_demod_kernel(const OfdmArgs a) {
  extern __shared__ float2 sm_fft[];                 // [syms_per_cta][2][pad16(fft_len)] then the twiddle tables
  const int N = a.fft_len;
  const int S = a.syms_per_cta;
  const int NP = pad16(N);                           // padded length of one ping-pong half
  const long long sym0 = static_cast<long long>(blockIdx.x) * S;
  const int n_here = static_cast<int>(min(static_cast<long long>(S), a.total_syms - sym0));
  const int sym_len = N + a.cp_len;
  const int lgN = 31 - __clz(N);                      // N is a power of two: shifts, not divisions
  float2* tw = sm_fft + static_cast<size_t>(S) * 2 * NP;
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
    sm_fft[static_cast<size_t>(s) * 2 * NP + pad16(i)] = v;
  }
  __syncthreads();

  int cur = 0;                                        // which ping-pong half holds the data
  int ns = 1;
  int tw_off = 0;
  // ---- radix-16 passes: 16 points per thread in registers
  for (; ns * 16 <= N; tw_off += 15 * ns, ns <<= 4) {
    const int q = N >> 4;
    const float2* w = tw + tw_off;                    // w[(r-1) * ns + k]
    for (int idx = threadIdx.x; idx < n_here * q; idx += blockDim.x) {
      const int s = idx >> (lgN - 4), j = idx & (q - 1);
      const float2* in = sm_fft + static_cast<size_t>(s) * 2 * NP + cur * NP;
      float2* out = sm_fft + static_cast<size_t>(s) * 2 * NP + (cur ^ 1) * NP;
      const int k = j & (ns - 1);
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = in[pad16(j + r * q)];
      if (ns > 1) {
#pragma unroll
        for (int r = 1; r < 16; ++r) v[r] = cmulf(v[r], w[(r - 1) * ns + k]);
      }
      dft16(v);
      const int d = ((j - k) << 4) + k;
#pragma unroll
      for (int r = 0; r < 16; ++r) out[pad16(d + r * ns)] = v[r];
    }
    __syncthreads();
    cur ^= 1;
  }
  // ---- at most one radix-4 pass
  if (ns * 4 <= N) {
    const int q = N >> 2;
    const float2* w = tw + tw_off;
    for (int idx = threadIdx.x; idx < n_here * q; idx += blockDim.x) {
      const int s = idx >> (lgN - 2), j = idx & (q - 1);
      const float2* in = sm_fft + static_cast<size_t>(s) * 2 * NP + cur * NP;
      float2* out = sm_fft + static_cast<size_t>(s) * 2 * NP + (cur ^ 1) * NP;
      const int k = j & (ns - 1);
      float2 v0 = in[pad16(j)];
      float2 v1 = cmulf(in[pad16(j + q)], w[k]);
      float2 v2 = cmulf(in[pad16(j + 2 * q)], w[ns + k]);
      float2 v3 = cmulf(in[pad16(j + 3 * q)], w[2 * ns + k]);
      dft4(v0, v1, v2, v3);
      const int d = ((j - k) << 2) + k;
      out[pad16(d)] = v0;
      out[pad16(d + ns)] = v1;
      out[pad16(d + 2 * ns)] = v2;
      out[pad16(d + 3 * ns)] = v3;
    }
    __syncthreads();
    cur ^= 1;
    tw_off += 3 * ns;
    ns <<= 2;
  }
  // ---- radix-2 passes for what is left (at most one: log2 N mod 4 odd)
  for (; ns < N; tw_off += ns, ns <<= 1) {
    const int half = N >> 1;
    const float2* w = tw + tw_off;
    for (int idx = threadIdx.x; idx < n_here * half; idx += blockDim.x) {
      const int s = idx >> (lgN - 1), j = idx & (half - 1);
      const float2* in = sm_fft + static_cast<size_t>(s) * 2 * NP + cur * NP;
      float2* out = sm_fft + static_cast<size_t>(s) * 2 * NP + (cur ^ 1) * NP;
      const int k = j & (ns - 1);
      const float2 u = in[pad16(j)];
      const float2 v = cmulf(in[pad16(j + half)], w[k]);
      const int d = ((j - k) << 1) + k;
      out[pad16(d)] = caddf(u, v);
      out[pad16(d + ns)] = csubf(u, v);
    }
    __syncthreads();
    cur ^= 1;
  }
  // symbols are contiguous in Y too ([stream][sym][k]): global symbol g writes at g * n_sc
  for (int s = 0; s < n_here; ++s) {
    const float2* res = sm_fft + static_cast<size_t>(s) * 2 * NP + cur * NP;
    float2* out = a.Y + static_cast<size_t>(sym0 + s) * a.n_sc;
    for (int k = threadIdx.x; k < a.n_sc; k += blockDim.x) out[k] = res[pad16(__ldg(a.bins + k))];
  }
}

}  // namespace mm
