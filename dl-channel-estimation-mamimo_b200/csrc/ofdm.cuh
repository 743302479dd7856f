// OFDM demodulation front-end: the step immediately before the LS estimate.
//
// Replaces the toolbox call  rxOFDM = ofdmdemod(inputRXSig, FFT, CP, symOffset, nullIdx, pilotIdx)
//   (packet_generation/phased_arr/generate_maMIMO_LTF.m:336-338) and its numpy mirror
//   massiveMIMO_dataGenerator.py:437-453 (F-order reshape into symbols, CP removal with symbol offset,
//   FFT, fftshift along frequency, removal of null + pilot carriers).
//
// One CTA per (packet, rx, OFDM symbol): coalesced load of the FFT window (rotated so the true symbol
// start comes first, dataGenerator.py:442), radix-2 Stockham autosort FFT in shared memory (ping-pong
// buffers, twiddles from a host-computed FP64->FP32 table), then a gather of the data carriers written
// straight into the LS stage's layout Y[pkt][rx][sym][k].  HBM-bound: (FFT+CP)*8 B in, Nsc*8 B out per symbol.
#pragma once
#include "ptx.cuh"

namespace mm {

struct OfdmArgs {
  const void* x;            // complex [n_pkt*n_rx][n_sym*(fft+cp)]  float2 or double2
  float2* Y;                // complex64 [n_pkt*n_rx][n_sym][n_sc]
  const float2* twiddle;    // [fft/2]  exp(-2*pi*i*k/fft)
  const int* bins;          // [n_sc]   natural-order FFT bin of each kept carrier
  int fft_len, log2_fft, cp_len, sym_offset, n_sym, n_sc;
  int x_double;
};

__global__ void __launch_bounds__(256) ofdm_demod_kernel(const OfdmArgs a) {
  extern __shared__ float2 sm_fft[];                 // [2][fft_len]
  const int N = a.fft_len;
  float2* buf0 = sm_fft;
  float2* buf1 = sm_fft + N;
  const int sym = blockIdx.x % a.n_sym;
  const size_t prx = blockIdx.x / a.n_sym;
  const int sym_len = N + a.cp_len;
  const size_t base = (prx * a.n_sym + sym) * static_cast<size_t>(sym_len);
  // window[i] = x[ix(i)],  ix = [cp, fft+off) ++ [off, cp)   (dataGenerator.py:442)
  const int first = N + a.sym_offset - a.cp_len;     // length of the first run
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int src = (i < first) ? (a.cp_len + i) : (a.sym_offset + (i - first));
    float2 v;
    if (a.x_double) {
      const double2 d = __ldg(reinterpret_cast<const double2*>(a.x) + base + src);
      v = make_float2(static_cast<float>(d.x), static_cast<float>(d.y));
    } else {
      v = __ldg(reinterpret_cast<const float2*>(a.x) + base + src);
    }
    buf0[i] = v;
  }
  __syncthreads();
  // Stockham radix-2: Ns = 1, 2, ..., N/2; output in natural order
  const int half = N >> 1;
  int tw_stride = half;                              // N / (2*Ns)
  for (int ns = 1; ns < N; ns <<= 1, tw_stride >>= 1) {
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int k = j & (ns - 1);
      const float2 w = a.twiddle[k * tw_stride];
      const float2 u = buf0[j];
      const float2 t = buf0[j + half];
      const float2 v = make_float2(t.x * w.x - t.y * w.y, t.x * w.y + t.y * w.x);
      const int d = ((j - k) << 1) + k;              // (j / ns) * 2ns + k
      buf1[d] = make_float2(u.x + v.x, u.y + v.y);
      buf1[d + ns] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
    float2* t = buf0; buf0 = buf1; buf1 = t;
  }
  float2* out = a.Y + (prx * a.n_sym + sym) * static_cast<size_t>(a.n_sc);
  for (int k = threadIdx.x; k < a.n_sc; k += blockDim.x) out[k] = buf0[a.bins[k]];
}

}  // namespace mm
