// Microbenchmark (2 GPUs, one process): which store shape fills NVLink when an SM writes output tiles into a PEER's
// row-major float plane [rows][1024]?  Decides the tile shape of the fused all-gather epilogue (fc.cuh).
//   st.v4      : plain 16-byte stores, a warp covers 512 contiguous bytes of one row
//   tma WxR    : one cp.async.bulk.tensor.2d store of a [R rows x W floats] box staged in shared memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_store_bw peer_store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kCols = 1024;

__global__ void st_v4_kernel(float4* dst, size_t n_vec, int iters) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int it = 0; it < iters; ++it)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// every CTA walks boxes of the plane; `bufs` staging tiles in flight (bulk groups), like the FC epilogue's ping-pong
template <int W, int R>
__global__ void tma_store_kernel(const __grid_constant__ CUtensorMap map, int n_rows, int iters, int bufs) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kTile = W * R * 4;
  for (int i = threadIdx.x; i < bufs * kTile / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int boxes_x = kCols / W, boxes_y = n_rows / R;
  const long long total = (long long)boxes_x * boxes_y;
  int b = 0;
  for (int it = 0; it < iters; ++it)
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      const int bx = (int)(t % boxes_x), by = (int)(t / boxes_x);
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(reinterpret_cast<unsigned long long>(&map)), "r"(smem_u32(smem + b * kTile)), "r"(bx * W), "r"(by * R) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      b = (b + 1) % bufs;
      if (bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else if (bufs == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int R>
void run_tma(PFN_encodeTiled enc, float* dst, int n_rows, int ctas, int bufs, const char* where) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {kCols, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {kCols * 4};
  const cuuint32_t box[2] = {W, R};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   W == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d for %dx%d\n", (int)r, W, R); return; }
  const int smem = bufs * W * R * 4 + 1024;
  CK(cudaFuncSetAttribute(tma_store_kernel<W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int iters = 4;
  tma_store_kernel<W, R><<<ctas, 128, smem>>>(map, n_rows, 1, bufs);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  tma_store_kernel<W, R><<<ctas, 128, smem>>>(map, n_rows, iters, bufs);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("%-6s tma %3dx%-3d (%4d B rows) ctas %3d bufs %d : %7.1f GB/s\n", where, W, R, W * 4, ctas, bufs,
         (double)iters * n_rows * kCols * 4 / ms / 1e6);
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  const int n_rows = 65536;                                  // 256 MB plane
  const size_t bytes = (size_t)n_rows * kCols * 4;
  float *local = nullptr, *peer = nullptr;
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&local, bytes));
  if (nd >= 2) {
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&peer, bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
  }
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(p);
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? local : peer;
    const char* where = pass == 0 ? "local" : "peer";
    if (!dst) continue;
    for (int ctas : {56, 148}) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      st_v4_kernel<<<ctas * 4, 256>>>(reinterpret_cast<float4*>(dst), bytes / 16, 1);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      st_v4_kernel<<<ctas * 4, 256>>>(reinterpret_cast<float4*>(dst), bytes / 16, 4);
      CK(cudaEventRecord(b));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("%-6s st.v4 (512 B per warp)      ctas %3d        : %7.1f GB/s\n", where, ctas, 4.0 * bytes / ms / 1e6);
      for (int bufs : {2, 4}) {
        run_tma<32, 128>(enc, dst, n_rows, ctas, bufs, where);
        run_tma<64, 64>(enc, dst, n_rows, ctas, bufs, where);
        run_tma<128, 32>(enc, dst, n_rows, ctas, bufs, where);
        run_tma<256, 16>(enc, dst, n_rows, ctas, bufs, where);
      }
    }
    if (dst == peer) {                                       // the copy engine, for reference
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      CK(cudaMemcpyPeer(peer, 1, local, 0, bytes));
      CK(cudaEventRecord(a));
      for (int i = 0; i < 4; ++i) CK(cudaMemcpyPeerAsync(peer, 1, local, 0, bytes, 0));
      CK(cudaEventRecord(b));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("peer   cudaMemcpyPeer                               : %7.1f GB/s\n", 4.0 * bytes / ms / 1e6);
    }
  }
  return 0;
}
