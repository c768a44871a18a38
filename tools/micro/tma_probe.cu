// Developer probe: which 4-D FP64 TMA tile loads the B200 accepts (box shape, data type, OOB coordinates).
// usage: tma_probe <variant>   (each variant in its own process: a fault poisons the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, double* out, int nelem, unsigned bytes, int x, int y, int z, int fence) {
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ unsigned long long bar;
  double* tile = reinterpret_cast<double*>(raw);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    if (fence & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (fence & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     s32(tile)),
                 "l"(&tmap), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z), "r"(0)
                 : "memory");
  }
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
  } while (!ok);
  for (int q = threadIdx.x; q < nelem; q += blockDim.x) out[q] = tile[q];
}

struct Pad { double x[40]; };  // 320 bytes in front of the tensor map, as StageArgs
__global__ void __launch_bounds__(256, 1) k_probe4(const __grid_constant__ Pad pad, const __grid_constant__ CUtensorMap tmap, double* out, int nelem,
                                                   unsigned bytes, int x, int y, int z, int stride_elems) {
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ unsigned long long other;
  __shared__ unsigned long long full[4];
  __shared__ unsigned long long empty;
  double* tile = reinterpret_cast<double*>(raw);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&other)), "r"(256) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty)), "r"(7) : "memory");
    for (int q = 0; q < 4; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[q])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const bool producer = (threadIdx.x >> 5) == 7 && (threadIdx.x & 31) == 0;
  if (producer) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[q])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                       s32(tile + q * stride_elems)),
                   "l"(&tmap), "r"(s32(&full[q])), "r"(x), "r"(y), "r"(z + q), "r"(0)
                   : "memory");
    }
  }
  for (int q = 0; q < 4; q++) {
    unsigned ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&full[q])), "r"(0) : "memory");
    } while (!ok);
  }
  if (blockIdx.x == 0)
    for (int q = threadIdx.x; q < nelem; q += blockDim.x) out[q] = tile[q] + pad.x[0];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const long sy = 64, ny = 30, nz = 24, nv = 5;
  const long sz = sy * ny, vs = sz * nz;
  std::vector<double> h(vs * nv);
  for (long q = 0; q < (long)h.size(); q++) h[q] = (double)q;
  double *d, *dout;
  cudaMalloc(&d, h.size() * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  unsigned cw = 36, rh = 11;
  CUtensorMapDataType dtp = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  int x = 14, y = 3, z = 2, fence = 3;
  if (variant == 1) dtp = CU_TENSOR_MAP_DATA_TYPE_UINT64;
  if (variant == 2) cw = 32;
  if (variant == 3) { y = -1; z = -1; }
  if (variant == 4) { x = 40; }   // box crosses the end of the row (OOB in x)
  if (variant == 5) fence = 0;
  if (variant == 6) { cw = 16; }
  if (variant == 7) { cw = 32; rh = 8; }
  if (variant == 8) { y = 25; }  // OOB rows at the high end
  if (variant == 13) { y = 25; x = 45; }
  if (variant == 14) { y = 25; x = 44; }
  if (variant == 15) { x = 45; }
  if (variant == 16) { x = 15; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)sy, (cuuint64_t)ny, (cuuint64_t)nz, (cuuint64_t)nv};
  const cuuint64_t strides[3] = {(cuuint64_t)sy * 8, (cuuint64_t)sz * 8, (cuuint64_t)vs * 8};
  const cuuint32_t box[4] = {cw, rh, 1u, (cuuint32_t)nv};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, dtp, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d: encode -> %d (box %u x %u x 1 x %ld, coords %d %d %d)\n", variant, (int)r, cw, rh, nv, x, y, z);
  if (r != CUDA_SUCCESS) return 1;
  const int nelem = cw * rh * nv;
  cudaMalloc(&dout, nelem * 8);
  const size_t smem = (size_t)nelem * 8 + 128;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (variant >= 10) {
    const int stride = (nelem * 8 + 127) / 128 * 128 / 8;
    const size_t smem4 = (size_t)4 * stride * 8 + (variant >= 12 ? 120000 : 0);
    cudaFuncSetAttribute(k_probe4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
    Pad pad; for (int q = 0; q < 40; q++) pad.x[q] = 0.0;
    k_probe4<<<variant >= 11 ? 300 : 1, 256, smem4>>>(pad, tm, dout, nelem, (unsigned)nelem * 8, x, y, z, stride);
  } else
  k_probe<<<1, 128, smem>>>(tm, dout, nelem, (unsigned)nelem * 8, x, y, z, fence);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  run -> %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<double> o(nelem);
  cudaMemcpy(o.data(), dout, nelem * 8, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (long v = 0; v < nv; v++)
    for (long rr = 0; rr < rh; rr++)
      for (long cc = 0; cc < cw; cc++) {
        const long gx = x + cc, gy = y + rr, gz = z;
        double want = 0.0;
        if (gx >= 0 && gx < sy && gy >= 0 && gy < ny && gz >= 0 && gz < nz) want = (double)(v * vs + gz * sz + gy * sy + gx);
        if (o[(v * rh + rr) * cw + cc] != want) bad++;
      }
  printf("  mismatches: %ld of %d\n", bad, nelem);
  return bad != 0;
}
