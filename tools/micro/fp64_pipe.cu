// Developer microbenchmark: FP64 pipe latency / throughput on sm_100a and accuracy of
// the MUFU-seeded reciprocal / square-root sequences used in physics.cuh.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../pion_b200/csrc/fastmath.cuh"

template <int ILP>
__global__ void k_chain(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; q++) x[q] = threadIdx.x * 1e-3 + q;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < ILP; q++) x[q] = fma(x[q], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; q++) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / iters / ILP;
}

__global__ void k_acc(const double* x, int n, double* err) {
  double er = 0, es = 0, ers = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = x[i];
    double r = pion::fast_rcp(v), r0 = 1.0 / v;
    er = fmax(er, fabs(r - r0) / fabs(r0));
    double av = fabs(v);
    double s = pion::fast_sqrt(av), s0 = sqrt(av);
    es = fmax(es, fabs(s - s0) / s0);
    double q = pion::fast_rsqrt(av), q0 = 1.0 / sqrt(av);
    ers = fmax(ers, fabs(q - q0) / q0);
  }
  atomicMax((unsigned long long*)&err[0], (unsigned long long)__double_as_longlong(er));
  atomicMax((unsigned long long*)&err[1], (unsigned long long)__double_as_longlong(es));
  atomicMax((unsigned long long*)&err[2], (unsigned long long)__double_as_longlong(ers));
}

template <int ILP>
void run(int threads, int blocks, const char* label) {
  double* d;
  cudaMalloc(&d, sizeof(double) * threads * blocks);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<ILP><<<blocks, threads>>>(d, iters, 0.999, 1e-3);
  cudaEventRecord(e0);
  k_chain<ILP><<<blocks, threads>>>(d, iters, 0.999, 1e-3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  double total = (double)iters * ILP * threads * blocks;
  printf("%-28s ILP=%d threads=%d blocks=%d : %.2f cycles/DFMA/warp-chain, %.2f TFLOP/s (FMA=2)\n", label, ILP, threads, blocks, cyc,
         2 * total / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  run<1>(32, 1, "latency (1 warp)");
  run<2>(32, 1, "1 warp ILP2");
  run<4>(32, 1, "1 warp ILP4");
  run<8>(32, 1, "1 warp ILP8");
  run<1>(128, 148, "1 warp/SMSP ILP1");
  run<1>(256, 148, "2 warps/SMSP ILP1");
  run<2>(256, 148, "2 warps/SMSP ILP2");
  run<1>(512, 148, "4 warps/SMSP ILP1");
  run<2>(512, 148, "4 warps/SMSP ILP2");
  run<1>(1024, 148, "8 warps/SMSP ILP1");
  run<4>(1024, 148 * 2, "16 warps/SMSP ILP4");
  // accuracy
  const int n = 1 << 22;
  double* h = new double[n];
  unsigned long long s = 88172645463325252ULL;
  for (int i = 0; i < n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double m = 1.0 + (double)(s >> 11) / 9007199254740992.0;
    int e = (int)((s >> 3) % 600) - 300;
    h[i] = ldexp(m, e) * ((s & 1) ? 1 : -1);
  }
  double *dx, *derr;
  cudaMalloc(&dx, 8 * n); cudaMalloc(&derr, 24);
  cudaMemcpy(dx, h, 8 * n, cudaMemcpyHostToDevice);
  cudaMemset(derr, 0, 24);
  k_acc<<<148 * 4, 256>>>(dx, n, derr);
  double e[3];
  cudaMemcpy(e, derr, 24, cudaMemcpyDeviceToHost);
  printf("max rel err vs IEEE: rcp %.3e  sqrt %.3e  rsqrt %.3e  (eps = 1.11e-16)\n", e[0], e[1], e[2]);
  return 0;
}
