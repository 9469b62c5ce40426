// Micro-benchmark: issue rates of the instruction classes of the trace kernels on one SM sub-partition, alone and mixed.
// 8 warps per sub-partition (1024 threads / SM, one block per SM), 8 independent chains per thread per class, so that neither
// latency nor occupancy limits: what is measured is the pipe / dispatch throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu && ./pipe_mix
// Output: warp instructions per cycle per sub-partition for every mix (1.0 = the issue port).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// NF FP64 FMAs, NL LOP3s, NW IMAD.WIDEs, NI IMADs (32-bit) per loop iteration and chain slot, interleaved
template <int NF, int NL, int NW, int NI>
__global__ void __launch_bounds__(1024, 1) mix(unsigned long long* out, int iters, double m, unsigned k1, unsigned k2) {
  double a[4];
  unsigned x[4], y[4], z[4];
  unsigned long long w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; x[i] = threadIdx.x * 2654435761u + i; y[i] = x[i] ^ k1; z[i] = y[i] + k2; w[i] = x[i]; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r < NF) a[i] = fma(a[i], m, 1e-9);
        if (r < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(k1), "r"(k2));
        if (r < NW) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)w[i]), "r"(0xD2511F53u));
        if (r < NI) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(k1), "r"(k2));
      }
    }
  }
  const long long t1 = clock64();
  unsigned long long acc = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc += (unsigned long long)__double_as_longlong(a[i]) + x[i] + w[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (unsigned long long)(t1 - t0);
}

template <int NF, int NL, int NW, int NI>
void run(const char* name, unsigned long long* d, int n_sm) {
  const int iters = 20000;
  mix<NF, NL, NW, NI><<<n_sm, 1024>>>(d, 100, 0.999999, 0x1234567u, 0x89abcdefu);
  CK(cudaDeviceSynchronize());
  mix<NF, NL, NW, NI><<<n_sm, 1024>>>(d, iters, 0.999999, 0x1234567u, 0x89abcdefu);
  CK(cudaDeviceSynchronize());
  unsigned long long cyc;
  CK(cudaMemcpy(&cyc, d + (size_t)n_sm * 1024, 8, cudaMemcpyDeviceToHost));
  const double per_it = (double)(NF + NL + NW + NI) * 4;           // instructions per thread and iteration
  const double warp_instr = per_it * iters * 8;                     // 8 warps per sub-partition
  printf("%-34s fp64 %d lop3 %d imad.wide %d imad %d : %.3f warp-instr/cycle/SMSP  (%.2f cycles per instruction)\n", name, NF, NL, NW, NI,
         warp_instr / (double)cyc, (double)cyc / warp_instr);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  unsigned long long* d;
  CK(cudaMalloc(&d, ((size_t)n_sm * 1024 + 1) * 8));
  run<4, 0, 0, 0>("DFMA only", d, n_sm);
  run<0, 4, 0, 0>("LOP3 only", d, n_sm);
  run<0, 0, 4, 0>("IMAD.WIDE only", d, n_sm);
  run<0, 0, 0, 4>("IMAD only", d, n_sm);
  run<2, 2, 0, 0>("DFMA + LOP3 1:1", d, n_sm);
  run<2, 0, 2, 0>("DFMA + IMAD.WIDE 1:1", d, n_sm);
  run<2, 0, 0, 2>("DFMA + IMAD 1:1", d, n_sm);
  run<0, 2, 2, 0>("LOP3 + IMAD.WIDE 1:1 (Philox)", d, n_sm);
  run<0, 2, 0, 2>("LOP3 + IMAD 1:1", d, n_sm);
  run<1, 1, 1, 1>("DFMA + LOP3 + IMAD.WIDE + IMAD", d, n_sm);
  run<2, 2, 2, 0>("DFMA + LOP3 + IMAD.WIDE", d, n_sm);
  run<1, 2, 0, 1>("DFMA + 2 LOP3 + IMAD", d, n_sm);
  run<1, 3, 0, 0>("DFMA + 3 LOP3", d, n_sm);
  run<3, 1, 0, 0>("3 DFMA + LOP3", d, n_sm);
  CK(cudaFree(d));
  return 0;
}
