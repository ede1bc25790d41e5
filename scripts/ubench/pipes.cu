// Micro-benchmark: per-SM-sub-partition throughput of MUFU.EX2, FFMA2, FADD2, F2FP, FMNMX for 1/2/4 warps per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (threadIdx.x + i) * 1e-3f;
  unsigned long long pa[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(pa[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1 && i < 8) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(pa[i]));
      if (OP == 2 && i < 8) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(pa[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]));
      if (OP == 5 && i < 8) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        a[2 * i] = __uint_as_float(r);
      }
      if (OP == 6) {   // mixed: MUFU + independent FFMA per element (can they co-issue?)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i < 8) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(pa[i]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)pa[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 8);
  const char* names[] = {"MUFU.EX2 x16", "FFMA2 x8", "FADD2 x8", "FFMA x16", "FMNMX x16", "F2FP x8", "MUFU x16 + FFMA2 x8"};
  const int nops[] = {16, 8, 8, 16, 16, 8, 16};
  const int iters = 2000;
  for (int op = 0; op < 7; ++op)
    for (int warps : {4, 8, 16, 32}) {   // warps per CTA, 1 CTA per SM -> warps/4 per SMSP
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: k<0><<<148, warps * 32>>>(out, cyc, iters); break;
          case 1: k<1><<<148, warps * 32>>>(out, cyc, iters); break;
          case 2: k<2><<<148, warps * 32>>>(out, cyc, iters); break;
          case 3: k<3><<<148, warps * 32>>>(out, cyc, iters); break;
          case 4: k<4><<<148, warps * 32>>>(out, cyc, iters); break;
          case 5: k<5><<<148, warps * 32>>>(out, cyc, iters); break;
          case 6: k<6><<<148, warps * 32>>>(out, cyc, iters); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double per_smsp_instr = (double)iters * nops[op] * (warps / 4);
      printf("%-22s warps/SMSP=%d  cycles per warp-instr per SMSP: %.2f  (per warp: %.2f)\n", names[op], warps / 4,
             h / per_smsp_instr, h / ((double)iters * nops[op]));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
