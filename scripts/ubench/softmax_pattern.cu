// Micro-benchmark of the attention softmax inner pattern: per pair FFMA2 -> 2 x MUFU.EX2 -> FADD2 (running sum) +
// F2FP (pack), with the consumers DIST pairs behind their producers.  Reports cycles per MUFU per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int DIST, bool WITH_FFMA2, bool WITH_MAX>
__global__ void k(float* out, long long* cyc, int iters, float c, float m) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (threadIdx.x * 32 + i) * 1e-4f;
  unsigned long long acc0 = 0, acc1 = 0;
  uint32_t pk[16];
  float mx = -1e30f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float e[32];
    // producers and consumers of one 32-element chunk, consumers DIST pairs behind
#pragma unroll
    for (int p = 0; p < 16 + DIST; ++p) {
      if (p < 16) {
        float x0 = v[2 * p], x1 = v[2 * p + 1];
        if (WITH_MAX) mx = fmaxf(mx, fmaxf(x0, x1));
        if (WITH_FFMA2) {
          unsigned long long xx, cc, mm;
          asm("mov.b64 %0, {%1,%2};" : "=l"(xx) : "f"(x0), "f"(x1));
          asm("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(c), "f"(c));
          asm("mov.b64 %0, {%1,%2};" : "=l"(mm) : "f"(m), "f"(m));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx) : "l"(cc), "l"(mm));
          asm("mov.b64 {%0,%1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
        }
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[2 * p]) : "f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[2 * p + 1]) : "f"(x1));
      }
      if (p >= DIST) {
        const int q = p - DIST;
        unsigned long long pp;
        asm("mov.b64 %0, {%1,%2};" : "=l"(pp) : "f"(e[2 * q]), "f"(e[2 * q + 1]));
        if (q & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc1) : "l"(pp));
        else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc0) : "l"(pp));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[q]) : "f"(e[2 * q + 1]), "f"(e[2 * q]));
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(pk[i & 15]) * 1e-9f;
  }
  long long t1 = clock64();
  float s = mx + (float)acc0 + (float)acc1;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float(pk[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int DIST, bool F, bool M>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      k<DIST, F, M><<<148, warps * 32>>>(out, cyc, iters, 1.3f, -2.f);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-34s warps/SMSP=%d  cycles per MUFU per SMSP: %.2f\n", name, warps / 4, h / ((double)iters * 32 * (warps / 4)));
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 8);
  run<1, false, false>("dist 1", out, cyc);
  run<2, false, false>("dist 2", out, cyc);
  run<4, false, false>("dist 4", out, cyc);
  run<8, false, false>("dist 8", out, cyc);
  run<2, true, false>("dist 2 + ffma2", out, cyc);
  run<4, true, true>("dist 4 + ffma2 + max", out, cyc);
  run<8, true, true>("dist 8 + ffma2 + max", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
