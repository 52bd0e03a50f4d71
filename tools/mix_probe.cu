// mix_probe: issue cost of packed f32x2 instructions inside mixed instruction streams (B200).
// Each test runs a loop body of independent chains and reports SM-cycles per loop body per warp
// (from clock64 over a long loop with 8 resident warps per SMSP), next to the naive expectations.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 1 << 14;

template <int T>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed)
{
  float a[8]; unsigned long long p[8]; unsigned u[8];
  const float m = 1.0f + 1e-7f, b = 1e-9f;
  unsigned long long q = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
#pragma unroll
  for (int c = 0; c < 8; ++c) { a[c] = seed + c * 0.01f + threadIdx.x * 1e-3f; p[c] = ((unsigned long long)__float_as_uint(a[c]) << 32) | __float_as_uint(a[c] + 1.f); u[c] = threadIdx.x + c; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (T == 0) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); }                                   // 8 FMUL
      if (T == 1) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); }                                   // 8 FMUL2
      if (T == 2) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)); }          // FMUL2 + FADD
      if (T == 3) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); }  // FMUL2 + 2 FADD
      if (T == 4) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[c]) : "r"(it)); }           // FMUL2 + LOP3
      if (T == 5) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[c]) : "r"(it)); }              // FMUL + LOP3
      if (T == 6) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[c])); }                  // FMUL2 + FRND
      if (T == 7) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[c])); }  // 2 FMUL + FRND
      if (T == 8) { asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[c])); }                                          // 8 FRND
      if (T == 9) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[c]) : "r"(it)); }  // FMUL2+FMUL+LOP3
    }
  }
  long long t1 = clock64();
  float s = 0; unsigned uu = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) { s += a[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32)); uu ^= u[c]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)uu;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T> void run(const char* name, int instrPerChain, double expectPipe, int sms, float* d, long long* c)
{
  const int grid = sms * 4;   // 4 CTAs x 8 warps = 32 warps/SM = 8 per SMSP
  k<T><<<grid, 256>>>(d, c, 1.0f); CK(cudaDeviceSynchronize());
  k<T><<<grid, 256>>>(d, c, 1.0f); CK(cudaDeviceSynchronize());
  long long* h = (long long*)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, c, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid; free(h);
  // per SMSP: 8 warps each executing ITERS*8 chains of the body
  double cyclesPerBody = avg / ((double)ITERS * 8.0) / 8.0;     // SMSP-cycles per chain-body per warp
  printf("{\"test\": \"%s\", \"instr_per_body\": %d, \"expected_fma_pipe_cycles\": %.1f, \"measured_smsp_cycles_per_body\": %.3f}\n", name, instrPerChain, expectPipe, cyclesPerBody);
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  float* d; long long* c;
  CK(cudaMalloc(&d, sizeof(float) * prop.multiProcessorCount * 4 * 256)); CK(cudaMalloc(&c, sizeof(long long) * prop.multiProcessorCount * 4));
  int sms = prop.multiProcessorCount;
  run<0>("fmul", 1, 1, sms, d, c);
  run<1>("fmul2", 1, 2, sms, d, c);
  run<2>("fmul2+fadd", 2, 3, sms, d, c);
  run<3>("fmul2+fadd+fadd", 3, 4, sms, d, c);
  run<4>("fmul2+lop3", 2, 2, sms, d, c);
  run<5>("fmul+lop3", 2, 1, sms, d, c);
  run<6>("fmul2+frnd", 2, 2, sms, d, c);
  run<7>("fmul+fmul+frnd", 3, 2, sms, d, c);
  run<8>("frnd", 1, 0, sms, d, c);
  run<9>("fmul2+fmul+lop3", 3, 3, sms, d, c);
  return 0;
}
