// mix_probe3: issue rate of packed f32x2 forms by operand kind (immediate / register / scalar broadcast).
// Cycles per instruction relative to scalar FMUL reg,imm (= 1.0), 8 warps per SMSP, independent chains.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 1 << 13;
typedef unsigned long long u64;

template <int T>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, const float* in)
{
  float a[8], r[8]; u64 p[8], q[8], z[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    a[c] = in[c] + threadIdx.x * 1e-3f; r[c] = in[8 + c];                       // runtime values: register operands
    p[c] = ((u64)__float_as_uint(a[c]) << 32) | __float_as_uint(a[c] + 1.f);
    q[c] = ((u64)__float_as_uint(in[16 + c]) << 32) | __float_as_uint(in[24 + c]);
    z[c] = ((u64)__float_as_uint(in[32 + c]) << 32) | __float_as_uint(in[40 + c]);
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (T == 0) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(a[c]));                         // FMUL r,imm
      if (T == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(r[c]));                      // FMUL r,r
      if (T == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(r[c]), "f"(r[(c + 1) & 7])); // FFMA r,r,r
      if (T == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(r[c]));                      // FADD r,r
      if (T == 4) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F800001, 0f3F800001}; mul.rn.f32x2 %0, %0, t;}" : "+l"(p[c]));   // FMUL2 r,imm
      if (T == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c]));                    // FMUL2 r,r
      if (T == 6) asm volatile("{.reg .b64 t; mov.b64 t, {%1, %1}; mul.rn.f32x2 %0, %0, t;}" : "+l"(p[c]) : "f"(r[c]));     // FMUL2 r, scalar broadcast
      if (T == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c]));                    // FADD2 r,r
      if (T == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(q[c]), "l"(z[c]));     // FFMA2 r,r,r
      if (T == 9) asm volatile("{.reg .b64 t; mov.b64 t, {0fBF800000, 0fBF800000}; fma.rn.f32x2 %0, %1, t, %0;}" : "+l"(p[c]) : "l"(q[c]));   // FFMA2 r,imm,r  (a - b)
      if (T == 10) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F800001, 0f3F800001}; fma.rn.f32x2 %0, %0, t, %1;}" : "+l"(p[c]) : "l"(q[c])); // FFMA2 acc*imm + r
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += a[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double base = 0;
template <int T> void run(const char* name, int sms, float* d, long long* c, const float* in)
{
  const int grid = sms * 4;
  k<T><<<grid, 256>>>(d, c, in); CK(cudaDeviceSynchronize());
  k<T><<<grid, 256>>>(d, c, in); CK(cudaDeviceSynchronize());
  long long* h = (long long*)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, c, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid; free(h);
  double perInstr = avg / ITERS / 8.0 / 8.0;
  if (T == 0) base = perInstr;
  printf("{\"form\": \"%s\", \"cycles_per_instr_rel\": %.3f}\n", name, perInstr / base);
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float* d; long long* c; float* in; float hin[48];
  for (int i = 0; i < 48; ++i) hin[i] = 1.0f + 1e-7f * (i % 5);
  CK(cudaMalloc(&d, sizeof(float) * sms * 4 * 256)); CK(cudaMalloc(&c, sizeof(long long) * sms * 4)); CK(cudaMalloc(&in, sizeof(hin)));
  CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
  run<0>("FMUL r,imm", sms, d, c, in);
  run<1>("FMUL r,r", sms, d, c, in);
  run<2>("FFMA r,r,r", sms, d, c, in);
  run<3>("FADD r,r", sms, d, c, in);
  run<4>("FMUL2 r,imm", sms, d, c, in);
  run<5>("FMUL2 r,r", sms, d, c, in);
  run<6>("FMUL2 r,scalar-bcast", sms, d, c, in);
  run<7>("FADD2 r,r", sms, d, c, in);
  run<8>("FFMA2 r,r,r", sms, d, c, in);
  run<9>("FFMA2 r,imm,r (a-b)", sms, d, c, in);
  run<10>("FFMA2 acc*imm+r", sms, d, c, in);
  return 0;
}
