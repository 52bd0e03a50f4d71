// mix_probe2: FMA-pipe cost of interleaving patterns of packed (P = mul.rn.f32x2) and scalar (S = mul.rn.f32)
// instructions, optional ALU filler (A = xor).  Reports cycles per pattern relative to pure S = 1 cycle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 1 << 15;
constexpr int REP = 8;
#define OPP(c) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q));
#define OPS(c) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m));
#define OPA(c) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[c]) : "r"(it));

template <int T>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed)
{
  float a[8]; unsigned long long p[8]; unsigned u[8];
  const float m = 1.0f + 1e-7f;
  unsigned long long q = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
#pragma unroll
  for (int c = 0; c < 8; ++c) { a[c] = seed + c * 0.01f + threadIdx.x * 1e-3f; p[c] = ((unsigned long long)__float_as_uint(a[c]) << 32) | __float_as_uint(a[c] + 1.f); u[c] = threadIdx.x + c; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
   for (int rep = 0; rep < REP; ++rep) {
    if (T == 0) { OPS(0) OPS(1) OPS(2) OPS(3) OPS(4) OPS(5) OPS(6) OPS(7) }
    if (T == 1) { OPP(0) OPP(1) OPP(2) OPP(3) OPP(4) OPP(5) OPP(6) OPP(7) }
    if (T == 2) { OPP(0) OPS(0) OPP(1) OPS(1) OPP(2) OPS(2) OPP(3) OPS(3) OPP(4) OPS(4) OPP(5) OPS(5) OPP(6) OPS(6) OPP(7) OPS(7) }          // PS x8
    if (T == 3) { OPP(0) OPP(1) OPS(0) OPS(1) OPP(2) OPP(3) OPS(2) OPS(3) OPP(4) OPP(5) OPS(4) OPS(5) OPP(6) OPP(7) OPS(6) OPS(7) }          // PPSS x4
    if (T == 4) { OPP(0) OPP(1) OPP(2) OPP(3) OPS(0) OPS(1) OPS(2) OPS(3) OPP(4) OPP(5) OPP(6) OPP(7) OPS(4) OPS(5) OPS(6) OPS(7) }          // PPPPSSSS x2
    if (T == 5) { OPP(0) OPP(1) OPP(2) OPP(3) OPP(4) OPP(5) OPP(6) OPP(7) OPS(0) OPS(1) OPS(2) OPS(3) OPS(4) OPS(5) OPS(6) OPS(7) }          // P8 S8
    if (T == 6) { OPP(0) OPS(0) OPS(1) OPP(1) OPS(2) OPS(3) OPP(2) OPS(4) OPS(5) OPP(3) OPS(6) OPS(7) }                                // PSS x4
    if (T == 7) { OPP(0) OPS(0) OPS(1) OPS(2) OPP(1) OPS(3) OPS(4) OPS(5) OPP(2) OPS(6) OPS(7) OPS(0) OPP(3) OPS(1) OPS(2) OPS(3) }          // PSSS x4
    if (T == 8) { OPP(0) OPA(0) OPS(0) OPP(1) OPA(1) OPS(1) OPP(2) OPA(2) OPS(2) OPP(3) OPA(3) OPS(3) OPP(4) OPA(4) OPS(4) OPP(5) OPA(5) OPS(5) OPP(6) OPA(6) OPS(6) OPP(7) OPA(7) OPS(7) }  // PAS x8
    if (T == 9) { OPP(0) OPP(1) OPS(0) OPP(2) OPP(3) OPS(1) OPP(4) OPP(5) OPS(2) OPP(6) OPP(7) OPS(3) }                                // PPS x4
    if (T == 10) { OPP(0) OPS(0) OPA(0) OPP(1) OPS(1) OPA(1) OPP(2) OPS(2) OPA(2) OPP(3) OPS(3) OPA(3) OPP(4) OPS(4) OPA(4) OPP(5) OPS(5) OPA(5) OPP(6) OPS(6) OPA(6) OPP(7) OPS(7) OPA(7) } // PSA x8
   }
  }
  long long t1 = clock64();
  float s = 0; unsigned uu = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) { s += a[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32)); uu ^= u[c]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)uu;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double base = 0;
template <int T> void run(const char* name, int nP, int nS, int nA, int sms, float* d, long long* c)
{
  const int grid = sms * 4;
  k<T><<<grid, 256>>>(d, c, 1.0f); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k<T><<<grid, 256>>>(d, c, 1.0f);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, c, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid; free(h);
  // absolute: SMSP cycles per pattern per warp at the sustained 1965 MHz SM clock (8 warps per SMSP)
  double cyclesEvent = (double)ms * 1e-3 * 1.965e9 / ((double)ITERS * REP * 8.0);
  double cyclesClock = avg / ((double)ITERS * REP * 8.0);
  (void)base;
  printf("{\"pattern\": \"%s\", \"P\": %d, \"S\": %d, \"A\": %d, \"issue_slots\": %d, \"pipe_cycles_if_P_is_2\": %d, \"cycles_by_event_at_1965MHz\": %.2f, \"cycles_by_clock64\": %.2f, \"ms\": %.2f}\n",
         name, nP, nS, nA, nP + nS + nA, 2 * nP + nS, cyclesEvent, cyclesClock, ms);
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  float* d; long long* c; int sms = prop.multiProcessorCount;
  CK(cudaMalloc(&d, sizeof(float) * sms * 4 * 256)); CK(cudaMalloc(&c, sizeof(long long) * sms * 4));
  run<0>("S x8", 0, 8, 0, sms, d, c);
  run<1>("P x8", 8, 0, 0, sms, d, c);
  run<2>("PS x8", 8, 8, 0, sms, d, c);
  run<3>("PPSS x4", 8, 8, 0, sms, d, c);
  run<4>("PPPPSSSS x2", 8, 8, 0, sms, d, c);
  run<5>("P8 S8", 8, 8, 0, sms, d, c);
  run<6>("PSS x4", 4, 8, 0, sms, d, c);
  run<7>("PSSS x4", 4, 12, 0, sms, d, c);
  run<8>("PAS x8", 8, 8, 8, sms, d, c);
  run<9>("PPS x4", 8, 4, 0, sms, d, c);
  run<10>("PSA x8", 8, 8, 8, sms, d, c);
  return 0;
}
