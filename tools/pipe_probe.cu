// pipe_probe: measures per-SM issue throughput of the FP32-pipe instructions the BC fitters are
// built from (FMUL/FADD without contraction, FFMA, the sm_100 packed f32x2 forms, FMNMX, MUFU.RCP,
// IEEE reciprocal, truncation, LDS.128, SHFL).  Output: one line per op with warp-instructions per
// clock per SM and the derived lane-op rate, so DESIGN.md's ALU roofline uses measured numbers.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;            // independent chains per thread

enum Op { FMUL, FADD, FMULADD, FFMA, FMUL2, FADD2, FFMA2, FMNMX, RCP_MUFU, RCP_IEEE, TRUNC, LDS128, SHFL, MIXCAND, NOPS };
static const char* kNames[NOPS] = {"fmul", "fadd", "fmul+fadd", "ffma", "fmul.f32x2", "fadd.f32x2", "ffma.f32x2",
                                   "fmnmx", "mufu.rcp", "rcp.ieee", "cvt.rzi(trunc)", "lds.128", "shfl.idx", "fmul2+fadd2"};

template <int OP>
__global__ void __launch_bounds__(256) probe(float* out, long long* cycles, float seed)
{
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(seed, seed + 1.f, seed + 2.f, seed + 3.f);
  __syncthreads();
  float a[CH], b[CH];
  unsigned long long p[CH], q[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    a[c] = seed + 0.001f * (threadIdx.x + c);
    b[c] = 1.0f + 1e-7f * (c + 1);
    float lo = a[c], hi = a[c] + 0.5f;
    p[c] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
    q[c] = ((unsigned long long)__float_as_uint(b[c]) << 32) | __float_as_uint(b[c]);
  }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == FMUL)      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c]));
      if (OP == FADD)      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c]));
      if (OP == FMULADD) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c]));
                           asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c])); }
      if (OP == FFMA)      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[c]) : "f"(b[c]));
      if (OP == FMUL2)     asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c]));
      if (OP == FADD2)     asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c]));
      if (OP == FFMA2)     asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[c]) : "l"(q[c]));
      if (OP == MIXCAND) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c]));
                           asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q[c])); }
      if (OP == FMNMX)     asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c]));
      if (OP == RCP_MUFU)  asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
      if (OP == RCP_IEEE)  a[c] = __frcp_rn(a[c]);
      if (OP == TRUNC)     asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[c]));
      if (OP == LDS128)  { float4 v = sm[(threadIdx.x + __float_as_int(a[c])) & 255]; a[c] = __int_as_float(__float_as_int(v.x) & 0xff); b[c] += v.y + v.z + v.w; }
      if (OP == SHFL)      a[c] = __shfl_sync(0xffffffffu, a[c], (threadIdx.x + c) & 31);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += a[c] + b[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, int ctas_per_sm, float* d_out, long long* d_cyc)
{
  int grid = sms * ctas_per_sm;
  probe<OP><<<grid, 256>>>(d_out, d_cyc, 1.0f);   // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe<OP><<<grid, 256>>>(d_out, d_cyc, 1.0f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
  free(h);
  int per_iter = (OP == FMULADD || OP == MIXCAND) ? 2 : 1;
  double warp_instr_per_sm = (double)ctas_per_sm * 8 /*warps*/ * ITERS * CH * per_iter;
  double ipc = warp_instr_per_sm / avg;            // warp-instructions / clk / SM
  double lanes = (OP == FMUL2 || OP == FADD2 || OP == FFMA2 || OP == MIXCAND) ? 64.0 : 32.0;
  double total_lane_ops = warp_instr_per_sm * sms * lanes;
  printf("{\"op\": \"%s\", \"ctas_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.1f, "
         "\"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"eff_clock_mhz\": %.0f}\n",
         kNames[OP], ctas_per_sm, ipc, ipc * lanes, ms, total_lane_ops / (ms * 1e-3), avg / (ms * 1e-3) / 1e6);
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", prop.name, sms, prop.major, prop.minor);
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * sms * 8 * 256));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
  for (int cps : {2, 4, 8}) {
    run<FMUL>(sms, cps, d_out, d_cyc);
    run<FADD>(sms, cps, d_out, d_cyc);
    run<FMULADD>(sms, cps, d_out, d_cyc);
    run<FFMA>(sms, cps, d_out, d_cyc);
    run<FMUL2>(sms, cps, d_out, d_cyc);
    run<FADD2>(sms, cps, d_out, d_cyc);
    run<FFMA2>(sms, cps, d_out, d_cyc);
    run<MIXCAND>(sms, cps, d_out, d_cyc);
    run<FMNMX>(sms, cps, d_out, d_cyc);
    run<RCP_MUFU>(sms, cps, d_out, d_cyc);
    run<RCP_IEEE>(sms, cps, d_out, d_cyc);
    run<TRUNC>(sms, cps, d_out, d_cyc);
    run<LDS128>(sms, cps, d_out, d_cyc);
    run<SHFL>(sms, cps, d_out, d_cyc);
  }
  return 0;
}
