// loop_probe: the EXHAUSTIVE 4-colour candidate loop of bc_device.cuh (run_sweep<4, ..., PRUNE = false>) in isolation -- every warp sweeps a private,
// pre-built prefix table over and over, no fixed work, no tile staging -- at a chosen number of resident warps per
// SM sub-partition.  Reports SMSP-cycles per row of 32 candidates next to a calibration loop of 128 independent
// scalar FMULs per row (= 128 FMA-pipe cycles), so that the figure does not depend on knowing the clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o tools/_bin/loop_probe tools/loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tileconv_b200/csrc/bc_device.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
using namespace tcb;

__global__ void __launch_bounds__(256) sweep_kernel(const uint2* __restrict__ cand, int ncand, int n, int reps, float* out, int ncol)
{
  extern __shared__ __align__(16) uint8_t smem[];
  WarpScratch& sc = reinterpret_cast<WarpScratch*>(smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  // a plausible ordered point list: weights 1, colours spread over [0,1]
  if (lane < n) {
    const float t = (float)(lane + 1) / (float)(n + 1), w = 1.0f + 0.25f * (float)((lane * 7 + (threadIdx.x >> 5)) & 3);
    sc.pw[lane] = make_float4(t * w, (1.0f - t) * w, (0.3f + 0.5f * t) * w, w);
  }
  __syncwarp();
  if (lane <= n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* row = sc.prefix + rowoff(lane);
    row[0] = acc;
    for (int e = lane; e < n; ++e) { const float4 v = sc.pw[e]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; row[e - lane + 1] = acc; }
  }
  __syncwarp();
  Variant vp = { { 1.f, 1.f, 1.f }, 0, nullptr, 0 };
  float acc = 0.f; int accC = 0;
  for (int r = 0; r < reps; ++r) {
    SweepResult s;
    if (ncol == 4) s = run_sweep<4, false, true, false>(sc, nullptr, lane, n, cand, ncand, FLT_MAX, vp);
    else s = run_sweep<3, false, true, false>(sc, nullptr, lane, n, cand, ncand, FLT_MAX, vp);
    acc += s.err; accC += s.c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)accC;
}

__global__ void __launch_bounds__(256) calib_kernel(int rows, float* out, float seed)
{
  float a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = seed + 0.001f * (float)(threadIdx.x + c);
  const float m = 1.0f + 1e-7f;
  for (int it = 0; it < rows; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int c = 0; c < 16; ++c) a[c] = a[c] * m;      // 128 FMULs per row
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(void (*launch)(void*), void* arg)
{
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(arg); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 3; ++i) {
    CK(cudaEventRecord(e0)); launch(arg); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

struct Args { const uint2* cand; int ncand, n, reps, grid, ncol; size_t smem; float* out; };
static void launch_sweep(void* p) { Args* a = (Args*)p; sweep_kernel<<<a->grid, 256, a->smem>>>(a->cand, a->ncand, a->n, a->reps, a->out, a->ncol); }
static void launch_calib(void* p) { Args* a = (Args*)p; calib_kernel<<<a->grid, 256>>>(a->reps, a->out, 1.0f); }

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float)));
  for (int ncol = 4; ncol >= 3; --ncol)
  for (int n : { 16, 12, 8 }) {
    std::vector<uint2> list;
    if (ncol == 4) {
      for (int i = 0; i < n; ++i) for (int j = i; j <= n; ++j) for (int k = (j == 0 ? 1 : j); k <= n; ++k)
        list.push_back(make_candidate(rowoff(0) + i, rowoff(i) + (j - i), rowoff(j) + (k - j), rowoff(k) + (n - k)));
    } else {
      for (int i = 0; i < n; ++i) for (int j = (i == 0 ? 1 : i); j <= n; ++j)
        list.push_back(make_candidate(rowoff(0) + i, rowoff(i) + (j - i), rowoff(j) + (n - j), 0));
    }
    uint2* dCand; CK(cudaMalloc(&dCand, list.size() * sizeof(uint2))); CK(cudaMemcpy(dCand, list.data(), list.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    const int rounds = ((int)list.size() + 31) / 32;
    for (int ctas : { 1, 2, 3, 4, 6, 8 }) {          // resident CTAs per SM (8 warps each) = 2*ctas warps per SMSP
      Args a; a.cand = dCand; a.ncand = (int)list.size(); a.n = n; a.ncol = ncol; a.reps = (ncol == 4 ? 4000 : 20000) * 16 / n; a.grid = sms * ctas; a.out = out;
      a.smem = sizeof(WarpScratch) * 8;
      // limit residency with dynamic shared memory: 227 KB / ctas, minus a margin
      size_t want = (size_t)(200 * 1024) / ctas;
      if (want > a.smem) a.smem = want;
      CK(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem));
      const float ms = time_ms(launch_sweep, &a);
      Args c = a; c.reps = a.reps * rounds;      // same number of rows, 128 FMUL each
      const float msCal = time_ms(launch_calib, &c);
      printf("{\"ncol\": %d, \"n\": %d, \"candidates\": %d, \"rounds\": %d, \"warps_per_smsp\": %d, \"ms\": %.3f, \"calib_ms_128_fmul_rows\": %.3f, \"cycles_per_row_vs_128\": %.1f}\n",
             ncol, n, (int)list.size(), rounds, 2 * ctas, ms, msCal, 128.0 * ms / msCal);
      fflush(stdout);
    }
    CK(cudaFree(dCand));
  }
  return 0;
}
