// peak_microbench.cu -- measured SM-side peaks of this pool's B200 for the rooflines bench.py reports (round-1 verdict,
// task 3): MEASURED_PEAKS.json (driver-written) holds only the HBM copy bandwidth and the cuBLAS bf16 rate; the MPPI
// kernels contain no contraction, their SM-side bound is the ISSUE SLOT (one warp instruction per SM sub-partition per
// clock) and, inside it, the FP32 pipe.  This program measures
//   ffma            scalar FFMA, 8 independent chains per thread      -> warp-inst/s == issue-slot peak, FP32 TFLOP/s
//   ffma2           packed fma.rn.f32x2 (sm_100: FFMA2)               -> flops per issue slot doubled?  (measured, not assumed)
//   ffma_iadd       FFMA interleaved with LOP3 (ALU pipe)             -> can two pipes issue in the same clock?
//   ffma2_lop       FFMA2 interleaved with LOP3                       -> does the packed form free issue slots for the other pipes?
//   dadd, f2f, mufu fp64 add, f32<->f64 conversion, MUFU.RSQ          -> the slow pipes the rollout touches
// and prints ONE JSON object.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o _build/peak_microbench
// peak_microbench.cu ; run on the GPU box; the result is committed as profiles/measured_sm_peaks.json.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x) do {cudaError_t e_ = (x); if (e_ != cudaSuccess) {fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);}} while (0)

constexpr int kThreads = 256;
constexpr int kChains = 8;
constexpr int kUnroll = 16;

__global__ void __launch_bounds__(kThreads) k_ffma(float * out, int iters, float a, float b)
{
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {x[c] = threadIdx.x * 1e-3f + c;}
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {x[c] = __fmaf_rn(x[c], a, b);}
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {s += x[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads) k_ffma2(float * out, int iters, float a, float b)
{
  unsigned long long x[kChains], aa, bb;
  {
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long *>(&t);
    bb = *reinterpret_cast<unsigned long long *>(&u);
  }
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    float2 t = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f + c);
    x[c] = *reinterpret_cast<unsigned long long *>(&t);
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {asm volatile ("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(aa), "l"(bb));}
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {float2 t = *reinterpret_cast<float2 *>(&x[c]); s += t.x + t.y;}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads) k_ffma_iadd(float * out, int iters, float a, float b, int k)
{
  float x[kChains];
  int n[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {x[c] = threadIdx.x * 1e-3f + c; n[c] = threadIdx.x + c;}
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        x[c] = __fmaf_rn(x[c], a, b);
        // three-input xor with a neighbouring chain: one LOP3 per FFMA that ptxas cannot fold away
        asm volatile ("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[c]) : "r"(k), "r"(n[(c + 1) % kChains]));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {s += x[c] + n[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

// FFMA2 + LOP3: does a packed FFMA (two pipe cycles) leave the issue slot of its second cycle to another pipe?
__global__ void __launch_bounds__(kThreads) k_ffma2_lop(float * out, int iters, float a, float b, int k)
{
  unsigned long long x[kChains], aa, bb;
  int n[kChains];
  {
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long *>(&t);
    bb = *reinterpret_cast<unsigned long long *>(&u);
  }
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    float2 t = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f + c);
    x[c] = *reinterpret_cast<unsigned long long *>(&t);
    n[c] = threadIdx.x + c;
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        asm volatile ("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(aa), "l"(bb));
        asm volatile ("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[c]) : "r"(k), "r"(n[(c + 1) % kChains]));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {float2 t = *reinterpret_cast<float2 *>(&x[c]); s += t.x + t.y + n[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads) k_dadd(float * out, int iters, double a)
{
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {x[c] = threadIdx.x * 1e-3 + c;}
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {x[c] = __dadd_rn(x[c], a);}
    }
  }
  double s = 0.;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {s += x[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = static_cast<float>(s);
}

// f32 -> f64 -> (+a) -> f32 : the pose add of the rollout (optimizer.cpp:339-342); counts 2 conversions + 1 DADD per element
__global__ void __launch_bounds__(kThreads) k_f2f(float * out, int iters, double a)
{
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {x[c] = threadIdx.x * 1e-3f + c;}
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {x[c] = static_cast<float>(__dadd_rn(static_cast<double>(x[c]), a));}
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {s += x[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads) k_mufu(float * out, int iters)
{
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {x[c] = threadIdx.x * 1e-3f + c + 1.0f;}
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {asm volatile ("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c]));}
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {s += x[c];}
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

template<typename F>
static double best_ms(F launch, int reps = 7)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, static_cast<double>(ms));
  }
  CK(cudaGetLastError());
  return best;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8;             // 2048 threads per SM: every sub-partition has 16 warps to pick from
  const int iters = 2000;
  float * out;
  CK(cudaMalloc(&out, sizeof(float) * blocks * kThreads));
  const double per_thread = static_cast<double>(iters) * kUnroll * kChains;
  const double warps = static_cast<double>(blocks) * kThreads / 32.0;

  const double t_ffma = best_ms([&] {k_ffma<<<blocks, kThreads>>>(out, iters, 0.999f, 1e-3f);});
  const double t_ffma2 = best_ms([&] {k_ffma2<<<blocks, kThreads>>>(out, iters, 0.999f, 1e-3f);});
  const double t_mix = best_ms([&] {k_ffma_iadd<<<blocks, kThreads>>>(out, iters, 0.999f, 1e-3f, 12345);});
  const double t_mix2 = best_ms([&] {k_ffma2_lop<<<blocks, kThreads>>>(out, iters, 0.999f, 1e-3f, 12345);});
  const double t_dadd = best_ms([&] {k_dadd<<<blocks, kThreads>>>(out, iters / 4, 1e-3);});
  const double t_f2f = best_ms([&] {k_f2f<<<blocks, kThreads>>>(out, iters / 16, 1e-3);});
  const double t_mufu = best_ms([&] {k_mufu<<<blocks, kThreads>>>(out, iters / 8);});

  const double wi_ffma = warps * per_thread / (t_ffma * 1e-3);
  const double wi_ffma2 = warps * per_thread / (t_ffma2 * 1e-3);
  const double wi_mix = warps * per_thread * 2.0 / (t_mix * 1e-3);
  const double wi_mix2 = warps * per_thread * 2.0 / (t_mix2 * 1e-3);
  const double wi_dadd = warps * per_thread / 4 / (t_dadd * 1e-3);
  const double el_f2f = warps * per_thread / 16 / (t_f2f * 1e-3);     // warp-level "pose adds" per second (3 instructions each)
  const double wi_mufu = warps * per_thread / 8 / (t_mufu * 1e-3);
  const double nominal = static_cast<double>(sms) * 4.0 * clock_khz * 1e3;
  printf("{\n");
  printf(" \"gpu_name\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d,\n", prop.name, sms, clock_khz);
  printf(" \"how\": \"scripts/peak_microbench.cu: %d blocks x %d threads, 8 independent chains per thread, best of 7 launches, CUDA events\",\n", blocks, kThreads);
  printf(" \"nominal_issue_slots_per_s\": %.6e,\n", nominal);
  printf(" \"ffma_warp_inst_per_s\": %.6e, \"fp32_ffma_tflops\": %.3f,\n", wi_ffma, wi_ffma * 64.0 / 1e12);
  printf(" \"ffma2_warp_inst_per_s\": %.6e, \"fp32_ffma2_tflops\": %.3f,\n", wi_ffma2, wi_ffma2 * 128.0 / 1e12);
  printf(" \"ffma_plus_int_warp_inst_per_s\": %.6e,\n", wi_mix);
  printf(" \"ffma2_plus_int_warp_inst_per_s\": %.6e, \"ffma2_plus_int_fp32_tflops\": %.3f,\n", wi_mix2, wi_mix2 / 2.0 * 128.0 / 1e12);
  printf(" \"dadd_warp_inst_per_s\": %.6e,\n", wi_dadd);
  printf(" \"f32_f64_add_f32_warp_ops_per_s\": %.6e,\n", el_f2f);
  printf(" \"mufu_rsq_warp_inst_per_s\": %.6e,\n", wi_mufu);
  printf(" \"warp_inst_per_s_peak\": %.6e,\n", std::max(wi_ffma, wi_mix));
  printf(" \"note\": \"warp_inst_per_s_peak is the issue-slot roofline bench.py uses: the larger of the pure-FFMA and the FFMA+integer rates\"\n");
  printf("}\n");
  CK(cudaFree(out));
  return 0;
}
