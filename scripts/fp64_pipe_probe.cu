// fp64_pipe_probe.cu -- does the FP64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA) have more throughput than the
// FP64 FMA pipe (DFMA) on this GPU?  That is the number behind the north star's "tensor cores only if ncu shows the
// small order-p contractions actually profit" for FP64 on sm_100a: the B / G contractions of order p are (p+2) x (p+1)
// matrices, padded to the 8 x 4 DMMA shape they use 6*5/(8*8) = 47 % (p=4) ... 8*7/(8*8) = 88 % (p=6) of the tensor
// instruction, so DMMA must be well above 1x DFMA per flop to pay.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/fp64_pipe_probe scripts/fp64_pipe_probe.cu
//   scripts/build/fp64_pipe_probe            (one JSON line)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
   double v[CHAINS];
#pragma unroll
   for (int i = 0; i < CHAINS; i++) { v[i] = threadIdx.x * 1e-3 + i; }
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) { v[i] = fma(v[i], a, b); }
   }
   double s = 0.0;
#pragma unroll
   for (int i = 0; i < CHAINS; i++) { s += v[i]; }
   if (s == 123.456) { out[threadIdx.x] = s; }
}

__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a0, double b0)
{
   double c[CHAINS][2];
#pragma unroll
   for (int i = 0; i < CHAINS; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
   const double a = a0 + threadIdx.x * 1e-9, b = b0;
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < CHAINS; i++)
      {
         asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                      : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
   }
   double s = 0.0;
#pragma unroll
   for (int i = 0; i < CHAINS; i++) { s += c[i][0] + c[i][1]; }
   if (s == 123.456) { out[threadIdx.x] = s; }
}

template <class F> float time_ms(F f)
{
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   f();                                     // warm-up
   cudaDeviceSynchronize();
   cudaEventRecord(e0);
   f();
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms = 0.f;
   cudaEventElapsedTime(&ms, e0, e1);
   return ms;
}

int main()
{
   cudaDeviceProp pr;
   cudaGetDeviceProperties(&pr, 0);
   int clk_khz = 0;
   cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
   double *out;
   cudaMalloc(&out, 4096);
   const int sms = pr.multiProcessorCount, blocks = sms * 8, threads = 256;
   const int it_f = 1 << 16, it_m = 1 << 14;
   const float ms_f = time_ms([&] { k_dfma<<<blocks, threads>>>(out, it_f, 1.0000001, 1e-9); });
   const float ms_m = time_ms([&] { k_dmma<<<blocks, threads>>>(out, it_m, 1.0000001, 1e-9); });
   const double flop_f = 2.0 * blocks * threads * (double)CHAINS * it_f;
   const double flop_m = 2.0 * 8 * 8 * 4 * (double)blocks * (threads / 32) * (double)CHAINS * it_m;
   const double tf_f = flop_f / ms_f * 1e-9, tf_m = flop_m / ms_m * 1e-9;
   const double ghz = clk_khz * 1e-6;
   printf("{\"gpu\": \"%s\", \"sms\": %d, \"boost_ghz\": %.3f, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
          "\"dfma_fma_per_sm_per_clk_at_boost\": %.1f, \"dmma_fma_per_sm_per_clk_at_boost\": %.1f, \"dmma_over_dfma\": %.3f, "
          "\"err\": \"%s\"}\n",
          pr.name, sms, ghz, tf_f, tf_m, tf_f * 1e3 / 2 / sms / ghz, tf_m * 1e3 / 2 / sms / ghz, tf_m / tf_f,
          cudaGetErrorString(cudaGetLastError()));
   return 0;
}
