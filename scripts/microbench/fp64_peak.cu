// fp64_peak.cu -- measured FP64 denominators for the roofline of the rollout kernels (SURVEY.md 8(d):
// "not in MEASURED_PEAKS.json - measure it"): DFMA and DMMA (mma.sync.m8n8k4.f64) throughput with every SM
// busy, and the dependent-issue latencies that bound a single warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
// prints one JSON line.
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) c0[c] = c1[c] = threadIdx.x + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) dmma884(c0[c], c1[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += c0[c] + c1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one warp, one dependent chain: cycles per instruction = latency
__global__ void latency_kernel(double *out, long long *cycles, int iters, double a, double b) {
  __shared__ double sh[64];
  sh[threadIdx.x] = threadIdx.x + 1;
  sh[threadIdx.x + 32] = 0;
  __syncwarp();
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = fma(x, a, b);
  long long t1 = clock64();
  double c0 = x, c1 = x;
  for (int i = 0; i < iters; i++) dmma884(c0, c1, a, b);
  long long t2 = clock64();
  // dependent shared-memory loads: pointer chase
  int idx = threadIdx.x;
  double v = 0;
  for (int i = 0; i < iters; i++) {
    v = sh[idx];
    idx = ((int)v) & 31;
  }
  long long t3 = clock64();
  double y = x;
  for (int i = 0; i < iters; i++) y = y * a + x / (y + 3.0);  // one division per trip
  long long t4 = clock64();
  double z = y;
  for (int i = 0; i < iters; i++) z = sqrt(z * z + 1.0);
  long long t5 = clock64();
  double w = z;
  for (int i = 0; i < iters; i++) w = tanh(w * 1e-3) + 0.5;
  long long t6 = clock64();
  double sn = w, cs = w;
  for (int i = 0; i < iters; i++) sincos(sn * 1e-3 + 0.3, &sn, &cs);
  long long t7 = clock64();
  out[threadIdx.x] = c0 + c1 + v + y + z + w + sn + cs;
  if (threadIdx.x == 0) {
    cycles[0] = t1 - t0;
    cycles[1] = t2 - t1;
    cycles[2] = t3 - t2;
    cycles[3] = t4 - t3;
    cycles[4] = t5 - t4;
    cycles[5] = t6 - t5;
    cycles[6] = t7 - t6;
  }
}

template <class F> float timeIt(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  long long *cycles;
  cudaMalloc(&cycles, 64);
  const int iters = 4096;
  // throughput: 8 CTAs x 256 threads per SM, 8 independent chains per thread
  const int blocks = sms * 8, threads = 256;
  float ms_fma = timeIt([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
  double fma_flops = 2.0 * blocks * threads * 8.0 * iters;
  float ms_mma = timeIt([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
  double mma_flops = 2.0 * 256.0 * (blocks * threads / 32) * 8.0 * iters;
  // one CTA of 8 warps per SM, one chain per warp: what the rollout kernel's occupancy can issue
  float ms_fma_8w = timeIt([&] { dfma_kernel<1><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); }, 10);
  double fma_flops_8w = 2.0 * sms * 256 * 1.0 * iters;
  latency_kernel<<<1, 32>>>(out, cycles, 1024, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h[7];
  cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.3f, \"fp64_dmma_tflops\": %.3f, "
              "\"dfma_tflops_8_warps_per_sm_1_chain\": %.3f, \"latency_cycles\": {\"dfma\": %.1f, \"dmma_m8n8k4\": %.1f, "
              "\"lds_dependent\": %.1f, \"fma_plus_div\": %.1f, \"sqrt_fma\": %.1f, \"tanh\": %.1f, \"sincos\": %.1f}, "
              "\"error\": \"%s\"}\n",
              prop.name, sms, fma_flops / ms_fma * 1e-9, mma_flops / ms_mma * 1e-9, fma_flops_8w / ms_fma_8w * 1e-9,
              h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0, h[6] / 1024.0,
              cudaGetErrorString(e));
  return 0;
}
