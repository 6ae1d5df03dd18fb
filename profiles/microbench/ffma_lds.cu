// Microbenchmarks that set the roofline denominators this project needs and that
// MEASURED_PEAKS.json does not carry:
//   1. FP32 FFMA peak (the RT path is FP32-pipe bound, not HBM or tensor bound)
//   2. shared-memory broadcast load rate (LDS.32/.64/.128, all lanes one address)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_lds ffma_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void ffma_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// three distinct register sources per FFMA (the shape of real code)
__global__ void ffma3_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  float c0 = a, c1 = b, c2 = a + b, c3 = a - b;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, c0, x4); x1 = fmaf(x1, c1, x5); x2 = fmaf(x2, c2, x6); x3 = fmaf(x3, c3, x7);
      x4 = fmaf(x4, c1, x0); x5 = fmaf(x5, c2, x1); x6 = fmaf(x6, c3, x2); x7 = fmaf(x7, c0, x3);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int WIDTH, bool UNIFORM>
__global__ void lds_kernel(float *out, int iters) {
  __shared__ float4 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float acc = 0.f;
  int idx = UNIFORM ? 0 : (threadIdx.x & 31);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = (idx + k * 32) & 1023;
      const unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[j]);
      if (WIDTH == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); acc += v; }
      else if (WIDTH == 8) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr)); acc += x + y; }
      else { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); acc += x + y + z + w; }
    }
    idx = (idx + 7) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 4096;
  for (int bps = 1; bps <= 2; ++bps)
    for (int th = 256; th <= 1024; th *= 2) {
      float ms = time_ms([&] { ffma_kernel<<<sms * bps, th>>>(out, iters, 1.0001f, 0.5f); });
      double fl = 2.0 * 128 * iters * (double)sms * bps * th;
      float ms3 = time_ms([&] { ffma3_kernel<<<sms * bps, th>>>(out, iters, 1.0001f, 0.5f); });
      printf("{\"bench\":\"ffma\",\"blocks_per_sm\":%d,\"threads\":%d,\"tflops_imm\":%.2f,\"tflops_3reg\":%.2f}\n", bps, th,
             fl / ms / 1e9, fl / ms3 / 1e9);
    }
  const int li = 2048;
  auto rep = [&](const char *name, float ms, int th) {
    double loads = 16.0 * li * (double)sms * 2 * th / 32;  // warp-level load instructions
    printf("{\"bench\":\"%s\",\"threads\":%d,\"warp_loads_per_clk_per_sm_at_1.9GHz\":%.3f,\"ms\":%.3f}\n", name, th,
           loads / sms / (ms * 1e-3) / 1.9e9, ms);
  };
  for (int th = 256; th <= 1024; th *= 2) {
    rep("lds32_uniform", time_ms([&] { lds_kernel<4, true><<<sms * 2, th>>>(out, li); }), th);
    rep("lds64_uniform", time_ms([&] { lds_kernel<8, true><<<sms * 2, th>>>(out, li); }), th);
    rep("lds128_uniform", time_ms([&] { lds_kernel<16, true><<<sms * 2, th>>>(out, li); }), th);
    rep("lds128_perlane", time_ms([&] { lds_kernel<16, false><<<sms * 2, th>>>(out, li); }), th);
  }
  return 0;
}
