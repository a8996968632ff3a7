// Microbenchmarks that decide the solver's design on B200: FP64 FMA dependent-issue latency and per-SM throughput
// as a function of resident warps, shared-memory load latency, and the cost of a 64-thread __syncthreads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench/ubench.cu ; run: ./ubench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, long long* cyc, int n, int ilp) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  if (ilp == 1) for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); }
  else if (ilp == 2) for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); }
  else if (ilp == 4) for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
  else for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
                                     a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ffma_chain(float* out, long long* cyc, int n) {
  float a0 = threadIdx.x * 1e-9f;
  const float b = 1.0000001f, c = 1e-9f;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a0 = fmaf(a0, b, c);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void lds_chase(int* out, long long* cyc, int n) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) p = nxt[p];
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void bar_loop(int* out, long long* cyc, int n) {
  __shared__ double buf[64];
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    buf[threadIdx.x] = v;
    __syncthreads();
    v = buf[(threadIdx.x + 33) & 63] + 1.0;     // cross-warp hand-off: store -> barrier -> load -> (dependent) store
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (int)v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* d; long long* c; float* f; int* ii;
  cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 1 << 20); cudaMalloc(&f, 1 << 24); cudaMalloc(&ii, 1 << 24);
  long long h[148 * 32];
  const int n = 4096;
  printf("FP64 FMA: cycles per loop iteration (one FMA per chain), 64-thread CTAs, k CTAs per SM\n");
  for (int ilp : {1, 2, 4, 8})
    for (int k : {1, 2, 4, 6, 8, 12, 16}) {
      dfma_chain<<<148 * k, 64>>>(d, c, n, ilp); cudaDeviceSynchronize();
      dfma_chain<<<148 * k, 64>>>(d, c, n, ilp); cudaDeviceSynchronize();
      cudaMemcpy(h, c, sizeof(long long) * 148 * k, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < 148 * k; ++i) m += h[i]; m /= 148 * k;
      // DFMA warp-instructions per SM per cycle = ilp * (2 warps * k) / (cycles per iteration)
      printf("  ilp %d  CTAs/SM %2d (warps/SM %2d): %.2f cyc/iter  -> %.3f DFMA warp-instr/clk/SM\n", ilp, k, 2 * k, m / n,
             ilp * 2.0 * k / (m / n));
    }
  ffma_chain<<<148, 64>>>(f, c, n); cudaDeviceSynchronize();
  cudaMemcpy(h, c, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
  printf("FP32 FMA dependent chain: %.2f cyc\n", (double)h[0] / n);
  lds_chase<<<148, 64>>>(ii, c, n); cudaDeviceSynchronize();
  cudaMemcpy(h, c, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
  printf("LDS dependent chain: %.2f cyc\n", (double)h[0] / n);
  for (int k : {1, 4, 6}) {
    bar_loop<<<148 * k, 64>>>(ii, c, 1024); cudaDeviceSynchronize();
    cudaMemcpy(h, c, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
    printf("store + __syncthreads + load + add + __syncthreads (64 threads, %d CTAs/SM): %.1f cyc per round\n", k, (double)h[0] / 1024);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
