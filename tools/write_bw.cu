// write_bw.cu -- measures the write-only / read-only / copy streaming bandwidth of the device.
// The metric kernel writes 5x more bytes than it reads, so the write-only number is the bound that matters for it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu && ./write_bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(double2* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += s) p[i] = make_double2(1.0, 2.0);
}
__global__ void k_write8(double* p, size_t n, int stride) {   // one 8-byte store per 32-byte sector when stride = 4
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  for (; i * stride < n; i += s) p[i * stride] = 1.0;
}
__global__ void k_read(const double2* p, size_t n, double* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  double a = 0;
  for (; i < n; i += s) { double2 v = p[i]; a += v.x + v.y; }
  if (a == 123.456) *out = a;
}
__global__ void k_copy(const double2* a, double2* b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += s) b[i] = a[i];
}
int main() {
  const size_t bytes = 4ull << 30;
  double2 *a, *b; double* o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 8);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t n = bytes / 16;
  auto run = [&](const char* name, auto f, double gb) {
    float best = 1e30f;
    for (int r = 0; r < 6; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms; }
    printf("%-28s %8.3f ms  %8.1f GB/s\n", name, best, gb / best * 1e3 / 1e9);
  };
  const int grid = 148 * 16;
  run("memset (driver)", [&] { cudaMemsetAsync(a, 1, bytes); }, (double)bytes);
  run("write st.v2.f64", [&] { k_write<<<grid, 512>>>(a, n); }, (double)bytes);
  run("read ld.v2.f64", [&] { k_read<<<grid, 512>>>(a, n, o); }, (double)bytes);
  run("copy (read+write bytes)", [&] { k_copy<<<grid, 512>>>(a, b, n); }, 2.0 * bytes);
  run("8B store per 32B sector", [&] { k_write8<<<grid, 512>>>((double*)a, bytes / 8, 4); }, (double)bytes / 4);
  run("8B store per 64B", [&] { k_write8<<<grid, 512>>>((double*)a, bytes / 8, 8); }, (double)bytes / 8);
  printf("(last two lines: useful bytes; sectors written per second = GB/s / 8e-9)\n");
  return 0;
}
