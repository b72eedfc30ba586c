// pcie_probe.cu — how long do the host<->device copies of one 64k-request step take on this box?  Diagnostic (profiles/ only).
// nvcc -O2 -o pcie_probe pcie_probe.cu ; prints one line per experiment.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  CK(cudaSetDevice(0));
  const size_t MAXB = 16u << 20;
  void *h_def, *h_wc, *d_a, *d_b;
  CK(cudaHostAlloc(&h_def, MAXB, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_wc, MAXB, cudaHostAllocWriteCombined));
  memset(h_def, 1, MAXB); memset(h_wc, 2, MAXB);
  CK(cudaMalloc(&d_a, MAXB)); CK(cudaMalloc(&d_b, MAXB));
  const int NS = 4;
  cudaStream_t s[NS];
  for (int i = 0; i < NS; i++) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
  const int reps = 400;
  const size_t sizes[] = {64u << 10, 512u << 10, 1u << 20, 2u << 20, 4u << 20, 8u << 20};
  for (int kind = 0; kind < 2; kind++) {
    char* h = (char*)(kind ? h_wc : h_def);
    for (size_t sz : sizes) {
      // H2D back-to-back on one stream
      CK(cudaDeviceSynchronize());
      double t0 = now_us();
      for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync(d_a, h, sz, cudaMemcpyHostToDevice, s[0]));
      double t_issue = now_us() - t0;
      CK(cudaDeviceSynchronize());
      double h2d = (now_us() - t0) / reps;
      // D2H
      t0 = now_us();
      for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync(h, d_b, sz, cudaMemcpyDeviceToHost, s[1]));
      CK(cudaDeviceSynchronize());
      double d2h = (now_us() - t0) / reps;
      // both directions at once
      t0 = now_us();
      for (int r = 0; r < reps; r++) {
        CK(cudaMemcpyAsync(d_a, h, sz, cudaMemcpyHostToDevice, s[0]));
        CK(cudaMemcpyAsync(h + MAXB / 2, d_b, sz, cudaMemcpyDeviceToHost, s[1]));
      }
      CK(cudaDeviceSynchronize());
      double both = (now_us() - t0) / reps;
      printf("%s size=%zuKiB h2d=%.1fus (%.1f GB/s, issue %.1fus) d2h=%.1fus (%.1f GB/s) both=%.1fus\n", kind ? "wc " : "def", sz >> 10, h2d,
             sz / h2d * 1e-3, t_issue / reps, d2h, sz / d2h * 1e-3, both);
    }
  }
  // 2 MiB H2D round-robin over k streams (do copies on different streams overlap?)
  for (int k = 1; k <= NS; k++) {
    CK(cudaDeviceSynchronize());
    double t0 = now_us();
    for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync((char*)d_a + (size_t)(r % k) * (2u << 20), (char*)h_def + (size_t)(r % k) * (2u << 20), 2u << 20, cudaMemcpyHostToDevice, s[r % k]));
    CK(cudaDeviceSynchronize());
    printf("h2d 2MiB round-robin over %d streams: %.1f us per copy\n", k, (now_us() - t0) / reps);
  }
  // 2 MiB split into 4 x 512 KiB on 4 streams
  {
    CK(cudaDeviceSynchronize());
    double t0 = now_us();
    for (int r = 0; r < reps; r++)
      for (int q = 0; q < 4; q++) CK(cudaMemcpyAsync((char*)d_a + (size_t)q * (512u << 10), (char*)h_def + (size_t)q * (512u << 10), 512u << 10, cudaMemcpyHostToDevice, s[q]));
    CK(cudaDeviceSynchronize());
    printf("h2d 2MiB as 4 x 512KiB on 4 streams: %.1f us per 2MiB\n", (now_us() - t0) / reps);
  }
  // isolated latency of one 2 MiB copy (sync after each)
  {
    double acc = 0;
    for (int r = 0; r < 50; r++) { double t0 = now_us(); CK(cudaMemcpyAsync(d_a, h_def, 2u << 20, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); acc += now_us() - t0; }
    printf("h2d 2MiB isolated (copy + sync): %.1f us\n", acc / 50);
    acc = 0;
    for (int r = 0; r < 50; r++) { double t0 = now_us(); CK(cudaMemcpyAsync(h_def, d_b, 2u << 20, cudaMemcpyDeviceToHost, s[0])); CK(cudaStreamSynchronize(s[0])); acc += now_us() - t0; }
    printf("d2h 2MiB isolated (copy + sync): %.1f us\n", acc / 50);
  }
  // a kernel reading mapped pinned memory directly (zero-copy gather of 2 MiB)
  return 0;
}
