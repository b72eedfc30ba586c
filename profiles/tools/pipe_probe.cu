// pipe_probe.cu — does H2D(b+1) | kernels(b) | D2H(b-1) overlap on this box the way the 3-stream pipeline assumes?  Diagnostic.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void spin_kernel(long long cycles, unsigned* sink) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void pdl_spin_kernel(long long cycles, unsigned* sink) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
// touches memory like the real path: reads 2 MiB, writes 2 MiB
__global__ void touch_kernel(const uint4* in, uint4* out, size_t n16, long long cycles) {
  const long long t0 = clock64();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
  while (clock64() - t0 < cycles) {}
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  const int DEPTH = 4, STEPS = 2000;
  const size_t SZ = 2u << 20;
  char *h_in[DEPTH], *h_out[DEPTH], *d_in[DEPTH], *d_out[DEPTH];
  for (int k = 0; k < DEPTH; k++) {
    CK(cudaHostAlloc((void**)&h_in[k], SZ, cudaHostAllocDefault)); CK(cudaHostAlloc((void**)&h_out[k], SZ, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&d_in[k], SZ)); CK(cudaMalloc((void**)&d_out[k], SZ));
  }
  cudaStream_t s_h2d, s_c, s_d2h;
  CK(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_c, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
  cudaEvent_t in_done[DEPTH], c_done[DEPTH], out_done[DEPTH];
  for (int k = 0; k < DEPTH; k++) {
    CK(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c_done[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&out_done[k], cudaEventDisableTiming));
  }
  const long long cyc40 = 40LL * 1965;  // ~40 us at 1965 MHz
  struct Cfg { const char* name; int nk; bool pdl, graph, copies, small_h2d; int blocks; };
  const Cfg cfgs[] = {
      {"1 kernel x 40us", 1, false, false, true, false, 148},
      {"5 kernels x 8us, plain launches", 5, false, false, true, false, 256},
      {"5 kernels x 8us, programmatic dependent launch", 5, true, false, true, false, 256},
      {"5 kernels x 8us, one CUDA graph", 5, false, true, true, false, 256},
      {"5 kernels x 8us, plain, no copies (event chain only)", 5, false, false, false, false, 256},
      {"2 kernels x 20us, plain", 2, false, false, true, false, 256},
      {"3 kernels x 13us, plain", 3, false, false, true, false, 256},
      {"5 kernels x 8us, plain + extra 1 KiB H2D", 5, false, false, true, true, 256},
      {"5 kernels x 8us, PDL + extra 1 KiB H2D", 5, true, false, true, true, 256},
      {"5 kernels x 8us, graph + extra 1 KiB H2D", 5, false, true, true, true, 256},
  };
  for (const Cfg& c : cfgs) {
    cudaGraphExec_t gexec = nullptr;
    if (c.graph) {
      cudaGraph_t gr;
      CK(cudaStreamBeginCapture(s_c, cudaStreamCaptureModeThreadLocal));
      for (int q = 0; q < c.nk; q++) spin_kernel<<<c.blocks, 256, 0, s_c>>>(cyc40 / c.nk, nullptr);
      CK(cudaStreamEndCapture(s_c, &gr));
      CK(cudaGraphInstantiate(&gexec, gr, 0));
    }
    bool busy[DEPTH] = {false, false, false, false};
    CK(cudaDeviceSynchronize());
    double t0 = now_us(), t_wait = 0, t_issue = 0;
    for (int b = 0; b < STEPS; b++) {
      const int k = b % DEPTH;
      if (busy[k]) { const double a = now_us(); CK(cudaEventSynchronize(out_done[k])); t_wait += now_us() - a; }
      const double a = now_us();
      if (c.copies) CK(cudaMemcpyAsync(d_in[k], h_in[k], SZ, cudaMemcpyHostToDevice, s_h2d));
      if (c.small_h2d) CK(cudaMemcpyAsync(d_out[k], h_in[k], 1024, cudaMemcpyHostToDevice, s_h2d));
      CK(cudaEventRecord(in_done[k], s_h2d));
      CK(cudaStreamWaitEvent(s_c, in_done[k], 0));
      if (c.graph) CK(cudaGraphLaunch(gexec, s_c));
      else for (int q = 0; q < c.nk; q++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c.blocks); cfg.blockDim = dim3(256); cfg.stream = s_c;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = (c.pdl && q > 0) ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, pdl_spin_kernel, cyc40 / c.nk, (unsigned*)nullptr));
      }
      CK(cudaEventRecord(c_done[k], s_c));
      CK(cudaStreamWaitEvent(s_d2h, c_done[k], 0));
      if (c.copies) CK(cudaMemcpyAsync(h_out[k], d_out[k], SZ, cudaMemcpyDeviceToHost, s_d2h));
      CK(cudaEventRecord(out_done[k], s_d2h));
      busy[k] = true;
      t_issue += now_us() - a;
    }
    CK(cudaDeviceSynchronize());
    printf("%-55s %.1f us per step (host: %.1f us issuing, %.1f us waiting)\n", c.name, (now_us() - t0) / STEPS, t_issue / STEPS, t_wait / STEPS);
  }
  return 0;
}
