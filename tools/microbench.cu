// Instruction-throughput and streaming-read probes for sm_100a (development tool, not shipped).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu && ./tools/microbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(256) tput(float* out, int n_iter, float seedf, unsigned seedu) {
  float f[ILP]; unsigned u[ILP]; float2 p[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { f[i] = seedf + i + threadIdx.x; u[i] = seedu * (i + 1) + threadIdx.x; p[i] = make_float2(f[i], f[i] + 1); }
  const float a = seedf * 0.5f, b = seedf * 0.25f;
  const unsigned w = seedu | 0x00010001u;
  for (int it = 0; it < n_iter; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) f[i] = fmaf(f[i], a, b);                                   // FFMA
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&p[i]) : "l"(*(unsigned long long*)&p[(i+1)%ILP]), "l"(*(unsigned long long*)&p[(i+2)%ILP]));  // FFMA2
      if (OP == 2) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(w), "r"(u[(i+1)%ILP]));
      if (OP == 3) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(w), "r"(u[(i+1)%ILP]));
      if (OP == 4) u[i] = __byte_perm(u[i], w, 0x5140 + (it & 1));            // PRMT
      if (OP == 5) f[i] = (float)(int)u[i] + f[i];                            // I2F (+FADD)
      if (OP == 6) f[i] = f[i] + a;                                           // FADD
      if (OP == 7) u[i] = u[i] * w + seedu;                                   // IMAD
      if (OP == 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(*(unsigned long long*)&p[i]) : "l"(*(unsigned long long*)&p[(i+1)%ILP]));
      if (OP == 9) { f[i] = fmaf(f[i], a, b); u[i] = __byte_perm(u[i], w, 0x5140 + (it & 1)); }   // FFMA + PRMT dual issue?
      if (OP == 10) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(w), "r"(u[(i+1)%ILP])); f[i] = fmaf(f[i], a, b); } // IDP + FFMA
      if (OP == 11) f[i] = __uint_as_float(__byte_perm(u[i], 0x4B000000u, 0x7650)) - 8388608.0f + f[i];   // PRMT+FADD+FADD cvt
      if (OP == 12) f[i] = (float)(u[i] & 0xff) + f[i];                       // compiler's u8->f32
    }
  }
  float s = 0; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { s += f[i] + p[i].x + p[i].y; t += u[i]; }
  if (s == 123.456f || t == 0x12345u) out[threadIdx.x] = s + t;
}

// streaming read: every thread loads 16B chunks, grid-stride, sums bytes (cheap) -> measures achievable read BW
template <int UNROLL>
__global__ void __launch_bounds__(256) stream_read(const uint4* __restrict__ src, size_t n16, unsigned* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) v[k] = __ldg(src + i + k * stride);
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  }
  if (acc == 0x1234567u) out[0] = acc;
}

template <int OP>
static int run(const char* name, int ops_per_iter_per_thread) {
  float* out; CK(cudaMalloc(&out, 4096));
  const int blocks = 148 * 8, threads = 256;
  tput<OP><<<blocks, threads>>>(out, 16, 1.0001f, 3u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  tput<OP><<<blocks, threads>>>(out, ITERS, 1.0001f, 3u);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter_per_thread;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-op/clk/SM at %d MHz nominal)\n", name, ms, ops / ms * 1e-9,
         ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
  cudaFree(out);
  return 0;
}

int main() {
  run<0>("FFMA", 1); run<1>("FFMA2 (x2 lanes)", 2); run<2>("DP2A", 1); run<3>("DP4A", 1); run<4>("PRMT", 1);
  run<5>("I2F+FADD", 1); run<6>("FADD", 1); run<7>("IMAD", 1); run<8>("FADD2 (x2 lanes)", 2);
  run<9>("FFMA+PRMT pair", 2); run<10>("DP2A+FFMA pair", 2); run<11>("PRMT+FADD+FADD cvt", 1); run<12>("(float)(u&0xff)+FADD", 1);
  // streaming read
  const size_t bytes = (size_t)2 << 30;
  uint4* src; unsigned* o; CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&o, 64)); CK(cudaMemset(src, 1, bytes));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int blocks = 148 * (cfg == 0 ? 4 : cfg == 1 ? 8 : 16);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      stream_read<8><<<blocks, 256>>>(src, bytes / 16, o);
      cudaEventRecord(b); CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 2) printf("stream_read unroll8 blocks=%d: %.3f ms  %.1f GB/s\n", blocks, ms, bytes / ms * 1e-6);
    }
  }
  return 0;
}
