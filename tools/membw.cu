// HBM direction microbenchmark for the roofline discussion in DESIGN.md: what does THIS GPU sustain for a read-only,
// a write-only and a copy kernel, at the sizes the tick's kernels move (90 MB head read, 157 MB letterbox write) and
// asymptotically, and do a read-only and a write-only kernel running side by side add up?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/scratch/membw tools/membw.cu && tools/scratch/membw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_read(const float4* __restrict__ src, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
  }
  for (; i < n4; i += stride) { float4 a = __ldcs(src + i); acc += a.x + a.y + a.z + a.w; }
  if (acc == 12345.678f) out[0] = acc;
}
__global__ void k_write(float4* __restrict__ dst, size_t n4, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 q = make_float4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(dst + i, q);
}
__global__ void k_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(dst + i, __ldcs(src + i));
}
__global__ void k_empty() {}

int main() {
  const size_t MAXB = (size_t)1 << 30;
  float4 *a[3], *b[3]; float* out;
  for (int s = 0; s < 3; ++s) { CK(cudaMalloc(&a[s], MAXB)); CK(cudaMalloc(&b[s], MAXB)); CK(cudaMemset(a[s], 0, MAXB)); CK(cudaMemset(b[s], 0, MAXB)); }
  CK(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
  const int sms = 148;
  auto timeit = [&](const char* name, double bytes, auto launch) {
    std::vector<float> ts;
    for (int it = 0; it < 12; ++it) {
      k_read<<<sms * 8, 256>>>(a[(it + 1) % 3], MAXB / 16, out);  // evicts the buffers of this iteration from L2, keeps the GPU busy
      cudaEventRecord(e0);
      launch(it % 3);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) ts.push_back(ms * 1e3f);
    }
    std::sort(ts.begin(), ts.end());
    printf("%-64s median %8.2f us  min %8.2f us -> %6.0f GB/s (median)\n", name, ts[ts.size() / 2], ts[0], bytes / (ts[ts.size() / 2] * 1e-6) / 1e9);
    fflush(stdout);
  };
  timeit("empty kernel (event to event)", 0, [&](int) { k_empty<<<1, 32>>>(); });
  char name[128];
  for (size_t mb : {90, 157, 360, 1024}) {
    const size_t n4 = mb * 1000000 / 16;
    for (int per : {4, 8, 16}) {
      snprintf(name, sizeof name, "read-only  %4zu MB, %2d CTAs/SM x 256 thr", mb, per);
      timeit(name, (double)n4 * 16, [&](int s) { k_read<<<sms * per, 256>>>(a[s], n4, out); });
    }
    snprintf(name, sizeof name, "write-only %4zu MB,  8 CTAs/SM x 256 thr", mb);
    timeit(name, (double)n4 * 16, [&](int s) { k_write<<<sms * 8, 256>>>(b[s], n4, 1.f); });
    snprintf(name, sizeof name, "copy       %4zu MB read + %4zu MB written", mb, mb);
    timeit(name, (double)n4 * 32, [&](int s) { k_copy<<<sms * 8, 256>>>(a[s], b[s], n4); });
  }
  // side by side: a 90 MB read-only kernel and a 157 MB write-only kernel on two streams (the tick's two HBM kernels are
  // 90 MB read-only and 66 MB read + 157 MB written)
  {
    const size_t nr = 90 * 1000000 / 16, nw = 157 * 1000000 / 16;
    timeit("read 90 MB || write 157 MB on two streams (total bytes)", (double)(nr + nw) * 16, [&](int s) {
      cudaEventRecord(f0, 0); cudaStreamWaitEvent(s1, f0, 0); cudaStreamWaitEvent(s2, f0, 0);
      k_read<<<sms * 4, 256, 0, s1>>>(a[s], nr, out);
      k_write<<<sms * 4, 256, 0, s2>>>(b[s], nw, 1.f);
      cudaEventRecord(f1, s1); cudaStreamWaitEvent(0, f1, 0); cudaEventRecord(f1, s2); cudaStreamWaitEvent(0, f1, 0);
    });
    timeit("read 90 MB then write 157 MB on one stream (total bytes)", (double)(nr + nw) * 16, [&](int s) {
      k_read<<<sms * 8, 256>>>(a[s], nr, out);
      k_write<<<sms * 8, 256>>>(b[s], nw, 1.f);
    });
    const size_t nr2 = 156 * 1000000 / 16;
    timeit("ONE kernel: 156 MB read + 157 MB written (copy-shaped)", (double)(nr2 + nw) * 16, [&](int s) { k_copy<<<sms * 8, 256>>>(a[s], b[s], nw); });
  }
  return 0;
}
