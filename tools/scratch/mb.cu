#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, long long* out, int* sink, int n) {
  __shared__ int s[256];
  __shared__ double sdv[64];
  int lane = threadIdx.x & 31;
  s[threadIdx.x] = threadIdx.x; sdv[lane] = lane * 0.5;
  __syncthreads();
  long long t0 = clock64();
  int acc = 0; double dacc = 1.0;
  for (int i = 0; i < n; ++i) {
    if (mode == 0) { __syncwarp(); acc += i; }
    else if (mode == 1) { acc += __popc(__ballot_sync(0xffffffffu, (acc + lane) & 1)); }
    else if (mode == 2) { if (lane == 0) s[(acc + i) & 255] = i; __syncwarp(); acc += s[(i * 7) & 255]; }
    else if (mode == 3) { acc = s[acc & 255] + 1; }
    else if (mode == 4) { int v = s[(acc + lane) & 255]; unsigned b = __ballot_sync(0xffffffffu, v & 1); int m = __shfl_sync(0xffffffffu, v, __ffs(b | 1) - 1); if (lane == 0) { s[(m + i) & 255] = m; s[(m + 2 * i) & 255] += 1; } __syncwarp(); acc += m; }
    else if (mode == 5) { dacc = dacc * 1.0000001 + 0.5; }
    else if (mode == 6) { dacc = 1.0 / (dacc + 1.5); }
    else if (mode == 7) { if (threadIdx.x < 32) { if (lane == 0) s[(acc + i) & 255] = i; __syncwarp(); acc += s[(i * 7) & 255]; } }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[mode] = (t1 - t0);
  sink[threadIdx.x] = acc + (int)dacc;
}
int main() {
  long long* out; int* sink; cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 1024 * 4);
  long long h[16];
  for (int threads : {32, 256}) {
    for (int m = 0; m < 8; ++m) { k<<<1, threads>>>(m, out, sink, 1000); }
    cudaDeviceSynchronize(); cudaMemcpy(h, out, 8 * 8, cudaMemcpyDeviceToHost);
    printf("threads %d cycles/iter: syncwarp %.1f ballot %.1f lane0-sts+syncwarp+lds %.1f lds-chain %.1f lookup-mix %.1f dfma %.1f ddiv %.1f warp0only-mix %.1f\n", threads,
      h[0]/1000., h[1]/1000., h[2]/1000., h[3]/1000., h[4]/1000., h[5]/1000., h[6]/1000., h[7]/1000.);
  }
  return 0;
}
