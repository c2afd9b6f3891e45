// a14: DFL box decode of a raw YOLOv8 Detect head (named by the north star).
//
// NOT in the reference tree: the tensors the reference's _postprocess consumes are already decoded; the
// decode lives in the third-party Ultralytics `Detect` head (ultralytics==8.3.209, pylock.toml:1432-1433,
// reached only through YOLO(...).predict, detector.py:129,147-155).  Its published algorithm, restated:
//   raw [B, 4*reg_max + nc, A];  per box side: softmax over reg_max bins, expectation with weights 0..reg_max-1;
//   anchor points = cell centres (x + 0.5, y + 0.5) of the 80x80 / 40x40 / 20x20 grids (A = 8400 at 640x640);
//   x1y1 = anchor - (l, t), x2y2 = anchor + (r, b);  (cx, cy) = (x1y1 + x2y2) / 2, (w, h) = x2y2 - x1y1;
//   all four times the level's stride;  class scores = sigmoid(logits).
// Output [B, 4 + nc, A] is exactly what b200va_postprocess reads (channel major).  Floating point with
// exp: parity is tolerance based (1e-5 relative, oracle/dfl.py), "parity unpinned" by the reference.
// HBM bound: (4*reg_max + nc) * A * 4 bytes read + (4 + nc) * A * 4 written per frame (7.7 MB at nc = 80).
#include "common.cuh"

namespace {

constexpr int kMaxLevels = 8;

struct DflParams {
  const float* raw;
  float* out;
  int nc, reg_max, A, n_levels;
  int off[kMaxLevels + 1];  // first anchor of each level
  int w[kMaxLevels];        // grid width of each level
  float stride[kMaxLevels];
};

template <int RM>
__device__ __forceinline__ float side_expectation(const float* __restrict__ base, size_t A, int a, int reg_max) {
  // softmax expectation over the bins of one box side, bins RM (compile time) or reg_max (run time)
  const int n = RM > 0 ? RM : reg_max;
  float x[RM > 0 ? RM : 1];
  float m = -INFINITY;
  if (RM > 0) {
#pragma unroll
    for (int k = 0; k < RM; ++k) {
      x[k] = __ldg(base + (size_t)k * A + a);
      m = fmaxf(m, x[k]);
    }
    float s = 0.f, e = 0.f;
#pragma unroll
    for (int k = 0; k < RM; ++k) {
      const float ek = expf(x[k] - m);
      s += ek;
      e += (float)k * ek;
    }
    return e / s;
  }
  for (int k = 0; k < n; ++k) m = fmaxf(m, __ldg(base + (size_t)k * A + a));
  float s = 0.f, e = 0.f;
  for (int k = 0; k < n; ++k) {
    const float ek = expf(__ldg(base + (size_t)k * A + a) - m);
    s += ek;
    e += (float)k * ek;
  }
  return e / s;
}

template <int RM>
__global__ void __launch_bounds__(128) k_dfl_decode(const __grid_constant__ DflParams p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.A) return;
  const size_t A = (size_t)p.A;
  const int cin = 4 * p.reg_max + p.nc, cout = 4 + p.nc;
  const float* __restrict__ in = p.raw + (size_t)blockIdx.y * cin * A;
  float* __restrict__ out = p.out + (size_t)blockIdx.y * cout * A;
  int l = 0;
  while (l + 1 < p.n_levels && a >= p.off[l + 1]) ++l;
  const int idx = a - p.off[l];
  const float ax = (float)(idx % p.w[l]) + 0.5f, ay = (float)(idx / p.w[l]) + 0.5f;
  const float st = p.stride[l];
  float d[4];
#pragma unroll
  for (int sd = 0; sd < 4; ++sd) d[sd] = side_expectation<RM>(in + (size_t)sd * p.reg_max * A, A, a, p.reg_max);
  const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
  out[a] = (x1 + x2) * 0.5f * st;
  out[A + a] = (y1 + y2) * 0.5f * st;
  out[2 * A + a] = (x2 - x1) * st;
  out[3 * A + a] = (y2 - y1) * st;
  const float* __restrict__ cl = in + (size_t)4 * p.reg_max * A;
  int c = 0;
  for (; c + 8 <= p.nc; c += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(cl + (size_t)(c + u) * A + a);
#pragma unroll
    for (int u = 0; u < 8; ++u) out[(size_t)(4 + c + u) * A + a] = 1.f / (1.f + expf(-v[u]));
  }
  for (; c < p.nc; ++c) out[(size_t)(4 + c) * A + a] = 1.f / (1.f + expf(-__ldg(cl + (size_t)c * A + a)));
}

}  // namespace

extern "C" int b200va_dfl_decode(b200va_handle h, const float* raw, int batch, int num_classes, int reg_max,
                                 const int* level_hw, const float* level_stride, int n_levels, float* out,
                                 void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, raw && out && level_hw && level_stride, "NULL argument");
  REQUIRE(h, batch >= 0 && num_classes >= 0 && reg_max >= 1 && reg_max <= 64, "bad head geometry");
  REQUIRE(h, n_levels >= 1 && n_levels <= kMaxLevels, "n_levels must be in [1, %d]", kMaxLevels);
  if (batch == 0) return B200VA_OK;
  DflParams p;
  memset(&p, 0, sizeof(p));
  int total = 0;
  for (int l = 0; l < n_levels; ++l) {
    REQUIRE(h, level_hw[2 * l] > 0 && level_hw[2 * l + 1] > 0, "level %d has an empty grid", l);
    p.off[l] = total;
    p.w[l] = level_hw[2 * l + 1];
    p.stride[l] = level_stride[l];
    total += level_hw[2 * l] * level_hw[2 * l + 1];
  }
  p.off[n_levels] = total;
  p.raw = raw;
  p.out = out;
  p.nc = num_classes;
  p.reg_max = reg_max;
  p.A = total;
  p.n_levels = n_levels;
  PhaseScope phase(h, B200VA_PHASE_DFL, (cudaStream_t)stream);
  dim3 grid((total + 127) / 128, batch);
  if (reg_max == 16) k_dfl_decode<16><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  else k_dfl_decode<0><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  LAUNCH_CHECK(h);
  return B200VA_OK;
}
