// a9 / a11: ROI polygon masks and the motion gate.
//
// Replaces utils.apply_roi (frame_filter.py:43-50: cv2.fillPoly per polygon + bitwise_and) and
// MotionFilter.should_process (frame_filter.py:26-40: BGR2GRAY -> GaussianBlur 5x5 -> absdiff ->
// threshold 25 -> count).  OpenCV's integer arithmetic is restated (oracle/cv_restate.py holds the
// CPU statement the tests compare against):
//   fillPoly   = 8-connected Bresenham boundary  U  even-odd scanline fill on 16.16 edge
//                crossings.  Both are evaluated in closed form per pixel / per line step, so the
//                rasteriser is fully parallel.
//   BGR2GRAY   = (B*3735 + G*19235 + R*9798 + 16384) >> 15
//   Gaussian   = separable [1 4 6 4 1], BORDER_REFLECT_101, one rounding (sum + 128) >> 8
//   motion     = count(|new - prev| > 25); the new blurred gray always replaces the state.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// ROI rasteriser
// ------------------------------------------------------------------------------------------
struct RoiEdge {  // scanline edge, active for y in [y0, y1)
  long long x, dx;  // 16.16 fixed point
  int y0, y1;
  int poly, pad_;
};
struct RoiLine {  // clipped boundary segment, left-to-right Bresenham
  int x1, y1, major, minor, sy, vert;
};

// cv::clipLine on a width x height image (OpenCV drawing.cpp, restated): double arithmetic,
// truncation toward zero.  Returns false when the segment misses the image.
static bool clip_line(long long width, long long height, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = width - 1, bottom = height - 1;
  if (width <= 0 || height <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

static long long trunc_div(long long a, long long b) { return a / b; }  // C++ division truncates toward zero

// Thread = 16 consecutive mask pixels of one row.  A pixel at X = x << 16 is inside a polygon
// when the number of active-edge crossings strictly left of X is odd, or a crossing sits exactly
// on X (which is what pairing the sorted crossings and filling [ceil(xa), floor(xb)] yields).
__global__ void __launch_bounds__(256) k_roi_fill(const RoiEdge* __restrict__ edges, int n_edges, int n_polys,
                                                  int height, int width, uint8_t* __restrict__ mask, int vec_ok) {
  const int chunks = (width + 15) >> 4;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= chunks * height) return;
  const int y = gid / chunks, x0 = (gid % chunks) << 4;
  uint32_t inside = 0;  // bit j: pixel x0 + j
  int e = 0;
  for (int poly = 0; poly < n_polys; ++poly) {
    uint32_t parity = 0, exact = 0;
    for (; e < n_edges && edges[e].poly == poly; ++e) {
      const RoiEdge ed = edges[e];
      if (y < ed.y0 || y >= ed.y1) continue;
      const long long xe = ed.x + (long long)(y - ed.y0) * ed.dx;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const long long X = (long long)(x0 + j) << 16;
        parity ^= (uint32_t)(xe < X) << j;
        exact |= (uint32_t)(xe == X) << j;
      }
    }
    inside |= parity | exact;
  }
  uint8_t* row = mask + (size_t)y * width + x0;
  if (vec_ok) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) v |= ((inside >> (4 * q + j)) & 1u) ? (0xffu << (8 * j)) : 0u;
      w[q] = v;
    }
    *reinterpret_cast<uint4*>(row) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    for (int j = 0; j < 16 && x0 + j < width; ++j) row[j] = ((inside >> j) & 1u) ? 255 : 0;
  }
}

// Thread = one step of one boundary line.  Step i of OpenCV's 8-connected line lies at major
// offset i and minor offset floor((2*minor*i + major - 1) / (2*major)).
__global__ void __launch_bounds__(256) k_roi_lines(const RoiLine* __restrict__ lines, int width, uint8_t* __restrict__ mask) {
  const RoiLine ln = lines[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > ln.major) return;
  const int m = ln.major > 0 ? (int)((2ll * ln.minor * i + ln.major - 1) / (2ll * ln.major)) : 0;
  const int x = ln.vert ? ln.x1 + m : ln.x1 + i;
  const int y = ln.vert ? ln.y1 + ln.sy * i : ln.y1 + ln.sy * m;
  mask[(size_t)y * width + x] = 255;
}

// dst = mask ? src : 0 (cv2.bitwise_and(frame, frame, mask=mask)); thread = 16 pixels.
__global__ void __launch_bounds__(256) k_apply_mask(const uint8_t* __restrict__ src, long long src_pitch,
                                                    const uint8_t* __restrict__ mask, int height, int width,
                                                    uint8_t* __restrict__ dst, long long dst_pitch, int vec_ok) {
  const int chunks = (width + 15) >> 4;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= chunks * height) return;
  const int y = gid / chunks, x0 = (gid % chunks) << 4;
  const uint8_t* s = src + (long long)y * src_pitch + 3ll * x0;
  uint8_t* d = dst + (long long)y * dst_pitch + 3ll * x0;
  const uint8_t* m = mask + (size_t)y * width + x0;
  if (vec_ok) {
    const uint4 mv = __ldg(reinterpret_cast<const uint4*>(m));
    const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
    uint32_t in[12], out[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(s) + q);
      in[4 * q] = v.x, in[4 * q + 1] = v.y, in[4 * q + 2] = v.z, in[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int wd = 0; wd < 12; ++wd) {
      uint32_t sel = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int px = (4 * wd + b) / 3;
        const uint32_t on = (mw[px >> 2] >> (8 * (px & 3))) & 0xffu;
        sel |= on ? (0xffu << (8 * b)) : 0u;
      }
      out[wd] = in[wd] & sel;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
      reinterpret_cast<uint4*>(d)[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
  } else {
    for (int j = 0; j < 16 && x0 + j < width; ++j) {
      const bool on = m[j] != 0;
      d[3 * j] = on ? s[3 * j] : 0;
      d[3 * j + 1] = on ? s[3 * j + 1] : 0;
      d[3 * j + 2] = on ? s[3 * j + 2] : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Motion gate
// ------------------------------------------------------------------------------------------
constexpr int kMotionWarps = 8;
constexpr int kStripPx = 256;             // pixels per warp-row (8 per lane)
constexpr int kRawBytes = 16 + 768 + 16;  // pixels [-2, 258) of a strip, on 16-byte chunk boundaries
constexpr int kMaskBytes = 16 + 256 + 16;
constexpr int kStagesM = 5;  // = the row-loop unroll factor, so stage indices are compile-time constants

struct MotionFrame {
  const uint8_t* src;
  const uint8_t* mask;
  const uint8_t* prev;
  uint8_t* next;
  long long pitch;
  int h, w;
  short has_prev, fast;
  int out_idx;
};
struct MotionParams {
  MotionFrame f[B200VA_LAUNCH_FRAMES];
  int32_t* changed;
  int rows_per_task;
};
static_assert(sizeof(MotionParams) <= 4000, "kernel parameter block too large");

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - i : i;
}
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
  return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// gray of the pixel whose B, G, R bytes sit in the low three bytes of `px` (the 4th byte is ignored):
// two 4-way byte dot products with the 15-bit coefficients split into high and low bytes.
__device__ __forceinline__ uint32_t gray_dp4a(uint32_t px) {
  const uint32_t lo = 151u | (35u << 8) | (70u << 16);  // 3735 = 14*256 + 151, 19235 = 75*256 + 35, 9798 = 38*256 + 70
  const uint32_t hi = 14u | (75u << 8) | (38u << 16);
  return (__dp4a(px, lo, 16384u) + (__dp4a(px, hi, 0u) << 8)) >> 15;
}

// One warp owns a 256-pixel strip (8 pixels per lane) and streams down its rows.  Per row: the BGR
// bytes arrive through a 4-stage cp.async ring; gray values are packed two per register (u16 pairs);
// the two halo pixels on either side come from the neighbouring lanes by shuffle; the horizontal and
// vertical [1 4 6 4 1] passes run on the packed pairs (sums stay below 65536, no carry between
// halves); five rows of horizontal sums live in registers (the row loop is unrolled by five so the
// ring index is static).
// BORDER_REFLECT_101 for indices at most n-1 outside [0, n): no division on the hot path
__device__ __forceinline__ int reflect_near(int i, int n) {
  if (n < 4) return reflect101(i, n);
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// SIMPLE = the common case (16-byte aligned frame, strip entirely inside the image, image at least 4
// pixels wide / high): no per-pixel reflection, no byte loaders, vector stores only.  Everything else
// (partial strips, tiny or unaligned images) runs the same arithmetic through the generic instance;
// keeping the two apart keeps the hot loop small enough for the instruction cache.
template <bool MASK, bool SIMPLE>
__device__ __noinline__ void motion_task(const MotionParams& p, const MotionFrame& f, int frame0, int task, int x0,
                                         int yb, int ye, uint8_t* raw_base, uint8_t* msk_base) {
  const int lane = threadIdx.x & 31;
  const int W = f.w, H = f.h;
  const bool fast = SIMPLE || f.fast != 0;
  const int row_bytes = 3 * W;
  const uint8_t* const src = f.src;
  const uint8_t* const mask = f.mask;
  const long long pitch = f.pitch;

  // per-lane copy plan of a row, fixed for the whole task: raw[16 + 3*(x - x0) + c] holds channel c
  // of pixel x, msk[16 + (x - x0)] its ROI flag; chunks that fall outside the row are skipped
  const int c0_off = 3 * x0 - 16 + 16 * lane, c1_off = c0_off + 512;
  const bool c0_ok = fast && c0_off >= 0 && c0_off + 16 <= row_bytes;
  const bool c1_ok = fast && lane + 32 < kRawBytes / 16 && c1_off >= 0 && c1_off + 16 <= row_bytes;
  const int m_off = x0 - 16 + 16 * lane;
  const bool m_ok = MASK && fast && lane < kMaskBytes / 16 && m_off >= 0 && m_off + 16 <= W;

  auto load_row = [&](int r, int stage) {
    const int rr = SIMPLE ? (r < 0 ? -r : (r >= H ? 2 * H - 2 - r : r)) : reflect_near(r, H);
    const uint8_t* g = src + (long long)rr * pitch;
    uint8_t* raw = raw_base + stage * kRawBytes;
    uint8_t* msk = msk_base + stage * kMaskBytes;
    if (SIMPLE || fast) {
      if (c0_ok) cp_async16(raw + 16 * lane, g + c0_off);
      if (c1_ok) cp_async16(raw + 16 * lane + 512, g + c1_off);
      if (m_ok) cp_async16(msk + 16 * lane, mask + (size_t)rr * W + m_off);
    } else {
      const int lo = max(0, x0 - 2), hi = min(W, x0 + kStripPx + 2);
      for (int i = 3 * lo + lane; i < 3 * hi; i += 32) raw[16 + i - 3 * x0] = __ldg(g + i);
      if (MASK)
        for (int i = lo + lane; i < hi; i += 32) msk[16 + i - x0] = __ldg(mask + (size_t)rr * W + i);
    }
    cp_async_commit();
  };

  const int r_first = yb - 2, r_last = ye + 1;  // rows whose horizontal pass this task needs
  const int nrows = r_last - r_first + 1;
#pragma unroll
  for (int k = 0; k < kStagesM - 1; ++k) {
    if (k < nrows) load_row(r_first + k, k);
    else cp_async_commit();
  }

  uint32_t ring[5][4];  // horizontal sums of 5 consecutive rows, 8 pixels as 4 packed u16 pairs
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) ring[a][b] = 0;
  int changed = 0;
  const int xl = x0 + 8 * lane;
  const bool interior = SIMPLE || xl + 8 <= W;  // all 8 pixels of the lane are real pixels
  // how the strip's outer halo pixels are obtained: 0 = they are staged pixels, 1 = mirror of the
  // lane's own pixels (image edge on a strip boundary), 2 = generic (reflect through the staged row)
  const int left_kind = x0 == 0 ? ((SIMPLE || W >= 4) ? 1 : 2) : 0;
  const int right_kind = x0 + kStripPx + 2 <= W ? 0 : ((SIMPLE || (x0 + kStripPx == W && W >= 4)) ? 1 : 2);
  const int lane_raw = 16 + 24 * lane, lane_msk = 16 + 8 * lane;
  const size_t out_col = (size_t)xl;
  const bool has_prev = f.has_prev != 0;
  const uint8_t* const prev = f.prev;
  uint8_t* const next = f.next;

  // gray of an arbitrary pixel of the staged row (slow path: partial strips, tiny images)
  auto gray_at = [&](const uint8_t* raw, const uint8_t* msk, int x) -> uint32_t {
    const int xr = reflect101(x, W);
    const uint8_t* px = raw + 16 + 3 * (xr - x0);
    if (MASK && !msk[16 + xr - x0]) return 0u;
    return gray_of(px[0], px[1], px[2]);
  };

  for (int k0 = 0; k0 < nrows; k0 += 5) {
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int k = k0 + u;
      if (k >= nrows) break;
      const int stage = u;  // k0 is a multiple of 5 = kStagesM: a constant after unrolling
      if (k + kStagesM - 1 < nrows) load_row(r_first + k + kStagesM - 1, (u + kStagesM - 1) % kStagesM);
      else cp_async_commit();
      cp_async_wait<kStagesM - 1>();
      __syncwarp();
      const uint8_t* raw = raw_base + stage * kRawBytes;
      const uint8_t* msk = msk_base + stage * kMaskBytes;

      // ---- gray of the lane's 8 pixels, packed as pairs P[q] = g[2q] | g[2q+1] << 16 ----
      uint32_t P[4];
      if (SIMPLE || interior) {
        const uint2* q = reinterpret_cast<const uint2*>(raw + lane_raw);
        const uint2 w0 = q[0], w1 = q[1], w2 = q[2];
        uint32_t g8[8];
        g8[0] = gray_dp4a(w0.x);
        g8[1] = gray_dp4a(__funnelshift_r(w0.x, w0.y, 24));
        g8[2] = gray_dp4a(__funnelshift_r(w0.y, w1.x, 16));
        g8[3] = gray_dp4a(__funnelshift_r(w1.x, w1.y, 8));
        g8[4] = gray_dp4a(w1.y);
        g8[5] = gray_dp4a(__funnelshift_r(w1.y, w2.x, 24));
        g8[6] = gray_dp4a(__funnelshift_r(w2.x, w2.y, 16));
        g8[7] = gray_dp4a(w2.y >> 8);
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) P[q2] = g8[2 * q2] | (g8[2 * q2 + 1] << 16);
        if (MASK) {  // masked pixels are black (bitwise_and with the mask): gray 0
          const uint2 mk = *reinterpret_cast<const uint2*>(msk + lane_msk);
          // any non-zero mask byte -> 0xff
          const uint32_t m0 = __vcmpne4(mk.x, 0u), m1 = __vcmpne4(mk.y, 0u);
          P[0] &= __byte_perm(m0, 0u, 0x1100);
          P[1] &= __byte_perm(m0, 0u, 0x3322);
          P[2] &= __byte_perm(m1, 0u, 0x1100);
          P[3] &= __byte_perm(m1, 0u, 0x3322);
        }
      } else {
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) {
          const int x = xl + 2 * q2;
          const uint32_t a = x < W + 2 ? gray_at(raw, msk, x) : 0u;
          const uint32_t b = x + 1 < W + 2 ? gray_at(raw, msk, x + 1) : 0u;
          P[q2] = a | (b << 16);
        }
      }
      // ---- halo pairs from the neighbouring lanes; the strip's outer halo from the staged row ----
      uint32_t L = __shfl_up_sync(0xffffffffu, P[3], 1);    // pixels xl-2, xl-1
      uint32_t R = __shfl_down_sync(0xffffffffu, P[0], 1);  // pixels xl+8, xl+9
      if (lane == 0) {
        if (left_kind == 0) {  // bytes 10..15 of the staged row: pixels x0-2 and x0-1
          const uint32_t a = *reinterpret_cast<const uint32_t*>(raw + 8), b = *reinterpret_cast<const uint32_t*>(raw + 12);
          uint32_t ga = gray_dp4a(__funnelshift_r(a, b, 16)), gb = gray_dp4a(b >> 8);
          if (MASK) {
            ga = msk[14] ? ga : 0u;
            gb = msk[15] ? gb : 0u;
          }
          L = ga | (gb << 16);
        } else if (SIMPLE || left_kind == 1) {  // reflect: g[-2] = g[2], g[-1] = g[1]
          L = (P[1] & 0xffffu) | (P[0] & 0xffff0000u);
        } else {
          L = gray_at(raw, msk, x0 - 2) | (gray_at(raw, msk, x0 - 1) << 16);
        }
      }
      if (lane == 31) {
        if (right_kind == 0) {  // bytes 784..789: pixels x0+256 and x0+257
          const uint32_t a = *reinterpret_cast<const uint32_t*>(raw + 784), b = *reinterpret_cast<const uint32_t*>(raw + 788);
          uint32_t ga = gray_dp4a(a), gb = gray_dp4a(__funnelshift_r(a, b, 24));
          if (MASK) {
            ga = msk[16 + kStripPx] ? ga : 0u;
            gb = msk[17 + kStripPx] ? gb : 0u;
          }
          R = ga | (gb << 16);
        } else if (SIMPLE || right_kind == 1) {  // reflect: g[W] = g[W-2], g[W+1] = g[W-3]
          R = (P[3] & 0xffffu) | (P[2] & 0xffff0000u);
        } else {
          const int x = x0 + kStripPx;
          R = (x < W + 2 ? gray_at(raw, msk, x) : 0u) | ((x + 1 < W + 2 ? gray_at(raw, msk, x + 1) : 0u) << 16);
        }
      }
      // ---- horizontal [1 4 6 4 1] on pairs: h_k = P[k-1] + P[k+1] + 4 (Q[k-1] + Q[k]) + 6 P[k], Q[k] = (g[2k+1], g[2k+2]) ----
      {
        const uint32_t Qm = __funnelshift_r(L, P[0], 16), Q0 = __funnelshift_r(P[0], P[1], 16),
                       Q1 = __funnelshift_r(P[1], P[2], 16), Q2 = __funnelshift_r(P[2], P[3], 16),
                       Q3 = __funnelshift_r(P[3], R, 16);
        ring[u][0] = L + P[1] + 4u * (Qm + Q0) + 6u * P[0];
        ring[u][1] = P[0] + P[2] + 4u * (Q0 + Q1) + 6u * P[1];
        ring[u][2] = P[1] + P[3] + 4u * (Q1 + Q2) + 6u * P[2];
        ring[u][3] = P[2] + R + 4u * (Q2 + Q3) + 6u * P[3];
      }
      __syncwarp();  // every lane is done with this stage before a later iteration refills it

      const int y = r_first + k - 2;  // the output row whose five inputs are now in the ring
      if (k >= 4 && y < ye) {
        uint32_t out[2];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t pq[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * hlf + e;
            // rows r-4 .. r sit in slots (u+1)%5 .. u; the kernel is symmetric so only the distance matters
            const uint32_t v = ring[(u + 1) % 5][c] + ring[u][c] + 4u * (ring[(u + 2) % 5][c] + ring[(u + 4) % 5][c]) +
                               6u * ring[(u + 3) % 5][c] + 0x00800080u;
            pq[e] = (v >> 8) & 0x00ff00ffu;
          }
          out[hlf] = __byte_perm(pq[0], pq[1], 0x6420);
        }
        const size_t o = (size_t)y * (size_t)W + out_col;
        if (SIMPLE || (fast && interior)) {
          if (has_prev) {
            const uint2 pv = __ldg(reinterpret_cast<const uint2*>(prev + o));
            const uint32_t d0 = __vcmpgtu4(__vabsdiffu4(out[0], pv.x), 0x19191919u);
            const uint32_t d1 = __vcmpgtu4(__vabsdiffu4(out[1], pv.y), 0x19191919u);
            changed += (__popc(d0) + __popc(d1)) >> 3;
          }
          *reinterpret_cast<uint2*>(next + o) = make_uint2(out[0], out[1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (xl + j < W) {
              const int nv = (out[j >> 2] >> (8 * (j & 3))) & 0xff;
              if (has_prev) {
                const int pv = prev[o + j];
                changed += (abs(nv - pv) > 25) ? 1 : 0;
              }
              next[o + j] = (uint8_t)nv;
            }
          }
        }
      }
    }
  }
  changed = warp_sum(changed);
  if (lane == 0) {
    if (has_prev) {
      if (changed) atomicAdd(p.changed + frame0 + f.out_idx, changed);
    } else if (task == 0) {
      p.changed[frame0 + f.out_idx] = -1;
    }
  }
}

template <bool MASK>
__global__ void __launch_bounds__(kMotionWarps * 32, 3) k_motion(const __grid_constant__ MotionParams p, int frame0) {
  __shared__ __align__(16) uint8_t s_raw[kMotionWarps][kStagesM][kRawBytes];
  __shared__ __align__(16) uint8_t s_msk[MASK ? kMotionWarps : 1][kStagesM][kMaskBytes];
  const int warp = threadIdx.x >> 5;
  const MotionFrame& f = p.f[blockIdx.y];
  const int W = f.w, H = f.h;
  const int strips = (W + kStripPx - 1) / kStripPx;
  const int row_tasks = (H + p.rows_per_task - 1) / p.rows_per_task;
  const int task = blockIdx.x * kMotionWarps + warp;
  if (task >= strips * row_tasks) return;
  const int x0 = (task % strips) * kStripPx;
  const int yb = (task / strips) * p.rows_per_task;
  const int ye = min(H, yb + p.rows_per_task);
  motion_task<MASK, false>(p, f, frame0, task, x0, yb, ye, &s_raw[warp][0][0], &s_msk[MASK ? warp : 0][0][0]);
}

// ------------------------------------------------------------------------------------------
// Tile kernel: the motion gate for 16-byte aligned frames (width a multiple of 16, at least 4 x 4)
// -- every real video format.  Same arithmetic as motion_task at about half the instructions:
//   * a CTA owns up to 8 adjacent 256-pixel strips x rows_per_task rows.  A ninth warp feeds the TMA
//     engine: per image row ONE cp.async.bulk each for the tile's BGR bytes, its ROI flags and the
//     previous blurred gray of the row that becomes final, into a 5-stage ring guarded by full /
//     empty mbarriers.  The eight compute warps never issue a global load and never wait for the
//     previous-gray read (36 % of all stall samples in the per-warp cp.async version).
//   * gray = two 16x8-bit dot products per pixel (IDP.2A) straight from the packed BGR words, no
//     funnel shifts to isolate a pixel; the coefficients are doubled so the gray byte lands in
//     bits 16..23 and one PRMT packs two pixels.
//   * the two halo pixels on either side of a lane's 8 pixels are read from the staged row (the
//     strips of a tile are contiguous in shared memory): no shuffles, no edge lanes.
//   * rounding + packing of the vertical pass is one PRMT; |diff| > 25 is three logic ops per four
//     pixels.
// ------------------------------------------------------------------------------------------
constexpr int kTileStrips = 8;
constexpr int kTileRaw = 16 + kTileStrips * 3 * kStripPx + 16;  // bytes [-16, 6160) of the tile's row piece
constexpr int kTileMsk = 16 + kTileStrips * kStripPx + 16;
constexpr int kTilePrev = kTileStrips * kStripPx;
constexpr int kTileStage = kTileRaw + kTileMsk + kTilePrev;  // 10304, a multiple of 16
constexpr int kTileStages = 5;                               // = the row-loop unroll factor
constexpr int kTileThreads = (kTileStrips + 1) * 32;
constexpr int kTileRowsMax = 64;  // rows_per_task never exceeds this
constexpr int kTileSmem = kTileStages * kTileStage + 2 * kTileStages * 8 + kTileRowsMax * 8;  // stages, barriers, fused-pass row table

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// per byte: bit 7 set where the byte is non-zero
__device__ __forceinline__ uint32_t nonzero_msb4(uint32_t m) { return ((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m; }

// doubled BGR2GRAY coefficients as 16-bit pairs for dp2a: 2*3735, 2*19235, 2*9798
constexpr uint32_t kBG = 7470u | (38470u << 16), kR0 = 19596u, k0B = 7470u << 16, kGR = 38470u | (19596u << 16);
constexpr uint32_t kRnd = 32768u;  // gray = (2*sum + 2*16384) >> 16 = (sum + 16384) >> 15

// grays (in bits 16..23) of the pixels packed at byte phase 0 / 3 / 6 / 9 of consecutive words
__device__ __forceinline__ uint32_t gray_ph0(uint32_t w0) { return __dp2a_hi(kR0, w0, __dp2a_lo(kBG, w0, kRnd)); }
__device__ __forceinline__ uint32_t gray_ph3(uint32_t w0, uint32_t w1) { return __dp2a_lo(kGR, w1, __dp2a_hi(k0B, w0, kRnd)); }
__device__ __forceinline__ uint32_t gray_ph6(uint32_t w1, uint32_t w2) { return __dp2a_lo(kR0, w2, __dp2a_hi(kBG, w1, kRnd)); }
__device__ __forceinline__ uint32_t gray_ph9(uint32_t w2) { return __dp2a_hi(kGR, w2, __dp2a_lo(k0B, w2, kRnd)); }

// ---- fused letterbox (b200va_motion_preprocess) ---------------------------------------------------
// The motion pass already stages every byte of the frame (and of its ROI mask) in shared memory, so the network
// input can be produced from the same staged rows instead of reading the tapped rows (and their mask rows) from
// HBM a second time: 11 MB of the 16 MB the masked 4K letterbox moves per frame.  When the row that has just
// arrived is the SECOND source row of a destination row d (rowmap[r - 1] = d; the first one is still in the
// previous stage: stages are released one row late in this variant), the warp interpolates the destination
// columns whose first tap lies in its 256-pixel strip -- OpenCV's two-tap fixed-point formula, unchanged -- and
// writes them to the planar output.  Pad rows / columns come from k_lb_pads.
struct LbFrame {
  int xtab, ytab, rowmap, colstart;  // offsets into the tap arena (16-byte entries)
  int out_idx;                       // position in the output batch
};
constexpr int kFuseFrames = 32;  // per launch: the parameter block stays below 4 KB
struct MotionLbParams {
  MotionFrame f[kFuseFrames];
  LbFrame lb[kFuseFrames];
  int32_t* changed;
  int rows_per_task;
  const int4* tabs;
  void* out;
  int dst_h, dst_w;
};
static_assert(sizeof(MotionLbParams) <= 4000, "kernel parameter block too large");

// one destination row, the columns [cb, ce) of this warp's strip; s0 / s1 = stages holding source rows y0 / y0 + 1
template <int LB, bool MASK>
__device__ __forceinline__ void lb_row(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1,
                                       const TapX* __restrict__ xt, const TapY ty, int cb, int ce, int lane, int tile_x0,
                                       void* out_row0, size_t plane) {
  const int b0 = ty.b0, b1 = ty.b1;
  const int base = 16 - 3 * tile_x0, mbase = kTileRaw + 16 - tile_x0;
  for (int c = cb + lane; c < ce; c += 32) {
    const int4 e = __ldg(reinterpret_cast<const int4*>(xt) + c);
    const TapX t = *reinterpret_cast<const TapX*>(&e);
    const int o0 = base + t.off0, o1 = base + t.off1;
    int a0 = t.a0, a1 = t.a1, c0 = a0, c1 = a1;
    if (MASK) {  // apply_roi zeroes masked source pixels before the resize (pipeline.py:149-154)
      const int m0 = mbase + (t.mx0 & 0xffff), m1 = mbase + ((unsigned)t.mx0 >> 16);
      a0 = s0[m0] ? a0 : 0;
      a1 = s0[m1] ? a1 : 0;
      c0 = s1[m0] ? c0 : 0;
      c1 = s1[m1] ? c1 : 0;
    }
    int v[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int r0 = (int)s0[o0 + ch] * a0 + (int)s0[o1 + ch] * a1;
      const int r1 = b1 ? (int)s1[o0 + ch] * c0 + (int)s1[o1 + ch] * c1 : 0;
      v[ch] = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    }
    if (LB == B200VA_OUT_F32_RGB_NCHW) {
      const float k = __int_as_float(0x3B808081);  // float32(1.0/255.0), detector.py:251
      float* o = (float*)out_row0 + c;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) o[pl * plane] = __fmul_rn((float)v[2 - pl], k);
    } else {
      const float k = 0.0039215087890625f;  // float(float16(1.0/255.0))
      __half* o = (__half*)out_row0 + c;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) o[pl * plane] = __float2half_rn(__fmul_rn((float)v[2 - pl], k));
    }
  }
}

// pad rows and pad columns of the fused pass's output (copyMakeBorder value 114, detector.py:233-241)
template <int LB>
__global__ void __launch_bounds__(256) k_lb_pads(const __grid_constant__ MotionLbParams p) {
  const LbFrame& L = p.lb[blockIdx.y];
  const int d = blockIdx.x;
  const TapX* xt = reinterpret_cast<const TapX*>(p.tabs + L.xtab);
  const TapY ty = *reinterpret_cast<const TapY*>(p.tabs + L.ytab + d);
  const size_t plane = (size_t)p.dst_h * p.dst_w;
  const bool pad_row = ty.y0 < 0;
  for (int c = threadIdx.x; c < p.dst_w; c += blockDim.x) {
    if (!pad_row && xt[c].off0 >= 0) continue;
    const size_t o = (size_t)L.out_idx * 3 * plane + (size_t)d * p.dst_w + c;
    if (LB == B200VA_OUT_F32_RGB_NCHW) {
      const float v = __fmul_rn(114.f, __int_as_float(0x3B808081));
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) ((float*)p.out)[o + pl * plane] = v;
    } else {
      const __half v = __float2half_rn(__fmul_rn(114.f, 0.0039215087890625f));
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) ((__half*)p.out)[o + pl * plane] = v;
    }
  }
}

template <bool MASK, int LB, typename P>
__global__ void __launch_bounds__(kTileThreads, 3) k_motion_tile(const __grid_constant__ P p) {
  extern __shared__ __align__(16) uint8_t s_tile[];
  uint64_t* const full = reinterpret_cast<uint64_t*>(s_tile + kTileStages * kTileStage);
  uint64_t* const empty = full + kTileStages;
  const MotionFrame& f = p.f[blockIdx.y];
  const int W = f.w, H = f.h;
  const int tiles_x = (W + kTileStrips * kStripPx - 1) / (kTileStrips * kStripPx);
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int yb = ty * p.rows_per_task;
  if (yb >= H) return;
  const int ye = min(H, yb + p.rows_per_task);
  const int tile_x0 = tx * kTileStrips * kStripPx;
  const int tile_w = min(W - tile_x0, kTileStrips * kStripPx);
  const int n_warps = (tile_w + kStripPx - 1) / kStripPx;  // compute warps with at least one pixel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_prev = f.has_prev != 0;
  const int r_first = yb - 2;
  const int nrows = ye - yb + 4;  // rows yb-2 .. ye+1 (reflected at the image border)

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kTileStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (uint32_t)n_warps);
    }
    mbar_fence_init();
  }
  // fused letterbox: which rows of this tile are the first source row of a destination row, and that row's
  // vertical weights -- looked up once per CTA, so the row loop never waits for a global load
  int2* const s_lbrow = reinterpret_cast<int2*>(empty + kTileStages);
  if constexpr (LB >= 0) {
    const LbFrame& L = p.lb[blockIdx.y];
    for (int i = threadIdx.x; i < ye - yb; i += kTileThreads) {
      const int d = __ldg(reinterpret_cast<const int*>(p.tabs + L.rowmap) + yb + i);
      int wts = 0;
      if (d >= 0) {
        const int4 te = __ldg(p.tabs + L.ytab + d);
        const TapY ty = *reinterpret_cast<const TapY*>(&te);
        wts = (int)(uint16_t)ty.b0 | ((int)(uint16_t)ty.b1 << 16);
      }
      s_lbrow[i] = make_int2(d, wts);
    }
  }
  __syncthreads();

  if (warp == kTileStrips) {
    // ---- producer: one lane drives the TMA engine ----
    if (lane != 0) return;
    const int raw_lo = max(0, 3 * tile_x0 - 16), raw_hi = min(3 * W, 3 * (tile_x0 + tile_w) + 16);
    const int msk_lo = max(0, tile_x0 - 16), msk_hi = min(W, tile_x0 + tile_w + 16);
    const uint32_t raw_dst = (uint32_t)(raw_lo - (3 * tile_x0 - 16)), raw_n = (uint32_t)(raw_hi - raw_lo);
    const uint32_t msk_dst = (uint32_t)(kTileRaw + msk_lo - (tile_x0 - 16)), msk_n = (uint32_t)(msk_hi - msk_lo);
    const uint8_t* const src0 = f.src + raw_lo;
    const uint8_t* const mask0 = MASK ? f.mask + msk_lo : nullptr;
    const uint8_t* const prev0 = has_prev ? f.prev + tile_x0 : nullptr;
    const long long pitch = f.pitch;
    for (int k = 0; k < nrows; ++k) {
      const int stage = k % kTileStages, use = k / kTileStages;
      if (use > 0) mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);
      const int r = r_first + k;
      const int rr = r < 0 ? -r : (r >= H ? 2 * H - 2 - r : r);
      uint8_t* st = s_tile + stage * kTileStage;
      const bool want_prev = has_prev && k >= 4;  // row r - 2 becomes final when row r has been consumed
      mbar_expect_tx(&full[stage], raw_n + (MASK ? msk_n : 0u) + (want_prev ? (uint32_t)tile_w : 0u));
      bulk_g2s(st + raw_dst, src0 + (long long)rr * pitch, raw_n, &full[stage]);
      if (MASK) bulk_g2s(st + msk_dst, mask0 + (size_t)rr * W, msk_n, &full[stage]);
      if (want_prev) bulk_g2s(st + kTileRaw + kTileMsk, prev0 + (size_t)(r - 2) * W, (uint32_t)tile_w, &full[stage]);
    }
    return;
  }
  if (warp >= n_warps) return;

  // ---- compute warps: 8 pixels per lane ----
  const int x0 = tile_x0 + warp * kStripPx;
  const int xl = x0 + 8 * lane;
  const bool live = xl < W;                       // W is a multiple of 8: a lane is all inside or all outside
  const bool left_reflect = xl == 0, right_reflect = xl + 8 == W;
  const int raw_off = 16 + 3 * (warp * kStripPx + 8 * lane);          // byte of the lane's first pixel
  const int msk_off = kTileRaw + 16 + warp * kStripPx + 8 * lane;
  const int prv_off = kTileRaw + kTileMsk + warp * kStripPx + 8 * lane;
  uint8_t* out_ptr = f.next + (size_t)yb * W + xl;
  // fused letterbox: this warp's destination columns and the tables of the frame
  const TapX* lb_xt = nullptr;
  const int4* lb_yt = nullptr;
  const int* lb_rowmap = nullptr;
  int lb_cb = 0, lb_ce = 0;
  size_t lb_plane = 0;
  uint8_t* lb_out = nullptr;
  if constexpr (LB >= 0) {
    const LbFrame& L = p.lb[blockIdx.y];
    lb_xt = reinterpret_cast<const TapX*>(p.tabs + L.xtab);
    lb_yt = p.tabs + L.ytab;
    lb_rowmap = reinterpret_cast<const int*>(p.tabs + L.rowmap);
    const int* cs = reinterpret_cast<const int*>(p.tabs + L.colstart);
    lb_cb = __ldg(cs + x0 / kStripPx);
    lb_ce = __ldg(cs + x0 / kStripPx + 1);
    lb_plane = (size_t)p.dst_h * p.dst_w;
    lb_out = (uint8_t*)p.out + (size_t)L.out_idx * 3 * lb_plane * (LB == B200VA_OUT_F32_RGB_NCHW ? 4 : 2);
  }

  uint32_t ring[5][4];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) ring[a][b] = 0;
  int changed = 0;
  uint32_t parity = 0;

  for (int k0 = 0; k0 < nrows; k0 += 5) {
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int k = k0 + u;
      if (k >= nrows) break;
      mbar_wait(&full[u], parity);
      const uint8_t* st = s_tile + u * kTileStage;

      // ---- gray of pixels xl-2 .. xl+9 as pairs: L | P[0..3] | R ----
      uint32_t P[4], L, R;
      {
        const uint2* q = reinterpret_cast<const uint2*>(st + raw_off - 8);
        const uint2 h0 = q[0], a = q[1], b = q[2], c = q[3], h1 = q[4];
        L = prmt(gray_ph6(h0.x, h0.y), gray_ph9(h0.y), 0x7632u);       // bytes -6 -5 -4 | -3 -2 -1
        P[0] = prmt(gray_ph0(a.x), gray_ph3(a.x, a.y), 0x7632u);
        P[1] = prmt(gray_ph6(a.y, b.x), gray_ph9(b.x), 0x7632u);
        P[2] = prmt(gray_ph0(b.y), gray_ph3(b.y, c.x), 0x7632u);
        P[3] = prmt(gray_ph6(c.x, c.y), gray_ph9(c.y), 0x7632u);
        R = prmt(gray_ph0(h1.x), gray_ph3(h1.x, h1.y), 0x7632u);       // bytes 24 25 26 | 27 28 29
      }
      if (MASK) {  // pixels outside the ROI are black (bitwise_and with the mask): gray 0
        const uint2 mk = *reinterpret_cast<const uint2*>(st + msk_off);
        const uint32_t m0 = nonzero_msb4(mk.x), m1 = nonzero_msb4(mk.y);
        P[0] &= prmt(m0, 0u, 0x9988u);  // selector bit 3: replicate the selected byte's msb
        P[1] &= prmt(m0, 0u, 0xbbaau);
        P[2] &= prmt(m1, 0u, 0x9988u);
        P[3] &= prmt(m1, 0u, 0xbbaau);
        const uint32_t ml = nonzero_msb4((uint32_t) * reinterpret_cast<const uint16_t*>(st + msk_off - 2));
        const uint32_t mr = nonzero_msb4((uint32_t) * reinterpret_cast<const uint16_t*>(st + msk_off + 8));
        L &= prmt(ml, 0u, 0x9988u);
        R &= prmt(mr, 0u, 0x9988u);
      }
      // BORDER_REFLECT_101: g[-2] = g[2], g[-1] = g[1];  g[W] = g[W-2], g[W+1] = g[W-3]
      if (left_reflect) L = prmt(P[1], P[0], 0x7610u);
      if (right_reflect) R = prmt(P[3], P[2], 0x7610u);
      // ---- horizontal [1 4 6 4 1] on pairs: h_k = P[k-1] + P[k+1] + 4 (Q[k-1] + Q[k]) + 6 P[k], Q[k] = (g[2k+1], g[2k+2]) ----
      {
        const uint32_t Qm = __funnelshift_r(L, P[0], 16), Q0 = __funnelshift_r(P[0], P[1], 16),
                       Q1 = __funnelshift_r(P[1], P[2], 16), Q2 = __funnelshift_r(P[2], P[3], 16),
                       Q3 = __funnelshift_r(P[3], R, 16);
        ring[u][0] = L + P[1] + 4u * (Qm + Q0) + 6u * P[0];
        ring[u][1] = P[0] + P[2] + 4u * (Q0 + Q1) + 6u * P[1];
        ring[u][2] = P[1] + P[3] + 4u * (Q1 + Q2) + 6u * P[2];
        ring[u][3] = P[2] + R + 4u * (Q2 + Q3) + 6u * P[3];
      }
      if (k >= 4) {  // output row yb + k - 4: its five inputs are in the ring
        uint32_t v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          v[c] = ring[(u + 1) % 5][c] + ring[u][c] + 4u * (ring[(u + 2) % 5][c] + ring[(u + 4) % 5][c]) +
                 6u * ring[(u + 3) % 5][c] + 0x00800080u;
        // (v + 128) >> 8 of both halves of two pairs = bytes 1 and 3 of each word
        const uint32_t o0 = prmt(v[0], v[1], 0x7531u), o1 = prmt(v[2], v[3], 0x7531u);
        if (has_prev) {
          const uint2 pv = *reinterpret_cast<const uint2*>(st + prv_off);
          const uint32_t d0 = __vabsdiffu4(o0, pv.x), d1 = __vabsdiffu4(o1, pv.y);
          // per byte d > 25  <=>  bit 7 of ((d & 0x7f) + 102) | d
          const uint32_t t0 = (((d0 & 0x7f7f7f7fu) + 0x66666666u) | d0) & 0x80808080u;
          const uint32_t t1 = (((d1 & 0x7f7f7f7fu) + 0x66666666u) | d1) & 0x80808080u;
          if (live) changed += __popc(t0) + __popc(t1);
        }
        if (live) *reinterpret_cast<uint2*>(out_ptr) = make_uint2(o0, o1);
        out_ptr += W;
      }
      if constexpr (LB >= 0) {
        const int rp = r_first + k - 1;  // the previous row (still staged): first source row of a destination row?
        if (k >= 1 && rp >= yb && rp < ye) {
          const int2 e = s_lbrow[rp - yb];
          if (e.x >= 0) {
            // everything else the rare path needs is re-derived from the parameter block here, so that none of it
            // stays live across the motion loop (hoisted, it cost 30 registers and one CTA per SM)
            const LbFrame& L = p.lb[blockIdx.y];
            const int d = e.x;
            TapY ty;
            ty.b0 = (short)(e.y & 0xffff);
            ty.b1 = (short)((unsigned)e.y >> 16);
            const int* cs = reinterpret_cast<const int*>(p.tabs + L.colstart) + x0 / kStripPx;
            const size_t esz = LB == B200VA_OUT_F32_RGB_NCHW ? 4 : 2;
            const size_t plane = (size_t)p.dst_h * p.dst_w;
            uint8_t* orow = (uint8_t*)p.out + ((size_t)L.out_idx * 3 * plane + (size_t)d * p.dst_w) * esz;
            lb_row<LB, MASK>(s_tile + ((u + 4) % 5) * kTileStage, st, reinterpret_cast<const TapX*>(p.tabs + L.xtab), ty,
                             __ldg(cs), __ldg(cs + 1), lane, tile_x0, orow, plane);
          }
        }
        __syncwarp();  // every lane has read this stage and the previous one
        if (lane == 0 && k >= 1) mbar_arrive(&empty[(u + 4) % 5]);  // released one row late
      } else {
        __syncwarp();  // every lane has read this stage
        if (lane == 0) mbar_arrive(&empty[u]);
      }
    }
    parity ^= 1u;
  }
  changed = warp_sum(changed);
  if (lane == 0) {
    if (has_prev) {
      if (changed) atomicAdd(p.changed + f.out_idx, changed);
    } else if (blockIdx.x == 0 && warp == 0) {
      p.changed[f.out_idx] = -1;
    }
  }
}

}  // namespace

int filters_configure(b200va_ctx* h) {  // called by b200va_create on the handle's device
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<true, -1, MotionParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<false, -1, MotionParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<true, 0, MotionLbParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<false, 0, MotionLbParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<true, 1, MotionLbParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  CUDA_TRY(h, cudaFuncSetAttribute(k_motion_tile<false, 1, MotionLbParams>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
  return B200VA_OK;
}

extern "C" int b200va_roi_rasterize(b200va_handle h, const int32_t* pts, const int* poly_sizes, int n_polys, int height,
                                    int width, uint8_t* mask_out, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  REQUIRE(h, mask_out && height > 0 && width > 0, "bad mask geometry %dx%d", width, height);
  REQUIRE(h, n_polys >= 0 && (n_polys == 0 || (pts && poly_sizes)), "NULL polygon arrays");
  std::vector<RoiEdge> edges;
  std::vector<RoiLine> lines;
  int max_major = 0;
  const int32_t* q = pts;
  for (int pi = 0; pi < n_polys; ++pi) {
    const int n = poly_sizes[pi];
    REQUIRE(h, n >= 0, "polygon %d has negative size", pi);
    if (n == 0) continue;
    long long px = q[2 * (n - 1)], py = q[2 * (n - 1) + 1];
    for (int i = 0; i < n; ++i) {
      const long long qx = q[2 * i], qy = q[2 * i + 1];
      // boundary segment (clipped to the image like cv::LineIterator does)
      {
        long long x1 = px, y1 = py, x2 = qx, y2 = qy;
        bool ok = true;
        if (!(x1 >= 0 && x1 < width && x2 >= 0 && x2 < width && y1 >= 0 && y1 < height && y2 >= 0 && y2 < height))
          ok = clip_line(width, height, x1, y1, x2, y2);
        if (ok) {
          long long dx = x2 - x1, dy = y2 - y1;
          RoiLine ln;
          ln.sy = 1;
          if (dx < 0) {  // left to right: start from the other end
            dx = -dx;
            dy = -dy;
            x1 = x2;
            y1 = y2;
          }
          if (dy < 0) {
            dy = -dy;
            ln.sy = -1;
          }
          ln.vert = dy > dx;
          ln.major = (int)(ln.vert ? dy : dx);
          ln.minor = (int)(ln.vert ? dx : dy);
          ln.x1 = (int)x1;
          ln.y1 = (int)y1;
          lines.push_back(ln);
          if (ln.major > max_major) max_major = ln.major;
        }
      }
      // scanline edge: slope from the clipped end points, start extrapolated to the unclipped top row
      if (py != qy) {
        long long c0x = px, c0y = py, c1x = qx, c1y = qy;
        if (!(px >= 0 && px < width && qx >= 0 && qx < width && py >= 0 && py < height && qy >= 0 && qy < height)) {
          long long t0x = px, t0y = py, t1x = qx, t1y = qy;
          clip_line(width, height, t0x, t0y, t1x, t1y);
          c0x = t0x;
          c1x = t1x;
          if (t0y != t1y) {
            c0y = t0y;
            c1y = t1y;
          }
        }
        RoiEdge e;
        e.dx = trunc_div((c1x - c0x) * 65536ll, c1y - c0y);
        e.poly = pi;
        e.pad_ = 0;
        if (py < qy) {
          e.y0 = (int)py;
          e.y1 = (int)qy;
          e.x = c0x * 65536ll + (py - c0y) * e.dx;
        } else {
          e.y0 = (int)qy;
          e.y1 = (int)py;
          e.x = c1x * 65536ll + (qy - c1y) * e.dx;
        }
        edges.push_back(e);
      }
      px = qx;
      py = qy;
    }
    q += 2 * n;
  }
  const size_t eb = edges.size() * sizeof(RoiEdge), lb = lines.size() * sizeof(RoiLine);
  const size_t lo = (eb + 255) & ~(size_t)255;
  if (lo + lb > ROI_SCRATCH_BYTES) return set_error(h, B200VA_ERR_CAPACITY, "polygons too complex: %zu edges", edges.size());
  uint8_t* scratch = (uint8_t*)h->roi_scratch;
  PhaseScope phase(h, B200VA_PHASE_ROI, st);
  if (eb) CUDA_TRY(h, cudaMemcpyAsync(scratch, edges.data(), eb, cudaMemcpyHostToDevice, st));
  if (lb) CUDA_TRY(h, cudaMemcpyAsync(scratch + lo, lines.data(), lb, cudaMemcpyHostToDevice, st));
  const int chunks = (width + 15) / 16;
  const long long threads = (long long)chunks * height;
  const int vec_ok = (width % 16 == 0) && ((uintptr_t)mask_out % 16 == 0);
  k_roi_fill<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const RoiEdge*)scratch, (int)edges.size(), n_polys, height,
                                                              width, mask_out, vec_ok);
  LAUNCH_CHECK(h);
  if (!lines.empty()) {
    dim3 grid((max_major + 1 + 255) / 256, (unsigned)lines.size());
    k_roi_lines<<<grid, 256, 0, st>>>((const RoiLine*)(scratch + lo), width, mask_out);
    LAUNCH_CHECK(h);
  }
  // the scratch buffer is reused by the next call: rasterisation is a once-per-stream setup step
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return B200VA_OK;
}

extern "C" int b200va_apply_mask(b200va_handle h, const uint8_t* src, int64_t src_pitch, const uint8_t* mask, int height,
                                 int width, uint8_t* dst, int64_t dst_pitch, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, src && mask && dst && height > 0 && width > 0, "bad arguments");
  REQUIRE(h, src_pitch >= 3ll * width && dst_pitch >= 3ll * width, "pitch smaller than 3*width");
  const int chunks = (width + 15) / 16;
  const long long threads = (long long)chunks * height;
  const int vec_ok = (width % 16 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) &&
                     ((uintptr_t)mask % 16 == 0) && (src_pitch % 16 == 0) && (dst_pitch % 16 == 0);
  PhaseScope phase(h, B200VA_PHASE_ROI, (cudaStream_t)stream);
  k_apply_mask<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, src_pitch, mask, height, width, dst,
                                                                                  dst_pitch, vec_ok);
  LAUNCH_CHECK(h);
  return B200VA_OK;
}

extern "C" int b200va_motion(b200va_handle h, const uint8_t* const* frames, const int* src_h, const int* src_w,
                             const int64_t* src_pitch, int batch, const uint8_t* const* roi_masks,
                             const uint8_t* const* prev_gray, uint8_t* const* next_gray, const int* has_prev,
                             int32_t* changed_out, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  REQUIRE(h, frames && src_h && src_w && next_gray && has_prev && changed_out, "NULL argument");
  REQUIRE(h, batch >= 0 && batch <= h->cfg.max_batch, "batch %d outside [0, %d]", batch, h->cfg.max_batch);
  if (batch == 0) return B200VA_OK;
  PhaseScope phase(h, B200VA_PHASE_MOTION, st);
  CUDA_TRY(h, cudaMemsetAsync(changed_out, 0, sizeof(int32_t) * batch, st));
  // frames with and without an ROI mask go to separate launches (the mask is a template switch), and so do
  // 16-byte aligned frames (tile kernel, TMA-fed) and everything else (generic kernel); `changed` is indexed
  // by the original batch position
  std::vector<int> order[4];
  std::vector<long long> pitches(batch);
  for (int b = 0; b < batch; ++b) {
    REQUIRE(h, frames[b] && next_gray[b], "frame %d: NULL frame or state buffer", b);
    REQUIRE(h, src_h[b] > 0 && src_w[b] > 0, "frame %d has bad size", b);
    REQUIRE(h, !has_prev[b] || (prev_gray && prev_gray[b]), "frame %d: has_prev without a previous buffer", b);
    REQUIRE(h, !has_prev[b] || prev_gray[b] != next_gray[b], "frame %d: prev_gray and next_gray alias", b);
    pitches[b] = src_pitch ? src_pitch[b] : 3ll * src_w[b];
    REQUIRE(h, pitches[b] >= 3ll * src_w[b], "frame %d: pitch smaller than 3*width", b);
    const bool with_mask = roi_masks && roi_masks[b];
    const bool aligned = (src_w[b] % 16 == 0) && ((uintptr_t)frames[b] % 16 == 0) && (pitches[b] % 16 == 0) &&
                         ((uintptr_t)next_gray[b] % 16 == 0) && (!has_prev[b] || (uintptr_t)prev_gray[b] % 16 == 0) &&
                         (!with_mask || (uintptr_t)roi_masks[b] % 16 == 0);
    const bool tile = aligned && src_w[b] >= 4 && src_h[b] >= 4;
    order[(with_mask ? 1 : 0) + (tile ? 2 : 0)].push_back(b);
  }
  for (int kind = 0; kind < 4; ++kind) {
    const int with_mask = kind & 1;
    const bool tile = (kind & 2) != 0;
    const std::vector<int>& idx = order[kind];
    for (size_t base = 0; base < idx.size(); base += B200VA_LAUNCH_FRAMES) {
      const int n = (int)std::min<size_t>(B200VA_LAUNCH_FRAMES, idx.size() - base);
      MotionParams p;
      memset(&p, 0, sizeof(p));
      // rows per task: tall enough that the 4 halo rows stay a small overhead, short enough that the
      // launch fills the SMs several times over
      long long total_px = 0;
      for (int i = 0; i < n; ++i) total_px += (long long)src_h[idx[base + i]] * src_w[idx[base + i]];
      int rows = 64;
      {
        const long long px_per_task = tile ? (long long)kTileStrips * kStripPx : kStripPx;
        const long long want_tasks = tile ? (long long)h->num_sms * 3 * 5 : (long long)h->num_sms * 24 * 3;
        while (rows > 8 && total_px / ((long long)rows * px_per_task) < want_tasks) rows >>= 1;
      }
      p.rows_per_task = rows;
      int max_tasks = 0;
      for (int i = 0; i < n; ++i) {
        const int b = idx[base + i];
        MotionFrame& f = p.f[i];
        f.src = frames[b];
        f.mask = with_mask ? roi_masks[b] : nullptr;
        f.prev = has_prev[b] ? prev_gray[b] : nullptr;
        f.next = next_gray[b];
        f.pitch = pitches[b];
        f.h = src_h[b];
        f.w = src_w[b];
        f.has_prev = has_prev[b] ? 1 : 0;
        f.out_idx = b;
        f.fast = tile ? 1 : 0;  // generic kernel: 0 selects its byte loaders
        const int per_row = tile ? (f.w + kTileStrips * kStripPx - 1) / (kTileStrips * kStripPx) : (f.w + kStripPx - 1) / kStripPx;
        const int row_tasks = (f.h + rows - 1) / rows;
        if (per_row * row_tasks > max_tasks) max_tasks = per_row * row_tasks;
      }
      p.changed = changed_out;
      if (tile) {
        dim3 grid(max_tasks, n);
        if (with_mask) k_motion_tile<true, -1, MotionParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
        else k_motion_tile<false, -1, MotionParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
      } else {
        dim3 grid((max_tasks + kMotionWarps - 1) / kMotionWarps, n);
        if (with_mask) k_motion<true><<<grid, kMotionWarps * 32, 0, st>>>(p, 0);
        else k_motion<false><<<grid, kMotionWarps * 32, 0, st>>>(p, 0);
      }
      LAUNCH_CHECK(h);
    }
  }
  return B200VA_OK;
}

// ---- a11 + a1 in one pass over the frame ----------------------------------------------------------
extern "C" int b200va_motion_preprocess(b200va_handle h, const uint8_t* const* frames, const int* src_h, const int* src_w,
                                        const int64_t* src_pitch, int batch, const uint8_t* const* roi_masks,
                                        const uint8_t* const* prev_gray, uint8_t* const* next_gray, const int* has_prev,
                                        int32_t* changed_out, void* out, int dst_h, int dst_w, int out_format,
                                        b200va_letterbox* meta_out, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  REQUIRE(h, frames && src_h && src_w && next_gray && has_prev && changed_out && out, "NULL argument");
  REQUIRE(h, batch >= 0 && batch <= h->cfg.max_batch, "batch %d outside [0, %d]", batch, h->cfg.max_batch);
  REQUIRE(h, dst_h > 0 && dst_w > 0 && dst_w < 65536, "bad destination size %dx%d", dst_w, dst_h);
  const int fmt = out_format & 0xff;
  REQUIRE(h, (fmt == B200VA_OUT_F32_RGB_NCHW || fmt == B200VA_OUT_F16_RGB_NCHW) &&
                 (out_format & ~(0xff | B200VA_OUT_FLAG_PADS_VALID)) == 0,
          "b200va_motion_preprocess writes B200VA_OUT_F32_RGB_NCHW or B200VA_OUT_F16_RGB_NCHW, not format %d", out_format);
  if (batch == 0) return B200VA_OK;
  const size_t esz = fmt == B200VA_OUT_F32_RGB_NCHW ? 4 : 2;
  const size_t frame_bytes = (size_t)3 * dst_h * dst_w * esz;
  PhaseScope phase(h, B200VA_PHASE_MOTION, st);
  CUDA_TRY(h, cudaMemsetAsync(changed_out, 0, sizeof(int32_t) * batch, st));

  std::vector<int> fused[2];  // by ROI mask
  std::vector<FusePlan> plans(batch);
  std::vector<b200va_letterbox> metas(batch);
  for (int b = 0; b < batch; ++b) {
    REQUIRE(h, frames[b] && next_gray[b], "frame %d: NULL frame or state buffer", b);
    REQUIRE(h, src_h[b] > 0 && src_w[b] > 0, "frame %d has bad size", b);
    REQUIRE(h, !has_prev[b] || (prev_gray && prev_gray[b]), "frame %d: has_prev without a previous buffer", b);
    REQUIRE(h, !has_prev[b] || prev_gray[b] != next_gray[b], "frame %d: prev_gray and next_gray alias", b);
    const int64_t pitch = src_pitch ? src_pitch[b] : (int64_t)3 * src_w[b];
    REQUIRE(h, pitch >= (int64_t)3 * src_w[b], "frame %d: pitch smaller than 3*width", b);
    b200va_letterbox& m = metas[b];
    REQUIRE(h, b200va_letterbox_meta(src_h[b], src_w[b], dst_h, dst_w, &m) == B200VA_OK && m.new_h > 0 && m.new_w > 0,
            "frame %d: bad size %dx%d", b, src_w[b], src_h[b]);
    if (meta_out) meta_out[b] = m;
    const bool with_mask = roi_masks && roi_masks[b];
    const bool aligned = (src_w[b] % 16 == 0) && ((uintptr_t)frames[b] % 16 == 0) && (pitch % 16 == 0) &&
                         ((uintptr_t)next_gray[b] % 16 == 0) && (!has_prev[b] || (uintptr_t)prev_gray[b] % 16 == 0) &&
                         (!with_mask || (uintptr_t)roi_masks[b] % 16 == 0);
    bool ok = aligned && src_w[b] >= 4 && src_h[b] >= 4;
    if (ok) {
      int rc = letterbox_fuse_plan(h, src_h[b], src_w[b], m.new_h, m.new_w, m.pad_top, m.pad_left, dst_h, dst_w, &plans[b]);
      if (rc) return rc;
      ok = plans[b].eligible;
    }
    if (ok) {
      fused[with_mask ? 1 : 0].push_back(b);
    } else {
      // not a fusable frame (unaligned, tiny, or an up-scaling geometry): the two separate kernels, same results
      const uint8_t* one_mask = with_mask ? roi_masks[b] : nullptr;
      const uint8_t* one_prev = has_prev[b] ? prev_gray[b] : nullptr;
      int rc = b200va_motion(h, frames + b, src_h + b, src_w + b, &pitch, 1, &one_mask, &one_prev, next_gray + b, has_prev + b,
                             changed_out + b, stream);
      if (rc) return rc;
      rc = b200va_preprocess(h, frames + b, src_h + b, src_w + b, &pitch, 1, &one_mask, (uint8_t*)out + (size_t)b * frame_bytes,
                             dst_h, dst_w, out_format, nullptr, stream);
      if (rc) return rc;
    }
  }
  for (int with_mask = 0; with_mask < 2; ++with_mask) {
    const std::vector<int>& idx = fused[with_mask];
    for (size_t base = 0; base < idx.size(); base += kFuseFrames) {
      const int n = (int)std::min<size_t>(kFuseFrames, idx.size() - base);
      MotionLbParams p;
      memset(&p, 0, sizeof(p));
      long long total_px = 0;
      for (int i = 0; i < n; ++i) total_px += (long long)src_h[idx[base + i]] * src_w[idx[base + i]];
      int rows = 64;
      while (rows > 8 && total_px / ((long long)rows * kTileStrips * kStripPx) < (long long)h->num_sms * 3 * 5) rows >>= 1;
      p.rows_per_task = rows;
      int max_tasks = 0;
      for (int i = 0; i < n; ++i) {
        const int b = idx[base + i];
        MotionFrame& f = p.f[i];
        f.src = frames[b];
        f.mask = with_mask ? roi_masks[b] : nullptr;
        f.prev = has_prev[b] ? prev_gray[b] : nullptr;
        f.next = next_gray[b];
        f.pitch = src_pitch ? src_pitch[b] : 3ll * src_w[b];
        f.h = src_h[b];
        f.w = src_w[b];
        f.has_prev = has_prev[b] ? 1 : 0;
        f.out_idx = b;
        f.fast = 1;
        p.lb[i] = LbFrame{plans[b].xtab, plans[b].ytab, plans[b].rowmap, plans[b].colstart, b};
        const int per_row = (f.w + kTileStrips * kStripPx - 1) / (kTileStrips * kStripPx);
        const int row_tasks = (f.h + rows - 1) / rows;
        if (per_row * row_tasks > max_tasks) max_tasks = per_row * row_tasks;
      }
      p.changed = changed_out;
      p.tabs = tap_arena(h);
      p.out = out;
      p.dst_h = dst_h;
      p.dst_w = dst_w;
      if (!(out_format & B200VA_OUT_FLAG_PADS_VALID)) {
        dim3 pgrid(dst_h, n);
        if (fmt == B200VA_OUT_F32_RGB_NCHW) k_lb_pads<0><<<pgrid, 256, 0, st>>>(p);
        else k_lb_pads<1><<<pgrid, 256, 0, st>>>(p);
        LAUNCH_CHECK(h);
      }
      dim3 grid(max_tasks, n);
      if (fmt == B200VA_OUT_F32_RGB_NCHW) {
        if (with_mask) k_motion_tile<true, 0, MotionLbParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
        else k_motion_tile<false, 0, MotionLbParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
      } else {
        if (with_mask) k_motion_tile<true, 1, MotionLbParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
        else k_motion_tile<false, 1, MotionLbParams><<<grid, kTileThreads, kTileSmem, st>>>(p);
      }
      LAUNCH_CHECK(h);
    }
  }
  return B200VA_OK;
}
