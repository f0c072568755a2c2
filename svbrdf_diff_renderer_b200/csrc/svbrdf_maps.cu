// Texture-map hand-off between resolutions on the device (include/svbrdf_b200.h, "texture-map hand-off"):
// 8-bit encode (SvbrdfIO.save_textures_th + imwrite), cv2-exact Lanczos-4 resize of the bytes, decode
// (imread + SvbrdfIO.load_textures_th).  Byte/integer work, HBM-bound; float steps use IEEE sqrt/div in the reference's
// operation order (the library is compiled with -fmad=false, so no product is fused into a following add).
//
// Restated from the formulas of /root/reference/src/svbrdf.py:150-189, src/imageio.py:11-76 and OpenCV's
// imgproc/resize.cpp; the CPU restatement the tests compare against is oracle/maps_port.py.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/svbrdf_b200.h"
#include "svbrdf_host.h"

namespace svbrdf_maps {

constexpr int kThreads = 256;

// (x * 255).astype(uint8) for x in [0,1]: truncation (imageio.py:70)
__device__ __forceinline__ unsigned char quant(float x01) { return (unsigned char)(int)(x01 * 255.0f); }
// ((t + 1) / 2).clip(0, 1)
__device__ __forceinline__ float half01(float t) { return fminf(fmaxf((t + 1.0f) / 2.0f, 0.0f), 1.0f); }

template <int V>   // texels per thread: 4 (vector path) or 1
__global__ void __launch_bounds__(kThreads) encode_kernel(const float* __restrict__ tex, long long stride, long long texels, int clamp_input,
                                                          unsigned char* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * kThreads + threadIdx.x) * V;
  if (i >= texels) return;
  float t[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(tex + k * stride + i));
      t[k][0] = q.x; t[k][1 % V] = q.y; t[k][2 % V] = q.z; t[k][3 % V] = q.w;
    } else {
      t[k][0] = __ldg(tex + k * stride + i);
    }
  }
  unsigned char b[10][V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = clamp_input ? fminf(fmaxf(t[k][j], -1.0f), 1.0f) : t[k][j];
    b[0][j] = quant(half01(c[0])); b[1][j] = quant(half01(c[1])); b[2][j] = quant(half01(c[2]));
    b[6][j] = quant(half01(c[5]));
    b[7][j] = quant(half01(c[6])); b[8][j] = quant(half01(c[7])); b[9][j] = quant(half01(c[8]));
    // SvbrdfIO.reconstruct_normal (svbrdf.py:102-108): clamp to 1 (not 1-eps), IEEE sqrt and divide
    const float x = fminf(fmaxf(c[3], -1.0f), 1.0f), y = fminf(fmaxf(c[4], -1.0f), 1.0f);
    const float xx = x * x, yy = y * y;
    const float z = __fsqrt_rn(1.0f - fminf(fmaxf(xx + yy, 0.0f), 1.0f));
    const float norm = __fsqrt_rn(xx + yy + z * z);
    const float n[3] = {__fdiv_rn(x, norm), __fdiv_rn(y, norm), __fdiv_rn(z, norm)};
#pragma unroll
    for (int k = 0; k < 3; ++k) b[3 + k][j] = quant((fminf(fmaxf(n[k], -1.0f), 1.0f) + 1.0f) / 2.0f);
  }
#pragma unroll
  for (int p = 0; p < 10; ++p) {
    if (V == 4) *reinterpret_cast<uchar4*>(out + p * texels + i) = make_uchar4(b[p][0], b[p][1 % V], b[p][2 % V], b[p][3 % V]);
    else out[p * texels + i] = b[p][0];
  }
}

// float(b) / 255 correctly rounded (bit-identical to the IEEE division of imageio.py:18-19)
__device__ __forceinline__ float unit(unsigned char x) { return __fdiv_rn(float(x), 255.0f); }

template <int V>
__global__ void __launch_bounds__(kThreads) decode_kernel(const unsigned char* __restrict__ in, long long texels, float* __restrict__ tex,
                                                          long long stride) {
  const long long i = ((long long)blockIdx.x * kThreads + threadIdx.x) * V;
  if (i >= texels) return;
  unsigned char b[10][V];
#pragma unroll
  for (int p = 0; p < 10; ++p) {
    if (V == 4) {
      const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(in + p * texels + i));
      b[p][0] = q.x; b[p][1 % V] = q.y; b[p][2 % V] = q.z; b[p][3 % V] = q.w;
    } else {
      b[p][0] = __ldg(in + p * texels + i);
    }
  }
  float o[9][V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    o[0][j] = unit(b[0][j]) * 2.0f - 1.0f; o[1][j] = unit(b[1][j]) * 2.0f - 1.0f; o[2][j] = unit(b[2][j]) * 2.0f - 1.0f;
    o[5][j] = unit(b[6][j]) * 2.0f - 1.0f;
    o[6][j] = unit(b[7][j]) * 2.0f - 1.0f; o[7][j] = unit(b[8][j]) * 2.0f - 1.0f; o[8][j] = unit(b[9][j]) * 2.0f - 1.0f;
    // imread "normal" (imageio.py:44-49): *2-1, divide by the float32 norm
    const float x = unit(b[3][j]) * 2.0f - 1.0f, y = unit(b[4][j]) * 2.0f - 1.0f, z = unit(b[5][j]) * 2.0f - 1.0f;
    const float norm = __fsqrt_rn(x * x + y * y + z * z);
    o[3][j] = __fdiv_rn(x, norm);
    o[4][j] = __fdiv_rn(y, norm);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (V == 4) *reinterpret_cast<float4*>(tex + k * stride + i) = make_float4(o[k][0], o[k][1 % V], o[k][2 % V], o[k][3 % V]);
    else tex[k * stride + i] = o[k][0];
  }
}

// ---------------------------------------------------------------------------------------------
// Lanczos-4 resize of uint8 planes, cv2-exact.
// One CTA produces a TW x TH tile of one destination plane: the source window is staged in shared memory (bytes),
// pass 1 filters it horizontally into int32 rows (what cv2's HResizeLanczos4 leaves in its row buffers), pass 2
// filters those vertically and applies FixedPtCast<int, uchar, 22>.  Source reads are coalesced row segments, each
// source byte is read from HBM once per tile; the window of a 2x upscale is (TW/2+8) x (TH/2+8) bytes.
// Windows that do not fit (strong downscales) take the direct kernel below.
// ---------------------------------------------------------------------------------------------
constexpr int TW = 128, TH = 32;           // destination tile (kThreads = 2 * TW: two row phases per column)
static_assert(kThreads == 2 * TW, "resize_tiled_kernel maps 2 threads to a destination column");
constexpr int kMaxWinW = 160, kMaxWinH = 64;   // source window budget: 160*64 B + 64*128*4 B = 42 KB static shared memory

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(kThreads) resize_tiled_kernel(const unsigned char* __restrict__ src, int sh, int sw, unsigned char* __restrict__ dst,
                                                                int dh, int dw, const int* __restrict__ xtap, const short* __restrict__ xco,
                                                                const int* __restrict__ ytap, const short* __restrict__ yco) {
  __shared__ unsigned char s_win[kMaxWinH][kMaxWinW];
  __shared__ __align__(16) int s_rows[kMaxWinH][TW];
  __shared__ int s_yidx[TH][8];            // window-relative source rows of each destination row's 8 taps
  __shared__ __align__(16) short s_yco[TH][8];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int x1 = min(x0 + TW, dw), y1 = min(y0 + TH, dh);
  const int tw = x1 - x0, th = y1 - y0;
  const unsigned char* plane = src + (size_t)blockIdx.z * sh * sw;
  // source window of this tile (taps are clamped to the image, so the window is too)
  const int wx0 = clampi(xtap[x0], 0, sw - 1), wx1 = clampi(xtap[x1 - 1] + 7, 0, sw - 1);
  const int wy0 = clampi(ytap[y0], 0, sh - 1), wy1 = clampi(ytap[y1 - 1] + 7, 0, sh - 1);
  const int ww = wx1 - wx0 + 1, wh = wy1 - wy0 + 1;
  for (int i = threadIdx.x; i < ww * wh; i += kThreads) {
    const int r = i / ww, c = i - r * ww;
    s_win[r][c] = __ldg(plane + (size_t)(wy0 + r) * sw + wx0 + c);
  }
  if (threadIdx.x < th) {
    const int t0 = ytap[y0 + threadIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) s_yidx[threadIdx.x][k] = clampi(t0 + k, 0, sh - 1) - wy0;
    const int4 q = __ldg(reinterpret_cast<const int4*>(yco + (size_t)(y0 + threadIdx.x) * 8));
    *reinterpret_cast<int4*>(&s_yco[threadIdx.x][0]) = q;
  }
  // pass 1 mapping: a thread owns one destination column (c) and every second window row (hh): its 8 taps and weights
  // are loaded once
  const int c = threadIdx.x % TW, hh = threadIdx.x / TW;          // kThreads = 2 * TW
  int ix[8], cx[8];
  if (c < tw) {
    const int t0 = xtap[x0 + c];
    const int4 q = __ldg(reinterpret_cast<const int4*>(xco + (size_t)(x0 + c) * 8));
    const short* cs = reinterpret_cast<const short*>(&q);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      ix[k] = clampi(t0 + k, 0, sw - 1) - wx0;
      cx[k] = cs[k];
    }
  }
  __syncthreads();
  // pass 1: horizontal filter of every window row (what cv2's HResizeLanczos4 leaves in its int row buffers)
  if (c < tw) {
    for (int r = hh; r < wh; r += kThreads / TW) {
      int acc = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += int(s_win[r][ix[k]]) * cx[k];
      s_rows[r][c] = acc;
    }
  } else {
    for (int r = hh; r < wh; r += kThreads / TW) s_rows[r][c] = 0;
  }
  __syncthreads();
  // pass 2: vertical filter + FixedPtCast<int, uchar, 22>.  A thread produces 4 adjacent bytes of one row at a time
  // (LDS.128 of the int rows, one 32-bit store); the row's taps and weights are warp-uniform shared-memory broadcasts.
  const int cq = (threadIdx.x % (TW / 4)) * 4, rr = threadIdx.x / (TW / 4);
  unsigned char* dplane = dst + (size_t)blockIdx.z * dh * dw;
  const bool vec_ok = (dw % 4 == 0) && ((reinterpret_cast<uintptr_t>(dplane) & 3) == 0);
  for (int r = rr; r < th; r += kThreads / (TW / 4)) {
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int4 v = *reinterpret_cast<const int4*>(&s_rows[s_yidx[r][k]][cq]);
      const int w = int(s_yco[r][k]);
      acc[0] += v.x * w; acc[1] += v.y * w; acc[2] += v.z * w; acc[3] += v.w * w;
    }
    unsigned char o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (unsigned char)clampi((acc[j] + (1 << 21)) >> 22, 0, 255);
    unsigned char* out = dplane + (size_t)(y0 + r) * dw + x0 + cq;
    if (vec_ok && cq + 4 <= tw) {
      *reinterpret_cast<uchar4*>(out) = make_uchar4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (cq + j < tw) out[j] = o[j];
    }
  }
}

// Direct form: one thread per destination byte, 64 taps from global memory (L1/L2 cached).  Any ratio.
__global__ void __launch_bounds__(kThreads) resize_direct_kernel(const unsigned char* __restrict__ src, int sh, int sw, unsigned char* __restrict__ dst,
                                                                 int dh, int dw, const int* __restrict__ xtap, const short* __restrict__ xco,
                                                                 const int* __restrict__ ytap, const short* __restrict__ yco) {
  const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  const unsigned char* plane = src + (size_t)blockIdx.z * sh * sw;
  const int tx = xtap[x], ty = ytap[y];
  const short* cx = xco + (size_t)x * 8;
  const short* cy = yco + (size_t)y * 8;
  int acc = 0;
  for (int j = 0; j < 8; ++j) {
    const unsigned char* row = plane + (size_t)clampi(ty + j, 0, sh - 1) * sw;
    int h = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) h += int(__ldg(row + clampi(tx + k, 0, sw - 1))) * int(cx[k]);
    acc += h * int(cy[j]);
  }
  const int v = (acc + (1 << 21)) >> 22;
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)clampi(v, 0, 255);
}

// cv::interpolateLanczos4 (float32 fraction in, double trig, float32 normalisation)
static void lanczos4_coeffs(float x, float* c) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  if (x < 1.1920929e-07f) {
    for (int i = 0; i < 8; ++i) c[i] = 0.f;
    c[3] = 1.f;
    return;
  }
  float sum = 0.f;
  const double y0 = -(double(x) + 3) * M_PI * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
  for (int i = 0; i < 8; ++i) {
    const double y = -(double(x) + 3 - i) * M_PI * 0.25;
    c[i] = float((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    sum += c[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) c[i] *= sum;
}

}  // namespace svbrdf_maps

using namespace svbrdf_maps;

extern "C" {

int svbrdf_maps_encode_u8(const float* tex, int64_t plane_stride, int32_t rows, int32_t cols, int32_t clamp_input, uint8_t* bytes,
                          svbrdf_stream_t stream) {
  if (!tex || !bytes || rows <= 0 || cols <= 0) return SVBRDF_E_BADARG;
  svbrdf::DeviceGuard guard(tex);
  const long long texels = (long long)rows * cols;
  const long long stride = plane_stride ? plane_stride : texels;
  if (stride < texels) return SVBRDF_E_BADARG;
  const bool vec = texels % 4 == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(tex) & 15) == 0 && (reinterpret_cast<uintptr_t>(bytes) & 3) == 0;
  if (vec) encode_kernel<4><<<unsigned((texels / 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(tex, stride, texels, clamp_input, bytes);
  else encode_kernel<1><<<unsigned((texels + kThreads - 1) / kThreads), kThreads, 0, stream>>>(tex, stride, texels, clamp_input, bytes);
  return int(cudaGetLastError());
}

int svbrdf_maps_decode_u8(const uint8_t* bytes, int32_t rows, int32_t cols, float* tex, int64_t plane_stride, svbrdf_stream_t stream) {
  if (!tex || !bytes || rows <= 0 || cols <= 0) return SVBRDF_E_BADARG;
  svbrdf::DeviceGuard guard(tex);
  const long long texels = (long long)rows * cols;
  const long long stride = plane_stride ? plane_stride : texels;
  if (stride < texels) return SVBRDF_E_BADARG;
  const bool vec = texels % 4 == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(tex) & 15) == 0 && (reinterpret_cast<uintptr_t>(bytes) & 3) == 0;
  if (vec) decode_kernel<4><<<unsigned((texels / 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(bytes, texels, tex, stride);
  else decode_kernel<1><<<unsigned((texels + kThreads - 1) / kThreads), kThreads, 0, stream>>>(bytes, texels, tex, stride);
  return int(cudaGetLastError());
}

int svbrdf_lanczos4_tables(int32_t src_size, int32_t dst_size, int32_t* first_tap, int16_t* coef) {
  if (src_size <= 0 || dst_size <= 0 || !first_tap || !coef) return SVBRDF_E_BADARG;
  const double scale = double(src_size) / double(dst_size);
  for (int d = 0; d < dst_size; ++d) {
    float fx = float((d + 0.5) * scale - 0.5);
    const int sx = int(std::floor(fx));
    fx -= float(sx);
    first_tap[d] = sx - 3;
    float c[8];
    lanczos4_coeffs(fx, c);
    for (int k = 0; k < 8; ++k) {
      long v = std::lrintf(c[k] * 2048.0f);               // saturate_cast<short>(float): round to nearest even, saturate
      coef[d * 8 + k] = int16_t(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
    }
  }
  return 0;
}

int svbrdf_resize_lanczos4_u8(const uint8_t* src, int32_t planes, int32_t src_rows, int32_t src_cols, uint8_t* dst, int32_t dst_rows,
                              int32_t dst_cols, const int32_t* x_tap, const int16_t* x_coef, const int32_t* y_tap, const int16_t* y_coef,
                              svbrdf_stream_t stream) {
  if (!src || !dst || planes <= 0 || src_rows <= 0 || src_cols <= 0 || dst_rows <= 0 || dst_cols <= 0) return SVBRDF_E_BADARG;
  svbrdf::DeviceGuard guard(src);
  if (src_rows == dst_rows && src_cols == dst_cols)      // cv::resize returns a copy
    return int(cudaMemcpyAsync(dst, src, size_t(planes) * src_rows * src_cols, cudaMemcpyDeviceToDevice, stream));
  if (!x_tap || !x_coef || !y_tap || !y_coef) return SVBRDF_E_BADARG;
  if (planes > 65535 || dst_rows > 65535 * TH) return SVBRDF_E_UNSUPPORTED;
  // the tiled kernel needs the source window of a tile to fit its shared-memory budget: scale <= ~1.1 per axis
  const bool fits = (long long)TW * src_cols <= (long long)(kMaxWinW - 10) * dst_cols && (long long)TH * src_rows <= (long long)(kMaxWinH - 10) * dst_rows;
  if (fits) {
    dim3 grid((dst_cols + TW - 1) / TW, (dst_rows + TH - 1) / TH, planes);
    resize_tiled_kernel<<<grid, kThreads, 0, stream>>>(src, src_rows, src_cols, dst, dst_rows, dst_cols, x_tap, reinterpret_cast<const short*>(x_coef),
                                                       y_tap, reinterpret_cast<const short*>(y_coef));
  } else {
    if (dst_rows > 65535) return SVBRDF_E_UNSUPPORTED;
    dim3 grid((dst_cols + kThreads - 1) / kThreads, dst_rows, planes);
    resize_direct_kernel<<<grid, kThreads, 0, stream>>>(src, src_rows, src_cols, dst, dst_rows, dst_cols, x_tap, reinterpret_cast<const short*>(x_coef),
                                                        y_tap, reinterpret_cast<const short*>(y_coef));
  }
  return int(cudaGetLastError());
}

}  // extern "C"
