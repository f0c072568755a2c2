// Per-texel arithmetic of the SVBRDF render / backward / Adam path.
//
// One texel is shaded under N point lights; everything a texel needs lives in
// registers: a prologue turns the 9 texture channels into material parameters,
// the light loop accumulates loss and parameter gradients, an epilogue chains
// the gradients back to the 9 channels (and, in the fused mode, applies Adam).
//
// The math restates /root/reference/src/microfacet.py:64-120 (forward) and the
// derivative torch autograd produces for it, plus torch.optim.Adam's update
// (torch/optim/adam.py:531-547) — written from the formulas, not from the code:
//   * the half vector is never materialised: with unit l, v
//         |l+v|^2 = 2 + 2 l.v,  n.h = (n.l + n.v)/|l+v|,  v.h = (1 + l.v)/|l+v|
//   * v and l are never materialised either: n.v = (n.V) * rsqrt(V.V)
//   * x^2.2 is evaluated as x^2 * x^0.2 so the lg2/ex2 error is scaled by 0.2,
//     and x^0.2 is reused for the derivative 2.2 x^1.2 = 2.2 * x * x^0.2
//   * constant factors of the image gradient (2/(N*3*H*W), 1/gamma) are applied
//     once per texel in the epilogue, not per light.
//
// The header is plain C++ when compiled without nvcc: tests/hostemu builds it
// with g++ (float and double) to check the analytic backward against the
// oracle on the CPU.  That host build is a test harness only; the product has
// no CPU path.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>

#if defined(__CUDACC__)
#define SV_HD __host__ __device__ __forceinline__
#define SV_D __device__ __forceinline__
#else
#define SV_HD inline
#define SV_D inline
#endif

namespace svbrdf {

constexpr double kPi = 3.14159265358979323846;
constexpr double kGamma = 2.2;
constexpr double kEps = 1e-6;                       // microfacet.py:14
constexpr double kFresA = -5.55473, kFresB = -6.98316;  // microfacet.py:44

// ---------------------------------------------------------------------------------------------
// Scalar math: MUFU approximations on the device (1-2 ulp, see DESIGN.md "Numerics"),
// libm on the host build.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Fm;

template <>
struct Fm<float> {
  typedef bool mask;                       // per-lane predicate type (a pair of bools for the packed type)
  static constexpr bool kRefine = true;    // MUFU.RSQ needs the Newton step (see rsqrt_dir)
  static SV_HD mask ge(float a, float b) { return a >= b; }
  static SV_HD mask le(float a, float b) { return a <= b; }
  static SV_HD mask eq(float a, float b) { return a == b; }
  static SV_HD mask mand(mask a, mask b) { return a && b; }
  static SV_HD mask mtrue() { return true; }
  static SV_HD float sel(mask m, float a, float b) { return m ? a : b; }
#if defined(__CUDA_ARCH__)
  static SV_D float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static SV_D float rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static SV_D float sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static SV_D float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static SV_D float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static SV_D float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static SV_D float mul(float a, float b) { return __fmul_rn(a, b); }     // a rounded product the compiler may not contract into an FMA
  static SV_D float sub(float a, float b) { return __fsub_rn(a, b); }     // a - b that never absorbs the product a came from
  static SV_D float max(float a, float b) { return fmaxf(a, b); }
  static SV_D float min(float a, float b) { return fminf(a, b); }
#else
  // Host build (tests only).  SV_EMU_MUFU_NOISE perturbs every approximated function by a
  // pseudo-random relative error of up to +-2^-22 (lg2: absolute), the documented bound of
  // the MUFU units, so CPU tests can check that the tolerances hold with device-like math.
#if defined(SV_EMU_MUFU_NOISE)
  static inline int jitter_mask() {
    static const int m = [] { const char* e = std::getenv("SV_EMU_NOISE_MASK"); return e ? std::atoi(e) : 31; }();
    return m;
  }
  static inline float jitter(float y, bool absolute = false, int which = 1) {
    // the device units are exact at their special points: lg2(1) = 0, ex2(0) = 1, rcp(1) = 1, ...
    if (!std::isfinite(y) || y == 0.0f || y == 1.0f || !(jitter_mask() & which)) return y;
    static thread_local uint32_t s = 0x9E3779B9u;
    s = s * 1664525u + 1013904223u;
    const float u = (float(s >> 8) * (1.0f / 8388608.0f) - 1.0f) * 2.3841858e-7f;   // [-2^-22, 2^-22)
    return absolute ? y + u * (std::fabs(y) > 1.0f ? std::fabs(y) : 1.0f) : y * (1.0f + u);
  }
#else
  static inline float jitter(float y, bool = false, int = 0) { return y; }
#endif
  static SV_HD float rcp(float x) { return jitter(1.0f / x, false, 1); }
  static SV_HD float rsqrt(float x) { return jitter(1.0f / std::sqrt(x), false, 2); }
  static SV_HD float sqrt(float x) { return jitter(std::sqrt(x), false, 4); }
  static SV_HD float lg2(float x) { return jitter(std::log2(x), true, 8); }
  static SV_HD float ex2(float x) { return jitter(std::exp2(x), false, 16); }
  static SV_HD float fma(float a, float b, float c) { return std::fma(a, b, c); }
  static SV_HD float mul(float a, float b) { return a * b; }
  static SV_HD float sub(float a, float b) { return a - b; }
  static SV_HD float max(float a, float b) { return a > b ? a : b; }
  static SV_HD float min(float a, float b) { return a < b ? a : b; }
#endif
};

template <>
struct Fm<double> {
  typedef bool mask;
  static constexpr bool kRefine = false;
  static SV_HD mask ge(double a, double b) { return a >= b; }
  static SV_HD mask le(double a, double b) { return a <= b; }
  static SV_HD mask eq(double a, double b) { return a == b; }
  static SV_HD mask mand(mask a, mask b) { return a && b; }
  static SV_HD mask mtrue() { return true; }
  static SV_HD double sel(mask m, double a, double b) { return m ? a : b; }
  static SV_HD double rcp(double x) { return 1.0 / x; }
  static SV_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static SV_HD double sqrt(double x) { return ::sqrt(x); }
  static SV_HD double lg2(double x) { return ::log2(x); }
  static SV_HD double ex2(double x) { return ::exp2(x); }
  static SV_HD double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static SV_HD double mul(double a, double b) { return a * b; }
  static SV_HD double sub(double a, double b) { return a - b; }
  static SV_HD double max(double a, double b) { return a > b ? a : b; }
  static SV_HD double min(double a, double b) { return a < b ? a : b; }
};

// ---------------------------------------------------------------------------------------------
// Packed pair of floats: two texels per thread.  On sm_100 the arithmetic maps to the packed
// FP32x2 instructions (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes' worth of work), which is
// what an issue-bound elementwise kernel wants; MUFU, min/max and selects stay per component.
// The host build (tests) implements the same type with scalar operations.
// ---------------------------------------------------------------------------------------------
struct V2 {
  float x, y;
  SV_HD V2() {}
  SV_HD V2(float a) : x(a), y(a) {}
  SV_HD V2(double a) : x(float(a)), y(float(a)) {}
  SV_HD V2(int a) : x(float(a)), y(float(a)) {}
  SV_HD V2(float a, float b) : x(a), y(b) {}
};
struct M2 {
  bool x, y;
};
#if defined(__CUDA_ARCH__)
SV_D float2 v2f(const V2& a) { return make_float2(a.x, a.y); }
SV_D V2 f2v(const float2& a) { return V2(a.x, a.y); }
SV_D V2 operator+(const V2& a, const V2& b) { return f2v(__fadd2_rn(v2f(a), v2f(b))); }
SV_D V2 operator-(const V2& a, const V2& b) { return f2v(__fadd2_rn(v2f(a), make_float2(-b.x, -b.y))); }
SV_D V2 operator*(const V2& a, const V2& b) { return f2v(__fmul2_rn(v2f(a), v2f(b))); }
SV_D V2 operator-(const V2& a) { return V2(-a.x, -a.y); }
#else
SV_HD V2 operator+(const V2& a, const V2& b) { return V2(a.x + b.x, a.y + b.y); }
SV_HD V2 operator-(const V2& a, const V2& b) { return V2(a.x - b.x, a.y - b.y); }
SV_HD V2 operator*(const V2& a, const V2& b) { return V2(a.x * b.x, a.y * b.y); }
SV_HD V2 operator-(const V2& a) { return V2(-a.x, -a.y); }
#endif
SV_HD V2& operator+=(V2& a, const V2& b) { a = a + b; return a; }
SV_HD V2& operator-=(V2& a, const V2& b) { a = a - b; return a; }

template <>
struct Fm<V2> {
  typedef M2 mask;
  typedef Fm<float> S;
  static constexpr bool kRefine = true;
  static SV_HD mask ge(const V2& a, const V2& b) { return M2{a.x >= b.x, a.y >= b.y}; }
  static SV_HD mask le(const V2& a, const V2& b) { return M2{a.x <= b.x, a.y <= b.y}; }
  static SV_HD mask eq(const V2& a, const V2& b) { return M2{a.x == b.x, a.y == b.y}; }
  static SV_HD mask mand(mask a, mask b) { return M2{a.x && b.x, a.y && b.y}; }
  static SV_HD mask mtrue() { return M2{true, true}; }
  static SV_HD V2 sel(mask m, const V2& a, const V2& b) { return V2(m.x ? a.x : b.x, m.y ? a.y : b.y); }
  static SV_HD V2 rcp(const V2& a) { return V2(S::rcp(a.x), S::rcp(a.y)); }
  static SV_HD V2 rsqrt(const V2& a) { return V2(S::rsqrt(a.x), S::rsqrt(a.y)); }
  static SV_HD V2 sqrt(const V2& a) { return V2(S::sqrt(a.x), S::sqrt(a.y)); }
  static SV_HD V2 lg2(const V2& a) { return V2(S::lg2(a.x), S::lg2(a.y)); }
  static SV_HD V2 ex2(const V2& a) { return V2(S::ex2(a.x), S::ex2(a.y)); }
  static SV_HD V2 max(const V2& a, const V2& b) { return V2(S::max(a.x, b.x), S::max(a.y, b.y)); }
  static SV_HD V2 min(const V2& a, const V2& b) { return V2(S::min(a.x, b.x), S::min(a.y, b.y)); }
  // NOTE: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false (checked with cuobjdump); where the
  // product must be rounded before a subtraction (`encoded - target`), the subtraction is written with sub() below.
  static SV_HD V2 mul(const V2& a, const V2& b) { return a * b; }
#if defined(__CUDA_ARCH__)
  static SV_D V2 fma(const V2& a, const V2& b, const V2& c) { return f2v(__ffma2_rn(v2f(a), v2f(b), v2f(c))); }
  // a - b as two scalar subtractions (opaque to the vectoriser): a packed product followed by a scalar FADD is not
  // contracted, so `encoded - target` sees the encoded value rounded exactly like the render kernel writes it and the L2
  // forward of a rendered target is exactly zero in the packed instantiations too.
  static SV_D V2 sub(const V2& a, const V2& b) {
    V2 r;
    asm("sub.rn.f32 %0, %1, %2;" : "=f"(r.x) : "f"(a.x), "f"(b.x));
    asm("sub.rn.f32 %0, %1, %2;" : "=f"(r.y) : "f"(a.y), "f"(b.y));
    return r;
  }
#else
  static SV_HD V2 fma(const V2& a, const V2& b, const V2& c) { return V2(S::fma(a.x, b.x, c.x), S::fma(a.y, b.y, c.y)); }
  static SV_HD V2 sub(const V2& a, const V2& b) { return V2(a.x - b.x, a.y - b.y); }
#endif
};

// rsqrt with one Newton step: y' = y + (y/2)(1 - x y^2).  MUFU.RSQ is good to ~2 ulp; the cosines
// n.v, n.h feed 1 - c^2 in the GGX denominator, where for the reference's default (smooth)
// roughness an error of 2e-7 in c is a 1 % error of the highlight.  The step brings the
// direction normalisations to <1 ulp, the level of the IEEE sqrt + divide the reference uses.
#ifndef SV_REFINE_RSQRT
#define SV_REFINE_RSQRT 1
#endif
template <typename T>
SV_HD T rsqrt_dir(T x) {
  T y = Fm<T>::rsqrt(x);
#if SV_REFINE_RSQRT
  if (Fm<T>::kRefine) {
    const T e = Fm<T>::fma(-x * y, y, T(1));
    y = Fm<T>::fma(T(0.5) * y, e, y);
  }
#endif
  return y;
}

// x^2.2 for x in [0,1] as x^2 * x^0.2; also returns r = x^0.2 for the derivative.
template <typename T>
SV_HD T pow22(T x, T& r) {
  r = Fm<T>::ex2(T(0.2) * Fm<T>::lg2(x));
  return x * x * r;
}

// ---------------------------------------------------------------------------------------------
// Per-texel state
// ---------------------------------------------------------------------------------------------
// 2^(-5.55473 - 6.98316): the Fresnel term of a co-located light/camera pair (v.h = 1).
constexpr double kSphgColoc = 1.6815857857725708e-4;

template <typename T>
struct Texel {
  T px, py;          // texel centre on the sample plane (z = 0), microfacet.py:16-19
  T n[3];            // unit shading normal, microfacet.py:64-70
  // the light power pw_c (microfacet.py:117) is folded into the per-texel albedo terms:
  T kdp[3];          // pw_c * d_c/pi * (1 - s_c)            (lambert, microfacet.py:102-103)
  T sp[3];           // pw_c * s_c
  T omsp[3];         // pw_c * (1 - s_c)                     (Fresnel: F_c*pw_c = sp + omsp*sphg)
  T Fp[3];           // pw_c * F_c for a co-located pair (constant per texel)
  T a2, k, omk;      // alpha^2, k = alpha/2 + eps, 1 - k  (alpha = rough^2), microfacet.py:30,51,106
  T a2q, a2m1x2;     // a2/4 and 2 (a2 - 1): per-texel constants of the co-located light body
  T mnP;             // -(n.P): n.V = n.C - n.P (SV_NV_FOLD)
};

template <typename T>
struct TexelAux {    // needed again only by the epilogue
  T dpow[7];         // d(x^2.2)/dt = 1.1 * x * x^0.2 for channels 0,1,2,5,6,7,8
  T d[3];            // diffuse albedo
  T oms[3];          // 1 - s_c
  T rough, alpha;
  T mx, my, mz;      // un-normalised normal (clamped nx, ny, reconstructed nz)
  T rlen;            // 1/|m|
  typename Fm<T>::mask in3, in4;     // inner clamp masks on nx, ny (microfacet.py:65-66)
  typename Fm<T>::mask planar_free;  // nx^2+ny^2 <= 1-eps: the nz path carries gradient (microfacet.py:67)
};

template <typename T>
struct Grads {       // accumulated over lights (without the constant image-gradient factor)
  T kdp[3];          // dL/d kdp_c
  T sF[3];           // sum of gfp_c * Q * (1 - sphg)   (co-located: sum of gfp_c * Q; (1-sphg) applied in the epilogue)
  T a2, k;           // general lights: dL/da2, dL/dk.  Co-located lights: SA = sum T u c^2 / 2 and SK = sum T (1-c)/gv (see shade_light_coloc)
  T sT;              // co-located lights: sum of T = gQ * Q (dL/da2 = sT/a2 - 2 SA)
  T n[3];
  T pw[3];           // sum of gfp_c * fp_c = pw_c * dL/dpw_c (only when requested)
  T loss;            // sum of squared differences (L2 modes)
  T loss_g;          // second lane of the packed loss accumulator (SV_PAIR_RG); grads_loss() returns the total
};

template <typename T>
SV_HD T grads_loss(const Grads<T>& g) { return g.loss + g.loss_g; }

// Texel centre: ((j + 0.5)/res - 0.5) * size in the reference's fp32 op order
// (microfacet.py:16-19; y is negated, rows index y).
template <typename T>
SV_HD void texel_position(int row, int col, int res, float size, T& px, T& py) {
  const float fx = ((float(col) + 0.5f) / float(res) - 0.5f) * size;
  const float fy = ((float(row) + 0.5f) / float(res) - 0.5f) * size;
  px = T(fx);
  py = T(-fy);
}

// The same value without the division instruction sequence: q = a/res correctly rounded from the correctly rounded
// reciprocal r = RN(1/res) by one residual correction (Markstein): q0 = RN(a r), rem = a - q0 res (exact in an FMA),
// q = RN(q0 + rem r).  Bit-identical to texel_position for every resolution (tests/test_host_logic.py checks all
// (index, res) pairs up to 4096 and a set of larger ones), power of two or not.
template <typename T>
SV_HD void texel_position_rcp(int row, int col, float res_f, float inv_res, float size, T& px, T& py) {
  typedef Fm<float> S;
  const float ax = float(col) + 0.5f, ay = float(row) + 0.5f;
  float qx = S::mul(ax, inv_res), qy = S::mul(ay, inv_res);
  qx = S::fma(S::fma(-qx, res_f, ax), inv_res, qx);
  qy = S::fma(S::fma(-qy, res_f, ay), inv_res, qy);
  px = T((qx - 0.5f) * size);
  py = T(-((qy - 0.5f) * size));
}

// Prologue: 9 channels -> material parameters.  `t` must already be clamped to [-1,1]
// by the caller when the outer clamp of svbrdf.py:60 applies.
template <typename T>
SV_HD void texel_prologue(const T t[9], const T pw[3], Texel<T>& tx, TexelAux<T>& ax) {
  const T inv_pi = T(1.0 / kPi);
  T r;
  T x;
  for (int c = 0; c < 3; ++c) {
    x = (t[c] + T(1)) * T(0.5);
    ax.d[c] = pow22(x, r);
    ax.dpow[c] = T(0.5 * kGamma) * x * r;
    x = (t[6 + c] + T(1)) * T(0.5);
    const T s = pow22(x, r);
    ax.dpow[4 + c] = T(0.5 * kGamma) * x * r;
    ax.oms[c] = T(1) - s;
    tx.kdp[c] = pw[c] * (ax.d[c] * inv_pi * ax.oms[c]);
    tx.sp[c] = pw[c] * s;
    tx.omsp[c] = pw[c] * ax.oms[c];
    tx.Fp[c] = Fm<T>::fma(tx.omsp[c], T(kSphgColoc), tx.sp[c]);
  }
  x = (t[5] + T(1)) * T(0.5);
  ax.rough = pow22(x, r);
  ax.dpow[3] = T(0.5 * kGamma) * x * r;
  ax.alpha = ax.rough * ax.rough;
  tx.a2 = ax.alpha * ax.alpha;
  tx.k = ax.alpha * T(0.5) + T(kEps);
  tx.omk = T(1) - tx.k;
  tx.a2q = tx.a2 * T(0.25);
  tx.a2m1x2 = (tx.a2 - T(1)) * T(2);

  ax.in3 = Fm<T>::mand(Fm<T>::ge(t[3], T(-1)), Fm<T>::le(t[3], T(1)));
  ax.in4 = Fm<T>::mand(Fm<T>::ge(t[4], T(-1)), Fm<T>::le(t[4], T(1)));
  ax.mx = Fm<T>::min(Fm<T>::max(t[3], T(-1)), T(1));
  ax.my = Fm<T>::min(Fm<T>::max(t[4], T(-1)), T(1));
  const T planar = ax.mx * ax.mx + ax.my * ax.my;
  const T cap = T(1) - T(kEps);
  ax.planar_free = Fm<T>::le(planar, cap);
  const T pc = Fm<T>::min(planar, cap);
  ax.mz = Fm<T>::sqrt(T(1) - pc);
  ax.rlen = rsqrt_dir(ax.mx * ax.mx + ax.my * ax.my + ax.mz * ax.mz);
  tx.n[0] = ax.mx * ax.rlen;
  tx.n[1] = ax.my * ax.rlen;
  tx.n[2] = ax.mz * ax.rlen;
  tx.mnP = -Fm<T>::fma(tx.n[0], tx.px, tx.n[1] * tx.py);      // the sample plane is z = 0 (microfacet.py:19)
}

template <typename T>
SV_HD void grads_zero(Grads<T>& g) {
  for (int c = 0; c < 3; ++c) g.kdp[c] = g.sF[c] = g.n[c] = g.pw[c] = T(0);
  g.a2 = g.k = g.sT = g.loss = g.loss_g = T(0);
}

// ---------------------------------------------------------------------------------------------
// One light.  Modes:
//   kRender  : forward only, returns the 3 gamma-encoded channels in `out`
//   kVjp     : `io` holds the upstream dL/d out_c (mode B backward)
//   kL2      : `io` holds the target image; accumulates (out - target)^2 and its gradient
// Gradients are accumulated WITHOUT the constant image-gradient factor (see epilogue).
// ---------------------------------------------------------------------------------------------
//   kVjpL2   : `io` holds an upstream dL/d out_c, `tgt` a target image and `l2w` a weight: the upstream is
//              io_c + l2w * (out_c - tgt_c) — an arbitrary image loss plus an L2 term in one pass (the combined
//              loss of materialgan.py:141-147); also accumulates (out - tgt)^2
enum LightMode { kRender = 0, kVjp = 1, kL2 = 2, kVjpL2 = 3 };

template <typename T>
struct LightGeom {   // per light, texture independent
  T cx, cy, cz, cz2; // camera position, cz^2
  T lx, ly, lz, lz2; // light position, lz^2 (unused when co-located)
};

// Radiance -> clamp -> gamma for the 3 channels, and the image gradient gI_c (without its
// constant factor) in the gradient modes.  fp_c = pw_c * f_c, w = n.l / d^2.
// SV_MUFU_LITE trades MUFU operations for multiplications where three reciprocals (or the three gamma slopes)
// can share one MUFU.RCP of a product:  1/a, 1/b, 1/c  from  r = 1/(abc):  1/a = r*(bc), ...
// 13 -> 9 MUFU per pixel.light on the co-located path at the cost of 12 multiplications — a win when the XU
// pipe, not the issue slot, is the limiter (the packed FP32x2 kernel).  All factors are bounded away from zero
// (eps-regularised denominators >= 1e-6, clamped radiance >= 1e-6), so the products stay >= 1e-18.
#ifndef SV_MUFU_LITE
#define SV_MUFU_LITE 0
#endif

// SV_PAIR_RG: in the scalar (one texel per thread) instantiation, the R and G channel chains of a light are evaluated
// together with the packed FP32x2 instructions and B stays scalar: FMA-pipe instructions per light drop by ~16, register
// pairing costs ~10 moves.  Measured (profiles/r01_variants.txt): 78.5 -> 76.2 us per epoch at 1024^2 x 9 lights,
// 1307 -> 1329 us at 2048^2 x 64 (ALU pipe pressure); on, because the 9-light captures are the reference's use case.
#ifndef SV_PAIR_RG
#define SV_PAIR_RG 1
#endif

// SV_OUT_FROM_SLOPE: in the gradient modes the gamma slope Icl^(1/g - 1) is needed anyway, so the encoded value
// Icl^(1/g) is formed as slope * Icl — one multiplication instead of a second MUFU.EX2 per channel: 13 -> 10 MUFU per
// pixel.light on the co-located path.  The extra rounding (one multiplication, and an ex2 argument 1.2x larger) is
// below the lg2/ex2 approximation error; the forward render (kRender) uses the same expression.
#ifndef SV_OUT_FROM_SLOPE
#define SV_OUT_FROM_SLOPE 1
#endif

#if SV_PAIR_RG && defined(__CUDA_ARCH__)
template <int MODE, bool WANT_POW>
SV_D void channels_rg(const float kdp[3], const float Fp[3], float w, float Q, const float io[3], float out[3], Grads<float>& g,
                      float& gw, float& gQ) {
  typedef Fm<float> S;
  typedef Fm<V2> F;
  const V2 QQ(Q), ww(w);
  const V2 FpRG(Fp[0], Fp[1]);
  const V2 fpRG = F::fma(QQ, FpRG, V2(kdp[0], kdp[1]));
  const float fpB = S::fma(Q, Fp[2], kdp[2]);
  const V2 IRG = fpRG * ww;
  const float IB = fpB * w;
  const V2 IclRG(S::min(S::max(IRG.x, float(kEps)), 1.f), S::min(S::max(IRG.y, float(kEps)), 1.f));
  const float IclB = S::min(S::max(IB, float(kEps)), 1.f);
  const V2 lgRG(S::lg2(IclRG.x), S::lg2(IclRG.y));
  const float lgB = S::lg2(IclB);
  const V2 es = lgRG * V2(1.0 / kGamma - 1.0);
  const V2 slopeRG(S::ex2(es.x), S::ex2(es.y));
  const float slopeB = S::ex2(lgB * float(1.0 / kGamma - 1.0));
  V2 upRG;
  float upB;
  if (MODE == kL2) {
#if SV_OUT_FROM_SLOPE
    // Icl^(1/g) = Icl^(1/g - 1) * Icl.  Two scalar FMULs, not one FMUL2: ptxas would fuse the packed product with the
    // subtraction below into an FFMA2 and the L2 forward would no longer reproduce the render bit for bit.
    const V2 oRG(S::mul(slopeRG.x, IclRG.x), S::mul(slopeRG.y, IclRG.y));
    const float oB = S::mul(slopeB, IclB);                     // rounded like the render's (no FMA contraction with the subtraction)
#else
    const V2 eo = lgRG * V2(1.0 / kGamma);
    const V2 oRG(S::ex2(eo.x), S::ex2(eo.y));
    const float oB = S::ex2(lgB * float(1.0 / kGamma));
#endif
    upRG = oRG - V2(io[0], io[1]);
    upB = oB - io[2];
    // R^2 and G^2 accumulate as a packed pair (loss, loss_g), B^2 joins the first lane: 2 issue slots instead of 4
    const V2 l2 = F::fma(upRG, upRG, V2(g.loss, g.loss_g));
    g.loss_g = l2.y;
    g.loss = S::fma(upB, upB, l2.x);
  } else {
    upRG = V2(io[0], io[1]);
    upB = io[2];
  }
  const V2 t = upRG * slopeRG;
  const V2 gIRG(IRG.x == IclRG.x ? t.x : 0.f, IRG.y == IclRG.y ? t.y : 0.f);
  const float gIB = (IB == IclB) ? upB * slopeB : 0.f;
  const V2 gfpRG = gIRG * ww;
  const float gfpB = gIB * w;
  {
    const V2 k2 = V2(g.kdp[0], g.kdp[1]) + gfpRG;
    g.kdp[0] = k2.x; g.kdp[1] = k2.y; g.kdp[2] += gfpB;
  }
  const V2 a = gIRG * fpRG, b = gfpRG * FpRG;
  gw = S::fma(gIB, fpB, a.x + a.y);
  gQ = S::fma(gfpB, Fp[2], b.x + b.y);
  if (WANT_POW) {
    const V2 c = gfpRG * fpRG;
    g.pw[0] += c.x; g.pw[1] += c.y;
    g.pw[2] = S::fma(gfpB, fpB, g.pw[2]);
  }
  out[0] = gfpRG.x; out[1] = gfpRG.y; out[2] = gfpB;
}
#endif

template <typename T, int MODE, bool WANT_POW>
SV_HD void channels(const T fp[3], const T Fp[3], T w, T Q, const T io[3], T out[3], Grads<T>& g, T& gw, T& gQ,
                    const T* tgt = nullptr, T l2w = T(0)) {
  typedef Fm<T> F;
  gw = T(0);
  gQ = T(0);
  T I[3], Icl[3], o[3];
  for (int c = 0; c < 3; ++c) {
    I[c] = fp[c] * w;                                         // microfacet.py:117
    Icl[c] = F::min(F::max(I[c], T(kEps)), T(1));             // microfacet.py:120
  }
  if (MODE == kRender) {
#if SV_OUT_FROM_SLOPE
    // same expression as the gradient modes below, so a rendered target is reproduced bit for bit by the L2 forward
    // (the ground truth is then an exact fixed point of the optimisation, as it is for the reference)
    for (int c = 0; c < 3; ++c) out[c] = F::mul(F::ex2(F::lg2(Icl[c]) * T(1.0 / kGamma - 1.0)), Icl[c]);
#else
    for (int c = 0; c < 3; ++c) out[c] = F::ex2(F::lg2(Icl[c]) * T(1.0 / kGamma));
#endif
    return;
  }
  // d out/d I = (1/gamma) Icl^(1/gamma - 1) = (1/gamma) out/Icl, zero outside the clamp (inclusive edges);
  // the 1/gamma factor is applied in the epilogue
  T slope[3];
#if SV_MUFU_LITE
  {
    const T p01 = Icl[0] * Icl[1];
    const T r = F::rcp(p01 * Icl[2]);
    const T t2 = r * Icl[2];
    const T rI[3] = {t2 * Icl[1], t2 * Icl[0], r * p01};
    for (int c = 0; c < 3; ++c) {
      o[c] = F::ex2(F::lg2(Icl[c]) * T(1.0 / kGamma));
      slope[c] = o[c] * rI[c];
    }
  }
#else
  for (int c = 0; c < 3; ++c) {
    const T lg = F::lg2(Icl[c]);
    slope[c] = F::ex2(lg * T(1.0 / kGamma - 1.0));
#if SV_OUT_FROM_SLOPE
    if (MODE == kL2 || MODE == kVjpL2) o[c] = F::mul(slope[c], Icl[c]);   // Icl^(1/g) = Icl^(1/g - 1) * Icl: a multiplication instead of a MUFU
#else
    if (MODE == kL2 || MODE == kVjpL2) o[c] = F::ex2(lg * T(1.0 / kGamma));
#endif
  }
#endif
  for (int c = 0; c < 3; ++c) {
    T up;
    if (MODE == kL2) {
      const T diff = F::sub(o[c], io[c]);
      g.loss = F::fma(diff, diff, g.loss);
      up = diff;
    } else if (MODE == kVjpL2) {
      const T diff = F::sub(o[c], tgt[c]);
      g.loss = F::fma(diff, diff, g.loss);
      up = F::fma(diff, l2w, io[c]);
    } else {
      up = io[c];
    }
    const T gI = F::sel(F::eq(I[c], Icl[c]), up * slope[c], T(0));
    const T gfp = gI * w;                                     // dL/d fp_c
    g.kdp[c] += gfp;
    gw = F::fma(gI, fp[c], gw);
    gQ = F::fma(gfp, Fp[c], gQ);
    if (WANT_POW) g.pw[c] = F::fma(gfp, fp[c], g.pw[c]);
    out[c] = gfp;                                             // handed back for the Fresnel-albedo gradient
  }
}

// SV_COLOC_V2: the co-located light body rewritten for fewer issued instructions (the kernel is issue-bound, DESIGN.md
// section 3.1): 109 -> ~97 SASS instructions per pixel.light.  Same formulas, different evaluation order:
//   * 1 - c^2 is formed as the cross-product identity (V.V - (n.V)^2)/V.V with ONE rounding of the difference (FMA): the
//     GGX denominator den = c^2 a2 + (1 - c^2) no longer cancels against the error of the MUFU.RSQ normalisation, so the
//     Newton step on rsqrt (4 instructions) is gone — and den is more accurate than the reference's own fp32 `1 - c*c`;
//   * q/4 = c^2 + eps/4 instead of q = 4 c^2 + eps (exact scaling by a power of two): 4/q comes out of the MUFU.RCP
//     for free where the derivative needs it, the 1/4 is folded into the per-texel constant a2/4;
//   * QR = Q/c is formed on the way to Q and reused for the 2Q/c term of dQ/dc;
//   * the material gradients are accumulated as sum T (T = gQ Q), sum T u c^2 / 2 and sum T (1-c)/gv; the per-texel
//     factors 1/a2, -2 are applied once in the epilogue;
//   * the accumulators take the image gradient gI_c directly (fma(gI, w, .), fma(gI, w Q, .)) instead of dL/dfp_c.
#ifndef SV_COLOC_V2
#define SV_COLOC_V2 1
#endif

// SV_PRED_ACC (device, float): the clamp mask of the image gradient predicates the accumulating FMAs instead of
// selecting a zero first (one FSEL per channel less).
#ifndef SV_PRED_ACC
#define SV_PRED_ACC 0
#endif
#ifndef SV_RCP_MERGE
#define SV_RCP_MERGE 0
#endif
// SV_ST_FROM_SF: the per-light accumulation of sum T (co-located body) is replaced by sum_c Fp_c sF_c in the epilogue: one
// instruction per light less (104 -> 103), 70.9 -> 70.2 us at 1024^2 x 9 but 1146 -> 1157 us at 2048^2 x 64 and 4492-4510 ->
// 4530 us at 4096^2 x 64 (the schedule ptxas finds matters more than the count): off
#ifndef SV_ST_FROM_SF
#define SV_ST_FROM_SF 0
#endif
#ifndef SV_NV_FOLD
#define SV_NV_FOLD 0
#endif

template <typename T, int MODE, bool WANT_POW>
SV_HD void channels_coloc(const Texel<T>& tx, T w, T Q, const T io[3], T out[3], Grads<T>& g, T& B, T& gw, const T* tgt, T l2w) {
  typedef Fm<T> F;
  T fp[3], I[3], Icl[3];
  for (int c = 0; c < 3; ++c) {
    fp[c] = F::fma(Q, tx.Fp[c], tx.kdp[c]);                   // pw_c f_c, microfacet.py:102-109
    I[c] = fp[c] * w;                                         // microfacet.py:117
    Icl[c] = F::min(F::max(I[c], T(kEps)), T(1));             // microfacet.py:120
  }
  if (MODE == kRender) {
    for (int c = 0; c < 3; ++c) out[c] = F::mul(F::ex2(F::lg2(Icl[c]) * T(1.0 / kGamma - 1.0)), Icl[c]);
    return;
  }
  const T wQ = w * Q;
  B = T(0);
  gw = T(0);
  for (int c = 0; c < 3; ++c) {
    const T slope = F::ex2(F::lg2(Icl[c]) * T(1.0 / kGamma - 1.0));   // Icl^(1/gamma - 1); 1/gamma applied in the epilogue
    T up;
    if (MODE == kL2) {
      const T diff = F::sub(F::mul(slope, Icl[c]), io[c]);   // Icl^(1/gamma) - target, rounded like the render
      g.loss = F::fma(diff, diff, g.loss);
      up = diff;
    } else if (MODE == kVjpL2) {
      const T diff = F::sub(F::mul(slope, Icl[c]), tgt[c]);
      g.loss = F::fma(diff, diff, g.loss);
      up = F::fma(diff, l2w, io[c]);
    } else {
      up = io[c];
    }
    const T gI = F::sel(F::eq(I[c], Icl[c]), up * slope, T(0));   // clamp masks are inclusive
    g.kdp[c] = F::fma(gI, w, g.kdp[c]);
    g.sF[c] = F::fma(gI, wQ, g.sF[c]);
    B = F::fma(gI, tx.Fp[c], B);
    gw = F::fma(gI, fp[c], gw);
    if (WANT_POW) g.pw[c] = F::fma(gI, I[c], g.pw[c]);       // gfp_c fp_c = gI_c I_c
  }
}

#if SV_PAIR_RG && defined(__CUDA_ARCH__)
// float instantiation on the device: the R and G chains share packed FP32x2 instructions, B stays scalar
template <int MODE, bool WANT_POW>
SV_D void channels_coloc_rg(const Texel<float>& tx, float w, float Q, const float io[3], Grads<float>& g, float& B, float& gw) {
  typedef Fm<float> S;
  typedef Fm<V2> F;
  const V2 QQ(Q), ww(w);
  const V2 FpRG(tx.Fp[0], tx.Fp[1]);
  const V2 fpRG = F::fma(QQ, FpRG, V2(tx.kdp[0], tx.kdp[1]));
  const float fpB = S::fma(Q, tx.Fp[2], tx.kdp[2]);
  const V2 IRG = fpRG * ww;
  const float IB = fpB * w;
  const V2 IclRG(S::min(S::max(IRG.x, float(kEps)), 1.f), S::min(S::max(IRG.y, float(kEps)), 1.f));
  const float IclB = S::min(S::max(IB, float(kEps)), 1.f);
  const V2 es = V2(S::lg2(IclRG.x), S::lg2(IclRG.y)) * V2(1.0 / kGamma - 1.0);
  const V2 slopeRG(S::ex2(es.x), S::ex2(es.y));
  const float slopeB = S::ex2(S::lg2(IclB) * float(1.0 / kGamma - 1.0));
  V2 upRG;
  float upB;
  if (MODE == kL2) {
    // two scalar FMULs, not one FMUL2: ptxas would fuse the packed product with the subtraction into an FFMA2 and the L2
    // forward would no longer reproduce the render bit for bit
    const V2 oRG(S::mul(slopeRG.x, IclRG.x), S::mul(slopeRG.y, IclRG.y));
    const float oB = S::mul(slopeB, IclB);
    upRG = oRG - V2(io[0], io[1]);
    upB = oB - io[2];
    const V2 l2 = F::fma(upRG, upRG, V2(g.loss, g.loss_g));
    g.loss_g = l2.y;
    g.loss = S::fma(upB, upB, l2.x);
  } else {
    upRG = V2(io[0], io[1]);
    upB = io[2];
  }
  const V2 t = upRG * slopeRG;
  const float tB = upB * slopeB;
  const float wQ = w * Q;
#if SV_PRED_ACC
  {
    // channel R selects (it initialises B and gw), G and B predicate their accumulations on the clamp mask
    const float gIR = (IRG.x == IclRG.x) ? t.x : 0.f;
    g.kdp[0] = S::fma(gIR, w, g.kdp[0]);
    g.sF[0] = S::fma(gIR, wQ, g.sF[0]);
    B = gIR * tx.Fp[0];
    gw = gIR * fpRG.x;
    if (WANT_POW) g.pw[0] = S::fma(gIR, IRG.x, g.pw[0]);
    asm("{\n .reg .pred p;\n setp.eq.f32 p, %4, %5;\n @p fma.rn.f32 %0, %6, %7, %0;\n @p fma.rn.f32 %1, %6, %8, %1;\n"
        " @p fma.rn.f32 %2, %6, %9, %2;\n @p fma.rn.f32 %3, %6, %10, %3;\n}"
        : "+f"(g.kdp[1]), "+f"(g.sF[1]), "+f"(B), "+f"(gw)
        : "f"(IRG.y), "f"(IclRG.y), "f"(t.y), "f"(w), "f"(wQ), "f"(tx.Fp[1]), "f"(fpRG.y));
    asm("{\n .reg .pred p;\n setp.eq.f32 p, %4, %5;\n @p fma.rn.f32 %0, %6, %7, %0;\n @p fma.rn.f32 %1, %6, %8, %1;\n"
        " @p fma.rn.f32 %2, %6, %9, %2;\n @p fma.rn.f32 %3, %6, %10, %3;\n}"
        : "+f"(g.kdp[2]), "+f"(g.sF[2]), "+f"(B), "+f"(gw)
        : "f"(IB), "f"(IclB), "f"(tB), "f"(w), "f"(wQ), "f"(tx.Fp[2]), "f"(fpB));
    if (WANT_POW) {
      g.pw[1] = S::fma((IRG.y == IclRG.y) ? t.y : 0.f, IRG.y, g.pw[1]);
      g.pw[2] = S::fma((IB == IclB) ? tB : 0.f, IB, g.pw[2]);
    }
  }
#else
  const V2 gIRG(IRG.x == IclRG.x ? t.x : 0.f, IRG.y == IclRG.y ? t.y : 0.f);
  const float gIB = (IB == IclB) ? tB : 0.f;
  {
    const V2 k2 = F::fma(gIRG, ww, V2(g.kdp[0], g.kdp[1]));
    g.kdp[0] = k2.x; g.kdp[1] = k2.y;
    g.kdp[2] = S::fma(gIB, w, g.kdp[2]);
    const V2 s2 = F::fma(gIRG, V2(wQ), V2(g.sF[0], g.sF[1]));
    g.sF[0] = s2.x; g.sF[1] = s2.y;
    g.sF[2] = S::fma(gIB, wQ, g.sF[2]);
  }
  const V2 a = gIRG * fpRG, b = gIRG * FpRG;
  gw = S::fma(gIB, fpB, a.x + a.y);
  B = S::fma(gIB, tx.Fp[2], b.x + b.y);
  if (WANT_POW) {
    const V2 c = gIRG * IRG;
    g.pw[0] += c.x; g.pw[1] += c.y;
    g.pw[2] = S::fma(gIB, IB, g.pw[2]);
  }
#endif
}
#endif

#if SV_COLOC_V2
template <typename T, int MODE, bool WANT_POW>
SV_HD void shade_light_coloc(const Texel<T>& tx, const LightGeom<T>& lg, const T io[3], T out[3], Grads<T>& g,
                             const T* tgt = nullptr, T l2w = T(0)) {
  typedef Fm<T> F;
  const T Vx = lg.cx - tx.px, Vy = lg.cy - tx.py, Vz = lg.cz;
  const T vv = F::fma(Vx, Vx, F::fma(Vy, Vy, lg.cz2));        // |V|^2, microfacet.py:60-62
  const T rv = F::rsqrt(vv);
#if SV_NV_FOLD
  const T nV = F::fma(tx.n[0], lg.cx, F::fma(tx.n[1], lg.cy, F::fma(tx.n[2], lg.cz, tx.mnP)));   // n.C - n.P: three FMAs
#else
  const T nV = F::fma(tx.n[0], Vx, F::fma(tx.n[1], Vy, tx.n[2] * Vz));
#endif
  const T c = F::max(nV * rv, T(0));                          // n.v = n.l = n.h, clamp(min=0) (microfacet.py:96-98)
  const T inv_d2 = rv * rv;
  const T w = c * inv_d2;                                     // n.l / d^2
  const T c2 = c * c;
  const T s2 = F::fma(-nV, nV, vv) * inv_d2;                  // 1 - (n.v)^2 without cancellation error (|n| = 1)
  const T den = F::fma(c2, tx.a2, s2);                        // c2 a2 + 1 - c2, microfacet.py:30
  const T pden = T(kPi) * den;
  const T Dd = F::fma(pden, den, T(kEps));                    // microfacet.py:31
  const T gv = F::fma(c, tx.omk, tx.k);                       // microfacet.py:49
  const T q4 = c2 + T(0.25 * kEps);                           // (4 n.v n.l + eps)/4, microfacet.py:109
#if SV_RCP_MERGE
  // one reciprocal of the product Dd gv^2 q/4 instead of three (8 instead of 10 MUFU per pixel.light); the backward
  // half rebuilds T/Dd, T/gv, T/q from T R and the partial products (4 more multiplications).  Dd gv^2 q/4 >= 2.5e-25.
  const T g2 = gv * gv;
  const T A = Dd * g2;
  const T R = F::rcp(A * q4);
  const T Rc = c * R;                                         // 4 Q / (a2 c)
#else
  const T rDd = F::rcp(Dd), rgv = F::rcp(gv), rq4 = F::rcp(q4);
  const T Rc = c * rDd * (rgv * rgv) * rq4;                   // 4 Q / (a2 c)
#endif
  const T QR = Rc * tx.a2q;                                   // Q / c
  const T Q = QR * c;                                         // D G / q
  T B, gw;
#if SV_PAIR_RG && defined(__CUDA_ARCH__)
  if (sizeof(T) == 4 && (MODE == kVjp || MODE == kL2)) {
    channels_coloc_rg<MODE, WANT_POW>(*reinterpret_cast<const Texel<float>*>(&tx), *reinterpret_cast<const float*>(&w),
                                      *reinterpret_cast<const float*>(&Q), reinterpret_cast<const float*>(io),
                                      *reinterpret_cast<Grads<float>*>(&g), *reinterpret_cast<float*>(&B), *reinterpret_cast<float*>(&gw));
  } else
#endif
  {
    channels_coloc<T, MODE, WANT_POW>(tx, w, Q, io, out, g, B, gw, tgt, l2w);
  }
  if (MODE == kRender) return;

  const T gQ = w * B;                                         // dL/dQ = sum_c dL/dfp_c Fp_c
  const T Tq = gQ * Q;
#if !SV_ST_FROM_SF
  g.sT += Tq;
#endif
#if SV_RCP_MERGE
  const T TR = Tq * R;
  const T E = TR * (g2 * q4 * pden);                          // T pden / Dd
  const T Ggv = TR * (Dd * q4 * gv);                          // T / gv
  const T Tq4 = TR * A;                                       // T / (q/4)
#else
  const T E = Tq * (rDd * pden);                              // T u/2, u = (dDd/dden)/Dd
  const T Ggv = Tq * rgv;
  const T Tq4 = Tq * rq4;
#endif
  g.a2 = F::fma(E, c2, g.a2);                                 // SA
  g.k = F::fma(Ggv, T(1) - c, g.k);                           // SK
  // dQ/dc = 2 Q/c - 2 Q [ c (u (a2-1) + 4/q) + (1-k)/gv ];  w = c/d2
  const T X = F::fma(E, tx.a2m1x2, Tq4);
  const T Y = F::fma(c, X, F::fma(Ggv, tx.omk, -(gQ * QR)));
  const T gc = F::fma(Y, T(-2), gw * inv_d2);
  // clamp(min=0): below the horizon c = 0 -> I = 0 -> clamped to eps -> gI = 0 -> gc = 0 already
  const T cV = gc * rv;
  g.n[0] = F::fma(cV, Vx, g.n[0]);
  g.n[1] = F::fma(cV, Vy, g.n[1]);
  g.n[2] = F::fma(cV, Vz, g.n[2]);
}
#else
// Co-located light and camera (everything the reference's capture emits): l = v = h, v.h = 1,
// n.v = n.l = n.h = c.  The specular lobe collapses to a function of c alone,
//     Q = D G / q = a2 c^2 / (Dd gv^2 q),   Dd = pi den^2 + eps, den = c2 a2 + 1 - c2,
//     gv = c (1-k) + k,  q = 4 c^2 + eps,
// and its gradient is taken through d ln Q.
template <typename T, int MODE, bool WANT_POW>
SV_HD void shade_light_coloc(const Texel<T>& tx, const LightGeom<T>& lg, const T io[3], T out[3], Grads<T>& g,
                             const T* tgt = nullptr, T l2w = T(0)) {
  typedef Fm<T> F;
  const T Vx = lg.cx - tx.px, Vy = lg.cy - tx.py, Vz = lg.cz;
  const T vv = F::fma(Vx, Vx, F::fma(Vy, Vy, lg.cz2));
  const T rv = rsqrt_dir(vv);
  const T c_raw = F::fma(tx.n[0], Vx, F::fma(tx.n[1], Vy, tx.n[2] * Vz)) * rv;
  const T c = F::max(c_raw, T(0));
  const T inv_d2 = rv * rv;
  const T w = c * inv_d2;
  const T c2 = c * c;
  const T den = F::fma(c2, tx.a2, T(1) - c2);
  const T Dd = F::fma(T(kPi) * den, den, T(kEps));
  const T gv = F::fma(c, tx.omk, tx.k);
  const T q = F::fma(c2, T(4), T(kEps));
#if SV_MUFU_LITE
  const T gq = gv * q;
  const T rP = F::rcp(Dd * gq);
  const T rDd = rP * gq;
  const T tP = rP * Dd;
  const T rgv = tP * q, rq = tP * gv;
#else
  const T rDd = F::rcp(Dd), rgv = F::rcp(gv), rq = F::rcp(q);
#endif
  const T Rc = c * rDd * (rgv * rgv) * rq;                    // Q / (a2 c)
  const T R0 = Rc * c;                                        // Q / a2
  const T Q = tx.a2 * R0;
  T fp[3], gfp[3], gw, gQ;
#if SV_PAIR_RG && defined(__CUDA_ARCH__)
  if (sizeof(T) == 4 && (MODE == kVjp || MODE == kL2)) {
    channels_rg<MODE, WANT_POW>(reinterpret_cast<const float*>(tx.kdp), reinterpret_cast<const float*>(tx.Fp), *reinterpret_cast<const float*>(&w),
                                *reinterpret_cast<const float*>(&Q), reinterpret_cast<const float*>(io), reinterpret_cast<float*>(gfp),
                                *reinterpret_cast<Grads<float>*>(&g), *reinterpret_cast<float*>(&gw), *reinterpret_cast<float*>(&gQ));
  } else
#endif
  {
    for (int ch = 0; ch < 3; ++ch) fp[ch] = F::fma(Q, tx.Fp[ch], tx.kdp[ch]);
    channels<T, MODE, WANT_POW>(fp, tx.Fp, w, Q, io, MODE == kRender ? out : gfp, g, gw, gQ, tgt, l2w);
  }
  if (MODE == kRender) return;

  for (int ch = 0; ch < 3; ++ch) g.sF[ch] = F::fma(gfp[ch], Q, g.sF[ch]);
  const T Tq = gQ * Q;
  const T u = rDd * (T(2.0 * kPi) * den);                     // (dDd/dden)/Dd
  g.a2 = F::fma(gQ, R0, F::fma(-(Tq * u), c2, g.a2));         // dQ/da2 = Q/a2 - Q u c2
  const T m2T = T(-2) * Tq;
  g.k = F::fma(m2T * (T(1) - c), rgv, g.k);                   // dQ/dk  = -2 Q (1-c)/gv
  // dQ/dc = 2 Q/c - 2 Q [ c (u (a2-1) + 4/q) + (1-k)/gv ];  w = c/d2
  const T s1 = F::fma(u, tx.a2 - T(1), T(4) * rq);
  const T S = F::fma(c, s1, tx.omk * rgv);
  const T gc = F::fma(gQ + gQ, tx.a2 * Rc, F::fma(m2T, S, gw * inv_d2));
  // clamp(min=0): below the horizon c = 0 -> I = 0 -> clamped to eps -> gI = 0 -> gc = 0 already
  const T cV = gc * rv;
  g.n[0] = F::fma(cV, Vx, g.n[0]);
  g.n[1] = F::fma(cV, Vy, g.n[1]);
  g.n[2] = F::fma(cV, Vz, g.n[2]);
}

#endif  // SV_COLOC_V2

// General light/camera pair.
template <typename T, int MODE, bool WANT_POW>
SV_HD void shade_light_general(const Texel<T>& tx, const LightGeom<T>& lg, const T io[3], T out[3], Grads<T>& g,
                               const T* tgt = nullptr, T l2w = T(0)) {
  typedef Fm<T> F;
  // --- geometry (microfacet.py:91-99) ---
  const T Vx = lg.cx - tx.px, Vy = lg.cy - tx.py, Vz = lg.cz;
  const T rv = rsqrt_dir(F::fma(Vx, Vx, F::fma(Vy, Vy, lg.cz2)));
  const T ndv_raw = F::fma(tx.n[0], Vx, F::fma(tx.n[1], Vy, tx.n[2] * Vz)) * rv;
  const T Lx = lg.lx - tx.px, Ly = lg.ly - tx.py, Lz = lg.lz;
  const T rl = rsqrt_dir(F::fma(Lx, Lx, F::fma(Ly, Ly, lg.lz2)));
  const T inv_d2 = rl * rl;
  const T ndl_raw = F::fma(tx.n[0], Lx, F::fma(tx.n[1], Ly, tx.n[2] * Lz)) * rl;
  const T lv = F::fma(Lx, Vx, F::fma(Ly, Vy, Lz * Vz)) * (rl * rv);
  const T opl = T(1) + lv;                                    // |l+v|^2 / 2
  const T rh = rsqrt_dir(opl + opl);
  const T ndh_raw = (ndl_raw + ndv_raw) * rh;
  const T vdh = F::max(opl * rh, T(0));
  const T ndv = F::max(ndv_raw, T(0));
  const T ndl = F::max(ndl_raw, T(0));
  const T ndh = F::max(ndh_raw, T(0));
  // --- GGX (microfacet.py:28-32) ---
  const T c2 = ndh * ndh;
  const T den = F::fma(c2, tx.a2, T(1) - c2);
  const T Dd = F::fma(T(kPi) * den, den, T(kEps));
  const T rDd = F::rcp(Dd);
  const T D = tx.a2 * rDd;
  // --- Fresnel (microfacet.py:43-45) ---
  const T sphg = F::ex2(F::fma(T(kFresA), vdh, T(kFresB)) * vdh);
  // --- Smith (microfacet.py:47-52) ---
  const T rgv = F::rcp(F::fma(ndv, tx.omk, tx.k));
  const T rgl = F::rcp(F::fma(ndl, tx.omk, tx.k));
  const T A = ndv * rgv, B = ndl * rgl;
  const T G = A * B;
  // --- specular lobe (microfacet.py:109) ---
  const T rq = F::rcp(F::fma(T(4) * ndv, ndl, T(kEps)));
  const T DGq = D * rq;
  const T Q = DGq * G;
  const T w = ndl * inv_d2;
  T fp[3], Fp[3], gfp[3], gw, gQ;
  for (int ch = 0; ch < 3; ++ch) {
    Fp[ch] = F::fma(tx.omsp[ch], sphg, tx.sp[ch]);
    fp[ch] = F::fma(Q, Fp[ch], tx.kdp[ch]);
  }
  channels<T, MODE, WANT_POW>(fp, Fp, w, Q, io, MODE == kRender ? out : gfp, g, gw, gQ, tgt, l2w);
  if (MODE == kRender) return;

  const T QomS = Q * (T(1) - sphg);
  for (int ch = 0; ch < 3; ++ch) g.sF[ch] = F::fma(gfp[ch], QomS, g.sF[ch]);
  // Q = D*G/q
  const T gD = gQ * G * rq;
  const T gG = gQ * DGq;
  const T gq = -gQ * Q * rq;
  // D = a2/Dd, Dd = pi*den^2 + eps, den = c2*a2 + 1 - c2
  const T gden = -gD * D * rDd * T(2.0 * kPi) * den;
  g.a2 += F::fma(gden, c2, gD * rDd);
  const T gndh = gden * (tx.a2 - T(1)) * (ndh + ndh);
  // G = A*B, A = ndv/gv, B = ndl/gl
  const T rgv2 = rgv * rgv, rgl2 = rgl * rgl;
  const T gA = gG * B, gB = gG * A;
  const T gndv = F::fma(gA * tx.k, rgv2, gq * T(4) * ndl);
  const T gndl = F::fma(gB * tx.k, rgl2, F::fma(gq * T(4), ndv, gw * inv_d2));
  g.k -= F::fma(gA * ndv * (T(1) - ndv), rgv2, gB * ndl * (T(1) - ndl) * rgl2);
  // clamp(min=0) masks are inclusive; n.h = (n.l + n.v) * rh
  const T mv = F::sel(F::ge(ndv_raw, T(0)), gndv, T(0));
  const T ml = F::sel(F::ge(ndl_raw, T(0)), gndl, T(0));
  const T mh = F::sel(F::ge(ndh_raw, T(0)), gndh * rh, T(0));
  const T cV = (mv + mh) * rv;
  const T cL = (ml + mh) * rl;
  g.n[0] = F::fma(cV, Vx, F::fma(cL, Lx, g.n[0]));
  g.n[1] = F::fma(cV, Vy, F::fma(cL, Ly, g.n[1]));
  g.n[2] = F::fma(cV, Vz, F::fma(cL, Lz, g.n[2]));
}

template <typename T, int MODE, bool COLOC, bool WANT_POW>
SV_HD void shade_light(const Texel<T>& tx, const LightGeom<T>& lg, const T io[3], T out[3], Grads<T>& g,
                       const T* tgt = nullptr, T l2w = T(0)) {
  if (COLOC) shade_light_coloc<T, MODE, WANT_POW>(tx, lg, io, out, g, tgt, l2w);
  else shade_light_general<T, MODE, WANT_POW>(tx, lg, io, out, g, tgt, l2w);
}

// ---------------------------------------------------------------------------------------------
// Epilogue: gradients w.r.t. material parameters -> gradients w.r.t. the 9 channels.
// `scale` is the constant image-gradient factor (2/(N*3*H*W*gamma) for L2, 1/gamma for VJP).
// `outer` = mask of the caller's clamp(-1,1) on the raw parameter (svbrdf.py:60); pass all-true
// when the clamp belongs to the caller's graph (mode B).
// ---------------------------------------------------------------------------------------------
template <typename T, bool COLOC>
SV_HD void texel_epilogue(const Texel<T>& tx, const TexelAux<T>& ax, const T pw[3], const Grads<T>& g, T scale,
                          const typename Fm<T>::mask outer[9], T gt[9]) {
  typedef Fm<T> F;
  const T inv_pi = T(1.0 / kPi);
  const T sf = COLOC ? T(1.0 - kSphgColoc) : T(1);
  for (int c = 0; c < 3; ++c) {
    const T gkd = g.kdp[c] * pw[c];                           // dL/d (d_c/pi (1-s_c))
    const T gd = gkd * ax.oms[c] * inv_pi;
    const T gs = F::fma(-gkd, ax.d[c] * inv_pi, g.sF[c] * (pw[c] * sf));
    gt[c] = gd * ax.dpow[c];
    gt[6 + c] = gs * ax.dpow[4 + c];
  }
  // a2 = alpha^2, k = alpha/2 + eps, alpha = rough^2
  T ga2 = g.a2, gk = g.k;
#if SV_COLOC_V2
  if (COLOC) {
    // dL/da2 = sT/a2 - 2 SA;  dL/dk = -2 SK.  Below a2 = 1e-30 the sT/a2 term is dropped: MUFU.RCP flushes a subnormal a2
    // to zero (-> inf, and a -inf roughness gradient), and at a2 = 0 the reference's dD/da2 = 1/Dd is finite but rough = 0
    // zeroes the channel's chain factor.  Dropping it is exact to fp32: Q/a2 = c^2/(Dd gv^2 q) <= 1e12, and the chain
    // factor of the roughness channel, 4 rough^3 dpow = 4 a2^(3/4) dpow, is < 2e-22 there.
#if SV_ST_FROM_SF
    // sum_l T_l = sum_l w_l Q_l sum_c gI_c Fp_c = sum_c Fp_c sF_c: the accumulators the Fresnel-albedo gradient needs anyway
    const T sT = F::fma(tx.Fp[0], g.sF[0], F::fma(tx.Fp[1], g.sF[1], tx.Fp[2] * g.sF[2]));
#else
    const T sT = g.sT;
#endif
    ga2 = F::fma(g.a2, T(-2), F::sel(F::ge(T(1e-30), tx.a2), T(0), sT * F::rcp(tx.a2)));
    gk = g.k * T(-2);
  }
#endif
  const T galpha = F::fma(ga2, ax.alpha + ax.alpha, gk * T(0.5));
  gt[5] = galpha * (ax.rough + ax.rough) * ax.dpow[3];
  // n = m/|m|
  const T ndg = F::fma(tx.n[0], g.n[0], F::fma(tx.n[1], g.n[1], tx.n[2] * g.n[2]));
  const T gmx = (g.n[0] - tx.n[0] * ndg) * ax.rlen;
  const T gmy = (g.n[1] - tx.n[1] * ndg) * ax.rlen;
  const T gmz = (g.n[2] - tx.n[2] * ndg) * ax.rlen;
  // mz = sqrt(1 - clamp(mx^2+my^2, 0, 1-eps))
  const T gplanar = F::sel(ax.planar_free, -gmz * T(0.5) * F::rcp(ax.mz), T(0));
  gt[3] = F::sel(ax.in3, F::fma(gplanar, ax.mx + ax.mx, gmx), T(0));
  gt[4] = F::sel(ax.in4, F::fma(gplanar, ax.my + ax.my, gmy), T(0));
  for (int kk = 0; kk < 9; ++kk) gt[kk] = F::sel(outer[kk], gt[kk] * scale, T(0));
}

// dL/d light_pow_c from the accumulated sum of gfp_c * fp_c (= pw_c * dL/dpw_c).
// pw_c == 0 returns 0, which is also the reference's value: every sample of that channel is I = 0, the clamp of
// microfacet.py:120 replaces it by eps, and autograd's clamp mask (eps <= I <= 1) blocks the gradient — so
// dL/dpw_c = sum gI_c f_c w = 0 there too, and Adam leaves a zero light-power channel where it is in both.
template <typename T>
SV_HD T pow_grad(T acc, T pw) { return pw != T(0) ? acc / pw : T(0); }   // scalar only (finalisation)

// ---------------------------------------------------------------------------------------------
// x / b with IEEE rounding for a divisor that is fixed for the whole launch (the per-channel std of Normalize): the
// correctly rounded reciprocal r = RN(1/b) is formed once, then q0 = RN(x r), rem = x - q0 b (exact in an FMA),
// q = RN(q0 + rem r) is the correctly rounded quotient (Markstein) — three FMA-pipe instructions instead of the ~10 plus a
// MUFU.RCP of the general division.  tests/test_host_logic.py runs this very code (host build) against IEEE division on
// millions of (x, b) pairs incl. all-ones significands (the one difference: a -0.0 numerator gives +0.0);
// tests/test_gpu_features.py compares the normalised image with torch's.  Launches whose divisors lie outside [1e-30, 1e30] (reciprocal or quotient could leave the normal range) keep
// the general division (norm_l2_kernel).
// ---------------------------------------------------------------------------------------------
struct FixedDiv {
  float b, r;
};
SV_HD FixedDiv make_fixed_div(float b) {
  FixedDiv d;
  d.b = b;
#if defined(__CUDA_ARCH__)
  d.r = __fdiv_rn(1.f, b);
#else
  d.r = 1.0f / b;
#endif
  return d;
}
SV_HD float div_rn(float x, const FixedDiv& d) {
  typedef Fm<float> S;
  const float q0 = S::mul(x, d.r);
  const float rem = S::fma(-q0, d.b, x);
  return S::fma(rem, d.r, q0);
}

// ---------------------------------------------------------------------------------------------
// Adam (torch/optim/adam.py:531-547, single-tensor path, amsgrad off, weight decay 0).
//   m <- m + (g - m)(1 - b1);  v <- v*b2 + (1 - b2) g^2;
//   p <- p - step_size * m / (sqrt(v)/sqrt(bc2) + eps),  step_size = lr/bc1
// The host passes step_size and 1/sqrt(bc2) computed in double like torch does.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct AdamStep {
  T one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps;
};

template <typename T>
SV_HD void adam_update(T& p, T& m, T& v, T g, const AdamStep<T>& a) {
  typedef Fm<T> F;
  m = F::fma(g - m, a.one_minus_b1, m);
  v = F::fma(a.one_minus_b2 * g, g, v * a.b2);
  const T denom = F::fma(F::sqrt(v), a.inv_sqrt_bc2, a.eps);
  p = F::fma(-a.step_size * m, F::rcp(denom), p);
}

}  // namespace svbrdf
