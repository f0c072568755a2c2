// sm_100a kernels + C ABI (include/svbrdf_b200.h) of the per-pixel SVBRDF optimisation path.
//
// One thread owns one texel for the whole pass: it loads the texel's 9 channels once
// (coalesced: a warp reads 128 contiguous bytes of each plane), keeps the material
// parameters and the gradient accumulators in registers, loops over all lights, and
// finishes with the texture-gradient epilogue and — in the fused mode — the Adam update.
// The rendered image is never written in the L2 modes; nothing image-sized is saved
// between forward and backward (the backward kernel recomputes the forward per light).
//
// Per-light geometry (2 x float4 per light) is staged in shared memory once per CTA and
// read with broadcast LDS.128.  Whether every light is co-located with its camera is
// detected while staging; co-located captures (everything the reference's capture code
// emits, capture.py:70-71) take a loop body with l = v = h folded.
//
// The arithmetic is elementwise and transcendental (MUFU) bound: no tensor cores.
#include <cuda_runtime.h>

#include "../../include/svbrdf_b200.h"
#include "svbrdf_core.cuh"

namespace svbrdf {

constexpr int kThreads = 256;
constexpr int kPrefetch = 2;   // lights of target/grad_out data in flight per thread

enum KernelMode { kModeRender = 0, kModeVjp = 1, kModeL2Grad = 2, kModeL2Adam = 3 };

struct Params {
  float* tex;            // [9] planes (read-only except kModeL2Adam)
  float* m;              // Adam first moment  (kModeL2Adam)
  float* v;              // Adam second moment (kModeL2Adam)
  const float* cam;      // [N,3]
  const float* light;    // [N,3]
  const float* pow;      // [3]
  const void* io;        // target (L2 modes) or grad_out (VJP): [N,3] planes
  float* out;            // rendered image [N,3] planes (render) or grad_tex [9] planes
  float* partials;       // [gridDim.x][4]: loss, gpow[3]
  long long stride;      // plane stride in elements
  long long texels;      // rows * res
  float size;
  int res;
  int row_offset;
  int n_lights;
  float scale;           // constant image-gradient factor
  AdamStep<float> adam;
};

template <int TGT>
struct IoLoad;
template <>
struct IoLoad<SVBRDF_TARGET_F32> {
  static __device__ __forceinline__ float at(const void* base, long long idx) {
    return __ldg(static_cast<const float*>(base) + idx);
  }
};
template <>
struct IoLoad<SVBRDF_TARGET_U8> {
  // float(b)/255 correctly rounded (bit-identical to the IEEE division of imageio.py:18-19):
  // q = b*r, one Newton correction with the exact residual.
  static __device__ __forceinline__ float at(const void* base, long long idx) {
    const float b = float(__ldg(static_cast<const unsigned char*>(base) + idx));
    const float r = 1.0f / 255.0f;
    const float q = b * r;
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, b), r, q);
  }
};

template <int MODE, bool COLOC, bool WANT_POW, int TGT>
__device__ __forceinline__ void light_loop(const Params& P, const float4* __restrict__ s_geo, const Texel<float>& tx,
                                            const float pw[3], long long p, bool valid, Grads<float>& g) {
  const int N = P.n_lights;
  const long long stride = P.stride;
  constexpr int LM = (MODE == kModeRender) ? kRender : (MODE == kModeVjp ? kVjp : kL2);

  float buf[kPrefetch][3];
  if (MODE != kModeRender) {
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        buf[j][c] = (valid && j < N) ? IoLoad<TGT>::at(P.io, (long long)(j * 3 + c) * stride + p) : 0.f;
    }
  }
  for (int i0 = 0; i0 < N; i0 += kPrefetch) {
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j) {
      const int i = i0 + j;
      if (i >= N) break;
      float in3[3] = {0.f, 0.f, 0.f}, o3[3];
      if (MODE != kModeRender) {
#pragma unroll
        for (int c = 0; c < 3; ++c) in3[c] = buf[j][c];
        const int nx = i + kPrefetch;
        if (nx < N && valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) buf[j][c] = IoLoad<TGT>::at(P.io, ((long long)nx * 3 + c) * stride + p);
        }
      }
      LightGeom<float> lg;
      const float4 a = s_geo[2 * i];
      lg.cx = a.x; lg.cy = a.y; lg.cz = a.z; lg.cz2 = a.w;
      if (!COLOC) {
        const float4 b = s_geo[2 * i + 1];
        lg.lx = b.x; lg.ly = b.y; lg.lz = b.z; lg.lz2 = b.w;
      }
      shade_light<float, LM, COLOC, WANT_POW>(tx, lg, pw, in3, o3, g);
      if (MODE == kModeRender && valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) P.out[((long long)i * 3 + c) * stride + p] = o3[c];
      }
    }
  }
}

template <int MODE, bool WANT_POW, int TGT>
__global__ void __launch_bounds__(kThreads) texel_kernel(const Params P) {
  extern __shared__ float4 s_geo[];
  __shared__ float s_red[kThreads / 32][4];

  // ---- stage per-light geometry, detect co-location ----
  int same = 1;
  for (int i = threadIdx.x; i < P.n_lights; i += kThreads) {
    const float cx = P.cam[3 * i], cy = P.cam[3 * i + 1], cz = P.cam[3 * i + 2];
    const float lx = P.light[3 * i], ly = P.light[3 * i + 1], lz = P.light[3 * i + 2];
    s_geo[2 * i] = make_float4(cx, cy, cz, cz * cz);
    s_geo[2 * i + 1] = make_float4(lx, ly, lz, lz * lz);
    same &= (cx == lx) & (cy == ly) & (cz == lz);
  }
  const bool coloc = __syncthreads_and(same) != 0;

  const long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool valid = p < P.texels;
  const long long pc = valid ? p : 0;

  float pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pw[c] = __ldg(P.pow + c);

  // ---- texel prologue ----
  float raw[9], t[9];
  bool outer[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    raw[k] = (MODE == kModeL2Adam) ? P.tex[k * P.stride + pc] : __ldg(P.tex + k * P.stride + pc);
    if (MODE == kModeL2Grad || MODE == kModeL2Adam) {   // the clamp of svbrdf.py:60
      outer[k] = raw[k] >= -1.f && raw[k] <= 1.f;
      t[k] = fminf(fmaxf(raw[k], -1.f), 1.f);
    } else {
      outer[k] = true;
      t[k] = raw[k];
    }
  }
  Texel<float> tx;
  TexelAux<float> ax;
  {
    const int row = int(pc / P.res);
    const int col = int(pc - (long long)row * P.res);
    texel_position(row + P.row_offset, col, P.res, P.size, tx.px, tx.py);
  }
  texel_prologue(t, tx, ax);
  Grads<float> g;
  grads_zero(g);

  // ---- all lights ----
  if (coloc) light_loop<MODE, true, WANT_POW, TGT>(P, s_geo, tx, pw, pc, valid, g);
  else light_loop<MODE, false, WANT_POW, TGT>(P, s_geo, tx, pw, pc, valid, g);
  if (MODE == kModeRender) return;

  // ---- epilogue ----
  float gt[9];
  texel_epilogue(tx, ax, g, P.scale, outer, gt);
  if (valid) {
    if (MODE == kModeL2Adam) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const long long idx = k * P.stride + p;
        float mk = P.m[idx], vk = P.v[idx], pk = raw[k];
        adam_update(pk, mk, vk, gt[k], P.adam);
        P.tex[idx] = pk;
        P.m[idx] = mk;
        P.v[idx] = vk;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) P.out[k * P.stride + p] = gt[k];
    }
  }

  // ---- block partials: loss and light-power gradient (fixed order => deterministic) ----
  constexpr bool kHasLoss = (MODE == kModeL2Grad || MODE == kModeL2Adam);
  if (kHasLoss || WANT_POW) {
    float r[4] = {valid ? g.loss : 0.f, valid ? g.pw[0] : 0.f, valid ? g.pw[1] : 0.f, valid ? g.pw[2] : 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c == 0 ? kHasLoss : WANT_POW) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s_red[warp][c] = r[c];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) acc += s_red[w][threadIdx.x];
      P.partials[(long long)blockIdx.x * 4 + threadIdx.x] = acc;
    }
  }
}

// Sums the block partials in double (fixed tree => run-to-run deterministic), writes the loss
// and the light-power gradient, and — for optim_light — applies Adam to light_pow[3].
struct FinalizeParams {
  const float* partials;
  int n_blocks;
  double loss_norm;       // 1/(n_total*3*res*res)
  float grad_scale;       // constant image-gradient factor for gpow
  float* loss_out;        // nullable
  float* grad_pow;        // nullable [3]
  float* pow;             // nullable: light_pow to update (optim_light)
  float* pow_state;       // m[3], v[3]
  AdamStep<float> adam;
};

__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeParams F) {
  __shared__ double s[256][4];
  double acc[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < F.n_blocks; b += 256) {
    const float4 q = reinterpret_cast<const float4*>(F.partials)[b];
    acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
  }
  for (int c = 0; c < 4; ++c) s[threadIdx.x][c] = acc[c];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int c = 0; c < 4; ++c) s[threadIdx.x][c] += s[threadIdx.x + w][c];
    __syncthreads();
  }
  if (threadIdx.x == 0 && F.loss_out) F.loss_out[0] = float(s[0][0] * F.loss_norm);
  if (threadIdx.x < 3) {
    const float gp = float(s[0][1 + threadIdx.x] * double(F.grad_scale));
    if (F.grad_pow) F.grad_pow[threadIdx.x] = gp;
    if (F.pow) {
      float p = F.pow[threadIdx.x], m = F.pow_state[threadIdx.x], v = F.pow_state[3 + threadIdx.x];
      // light_pow has 3 elements: IEEE sqrt/div here, cost is nil
      m = m + (gp - m) * F.adam.one_minus_b1;
      v = v * F.adam.b2 + F.adam.one_minus_b2 * gp * gp;
      const float denom = sqrtf(v) * F.adam.inv_sqrt_bc2 + F.adam.eps;
      p = p - F.adam.step_size * m / denom;
      F.pow[threadIdx.x] = p;
      F.pow_state[threadIdx.x] = m;
      F.pow_state[3 + threadIdx.x] = v;
    }
  }
}

__global__ void __launch_bounds__(256) adam_apply_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ g, size_t n, const AdamStep<float> a) {
  const size_t n4 = n / 4;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nth = size_t(gridDim.x) * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  size_t done = 0;
  if (aligned) {
    for (size_t i = tid; i < n4; i += nth) {
      float4 P4 = reinterpret_cast<float4*>(p)[i], M4 = reinterpret_cast<float4*>(m)[i], V4 = reinterpret_cast<float4*>(v)[i];
      const float4 G4 = __ldg(reinterpret_cast<const float4*>(g) + i);
      adam_update(P4.x, M4.x, V4.x, G4.x, a);
      adam_update(P4.y, M4.y, V4.y, G4.y, a);
      adam_update(P4.z, M4.z, V4.z, G4.z, a);
      adam_update(P4.w, M4.w, V4.w, G4.w, a);
      reinterpret_cast<float4*>(p)[i] = P4;
      reinterpret_cast<float4*>(m)[i] = M4;
      reinterpret_cast<float4*>(v)[i] = V4;
    }
    done = n4 * 4;
  }
  for (size_t i = done + tid; i < n; i += nth) adam_update(p[i], m[i], v[i], g[i], a);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static AdamStep<float> make_adam(const svbrdf_adam_t& h, int64_t step) {
  // torch/optim/adam.py:457-547: scalars are formed in double precision on the host
  const double bc1 = 1.0 - std::pow(h.beta1, double(step));
  const double bc2 = 1.0 - std::pow(h.beta2, double(step));
  AdamStep<float> a;
  a.one_minus_b1 = float(1.0 - h.beta1);
  a.b2 = float(h.beta2);
  a.one_minus_b2 = float(1.0 - h.beta2);
  a.step_size = float(h.lr / bc1);
  a.inv_sqrt_bc2 = float(1.0 / std::sqrt(bc2));
  a.eps = float(h.eps);
  return a;
}

static int check_geom(const svbrdf_geom_t* g) {
  if (!g || !g->camera_pos || !g->light_pos || !g->light_pow) return SVBRDF_E_BADARG;
  if (g->res <= 0 || g->rows <= 0 || g->n_lights <= 0 || g->row_offset < 0) return SVBRDF_E_BADARG;
  if (g->plane_stride != 0 && g->plane_stride < (long long)g->rows * g->res) return SVBRDF_E_BADARG;
  if (size_t(g->n_lights) * 2 * sizeof(float4) > 200 * 1024) return SVBRDF_E_UNSUPPORTED;
  return 0;
}

static Params base_params(const svbrdf_geom_t* g) {
  Params P{};
  P.cam = g->camera_pos;
  P.light = g->light_pos;
  P.pow = g->light_pow;
  P.texels = (long long)g->rows * g->res;
  P.stride = g->plane_stride ? g->plane_stride : P.texels;
  P.size = g->size;
  P.res = g->res;
  P.row_offset = g->row_offset;
  P.n_lights = g->n_lights;
  return P;
}

static inline int n_blocks(const Params& P) { return int((P.texels + kThreads - 1) / kThreads); }

template <int MODE, bool WANT_POW, int TGT>
static int launch(const Params& P, cudaStream_t st) {
  const size_t smem = size_t(P.n_lights) * 2 * sizeof(float4);
  auto kern = texel_kernel<MODE, WANT_POW, TGT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
  }
  kern<<<n_blocks(P), kThreads, smem, st>>>(P);
  return int(cudaGetLastError());
}

template <int MODE>
static int launch_l2(const Params& P, bool want_pow, int tgt, cudaStream_t st) {
  if (tgt == SVBRDF_TARGET_F32)
    return want_pow ? launch<MODE, true, SVBRDF_TARGET_F32>(P, st) : launch<MODE, false, SVBRDF_TARGET_F32>(P, st);
  if (tgt == SVBRDF_TARGET_U8)
    return want_pow ? launch<MODE, true, SVBRDF_TARGET_U8>(P, st) : launch<MODE, false, SVBRDF_TARGET_U8>(P, st);
  return SVBRDF_E_UNSUPPORTED;
}

static double l2_scale(const svbrdf_geom_t* g, int n_total) {
  return 2.0 / (double(n_total) * 3.0 * double(g->res) * double(g->res) * kGamma);
}

}  // namespace svbrdf

using namespace svbrdf;

extern "C" {

int svbrdf_abi_version(void) { return SVBRDF_B200_ABI_VERSION; }

const char* svbrdf_error_string(int code) {
  if (code == 0) return "success";
  if (code == SVBRDF_E_BADARG) return "svbrdf_b200: bad argument (null pointer or non-positive size)";
  if (code == SVBRDF_E_UNSUPPORTED) return "svbrdf_b200: unsupported target dtype or light count";
  if (code > 0) return cudaGetErrorString(cudaError_t(code));
  return "svbrdf_b200: unknown error";
}

size_t svbrdf_workspace_bytes(int32_t res, int32_t rows) {
  if (res <= 0 || rows <= 0) return 0;
  const long long texels = (long long)res * rows;
  return size_t((texels + kThreads - 1) / kThreads) * 4 * sizeof(float);
}

int svbrdf_render_fwd(const svbrdf_geom_t* geom, const float* tex, float* out, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !out) return SVBRDF_E_BADARG;
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.out = out;
  return launch<kModeRender, false, SVBRDF_TARGET_F32>(P, stream);
}

int svbrdf_render_bwd(const svbrdf_geom_t* geom, const float* tex, const float* grad_out, float* grad_tex, float* grad_pow,
                      void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !grad_out || !grad_tex || (grad_pow && !workspace)) return SVBRDF_E_BADARG;
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = grad_out;
  P.out = grad_tex;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(1.0 / kGamma);
  int e = grad_pow ? launch<kModeVjp, true, SVBRDF_TARGET_F32>(P, stream) : launch<kModeVjp, false, SVBRDF_TARGET_F32>(P, stream);
  if (e || !grad_pow) return e;
  FinalizeParams F{};
  F.partials = P.partials;
  F.n_blocks = n_blocks(P);
  F.grad_scale = P.scale;
  F.grad_pow = grad_pow;
  finalize_kernel<<<1, 256, 0, stream>>>(F);
  return int(cudaGetLastError());
}

int svbrdf_l2_grad(const svbrdf_geom_t* geom, const float* tex, const void* target, int32_t target_dtype, int32_t n_total,
                   float* grad_tex, float* loss_out, float* grad_pow, void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !target || !grad_tex || !workspace || n_total < geom->n_lights) return SVBRDF_E_BADARG;
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = target;
  P.out = grad_tex;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(l2_scale(geom, n_total));
  if (int e = launch_l2<kModeL2Grad>(P, grad_pow != nullptr, target_dtype, stream)) return e;
  FinalizeParams F{};
  F.partials = P.partials;
  F.n_blocks = n_blocks(P);
  F.loss_norm = 1.0 / (double(n_total) * 3.0 * double(geom->res) * double(geom->res));
  F.grad_scale = P.scale;
  F.loss_out = loss_out;
  F.grad_pow = grad_pow;
  finalize_kernel<<<1, 256, 0, stream>>>(F);
  return int(cudaGetLastError());
}

int svbrdf_l2_adam_run(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target, int32_t target_dtype,
                       const svbrdf_adam_t* adam, int32_t epochs, float* loss_curve, float* pow_state, void* workspace,
                       svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !m || !v || !target || !adam || !workspace || epochs < 0 || adam->step < 1) return SVBRDF_E_BADARG;
  Params P = base_params(geom);
  P.tex = tex;
  P.m = m;
  P.v = v;
  P.io = target;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(l2_scale(geom, geom->n_lights));
  FinalizeParams F{};
  F.partials = P.partials;
  F.n_blocks = n_blocks(P);
  F.loss_norm = 1.0 / (double(geom->n_lights) * 3.0 * double(geom->res) * double(geom->res));
  F.grad_scale = P.scale;
  if (pow_state) {
    F.pow = const_cast<float*>(geom->light_pow);
    F.pow_state = pow_state;
  }
  for (int e = 0; e < epochs; ++e) {
    P.adam = make_adam(*adam, adam->step + e);
    if (int err = launch_l2<kModeL2Adam>(P, pow_state != nullptr, target_dtype, stream)) return err;
    F.adam = P.adam;
    F.loss_out = loss_curve ? loss_curve + e : nullptr;
    if (F.loss_out || F.pow) {
      finalize_kernel<<<1, 256, 0, stream>>>(F);
      if (cudaError_t err = cudaGetLastError()) return int(err);
    }
  }
  return 0;
}

int svbrdf_l2_adam_step(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target, int32_t target_dtype,
                        const svbrdf_adam_t* adam, float* loss_out, float* pow_state, void* workspace,
                        svbrdf_stream_t stream) {
  return svbrdf_l2_adam_run(geom, tex, m, v, target, target_dtype, adam, 1, loss_out, pow_state, workspace, stream);
}

int svbrdf_adam_apply(float* param, float* m, float* v, const float* grad, size_t count, const svbrdf_adam_t* adam,
                      svbrdf_stream_t stream) {
  if (!param || !m || !v || !grad || !adam || adam->step < 1) return SVBRDF_E_BADARG;
  if (count == 0) return 0;
  const AdamStep<float> a = make_adam(*adam, adam->step);
  const size_t want = (count / 4 + 255) / 256 + 1;
  const int blocks = int(want < size_t(148 * 16) ? want : size_t(148 * 16));
  adam_apply_kernel<<<blocks, 256, 0, stream>>>(param, m, v, grad, count, a);
  return int(cudaGetLastError());
}

}  // extern "C"
