// TEST HARNESS ONLY — host build of csrc/svbrdf_core.cuh (the per-texel math the CUDA
// kernels instantiate), in float (libm in place of MUFU) and in double.
//
// Purpose: check the analytic backward / Adam restatement against the oracle on a
// machine without a GPU (`pytest -m "not gpu"`).  Nothing in the product imports,
// links or loads this file; the product has no CPU path.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../svbrdf_diff_renderer_b200/csrc/svbrdf_core.cuh"

using namespace svbrdf;

namespace {

template <typename T>
LightGeom<T> geom(const T* cam, const T* light, int i) {
  LightGeom<T> g;
  g.cx = cam[3 * i]; g.cy = cam[3 * i + 1]; g.cz = cam[3 * i + 2]; g.cz2 = g.cz * g.cz;
  g.lx = light[3 * i]; g.ly = light[3 * i + 1]; g.lz = light[3 * i + 2]; g.lz2 = g.lz * g.lz;
  return g;
}

// mode: 0 render, 1 vjp, 2 l2 grad, 3 l2 + adam, 4 vjp + l2 term (LightMode kVjpL2: `io` upstream gradient, `m` doubles
// as the target image and adam[0] as the L2 weight — the arguments of svbrdf_render_norm_l2_bwd's inner loop)
template <typename T, bool COLOC>
void run(int mode, T* tex, const T* cam, const T* light, const T* pw, float size, int res, int row0, int rows, int W,
         int N, int n_total, const T* io, T* out, T* grad_tex, T* grad_pow, double* loss, int outer_clamp, T* m, T* v,
         const double* adam) {
  const size_t plane = size_t(rows) * W;
  double loss_acc = 0.0, gp_acc[3] = {0, 0, 0};
  T scale;
  if (mode == 1 || mode == 4) scale = T(1.0 / kGamma);
  else scale = T(2.0 / (double(n_total) * 3.0 * double(res) * double(res) * kGamma));
  for (int r = 0; r < rows; ++r) {
    for (int c = 0; c < W; ++c) {
      const size_t p = size_t(r) * W + c;
      T t[9];
      bool outer[9];
      for (int k = 0; k < 9; ++k) {
        const T raw = tex[k * plane + p];
        if (outer_clamp) {
          outer[k] = raw >= T(-1) && raw <= T(1);
          t[k] = Fm<T>::min(Fm<T>::max(raw, T(-1)), T(1));
        } else {
          outer[k] = true;
          t[k] = raw;
        }
      }
      Texel<T> tx;
      TexelAux<T> ax;
      texel_position(row0 + r, c, res, size, tx.px, tx.py);
      texel_prologue(t, pw, tx, ax);
      Grads<T> g;
      grads_zero(g);
      for (int i = 0; i < N; ++i) {
        const LightGeom<T> lg = geom(cam, light, i);
        T in3[3] = {0, 0, 0}, o3[3];
        if (mode >= 1)
          for (int ch = 0; ch < 3; ++ch) in3[ch] = io[(size_t(i) * 3 + ch) * plane + p];
        if (mode == 0) {
          shade_light<T, kRender, COLOC, false>(tx, lg, in3, o3, g);
          for (int ch = 0; ch < 3; ++ch) out[(size_t(i) * 3 + ch) * plane + p] = o3[ch];
        } else if (mode == 1) {
          shade_light<T, kVjp, COLOC, true>(tx, lg, in3, o3, g);
        } else if (mode == 4) {
          T tg[3];
          for (int ch = 0; ch < 3; ++ch) tg[ch] = m[(size_t(i) * 3 + ch) * plane + p];
          shade_light<T, kVjpL2, COLOC, true>(tx, lg, in3, o3, g, tg, T(adam[0]));
        } else {
          shade_light<T, kL2, COLOC, true>(tx, lg, in3, o3, g);
        }
      }
      if (mode == 0) continue;
      T gt[9];
      texel_epilogue<T, COLOC>(tx, ax, pw, g, scale, outer, gt);
      loss_acc += double(g.loss) + double(g.loss_g);
      for (int ch = 0; ch < 3; ++ch) gp_acc[ch] += double(g.pw[ch]) * double(scale);
      if (mode == 3) {
        AdamStep<T> a;
        a.one_minus_b1 = T(adam[0]); a.b2 = T(adam[1]); a.one_minus_b2 = T(adam[2]);
        a.step_size = T(adam[3]); a.inv_sqrt_bc2 = T(adam[4]); a.eps = T(adam[5]);
        for (int k = 0; k < 9; ++k) adam_update(tex[k * plane + p], m[k * plane + p], v[k * plane + p], gt[k], a);
      } else {
        for (int k = 0; k < 9; ++k) grad_tex[k * plane + p] = gt[k];
      }
    }
  }
  if (loss) *loss = loss_acc / (double(n_total) * 3.0 * double(res) * double(res));
  if (grad_pow) for (int ch = 0; ch < 3; ++ch) grad_pow[ch] = pow_grad(T(gp_acc[ch]), pw[ch]);   // acc = pw_c * dL/dpw_c
}

// Packed instantiation (T = V2): two horizontally adjacent texels per "thread", as the packed CUDA kernel does.
template <bool COLOC>
void run_v2(int mode, float* tex, const float* cam, const float* light, const float* pwf, float size, int res, int row0, int rows,
            int W, int N, int n_total, const float* io, float* out, float* grad_tex, float* grad_pow, double* loss, int outer_clamp,
            float* m, float* v, const double* adam) {
  typedef V2 T;
  const size_t plane = size_t(rows) * W;
  double loss_acc = 0.0, gp_acc[3] = {0, 0, 0};
  const T scale = (mode == 1) ? T(1.0 / kGamma) : T(2.0 / (double(n_total) * 3.0 * double(res) * double(res) * kGamma));
  const T pw[3] = {T(pwf[0]), T(pwf[1]), T(pwf[2])};
  for (size_t p = 0; p + 1 < plane + 1 && p < plane; p += 2) {          // plane is even (checked by the caller)
    T t[9];
    Fm<T>::mask outer[9];
    for (int k = 0; k < 9; ++k) {
      const T raw(tex[k * plane + p], tex[k * plane + p + 1]);
      if (outer_clamp) {
        outer[k] = Fm<T>::mand(Fm<T>::ge(raw, T(-1)), Fm<T>::le(raw, T(1)));
        t[k] = Fm<T>::min(Fm<T>::max(raw, T(-1)), T(1));
      } else {
        outer[k] = Fm<T>::mtrue();
        t[k] = raw;
      }
    }
    Texel<T> tx;
    TexelAux<T> ax;
    float px0, py0, px1, py1;
    texel_position(row0 + int(p / W), int(p % W), res, size, px0, py0);
    texel_position(row0 + int((p + 1) / W), int((p + 1) % W), res, size, px1, py1);
    tx.px = T(px0, px1);
    tx.py = T(py0, py1);
    texel_prologue(t, pw, tx, ax);
    Grads<T> g;
    grads_zero(g);
    for (int i = 0; i < N; ++i) {
      const LightGeom<float> lf = geom(cam, light, i);
      LightGeom<T> lg;
      lg.cx = T(lf.cx); lg.cy = T(lf.cy); lg.cz = T(lf.cz); lg.cz2 = T(lf.cz2);
      lg.lx = T(lf.lx); lg.ly = T(lf.ly); lg.lz = T(lf.lz); lg.lz2 = T(lf.lz2);
      T in3[3] = {T(0.f), T(0.f), T(0.f)}, o3[3];
      if (mode >= 1)
        for (int ch = 0; ch < 3; ++ch) in3[ch] = T(io[(size_t(i) * 3 + ch) * plane + p], io[(size_t(i) * 3 + ch) * plane + p + 1]);
      if (mode == 0) {
        shade_light<T, kRender, COLOC, false>(tx, lg, in3, o3, g);
        for (int ch = 0; ch < 3; ++ch) {
          out[(size_t(i) * 3 + ch) * plane + p] = o3[ch].x;
          out[(size_t(i) * 3 + ch) * plane + p + 1] = o3[ch].y;
        }
      } else if (mode == 1) {
        shade_light<T, kVjp, COLOC, true>(tx, lg, in3, o3, g);
      } else {
        shade_light<T, kL2, COLOC, true>(tx, lg, in3, o3, g);
      }
    }
    if (mode == 0) continue;
    T gt[9];
    texel_epilogue<T, COLOC>(tx, ax, pw, g, scale, outer, gt);
    loss_acc += double(g.loss.x) + double(g.loss.y) + double(g.loss_g.x) + double(g.loss_g.y);
    for (int ch = 0; ch < 3; ++ch) gp_acc[ch] += (double(g.pw[ch].x) + double(g.pw[ch].y)) * double(scale.x);
    if (mode == 3) {
      AdamStep<T> a;
      a.one_minus_b1 = T(adam[0]); a.b2 = T(adam[1]); a.one_minus_b2 = T(adam[2]);
      a.step_size = T(adam[3]); a.inv_sqrt_bc2 = T(adam[4]); a.eps = T(adam[5]);
      for (int k = 0; k < 9; ++k) {
        T pp(tex[k * plane + p], tex[k * plane + p + 1]), mm(m[k * plane + p], m[k * plane + p + 1]), vv(v[k * plane + p], v[k * plane + p + 1]);
        adam_update(pp, mm, vv, gt[k], a);
        tex[k * plane + p] = pp.x; tex[k * plane + p + 1] = pp.y;
        m[k * plane + p] = mm.x; m[k * plane + p + 1] = mm.y;
        v[k * plane + p] = vv.x; v[k * plane + p + 1] = vv.y;
      }
    } else {
      for (int k = 0; k < 9; ++k) { grad_tex[k * plane + p] = gt[k].x; grad_tex[k * plane + p + 1] = gt[k].y; }
    }
  }
  if (loss) *loss = loss_acc / (double(n_total) * 3.0 * double(res) * double(res));
  if (grad_pow) for (int ch = 0; ch < 3; ++ch) grad_pow[ch] = pow_grad(float(gp_acc[ch]), pwf[ch]);
}

template <typename T>
void dispatch(int coloc, int mode, T* tex, const T* cam, const T* light, const T* pw, float size, int res, int row0,
              int rows, int W, int N, int n_total, const T* io, T* out, T* grad_tex, T* grad_pow, double* loss,
              int outer_clamp, T* m, T* v, const double* adam) {
  if (coloc) run<T, true>(mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
  else run<T, false>(mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
}

}  // namespace

extern "C" {

// x[i] / b through the kernels' fixed-divisor sequence (svbrdf_core.cuh div_rn) and through the IEEE division
void emu_fixed_div(int n, const float* x, float b, float* by_sequence, float* by_division) {
  const FixedDiv d = make_fixed_div(b);
  for (int i = 0; i < n; ++i) {
    by_sequence[i] = div_rn(x[i], d);
    volatile float q = x[i] / b;
    by_division[i] = q;
  }
}

// texel centres of one row/column: the reference's division (texel_position) and the kernels' corrected-reciprocal form
void emu_positions(int res, float size, float* by_division, float* by_reciprocal) {
  const float inv = 1.0f / float(res);
  for (int j = 0; j < res; ++j) {
    float px, py;
    texel_position(j, j, res, size, px, py);
    by_division[2 * j] = px; by_division[2 * j + 1] = py;
    texel_position_rcp(j, j, float(res), inv, size, px, py);
    by_reciprocal[2 * j] = px; by_reciprocal[2 * j + 1] = py;
  }
}

void emu_run_f32(int coloc, int mode, float* tex, const float* cam, const float* light, const float* pw, float size,
                 int res, int row0, int rows, int W, int N, int n_total, const float* io, float* out, float* grad_tex,
                 float* grad_pow, double* loss, int outer_clamp, float* m, float* v, const double* adam) {
  dispatch<float>(coloc, mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
}

void emu_run_v2(int coloc, int mode, float* tex, const float* cam, const float* light, const float* pw, float size,
                int res, int row0, int rows, int W, int N, int n_total, const float* io, float* out, float* grad_tex,
                float* grad_pow, double* loss, int outer_clamp, float* m, float* v, const double* adam) {
  if (coloc) run_v2<true>(mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
  else run_v2<false>(mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
}

void emu_run_f64(int coloc, int mode, double* tex, const double* cam, const double* light, const double* pw, float size,
                 int res, int row0, int rows, int W, int N, int n_total, const double* io, double* out, double* grad_tex,
                 double* grad_pow, double* loss, int outer_clamp, double* m, double* v, const double* adam) {
  dispatch<double>(coloc, mode, tex, cam, light, pw, size, res, row0, rows, W, N, n_total, io, out, grad_tex, grad_pow, loss, outer_clamp, m, v, adam);
}

}  // extern "C"
