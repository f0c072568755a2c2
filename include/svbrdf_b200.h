/* svbrdf_b200 — C ABI of the B200-native per-pixel SVBRDF optimisation path.
 *
 * The reference (tflsguoyu/svbrdf-diff-renderer) is pure Python and exposes no FFI; its
 * boundary for this path is the Python API of src/microfacet.py, src/svbrdf.py and
 * src/scripts.py::optim_perpixel.  This header is the native boundary a replacement binds
 * instead of the torch-eager op sequence; every entry point names the reference interface it
 * stands in for.  The Python classes in svbrdf_diff_renderer_b200/ bind it with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous memory owned by the caller, except
 *    `const svbrdf_geom_t*`, which is a host struct read during the call;
 *  - images are planar: textures [9, rows, res] (channel order of svbrdf.py:38: diffuse 0:3,
 *    normal-xy 3:5, roughness 5, specular 6:9), image stacks [N, 3, rows, res];
 *    `plane_stride` (elements) is the distance between consecutive planes;
 *  - calls enqueue work on `stream` and return; they never allocate, free or synchronise;
 *  - return value: 0 on success, a positive cudaError_t, or a negative SVBRDF_E_* code;
 *    svbrdf_error_string() maps either to text;  no exceptions cross the boundary;
 *  - re-entrant: no global mutable state; concurrent calls on different streams/devices are
 *    safe as long as they use different workspaces.
 */
#ifndef SVBRDF_B200_H_
#define SVBRDF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVBRDF_B200_ABI_VERSION 1

typedef struct CUstream_st* svbrdf_stream_t; /* == cudaStream_t */

enum {
  SVBRDF_E_BADARG = -1,    /* null pointer, non-positive size, misaligned pointer */
  SVBRDF_E_UNSUPPORTED = -2 /* unknown target dtype / too many lights for shared memory */
};

/* dtype of the target image stack handed to the L2 entry points */
enum {
  SVBRDF_TARGET_F32 = 0, /* what SvbrdfIO.load_images_th returns (svbrdf.py:191-204) */
  SVBRDF_TARGET_U8 = 1,  /* the PNG bytes themselves; decoded as float(b)/255 in-kernel (imageio.py:18-19) */
  SVBRDF_TARGET_F16 = 2  /* IEEE binary16 targets (half the bytes of f32; the conversion to float is exact); accepted by
                            svbrdf_l2_grad, svbrdf_l2_grad_push, svbrdf_l2_adam_step/_run */
};

/* Capture geometry: the arguments of Microfacet.__init__ (microfacet.py:10-26). */
typedef struct svbrdf_geom_t {
  const float* camera_pos; /* [n_lights,3]  cl[0] */
  const float* light_pos;  /* [n_lights,3]  cl[1] */
  const float* light_pow;  /* [3]           cl[2]; read at kernel time, so update_light() is a pointer swap */
  float size;              /* im_size in cm (microfacet.py:17) */
  int32_t res;             /* full image resolution (square, microfacet.py:85-86) */
  int32_t rows;            /* rows held by the buffers of this call (== res, or a row band) */
  int32_t row_offset;      /* global row index of buffer row 0 */
  int32_t n_lights;        /* N of this call (a light shard in view-sharded runs) */
  int64_t plane_stride;    /* elements between planes; 0 means rows*res */
} svbrdf_geom_t;

/* Adam hyper-parameters of torch.optim.Adam as SvbrdfOptim.optim builds it (svbrdf.py:50,52):
 * amsgrad off, weight decay 0.  `step` is 1-based (the value torch's state['step'] has after
 * its increment). */
typedef struct svbrdf_adam_t {
  double lr, beta1, beta2, eps;
  int64_t step;
} svbrdf_adam_t;

int svbrdf_abi_version(void);
const char* svbrdf_error_string(int code);

/* Bytes of scratch (per concurrent call) the entry points below need: block partials plus a finish
 * counter.  The caller zero-fills it ONCE after allocation (cudaMemset); every call leaves the counter
 * at zero again, so the same workspace can be reused by consecutive calls on one stream. */
size_t svbrdf_workspace_bytes(int32_t res, int32_t rows);

/* Microfacet.eval forward (microfacet.py:84-120):  tex [9,rows,res] -> out [N,3,rows,res]. */
int svbrdf_render_fwd(const svbrdf_geom_t* geom, const float* tex, float* out, svbrdf_stream_t stream);

/* Vector-Jacobian product of Microfacet.eval (what loss.backward() runs through the autograd
 * graph of microfacet.py:84-120): grad_out [N,3,rows,res] -> grad_tex [9,rows,res], and, when
 * grad_pow != NULL, d/d light_pow [3] (the update_light path, microfacet.py:81-82).
 * Recomputes the forward; nothing image-sized is saved between fwd and bwd. */
int svbrdf_render_bwd(const svbrdf_geom_t* geom, const float* tex, const float* grad_out, float* grad_tex,
                      float* grad_pow, void* workspace, svbrdf_stream_t stream);

/* clamp(-1,1) -> eval -> MSELoss -> backward (svbrdf.py:60-70) without the optimiser step.
 * n_total is the light count of the WHOLE problem (the MSE mean divides by n_total*3*res*res);
 * a view shard passes its own lights in geom and the global count here, so shard results add.
 * loss_out[0] receives this call's share of the mean-squared error. */
int svbrdf_l2_grad(const svbrdf_geom_t* geom, const float* tex, const void* target, int32_t target_dtype,
                   int32_t n_total, float* grad_tex, float* loss_out, float* grad_pow, void* workspace,
                   svbrdf_stream_t stream);

/* One iteration of SvbrdfOptim.optim's loop body (svbrdf.py:60-71) in a single pass:
 * clamp -> eval -> MSE -> backward -> Adam.step, updating tex/m/v [9,rows,res] in place.
 * The rendered image is never materialised.  loss_out[0] receives the loss BEFORE the update
 * (what svbrdf.py:64 logs).  When pow_state != NULL (optim_light=True, svbrdf.py:48-50) it
 * points at float[6] = Adam m[3], v[3] of light_pow, and geom->light_pow is updated in place. */
int svbrdf_l2_adam_step(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target,
                        int32_t target_dtype, const svbrdf_adam_t* adam, float* loss_out, float* pow_state,
                        void* workspace, svbrdf_stream_t stream);

/* `epochs` iterations of the above, enqueued back to back; loss_curve[e] is epoch e's loss.
 * adam->step is the step number of the first iteration. */
int svbrdf_l2_adam_run(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target,
                       int32_t target_dtype, const svbrdf_adam_t* adam, int32_t epochs, float* loss_curve,
                       float* pow_state, void* workspace, svbrdf_stream_t stream);

/* ---- view-sharded runs over NVLink peer memory (one process per GPU, <= 8 GPUs of one NVSwitch box) ----
 * Texel p is OWNED by rank p / chunk (chunk: a multiple of 480 texels).  Every rank maps every peer's buffers
 * (CUDA IPC / torch symmetric memory); the pointers below are peer-mapped DEVICE pointers, index = rank.
 *   recv[o]: rank o's receive buffer  [world, 9, chunk]  — slot [r] takes rank r's partial gradient of o's texels
 *   tex[q] : rank q's replica of the textures [9, texels]                                                     */
typedef struct svbrdf_peers_t {
  int32_t world, rank;
  int64_t chunk;
  float* recv[8];
  float* tex[8];
  float* tex_multicast; /* optional NVSwitch multicast mapping of the tex replicas (NVLS); NULL = unicast peer stores */
  int32_t pull_tex;     /* 1: no all-gather at all — the gradient kernel TMA-loads each tile's textures from the OWNER's
                           replica over NVLink (hidden by the shared-memory ring) and svbrdf_reduce_adam_push stores the
                           new parameters only into the owner's replica; a rank's replica is then authoritative only for
                           the texels it owns */
} svbrdf_peers_t;

/* svbrdf_l2_grad fused with the reduce-scatter: the partial gradient of every tile is STORED STRAIGHT INTO THE
 * OWNER'S receive slot over NVLink while the next tile is being shaded (no local gradient tensor, no NCCL call).
 * `tex` is this rank's replica (peers->tex[peers->rank]); geom holds this rank's light shard.  geom may describe a row
 * band (rows < res, row_offset): the band's rows*res texels are then what the peer group shares — ownership, receive
 * slots and the replicas are all band-relative (2-D decomposition: row bands x light shards, one peer group per band). */
int svbrdf_l2_grad_push(const svbrdf_geom_t* geom, const float* tex, const void* target, int32_t target_dtype,
                        int32_t n_total, const svbrdf_peers_t* peers, float* loss_out, void* workspace,
                        svbrdf_stream_t stream);

/* After all ranks' pushes have landed (cross-rank barrier): for the texels this rank owns, sum the `world`
 * partial gradients in rank order (deterministic, every texel reduced exactly once => replicas stay bit-identical),
 * apply torch.optim.Adam.step (m, v: [9, chunk], owner-local) and store the new parameters into EVERY rank's
 * replica (all-gather by peer stores; with peers->tex_multicast one multimem.st per value, replicated by the switch).
 * texels = res*res. */
int svbrdf_reduce_adam_push(const svbrdf_peers_t* peers, int64_t texels, float* m, float* v, const svbrdf_adam_t* adam,
                            svbrdf_stream_t stream);

/* torch.optim.Adam.step on a flat fp32 array (adam.py:531-547), used after the gradient
 * all-reduce of a view-sharded run. */
int svbrdf_adam_apply(float* param, float* m, float* v, const float* grad, size_t count, const svbrdf_adam_t* adam,
                      svbrdf_stream_t stream);

/* ---- mode-B consumers: the render as the MaterialGAN latent optimiser and the VGG descriptor loss use it --------------
 * (materialgan.py:136-147, descriptor.py:65-79, optimization.py:28-29).
 * Forward: Microfacet.eval, then VGGLoss.normalize — (x - mean[c]) / std[c] per channel, torchvision Normalize — written
 * to `out` [N,3,rows,res], and, when target != NULL, MSELoss(rendered, target) of the UN-normalised image in loss_out[0]:
 * one pass instead of render + clone + per-image normalise loop + MSE.  mean/std are HOST float[3]. */
int svbrdf_render_norm_l2_fwd(const svbrdf_geom_t* geom, const float* tex, const float* mean, const float* std_,
                              const void* target, int32_t target_dtype, float* out, float* loss_out, void* workspace,
                              svbrdf_stream_t stream);

/* Backward of the above in one pass: grad_out [N,3,rows,res] is dLoss/d(normalised image) (from the feature network's
 * backward), l2_grad a DEVICE scalar dLoss/d(loss_out[0]) (NULL or target == NULL: no L2 term).  The upstream of the
 * render VJP is grad_out/std[c] + l2_grad * 2 (rendered - target) / (N*3*res*res), formed per sample in registers. */
int svbrdf_render_norm_l2_bwd(const svbrdf_geom_t* geom, const float* tex, const float* std_, const float* grad_out,
                              const void* target, int32_t target_dtype, const float* l2_grad, float* grad_tex,
                              float* grad_pow, void* workspace, svbrdf_stream_t stream);

/* ---- texture-map hand-off between resolutions -------------------------------
 * The reference carries maps from one optim_perpixel call to the next (256 -> 512 -> 1024, run.py:55-56) through
 * 8-bit PNG files: SvbrdfIO.save_textures_th (svbrdf.py:168-189) quantises, SvbrdfIO.load_textures_th
 * (svbrdf.py:150-166) decodes and resizes THE BYTES with cv2.resize(INTER_LANCZOS4) (imageio.py:14-15,75-76).
 * The three entry points below are that round trip on the device, bit for bit (integer/byte work; float steps use
 * IEEE sqrt/div in the reference's operation order):  encode -> resize -> decode  ==  save -> load(res).
 * Byte planes are planar, RGB order: [dif r,g,b | nom x,y,z | rgh | spe r,g,b]  (10 planes of rows*cols bytes). */
#define SVBRDF_MAP_PLANES 10

/* save_textures_th + imwrite: tex [9,rows,cols] (plane_stride elements apart; 0 = rows*cols) -> bytes [10,rows,cols].
 * clamp_input != 0 applies the caller's textures.clamp(-1,1) (scripts.py:91) first. */
int svbrdf_maps_encode_u8(const float* tex, int64_t plane_stride, int32_t rows, int32_t cols, int32_t clamp_input,
                          uint8_t* bytes, svbrdf_stream_t stream);

/* imread + load_textures_th: bytes [10,rows,cols] -> tex [9,rows,cols] in the parameter range [-1,1]. */
int svbrdf_maps_decode_u8(const uint8_t* bytes, int32_t rows, int32_t cols, float* tex, int64_t plane_stride,
                          svbrdf_stream_t stream);

/* HOST function: the per-destination-index tables of cv2.resize(INTER_LANCZOS4) on 8-bit data for one axis
 * (OpenCV imgproc/resize.cpp: interpolateLanczos4, INTER_RESIZE_COEF_BITS = 11): first_tap[d] = floor(fx) - 3 (may be
 * out of range: taps are clamped to the border), coef[d][8] = saturate_cast<short>(w * 2048).  The caller uploads them. */
int svbrdf_lanczos4_tables(int32_t src_size, int32_t dst_size, int32_t* first_tap, int16_t* coef);

/* cv2.resize(src, (dst_cols, dst_rows), interpolation=INTER_LANCZOS4) on `planes` independent uint8 planes:
 * horizontal pass in int32, vertical pass in int32, one rounding shift by 22 bits, replicated borders — bit-identical
 * to OpenCV.  x_tap/x_coef, y_tap/y_coef: DEVICE copies of the tables above for the two axes. */
int svbrdf_resize_lanczos4_u8(const uint8_t* src, int32_t planes, int32_t src_rows, int32_t src_cols, uint8_t* dst,
                              int32_t dst_rows, int32_t dst_cols, const int32_t* x_tap, const int16_t* x_coef,
                              const int32_t* y_tap, const int16_t* y_coef, svbrdf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SVBRDF_B200_H_ */
