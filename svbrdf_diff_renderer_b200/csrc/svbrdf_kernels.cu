// sm_100a kernels + C ABI (include/svbrdf_b200.h) of the per-pixel SVBRDF optimisation path.
//
// One thread owns one texel for the whole pass: the texel's 9 channels are read once, the
// material parameters and gradient accumulators stay in registers while the thread loops over
// all lights, and the pass ends with the texture-gradient epilogue and — in the fused mode —
// the Adam update.  The rendered image is never written in the L2 modes and nothing image-sized
// is saved between forward and backward (the backward kernel recomputes the forward per light).
// The arithmetic is elementwise and transcendental (MUFU) bound: no tensor cores.
//
// Two kernels implement that pass:
//
//  * tile_kernel  (the hot path) — persistent, warp-specialised.  Each CTA walks over 480-texel
//    tiles.  A producer warp streams every byte the tile needs (texture planes, 3 or 4 lights of
//    targets per chunk, Adam m and v) from HBM into a shared-memory ring with 1-D TMA bulk copies
//    (cp.async.bulk ... mbarrier::complete_tx); 15 consumer warps wait on the slot's mbarrier,
//    pull their texel's values with conflict-free LDS, release the slot and compute.  The ring
//    keeps 7-9 slots of 17-23 KB per CTA in flight without costing the consumers a single register,
//    which is what the first version of this kernel (plain LDG with a 2-light register prefetch)
//    could not do: ncu showed 6.8 warps per issue slot stalled on long-scoreboard at 24 %
//    occupancy (profiles/r01_v1_ldg_*.txt).  Results leave through coalesced STG.
//
//  * texel_kernel — one thread per texel with direct LDG/STG.  Used for the forward render
//    (store-dominated: nothing to stage) and as the general fallback whenever the TMA alignment
//    rules (16-byte aligned plane segments) do not hold.
//
// Per-light geometry (2 x float4 per light) is staged in shared memory once per CTA.  Whether
// every light is co-located with its camera is detected while staging; co-located captures
// (everything the reference's capture code emits, capture.py:70-71) take the folded loop body.
#include <cuda.h>              // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/svbrdf_b200.h"
#include "svbrdf_core.cuh"
#include "svbrdf_host.h"

namespace svbrdf {

constexpr int kThreads = 256;          // texel_kernel: threads (= texels) per CTA
constexpr int kPrefetch = 2;           // texel_kernel: lights of io data in flight per thread
// tile_kernel shapes: CW consumer warps (+1 producer warp) per CTA, LANES texels per consumer thread.
// Registers are allocated to a CTA in units of 4 warps, so CW+1 is kept a multiple of 4.
//  * scalar (LANES = 1): 15+1 warps, one CTA per SM, 128 registers/thread, no spills.  Measured on B200 at
//    1024^2 x 9 lights (profiles/r01_variants.txt): beats 2 x (7+1) warps and 19+1 / 23+1 warps at 96 / 80
//    registers: 80.0 vs 89.1 / 84.7 / 89.7 us per step.
//  * packed (LANES = 2): two texels per thread in FP32x2 registers (FFMA2/FMUL2/FADD2), 7+1 warps at up to
//    255 registers.
//  * CHUNK = lights per ring slot: the consumer loads the targets of CHUNK lights, releases the slot and shades them
//    in one unrolled, unguarded block (CHUNK independent dependency chains interleave).  Measured (profiles/
//    r01_s3_variants_chunk_lights.txt): 4 lights per slot are 4.1 % faster than 3 at 64 and at 16 lights, 3 are 2.7 %
//    faster at 9 lights (9 = 4+4+1: the odd light takes the guarded partial-slot path); 5, 6, 8 are slower than 4.
//    The launcher picks 3 or 4 from the light count (pick_chunk).
//  * PW = warps of the producer's slice of the CTA.  PW = 1: the producer is the last warp of the last warpgroup and all
//    warps keep the launch-time register count.  PW = 4: the producer sits in a warpgroup of its own (3 of its warps
//    retire at once), so the register file can be re-split with setmaxnreg — the producer warpgroup drops to 24
//    registers, the CW consumer warps (a multiple of 4: the same number on each of the 4 SM sub-partitions, which 15 + 1
//    is not) rise from the launch-time MAXNREG to CREGS.
template <int CW, int LANES, int MAXNREG, int CHUNK = 3, int PW = 1, int CREGS = 0>
struct TileShape {
  static constexpr int kCW = CW;
  static constexpr int kLanes = LANES;
  static constexpr int kPW = PW;
  static constexpr int kConsumerRegs = CREGS;        // setmaxnreg.inc target of the consumer warpgroups (PW = 4)
  static constexpr int kTile = 32 * CW * LANES;      // texels per tile
  static constexpr int kConsumers = 32 * CW;         // consumer threads
  static constexpr int kThreads = 32 * (CW + PW);    // + the producer warp(group)
  static constexpr int kChunk = CHUNK;               // lights per ring slot
  static constexpr int kSlotPlanes = 3 * CHUNK > 9 ? 3 * CHUNK : 9;
  static constexpr int kSlotBytes = kSlotPlanes * kTile * 4;   // one ring slot: 9 plane segments (12 with 4-light slots)
  static constexpr int kMaxReg = MAXNREG;
};
#ifndef SV_CONSUMER_WARPS
#define SV_CONSUMER_WARPS 15
#endif
#ifndef SV_TILE_MAXNREG
#define SV_TILE_MAXNREG 128
#endif
#ifndef SV_PACKED_WARPS
#define SV_PACKED_WARPS 7
#endif
#ifndef SV_PACKED_MAXNREG
#define SV_PACKED_MAXNREG 255
#endif
#ifndef SV_ENABLE_PACKED
#define SV_ENABLE_PACKED 1
#endif
#ifndef SV_PROBE_AHEAD
#define SV_PROBE_AHEAD 0               // measured slower (profiles/r01_s3_variants_probe_ahead.txt): off
#endif
#ifndef SV_ENABLE_TS
#define SV_ENABLE_TS 1                 // the store-back kernel is written for 480-texel tiles: 0 for builds with other SV_CONSUMER_WARPS
#endif
#ifndef SV_PACKED_DEFAULT
#define SV_PACKED_DEFAULT 0
#endif
#ifndef SV_PRODUCER_WARPS
#define SV_PRODUCER_WARPS 1
#endif
#ifndef SV_CONSUMER_REGS
#define SV_CONSUMER_REGS 0
#endif
typedef TileShape<SV_CONSUMER_WARPS, 1, SV_TILE_MAXNREG, 3, SV_PRODUCER_WARPS, SV_CONSUMER_REGS> ScalarShape;
typedef TileShape<SV_CONSUMER_WARPS, 1, SV_TILE_MAXNREG, 4, SV_PRODUCER_WARPS, SV_CONSUMER_REGS> ScalarShape4;
typedef TileShape<SV_PACKED_WARPS, 2, SV_PACKED_MAXNREG, 3> PackedShape;
// Self-fed ring (PW = 0): NO producer warp.  16 consumer warps x 128 registers fill the register file and put the same
// number of warps (4) on each of the SM's 4 sub-partitions — with 15 + 1 three sub-partitions carry 4/15 of the texels each
// and the fourth 3/15 plus the producer's polling.  The ring is refilled by whichever consumer warp hands a slot back LAST:
// mbarrier.arrive returns the barrier state before the arrival, mbarrier.pending_count == 1 identifies the last of the 16
// arrivals, and that warp (in uniform control flow, one elected lane issuing) streams the chunk NS positions further
// down the CTA's chunk stream into the slot before it waits for its own next chunk.  Nobody polls an `empty` barrier.
#ifndef SV_SELF_WARPS
#define SV_SELF_WARPS 16
#endif
typedef TileShape<SV_SELF_WARPS, 1, SV_TILE_MAXNREG, 3, 0> SelfShape;
typedef TileShape<SV_SELF_WARPS, 1, SV_TILE_MAXNREG, 4, 0> SelfShape4;
// Measured (profiles/r02_selffed_ring.md): correct (all GPU tests pass through it) and perfectly balanced — every scheduler
// issues on 63-68 % of its cycles — but 19 % SLOWER than 15 + 1 (5.35 vs 4.49 ms at 4096^2 x 64, 86.3 vs 70.7 us at 1024^2 x 9):
// the refill lands on the critical path of the slowest warp, which therefore stays the slowest and ends up doing all of
// them, and the hand-back bookkeeping costs every warp ~20 instructions per slot — as many as the producer warp's polling
// wasted.  Compiled only with -DSV_ENABLE_SELF=1 and selected with SVBRDF_B200_SELFFED=1.
#ifndef SV_ENABLE_SELF
#define SV_ENABLE_SELF 0
#endif
#ifndef SV_SELFFED_DEFAULT
#define SV_SELFFED_DEFAULT 0
#endif
// SV_SELF_EAGER: the last warp to hand a slot back refills it right there (1) instead of before its next wait (0)
#ifndef SV_SELF_EAGER
#define SV_SELF_EAGER 1
#endif
// SV_STASH: park the values only the epilogue needs (raw texel, gamma derivatives, normal reconstruction:
// 29 floats per texel) in shared memory while the light loop runs, instead of in registers: the compiler
// then stops rematerialising per-texel constants inside the loop.  Measured (profiles/r01_variants.txt):
// 1403 -> 1285 us at 2048^2 x 64 lights, neutral at 9 lights.
#ifndef SV_STASH
#define SV_STASH 1
#endif
constexpr int kStashFloats = 28;          // raw[9] dpow[7] d[3] oms[3] rough alpha mx my mz rlen
// SV_PAIR_CHANNELS: inside one texel, independent channels (RGB radiance chain, the 7 gamma-encoded texture channels,
// the 9 Adam updates) are processed two at a time with the packed FP32x2 instructions — no extra registers, fewer
// issue slots for the same FMA-pipe work (the scalar kernel is issue-bound).
#ifndef SV_PAIR_CHANNELS
#define SV_PAIR_CHANNELS 1
#endif
// Development switch: SV_STREAM_ONLY=1 builds a kernel that moves exactly the same bytes through the same TMA ring
// and stores but skips the shading math — the streaming ceiling of the pipeline design (profiles/r01_variants.txt).
#ifndef SV_STREAM_ONLY
#define SV_STREAM_ONLY 0
#endif

// kModeVjpL2: the consumer backward (svbrdf_render_norm_l2_bwd) on the TMA ring — per light chunk TWO ring slots arrive, the
// upstream gradient of the normalised image (fp32) and the targets (fp32 / uint8), and the light body forms
// grad/std + l2_grad * 2 (out - target)/n per sample in registers (LightMode kVjpL2).
// kModeNormFwd: the consumer forward (svbrdf_render_norm_l2_fwd) on the ring — the targets stream through the slots, each
// light is rendered, compared with its target (L2 partial sums) and stored normalised; no backward half, no epilogue.
enum KernelMode { kModeRender = 0, kModeVjp = 1, kModeL2Grad = 2, kModeL2Adam = 3, kModeVjpL2 = 4, kModeNormFwd = 5 };

struct Params {
  // tile_kernel_ts: 2-D TMA descriptors {texel, plane} of the planar arrays (box = 160 texels x 9 planes)
  alignas(64) CUtensorMap tm_tex;   // textures [9 planes]; also the store target of the fused mode
  alignas(64) CUtensorMap tm_io;    // targets / grad_out [3N planes], f32 or u8
  alignas(64) CUtensorMap tm_m;     // Adam moments [9 planes]
  alignas(64) CUtensorMap tm_v;
  alignas(64) CUtensorMap tm_out;   // gradient output [9 planes] (L2-grad / VJP)
  float* tex;            // [9] planes (read-only except kModeL2Adam)
  float* m;              // Adam first moment  (kModeL2Adam)
  float* v;              // Adam second moment (kModeL2Adam)
  const float* cam;      // [N,3]
  const float* light;    // [N,3]
  float* pow;            // [3] (updated in place by the fused optim_light path)
  const void* io;        // target (L2 modes) or grad_out (VJP): [N,3] planes
  float* out;            // rendered image [N,3] planes (render) or grad_tex [9] planes
  float* partials;       // [blocks][4]: loss, gpow[3]; then the finish counter
  long long stride;      // plane stride in elements
  long long texels;      // rows * res
  float size;
  float inv_res;          // 1/res
  unsigned long long res_magic;   // floor(2^64/res)+1: row = umul64hi(p, magic) for p < 2^32
  int res;
  int row_offset;
  int n_lights;
  float scale;           // constant image-gradient factor
  AdamStep<float> adam;
  // mode-B consumers (norm_l2_kernel): per-channel normalisation (out - mean)/std of the rendered image, an optional
  // target for the L2 term and the upstream gradient of the L2 loss value (device scalar)
  const void* io2;        // target image [N,3] planes (f32 or u8), nullable
  float aff_mean[3], aff_std[3];
  const float* l2_up;     // nullable: dLoss/d(l2 loss value)
  // finalisation by the last CTA to finish (tile_kernel)
  double loss_norm;      // 1/(n_total*3*res*res)
  float* loss_out;       // nullable
  float* grad_pow;       // nullable [3]
  float* pow_state;      // nullable: Adam m[3], v[3] of light_pow (optim_light)
  unsigned int* counters; // [kMaxEpochs] finish tickets (zero before and after every launch)
  int slots;             // ring depth
  // tile_kernel: `epochs` Adam iterations in ONE launch (texels never interact, so a CTA streams its own tiles
  // again for the next epoch without any grid-wide synchronisation); per-epoch bias-correction scalars:
  int epochs;
  float step_size[64];
  float inv_sqrt_bc2[64];
  // view-sharded push mode (svbrdf_l2_grad_push): gradients go to the owner's receive slot in peer memory
  int push_world;         // 0 = off
  int push_rank;
  long long push_chunk;   // texels owned per rank (multiple of the tile size)
  float* push_recv[8];    // peer-mapped receive buffers [world, 9, chunk]
  const float* push_tex[8];  // pull mode: every rank's texture replica; a tile is read from its owner's (nullable)
  int push_pull;
  // Each rank starts its sweep over the tiles at its own chunk, so at any moment the ranks talk to DIFFERENT owners
  // (with identical sweeps all ranks would hit one owner's NVLink port at a time: measured 5.0 instead of 2.8 ms
  // for the gradient kernel on 8 GPUs).
  long long tile_rotate;
  long long span;         // tile_kernel: texels per CTA (contiguous range), 0 = round-robin tiles (see TileMap)
};

__device__ __forceinline__ long long rotate_tile(long long tile, long long rot, long long n_tiles) {
  const long long t = tile + rot;
  return t >= n_tiles ? t - n_tiles : t;
}
// Which texels a CTA of tile_kernel owns.  span == 0: tiles dealt round-robin (tile = cta + j * grid; the peer-sharded
// modes need tile-aligned ownership).  span > 0: ONE contiguous range of `span` texels per CTA, cut into tiles with a
// partial last tile — every CTA then carries the same load (ceil(texels/grid) texels) instead of floor or ceil of
// tiles/grid whole tiles: at 1024^2 (2185 tiles on 148 CTAs) 113 CTAs had 15 tiles and 35 had 14; at 512^2 it was 4 vs 3.
struct TileMap {
  long long lo, hi;       // span mode: this CTA's range
  unsigned count;         // tiles of this CTA per epoch
  bool spans;
};
template <int TILE>
__device__ __forceinline__ TileMap tile_map(const Params& P, long long span, long long n_tiles) {
  TileMap m;
  m.spans = span > 0;
  if (m.spans) {
    m.lo = (long long)blockIdx.x * span;
    m.hi = m.lo + span < P.texels ? m.lo + span : P.texels;
    m.count = m.hi > m.lo ? unsigned((m.hi - m.lo + TILE - 1) / TILE) : 0u;
  } else {
    m.lo = 0;
    m.hi = P.texels;
    m.count = unsigned((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  }
  return m;
}
// first texel of the CTA's j-th tile
template <int TILE>
__device__ __forceinline__ long long tile_first(const TileMap& m, const Params& P, unsigned j, long long n_tiles) {
  if (m.spans) return m.lo + (long long)j * TILE;
  return rotate_tile((long long)blockIdx.x + (long long)j * gridDim.x, P.tile_rotate, n_tiles) * TILE;
}
constexpr int kMaxEpochs = 64;          // per launch
constexpr int kRowsPerEpoch = 4096;     // workspace rows ([loss, gpow x3]) per epoch: one per consumer warp

#ifndef SV_U8_MAGIC
#define SV_U8_MAGIC 0
#endif
#ifndef SV_U8_I2FP
#define SV_U8_I2FP 1                   // measured (profiles/r01_s3_variants_u8_i2fp.txt): 1371 -> 1283 us at 2048^2 x 64, 77.7 -> 74.8 us at 1024^2 x 9
#endif
template <int TGT>
struct IoLoad;
template <>
struct IoLoad<SVBRDF_TARGET_F32> {
  typedef float elem;
  static __device__ __forceinline__ float decode(float x) { return x; }
};
template <>
struct IoLoad<SVBRDF_TARGET_F16> {
  typedef __half elem;
  static __device__ __forceinline__ float decode(__half x) { return __half2float(x); }   // exact
};
template <>
struct IoLoad<SVBRDF_TARGET_U8> {
  typedef unsigned char elem;
  // float(b)/255 correctly rounded (bit-identical to the IEEE division of imageio.py:18-19) in three instructions:
  // 1/255 split into hi + lo floats, q = fma(b, hi, b*lo).  b*hi is exact inside the FMA (8 x 24 bits) and b*lo is a
  // 2^-25-relative correction, so the single rounding of the FMA is the rounding of b/255; checked exhaustively for the
  // 256 inputs with exact rational arithmetic and on the GPU (test_uint8_targets_bit_exact_with_float_decode).
  static __device__ __forceinline__ float decode(unsigned char x) {
#if SV_U8_MAGIC
    // integer -> float without I2F (quarter-rate conversion pipe): 2^23 + b as a bit pattern, minus 2^23 (exact).
    // Measured: 77.4 vs 74.8 us per epoch at 1024^2 x 9 with uint8 targets — the extra issue slot costs more; off.
    const float b = __uint_as_float(0x4B000000u | unsigned(x)) - 8388608.0f;
#elif SV_U8_I2FP
    // 32-bit conversion: I2FP.F32.U32 runs on the ALU pipe, the 16-bit form the compiler picks for a byte (I2F.U16) on the
    // quarter-rate conversion (XU) pipe next to the 10 MUFU operations of the light body
    float b;
    asm("cvt.rn.f32.u32 %0, %1;" : "=f"(b) : "r"(unsigned(x)));
#else
    const float b = float(x);
#endif
    return __fmaf_rn(b, 0x1.010102p-8f, __fmul_rn(b, -0x1.fdfdfep-33f));
  }
};

// ---------------------------------------------------------------------------------------------
// shared pieces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool stage_lights(const Params& P, float4* s_geo, int tid, int nthreads) {
  int same = 1;
  for (int i = tid; i < P.n_lights; i += nthreads) {
    const float cx = P.cam[3 * i], cy = P.cam[3 * i + 1], cz = P.cam[3 * i + 2];
    const float lx = P.light[3 * i], ly = P.light[3 * i + 1], lz = P.light[3 * i + 2];
    s_geo[2 * i] = make_float4(cx, cy, cz, cz * cz);
    s_geo[2 * i + 1] = make_float4(lx, ly, lz, lz * lz);
    same &= (cx == lx) & (cy == ly) & (cz == lz);
  }
  return __syncthreads_and(same) != 0;
}

template <bool COLOC>
__device__ __forceinline__ LightGeom<float> load_geom(const float4* __restrict__ s_geo, int i) {
  LightGeom<float> lg;
  const float4 a = s_geo[2 * i];
  lg.cx = a.x; lg.cy = a.y; lg.cz = a.z; lg.cz2 = a.w;
  if (!COLOC) {
    const float4 b = s_geo[2 * i + 1];
    lg.lx = b.x; lg.ly = b.y; lg.lz = b.z; lg.lz2 = b.w;
  }
  return lg;
}

template <int MODE>
__device__ __forceinline__ void clamp_outer(const float raw[9], float t[9], bool outer[9]) {
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (MODE == kModeL2Grad || MODE == kModeL2Adam) {   // the clamp of svbrdf.py:60
      outer[k] = raw[k] >= -1.f && raw[k] <= 1.f;
      t[k] = fminf(fmaxf(raw[k], -1.f), 1.f);
    } else {
      outer[k] = true;
      t[k] = raw[k];
    }
  }
}

// Sums [n][4] partials in double in a fixed order (run-to-run deterministic), writes the loss and
// the light-power gradient, and — for optim_light — applies Adam to light_pow[3].
// Every thread of the CTA must call it (it synchronises); threads with tid >= 256 only synchronise.
__device__ __forceinline__ void finalize_block(const Params& P, int n, double (*s)[4], int tid, int nthreads) {
  // `nthreads` = CTA size; the first min(nthreads, 256) threads accumulate into s[256][4]
  const int np = nthreads < 256 ? nthreads : 256;
  for (int z = tid; z < 256; z += nthreads) {
#pragma unroll
    for (int c = 0; c < 4; ++c) s[z][c] = 0.0;
  }
  __syncthreads();
  if (tid < np) {
    double acc[4] = {0, 0, 0, 0};
    for (int b = tid; b < n; b += np) {
      const float4 q = __ldcg(reinterpret_cast<const float4*>(P.partials) + b);
      acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) s[tid][c] = acc[c];
  }
  __syncthreads();
  double tot = 0;
  if (tid < 4) {
    for (int i = 0; i < 256; ++i) tot += s[i][tid];
  }
  __syncthreads();
  if (tid < 4) s[0][tid] = tot;
  __syncthreads();
  if (tid == 0 && P.loss_out) P.loss_out[0] = float(s[0][0] * P.loss_norm);
  if (tid < 3 && (P.grad_pow || P.pow_state)) {
    const float pw = P.pow[tid];
    const float gp = pow_grad(float(s[0][1 + tid] * double(P.scale)), pw);
    if (P.grad_pow) P.grad_pow[tid] = gp;
    if (P.pow_state) {
      float p = pw, m = P.pow_state[tid], v = P.pow_state[3 + tid];
      // 3 elements: IEEE sqrt/div, cost is nil (adam.py:531-547)
      m = m + (gp - m) * P.adam.one_minus_b1;
      v = v * P.adam.b2 + P.adam.one_minus_b2 * gp * gp;
      const float denom = sqrtf(v) * P.adam.inv_sqrt_bc2 + P.adam.eps;
      p = p - P.adam.step_size * m / denom;
      P.pow[tid] = p;
      P.pow_state[tid] = m;
      P.pow_state[3 + tid] = v;
    }
  }
}

// tile_kernel, end of an epoch: the consumer warps of a CTA combine their [loss, gpow x3] sums through shared memory
// (consumer-only named barrier — the producer warp is already prefetching the next epoch), the CTA publishes one row
// and takes a ticket; the CTA that draws the last ticket of epoch `e` sums all rows in a fixed order in double
// (deterministic), writes the epoch's loss / light-power gradient and (optim_light) applies Adam to light_pow.
template <int CW>
__device__ __forceinline__ void epoch_end(const Params& P, int e, float acc[4], float (*s_red)[CW][4], int warp, int lane) {
  float r[4] = {acc[0], acc[1], acc[2], acc[3]};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
  }
  float(*buf)[4] = s_red[e & 1];                           // double-buffered: no second barrier needed
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) buf[warp][c] = r[c];
  }
  asm volatile("bar.sync 1, %0;" ::"r"(CW * 32) : "memory");
  if (warp != 0) return;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int w = 0; w < CW; ++w) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] += buf[w][c];
  }
  const int rows = int(gridDim.x);
  float* base = P.partials + size_t(e) * kRowsPerEpoch * 4;
  unsigned ticket = 0;
  if (lane == 0) {
    __stcg(reinterpret_cast<float4*>(base) + blockIdx.x, make_float4(a[0], a[1], a[2], a[3]));
    __threadfence();
    ticket = atomicAdd(P.counters + e, 1u);
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != unsigned(rows - 1)) return;
  __threadfence();
  double t[4] = {0, 0, 0, 0};
  for (int b = lane; b < rows; b += 32) {
    const float4 q = __ldcg(reinterpret_cast<const float4*>(base) + b);
    t[0] += q.x; t[1] += q.y; t[2] += q.z; t[3] += q.w;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[c] += __shfl_xor_sync(0xffffffffu, t[c], o);
  }
  if (lane == 0) {
    if (P.loss_out) P.loss_out[e] = float(t[0] * P.loss_norm);
    P.counters[e] = 0u;                                    // leave the workspace ready for the next launch
  }
  if (lane < 3 && (P.grad_pow || P.pow_state)) {
    const float pw = P.pow[lane];
    const double tl = lane == 0 ? t[1] : (lane == 1 ? t[2] : t[3]);
    const float gp = pow_grad(float(tl * double(P.scale)), pw);
    if (P.grad_pow) P.grad_pow[lane] = gp;
    if (P.pow_state) {                                      // adam.py:531-547 on 3 elements, IEEE sqrt/div
      float p = pw, m = P.pow_state[lane], v = P.pow_state[3 + lane];
      m = m + (gp - m) * P.adam.one_minus_b1;
      v = v * P.adam.b2 + P.adam.one_minus_b2 * gp * gp;
      const float denom = sqrtf(v) * P.inv_sqrt_bc2[e] + P.adam.eps;
      p = p - P.step_size[e] * m / denom;
      P.pow[lane] = p;
      P.pow_state[lane] = m;
      P.pow_state[3 + lane] = v;
    }
  }
}

// Multi-epoch launches: a tile stored in epoch e is re-read by this CTA's TMA loads in epoch e+1.  The stores are
// made visible to the async proxy and the warp's progress is published one tile LATE — right before the next tile's
// stores — so the membar finds the previous tile's stores long complete and does not stall the warp.
__device__ __forceinline__ void publish_tiles(volatile unsigned* s_done, int warp, int lane, unsigned count) {
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  if (lane == 0) s_done[warp] = count;
}

__device__ __forceinline__ void warp_reduce4(float r[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
  }
}

// ---------------------------------------------------------------------------------------------
// texel_kernel: one thread per texel, direct LDG/STG
// ---------------------------------------------------------------------------------------------
template <int MODE, bool COLOC, bool WANT_POW, int TGT>
__device__ __forceinline__ void light_loop(const Params& P, const float4* __restrict__ s_geo, const Texel<float>& tx, long long p,
                                            bool valid, Grads<float>& g) {
  typedef typename IoLoad<TGT>::elem elem;
  const int N = P.n_lights;
  constexpr int LM = (MODE == kModeRender) ? kRender : (MODE == kModeVjp ? kVjp : kL2);
  const long long step = 3 * P.stride;
  const elem* __restrict__ src = static_cast<const elem*>(P.io) + p;    // light 0, channel 0
  float* __restrict__ dst = P.out + p;

  float buf[kPrefetch][3];
  if (MODE != kModeRender) {
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) buf[j][c] = (valid && j < N) ? IoLoad<TGT>::decode(__ldg(src + j * step + c * P.stride)) : 0.f;
    }
    src += kPrefetch * step;
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
        if (i + kPrefetch < N && valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) buf[j][c] = IoLoad<TGT>::decode(__ldg(src + c * P.stride));
        }
        src += step;
      }
      const LightGeom<float> lg = load_geom<COLOC>(s_geo, i);
      shade_light<float, LM, COLOC, WANT_POW>(tx, lg, in3, o3, g);
      if (MODE == kModeRender) {
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) dst[c * P.stride] = o3[c];
        }
        dst += step;
      }
    }
  }
}

template <int MODE, bool WANT_POW, int TGT>
__global__ void __launch_bounds__(kThreads) texel_kernel(const Params P) {
  extern __shared__ float4 s_geo[];
  __shared__ float s_red[kThreads / 32][4];
  const bool coloc = stage_lights(P, s_geo, threadIdx.x, kThreads);

  const long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool valid = p < P.texels;
  const long long pc = valid ? p : 0;

  float pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pw[c] = P.pow[c];

  float raw[9], t[9];
  bool outer[9];
  const float* tex_src = P.tex;
  if (MODE == kModeL2Grad && P.push_pull) tex_src = P.push_tex[pc / P.push_chunk];
#pragma unroll
  for (int k = 0; k < 9; ++k) raw[k] = (MODE == kModeL2Adam) ? P.tex[k * P.stride + pc] : __ldg(tex_src + k * P.stride + pc);
  clamp_outer<MODE>(raw, t, outer);
  Texel<float> tx;
  TexelAux<float> ax;
  {
    const int row = int(pc / P.res);
    const int col = int(pc - (long long)row * P.res);
    texel_position_rcp(row + P.row_offset, col, float(P.res), P.inv_res, P.size, tx.px, tx.py);
  }
  texel_prologue(t, pw, tx, ax);
  Grads<float> g;
  grads_zero(g);

  if (coloc) light_loop<MODE, true, WANT_POW, TGT>(P, s_geo, tx, pc, valid, g);
  else light_loop<MODE, false, WANT_POW, TGT>(P, s_geo, tx, pc, valid, g);
  if (MODE == kModeRender) return;

  float gt[9];
  if (coloc) texel_epilogue<float, true>(tx, ax, pw, g, P.scale, outer, gt);
  else texel_epilogue<float, false>(tx, ax, pw, g, P.scale, outer, gt);
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
    } else if (MODE == kModeL2Grad && P.push_world > 0) {
      const long long owner = p / P.push_chunk;
      float* po = P.push_recv[owner] + (size_t(P.push_rank) * 9) * P.push_chunk + (p - owner * P.push_chunk);
#pragma unroll
      for (int k = 0; k < 9; ++k) po[k * P.push_chunk] = gt[k];
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) P.out[k * P.stride + p] = gt[k];
    }
  }

  // block partials: loss and light-power gradient (fixed order => deterministic)
  constexpr bool kHasLoss = (MODE == kModeL2Grad || MODE == kModeL2Adam);
  if (kHasLoss || WANT_POW) {
    float r[4] = {valid ? grads_loss(g) : 0.f, valid ? g.pw[0] : 0.f, valid ? g.pw[1] : 0.f, valid ? g.pw[2] : 0.f};
    warp_reduce4(r);
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

// ---------------------------------------------------------------------------------------------
// norm_l2_kernel: the render as the mode-B consumers use it (materialgan.py:136-147, descriptor.py:65-79): the image is
// wanted NORMALISED for the feature network ((x - mean)/std per channel, descriptor.py:65-75) and the image loss is an
// L2 against the targets (optimization.py:28-29).  Forward: one pass writes the normalised image and the L2 partial
// sums (instead of render + clone + per-image normalise loop + MSE).  Backward: one pass takes the gradient w.r.t. the
// normalised image, adds the L2 term's gradient from the target and chains to the textures (instead of 4 image-sized
// passes + the render VJP).  One thread per texel, direct LDG/STG.
// ---------------------------------------------------------------------------------------------
// FixedDiv / make_fixed_div / div_rn (x / b with IEEE rounding for a divisor fixed per launch) live in svbrdf_core.cuh.
// the launcher's side of the contract: every divisor and every mean inside the range where neither r nor q can leave the
// normal range for a rendered value in [0, 1]
static bool fixed_div_ok(const float* mean, const float* std_) {
  for (int c = 0; c < 3; ++c) {
    const float a = fabsf(std_[c]);
    if (!(a > 1e-30f && a < 1e30f) || !(fabsf(mean[c]) < 1e30f)) return false;
  }
  return true;
}

template <bool BWD, bool COLOC, bool WANT_POW, int TGT>
__device__ __forceinline__ void norm_l2_lights(const Params& P, const float4* __restrict__ s_geo, const Texel<float>& tx, long long p, bool valid,
                                                Grads<float>& g) {
  typedef typename IoLoad<TGT>::elem elem;
  const int N = P.n_lights;
  const long long step = 3 * P.stride;
  const float* __restrict__ gsrc = static_cast<const float*>(P.io) + p;           // BWD: upstream gradient of the normalised image
  const elem* __restrict__ tsrc = static_cast<const elem*>(P.io2) + p;            // target (nullable)
  float* __restrict__ dst = P.out + p;
  const bool has_t = P.io2 != nullptr;
  float l2w = 0.f;
  if (BWD && has_t && P.l2_up) l2w = __ldg(P.l2_up) * float(2.0 * P.loss_norm);   // d mse / d out = 2 (out - t) / n_elems
  for (int i = 0; i < N; ++i) {
    float tg[3] = {0.f, 0.f, 0.f}, up[3] = {0.f, 0.f, 0.f}, o3[3];
    if (valid && has_t) {
#pragma unroll
      for (int c = 0; c < 3; ++c) tg[c] = IoLoad<TGT>::decode(__ldg(tsrc + c * P.stride));
    }
    if (BWD && valid) {
#pragma unroll
      for (int c = 0; c < 3; ++c) up[c] = __fdiv_rn(__ldg(gsrc + c * P.stride), P.aff_std[c]);   // d((x - mean)/std)/dx
    }
    const LightGeom<float> lg = load_geom<COLOC>(s_geo, i);
    if (BWD) {
      if (has_t) shade_light<float, kVjpL2, COLOC, WANT_POW>(tx, lg, up, o3, g, tg, l2w);
      else shade_light<float, kVjp, COLOC, WANT_POW>(tx, lg, up, o3, g);
    } else {
      shade_light<float, kRender, COLOC, false>(tx, lg, up, o3, g);
      if (valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (has_t) {
            const float diff = o3[c] - tg[c];
            g.loss = __fmaf_rn(diff, diff, g.loss);
          }
          dst[c * P.stride] = __fdiv_rn(o3[c] - P.aff_mean[c], P.aff_std[c]);    // torchvision Normalize: sub_(mean).div_(std)
        }
      }
    }
    gsrc += step; tsrc += step; dst += step;
  }
}

template <bool BWD, bool WANT_POW, int TGT>
__global__ void __launch_bounds__(kThreads) norm_l2_kernel(const Params P) {
  extern __shared__ float4 s_geo[];
  __shared__ float s_red[kThreads / 32][4];
  const bool coloc = stage_lights(P, s_geo, threadIdx.x, kThreads);
  const long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool valid = p < P.texels;
  const long long pc = valid ? p : 0;
  float pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pw[c] = P.pow[c];
  float raw[9];
  bool outer[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { raw[k] = __ldg(P.tex + k * P.stride + pc); outer[k] = true; }
  Texel<float> tx;
  TexelAux<float> ax;
  {
    const int row = int(pc / P.res);
    const int col = int(pc - (long long)row * P.res);
    texel_position_rcp(row + P.row_offset, col, float(P.res), P.inv_res, P.size, tx.px, tx.py);
  }
  texel_prologue(raw, pw, tx, ax);
  Grads<float> g;
  grads_zero(g);
  if (coloc) norm_l2_lights<BWD, true, WANT_POW, TGT>(P, s_geo, tx, pc, valid, g);
  else norm_l2_lights<BWD, false, WANT_POW, TGT>(P, s_geo, tx, pc, valid, g);
  if (BWD) {
    float gt[9];
    if (coloc) texel_epilogue<float, true>(tx, ax, pw, g, P.scale, outer, gt);
    else texel_epilogue<float, false>(tx, ax, pw, g, P.scale, outer, gt);
    if (valid) {
#pragma unroll
      for (int k = 0; k < 9; ++k) P.out[k * P.stride + p] = gt[k];
    }
  }
  // block partials: L2 sum (forward) / light-power gradient (backward)
  float r[4] = {valid ? grads_loss(g) : 0.f, valid ? g.pw[0] : 0.f, valid ? g.pw[1] : 0.f, valid ? g.pw[2] : 0.f};
  warp_reduce4(r);
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

// Finalisation for texel_kernel launches (one small CTA).
__global__ void __launch_bounds__(256) finalize_kernel(const Params P, int n_blocks) {
  __shared__ double s[256][4];
  finalize_block(P, n_blocks, s, threadIdx.x, 256);
}

// ---------------------------------------------------------------------------------------------
// tile_kernel: persistent CTAs, TMA bulk copies into a shared-memory ring, mbarrier pipeline
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive and report whether this arrival was the last one of the phase (the state returned by mbarrier.arrive is the one
// BEFORE the arrival: one pending arrival left = ours).  SV_SELF_LAST is the pending count that identifies it.
#ifndef SV_SELF_LAST
#define SV_SELF_LAST 1
#endif
// Lane 0 arrives (predicated, no divergent branch); the other lanes get 0.
__device__ __forceinline__ unsigned mbar_arrive_is_last(unsigned long long* bar, int lane) {
  unsigned last;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 st;\n"
      ".reg .u32 c;\n"
      "setp.eq.u32 p, %1, 0;\n"
      "mov.u32 c, 0;\n"
      "@p mbarrier.arrive.shared::cta.b64 st, [%2];\n"
      "@p mbarrier.pending_count.b64 c, st;\n"
      "setp.eq.u32 q, c, %3;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(last)
      : "r"(lane), "r"(smem_u32(bar)), "n"(SV_SELF_LAST)
      : "memory");
  return last;
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)      // suspend-time hint: park the warp instead of polling
      : "memory");
  return ok != 0;
}
// Non-blocking probe (acquire): true if the phase with this parity has completed.
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


// The producer role is executed by a WHOLE warp in uniform control flow; only the instructions with side effects
// (bulk copies, mbarrier arrivals, bulk-group commits/waits) are predicated on one elected lane.  With a single thread
// running the role (`if (tid == X)`), ptxas cannot prove that the operands of the uniform-datapath UBLKCP instruction
// are warp-uniform and wraps every copy in a ~14-instruction lane-serialisation loop (PLOP3 / R2UR / BRA.U.ANY): the
// producer thread then executes ~15 dependent instructions per plane segment — 54-81 segments per tile — and becomes
// the bottleneck of the pipeline (ncu source view, profiles/r01_s2_ts_producer_bound.txt).
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Warp-uniform wait: every lane polls, the vote makes the loop exit provably uniform for the compiler.
__device__ __forceinline__ void mbar_wait_uniform(unsigned long long* bar, unsigned parity) {
  while (!__any_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
  }
}

// Chunk stream of one tile:  [tex] [lights 0..L-1] [lights L..2L-1] ... ([m] [v] in the fused mode), L = SH::kChunk.
template <int MODE>
__host__ __device__ __forceinline__ int chunks_per_tile(int n_lights, int chunk) {
  return 1 + (n_lights + chunk - 1) / chunk * (MODE == kModeVjpL2 ? 2 : 1) + (MODE == kModeL2Adam ? 2 : 0);
}

// Self-fed ring: stream chunk number `chunk` of this CTA's chunk stream (epochs x tiles x chunks_per_tile, the order the
// consumers walk it) into ring slot `slot`.  Executed by a whole warp in uniform control flow; one elected lane issues.
// Cross-epoch hazard (the texture / m / v planes of a tile were stored by this CTA one epoch earlier): every consumer warp
// fences its stores (publish_tiles) at least every `publish_every` tiles, and a chunk is only requested once ALL warps have
// handed back the chunk `slots` positions ahead of it — with >= 4 tiles per CTA (the launcher's condition for multi-epoch
// self-fed launches) that hand-back is later than the fence of the tile in question for every warp.
// The refill sits on the critical path of the slowest warp (it is the last to hand a slot back), so it has to be short:
// a chunk moves as kTile/256 two-dimensional TMA boxes ([planes][256 texels] through a tensor map of the planar array) —
// two instructions instead of 9-12 one-plane bulk copies (a first version with those spent ~2300 cycles per refill in the
// warp everybody else was waiting for: 7.08 instead of 4.48 ms at 4096^2 x 64) — and the chunk index is split into
// (epoch, tile, stage) with multiply-high by precomputed reciprocals.  Texels past the end of the array and planes past
// the last light are zero-filled by the TMA unit and count towards the transaction bytes, so every chunk of a kind has
// the same size.  Its constants live in shared memory (filled once per CTA), so the light loop carries no extra registers
// for it; not inlined — it runs once per slot per CTA, from eight call sites in the consumer.
constexpr int kSelfBoxW = 256;                             // texels per TMA box (the box-dimension limit)
struct SelfFeed {
  const CUtensorMap* tm_tex; const CUtensorMap* tm_io; const CUtensorMap* tm_m; const CUtensorMap* tm_v;
  long long lo;                                            // first texel of the CTA's first tile
  long long tile_step;                                     // texels from one of the CTA's tiles to the next
  unsigned ring, full;                                     // shared-memory addresses of the ring and of full[0]
  unsigned my_tiles, cpt, rcp_tiles, rcp_cpt;              // rcp_x = ceil(2^32 / x): n / x = umulhi(n, rcp_x) for n x < 2^32
  int n_light_chunks, epochs;
};
template <int MODE, int TGT, typename SH>
__device__ __noinline__ void self_fill(const SelfFeed* __restrict__ fp, unsigned chunk, unsigned slot) {
  const SelfFeed f = *fp;                                  // shared memory: the consumers keep none of this in registers
  typedef typename IoLoad<TGT>::elem elem;
  static_assert(MODE != kModeVjpL2, "the self-fed ring streams one slot per light chunk");
  static_assert(SH::kTile % kSelfBoxW == 0, "whole boxes per tile");
  const unsigned tile_idx = f.cpt == 1u ? chunk : __umulhi(chunk, f.rcp_cpt);          // running tile index over the epochs of this launch
  const unsigned stage = chunk - tile_idx * f.cpt;         // 0: textures, 1..: light chunks, then m, v
  const unsigned e = f.my_tiles == 1u ? tile_idx : __umulhi(tile_idx, f.rcp_tiles);
  if (e >= unsigned(f.epochs)) return;                     // past the end of the stream
  const unsigned local = tile_idx - e * f.my_tiles;
  const int x0 = int(f.lo + (long long)local * f.tile_step);
  const CUtensorMap* tm = f.tm_tex;
  int y0 = 0;
  unsigned box_bytes = 9u * kSelfBoxW * 4u;
  if (stage != 0u) {
    if (int(stage) <= f.n_light_chunks) {
      tm = f.tm_io;
      y0 = (int(stage) - 1) * SH::kChunk * 3;
      box_bytes = 3u * SH::kChunk * kSelfBoxW * unsigned(sizeof(elem));
    } else {
      tm = int(stage) == f.n_light_chunks + 1 ? f.tm_m : f.tm_v;
    }
  }
  const unsigned dst = f.ring + slot * unsigned(SH::kSlotBytes);
  const unsigned bar = f.full + slot * 8u;
  if (elect_one()) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * unsigned(SH::kTile / kSelfBoxW)) : "memory");
#pragma unroll
    for (int q = 0; q < SH::kTile / kSelfBoxW; ++q)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst + q * box_bytes),
                   "l"(reinterpret_cast<unsigned long long>(tm)), "r"(x0 + q * kSelfBoxW), "r"(y0), "r"(bar)
                   : "memory");
  }
}

template <int MODE, bool COLOC, bool WANT_POW, int TGT, typename SH>
__device__ __forceinline__ void tile_consumer(const Params& P, const float4* __restrict__ s_geo, unsigned char* ring,
                                               unsigned long long* full, unsigned long long* empty, float* stash, volatile unsigned* s_done, float (*s_red)[SH::kCW][4],
                                               SelfFeed* sfeed = nullptr) {
  typedef typename IoLoad<TGT>::elem elem;
  constexpr int LM = (MODE == kModeVjp) ? kVjp : (MODE == kModeVjpL2 ? kVjpL2 : (MODE == kModeNormFwd ? kRender : kL2));
  float l2w = 0.f;
  FixedDiv dstd[3];
  float amean[3] = {0.f, 0.f, 0.f};
  if (MODE == kModeNormFwd || MODE == kModeVjpL2) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { dstd[c] = make_fixed_div(P.aff_std[c]); amean[c] = MODE == kModeNormFwd ? P.aff_mean[c] : 0.f; }
  }
  if (MODE == kModeVjpL2) {
    l2w = __ldg(P.l2_up) * float(2.0 * P.loss_norm);        // d mse / d out = 2 (out - target) / n_elems, times the upstream scalar
  }
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = P.n_lights, S = P.slots;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  float pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pw[c] = P.pow[c];

  unsigned it = 0;        // running chunk index of this CTA
  unsigned slot = 0, phase = 0;
  // SV_PROBE_AHEAD (off): right after a slot is handed back, the NEXT slot's `full` barrier is probed without blocking, so
  // the common case (data already there) starts the next chunk without the try_wait round trip.  The ncu source view at
  // 4096^2 x 64 attributes 9 % of the warp samples to the branch on the try_wait result and its convergence bookkeeping,
  // yet hiding it is 1.8 % SLOWER at 64 lights and 2.1 % at 9: those samples are slack, not the critical path.
  bool ready = false;
#if SV_PROBE_AHEAD
  auto advance = [&]() { ++it; if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } ready = mbar_test(&full[slot], phase); };
  auto wait_full = [&]() { if (!ready) mbar_wait(&full[slot], phase); };
#else
  auto advance = [&]() { ++it; if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } };
#endif
  // multi-epoch launches: progress is published (with a fence) a few times per epoch, not per tile
  const TileMap tmap = tile_map<SH::kTile>(P, P.span, n_tiles);
  const unsigned my_tiles = tmap.count;
  constexpr bool kSelf = SH::kPW == 0;                     // self-fed ring: no producer warp (see SelfShape)
  // Where this thread's value of plane k sits in a ring slot.  Producer-fed ring: planes of kTile texels back to back.
  // Self-fed ring: a chunk arrives as kTile/256 two-dimensional TMA boxes of [planes][256 texels], one after the other.
  constexpr int kPS = kSelf ? kSelfBoxW : SH::kTile;       // plane stride (elements)
  const int tix9 = kSelf ? (tid / kSelfBoxW) * (9 * kSelfBoxW) + (tid % kSelfBoxW) : tid;                  // texture / m / v chunks
  const int tixL = kSelf ? (tid / kSelfBoxW) * (3 * SH::kChunk * kSelfBoxW) + (tid % kSelfBoxW) : tid;     // light chunks
  const unsigned cpt = unsigned(chunks_per_tile<MODE>(N, SH::kChunk));
  unsigned was_last = 0;                                   // lane 0: our hand-back of the previous chunk's slot completed its phase
  if constexpr (kSelf) {
    if (tid == 0) {                                        // the refill's constants (read by self_fill only)
      SelfFeed f;
      f.tm_tex = &P.tm_tex; f.tm_io = &P.tm_io; f.tm_m = &P.tm_m; f.tm_v = &P.tm_v;
      f.lo = tile_first<SH::kTile>(tmap, P, 0u, n_tiles);
      f.tile_step = tmap.spans ? (long long)SH::kTile : (long long)SH::kTile * gridDim.x;
      f.ring = smem_u32(ring); f.full = smem_u32(full);
      f.my_tiles = my_tiles; f.cpt = cpt;
      f.rcp_cpt = cpt > 1u ? unsigned(((1ull << 32) + cpt - 1u) / cpt) : 0u;
      f.rcp_tiles = my_tiles > 1u ? unsigned(((1ull << 32) + my_tiles - 1u) / my_tiles) : 0u;
      f.n_light_chunks = (N + SH::kChunk - 1) / SH::kChunk; f.epochs = P.epochs;
      *sfeed = f;
      __threadfence_block();
    }
    __syncwarp();                                          // warp 0 primes the ring below: lane 0's stores first
  }
  if constexpr (kSelf) {
    if (tid < 32 && my_tiles > 0u) {                       // warp 0 primes the ring: chunks 0 .. S-1
      for (unsigned c = 0; c < unsigned(S); ++c) self_fill<MODE, TGT, SH>(sfeed, c, c);
    }
  }
  // before waiting for chunk `it`: if this warp was the last of the CTA to hand back chunk it-1, its slot is free —
  // refill it with chunk it-1+S (the try_wait on the completed phase is the acquire side of the other warps' arrivals)
  auto service = [&]() {
    if constexpr (kSelf) {
      if (__any_sync(0xffffffffu, was_last != 0u)) {
        was_last = 0u;
        const unsigned ps = slot == 0u ? unsigned(S) - 1u : slot - 1u;
        mbar_wait_uniform(&empty[ps], slot == 0u ? phase ^ 1u : phase);
        self_fill<MODE, TGT, SH>(sfeed, it - 1u + unsigned(S), ps);
      }
    }
  };
#if !SV_PROBE_AHEAD
  auto wait_full = [&]() {
#if !SV_SELF_EAGER
    if constexpr (kSelf) service();
#endif
    mbar_wait(&full[slot], phase);
  };
#endif
  auto release = [&](unsigned s) {
    __syncwarp();
    if constexpr (kSelf) {
      was_last = mbar_arrive_is_last(&empty[s], lane);
#if SV_SELF_EAGER
      // refill at once (the round trip of the arrival overlaps the LDS latency the warp is about to wait for anyway):
      // chunk `it` was just handed back, chunk it + S goes into its slot
      if (__any_sync(0xffffffffu, was_last != 0u)) {
        mbar_wait_uniform(&empty[s], phase);
        self_fill<MODE, TGT, SH>(sfeed, it + unsigned(S), s);
      }
      was_last = 0u;
#endif
    } else {
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  };

  const unsigned publish_every = my_tiles >= 6 ? my_tiles / 3 : 1;
  unsigned tiles_finished = 0;
  for (int e = 0; e < P.epochs; ++e) {
  AdamStep<float> adam_e = P.adam;
  adam_e.step_size = P.step_size[e];
  adam_e.inv_sqrt_bc2 = P.inv_sqrt_bc2[e];
  float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (unsigned jt = 0; jt < my_tiles; ++jt) {
    const long long p0t = tile_first<SH::kTile>(tmap, P, jt, n_tiles);
    const long long rtile = p0t / SH::kTile;                  // round-robin mode only (peer-push ownership)
    const long long p = p0t + tid;
    const bool valid = p < tmap.hi;

    // ---- texel prologue ----
    float raw[9], t[9];
    bool outer[9];
    wait_full();
    {
      const float* s = reinterpret_cast<const float*>(ring + size_t(slot) * SH::kSlotBytes);
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = s[k * kPS + tix9];
    }
    release(slot);
    advance();
    if (!valid) {
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = 0.f;
    }
    clamp_outer<MODE>(raw, t, outer);
    Texel<float> tx;
    TexelAux<float> ax;
    {
      const long long pc = valid ? p : 0;
      int row, col;
      if (P.texels < (1ll << 32)) {
        row = int(__umul64hi((unsigned long long)pc, P.res_magic));        // exact floor(pc/res) for pc < 2^32
        col = int(unsigned(pc) - unsigned(row) * unsigned(P.res));
      } else {
        row = int(pc / P.res);
        col = int(pc - (long long)row * P.res);
      }
      texel_position_rcp(row + P.row_offset, col, float(P.res), P.inv_res, P.size, tx.px, tx.py);
    }
#if SV_STREAM_ONLY
    tx.px = t[0]; ax.mx = t[1];
#else
    texel_prologue(t, pw, tx, ax);
#endif
    if constexpr (MODE != kModeNormFwd) {
#if SV_STASH && !SV_STREAM_ONLY
    {
      float* st = stash + tid;
#pragma unroll
      for (int k = 0; k < 9; ++k) st[k * SH::kTile] = raw[k];
#pragma unroll
      for (int k = 0; k < 7; ++k) st[(9 + k) * SH::kTile] = ax.dpow[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) { st[(16 + k) * SH::kTile] = ax.d[k]; st[(19 + k) * SH::kTile] = ax.oms[k]; }
      st[22 * SH::kTile] = ax.rough; st[23 * SH::kTile] = ax.alpha;
      st[24 * SH::kTile] = ax.mx; st[25 * SH::kTile] = ax.my; st[26 * SH::kTile] = ax.mz; st[27 * SH::kTile] = ax.rlen;
    }
#endif
    }
    Grads<float> g;
    grads_zero(g);
    // kModeNormFwd: what norm_l2_kernel does per light after the render — (out - target)^2 into the L2 sum, and
    // torchvision's Normalize (sub_(mean).div_(std), IEEE division) of the rendered value into the output image
    auto norm_store = [&](const float o3[3], const float t3[3], int light) {
      if (valid) {
        float* __restrict__ dst = P.out + (size_t(light) * 3) * P.stride + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float diff = o3[c] - t3[c];
          g.loss = __fmaf_rn(diff, diff, g.loss);
          dst[c * P.stride] = div_rn(o3[c] - amean[c], dstd[c]);
        }
      }
    };

    // ---- lights, SH::kChunk per ring slot ----
    for (int i0 = 0; i0 < N; i0 += SH::kChunk) {
      float in[SH::kChunk][3];
      float tg[MODE == kModeVjpL2 ? SH::kChunk : 1][3];
      if constexpr (MODE == kModeVjpL2) {
        // first slot of the chunk: the upstream gradient of the normalised image, d((x - mean)/std)/dx = 1/std applied here
        // (correctly rounded like the unfused torch route's division: the fixed-divisor FMA sequence, see FixedDiv — it can
        // differ from IEEE division by an ulp only where the quotient or the residual is subnormal, |grad| < ~1e-30);
        // lights past N were not loaded: their values are never used
        wait_full();
        const float* sg = reinterpret_cast<const float*>(ring + size_t(slot) * SH::kSlotBytes);
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = (i0 + j < N) ? div_rn(sg[(j * 3 + c) * kPS + tixL], dstd[c]) : 0.f;
        }
        release(slot);
        advance();
      }
      wait_full();
      const elem* s = reinterpret_cast<const elem*>(ring + size_t(slot) * SH::kSlotBytes);
      if constexpr (MODE == kModeVjpL2) {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) tg[j][c] = (i0 + j < N) ? IoLoad<TGT>::decode(s[(j * 3 + c) * kPS + tixL]) : 0.f;
        }
        release(slot);
        advance();
        if (i0 + SH::kChunk <= N) {
#pragma unroll
          for (int j = 0; j < SH::kChunk; ++j) {
            float o3[3];
            shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g, tg[j], l2w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < SH::kChunk - 1; ++j) {
            if (i0 + j < N) {
              float o3[3];
              shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g, tg[j], l2w);
            }
          }
        }
        continue;
      }
      if (i0 + SH::kChunk <= N) {
        // full chunk: no per-light guards, so the three independent lights can be interleaved
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = IoLoad<TGT>::decode(s[(j * 3 + c) * kPS + tixL]);
        }
        release(slot);
        advance();
#if SV_STREAM_ONLY
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) g.loss += in[j][0] + in[j][1] + in[j][2];
#else
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
          float o3[3];
          shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g);
          if constexpr (MODE == kModeNormFwd) norm_store(o3, in[j], i0 + j);
        }
#endif
      } else {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = (i0 + j < N) ? IoLoad<TGT>::decode(s[(j * 3 + c) * kPS + tixL]) : 0.f;
        }
        release(slot);
        advance();
#pragma unroll
        for (int j = 0; j < SH::kChunk - 1; ++j) {       // a partial chunk holds at most SH::kChunk-1 lights
          if (i0 + j < N) {
            float o3[3];
            shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g);
            if constexpr (MODE == kModeNormFwd) norm_store(o3, in[j], i0 + j);
          }
        }
      }
    }

    if constexpr (MODE != kModeNormFwd) {
    // ---- epilogue ----
#if SV_STASH && !SV_STREAM_ONLY
    {
      const volatile float* st = stash + tid;
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = st[k * SH::kTile];
#pragma unroll
      for (int k = 0; k < 7; ++k) ax.dpow[k] = st[(9 + k) * SH::kTile];
#pragma unroll
      for (int k = 0; k < 3; ++k) { ax.d[k] = st[(16 + k) * SH::kTile]; ax.oms[k] = st[(19 + k) * SH::kTile]; }
      ax.rough = st[22 * SH::kTile]; ax.alpha = st[23 * SH::kTile];
      ax.mx = st[24 * SH::kTile]; ax.my = st[25 * SH::kTile]; ax.mz = st[26 * SH::kTile]; ax.rlen = st[27 * SH::kTile];
      clamp_outer<MODE>(raw, t, outer);
      ax.in3 = (t[3] >= -1.f) && (t[3] <= 1.f);
      ax.in4 = (t[4] >= -1.f) && (t[4] <= 1.f);
      ax.planar_free = (ax.mx * ax.mx + ax.my * ax.my) <= (1.f - float(kEps));
    }
#endif
    float gt[9];
#if SV_STREAM_ONLY
#pragma unroll
    for (int k = 0; k < 9; ++k) gt[k] = g.loss + tx.px;
#else
    texel_epilogue<float, COLOC>(tx, ax, pw, g, P.scale, outer, gt);
#endif
    if (MODE == kModeL2Adam) {
      float mk[9], vk[9];
      wait_full();
      {
        const float* s = reinterpret_cast<const float*>(ring + size_t(slot) * SH::kSlotBytes);
#pragma unroll
        for (int k = 0; k < 9; ++k) mk[k] = s[k * kPS + tix9];
      }
      release(slot);
      advance();
      wait_full();
      {
        const float* s = reinterpret_cast<const float*>(ring + size_t(slot) * SH::kSlotBytes);
#pragma unroll
        for (int k = 0; k < 9; ++k) vk[k] = s[k * kPS + tix9];
      }
      release(slot);
      advance();
      if (P.epochs > 1 && tiles_finished > 0 && tiles_finished % publish_every == 0) publish_tiles(s_done, tid >> 5, lane, tiles_finished);
      if (valid) {
        float* __restrict__ pt = P.tex + p;
        float* __restrict__ pm = P.m + p;
        float* __restrict__ pv = P.v + p;
#if SV_STREAM_ONLY
#pragma unroll
        for (int k = 0; k < 9; ++k) { raw[k] += gt[k]; mk[k] += 1.f; vk[k] += 1.f; }
#elif SV_PAIR_CHANNELS
        {
          // channels are updated two at a time with the packed FP32x2 instructions (same arithmetic, half the FMA-pipe
          // issue slots); MUFU sqrt/rcp stay per component
          AdamStep<V2> a2;
          a2.one_minus_b1 = V2(adam_e.one_minus_b1); a2.b2 = V2(adam_e.b2); a2.one_minus_b2 = V2(adam_e.one_minus_b2);
          a2.step_size = V2(adam_e.step_size); a2.inv_sqrt_bc2 = V2(adam_e.inv_sqrt_bc2); a2.eps = V2(adam_e.eps);
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            V2 pp(raw[k], raw[k + 1]), mm(mk[k], mk[k + 1]), vv(vk[k], vk[k + 1]);
            adam_update(pp, mm, vv, V2(gt[k], gt[k + 1]), a2);
            raw[k] = pp.x; raw[k + 1] = pp.y; mk[k] = mm.x; mk[k + 1] = mm.y; vk[k] = vv.x; vk[k + 1] = vv.y;
          }
          adam_update(raw[8], mk[8], vk[8], gt[8], adam_e);
        }
#else
#pragma unroll
        for (int k = 0; k < 9; ++k) adam_update(raw[k], mk[k], vk[k], gt[k], adam_e);
#endif
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          *pt = raw[k];
          *pm = mk[k];
          *pv = vk[k];
          pt += P.stride;
          pm += P.stride;
          pv += P.stride;
        }
      }
    } else if (valid) {
      float* __restrict__ po;
      long long ostride;
      if (MODE == kModeL2Grad && P.push_world > 0) {
        // reduce-scatter by push: this tile belongs to rank `owner`; write our partial into its slot [push_rank]
        const long long owner = (rtile * SH::kTile) / P.push_chunk;       // uniform per tile (chunk % tile == 0)
        po = P.push_recv[owner] + (size_t(P.push_rank) * 9) * P.push_chunk + (p - owner * P.push_chunk);
        ostride = P.push_chunk;
      } else {
        po = P.out + p;
        ostride = P.stride;
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        *po = gt[k];
        po += ostride;
      }
    }
    }
    if (valid) {
      loss_acc[0] += grads_loss(g);
#pragma unroll
      for (int c = 0; c < 3; ++c) loss_acc[1 + c] += g.pw[c];
    }
    ++tiles_finished;
  }
  epoch_end<SH::kCW>(P, e, loss_acc, s_red, tid >> 5, lane);
  }
}

// Two texels per thread, packed FP32x2 arithmetic (core instantiated with T = V2).  Thread `tid` owns texels
// 2*tid and 2*tid+1 of the tile: LDS.64 from the ring, STG.64 to global.
template <int TGT>
struct IoLoad2;
template <>
struct IoLoad2<SVBRDF_TARGET_F32> {
  static __device__ __forceinline__ V2 at(const unsigned char* slot, int plane, int tile, int tid) {
    const float2 q = reinterpret_cast<const float2*>(slot + size_t(plane) * tile * 4)[tid];
    return V2(q.x, q.y);
  }
};
template <>
struct IoLoad2<SVBRDF_TARGET_F16> {
  static __device__ __forceinline__ V2 at(const unsigned char* slot, int plane, int tile, int tid) {
    const float2 q = __half22float2(reinterpret_cast<const __half2*>(slot + size_t(plane) * tile * 2)[tid]);
    return V2(q.x, q.y);
  }
};
template <>
struct IoLoad2<SVBRDF_TARGET_U8> {
  static __device__ __forceinline__ V2 at(const unsigned char* slot, int plane, int tile, int tid) {
    const uchar2 q = reinterpret_cast<const uchar2*>(slot + size_t(plane) * tile)[tid];
    return V2(IoLoad<SVBRDF_TARGET_U8>::decode(q.x), IoLoad<SVBRDF_TARGET_U8>::decode(q.y));
  }
};

template <bool COLOC>
__device__ __forceinline__ LightGeom<V2> load_geom2(const float4* __restrict__ s_geo, int i) {
  LightGeom<V2> lg;
  const float4 a = s_geo[2 * i];
  lg.cx = V2(a.x); lg.cy = V2(a.y); lg.cz = V2(a.z); lg.cz2 = V2(a.w);
  if (!COLOC) {
    const float4 b = s_geo[2 * i + 1];
    lg.lx = V2(b.x); lg.ly = V2(b.y); lg.lz = V2(b.z); lg.lz2 = V2(b.w);
  }
  return lg;
}

template <int MODE, bool COLOC, bool WANT_POW, int TGT, typename SH>
__device__ __forceinline__ void tile_consumer2(const Params& P, const float4* __restrict__ s_geo, unsigned char* ring,
                                                unsigned long long* full, unsigned long long* empty, float* stash, volatile unsigned* s_done, float (*s_red)[SH::kCW][4]) {
  typedef V2 T;
  typedef Fm<V2> F;
  constexpr int LM = (MODE == kModeVjp) ? kVjp : kL2;
  constexpr bool kClamp = (MODE == kModeL2Grad || MODE == kModeL2Adam);
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = P.n_lights, S = P.slots;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  const T pw[3] = {T(P.pow[0]), T(P.pow[1]), T(P.pow[2])};
  const T scale(P.scale);

  unsigned slot = 0, phase = 0;
  auto advance = [&]() { if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } };
  auto release = [&](unsigned s) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[s]); };
  auto clamp9 = [&](const T raw[9], T t[9], M2 outer[9]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (kClamp) {   // the clamp of svbrdf.py:60
        outer[k] = F::mand(F::ge(raw[k], T(-1.f)), F::le(raw[k], T(1.f)));
        t[k] = F::min(F::max(raw[k], T(-1.f)), T(1.f));
      } else {
        outer[k] = F::mtrue();
        t[k] = raw[k];
      }
    }
  };

  // multi-epoch launches: progress is published (with a fence) a few times per epoch, not per tile
  const TileMap tmap = tile_map<SH::kTile>(P, P.span, n_tiles);
  const unsigned my_tiles = tmap.count;
  const unsigned publish_every = my_tiles >= 6 ? my_tiles / 3 : 1;
  unsigned tiles_finished = 0;
  for (int e = 0; e < P.epochs; ++e) {
  AdamStep<float> adam_e = P.adam;
  adam_e.step_size = P.step_size[e];
  adam_e.inv_sqrt_bc2 = P.inv_sqrt_bc2[e];
  float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (unsigned jt = 0; jt < my_tiles; ++jt) {
    const long long p = tile_first<SH::kTile>(tmap, P, jt, n_tiles) + 2 * tid;   // first of this thread's two texels
    const bool valid = p < tmap.hi;                           // range ends are multiples of 4: both or neither

    // ---- texel prologue ----
    T raw[9], t[9];
    M2 outer[9];
    mbar_wait(&full[slot], phase);
    {
      const unsigned char* s = ring + size_t(slot) * SH::kSlotBytes;
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = IoLoad2<SVBRDF_TARGET_F32>::at(s, k, SH::kTile, tid);
    }
    release(slot);
    advance();
    if (!valid) {
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = T(0.f);
    }
    clamp9(raw, t, outer);
    Texel<T> tx;
    TexelAux<T> ax;
    {
      const long long pc = valid ? p : 0;
      float px[2], py[2];
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const long long q = pc + l;
        int row, col;
        if (P.texels < (1ll << 32)) {
          row = int(__umul64hi((unsigned long long)q, P.res_magic));
          col = int(unsigned(q) - unsigned(row) * unsigned(P.res));
        } else {
          row = int(q / P.res);
          col = int(q - (long long)row * P.res);
        }
        texel_position_rcp(row + P.row_offset, col, float(P.res), P.inv_res, P.size, px[l], py[l]);
      }
      tx.px = T(px[0], px[1]);
      tx.py = T(py[0], py[1]);
    }
    texel_prologue(t, pw, tx, ax);
#if SV_STASH
    {
      float2* st = reinterpret_cast<float2*>(stash) + tid;
      auto put = [&](int k, const T& x) { st[k * SH::kConsumers] = make_float2(x.x, x.y); };
#pragma unroll
      for (int k = 0; k < 9; ++k) put(k, raw[k]);
#pragma unroll
      for (int k = 0; k < 7; ++k) put(9 + k, ax.dpow[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) { put(16 + k, ax.d[k]); put(19 + k, ax.oms[k]); }
      put(22, ax.rough); put(23, ax.alpha); put(24, ax.mx); put(25, ax.my); put(26, ax.mz); put(27, ax.rlen);
    }
#endif
    Grads<T> g;
    grads_zero(g);

    // ---- lights, SH::kChunk per ring slot ----
    for (int i0 = 0; i0 < N; i0 += SH::kChunk) {
      T in[SH::kChunk][3];
      mbar_wait(&full[slot], phase);
      const unsigned char* s = ring + size_t(slot) * SH::kSlotBytes;
      if (i0 + SH::kChunk <= N) {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = IoLoad2<TGT>::at(s, j * 3 + c, SH::kTile, tid);
        }
        release(slot);
        advance();
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
          T o3[3];
          shade_light<T, LM, COLOC, WANT_POW>(tx, load_geom2<COLOC>(s_geo, i0 + j), in[j], o3, g);
        }
      } else {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = (i0 + j < N) ? IoLoad2<TGT>::at(s, j * 3 + c, SH::kTile, tid) : T(0.f);
        }
        release(slot);
        advance();
#pragma unroll
        for (int j = 0; j < SH::kChunk - 1; ++j) {
          if (i0 + j < N) {
            T o3[3];
            shade_light<T, LM, COLOC, WANT_POW>(tx, load_geom2<COLOC>(s_geo, i0 + j), in[j], o3, g);
          }
        }
      }
    }

    // ---- epilogue ----
#if SV_STASH
    {
      const volatile float2* st = reinterpret_cast<const volatile float2*>(stash) + tid;
      auto get = [&](int k) { const float2 q = make_float2(st[k * SH::kConsumers].x, st[k * SH::kConsumers].y); return T(q.x, q.y); };
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = get(k);
#pragma unroll
      for (int k = 0; k < 7; ++k) ax.dpow[k] = get(9 + k);
#pragma unroll
      for (int k = 0; k < 3; ++k) { ax.d[k] = get(16 + k); ax.oms[k] = get(19 + k); }
      ax.rough = get(22); ax.alpha = get(23); ax.mx = get(24); ax.my = get(25); ax.mz = get(26); ax.rlen = get(27);
      clamp9(raw, t, outer);
      ax.in3 = F::mand(F::ge(t[3], T(-1.f)), F::le(t[3], T(1.f)));
      ax.in4 = F::mand(F::ge(t[4], T(-1.f)), F::le(t[4], T(1.f)));
      ax.planar_free = F::le(ax.mx * ax.mx + ax.my * ax.my, T(1.f - float(kEps)));
    }
#endif
    T gt[9];
    texel_epilogue<T, COLOC>(tx, ax, pw, g, scale, outer, gt);
    if (MODE == kModeL2Adam) {
      T mk[9], vk[9];
      mbar_wait(&full[slot], phase);
      {
        const unsigned char* s = ring + size_t(slot) * SH::kSlotBytes;
#pragma unroll
        for (int k = 0; k < 9; ++k) mk[k] = IoLoad2<SVBRDF_TARGET_F32>::at(s, k, SH::kTile, tid);
      }
      release(slot);
      advance();
      mbar_wait(&full[slot], phase);
      {
        const unsigned char* s = ring + size_t(slot) * SH::kSlotBytes;
#pragma unroll
        for (int k = 0; k < 9; ++k) vk[k] = IoLoad2<SVBRDF_TARGET_F32>::at(s, k, SH::kTile, tid);
      }
      release(slot);
      advance();
      if (P.epochs > 1 && tiles_finished > 0 && tiles_finished % publish_every == 0) publish_tiles(s_done, tid >> 5, lane, tiles_finished);
      if (valid) {
        AdamStep<T> a;
        a.one_minus_b1 = T(adam_e.one_minus_b1); a.b2 = T(adam_e.b2); a.one_minus_b2 = T(adam_e.one_minus_b2);
        a.step_size = T(adam_e.step_size); a.inv_sqrt_bc2 = T(adam_e.inv_sqrt_bc2); a.eps = T(adam_e.eps);
        float* __restrict__ pt = P.tex + p;
        float* __restrict__ pm = P.m + p;
        float* __restrict__ pv = P.v + p;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          T pk = raw[k];
          adam_update(pk, mk[k], vk[k], gt[k], a);
          *reinterpret_cast<float2*>(pt) = make_float2(pk.x, pk.y);
          *reinterpret_cast<float2*>(pm) = make_float2(mk[k].x, mk[k].y);
          *reinterpret_cast<float2*>(pv) = make_float2(vk[k].x, vk[k].y);
          pt += P.stride;
          pm += P.stride;
          pv += P.stride;
        }
      }
    } else if (valid) {
      float* __restrict__ po = P.out + p;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        *reinterpret_cast<float2*>(po) = make_float2(gt[k].x, gt[k].y);
        po += P.stride;
      }
    }
    if (valid) {
      loss_acc[0] += g.loss.x + g.loss.y + g.loss_g.x + g.loss_g.y;
#pragma unroll
      for (int c = 0; c < 3; ++c) loss_acc[1 + c] += g.pw[c].x + g.pw[c].y;
    }
    ++tiles_finished;
  }
  epoch_end<SH::kCW>(P, e, loss_acc, s_red, tid >> 5, lane);
  }
}

template <int MODE, int TGT, typename SH>
__device__ __forceinline__ void tile_producer(const Params& P, unsigned char* ring, unsigned long long* full, unsigned long long* empty,
                                              volatile unsigned* s_done) {
  typedef typename IoLoad<TGT>::elem elem;
  const int N = P.n_lights, S = P.slots;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  unsigned slot = 0, phase = 0;
  auto advance = [&]() { if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } };
  const bool leader = elect_one();                       // the whole warp runs this role; one lane issues

  const TileMap tmap = tile_map<SH::kTile>(P, P.span, n_tiles);
  const unsigned my_tiles = tmap.count;
  for (int e = 0; e < P.epochs; ++e) {
  for (unsigned local = 0; local < my_tiles; ++local) {
    if (e > 0) {
      // this tile was written by this CTA's consumers in epoch e-1: wait until every warp has stored (and fenced) it
      const unsigned publish_every = my_tiles >= 6 ? my_tiles / 3 : 1;
      const unsigned need = (unsigned(e - 1) * my_tiles + local + publish_every) / publish_every * publish_every;
      for (int w = 0; w < SH::kCW; ++w)
        while (!__any_sync(0xffffffffu, s_done[w] >= need)) {
        }
    }
    const long long p0 = tile_first<SH::kTile>(tmap, P, local, n_tiles);
    const unsigned len = unsigned(min((long long)SH::kTile, tmap.hi - p0));       // texels in this tile (multiple of 4)
    auto fill = [&](const void* base, unsigned elem_bytes, int planes) {
      // `planes` plane segments of `len` elements each, starting at element p0 of consecutive planes of `base`
      mbar_wait_uniform(&empty[slot], phase ^ 1);
      unsigned dst = smem_u32(ring) + slot * unsigned(SH::kSlotBytes);
      const unsigned bar = smem_u32(&full[slot]);
      const unsigned seg = len * elem_bytes;
      if (leader) mbar_expect_tx(&full[slot], seg * planes);
      const unsigned char* src = static_cast<const unsigned char*>(base) + size_t(p0) * elem_bytes;
      const size_t src_step = size_t(P.stride) * elem_bytes;
      const unsigned dst_step = unsigned(SH::kTile) * elem_bytes;
      for (int j = 0; j < planes; ++j) {
        if (leader)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                       "r"(seg), "r"(bar)
                       : "memory");
        dst += dst_step;
        src += src_step;
      }
      advance();
    };
    // view-sharded pull mode: the owner's replica is the authoritative copy of this tile's textures
    const float* tex_src = P.tex;
    if (MODE == kModeL2Grad && P.push_pull) tex_src = P.push_tex[p0 / P.push_chunk];
    fill(tex_src, 4, 9);
    for (int i0 = 0; i0 < N; i0 += SH::kChunk) {
      const int nl = min(SH::kChunk, N - i0);
      if (MODE == kModeVjpL2) {
        // two slots per chunk: upstream gradient (always fp32), then the targets
        fill(static_cast<const float*>(P.io) + size_t(i0) * 3 * P.stride, 4, 3 * nl);
        fill(static_cast<const elem*>(P.io2) + size_t(i0) * 3 * P.stride, sizeof(elem), 3 * nl);
      } else {
        fill(static_cast<const elem*>(P.io) + size_t(i0) * 3 * P.stride, sizeof(elem), 3 * nl);
      }
    }
    if (MODE == kModeL2Adam) {
      fill(P.m, 4, 9);
      fill(P.v, 4, 9);
    }
  }
  }
}

// ---------------------------------------------------------------------------------------------
// tile_kernel_ts: the same pipeline with the results leaving through TMA as well ("store-back").
//
// In tile_kernel every consumer thread ends a tile with up to 27 STG, each with its own 64-bit pointer increment
// (IADD3 + IADD3.X) — 81 issue slots per texel in an issue-bound kernel — and multi-epoch launches need generic->async
// proxy fences so the next epoch's TMA loads see those stores.  Here
//   * the tile's 9 texture planes live in a dedicated double-buffered shared-memory segment T[2] for the whole tile
//     (TMA-loaded, read by the prologue, re-read by the epilogue instead of a register/stash copy),
//   * the consumers overwrite T in place with the new parameters (fused mode) or the gradient (L2-grad / VJP), and the
//     Adam moments in place in the ring slots they arrived in: 27 STS with immediate offsets,
//   * one fence.proxy.async per thread + one mbarrier arrival per warp hands the tile to the producer thread, which
//     writes it out with 9-27 bulk copies (cp.async.bulk.global.shared::cta, SASS UBLKCP) and recycles the slots once
//     the copies have read shared memory.
// Loads and stores of one tile are now issued by the same thread in the same (async) proxy: the cross-epoch hazard is
// covered by cp.async.bulk.wait_group before a tile is loaded again.
// ---------------------------------------------------------------------------------------------
constexpr int kStashTs = 19;            // the raw texel stays in T: only the 19 derived values are parked
// 2-D TMA boxes: a tile (480 texels x 9 planes) moves as 3 boxes of 160 texels x 9 planes — 3 instructions per ring slot
// instead of 9 one-plane bulk copies, 27 per tile instead of 81 with the store-back.  160 texels = 5 whole warps, and a
// box of 9 x 160 floats is 5760 B = 45 x 128 B, so boxes are 128-byte aligned back to back.  Out-of-range texels of the
// last tile and planes past the last light are zero-filled on load and clipped on store by the TMA unit.
constexpr int kBoxW = 160, kBoxes = 3;
// Texture segments in flight per CTA.  With 2, a warp that is a tile ahead of the slowest warp of its CTA stalls at the
// next tile's segment (its buffer is released only when the tile two back has been written out): 15 % of all samples
// in the ncu source view; 3 lets the warps drift like the ring does.
#ifndef SV_TS_TBUF
#define SV_TS_TBUF 3
#endif
constexpr int kTBuf = SV_TS_TBUF;
template <typename E>
__host__ __device__ constexpr int box_stride_bytes() { return (9 * kBoxW * int(sizeof(E)) + 127) / 128 * 128; }

__device__ __forceinline__ void tma_load_box(const CUtensorMap* tm, unsigned dst_smem, int x, int y, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(reinterpret_cast<unsigned long long>(tm)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, unsigned src_smem, int x, int y) {
#if !defined(SV_TS_NOSTORE)
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<unsigned long long>(tm)),
               "r"(x), "r"(y), "r"(src_smem)
               : "memory");
#endif
}

__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_ns(unsigned long long* bar, unsigned parity, unsigned ns) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_n(unsigned long long* bar, unsigned n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// 1-D TMA bulk copy shared -> global (bytes % 16 == 0, both addresses 16-byte aligned), tracked by bulk async-groups.
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, unsigned src_smem, unsigned bytes) {
#if !defined(SV_TS_NOSTORE)
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
#endif
}

struct TsBars {
  unsigned long long* full;     // [slots] ring slot loaded (TMA complete_tx)
  unsigned long long* empty;    // [slots] ring slot free (CW consumer-warp arrivals, or CW arrivals by the producer after a store-back)
  unsigned long long* t_full;   // [kTBuf] texture segment loaded
  unsigned long long* t_empty;  // [kTBuf] texture segment stored and free
  unsigned long long* t_done;   // [kTBuf] all consumer warps have written the tile's results (CW arrivals)
};

template <int MODE, bool COLOC, bool WANT_POW, int TGT, typename SH>
__device__ __forceinline__ void tile_consumer_ts(const Params& P, const float4* __restrict__ s_geo, unsigned char* ring, float* T,
                                                  const TsBars B, float* stash, float (*s_red)[SH::kCW][4]) {
  typedef typename IoLoad<TGT>::elem elem;
  constexpr int LM = (MODE == kModeVjp) ? kVjp : kL2;
  static_assert(SH::kTile == kBoxW * kBoxes, "tile = 3 boxes of 160 texels");
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = P.n_lights, S = P.slots;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  float pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pw[c] = P.pow[c];
  // this thread's column inside a slot: box (5 warps each), position in the box; planes are kBoxW elements apart
  const int box = tid / kBoxW, wbox = tid - box * kBoxW;
  const int off_f = box * (box_stride_bytes<float>() / 4) + wbox;                // in floats
  const int off_e = box * box_stride_bytes<elem>() + wbox * int(sizeof(elem));  // in bytes, target element type

  unsigned slot = 0, phase = 0;
  auto advance = [&]() { if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } };
  auto release = [&](unsigned s) { __syncwarp(); if (lane == 0) mbar_arrive(&B.empty[s]); };

  unsigned t = 0;                                       // tiles of this CTA so far, over all epochs
  for (int e = 0; e < P.epochs; ++e) {
  AdamStep<float> adam_e = P.adam;
  adam_e.step_size = P.step_size[e];
  adam_e.inv_sqrt_bc2 = P.inv_sqrt_bc2[e];
  float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
    const long long p = tile * SH::kTile + tid;
    const bool valid = p < P.texels;
    const unsigned b = t % unsigned(kTBuf);
    float* Tb = T + size_t(b) * (SH::kSlotBytes / 4) + off_f;   // this thread's column of the tile's texture segment

    // ---- texel prologue ----
    float raw[9], tt[9];
    bool outer[9];
    mbar_wait(&B.t_full[b], (t / unsigned(kTBuf)) & 1u);
#pragma unroll
    for (int k = 0; k < 9; ++k) raw[k] = valid ? Tb[k * kBoxW] : 0.f;
    clamp_outer<MODE>(raw, tt, outer);
    Texel<float> tx;
    TexelAux<float> ax;
    {
      const long long pc = valid ? p : 0;
      int row, col;
      if (P.texels < (1ll << 32)) {
        row = int(__umul64hi((unsigned long long)pc, P.res_magic));        // exact floor(pc/res) for pc < 2^32
        col = int(unsigned(pc) - unsigned(row) * unsigned(P.res));
      } else {
        row = int(pc / P.res);
        col = int(pc - (long long)row * P.res);
      }
      texel_position_rcp(row + P.row_offset, col, float(P.res), P.inv_res, P.size, tx.px, tx.py);
    }
#if SV_STREAM_ONLY
    tx.px = tt[0]; ax.mx = tt[1];
#else
    texel_prologue(tt, pw, tx, ax);
    {
      float* st = stash + tid;
#pragma unroll
      for (int k = 0; k < 7; ++k) st[k * SH::kTile] = ax.dpow[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) { st[(7 + k) * SH::kTile] = ax.d[k]; st[(10 + k) * SH::kTile] = ax.oms[k]; }
      st[13 * SH::kTile] = ax.rough; st[14 * SH::kTile] = ax.alpha;
      st[15 * SH::kTile] = ax.mx; st[16 * SH::kTile] = ax.my; st[17 * SH::kTile] = ax.mz; st[18 * SH::kTile] = ax.rlen;
    }
#endif
    Grads<float> g;
    grads_zero(g);

    // ---- lights, SH::kChunk per ring slot ----
    for (int i0 = 0; i0 < N; i0 += SH::kChunk) {
      float in[SH::kChunk][3];
      mbar_wait(&B.full[slot], phase);
      const elem* s = reinterpret_cast<const elem*>(ring + size_t(slot) * SH::kSlotBytes + off_e);
      if (i0 + SH::kChunk <= N) {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = IoLoad<TGT>::decode(s[(j * 3 + c) * kBoxW]);
        }
        release(slot);
        advance();
#if SV_STREAM_ONLY
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) g.loss += in[j][0] + in[j][1] + in[j][2];
#else
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
          float o3[3];
          shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g);
        }
#endif
      } else {
#pragma unroll
        for (int j = 0; j < SH::kChunk; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) in[j][c] = (i0 + j < N) ? IoLoad<TGT>::decode(s[(j * 3 + c) * kBoxW]) : 0.f;
        }
        release(slot);
        advance();
#pragma unroll
        for (int j = 0; j < SH::kChunk - 1; ++j) {       // a partial chunk holds at most SH::kChunk-1 lights
          if (i0 + j < N) {
            float o3[3];
            shade_light<float, LM, COLOC, WANT_POW>(tx, load_geom<COLOC>(s_geo, i0 + j), in[j], o3, g);
          }
        }
      }
    }

    // ---- epilogue ----
    float gt[9];
#if SV_STREAM_ONLY
#pragma unroll
    for (int k = 0; k < 9; ++k) gt[k] = g.loss + tx.px;
#else
    {
      const volatile float* st = stash + tid;
      const volatile float* tv = Tb;
#pragma unroll
      for (int k = 0; k < 9; ++k) raw[k] = valid ? tv[k * kBoxW] : 0.f;
#pragma unroll
      for (int k = 0; k < 7; ++k) ax.dpow[k] = st[k * SH::kTile];
#pragma unroll
      for (int k = 0; k < 3; ++k) { ax.d[k] = st[(7 + k) * SH::kTile]; ax.oms[k] = st[(10 + k) * SH::kTile]; }
      ax.rough = st[13 * SH::kTile]; ax.alpha = st[14 * SH::kTile];
      ax.mx = st[15 * SH::kTile]; ax.my = st[16 * SH::kTile]; ax.mz = st[17 * SH::kTile]; ax.rlen = st[18 * SH::kTile];
      clamp_outer<MODE>(raw, tt, outer);
      ax.in3 = (tt[3] >= -1.f) && (tt[3] <= 1.f);
      ax.in4 = (tt[4] >= -1.f) && (tt[4] <= 1.f);
      ax.planar_free = (ax.mx * ax.mx + ax.my * ax.my) <= (1.f - float(kEps));
    }
    texel_epilogue<float, COLOC>(tx, ax, pw, g, P.scale, outer, gt);
#endif
    if (MODE == kModeL2Adam) {
      float mk[9], vk[9];
      mbar_wait(&B.full[slot], phase);
      float* sm = reinterpret_cast<float*>(ring + size_t(slot) * SH::kSlotBytes) + off_f;
#pragma unroll
      for (int k = 0; k < 9; ++k) mk[k] = sm[k * kBoxW];
      advance();                                           // not released: the slot is written back in place below
      mbar_wait(&B.full[slot], phase);
      float* sv = reinterpret_cast<float*>(ring + size_t(slot) * SH::kSlotBytes) + off_f;
#pragma unroll
      for (int k = 0; k < 9; ++k) vk[k] = sv[k * kBoxW];
      advance();
#if SV_STREAM_ONLY
#pragma unroll
      for (int k = 0; k < 9; ++k) { raw[k] += gt[k]; mk[k] += 1.f; vk[k] += 1.f; }
#elif SV_PAIR_CHANNELS
      {
        AdamStep<V2> a2;
        a2.one_minus_b1 = V2(adam_e.one_minus_b1); a2.b2 = V2(adam_e.b2); a2.one_minus_b2 = V2(adam_e.one_minus_b2);
        a2.step_size = V2(adam_e.step_size); a2.inv_sqrt_bc2 = V2(adam_e.inv_sqrt_bc2); a2.eps = V2(adam_e.eps);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          V2 pp(raw[k], raw[k + 1]), mm(mk[k], mk[k + 1]), vv(vk[k], vk[k + 1]);
          adam_update(pp, mm, vv, V2(gt[k], gt[k + 1]), a2);
          raw[k] = pp.x; raw[k + 1] = pp.y; mk[k] = mm.x; mk[k + 1] = mm.y; vk[k] = vv.x; vk[k + 1] = vv.y;
        }
        adam_update(raw[8], mk[8], vk[8], gt[8], adam_e);
      }
#else
#pragma unroll
      for (int k = 0; k < 9; ++k) adam_update(raw[k], mk[k], vk[k], gt[k], adam_e);
#endif
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        Tb[k * kBoxW] = raw[k];
        sm[k * kBoxW] = mk[k];
        sv[k * kBoxW] = vk[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) Tb[k * kBoxW] = gt[k];
    }
    // hand the tile to the producer: generic-proxy writes -> visible to the async proxy, then one arrival per warp
#if !defined(SV_TS_NOFENCE)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(&B.t_done[b]);
    if (valid) {
      loss_acc[0] += grads_loss(g);
#pragma unroll
      for (int c = 0; c < 3; ++c) loss_acc[1 + c] += g.pw[c];
    }
  }
  epoch_end<SH::kCW>(P, e, loss_acc, s_red, tid >> 5, lane);
  }
}

template <int MODE, int TGT, typename SH>
__device__ __forceinline__ void tile_producer_ts(const Params& P, unsigned char* ring, float* T, const TsBars B) {
  typedef typename IoLoad<TGT>::elem elem;
  const int N = P.n_lights, S = P.slots;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  const unsigned CL = unsigned((N + SH::kChunk - 1) / SH::kChunk);        // light chunks per tile
  const unsigned C = CL + (MODE == kModeL2Adam ? 2u : 0u);                     // ring chunks per tile
  const unsigned M = unsigned((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const unsigned total = M * unsigned(P.epochs);
  const CUtensorMap* const tm_store = (MODE == kModeL2Adam) ? &P.tm_tex : &P.tm_out;
  constexpr unsigned kBoxF = unsigned(box_stride_bytes<float>()), kBoxE = unsigned(box_stride_bytes<elem>());
  constexpr unsigned kBytesF = 9u * SH::kTile * 4u, kBytesE = 9u * SH::kTile * unsigned(sizeof(elem));   // per slot, OOB included
  unsigned slot = 0, phase = 0;
  auto advance = [&]() { if (++slot == unsigned(S)) { slot = 0; phase ^= 1; } };
  const bool leader = elect_one();                       // the whole warp runs this role; one lane issues (and owns the bulk groups)

  unsigned stored = 0;                                    // tiles written back so far
  // Write tile `stored` back if all consumer warps are done with it (non-blocking otherwise).
  auto service = [&]() -> bool {
    if (stored >= total) return false;
    const unsigned b = stored % unsigned(kTBuf);
    if (!__any_sync(0xffffffffu, mbar_test_wait(&B.t_done[b], (stored / unsigned(kTBuf)) & 1u))) return false;
    const long long tile = (long long)blockIdx.x + (long long)(stored % M) * gridDim.x;
    const int x0 = int(tile * SH::kTile);
    const unsigned src = smem_u32(T) + b * unsigned(SH::kSlotBytes);
    unsigned sm = 0, sv = 0;
    if (MODE == kModeL2Adam) {
      sm = (stored * C + CL) % unsigned(S);
      sv = sm + 1 == unsigned(S) ? 0u : sm + 1;
    }
    if (leader) {
#pragma unroll
      for (int q = 0; q < kBoxes; ++q) tma_store_box(tm_store, src + q * kBoxF, x0 + q * kBoxW, 0);
      if (MODE == kModeL2Adam) {
        const unsigned srcm = smem_u32(ring) + sm * unsigned(SH::kSlotBytes), srcv = smem_u32(ring) + sv * unsigned(SH::kSlotBytes);
#pragma unroll
        for (int q = 0; q < kBoxes; ++q) {
          tma_store_box(&P.tm_m, srcm + q * kBoxF, x0 + q * kBoxW, 0);
          tma_store_box(&P.tm_v, srcv + q * kBoxF, x0 + q * kBoxW, 0);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory has been read: the segments may be reused
      mbar_arrive(&B.t_empty[b]);
      if (MODE == kModeL2Adam) {
        mbar_arrive_n(&B.empty[sm], SH::kCW);
        mbar_arrive_n(&B.empty[sv], SH::kCW);
      }
    }
    ++stored;
    return true;
  };
  auto wait_serving = [&](unsigned long long* bar, unsigned parity) {
    while (!__any_sync(0xffffffffu, mbar_try_wait_ns(bar, parity, 400u))) service();
  };

  unsigned t = 0;
  for (int e = 0; e < P.epochs; ++e) {
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
    const int x0 = int(tile * SH::kTile);
    const unsigned b = t % unsigned(kTBuf);
    wait_serving(&B.t_empty[b], ((t / unsigned(kTBuf)) & 1u) ^ 1u);
    if (leader) {
      // e > 0: this tile was written back M >= 4 store groups ago; all but the most recent group are complete after this wait
      if (e > 0) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
      const unsigned dst = smem_u32(T) + b * unsigned(SH::kSlotBytes);
      const unsigned bar = smem_u32(&B.t_full[b]);
      mbar_expect_tx(&B.t_full[b], kBytesF);
#pragma unroll
      for (int q = 0; q < kBoxes; ++q) tma_load_box(&P.tm_tex, dst + q * kBoxF, x0 + q * kBoxW, 0, bar);
    }
    for (int i0 = 0; i0 < N; i0 += SH::kChunk) {
      wait_serving(&B.empty[slot], phase ^ 1);
      if (leader) {
        const unsigned dst = smem_u32(ring) + slot * unsigned(SH::kSlotBytes);
        const unsigned bar = smem_u32(&B.full[slot]);
        mbar_expect_tx(&B.full[slot], kBytesE);
#pragma unroll
        for (int q = 0; q < kBoxes; ++q) tma_load_box(&P.tm_io, dst + q * kBoxE, x0 + q * kBoxW, 3 * i0, bar);
      }
      advance();
      service();
    }
    if (MODE == kModeL2Adam) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        wait_serving(&B.empty[slot], phase ^ 1);
        if (leader) {
          const unsigned dst = smem_u32(ring) + slot * unsigned(SH::kSlotBytes);
          const unsigned bar = smem_u32(&B.full[slot]);
          mbar_expect_tx(&B.full[slot], kBytesF);
#pragma unroll
          for (int q = 0; q < kBoxes; ++q) tma_load_box(w == 0 ? &P.tm_m : &P.tm_v, dst + q * kBoxF, x0 + q * kBoxW, 0, bar);
        }
        advance();
      }
    }
  }
  }
  // drain: the last tiles are written back as their consumers finish
  while (stored < total) {
    const unsigned b = stored % unsigned(kTBuf);
    while (!__any_sync(0xffffffffu, mbar_try_wait_ns(&B.t_done[b], (stored / unsigned(kTBuf)) & 1u, 2000u))) {
    }
    service();
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, bool WANT_POW, int TGT, typename SH>
__global__ void __maxnreg__(SH::kMaxReg) tile_kernel_ts(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: ring [slots] | T [kTBuf] | barriers | light geometry 2*N float4 | stash [19 planes]
  unsigned char* ring = smem;
  float* T = reinterpret_cast<float*>(smem + size_t(P.slots) * SH::kSlotBytes);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + size_t(P.slots + kTBuf) * SH::kSlotBytes);
  TsBars B;
  B.full = bars;
  B.empty = B.full + P.slots;
  B.t_full = B.empty + P.slots;
  B.t_empty = B.t_full + 4;
  B.t_done = B.t_empty + 4;
  float4* s_geo = reinterpret_cast<float4*>(B.t_done + 4);
  float* stash = reinterpret_cast<float*>(s_geo + 2 * P.n_lights);
  __shared__ float s_red[2][SH::kCW][4];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) {
      mbar_init(&B.full[s], 1);
      mbar_init(&B.empty[s], SH::kCW);
    }
    for (int b = 0; b < kTBuf; ++b) {
      mbar_init(&B.t_full[b], 1);
      mbar_init(&B.t_empty[b], 1);
      mbar_init(&B.t_done[b], SH::kCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const bool coloc = stage_lights(P, s_geo, tid, SH::kThreads);     // includes __syncthreads

  if (tid >= SH::kConsumers) {
    tile_producer_ts<MODE, TGT, SH>(P, ring, T, B);               // whole warp, uniform control flow
  } else {
    if (coloc) tile_consumer_ts<MODE, true, WANT_POW, TGT, SH>(P, s_geo, ring, T, B, stash, s_red);
    else tile_consumer_ts<MODE, false, WANT_POW, TGT, SH>(P, s_geo, ring, T, B, stash, s_red);
  }
}

template <int MODE, bool WANT_POW, int TGT, typename SH>
__global__ void __maxnreg__(SH::kMaxReg) tile_kernel(const Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: ring [slots * 9216] | barriers full[slots], empty[slots] | light geometry 2*N float4 | reduction scratch
  unsigned char* ring = smem;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + size_t(P.slots) * SH::kSlotBytes);
  unsigned long long* empty = full + P.slots;
  float4* s_geo = reinterpret_cast<float4*>(empty + P.slots);
  float* stash = reinterpret_cast<float*>(s_geo + 2 * P.n_lights);
  __shared__ unsigned s_done[SH::kCW];                              // per consumer warp: tiles stored (and fenced) so far
  __shared__ float s_red[2][SH::kCW][4];

  const int tid = threadIdx.x;
  if (tid < SH::kCW) s_done[tid] = 0u;
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SH::kCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const bool coloc = stage_lights(P, s_geo, tid, SH::kThreads);     // includes __syncthreads

  bool producer = false;
  if constexpr (SH::kPW > 0) producer = tid >= SH::kConsumers;     // PW = 0: self-fed ring, every warp is a consumer
  if (producer) {
    if (SH::kPW == 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
      if (tid >= SH::kConsumers + 32) return;                      // only the first warp of the producer warpgroup works
    }
    if constexpr (SH::kPW > 0) tile_producer<MODE, TGT, SH>(P, ring, full, empty, s_done);    // whole warp, uniform control flow
  } else {
    if (SH::kPW == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SH::kConsumerRegs));
    if constexpr (SH::kLanes == 2) {
      if (coloc) tile_consumer2<MODE, true, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red);
      else tile_consumer2<MODE, false, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red);
    } else {
      if (coloc) tile_consumer<MODE, true, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red);
      else tile_consumer<MODE, false, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red);
    }
  }
}

// Self-fed shapes (PW = 0): every warp is a consumer.  Params is __grid_constant__ so that the tensor maps inside it can be
// addressed in place (cp.async.bulk.tensor takes a generic pointer to the descriptor).
template <int MODE, bool WANT_POW, int TGT, typename SH>
__global__ void __maxnreg__(SH::kMaxReg) tile_kernel_self(const __grid_constant__ Params P) {
  static_assert(SH::kPW == 0 && SH::kLanes == 1, "self-fed scalar shape");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + size_t(P.slots) * SH::kSlotBytes);
  unsigned long long* empty = full + P.slots;
  float4* s_geo = reinterpret_cast<float4*>(empty + P.slots);
  float* stash = reinterpret_cast<float*>(s_geo + 2 * P.n_lights);
  __shared__ unsigned s_done[SH::kCW];
  __shared__ float s_red[2][SH::kCW][4];
  __shared__ SelfFeed s_feed;
  const int tid = threadIdx.x;
  if (tid < SH::kCW) s_done[tid] = 0u;
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SH::kCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const bool coloc = stage_lights(P, s_geo, tid, SH::kThreads);     // includes __syncthreads
  if (coloc) tile_consumer<MODE, true, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red, &s_feed);
  else tile_consumer<MODE, false, WANT_POW, TGT, SH>(P, s_geo, ring, full, empty, stash, s_done, s_red, &s_feed);
}

// View-sharded push mode, second half: owner-side reduction of the `world` partial gradients (fixed rank order),
// Adam on the owned texels, new parameters stored into every rank's replica over NVLink (all-gather by push).
struct PushAdamParams {
  int world, rank;
  long long chunk, texels;
  const float* recv;      // this rank's receive buffer [world, 9, chunk]
  float* tex[8];          // peer-mapped replicas [9, texels]
  float* tex_mc;          // NVSwitch multicast mapping of the replicas (nullable)
  int local_only;         // pull mode: peers read the owner's replica themselves
  float* m;               // [9, chunk]
  float* v;
  AdamStep<float> adam;
};

__global__ void __launch_bounds__(256) reduce_adam_push_kernel(const PushAdamParams Q) {
  const long long first = (long long)Q.rank * Q.chunk;
  long long owned = Q.texels - first;
  if (owned > Q.chunk) owned = Q.chunk;
  if (owned <= 0) return;
  const long long n4 = owned / 4;                           // chunk, texels are multiples of 4
  const long long total = n4 * 9;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long k = idx / n4, i4 = idx - k * n4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < Q.world; ++r) {                     // fixed order: deterministic, identical on every rank
      const float4 q = __ldcs(reinterpret_cast<const float4*>(Q.recv + (size_t(r) * 9 + k) * Q.chunk) + i4);
      g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
    }
    float4* pm = reinterpret_cast<float4*>(Q.m + k * Q.chunk) + i4;
    float4* pv = reinterpret_cast<float4*>(Q.v + k * Q.chunk) + i4;
    const size_t toff = size_t(k) * Q.texels + first;
    float4 p4 = reinterpret_cast<const float4*>(Q.tex[Q.rank] + toff)[i4], m4 = *pm, v4 = *pv;
    adam_update(p4.x, m4.x, v4.x, g.x, Q.adam);
    adam_update(p4.y, m4.y, v4.y, g.y, Q.adam);
    adam_update(p4.z, m4.z, v4.z, g.z, Q.adam);
    adam_update(p4.w, m4.w, v4.w, g.w, Q.adam);
    *pm = m4;
    *pv = v4;
    if (Q.local_only) {
      reinterpret_cast<float4*>(Q.tex[Q.rank] + toff)[i4] = p4;
    } else if (Q.tex_mc) {
      // one store, replicated to every GPU's replica by the NVSwitch (NVLS multicast)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4*>(Q.tex_mc + toff) + i4),
                   "f"(p4.x), "f"(p4.y), "f"(p4.z), "f"(p4.w)
                   : "memory");
    } else {
      for (int q = 0; q < Q.world; ++q) reinterpret_cast<float4*>(Q.tex[q] + toff)[i4] = p4;    // own replica + peers
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
  if (size_t(g->n_lights) * 2 * sizeof(float4) > 64 * 1024) return SVBRDF_E_UNSUPPORTED;   // 2048 lights per call
  return 0;
}

static Params base_params(const svbrdf_geom_t* g) {
  Params P{};
  P.cam = g->camera_pos;
  P.light = g->light_pos;
  P.pow = const_cast<float*>(g->light_pow);
  P.texels = (long long)g->rows * g->res;
  P.stride = g->plane_stride ? g->plane_stride : P.texels;
  P.size = g->size;
  P.res = g->res;
  P.inv_res = 1.0f / float(g->res);
  P.res_magic = ~0ull / (unsigned long long)g->res + 1ull;
  P.row_offset = g->row_offset;
  P.n_lights = g->n_lights;
  P.epochs = 1;
  return P;
}

static inline int texel_blocks(const Params& P) { return int((P.texels + kThreads - 1) / kThreads); }

// Workspace layout: [max(texel blocks, persistent CTAs)][4] floats of partials, then the finish counter.
static inline size_t partial_rows(long long texels) {
  const size_t tb = size_t((texels + kThreads - 1) / kThreads);          // texel_kernel: one row per CTA
  const size_t tile_rows = size_t(kMaxEpochs) * kRowsPerEpoch;             // tile_kernel: one row per warp and epoch
  return tb > tile_rows ? tb : tile_rows;
}

struct Device {
  int index = 0;
  int sms = 0;
  int smem_optin = 0;
};
// Attributes of the current device, queried once per device and process (the single-epoch launch path is called every
// ~70 us; with one process per GPU on an 8-GPU box the per-launch driver queries showed up as a 17 % longer step).
// Racing first calls write the same values.
static int device_info(Device* d) {
  constexpr int kMaxDev = 64;
  static int cached_sms[kMaxDev] = {0}, cached_smem[kMaxDev] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return int(e);
  d->index = dev;
  if (dev >= 0 && dev < kMaxDev && cached_sms[dev] > 0 && cached_smem[dev] > 0) {
    d->sms = cached_sms[dev];
    d->smem_optin = cached_smem[dev];
    return 0;
  }
  if ((e = cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return int(e);
  if ((e = cudaDeviceGetAttribute(&d->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return int(e);
  if (dev >= 0 && dev < kMaxDev) {
    cached_smem[dev] = d->smem_optin;
    cached_sms[dev] = d->sms;
  }
  return 0;
}

// cudaFuncSetAttribute(max dynamic shared memory, carve-out) once per (kernel instantiation, device, size) instead of on
// every launch.  `slot` is a per-instantiation static array indexed by device.
static int set_smem_once(const void* kern, int dev, size_t smem, int* slot) {
  if (dev >= 0 && dev < 64 && slot[dev] == int(smem)) return 0;
  if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return int(e);
  if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) return int(e);
  if (dev >= 0 && dev < 64) slot[dev] = int(smem);
  return 0;
}

template <int MODE, bool WANT_POW, int TGT>
static int launch_texel_one(const Params& P, cudaStream_t st);

// texel_kernel handles one epoch per launch: a multi-epoch request is unrolled on the host.
template <int MODE, bool WANT_POW, int TGT>
static int launch_texel(const Params& P, cudaStream_t st) {
  if (P.epochs <= 1) {
    Params Q = P;
    Q.adam.step_size = P.step_size[0];
    Q.adam.inv_sqrt_bc2 = P.inv_sqrt_bc2[0];
    return launch_texel_one<MODE, WANT_POW, TGT>(Q, st);
  }
  for (int e = 0; e < P.epochs; ++e) {
    Params Q = P;
    Q.epochs = 1;
    Q.adam.step_size = Q.step_size[0] = P.step_size[e];
    Q.adam.inv_sqrt_bc2 = Q.inv_sqrt_bc2[0] = P.inv_sqrt_bc2[e];
    Q.loss_out = P.loss_out ? P.loss_out + e : nullptr;
    if (int err = launch_texel_one<MODE, WANT_POW, TGT>(Q, st)) return err;
  }
  return 0;
}

template <int MODE, bool WANT_POW, int TGT>
static int launch_texel_one(const Params& P, cudaStream_t st) {
  const size_t smem = size_t(P.n_lights) * 2 * sizeof(float4);
  auto kern = texel_kernel<MODE, WANT_POW, TGT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
  }
  kern<<<texel_blocks(P), kThreads, smem, st>>>(P);
  if (cudaError_t e = cudaGetLastError()) return int(e);
  constexpr bool kHasLoss = (MODE == kModeL2Grad || MODE == kModeL2Adam);
  if ((kHasLoss && P.loss_out) || P.grad_pow || P.pow_state) {
    finalize_kernel<<<1, 256, 0, st>>>(P, texel_blocks(P));
    return int(cudaGetLastError());
  }
  return 0;
}

// The TMA path needs 16-byte aligned plane segments: base pointers, plane stride and tile length.
template <int TGT>
static bool tma_ok(const Params& P) {
  const size_t eb = sizeof(typename IoLoad<TGT>::elem);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(P.tex) || !al(P.io) || (P.m && !al(P.m)) || (P.v && !al(P.v))) return false;
  if ((P.stride * 4) % 16 != 0 || (size_t(P.stride) * eb) % 16 != 0) return false;
  if ((P.texels * 4) % 16 != 0 || (size_t(P.texels) * eb) % 16 != 0) return false;
  return true;
}

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

// What a tile launch does when the ring cannot be set up (too many lights for the shared-memory budget, more CTAs than
// workspace rows): the one-thread-per-texel kernel — except for the consumer modes, whose LDG kernels live in
// launch_norm_l2; those get kNoTilePath back and take them.
constexpr int kNoTilePath = -12345;
template <int MODE, bool WANT_POW, int TGT>
static int tile_fallback(const Params& P, cudaStream_t st) {
  if constexpr (MODE == kModeVjpL2 || MODE == kModeNormFwd) return kNoTilePath;
  else return launch_texel<MODE, WANT_POW, TGT>(P, st);
}

template <int MODE, bool WANT_POW, int TGT, typename SH>
static int launch_tile_shape(Params P, cudaStream_t st) {
  const bool trace = env_int("SVBRDF_B200_TRACE", 0) != 0;
  Device d;
  if (int e = device_info(&d)) return e;
  const int ctas_per_sm = env_int("SVBRDF_B200_CTAS_PER_SM", 1);
  const size_t geo = size_t(P.n_lights) * 2 * sizeof(float4);
  const size_t stash_bytes = (SV_STASH && MODE != kModeNormFwd) ? size_t(kStashFloats) * SH::kTile * 4 : 0;   // the forward-only mode parks nothing
  const size_t fixed = geo + stash_bytes + 64 + 2 * 8 * 64;            // barriers (<= 64 slots) + padding
  const size_t static_smem = 1024;                                     // s_done + s_red (static __shared__)
  const size_t budget = size_t(d.smem_optin + 1024) / ctas_per_sm - 1024 - static_smem;
  int slots = env_int("SVBRDF_B200_SLOTS", 0);
  if (slots <= 0) slots = int((budget - fixed) / SH::kSlotBytes);
  const int need = chunks_per_tile<MODE>(P.n_lights, SH::kChunk);
  if (slots > 2 * need) slots = 2 * need;                               // two whole tiles in flight is plenty
  if (slots > 64) slots = 64;
  if (slots < 2) return tile_fallback<MODE, WANT_POW, TGT>(P, st);
  P.slots = slots;
  const size_t smem = size_t(slots) * SH::kSlotBytes + size_t(slots) * 16 + geo + stash_bytes + 16;
  if (smem + static_smem > size_t(d.smem_optin)) return tile_fallback<MODE, WANT_POW, TGT>(P, st);
  auto kern = tile_kernel<MODE, WANT_POW, TGT, SH>;
  static int smem_set[64] = {0};
  if (int e = set_smem_once(reinterpret_cast<const void*>(kern), d.index, smem, smem_set)) return e;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  long long grid = (long long)d.sms * ctas_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  // equal load per CTA: one contiguous range of ceil(texels/grid) texels (rounded to 32) instead of whole tiles dealt
  // round-robin; the peer-sharded modes keep tile-aligned ownership
  P.span = 0;
  // measured (profiles/r01_s2_variants_spans.txt): 72.3 -> 71.3 us at 1024^2 x 9, 265.9 -> 264.0 us at 2048^2 x 9, but 23.6 -> 26.0 us
  // at 512^2 (3.7 tiles per CTA: a partial tile costs almost a whole tile's latency), hence only from 8 tiles per CTA up
  if (P.push_world == 0 && P.tile_rotate == 0 && n_tiles >= 8 * grid && env_int("SVBRDF_B200_SPANS", 1)) {
    const long long ctas = (long long)d.sms * ctas_per_sm;
    long long span = (P.texels + ctas - 1) / ctas;
    span = (span + 31) / 32 * 32;
    if (span < 32 * 4) span = 32 * 4;                                   // at least 4 warps' worth per CTA
    P.span = span;
    grid = (P.texels + span - 1) / span;
  }
  // fewest tiles any CTA gets (span mode: the last CTA takes what is left)
  long long tiles_per_cta = n_tiles / grid;
  if (P.span) {
    const long long last = P.texels - (grid - 1) * P.span;
    tiles_per_cta = (last + SH::kTile - 1) / SH::kTile;
    if (grid > 1 && (P.span + SH::kTile - 1) / SH::kTile < tiles_per_cta) tiles_per_cta = (P.span + SH::kTile - 1) / SH::kTile;
  }
  P.counters = reinterpret_cast<unsigned int*>(P.partials + partial_rows(P.texels) * 4);
  if (grid > kRowsPerEpoch) return tile_fallback<MODE, WANT_POW, TGT>(P, st);
  if (P.epochs > 1 && tiles_per_cta < 3) {                            // too few tiles per CTA for the lagging publication: one launch per epoch
    Params Q = P;
    for (int e = 0; e < P.epochs; ++e) {
      Q.epochs = 1;
      Q.step_size[0] = P.step_size[e];
      Q.inv_sqrt_bc2[0] = P.inv_sqrt_bc2[e];
      Q.loss_out = P.loss_out ? P.loss_out + e : nullptr;
      if (int err = launch_tile_shape<MODE, WANT_POW, TGT, SH>(Q, st)) return err;
    }
    return 0;
  }
  if (trace) fprintf(stderr, "[svbrdf] tile_kernel mode %d lanes %d tile %d slots %d smem %zu grid %lld\n", MODE, SH::kLanes, SH::kTile, slots, smem, grid);
  // the finish tickets must be zero when the launch starts: the last CTA of every epoch re-zeroes its own, but a workspace
  // a C-ABI caller did not clear, or one an aborted launch left dirty, would silently corrupt the loss — 256 bytes per launch
  if (cudaError_t e = cudaMemsetAsync(P.counters, 0, sizeof(unsigned int) * kMaxEpochs, st)) return int(e);
  kern<<<int(grid), SH::kThreads, smem, st>>>(P);
  return int(cudaGetLastError());
}

// 2-D TMA descriptor of a planar array: dim0 = texels of the band (contiguous), dim1 = planes (plane stride apart),
// box = 160 texels x 9 planes.  cuTensorMapEncodeTiled is a pure host-side encoder fetched through the runtime, so the
// library does not link libcuda.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}
static bool make_plane_map(CUtensorMap* tm, const void* base, int elem_bytes, long long texels, long long planes, long long stride_elems,
                           int box_w = kBoxW, int box_h = 9) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc || !base) return false;
  const cuuint64_t eb = cuuint64_t(elem_bytes);             // 1 = uint8, 2 = float16, 4 = float32
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  const cuuint64_t dims[2] = {cuuint64_t(texels), cuuint64_t(planes)};
  const cuuint64_t strides[1] = {cuuint64_t(stride_elems) * eb};
  const cuuint32_t box[2] = {cuuint32_t(box_w), cuuint32_t(box_h)};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Self-fed ring (SelfShape): 16 consumer warps, no producer warp, chunks loaded as 2-D TMA boxes through tensor maps of the
// planar arrays.  `FB` is the producer-fed shape to fall back to when the maps cannot be built (>= 2^31 texels, no driver
// entry point) or the ring would be too shallow.
template <int MODE, bool WANT_POW, int TGT, typename SH, typename FB>
static int launch_tile_self(Params P, cudaStream_t st) {
  const bool trace = env_int("SVBRDF_B200_TRACE", 0) != 0;
  Device d;
  if (int e = device_info(&d)) return e;
  const size_t geo = size_t(P.n_lights) * 2 * sizeof(float4);
  const size_t stash_bytes = SV_STASH ? size_t(kStashFloats) * SH::kTile * 4 : 0;
  const size_t fixed = geo + stash_bytes + 64 + 2 * 8 * 64;
  const size_t static_smem = 1024;
  const size_t budget = size_t(d.smem_optin + 1024) - 1024 - static_smem;
  int slots = env_int("SVBRDF_B200_SLOTS", 0);
  if (slots <= 0) slots = int((budget - fixed) / SH::kSlotBytes);
  const int need = chunks_per_tile<MODE>(P.n_lights, SH::kChunk);
  if (slots > 2 * need) slots = 2 * need;
  if (slots > 64) slots = 64;
  const size_t smem = size_t(slots) * SH::kSlotBytes + size_t(slots) * 16 + geo + stash_bytes + 16;
  if (slots < 3 || smem + static_smem > size_t(d.smem_optin)) return launch_tile_shape<MODE, WANT_POW, TGT, FB>(P, st);
  P.slots = slots;
  {
    bool ok = P.texels < (1ll << 31) - SH::kTile &&
              make_plane_map(&P.tm_tex, P.tex, 4, P.texels, 9, P.stride, kSelfBoxW, 9) &&
              make_plane_map(&P.tm_io, P.io, int(sizeof(typename IoLoad<TGT>::elem)), P.texels, 3ll * P.n_lights, P.stride, kSelfBoxW, 3 * SH::kChunk);
    if (MODE == kModeL2Adam)
      ok = ok && make_plane_map(&P.tm_m, P.m, 4, P.texels, 9, P.stride, kSelfBoxW, 9) &&
           make_plane_map(&P.tm_v, P.v, 4, P.texels, 9, P.stride, kSelfBoxW, 9);
    if (!ok) return launch_tile_shape<MODE, WANT_POW, TGT, FB>(P, st);
  }
  auto kern = tile_kernel_self<MODE, WANT_POW, TGT, SH>;
  static int smem_set[64] = {0};
  if (int e = set_smem_once(reinterpret_cast<const void*>(kern), d.index, smem, smem_set)) return e;
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  long long grid = d.sms;
  if (grid > n_tiles) grid = n_tiles;
  P.span = 0;
  if (n_tiles >= 8 * grid && env_int("SVBRDF_B200_SPANS", 1)) {          // equal load per CTA, as in launch_tile_shape
    long long span = (P.texels + d.sms - 1) / d.sms;
    span = (span + 31) / 32 * 32;
    if (span < 32 * 4) span = 32 * 4;
    P.span = span;
    grid = (P.texels + span - 1) / span;
  }
  long long tiles_per_cta = n_tiles / grid;
  if (P.span) {
    const long long last = P.texels - (grid - 1) * P.span;
    tiles_per_cta = (last + SH::kTile - 1) / SH::kTile;
    if (grid > 1 && (P.span + SH::kTile - 1) / SH::kTile < tiles_per_cta) tiles_per_cta = (P.span + SH::kTile - 1) / SH::kTile;
  }
  P.counters = reinterpret_cast<unsigned int*>(P.partials + partial_rows(P.texels) * 4);
  if (grid > kRowsPerEpoch) return launch_texel<MODE, WANT_POW, TGT>(P, st);
  // multi-epoch launches need >= 4 tiles per CTA: a tile's reload is requested when the chunk `slots` positions ahead of
  // it (<= 2 tiles) has been handed back by every warp, and every warp must have fenced the tile's stores before that
  if (P.epochs > 1 && tiles_per_cta < 4) {
    Params Q = P;
    for (int e = 0; e < P.epochs; ++e) {
      Q.epochs = 1;
      Q.step_size[0] = P.step_size[e];
      Q.inv_sqrt_bc2[0] = P.inv_sqrt_bc2[e];
      Q.loss_out = P.loss_out ? P.loss_out + e : nullptr;
      if (int err = launch_tile_self<MODE, WANT_POW, TGT, SH, FB>(Q, st)) return err;
    }
    return 0;
  }
  if (trace) fprintf(stderr, "[svbrdf] tile_kernel_self mode %d tile %d slots %d smem %zu grid %lld\n", MODE, SH::kTile, slots, smem, grid);
  if (cudaError_t e = cudaMemsetAsync(P.counters, 0, sizeof(unsigned int) * kMaxEpochs, st)) return int(e);
  kern<<<int(grid), SH::kThreads, smem, st>>>(P);
  return int(cudaGetLastError());
}

// tile_kernel_ts (results written back by TMA): scalar shape, local outputs only (the peer-push mode keeps per-thread
// stores to the owner's memory).
template <int MODE, bool WANT_POW, int TGT, typename SH>
static int launch_tile_ts(Params P, cudaStream_t st) {
  const bool trace = env_int("SVBRDF_B200_TRACE", 0) != 0;
  Device d;
  if (int e = device_info(&d)) return e;
  const size_t geo = size_t(P.n_lights) * 2 * sizeof(float4);
  const size_t stash_bytes = size_t(kStashTs) * SH::kTile * 4;
  const size_t static_smem = 1024;                                     // s_red (static __shared__) + slack
  const size_t fixed = size_t(kTBuf) * SH::kSlotBytes + geo + stash_bytes + (2 * 64 + 12) * 8 + 64;
  if (size_t(d.smem_optin) < fixed + static_smem + 3 * size_t(SH::kSlotBytes)) return launch_tile_shape<MODE, WANT_POW, TGT, SH>(P, st);
  int slots = env_int("SVBRDF_B200_SLOTS", 0);
  if (slots <= 0) slots = int((size_t(d.smem_optin) - static_smem - fixed) / SH::kSlotBytes);
  const int need = chunks_per_tile<MODE>(P.n_lights, SH::kChunk) - 1;              // the texture chunk is not in the ring
  if (slots > 2 * need) slots = 2 * need;
  if (slots > 64) slots = 64;
  if (slots < 3) return launch_tile_shape<MODE, WANT_POW, TGT, SH>(P, st);
  P.slots = slots;
  const size_t smem = size_t(slots + kTBuf) * SH::kSlotBytes + size_t(2 * slots + 12) * 8 + geo + stash_bytes + 16;
  if (smem + static_smem > size_t(d.smem_optin)) return launch_tile_shape<MODE, WANT_POW, TGT, SH>(P, st);
  {
    const bool u8 = TGT == SVBRDF_TARGET_U8;
    bool ok = P.texels < (1ll << 31) && make_plane_map(&P.tm_tex, P.tex, 4, P.texels, 9, P.stride) &&
              make_plane_map(&P.tm_io, P.io, u8 ? 1 : 4, P.texels, 3ll * P.n_lights, P.stride);
    if (MODE == kModeL2Adam) ok = ok && make_plane_map(&P.tm_m, P.m, 4, P.texels, 9, P.stride) && make_plane_map(&P.tm_v, P.v, 4, P.texels, 9, P.stride);
    else ok = ok && make_plane_map(&P.tm_out, P.out, 4, P.texels, 9, P.stride);
    if (!ok) return launch_tile_shape<MODE, WANT_POW, TGT, SH>(P, st);
  }
  auto kern = tile_kernel_ts<MODE, WANT_POW, TGT, SH>;
  if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return int(e);
  if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) return int(e);
  const long long n_tiles = (P.texels + SH::kTile - 1) / SH::kTile;
  long long grid = d.sms;
  if (grid > n_tiles) grid = n_tiles;
  P.counters = reinterpret_cast<unsigned int*>(P.partials + partial_rows(P.texels) * 4);
  if (grid > kRowsPerEpoch) return launch_texel<MODE, WANT_POW, TGT>(P, st);
  if (P.epochs > 1 && n_tiles / grid < 4) {          // a tile must be >= 3 store groups old before it is loaded again
    Params Q = P;
    for (int e = 0; e < P.epochs; ++e) {
      Q.epochs = 1;
      Q.step_size[0] = P.step_size[e];
      Q.inv_sqrt_bc2[0] = P.inv_sqrt_bc2[e];
      Q.loss_out = P.loss_out ? P.loss_out + e : nullptr;
      if (int err = launch_tile_ts<MODE, WANT_POW, TGT, SH>(Q, st)) return err;
    }
    return 0;
  }
  if (trace) fprintf(stderr, "[svbrdf] tile_kernel_ts mode %d tile %d slots %d smem %zu grid %lld\n", MODE, SH::kTile, slots, smem, grid);
  if (cudaError_t e = cudaMemsetAsync(P.counters, 0, sizeof(unsigned int) * kMaxEpochs, st)) return int(e);
  kern<<<int(grid), SH::kThreads, smem, st>>>(P);
  return int(cudaGetLastError());
}

// Lights per ring slot (TileShape CHUNK) for a light count: whole slots are shaded unguarded and interleaved, the lights
// left over in a partial slot one by one behind guards — so fewest left-over lights first, then the larger slot.
// SVBRDF_B200_CHUNK=3|4 overrides (development).
static int pick_chunk(int n_lights) {
  const int forced = env_int("SVBRDF_B200_CHUNK", 0);
  if (forced == 3 || forced == 4) return forced;
  if (n_lights < 4) return 3;
  return (n_lights % 4) <= (n_lights % 3) ? 4 : 3;
}

template <int MODE, bool WANT_POW, int TGT>
static int launch_tile(Params P, cudaStream_t st) {
  if (env_int("SVBRDF_B200_FORCE_LDG", 0) || !tma_ok<TGT>(P)) return launch_texel<MODE, WANT_POW, TGT>(P, st);
  const bool out_ok = MODE == kModeL2Adam || (reinterpret_cast<uintptr_t>(P.out) & 15) == 0;
  // opt-in (SVBRDF_B200_TSTORE=1): measured 72.3 vs 71.2 us per epoch at 1024^2 x 9 and 1320 vs 1232 us at 2048^2 x 64 against the
  // per-thread-store kernel below, although it executes 8 % fewer instructions
#if SV_ENABLE_TS
  if constexpr (TGT != SVBRDF_TARGET_F16)                   // the store-back kernel's 2-D tensor maps cover f32 and u8 targets
  if (P.push_world == 0 && P.tile_rotate == 0 && out_ok && !env_int("SVBRDF_B200_PACKED", SV_PACKED_DEFAULT) && env_int("SVBRDF_B200_TSTORE", 0))
    return launch_tile_ts<MODE, WANT_POW, TGT, ScalarShape>(P, st);
#endif
#if SV_ENABLE_PACKED
  // the peer-sharded modes own texels in units of ScalarShape::kTile (check_peers): the packed shape's 448-texel tiles would
  // straddle owners, so the development override does not apply to them
  if (P.push_world == 0 && env_int("SVBRDF_B200_PACKED", SV_PACKED_DEFAULT)) return launch_tile_shape<MODE, WANT_POW, TGT, PackedShape>(P, st);
#endif
  static_assert(ScalarShape4::kTile == ScalarShape::kTile, "peer ownership granularity is one tile of either scalar shape");
  // self-fed ring (16 consumer warps, no producer warp): not for the peer-sharded modes (480-texel ownership granularity)
#if SV_ENABLE_SELF
  if (P.push_world == 0 && P.tile_rotate == 0 && env_int("SVBRDF_B200_SELFFED", SV_SELFFED_DEFAULT)) {
    if (pick_chunk(P.n_lights) == 4) return launch_tile_self<MODE, WANT_POW, TGT, SelfShape4, ScalarShape4>(P, st);
    return launch_tile_self<MODE, WANT_POW, TGT, SelfShape, ScalarShape>(P, st);
  }
#endif
  if (pick_chunk(P.n_lights) == 4) return launch_tile_shape<MODE, WANT_POW, TGT, ScalarShape4>(P, st);
  return launch_tile_shape<MODE, WANT_POW, TGT, ScalarShape>(P, st);
}

template <int MODE>
static int launch_l2(const Params& P, bool want_pow, int tgt, cudaStream_t st) {
  if (tgt == SVBRDF_TARGET_F32)
    return want_pow ? launch_tile<MODE, true, SVBRDF_TARGET_F32>(P, st) : launch_tile<MODE, false, SVBRDF_TARGET_F32>(P, st);
  if (tgt == SVBRDF_TARGET_U8)
    return want_pow ? launch_tile<MODE, true, SVBRDF_TARGET_U8>(P, st) : launch_tile<MODE, false, SVBRDF_TARGET_U8>(P, st);
  if (tgt == SVBRDF_TARGET_F16)
    return want_pow ? launch_tile<MODE, true, SVBRDF_TARGET_F16>(P, st) : launch_tile<MODE, false, SVBRDF_TARGET_F16>(P, st);
  return SVBRDF_E_UNSUPPORTED;
}

// The consumer backward with targets on the TMA ring (kModeVjpL2): needs what every tile launch needs (16-byte aligned
// plane segments) for BOTH image-sized inputs.
template <bool WANT_POW, int TGT>
static int launch_vjp_l2_tile(const Params& P, cudaStream_t st) {
  if (pick_chunk(P.n_lights) == 4) return launch_tile_shape<kModeVjpL2, WANT_POW, TGT, ScalarShape4>(P, st);
  return launch_tile_shape<kModeVjpL2, WANT_POW, TGT, ScalarShape>(P, st);
}

template <bool BWD>
static int launch_norm_l2(Params& P, bool want_pow, int tgt, cudaStream_t st) {
  const float no_mean[3] = {0.f, 0.f, 0.f};
  if (BWD && P.io2 && P.l2_up && !env_int("SVBRDF_B200_FORCE_LDG", 0) && (reinterpret_cast<uintptr_t>(P.io2) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(P.out) & 15) == 0 && fixed_div_ok(no_mean, P.aff_std)) {
    // one thread per texel with direct LDG ran at 24 % occupancy with 11.6 warps per issued instruction waiting on the
    // long scoreboard (260 us at 1024^2 x 9, profiles/r02_modeb_kernels_1024x9_summary.txt): both streams go through the ring
    int r = kNoTilePath;
    if (tgt == SVBRDF_TARGET_U8 && tma_ok<SVBRDF_TARGET_U8>(P))
      r = want_pow ? launch_vjp_l2_tile<true, SVBRDF_TARGET_U8>(P, st) : launch_vjp_l2_tile<false, SVBRDF_TARGET_U8>(P, st);
    else if (tgt == SVBRDF_TARGET_F32 && tma_ok<SVBRDF_TARGET_F32>(P))
      r = want_pow ? launch_vjp_l2_tile<true, SVBRDF_TARGET_F32>(P, st) : launch_vjp_l2_tile<false, SVBRDF_TARGET_F32>(P, st);
    if (r != kNoTilePath) return r;
  }
  if (!BWD && P.io2 && !env_int("SVBRDF_B200_FORCE_LDG", 0) && fixed_div_ok(P.aff_mean, P.aff_std)) {
    // forward with targets: the targets stream through the ring (kModeNormFwd); one thread per texel with direct LDG and no
    // prefetch ran at 76.7 us at 1024^2 x 9 (53 % of the HBM roof, profiles/r02_modeb_kernels_1024x9_summary_ring.txt)
    Params Q = P;
    Q.io = P.io2;
    Q.epochs = 1;
    int r = kNoTilePath;
    // (23 + 1 warps at 80 registers — the forward-only body fits — were measured too: 57.6 vs 58.5 us, issue slots 75 vs 71 %;
    // not worth a second shape)
    const bool four = pick_chunk(Q.n_lights) == 4;
    if (tgt == SVBRDF_TARGET_U8 && tma_ok<SVBRDF_TARGET_U8>(Q))
      r = four ? launch_tile_shape<kModeNormFwd, false, SVBRDF_TARGET_U8, ScalarShape4>(Q, st)
               : launch_tile_shape<kModeNormFwd, false, SVBRDF_TARGET_U8, ScalarShape>(Q, st);
    else if (tgt == SVBRDF_TARGET_F32 && tma_ok<SVBRDF_TARGET_F32>(Q))
      r = four ? launch_tile_shape<kModeNormFwd, false, SVBRDF_TARGET_F32, ScalarShape4>(Q, st)
               : launch_tile_shape<kModeNormFwd, false, SVBRDF_TARGET_F32, ScalarShape>(Q, st);
    if (r != kNoTilePath) return r;
  }
  const size_t smem = size_t(P.n_lights) * 2 * sizeof(float4);
  const int blocks = texel_blocks(P);
#define SV_LAUNCH_NL2(WP, TG)                                                                                       \
  do {                                                                                                              \
    auto kern = norm_l2_kernel<BWD, WP, TG>;                                                                        \
    if (smem > 48 * 1024)                                                                                           \
      if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return int(e); \
    kern<<<blocks, kThreads, smem, st>>>(P);                                                                        \
  } while (0)
  if (tgt == SVBRDF_TARGET_U8) {
    if (want_pow) SV_LAUNCH_NL2(true, SVBRDF_TARGET_U8); else SV_LAUNCH_NL2(false, SVBRDF_TARGET_U8);
  } else if (tgt == SVBRDF_TARGET_F32) {
    if (want_pow) SV_LAUNCH_NL2(true, SVBRDF_TARGET_F32); else SV_LAUNCH_NL2(false, SVBRDF_TARGET_F32);
  } else {
    return SVBRDF_E_UNSUPPORTED;
  }
#undef SV_LAUNCH_NL2
  if (cudaError_t e = cudaGetLastError()) return int(e);
  if (P.loss_out || P.grad_pow) {
    finalize_kernel<<<1, 256, 0, st>>>(P, blocks);
    return int(cudaGetLastError());
  }
  return 0;
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
  return partial_rows((long long)res * rows) * 4 * sizeof(float) + kMaxEpochs * sizeof(unsigned) + 16;
}

int svbrdf_render_fwd(const svbrdf_geom_t* geom, const float* tex, float* out, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !out) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.out = out;
  return launch_texel<kModeRender, false, SVBRDF_TARGET_F32>(P, stream);
}

int svbrdf_render_bwd(const svbrdf_geom_t* geom, const float* tex, const float* grad_out, float* grad_tex, float* grad_pow,
                      void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !grad_out || !grad_tex || !workspace) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = grad_out;
  P.out = grad_tex;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(1.0 / kGamma);
  P.grad_pow = grad_pow;
  return grad_pow ? launch_tile<kModeVjp, true, SVBRDF_TARGET_F32>(P, stream) : launch_tile<kModeVjp, false, SVBRDF_TARGET_F32>(P, stream);
}

int svbrdf_l2_grad(const svbrdf_geom_t* geom, const float* tex, const void* target, int32_t target_dtype, int32_t n_total,
                   float* grad_tex, float* loss_out, float* grad_pow, void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !target || !grad_tex || !workspace || n_total < geom->n_lights) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = target;
  P.out = grad_tex;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(l2_scale(geom, n_total));
  P.loss_norm = 1.0 / (double(n_total) * 3.0 * double(geom->res) * double(geom->res));
  P.loss_out = loss_out;
  P.grad_pow = grad_pow;
  return launch_l2<kModeL2Grad>(P, grad_pow != nullptr, target_dtype, stream);
}

int svbrdf_l2_adam_run(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target, int32_t target_dtype,
                       const svbrdf_adam_t* adam, int32_t epochs, float* loss_curve, float* pow_state, void* workspace,
                       svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !m || !v || !target || !adam || !workspace || epochs < 0 || adam->step < 1) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = tex;
  P.m = m;
  P.v = v;
  P.io = target;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(l2_scale(geom, geom->n_lights));
  P.loss_norm = 1.0 / (double(geom->n_lights) * 3.0 * double(geom->res) * double(geom->res));
  P.pow_state = pow_state;
  // optim_light couples all texels through light_pow every epoch -> one launch per epoch; otherwise up to
  // kMaxEpochs epochs run inside one persistent launch.
  // Long epochs (>= ~1 ms: launch, ramp-up and tail are already < 1 %) gain nothing from persistence and pay ~2 % for
  // the progress fences (measured at 4096^2 x 64), so they keep one launch per epoch.
  const bool long_epoch = double(P.texels) * double(geom->n_lights) >= 2.0e8;
  int per_launch = env_int("SVBRDF_B200_EPOCHS_PER_LAUNCH", long_epoch ? 1 : kMaxEpochs);
  if (pow_state || per_launch < 1) per_launch = 1;
  if (per_launch > kMaxEpochs) per_launch = kMaxEpochs;
  for (int e0 = 0; e0 < epochs; e0 += per_launch) {
    const int n = epochs - e0 < per_launch ? epochs - e0 : per_launch;
    for (int i = 0; i < n; ++i) {
      const AdamStep<float> a = make_adam(*adam, adam->step + e0 + i);
      if (i == 0) P.adam = a;
      P.step_size[i] = a.step_size;
      P.inv_sqrt_bc2[i] = a.inv_sqrt_bc2;
    }
    P.epochs = n;
    P.loss_out = loss_curve ? loss_curve + e0 : nullptr;
    if (int err = launch_l2<kModeL2Adam>(P, pow_state != nullptr, target_dtype, stream)) return err;
  }
  return 0;
}

int svbrdf_l2_adam_step(const svbrdf_geom_t* geom, float* tex, float* m, float* v, const void* target, int32_t target_dtype,
                        const svbrdf_adam_t* adam, float* loss_out, float* pow_state, void* workspace,
                        svbrdf_stream_t stream) {
  return svbrdf_l2_adam_run(geom, tex, m, v, target, target_dtype, adam, 1, loss_out, pow_state, workspace, stream);
}

static int check_peers(const svbrdf_peers_t* p) {
  if (!p || p->world < 1 || p->world > 8 || p->rank < 0 || p->rank >= p->world || p->chunk <= 0 || p->chunk % ScalarShape::kTile != 0) return SVBRDF_E_BADARG;
  for (int r = 0; r < p->world; ++r)
    if (!p->recv[r] || !p->tex[r]) return SVBRDF_E_BADARG;
  return 0;
}

int svbrdf_l2_grad_push(const svbrdf_geom_t* geom, const float* tex, const void* target, int32_t target_dtype, int32_t n_total,
                        const svbrdf_peers_t* peers, float* loss_out, void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (int e = check_peers(peers)) return e;
  // a row band (rows < res, row_offset > 0) is legal: the band's texels are what the peers share (2-D decomposition: row
  // bands x light shards, one peer group per band); ownership, receive slots and `tex` are all relative to the band
  if (!tex || !target || !workspace || n_total < geom->n_lights || geom->plane_stride != 0) return SVBRDF_E_BADARG;
  if ((long long)peers->world * peers->chunk < (long long)geom->rows * geom->res) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = target;
  P.out = nullptr;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(l2_scale(geom, n_total));
  P.loss_norm = 1.0 / (double(n_total) * 3.0 * double(geom->res) * double(geom->res));
  P.loss_out = loss_out;
  P.push_world = peers->world;
  P.push_rank = peers->rank;
  P.push_chunk = peers->chunk;
  for (int r = 0; r < peers->world; ++r) {
    P.push_recv[r] = peers->recv[r];
    P.push_tex[r] = peers->tex[r];
  }
  P.push_pull = peers->pull_tex ? 1 : 0;
  {
    const long long n_tiles = (P.texels + ScalarShape::kTile - 1) / ScalarShape::kTile;
    P.tile_rotate = ((long long)peers->rank * (peers->chunk / ScalarShape::kTile)) % n_tiles;
  }
  return launch_l2<kModeL2Grad>(P, false, target_dtype, stream);
}

int svbrdf_reduce_adam_push(const svbrdf_peers_t* peers, int64_t texels, float* m, float* v, const svbrdf_adam_t* adam,
                            svbrdf_stream_t stream) {
  if (int e = check_peers(peers)) return e;
  if (!m || !v || !adam || adam->step < 1 || texels <= 0 || texels % 4 != 0) return SVBRDF_E_BADARG;
  DeviceGuard guard(m);
  PushAdamParams Q{};
  Q.world = peers->world;
  Q.rank = peers->rank;
  Q.chunk = peers->chunk;
  Q.texels = texels;
  Q.recv = peers->recv[peers->rank];
  for (int r = 0; r < peers->world; ++r) Q.tex[r] = peers->tex[r];
  Q.tex_mc = peers->tex_multicast;
  Q.local_only = peers->pull_tex ? 1 : 0;
  Q.m = m;
  Q.v = v;
  Q.adam = make_adam(*adam, adam->step);
  const long long work = peers->chunk / 4 * 9;
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reduce_adam_push_kernel<<<int(blocks), 256, 0, stream>>>(Q);
  return int(cudaGetLastError());
}

int svbrdf_render_norm_l2_fwd(const svbrdf_geom_t* geom, const float* tex, const float* mean, const float* std_, const void* target,
                              int32_t target_dtype, float* out, float* loss_out, void* workspace, svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !out || !mean || !std_ || !workspace) return SVBRDF_E_BADARG;
  if (loss_out && !target) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.out = out;
  P.io2 = target;
  for (int c = 0; c < 3; ++c) { P.aff_mean[c] = mean[c]; P.aff_std[c] = std_[c]; }
  P.partials = static_cast<float*>(workspace);
  P.loss_norm = 1.0 / (double(geom->n_lights) * 3.0 * double(geom->res) * double(geom->res));
  P.loss_out = loss_out;
  return launch_norm_l2<false>(P, false, target ? target_dtype : SVBRDF_TARGET_F32, stream);
}

int svbrdf_render_norm_l2_bwd(const svbrdf_geom_t* geom, const float* tex, const float* std_, const float* grad_out, const void* target,
                              int32_t target_dtype, const float* l2_grad, float* grad_tex, float* grad_pow, void* workspace,
                              svbrdf_stream_t stream) {
  if (int e = check_geom(geom)) return e;
  if (!tex || !grad_out || !grad_tex || !std_ || !workspace) return SVBRDF_E_BADARG;
  DeviceGuard guard(tex);
  Params P = base_params(geom);
  P.tex = const_cast<float*>(tex);
  P.io = grad_out;
  P.io2 = (target && l2_grad) ? target : nullptr;
  P.l2_up = l2_grad;
  for (int c = 0; c < 3; ++c) { P.aff_mean[c] = 0.f; P.aff_std[c] = std_[c]; }
  P.out = grad_tex;
  P.partials = static_cast<float*>(workspace);
  P.scale = float(1.0 / kGamma);
  P.loss_norm = 1.0 / (double(geom->n_lights) * 3.0 * double(geom->res) * double(geom->res));
  P.grad_pow = grad_pow;
  return launch_norm_l2<true>(P, grad_pow != nullptr, P.io2 ? target_dtype : SVBRDF_TARGET_F32, stream);
}

int svbrdf_adam_apply(float* param, float* m, float* v, const float* grad, size_t count, const svbrdf_adam_t* adam,
                      svbrdf_stream_t stream) {
  if (!param || !m || !v || !grad || !adam || adam->step < 1) return SVBRDF_E_BADARG;
  if (count == 0) return 0;
  DeviceGuard guard(param);
  const AdamStep<float> a = make_adam(*adam, adam->step);
  const size_t want = (count / 4 + 255) / 256 + 1;
  const int blocks = int(want < size_t(148 * 16) ? want : size_t(148 * 16));
  adam_apply_kernel<<<blocks, 256, 0, stream>>>(param, m, v, grad, count, a);
  return int(cudaGetLastError());
}

}  // extern "C"
