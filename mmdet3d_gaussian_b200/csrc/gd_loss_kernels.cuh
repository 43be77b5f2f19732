// Fused forward+backward Gaussian-distance loss kernels for B200 (sm_100a).
//
// Replaces, in ONE pass over HBM (88 algorithmic bytes per box pair: pred 28 +
// target 28 + weight 4 read, grad 28 written), what the reference does with
// ~160-250 eager torch ops and autograd: GDLoss.forward
// (mmdet3d_gaussian/models/losses/gaussian_distance_loss.py:280-310, "ref")
// = preprocess x2 (ref:8-21) -> distance (ref:42-248) -> postprocess (ref:24-39)
// -> weighted reduction (mmdet weight_reduce_loss) -> x loss_weight, plus
// d loss / d pred.
//
// Two kernels, same per-row math (gd_math.cuh):
//   * gd_warp_kernel   -- persistent, one CTA per SM.  Every WARP runs its own
//                         2-stage ring of [32*R-row] tiles (R rows per lane):
//                         lane 0 fills a stage with 1-D bulk async copies (TMA
//                         engine, cp.async.bulk + mbarrier complete_tx), lanes read
//                         their rows from shared memory at stride 7 (odd => no bank
//                         conflicts), stage the gradient rows in shared memory and
//                         lane 0 writes them back with a bulk store.  No CTA-wide
//                         barrier in the loop; per-tile bookkeeping is amortised
//                         over R rows.  fun / tau / flag / weight mode are
//                         compile-time for the common configurations.  Needs
//                         contiguous, 16-byte aligned tensors.
//   * gd_staged_kernel -- CTA tiles with plain (vector when possible) loads and
//                         stores; takes any row stride / alignment (the
//                         CenterGDHead call site passes row-strided views).
// The AoS [N,7] layout is the reference's contract; the transpose to
// one-row-per-thread happens in shared memory, never in HBM.
//
// Loss sum: per-thread fp32 -> warp shuffle -> per-CTA fp64 partial -> the last
// CTA (atomic ticket, one atomic per CTA) adds the partials in fixed order, so
// the result is deterministic for a given grid.
#pragma once
#include "gd_common.cuh"
#include "gd_packed.cuh"

// Tuning knobs of gd_warp_kernel.  The production library bakes GD_TUNE_DEFAULT in
// at compile time (dead branches vanish); `-DGD_TUNE=1` builds
// libgdloss_b200_tune.so, where the same knobs are read per launch from the
// environment (GD_TUNE_FLAGS, GD_TUNE_WARPS) so one GPU session can measure all of
// them (tools/tune_sweep.py).
#ifndef GD_TUNE
#define GD_TUNE 0
#endif
#ifndef GD_TUNE_DEFAULT
#define GD_TUNE_DEFAULT 0
#endif

namespace gdk {

enum : int {
  kTuneLateWait = 1,       // wait for the previous tile's bulk store after the math, not before
  kTuneLoadNormal = 2,     // loads with L2 evict_normal instead of evict_first
  kTuneStoreHint = 4,      // bulk stores carry an L2 evict_first hint
  kTuneNoMath = 8,         // (measurement only) replace the math by p + t*w: pipeline ceiling
  kTuneGenericStore = 16,  // all lanes write the gradient tile with st.global.cs.v4
  kTuneHalfMath = 32,      // (measurement only) math on half of each lane's rows
  kTuneLane0 = 64,         // copy commands issued by `lane == 0` instead of elect.sync
  kTuneStrided = 128,      // lane owns rows lane + 32 k (32-bit shared-memory accesses)
  // Compile-time only (GD_TUNE_DEFAULT; ignored in GD_TUNE_FLAGS): instruction diets of the
  // FAST math (gd_math.cuh, gd::Diet) and the packed math as the default of 'bulk'.
  kTuneGuardMinMax = 256,  // nice-row screen with FMNMX3.NAN instead of 16 compares
  kTuneStd = 512,          // specialised kernels assume alpha == 1, center_offset == (0,0,.5);
                           // any other value takes the run-time-parameter kernel
  kTunePacked = 1024,      // GD_VARIANT_AUTO runs the packed-FP32 math where it exists
};
constexpr int kDietGuards = (GD_TUNE_DEFAULT & kTuneGuardMinMax) ? gd::kDietGuards : 0;
struct FullTile { static constexpr bool value = true; };
struct PartTile { static constexpr bool value = false; };

struct LossArgs {
  const float* pred;
  const float* target;
  const float* weight;
  long long pstride, tstride, wstride;   // row strides in elements
  long long n;
  int wmode;
  int mask_zero_w;                        // GD_FLAG_MASK_ZERO_WEIGHT
  float scale;
  float* loss_sum;
  float* row_loss;
  float* grad;
  double* partials;                       // [grid]
  unsigned int* ticket;                   // zero on entry, zero again on exit; ticket[1] = any-positive word
  gd::PairParams<float> pp;
  int tune = GD_TUNE_DEFAULT;             // only read by -DGD_TUNE=1 builds
  // effective scale = scale / *scale_div when scale_div is given (avg_factor living on the
  // device: gd_centerpoint_head.py:407 without its .item())
  const float* scale_div = nullptr;
  // nullable, 1 fp32 := any(weight > 0) ? 1 : 0 over every weight ELEMENT (ref:290), written by
  // the last CTA together with loss_sum
  float* status = nullptr;
  // early return of GDLoss.forward decided on the device (ref:290-292): when no weight
  // element is > 0 the last CTA replaces the outputs by (pred * weight).sum() and its gradient
  // `weight`, with weight(i, c) = weight[i * er_wrow + c * er_wcol]
  int early_return = 0;
  long long er_wrow = 0, er_wcol = 0;
  // nullable, PINNED host memory (device-addressable under UVA), zeroed by the caller before the
  // launch: any(weight > 0) reported to a waiting host thread from inside THIS launch -- 1 as
  // soon as the first tile of the first warp holds a positive weight (a few microseconds after
  // the kernel starts), else 1 / 2 (some / none positive) from the last CTA.  Replaces the
  // separate probe launch of the host-visible early-return check (ref:290-292).
  int* host_flag = nullptr;
  // gd_warp_kernel<..., ANY = true> (row-strided and/or 16-byte-unaligned inputs): rows
  // [row_lo, row_lo + n_bulk) go through the bulk-copy tiles, the <= 8 rows around them are
  // read straight from global memory; *shift = words between the 16-byte aligned copy
  // window and the first element of a tile (constant: tiles are multiples of 4 rows)
  long long row_lo = 0, n_bulk = 0;
  int pshift = 0, tshift = 0, wshift = 0;
  // cross-GPU sum of loss_sum over peer memory (world <= 1: none); include/gd_loss_b200.h
  gd_peer_sum peer = {};
};

// Exchange buffer of one rank (gd_peer_sum_buffer_bytes()).  Two value / flag sets, used
// alternately by consecutive calls: a rank can be at most one call ahead of its slowest peer
// (it cannot finish call s + 1 without that peer's flag for s + 1), so the set of call s is
// never overwritten before everybody has read it.
struct PeerBuf {
  unsigned int seq;                          // calls completed by the owning rank
  unsigned int pad[3];
  double val[2][GD_MAX_PEERS];
  unsigned int flag[2][GD_MAX_PEERS];
};

// Called by every thread of the LAST CTA of a launch (uniformly).  Returns the global sum.
__device__ __forceinline__ double peer_exchange_sum(const gd_peer_sum& ps, double mine) {
  __shared__ double s_total;
  const int tid = threadIdx.x;
  if (tid == 0) s_total = mine;              // thread 0 holds the partial; every talker sends it
  __syncthreads();
  mine = s_total;
  __syncthreads();
  PeerBuf* own = reinterpret_cast<PeerBuf*>(ps.peer_buf[ps.rank]);
  const unsigned int seq = own->seq + 1u;    // written only by this rank's previous launch
  const int par = (int)(seq & 1u);
  if (tid < ps.world) {                      // thread p talks to peer p
    PeerBuf* dst = reinterpret_cast<PeerBuf*>(ps.peer_buf[tid]);
    st_relaxed_sys(&dst->val[par][ps.rank], mine);
    st_release_sys(&dst->flag[par][ps.rank], seq);      // orders the value before the flag
    while (ld_acquire_sys(&own->flag[par][tid]) != seq) {
    }
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int r = 0; r < ps.world; ++r) tot += ld_relaxed_sys(&own->val[par][r]);   // rank order
    own->seq = seq;
    s_total = tot;
  }
  __syncthreads();
  return s_total;
}

__device__ __forceinline__ float effective_scale(const LossArgs& a) {
  return a.scale_div ? a.scale / __ldg(a.scale_div) : a.scale;
}

// ---------------------------------------------------------------------------
// deterministic grid-wide sum
// ---------------------------------------------------------------------------
constexpr int kMaxWarps = 32;

__device__ __forceinline__ void finish_sum(float acc, const LossArgs& a, float scale,
                                           bool any_pos = false) {
  __shared__ float s_warp[kMaxWarps];
  __shared__ double s_dwarp[kMaxWarps];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = (int)(blockDim.x >> 5), nthreads = (int)blockDim.x;
  acc = warp_sum(acc);
  if (lane == 0) s_warp[warp] = acc;
  const int cta_any = __syncthreads_or(any_pos ? 1 : 0);     // also orders s_warp
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += (double)s_warp[w];
    a.partials[blockIdx.x] = s;
    if ((a.status || a.early_return || a.host_flag) && cta_any) atomicOr(a.ticket + 1, 1u);
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double s = 0.0;
    for (unsigned int i = tid; i < gridDim.x; i += nthreads) s += __ldcg(a.partials + i);
    s = warp_sum(s);
    if (lane == 0) s_dwarp[warp] = s;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0) {
      for (int w = 0; w < nwarps; ++w) tot += s_dwarp[w];
      tot *= (double)scale;
    }
    if (a.peer.world > 1) tot = peer_exchange_sum(a.peer, tot);    // thread 0's partial is sent
    const bool probe = a.status != nullptr || a.early_return != 0 || a.host_flag != nullptr;
    // ONE thread reads the any-positive word (it is reset below) and the vote makes the
    // decision CTA-uniform
    const bool none_positive =
        probe && __syncthreads_or(tid == 0 && __ldcg(a.ticket + 1) == 0u ? 1 : 0) != 0;
    if (a.early_return && none_positive) {
      // ref:292, the rare branch (in practice an EMPTY batch): every other CTA has finished
      // (ticket), so this one rewrites the gradient and sums pred * weight on its own
      __syncthreads();
      double acc = 0.0;
      const long long nel = a.n * 7;
      for (long long i = tid; i < nel; i += nthreads) {
        const long long r = i / 7;
        const int c = (int)(i - r * 7);
        const float w = a.weight[r * a.er_wrow + c * a.er_wcol];
        acc += (double)(a.pred[r * a.pstride + c] * w);
        if (a.grad) a.grad[i] = w;
      }
      acc = warp_sum(acc);
      if (lane == 0) s_dwarp[warp] = acc;
      __syncthreads();
      if (tid == 0) {
        tot = 0.0;
        for (int w = 0; w < nwarps; ++w) tot += s_dwarp[w];
      }
    }
    if (tid == 0) {
      *a.loss_sum = (float)tot;
      if (a.status) *a.status = none_positive ? 0.0f : 1.0f;
      if (a.host_flag) {
        *reinterpret_cast<volatile int*>(a.host_flag) = none_positive ? 2 : 1;
        __threadfence_system();
      }
      if (probe) a.ticket[1] = 0u;
      *a.ticket = 0u;                     // leave the workspace reusable
    }
  }
}

// weight of one row: None -> 1, [N] -> w, [N,7] -> mean(-1)          ref:295-296
__device__ __forceinline__ float row_weight_smem(const float* sw, int wmode, int r) {
  if (wmode == GD_WEIGHT_ROW) return sw[r];
  if (wmode == GD_WEIGHT_ROW7) {
    const float* w = sw + 7 * r;
    return (((((w[0] + w[1]) + w[2]) + w[3]) + w[4]) + w[5] + w[6]) / 7.0f;
  }
  return 1.0f;
}

// One row, registers only.  Returns w_i * loss_i (unscaled) for the sum;
// g[] <- scale * w_i * dloss_i/dpred_i, *rl <- scale * w_i * loss_i.
template <int LOSS, bool GRAD>
__device__ __forceinline__ float eval_row(const float* p, const float* t, float w,
                                          const gd::PairParams<float>& pp, float scale,
                                          bool mask_zero_w, float* g, float* rl) {
  const float ws = w * scale;
  float l = gd::pair_eval<float, LOSS, GRAD>(p, t, pp, ws, g);
  if (mask_zero_w && w == 0.0f) {          // masked row: exactly zero, nan/inf do not leak
    l = 0.0f;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) g[c] = 0.0f;
    }
  }
  *rl = l * ws;
  return l * w;
}

// ---------------------------------------------------------------------------
// staged kernel: any stride / alignment
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_rows(float* __restrict__ s, const float* __restrict__ g,
                                          long long stride, int cols, long long row0, int rows,
                                          int tid) {
  const int nel = rows * cols;
  if (stride == cols) {
    const float* base = g + row0 * cols;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
      const int nv = nel >> 2;
      const float4* b4 = reinterpret_cast<const float4*>(base);
      float4* s4 = reinterpret_cast<float4*>(s);
      for (int i = tid; i < nv; i += kThreads) s4[i] = __ldcs(b4 + i);
      for (int i = (nv << 2) + tid; i < nel; i += kThreads) s[i] = __ldcs(base + i);
    } else {
      for (int i = tid; i < nel; i += kThreads) s[i] = __ldcs(base + i);
    }
  } else {
    for (int i = tid; i < nel; i += kThreads) {
      const int r = i / cols, c = i - r * cols;
      s[i] = __ldcs(g + (row0 + r) * stride + c);
    }
  }
}

__device__ __forceinline__ void store_rows(float* __restrict__ g, const float* __restrict__ s,
                                           int cols, long long row0, int rows, int tid) {
  const int nel = rows * cols;
  float* base = g + row0 * cols;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
    const int nv = nel >> 2;
    float4* b4 = reinterpret_cast<float4*>(base);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = tid; i < nv; i += kThreads) __stcs(b4 + i, s4[i]);
    for (int i = (nv << 2) + tid; i < nel; i += kThreads) __stcs(base + i, s[i]);
  } else {
    for (int i = tid; i < nel; i += kThreads) __stcs(base + i, s[i]);
  }
}

template <int LOSS, bool GRAD>
__global__ void __launch_bounds__(kThreads) gd_staged_kernel(const LossArgs a) {
  __shared__ __align__(16) float s_pred[kTile * 7];    // reused for the gradient tile
  __shared__ __align__(16) float s_tgt[kTile * 7];
  __shared__ __align__(16) float s_w[kTile * 7];       // [N,7] weights only
  const int tid = threadIdx.x;
  const long long ntiles = (a.n + kTile - 1) / kTile;
  const float scale = effective_scale(a);
  const bool probe = a.status != nullptr || a.early_return != 0 ||
                     a.host_flag != nullptr;   // any(weight > 0) over every weight element, ref:290
  bool anyp = false;
  float acc = 0.0f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kTile;
    const int rows = (int)min((long long)kTile, a.n - row0);
    load_rows(s_pred, a.pred, a.pstride, 7, row0, rows, tid);
    load_rows(s_tgt, a.target, a.tstride, 7, row0, rows, tid);
    if (a.wmode == GD_WEIGHT_ROW7) load_rows(s_w, a.weight, a.wstride, 7, row0, rows, tid);
    float w1 = 1.0f;
    if (a.wmode == GD_WEIGHT_ROW && tid < rows) w1 = __ldcs(a.weight + (row0 + tid) * a.wstride);
    __syncthreads();
    if (tid < rows) {
      float p[7], t[7], g[7], rl;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        p[c] = s_pred[7 * tid + c];
        t[c] = s_tgt[7 * tid + c];
      }
      const float w = (a.wmode == GD_WEIGHT_ROW7) ? row_weight_smem(s_w, a.wmode, tid) : w1;
      if (probe) {
        if (a.wmode == GD_WEIGHT_ROW7) {
#pragma unroll
          for (int c = 0; c < 7; ++c) anyp |= s_w[7 * tid + c] > 0.0f;
        } else if (a.wmode == GD_WEIGHT_ROW) {
          anyp |= w > 0.0f;
        }
      }
      acc += eval_row<LOSS, GRAD>(p, t, w, a.pp, scale, a.mask_zero_w != 0, g, &rl);
      if (a.row_loss) __stcs(a.row_loss + row0 + tid, rl);
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < 7; ++c) s_pred[7 * tid + c] = g[c];   // own row only: no hazard
      }
    }
    __syncthreads();
    if (GRAD) store_rows(a.grad, s_pred, 7, row0, rows, tid);
    __syncthreads();
  }
  if (a.loss_sum) finish_sum(acc, a, scale, anyp);
}

// ---------------------------------------------------------------------------
// warp pipeline kernel: persistent, every warp owns a 2-stage ring of tiles
// ---------------------------------------------------------------------------
constexpr int kWarpStages = 2;

struct WarpLayout {
  int ptile;      // bytes of one pred tile
  int ttile;      // bytes of one target tile
  int wtile;      // bytes of one weight tile
  int stage;      // bytes of one input stage (pred | target | weight)
  int out;        // bytes of the output buffer (grad | row loss)
  long long per_warp;   // shared memory bytes one warp needs
};

// Contiguous 16-byte aligned inputs (ANY = false): a tile of an array is `trows` packed
// rows.  ANY = true: a tile is the 16-byte aligned window around `trows` rows of `stride`
// floats each, up to 12 bytes of lead-in and 12 of round-up included.
__host__ __device__ inline int any_tile_bytes(int trows, long long stride) {
  return (int)(((long long)trows * stride * 4 + 16 + 15) & ~15LL);
}
__host__ __device__ inline WarpLayout warp_layout(int rows_per_lane, int wmode, bool grad,
                                                  bool rows, bool any = false, long long ps = 7,
                                                  long long ts = 7, long long wsd = 0) {
  const int trows = 32 * rows_per_lane;
  WarpLayout L;
  if (any) {
    L.ptile = any_tile_bytes(trows, ps);
    L.ttile = any_tile_bytes(trows, ts);
    L.wtile = wmode != GD_WEIGHT_NONE ? any_tile_bytes(trows, wsd) : 0;
  } else {
    L.ptile = L.ttile = trows * kRowBytes;
    L.wtile = wmode == GD_WEIGHT_ROW7 ? trows * kRowBytes : (wmode == GD_WEIGHT_ROW ? trows * 4 : 0);
  }
  L.stage = L.ptile + L.ttile + L.wtile;
  L.out = (grad ? trows * kRowBytes : 0) + (rows ? trows * 4 : 0);
  L.per_warp = (long long)kWarpStages * L.stage + L.out + 16;      // + two mbarriers
  return L;
}

// SPEC < 0: fun / tau_on / flag / mask come from the arguments at run time.
// SPEC >= 0: bits [1:0] fun, [2] tau_on, [3] flag are compile-time constants (the
// per-row branches and constant loads disappear).
// WM < 0: weight mode at run time, else compile-time.
// N consecutive floats shared <-> registers with the widest aligned vector access
// (N = 7 R per lane: R = 4 -> 7 x 128-bit, lane stride 112 B; R = 2 -> 7 x 64-bit, lane
// stride 56 B; both strides are bank-conflict free within a quarter / half warp).
template <int N>
__device__ __forceinline__ void smem_load(float* dst, const float* src) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
      const float4 v = reinterpret_cast<const float4*>(src)[j];
      dst[4 * j] = v.x; dst[4 * j + 1] = v.y; dst[4 * j + 2] = v.z; dst[4 * j + 3] = v.w;
    }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const float2 v = reinterpret_cast<const float2*>(src)[j];
      dst[2 * j] = v.x; dst[2 * j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = src[j];
  }
}
template <int N>
__device__ __forceinline__ void smem_store(float* dst, const float* src) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int j = 0; j < N / 4; ++j)
      reinterpret_cast<float4*>(dst)[j] =
          make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j)
      reinterpret_cast<float2*>(dst)[j] = make_float2(src[2 * j], src[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = src[j];
  }
}

// PACK: the FAST math of two rows of a lane runs as one packed-FP32 (f32x2) stream
// (gd_packed.cuh); needs an even R and one of the three headline losses.
// ANY: inputs with any row stride (the CenterGDHead call site passes `[..., :7]` views of
// 9- / 11-wide rows, gd_centerpoint_head.py:413-423) and any 4-byte alignment.  4 rows of
// `stride` floats are 16 * stride bytes, so every tile of whole WIDE rows is still one
// 1-D bulk copy once its start is rounded down to 16 bytes; the lanes pick columns 0..6 out
// of shared memory (row lane + 32 k: bank = lane * stride mod 32, conflict free for odd
// strides).  The gradient is contiguous [n,7] as always.
template <int LOSS, bool GRAD, int R, int SPEC, int WM, bool PACK = false, bool ANY = false>
__global__ void __launch_bounds__(R >= 4 ? 384 : 768, 1) gd_warp_kernel(const LossArgs a) {
#if defined(GD_HOST_EMULATION)
  unsigned char* smem = emu_dynamic_smem();   // tests/host_math/loss_emul.cpp
#else
  extern __shared__ __align__(128) unsigned char smem[];
#endif
  constexpr int kTileRows = 32 * R;
  constexpr int kDiet = kDietGuards | ((SPEC >= 0 && (GD_TUNE_DEFAULT & kTuneStd)) ? gd::kDietStd : 0);
  const int tid = threadIdx.x, lane = tid & 31;
  // The warp index goes through a shuffle so the compiler knows it is warp-uniform:
  // every address of the copy engine commands below then lives in uniform registers
  // and the one-lane issue needs no per-lane "waterfall" loop around UBLKCP.
  const int warp = (GD_TUNE_DEFAULT & kTuneLane0) ? (tid >> 5) : __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wmode = WM >= 0 ? WM : a.wmode;
  const bool want_rows = a.row_loss != nullptr;
  gd::PairParams<float> pp = a.pp;
  const bool mask_zero = a.mask_zero_w != 0;
  const float scale = effective_scale(a);
  const bool probe = a.status != nullptr || a.early_return != 0 ||
                     a.host_flag != nullptr;                         // any(weight > 0), ref:290
  bool anyp = false;
  const int wcols = wmode == GD_WEIGHT_ROW7 ? 7 : 1;
  // row strides in floats / lead-in words of a tile in shared memory
  const int ps = ANY ? (int)a.pstride : 7, ts = ANY ? (int)a.tstride : 7;
  const int wsd = ANY ? (int)a.wstride : wcols;
  const int psh = ANY ? a.pshift : 0, tsh = ANY ? a.tshift : 0, wsh = ANY ? a.wshift : 0;
  const long long row_lo = ANY ? a.row_lo : 0;
  if (SPEC >= 0) {
    pp.fun = SPEC & 3;
    pp.tau_on = (SPEC >> 2) & 1;
    pp.flag = (SPEC >> 3) & 1;
  }
  [[maybe_unused]] gd::PairParams<gd::f2> pp2;
  if constexpr (PACK) pp2 = gd::broadcast_params(pp);
  const WarpLayout L = warp_layout(R, wmode, GRAD, want_rows, ANY, ps, ts, wsd);
  unsigned char* base = smem + (size_t)warp * L.per_warp;
  unsigned char* out_base = base + kWarpStages * L.stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_base + L.out);
  float* og = reinterpret_cast<float*>(out_base);
  float* orow = reinterpret_cast<float*>(out_base + (GRAD ? kTileRows * kRowBytes : 0));

  // rows the bulk path can move: a multiple of 4 rows keeps every copy a multiple of 16 B
  // (ANY: rows [row_lo, row_lo + n_main); tile row indices below are relative to row_lo)
  const long long n_main = ANY ? a.n_bulk : (a.n & ~3LL);
  const int cta_warps = (int)(blockDim.x >> 5);
  const long long gwarp = (long long)blockIdx.x * cta_warps + warp;
  const long long nwarps = (long long)gridDim.x * cta_warps;
  // Tile schedule: `full_rounds` rounds in which every warp takes one full tile
  // (tile index gwarp + round * nwarps: the grid sweeps a contiguous window of each
  // array), then ONE balanced round in which the remaining rows are split evenly over
  // all warps (a multiple of 4 rows each), so that every warp finishes together instead
  // of a fraction of them running one more full tile while the rest idle.
  const long long round_rows = nwarps * kTileRows;
  const int full_rounds = (int)(n_main / round_rows);
  const long long tail_base = (long long)full_rounds * round_rows;
  const int last_rows = (int)((((n_main - tail_base) + nwarps - 1) / nwarps + 3) & ~3LL);  // <= kTileRows
  const int my_n = full_rounds + ((tail_base + gwarp * last_rows < n_main) ? 1 : 0);
  auto tile_of = [&](int i, long long* row0) -> int {     // rows of this warp's i-th tile
    if (i < full_rounds) {
      *row0 = (gwarp + (long long)i * nwarps) * kTileRows;
      return kTileRows;
    }
    *row0 = tail_base + gwarp * last_rows;
    return (int)min((long long)last_rows, n_main - *row0);
  };
#if GD_TUNE
  const int tune = a.tune;
#else
  constexpr int tune = GD_TUNE_DEFAULT;
#endif
  const uint64_t policy = (tune & kTuneLoadNormal) ? policy_evict_normal() : policy_evict_first();
  const bool early_wait = !(tune & (kTuneLateWait | kTuneGenericStore));
  // One elected lane (always the same one: elect.sync is deterministic for a full
  // mask) issues, commits and waits for the warp's copy-engine commands.
  const bool leader = (tune & kTuneLane0) ? (lane == 0) : elect_one();
  auto late_wait = [&](int i) {           // out buffer must be free before it is rewritten
    if ((tune & kTuneLateWait) && !(tune & kTuneGenericStore)) {
      if (leader && L.out && i > 0) bulk_wait_read<0>();
      __syncwarp();
    }
  };

  auto issue = [&](int i) {               // leader only
    long long row0;
    const uint32_t rows = (uint32_t)tile_of(i, &row0);
    const int s = i & (kWarpStages - 1);
    unsigned char* st = base + s * L.stage;
    if constexpr (ANY) {
      // 16-byte aligned window around the tile's wide rows: starts `shift` words before the
      // first element and is rounded up; the bytes past the last row of the tile belong to
      // the next row of the array, which exists (the host keeps >= 1 row behind n_bulk)
      const long long g0 = row_lo + row0;
      const uint32_t pb = (uint32_t)(((psh + rows * ps) * 4 + 15) & ~15);
      const uint32_t tb = (uint32_t)(((tsh + rows * ts) * 4 + 15) & ~15);
      const uint32_t wb = wmode ? (uint32_t)(((wsh + rows * wsd) * 4 + 15) & ~15) : 0u;
      mbar_arrive_expect_tx(&bars[s], pb + tb + wb);
      bulk_load(st, a.pred + g0 * ps - psh, pb, &bars[s], policy);
      bulk_load(st + L.ptile, a.target + g0 * ts - tsh, tb, &bars[s], policy);
      if (wmode) bulk_load(st + L.ptile + L.ttile, a.weight + g0 * wsd - wsh, wb, &bars[s], policy);
    } else {
      const uint32_t box_bytes = rows * kRowBytes;
      const uint32_t w_bytes = wmode ? rows * wcols * 4u : 0u;
      mbar_arrive_expect_tx(&bars[s], 2 * box_bytes + w_bytes);
      bulk_load(st, a.pred + row0 * 7, box_bytes, &bars[s], policy);
      bulk_load(st + kTileRows * kRowBytes, a.target + row0 * 7, box_bytes, &bars[s], policy);
      if (wmode)
        bulk_load(st + 2 * kTileRows * kRowBytes, a.weight + row0 * wcols, w_bytes, &bars[s], policy);
    }
  };

  if (leader) {
#pragma unroll
    for (int s = 0; s < kWarpStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    const int pre = my_n < kWarpStages ? my_n : kWarpStages;
    for (int i = 0; i < pre; ++i) issue(i);
  }
  __syncwarp();

  float acc = 0.0f;
  for (int i = 0; i < my_n; ++i) {
    const int s = i & (kWarpStages - 1);
    long long row0;
    const int rows = tile_of(i, &row0);
    const unsigned char* st = base + s * L.stage;
    const float* sp = reinterpret_cast<const float*>(st) + psh;
    const float* stg = reinterpret_cast<const float*>(st + L.ptile) + tsh;
    const float* sw = reinterpret_cast<const float*>(st + L.ptile + L.ttile) + wsh;

    mbar_wait(&bars[s], (uint32_t)((i / kWarpStages) & 1));
    // Lane l owns the R consecutive rows R*l .. R*l + R - 1 of the tile: 28 R contiguous
    // bytes per array, moved with 128-bit (R = 4) shared-memory accesses.  Rows past a
    // partial tile's end read stale shared memory; they are never redone or stored.
    float p[R][7], t[R][7], w[R];
    bool wpos[R];                          // any element of the row's weight > 0 (probe only)
    const bool strided = ANY || (tune & kTuneStrided);
    if (strided) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int r = lane + 32 * k;       // word s*r+c -> bank (s lane + c) mod 32: conflict free, s odd
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          p[k][c] = sp[ps * r + c];
          t[k][c] = stg[ts * r + c];
        }
        if (wmode == GD_WEIGHT_ROW7) {
          float w7[7];
#pragma unroll
          for (int c = 0; c < 7; ++c) w7[c] = sw[wsd * r + c];
          w[k] = row_weight_smem(w7, GD_WEIGHT_ROW7, 0);
          wpos[k] = false;
          if (probe) {
#pragma unroll
            for (int c = 0; c < 7; ++c) wpos[k] |= w7[c] > 0.0f;
          }
        } else {
          w[k] = wmode == GD_WEIGHT_ROW ? sw[wsd * r] : 1.0f;
          wpos[k] = w[k] > 0.0f;
        }
      }
    } else {
      smem_load<7 * R>(&p[0][0], sp + 7 * R * lane);
      smem_load<7 * R>(&t[0][0], stg + 7 * R * lane);
      if (wmode == GD_WEIGHT_ROW) {
        smem_load<R>(w, sw + R * lane);
      } else if (wmode == GD_WEIGHT_ROW7) {
        float w7[R][7];
        smem_load<7 * R>(&w7[0][0], sw + 7 * R * lane);
#pragma unroll
        for (int k = 0; k < R; ++k) {
          w[k] = row_weight_smem(&w7[k][0], GD_WEIGHT_ROW7, 0);
          wpos[k] = false;
          if (probe) {
#pragma unroll
            for (int c = 0; c < 7; ++c) wpos[k] |= w7[k][c] > 0.0f;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < R; ++k) w[k] = 1.0f;
      }
      if (wmode != GD_WEIGHT_ROW7) {
#pragma unroll
        for (int k = 0; k < R; ++k) wpos[k] = w[k] > 0.0f;
      }
    }
    if (a.host_flag != nullptr && i == 0 && gwarp == 0) {
      // a host thread is waiting for any(weight > 0): the common answer is in the first tile
      bool pos = false;
#pragma unroll
      for (int k = 0; k < R; ++k) pos |= wpos[k] && ((strided ? lane + 32 * k : R * lane + k) < rows);
      if (__any_sync(0xffffffffu, pos) && lane == 0) {
        *reinterpret_cast<volatile int*>(a.host_flag) = 1;
        __threadfence_system();
      }
    }
    // the previous tile's store must have finished READING the output buffer
    if (early_wait && leader && L.out && i > 0) bulk_wait_read<0>();
    __syncwarp();                          // stage s consumed by every lane; out buffer free
    if (leader && i + kWarpStages < my_n) issue(i + kWarpStages);

    // The R rows of this lane go through the branch-free FAST math as one
    // straight-line block, so their instruction streams interleave; rows it flags
    // (clamped / degenerate extents, huge yaw, masked weight, ...) are redone on the
    // robust path.  FULL: every lane row is valid (all tiles but the warp's last one).
    auto row_of = [&](int k) { return strided ? lane + 32 * k : R * lane + k; };
    auto eval_tile = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float g[R][7], rl[R], lw[R];
      bool rare[R];
      if constexpr (PACK) {
        static_assert(R % 2 == 0, "packed math pairs the rows of a lane");
#pragma unroll
        for (int k = 0; k < R; k += 2) {
          rare[k] = mask_zero && w[k] == 0.0f;
          rare[k + 1] = mask_zero && w[k + 1] == 0.0f;
          const float ws0 = w[k] * scale, ws1 = w[k + 1] * scale;
          float l0, l1;
          gd::pair_eval_fast2<LOSS, GRAD, kDiet>(p[k], t[k], p[k + 1], t[k + 1], pp2, ws0, ws1, g[k],
                                          g[k + 1], &rare[k], &rare[k + 1], &l0, &l1);
          rl[k] = l0 * ws0;
          lw[k] = l0 * w[k];
          rl[k + 1] = l1 * ws1;
          lw[k + 1] = l1 * w[k + 1];
          if (!FULL) {                         // stale rows: never redone
            rare[k] = rare[k] && (row_of(k) < rows);
            rare[k + 1] = rare[k + 1] && (row_of(k + 1) < rows);
          }
        }
      } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        rare[k] = mask_zero && w[k] == 0.0f;
        const float ws = w[k] * scale;
        float l;
        if (GD_TUNE && ((tune & kTuneNoMath) || ((tune & kTuneHalfMath) && k >= R / 2))) {
          l = p[k][0];
#pragma unroll
          for (int c = 0; c < 7; ++c) g[k][c] = p[k][c] + t[k][c] * ws;
        } else {
          l = gd::pair_eval_fast<float, LOSS, GRAD, kDiet>(p[k], t[k], pp, ws, g[k], &rare[k]);
        }
        rl[k] = l * ws;
        lw[k] = l * w[k];
        if (!FULL) rare[k] = rare[k] && (row_of(k) < rows);   // stale rows: never redone
      }
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (rare[k])
          lw[k] = eval_row<LOSS, GRAD>(p[k], t[k], w[k], pp, scale, mask_zero, g[k], &rl[k]);
      }
      late_wait(i);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (FULL || row_of(k) < rows) {
          acc += lw[k];
          anyp |= wpos[k];
        }
      }
      // rows past the end of a partial tile are written to the staging buffer too (it
      // has room for a full tile); the bulk store below only moves `rows` rows
      if (strided) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int r = lane + 32 * k;
          if (GRAD) {
#pragma unroll
            for (int c = 0; c < 7; ++c) og[7 * r + c] = g[k][c];
          }
          if (want_rows) orow[r] = rl[k];
        }
      } else {
        if (GRAD) smem_store<7 * R>(og + 7 * R * lane, &g[0][0]);
        if (want_rows) smem_store<R>(orow + R * lane, rl);
      }
    };
    if (rows == kTileRows) eval_tile(FullTile{});
    else eval_tile(PartTile{});

    if (L.out && (tune & kTuneGenericStore)) {
      __syncwarp();                        // tile complete in shared memory
      if (GRAD) {                          // tile base is 16-B aligned, rows*7 a multiple of 4
        const float4* s4 = reinterpret_cast<const float4*>(og);
        float4* g4 = reinterpret_cast<float4*>(a.grad + (row_lo + row0) * 7);
        const int nv = (rows * 7) >> 2;
        for (int j = lane; j < nv; j += 32) __stcs(g4 + j, s4[j]);
      }
      if (want_rows)
        for (int j = lane; j < rows; j += 32) __stcs(a.row_loss + row_lo + row0 + j, orow[j]);
      __syncwarp();                        // reads done before the next tile rewrites og
    } else if (L.out) {
      fence_proxy_async_smem();            // generic-proxy writes -> visible to the bulk engine
      __syncwarp();
      if (leader) {
        if (tune & kTuneStoreHint) {
          if (GRAD) bulk_store_hint(a.grad + (row_lo + row0) * 7, og, (uint32_t)rows * kRowBytes, policy);
          if (want_rows) bulk_store_hint(a.row_loss + row_lo + row0, orow, (uint32_t)rows * 4u, policy);
        } else {
          if (GRAD) bulk_store(a.grad + (row_lo + row0) * 7, og, (uint32_t)rows * kRowBytes);
          if (want_rows) bulk_store(a.row_loss + row_lo + row0, orow, (uint32_t)rows * 4u);
        }
        bulk_commit();
      }
    }
  }

  // leftover rows -- n % 4 <= 3 of them, or (ANY) the <= 4 rows in front of row_lo and the
  // <= 4 behind row_lo + n_main: block 0 warp 0, straight from global memory
  if (blockIdx.x == 0 && tid < (int)(a.n - n_main)) {
    const long long r = tid < row_lo ? tid : n_main + tid;
    float p[7], t[7], g[7], rl, w = 1.0f;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      p[c] = a.pred[r * ps + c];
      t[c] = a.target[r * ts + c];
    }
    if (wmode == GD_WEIGHT_ROW) {
      w = a.weight[r * wsd];
      anyp |= w > 0.0f;
    }
    if (wmode == GD_WEIGHT_ROW7) {
      w = row_weight_smem(a.weight + r * wsd, GD_WEIGHT_ROW7, 0);
      if (probe) {
#pragma unroll
        for (int c = 0; c < 7; ++c) anyp |= a.weight[r * wsd + c] > 0.0f;
      }
    }
    acc += eval_row<LOSS, GRAD>(p, t, w, pp, scale, mask_zero, g, &rl);
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) a.grad[r * 7 + c] = g[c];
    }
    if (want_rows) a.row_loss[r] = rl;
  }
  if (leader && L.out) bulk_wait_all<0>();
  if (a.loss_sum) finish_sum(acc, a, scale, anyp);
}

// ---------------------------------------------------------------------------
// host-side dispatch (left out of the CPU emulation build, which has its own launch loop)
// ---------------------------------------------------------------------------
constexpr int kSmemBudget = 227 * 1024 - 1024;   // opt-in max per CTA minus static + slack
#if !defined(GD_HOST_EMULATION)
template <int LOSS, bool GRAD>
int launch_staged(const LossArgs& a, int grid, cudaStream_t stream) {
  gd_staged_kernel<LOSS, GRAD><<<grid, kThreads, 0, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

// CTAs of the persistent warp kernel (at most one per SM).
//
// The board runs this kernel at its 1000 W power cap, and the delivered bandwidth is not
// monotone in the number of busy SMs: interleaved A/B runs on the pool's boxes
// (profiles/r03_grid.md) show it falling slowly from 148 CTAs (5.70 TB/s) to 134 (5.50), JUMPING
// to 6.1-6.2 TB/s at 132 (one box: 130) and falling slowly again (128: 6.0-6.1, 120: 5.8-5.9);
// 4 s of back-to-back launches keep the gap.  On most boxes 128 CTAs are 5-6 % faster than one
// per SM -- but the position of the jump moves from chip to chip, on one box of seven 128 CTAs
// sat on the wrong side of it (5.61 instead of 5.75 TB/s), and with all 8 GPUs of a box busy
// the slowest GPU does too (1.125 instead of 1.084 ms per step).  So the number is MEASURED:
// the first launch of a process with >= 2^22 rows on a contiguous [N] / no-weight layout times
// the launch it was asked for on both grids (gd_loss_api.cu: calibrate_grid, ~13 ms once per
// device, host-synchronising) and keeps the faster one.  No calibration -- one CTA per SM --
//   * for the instantiations with fewer, heavier warps per CTA ([N,7] weights, row-strided /
//     unaligned inputs): bound by their instruction stream, they lose bandwidth with every SM
//     taken away, no jump (profiles/r03_grid.md);
//   * for ranks of a multi-process job (WORLD_SIZE > 1, as torchrun / dist_train.sh export it);
//   * while the stream is being captured into a CUDA graph, and for smaller launches as long as
//     nothing has been calibrated;
//   * when gd_set_loss_grid(n > 0) pins the grid.
extern std::atomic<int> g_loss_grid;                      // gd_set_loss_grid
extern std::atomic<int> g_loss_grid_cal[kMaxDevices];     // calibrated CTAs per device, 0 = not yet
extern thread_local int t_loss_grid_try;                  // set by the calibration around its launches
inline bool loss_multi_process() {
  static const bool multi = [] {
    const char* e = getenv("WORLD_SIZE");
    return e != nullptr && atoi(e) > 1;
  }();
  return multi;
}
inline long long warp_kernel_ctas(long long sms, bool light) {
  if (t_loss_grid_try > 0) return t_loss_grid_try < sms ? t_loss_grid_try : sms;
  const int forced = g_loss_grid.load(std::memory_order_relaxed);
  if (forced > 0) return forced < sms ? forced : sms;
  if (!light || loss_multi_process()) return sms;
  const int cal = g_loss_grid_cal[current_device()].load(std::memory_order_relaxed);
  return cal > 0 && cal < sms ? cal : sms;
}

template <int LOSS, bool GRAD, int R, int SPEC, int WM, bool PACK = false, bool ANY = false>
int launch_warp_inst(const LossArgs& a_in, int max_grid, cudaStream_t stream) {
#if GD_TUNE
  LossArgs a = a_in;
  if (const char* e = getenv("GD_TUNE_FLAGS")) a.tune = atoi(e);
#else
  const LossArgs& a = a_in;
#endif
  auto kern = gd_warp_kernel<LOSS, GRAD, R, SPEC, WM, PACK, ANY>;
  // opt in to the large dynamic shared memory once per device (function attributes are
  // per context); a racing second call sets the same value
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    const cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const WarpLayout L = warp_layout(R, a.wmode, GRAD, a.row_loss != nullptr, ANY, a.pstride,
                                   a.tstride, a.wstride);
  constexpr int kWarpCap = R >= 4 ? 12 : 24;
  int warps = (int)(kSmemBudget / L.per_warp);
  if (warps > kWarpCap) warps = kWarpCap;
#if GD_TUNE
  if (const char* e = getenv("GD_TUNE_WARPS")) {
    const int w = atoi(e);
    if (w >= 1 && w < warps) warps = w;
  }
#endif
  if (warps < 1) return GD_ERR_BAD_ARG;
  // Persistent: at most one CTA per SM with `warps` warps.  A batch too small to give
  // every such warp 32 rows uses fewer warps, spread over as many SMs as possible.
  const long long sms = warp_kernel_ctas(device_info().sm_count, !ANY && a.wmode != GD_WEIGHT_ROW7);
  const long long n_main = ANY ? a.n_bulk : (a.n & ~3LL);
  long long want = (n_main + 31) / 32;                    // warps that would get >= 32 rows
  if (want < 1) want = 1;
  long long grid = sms < max_grid ? sms : max_grid;
  if (want < grid * warps) {
    const long long per_cta = (want + grid - 1) / grid;   // <= warps
    warps = (int)per_cta;
    grid = (want + per_cta - 1) / per_cta;
  }
  if (grid < 1) grid = 1;
#if GD_TUNE
  // measurement only: fewer active SMs (the board is power capped: does a smaller grid at a
  // higher SM clock move more bytes?)
  if (const char* e = getenv("GD_TUNE_GRID")) {
    const int g = atoi(e);
    if (g >= 1 && g < grid) grid = g;
  }
#endif
  kern<<<(int)grid, warps * 32, (size_t)warps * L.per_warp, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

// Specialised instantiations exist for the shipped configurations (fun in
// {none, log1p}, flag = default true, weights None / [N]); anything else takes the
// run-time-parameter instantiation of the same kernel.
template <int LOSS, bool GRAD, int R, bool PACK = false>
int launch_warp(const LossArgs& a, int max_grid, cudaStream_t stream) {
  constexpr bool kHasSpec = GRAD && (LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd);
  static_assert(!PACK || kHasSpec, "packed math exists for the specialised instantiations only");
  if constexpr (kHasSpec) {
    const gd::PairParams<float>& pp = a.pp;
    bool spec_ok = pp.flag == 1 && (pp.fun == gd::kFunNone || pp.fun == gd::kFunLog1p);
    if (GD_TUNE_DEFAULT & kTuneStd)       // the specialised kernels bake these values in
      spec_ok = spec_ok && pp.alpha2 == 1.0f && pp.inv_alpha2 == 1.0f && pp.off[0] == 0.0f &&
                pp.off[1] == 0.0f && pp.off[2] == 0.5f;
    if (spec_ok) {
      const int spec = pp.fun | (pp.tau_on << 2) | (1 << 3);
#define GD_SPEC_CASE(S)                                                                 \
  case S:                                                                               \
    if constexpr (!PACK) {                                                              \
      if (a.wmode == GD_WEIGHT_ROW7)                                                    \
        return launch_warp_inst<LOSS, GRAD, R, S, 2, PACK>(a, max_grid, stream);        \
    }                                                                                   \
    return a.wmode == GD_WEIGHT_ROW                                                     \
               ? launch_warp_inst<LOSS, GRAD, R, S, 1, PACK>(a, max_grid, stream)       \
               : launch_warp_inst<LOSS, GRAD, R, S, 0, PACK>(a, max_grid, stream);
      switch (spec) {
        GD_SPEC_CASE(8) GD_SPEC_CASE(9) GD_SPEC_CASE(12) GD_SPEC_CASE(13)
        default: break;
      }
#undef GD_SPEC_CASE
    }
  }
  return launch_warp_inst<LOSS, GRAD, R, -1, -1>(a, max_grid, stream);
}

template <int LOSS>
int launch_loss(const LossArgs& a, int variant, int max_grid, cudaStream_t stream) {
  const bool grad = a.grad != nullptr;
  if (variant == GD_VARIANT_BULK) {
    return grad ? launch_warp<LOSS, true, 4>(a, max_grid, stream)
                : launch_warp<LOSS, false, 4>(a, max_grid, stream);
  }
  if (variant == GD_VARIANT_BULK_PACKED) {
    // packed math: gradient + one of the three headline losses; anything else (and any
    // configuration without a specialised instantiation) runs the scalar bulk kernel
    if constexpr (LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd) {
      if (grad && a.wmode != GD_WEIGHT_ROW7 && !a.row_loss)
        return launch_warp<LOSS, true, 4, true>(a, max_grid, stream);
    }
    return grad ? launch_warp<LOSS, true, 4>(a, max_grid, stream)
                : launch_warp<LOSS, false, 4>(a, max_grid, stream);
  }
  if (variant == GD_VARIANT_BULK_R2) {
    return grad ? launch_warp<LOSS, true, 2>(a, max_grid, stream)
                : launch_warp<LOSS, false, 2>(a, max_grid, stream);
  }
  if (variant == GD_VARIANT_BULK_ANY) {
    // Row-strided / unaligned inputs.  The CenterGDHead call (no weight, gradient, a shipped
    // fun / tau / flag configuration of one of the three headline losses) gets the specialised
    // packed-FP32 instantiation: this layout runs 9 warps per CTA and is bound by its
    // instruction stream (profiles/r03_grid.md), and the run-time-parameter kernel executes
    // 55 % more instructions per row.  Everything else: run-time parameters, scalar math.
    if constexpr ((LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd) &&
                  (GD_TUNE_DEFAULT & kTunePacked) != 0) {
      const gd::PairParams<float>& pp = a.pp;
      bool spec_ok = grad && !a.row_loss && a.wmode == GD_WEIGHT_NONE && pp.flag == 1 &&
                     (pp.fun == gd::kFunNone || pp.fun == gd::kFunLog1p);
      if (GD_TUNE_DEFAULT & kTuneStd)
        spec_ok = spec_ok && pp.alpha2 == 1.0f && pp.inv_alpha2 == 1.0f && pp.off[0] == 0.0f &&
                  pp.off[1] == 0.0f && pp.off[2] == 0.5f;
      if (spec_ok) {
        switch (pp.fun | (pp.tau_on << 2) | (1 << 3)) {
          case 8: return launch_warp_inst<LOSS, true, 4, 8, 0, true, true>(a, max_grid, stream);
          case 9: return launch_warp_inst<LOSS, true, 4, 9, 0, true, true>(a, max_grid, stream);
          case 12: return launch_warp_inst<LOSS, true, 4, 12, 0, true, true>(a, max_grid, stream);
          case 13: return launch_warp_inst<LOSS, true, 4, 13, 0, true, true>(a, max_grid, stream);
          default: break;
        }
      }
    }
    return grad ? launch_warp_inst<LOSS, true, 4, -1, -1, false, true>(a, max_grid, stream)
                : launch_warp_inst<LOSS, false, 4, -1, -1, false, true>(a, max_grid, stream);
  }
  long long grid = (a.n + kTile - 1) / kTile;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  return grad ? launch_staged<LOSS, true>(a, (int)grid, stream)
              : launch_staged<LOSS, false>(a, (int)grid, stream);
}

#endif  // !GD_HOST_EMULATION

// Tile plan of gd_warp_kernel<..., ANY = true>: rows [4, 4 + n_bulk) in tiles (>= 1 row stays
// behind them: the 16-byte aligned copy window of the last tile may reach 12 bytes into the
// next row, and the window of the first tile up to 12 bytes into row 3), the <= 8 rows around
// them straight from global memory; the lead-in words of each array are constant because
// tiles start at multiples of 4 rows (16 * stride bytes).  Needs n >= 16.
inline void plan_any(LossArgs* a) {
  a->row_lo = 4;
  a->n_bulk = (a->n - 1 - a->row_lo) & ~3LL;
  auto shift = [&](const float* base, long long stride) {
    return (int)((reinterpret_cast<uintptr_t>(base + a->row_lo * stride) & 15u) >> 2);
  };
  a->pshift = shift(a->pred, a->pstride);
  a->tshift = shift(a->target, a->tstride);
  a->wshift = a->wmode == GD_WEIGHT_NONE ? 0 : shift(a->weight, a->wstride);
}

constexpr int kMaxGrid = 65536;           // partials capacity of the workspace

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace gdk

