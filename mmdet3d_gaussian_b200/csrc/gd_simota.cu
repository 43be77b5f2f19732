// SimOTA-style consumer of the pairwise Gaussian distance WITHOUT the N x M matrix
// (second branch).
//
// The reference's SimOTA (mmdet3d_gaussian/core/bbox/assigners/sim_ota_3d_assigner.py, "ref")
// materialises an IoU matrix (ref:91-93), builds a cost from it (ref:94-107) and runs
// dynamic_k_matching (ref:184-211):
//     topk_ious = topk(pairwise_ious, candidate_topk, dim=0)            ref:187-188
//     dynamic_ks = clamp(topk_ious.sum(0).int(), min=1)                 ref:190
//     per GT: the dynamic_k rows of lowest cost are matched             ref:191-194
//     a row matched by several GTs keeps argmin_j cost[row, j]          ref:198-203
// With the Gaussian similarity s = 1 - D (D in [0, 1) for tau >= 1) in the role of the IoU and
// the cost monotone in D, everything the matching reads lives in (a) the candidate_topk
// smallest D of every COLUMN and (b) the (min, argmin) of every ROW.  Both are reductions of
// the pairwise kernel's value stream:
//   gd_pairwise_topk_kernel  -- the pairwise value loop (lane <-> column box in registers, row
//                               Gaussians broadcast from shared memory) with a K-deep sorted
//                               list of (value key, row) per lane and the row minima of the
//                               assign kernel; lists of all CTAs / row phases go to a scratch
//   gd_topk_merge_kernel     -- one warp per column merges the scratch lists
//   gd_simota_cols_kernel / gd_simota_rows_kernel -- dynamic k, matching, conflict rule
// Keys are the pairwise kernel's: order-preserving 32-bit value key (NaN lowest) in the high
// word, row index in the low word -- ties go to the lowest row, deterministically (torch.topk
// leaves tie order unspecified).
#include "gd_pairwise.cuh"

namespace gdk {

constexpr int kTopK = 16;                    // list depth kept per column (candidate_topk <= 16)

struct TopkArgs {
  unsigned long long* cand;                  // [slots][kTopK][m] scratch
  long long slots;
  // Sample pass (row_step > 1): the kernel sees rows 0, row_step, 2 row_step, ... (a.n of them)
  // and writes no row minima.  Main pass: thr[j] (nullable) = the k-th smallest (key, row) of
  // column j over the sample -- an upper bound of the k-th smallest over all rows, so only
  // candidates <= thr[j] can belong to the final top k and only they are inserted.
  long long row_step;
  const unsigned long long* thr;
};

// sorted insert of x into the ascending list best[0..K): branch-free bubble
template <int K>
__device__ __forceinline__ void topk_insert(unsigned long long (&best)[K], unsigned long long x) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const unsigned long long lo = best[i] < x ? best[i] : x;
    x = best[i] < x ? x : best[i];
    best[i] = lo;
  }
}

template <int LOSS>
__global__ void __launch_bounds__(kThreads, 3) gd_pairwise_topk_kernel(const PairwiseArgs a,
                                                                       const TopkArgs tk) {
  __shared__ gd::BoxGauss<float> s_rows[kRowsPerCta];
  __shared__ unsigned long long s_best[kRowsPerCta][kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gd::PairParams<float> pp = a.pp;
  pp.lean = 1;
  const int wx = pairwise_wx(a.m, 32);
  const int wy = kWarps / wx;
  const int cgrp = warp % wx, ry = warp / wx;
  const long long chunk = 32LL * wx;
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  const long long slot = (long long)blockIdx.x * wy + ry;

  // columns outer (each lane's list lives across all row tiles of this CTA), tiles inner
  for (long long c0 = 0; c0 < a.m; c0 += chunk) {
    const long long j = c0 + 32LL * cgrp + lane;
    const bool warp_live = c0 + 32LL * cgrp < a.m;
    const bool live = j < a.m;
    unsigned long long best[kTopK];
#pragma unroll
    for (int i = 0; i < kTopK; ++i) best[i] = ~0ull;
    // insert limit: strictly below it a candidate may still belong to the top k
    unsigned long long cap = ~0ull;
    if (tk.thr != nullptr && live) {
      const unsigned long long t = tk.thr[j];
      cap = t == ~0ull ? t : t + 1;
    }
    unsigned long long lim = cap;
    gd::BoxGauss<float> t;
    if (warp_live) t = gd::box_gauss(a.b2 + (live ? j : 0) * 7, pp);   // dead lanes: any valid box
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * kRowsPerCta;
      const int rows = (int)min((long long)kRowsPerCta, a.n - row0);
      __syncthreads();                         // previous tile fully consumed
      if (tid < rows) s_rows[tid] = gd::box_gauss(a.b1 + (row0 + tid) * tk.row_step * 7, pp);
      for (int i = tid; i < kRowsPerCta * kWarps; i += kThreads) (&s_best[0][0])[i] = ~0ull;
      __syncthreads();
      if (warp_live) {
        for (int r = ry; r < rows; r += wy) {
          const float v = gd::pair_value_auto<float, LOSS>(s_rows[r], t, pp);
          if (a.out != nullptr && live) __stcs(a.out + (row0 + r) * a.out_stride + j, v);
          const unsigned int key = live ? order_key(v) : 0xffffffffu;
          const unsigned int mn = __reduce_min_sync(0xffffffffu, key);
          const unsigned int who = __ballot_sync(0xffffffffu, key == mn);
          if (lane == 0 && mn != 0xffffffffu)
            s_best[r][warp] = ((unsigned long long)mn << 32) |
                              (unsigned int)(c0 + 32LL * cgrp + (__ffs(who) - 1));
          if (live) {
            const unsigned long long k64 =
                ((unsigned long long)key << 32) | (unsigned int)((row0 + r) * tk.row_step);
            if (k64 < lim) {
              topk_insert<kTopK>(best, k64);
              lim = best[kTopK - 1] < cap ? best[kTopK - 1] : cap;
            }
          }
        }
      }
      // row minima of this tile over this chunk of columns, merged into the outputs (chunks are
      // walked in ascending column order by every CTA, so "strictly smaller wins" keeps the
      // lowest column on ties)
      __syncthreads();
      if (tid < rows && a.row_min != nullptr) {
        unsigned long long k = s_best[tid][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) k = s_best[tid][w] < k ? s_best[tid][w] : k;
        const unsigned int key = (unsigned int)(k >> 32);
        if (c0 == 0) {
          a.row_min[row0 + tid] = key_value(key);
          a.row_argmin[row0 + tid] = (int)(unsigned int)(k & 0xffffffffu);
        } else if (k != ~0ull && key < order_key(a.row_min[row0 + tid])) {
          a.row_min[row0 + tid] = key_value(key);
          a.row_argmin[row0 + tid] = (int)(unsigned int)(k & 0xffffffffu);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < kTopK; ++i) tk.cand[(slot * kTopK + i) * a.m + j] = best[i];
    }
  }
}

// Merge of the per-slot candidate lists, coalesced and in two stages.  Lane <-> column (a warp
// reads 32 consecutive columns of one list entry: 256 contiguous bytes), the 8 warps of a CTA
// and the gridDim.y CTA groups split the slots; every slot list is sorted ascending, so a lane
// stops walking a list at the first entry that cannot enter its own.  The warps' lists meet in
// shared memory and warp 0 merges them; with `cand_out` the CTA writes its K-list in the input
// layout ([gridDim.y][K][m]: stage A, many CTAs), without it the final values / rows (stage B).
// (Round 2 ran one warp per column over strided 8-byte loads: 256 warps, ~0.15 ms at C4.)
constexpr int kMergeGroups = 16;
__global__ void __launch_bounds__(kThreads) gd_topk_merge_kernel(
    const unsigned long long* __restrict__ cand, long long slots, long long m, int k,
    unsigned long long* __restrict__ cand_out, float* __restrict__ topk_val,
    int* __restrict__ topk_row, unsigned long long* __restrict__ thr_out) {
  __shared__ unsigned long long s_list[kWarps][kTopK][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long j = (long long)blockIdx.x * 32 + lane;
  const bool live = j < m;
  unsigned long long best[kTopK];
#pragma unroll
  for (int i = 0; i < kTopK; ++i) best[i] = ~0ull;
  if (live) {
    for (long long s = (long long)blockIdx.y * kWarps + warp; s < slots;
         s += (long long)gridDim.y * kWarps) {
      const unsigned long long* list = cand + s * kTopK * m + j;
      // the whole list in flight at once (an early exit would serialise 16 DRAM latencies)
      unsigned long long x[kTopK];
#pragma unroll
      for (int i = 0; i < kTopK; ++i) x[i] = __ldcs(list + (long long)i * m);
#pragma unroll
      for (int i = 0; i < kTopK; ++i) {
        if (x[i] < best[kTopK - 1]) topk_insert<kTopK>(best, x[i]);   // sorted: later ones fail too
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kTopK; ++i) s_list[warp][i][lane] = best[i];
  __syncthreads();
  if (warp != 0 || !live) return;
  for (int w = 1; w < kWarps; ++w) {
    for (int i = 0; i < kTopK; ++i) {
      const unsigned long long x = s_list[w][i][lane];
      if (!(x < best[kTopK - 1])) break;
      topk_insert<kTopK>(best, x);
    }
  }
  if (cand_out != nullptr) {
#pragma unroll
    for (int i = 0; i < kTopK; ++i)
      cand_out[((long long)blockIdx.y * kTopK + i) * m + j] = best[i];
    return;
  }
  if (thr_out != nullptr) {                    // sample pass: only the k-th key is wanted
    unsigned long long t = ~0ull;
#pragma unroll
    for (int i = 0; i < kTopK; ++i)
      if (i == k - 1) t = best[i];
    thr_out[j] = t;
    return;
  }
#pragma unroll
  for (int i = 0; i < kTopK; ++i) {
    if (i < k) {
      const unsigned long long head = best[i];
      const bool none = head == ~0ull;
      topk_val[(long long)i * m + j] =
          none ? __uint_as_float(0x7f800000u) : key_value((unsigned int)(head >> 32));
      topk_row[(long long)i * m + j] = none ? -1 : (int)(unsigned int)(head & 0xffffffffu);
    }
  }
}

// ref:187-194 per GT column: dynamic k from the k largest similarities (= 1 - the k smallest
// distances), then the dynamic_k lowest-cost rows are matched
__global__ void __launch_bounds__(kThreads) gd_simota_cols_kernel(
    const float* __restrict__ topk_val, const int* __restrict__ topk_row, long long m, int k,
    int* __restrict__ match_count, int* __restrict__ match_gt, float* __restrict__ match_val,
    int* __restrict__ dynamic_ks) {
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (j >= m) return;
  float sum = 0.0f;
  int have = 0;
  for (int i = 0; i < k; ++i) {
    if (topk_row[(long long)i * m + j] < 0) break;
    sum += 1.0f - topk_val[(long long)i * m + j];                  // topk_ious.sum(0)      ref:188-190
    ++have;
  }
  int dk = (int)sum;                                               // .int(): truncation     ref:190
  dk = dk < 1 ? 1 : dk;                                            // clamp(min=1)
  dk = dk > have ? have : dk;
  if (dynamic_ks) dynamic_ks[j] = dk;
  for (int i = 0; i < dk; ++i) {                                   // topk(cost[:, j], k=dk, largest=False)  ref:191-194
    const int row = topk_row[(long long)i * m + j];
    atomicAdd(match_count + row, 1);
    match_gt[row] = (int)j;                   // read back only where the count is exactly 1
    match_val[row] = topk_val[(long long)i * m + j];
  }
}

// ref:198-211 per prior row: several GTs -> the row's own argmin over ALL GTs; one -> that GT
__global__ void __launch_bounds__(kThreads) gd_simota_rows_kernel(
    const int* __restrict__ match_count, const int* __restrict__ match_gt,
    const float* __restrict__ match_val, const float* __restrict__ row_min,
    const int* __restrict__ row_argmin, long long n, long long* __restrict__ assigned_gt_inds,
    float* __restrict__ matched_sim, float unmatched_sim) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int c = match_count[i];
  long long gt = 0;                                                // background          ref:64-66
  float sim = unmatched_sim;
  if (c > 1) {                                                     // ref:198-203
    gt = (long long)row_argmin[i] + 1;
    sim = 1.0f - row_min[i];
  } else if (c == 1) {
    gt = (long long)match_gt[i] + 1;
    sim = 1.0f - match_val[i];
  }
  assigned_gt_inds[i] = gt;                                        // matched_gt_inds + 1   ref:112
  if (matched_sim) matched_sim[i] = sim;                           // matched_pred_ious     ref:209-210
}

template <int LOSS>
static int launch_topk(const PairwiseArgs& a, const TopkArgs& tk, long long gx, cudaStream_t st) {
  gd_pairwise_topk_kernel<LOSS><<<(unsigned)gx, kThreads, 0, st>>>(a, tk);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

constexpr long long kSampleRows = 2048;      // rows of the threshold sample (strided over all rows)
constexpr long long kSampleMinRows = 16 * kSampleRows;   // smaller inputs: no sample pass

static long long topk_grid(long long n) {
  const long long ntiles = (n + kRowsPerCta - 1) / kRowsPerCta;
  long long gx = (long long)device_info().sm_count * 3;            // persistent: bounds the scratch
  if (gx > ntiles) gx = ntiles;
  return gx < 1 ? 1 : gx;
}

}  // namespace gdk

extern "C" {

size_t gd_pairwise_topk_workspace_bytes(int64_t n, int64_t m) {
  using namespace gdk;
  if (n < 0 || m < 0) return 0;
  // main lists (wy <= kWarps row phases) + stage-A lists + the sample pass's lists + thresholds
  const long long slots = topk_grid(n) * kWarps + kMergeGroups + topk_grid(kSampleRows) * kWarps + 1;
  return 256 + sizeof(unsigned long long) * (size_t)(slots * kTopK * (m > 0 ? m : 0));
}

int gd_pairwise_col_topk(const gd_loss_config* cfg, const float* boxes1, int64_t n,
                         const float* boxes2, int64_t m, int32_t k, float* row_min,
                         int32_t* row_argmin, float* topk_val, int32_t* topk_row, float* out,
                         int64_t out_row_stride, void* workspace, size_t workspace_bytes,
                         void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m <= 0 || k < 1 || k > kTopK) return GD_ERR_BAD_ARG;
  if (out && out_row_stride < m) return GD_ERR_BAD_ARG;
  if (n > 0x7fffffffLL || m > 0x7fffffffLL) return GD_ERR_BAD_ARG;
  if (!topk_val || !topk_row || (n > 0 && (!boxes1 || !boxes2 || !row_min || !row_argmin)))
    return GD_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < gd_pairwise_topk_workspace_bytes(n, m)) return GD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PairwiseArgs a{};
  a.b1 = boxes1;
  a.n = n;
  a.b2 = boxes2;
  a.m = m;
  a.out = out;
  a.out_stride = out ? out_row_stride : m;
  a.row_min = row_min;
  a.row_argmin = row_argmin;
  a.pp = make_pair_params(*cfg);
  unsigned long long* ws64 =
      reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  const long long gx = topk_grid(n);
  const int wy = kWarps / pairwise_wx(m, 32);
  TopkArgs tk;
  tk.cand = ws64;
  tk.slots = n > 0 ? gx * wy : 0;
  tk.row_step = 1;
  tk.thr = nullptr;
  unsigned long long* mid = tk.cand + (long long)topk_grid(n) * kWarps * kTopK * m;   // stage-A lists
  const unsigned mx = (unsigned)((m + 31) / 32);
  auto launch = [&](const PairwiseArgs& pa, const TopkArgs& pt, long long grid) -> int {
    switch (cfg->loss_type) {
      case GD_LOSS_GWD3D: return launch_topk<gd::kGwd>(pa, pt, grid, st);
      case GD_LOSS_KLD3D: return launch_topk<gd::kKld>(pa, pt, grid, st);
      case GD_LOSS_JD3D: return launch_topk<gd::kJd>(pa, pt, grid, st);
      case GD_LOSS_KLD3D_SYMMAX: return launch_topk<gd::kSymMax>(pa, pt, grid, st);
      case GD_LOSS_KLD3D_SYMMIN: return launch_topk<gd::kSymMin>(pa, pt, grid, st);
      case GD_LOSS_BD3D: return launch_topk<gd::kBd>(pa, pt, grid, st);
      case GD_LOSS_KFIOU3D: return launch_topk<gd::kKfiou>(pa, pt, grid, st);
    }
    return GD_ERR_BAD_ARG;
  };
  // merge of `slots` lists at `lists`: two coalesced stages when there are many
  auto merge = [&](const unsigned long long* lists, long long slots, float* val, int* row,
                   unsigned long long* thr_out) {
    if (slots > 2 * kWarps) {
      gd_topk_merge_kernel<<<dim3(mx, kMergeGroups), kThreads, 0, st>>>(lists, slots, m, k, mid,
                                                                       nullptr, nullptr, nullptr);
      gd_topk_merge_kernel<<<dim3(mx, 1), kThreads, 0, st>>>(mid, kMergeGroups, m, k, nullptr, val,
                                                            row, thr_out);
      g_launches.fetch_add(2, std::memory_order_relaxed);
    } else {
      gd_topk_merge_kernel<<<dim3(mx, 1), kThreads, 0, st>>>(lists, slots, m, k, nullptr, val, row,
                                                            thr_out);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
  };
  int rc = 0;
  if (n >= kSampleMinRows) {
    // Threshold pass.  The row tiles of the main pass are spread over ~450 CTAs, each with its own
    // K-deep list per column: without a bound every list accepts ~K ln(rows / K) candidates and
    // nearly every warp iteration runs the sorted insert (0.49 ms at C4).  The k-th smallest key of
    // a 2048-row strided sample bounds the k-th smallest of the column from above, and with it
    // as the insert limit a list accepts a handful.
    PairwiseArgs sa = a;
    TopkArgs sk = tk;
    sa.n = kSampleRows;
    sa.out = nullptr;
    sa.row_min = nullptr;
    sa.row_argmin = nullptr;
    sk.row_step = n / kSampleRows;
    sk.cand = mid + (long long)kMergeGroups * kTopK * m;
    const long long sgx = topk_grid(kSampleRows);
    sk.slots = sgx * wy;
    unsigned long long* thr = sk.cand + (long long)topk_grid(kSampleRows) * kWarps * kTopK * m;
    rc = launch(sa, sk, sgx);
    if (rc != 0) return rc;
    merge(sk.cand, sk.slots, nullptr, nullptr, thr);
    tk.thr = thr;
  }
  if (n > 0) {
    rc = launch(a, tk, gx);
    if (rc != 0) return rc;
  }
  merge(tk.cand, tk.slots, topk_val, topk_row, nullptr);
  return (int)cudaGetLastError();
}

int gd_simota_from_topk(const float* topk_val, const int32_t* topk_row, int64_t m, int32_t k,
                        const float* row_min, const int32_t* row_argmin, int64_t n,
                        int64_t* assigned_gt_inds, float* matched_sim, int32_t* dynamic_ks,
                        float unmatched_sim, int32_t* scratch, void* stream) {
  using namespace gdk;
  if (n < 0 || m < 0 || k < 1 || k > kTopK) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!assigned_gt_inds || !scratch || !row_min || !row_argmin || (m > 0 && (!topk_val || !topk_row)))
    return GD_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int* count = scratch;
  int* mgt = scratch + n;
  float* mval = reinterpret_cast<float*>(scratch + 2 * n);
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int) * (size_t)n, st);
  if (e != cudaSuccess) return (int)e;
  if (m > 0) {
    gd_simota_cols_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        topk_val, topk_row, m, k, count, mgt, mval, dynamic_ks);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  gd_simota_rows_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(
      count, mgt, mval, row_min, row_argmin, n, reinterpret_cast<long long*>(assigned_gt_inds),
      matched_sim, unmatched_sim);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

}  // extern "C"
