// Host-buffer entry point: the whole GDLoss forward+backward for callers whose
// boxes live in HOST memory (bench `e2e`).
//
// Rows are cut into chunks; slot s = chunk % kSlots owns one stream and one set
// of device buffers, so H2D of chunk k+1, the kernel of chunk k and D2H of chunk
// k-1 run concurrently (PCIe is full duplex; the kernel is ~100x faster than the
// link).  Per-chunk loss partials come back through pinned memory and are added
// on the host in chunk order in fp64 => deterministic.  This is the one entry
// point that owns device memory (a per-device cache that only grows).
#include <mutex>
#include <vector>

#include "gd_common.cuh"

namespace gdk {

constexpr int kSlots = 3;
constexpr int kMaxChunks = 1 << 16;

struct Slot {
  cudaStream_t stream = nullptr;
  float* pred = nullptr;
  float* target = nullptr;
  float* weight = nullptr;
  float* grad = nullptr;
  float* loss = nullptr;      // [1]
  void* ws = nullptr;
  size_t ws_bytes = 0;
  long long cap_rows = 0;     // capacity of pred/target/grad
  long long cap_w = 0;        // capacity of weight in floats
};

struct Pipe {
  bool init = false;
  Slot slot[kSlots];
  float* loss_pinned = nullptr;   // [kMaxChunks]
};

static Pipe g_pipe[64];
static std::mutex g_mu[64];          // one pipeline per device: callers on different GPUs do not serialise

#define GD_TRY(expr)                          \
  do {                                        \
    const cudaError_t e__ = (expr);           \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

static int ensure(Pipe& p, long long chunk_rows, long long w_floats) {
  if (!p.init) {
    // every resource is created at most once: a call that fails half way leaves what it got in
    // place and the next call continues from there (nothing is leaked, nothing is created twice)
    for (int s = 0; s < kSlots; ++s) {
      Slot& sl = p.slot[s];
      if (!sl.stream) GD_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
      if (!sl.loss) GD_TRY(cudaMalloc(&sl.loss, sizeof(float)));
      if (!sl.ws) {
        const size_t bytes = gd_loss_workspace_bytes(chunk_rows);
        void* ws = nullptr;
        GD_TRY(cudaMalloc(&ws, bytes));
        const cudaError_t e = cudaMemset(ws, 0, bytes);
        if (e != cudaSuccess) {
          cudaFree(ws);
          return (int)e;
        }
        sl.ws = ws;
        sl.ws_bytes = bytes;
      }
    }
    if (!p.loss_pinned) GD_TRY(cudaMallocHost(&p.loss_pinned, sizeof(float) * kMaxChunks));
    p.init = true;
  }
  // staging buffers only ever grow (to the largest chunk this process has used): a steady
  // workload allocates once; cudaFree / cudaMalloc below are the documented sync points of a
  // chunk-size increase
  for (int s = 0; s < kSlots; ++s) {
    Slot& sl = p.slot[s];
    if (sl.cap_rows < chunk_rows) {
      cudaFree(sl.pred);
      cudaFree(sl.target);
      cudaFree(sl.grad);
      sl.pred = sl.target = sl.grad = nullptr;
      sl.cap_rows = 0;
      GD_TRY(cudaMalloc(&sl.pred, (size_t)chunk_rows * kRowBytes));
      GD_TRY(cudaMalloc(&sl.target, (size_t)chunk_rows * kRowBytes));
      GD_TRY(cudaMalloc(&sl.grad, (size_t)chunk_rows * kRowBytes));
      sl.cap_rows = chunk_rows;
    }
    if (sl.cap_w < w_floats) {
      cudaFree(sl.weight);
      sl.weight = nullptr;
      sl.cap_w = 0;
      GD_TRY(cudaMalloc(&sl.weight, (size_t)w_floats * sizeof(float)));
      sl.cap_w = w_floats;
    }
  }
  return 0;
}

}  // namespace gdk

extern "C" int64_t gd_host_chunk_plan(int64_t n, int64_t chunk_rows, int64_t* starts, int64_t* rows,
                                      int64_t capacity) {
  using namespace gdk;
  if (n < 0 || capacity < 0 || (capacity > 0 && (!starts || !rows))) return GD_ERR_BAD_ARG;
  if (chunk_rows <= 0) chunk_rows = 1 << 20;
  chunk_rows = (chunk_rows + 255) & ~255LL;            // whole tiles; keeps chunks 16 B aligned
  long long piece_min = (chunk_rows / 8 + 255) & ~255LL;
  if (piece_min < 256) piece_min = 256;
  const bool taper = n > chunk_rows;                   // single-chunk inputs stay one launch
  int64_t count = 0;
  long long r = 0;
  while (r < n) {
    long long take = n - r;
    if (take > chunk_rows) {
      take = chunk_rows;                               // a full chunk, more rows follow
    } else if (taper && take > piece_min) {
      // inside the last chunk: halve (rounded up to 256 rows) until piece_min is reached
      long long half = ((take + 1) / 2 + 255) & ~255LL;
      if (half < piece_min) half = piece_min;
      if (half < take) take = half;
    }
    if (count >= kMaxChunks) return GD_ERR_BAD_ARG;
    if (count < capacity) {
      starts[count] = r;
      rows[count] = take;
    } else if (capacity > 0) {
      return GD_ERR_BAD_ARG;
    }
    ++count;
    r += take;
  }
  return count;
}

extern "C" int gd_loss_fwd_bwd_host(const gd_loss_config* cfg, const float* pred_host,
                                    const float* target_host, const float* weight_host,
                                    int32_t weight_mode, int64_t n, float scale,
                                    float* loss_host, float* grad_host, int32_t device,
                                    int64_t chunk_rows) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || !loss_host || device < 0 || device >= 64 ||
      weight_mode < GD_WEIGHT_NONE || weight_mode > GD_WEIGHT_ROW7)
    return GD_ERR_BAD_ARG;
  if (n > 0 && (!pred_host || !target_host || (weight_mode != GD_WEIGHT_NONE && !weight_host)))
    return GD_ERR_BAD_ARG;
  if (n == 0) {
    *loss_host = 0.0f;
    return 0;
  }
  if (chunk_rows <= 0) chunk_rows = 1 << 20;
  chunk_rows = (chunk_rows + 255) & ~255LL;            // whole tiles; keeps chunks 16 B aligned
  if (chunk_rows > n) chunk_rows = (n + 255) & ~255LL;
  // chunk boundaries: full chunks, the last one tapered (see gd_host_chunk_plan)
  const int64_t nplan = gd_host_chunk_plan(n, chunk_rows, nullptr, nullptr, 0);
  if (nplan < 0) return (int)nplan;
  std::vector<int64_t> c_start((size_t)nplan), c_rows((size_t)nplan);
  if (gd_host_chunk_plan(n, chunk_rows, c_start.data(), c_rows.data(), nplan) != nplan)
    return GD_ERR_BAD_ARG;
  const long long nchunks = nplan;
  const int wcols = weight_mode == GD_WEIGHT_ROW7 ? 7 : (weight_mode == GD_WEIGHT_ROW ? 1 : 0);

  std::lock_guard<std::mutex> lock(g_mu[device]);
  int prev_dev = 0;
  GD_TRY(cudaGetDevice(&prev_dev));
  GD_TRY(cudaSetDevice(device));
  Pipe& p = g_pipe[device];
  int rc = ensure(p, chunk_rows, chunk_rows * (wcols ? wcols : 1));
  if (rc != 0) {
    cudaSetDevice(prev_dev);
    return rc;
  }
  for (long long c = 0; c < nchunks && rc == 0; ++c) {
    Slot& sl = p.slot[c % kSlots];
    const long long r0 = c_start[(size_t)c];
    const long long rows = c_rows[(size_t)c];
    cudaError_t e = cudaMemcpyAsync(sl.pred, pred_host + r0 * 7, (size_t)rows * kRowBytes,
                                    cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(sl.target, target_host + r0 * 7, (size_t)rows * kRowBytes,
                          cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess && wcols)
      e = cudaMemcpyAsync(sl.weight, weight_host + r0 * wcols, (size_t)rows * wcols * 4,
                          cudaMemcpyHostToDevice, sl.stream);
    if (e != cudaSuccess) {
      rc = (int)e;
      break;
    }
    rc = gd_loss_fwd_bwd(cfg, sl.pred, 7, sl.target, 7, sl.weight, weight_mode, wcols, rows, scale,
                         sl.loss, nullptr, grad_host ? sl.grad : nullptr, sl.ws, sl.ws_bytes,
                         GD_VARIANT_AUTO, 0, sl.stream);
    if (rc != 0) break;
    if (grad_host)
      e = cudaMemcpyAsync(grad_host + r0 * 7, sl.grad, (size_t)rows * kRowBytes,
                          cudaMemcpyDeviceToHost, sl.stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(p.loss_pinned + c, sl.loss, sizeof(float), cudaMemcpyDeviceToHost,
                          sl.stream);
    if (e != cudaSuccess) rc = (int)e;
  }
  for (int s = 0; s < kSlots; ++s) {
    const cudaError_t e = cudaStreamSynchronize(p.slot[s].stream);
    if (rc == 0 && e != cudaSuccess) rc = (int)e;
  }
  if (rc == 0) {
    double tot = 0.0;
    for (long long c = 0; c < nchunks; ++c) tot += (double)p.loss_pinned[c];
    *loss_host = (float)tot;
  }
  cudaSetDevice(prev_dev);
  return rc;
}
