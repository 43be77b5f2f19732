// Torch C++ extension over the C ABI of libgdloss_b200.so (include/gd_loss_b200.h).
//
// This is the "thin shim" of the drop-in boundary: everything GDLoss.forward
// (mmdet3d_gaussian/models/losses/gaussian_distance_loss.py:280-310, "ref") does around the
// arithmetic -- argument checks, the [..., 7] -> [N, 7] view, weight-shape dispatch
// (ref:295-296), the weight_reduce_loss scalar folding (mmdet), the early return
// (ref:290-292), output allocation, the current stream, the autograd edge -- in C++, so a
// module call costs a few microseconds of host time instead of ~120 us of Python + ctypes.
// No arithmetic happens here: every number comes out of the CUDA library, which is
// dlopen()ed by path (so the tests can point it at the IEEE-math build) and called through
// the exported C symbols only.  There is no fallback: if the library is not bound, every
// entry point throws.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <c10/cuda/CUDAGuard.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include <torch/csrc/autograd/function.h>
#include <torch/csrc/autograd/functions/utils.h>
#include <torch/csrc/autograd/python_variable.h>
#include <torch/extension.h>

#include <cmath>
#include <mutex>
#include <unordered_map>

#include "../../include/gd_loss_b200.h"

namespace py = pybind11;
using at::Tensor;

namespace {

// ---------------------------------------------------------------------------
// the C ABI, resolved with dlsym
// ---------------------------------------------------------------------------
struct Abi {
  void* handle = nullptr;
  std::string path;
#define GD_SYM(name) decltype(&name) name = nullptr;
  GD_SYM(gd_abi_version)
  GD_SYM(gd_loss_workspace_bytes)
  GD_SYM(gd_loss_launch)
  GD_SYM(gd_scale_grad)
  GD_SYM(gd_scale_buffer)
  GD_SYM(gd_scale_grad_rows)
  GD_SYM(gd_host_flag_wait)
  GD_SYM(gd_count_positive_labels)
  GD_SYM(gd_peer_sum_buffer_bytes)
  GD_SYM(gd_anchor_decoded_loss_fwd_bwd)
  GD_SYM(gd_center_decoded_loss_fwd_bwd)
  GD_SYM(gd_error_string)
#undef GD_SYM
};
Abi g_abi;

void bind(const std::string& path) {
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error(std::string("gd_loss_b200: cannot load ") + path + ": " + dlerror());
  Abi a;
  a.handle = h;
  a.path = path;
#define GD_SYM(name)                                                                         \
  a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, #name));                              \
  if (!a.name) throw std::runtime_error("gd_loss_b200: " + path + " does not export " #name);
  GD_SYM(gd_abi_version)
  GD_SYM(gd_loss_workspace_bytes)
  GD_SYM(gd_loss_launch)
  GD_SYM(gd_scale_grad)
  GD_SYM(gd_scale_buffer)
  GD_SYM(gd_scale_grad_rows)
  GD_SYM(gd_host_flag_wait)
  GD_SYM(gd_count_positive_labels)
  GD_SYM(gd_peer_sum_buffer_bytes)
  GD_SYM(gd_anchor_decoded_loss_fwd_bwd)
  GD_SYM(gd_center_decoded_loss_fwd_bwd)
  GD_SYM(gd_error_string)
#undef GD_SYM
  if (a.gd_abi_version() != GD_ABI_VERSION)
    throw std::runtime_error("gd_loss_b200: ABI version mismatch in " + path);
  g_abi = a;
}

inline const Abi& gd_abi() {
  if (!g_abi.handle)
    throw std::runtime_error("gd_loss_b200: the CUDA library is not bound (no CPU fallback exists)");
  return g_abi;
}

inline void check(int code, const char* what) {
  if (code != 0)
    throw std::runtime_error(std::string("gd_loss_b200.") + what + " failed (" +
                             std::to_string(code) + "): " + gd_abi().gd_error_string(code));
}

[[noreturn]] void raise_not_implemented(const char* msg) {
  PyErr_SetString(PyExc_NotImplementedError, msg);
  throw py::error_already_set();
}

// ---------------------------------------------------------------------------
// per (device, stream) scratch: zero-initialised once, left zeroed by the kernels
// ---------------------------------------------------------------------------
struct StreamState {
  Tensor workspace;            // ticket + per-CTA partials
  Tensor probe_host;           // int32[kProbeSlots] pinned: any(weight > 0) words written by the launches
  unsigned probe_next = 0;     // guarded by g_mutex
};
constexpr int kProbeSlots = 16;
std::mutex g_mutex;
std::unordered_map<uint64_t, std::shared_ptr<StreamState>> g_state;
constexpr size_t kMaxStreamStates = 256;

std::shared_ptr<StreamState> stream_state(const c10::cuda::CUDAStream& s) {
  const uint64_t key = (static_cast<uint64_t>(s.device_index()) << 56) ^
                       reinterpret_cast<uint64_t>(s.stream());
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g_state.find(key);
  if (it != g_state.end()) return it->second;
  // bounded: a process that keeps creating streams starts over (callers hold their entry by
  // shared_ptr; the buffers belong to the stream-ordered caching allocator, so dropping an
  // entry whose kernels are still queued is safe)
  if (g_state.size() >= kMaxStreamStates) g_state.clear();
  auto st = std::make_shared<StreamState>();
  const auto nbytes = static_cast<int64_t>(gd_abi().gd_loss_workspace_bytes(0));
  st->workspace = at::zeros({nbytes}, at::TensorOptions().dtype(at::kByte).device(at::kCUDA, s.device_index()));
  g_state.emplace(key, st);
  return st;
}

struct Config {
  gd_loss_config c;
};

// NVTX range around every entry point (the reference inherits mmcv's
// profiler hooks; here a timeline tool sees "gd_loss_b200::<entry>" on the host thread).
// Header-only NVTX v3: a few nanoseconds when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

inline void require_cuda(const Tensor& t, const char* name) {
  if (!t.is_cuda())
    throw std::runtime_error(std::string("gd_loss_b200: ") + name + " is on " + t.device().str() +
                             "; this implementation is CUDA-only (sm_100a) and has no CPU fallback");
}

// [..., 7] -> [N, 7] fp32 with unit inner stride (row stride free).  Dispatched torch ops, so
// autograd tracks whatever reshaping / casting was needed.
inline Tensor rows7(const Tensor& t) {
  Tensor r = t.dim() == 2 ? t : t.reshape({-1, 7});
  if (r.scalar_type() != at::kFloat) r = r.to(at::kFloat);
  if (r.size(0) > 0 && r.stride(1) != 1) r = r.contiguous();
  return r;
}
inline int64_t row_stride(const Tensor& t) { return t.size(0) > 1 ? t.stride(0) : 7; }
inline const float* fptr(const Tensor& t) { return t.defined() ? t.const_data_ptr<float>() : nullptr; }
inline float* fptr_mut(const Tensor& t) { return t.defined() ? t.mutable_data_ptr<float>() : nullptr; }

// ---------------------------------------------------------------------------
// fused loss: launch + autograd node
// ---------------------------------------------------------------------------
struct LossCall {
  gd_loss_config cfg;
  Tensor pred, target, weight, scale_div;     // [N,7], [N,7], [N] | [N,7] | undefined
  int wmode = GD_WEIGHT_NONE;
  float scale = 1.0f;
  bool rows_out = false;
  int variant = GD_VARIANT_AUTO;
  int flags = 0;
  bool device_early_return = false;           // ref:290-292 decided on the device
  int64_t er_wrow = 0, er_wcol = 0;           // weight(i, c) strides for (pred * weight)
  gd_peer_sum peer = {};                      // world > 1: loss summed over the box in-kernel
};

struct LossOut {
  Tensor loss, rows, grad;
};

LossOut launch_loss(const LossCall& k, bool want_grad, int32_t* any_positive_host = nullptr) {
  const Abi& lib = gd_abi();
  const int64_t n = k.pred.size(0);
  const auto opts = k.pred.options();
  LossOut o;
  const bool want_sum = !k.rows_out;
  if (want_sum) o.loss = at::empty({}, opts);
  if (k.rows_out) o.rows = at::empty({n}, opts);
  if (want_grad) o.grad = at::empty({n, 7}, opts);
  const auto stream = at::cuda::getCurrentCUDAStream(k.pred.get_device());
  gd_loss_io io{};
  io.pred = fptr(k.pred);
  io.pred_row_stride = row_stride(k.pred);
  io.target = fptr(k.target);
  io.target_row_stride = row_stride(k.target);
  io.weight = fptr(k.weight);
  io.weight_mode = k.wmode;
  if (k.wmode == GD_WEIGHT_ROW) io.weight_row_stride = n > 1 ? k.weight.stride(0) : 1;
  if (k.wmode == GD_WEIGHT_ROW7) io.weight_row_stride = row_stride(k.weight);
  io.n = n;
  io.scale = k.scale;
  io.scale_div = fptr(k.scale_div);
  io.loss_sum = fptr_mut(o.loss);
  io.row_loss = fptr_mut(o.rows);
  io.grad_pred = fptr_mut(o.grad);
  if (k.device_early_return) {
    io.early_return = 1;
    io.er_weight_row_stride = k.er_wrow;
    io.er_weight_col_stride = k.er_wcol;
  }
  std::shared_ptr<StreamState> st;
  if (want_sum) {
    st = stream_state(stream);
    io.workspace = st->workspace.mutable_data_ptr();
    io.workspace_bytes = static_cast<size_t>(st->workspace.numel());
  }
  io.variant = k.variant;
  io.flags = k.flags;
  io.peer_sum = (k.peer.world > 1 && want_sum) ? &k.peer : nullptr;
  io.any_positive_host = any_positive_host;
  check(lib.gd_loss_launch(&k.cfg, &io, stream.stream()), "gd_loss_launch");
  return o;
}

// d loss / d pred was produced by the forward launch; backward folds grad_output in.
struct LossNode : public torch::autograd::Node {
  LossCall call;
  Tensor grad_buf;
  bool released = false;

  torch::autograd::variable_list apply(torch::autograd::variable_list&& grads) override {
    NvtxRange nvtx("gd_loss_b200::gd_loss.backward");
    torch::autograd::variable_list out(1);
    if (grads.empty() || !grads[0].defined() || !task_should_compute_output(0)) return out;
    if (released)
      throw std::runtime_error(
          "Trying to backward through the graph a second time (gd_loss_b200: the saved tensors "
          "of the fused loss were freed; pass retain_graph=True to the first backward)");
    const Abi& lib = gd_abi();
    c10::cuda::CUDAGuard guard(call.pred.device());
    Tensor grad = std::move(grad_buf);
    grad_buf = Tensor();
    if (!grad.defined())       // second backward (retain_graph=True): regenerate, one fused launch
      grad = launch_loss(call, true).grad;
    Tensor go = grads[0];
    if (go.scalar_type() != at::kFloat) go = go.to(at::kFloat);
    const auto stream = at::cuda::getCurrentCUDAStream(grad.get_device());
    const int64_t n = grad.size(0);
    if (call.rows_out) {
      go = go.reshape({-1});
      check(lib.gd_scale_grad_rows(fptr_mut(grad), n, fptr(go), n > 1 ? go.stride(0) : 1,
                                   stream.stream()),
            "gd_scale_grad_rows");
    } else {
      check(lib.gd_scale_grad(fptr_mut(grad), n, fptr(go), stream.stream()), "gd_scale_grad");
    }
    out[0] = std::move(grad);
    return out;
  }

  void release_variables() override {
    std::lock_guard<std::mutex> lock(mutex_);
    released = true;
    grad_buf = Tensor();
    call.pred = Tensor();
    call.target = Tensor();
    call.weight = Tensor();
    call.scale_div = Tensor();
  }
};

Tensor attach_history(Tensor out, const Tensor& pred, const LossCall& call, Tensor grad) {
  auto node = std::shared_ptr<LossNode>(new LossNode(), torch::autograd::deleteNode);
  node->call = call;
  node->grad_buf = std::move(grad);
  node->set_next_edges(torch::autograd::collect_next_edges(pred));
  torch::autograd::set_history(out, node);
  return out;
}

struct PeerSum {
  gd_peer_sum c;
};

enum Reduction { kNone = 0, kMean = 1, kSum = 2 };
enum SyncMode {
  kMaskZero = 0,   // host_sync=False: never probe, rows with weight exactly 0 masked in-kernel
  kExact = 1       // reference semantics of ref:290-292 (device-side where the shapes allow)
};

// GDLoss.forward after the Python-only steps (reduction_override assert, kwargs merge).
Tensor gd_loss(const Tensor& pred_in, const Tensor& target_in, const c10::optional<Tensor>& weight_in,
               const Config& cfg, double loss_weight, int reduction, const py::object& avg_factor,
               int variant, int sync_mode, const PeerSum* peer) {
  NvtxRange nvtx("gd_loss_b200::gd_loss");
  require_cuda(pred_in, "pred");
  require_cuda(target_in, "target");
  if (target_in.requires_grad())
    raise_not_implemented(
        "gd_loss_b200: gradients w.r.t. `target` are not produced (the reference call sites build "
        "targets without grad); detach the target");
  if (pred_in.dim() < 1 || pred_in.size(-1) != 7)
    throw py::value_error("pred must be [..., 7], got " + c10::str(pred_in.sizes()));
  const bool has_weight = weight_in.has_value() && weight_in->defined();
  LossCall k;
  k.cfg = cfg.c;
  k.variant = variant;
  k.pred = rows7(pred_in);
  k.target = rows7(target_in.requires_grad() ? target_in.detach() : target_in);
  if (k.pred.sizes() != k.target.sizes())
    throw py::value_error("pred " + c10::str(pred_in.sizes()) + " and target " +
                          c10::str(target_in.sizes()) + " differ");
  const int64_t n = k.pred.size(0);
  c10::cuda::CUDAGuard guard(k.pred.device());

  if (has_weight) {
    require_cuda(*weight_in, "weight");
    Tensor w = weight_in->requires_grad() ? weight_in->detach() : *weight_in;
    if (w.scalar_type() != at::kFloat) w = w.to(at::kFloat);
    const auto psz = pred_in.sizes();
    if (w.sizes() == psz) {                                                   // ref:295-296
      k.wmode = GD_WEIGHT_ROW7;
      k.weight = w.dim() == 2 ? w : w.reshape({-1, 7});
      if (n > 0 && k.weight.stride(1) != 1) k.weight = k.weight.contiguous();
    } else if (w.numel() == n && w.sizes() == psz.slice(0, psz.size() - 1)) {
      k.wmode = GD_WEIGHT_ROW;
      k.weight = w.dim() == 1 ? w : w.reshape({-1});
    } else {
      throw py::value_error("weight shape " + c10::str(w.sizes()) + " must be " + c10::str(psz) +
                            " or " + c10::str(psz.slice(0, psz.size() - 1)));
    }
  }

  // mmdet weight_reduce_loss folded into one scalar
  const bool have_avg = !avg_factor.is_none();
  if (!have_avg) {
    if (reduction == kMean) k.scale = n > 0 ? static_cast<float>(loss_weight / static_cast<double>(n)) : NAN;
    else k.scale = static_cast<float>(loss_weight);
    k.rows_out = reduction == kNone;
  } else {
    if (reduction == kSum) throw py::value_error("avg_factor can not be used with reduction=\"sum\"");
    k.rows_out = reduction == kNone;
    if (THPVariable_Check(avg_factor.ptr()) && reduction == kMean) {
      // avg_factor living on the device: divided in-kernel, no .item()
      Tensor a = THPVariable_Unpack(avg_factor.ptr());
      if (a.numel() != 1) throw py::value_error("a tensor avg_factor must have one element");
      if (a.is_cuda()) {
        if (a.requires_grad()) a = a.detach();
        if (a.scalar_type() != at::kFloat) a = a.to(at::kFloat);
        k.scale_div = a;
        k.scale = static_cast<float>(loss_weight);
      } else {
        k.scale = static_cast<float>(loss_weight / a.item<double>());
      }
    } else if (reduction == kMean) {
      k.scale = static_cast<float>(loss_weight / avg_factor.cast<double>());
    } else {
      k.scale = static_cast<float>(loss_weight);
    }
  }

  // early return (ref:290-292): weight given, reduction != 'none', no weight element > 0
  // -> (pred * weight).sum().  The decision is taken on the DEVICE whenever that expression
  // is shape-valid with the shape of pred ([N,7] weights -- the KITTI call pattern --, or a
  // [N] weight that broadcasts against [N,7]: N == 7 / N == 1).  For other [N] weights the
  // reference raises a broadcasting error when the branch is taken; to raise as well the
  // host has to know, which costs a wait for a 4-byte probe -- but the fused launch is
  // queued BEFORE that wait, so the GPU never idles.
  bool host_probe = false;
  if (sync_mode == kExact && has_weight && !k.rows_out) {
    if (k.wmode == GD_WEIGHT_ROW7) {
      k.device_early_return = true;
      k.er_wrow = row_stride(k.weight);
      k.er_wcol = 1;
    } else if (pred_in.dim() == 2 && (n == 7 || n == 1)) {
      k.device_early_return = true;            // [7] against [7,7]: broadcast over columns
      k.er_wrow = 0;
      k.er_wcol = n == 7 ? k.weight.stride(0) : 0;
    } else {
      host_probe = true;
    }
    // any(weight > 0) of an empty weight is False without looking: the branch is taken
    if (n == 0) return (pred_in * *weight_in).sum();
  }
  if (sync_mode == kMaskZero) k.flags |= GD_FLAG_MASK_ZERO_WEIGHT;
  if (peer != nullptr && peer->c.world > 1 && !k.rows_out) {
    // the early-return branch is a per-process decision in the reference (one process per
    // GPU); its replacement value would bypass the exchange, so the fused cross-GPU sum is
    // offered for the sync-free semantics only
    if (k.device_early_return || host_probe)
      throw py::value_error(
          "gd_loss_b200: the in-kernel cross-GPU sum needs weight=None or GDLoss(host_sync=False)");
    k.peer = peer->c;
  }

  const bool need_grad = at::GradMode::is_enabled() && k.pred.requires_grad();
  int32_t* probe_host = nullptr;
  if (host_probe) {
    if (c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None)
      throw std::runtime_error(
          "gd_loss_b200: GDLoss with [N] weights keeps the reference's early-return check "
          "(gaussian_distance_loss.py:290), whose outcome (an exception) needs a host wait and "
          "cannot be captured in a CUDA graph; pass [N,7] weights, weight=None or host_sync=False");
    const auto stream = at::cuda::getCurrentCUDAStream(k.pred.get_device());
    std::shared_ptr<StreamState> st = stream_state(stream);
    {
      std::lock_guard<std::mutex> lock(g_mutex);
      if (!st->probe_host.defined())
        st->probe_host = at::zeros({kProbeSlots}, at::TensorOptions().dtype(at::kInt).pinned_memory(true));
      probe_host = st->probe_host.mutable_data_ptr<int32_t>() + (st->probe_next++ % kProbeSlots);
    }
    __atomic_store_n(probe_host, 0, __ATOMIC_RELEASE);
  }

  // any(weight > 0) comes back from INSIDE the fused launch (its first warp reports a positive
  // weight of the first tile microseconds after the kernel starts): no probe launch, and the
  // GPU is already working on the loss while the host waits for the word.
  LossOut o = launch_loss(k, need_grad, probe_host);

  if (host_probe) {
    int code;
    {
      py::gil_scoped_release nogil;
      code = gd_abi().gd_host_flag_wait(
          probe_host, at::cuda::getCurrentCUDAStream(k.pred.get_device()).stream());
    }
    check(code, "gd_host_flag_wait");
    if (__atomic_load_n(probe_host, __ATOMIC_ACQUIRE) == 2)
      return (pred_in * *weight_in).sum();      // ref:292 (raises like the reference)
  }

  Tensor out = k.rows_out ? o.rows : o.loss;
  if (need_grad) out = attach_history(out, k.pred, k, o.grad);
  const auto in_dtype = pred_in.scalar_type();
  if (in_dtype != at::kFloat && at::isFloatingType(in_dtype)) out = out.to(in_dtype);
  return out;
}

// ---------------------------------------------------------------------------
// head front ends (f1 / f4): one launch, gradient w.r.t. the raw head outputs
// ---------------------------------------------------------------------------
struct BufferNode : public torch::autograd::Node {
  Tensor grad_buf;
  torch::autograd::variable_list apply(torch::autograd::variable_list&& grads) override {
    torch::autograd::variable_list out(1);
    if (grads.empty() || !grads[0].defined() || !task_should_compute_output(0)) return out;
    if (!grad_buf.defined())
      throw std::runtime_error(
          "gd_loss_b200: the fused head loss supports one backward pass per forward "
          "(retain_graph re-use is not supported)");
    Tensor grad = std::move(grad_buf);
    grad_buf = Tensor();
    c10::cuda::CUDAGuard guard(grad.device());
    Tensor go = grads[0];
    if (go.scalar_type() != at::kFloat) go = go.to(at::kFloat);
    const auto stream = at::cuda::getCurrentCUDAStream(grad.get_device());
    check(gd_abi().gd_scale_buffer(fptr_mut(grad), grad.numel(), fptr(go), stream.stream()),
          "gd_scale_buffer");
    out[0] = std::move(grad);
    return out;
  }
  void release_variables() override {
    std::lock_guard<std::mutex> lock(mutex_);
    grad_buf = Tensor();
  }
};

Tensor attach_buffer(Tensor out, const Tensor& input, Tensor grad) {
  auto node = std::shared_ptr<BufferNode>(new BufferNode(), torch::autograd::deleteNode);
  node->grad_buf = std::move(grad);
  node->set_next_edges(torch::autograd::collect_next_edges(input));
  torch::autograd::set_history(out, node);
  return out;
}

inline Tensor rows_f32(const Tensor& t, const char* name, int64_t min_cols) {
  require_cuda(t, name);
  if (t.dim() != 2 || t.size(1) < min_cols)
    throw py::value_error(std::string(name) + " must be [K,>=" + std::to_string(min_cols) + "], got " +
                          c10::str(t.sizes()));
  Tensor r = t.scalar_type() == at::kFloat ? t : t.to(at::kFloat);
  if (r.size(0) > 0 && r.stride(1) != 1) r = r.contiguous();
  return r;
}
inline int64_t stride0(const Tensor& t) { return t.size(0) > 1 ? t.stride(0) : t.size(1); }

// scale (host) and scale_div (device) from loss_weight / avg_factor for the reduced head losses
void head_scale(double loss_weight, const py::object& avg_factor, float* scale, Tensor* scale_div) {
  if (THPVariable_Check(avg_factor.ptr())) {
    Tensor a = THPVariable_Unpack(avg_factor.ptr());
    if (a.numel() != 1) throw py::value_error("a tensor avg_factor must have one element");
    if (a.is_cuda()) {
      if (a.requires_grad()) a = a.detach();
      if (a.scalar_type() != at::kFloat) a = a.to(at::kFloat);
      *scale_div = a;
      *scale = static_cast<float>(loss_weight);
      return;
    }
    *scale = static_cast<float>(loss_weight / a.item<double>());
    return;
  }
  *scale = static_cast<float>(loss_weight / avg_factor.cast<double>());
}

// GD branch of GDAnchor3DHead.loss_single (gd_anchor3d_head.py:102-141).
// scale_mode: 0 = `scale` is final; 1 = divide by avg_factor (number or device tensor);
// 2 = labels mode, reduction='mean' without avg_factor: divide by max(#positives, 1) counted
// on the device (the reference's `loss.mean()` over the positives, 0 when there are none).
Tensor anchor_decoded_loss(const Tensor& anchors_in, const Tensor& bbox_pred, const Tensor& bbox_targets,
                           const c10::optional<Tensor>& bbox_weights,
                           const c10::optional<std::vector<double>>& decode_weight,
                           const c10::optional<Tensor>& pos_inds_in, const c10::optional<Tensor>& labels_in,
                           int64_t num_classes, const Config& cfg, double loss_weight, int scale_mode,
                           const py::object& avg_factor, bool mask_zero_weight) {
  NvtxRange nvtx("gd_loss_b200::anchor_decoded_loss");
  const Abi& lib = gd_abi();
  const bool index_mode = pos_inds_in.has_value() && pos_inds_in->defined();
  const bool label_mode = labels_in.has_value() && labels_in->defined();
  if (index_mode == label_mode) throw py::value_error("pass exactly one of pos_inds / labels");
  Tensor anchors = rows_f32(anchors_in.requires_grad() ? anchors_in.detach() : anchors_in, "anchors", 7);
  if (anchors.size(1) != 7 || !anchors.is_contiguous()) anchors = anchors.slice(1, 0, 7).contiguous();
  Tensor bp = rows_f32(bbox_pred, "bbox_pred", 7);
  Tensor bt = rows_f32(bbox_targets.requires_grad() ? bbox_targets.detach() : bbox_targets, "bbox_targets", 7);
  if (bp.size(0) != bt.size(0)) throw py::value_error("bbox_pred and bbox_targets row counts differ");
  if (anchors.size(0) == 0) throw py::value_error("anchors must not be empty");
  const int64_t total = bp.size(0);
  c10::cuda::CUDAGuard guard(bp.device());
  Tensor bw;
  float dw[7];
  const bool weighted = decode_weight.has_value() && bbox_weights.has_value() && bbox_weights->defined();
  if (weighted) {
    bw = rows_f32(bbox_weights->requires_grad() ? bbox_weights->detach() : *bbox_weights, "bbox_weights", 7);
    if (bw.size(0) != total) throw py::value_error("bbox_weights row count differs from bbox_pred");
    if (decode_weight->size() != 7) throw py::value_error("decode_weight must be a scalar or 7 values");
    for (int c = 0; c < 7; ++c) dw[c] = static_cast<float>((*decode_weight)[c]);
  }
  Tensor pos, labels;
  if (index_mode) {
    require_cuda(*pos_inds_in, "pos_inds");
    pos = pos_inds_in->reshape({-1});
    if (pos.scalar_type() != at::kLong) pos = pos.to(at::kLong);
    if (!pos.is_contiguous()) pos = pos.contiguous();
  } else {
    require_cuda(*labels_in, "labels");
    labels = labels_in->reshape({-1});
    if (labels.scalar_type() != at::kLong) labels = labels.to(at::kLong);
    if (!labels.is_contiguous()) labels = labels.contiguous();
    if (labels.numel() != total) throw py::value_error("labels must have one entry per bbox_pred row");
  }
  const auto stream = at::cuda::getCurrentCUDAStream(bp.get_device());
  const auto st_ptr = stream_state(stream);
  StreamState& st = *st_ptr;
  float scale = static_cast<float>(loss_weight);
  Tensor scale_div;
  if (scale_mode == 1) {
    head_scale(loss_weight, avg_factor, &scale, &scale_div);
  } else if (scale_mode == 2) {
    if (!label_mode) throw py::value_error("scale_mode 2 is the labels mode");
    scale_div = at::empty({1}, bp.options());
    check(lib.gd_count_positive_labels(labels.const_data_ptr<int64_t>(), total, num_classes,
                                       fptr_mut(scale_div), st.workspace.mutable_data_ptr(),
                                       static_cast<size_t>(st.workspace.numel()), stream.stream()),
          "gd_count_positive_labels");
  }
  const bool need_grad = at::GradMode::is_enabled() && bp.requires_grad();
  Tensor loss = at::empty({}, bp.options());
  Tensor grad;
  int mode = GD_GRAD_NONE;
  if (need_grad) {
    if (label_mode) {
      grad = at::empty({total, 7}, bp.options());
      mode = GD_GRAD_DENSE;
    } else {
      grad = at::zeros({total, 7}, bp.options());
      mode = GD_GRAD_SCATTER;
    }
  }
  check(lib.gd_anchor_decoded_loss_fwd_bwd(
            &cfg.c, fptr(anchors), anchors.size(0), fptr(bp), stride0(bp), fptr(bt), stride0(bt),
            fptr(bw), weighted ? stride0(bw) : 7, weighted ? dw : nullptr,
            index_mode ? pos.const_data_ptr<int64_t>() : nullptr, index_mode ? pos.numel() : 0,
            label_mode ? labels.const_data_ptr<int64_t>() : nullptr, num_classes, total, scale,
            fptr(scale_div), fptr_mut(loss), fptr_mut(grad), mode, st.workspace.mutable_data_ptr(),
            static_cast<size_t>(st.workspace.numel()), mask_zero_weight ? GD_FLAG_MASK_ZERO_WEIGHT : 0,
            stream.stream()),
        "gd_anchor_decoded_loss_fwd_bwd");
  if (need_grad) loss = attach_buffer(loss, bp, grad);
  return loss;
}

struct CenterCoder {
  gd_center_coder c;
};

// GD branch of CenterGDHead.loss (gd_centerpoint_head.py:413-434).
Tensor center_decoded_loss(const Tensor& pred_in, const Tensor& pos_ind, const Tensor& target_box,
                           const c10::optional<Tensor>& weight_in, const CenterCoder& coder,
                           const Config& cfg, double loss_weight, int scale_mode,
                           const py::object& avg_factor, bool mask_zero_weight) {
  NvtxRange nvtx("gd_loss_b200::center_decoded_loss");
  const Abi& lib = gd_abi();
  Tensor p = rows_f32(pred_in, "pred", 7);
  Tensor t = rows_f32(target_box.requires_grad() ? target_box.detach() : target_box, "target_box", 7);
  require_cuda(pos_ind, "pos_ind");
  const int64_t n = p.size(0), cols = p.size(1);
  if (pos_ind.dim() != 2 || pos_ind.size(1) != 3 || pos_ind.size(0) != n)
    throw py::value_error("pos_ind must be [" + std::to_string(n) + ",3] (batch, x, y)");
  if (t.size(0) != n) throw py::value_error("pred and target_box row counts differ");
  c10::cuda::CUDAGuard guard(p.device());
  Tensor locs = pos_ind.scalar_type() == at::kLong ? pos_ind : pos_ind.to(at::kLong);
  locs = locs.slice(1, 1, 3);                       // (x_ind, y_ind), a strided view
  if (locs.stride(1) != 1) locs = locs.contiguous();
  int wmode = GD_WEIGHT_NONE;
  Tensor w;
  int64_t wstride = 0;
  if (weight_in.has_value() && weight_in->defined()) {
    require_cuda(*weight_in, "weight");
    w = weight_in->requires_grad() ? weight_in->detach() : *weight_in;
    if (w.scalar_type() != at::kFloat) w = w.to(at::kFloat);
    if (w.dim() == 2 && w.size(0) == n && w.size(1) == 7) {
      wmode = GD_WEIGHT_ROW7;
      if (w.stride(1) != 1) w = w.contiguous();
      wstride = stride0(w);
    } else if (w.numel() == n) {
      wmode = GD_WEIGHT_ROW;
      w = w.reshape({-1});
      wstride = n > 1 ? w.stride(0) : 1;
    } else {
      throw py::value_error("weight must be [P] or [P,7]");
    }
  }
  float scale = static_cast<float>(loss_weight);
  Tensor scale_div;
  if (scale_mode == 1) head_scale(loss_weight, avg_factor, &scale, &scale_div);
  else if (scale_mode == 2) scale = n > 0 ? static_cast<float>(loss_weight / static_cast<double>(n)) : NAN;
  const auto stream = at::cuda::getCurrentCUDAStream(p.get_device());
  const auto st_ptr = stream_state(stream);
  StreamState& st = *st_ptr;
  const bool need_grad = at::GradMode::is_enabled() && p.requires_grad();
  Tensor loss = at::empty({}, p.options());
  Tensor grad;
  if (need_grad) grad = at::empty({n, cols}, p.options());
  check(lib.gd_center_decoded_loss_fwd_bwd(
            &cfg.c, &coder.c, fptr(p), stride0(p), locs.const_data_ptr<int64_t>(),
            n > 1 ? locs.stride(0) : 2, fptr(t), stride0(t), fptr(w), wmode, wstride, n, scale,
            fptr(scale_div), fptr_mut(loss), fptr_mut(grad), cols, static_cast<int32_t>(cols),
            st.workspace.mutable_data_ptr(), static_cast<size_t>(st.workspace.numel()),
            mask_zero_weight ? GD_FLAG_MASK_ZERO_WEIGHT : 0, stream.stream()),
        "gd_center_decoded_loss_fwd_bwd");
  if (need_grad) loss = attach_buffer(loss, p, grad);
  return loss;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torch C++ shim over the C ABI of libgdloss_b200.so";
  m.def("bind", &bind, "dlopen the CUDA library and resolve the C ABI");
  m.def("bound_path", []() { return g_abi.path; });
  py::class_<Config>(m, "LossConfig")
      .def(py::init([](int loss_type, int fun, bool flag, double tau, double alpha,
                       const std::vector<double>& off) {
        if (off.size() != 3) throw py::value_error("center_offset must have 3 values");
        Config c{};
        c.c.loss_type = loss_type;
        c.c.fun = fun;
        c.c.flag = flag ? 1 : 0;
        c.c.tau = static_cast<float>(tau);
        c.c.alpha = static_cast<float>(alpha);
        for (int i = 0; i < 3; ++i) c.c.center_offset[i] = static_cast<float>(off[i]);
        return c;
      }));
  py::class_<CenterCoder>(m, "CenterCoder")
      .def(py::init([](const std::vector<double>& pc_range, int out_size_factor,
                       const std::vector<double>& voxel_size, bool norm_bbox) {
        if (pc_range.size() < 2 || voxel_size.size() < 2)
          throw py::value_error("pc_range / voxel_size need at least 2 values");
        CenterCoder c{};
        for (int i = 0; i < 2; ++i) {
          c.c.pc_range[i] = pc_range[i];
          c.c.voxel_size[i] = voxel_size[i];
        }
        c.c.out_size_factor = out_size_factor;
        c.c.norm_bbox = norm_bbox ? 1 : 0;
        return c;
      }))
      .def_property_readonly("norm_bbox", [](const CenterCoder& c) { return c.c.norm_bbox; })
      .def_property_readonly("out_size_factor", [](const CenterCoder& c) { return c.c.out_size_factor; });
  py::class_<PeerSum>(m, "PeerSum")
      .def(py::init([](int world, int rank, const std::vector<uint64_t>& bufs) {
        if (world < 1 || world > GD_MAX_PEERS || rank < 0 || rank >= world ||
            static_cast<int>(bufs.size()) != world)
          throw py::value_error("PeerSum(world, rank, one buffer pointer per rank)");
        PeerSum p{};
        p.c.world = world;
        p.c.rank = rank;
        for (int r = 0; r < world; ++r) p.c.peer_buf[r] = reinterpret_cast<void*>(bufs[r]);
        return p;
      }))
      .def_property_readonly("world", [](const PeerSum& p) { return p.c.world; })
      .def_property_readonly("rank", [](const PeerSum& p) { return p.c.rank; });
  m.def("peer_sum_buffer_bytes", []() { return gd_abi().gd_peer_sum_buffer_bytes(); });
  m.def("gd_loss", &gd_loss, py::arg("pred"), py::arg("target"), py::arg("weight"), py::arg("cfg"),
        py::arg("loss_weight"), py::arg("reduction"), py::arg("avg_factor"), py::arg("variant"),
        py::arg("sync_mode"), py::arg("peer_sum") = nullptr);
  m.def("anchor_decoded_loss", &anchor_decoded_loss);
  m.def("center_decoded_loss", &center_decoded_loss);
  m.def("stream_states", []() {
    std::lock_guard<std::mutex> lock(g_mutex);
    return g_state.size();
  });
}
