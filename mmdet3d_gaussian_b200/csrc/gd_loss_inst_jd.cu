// Instantiates the fused kernels for one loss type (see gd_loss_kernels.cuh).
#include "gd_loss_kernels.cuh"

namespace gdk {
template int launch_loss<gd::kJd>(const LossArgs&, int, int, cudaStream_t);
}  // namespace gdk
