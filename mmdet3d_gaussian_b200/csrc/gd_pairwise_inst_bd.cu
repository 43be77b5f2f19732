// One pairwise instantiation per translation unit so the build parallelises.
#include "gd_pairwise.cuh"
namespace gdk {
template int launch_pairwise<gd::kBd>(const PairwiseArgs&, cudaStream_t);
template int launch_filter<gd::kBd>(const PairwiseArgs&, const FilterArgs&, cudaStream_t);
}  // namespace gdk
