// One pairwise instantiation per translation unit so the build parallelises.
#include "gd_pairwise.cuh"
namespace gdk {
template int launch_pairwise<gd::kKld>(const PairwiseArgs&, cudaStream_t);
template int launch_filter<gd::kKld>(const PairwiseArgs&, const FilterArgs&, cudaStream_t);
}  // namespace gdk
