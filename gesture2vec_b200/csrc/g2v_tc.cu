// placeholder: tensor-core search (filled in next)
#include "g2v_common.cuh"
namespace g2v {
bool tc_supported(int, int) { return false; }
size_t tc_workspace_bytes(int64_t, int, int, int) { return 0; }
int launch_search_tc(const void*, int, const float*, const void*, int64_t, int, int, int32_t*,
                     unsigned long long*, void*, size_t, unsigned, cudaStream_t) {
  set_error_detail("tensor-core search not built");
  return G2V_ERR_UNSUPPORTED;
}
}  // namespace g2v
