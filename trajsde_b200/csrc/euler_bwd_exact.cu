// placeholder: exact backward lands here
#include "common.cuh"
namespace trajsde {
int64_t euler_bwd_exact_workspace_bytes(int64_t, int32_t, int32_t) { return 0; }
int launch_euler_bwd_exact(const TrajsdeEulerBwdArgs&, cudaStream_t) { return set_error(TRAJSDE_ERR_UNSUPPORTED, "backward not built yet"); }
}
