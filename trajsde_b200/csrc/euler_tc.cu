// placeholder: tcgen05 forward lands here
#include "common.cuh"
namespace trajsde {
int64_t euler_fwd_tc_workspace_bytes(int64_t, int32_t, int32_t) { return 0; }
int launch_euler_fwd_tc(const TrajsdeEulerFwdArgs&, cudaStream_t) { return set_error(TRAJSDE_ERR_UNSUPPORTED, "TC mode not built yet"); }
}
