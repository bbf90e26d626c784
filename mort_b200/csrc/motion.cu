// motion.cu — the block wavefront and the parity hook compiled with motion-aware node boxes (rt_core.cuh, MORT_MOTION_BOUNDS):
// sc.nodes holds every child box at ray time 0, sc.node_dt its change to time 1, and a node test interpolates them at the ray's
// time.  A separate unit so that the kernels every other scene runs are not touched by it (registers, spills); launched only
// for scenes committed with mort_build_opts.motion_bounds.  Exports pool_query_motion / pool_launch_motion / trace_launch_motion.
#define MORT_MOTION_BOUNDS 1
#include "pool.cu"
