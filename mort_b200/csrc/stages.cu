// stages.cu — the block wavefront compiled with world::hit's general medium order (rt_core.cuh, MORT_GENERAL_MEDIA): a
// constant_medium reached through translate / rotate_y / list wrappers (hitDispatch has the case, objects.cuh:875-877; no shipped
// scene does it) is one stage of the visit order — clipped against what was found before it, followed by a windowed closest-hit
// pass over the leaves behind it.  A separate unit, one block shape, so that the kernels every shipped scene runs stay as they are;
// mort_render_device launches it for scenes whose flattening found such a medium (FlatScene::two_pass == 2), whatever opts.mode says.
// Exports pool_query_stages / pool_launch_stages.
#define MORT_GENERAL_MEDIA 1
#include "pool.cu"
