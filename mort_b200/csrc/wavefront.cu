// wavefront.cu — wavefront variant of the shading stage (placeholder until the queue kernels land).
#include "render.hpp"
namespace mort {
struct WavefrontBuffers { int n; };
cudaError_t wavefront_alloc(WavefrontBuffers** out, int) { *out = nullptr; return cudaErrorNotSupported; }
void wavefront_free(WavefrontBuffers*) {}
size_t wavefront_bytes(int) { return 0; }
cudaError_t wavefront_render(const FrameParams&, WavefrontBuffers*, int, int, cudaStream_t, uint64_t*) { return cudaErrorNotSupported; }
}
