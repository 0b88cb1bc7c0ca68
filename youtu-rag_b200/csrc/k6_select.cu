// K6 placeholder (radix select for k > 128).
#include "kernels.h"
namespace yrb {
size_t select_scratch_bytes(int64_t, int) { return 1; }
cudaError_t launch_select(const float*, int64_t, int, uint64_t*, void*, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace yrb
