// K2 placeholder while K1 is validated on hardware; replaced by the tcgen05 kernel.
#include "k2_batched.h"

#include "kernels.h"
namespace yrb {
struct K2State {};
K2State* k2_create() { return new K2State(); }
void k2_destroy(K2State* s) { delete s; }
void k2_invalidate(K2State*) {}
int k2_parts(int sm_count) { return sm_count; }
bool k2_supported(int, int, int) { return false; }
int k2_search(K2State*, const void*, int64_t, int64_t, int, int, const void*, int, int, const uint32_t*, int,
              const float*, const float*, uint64_t*, uint64_t*, int, cudaStream_t, int*, std::string& err) {
    err = "K2 not built";
    return -4;
}
}  // namespace yrb
