// TEST INFRASTRUCTURE (see cuda_emul.h).
#include "cuda_emul.h"

namespace pmp {
// the shared-memory arena of the one CTA that runs at a time (1 MB); the kernels' block-scope
// `extern float4 sm[]` declarations name this variable of the enclosing namespace
float4 sm[1 << 16];
}  // namespace pmp
