/* TEST INFRASTRUCTURE.  The reference's main.cpp references ispc::trace (src/main.cpp:619-624), which is generated
 * by the ISPC compiler from src/ispc/trace.ispc; there is no ispc in this image, so the drop-in proof binary
 * (oracle/_ref/ESCViewer2021_cuda) links this stand-in.  --ispc therefore aborts; --cuda and the serial path work. */
#include <cstdio>
#include <cstdlib>

#include "trace_ispc.h"

namespace ispc {
extern "C" void trace(int32_t, int32_t, struct ispc_cam &, int32_t, struct ispc_triangle *, int32_t, struct ispc_light *, int32_t,
                      struct ispc_triangle *, float *, int32_t, int32_t) {
    std::fprintf(stderr, "--ispc: not available (no ispc toolchain in this image)\n");
    std::abort();
}
}  // namespace ispc
