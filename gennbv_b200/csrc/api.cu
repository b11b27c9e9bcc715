// api.cu -- error reporting + ABI version of libgennbv_b200.
#include "common.cuh"

namespace gnbv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gnbv

extern "C" int gnbv_abi_version(void) { return GNBV_ABI_VERSION; }
extern "C" const char* gnbv_last_error(void) { return gnbv::g_err; }

// ---- optional stage timing: when enabled, the multi-kernel entry points (gnbv_encoder_forward / _backward) record a
// CUDA event on their stream at each stage boundary, so that bench.py can report the live duration of individual
// kernels inside the timed region (the roofline rule: "measured live ... on the stream the kernel is launched on").
namespace gnbv {
static cudaEvent_t g_ev[GNBV_MAX_STAGES];
static bool g_ev_created = false, g_ev_enabled = false;
static int g_ev_used[GNBV_MAX_STAGES];
void stage_mark(int id, cudaStream_t stream) {
    if (!g_ev_enabled || id < 0 || id >= GNBV_MAX_STAGES) return;
    cudaEventRecord(g_ev[id], stream);
    g_ev_used[id] = 1;
}
}  // namespace gnbv

extern "C" int gnbv_profile_enable(int on) {
    using namespace gnbv;
    if (on && !g_ev_created) {
        for (int i = 0; i < GNBV_MAX_STAGES; ++i) GNBV_CUDA_CHECK(cudaEventCreate(&g_ev[i]));
        g_ev_created = true;
    }
    if (on) for (int i = 0; i < GNBV_MAX_STAGES; ++i) g_ev_used[i] = 0;
    g_ev_enabled = on != 0;
    return GNBV_OK;
}

extern "C" int gnbv_profile_elapsed_ms(int stage_from, int stage_to, float* ms) {
    using namespace gnbv;
    GNBV_REQUIRE(ms && g_ev_created && stage_from >= 0 && stage_to < GNBV_MAX_STAGES && stage_from < stage_to,
                 "gnbv_profile_elapsed_ms: bad arguments");
    GNBV_REQUIRE(g_ev_used[stage_from] && g_ev_used[stage_to], "gnbv_profile_elapsed_ms: stage %d or %d was not recorded",
                 stage_from, stage_to);
    GNBV_CUDA_CHECK(cudaEventSynchronize(g_ev[stage_to]));
    GNBV_CUDA_CHECK(cudaEventElapsedTime(ms, g_ev[stage_from], g_ev[stage_to]));
    return GNBV_OK;
}
