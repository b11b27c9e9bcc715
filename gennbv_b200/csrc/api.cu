// api.cu -- error reporting + ABI version of libgennbv_b200.
#include "common.cuh"

#include <cstdlib>
#include <mutex>

namespace gnbv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gnbv

// ---- kernel-variant switches.  Every variant is parity-tested (tests/test_policy_gpu.py re-runs the encoder parity tests
// under each setting); the defaults are the fastest measured on B200 at B = 256, 64^3 (profiles/).
//   GNBV_CONV2_TC  (bit mask) 2 = mma.sync forward, 4 = mma.sync data gradient, 8 = mma.sync weight gradient (register
//                  path), 16 = TMA-staged weight gradient (overrides 8); 32 = tcgen05 TS-form forward and 64 = tcgen05 TS-form
//                  data gradient (conv2_ts.cu: A operand in tensor memory, TMA bulk staging, warp-specialised; parity-green,
//                  measured 0.31 / 0.66 ms against 0.27 / 0.44 ms for the mma.sync kernels at B = 256 -- with N = 16 output
//                  channels a tcgen05.mma is bound by its A-operand fetch (16 cycles per M128 x N16 x K8 instruction,
//                  scripts/micro/mma_rate.cu) and the staging traffic saturates shared memory);
//                  1 = tcgen05 SS-form forward (conv2_tc.cu; slower still); 0 = fp32 CUDA-core kernels.
//   GNBV_CONV1_MMA (bit mask) 1 = conv1 forward on the tensor cores, 2 = conv1 weight gradient; 0 = CUDA-core TMA kernels.
//   GNBV_GEMM_MMA  (bit mask) 1 = mma.sync 3xTF32 GEMM for the Linear layers (0 = fp32 CUDA-core GEMM), 2 = the large GEMMs
//                  (grid Linear forward / dX / dW) on the tcgen05 warp-specialised pipeline of tc_gemm.cu.  Default 3.
namespace gnbv {
static int env_mode(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
int conv2_tc_mode() { static const int m = env_mode("GNBV_CONV2_TC", 30); return m; }
int conv1_mma_mode() { static const int m = env_mode("GNBV_CONV1_MMA", 3); return m; }
int gemm_mma_mode() { static const int m = env_mode("GNBV_GEMM_MMA", 3); return m; }
}  // namespace gnbv

namespace gnbv {
// (kernel, device) -> largest dynamic shared-memory size already granted
int ensure_dyn_smem_impl(const void* func, size_t bytes) {
    struct Entry { const void* f; int dev; size_t bytes; };
    static Entry table[256];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    GNBV_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    Entry* e = nullptr;
    for (int i = 0; i < used; ++i)
        if (table[i].f == func && table[i].dev == dev) { e = &table[i]; break; }
    if (e && e->bytes >= bytes) return GNBV_OK;
    GNBV_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (!e && used < 256) { e = &table[used++]; e->f = func; e->dev = dev; }
    if (e) e->bytes = bytes;
    return GNBV_OK;
}
}  // namespace gnbv

extern "C" int gnbv_kernel_mode(int which) {
    return which == 0 ? gnbv::conv2_tc_mode() : which == 1 ? gnbv::conv1_mma_mode() : which == 2 ? gnbv::gemm_mma_mode() : -1;
}

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
