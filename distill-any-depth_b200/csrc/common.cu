#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.h"

namespace dad {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

const char* last_error() { return g_err; }

bool pdl_enabled() {
    static const bool on = getenv("DAD_NO_PDL") == nullptr;
    return on;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------ profiling / launch counting
static thread_local char g_label[128] = "";
void debug_label(const char* s) {
    static const bool dbg = getenv("DAD_DEBUG_SYNC") != nullptr || getenv("DAD_DEBUG_TIME") != nullptr;
    if (!dbg) return;
    snprintf(g_label, sizeof(g_label), "%s", s);
}
namespace {
struct ProfRec { int cls; double work; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_recs;
std::atomic<long long> g_launches{0};
bool g_prof = false;
std::mutex g_mu;
}  // namespace

static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
static bool debug_time() {
    static const bool on = getenv("DAD_DEBUG_TIME") != nullptr;
    return on;
}

ProfScope::ProfScope(int cls, double work, cudaStream_t stream, int launches) : idx(-1), st(stream) {
    g_launches += launches;
    if (debug_time()) {
        if (!g_t0) { cudaEventCreate(&g_t0); cudaEventCreate(&g_t1); }
        cudaEventRecord(g_t0, st);
    }
    if (!g_prof) return;
    std::lock_guard<std::mutex> lk(g_mu);
    ProfRec r{cls, work, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
    idx = static_cast<int>(g_recs.size()) - 1;
}
ProfScope::~ProfScope() {
    if (debug_time()) {  // per-launch device time with the current label (serialises the stream)
        cudaEventRecord(g_t1, st);
        cudaEventSynchronize(g_t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_t0, g_t1);
        fprintf(stderr, "dad[time] %8.3f ms  %s\n", ms, g_label);
    }
    static const bool dbg = getenv("DAD_DEBUG_SYNC") != nullptr;
    if (dbg) {  // bisecting aid: surface the first faulting launch with its label
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            fprintf(stderr, "dad[debug]: launch after label '%s' failed: %s\n", g_label, cudaGetErrorString(e));
            fflush(stderr);
        }
    }
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_recs[idx].e1, st);
}

void profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& r : g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_recs.clear();
    g_prof = on != 0;
}
int profile_get(int cls, double* ms, double* work, long long* launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    double t = 0, w = 0;
    long long n = 0;
    for (auto& r : g_recs) {
        if (r.cls != cls) continue;
        if (cudaEventSynchronize(r.e1) != cudaSuccess) return set_error(DAD_ERR_CUDA, "profile: event sync failed");
        float f = 0;
        if (cudaEventElapsedTime(&f, r.e0, r.e1) != cudaSuccess) return set_error(DAD_ERR_CUDA, "profile: elapsed failed");
        t += f; w += r.work; ++n;
    }
    if (ms) *ms = t;
    if (work) *work = w;
    if (launches) *launches = n;
    return DAD_OK;
}
long long launch_count() { return g_launches.load(); }

}  // namespace dad

extern "C" const char* dad_last_error() { return dad::last_error(); }

extern "C" void dad_profile_enable(int on) { dad::profile_enable(on); }
extern "C" int dad_profile_get(int cls, double* ms, double* work, long long* launches) {
    return dad::profile_get(cls, ms, work, launches);
}
extern "C" long long dad_launch_count() { return dad::launch_count(); }
