// Library runtime: error messages, backend switch, launch counter.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace vsw {

static thread_local char g_err[512] = {0};
static std::atomic<int> g_backend{VSW_GEMM_AUTO};
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return VSW_ERR_CUDA;
    }
    return VSW_OK;
}

int backend() { return g_backend.load(); }

}  // namespace vsw

extern "C" {

int vsw_version(void) { return 100; }  // 0.1.0

int vsw_last_error(char* buf, size_t n) {
    size_t len = strlen(vsw::g_err);
    if (buf && n) {
        size_t c = len < n - 1 ? len : n - 1;
        memcpy(buf, vsw::g_err, c);
        buf[c] = 0;
    }
    return (int)len;
}

int vsw_set_gemm_backend(int backend) {
    if (backend < VSW_GEMM_AUTO || backend > VSW_GEMM_TCGEN05) return VSW_ERR_ARG;
    return vsw::g_backend.exchange(backend);
}
int vsw_get_gemm_backend(void) { return vsw::g_backend.load(); }

long long vsw_launch_count(void) { return vsw::g_launches.load(); }

}  // extern "C"
