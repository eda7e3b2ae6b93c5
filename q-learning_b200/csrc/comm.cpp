// comm.cpp — see comm.h. Plain C++ (host compiler only).
#include "comm.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace qlc_comm {

namespace {
Api g_api;
bool g_ok = false;
std::string g_why;
std::once_flag g_once;

void load() {
    const char* names[] = {getenv("QLC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);      // an already loaded libnccl.so.2 (same SONAME) is reused, not loaded twice
        if (h) { g_api.path = n; break; }
        g_why = dlerror();
    }
    if (!h) { g_why = "cannot load NCCL (" + g_why + "); set QLC_NCCL_LIB to the full path of libnccl.so.2"; return; }
    struct { const char* name; void** slot; } syms[] = {
        {"ncclGetVersion", (void**)&g_api.GetVersion},     {"ncclGetUniqueId", (void**)&g_api.GetUniqueId},
        {"ncclCommInitRank", (void**)&g_api.CommInitRank}, {"ncclCommDestroy", (void**)&g_api.CommDestroy},
        {"ncclCommCount", (void**)&g_api.CommCount},       {"ncclAllGather", (void**)&g_api.AllGather},
        {"ncclAllReduce", (void**)&g_api.AllReduce},       {"ncclGetErrorString", (void**)&g_api.GetErrorString},
    };
    for (auto& s : syms) {
        *s.slot = dlsym(h, s.name);
        if (!*s.slot) { g_why = std::string("NCCL symbol missing: ") + s.name; return; }
    }
    g_api.CommInitRankConfig = (int (*)(Comm*, int, UniqueId, int, void*))dlsym(h, "ncclCommInitRankConfig");   // optional
    g_ok = true;
}
}  // namespace

const Api* api(std::string* why) {
    std::call_once(g_once, load);
    if (!g_ok) { if (why) *why = g_why; return nullptr; }
    return &g_api;
}

}  // namespace qlc_comm
