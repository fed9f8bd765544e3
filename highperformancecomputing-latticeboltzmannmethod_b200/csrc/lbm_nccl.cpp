// lbm_nccl.cpp -- run-time binding of NCCL (see lbm_nccl.h).
#include "lbm_nccl.h"

#include <dlfcn.h>

#include <mutex>

namespace lbm {

namespace {
NcclApi g_api;
std::once_flag g_once;

template <class F>
bool bind(void* lib, const char* name, F& slot) {
    slot = reinterpret_cast<F>(dlsym(lib, name));
    return slot != nullptr;
}

void load() {
    // RTLD_NOLOAD first: reuse the copy already mapped into the process (torch ships its own
    // libnccl.so.2 under the same SONAME), then fall back to the system library.
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        g_api.why = "libnccl.so.2 not found";
        return;
    }
    bool ok = true;
    ok &= bind(lib, "ncclGetUniqueId", g_api.GetUniqueId);
    ok &= bind(lib, "ncclCommInitRank", g_api.CommInitRank);
    ok &= bind(lib, "ncclCommDestroy", g_api.CommDestroy);
    ok &= bind(lib, "ncclSend", g_api.Send);
    ok &= bind(lib, "ncclRecv", g_api.Recv);
    ok &= bind(lib, "ncclGroupStart", g_api.GroupStart);
    ok &= bind(lib, "ncclGroupEnd", g_api.GroupEnd);
    ok &= bind(lib, "ncclAllReduce", g_api.AllReduce);
    ok &= bind(lib, "ncclGetErrorString", g_api.GetErrorString);
    ok &= bind(lib, "ncclGetVersion", g_api.GetVersion);
    g_api.ok = ok;
    if (!ok) g_api.why = "libnccl.so.2 lacks a required symbol";
}
}  // namespace

const NcclApi& nccl_api() {
    std::call_once(g_once, load);
    return g_api;
}

}  // namespace lbm
