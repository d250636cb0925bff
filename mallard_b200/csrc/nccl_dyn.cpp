#include "nccl_dyn.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace mlb {

const NcclApi & nccl() {
    static NcclApi api{};
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        void * h = nullptr;
        const char * env = getenv("MLB_NCCL_LIB");
        static std::string loaded;
        if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); loaded = env; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); loaded = "libnccl.so.2 (already loaded in this process)"; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); loaded = "libnccl.so.2"; }
        if (!h) { error = std::string("mallard_b200: NCCL is not available (") + dlerror() + "); set MLB_NCCL_LIB to a libnccl.so.2"; return; }
        auto sym = [&](const char * name) -> void * {
            void * p = dlsym(h, name);
            if (!p && error.empty()) error = std::string("mallard_b200: NCCL symbol missing: ") + name;
            return p;
        };
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(sym("ncclCommSplit"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.path = loaded.c_str();
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}

}  // namespace mlb
