// Test and measurement switches, kept off the production call path.
//
// Which kernel serves a call is decided from the shape alone.  The switches below force a path so that the
// parity tests can run each kernel on shapes another one normally takes and so that profiles/ can compare
// variants.  They are set through the test-only entry point maxpro_debug_set(); MAXPRO_* environment variables
// are read ONCE, when the library first looks at a switch, never per call -- an environment that changes
// later cannot change which kernel serves a call or how large its workspace is.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "kernels.h"

extern char** environ;

namespace mp {

namespace {
std::mutex g_cfg_mutex;
std::map<std::string, std::string> g_cfg;
bool g_cfg_loaded = false;

void load_env_once() {
    if (g_cfg_loaded) return;
    g_cfg_loaded = true;
    for (char** e = environ; e && *e; ++e) {
        if (strncmp(*e, "MAXPRO_", 7) != 0) continue;
        const char* eq = strchr(*e, '=');
        if (!eq) continue;
        g_cfg.emplace(std::string(*e, eq - *e), std::string(eq + 1));
    }
}
}  // namespace

const char* cfg(const char* name) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    load_env_once();
    auto it = g_cfg.find(name);
    return it == g_cfg.end() ? nullptr : it->second.c_str();
}

static std::atomic<int> g_stream_version{2};
int stream_version() { return g_stream_version.load(std::memory_order_relaxed); }
void set_stream_version(int v) { g_stream_version.store(v, std::memory_order_relaxed); }

void cfg_set(const char* name, const char* value) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    load_env_once();
    if (value) g_cfg[name] = value;
    else g_cfg.erase(name);
}

}  // namespace mp
