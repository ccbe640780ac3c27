// In-memory Android system-property store for Linux builds (see shim/sys/system_properties.h).
#include <sys/system_properties.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace {
std::mutex g_mu;
std::map<std::string, std::string> &store() { static std::map<std::string, std::string> s; return s; }
}

extern "C" int __system_property_get(const char *name, char *value)
{
    if (!name || !value) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = store().find(name);
    std::string v;
    if (it != store().end()) v = it->second;
    else {
        std::string env = std::string("PROP_") + name;
        for (auto &c : env) if (c == '.') c = '_';
        if (const char *e = getenv(env.c_str())) v = e;
    }
    strncpy(value, v.c_str(), PROP_VALUE_MAX - 1);
    value[PROP_VALUE_MAX - 1] = '\0';
    return (int)strlen(value);
}

extern "C" int __system_property_set(const char *name, const char *value)
{
    if (!name || !value || strlen(value) >= PROP_VALUE_MAX) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    store()[name] = value;
    return 0;
}
