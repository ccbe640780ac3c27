#include "MediaLog.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

namespace {
std::atomic<MediaLogCallbackFunc> g_cb{ nullptr };
int min_level()
{
    static int lv = [] { const char *e = getenv("B200_LOG_LEVEL"); return e ? atoi(e) : (int)LOG_LEVEL_WARN; }();
    return lv;
}
}
void SetMediaLogCallback(MediaLogCallbackFunc cb) { g_cb.store(cb); }
void MediaLogPrint(int level, const char *tag, const char *fmt, ...)
{
    MediaLogCallbackFunc cb = g_cb.load();
    if (!cb && level < min_level()) return;
    char msg[512];
    va_list ap; va_start(ap, fmt); vsnprintf(msg, sizeof msg, fmt, ap); va_end(ap);
    char full[64]; snprintf(full, sizeof full, "Media_%s", tag);
    if (cb) cb(level, full, msg);
    else fprintf(stderr, "[%d] %s: %s\n", level, full, msg);
}
