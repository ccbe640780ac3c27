// Logging macros with the names the reference's adapters use (common/log/MediaLog.h:41-47): level filter,
// "Media_<LOG_TAG>" prefix (common/log/MediaLogManager.cpp:78-82), pluggable callback.
#ifndef B200_MEDIA_LOG_H
#define B200_MEDIA_LOG_H
#ifndef LOG_TAG
#define LOG_TAG "Media"
#endif
enum { LOG_LEVEL_DEBUG = 3, LOG_LEVEL_INFO = 4, LOG_LEVEL_WARN = 5, LOG_LEVEL_ERROR = 6, LOG_LEVEL_FATAL = 7 };
typedef void (*MediaLogCallbackFunc)(int level, const char *tag, const char *msg);
void SetMediaLogCallback(MediaLogCallbackFunc cb);
void MediaLogPrint(int level, const char *tag, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
#define DBG(fmt, ...) MediaLogPrint(LOG_LEVEL_DEBUG, LOG_TAG, fmt, ##__VA_ARGS__)
#define INFO(fmt, ...) MediaLogPrint(LOG_LEVEL_INFO, LOG_TAG, fmt, ##__VA_ARGS__)
#define WARN(fmt, ...) MediaLogPrint(LOG_LEVEL_WARN, LOG_TAG, fmt, ##__VA_ARGS__)
#define ERR(fmt, ...) MediaLogPrint(LOG_LEVEL_ERROR, LOG_TAG, fmt, ##__VA_ARGS__)
#define FATAL(fmt, ...) MediaLogPrint(LOG_LEVEL_FATAL, LOG_TAG, fmt, ##__VA_ARGS__)
#endif
