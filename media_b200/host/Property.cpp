#include "Property.h"
#include <sstream>
#include <sys/system_properties.h>

std::string GetStrEncParam(const char *key)
{
    char buf[PROP_VALUE_MAX] = { 0 };
    __system_property_get(key, buf);
    return std::string(buf);
}
int32_t StrToInt(std::string value)
{
    std::stringstream ss(value);
    int32_t r = -1;
    ss >> r;
    return r;
}
int32_t GetIntEncParam(const char *key) { return StrToInt(GetStrEncParam(key)); }
void SetEncParam(const char *key, const char *value) { __system_property_set(key, value); }
