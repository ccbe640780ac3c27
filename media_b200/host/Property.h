// Property helpers with the semantics of the reference's common/prop/Property.{h,cpp}:6-9 / :8-44
// (unset or non-numeric -> -1 / 0 as a stringstream extraction gives).
#ifndef B200_PROPERTY_H
#define B200_PROPERTY_H
#include <cstdint>
#include <string>
int32_t GetIntEncParam(const char *key);
std::string GetStrEncParam(const char *key);
void SetEncParam(const char *key, const char *value);
int32_t StrToInt(std::string value);
#endif
