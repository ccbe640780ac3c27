/* Linux stand-in for bionic's <sys/system_properties.h>: the reference reads all of its configuration through
 * __system_property_get/set (common/prop/Property.cpp:8-44). Backed by an in-memory store (PropertyStore.cpp);
 * initial values can also come from the environment: key "a.b.c" <- variable "PROP_a_b_c". */
#ifndef B200_SYSTEM_PROPERTIES_SHIM_H
#define B200_SYSTEM_PROPERTIES_SHIM_H
#define PROP_VALUE_MAX 92
#ifdef __cplusplus
extern "C" {
#endif
int __system_property_get(const char *name, char *value);
int __system_property_set(const char *name, const char *value);
#ifdef __cplusplus
}
#endif
#endif
