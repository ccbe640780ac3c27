/* oracle/ref_shim/sys/system_properties.h -- the two bionic property calls the reference's common/prop/Property.cpp uses, declared for a
 * Linux build of the UNMODIFIED reference adapter (oracle/ref_adapter.mk). The definitions are the in-memory store of
 * media_b200/host/PropertyStore.cpp. TEST INFRASTRUCTURE ONLY. */
#ifndef B200_REF_SHIM_SYSTEM_PROPERTIES_H
#define B200_REF_SHIM_SYSTEM_PROPERTIES_H
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
