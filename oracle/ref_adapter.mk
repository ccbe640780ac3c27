# oracle/ref_adapter.mk -- TEST INFRASTRUCTURE ONLY: builds the reference's own encoder adapter, UNMODIFIED, from the sources where they lie
# under /root/reference into oracle/_ref/libVideoCodecRef.so (git-ignored; travels to the GPU box with the snapshot). Nothing is copied.
#   video_codec/{VideoCodecApi,VideoEncoderOpenH264,VideoEncoderNetint}.cpp + common/{log,prop}/*.cpp, g++ -std=c++14, plus a property store
#   (media_b200/host/PropertyStore.cpp) standing in for bionic's __system_property_get/set.
# The adapter holds no arithmetic: it dlopen()s "libopenh264.so" at run time (VideoEncoderOpenH264.cpp:46,203). With cisco's library on
# LD_LIBRARY_PATH / in baseline/_ref/ this is the real CPU baseline (bench.py --impl reference); with media_b200/shim/libopenh264.so on the
# path the same unmodified adapter drives the B200 encoder (tests/test_gpu_parity.py::test_unmodified_reference_adapter_drives_the_gpu).
REF ?= /root/reference
CXX := $(shell command -v /usr/bin/g++ || echo g++)
OUT = _ref/libVideoCodecRef.so
SRCS = $(REF)/video_codec/VideoCodecApi.cpp $(REF)/video_codec/VideoEncoderOpenH264.cpp $(REF)/video_codec/VideoEncoderNetint.cpp \
       $(REF)/common/log/MediaLog.cpp $(REF)/common/log/MediaLogManager.cpp $(REF)/common/prop/Property.cpp ../media_b200/host/PropertyStore.cpp ref_capi.cpp
INC = -Iref_shim -I$(REF)/video_codec -I$(REF)/common/log -I$(REF)/common/prop -I$(REF)/vendor/openh264 -I$(REF)/vendor/netint
$(OUT): $(SRCS) ref_adapter.mk
	mkdir -p _ref
	$(CXX) -std=c++14 -O2 -fPIC -shared $(INC) -o $@ $(SRCS) -ldl -lpthread
