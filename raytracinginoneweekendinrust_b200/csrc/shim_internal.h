// shim_internal.h — shared between the CUDA-free builder entry points and the device driver.
#pragma once
#include <string>
#include "../../include/shimmer_b200.h"
#include "shim_scene.h"

#define SHIM_API extern "C" __attribute__((visibility("default")))

namespace shim {
struct DeviceState;                       // device buffers + wavefront pool (shim_api.cu)
void device_state_release(DeviceState*);  // defined by whoever owns the device side
int set_err(int code, const std::string& m);
}  // namespace shim

struct shim_scene {
    shim::SceneBuilder sb;
    shim::FlatScene flat;
    bool committed = false;
    bool has_media = false;
    shim::DeviceState* dev = nullptr;
};

#define NEED(s) if (!(s)) return shim::set_err(SHIM_ERR_INVALID, "null scene")
#define MUTABLE(s) NEED(s); if ((s)->committed) return shim::set_err(SHIM_ERR_STATE, "scene already committed")
