// api_internal.hpp -- glue shared by rt_api.cpp (host) and render.cu (device).
#pragma once

#include <string>

#include "scene.hpp"

namespace rtb {

void set_error(const std::string& msg);
HostScene* host_of(rt_scene* s);
const HostScene* host_of(const rt_scene* s);

// render.cu
void device_release(HostScene& h);     // frees everything on the device
void device_invalidate(HostScene& h);  // marks the device copy of the scene stale (next use uploads again)

}  // namespace rtb
