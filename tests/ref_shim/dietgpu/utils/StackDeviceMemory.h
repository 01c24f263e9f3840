// shim: StackDeviceMemory / GpuMemoryReservation / makeStackMemory come from the product's drop-in
// layer; the reference header also drags in its device utilities, static helpers and glog.
#pragma once
#include "dietgpu_b200_compat.hpp"
#include "dietgpu/utils/DeviceUtils.h"
#include "dietgpu/utils/StaticUtils.h"
#include "glog/logging.h"
