// Umbrella header of the source-compatible subset (reference: nvblox/include/nvblox/nvblox.h).
#pragma once
#include "nvblox/core/types.h"
#include "nvblox/geometry/plane.h"
#include "nvblox/integrators/esdf_slicer.h"
#include "nvblox/integrators/weighting_function.h"
#include "nvblox/map/blox.h"
#include "nvblox/map/layer.h"
#include "nvblox/map/voxels.h"
#include "nvblox/mapper/mapper.h"
#include "nvblox/sensors/camera.h"
#include "nvblox/sensors/image.h"
