// nvblox/serialization/mesh_serializer_gpu.h -- see layer_serializer_gpu.h (both serializers live there).
#pragma once
#include "nvblox/serialization/layer_serializer_gpu.h"
