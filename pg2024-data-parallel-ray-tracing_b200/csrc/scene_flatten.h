// scene_flatten.h -- indexed, instanced meshes -> flat per-corner streams (see scene_flatten.cpp).
#pragma once
#include <cstdint>
#include "dprt_types.h"

namespace dprt {

// total triangles over all instances; < 0: invalid description
int64_t flatten_count(const dprt_mesh_desc* meshes, int nMeshes, const dprt_instance_desc* instances, int64_t nInstances);
// fills verts9 / normals9 (9 floats per flattened triangle), uv6 (6 per triangle, may be null), matIds (may be null);
// *hasUv (may be null) = some mesh carries texture coordinates. 0 = ok.
int flatten_instances(const dprt_mesh_desc* meshes, int nMeshes, const dprt_instance_desc* instances, int64_t nInstances,
                      float* verts9, float* normals9, float* uv6, int32_t* matIds, int* hasUv);

}  // namespace dprt
