/* oracle/shim -- TEST INFRASTRUCTURE ONLY. The shim mesh is always triangulated; repair is never reached. */
#pragma once
#include <geogram/mesh/mesh.h>
namespace GEO {
enum MeshRepairMode { MESH_REPAIR_TRIANGULATE = 16, MESH_REPAIR_QUIET = 32 };
inline void mesh_repair(Mesh&, MeshRepairMode) {}
}
