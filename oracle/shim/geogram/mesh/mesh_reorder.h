/* oracle/shim -- TEST INFRASTRUCTURE ONLY. The wrapper pre-sorts facets and constructs the tree with reorder=false;
 * a call with reorder=true keeps the given order (results do not depend on facet order). */
#pragma once
#include <geogram/mesh/mesh.h>
namespace GEO {
enum MeshOrder { MESH_ORDER_HILBERT, MESH_ORDER_MORTON };
inline void mesh_reorder(Mesh&, MeshOrder) {}
}
