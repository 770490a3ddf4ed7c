/* oracle/shim -- TEST INFRASTRUCTURE ONLY; see geogram/basic/common.h in this directory. */
#pragma once
#include <geogram/basic/geometry.h>
namespace GEO {
class Mesh {
public:
    struct Vertices {
        std::vector<double> xyz;
        index_t nb() const { return index_t(xyz.size() / 3); }
        const double* point_ptr(index_t v) const { return &xyz[3 * size_t(v)]; }
        double* point_ptr(index_t v) { return &xyz[3 * size_t(v)]; }
    } vertices;
    struct Facets {
        index_t n = 0;
        index_t nb() const { return n; }
        index_t corners_begin(index_t f) const { return 3 * f; }
        index_t corners_end(index_t f) const { return 3 * f + 3; }
        index_t nb_vertices(index_t) const { return 3; }
        bool are_simplices() const { return true; }
    } facets;
    struct FacetCorners {
        std::vector<index_t> v;
        index_t vertex(index_t c) const { return v[c]; }
    } facet_corners;
};
namespace Geom {
inline const vec3& mesh_vertex(const Mesh& M, index_t v) { return *reinterpret_cast<const vec3*>(M.vertices.point_ptr(v)); }
}
}
