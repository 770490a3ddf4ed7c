/* oracle/shim -- TEST INFRASTRUCTURE ONLY; see geogram/basic/common.h in this directory. */
#pragma once
#include <geogram/basic/geometry.h>
extern "C" double ora_point_triangle_sqdist(const double* p, const double* v0, const double* v1, const double* v2, double* nearest3);
namespace GEO { namespace Geom {
/* geogram's Eberly point-triangle distance is not under /root/reference: forwarded to the oracle's restatement. */
inline double point_triangle_squared_distance(const vec3& p, const vec3& p1, const vec3& p2, const vec3& p3,
                                              vec3& nearest, double& l1, double& l2, double& l3) {
    l1 = l2 = l3 = 0.0;
    return ora_point_triangle_sqdist(p.data(), p1.data(), p2.data(), p3.data(), nearest.data());
}
} }
