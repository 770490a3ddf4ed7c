/*
 * oracle/shim -- TEST INFRASTRUCTURE ONLY.
 * A minimal stand-in for the subset of the geogram API (Yixin-Hu fork b613750, not vendored in /root/reference)
 * that the reference's src/tetwild/geogram/mesh_AABB.{h,cpp} and Common.cpp::sampleTriangle use, so that those
 * reference sources can be compiled UNMODIFIED, from where they lie, into oracle/_ref (see oracle/ref_build.sh).
 * Everything here is our own code written against the API names only; vec3 arithmetic follows geogram's
 * documented operator semantics (RECOLLECTED, see oracle/envelope.c header).
 */
#pragma once
#include <vector>
#include <cassert>
#include <cmath>
#include <algorithm>
#define GEOGRAM_API
#define geo_debug_assert(x) assert(x)
#define geo_assert(x) assert(x)
namespace GEO {
typedef unsigned int index_t;
typedef unsigned char coord_index_t;
static const index_t NO_FACET = index_t(-1);
template <class T> class vector : public std::vector<T> {
public:
    using std::vector<T>::vector;
};
template <class T> inline T geo_sqr(T x) { return x * x; }
template <class T> inline int geo_sgn(const T& x) { return (x > 0) ? 1 : ((x < 0) ? -1 : 0); }
}
