/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <geogram/mesh/mesh.h>
#include <geogram/basic/geometry_nd.h> /* real geogram reaches point_triangle_squared_distance through this header too */
