/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <geogram/mesh/mesh.h>
