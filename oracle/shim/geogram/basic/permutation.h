/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#pragma once
