// TEST INFRASTRUCTURE: stand-in for the OptiX SDK header of the same name (see optix_shim.h)
#pragma once
#include "../optix_shim.h"
