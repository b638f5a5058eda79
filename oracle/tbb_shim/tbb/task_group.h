// TBB shim forwarding header (test infrastructure, see ../tbb_shim.h)
#pragma once
#include "../tbb_shim.h"
