// drop-in replacement for the reference's src/common/Material.h: the B200 shim (rfx_shim.hpp) provides the class surface
#pragma once
#include "rfx_shim.hpp"
