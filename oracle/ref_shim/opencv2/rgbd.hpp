// stand-in: see ../cv_shim.h (oracle/ref_shim) -- test infrastructure only
#pragma once
#include "../cv_shim.h"
