/* stand-in, see fw_stub.h (test infrastructure) */
#pragma once
#include "fw_stub.h"
