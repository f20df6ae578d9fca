/*
 * orc_float.c — float32 oracle (TEST INFRASTRUCTURE; see dspi_oracle.h).
 * Instantiates orc_float.inc for the strict and the fused arithmetic flavour.
 * Build: -O2 -mfma -ffp-contract=off -fno-fast-math (oracle/Makefile).
 */
#include <math.h>
#include <string.h>
#include "dspi_oracle.h"
#include "orc_internal.h"

#define ORC_FUSED 0
#include "orc_float.inc"
#undef ORC_FUSED

#define ORC_FUSED 1
#include "orc_float.inc"
#undef ORC_FUSED
