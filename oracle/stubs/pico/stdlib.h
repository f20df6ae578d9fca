/* Stand-in for pico/stdlib.h: only the C basics flash_storage.c needs.  Test infrastructure. */
#pragma once
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
