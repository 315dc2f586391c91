// Forwarding header: Generate / BoundGenerate live in supersonic/cursor.h of this implementation.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_GENERATE_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_GENERATE_H_
#include "supersonic/cursor.h"
#endif
