// Forwarding header: Coalesce / BoundCoalesce live in supersonic/cursor.h of this implementation.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_COALESCE_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_COALESCE_H_
#include "supersonic/cursor.h"
#endif
