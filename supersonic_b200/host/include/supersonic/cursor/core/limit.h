// Forwarding header: Limit / BoundLimit live in supersonic/cursor.h of this implementation.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_LIMIT_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_LIMIT_H_
#include "supersonic/cursor.h"
#endif
