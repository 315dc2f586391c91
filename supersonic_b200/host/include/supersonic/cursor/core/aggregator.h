// Forwarding header: the reference declares Aggregator in supersonic/cursor/core/aggregator.h
// (not pulled in by supersonic.h); here it lives with the other cursor-level declarations.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_AGGREGATOR_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_AGGREGATOR_H_
#include "supersonic/cursor.h"
#endif
