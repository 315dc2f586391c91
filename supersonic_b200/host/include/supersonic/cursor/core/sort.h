// Forwarding header: Sort / ExtendedSort / BoundSort (supersonic/cursor/core/sort.h:65-131) are declared with the
// other operation factories in supersonic/cursor.h here.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_SORT_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_SORT_H_
#include "supersonic/cursor.h"
#endif
