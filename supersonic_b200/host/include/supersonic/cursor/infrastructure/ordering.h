// Forwarding header: SortOrder / BoundSortOrder / ColumnOrder (supersonic/cursor/infrastructure/ordering.h:48-215)
// are declared in supersonic/cursor.h here.
#ifndef SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_ORDERING_H_
#define SUPERSONIC_B200_HOST_CURSOR_INFRASTRUCTURE_ORDERING_H_
#include "supersonic/cursor.h"
#endif
