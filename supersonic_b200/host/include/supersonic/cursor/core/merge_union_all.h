// Forwarding header: MergeUnionAll / BoundMergeUnionAll (supersonic/cursor/core/merge_union_all.h) are declared
// with the other operation factories in supersonic/cursor.h here.
#ifndef SUPERSONIC_B200_HOST_CURSOR_CORE_MERGE_UNION_ALL_H_
#define SUPERSONIC_B200_HOST_CURSOR_CORE_MERGE_UNION_ALL_H_
#include "supersonic/cursor.h"
#endif
