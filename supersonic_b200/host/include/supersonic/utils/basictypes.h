// supersonic/utils/basictypes.h:14-17 (the part client code of the hot path uses)
#ifndef SUPERSONIC_B200_HOST_UTILS_BASICTYPES_H_
#define SUPERSONIC_B200_HOST_UTILS_BASICTYPES_H_
#include "supersonic/base.h"
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
#endif
