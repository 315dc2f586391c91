// Forwarding header: plan code includes supersonic/utils/basictypes.h for `Ownership`, which lives in supersonic/base.h here.
#ifndef SUPERSONIC_B200_HOST_UTILS_BASICTYPES_H_
#define SUPERSONIC_B200_HOST_UTILS_BASICTYPES_H_
#include "supersonic/base.h"
#endif
