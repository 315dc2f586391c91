// Forwarding header: StringPiece (supersonic/utils/strings/stringpiece.h) lives in supersonic/base.h here.
#ifndef SUPERSONIC_B200_HOST_UTILS_STRINGS_STRINGPIECE_H_
#define SUPERSONIC_B200_HOST_UTILS_STRINGS_STRINGPIECE_H_
#include "supersonic/base.h"
#endif
