// supersonic/supersonic.h -- umbrella header of the B200 implementation; the include that
// plan code written against the reference's supersonic/supersonic.h:20-69 keeps using.
#ifndef SUPERSONIC_B200_HOST_SUPERSONIC_H_
#define SUPERSONIC_B200_HOST_SUPERSONIC_H_
#include "supersonic/base.h"
#include "supersonic/projector.h"
#include "supersonic/expression.h"
#include "supersonic/cursor.h"
#endif
