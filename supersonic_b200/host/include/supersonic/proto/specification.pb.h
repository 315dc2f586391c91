// Forwarding header: plan code that includes the reference's generated
// "supersonic/proto/specification.pb.h" (ExtendedSortSpecification) compiles unchanged against the mirror.
#ifndef SUPERSONIC_B200_HOST_PROTO_SPECIFICATION_PB_H_
#define SUPERSONIC_B200_HOST_PROTO_SPECIFICATION_PB_H_
#include "supersonic/cursor.h"
#endif  // SUPERSONIC_B200_HOST_PROTO_SPECIFICATION_PB_H_
