/* ArrayLDPC.h -- the reference's public include (an empty guard there, ArrayLDPC.h:1-6); here it forwards to
 * the facade classes and the C ABI of the B200 engine. */
#ifndef PROTO_LDPC_H
#define PROTO_LDPC_H
#include "ArrayLDPCMacro.h"
#include "ldpc_capi.h"
#endif
