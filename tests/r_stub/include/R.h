/* see Rinternals.h in this directory: stand-in for R's headers (test infrastructure only) */
#ifndef MCB_STUB_R_H
#define MCB_STUB_R_H
#include "Rinternals.h"
#endif
