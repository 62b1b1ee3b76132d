/* TEST INFRASTRUCTURE ONLY: see Rinternals.h in this directory. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
#endif
