/* Drop-in shim: code that says `#include <nlopt.h>` (reference: src/api/nlopt.h)
 * gets the B200 library's declarations, which carry the same names, enum values
 * and signatures for the whole nlopt_* object API. */
#ifndef NLOPT_H
#define NLOPT_H
#include "nlopt_b200.h"
#endif
