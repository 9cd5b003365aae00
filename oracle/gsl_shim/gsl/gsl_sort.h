/* forwarding header of the minimal GSL-compatible shim (test infrastructure, see ../gsl_shim.h) */
#include "../gsl_shim.h"
