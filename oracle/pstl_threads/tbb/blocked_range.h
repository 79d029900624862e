/* see tbb.h in this directory (test infrastructure) */
#include "tbb.h"
