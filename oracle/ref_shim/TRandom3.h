// Stand-in for ROOT <TRandom3.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "root_shim.h"
