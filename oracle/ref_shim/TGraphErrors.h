// Stand-in for ROOT <TGraphErrors.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "root_shim.h"
