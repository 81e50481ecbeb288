// Stand-in for ROOT <TGraph.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "root_shim.h"
