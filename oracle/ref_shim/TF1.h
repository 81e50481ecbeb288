// Stand-in for ROOT <TF1.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "root_shim.h"
