// Stand-in for ROOT <TFile.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "root_shim.h"
