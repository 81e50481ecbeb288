// Stand-in for ROOT <Minuit2/FunctionMinimum.h>: ORACLE / TEST INFRASTRUCTURE, see FCNBase.h next to this file.
#include "FCNBase.h"
