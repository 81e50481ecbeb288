// Stand-in for ROOT <Math/PdfFuncMathCore.h>: ORACLE / TEST INFRASTRUCTURE, see root_shim.h.
#include "../root_shim.h"
