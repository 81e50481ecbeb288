// Stand-in for <nlopt.hpp>: ORACLE / TEST INFRASTRUCTURE.  src/ISRSolverTikhonov.cpp and src/Utils.cpp include it but the
// code compiled here (solve, the L-curve functionals, lambdaObjective with NLopt's plain C callback signature) uses no
// NLopt symbol; the optimiser loop itself lives in the reference's executables, which are not built.
#ifndef ISR_REF_SHIM_NLOPT_HPP
#define ISR_REF_SHIM_NLOPT_HPP
#endif
