// Stand-in for <boost/lexical_cast.hpp>: ORACLE / TEST INFRASTRUCTURE.  src/ISRSolverSLE.cpp includes it without using it.
#ifndef ISR_REF_SHIM_BOOST_LEXICAL_CAST
#define ISR_REF_SHIM_BOOST_LEXICAL_CAST
#endif
