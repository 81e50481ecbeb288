// Stand-in for <boost/accumulators/statistics/mean.hpp>: ORACLE / TEST INFRASTRUCTURE.
#include "../accumulators.hpp"
