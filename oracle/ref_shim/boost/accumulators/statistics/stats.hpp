// Stand-in for <boost/accumulators/statistics/stats.hpp>: ORACLE / TEST INFRASTRUCTURE.
#include "../accumulators.hpp"
