// Stand-in for <boost/accumulators/...>: ORACLE / TEST INFRASTRUCTURE.  src/Chi2Test.cpp (chi2TestData) averages the
// toy solutions with accumulator_set<double, stats<tag::mean>>; Boost's tag::mean is sum / count.
#ifndef ISR_REF_SHIM_BOOST_ACCUMULATORS
#define ISR_REF_SHIM_BOOST_ACCUMULATORS
#include <cstddef>
namespace boost {
namespace accumulators {
namespace tag {
struct mean {};
}  // namespace tag
template <typename... Tags>
struct stats {};
template <typename T, typename Stats>
class accumulator_set {
 public:
  accumulator_set() : sum_(0), count_(0) {}
  void operator()(const T& v) {
    sum_ += v;
    ++count_;
  }
  T sum_;
  std::size_t count_;
};
template <typename T, typename Stats>
T mean(const accumulator_set<T, Stats>& a) {
  return a.sum_ / static_cast<T>(a.count_);
}
}  // namespace accumulators
}  // namespace boost
#endif
