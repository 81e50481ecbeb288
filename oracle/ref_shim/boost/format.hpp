// Stand-in for <boost/format.hpp>: ORACLE / TEST INFRASTRUCTURE.  src/RatioTest.cpp builds histogram names with
// boost::str(boost::format("bias_pt_%1%") % i); positional %N% directives are all that is needed.  Arguments are
// rendered with std::to_string, not through an ostream: the oracle library may carry a statically linked libstdc++ whose
// locale machinery is not initialised inside a foreign process (Python).
#ifndef ISR_REF_SHIM_BOOST_FORMAT
#define ISR_REF_SHIM_BOOST_FORMAT
#include <type_traits>
#include <string>
#include <vector>
namespace boost {
class format {
 public:
  explicit format(const std::string& f) : f_(f) {}
  template <typename T>
  format& operator%(const T& v) {
    if constexpr (std::is_arithmetic<T>::value) args_.push_back(std::to_string(v));
    else args_.push_back(std::string(v));
    return *this;
  }
  std::string str() const {
    std::string out;
    for (std::size_t i = 0; i < f_.size(); ++i) {
      if (f_[i] == '%') {
        std::size_t j = i + 1, k = 0;
        while (j < f_.size() && f_[j] >= '0' && f_[j] <= '9') k = 10 * k + (std::size_t)(f_[j++] - '0');
        if (j < f_.size() && f_[j] == '%' && j > i + 1 && k >= 1 && k <= args_.size()) {
          out += args_[k - 1];
          i = j;
          continue;
        }
      }
      out += f_[i];
    }
    return out;
  }

 private:
  std::string f_;
  std::vector<std::string> args_;
};
inline std::string str(const format& f) { return f.str(); }
}  // namespace boost
#endif
