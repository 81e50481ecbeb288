// Stand-in for ROOT's <Minuit2/FCNBase.h>, <Minuit2/FunctionMinimizer.h>, <Minuit2/FunctionMinimum.h>: ORACLE / TEST
// INFRASTRUCTURE.  src/ISRSolverVCSFitter.cpp derives its chi2 objective from FCNBase and reads the fitted parameters back
// through FunctionMinimum::UserParameters().Params(); the minimiser itself lives in the reference's executables and in
// PyFitVCS, which are not built.  Only those three names are provided.
#ifndef ISR_REF_SHIM_MINUIT2
#define ISR_REF_SHIM_MINUIT2
#include <vector>
namespace ROOT {
namespace Minuit2 {
class FCNBase {
 public:
  virtual ~FCNBase() {}
  virtual double operator()(const std::vector<double>&) const = 0;
  virtual double Up() const = 0;
};
class MnUserParameters {
 public:
  explicit MnUserParameters(const std::vector<double>& p) : p_(p) {}
  std::vector<double> Params() const { return p_; }

 private:
  std::vector<double> p_;
};
class FunctionMinimum {
 public:
  explicit FunctionMinimum(const std::vector<double>& p) : u_(p) {}
  const MnUserParameters& UserParameters() const { return u_; }

 private:
  MnUserParameters u_;
};
class FunctionMinimizer {};
}  // namespace Minuit2
}  // namespace ROOT
#endif
