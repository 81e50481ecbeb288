// Stand-ins for the ROOT classes the reference's solver / toy-test sources touch: ORACLE / TEST INFRASTRUCTURE.
//
// ROOT is a third-party dependency of the reference (CMakeLists.txt: find_package(ROOT)) that is absent from this image.
// The reference uses it on this path for I/O and book-keeping only, never for the numbers:
//   TFile / TGraphErrors / TMatrixD / TF1         containers read from and written to .root files
//   TEfficiency                                   a binned table looked up by FindFixBin + GetEfficiency
//   TH1F                                          Fill, GetMean, GetMeanError, Fit (the fit result is only drawn)
//   gRandom->Gaus                                 the toy generator (unseeded in the reference)
// Here files are an in-memory registry (name -> object) shared by every "file": the glue (oracle/ref_solver_glue.cpp)
// registers the model graphs before the reference's chi2TestModel / ratioTestModel open their input "file", and reads
// what those functions Write() afterwards.  TH1F keeps ROOT's documented bin arithmetic (TAxis::FindBin for fixed bins,
// the restated orc_find_bin of oracle/isr_oracle.c) and ROOT's statistics rule (entries outside the axis range count in
// the under/overflow bins but not in the mean / RMS sums); bin contents are float like TH1F's.  gRandom->Gaus returns the
// next value of a queue the glue fills, so the reference's own toy loops run on exactly the toys the CUDA path is given.
#ifndef ISR_REF_SHIM_ROOT_H
#define ISR_REF_SHIM_ROOT_H
#include <cmath>
#include <cstddef>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" int orc_find_bin(int nbins, double lo, double hi, const double* edges, double x);

class TObject {
 public:
  virtual ~TObject() {}
  virtual TObject* Clone(const char* = "") const { return nullptr; }
  virtual int Write(const char* = nullptr) { return 0; }
};

namespace root_shim {
// every Open()ed "file" reads from and writes to this one registry; entries are owned by it
std::map<std::string, std::shared_ptr<TObject>>& registry();
// values the next gRandom->Gaus calls return, in order
std::deque<double>& gaus_queue();
}  // namespace root_shim

class TGraphErrors : public TObject {
 public:
  TGraphErrors() {}
  TGraphErrors(int n, const double* x, const double* y, const double* ex = nullptr, const double* ey = nullptr)
      : x_(x, x + n), y_(y, y + n), ex_((std::size_t)n, 0.0), ey_((std::size_t)n, 0.0) {
    if (ex) ex_.assign(ex, ex + n);
    if (ey) ey_.assign(ey, ey + n);
  }
  int GetN() const { return (int)x_.size(); }
  double* GetX() { return x_.data(); }
  double* GetY() { return y_.data(); }
  double* GetEX() { return ex_.data(); }
  double* GetEY() { return ey_.data(); }
  void SetMarkerStyle(int) {}
  TObject* Clone(const char* = "") const override { return new TGraphErrors(*this); }
  int Write(const char* name = nullptr) override {
    root_shim::registry()[name ? name : "Graph"] = std::make_shared<TGraphErrors>(*this);
    return 0;
  }

 private:
  std::vector<double> x_, y_, ex_, ey_;
};

class TGraph : public TObject {
 public:
  TGraph(int n, const double* x, const double* y) : x_(x, x + n), y_(y, y + n) {}
  int GetN() const { return (int)x_.size(); }
  double* GetX() { return x_.data(); }
  double* GetY() { return y_.data(); }
  void SetMarkerStyle(int) {}
  int Write(const char* name = nullptr) override {
    root_shim::registry()[name ? name : "Graph"] = std::make_shared<TGraph>(*this);
    return 0;
  }

 private:
  std::vector<double> x_, y_;
};

class TMatrixD : public TObject {
 public:
  TMatrixD(int r, int c) : r_(r), c_(c), v_((std::size_t)r * c, 0.0) {}
  void SetMatrixArray(const double* p) { v_.assign(p, p + v_.size()); }   // row-major, like ROOT
  int GetNrows() const { return r_; }
  int GetNcols() const { return c_; }
  const double* GetMatrixArray() const { return v_.data(); }
  int Write(const char* name = nullptr) override {
    root_shim::registry()[name ? name : "TMatrixT<double>"] = std::make_shared<TMatrixD>(*this);
    return 0;
  }

 private:
  int r_, c_;
  std::vector<double> v_;
};

class TF1 : public TObject {
 public:
  TF1(const char* name, const std::function<double(double*, double*)>& f, double lo, double hi, int npar)
      : name_(name), f_(f), lo_(lo), hi_(hi), par_((std::size_t)npar, 0.0) {}
  void SetNpx(int) {}
  void SetParameter(int i, double v) { par_[(std::size_t)i] = v; }
  double GetParameter(int i) const { return par_[(std::size_t)i]; }
  double Eval(double x) {
    double xx[1] = {x};
    return f_(xx, par_.data());
  }
  int Write(const char* name = nullptr) override {
    root_shim::registry()[name ? name : name_] = std::make_shared<TF1>(*this);
    return 0;
  }

 private:
  std::string name_;
  std::function<double(double*, double*)> f_;
  double lo_, hi_;
  std::vector<double> par_;
};

// A binned efficiency table.  Global bin numbering as in ROOT: bin = binx + (nx + 2) * biny, 0 / n+1 = under / overflow.
class TEfficiency : public TObject {
 public:
  TEfficiency(int dim, int nx, double xlo, double xhi, const double* xedges, int ny, double ylo, double yhi, const double* yedges,
              const double* values)
      : dim_(dim), nx_(nx), ny_(ny), xlo_(xlo), xhi_(xhi), ylo_(ylo), yhi_(yhi) {
    if (xedges) xe_.assign(xedges, xedges + nx + 1);
    if (yedges) ye_.assign(yedges, yedges + ny + 1);
    const std::size_t cells = dim == 1 ? (std::size_t)(nx + 2) : (std::size_t)(nx + 2) * (ny + 2);
    val_.assign(values, values + cells);
  }
  int GetDimension() const { return dim_; }
  int FindFixBin(double x) const { return orc_find_bin(nx_, xlo_, xhi_, xe_.empty() ? nullptr : xe_.data(), x); }
  int FindFixBin(double x, double y) const {
    const int bx = orc_find_bin(nx_, xlo_, xhi_, xe_.empty() ? nullptr : xe_.data(), x);
    const int by = orc_find_bin(ny_, ylo_, yhi_, ye_.empty() ? nullptr : ye_.data(), y);
    return bx + (nx_ + 2) * by;
  }
  double GetEfficiency(int bin) const { return val_[(std::size_t)bin]; }
  TObject* Clone(const char* = "") const override { return new TEfficiency(*this); }

 private:
  int dim_, nx_, ny_;
  double xlo_, xhi_, ylo_, yhi_;
  std::vector<double> xe_, ye_, val_;
};

class TH1F : public TObject {
 public:
  TH1F(const char* name, const char*, int nbins, double lo, double hi)
      : name_(name), nbins_(nbins), lo_(lo), hi_(hi), counts_((std::size_t)nbins + 2, 0.0f), entries_(0), sw_(0), swx_(0), swx2_(0) {}
  int Fill(double x) {
    ++entries_;
    fills_.push_back(x);
    const int bin = orc_find_bin(nbins_, lo_, hi_, nullptr, x);
    counts_[(std::size_t)bin] += 1.0f;
    if (bin == 0 || bin > nbins_) return -1;     // under/overflow: not part of the statistics
    sw_ += 1.0;
    swx_ += x;
    swx2_ += x * x;
    return bin;
  }
  double GetMean() const { return sw_ != 0.0 ? swx_ / sw_ : 0.0; }
  double GetStdDev() const {
    if (sw_ == 0.0) return 0.0;
    const double m = swx_ / sw_;
    return std::sqrt(std::fabs(swx2_ / sw_ - m * m));
  }
  double GetMeanError() const { return sw_ > 0.0 ? GetStdDev() / std::sqrt(sw_) : 0.0; }   // unit weights: neff = sum w
  template <typename F>
  void Fit(F*, const char* = "") {}               // the fitted curve is only drawn, never read back
  int Write(const char* name = nullptr) override {
    root_shim::registry()[name ? name : name_] = std::make_shared<TH1F>(*this);
    return 0;
  }
  // shim-side read-back
  int nbins() const { return nbins_; }
  double lo() const { return lo_; }
  double hi() const { return hi_; }
  const std::vector<float>& counts() const { return counts_; }
  const std::vector<double>& fills() const { return fills_; }
  double stat(int i) const { return i == 0 ? sw_ : (i == 1 ? swx_ : swx2_); }

 private:
  std::string name_;
  int nbins_;
  double lo_, hi_;
  std::vector<float> counts_;
  std::vector<double> fills_;
  long entries_;
  double sw_, swx_, swx2_;
};

// One efficiency histogram per energy point (BaseISRSolver2): fixed or variable bins, contents incl. under/overflow.
class TH1D : public TObject {
 public:
  TH1D(const char* name, const char*, int nbins, double lo, double hi) : name_(name), nbins_(nbins), lo_(lo), hi_(hi), c_((std::size_t)nbins + 2, 0.0) {}
  TH1D(const char* name, const char*, int nbins, const double* edges)
      : name_(name), nbins_(nbins), lo_(edges[0]), hi_(edges[nbins]), edges_(edges, edges + nbins + 1), c_((std::size_t)nbins + 2, 0.0) {}
  void SetBinContent(int bin, double v) { c_[(std::size_t)bin] = v; }
  double GetBinContent(int bin) const { return c_[(std::size_t)bin]; }
  int FindBin(double x) const { return orc_find_bin(nbins_, lo_, hi_, edges_.empty() ? nullptr : edges_.data(), x); }
  void SetDirectory(void*) {}
  TObject* Clone(const char* = "") const override { return new TH1D(*this); }

 private:
  std::string name_;
  int nbins_;
  double lo_, hi_;
  std::vector<double> edges_, c_;
};

class TFile {
 public:
  static TFile* Open(const char*, const char* = "") { return new TFile(); }
  void cd() {}
  void Close() {}
  TObject* Get(const char* name) {
    auto it = root_shim::registry().find(name);
    return it == root_shim::registry().end() ? nullptr : it->second.get();
  }
};

class TRandom {
 public:
  virtual ~TRandom() {}
  virtual double Gaus(double = 0.0, double = 1.0) {
    auto& q = root_shim::gaus_queue();
    if (q.empty()) throw std::runtime_error("root_shim: gRandom->Gaus called with an empty toy queue");
    const double v = q.front();
    q.pop_front();
    return v;
  }
};
extern TRandom* gRandom;

namespace ROOT {
namespace Math {
inline double chisquared_pdf(double x, double r, double x0 = 0.0) {
  if ((x - x0) < 0.0) return 0.0;
  const double a = r / 2.0 - 1.0;
  if (x == x0 && a == 0.0) return 0.5;
  return std::exp((r / 2.0 - 1.0) * std::log((x - x0) / 2.0) - ((x - x0) / 2.0) - std::lgamma(r / 2.0)) / 2.0;
}
}  // namespace Math
}  // namespace ROOT
#endif
