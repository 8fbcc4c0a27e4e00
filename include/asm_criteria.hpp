// asm_criteria.hpp -- host side above the C ABI: a C++ mirror of the reference's plugin interface for
// the hot path, same names, same inputs and defaults, same state machine, same error behaviour.
//
// The reference is Java (package asm.inference); no JDK exists in this image, so the host side is C++
// (task rule: "C++ where the reference is compiled code").  The Java binding a maintainer adds is in
// INTEGRATION.md / java/.  Every arithmetic step is delegated to libasmgpu (include/asmgpu.h); what is
// kept here is exactly what the reference keeps in Java: thresholds, `delta` doubling, `start0`/`prev`
// bookkeeping, the NaN-propagating min, column mapping.
//
//   reference                                   here
//   MCMCConvergenceCriterion.java:8-31          asmgpu::MCMCConvergenceCriterion
//   TraceInfo.java:20-127                       asmgpu::TraceInfo (owns the GPU context + clade dictionary)
//   TreeESS.java:17-204                         asmgpu::TreeESS
//   TreePSRF.java:14-332                        asmgpu::TreePSRF
//   TraceESS.java:12-204                        asmgpu::TraceESS
//   AutoStopMCMC.readLogLines/readTreeLogLine   asmgpu::TraceInfo::readLogLine / readTreeLogLine
#pragma once
#include <cstdint>
#include <map>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "asmgpu.h"

namespace asmgpu {

// java.lang.IllegalArgumentException
struct IllegalArgumentException : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};

class TraceInfo {  // TraceInfo.java
 public:
  explicit TraceInfo(int chainCount, int device = 0);  // TraceInfo.java:31-41
  ~TraceInfo();
  TraceInfo(const TraceInfo&) = delete;
  TraceInfo& operator=(const TraceInfo&) = delete;

  int chainCount() const { return (int)nTrees_.size(); }  // :107-109

  // ---- what the runner appends (AutoStopMCMC.java:430-538) ----
  // One line of a trace log.  Returns true if a sample was appended (readLogLines :430-479): comment
  // lines and single-token lines are skipped, the line starting with "Sample" sets the labels, an
  // unparsable cell becomes 0.0.
  bool readLogLine(int chain, const std::string& line);
  // One line of a tree log.  Returns true if a tree was appended (readTreeLogLine :486-538): the first
  // `translate` block fixes the taxon numbering; `tree STATE...` lines are parsed from the first '('
  // with label offset 1.
  bool readTreeLogLine(int chain, const std::string& line);
  void setLabels(const std::vector<std::string>& labels);  // :125-127
  void addLogLine(int chain, const double* values, int n);
  void setTaxonCount(int nTaxa);
  void addTreeNewick(int chain, const std::string& newick, int labelOffset = 1);
  void addTree(int chain, const int32_t* left, const int32_t* right, int32_t root);

  // ---- TraceInfo.java:52-104 ----
  const std::vector<int>& getMap();
  void setUpMap(const std::string& tracesString);
  const std::vector<std::string>& getTraceLabels();

  int64_t treeCount(int chain) const { return nTrees_[(size_t)chain]; }
  int64_t logCount(int chain) const { return nLines_[(size_t)chain]; }
  // logLines[chain][col].get(x): the host copy the burn-in detector reads (the criteria read the GPU copy)
  double logValue(int chain, int col, int64_t x) const { return hostLogs_[(size_t)chain][(size_t)x * (size_t)hostCols_[(size_t)chain] + (size_t)col]; }
  int taxonCount() const { return nTaxa_; }
  int64_t cladeCount() const;
  // The GPU context with everything appended so far uploaded: trees and log lines are staged on the host as they
  // arrive and go up in one copy per chain here, i.e. once per convergence check instead of once per sample.
  asm_ctx* ctx() {
    flush();
    return ctx_;
  }
  void flush();

 private:
  asm_ctx* ctx_ = nullptr;
  asm_encoder* enc_ = nullptr;
  int nTaxa_ = 0;
  std::vector<int64_t> nTrees_, nLines_;
  std::vector<std::vector<double>> hostLogs_;  // [chain][line * cols + col]
  std::vector<int> hostCols_;
  std::vector<std::string> labels_, traces_;
  std::vector<int> map_;
  bool haveMap_ = false, haveLabels_ = false;
  std::string tracesString_;
  bool haveTracesString_ = false;
  bool inTranslate_ = false, haveTaxa_ = false;
  int translateCount_ = 0;
  std::vector<int32_t> idsScratch_;
  std::vector<uint64_t> bitsScratch_;
  std::vector<uint16_t> ids16Scratch_;
  std::vector<std::vector<int32_t>> pendingIds_;  // [chain]: clade ids of the trees not uploaded yet
  std::vector<int64_t> linesUploaded_;            // [chain]: lines of hostLogs_ already on the GPU
  void appendIds_(int chain);
};

// BurnInStrategy.java:3-37
struct BurnInStrategy {
  enum Kind { Running, Automatic } kind = Automatic;
  int burnInPercent = 0;  // Automatic: 0; Running: 10 (DEFAULT_BURN_IN_PERCENT :11)
  static BurnInStrategy running(int percent = 10) {
    BurnInStrategy s;
    s.kind = Running;
    s.setBurnInPercent(percent);
    return s;
  }
  static BurnInStrategy automatic() { return BurnInStrategy(); }
  void setBurnInPercent(int p) {  // :30-36
    if (p < 0 || p > 99) throw IllegalArgumentException("BurnInPercent must be between 0 and 99");
    burnInPercent = p;
  }
};

// BurnInDetector.java:12-80 -- supplies burnin[] to every converged() call (AutoStopMCMC.java:397-409).  Host code, as
// in the reference: it is the caller of the path, O(end) per column per check, not part of it.
class BurnInDetector {
 public:
  BurnInDetector(TraceInfo* traceInfo, BurnInStrategy burnInStrat) : traceInfo(traceInfo), burnInStrat(burnInStrat) {}
  std::vector<int> burnIn(int end);  // :21-36
  static constexpr int WINDOW_SIZE = 10;  // :49

 private:
  TraceInfo* traceInfo;
  BurnInStrategy burnInStrat;
  int burnIn(int chainNr, int end);           // :38-47
  int burnIn(int chainNr, int col, int end);  // :51-78
};

class MCMCConvergenceCriterion {  // MCMCConvergenceCriterion.java:8-31
 public:
  virtual ~MCMCConvergenceCriterion() = default;
  virtual bool converged(const int* burnin, int end) = 0;     // :14
  virtual void setup(int nChains, TraceInfo* traceInfo) = 0;  // :20
  virtual std::string header() const = 0;                     // :26 getClass().getSimpleName()
  virtual void initAndValidate() {}                           // :30
};

class TreeESS : public MCMCConvergenceCriterion {  // TreeESS.java
 public:
  // Inputs (TreeESS.java:18-26)
  int targetESSInput = 100;
  double smoothingInput = 1.0;
  int cacheLimitInput = 1024;
  int ESSSampleSizeInput = 10;  // "sampleSize"
  // not in the reference: which distance kernel the GPU uses (ASM_RF_POPC / ASM_RF_I8 / ASM_RF_FP4; same integers)
  int methodInput = ASM_RF_I8;

  void initAndValidate() override;                          // :45-76 (as written: rejects every smoothing >= 0)
  bool converged(const int* burnin, int end) override;      // :79-112
  void setup(int nChains, TraceInfo* traceInfo) override;   // :189-199
  std::string header() const override { return "TreeESS"; }
  int getNumChains() const { return numChains; }            // :31-33
  int getDelta() const { return delta; }
  std::string lastLog;  // what the reference prints through Log.info / Log.warning during the last call

 protected:
  TraceInfo* traceInfo = nullptr;
  int numChains = 0;
  int targetESS = 0;
  double smoothing = 0;
  int cacheLimit = 0;
  int delta = 1;  // :40
  int N = 0;      // :41
  std::vector<int> indices;  // :115
  bool haveIndices = false;
  bool allowManyChains = false;  // TreePSRF's generalisation to C > 2
  double pseudoESS(int treeSet, int cutStart, int cutEnd);  // :117-166
};

class TreePSRF : public TreeESS {  // TreePSRF.java
 public:
  // Inputs (TreePSRF.java:15-20)
  double bInput = 0.05;
  bool twoSidedInput = true;
  bool checkESSInput = false;

  TreePSRF() { allowManyChains = true; }
  void initAndValidate() override;                         // :54-93
  bool converged(const int* burnin, int end) override;     // :96-105
  void setup(int nChains, TraceInfo* traceInfo) override;  // :304-318
  std::string header() const override { return "TreePSRF"; }
  const std::vector<double>& getGrValues() const { return grValues; }  // :43-45
  std::map<std::string, double> getLog() const;                        // :320-328
  int getStart0() const { return start0; }
  bool threw = false;  // the last call went through the catch (Throwable) of :188-194

 private:
  int start0 = -1;  // :28
  double upper = 0, lower = 0;
  bool prev = false;  // :31
  bool twoSided = true, checkESS = false;
  std::vector<double> grValues;  // :51; [2] for two chains, [C*C] for the generalisation
  std::vector<double> means_;
  bool converged2sided(const int* burnin, int end);            // :107-197
  bool converged1sided(int side, const int* burnin, int end);  // :200-261
  bool convergedMulti(const int* burnin, int end);             // SURVEY Appendix A.1, C > 2
  void evalMeans(const int* burnin, int start, int end);
};

class TraceESS : public MCMCConvergenceCriterion {  // TraceESS.java
 public:
  int targetESSInput = 100;                              // :14
  std::string tracesInput = "posterior,prior,likelihood";  // :15
  static constexpr int MAX_LAG = 2000;                   // :21

  void initAndValidate() override { targetESS = targetESSInput; }  // :27-29
  bool converged(const int* burnin, int end) override { return converged(burnin, end, true); }  // :34-36
  bool converged(const int* burnin, int end, bool useMapping);                                   // :38-69
  void setup(int nChains, TraceInfo* traceInfo) override;                                        // :182-188
  std::string header() const override { return "TraceESS"; }
  std::map<std::string, double> getLogMap();                                                     // :191-198
  const std::vector<double>& currentESS() const { return currentESSs; }
  double lastMinESS = 0;

 protected:
  TraceInfo* traceInfo = nullptr;
  int nChains = 0;
  int targetESS = 0;
  std::vector<double> currentESSs;  // :24
};

class GelmanRubin : public MCMCConvergenceCriterion {  // GelmanRubin.java
 public:
  double acceptedThresholdInput = 1.05;  // "threshold" :14
  void initAndValidate() override { acceptedThreshold = acceptedThresholdInput; }  // :27-29
  bool converged(const int* burnin, int available) override;                       // :32-51
  void setup(int nChains_, TraceInfo* ti) override {                               // :101-104
    nChains = nChains_;
    traceInfo = ti;
  }
  std::string header() const override { return "GelmanRubin"; }
  double lastMaxGR = 0;  // largest statistic seen by the last call (NaNs skipped), for logging/tests

 private:
  TraceInfo* traceInfo = nullptr;
  int nChains = 0;
  double acceptedThreshold = 1.05;
};

class CladeDifference : public MCMCConvergenceCriterion {  // CladeDifference.java
 public:
  double acceptedThresholdInput = 0.25;  // "threshold" :16
  void initAndValidate() override { acceptedThreshold = acceptedThresholdInput; }  // :32-35
  bool converged(const int* burnin, int available) override;                       // :41-65
  void setup(int nChains_, TraceInfo* ti) override {                               // :68-77
    nChains = nChains_;
    traceInfo = ti;
    m_fMaxCladeProbDiffs.clear();
  }
  std::string header() const override { return "CladeDifference"; }
  std::vector<double> m_fMaxCladeProbDiffs;  // :19

 private:
  TraceInfo* traceInfo = nullptr;
  int nChains = 0;
  double acceptedThreshold = 0.25;
};

// AutoStopMCMC.combinelogs (AutoStopMCMC.java:247-317) for one logger: appends the samples [burnIn[i],
// lastReported) of every chain's log file to `out`, renumbered 0, sampleDelta, 2*sampleDelta, ... across the
// chains; treeMode selects the tree-log branch (:260-285) or the trace-log branch (:287-312).  Returns the
// number of samples written.  A file that ends early throws (the reference's NullPointerException).  The header
// BEAST's Logger.init() prints (:254) is the caller's business.
int64_t combineLogs(bool treeMode, const std::vector<std::string>& chainFiles, const int* burnIn, int lastReported,
                    int64_t sampleDelta, std::ostream& out);

}  // namespace asmgpu
