// host_criteria.cpp -- implementation of include/asm_criteria.hpp and its flat C wrapper
// (include/asmgpu_host.h).  Host logic only; every number comes from libasmgpu's kernels.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <locale.h>
#include <fstream>
#include <sstream>

#include "../../include/asm_criteria.hpp"
#include "../../include/asmgpu_host.h"

namespace asmgpu {

namespace {

[[noreturn]] void gpu_fail(asm_ctx* ctx, int32_t rc, const char* what) {
  std::string m = std::string(what) + ": " + (ctx ? asm_last_error(ctx) : "") + " (code " + std::to_string(rc) + ")";
  if (rc == ASM_E_RANGE) throw std::out_of_range(m);  // the Java's IndexOutOfBoundsException
  throw std::runtime_error(m);
}
#define GPU_CHECK(ctx, call)                \
  do {                                      \
    int32_t rc__ = (call);                  \
    if (rc__ != ASM_OK) gpu_fail((ctx), rc__, #call); \
  } while (0)

// String.split("\\s+"): a leading run of whitespace yields an empty first token; trailing empties are dropped.
std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0, n = s.size();
  if (n && std::isspace((unsigned char)s[0])) {
    out.emplace_back();
    while (i < n && std::isspace((unsigned char)s[i])) i++;
  }
  while (i < n) {
    size_t j = i;
    while (j < n && !std::isspace((unsigned char)s[j])) j++;
    out.emplace_back(s.substr(i, j - i));
    i = j;
    while (i < n && std::isspace((unsigned char)s[i])) i++;
  }
  while (!out.empty() && out.back().empty()) out.pop_back();
  return out;
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

// Double.parseDouble's grammar (FloatingDecimal.readJavaFormatString): surrounding whitespace trimmed, optional sign,
// then exactly "NaN", exactly "Infinity", a decimal literal (digits with an optional point, optional exponent) or a
// hexadecimal one (0x, hex digits with an optional point, mandatory binary exponent), optional type suffix f/F/d/D.
// Anything else -- "inf", "nan", "0x10", "1,5", "1e" -- is a NumberFormatException there and false here.  The value
// comes from strtod_l under the "C" locale: correctly rounded like Java's, overflow to Infinity, underflow towards 0,
// and independent of the embedding process's LC_NUMERIC.
bool java_parse_double(const std::string& raw, double* out) {
  size_t a = 0, b = raw.size();
  while (a < b && (unsigned char)raw[a] <= ' ') a++;  // String.trim()
  while (b > a && (unsigned char)raw[b - 1] <= ' ') b--;
  if (a == b) return false;
  const std::string t = raw.substr(a, b - a);
  size_t i = 0;
  bool neg = false;
  if (t[i] == '+' || t[i] == '-') neg = t[i++] == '-';
  if (t.compare(i, std::string::npos, "NaN") == 0) {
    *out = std::numeric_limits<double>::quiet_NaN();
    return true;
  }
  if (t.compare(i, std::string::npos, "Infinity") == 0) {
    *out = neg ? -std::numeric_limits<double>::infinity() : std::numeric_limits<double>::infinity();
    return true;
  }
  const size_t body = i;
  auto digits = [&](bool hex) {
    size_t n = 0;
    while (i < t.size() && (hex ? std::isxdigit((unsigned char)t[i]) : std::isdigit((unsigned char)t[i]))) i++, n++;
    return n;
  };
  const bool hex = i + 1 < t.size() && t[i] == '0' && (t[i + 1] == 'x' || t[i + 1] == 'X');
  if (hex) i += 2;
  size_t nd = digits(hex);
  if (i < t.size() && t[i] == '.') {
    i++;
    nd += digits(hex);
  }
  if (nd == 0) return false;
  const bool has_exp = i < t.size() && (hex ? (t[i] == 'p' || t[i] == 'P') : (t[i] == 'e' || t[i] == 'E'));
  if (has_exp) {
    i++;
    if (i < t.size() && (t[i] == '+' || t[i] == '-')) i++;
    if (digits(false) == 0) return false;
  } else if (hex) {
    return false;  // the binary exponent is mandatory
  }
  const size_t num_end = i;
  if (i < t.size() && (t[i] == 'f' || t[i] == 'F' || t[i] == 'd' || t[i] == 'D')) i++;
  if (i != t.size()) return false;
  static const locale_t c_locale = newlocale(LC_ALL_MASK, "C", (locale_t)0);
  const std::string num = t.substr(body, num_end - body);
  char* end = nullptr;
  const double v = c_locale ? strtod_l(num.c_str(), &end, c_locale) : std::strtod(num.c_str(), &end);
  if (end != num.c_str() + num.size()) return false;
  *out = neg ? -v : v;
  return true;
}

// Double.parseDouble, NumberFormatException -> 0.0 (AutoStopMCMC.java:461-465)
double parse_double_or_zero(const std::string& t) {
  double v = 0.0;
  return java_parse_double(t, &v) ? v : 0.0;
}

// Java Math.min(double, double): NaN if either is NaN
double java_min(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return a < b ? a : b;
}

std::string fmt3(double v) {
  char b[64];
  std::snprintf(b, sizeof b, "%.3f", v);
  return b;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ TraceInfo
TraceInfo::TraceInfo(int chainCount, int device) {
  if (chainCount < 1) throw IllegalArgumentException("chainCount must be positive");
  int32_t rc = asm_create(&ctx_, device, chainCount);
  if (rc != ASM_OK) throw std::runtime_error(std::string("asm_create: ") + asm_last_error(nullptr));
  nTrees_.assign((size_t)chainCount, 0);
  nLines_.assign((size_t)chainCount, 0);
  pendingIds_.resize((size_t)chainCount);
  linesUploaded_.assign((size_t)chainCount, 0);
}

TraceInfo::~TraceInfo() {
  if (enc_) asm_encoder_destroy(enc_);
  if (ctx_) asm_destroy(ctx_);
}

int64_t TraceInfo::cladeCount() const { return enc_ ? asm_encoder_num_clades(enc_) : 0; }

void TraceInfo::setLabels(const std::vector<std::string>& labels) {
  labels_ = labels;
  haveLabels_ = true;
}

void TraceInfo::addLogLine(int chain, const double* values, int n) {
  if (chain < 0 || chain >= chainCount()) throw IllegalArgumentException("chain out of range");
  if (n < 1 || !values) throw IllegalArgumentException("empty log line");
  if (hostLogs_.empty()) hostLogs_.resize(nLines_.size()), hostCols_.assign(nLines_.size(), 0);
  if (hostCols_[(size_t)chain] == 0) hostCols_[(size_t)chain] = n;
  if (hostCols_[(size_t)chain] != n)  // what asm_traces_append would answer when the line is uploaded
    throw IllegalArgumentException("log line with " + std::to_string(n) + " columns, chain " + std::to_string(chain) + " has " +
                                   std::to_string(hostCols_[(size_t)chain]));
  hostLogs_[(size_t)chain].insert(hostLogs_[(size_t)chain].end(), values, values + n);  // asm_traces_append checked n
  nLines_[(size_t)chain]++;
}

bool TraceInfo::readLogLine(int chain, const std::string& sStr) {
  std::vector<std::string> sStrs = split_ws(sStr);
  if (sStr.rfind("#", 0) == 0 || sStrs.size() == 1 || sStrs.empty()) return false;  // :434-440 comment lines
  if (sStr.rfind("Sample", 0) == 0) {                                                // :443-446
    setLabels(sStrs);
    return false;
  }
  std::vector<double> v(sStrs.size());
  for (size_t i = 0; i < sStrs.size(); i++) v[i] = parse_double_or_zero(sStrs[i]);  // :460-466
  addLogLine(chain, v.data(), (int)v.size());
  return true;
}

void TraceInfo::setTaxonCount(int nTaxa) {
  if (haveTaxa_) {
    if (nTaxa != nTaxa_) throw IllegalArgumentException("taxon count already fixed");
    return;
  }
  if (asm_encoder_create(&enc_, nTaxa) != ASM_OK) throw IllegalArgumentException("bad taxon count");
  nTaxa_ = nTaxa;
  haveTaxa_ = true;
  idsScratch_.resize((size_t)std::max(1, nTaxa - 1));
}

// The encoded tree waits on the host (its clade ids: a bit row is only as wide as the dictionary at upload time, so
// rows are built then, all at the final width, and the chains already on the GPU are re-strided at most once).
void TraceInfo::appendIds_(int chain) {
  std::vector<int32_t>& p = pendingIds_[(size_t)chain];
  p.insert(p.end(), idsScratch_.begin(), idsScratch_.begin() + (nTaxa_ - 1));
  nTrees_[(size_t)chain]++;
}

void TraceInfo::flush() {
  for (int c = 0; c < chainCount(); c++) {
    std::vector<int32_t>& p = pendingIds_[(size_t)c];
    if (!p.empty()) {
      const int per = nTaxa_ - 1;
      const int64_t n = (int64_t)(p.size() / (size_t)per);
      const int64_t D = cladeCount();
      if (D <= 0xFFFF) {  // the ids themselves go up (half the bytes of the bit rows); the GPU builds the rows
        ids16Scratch_.resize(p.size());
        for (size_t k = 0; k < p.size(); k++) ids16Scratch_[k] = p[k] < 0 ? (uint16_t)ASM_NO_CLADE : (uint16_t)p[k];
        GPU_CHECK(ctx_, asm_trees_append_ids(ctx_, c, n, per, ids16Scratch_.data(), (int32_t)D));
      } else {
        const int words = asm_encoder_words(enc_);
        bitsScratch_.resize((size_t)words * (size_t)n);
        if (asm_encode_bits(p.data(), n, per, words, bitsScratch_.data()) != ASM_OK) throw std::runtime_error("asm_encode_bits failed");
        GPU_CHECK(ctx_, asm_trees_append(ctx_, c, n, words, bitsScratch_.data()));
      }
      p.clear();
    }
    const int64_t up = linesUploaded_[(size_t)c], have = nLines_[(size_t)c];
    if (have > up) {
      const int cols = hostCols_[(size_t)c];
      GPU_CHECK(ctx_, asm_traces_append(ctx_, c, cols, have - up, hostLogs_[(size_t)c].data() + (size_t)up * (size_t)cols));
      linesUploaded_[(size_t)c] = have;
    }
  }
}

void TraceInfo::addTreeNewick(int chain, const std::string& newick, int labelOffset) {
  if (!haveTaxa_) throw IllegalArgumentException("taxon set unknown: no translate block seen yet");
  if (chain < 0 || chain >= chainCount()) throw IllegalArgumentException("chain out of range");
  if (asm_encoder_add_newick(enc_, newick.c_str(), labelOffset, idsScratch_.data()) != ASM_OK)
    throw IllegalArgumentException("cannot parse tree: " + newick.substr(0, 60));
  appendIds_(chain);
}

void TraceInfo::addTree(int chain, const int32_t* left, const int32_t* right, int32_t root) {
  if (!haveTaxa_) throw IllegalArgumentException("taxon count unknown");
  if (chain < 0 || chain >= chainCount()) throw IllegalArgumentException("chain out of range");
  if (asm_encoder_add_topologies(enc_, 1, left, right, &root, idsScratch_.data()) != ASM_OK)
    throw IllegalArgumentException("not a binary tree on the taxon set");
  appendIds_(chain);
}

bool TraceInfo::readTreeLogLine(int chain, const std::string& sStr) {
  if (inTranslate_) {  // :505-519
    if (sStr.find(';') != std::string::npos) {
      inTranslate_ = false;
      setTaxonCount(translateCount_);  // :520
    } else {
      std::string t = trim(sStr);
      if (!t.empty()) translateCount_++;  // the taxon label itself is not needed for clade identity
    }
    return false;
  }
  if (!haveTaxa_) {
    std::string low = trim(sStr);
    std::transform(low.begin(), low.end(), low.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (low == "translate") {  // :503
      inTranslate_ = true;
      translateCount_ = 0;
      return false;
    }
  }
  if (sStr.rfind("tree STATE", 0) != 0) return false;  // :522 ignore non-tree lines
  size_t p = sStr.find('(');                            // :524
  if (p == std::string::npos) return false;
  addTreeNewick(chain, sStr.substr(p), 1);  // :525-536, offset 1 (:527)
  return true;
}

const std::vector<int>& TraceInfo::getMap() {  // :52-68
  if (haveMap_) return map_;
  if (haveLabels_) {
    setUpMap(haveTracesString_ ? tracesString_ : std::string("posterior,likelihood,prior"));
    return map_;
  }
  traces_ = {"posterior", "likelihood", "prior"};
  static const std::vector<int> DEFAULT_MAP = {1, 2, 3};  // :46
  return DEFAULT_MAP;
}

void TraceInfo::setUpMap(const std::string& tracesString) {  // :70-88
  if (!haveLabels_) {
    tracesString_ = tracesString;
    haveTracesString_ = true;
    return;
  }
  traces_.clear();
  std::stringstream ss(tracesString);
  std::string tok;
  while (std::getline(ss, tok, ',')) traces_.push_back(tok);
  map_.assign(traces_.size(), -1);
  for (size_t k = 0; k < traces_.size(); k++) {
    std::string want = trim(traces_[k]);
    for (size_t i = 0; i < labels_.size(); i++)
      if (labels_[i] == want) {
        map_[k] = (int)i;
        break;
      }
    if (map_[k] == -1) {
      haveMap_ = false;
      throw IllegalArgumentException("Could not find label " + traces_[k] + " in trace log.");  // :82-85
    }
  }
  haveMap_ = true;
}

const std::vector<std::string>& TraceInfo::getTraceLabels() {  // :90-96
  if (traces_.empty()) getMap();
  return traces_;
}

// ------------------------------------------------------------------------------------------------ TreeESS
void TreeESS::initAndValidate() {  // TreeESS.java:45-76, as written
  smoothing = smoothingInput;
  // `!(compare(smoothing, 0.0) < 0 && compare(1.0, smoothing) > 0)` holds for every smoothing >= 0 (:47)
  if (!(smoothing < 0.0 && 1.0 > smoothing))
    throw IllegalArgumentException("smoothing should be between 0 and 1, not " + std::to_string(smoothing));
  targetESS = targetESSInput;
  if (targetESS <= 0) throw IllegalArgumentException("targetESS should be positive");
  cacheLimit = cacheLimitInput;
  if (cacheLimit <= 0) throw IllegalArgumentException("cacheLimit should be positive");
  if (cacheLimit <= targetESS) throw IllegalArgumentException("cacheLimit should be at least as big as targetESS");
  N = ESSSampleSizeInput;
  if (N <= 0) throw IllegalArgumentException("sampleSize should be positive");
  if (N >= targetESS) throw IllegalArgumentException("sampleSize should be less than targetESS");
}

void TreeESS::setup(int nChains, TraceInfo* ti) {  // :189-199 / TreePSRF.java:304-313
  traceInfo = ti;
  numChains = nChains;
  if (nChains != 2 && !(allowManyChains && nChains > 2 && nChains <= ASM_MAX_CHAINS))
    throw IllegalArgumentException("Only 2 chains can be handled by " + header() + ", not " + std::to_string(nChains));
  if (ti->chainCount() != nChains) throw IllegalArgumentException("TraceInfo holds a different number of chains");
}

bool TreeESS::converged(const int* burnin, int end) {  // :79-112
  lastLog.clear();
  end = end - end % delta;  // :84
  if (end / delta > cacheLimit) {  // :92-101
    delta *= 2;
    lastLog += "TreeESS Delta=" + std::to_string(delta) + " ";
    return converged(burnin, end);
  }
  int start = (int)(end * (1.0 - smoothing));  // :103
  start = start - start % delta;               // :104
  for (int i = 0; i < numChains; i++)          // :106-110
    if (pseudoESS(i, start, end) < targetESS) return false;
  return true;
}

double TreeESS::pseudoESS(int treeSet, int cutStart, int cutEnd) {  // :117-166
  if (!haveIndices) {  // :120-126
    indices.assign((size_t)N, 0);
    for (int i = 0; i < N; i++) {
      int offset = (cutEnd - cutStart) * i / N;
      offset = offset - offset % delta;
      indices[(size_t)i] = cutStart + offset;
    }
    haveIndices = true;
  } else {  // :127-132
    while (indices[0] < cutStart) {
      std::rotate(indices.begin(), indices.begin() + 1, indices.end());
      indices[(size_t)N - 1] = cutEnd - cutEnd % delta;
    }
  }
  const int len = (cutEnd - cutStart) / delta;  // :137
  if (len <= 0) return std::numeric_limits<double>::quiet_NaN();
  asm_ctx* ctx = traceInfo->ctx();
  std::vector<uint16_t> d((size_t)(cutEnd - cutStart));
  std::vector<double> trace((size_t)N * (size_t)len);
  for (int j = 0; j < N; j++) {  // :139-148, one reference tree at a time
    GPU_CHECK(ctx, asm_rf_block(ctx, treeSet, indices[(size_t)j], indices[(size_t)j] + 1, treeSet, cutStart, cutEnd,
                                methodInput == ASM_RF_FP4 ? ASM_RF_I8 : methodInput, d.data()));  // fp4 has no distance dump
    int k = 0;
    for (int i = cutStart; i < cutEnd && k < len; i += delta)
      trace[(size_t)j * len + k++] = (i != indices[(size_t)j]) ? (double)((float)d[(size_t)(i - cutStart)] + 1) : 0.0;
  }
  std::vector<double> ess((size_t)N);  // :151-156: ESS.calcESS(trace[j], 1) = len / ACT
  GPU_CHECK(ctx, asm_ess_batch(ctx, N, len, len, trace.data(), ASM_MAX_LAG, nullptr, ess.data()));
  double sum = 0;
  for (double e : ess) sum += e;  // :158-161
  const double meanESS = sum / N;  // :162
  char b[64];
  std::snprintf(b, sizeof b, "pseudoESS = %.1f ", meanESS);
  lastLog += b;
  return meanESS;
}

// ------------------------------------------------------------------------------------------------ TreePSRF
void TreePSRF::initAndValidate() {  // TreePSRF.java:54-93
  smoothing = smoothingInput;
  if (!(smoothing >= 0.0 && 1.0 >= smoothing))  // :56-58 (NaN fails both compares in Java as well)
    throw IllegalArgumentException("smoothing should be between 0 and 1, not " + std::to_string(smoothing));
  targetESS = targetESSInput;
  if (targetESS <= 0) throw IllegalArgumentException("targetESS should be positive, not " + std::to_string(targetESS));
  cacheLimit = cacheLimitInput;
  if (cacheLimit <= 0) throw IllegalArgumentException("cacheLimit should be positive, not " + std::to_string(cacheLimit));
  if (cacheLimit <= targetESS)
    throw IllegalArgumentException("cacheLimit (" + std::to_string(cacheLimit) + ") should be at least as big as targetESS");
  N = ESSSampleSizeInput;
  if (N <= 0) throw IllegalArgumentException("sampleSize should be positive, not " + std::to_string(N));
  if (N >= targetESS) throw IllegalArgumentException("sampleSize should be less than targetESS");
  twoSided = twoSidedInput;
  checkESS = checkESSInput;
  const double b = bInput;
  upper = 1.0 + b;  // :89-91
  lower = 1.0 - b;
}

void TreePSRF::setup(int nChains, TraceInfo* ti) {  // :304-318
  TreeESS::setup(nChains, ti);
  grValues.assign(nChains == 2 ? 2 : (size_t)nChains * nChains, -2.0);  // :315-317
  means_.assign((size_t)nChains * nChains, 0.0);
}

std::map<std::string, double> TreePSRF::getLog() const {  // :320-328
  std::map<std::string, double> m;
  for (int i = 0; i < numChains && i < (int)grValues.size(); i++) m["GRT-" + std::to_string(i)] = grValues[(size_t)i];
  return m;
}

bool TreePSRF::converged(const int* burnin, int end) {  // :96-105
  if (numChains != 2) {
    prev = convergedMulti(burnin, end);
    return prev;
  }
  if (twoSided) {
    prev = converged2sided(burnin, end);
    return prev;
  } else {
    prev = converged1sided(0, burnin, end);
    return prev;
  }
}

// One device call replaces both row loops (:141-144 and :154-156): psrf_mean[a][b] for every ordered pair.
void TreePSRF::evalMeans(const int* burnin, int start, int end) {
  asm_ctx* ctx = traceInfo->ctx();
  GPU_CHECK(ctx, asm_psrf_eval(ctx, burnin, start, end, delta, methodInput, nullptr, means_.data()));
}

bool TreePSRF::converged2sided(const int* burnin, int end) {  // :107-197
  threw = false;
  try {
    if (end % delta != 0) return prev;           // :113-116
    int start = (int)(end * (1.0 - smoothing));  // :117
    start = start - start % delta;               // :118
    if (end / delta > cacheLimit) {              // :120-124
      delta *= 2;
      lastLog += "Delta=" + std::to_string(delta) + " ";
      return converged(burnin, end);
    }
    lastLog.clear();
    evalMeans(burnin, start, end);
    const double psrf1mean = means_[0 * 2 + 1];  // :141-145
    grValues[0] = psrf1mean;                     // :146
    double psrf2mean = 0;                        // :147
    lastLog += "psrf1mean = " + fmt3(psrf1mean) + " ";
    if (lower < psrf1mean && psrf1mean < upper) {  // :150
      if (start0 < 0) start0 = start;              // :151-153
      psrf2mean = means_[1 * 2 + 0];               // :154-157
      grValues[1] = psrf2mean;                     // :158
      lastLog += "psrf2mean = " + fmt3(psrf2mean) + " ";
      if (lower < psrf2mean && psrf2mean < upper) {  // :161
        if (!checkESS) return true;                  // :162-164
        if ((end - start0) <= targetESS) {           // :168-171
          lastLog += "not enough samples ";
          return false;
        }
        // :176 `checkESS || pseudoESS(...) && ...` -- checkESS is true here, pseudoESS is never evaluated
        return true;
      }
    }
    if (lower <= psrf1mean || psrf1mean >= upper || lower <= psrf2mean || psrf2mean >= upper) {  // :184-187
      lastLog += "reset start ";
      start0 = -1;
    }
  } catch (const std::exception& e) {  // :188-194 catch (Throwable)
    lastLog += std::string("Programmer error: Something is WRONG and needs fixing: ") + e.what();
    start0 = -1;
    threw = true;
    return false;
  }
  lastLog += std::to_string(start0) + " ";  // :195
  return false;
}

bool TreePSRF::converged1sided(int side, const int* burnin, int end) {  // :200-261
  threw = false;
  try {
    if (end % delta != 0) return prev;           // :206-209
    int start = (int)(end * (1.0 - smoothing));  // :210
    start = start - start % delta;               // :211
    if (end / delta > cacheLimit) {              // :213-217
      delta *= 2;
      return converged(burnin, end);
    }
    lastLog.clear();
    evalMeans(burnin, start, end);
    const double psrf1mean = means_[(size_t)(side - 0) * 2 + (1 - side)];  // :219-223
    grValues[0] = psrf1mean;                                               // :224
    if (lower < psrf1mean && psrf1mean < upper) {                          // :227
      if (!checkESS) return true;                                          // :228-230
      if (start0 < 0) start0 = start;                                      // :232-234
      if ((end - start0) <= targetESS) return false;                       // :238-240
      int cutStart = start0;
      cutStart = cutStart - cutStart % delta;  // :241-242
      const int cutEnd = end;
      if (pseudoESS(side - 0, cutStart, cutEnd) >= targetESS) return true;  // :244-246
      return false;
    }
    if (lower >= psrf1mean || psrf1mean >= upper) start0 = -1;  // :249-251
  } catch (const std::exception& e) {  // :252-258
    lastLog += std::string("Programmer error: Something is WRONG and needs fixing: ") + e.what();
    start0 = -1;
    threw = true;
    return false;
  }
  return false;
}

// C > 2: the two-chain statistic for every ordered chain pair (tools/TreeConvergenceCheck.java:52-77 does
// the same over files); converged iff every mean lies strictly inside (lower, upper).
bool TreePSRF::convergedMulti(const int* burnin, int end) {
  threw = false;
  try {
    const int C = numChains;
    if (end % delta != 0) return prev;
    int start = (int)(end * (1.0 - smoothing));
    start = start - start % delta;
    if (end / delta > cacheLimit) {
      delta *= 2;
      return converged(burnin, end);
    }
    lastLog.clear();
    evalMeans(burnin, start, end);
    bool all_in = true;
    for (int a = 0; a < C; a++)
      for (int b = 0; b < C; b++) {
        if (a == b) continue;
        const double m = means_[(size_t)a * C + b];
        grValues[(size_t)a * C + b] = m;
        if (!(lower < m && m < upper)) all_in = false;
      }
    if (all_in) {
      if (start0 < 0) start0 = start;
      if (!checkESS) return true;
      if ((end - start0) <= targetESS) return false;
      return true;
    }
    start0 = -1;
  } catch (const std::exception& e) {
    lastLog += std::string("Programmer error: Something is WRONG and needs fixing: ") + e.what();
    start0 = -1;
    threw = true;
    return false;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ TraceESS
void TraceESS::setup(int nChains_, TraceInfo* ti) {  // :182-188
  traceInfo = ti;
  nChains = nChains_;
  ti->setUpMap(tracesInput);
}

bool TraceESS::converged(const int* burnin, int end, bool useMapping) {  // :38-69
  const std::vector<int>& map = traceInfo->getMap();  // :51
  std::vector<int> cols(map.size());
  for (size_t j = 0; j < map.size(); j++) cols[j] = useMapping ? map[j] : (int)j;  // :54
  currentESSs.assign(map.size(), 0.0);                                              // :52
  asm_ctx* ctx = traceInfo->ctx();
  // :55-62 -- chain concatenation, ACT and ESS of every selected column in one device call
  int32_t rc = asm_ess_eval(ctx, burnin, end, cols.data(), (int)cols.size(), currentESSs.data());
  if (rc != ASM_OK) gpu_fail(ctx, rc, "asm_ess_eval");  // the Java would throw IndexOutOfBounds out of converged()
  double minESS = std::numeric_limits<double>::infinity();                // :50
  for (double ess : currentESSs) minESS = java_min(ess, minESS);         // :63
  lastMinESS = minESS;
  return minESS >= (double)(targetESS * nChains);  // :68
}

std::map<std::string, double> TraceESS::getLogMap() {  // :191-198
  const std::vector<std::string>& traces = traceInfo->getTraceLabels();
  std::map<std::string, double> m;
  for (size_t i = 0; i < traces.size() && i < currentESSs.size(); i++) m[traces[i]] = currentESSs[i];
  return m;
}

// ------------------------------------------------------------------------------------------------ GelmanRubin
bool GelmanRubin::converged(const int*, int available) {  // GelmanRubin.java:32-51
  lastMaxGR = 0;
  if (nChains < 2) return true;  // no pair to look at (:40-41): true whatever `available` is, nothing is read
  asm_ctx* ctx = traceInfo->ctx();
  int32_t nItems = 0;
  asm_traces_count(ctx, 0, nullptr, &nItems);
  std::vector<double> sums((size_t)nChains * nItems), sumsq((size_t)nChains * nItems);
  // the two accumulation loops of calcGRStat (:61-73) for every chain and item, in Java order, on the device;
  // sampleCount > trace.size() is the reference's IllegalArgumentException (:56-58)
  int32_t rc = asm_trace_moments(ctx, available, sums.data(), sumsq.data());
  if (rc == ASM_E_RANGE) throw IllegalArgumentException("Expected traces of sufficient length");
  if (rc != ASM_OK) gpu_fail(ctx, rc, "asm_trace_moments");
  const int sampleCount = available;
  lastMaxGR = 0;
  bool result = true, decided = false;
  for (int i = 0; i < nChains; i++)
    for (int j = i + 1; j < nChains; j++)
      for (int k = 0; k < nItems; k++) {
        double mean1 = sums[(size_t)i * nItems + k], sumsq1 = sumsq[(size_t)i * nItems + k];
        double mean2 = sums[(size_t)j * nItems + k], sumsq2 = sumsq[(size_t)j * nItems + k];
        mean1 /= sampleCount;  // :67
        mean2 /= sampleCount;  // :73
        const double var1 = (sumsq1 - mean1 * mean1 * sampleCount) / (sampleCount - 1);  // :76
        const double var2 = (sumsq2 - mean2 * mean2 * sampleCount) / (sampleCount - 1);  // :77
        const double fW = (var1 + var2) / 2;                                             // :80
        double GRstat;
        if (fW == 0) {
          GRstat = 1;  // :81-83
        } else {
          const double totalMean = (mean1 + mean2) / 2;            // :86
          const double totalSq = mean1 * mean1 + mean2 * mean2;    // :87
          const double fB = (totalSq - totalMean * totalMean * 2);  // :90
          const double varR = ((sampleCount - 1.0) / sampleCount) + (fB / fW) * (1.0 / sampleCount);  // :93
          GRstat = std::sqrt(varR);                                // :94
        }
        if (GRstat > lastMaxGR) lastMaxGR = GRstat;
        if (!decided && GRstat > acceptedThreshold) {  // :44-46 (the Java returns here)
          result = false;
          decided = true;
        }
      }
  return result;
}

// ------------------------------------------------------------------------------------------------ CladeDifference
bool CladeDifference::converged(const int*, int available) {  // CladeDifference.java:41-65
  asm_ctx* ctx = traceInfo->ctx();
  for (int c = 0; c < nChains; c++)  // trees[j].get(i) for i < available (:48-52) throws past the end
    if (available > traceInfo->treeCount(c)) throw std::out_of_range("CladeDifference: available past the stored trees");
  int32_t words = 0;
  asm_trees_count(ctx, 0, nullptr, &words);
  const size_t D = (size_t)words * 64;
  std::vector<std::vector<int32_t>> counts((size_t)nChains, std::vector<int32_t>(D ? D : 1, 0));
  for (int c = 0; c < nChains && D; c++)  // the clade maps of process() (:147-159) = column sums of the bit matrix
    GPU_CHECK(ctx, asm_clade_counts(ctx, c, 0, available, counts[(size_t)c].data()));
  const int m_nClades = available > 0 ? traceInfo->taxonCount() : 1;  // :25, :149 (root clade length = taxa)
  auto java_max = [](double a, double b) { return a != a ? a : (b != b ? b : (a > b ? a : b)); };
  double fMaxCladeProbDiff = 0;
  for (int i = 0; i < nChains; i++)
    for (int k = 0; k < nChains; k++) {
      double diff = 0;
      if (i != k) {  // calcMaxCladeDifference :85-103
        int nTotal = 0, nMax = 0;
        for (size_t id = 0; id < D; id++) {
          const int i1 = counts[(size_t)i][id];
          if (i1 == 0) continue;  // map1.keySet()
          const int i2 = counts[(size_t)k][id];
          nTotal += i1;
          nMax = std::max(nMax, std::abs(i1 - i2));
        }
        diff = nMax / (double)(nTotal / m_nClades);  // int division inside, :102
      }
      fMaxCladeProbDiff = java_max(fMaxCladeProbDiff, diff);  // :60
    }
  m_fMaxCladeProbDiffs.push_back(fMaxCladeProbDiff);  // :63
  return fMaxCladeProbDiff < acceptedThreshold;       // :64
}

}  // namespace asmgpu

// ================================================================================================
// Flat C wrapper (include/asmgpu_host.h): lets tests written in Python read like the reference's own
// would -- create a TraceInfo, feed it log lines, call setup() / converged().
// ================================================================================================
using namespace asmgpu;

struct asmh_criterion {
  MCMCConvergenceCriterion* c = nullptr;
  TreePSRF* psrf = nullptr;
  TraceESS* tess = nullptr;
  TreeESS* tree = nullptr;
  GelmanRubin* gr = nullptr;
  CladeDifference* cd = nullptr;
};


// ---- BurnInDetector (BurnInDetector.java:12-80) ------------------------------------------------------------
std::vector<int> asmgpu::BurnInDetector::burnIn(int end) {
  const int chainCount = traceInfo->chainCount();
  std::vector<int> burnin((size_t)chainCount, 0);
  if (burnInStrat.kind == BurnInStrategy::Automatic) {
    for (int i = 0; i < chainCount; i++) burnin[(size_t)i] = burnIn(i, end);
  } else {
    const double burnPercent = (double)burnInStrat.burnInPercent / 100;  // :32
    std::fill(burnin.begin(), burnin.end(), (int)std::ceil(burnPercent * end));
  }
  return burnin;
}

int asmgpu::BurnInDetector::burnIn(int chainNr, int end) {
  int maxBurnin = 0;
  const std::vector<int>& map = traceInfo->getMap();
  for (size_t j = 0; j < map.size(); j++) maxBurnin = std::max(burnIn(chainNr, map[j], end), maxBurnin);
  return maxBurnin;
}

int asmgpu::BurnInDetector::burnIn(int chainNr, int col, int end) {
  if (end > traceInfo->logCount(chainNr)) throw std::out_of_range("IndexOutOfBoundsException: burn-in past the end of the trace");
  // mean and stdev of the last 25 % (:52-64), sequential sums as written
  double m2 = 0, sq2 = 0;
  const int lb = 3 * end / 4;
  for (int i = lb; i < end; i++) m2 += traceInfo->logValue(chainNr, col, i);
  m2 = m2 / (end - lb);
  for (int i = lb; i < end; i++) {
    const double d = traceInfo->logValue(chainNr, col, i);
    sq2 += (d - m2) * (d - m2);
  }
  const double stdev2 = std::sqrt(sq2 / (end - lb - 1));
  // first moving average inside m2 +/- stdev2 (:67-76)
  for (int i = 0; i + WINDOW_SIZE < lb; i++) {
    double sum = 0;
    for (int j = 0; j < WINDOW_SIZE; j++) sum += traceInfo->logValue(chainNr, col, i + j);
    sum /= WINDOW_SIZE;
    if (m2 - stdev2 < sum && sum < m2 + stdev2) return i + WINDOW_SIZE;
  }
  return lb;
}

// ---- AutoStopMCMC.combinelogs (AutoStopMCMC.java:247-317) ------------------------------------------------------
// The output format on the far side of the path: once the criteria have agreed, the post-burn-in samples of every
// chain's log are appended to one file and renumbered 0, delta, 2*delta, ...  BEAST's Logger (un-vendored) writes
// the header (logger.init(), :254) and the footer; here the caller passes both as text.
namespace {

// BufferedReader.readLine: a line ends at \n, \r or \r\n; returns false at end of stream (Java: null)
bool java_read_line(std::istream& in, std::string& out) {
  out.clear();
  int c = in.get();
  if (c == EOF) return false;
  for (; c != EOF; c = in.get()) {
    if (c == '\n') return true;
    if (c == '\r') {
      if (in.peek() == '\n') in.get();
      return true;
    }
    out.push_back((char)c);
  }
  return true;
}
bool java_ready(std::istream& in) { return in.peek() != EOF; }  // BufferedReader.ready() on a file
bool is_java_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\x0B' || c == '\f' || c == '\r'; }
bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

// str.replaceFirst("[0-9]+\t", "") (:308): the leftmost run of digits that is followed by a tab, wherever it is
std::string drop_first_number_tab(const std::string& s) {
  for (size_t i = 0; i < s.size();) {
    if (s[i] < '0' || s[i] > '9') {
      i++;
      continue;
    }
    size_t j = i;
    while (j < s.size() && s[j] >= '0' && s[j] <= '9') j++;
    if (j < s.size() && s[j] == '\t') return s.substr(0, i) + s.substr(j + 1);
    i = j;
  }
  return s;
}

struct NullPointer : std::out_of_range {  // str.matches(...) / str.replaceFirst(...) on the null of a finished stream
  using std::out_of_range::out_of_range;
};

}  // namespace

int64_t asmgpu::combineLogs(bool treeMode, const std::vector<std::string>& chainFiles, const int* burnIn, int lastReported,
                            int64_t sampleDelta, std::ostream& out) {
  int64_t sampleNr = 0, written = 0;  // :252
  for (size_t i = 0; i < chainFiles.size(); i++) {
    std::ifstream fin(chainFiles[i], std::ios::binary);
    if (!fin) throw std::runtime_error("FileNotFoundException: " + chainFiles[i]);
    std::string str;
    bool have = false;  // Java: str == null
    if (treeMode) {
      int skipped = 0;  // :266-272: burn-in counts trees, other lines are passed over
      while (java_ready(fin) && skipped < burnIn[i]) {
        have = java_read_line(fin, str);
        if (starts_with(str, "tree STATE")) skipped++;  // str.matches("^tree STATE.*")
      }
      for (int j = burnIn[i]; j < lastReported; j++) {  // :276-284: one LINE per j, printed only if it is a tree
        if (!java_read_line(fin, str)) throw NullPointer("NullPointerException: " + chainFiles[i] + " ended before sample " + std::to_string(j));
        if (starts_with(str, "tree STATE")) {
          if (starts_with(str, "tree STATE_")) {  // replaceAll("^tree STATE_[^\\s]*", "")
            size_t k = 11;
            while (k < str.size() && !is_java_space(str[k])) k++;
            str = str.substr(k);
          }
          out << "tree STATE_" << sampleNr << str << "\n";
          sampleNr += sampleDelta;
          written++;
        }
      }
    } else {
      while (java_ready(fin) && (!have || !starts_with(str, "Sample"))) have = java_read_line(fin, str);  // :292-294
      int skipped = 0;  // :297-301
      while (java_ready(fin) && skipped < burnIn[i]) {
        java_read_line(fin, str);
        skipped++;
      }
      for (int j = burnIn[i]; j < lastReported; j++) {  // :305-310
        if (!java_read_line(fin, str)) throw NullPointer("NullPointerException: " + chainFiles[i] + " ended before sample " + std::to_string(j));
        out << sampleNr << "\t" << drop_first_number_tab(str) << "\n";
        sampleNr += sampleDelta;
        written++;
      }
    }
  }
  return written;
}

static thread_local std::string g_host_err;
static int32_t host_fail(const std::exception& e) {
  g_host_err = e.what();
  if (dynamic_cast<const IllegalArgumentException*>(&e)) return ASM_E_INVALID;
  if (dynamic_cast<const std::out_of_range*>(&e)) return ASM_E_RANGE;
  return ASM_E_CUDA;
}
#define HOST_TRY(body)            \
  try {                           \
    body;                         \
    return ASM_OK;                \
  } catch (const std::exception& e) { \
    return host_fail(e);          \
  }

extern "C" {

const char* asmh_last_error(void) { return g_host_err.c_str(); }

int32_t asmh_trace_info_create(asmh_trace_info** out, int32_t chain_count, int32_t device) {
  if (!out) return ASM_E_INVALID;
  try {
    *out = reinterpret_cast<asmh_trace_info*>(new TraceInfo(chain_count, device));
    return ASM_OK;
  } catch (const IllegalArgumentException& e) {
    g_host_err = e.what();
    return ASM_E_INVALID;
  } catch (const std::exception& e) {
    g_host_err = e.what();
    return ASM_E_NO_DEVICE;
  }
}
int32_t asmh_trace_info_destroy(asmh_trace_info* ti) {
  delete reinterpret_cast<TraceInfo*>(ti);
  return ASM_OK;
}
#define TI(p) (reinterpret_cast<TraceInfo*>(p))
int32_t asmh_read_log_line(asmh_trace_info* ti, int32_t chain, const char* line, int32_t* appended) {
  HOST_TRY({
    bool r = TI(ti)->readLogLine(chain, line);
    if (appended) *appended = r;
  })
}
double asmh_parse_log_cell(const char* token, int32_t* is_number) {
  double v = 0.0;
  const bool ok = token && java_parse_double(token, &v);
  if (is_number) *is_number = ok;
  return ok ? v : 0.0;
}
int32_t asmh_flush(asmh_trace_info* ti) {
  HOST_TRY({ TI(ti)->flush(); })
}
int32_t asmh_read_tree_log_line(asmh_trace_info* ti, int32_t chain, const char* line, int32_t* appended) {
  HOST_TRY({
    bool r = TI(ti)->readTreeLogLine(chain, line);
    if (appended) *appended = r;
  })
}
int32_t asmh_set_taxon_count(asmh_trace_info* ti, int32_t n_taxa) { HOST_TRY(TI(ti)->setTaxonCount(n_taxa)) }
int32_t asmh_add_tree(asmh_trace_info* ti, int32_t chain, const int32_t* left, const int32_t* right, int32_t root) {
  HOST_TRY(TI(ti)->addTree(chain, left, right, root))
}
int32_t asmh_add_log_line(asmh_trace_info* ti, int32_t chain, const double* values, int32_t n) {
  HOST_TRY(TI(ti)->addLogLine(chain, values, n))
}
int32_t asmh_set_labels(asmh_trace_info* ti, const char* whitespace_separated) {
  HOST_TRY(TI(ti)->setLabels(split_ws(whitespace_separated)))
}
int32_t asmh_get_map(asmh_trace_info* ti, int32_t* map_out, int32_t cap, int32_t* n_out) {
  HOST_TRY({
    const std::vector<int>& m = TI(ti)->getMap();
    if (n_out) *n_out = (int32_t)m.size();
    for (size_t i = 0; i < m.size() && (int32_t)i < cap; i++) map_out[i] = m[i];
  })
}
int64_t asmh_tree_count(const asmh_trace_info* ti, int32_t chain) { return reinterpret_cast<const TraceInfo*>(ti)->treeCount(chain); }
int64_t asmh_log_count(const asmh_trace_info* ti, int32_t chain) { return reinterpret_cast<const TraceInfo*>(ti)->logCount(chain); }
int64_t asmh_clade_count(const asmh_trace_info* ti) { return reinterpret_cast<const TraceInfo*>(ti)->cladeCount(); }
asm_ctx* asmh_ctx(asmh_trace_info* ti) { return TI(ti)->ctx(); }

int32_t asmh_tree_psrf_create(asmh_criterion** out, double b, int32_t two_sided, int32_t check_ess, int32_t target_ess,
                              double smoothing, int32_t cache_limit, int32_t sample_size, int32_t method) {
  if (!out) return ASM_E_INVALID;
  auto* p = new TreePSRF();
  p->bInput = b;
  p->twoSidedInput = two_sided != 0;
  p->checkESSInput = check_ess != 0;
  p->targetESSInput = target_ess;
  p->smoothingInput = smoothing;
  p->cacheLimitInput = cache_limit;
  p->ESSSampleSizeInput = sample_size;
  p->methodInput = method;
  try {
    p->initAndValidate();
  } catch (const std::exception& e) {
    delete p;
    return host_fail(e);
  }
  auto* h = new asmh_criterion();
  h->c = p;
  h->psrf = p;
  h->tree = p;
  *out = h;
  return ASM_OK;
}
int32_t asmh_tree_ess_create(asmh_criterion** out, int32_t target_ess, double smoothing, int32_t cache_limit,
                             int32_t sample_size, int32_t method) {
  if (!out) return ASM_E_INVALID;
  auto* p = new TreeESS();
  p->targetESSInput = target_ess;
  p->smoothingInput = smoothing;
  p->cacheLimitInput = cache_limit;
  p->ESSSampleSizeInput = sample_size;
  p->methodInput = method;
  try {
    p->initAndValidate();
  } catch (const std::exception& e) {
    delete p;
    return host_fail(e);
  }
  auto* h = new asmh_criterion();
  h->c = p;
  h->tree = p;
  *out = h;
  return ASM_OK;
}
int32_t asmh_trace_ess_create(asmh_criterion** out, int32_t target_ess, const char* traces) {
  if (!out) return ASM_E_INVALID;
  auto* p = new TraceESS();
  p->targetESSInput = target_ess;
  if (traces) p->tracesInput = traces;
  p->initAndValidate();
  auto* h = new asmh_criterion();
  h->c = p;
  h->tess = p;
  *out = h;
  return ASM_OK;
}
int32_t asmh_gelman_rubin_create(asmh_criterion** out, double threshold) {
  if (!out) return ASM_E_INVALID;
  auto* p = new GelmanRubin();
  p->acceptedThresholdInput = threshold;
  p->initAndValidate();
  auto* h = new asmh_criterion();
  h->c = p;
  h->gr = p;
  *out = h;
  return ASM_OK;
}
int32_t asmh_clade_difference_create(asmh_criterion** out, double threshold) {
  if (!out) return ASM_E_INVALID;
  auto* p = new CladeDifference();
  p->acceptedThresholdInput = threshold;
  p->initAndValidate();
  auto* h = new asmh_criterion();
  h->c = p;
  h->cd = p;
  *out = h;
  return ASM_OK;
}
/* the value the last call logged: GelmanRubin -> largest statistic; CladeDifference -> fMaxCladeProbDiff (:63) */
int32_t asmh_criterion_last_value(asmh_criterion* c, double* out) {
  if (!c || !out) return ASM_E_INVALID;
  if (c->gr) *out = c->gr->lastMaxGR;
  else if (c->cd) *out = c->cd->m_fMaxCladeProbDiffs.empty() ? 0.0 : c->cd->m_fMaxCladeProbDiffs.back();
  else return ASM_E_INVALID;
  return ASM_OK;
}
int32_t asmh_criterion_destroy(asmh_criterion* c) {
  if (c) {
    delete c->c;
    delete c;
  }
  return ASM_OK;
}
int32_t asmh_criterion_setup(asmh_criterion* c, int32_t n_chains, asmh_trace_info* ti) { HOST_TRY(c->c->setup(n_chains, TI(ti))) }
int32_t asmh_criterion_converged(asmh_criterion* c, const int32_t* burnin, int32_t end, int32_t* converged) {
  HOST_TRY({
    bool r = c->c->converged(burnin, end);
    if (converged) *converged = r;
  })
}
const char* asmh_criterion_header(asmh_criterion* c) {
  static thread_local std::string s;
  s = c->c->header();
  return s.c_str();
}
int32_t asmh_tree_psrf_state(asmh_criterion* c, int32_t* delta, int32_t* start0, int32_t* threw, double* gr_values, int32_t cap,
                             int32_t* n_gr) {
  if (!c || !c->psrf) return ASM_E_INVALID;
  if (delta) *delta = c->psrf->getDelta();
  if (start0) *start0 = c->psrf->getStart0();
  if (threw) *threw = c->psrf->threw;
  const std::vector<double>& g = c->psrf->getGrValues();
  if (n_gr) *n_gr = (int32_t)g.size();
  for (size_t i = 0; i < g.size() && (int32_t)i < cap; i++) gr_values[i] = g[i];
  return ASM_OK;
}
int32_t asmh_trace_ess_current(asmh_criterion* c, double* ess_out, int32_t cap, int32_t* n_out, double* min_out) {
  if (!c || !c->tess) return ASM_E_INVALID;
  const std::vector<double>& e = c->tess->currentESS();
  if (n_out) *n_out = (int32_t)e.size();
  for (size_t i = 0; i < e.size() && (int32_t)i < cap; i++) ess_out[i] = e[i];
  if (min_out) *min_out = c->tess->lastMinESS;
  return ASM_OK;
}
/* BurnInDetector.burnIn(end) (BurnInDetector.java:21-36): strategy 0 = Automatic, 1 = Running(percent) */
int32_t asmh_burn_in(asmh_trace_info* ti, int32_t strategy, int32_t percent, int32_t end, int32_t* burnin_out) {
  if (!ti || !burnin_out || (strategy != 0 && strategy != 1) || end < 0) {
    g_host_err = "burn_in: bad argument";
    return ASM_E_INVALID;
  }
  HOST_TRY({
    BurnInStrategy st = strategy == 1 ? BurnInStrategy::running(percent) : BurnInStrategy::automatic();
    std::vector<int> b = BurnInDetector(TI(ti), st).burnIn(end);
    for (size_t i = 0; i < b.size(); i++) burnin_out[i] = b[i];
  })
}

int32_t asmh_combine_logs(int32_t mode, int32_t n_chains, const char* const* chain_paths, const int32_t* burnin,
                          int32_t last_reported, int64_t sample_delta, const char* header, const char* footer,
                          const char* out_path, int64_t* lines_out) {
  if ((mode != 0 && mode != 1) || n_chains < 0 || (n_chains > 0 && (!chain_paths || !burnin)) || !out_path) {
    g_host_err = "combine_logs: bad argument";
    return ASM_E_INVALID;
  }
  try {
    std::vector<std::string> files;
    for (int i = 0; i < n_chains; i++) files.emplace_back(chain_paths[i] ? chain_paths[i] : "");
    std::ostringstream body;  // nothing reaches out_path if a chain file ends early
    const int64_t n = combineLogs(mode == 1, files, burnin, last_reported, sample_delta, body);
    std::ofstream out(out_path, std::ios::binary | std::ios::trunc);
    if (!out) throw std::runtime_error(std::string("cannot write ") + out_path);
    if (header) out << header;
    out << body.str();
    if (footer) out << footer;
    if (lines_out) *lines_out = n;
    return ASM_OK;
  } catch (const std::exception& e) {
    return host_fail(e);
  }
}

const char* asmh_criterion_last_log(asmh_criterion* c) {
  static thread_local std::string s;
  s = c->tree ? c->tree->lastLog : std::string();
  return s.c_str();
}

}  // extern "C"
