// report_host.hpp -- the report half of elector/computeStats.py: computeMetrics (:519-675), outputMetrics
// (:444-468), outputRecallPrecision (:196-264) and outputReadSizeDistribution (:273-288) fed by the per-record integer counters
// of the device tally (tally_kernel.cuh) instead of a second pass of Python over msa.fa.  The reference needs 14.9 s for the
// 459 reads of its example; what is left here is a loop over records: sums, ratios, Python's float formatting.
//
// Two things are not in the counters and are worked out here from the merged rows of the few records that need them:
//   - the "missing" size of a SPLIT read (several consecutive msa.fa records with one header, :564-615): the reference columns of
//     its last fragment that no fragment's mask keeps (:595-599) -- from the fragments' masks (gapsLeft / gapsRight + the border
//     gap stretches the tally kernel returns) and the last fragment's reference row;
//   - the homopolymer ratio (:298-363 inside indels(), :416-421): the reference resets the list per read (:560), so the figure of
//     the summary is that of the LAST read of the file alone -- one pass over that read's rows.
#pragma once
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/elector_poa.h"

namespace elector {

// str(float) of Python 3: the shortest digits that round-trip, fixed notation for 1e-4 <= |v| < 1e16
inline std::string py_float_str(double v) {
  if (std::isnan(v)) return "nan";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
  std::string s(buf, res.ptr);                       // d[.ddd]e[+-]XX
  const bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  const size_t e = s.find('e');
  std::string digits = s.substr(0, e);
  const int exp10 = atoi(s.c_str() + e + 1);
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  std::string out;
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 < 0) out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    else if ((int)digits.size() <= exp10 + 1) out = digits + std::string((size_t)(exp10 + 1 - (int)digits.size()), '0') + ".0";
    else out = digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
  } else {
    out = digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    char eb[16];
    snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
    out += eb;
  }
  return neg ? "-" + out : out;
}

// round(v, nd) of Python 3: the decimal expansion of the double itself, rounded half to even at nd digits (what glibc's %.*f prints)
inline double py_round(double v, int nd) {
  if (!std::isfinite(v)) return v;
  char buf[512];
  snprintf(buf, sizeof buf, "%.*f", nd, v);
  return strtod(buf, nullptr);
}

// a ratio a / b printed like Python prints `a / b if b != 0 else 0`: the int 0 has no ".0"
struct PyRatio {
  double v = 0; bool is_int_zero = true;
  static PyRatio of(int64_t a, int64_t b) { PyRatio r; if (b != 0) { r.v = (double)a / (double)b; r.is_int_zero = false; } return r; }
  std::string str() const { return is_int_zero ? "0" : py_float_str(v); }
};

// statistics.mean of doubles: the exact sum (Python adds Fractions) divided by n, rounded once.  The values are ratios rounded to two
// decimals (well below 2^60, multiples of 2^-70 as doubles of magnitude >= 2^-17 or zero), so a scaled 128-bit integer holds the sum.
inline double exact_mean(const std::vector<double> &x) {
  if (x.empty()) return 0;
  __int128 sum = 0;
  bool exact = true;
  for (double v : x) {
    const double s = std::ldexp(v, 70);
    if (!(std::fabs(s) < 1.7e38 / (double)x.size()) || s != std::floor(s)) { exact = false; break; }
    int e; const double m = std::frexp(s, &e);             // s = m * 2^e, |m| in [0.5, 1)
    if (s == 0) continue;
    const int64_t mi = (int64_t)std::ldexp(m, 53);         // 53-bit integer mantissa
    sum += (e - 53 >= 0) ? ((__int128)mi << (e - 53)) : ((__int128)mi >> (53 - e));   // s is an integer: the shifted-out bits are zero
  }
  if (!exact) { long double t = 0; for (double v : x) t += v; return (double)(t / (long double)x.size()); }
  // sum / (n * 2^70), rounded to nearest even: long division to 64 significant bits + sticky
  const bool neg = sum < 0;
  unsigned __int128 num = neg ? (unsigned __int128)(-sum) : (unsigned __int128)sum;
  const unsigned __int128 den = (unsigned __int128)x.size();
  if (num == 0) return 0;
  int shift = 0;
  while ((num >> 100) == 0) { num <<= 1; ++shift; }        // keep plenty of quotient bits
  unsigned __int128 q = num / den;
  const bool sticky = (num % den) != 0;
  int qbits = 0;
  for (unsigned __int128 t = q; t; t >>= 1) ++qbits;
  int drop = qbits - 53;
  int exp2 = -70 - shift;
  if (drop > 0) {
    const unsigned __int128 rem = q & (((unsigned __int128)1 << drop) - 1), half = (unsigned __int128)1 << (drop - 1);
    q >>= drop; exp2 += drop;
    if (rem > half || (rem == half && (sticky || (q & 1)))) ++q;
  }
  const double r = std::ldexp((double)(uint64_t)q, exp2);
  return neg ? -r : r;
}

struct ReportRecords {          // the merged records of one msa.fa, in file order
  int64_t n = 0;
  const char *headers = nullptr; const int64_t *header_off = nullptr;   // header text of record i WITHOUT '>' (Donatello's: ends in a blank)
  const int64_t *counters = nullptr;                                    // [n][ELECTOR_TALLY_K]
  const int32_t *stretches = nullptr;                                   // [n][ELECTOR_STRETCH_K]: count, then (first, last) column pairs
  const char *m_ref = nullptr, *m_cor = nullptr; const int64_t *m_off = nullptr;   // rows of record i: NCOLS columns from m_off[i]
  std::string header(int64_t i) const { return std::string(headers + header_off[i], (size_t)(header_off[i + 1] - header_off[i])); }
  int64_t c(int64_t i, int k) const { return counters[i * ELECTOR_TALLY_K + k]; }
  // existingCorrectedPositions of record i (getCorrectedPositions, :712-752, without clips)
  void mask(int64_t i, std::vector<uint8_t> &m) const {
    const int64_t L = c(i, ELECTOR_T_NCOLS), gl = c(i, ELECTOR_T_GAPSLEFT), gr = c(i, ELECTOR_T_GAPSRIGHT);
    m.assign((size_t)L, 1);
    if (gl >= 5) for (int64_t p = 0; p < gl && p < L; ++p) m[(size_t)p] = 0;
    if (gr >= 5) for (int64_t p = L - 1; p > L - gr && p >= 0; --p) m[(size_t)p] = 0;
    const int32_t *s = stretches + i * ELECTOR_STRETCH_K;
    for (int k = 0; k < s[0]; ++k) for (int64_t p = s[1 + 2 * k]; p <= s[2 + 2 * k] && p < L; ++p) if (p >= 0) m[(size_t)p] = 0;
  }
};

// the homopolymer bookkeeping of indels() + getTPFNFP() (:298-363, :416-421) over one record; appends to ratios
inline void homopolymer_ratios(const char *R, const char *C, const std::vector<uint8_t> &mask, int threshold, std::vector<double> &ratios) {
  std::string rep0 = "x", rep1 = "x";
  bool ok_to_report = false, end_ref = false;
  const size_t L = mask.size();
  for (size_t pos = 0; pos < L; ++pos) {
    const char r = R[pos], c = C[pos];
    bool end_res = true, app_r = false, app_c = false;
    if (mask[pos]) {
      if (r != '.') {
        if (r == rep0.back()) { app_r = true; if ((int)rep0.size() + 1 >= threshold) ok_to_report = true; }
        else if (ok_to_report) end_ref = true;
      }
      if (c != '.' && c == rep1.back()) { app_c = true; end_res = false; }
    }
    if (app_c || app_r) { rep0 += r; rep1 += c; }
    else if (!(end_ref && end_res) && !end_ref && r != '.') { rep0.assign(1, r); rep1.assign(1, c); }
    if (end_ref && end_res) {
      // the most frequent letter of the reference side (the reference asks a Python set: ties have no defined winner there; the
      // first letter to reach the count wins here), a gap only when nothing else is there
      auto most = [](const std::string &s, bool skip_dots) {
        char best = 0; size_t cnt = 0;
        for (size_t i = 0; i < s.size(); ++i) {
          if (skip_dots && s[i] == '.') continue;
          const size_t n = (size_t)std::count(s.begin(), s.end(), s[i]);
          if (n > cnt) { cnt = n; best = s[i]; }
        }
        return best;
      };
      char h = most(rep0, false);
      if (h == '.') h = most(rep0, true);
      int cur_r = 0, max_r = 0, cur_c = 0, max_c = 0;
      for (size_t i = 0; i < rep0.size(); ++i) {
        if (rep0[i] == h) ++cur_r; else if (rep0[i] != '.') { max_r = std::max(max_r, cur_r); cur_r = 0; }
        if (rep1[i] == h) ++cur_c; else if (rep1[i] != '.') { max_c = std::max(max_c, cur_c); cur_c = 0; }
      }
      max_r = std::max(max_r, cur_r); max_c = std::max(max_c, cur_c);
      ok_to_report = false; end_ref = false;
      ratios.push_back(py_round((double)max_c * 1.0 / (double)max_r, 2));
      rep0.assign(1, r); rep1.assign(1, c);
    }
  }
}

struct ReportResult {
  elector_report_summary s{};
  std::string per_read_metrics, size_distribution, log_text, stdout_text, error;
};

// computeMetrics + outputRecallPrecision.  corrected_fasta: the corrected reads file of outputReadSizeDistribution (read only when
// there are trimmed or split reads; may be null: then the "sequences" lines are left out and the result says so).
// compensated_sum: Python >= 3.12 adds floats in sum() with Neumaier's compensation (the means differ in the last digit from the
// plain left-to-right sum of the Python versions before, which is what the README example of the reference shows).
inline int report_compute(const ReportRecords &in, int small_reads, int wrongly_cor_reads, double size_threshold, int homopolymer_threshold,
                          const char *corrected_fasta, const char *soft, bool compensated_sum, ReportResult &out) {
  // getSplit (:45-57): `grep ">" | uniq -c` -- runs of identical header lines, three lines per record; a later run of the same
  // header replaces an earlier one
  std::map<std::string, int64_t> reads_to_split;
  auto squeeze = [](std::string h) { h.erase(std::remove_if(h.begin(), h.end(), [](char ch) { return ch == ' ' || ch == '\t'; }), h.end()); return h; };
  auto first_token = [](const std::string &h) { const size_t p = h.find(' '); return p == std::string::npos ? h : h.substr(0, p); };
  for (int64_t i = 0; i < in.n;) {
    int64_t j = i + 1;
    const std::string h = in.header(i);
    while (j < in.n && in.header(j) == h) ++j;
    std::string key = h;
    const size_t gt = key.find('>');
    if (gt != std::string::npos) key.erase(gt);
    reads_to_split[squeeze(key)] = j - i;
    i = j;
  }
  std::vector<double> recall, precision, cor_rate, uncor_cor_rate, gc_ref, gc_cor, ratios;
  std::vector<int64_t> missing_sizes, len_corrected;
  int64_t n_reads = 0, count_split = 0, count_ext = 0, count_trim = 0, ext_bases = 0, thr_unc = 0, total_cor = 0, total_uncor = 0;
  int64_t idc[3] = {0, 0, 0}, idu[3] = {0, 0, 0};
  out.per_read_metrics = "score metric\n";
  std::vector<uint8_t> m;
  for (int64_t i = 0; i < in.n;) {
    std::string hn = in.header(i);
    { const size_t gt = hn.find('>'); if (gt != std::string::npos) hn.erase(gt); }
    hn = first_token(hn);
    const auto it = reads_to_split.find(hn);
    if (it == reads_to_split.end()) { out.error = "header '" + hn + "' is not a key of the split table (computeStats.getSplit needs headers without blanks inside)"; return ELECTOR_EINVAL; }
    const int64_t frags = it->second;
    if (i + (frags > 1 ? frags : 1) > in.n) { out.error = "split read '" + hn + "' runs past the end of the records"; return ELECTOR_EINVAL; }
    ratios.clear();
    const bool last_group = i + (frags > 1 ? frags : 1) >= in.n;   // the list of the last read is the one that counts (:560)
    int64_t sFN = 0, sTP = 0, sFP = 0, sCor = 0, sUncor = 0, sUC = 0, sUU = 0, missing_in_read = 0;
    bool any = false, extended = false;
    double gcr = 0, gcc = 0;
    auto take = [&](int64_t r) {   // gapsAndExtensions + nucleotideMetrics of one assessed record
      if (in.c(r, ELECTOR_T_EXTENDED) >= 0) { extended = true; ext_bases += in.c(r, ELECTOR_T_EXTENDED); }
      idc[0] += in.c(r, ELECTOR_T_INSC); idc[1] += in.c(r, ELECTOR_T_DELC); idc[2] += in.c(r, ELECTOR_T_SUBSC);
      idu[0] += in.c(r, ELECTOR_T_INSU); idu[1] += in.c(r, ELECTOR_T_DELU); idu[2] += in.c(r, ELECTOR_T_SUBSU);
      sFN += in.c(r, ELECTOR_T_FN); sTP += in.c(r, ELECTOR_T_TP); sFP += in.c(r, ELECTOR_T_FP);
      sCor += in.c(r, ELECTOR_T_COR); sUncor += in.c(r, ELECTOR_T_UNCOR); sUC += in.c(r, ELECTOR_T_UNCORCOR); sUU += in.c(r, ELECTOR_T_UNCORUNCOR);
      len_corrected.push_back(in.c(r, ELECTOR_T_LENCOR));
      gcr = py_round((double)in.c(r, ELECTOR_T_GCREF) * 1.0 / (double)in.c(r, ELECTOR_T_LENREF), 3);
      gcc = py_round((double)in.c(r, ELECTOR_T_GCCOR) * 1.0 / (double)in.c(r, ELECTOR_T_LENCOR), 3);
      any = true;
      if (last_group && in.m_ref && in.m_cor) { in.mask(r, m); homopolymer_ratios(in.m_ref + in.m_off[r], in.m_cor + in.m_off[r], m, homopolymer_threshold, ratios); }
    };
    auto output_metrics = [&]() {   // outputMetrics (:444-468)
      if (any) {
        const PyRatio rec = PyRatio::of(sTP, sTP + sFN), prec = PyRatio::of(sTP, sTP + sFP), cr = PyRatio::of(sCor, sCor + sUncor), ur = PyRatio::of(sUC, sUC + sUU);
        if (missing_in_read != 0) missing_sizes.push_back(missing_in_read);
        out.per_read_metrics += rec.str() + " recall\n" + prec.str() + " precision\n" + cr.str() + " correct_rate\n";
        recall.push_back(rec.v); precision.push_back(prec.v); cor_rate.push_back(cr.v); uncor_cor_rate.push_back(ur.v);
        total_cor += sCor; total_uncor += sUncor;
      }
      gc_ref.push_back(gcr); gc_cor.push_back(gcc);
      if (extended) ++count_ext;
      ++n_reads;
    };
    if (frags > 1) {                                           // a split read (:564-615)
      ++count_split;
      std::vector<uint8_t> kept;                               // realNotMissing as a column set
      for (int64_t f = 0; f < frags; ++f) {
        const int64_t r = i + f;
        if (!in.c(r, ELECTOR_T_ASSESSED)) continue;            // len(reference) <= 10 (:577)
        if (f == 0) thr_unc += in.c(r, ELECTOR_T_LENUNC);
        take(r);
        in.mask(r, m);
        if (kept.size() < m.size()) kept.resize(m.size(), 0);
        for (size_t p = 0; p < m.size(); ++p) kept[p] |= m[p];
        if (f == frags - 1) {
          missing_in_read = 0;
          if (!in.m_ref) { out.error = "split reads need the merged reference rows"; return ELECTOR_EINVAL; }
          const char *R = in.m_ref + in.m_off[r];
          const int64_t L = in.c(r, ELECTOR_T_NCOLS);
          for (int64_t p = 0; p < L; ++p) if (!kept[(size_t)p] && R[p] != '.') ++missing_in_read;
          output_metrics();
        }
      }
      i += frags;
    } else {
      if (in.c(i, ELECTOR_T_ASSESSED)) {
        thr_unc += in.c(i, ELECTOR_T_LENUNC);
        take(i);
        missing_in_read = in.c(i, ELECTOR_T_MISSING);
        output_metrics();
        if (missing_in_read > 5) ++count_trim;
      }
      ++i;
    }
  }
  if (gc_ref.empty() || total_cor + total_uncor == 0) { out.error = "no assessed read (the reference divides by zero here, computeStats.py:661-669)"; return ELECTOR_EINVAL; }
  auto fsum = [compensated_sum](const std::vector<double> &v) {   // Python's sum()
    double s = 0, c = 0;
    for (double x : v) {
      const double t = s + x;
      if (compensated_sum) c += std::fabs(s) >= std::fabs(x) ? (s - t) + x : (x - t) + s;
      s = t;
    }
    return (c != 0 && std::isfinite(c)) ? s + c : s;
  };
  elector_report_summary &s = out.s;
  const double gcR = py_round(fsum(gc_ref) / (double)gc_ref.size(), 3), gcC = py_round(fsum(gc_cor) / (double)gc_cor.size(), 3);
  const bool have = n_reads != 0;
  const double rec = have ? fsum(recall) * 1.0 / (double)n_reads : 0, prec = have ? fsum(precision) * 1.0 / (double)n_reads : 0,
               cbr = have ? fsum(cor_rate) * 1.0 / (double)n_reads : 0, ucbr = have ? fsum(uncor_cor_rate) * 1.0 / (double)n_reads : 0;
  int64_t thr_cor = 0;
  for (int64_t v : len_corrected) thr_cor += v;
  const double homopol = ratios.size() > 1 ? exact_mean(ratios) : 1;
  const int64_t trim_split = count_split + count_trim;
  int64_t miss_sum = 0;
  for (int64_t v : missing_sizes) miss_sum += v;
  // outputReadSizeDistribution (:273-288)
  out.size_distribution = "size type\n";
  for (int64_t v : len_corrected) out.size_distribution += std::to_string(v) + " reads\n";
  s.size_distribution_complete = 1;
  if (trim_split != 0) {
    FILE *f = corrected_fasta ? fopen(corrected_fasta, "rb") : nullptr;
    if (!f) s.size_distribution_complete = 0;
    else {
      std::string line;
      auto getl = [&](std::string &l) { l.clear(); int ch; while ((ch = fgetc(f)) != EOF) { l += (char)ch; if (ch == '\n') break; } };
      getl(line);
      while (!line.empty()) {
        getl(line);
        out.size_distribution += std::to_string(line.empty() ? 0 : line.size() - 1) + " sequences\n";
        getl(line);
      }
      fclose(f);
    }
  }
  s.assessed_reads = n_reads; s.throughput_uncorrected = thr_unc; s.throughput_corrected = thr_cor;
  s.recall = py_round(rec, 7); s.precision = py_round(prec, 7);
  s.correct_rate_uncorrected = ucbr; s.correct_rate_corrected = py_round(cbr, 7);
  s.error_rate = py_round(1 - (double)total_cor / (double)(total_cor + total_uncor), 7);
  s.trimmed_or_split = trim_split; s.split_reads = count_split; s.trimmed_reads = count_trim;
  s.mean_missing = trim_split > 0 ? py_round((double)miss_sum / (double)trim_split, 1) : 0;
  s.extended_reads = count_ext;
  s.mean_extension = count_ext > 0 ? py_round((double)ext_bases / (double)count_ext, 1) : 0;
  s.gc_ref = py_round(gcR * 100, 7); s.gc_cor = py_round(gcC * 100, 7);
  s.small_reads = small_reads; s.wrongly_cor_reads = wrongly_cor_reads;
  s.ins_u = idu[0]; s.del_u = idu[1]; s.subs_u = idu[2]; s.ins_c = idc[0]; s.del_c = idc[1]; s.subs_c = idc[2];
  s.homopolymer_ratio = homopol;
  // Python prints the ints 0 of the "if ... else 0" branches without ".0"
  auto f_or_int = [](double v, bool is_float) { return is_float ? py_float_str(v) : std::string("0"); };
  const std::string sRec = have ? py_float_str(s.recall) : "0", sPrec = have ? py_float_str(s.precision) : "0", sCbr = have ? py_float_str(s.correct_rate_corrected) : "0",
                    sUcbr = f_or_int(ucbr, have), sErrU = have ? py_float_str(1 - ucbr) : "1", sErrC = have ? py_float_str(1 - s.correct_rate_corrected) : "1",
                    sMiss = trim_split > 0 ? py_float_str(s.mean_missing) : "0", sExt = count_ext > 0 ? py_float_str(s.mean_extension) : "0",
                    sHom = ratios.size() > 1 ? py_float_str(homopol) : "1", sThr = py_float_str(size_threshold * 100);
  auto I = [](int64_t v) { return std::to_string(v); };
  out.log_text = "*********** SUMMARY ***********\nAssessed reads: " + I(n_reads) + "\nThroughput (uncorrected): " + I(thr_unc) + "\nThroughput (corrected): " + I(thr_cor) +
                 "\nRecall (computed only on corrected bases):" + sRec + "\nPrecision (computed only on corrected bases):" + sPrec + "\nAverage correct bases rate (uncorrected):" + sUcbr +
                 "\nError rate (uncorrected): " + sErrU + "\nAverage correct bases rate (corrected):" + sCbr + "\nError rate (corrected): " + sErrC +
                 "\nNumber of trimmed/split reads:" + I(trim_split) + "\nMean missing size in trimmed/split reads:" + sMiss + "\nNumber of over-corrected reads by extention: " + I(count_ext) +
                 "\nMean extension size in over-corrected reads: " + sExt + "\n%GC in reference reads: " + py_float_str(s.gc_ref) + "\n%GC in corrected reads: " + py_float_str(s.gc_cor) +
                 "\nNumber of corrected reads which length is <" + sThr + "% of the original read:" + I(small_reads) + "\nNumber of very low quality corrected reads: " + I(wrongly_cor_reads) +
                 "\nNumber of insertions in uncorrected: " + I(idu[0]) + "\nNumber of insertions in corrected: " + I(idc[0]) + "\nNumber of deletions in uncorrected: " + I(idu[1]) +
                 "\nNumber of deletions in corrected: " + I(idc[1]) + "\nNumber of substitutions in uncorrected: " + I(idu[2]) + "\nNumber of substitutions in corrected: " + I(idc[2]) +
                 "\nRatio of homopolymer sizes in corrected vs reference: " + sHom + "\n";
  const std::string softs = soft ? std::string(soft) : std::string("None");
  out.stdout_text = softs + "\n" + (soft ? softs + "\n" : std::string()) + "*********** SUMMARY ***********\nAssessed reads:  " + I(n_reads) + "\nThroughput (uncorrected) " + I(thr_unc) +
                    "\nThroughput (corrected):  " + I(thr_cor) + "\nRecall: " + sRec + "\nPrecision: " + sPrec + "\nAverage correct bases rate (uncorrected):  " + sUcbr +
                    "\nError rate (uncorrected): " + sErrU + "\nAverage correct bases rate (corrected):  " + sCbr + "\nError rate (corrected): " + sErrC +
                    "\nNumber of trimmed/split reads: " + I(trim_split) + "\nMean missing size in trimmed/split reads: " + sMiss + "\nNumber of over-corrected reads by extention:  " + I(count_ext) +
                    "\nMean extension size in over-corrected reads:  " + sExt + "\n%GC in reference reads:  " + py_float_str(s.gc_ref) + "\n%GC in corrected reads:  " + py_float_str(s.gc_cor) +
                    "\nNumber of corrected reads which length is < " + sThr + " % of the original read: " + I(small_reads) + "\nNumber of very low quality corrected reads:  " + I(wrongly_cor_reads) +
                    "\nNumber of insertions in uncorrected:  " + I(idu[0]) + "\nNumber of insertions in corrected:  " + I(idc[0]) + "\nNumber of deletions in uncorrected:  " + I(idu[1]) +
                    "\nNumber of deletions in corrected:  " + I(idc[1]) + "\nNumber of substitutions in uncorrected:  " + I(idu[2]) + "\nNumber of substitutions in corrected:  " + I(idc[2]) +
                    "\nRatio of homopolymer sizes in corrected vs reference: " + sHom + "\n";
  return ELECTOR_OK;
}

}  // namespace elector
