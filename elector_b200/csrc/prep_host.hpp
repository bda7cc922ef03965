// prep_host.hpp -- the file preparation in front of the splitter: readAndSortFasta (elector/readAndSortFiles.py:150-166)
// and duplicateRefReads (:171-191) without Biopython objects.  The reference parses every record into a SeqRecord, sorts the list
// and writes it back (three files, whole in memory as Python objects); here a file is read once into one buffer, the records are
// (offset, length) pairs into it, the sort moves indices, and the output is written from the buffer.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace elector {

inline bool prep_read_file(const char *path, std::string &buf) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  char tmp[1 << 16];
  size_t n;
  buf.clear();
  while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
  const bool ok = !ferror(f);
  fclose(f);
  return ok;
}

struct PrepRecord { size_t desc, desc_len, seq, seq_end; };   // description (SeqRecord.description) and the raw span of its sequence lines

// what str.rstrip() strips
inline bool prep_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Bio.SeqIO.parse(handle, "fasta") (SimpleFastaParser): lines before the first '>' are skipped; title = header line without '>' and
// trailing white space; the sequence is every following line, right-stripped, joined, blanks and '\r' removed
inline void prep_parse(const std::string &b, std::vector<PrepRecord> &recs) {
  size_t p = 0;
  const size_t n = b.size();
  while (p < n && b[p] != '>') { const size_t e = b.find('\n', p); p = e == std::string::npos ? n : e + 1; }
  while (p < n) {
    size_t e = b.find('\n', p);
    if (e == std::string::npos) e = n;
    size_t de = e;
    while (de > p + 1 && prep_space(b[de - 1])) --de;
    PrepRecord r{p + 1, de - (p + 1), e < n ? e + 1 : n, n};
    size_t q = r.seq;
    while (q < n && b[q] != '>') { const size_t le = b.find('\n', q); q = le == std::string::npos ? n : le + 1; }
    r.seq_end = q;
    recs.push_back(r);
    p = q;
  }
}

inline void prep_append_seq(const std::string &b, const PrepRecord &r, std::string &out) {
  size_t q = r.seq;
  while (q < r.seq_end) {
    size_t le = b.find('\n', q);
    if (le == std::string::npos || le > r.seq_end) le = r.seq_end;
    size_t e = le;
    while (e > q && prep_space(b[e - 1])) --e;
    for (size_t i = q; i < e; ++i) if (b[i] != ' ' && b[i] != '\r') out += b[i];
    q = le + 1;
  }
}

// readAndSortFasta: records sorted by description (sorted() is stable; str order is code-point order, which UTF-8 bytes keep),
// written as ">description\nsequence\n".  Returns the number of records, -1 on I/O errors.
inline int64_t prep_sort_fasta(const char *in_path, const char *out_path) {
  std::string b;
  if (!prep_read_file(in_path, b)) return -1;
  std::vector<PrepRecord> recs;
  prep_parse(b, recs);
  std::vector<uint32_t> order(recs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    const PrepRecord &a = recs[x], &c = recs[y];
    const int d = memcmp(b.data() + a.desc, b.data() + c.desc, std::min(a.desc_len, c.desc_len));
    return d < 0 || (d == 0 && a.desc_len < c.desc_len);
  });
  FILE *o = fopen(out_path, "wb");
  if (!o) return -1;
  std::string out;
  for (uint32_t i : order) {
    out += '>'; out.append(b, recs[i].desc, recs[i].desc_len); out += '\n';
    prep_append_seq(b, recs[i], out);
    out += '\n';
    if (out.size() > (1u << 22)) { if (fwrite(out.data(), 1, out.size(), o) != out.size()) { fclose(o); return -1; } out.clear(); }
  }
  const bool ok = fwrite(out.data(), 1, out.size(), o) == out.size();
  return (fclose(o) == 0 && ok) ? (int64_t)recs.size() : -1;
}

// duplicateRefReads on the three sorted files: every reference / uncorrected record whose header is that of k corrected records is
// written k times as header_0 .. header_(k-1); records without a corrected read are dropped.  The two files are walked line by line
// side by side like the reference's zip() (a line with a '>' anywhere is a header line, :180-190).  Returns the triplets written.
inline int64_t prep_duplicate(const char *sorted_ref, const char *sorted_unc, const char *sorted_cor, const char *new_ref, const char *new_unc) {
  std::string br, bu, bc;
  if (!prep_read_file(sorted_ref, br) || !prep_read_file(sorted_unc, bu) || !prep_read_file(sorted_cor, bc)) return -1;
  std::unordered_map<std::string, int64_t> occ;
  {
    std::vector<PrepRecord> recs;
    prep_parse(bc, recs);
    for (const PrepRecord &r : recs) ++occ[bc.substr(r.desc, r.desc_len)];
  }
  FILE *fr = fopen(new_ref, "wb"), *fu = fopen(new_unc, "wb");
  if (!fr || !fu) { if (fr) fclose(fr); if (fu) fclose(fu); return -1; }
  auto line_of = [](const std::string &b, size_t &p, size_t &s, size_t &e) {   // next line [s, e) with its newline dropped; false at the end
    if (p >= b.size()) return false;
    s = p;
    const size_t nl = b.find('\n', p);
    e = nl == std::string::npos ? b.size() : nl;
    p = nl == std::string::npos ? b.size() : nl + 1;
    return true;
  };
  auto rstrip = [](const std::string &b, size_t s, size_t e) { while (e > s && prep_space(b[e - 1])) --e; return e; };
  size_t pr = 0, pu = 0, rs, re, us, ue;
  std::string header, outr, outu;
  bool have = false;
  int64_t written = 0;
  while (line_of(bu, pu, us, ue) && line_of(br, pr, rs, re)) {
    if (memchr(br.data() + rs, '>', re - rs)) { const size_t e = rstrip(br, rs, re); header = e > rs + 1 ? br.substr(rs + 1, e - rs - 1) : std::string(); have = true; continue; }
    if (!have) return -2;   // the reference fails here with an unbound `header`
    const auto it = occ.find(header);
    if (it == occ.end()) continue;
    const size_t er = rstrip(br, rs, re), eu = rstrip(bu, us, ue);
    for (int64_t t = 0; t < it->second; ++t) {
      const std::string h = ">" + header + "_" + std::to_string(t) + "\n";
      outr += h; outr.append(br, rs, er - rs); outr += '\n';
      outu += h; outu.append(bu, us, eu - us); outu += '\n';
      ++written;
    }
    if (outr.size() > (1u << 22)) { fwrite(outr.data(), 1, outr.size(), fr); fwrite(outu.data(), 1, outu.size(), fu); outr.clear(); outu.clear(); }
  }
  const bool ok = fwrite(outr.data(), 1, outr.size(), fr) == outr.size() && fwrite(outu.data(), 1, outu.size(), fu) == outu.size();
  const bool c1 = fclose(fr) == 0, c2 = fclose(fu) == 0;
  return (ok && c1 && c2) ? written : -1;
}

}  // namespace elector
