// RSGV: a trivial container of named uint64 arrays, used to move golden vectors between the
// reference harness (oracle/_ref), the C oracle and the Python tests.  TEST INFRASTRUCTURE.
//
//   file   := magic "RSGV0001" | u64 n_entries | entry*
//   entry  := u64 name_len | name bytes, zero-padded to a multiple of 8 | u64 n_words | u64 words[n_words]
// All integers little-endian.  Python twin: tests/rsgv.py.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rsgv {

class Writer {
 public:
  void put(const std::string &name, const std::vector<uint64_t> &v) { entries_.emplace_back(name, v); }
  void put(const std::string &name, const uint64_t *p, size_t n) {
    entries_.emplace_back(name, std::vector<uint64_t>(p, p + n));
  }
  void put1(const std::string &name, uint64_t v) { entries_.emplace_back(name, std::vector<uint64_t>{v}); }
  void save(const std::string &path) const {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    fwrite("RSGV0001", 1, 8, f);
    uint64_t n = entries_.size();
    fwrite(&n, 8, 1, f);
    for (const auto &e : entries_) {
      uint64_t len = e.first.size();
      fwrite(&len, 8, 1, f);
      std::string padded = e.first;
      padded.resize((len + 7) / 8 * 8, '\0');
      fwrite(padded.data(), 1, padded.size(), f);
      uint64_t nw = e.second.size();
      fwrite(&nw, 8, 1, f);
      if (nw) fwrite(e.second.data(), 8, nw, f);
    }
    fclose(f);
  }

 private:
  std::vector<std::pair<std::string, std::vector<uint64_t>>> entries_;
};

inline std::map<std::string, std::vector<uint64_t>> load(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "RSGV0001", 8) != 0) throw std::runtime_error("bad magic");
  uint64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) throw std::runtime_error("truncated");
  std::map<std::string, std::vector<uint64_t>> out;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t len = 0, nw = 0;
    if (fread(&len, 8, 1, f) != 1) throw std::runtime_error("truncated");
    std::string name((len + 7) / 8 * 8, '\0');
    if (!name.empty() && fread(&name[0], 1, name.size(), f) != name.size()) throw std::runtime_error("truncated");
    name.resize(len);
    if (fread(&nw, 8, 1, f) != 1) throw std::runtime_error("truncated");
    std::vector<uint64_t> v(nw);
    if (nw && fread(v.data(), 8, nw, f) != nw) throw std::runtime_error("truncated");
    out.emplace(std::move(name), std::move(v));
  }
  fclose(f);
  return out;
}

}  // namespace rsgv
