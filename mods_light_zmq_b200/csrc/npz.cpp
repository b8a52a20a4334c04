// npz.cpp -- minimal .npz (zip of .npy) reader for the network weights.
// Plays the role cnpy::npz_load (cnpy/cnpy.h:71-72) plays in the reference, restricted to what
// tools/export_weights.py writes: little-endian float32, C order, stored or deflated members.
#include "npz.h"
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace {
uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

bool parse_npy_raw(const std::vector<unsigned char>& raw, NpzRaw& a, std::string& err) {
  if (raw.size() < 10 || memcmp(raw.data(), "\x93NUMPY", 6) != 0) { err = "not an npy member"; return false; }
  int major = raw[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = rd16(&raw[8]); hoff = 10; }
  else { hlen = rd32(&raw[8]); hoff = 12; }
  if (hoff + hlen > raw.size()) { err = "truncated npy header"; return false; }
  std::string hdr((const char*)&raw[hoff], hlen);
  size_t dp = hdr.find("'descr':");
  if (dp == std::string::npos) { err = "npy header without descr"; return false; }
  size_t q1 = hdr.find('\'', dp + 8), q2 = q1 == std::string::npos ? q1 : hdr.find('\'', q1 + 1);
  if (q2 == std::string::npos) { err = "bad descr"; return false; }
  a.descr = hdr.substr(q1 + 1, q2 - q1 - 1);
  if (a.descr.size() < 3) { err = "unsupported npy dtype " + a.descr; return false; }
  const int item = atoi(a.descr.c_str() + 2);
  if (item <= 0 || (a.descr[0] == '>' && item > 1)) { err = "unsupported npy dtype " + a.descr; return false; }
  if (hdr.find("'fortran_order': False") == std::string::npos) { err = "npy is fortran ordered"; return false; }
  size_t sp = hdr.find("'shape':");
  if (sp == std::string::npos) { err = "npy header without shape"; return false; }
  size_t lp = hdr.find('(', sp), rp = hdr.find(')', sp);
  if (lp == std::string::npos || rp == std::string::npos) { err = "bad shape"; return false; }
  a.shape.clear();
  size_t n = 1;
  const char* s = hdr.c_str() + lp + 1;
  const char* e = hdr.c_str() + rp;
  while (s < e) {
    while (s < e && (*s == ' ' || *s == ',')) s++;
    if (s >= e) break;
    char* q;
    long v = strtol(s, &q, 10);
    if (q == s || v < 0) break;
    a.shape.push_back((int)v);
    n *= (size_t)v;
    s = q;
  }
  size_t doff = hoff + hlen;
  if (doff + n * (size_t)item > raw.size()) { err = "npy payload shorter than its shape"; return false; }
  a.bytes.assign(raw.begin() + doff, raw.begin() + doff + n * (size_t)item);
  return true;
}

}  // namespace

namespace {
bool read_members(const char* path, std::map<std::string, NpzRaw>& out, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return false; }
  fseek(f, 0, SEEK_END);
  long fsz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> buf((size_t)fsz);
  if (fsz <= 22 || fread(buf.data(), 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); err = "short read"; return false; }
  fclose(f);
  long eocd = -1;
  for (long i = fsz - 22; i >= 0 && i >= fsz - 22 - 65536; i--)
    if (rd32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
  if (eocd < 0) { err = "no zip end-of-central-directory record"; return false; }
  int count = rd16(&buf[eocd + 10]);
  size_t cd = rd32(&buf[eocd + 16]);
  for (int k = 0; k < count; k++) {
    if (cd + 46 > (size_t)fsz || rd32(&buf[cd]) != 0x02014b50u) { err = "bad central directory"; return false; }
    int method = rd16(&buf[cd + 10]);
    size_t csize = rd32(&buf[cd + 20]), usize = rd32(&buf[cd + 24]);
    int nlen = rd16(&buf[cd + 28]), xlen = rd16(&buf[cd + 30]), clen = rd16(&buf[cd + 32]);
    size_t lho = rd32(&buf[cd + 42]);
    std::string name((const char*)&buf[cd + 46], nlen);
    cd += 46 + nlen + xlen + clen;
    if (lho + 30 > (size_t)fsz || rd32(&buf[lho]) != 0x04034b50u) { err = "bad local header"; return false; }
    size_t doff = lho + 30 + rd16(&buf[lho + 26]) + rd16(&buf[lho + 28]);
    if (doff + csize > (size_t)fsz) { err = "member beyond end of file"; return false; }
    std::vector<unsigned char> raw;
    if (method == 0) raw.assign(buf.begin() + doff, buf.begin() + doff + csize);
    else if (method == 8) {
      raw.resize(usize);
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) { err = "zlib init"; return false; }
      zs.next_in = &buf[doff]; zs.avail_in = (uInt)csize;
      zs.next_out = raw.data(); zs.avail_out = (uInt)usize;
      int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END) { err = "inflate failed for " + name; return false; }
    } else { err = "unsupported zip method in " + name; return false; }
    if (name.size() > 4 && name.substr(name.size() - 4) == ".npy") name = name.substr(0, name.size() - 4);
    NpzRaw a;
    if (!parse_npy_raw(raw, a, err)) { err += " (" + name + ")"; return false; }
    out[name] = std::move(a);
  }
  return true;
}
}  // namespace

bool npz_load_raw(const char* path, std::map<std::string, NpzRaw>& out, std::string& err) { return read_members(path, out, err); }

bool npz_load(const char* path, std::map<std::string, NpzArray>& out, std::string& err) {
  std::map<std::string, NpzRaw> raw;
  if (!read_members(path, raw, err)) return false;
  for (auto& kv : raw) {
    if (kv.second.descr != "<f4") { err = "npy dtype is not <f4 (" + kv.first + ")"; return false; }
    NpzArray a;
    a.shape = kv.second.shape;
    a.data.resize(kv.second.bytes.size() / 4);
    memcpy(a.data.data(), kv.second.bytes.data(), a.data.size() * 4);
    out[kv.first] = std::move(a);
  }
  return true;
}

size_t NpzRaw::count() const { size_t n = 1; for (int v : shape) n *= (size_t)v; return n; }

// element i as double, for the dtypes numpy / cnpy write for region files (f8, f4, u1, i4, i8, u2 ...)
double NpzRaw::at(size_t i) const {
  const char k = descr[1];
  const int item = atoi(descr.c_str() + 2);
  const unsigned char* p = bytes.data() + i * (size_t)item;
  if (k == 'f' && item == 8) { double v; memcpy(&v, p, 8); return v; }
  if (k == 'f' && item == 4) { float v; memcpy(&v, p, 4); return v; }
  if (k == 'u' && item == 1) return *p;
  if (k == 'i' && item == 1) return (signed char)*p;
  if (k == 'b' && item == 1) return *p != 0;
  if (k == 'u' && item == 2) { uint16_t v; memcpy(&v, p, 2); return v; }
  if (k == 'i' && item == 2) { int16_t v; memcpy(&v, p, 2); return v; }
  if (k == 'u' && item == 4) { uint32_t v; memcpy(&v, p, 4); return v; }
  if (k == 'i' && item == 4) { int32_t v; memcpy(&v, p, 4); return v; }
  if (k == 'u' && item == 8) { uint64_t v; memcpy(&v, p, 8); return (double)v; }
  if (k == 'i' && item == 8) { int64_t v; memcpy(&v, p, 8); return (double)v; }
  return 0.0;
}
bool NpzRaw::numeric() const {
  const char k = descr[1];
  const int item = atoi(descr.c_str() + 2);
  return (k == 'f' && (item == 4 || item == 8)) || ((k == 'u' || k == 'i') && (item == 1 || item == 2 || item == 4 || item == 8)) || (k == 'b' && item == 1);
}


// ---- writer ---------------------------------------------------------------------------------------------------
#include <cstdio>
namespace {
void put16(std::vector<unsigned char>& b, unsigned v) { b.push_back(v & 255); b.push_back((v >> 8) & 255); }
void put32(std::vector<unsigned char>& b, unsigned v) { for (int i = 0; i < 4; i++) b.push_back((v >> (8 * i)) & 255); }
}  // namespace

NpzWriter::NpzWriter(const std::string& path) { f_ = fopen(path.c_str(), "wb"); }
NpzWriter::~NpzWriter() { if (f_) fclose((FILE*)f_); }

bool NpzWriter::add(const std::string& name, const char* descr, const std::vector<size_t>& shape, const void* data, size_t bytes) {
  if (!f_) return false;
  FILE* f = (FILE*)f_;
  // .npy v1.0 header, padded so that the data starts at a multiple of 64 bytes
  std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  dict += "), }";
  size_t hl = dict.size() + 1;
  while ((10 + hl) % 64) hl++;
  dict.resize(hl - 1, ' ');
  dict += '\n';
  std::vector<unsigned char> npy = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  put16(npy, (unsigned)dict.size());
  npy.insert(npy.end(), dict.begin(), dict.end());
  const std::string fname = name + ".npy";
  const unsigned long long total = npy.size() + bytes;
  if (total > 0xffffffffull) { good_ = false; return false; }   // zip64 not needed on this path
  unsigned crc = crc32(0L, npy.data(), (uInt)npy.size());
  const unsigned char* p = (const unsigned char*)data;
  for (size_t off = 0; off < bytes; off += (size_t)1 << 30) crc = crc32(crc, p + off, (uInt)std::min<size_t>(bytes - off, (size_t)1 << 30));
  Entry e{fname, crc, total, (unsigned long long)ftell(f)};
  std::vector<unsigned char> lh;
  put32(lh, 0x04034b50); put16(lh, 20); put16(lh, 0); put16(lh, 0); put16(lh, 0); put16(lh, 0);
  put32(lh, crc); put32(lh, (unsigned)total); put32(lh, (unsigned)total); put16(lh, (unsigned)fname.size()); put16(lh, 0);
  lh.insert(lh.end(), fname.begin(), fname.end());
  good_ = good_ && fwrite(lh.data(), 1, lh.size(), f) == lh.size() && fwrite(npy.data(), 1, npy.size(), f) == npy.size() &&
          (bytes == 0 || fwrite(data, 1, bytes, f) == bytes);
  entries_.push_back(e);
  return good_;
}

bool NpzWriter::close() {
  if (!f_) return false;
  FILE* f = (FILE*)f_;
  const unsigned long long cd_off = (unsigned long long)ftell(f);
  std::vector<unsigned char> cd;
  for (const Entry& e : entries_) {
    put32(cd, 0x02014b50); put16(cd, 20); put16(cd, 20); put16(cd, 0); put16(cd, 0); put16(cd, 0); put16(cd, 0);
    put32(cd, e.crc); put32(cd, (unsigned)e.size); put32(cd, (unsigned)e.size); put16(cd, (unsigned)e.name.size());
    put16(cd, 0); put16(cd, 0); put16(cd, 0); put16(cd, 0); put32(cd, 0); put32(cd, (unsigned)e.offset);
    cd.insert(cd.end(), e.name.begin(), e.name.end());
  }
  const unsigned cd_size = (unsigned)cd.size();
  put32(cd, 0x06054b50); put16(cd, 0); put16(cd, 0); put16(cd, (unsigned)entries_.size()); put16(cd, (unsigned)entries_.size());
  put32(cd, cd_size); put32(cd, (unsigned)cd_off); put16(cd, 0);
  good_ = good_ && fwrite(cd.data(), 1, cd.size(), f) == cd.size();
  good_ = (fclose(f) == 0) && good_;
  f_ = nullptr;
  return good_;
}
