// npz.h -- minimal float32 .npz reader (see npz.cpp)
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct NpzArray {
  std::vector<int> shape;
  std::vector<float> data;
};
bool npz_load(const char* path, std::map<std::string, NpzArray>& out, std::string& err);

// Any-dtype member (region files: float64 xy/scales/responses/A/angles, uint8 descs -- cnpy::NpyArray's role in
// PreLoadRegionsNPZ, imagerepresentation.cpp:1355-1507)
struct NpzRaw {
  std::vector<int> shape;
  std::string descr;                 // e.g. "<f8", "|u1"
  std::vector<unsigned char> bytes;
  size_t count() const;
  bool numeric() const;
  double at(size_t i) const;
};
bool npz_load_raw(const char* path, std::map<std::string, NpzRaw>& out, std::string& err);

// Minimal .npz writer (zip archive of .npy members, stored uncompressed like cnpy::npz_save does):
// NpzWriter w(path); w.add("xy", "<f8", {n, 2}, ptr, bytes); ... w.close();
struct NpzWriter {
  explicit NpzWriter(const std::string& path);
  ~NpzWriter();
  bool add(const std::string& name, const char* descr, const std::vector<size_t>& shape, const void* data, size_t bytes);
  bool close();
  bool ok() const { return f_ != nullptr && good_; }
 private:
  struct Entry { std::string name; unsigned crc; unsigned long long size, offset; };
  void* f_ = nullptr;
  bool good_ = true;
  std::vector<Entry> entries_;
};
