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
