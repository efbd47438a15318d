// The reference's integration-test driver (tests/simple_model_wiki/main.cpp:26-42), parameterised:
// for every key of the data set, |lookup(key) - lower_bound(key)| <= err.   KEY_T / WITH_ERR via -D.
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <vector>
#include "rmi.h"
#ifndef KEY_T
#define KEY_T uint64_t
#endif
#ifndef FILE_T
#define FILE_T KEY_T
#endif
int main(int argc, char** argv) {
  std::vector<FILE_T> data;
  std::ifstream in(argv[1], std::ios::binary);
  uint64_t size;
  in.read(reinterpret_cast<char*>(&size), sizeof(uint64_t));
  data.resize(size);
  in.read(reinterpret_cast<char*>(data.data()), size * sizeof(FILE_T));
  if (!rmi::load(argv[2])) { std::cout << "load failed" << std::endl; return 3; }
  uint64_t worst = 0;
  for (uint64_t i = 0; i < size; i++) {
    KEY_T key = (KEY_T)data[i];
    uint64_t true_index = (uint64_t)std::distance(data.begin(), std::lower_bound(data.begin(), data.end(), data[i]));
#ifdef NO_ERR
    uint64_t guess = rmi::lookup(key);
    (void)guess; (void)true_index;
#else
    size_t err;
    uint64_t guess = rmi::lookup(key, &err);
    uint64_t diff = guess > true_index ? guess - true_index : true_index - guess;
    if (diff > err) {
      std::cout << "key index " << i << " guess " << guess << " +/- " << err << " true " << true_index << std::endl;
      return 1;
    }
    worst = std::max<uint64_t>(worst, diff);
#endif
  }
  rmi::cleanup();
  std::cout << "ok worst " << worst << " size " << rmi::RMI_SIZE << " name " << rmi::NAME << std::endl;
  return 0;
}
