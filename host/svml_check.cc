// Test helper (no GPU needed): reads an SVMLight file with the host reader and prints its shape and
// FNV-1a checksums of the labels, query offsets and feature matrix, so that tests can compare the
// multi-threaded parse with the single-threaded one and with an independent parse.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "quickrank_host.h"

static uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char *b = (const unsigned char *) p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char **argv) {
  if (argc < 2) { std::cerr << "usage: svml_check file" << std::endl; return 1; }
  quickrank::io::Svml reader;
  auto ds = reader.read_horizontal(argv[1]);
  const size_t n = ds->num_instances(), f = ds->num_features(), q = ds->num_queries();
  std::vector<uint64_t> off(ds->offsets().begin(), ds->offsets().end());
  printf("%zu %zu %zu %016llx %016llx %016llx\n", n, f, q, (unsigned long long) fnv(ds->labels(), n * sizeof(float)),
         (unsigned long long) fnv(off.data(), off.size() * sizeof(uint64_t)),
         (unsigned long long) fnv(ds->data(), n * f * sizeof(float)));
  return 0;
}
