#include "flat_map.hpp"
#include <unordered_map>
#include <random>
#include <cassert>
#include <cstdio>
int main() {
  batching_pomp::util::FlatMap64<uint32_t> m; std::unordered_map<uint64_t,uint32_t> r;
  std::mt19937_64 g(1);
  for (int it = 0; it < 2000000; ++it) {
    uint64_t k = g() % 50000; int op = g() % 10;
    if (op < 6) { bool a = m.emplace(k, (uint32_t)it); bool b = r.emplace(k, (uint32_t)it).second; assert(a == b); }
    else if (op < 8) { bool a = m.erase(k); bool b = r.erase(k) > 0; assert(a == b); }
    else { auto* p = m.find(k); auto q = r.find(k); assert((p != nullptr) == (q != r.end())); if (p) assert(*p == q->second); }
    if (it % 500000 == 499999) { m.clear(); r.clear(); }
  }
  assert(m.size() == r.size());
  printf("flat map ok %zu\n", m.size());
}
