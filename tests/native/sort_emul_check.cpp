// Checks srrg2_proslam_b200/csrc/libstdcxx_sort.h (host build) against the real libstdc++
// std::sort: identical permutation for one-field comparators on many key distributions,
// including few-distinct-key inputs (ties) and median-of-3 killers (heap-sort fallback).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "../../srrg2_proslam_b200/csrc/libstdcxx_sort.h"

struct Item {
  int key;
  int id;
};

static std::vector<int> killer(int n) {  // Musser's median-of-3 killer
  std::vector<int> v(n);
  int k = n / 2;
  for (int i = 0; i < k; ++i) {
    if (i % 2 == 0) {
      v[i] = i + 1;
    } else {
      v[i] = k + i + (k % 2 == 0 ? 0 : 1);
    }
    v[k + i] = 2 * (i + 1);
  }
  return v;
}

static long check(const std::vector<int>& keys, bool desc) {
  const int n = (int) keys.size();
  std::vector<Item> a(n), b(n);
  for (int i = 0; i < n; ++i) a[i] = b[i] = Item{keys[i], i};
  if (desc) {
    std::sort(a.begin(), a.end(), [](const Item& x, const Item& y) { return x.key > y.key; });
    pslam_sort::std_sort(b.data(), n, [](const Item& x, const Item& y) { return x.key > y.key; });
  } else {
    std::sort(a.begin(), a.end(), [](const Item& x, const Item& y) { return x.key < y.key; });
    pslam_sort::std_sort(b.data(), n, [](const Item& x, const Item& y) { return x.key < y.key; });
  }
  long bad = 0;
  for (int i = 0; i < n; ++i) bad += (a[i].id != b[i].id);
  // prefix-pruned variant: first `need` positions must equal the full std::sort
  for (int need : {0, 1, n / 9, n / 2, n - 1, n}) {
    if (need < 0 || need > n) continue;
    std::vector<Item> c(n);
    for (int i = 0; i < n; ++i) c[i] = Item{keys[i], i};
    if (desc)
      pslam_sort::std_sort_prefix(c.data(), n, need, [](const Item& x, const Item& y) { return x.key > y.key; });
    else
      pslam_sort::std_sort_prefix(c.data(), n, need, [](const Item& x, const Item& y) { return x.key < y.key; });
    for (int i = 0; i < need; ++i) bad += (a[i].id != c[i].id);
  }
  return bad;
}

int main() {
  std::mt19937 rng(12345);
  long bad = 0, cases = 0;
  const int sizes[] = {0, 1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 100, 111, 112, 500, 701, 1000, 4096, 20000};
  for (int n : sizes) {
    for (int distinct : {1, 2, 3, 5, 16, 64, 255, 100000}) {
      for (int rep = 0; rep < 6; ++rep) {
        std::vector<int> k(n);
        for (int& v : k) v = (int) (rng() % (unsigned) distinct);
        bad += check(k, true);
        bad += check(k, false);
        cases += 2;
      }
    }
    std::vector<int> asc(n), dsc(n);
    for (int i = 0; i < n; ++i) {
      asc[i] = i;
      dsc[i] = n - i;
    }
    bad += check(asc, true) + check(asc, false) + check(dsc, true) + check(dsc, false);
    std::vector<int> kl = killer(n);
    bad += check(kl, false) + check(kl, true);
    std::vector<int> organ(n);
    for (int i = 0; i < n; ++i) organ[i] = std::min(i, n - i);
    bad += check(organ, false) + check(organ, true);
    cases += 8;
  }
  // sawtooth / many killers concatenated to force the depth limit
  for (int n : {4096, 65536}) {
    std::vector<int> k = killer(n);
    for (int round = 0; round < 3; ++round) {
      bad += check(k, false);
      std::rotate(k.begin(), k.begin() + n / 3, k.end());
      ++cases;
    }
  }
  std::printf("%s cases=%ld mismatches=%ld\n", bad ? "FAIL" : "OK", cases, bad);
  return bad ? 1 : 0;
}
