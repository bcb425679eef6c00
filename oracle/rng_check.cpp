// oracle/rng_check.cpp -- TEST INFRASTRUCTURE.  Prints the first N values of OpenGV's rnd()
// (= std::uniform_int_distribution<>(0, INT_MAX) bound to std::mt19937 seeded 12345u, see
// opengv/sac/implementation/SampleConsensusProblem.hpp) as produced by THIS image's libstdc++,
// plus the all-equal-keys std::sort permutation cv::sortIdx relies on.
// Used by tests/test_oracle_ransac.py to pin oracle/ransac.py:rnd_table("lemire").
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  if (argc > 2 && std::string(argv[1]) == "rnd") {
    int n = std::atoi(argv[2]);
    std::uniform_int_distribution<> dist(0, std::numeric_limits<int>::max());
    std::mt19937 alg;
    alg.seed(12345u);
    std::function<int()> gen = std::bind(dist, alg);
    for (int i = 0; i < n; ++i) std::printf("%d\n", gen());
    return 0;
  }
  if (argc > 2 && std::string(argv[1]) == "sortperm") {
    // cv::sortIdx(SORT_DESCENDING) on all-equal int keys: std::sort of indices with a key
    // comparator (always false), then reversed.
    int n = std::atoi(argv[2]);
    std::vector<int> keys(n, 0), idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    std::reverse(idx.begin(), idx.end());
    for (int i = 0; i < n; ++i) std::printf("%d\n", idx[i]);
    return 0;
  }
  std::fprintf(stderr, "usage: rng_check rnd N | sortperm N\n");
  return 2;
}
