// oracle/ref_rng.cpp -- TEST INFRASTRUCTURE.  The random streams the reference's tests/testTracker.cpp draws its
// synthetic scenes from, produced by the very library calls the test makes:
//   rand  N        srand(3) (testTracker.cpp:59) then N x rand()                       -- glibc
//   normal S N     default_random_engine g; normal_distribution<double> d(0, S); N x d(g)  (testTracker.cpp:203-204,515-516)
// Build: g++ -O2 -std=c++17 -o ref_rng ref_rng.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "rand")) {
    srand(3);
    const long n = atol(argv[2]);
    for (long i = 0; i < n; ++i) printf("%d\n", rand());
    printf("RAND_MAX %d\n", RAND_MAX);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "normal")) {
    std::default_random_engine g;
    std::normal_distribution<double> d(0, atof(argv[2]));
    const long n = atol(argv[3]);
    for (long i = 0; i < n; ++i) printf("%.17g\n", d(g));
    return 0;
  }
  fprintf(stderr, "usage: ref_rng rand N | normal SIGMA N\n");
  return 2;
}
