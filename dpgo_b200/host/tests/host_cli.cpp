// CPU-only command line probe of the host-side (non-CUDA) parts of the drop-in API, used by the
// "not gpu" tests:  host_cli parse <file.g2o>   |   host_cli chordal <file.g2o>
#include <DPGO/DPGO_solver.h>
#include <DPGO/DPGO_utils.h>

#include <cstdio>
#include <cstring>

using namespace DPGO;

int main(int argc, char **argv) {
  if (argc != 3) {
    std::fprintf(stderr, "usage: %s parse|chordal <file.g2o>\n", argv[0]);
    return 2;
  }
  size_t n = 0;
  const std::vector<RelativeSEMeasurement> ms = read_g2o_file(argv[2], n);
  const size_t d = ms.empty() ? 0 : static_cast<size_t>(ms[0].t.size());
  if (!std::strcmp(argv[1], "parse")) {
    std::printf("%zu %zu %zu\n", n, ms.size(), d);
    for (const auto &m : ms) {
      std::printf("%zu %zu %d", m.p1, m.p2, m.fixedWeight ? 1 : 0);
      for (size_t a = 0; a < d; ++a)
        for (size_t b = 0; b < d; ++b) std::printf(" %.17g", m.R(a, b));
      for (size_t a = 0; a < d; ++a) std::printf(" %.17g", m.t(a, 0));
      std::printf(" %.17g %.17g\n", m.kappa, m.tau);
    }
    return 0;
  }
  if (!std::strcmp(argv[1], "chordal")) {
    const Matrix T = chordalInitialization(ms).getData();
    std::printf("%td %td\n", T.rows(), T.cols());
    for (std::ptrdiff_t j = 0; j < T.cols(); ++j)
      for (std::ptrdiff_t i = 0; i < T.rows(); ++i) std::printf("%.17g\n", T(i, j));
    return 0;
  }
  return 2;
}
