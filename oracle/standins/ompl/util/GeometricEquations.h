// OMPL API stand-in (oracle/standins/README.md): unitNBallMeasure = pi^(N/2) / Gamma(N/2 + 1)  
#pragma once
#include <cmath>
namespace ompl {
inline double unitNBallMeasure(unsigned int N) {
  return std::pow(std::sqrt(M_PI), static_cast<double>(N)) / std::tgamma(static_cast<double>(N) / 2.0 + 1.0);
}
}  // namespace ompl
