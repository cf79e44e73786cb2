// integration/check_spline_adapter.cpp -- instantiates SplineB200Core for the four (storage, value) pairs the reference
// builds (real / complex x mixed / full) against the reference's einspline structs and containers; `g++ -fsyntax-only`.
#include <complex>
#include <cstddef>
#include <vector>
#include "config.h"
#include "type_traits/template_types.hpp"
#include "SplineB200.h"
namespace qmcplusplus
{
template class SplineB200Core<float, float>;
template class SplineB200Core<double, double>;
template class SplineB200Core<float, std::complex<float>>;
template class SplineB200Core<double, std::complex<double>>;
} // namespace qmcplusplus
